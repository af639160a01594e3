/*
 * ptk.h -- C ABI of libptk.so, the B200 (sm_100a) stereo-reconstruction hot path.
 *
 * The reference (ANP-Granular/ParticleTracking, package ParticleDetection) is pure
 * Python and has no FFI of its own, so the boundary a maintainer would
 * bind is the body of three functions.  Each entry point below names the reference
 * lines it replaces; paths are relative to
 *   ParticleDetection/src/ParticleDetection/reconstruct_3D/
 * INTEGRATION.md shows the ctypes stub for each.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch types.  Functions whose names end in
 *    `_host` take HOST pointers and do their own (pinned, chunked, overlapped) copies;
 *    all others take DEVICE pointers owned by the caller, never allocate, and only
 *    enqueue work on `stream` (a cudaStream_t passed as void*; NULL = default stream).
 *  - return 0 on success, a negative ptk_status otherwise; never throw, never exit.
 *  - all floating point is IEEE binary64; indices are int32; a "problem" is one
 *    (frame, colour) matching problem; arrays are dense, row-major, padded to Nmax.
 *  - re-entrant: no global mutable state; a ptk_ctx must not be used from two threads
 *    at once (the reference's callers run one worker per colour -- give each its own).
 */
#ifndef PTK_H_
#define PTK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTK_VERSION 100 /* 0.1.0 */
#define PTK_MAX_RODS 256 /* largest N1, N2, N per problem (BASELINE config 5) */
#define PTK_ROW_COLS 18  /* x1 y1 z1 x2 y2 z2 x y z l + 4 cam1 + 4 cam2 (match2D.py:963,986) */

typedef enum ptk_status {
    PTK_OK = 0,
    PTK_ERR_INVALID_ARG = -1, /* NULL pointer, negative size, N > PTK_MAX_RODS */
    PTK_ERR_CUDA = -2,        /* a CUDA runtime call failed; see ptk_last_cuda_error() */
    PTK_ERR_WORKSPACE = -3,   /* workspace too small; see ptk_workspace_bytes() */
    PTK_ERR_NO_DEVICE = -4,   /* no CUDA device / not an sm_100 device */
    PTK_ERR_NUMERIC = -5      /* host API only: some problem had NaN/-inf costs or was infeasible
                                 (scipy raises ValueError there); per-problem detail in `status` */
} ptk_status;

/* per-problem status written by the solver kernels (mirrors scipy's rectangular_lsap codes,
 * which match2D.py:933 and tracking.py:122 surface as ValueError) */
#define PTK_PROBLEM_OK 0
#define PTK_PROBLEM_INVALID 1    /* cost matrix contains NaN or -inf */
#define PTK_PROBLEM_INFEASIBLE 2 /* no finite assignment */
#define PTK_PROBLEM_EMPTY 3      /* n1 == 0 or n2 == 0: match_frame returns None (match2D.py:838-840) */

/*
 * The rig: everything match_frame receives besides the rod data (match2D.py:741-757).
 * The reference passes P1, P2, r1, r2, t1, t2 redundantly with `calibration`; they are
 * honoured as passed, not recomputed.
 */
typedef struct ptk_rig {
    double CM1[9], CM2[9];       /* calibration["CM1"], ["CM2"], row-major 3x3 */
    double dist1[14], dist2[14]; /* calibration["dist1"], ["dist2"], zero padded (k1 k2 p1 p2 k3 k4 k5 k6 s1..s4 tx ty) */
    double R1[9], t1[3];         /* r1, t1 given to cv2.projectPoints for camera 1 (match2D.py:898) */
    double R2[9], t2[3];         /* r2, t2 for camera 2 (match2D.py:901) */
    double P1[12], P2[12];       /* 3x4 projection matrices given to cv2.triangulatePoints (match2D.py:889) */
    double Rw[9], tw[3];         /* world transform: X_w = Rw X + tw == rot.apply(X) + trans (match2D.py:913) */
    int32_t agg_sum;             /* 0: mean over the two cameras (match2D.py:910); 1: sum (matchND.py:525) */
    int32_t reserved;
} ptk_rig;

/* ---- library ----------------------------------------------------------------------- */
int ptk_version(void);
const char* ptk_strerror(int status);
/* cudaGetErrorString of the last CUDA failure seen by the calling thread ("" if none). */
const char* ptk_last_cuda_error(void);
/* number of CUDA devices usable by this library (compute capability 10.x); 0 if none. */
int ptk_device_count(void);

/* ---- K0: undistortion with the reference's reshape quirk ----------------------------
 * Replaces match2D.py:846-863 (== matchND.py:473-490): cv2.undistortImagePoints on
 * rods.reshape(2,-1) plus the Python write-back loop.
 * cam[P,Nmax,2(end),2(xy)], n[P] -> und[P,Nmax,2,2]; rows >= n[p] are not written. */
int ptk_undistort_batch(const ptk_rig* rig, int P, int Nmax,
                        const double* cam1, const double* cam2,
                        const int32_t* n1, const int32_t* n2,
                        double* und1, double* und2, void* stream);

/* ---- K0+K1: cost matrix ---------------------------------------------------------------
 * Replaces match2D.py:846-932/936-939 (all cam1 x cam2 rod pairs x 4 end-point combos ->
 * cv2.triangulatePoints -> 2x cv2.projectPoints -> error -> cost, choice).
 * cost[P,Nmax,Nmax] (row = cam1 rod); choice (uint8, optional, may be NULL) = (e0+e3 <= e1+e2).
 * Optional per-combo outputs for matchND.create_weights (matchND.py:549-551), may be NULL:
 *   p_world[P,Nmax,Nmax,4,3] world-transformed points, rep_errs[P,Nmax,Nmax,4,2].
 * workspace: ptk_workspace_bytes(PTK_WS_COST, P, Nmax). */
int ptk_cost_batch(const ptk_rig* rig, int P, int Nmax,
                   const double* cam1, const double* cam2,
                   const int32_t* n1, const int32_t* n2,
                   double* cost, uint8_t* choice, double* p_world, double* rep_errs,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ---- K2: batched rectangular linear-sum assignment ----------------------------------
 * Replaces scipy.optimize.linear_sum_assignment at match2D.py:933 and tracking.py:122,
 * with SciPy's scan order and tie rules so that results are identical
 * also on tied costs.  cost[P, ld_rows, ld_cols] (only the leading nr[p] x nc[p] block is
 * read; nr/nc may be NULL = ld_rows/ld_cols for every problem).
 * row_ind, col_ind[P, min(ld_rows,ld_cols)]: the first n_out[p] = min(nr,nc) entries are
 * SciPy's (row_ind sorted ascending); the rest are -1.  status[P]: PTK_PROBLEM_*. */
int ptk_lsap_batch(int P, int ld_rows, int ld_cols, const double* cost,
                   const int32_t* nr, const int32_t* nc,
                   int32_t* row_ind, int32_t* col_ind, int32_t* n_out, int32_t* status,
                   void* stream);

/* ---- K0..K3: the whole match_frame body for P problems ------------------------------
 * Replaces match2D.py:843-983 with renumber=True, per (frame, colour) problem:
 * undistort -> pairs -> DLT -> reproject -> cost -> LSAP -> 18-column rows, including the
 * loop-index quirk for N1 > N2 and the end-point reversal.
 *   rows[P,Nmax,18], match_cost[P,Nmax] (== costs, match2D.py:934), row_ind/col_ind[P,Nmax],
 *   n_out[P] = min(n1,n2), status[P].
 * Optional (NULL to skip): cost[P,Nmax,Nmax], choice[P,Nmax,Nmax].
 * Extension (no reference counterpart): keep_mask[P,Nmax] (uint8, may be NULL)
 * = match_cost <= max_repr_err; pass +inf to switch the threshold off.  Nothing else depends on it.
 * workspace: ptk_workspace_bytes(PTK_WS_MATCH, P, Nmax). */
int ptk_match_batch(const ptk_rig* rig, int P, int Nmax,
                    const double* cam1, const double* cam2,
                    const int32_t* n1, const int32_t* n2,
                    double* cost, uint8_t* choice,
                    int32_t* row_ind, int32_t* col_ind, int32_t* n_out, int32_t* status,
                    double* rows, double* match_cost,
                    double max_repr_err, uint8_t* keep_mask,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ---- K0+K3: rows for GIVEN assignments ------------------------------------------------
 * Replaces match2D.py:876-884,941-983 with renumber=False: the rod pairing is kept
 * (col_ind[p,m] = cam2 rod paired with cam1 rod m; the reference uses m itself after dropping
 * rows not present in both cameras, match2D.py:831-837) and only the end-point combination is
 * (re-)evaluated.  rows[P,Nmax,18]; pair_cost[P,Nmax] (optional) = min(e0+e3, e1+e2) of each
 * pair == the reference's `costs` for renumber=False (match2D.py:941-947).
 * workspace: ptk_workspace_bytes(PTK_WS_COST, P, Nmax). */
int ptk_rows_batch(const ptk_rig* rig, int P, int Nmax,
                   const double* cam1, const double* cam2,
                   const int32_t* n1, const int32_t* n2,
                   const int32_t* col_ind, const int32_t* n_out,
                   double* rows, double* pair_cost,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ---- glue: matched rows -> tracking layout ----------------------------------------------
 * Replaces tracking.py:97-98 (data[["x1","y1","z1"]] / [["x2","y2","z2"]] .reshape(F,-1,3)) for C
 * colours: rows[F,C,Nmax,18] (problem index f*C+c) -> ends[C,F,N,2,3], first N rows of each problem. */
int ptk_rows_to_ends(int F, int C, int N, int Nmax, const double* rows, double* ends, void* stream);

/* ---- K4 (+K2): tracking_global_assignment over all consecutive frame pairs ----------
 * Replaces tracking.py:97-123,136-137 for C colours at once.
 *   ends[C,F,N,2,3]: (x1,y1,z1),(x2,y2,z2) per rod in DataFrame row order.
 *   cost[C,F-1,N,N] optional (NULL: kept in workspace only); col_ind[C,F-1,N];
 *   total_cost[C,F-1]; status[C,F-1].
 * The cost arithmetic follows numpy's operation order exactly (bit-identical costs for
 * identical inputs; the reference's cost is massively tied, so this is what makes col_ind
 * reproducible).  workspace: ptk_workspace_bytes(PTK_WS_TRACK, C*(F-1), N). */
int ptk_track_batch(int C, int F, int N, const double* ends,
                    double* cost, int32_t* col_ind, double* total_cost, int32_t* status,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ---- K5: chain pass (extension; SURVEY.md App. D.3, no reference counterpart) --------
 * ids[c,0] = ids_in[c] (or arange(N) if ids_in is NULL); ids[c,f+1][col_ind[c,f][a]] = ids[c,f][a].
 * col_ind[C,F-1,N] -> track_id[C,F,N].  Entries of col_ind outside 0..N-1 (a failed assignment:
 * ptk_track_batch writes -1) break the chain: unreached positions get id -1 and -1 propagates. */
int ptk_chain_tracks(int C, int F, int N, const int32_t* col_ind, const int32_t* ids_in,
                     int32_t* track_id, void* stream);

/* Multi-GPU part of the chain pass: every rank chains its own frames from the identity
 * (ptk_chain_tracks with ids_in = NULL, last frame = next rank's first frame), the per-rank boundary
 * maps are all-gathered (maps[W,C,N]: position in rank r+1's first frame -> local id on rank r), and
 * this call composes them up to `rank` (start[C,N], written) and re-bases track_id[C,F,N] in place. */
int ptk_chain_rebase(int W, int C, int F, int N, int rank, const int32_t* maps, int32_t* start,
                     int32_t* track_id, void* stream);

/* ---- K6: matchND.create_weights ("next" row) --------------------------
 * Replaces matchND.py:141-217.  p_3D[N1,N2,4,3], p_prev[Np,2,3], rep_errs[N1,N2,4,2] ->
 * weights[Np,N1,N2] (= max(w) - w), choices[Np,N1,N2] (float64 0..3, as the reference).
 * workspace: ptk_workspace_bytes(PTK_WS_WEIGHTS, Np*N1*N2, 0). */
int ptk_create_weights(int Np, int N1, int N2, const double* p_3D, const double* p_prev,
                       const double* rep_errs, double* weights, double* choices,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- K7: end-point re-ordering ("next" row) ------------------------------
 * Replaces the numeric body of match2D.reorder_endpoints_csv (match2D.py:1055-1137): per particle,
 * frame after frame, the two end points are swapped when the swapped displacement to the previous,
 * already re-ordered frame is strictly smaller (match2D.py:1102).
 *   ends[C,F,N,2,3] (rows sorted by particle within each frame); pts2d[C,F,N,2(cam),2(end),2(xy)]
 *   optional -> out_ends, out_2d (same shapes), swapped[C,F,N] (uint8; frame 0 is never swapped),
 *   mincost[C,F-1,N] = min(comb1, comb2) (the reference's mincosts_T, transposed). */
int ptk_reorder_endpoints(int C, int F, int N, const double* ends, const double* pts2d,
                          double* out_ends, double* out_2d, uint8_t* swapped, double* mincost, void* stream);

/* ---- K8 / K9: stand-alone triangulation / projection ("next" row) ----------
 * ptk_triangulate_points replaces calibrate_cameras.project_points (calibrate_cameras.py:137-199):
 *   cv2.triangulatePoints(P1, P2, p1, p2) on the points as given (no undistortion), p[0:3]/p[3],
 *   optional world transform.  p1, p2 [n,2] -> out [n,3].
 * ptk_project_points replaces calibrate_cameras.reproject_points (calibrate_cameras.py:202-262):
 *   optional inverse world transform, then cv2.projectPoints into both cameras.  X [n,3] -> uv1, uv2 [n,2]. */
int ptk_triangulate_points(const ptk_rig* rig, long long n, const double* p1, const double* p2,
                           int to_world, double* out, void* stream);
int ptk_project_points(const ptk_rig* rig, long long n, const double* X, int from_world,
                       double* uv1, double* uv2, void* stream);

/* ---- K10: 3-index assignment ("next" row) ----------------------------------
 * Replaces matchND.npartite_matching (matchND.py:45-138; a 0/1 programme handed to pulp/CBC) for
 * the three-group case matchND.match_frame poses (matchND.py:562): B independent problems
 * weights[B,D0,D1,D2] (actual sizes n0,n1,n2[B], NULL = full), every axis <= 64.  Chooses triples
 * (i,j,k), every index of every axis used at most once, with maximal summed weight.
 *   sol[B,3,min(D0,D1,D2)]  index arrays per axis (the reference's np.where(x): sorted by i), -1 padded
 *   n_sol[B]                number of triples (min(n0,n1,n2) when all weights are >= 0)
 *   value[B], bound[B]      summed weight of sol and the proven upper bound
 *   status[B]               PTK_AP3_OK: proven optimal (bound - value within rounding);
 *                           PTK_AP3_GAP: work limits hit, sol is the best found and bound its bound;
 *                           PTK_AP3_INVALID: NaN/inf weights or sizes out of range; PTK_AP3_EMPTY
 *   stats[B,4] (optional)   bounding steps (dual descent + subgradient iterations), search nodes,
 *                           candidate triples, step that proved it (-1: the search did, or nothing)
 * max_steps (dual-descent steps before the subgradient phase, which adds up to 1200 iterations) and
 * node_limit (reduced-cost search) <= 0 select the defaults (60 / 2e6).
 * workspace: ptk_workspace_bytes(PTK_WS_AP3, B, 0). */
#define PTK_AP3_OK 0
#define PTK_AP3_GAP 1
#define PTK_AP3_INVALID 2
#define PTK_AP3_EMPTY 3
int ptk_npartite3_batch(int B, int D0, int D1, int D2, const double* weights,
                        const int32_t* n0, const int32_t* n1, const int32_t* n2,
                        int32_t* sol, int32_t* n_sol, double* value, double* bound,
                        int32_t* status, int32_t* stats, int max_steps, long long node_limit,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---- matchND.assign: a whole frame sequence, enqueue-only ------------------------------------
 * Replaces the frame loop of matchND.assign (matchND.py:318-360) for C colours at once: frame 0
 * through the match2D path (ptk_match_batch, mean error), every later frame f through
 * matchND.match_frame (matchND.py:372-632) against frame f-1's rows: K0 + K1 with the summed error
 * and per-combo outputs -> create_weights (matchND.py:141-217) -> 3-index assignment (K10) -> index
 * bookkeeping and rows (matchND.py:568-632).  All sizes are read on the device, so nothing
 * synchronises; frames are sequential by construction (5 launches per frame).
 *   cam1, cam2 [F,C,Nmax,2,2]; n1, n2 [F,C]; Nmax <= 64
 *   rows[F,C,Nmax,18] (NaN beyond n_rows), costs[F,C,Nmax], n_rows[F,C]
 *   status[F,C]: PTK_PROBLEM_* for frame 0; later frames PTK_ND_OK, PTK_ND_INVALID (NaN weights),
 *     PTK_ND_EMPTY (a camera without rods: match_frame returns None), PTK_ND_UNBALANCED (the
 *     reference's ValueError, matchND.py:592-598), PTK_ND_GAP (solver work limit: rows are the best
 *     matching found), PTK_ND_ABORTED (an earlier frame of this colour failed)
 *   solver_stats[C,4] (optional): summed descent steps, search nodes, candidates, problems that
 *     needed the search
 * workspace: ptk_matchnd_sequence_workspace_bytes(F, C, Nmax). */
#define PTK_ND_OK 0
#define PTK_ND_INVALID 1
#define PTK_ND_EMPTY 3
#define PTK_ND_UNBALANCED 6
#define PTK_ND_GAP 7
#define PTK_ND_ABORTED 8
size_t ptk_matchnd_sequence_workspace_bytes(int F, int C, int Nmax);
int ptk_matchnd_sequence(const ptk_rig* rig, int F, int C, int Nmax,
                         const double* cam1, const double* cam2, const int32_t* n1, const int32_t* n2,
                         double* rows, double* costs, int32_t* n_rows, int32_t* status,
                         int32_t* solver_stats, void* workspace, size_t workspace_bytes, void* stream);

/* ---- CSV writer (host; "next" row) ---------------------------------------
 * Replaces df_out.to_csv(...) at match2D.py:626-629 for the output layout of match_frame
 * (match2D.py:986-997): an unnamed RangeIndex column, `n_float_cols` float columns (Python repr()
 * formatting, NaN -> empty field), frame, color, particle, and `n_seen` columns that are all 1.
 * `header` is the full first line (starting with the empty index name, i.e. with a comma).
 * ptk_format_double is the float formatter alone (test hook; buf >= 32 bytes, returns the length). */
int ptk_csv_write_rods(const char* path, const char* header, long long n_rows, int n_float_cols,
                       const double* values, const int64_t* frame, const char* color,
                       const int64_t* particle, int n_seen);
int ptk_format_double(double x, char* buf);

/* ---- CSV reader (host; "next" row) ---------------------------------------
 * Replaces pd.read_csv(f_in, sep=",", index_col=0) at match2D.py:601-602, 1053 and matchND.py:317
 * with a memory-mapped, multi-threaded columnar parse.  Column kinds follow pandas' inference
 * (int64 / float64 with NA strings -> NaN / bool / string); floats go through the arithmetic of
 * pandas' default parser so values are bit-identical to what the reference reads.
 * ptk_csv_open parses the whole file (n_threads <= 0: all cores) and returns a handle that must be
 * closed with ptk_csv_close, also after a failure (ptk_csv_error says what went wrong).
 * ptk_csv_col_copy copies a column into dst: int64[n_rows] (PTK_CSV_INT64, PTK_CSV_BOOL as 0/1),
 * double[n_rows] (PTK_CSV_FLOAT64) or int32 codes[n_rows] into the column's dictionary
 * (PTK_CSV_STRING; -1 = missing).  Column 0 is the index column of the rods_df layout.
 * ptk_parse_double is the float parser alone (test hook; returns 1 if the field is a number). */
#define PTK_CSV_INT64 0
#define PTK_CSV_FLOAT64 1
#define PTK_CSV_STRING 2
#define PTK_CSV_BOOL 3
typedef struct ptk_csv ptk_csv;
int ptk_csv_open(const char* path, int n_threads, ptk_csv** out);
void ptk_csv_close(ptk_csv* h);
const char* ptk_csv_error(const ptk_csv* h);
long long ptk_csv_n_rows(const ptk_csv* h);
int ptk_csv_n_cols(const ptk_csv* h);
const char* ptk_csv_col_name(const ptk_csv* h, int col);
int ptk_csv_col_kind(const ptk_csv* h, int col);
int ptk_csv_col_copy(const ptk_csv* h, int col, void* dst);
int ptk_csv_col_dict_size(const ptk_csv* h, int col);
const char* ptk_csv_col_dict(const ptk_csv* h, int col, int k, int* len);
int ptk_parse_double(const char* s, int len, double* out);

/* ---- workspace sizing ---------------------------------------------------------------- */
#define PTK_WS_COST 1
#define PTK_WS_MATCH 2
#define PTK_WS_TRACK 3
#define PTK_WS_WEIGHTS 4
#define PTK_WS_AP3 5
size_t ptk_workspace_bytes(int kind, long long n_problems, int Nmax);

/* ---- the whole path on DEVICE buffers, one call ----------------------------------------
 * ptk_match_batch (whose row kernel also writes the tracking layout of ptk_rows_to_ends) -> ptk_track_batch ->
 * chain pass, enqueued on `stream`
 * (replaces the match2D.py:600-625 driver loop + tracking.py:70-138 for F frames x C colours;
 * problem index = f*C + c).  Outputs as in ptk_reconstruct_host but in device memory; tracking
 * (do_track != 0) expects n_out == Nmax for every problem, like tracking.py:97.  Only kernel
 * launches are enqueued (no allocation, no synchronisation), so the call can be captured in a
 * CUDA graph.  workspace: ptk_reconstruct_device_workspace_bytes(F, C, Nmax, do_track). */
size_t ptk_reconstruct_device_workspace_bytes(int F, int C, int Nmax, int do_track);
int ptk_reconstruct_device(const ptk_rig* rig, int F, int C, int Nmax,
                           const double* cam1, const double* cam2,
                           const int32_t* n1, const int32_t* n2,
                           double* rows, double* match_cost,
                           int32_t* row_ind, int32_t* col_ind, int32_t* n_out, int32_t* status,
                           double max_repr_err, uint8_t* keep_mask,
                           int do_track, int32_t* track_col, double* track_total,
                           int32_t* track_id, int32_t* track_status,
                           void* workspace, size_t workspace_bytes, void* stream);

/* ---- one rank's frame range of a sharded run, one call ------------------------
 * As ptk_reconstruct_device with tracking, for F own frames followed by n_halo (0 or 1) frames that
 * belong to the next rank: cam1/cam2/n1/n2 and rows/match_cost/row_ind/col_ind/n_out/status hold
 * F + n_halo frames.  The halo frame is matched again here (the matching is a pure function of its
 * inputs: same bits as on its owner), so the boundary pair (F-1, F) is tracked locally and no halo
 * exchange is needed (the "recompute" option of SURVEY.md 8e; replaces the per-step send/recv).
 * track_col[C,F+n_halo-1,N], track_total / track_status[C,F+n_halo-1]; track_id[C,F,N] are LOCAL ids
 * (relative to this rank's first frame); halo_map[C,N] (required if n_halo) = local id of every position
 * of the halo frame, i.e. the map from the next rank's local ids to this rank's.  Kernel launches
 * only: capturable in a CUDA graph.  workspace: ptk_reconstruct_shard_workspace_bytes(F, n_halo, C, Nmax). */
size_t ptk_reconstruct_shard_workspace_bytes(int F, int n_halo, int C, int Nmax);
int ptk_reconstruct_shard(const ptk_rig* rig, int F, int n_halo, int C, int Nmax,
                          const double* cam1, const double* cam2, const int32_t* n1, const int32_t* n2,
                          double* rows, double* match_cost, int32_t* row_ind, int32_t* col_ind,
                          int32_t* n_out, int32_t* status, double max_repr_err, uint8_t* keep_mask,
                          int32_t* track_col, double* track_total, int32_t* track_status,
                          int32_t* track_id, int32_t* halo_map,
                          void* workspace, size_t workspace_bytes, void* stream);

/* Root side of the sharded chain pass, after the final gather: block r (r = 0..W-1) starts at
 * blocks + r*block_stride bytes and holds rank r's local ids track_id[C,frames_per_rank[r],N] at byte
 * offset tid_off and its halo_map[C,N] at map_off.  Composes the maps (start_r = start_{r-1} o map_{r-1},
 * start_0 = identity) and re-bases every rank's ids in place, one kernel launch for all ranks; ids of a
 * broken chain stay -1.  frames_per_rank is a HOST array; frames_dev[W] device scratch for its copy. */
int ptk_chain_rebase_gathered(int W, int C, int N, const int32_t* frames_per_rank, void* blocks,
                              size_t block_stride, size_t tid_off, size_t map_off,
                              int32_t* frames_dev, void* stream);


/* ---- host-buffer pipeline (what bench.py's `e2e` and match_csv_complex call) ----------
 * A ptk_ctx owns a device, streams, pinned staging and device buffers sized on demand. */
typedef struct ptk_ctx ptk_ctx;
int ptk_ctx_create(int device, ptk_ctx** out);
void ptk_ctx_destroy(ptk_ctx* ctx);

/* match (+ optional track + chain) for a whole experiment held in HOST memory.
 *   cam1, cam2 [F,C,Nmax,2,2]; n1, n2 [F,C]   (problem index = f*C + c)
 * outputs (host; any may be NULL except rows/n_out/status):
 *   rows[F,C,Nmax,18], match_cost[F,C,Nmax], row_ind/col_ind[F,C,Nmax], n_out[F,C], status[F,C],
 *   keep_mask[F,C,Nmax]
 * if do_track != 0 (requires n_out == N for every problem, as tracking.py:97 does):
 *   track_col[C,F-1,N], track_total[C,F-1], track_id[C,F,N], track_status[C,F-1]
 * Copies are chunked over frames and overlapped with compute on separate streams.
 * Returns PTK_ERR_NUMERIC if any status != OK/EMPTY (outputs are still filled). */
int ptk_reconstruct_host(ptk_ctx* ctx, const ptk_rig* rig, int F, int C, int Nmax,
                         const double* cam1, const double* cam2,
                         const int32_t* n1, const int32_t* n2,
                         double* rows, double* match_cost,
                         int32_t* row_ind, int32_t* col_ind, int32_t* n_out, int32_t* status,
                         double max_repr_err, uint8_t* keep_mask,
                         int do_track, int32_t* track_col, double* track_total,
                         int32_t* track_id, int32_t* track_status);

/* The same with a COMPACT result: instead of rows[F,C,Nmax,18] (36.9 MB per 8000 x 32 problems) only what the
 * device computed travels back: ends6[F,C,Nmax,6] = x1 y1 z1 x2 y2 z2 and straight[F,C,Nmax] (1: the straight
 * end-point pairing won, 0: the crossed one, match2D.py:935-983) -- 12.5 MB.  The other 12 columns are functions of
 * these, of col_ind (required here) and of the caller's own cam1 / cam2: ptk_expand_rows_host rebuilds the
 * 18-column rows bit for bit.  Why: device-to-host writes are what bounds this path on a multi-GPU host
 * (profiles/r02g_copy_probe_8gpu.json: 75 GB/s per group of four GPUs, whatever the number of GPUs writing). */
int ptk_reconstruct_host_compact(ptk_ctx* ctx, const ptk_rig* rig, int F, int C, int Nmax,
                                 const double* cam1, const double* cam2,
                                 const int32_t* n1, const int32_t* n2,
                                 double* ends6, uint8_t* straight, double* match_cost,
                                 int32_t* row_ind, int32_t* col_ind, int32_t* n_out, int32_t* status,
                                 double max_repr_err, uint8_t* keep_mask,
                                 int do_track, int32_t* track_col, double* track_total,
                                 int32_t* track_id, int32_t* track_status);

/* rows[P,Nmax,18] from the compact result (match2D.py:963-983), HOST code on n_threads workers (0: all cores):
 * centre = (end1 + end2) / 2 and length = |end2 - end1| with the operations K3 performs (same bits), the 2D end
 * points of rod m of camera 1 and rod col_ind[m] of camera 2 from cam1 / cam2[P,Nmax,2,2], both reversed where
 * straight == 0; rows m >= n_out[p] are NaN. */
int ptk_expand_rows_host(long long P, int Nmax, const double* cam1, const double* cam2,
                         const double* ends6, const uint8_t* straight, const int32_t* col_ind,
                         const int32_t* n_out, double* rows, int n_threads);


/* ---- measurement helper: FP64 FMA peak of the device (roofline denominator) ----------
 * Runs a dependent-free DFMA loop on every SM and returns achieved FLOP/s in *flops
 * (2 flop per DFMA), timed with CUDA events on `stream`. */
int ptk_measure_fp64_peak(int iters, double* flops, void* stream);

/* Per-kernel timing for bench.py's roofline: when enabled, CUDA event pairs are recorded on the
 * launching stream around every kernel launch, by kind.  ptk_timing_read sums durations and counts
 * per kind into arrays of PTK_TIMING_KINDS entries (call after synchronising); off by default.
 * ptk_timing_enable: 0 off, 1 every launch, 2 the dominant kernel (kind 1) only -- two events per step instead of
 * twenty, for a timed region whose total is itself a reported figure.
 * kinds: 0 undistort (+ k_prep2), 1 cost (K1), 2 lsap (matching), 3 rows, 4 rows_to_ends, 5 track_cost,
 *        6 lsap (tracking), 7 chain (3 launches), 8 create_weights, 9 npartite3 */
#define PTK_TIMING_KINDS 10
/* Bounds-checked debug build (make -C particletracking_b200/csrc debug -> libptk_dbg.so, selected with
 * PTK_LIB=<path>): every data-derived index is checked on the device; violations are counted, not trapped.
 * enabled = 1 only in that build; count / first_* describe the log (site: see ptk_bounds.cuh call sites). */
int ptk_bounds_report(int* enabled, long long* count, int* first_site, long long* first_idx,
                      long long* first_bound, int reset);

void ptk_timing_enable(int on);
int ptk_timing_read(double* ms_by_kind, long long* launches_by_kind, int reset);

/* number of kernel launches enqueued by this library since the last call with reset != 0
 * (process-wide, atomic) -- bench.py's `gpu_launches`. */
long long ptk_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* PTK_H_ */
