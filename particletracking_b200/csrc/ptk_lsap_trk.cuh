// ptk_lsap_trk.cuh -- K2T: the tracking assignment (tracking.py:100-123) without a cost matrix.
//
// tracking_global_assignment's cost is  cost[a,b] = min(d0, d1)  with
//     d0 = |p1[f+1,a] - p1[f,a]| + |p2[f+1,b] - p2[f,b]|  =  A[a] + B[b]      (separable)
//     d1 = |p1[f,a] - p2[f+1,b]| + |p2[f,b] - p1[f+1,a]|                      (cross term)
// and d0 wins everywhere except where an end point of rod a lies next to the *other* end point of
// rod b one frame later (in practice: a == b after the detector flipped a rod's end points).  A
// separable matrix makes every permutation optimal up to rounding, so SciPy's answer is decided by
// the last bits of its duals, takes n(n+1)/2 Dijkstra steps per problem, and can only be reproduced
// by running SciPy's algorithm on bit-identical costs.  What it does NOT need is the matrix:
//
//  * A lives beside the row dual u in one 16-byte shared-memory record (one broadcast LDS.128 per
//    step fetches both), B[j] lives in the register of the lane that owns column j; the entry is
//    one DADD -- the very operation numpy performed, hence the same bits;
//  * the exceptions (entries where min(d0, d1) != d0) are found once per problem with a cheap lower
//    bound on d1 (6 subtractions, no square root; d1 itself only where the bound does not decide):
//    the first one of a column is kept in two registers (row index, value), any further ones set a
//    bit in a per-row mask word fetched with the same record and are re-evaluated on demand;
//  * per problem that is 44 N bytes of shared memory instead of 8 N^2 + state, so occupancy is set
//    by registers (12 warps per scheduler at N <= 32), not by shared memory (6 before), and
//    k_track_cost and its 65 MB matrix are gone for every N.
//
// The Dijkstra step itself is the k_lsap one (ptk_kernels.cuh) trimmed for issue slots, which is
// what bounds it: an order-preserving *signed* integer key (2 LOP3 + 1 SHF each way, an involution),
// the tie-rule candidate as ONE IMAD from two per-column constants that only change on
// augmentation, the next row's shared-memory address straight from the REDUX result.
#pragma once
#include "ptk_kernels.cuh"

namespace ptk {

struct TrkArgs {
    int N, F;
    long long Q;            // problems = C * (F - 1), problem q = c * (F - 1) + f
    const double* ends;     // [C,F,N,2,3]
    int32_t* col_ind;       // [Q,N]
    double* total;          // [Q]
    int32_t* status;        // optional [Q]
};

// One tracking-cost entry from the two frames in global memory (L1 resident), bit-identical to
// k_track_cost: np.min(d0, d1) with NaN propagation.  Out of line: it is only reached where the
// cheap bound below does not decide (a handful of entries per problem), and its two square roots
// must not set the register budget of the Dijkstra loop.
__device__ __noinline__ double trk_cross_min(const double* __restrict__ cur, const double* __restrict__ nxt, double d0,
                                             int a, int b)
{
    const double* p1a = cur + 6 * a;
    const double* p2b = cur + 6 * b + 3;
    const double* q1a = nxt + 6 * a;
    const double* q2b = nxt + 6 * b + 3;
    const double d1 = norm3(p1a[0], p1a[1], p1a[2], q2b[0], q2b[1], q2b[2]) +
                      norm3(p2b[0], p2b[1], p2b[2], q1a[0], q1a[1], q1a[2]);
    return (d0 != d0) ? d0 : ((d1 != d1) ? d1 : (d1 < d0 ? d1 : d0));
}

// d1 is skipped only where a lower bound on it (largest coordinate difference of each of its two
// norms) exceeds d0 by more than rounding could bridge: then min(d0, d1) == d0 exactly.
__device__ __forceinline__ double trk_entry(const double* __restrict__ cur, const double* __restrict__ nxt, double d0,
                                            int a, int b)
{
    const double* p1a = cur + 6 * a;
    const double* p2b = cur + 6 * b + 3;
    const double* q1a = nxt + 6 * a;
    const double* q2b = nxt + 6 * b + 3;
    const double la = fmax(fmax(fabs(p1a[0] - q2b[0]), fabs(p1a[1] - q2b[1])), fabs(p1a[2] - q2b[2]));
    const double lb = fmax(fmax(fabs(p2b[0] - q1a[0]), fabs(p2b[1] - q1a[1])), fabs(p2b[2] - q1a[2]));
    if ((la + lb) * 0.99999999999999 > d0) return d0;
    return trk_cross_min(cur, nxt, d0, a, b);
}

// The same from inside the Dijkstra loop (a masked exception): the frame pointers are recomputed from
// the problem index instead of being kept live across the loop (the loop runs in 40 registers).
__device__ __noinline__ double trk_cross_min_q(const TrkArgs& a, long long p, double d0, int i, int j)
{
    const int cc = (int)(p / (a.F - 1)), ff = (int)(p - (long long)cc * (a.F - 1));
    const double* cur = a.ends + ((size_t)cc * a.F + ff) * a.N * 6;
    return trk_cross_min(cur, cur + (size_t)a.N * 6, d0, i, j);
}

// order-preserving map double -> int64 (signed compare), an involution: negative values have their
// low 63 bits flipped
__device__ __forceinline__ void skey(double x, int& hi, unsigned& lo)
{
    const int h = __double2hiint(x);
    const int m = h >> 31;
    hi = h ^ (m & 0x7fffffff);
    lo = (unsigned)__double2loint(x) ^ (unsigned)m;
}
__device__ __forceinline__ double skey_value(int hi, unsigned lo)
{
    const int m = hi >> 31;
    return __hiloint2double(hi ^ (m & 0x7fffffff), (int)(lo ^ (unsigned)m));
}

// ---- per-problem record written by k_trk_prepare and read by k_lsap_trk (global memory, L2 resident)
//   double A[N] | double B[N] | double xval[N] | int32 xrow[N] | uint32 xm[N][W] | int32 bad, pad[3]
// W = mask words per row = ceil(N / 32).
__host__ __device__ inline int trk_mask_words(int N) { return (N + 31) / 32; }
__host__ __device__ inline size_t trk_prep_bytes(int N)
{
    return ((size_t)N * (3 * sizeof(double) + sizeof(int32_t) + sizeof(uint32_t) * trk_mask_words(N)) + 16 + 15) / 16 * 16;
}

struct TrkPrep {
    double* A;
    double* B;
    double* xval;
    int32_t* xrow;
    uint32_t* xm;
    int32_t* bad;
};
__host__ __device__ inline TrkPrep trk_prep_at(void* scratch, long long q, int N)
{
    unsigned char* b = static_cast<unsigned char*>(scratch) + (size_t)q * trk_prep_bytes(N);
    TrkPrep t;
    t.A = reinterpret_cast<double*>(b);
    t.B = t.A + N;
    t.xval = t.B + N;
    t.xrow = reinterpret_cast<int32_t*>(t.xval + N);
    t.xm = reinterpret_cast<uint32_t*>(t.xrow + N);
    t.bad = reinterpret_cast<int32_t*>(t.xm + (size_t)N * trk_mask_words(N));
    return t;
}

// K2T-prepare: one warp per problem (4 per CTA).  Stages the two frames in shared memory, computes the
// separable norms A, B (numpy's arithmetic: same bits as k_track_cost), then lane = column sweeps the
// rows in order with the lower-bound test, so "the first exception of a column" is well defined.
constexpr int kTrkPrepWarps = 4;
__global__ void __launch_bounds__(32 * kTrkPrepWarps) k_trk_prepare(const __grid_constant__ TrkArgs a, void* scratch)
{
    extern __shared__ __align__(16) double prep_smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long p = (long long)blockIdx.x * kTrkPrepWarps + warp;
    if (p >= a.Q) return;
    const int N = a.N, W = trk_mask_words(N);
    double* cur = prep_smem + (size_t)warp * 14 * N;  // cur[6N] | nxt[6N] | A[N] | B[N]
    double* nxt = cur + 6 * N;
    double* As = cur + 12 * N;
    double* Bs = As + N;
    const int cc = (int)(p / (a.F - 1)), ff = (int)(p - (long long)cc * (a.F - 1));
    const double* __restrict__ g = a.ends + ((size_t)cc * a.F + ff) * N * 6;
    for (int k = lane; k < 12 * N; k += 32) cur[k] = g[k];  // frames f and f+1 are contiguous
    __syncwarp();
    const TrkPrep t = trk_prep_at(scratch, p, N);
    for (int k = lane; k < 2 * N; k += 32) {
        const int r = k >> 1, e = (k & 1) * 3;  // rod r, end point 1 or 2
        const double v = norm3(nxt[6 * r + e], nxt[6 * r + e + 1], nxt[6 * r + e + 2],
                               cur[6 * r + e], cur[6 * r + e + 1], cur[6 * r + e + 2]);
        if (k & 1) { Bs[r] = v; t.B[r] = v; } else { As[r] = v; t.A[r] = v; }
    }
    __syncwarp();
    bool bad = false, masked = false;
    for (int s = 0; s < W; s++) {
        const int j = lane + 32 * s;
        const bool own = j < N;
        const int jj = own ? j : N - 1;
        const double Bj = Bs[jj];
        int xrow = -1;
        double xval = 0.0;
        for (int i = 0; i < N; i++) {
            const double d0 = As[i] + Bj;
            const double c = trk_entry(cur, nxt, d0, i, jj);
            bool flag = false;
            if (own) {
                bad |= (c != c) || (c == -INFINITY);  // SciPy: "matrix contains invalid numeric entries"
                if (__double_as_longlong(c) != __double_as_longlong(d0)) {
                    if (xrow < 0) { xrow = i; xval = c; }
                    else flag = true;
                }
            }
            const unsigned w = __ballot_sync(kFull, flag);
            masked |= w != 0u;
            if (lane == 0) t.xm[(size_t)i * W + s] = w;
        }
        if (own) { t.xrow[j] = xrow; t.xval[j] = xval; }
    }
    const bool anybad = __any_sync(kFull, bad);
    if (lane == 0) t.bad[0] = (anybad ? 1 : 0) | (masked ? 2 : 0);
}

template <int CPL>
struct TrkRow {
    double u;  // SciPy's row dual
    double A;  // |p1[f+1,a] - p1[f,a]|
    uint32_t xm[CPL <= 4 ? 4 : 8];  // bit `lane` of xm[s]: entry (row, lane + 32 s) is a masked exception
};

template <int CPL>
__host__ __device__ inline size_t trk_warp_bytes(int N)
{
    return (((size_t)N * (sizeof(TrkRow<CPL>) + sizeof(int32_t))) + 15) / 16 * 16;
}

// WARPS problems per CTA, one warp each; no block-level synchronisation anywhere.
// MASKED = false: problems without masked exceptions (at most one exception per column: all of them in
// practice); the others return at once and are solved by the MASKED = true instantiation, whose loop
// carries the out-of-line call -- and the spills around it -- that the common case must not pay for.
template <int CPL, int WARPS, int MINB, bool MASKED>
__global__ void __launch_bounds__(32 * WARPS, MINB) k_lsap_trk(const __grid_constant__ TrkArgs a, const void* scratch)
{
    extern __shared__ __align__(16) unsigned char trk_smem[];
    typedef TrkRow<CPL> Row;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long p = (long long)blockIdx.x * WARPS + warp;
    if (p >= a.Q) return;
    const int N = a.N;
    unsigned char* base = trk_smem + (size_t)warp * trk_warp_bytes<CPL>(N);
    Row* rows = reinterpret_cast<Row*>(base);
    int32_t* c4r = reinterpret_cast<int32_t*>(rows + N);

    // ---- k_trk_prepare's record: A and the mask words go to shared memory beside u, the column side
    // (B, the first exception of the column) into registers
    int st;
    double Bj[CPL], xval[CPL];
    int xrow[CPL];
    {
    const TrkPrep t = trk_prep_at(const_cast<void*>(scratch), p, N);
    const int W = trk_mask_words(N);
    for (int r = lane; r < N; r += 32) {
        Row& R = rows[r];
        R.u = 0.0;
        R.A = t.A[r];
#pragma unroll
        for (int s = 0; s < CPL; s++) R.xm[s] = (MASKED && s < W) ? t.xm[(size_t)r * W + s] : 0u;
        c4r[r] = -1;
    }
#pragma unroll
    for (int s = 0; s < CPL; s++) {
        const int j = lane + 32 * s;
        const int jj = j < N ? j : N - 1;
        Bj[s] = t.B[jj];
        xrow[s] = j < N ? t.xrow[jj] : -1;
        xval[s] = t.xval[jj];
    }
    const int flags = t.bad[0];  // bit 0: invalid entries, bit 1: masked exceptions present
    if (((flags >> 1) & 1) != (MASKED ? 1 : 0)) return;  // warp-uniform
    st = (flags & 1) ? PTK_PROBLEM_INVALID : PTK_PROBLEM_OK;
    }
    __syncwarp();

    double v[CPL], spc[CPL];
    int path[CPL], r4c[CPL], pos[CPL];
    unsigned ka[CPL], kb[CPL];  // tie-rule candidate = ka * pos + kb (see below)
#pragma unroll
    for (int s = 0; s < CPL; s++) {
        v[s] = 0.0; spc[s] = 0.0; path[s] = -1; r4c[s] = -1; pos[s] = -1;
        // unassigned column: rank = 511 - pos, so that the LARGEST list position wins a tie
        ka[s] = 0u - (1u << 18);
        kb[s] = (511u << 18) | (unsigned)(lane + 32 * s);
    }
    const int kInfHi = 0x7ff00000;

    if (st == PTK_PROBLEM_OK) {
#pragma unroll 1
        for (int curRow = 0; curRow < N; curRow++) {
#pragma unroll
            for (int s = 0; s < CPL; s++) {
                const int j = lane + 32 * s;
                spc[s] = INFINITY;
                pos[s] = (j < N) ? (N - 1 - j) : -1;  // remaining[it] = nc-1-it  <=>  pos[j] = nc-1-j
            }
            double minVal = 0.0;
            int i = curRow, num_rem = N, sink = -1;
            bool infeasible = false;
#pragma unroll 1
            while (sink == -1) {
                const Row& R = rows[i];
                const double ui = R.u, Ai = R.A;
                int khi[CPL], mhi;
                unsigned klo[CPL], mlo;
#pragma unroll
                for (int s = 0; s < CPL; s++) {
                    // branch-free: every lane evaluates, only columns still in `remaining` update
                    const bool live = pos[s] >= 0;
                    double c = Ai + Bj[s];
                    if (i == xrow[s]) c = xval[s];
                    if (MASKED && live && ((R.xm[s] >> lane) & 1u)) c = trk_cross_min_q(a, p, c, i, lane + 32 * s);
                    const double r = ((minVal + c) - ui) - v[s];
                    if (live && r < spc[s]) { spc[s] = r; path[s] = i; }
                    int h;
                    unsigned l;
                    skey(spc[s], h, l);
                    khi[s] = live ? h : 0x7fffffff;
                    klo[s] = live ? l : 0xffffffffu;
                    if (s == 0) { mhi = khi[0]; mlo = klo[0]; }
                    else if (khi[s] < mhi || (khi[s] == mhi && klo[s] < mlo)) { mhi = khi[s]; mlo = klo[s]; }
                }
                // warp argmin with SciPy's tie rule in three REDUX ops: (1) high and (2) low word of
                // the order-preserving key, then (3) one packed integer whose order is the tie rule
                // -- unassigned columns first (largest list position wins), then assigned ones
                // (smallest position wins) -- and whose low bits carry row4col+1 and the column.
                const int mh = __reduce_min_sync(kFull, mhi);
                const unsigned ml = __reduce_min_sync(kFull, mhi == mh ? mlo : 0xffffffffu);
                // all remaining columns at +inf (spc is never NaN: a NaN candidate fails r < spc)
                if (mh >= kInfHi) { infeasible = true; break; }
                unsigned key3 = 0xffffffffu;
#pragma unroll
                for (int s = 0; s < CPL; s++) {
                    const bool win = (khi[s] == mh) & (klo[s] == ml);
                    const unsigned cand = ka[s] * (unsigned)pos[s] + kb[s];
                    key3 = min(key3, win ? cand : 0xffffffffu);
                }
                key3 = __reduce_min_sync(kFull, key3);
                const int jsel = (int)(key3 & 511u);
                const int r4p1 = (int)((key3 >> 9) & 511u);  // row4col + 1
                int index;
                if (CPL == 1) {
                    index = __shfl_sync(kFull, pos[0], jsel);
                } else {
                    const unsigned rank = key3 >> 18;
                    index = rank < 512u ? (int)(511u - rank) : (int)(rank - 512u);
                }
                minVal = skey_value(mh, ml);
                if (r4p1 == 0) sink = jsel; else i = r4p1 - 1;
                // remaining[index] = remaining[--num_remaining]; -2 marks "visited" (SC)
                num_rem--;
#pragma unroll
                for (int s = 0; s < CPL; s++) {
                    const int ps = pos[s];
                    pos[s] = (ps == index) ? -2 : ((ps == num_rem) ? index : ps);
                }
            }
            if (infeasible) { st = PTK_PROBLEM_INFEASIBLE; break; }
            // dual update (see k_lsap): rows other than curRow are updated from their assigned column
            if (lane == 0) rows[curRow].u += minVal;
#pragma unroll
            for (int s = 0; s < CPL; s++) {
                if (pos[s] == -2) {
                    const double d = minVal - spc[s];
                    if (r4c[s] >= 0) rows[r4c[s]].u += d;
                    v[s] -= d;
                }
            }
            __syncwarp();
            // augment along the path
            int j = sink;
            while (true) {
                const int r = lane_array_get<CPL>(path, j);
#pragma unroll
                for (int s = 0; s < CPL; s++)
                    if (lane == (j & 31) && s == (j >> 5)) {
                        r4c[s] = r;
                        // assigned column: rank = 512 + pos, the SMALLEST list position wins a tie
                        ka[s] = 1u << 18;
                        kb[s] = (512u << 18) | ((unsigned)(r + 1) << 9) | (unsigned)j;
                    }
                const int t = c4r[r];
                __syncwarp();
                if (lane == 0) c4r[r] = j;
                j = t;
                if (r == curRow) break;
            }
            __syncwarp();
        }
    }

    // ---- outputs (pointers are formed here, after the loop, to keep them out of its live set)
    {
        const int cc = (int)(p / (a.F - 1)), ff = (int)(p - (long long)cc * (a.F - 1));
        const double* __restrict__ cur = a.ends + ((size_t)cc * a.F + ff) * N * 6;
        const double* __restrict__ nxt = cur + (size_t)N * 6;
        const TrkPrep t = trk_prep_at(const_cast<void*>(scratch), p, N);
        int32_t* ocol = a.col_ind + (size_t)p * N;
        if (st == PTK_PROBLEM_OK) {
            double tsum = 0.0;
            for (int r = lane; r < N; r += 32) {
                const int cj = c4r[r];
                ocol[r] = cj;
                tsum += trk_entry(cur, nxt, rows[r].A + t.B[cj], r, cj);
            }
            tsum = warp_sum(tsum);
            if (lane == 0) a.total[p] = tsum;
        } else {
            for (int r = lane; r < N; r += 32) ocol[r] = -1;
            if (lane == 0) a.total[p] = __longlong_as_double(0x7ff8000000000000LL);
        }
        if (lane == 0 && a.status) a.status[p] = st;
    }
}

}  // namespace ptk
