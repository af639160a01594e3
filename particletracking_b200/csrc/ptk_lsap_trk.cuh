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

// d1 is skipped only where a lower bound on it exceeds d0 by more than rounding could bridge: then
// min(d0, d1) == d0 exactly.  The bound on each of d1's two norms is (|dx| + |dy| + |dz|) / sqrt(3) <= |d|
// (Cauchy-Schwarz) -- two additions per norm where the largest coordinate difference cost two FP64 maxima
// (a compare and two selects each); which entries take the exact path changes, their values do not.
__device__ __forceinline__ bool trk_bound_decides(double p1a0, double p1a1, double p1a2, double q1a0, double q1a1, double q1a2,
                                                  double p2b0, double p2b1, double p2b2, double q2b0, double q2b1, double q2b2,
                                                  double d0)
{
    const double la = (fabs(p1a0 - q2b0) + fabs(p1a1 - q2b1)) + fabs(p1a2 - q2b2);
    const double lb = (fabs(p2b0 - q1a0) + fabs(p2b1 - q1a1)) + fabs(p2b2 - q1a2);
    return (la + lb) * 0.57735026918962 > d0;  // just below 1 / sqrt(3)
}
__device__ __forceinline__ double trk_entry(const double* __restrict__ cur, const double* __restrict__ nxt, double d0,
                                            int a, int b)
{
    const double* p1a = cur + 6 * a;
    const double* p2b = cur + 6 * b + 3;
    const double* q1a = nxt + 6 * a;
    const double* q2b = nxt + 6 * b + 3;
    if (trk_bound_decides(p1a[0], p1a[1], p1a[2], q1a[0], q1a[1], q1a[2], p2b[0], p2b[1], p2b[2], q2b[0], q2b[1], q2b[2], d0))
        return d0;
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
    bool bad = false, masked = false, anyexc = false;
    for (int s = 0; s < W; s++) {
        const int j = lane + 32 * s;
        const bool own = j < N;
        const int jj = own ? j : N - 1;
        const double Bj = Bs[jj];
        // the column's end points stay in registers across the rows
        const double p2b0 = cur[6 * jj + 3], p2b1 = cur[6 * jj + 4], p2b2 = cur[6 * jj + 5];
        const double q2b0 = nxt[6 * jj + 3], q2b1 = nxt[6 * jj + 4], q2b2 = nxt[6 * jj + 5];
        int xrow = -1;
        double xval = 0.0;
        for (int i = 0; i < N; i++) {
            const double d0 = As[i] + Bj;
            const double* p1a = cur + 6 * i;
            const double* q1a = nxt + 6 * i;
            const double c = trk_bound_decides(p1a[0], p1a[1], p1a[2], q1a[0], q1a[1], q1a[2], p2b0, p2b1, p2b2, q2b0, q2b1,
                                               q2b2, d0)
                                 ? d0
                                 : trk_cross_min(cur, nxt, d0, i, jj);
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
        anyexc |= __any_sync(kFull, own && xrow >= 0);
    }
    const bool anybad = __any_sync(kFull, bad);
    if (lane == 0) t.bad[0] = (anybad ? 1 : 0) | (anyexc ? 2 : 0) | (masked ? 4 : 0);
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

// ---- the solver: one warp per problem, WARPS problems per CTA, no block-level synchronisation.
//
// EXC = false: problems whose matrix is purely separable (no exception at all: every problem of a
// run without end-point flips).  EXC = true: the loop also checks the column's register exception and
// can re-evaluate masked entries through an out-of-line call.  Both live in ONE kernel (the first half
// of the grid takes the EXC problems, the second half the others; a warp whose problem belongs to the
// other half returns after one 4-byte load), so that the few slow EXC problems start first and overlap
// the bulk instead of forming a serial tail.
//
// The Dijkstra step is bound by the ALU pipe (16 lanes per scheduler: every integer / select / compare
// instruction costs two issue cycles; ncu r2a: ALU 81 % busy at 36 such instructions per step), so the
// step is written to keep integer work off it:
//  * keys are the RAW bits of the shortest-path cost: for non-negative doubles the signed integer
//    order of the bits is the value order, so the key costs nothing to form and the minimum nothing to
//    convert back.  One unsigned compare on the first REDUX result (>= 0x7ff00000) detects both rare
//    cases -- every remaining column at +inf (infeasible) and a negative candidate (possible through
//    rounding only) -- and the latter is redone with the order-preserving transform;
//  * a visited column's high word becomes that of +NaN (its real high word moves to `seen_hi`):
//    `r < NaN` is false (no update), NaN's bits are above +inf's (never the minimum), NaN == min is
//    false (never the winner) -- no "live" predicate anywhere in the loop.  Lanes beyond N start as NaN;
//  * the winner test is DSETP.EQ on the FP64 pipe (which also makes -0.0 tie with +0.0 as in SciPy);
//  * the next row is not decoded from the REDUX result: every assigned column keeps the shared-memory
//    address of its row in a register (0 = unassigned) and the winner's is fetched with one SHFL, as
//    is its position in SciPy's `remaining` list.
template <int CPL, bool EXC>
__device__ __forceinline__ void trk_solve(const TrkArgs& a, const void* scratch, long long p, int lane, unsigned char* base)
{
    typedef TrkRow<CPL> Row;
    const int N = a.N;
    Row* rows = reinterpret_cast<Row*>(base);
    int32_t* c4r = reinterpret_cast<int32_t*>(rows + N);
    const unsigned rows_addr = (unsigned)__cvta_generic_to_shared(rows);

    // ---- k_trk_prepare's record: A (and the mask words) go to shared memory beside u, the column side
    // (B, the first exception of the column) into registers
    int st;
    double Bj[CPL], xval[CPL];
    unsigned xrow_addr[CPL];
    {
        const TrkPrep t = trk_prep_at(const_cast<void*>(scratch), p, N);
        const int flags = t.bad[0];  // bit 0: invalid entries, bit 1: exceptions present, bit 2: masked ones too
        if (((flags >> 1) & 1) != (EXC ? 1 : 0)) return;  // warp-uniform: the other half's problem
        st = (flags & 1) ? PTK_PROBLEM_INVALID : PTK_PROBLEM_OK;
        const int W = trk_mask_words(N);
        for (int r = lane; r < N; r += 32) {
            Row& R = rows[r];
            R.u = 0.0;
            R.A = t.A[r];
#pragma unroll
            for (int s = 0; s < CPL; s++) R.xm[s] = (EXC && s < W) ? t.xm[(size_t)r * W + s] : 0u;
            c4r[r] = -1;
        }
#pragma unroll
        for (int s = 0; s < CPL; s++) {
            const int j = lane + 32 * s;
            const int jj = j < N ? j : N - 1;
            Bj[s] = t.B[jj];
            const int xr = (EXC && j < N) ? t.xrow[jj] : -1;
            xrow_addr[s] = xr >= 0 ? rows_addr + (unsigned)xr * (unsigned)sizeof(Row) : 0u;
            xval[s] = EXC ? t.xval[jj] : 0.0;
        }
    }
    __syncwarp();

    const int kDeadHi = 0x7ff80000;  // high word of +NaN: a visited (or absent) column
    double v[CPL];
    int spc_hi[CPL], spc_lo[CPL], seen_hi[CPL];  // shortest-path cost as two words (see above)
    unsigned path[CPL];                          // shared-memory address of the row the cost came from
    unsigned r4c_addr[CPL];                      // ... of the row the column is assigned to, 0 = unassigned
    int pos[CPL];
    unsigned ka[CPL], kb[CPL];                   // tie-rule candidate = ka * pos + kb (see below)
#pragma unroll
    for (int s = 0; s < CPL; s++) {
        v[s] = 0.0; spc_hi[s] = kDeadHi; spc_lo[s] = 0; seen_hi[s] = kDeadHi; path[s] = 0u; r4c_addr[s] = 0u; pos[s] = -1;
        // unassigned column: rank = 511 - pos, so that the LARGEST list position wins a tie
        ka[s] = 0u - (1u << 9);
        kb[s] = (511u << 9) | (unsigned)(lane + 32 * s);
    }

    if (st == PTK_PROBLEM_OK) {
#pragma unroll 1
        for (int curRow = 0; curRow < N; curRow++) {
#pragma unroll
            for (int s = 0; s < CPL; s++) {
                const int j = lane + 32 * s;
                spc_hi[s] = (j < N) ? 0x7ff00000 : kDeadHi;  // +inf
                spc_lo[s] = 0;
                seen_hi[s] = kDeadHi;                 // NaN = not visited in this row
                pos[s] = (j < N) ? (N - 1 - j) : -1;  // remaining[it] = nc-1-it  <=>  pos[j] = nc-1-j
            }
            double minVal = 0.0;
            unsigned iaddr = rows_addr + (unsigned)curRow * (unsigned)sizeof(Row);
            int num_rem = N, sink = -1;
            bool infeasible = false;
#pragma unroll 1
            while (sink == -1) {
                double ui, Ai;
                unsigned xmw[CPL];
                // (debug build) the row address comes out of a SHFL of per-column state: it must be a record of this problem
                if (!PTK_OK_IDX((long long)iaddr - (long long)rows_addr, (long long)N * (long long)sizeof(Row), 40) ||
                    !PTK_OK_IDX(((long long)iaddr - (long long)rows_addr) % (long long)sizeof(Row), 1, 41)) {
                    infeasible = true;
                    break;
                }
                {
                    unsigned long long w0, w1;
                    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "r"(iaddr));
                    ui = __longlong_as_double((long long)w0);
                    Ai = __longlong_as_double((long long)w1);
                    if (EXC) {
#pragma unroll
                        for (int s = 0; s < CPL; s++) asm volatile("ld.shared.b32 %0, [%1];" : "=r"(xmw[s]) : "r"(iaddr + 16u + 4u * s));
                    }
                }
                int mhi;
                unsigned mlo;
#pragma unroll
                for (int s = 0; s < CPL; s++) {
                    double c = Ai + Bj[s];
                    if (EXC) {
                        if (iaddr == xrow_addr[s]) c = xval[s];
                        if ((xmw[s] >> lane) & 1u)
                            c = trk_cross_min_q(a, p, c, (int)((iaddr - rows_addr) / (unsigned)sizeof(Row)), lane + 32 * s);
                    }
                    const double r = ((minVal + c) - ui) - v[s];
                    if (r < __hiloint2double(spc_hi[s], spc_lo[s])) {  // false for a visited column (NaN)
                        spc_hi[s] = __double2hiint(r);
                        spc_lo[s] = __double2loint(r);
                        path[s] = iaddr;
                    }
                    const int h = spc_hi[s];
                    const unsigned l = (unsigned)spc_lo[s];
                    if (s == 0 || h < mhi || (h == mhi && l < mlo)) { mhi = h; mlo = l; }
                }
                // warp argmin with SciPy's tie rule in three REDUX ops: (1) high and (2) low word of the
                // cost's bits, then (3) one packed integer whose order is the tie rule -- unassigned
                // columns first (largest list position wins), then assigned ones (smallest position wins).
                const int mh = __reduce_min_sync(kFull, mhi);
                if ((unsigned)mh >= 0x7ff00000u) {  // rare: +inf everywhere, or a negative candidate
                    if (mh >= 0) { infeasible = true; break; }
                    int h2 = 0x7fffffff;
                    unsigned l2 = 0xffffffffu;
#pragma unroll
                    for (int s = 0; s < CPL; s++) {
                        int h;
                        unsigned l;
                        skey(__hiloint2double(spc_hi[s], spc_lo[s]), h, l);
                        if (h < h2 || (h == h2 && l < l2)) { h2 = h; l2 = l; }
                    }
                    const int mh2 = __reduce_min_sync(kFull, h2);
                    const unsigned ml2 = __reduce_min_sync(kFull, h2 == mh2 ? l2 : 0xffffffffu);
                    minVal = skey_value(mh2, ml2);
                } else {
                    const unsigned ml = __reduce_min_sync(kFull, mhi == mh ? mlo : 0xffffffffu);
                    minVal = __hiloint2double(mh, (int)ml);
                }
                unsigned key3 = 0xffffffffu;
#pragma unroll
                for (int s = 0; s < CPL; s++) {
                    const unsigned cand = ka[s] * (unsigned)pos[s] + kb[s];
                    key3 = min(key3, (__hiloint2double(spc_hi[s], spc_lo[s]) == minVal) ? cand : 0xffffffffu);
                }
                key3 = __reduce_min_sync(kFull, key3);
                const int jsel = (int)(key3 & 511u);
                // the winner's row address (0: unassigned -> sink) and list position, straight from its lane
                unsigned naddr;
                int index;
                if (CPL == 1) {  // jsel < 32: the winner's lane
                    naddr = __shfl_sync(kFull, r4c_addr[0], jsel);
                    index = __shfl_sync(kFull, pos[0], jsel);
                } else {
                    naddr = lane_array_get<CPL>(r4c_addr, jsel);
                    index = lane_array_get<CPL>(pos, jsel);
                }
                if (naddr == 0u) sink = jsel; else iaddr = naddr;
                // remaining[index] = remaining[--num_remaining]: the column at the last position takes
                // the freed one; the selected column becomes visited (SC): its high word moves to `seen_hi`
                num_rem--;
#pragma unroll
                for (int s = 0; s < CPL; s++) {
                    if (lane + 32 * s == jsel) { seen_hi[s] = spc_hi[s]; spc_hi[s] = kDeadHi; }
                    else if (pos[s] == num_rem) pos[s] = index;
                }
            }
            if (infeasible) { st = PTK_PROBLEM_INFEASIBLE; break; }
            // dual update (see k_lsap): rows other than curRow are updated from their assigned column
            if (lane == 0) rows[curRow].u += minVal;
#pragma unroll
            for (int s = 0; s < CPL; s++) {
                if (seen_hi[s] != kDeadHi) {  // visited in this row
                    const double d = minVal - __hiloint2double(seen_hi[s], spc_lo[s]);
                    if (r4c_addr[s] != 0u) {
                        Row* R = rows + (r4c_addr[s] - rows_addr) / (unsigned)sizeof(Row);
                        R->u += d;
                    }
                    v[s] -= d;
                }
            }
            __syncwarp();
            // augment along the path
            int j = sink;
            while (true) {
                const unsigned raddr = lane_array_get<CPL>(path, j);
                const int r = (int)((raddr - rows_addr) / (unsigned)sizeof(Row));
                if (!PTK_OK_IDX(r, N, 42) || !PTK_OK_IDX(j, N, 43)) break;  // (debug build) a corrupt augmenting path
#pragma unroll
                for (int s = 0; s < CPL; s++)
                    if (lane == (j & 31) && s == (j >> 5)) {
                        r4c_addr[s] = raddr;
                        // assigned column: rank = 512 + pos, the SMALLEST list position wins a tie
                        ka[s] = 1u << 9;
                        kb[s] = (512u << 9) | (unsigned)j;
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
                if (!PTK_OK_IDX(cj, N, 44)) continue;
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

// grid = 2 * ceil(Q / WARPS): the first half solves the problems with exceptions, the second the rest
template <int CPL, int WARPS, int MINB>
__global__ void __launch_bounds__(32 * WARPS, MINB) k_lsap_trk(const __grid_constant__ TrkArgs a, const void* scratch)
{
    extern __shared__ __align__(16) unsigned char trk_smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned half = gridDim.x >> 1;
    const bool exc = blockIdx.x < half;
    const long long p = (long long)(exc ? blockIdx.x : blockIdx.x - half) * WARPS + warp;
    if (p >= a.Q) return;
    unsigned char* base = trk_smem + (size_t)warp * trk_warp_bytes<CPL>(a.N);
    if (exc) trk_solve<CPL, true>(a, scratch, p, lane, base);
    else trk_solve<CPL, false>(a, scratch, p, lane, base);
}

}  // namespace ptk
