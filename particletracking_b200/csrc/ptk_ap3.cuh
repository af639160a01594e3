// ptk_ap3.cuh -- K10 k_ap3: exact 3-index (axial) assignment, one CTA per problem.
//
// Replaces matchND.npartite_matching (matchND.py:45-138) for the three-group case that
// matchND.match_frame poses (:562): choose triples t = (i, j, k) from w[n0, n1, n2] such that every
// index of every axis is used at most once and the summed weight is maximal.  The reference writes
// this as a 0/1 programme with n0*n1*n2 variables and shells out to CBC (seconds per frame at 32
// rods); here the structure of the programme is used instead:
//
//  1. Dual block-coordinate descent.  The LP dual is  min sum(y0)+sum(y1)+sum(y2)  s.t.
//     y0[i]+y1[j]+y2[k] >= w[i,j,k], y >= 0.  With one block fixed, e.g. y0, the best other two
//     blocks are the duals of a 2-D assignment on c[j,k] = max(0, max_i (w[i,j,k] - y0[i])), so every
//     step is: a max-reduction over one axis (all threads), then a shortest-augmenting-path LSAP by
//     one warp (columns in registers, two per lane).  The upper bound never increases.
//  2. Every step's pairs are completed to a feasible solution by a second LSAP over the remaining
//     axis (lower bound).  On rod data the first step already closes the gap: the pair (j,k) that
//     triangulates a rod also prefers the right previous rod, so bound == value and the solution is
//     proven optimal without any search.
//  3. If the descent stalls with a gap, only triples whose dual slack y0+y1+y2-w is below the gap can
//     be part of a better solution; they are listed (sorted per i) and a depth-first search over the
//     first axis with slack-sum pruning finishes the proof.  Work limits (candidates, nodes) turn into
//     status PTK_AP3_GAP with the best solution found and its bound, never into a silent answer.
//
// Sizes: every axis <= 64 (w <= 2 MB per problem, L2-resident); B problems per launch (colours).
#pragma once
#include <cfloat>
#include <cstdint>

namespace ptk {

constexpr int AP3_MAXN = 64;
constexpr int AP3_THREADS = 1024;
constexpr int AP3_OK = 0, AP3_GAP = 1, AP3_INVALID = 2, AP3_EMPTY = 3;
constexpr int AP3_SG_ROUND = 300, AP3_SG_PATIENCE = 30;
constexpr double AP3_SG_DEFLECT = 0.7;
constexpr double AP3_SG_LAMBDA = 2.0;  // Polyak step factor at the start of a round (tools/ap3_proto.py: the sweep behind these three)

struct Ap3Cand {
    double slack;
    int16_t j, k;
    int32_t pad;
};

struct Ap3Args {
    const double* w;  // [B, D0, D1, D2]
    int D0, D1, D2;
    const int32_t* n0;  // [B] actual sizes (nullptr: D0 / D1 / D2)
    const int32_t* n1;
    const int32_t* n2;
    int32_t* sol;     // [B, 3, Dmin], -1 beyond n_sol; triples sorted by the first index
    int32_t* n_sol;   // [B]
    double* value;    // [B] summed weight of the returned triples
    double* bound;    // [B] proven upper bound
    int32_t* status;  // [B]
    int32_t* stats;   // [B, 4] descent steps, search nodes (saturating), candidates, proven-at-step (optional)
    Ap3Cand* cand;    // workspace [B, 2, cand_cap]
    int cand_cap;
    int max_steps;         // dual block-coordinate descent steps
    int max_sg;            // subgradient iterations after the descent has stalled
    long long node_limit;
};

struct Ap3Smem {
    double cneg[AP3_MAXN * AP3_MAXN];  // LSAP cost (negated benefit), row-major, leading dimension m
    double y[3][AP3_MAXN];             // duals per axis
    double ybest[3][AP3_MAXN];         // the duals behind the best bound so far (sum = ub)
    int gcnt[AP3_MAXN];                // how many of the step's pairs prefer index r of the fixed axis
    double dir[AP3_MAXN];              // deflected subgradient direction on the fixed axis
    int dir_valid;
    double uu[AP3_MAXN], vv[AP3_MAXN], sp[AP3_MAXN];
    double ub_hist[4];
    double red[AP3_THREADS / 32];
    int path[AP3_MAXN], col4row[AP3_MAXN], row4col[AP3_MAXN];
    int pair_p[AP3_MAXN], pair_q[AP3_MAXN];
    int16_t best[3][AP3_MAXN];
    double vals[AP3_MAXN];
    int16_t cur3[3][AP3_MAXN];
    uint8_t argr[AP3_MAXN * AP3_MAXN];  // argmax over the fixed axis behind every c[p,q] (255: none positive)
    int cnt[AP3_MAXN + 1], off[AP3_MAXN + 1], order[AP3_MAXN];
    // search
    double acc[AP3_MAXN + 1], usum[AP3_MAXN + 1], suffix[AP3_MAXN + 1];
    int pos[AP3_MAXN + 1];
    int16_t chj[AP3_MAXN + 1], chk[AP3_MAXN + 1];
    int n_best, n_pairs, flag, total_cand, improved, signif;
    double best_val, ub, curL, gg;
};

__device__ __forceinline__ double ap3_inf() { return __longlong_as_double(0x7ff0000000000000LL); }

// Min-cost square assignment on cost[m x m] (shared memory): shortest augmenting paths by one warp
// with duals uu (rows), vv (columns); lanes own columns lane and lane+32.  Ties go to the lowest
// column.  Result: col4row / row4col and the duals (cost[i][j] - uu[i] - vv[j] >= 0, = 0 on the
// assignment).
__device__ void ap3_lsap_augment(const double* __restrict__ cost, int m, Ap3Smem& s);

// Called by the whole CTA.  Row reduction first (every warp takes rows: uu[i] = min_j cost[i][j], row i
// keeps its arg-min column if no lower row wants it -- on rod data that already is the assignment),
// then warp 0 augments the rows left over.
__device__ void ap3_lsap(const double* __restrict__ cost, int m, Ap3Smem& s)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double INF = ap3_inf();
    for (int j = tid; j < AP3_MAXN; j += AP3_THREADS) {
        s.uu[j] = 0.0;
        s.vv[j] = 0.0;
        s.col4row[j] = -1;
        s.row4col[j] = 0x7fffffff;
    }
    __syncthreads();
    for (int i = warp; i < m; i += AP3_THREADS / 32) {
        double bv = lane < m ? cost[i * m + lane] : INF;
        int bj = lane;
        if (lane + 32 < m) {
            const double c1 = cost[i * m + lane + 32];
            if (c1 < bv) { bv = c1; bj = lane + 32; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
            if (ov < bv || (ov == bv && oj < bj)) { bv = ov; bj = oj; }
        }
        if (lane == 0) {
            s.uu[i] = bv;
            s.path[i] = bj;  // scratch: the row's arg-min column
            atomicMin(&s.row4col[bj], i);
        }
    }
    __syncthreads();
    for (int i = tid; i < m; i += AP3_THREADS)
        if (s.row4col[s.path[i]] == i) s.col4row[i] = s.path[i];
    __syncthreads();
    for (int j = tid; j < AP3_MAXN; j += AP3_THREADS)
        if (s.row4col[j] == 0x7fffffff) s.row4col[j] = -1;
    __syncthreads();
    if (warp == 0) ap3_lsap_augment(cost, m, s);
    __syncthreads();
}

__device__ void ap3_lsap_augment(const double* __restrict__ cost, int m, Ap3Smem& s)
{
    const int lane = threadIdx.x & 31;
    const double INF = ap3_inf();
    const int j0 = lane, j1 = lane + 32;
    const bool in0 = j0 < m, in1 = j1 < m;
    for (int cur = 0; cur < m; cur++) {
        if (s.col4row[cur] >= 0) continue;  // assigned by the row reduction (uniform)
        double sp0 = INF, sp1 = INF;
        unsigned long long SC = 0ull, SR = 0ull;
        int i = cur, sink = -1;
        double minv = 0.0;
        const double v0 = s.vv[j0], v1 = s.vv[j1];
        while (sink < 0) {
            SR |= 1ull << i;
            const double ui = s.uu[i];
            double c0 = INF, c1 = INF;
            if (in0 && !((SC >> j0) & 1ull)) {
                const double r = minv + cost[i * m + j0] - ui - v0;
                if (r < sp0) { sp0 = r; s.path[j0] = i; }
                c0 = sp0;
            }
            if (in1 && !((SC >> j1) & 1ull)) {
                const double r = minv + cost[i * m + j1] - ui - v1;
                if (r < sp1) { sp1 = r; s.path[j1] = i; }
                c1 = sp1;
            }
            double bv = c0;
            int bj = j0;
            if (c1 < bv) { bv = c1; bj = j1; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
                if (ov < bv || (ov == bv && oj < bj)) { bv = ov; bj = oj; }
            }
            if (!(bv < INF)) { sink = bj; minv = bv; break; }  // cannot happen for finite costs
            minv = bv;
            SC |= 1ull << bj;
            const int r4 = s.row4col[bj];
            if (r4 < 0) sink = bj;
            else i = r4;
        }
        if (in0) s.sp[j0] = sp0;
        if (in1) s.sp[j1] = sp1;
        __syncwarp();
        for (int r = lane; r < m; r += 32)
            if ((SR >> r) & 1ull) s.uu[r] += (r == cur) ? minv : (minv - s.sp[s.col4row[r]]);
        if (in0 && ((SC >> j0) & 1ull)) s.vv[j0] = v0 - (minv - sp0);
        if (in1 && ((SC >> j1) & 1ull)) s.vv[j1] = v1 - (minv - sp1);
        __syncwarp();
        if (lane == 0) {
            int j = sink;
            while (true) {
                const int i2 = s.path[j];
                s.row4col[j] = i2;
                const int t = s.col4row[i2];
                s.col4row[i2] = j;
                j = t;
                if (i2 == cur) break;
            }
        }
        __syncwarp();
    }
}

__device__ __forceinline__ double ap3_block_max(double v, Ap3Smem& s)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) s.red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = s.red[0];
    for (int k = 1; k < AP3_THREADS / 32; k++) r = fmax(r, s.red[k]);
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(AP3_THREADS) k_ap3(const __grid_constant__ Ap3Args a)
{
    extern __shared__ __align__(16) unsigned char ap3_smem_raw[];
    Ap3Smem& s = *reinterpret_cast<Ap3Smem*>(ap3_smem_raw);
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n[3] = {a.n0 ? a.n0[b] : a.D0, a.n1 ? a.n1[b] : a.D1, a.n2 ? a.n2[b] : a.D2};
    const long long st[3] = {(long long)a.D1 * a.D2, a.D2, 1};
    const double* __restrict__ w = a.w + (size_t)b * a.D0 * a.D1 * a.D2;
    const int Dmin = min(a.D0, min(a.D1, a.D2));
    int32_t* sol = a.sol + (size_t)b * 3 * Dmin;
    const int nmin = min(n[0], min(n[1], n[2]));
    const double INF = ap3_inf();

    for (int t = tid; t < 3 * Dmin; t += AP3_THREADS) sol[t] = -1;
    if (tid == 0) {
        a.n_sol[b] = 0;
        a.value[b] = 0.0;
        a.bound[b] = 0.0;
        if (a.stats) { a.stats[b * 4 + 0] = 0; a.stats[b * 4 + 1] = 0; a.stats[b * 4 + 2] = 0; a.stats[b * 4 + 3] = -1; }
    }
    if (n[0] < 0 || n[1] < 0 || n[2] < 0 || n[0] > a.D0 || n[1] > a.D1 || n[2] > a.D2) {
        if (tid == 0) a.status[b] = AP3_INVALID;
        return;
    }
    if (nmin == 0) {
        if (tid == 0) a.status[b] = AP3_EMPTY;
        return;
    }
    if (tid == 0) {
        s.best_val = -1.0;
        s.n_best = 0;
        s.ub = INF;
        s.flag = 0;
        s.dir_valid = 0;
        for (int q = 0; q < 4; q++) s.ub_hist[q] = INF;
    }
    for (int t = tid; t < 3 * AP3_MAXN; t += AP3_THREADS) s.y[t / AP3_MAXN][t % AP3_MAXN] = 0.0;
    __syncthreads();
    double eps = 0.0, wabs = 0.0;
    bool bad = false;

    int steps = 0, proven_at = -1;
    // mode 0: descent (the fixed block cycles); mode 1: subgradient on the fixed block r_sg (Polyak
    // steps towards the incumbent, halved after AP3_SG_PATIENCE steps without a better bound; the
    // block changes every AP3_SG_ROUND iterations).  The descent can stall at a non-optimal dual
    // when rows of the cube are identical (the detector's dummy rods); the subgradient phase does
    // not, and every iteration is the same reduction + 2-D assignment, so the incumbent heuristics
    // keep running on its duals.
    int mode = 0, r_sg = 0, sg_it = 0, since = 0;
    double lam = 1.0;
    const int max_iter = a.max_steps + a.max_sg;
    for (int step = 0; step < max_iter; step++) {
        // axis r is the fixed block, (p, q) the two that are re-optimised
        const int r = mode == 0 ? step % 3 : r_sg, p = r == 0 ? 1 : 0, q = r == 2 ? 1 : 2;
        const int np = n[p], nq = n[q], nr = n[r];
        const int m = max(np, nq);
        // ---- c[p,q] = max(0, max_r (w - y_r)), stored negated, padded with zeros to m x m
        for (int e = tid; e < m * m; e += AP3_THREADS) {
            const int iq = e % m, ip = e / m;
            double best = 0.0;
            int arg = 255;
            if (ip < np && iq < nq) {
                const double* base = w + ip * st[p] + iq * st[q];
#pragma unroll 4
                for (int ir = 0; ir < nr; ir++) {
                    const double wv = base[ir * st[r]];
                    if (step == 0) {
                        if (!(fabs(wv) < INF)) bad = true;
                        wabs = fmax(wabs, fabs(wv));
                    }
                    const double v = wv - s.y[r][ir];
                    if (v > best) { best = v; arg = ir; }
                }
            }
            s.cneg[e] = -best;
            s.argr[e] = (uint8_t)arg;
        }
        __syncthreads();
        if (step == 0) {  // the first reduction saw every weight: validate and scale here
            const double mx = ap3_block_max(bad ? INF : wabs, s);
            if (!(mx < INF)) {
                if (tid == 0) a.status[b] = AP3_INVALID;
                return;
            }
            eps = 16.0 * DBL_EPSILON * (double)nmin * fmax(mx, 1e-300);
        }
        ap3_lsap(s.cneg, m, s);
        if (warp == 0) {
            // duals of the max problem: a = -uu, b = -vv; shift so that both are >= 0 (possible
            // because every c >= 0); dummy rows / columns end at zero
            double amin = INF;
            for (int j = lane; j < m; j += 32) amin = fmin(amin, -s.uu[j]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) amin = fmin(amin, __shfl_xor_sync(0xffffffffu, amin, o));
            double sum = 0.0;
            for (int j = lane; j < AP3_MAXN; j += 32) {
                const double ya = j < np ? fmax(0.0, -s.uu[j] - amin) : 0.0;
                const double yb = j < nq ? fmax(0.0, -s.vv[j] + amin) : 0.0;
                s.y[p][j] = ya;
                s.y[q][j] = yb;
                sum += ya + yb + (j < nr ? s.y[r][j] : 0.0);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            for (int j = lane; j < AP3_MAXN; j += 32) s.gcnt[j] = 0;
            __syncwarp();
            if (lane == 0) {
                s.curL = sum;
                s.improved = sum < s.ub;
                s.signif = sum < s.ub - 1e-13 * fabs(sum);
                s.ub = fmin(s.ub, sum);
                int np_ = 0, nt = 0;
                unsigned long long used = 0ull;
                bool distinct = true;
                double v = 0.0;
                for (int ip = 0; ip < np; ip++) {
                    const int iq = s.col4row[ip];
                    if (iq < 0 || iq >= nq) continue;
                    s.pair_p[np_] = ip;
                    s.pair_q[np_] = iq;
                    np_++;
                    // the pair's own best partner on the fixed axis: if these are all different the
                    // pairs are a feasible solution as they stand (w = c + y up to rounding)
                    const int ar = s.argr[ip * m + iq];
                    if (ar == 255) continue;
                    s.gcnt[ar]++;
                    if ((used >> ar) & 1ull) distinct = false;
                    used |= 1ull << ar;
                    v += -s.cneg[ip * m + iq] + s.y[r][ar];
                    s.cur3[r][nt] = (int16_t)ar; s.cur3[p][nt] = (int16_t)ip; s.cur3[q][nt] = (int16_t)iq;
                    nt++;
                }
                s.n_pairs = np_;
                if (distinct && v > s.best_val) {
                    for (int x = 0; x < nt; x++) { s.best[0][x] = s.cur3[0][x]; s.best[1][x] = s.cur3[1][x]; s.best[2][x] = s.cur3[2][x]; }
                    s.n_best = nt;
                    s.best_val = v;
                }
            }
            __syncwarp();
            // subgradient 1 - gcnt over the fixed axis, deflected by the previous direction
            // (Camerini-Fratta-Maffioli: damps the zig-zag that otherwise leaves some cubes unconverged)
            double gg = 0.0;
            const double beta = s.dir_valid ? AP3_SG_DEFLECT : 0.0;
            for (int j = lane; j < nr; j += 32) {
                double g = 1.0 - (double)s.gcnt[j];
                if (beta != 0.0) g += beta * s.dir[j];
                s.dir[j] = g;
                gg += g * g;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) gg += __shfl_xor_sync(0xffffffffu, gg, o);
            if (lane == 0) s.gg = gg;
        }
        __syncthreads();
        if (s.improved)
            for (int t = tid; t < 3 * AP3_MAXN; t += AP3_THREADS) s.ybest[t / AP3_MAXN][t % AP3_MAXN] = s.y[t / AP3_MAXN][t % AP3_MAXN];
        // ---- complete the pairs over axis r: second assignment on max(w[r, pair], 0)
        const int K = s.n_pairs;
        if (s.ub - s.best_val <= eps) {
            // proven by the pairs themselves
        } else if (K > 0) {
            const int m2 = max(K, nr);
            for (int e = tid; e < m2 * m2; e += AP3_THREADS) {
                const int ir = e % m2, ik = e / m2;
                double v = 0.0;
                if (ik < K && ir < nr) v = fmax(0.0, w[ir * st[r] + s.pair_p[ik] * st[p] + s.pair_q[ik] * st[q]]);
                s.cneg[e] = -v;
            }
            __syncthreads();
            ap3_lsap(s.cneg, m2, s);
            if (warp == 0) {
                double val = 0.0;
                for (int ik = lane; ik < K; ik += 32) {
                    const int ir = s.col4row[ik];
                    if (ir >= 0 && ir < nr) val += -s.cneg[ik * m2 + ir];
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
                if (val > s.best_val) {  // uniform
                    __syncwarp();
                    if (lane == 0) {
                        int nb = 0;
                        for (int ik = 0; ik < K; ik++) {
                            const int ir = s.col4row[ik];
                            if (ir < 0 || ir >= nr) continue;
                            if (!(s.cneg[ik * m2 + ir] < 0.0)) continue;  // w <= 0: leave it out (zero-weight triples are re-added at the end)
                            s.best[r][nb] = (int16_t)ir;
                            s.best[p][nb] = (int16_t)s.pair_p[ik];
                            s.best[q][nb] = (int16_t)s.pair_q[ik];
                            nb++;
                        }
                        s.n_best = nb;
                        s.best_val = val;
                    }
                }
            }
        } else if (tid == 0 && s.best_val < 0.0) {
            s.best_val = 0.0;
            s.n_best = 0;
        }
        __syncthreads();
        steps = step + 1;
        // ---- stop rules (uniform: everything read from shared memory)
        const double gap = s.ub - s.best_val;
        if (gap <= eps) { proven_at = step; break; }
        bool restore = false;
        if (mode == 0) {
            const double prev3 = s.ub_hist[(step + 1) & 3];  // bound three steps ago (INF until then)
            __syncthreads();
            if (tid == 0) s.ub_hist[step & 3] = s.ub;
            __syncthreads();
            const bool stalled = step >= 5 && !(prev3 - s.ub > 1e-13 * fabs(s.ub) + eps);  // one full cycle without progress
            if (!stalled && step + 1 < a.max_steps) continue;
            if (a.max_sg <= 0) break;
            mode = 1; r_sg = r; lam = AP3_SG_LAMBDA; since = 0; sg_it = 0;  // this step doubles as the first subgradient iteration
        } else {
            since = s.signif ? 0 : since + 1;
            sg_it++;
            if (sg_it >= a.max_sg || lam < AP3_SG_LAMBDA / 1024.0 || gap <= 1e-10 * fabs(s.ub)) break;
            if (sg_it % AP3_SG_ROUND == 0) { r_sg = (r_sg + 1) % 3; lam = AP3_SG_LAMBDA; since = 0; restore = true; }
            else if (since >= AP3_SG_PATIENCE) { lam *= 0.5; since = 0; restore = true; }
        }
        if (restore || !(s.gg > 0.0)) {
            if (!restore) lam *= 0.5;  // every index of the fixed axis used once, yet a gap: rounding of c; shrink and go on
            if (restore && tid == 0) s.dir_valid = 0;
            __syncthreads();
            for (int t = tid; t < 3 * AP3_MAXN; t += AP3_THREADS) s.y[t / AP3_MAXN][t % AP3_MAXN] = s.ybest[t / AP3_MAXN][t % AP3_MAXN];
            __syncthreads();
            continue;
        }
        {
            const double tstep = lam * (s.curL - s.best_val) / s.gg;
            __syncthreads();
            for (int j = tid; j < nr; j += AP3_THREADS) s.y[r][j] = fmax(0.0, s.y[r][j] - tstep * s.dir[j]);
            if (tid == 0) s.dir_valid = 1;
            __syncthreads();
        }
    }
    if (proven_at < 0) {  // the search below works on the duals of the best bound
        __syncthreads();
        for (int t = tid; t < 3 * AP3_MAXN; t += AP3_THREADS) s.y[t / AP3_MAXN][t % AP3_MAXN] = s.ybest[t / AP3_MAXN][t % AP3_MAXN];
        __syncthreads();
    }

    int status = AP3_OK;
    long long nodes = 0;
    int total_cand = 0;
    if (proven_at < 0) {
        // ---- reduced-cost search: only triples with slack below the gap can improve the incumbent
        const double gap = s.ub - s.best_val;
        const double thr = gap + eps;
        Ap3Cand* c0 = a.cand + (size_t)b * 2 * a.cand_cap;
        Ap3Cand* c1 = c0 + a.cand_cap;
        for (int t = tid; t <= AP3_MAXN; t += AP3_THREADS) s.cnt[t] = 0;
        __syncthreads();
        const int tot = n[0] * n[1] * n[2];
        for (int e = tid; e < tot; e += AP3_THREADS) {
            const int k = e % n[2], j = (e / n[2]) % n[1], i = e / (n[2] * n[1]);
            const double sl = s.y[0][i] + s.y[1][j] + s.y[2][k] - w[i * st[0] + j * st[1] + k];
            if (sl < thr) atomicAdd(&s.cnt[i], 1);
        }
        __syncthreads();
        if (tid == 0) {
            int o = 0;
            for (int i = 0; i < n[0]; i++) { s.off[i] = o; o += s.cnt[i]; }
            s.off[n[0]] = o;
            s.total_cand = o;
        }
        __syncthreads();
        total_cand = s.total_cand;
        if (total_cand > a.cand_cap) {
            status = AP3_GAP;
        } else {
            for (int t = tid; t <= AP3_MAXN; t += AP3_THREADS) s.cnt[t] = 0;
            __syncthreads();
            for (int e = tid; e < tot; e += AP3_THREADS) {
                const int k = e % n[2], j = (e / n[2]) % n[1], i = e / (n[2] * n[1]);
                const double sl = s.y[0][i] + s.y[1][j] + s.y[2][k] - w[i * st[0] + j * st[1] + k];
                if (sl < thr) {
                    const int at = s.off[i] + atomicAdd(&s.cnt[i], 1);
                    c0[at].slack = fmax(sl, 0.0);
                    c0[at].j = (int16_t)j;
                    c0[at].k = (int16_t)k;
                }
            }
            __syncthreads();
            // rank sort inside every i (ascending slack, then (j,k)): makes the list deterministic
            for (int i = 0; i < n[0]; i++) {
                const int lo = s.off[i], L = s.cnt[i];
                for (int e = tid; e < L; e += AP3_THREADS) {
                    const Ap3Cand me = c0[lo + e];
                    const int mekey = me.j * AP3_MAXN + me.k;
                    int rank = 0;
                    for (int f = 0; f < L; f++) {
                        const Ap3Cand ot = c0[lo + f];
                        const int okey = ot.j * AP3_MAXN + ot.k;
                        rank += (ot.slack < me.slack || (ot.slack == me.slack && okey < mekey)) ? 1 : 0;
                    }
                    c1[lo + rank] = me;
                }
            }
            __syncthreads();
            if (tid == 0) {
                // depth order: fewest candidates first
                for (int i = 0; i < n[0]; i++) s.order[i] = i;
                for (int x = 1; x < n[0]; x++) {
                    const int v = s.order[x];
                    int y = x - 1;
                    while (y >= 0 && s.cnt[s.order[y]] > s.cnt[v]) { s.order[y + 1] = s.order[y]; y--; }
                    s.order[y + 1] = v;
                }
                s.suffix[n[0]] = 0.0;
                for (int d = n[0] - 1; d >= 0; d--) {
                    const int i = s.order[d];
                    double mn = s.y[0][i];  // leaving i unused costs its dual
                    if (s.cnt[i] > 0) mn = fmin(mn, c1[s.off[i]].slack);
                    s.suffix[d] = s.suffix[d + 1] + mn;
                }
                double ysum12 = 0.0;
                for (int j = 0; j < n[1]; j++) ysum12 += s.y[1][j];
                for (int k = 0; k < n[2]; k++) ysum12 += s.y[2][k];
                unsigned long long usedJ = 0ull, usedK = 0ull;
                double bestslack = gap;  // a better solution has total slack < gap
                const int D = n[0];
                int d = 0;
                s.acc[0] = 0.0;
                s.usum[0] = 0.0;
                s.pos[0] = 0;
                s.chj[0] = -1;
                bool limit = false;
                while (d >= 0) {
                    if (++nodes > a.node_limit) { limit = true; break; }
                    if (d == D) {
                        const double total = s.acc[d] + (ysum12 - s.usum[d]);
                        if (total < bestslack - eps) {
                            // candidate solution: take it if its true value beats the incumbent
                            double val = 0.0;
                            int nb = 0;
                            for (int x = 0; x < D; x++)
                                if (s.chj[x] >= 0) val += w[s.order[x] * st[0] + s.chj[x] * st[1] + s.chk[x]];
                            if (val > s.best_val) {
                                for (int x = 0; x < D; x++)
                                    if (s.chj[x] >= 0) {
                                        s.best[0][nb] = (int16_t)s.order[x];
                                        s.best[1][nb] = s.chj[x];
                                        s.best[2][nb] = s.chk[x];
                                        nb++;
                                    }
                                s.n_best = nb;
                                s.best_val = val;
                                bestslack = fmin(total, s.ub - val);
                            }
                        }
                        d--;
                        continue;
                    }
                    const int i = s.order[d];
                    if (s.chj[d] >= 0) {  // undo the choice this depth holds
                        usedJ &= ~(1ull << s.chj[d]);
                        usedK &= ~(1ull << s.chk[d]);
                        s.chj[d] = -1;
                    }
                    const int L = s.cnt[i], lo = s.off[i];
                    int t = s.pos[d];
                    bool went = false;
                    while (t < L) {
                        const Ap3Cand c = c1[lo + t];
                        if (s.acc[d] + c.slack + s.suffix[d + 1] >= bestslack) { t = L; break; }  // sorted: nothing later fits
                        if (!((usedJ >> c.j) & 1ull) && !((usedK >> c.k) & 1ull)) {
                            usedJ |= 1ull << c.j;
                            usedK |= 1ull << c.k;
                            s.chj[d] = c.j;
                            s.chk[d] = c.k;
                            s.pos[d] = t + 1;
                            s.acc[d + 1] = s.acc[d] + c.slack;
                            s.usum[d + 1] = s.usum[d] + s.y[1][c.j] + s.y[2][c.k];
                            went = true;
                            break;
                        }
                        t++;
                    }
                    if (!went && t == L) {  // last option: leave i unused
                        if (s.acc[d] + s.y[0][i] + s.suffix[d + 1] < bestslack) {
                            s.pos[d] = L + 1;
                            s.acc[d + 1] = s.acc[d] + s.y[0][i];
                            s.usum[d + 1] = s.usum[d];
                            went = true;
                        }
                    }
                    if (went) {
                        d++;
                        if (d < D) { s.pos[d] = 0; s.chj[d] = -1; }
                    } else {
                        d--;
                    }
                }
                s.flag = limit ? 1 : 0;
            }
            __syncthreads();
            if (s.flag) status = AP3_GAP;
        }
    }

    // ---- output: triples sorted by the first index, then completed with unused indices (weight
    // >= 0 only) up to min(n0, n1, n2) -- at an optimum these can only be zero-weight triples
    __syncthreads();
    if (tid == 0) {
        const int nb = s.n_best;
        unsigned long long u0 = 0ull, u1 = 0ull, u2 = 0ull;
        for (int x = 0; x < nb; x++) { u0 |= 1ull << s.best[0][x]; u1 |= 1ull << s.best[1][x]; u2 |= 1ull << s.best[2][x]; }
        int cnt_ = nb;
        for (int i = 0; i < n[0] && cnt_ < nmin; i++) {
            if ((u0 >> i) & 1ull) continue;
            for (int j = 0; j < n[1] && !((u0 >> i) & 1ull); j++) {
                if ((u1 >> j) & 1ull) continue;
                for (int k = 0; k < n[2]; k++) {
                    if ((u2 >> k) & 1ull) continue;
                    if (w[i * st[0] + j * st[1] + k] < 0.0) continue;
                    s.best[0][cnt_] = (int16_t)i; s.best[1][cnt_] = (int16_t)j; s.best[2][cnt_] = (int16_t)k;
                    cnt_++;
                    u0 |= 1ull << i; u1 |= 1ull << j; u2 |= 1ull << k;
                    break;
                }
            }
        }
        // insertion sort by first index
        for (int x = 1; x < cnt_; x++) {
            const int16_t i0 = s.best[0][x], i1 = s.best[1][x], i2 = s.best[2][x];
            int y = x - 1;
            while (y >= 0 && s.best[0][y] > i0) {
                s.best[0][y + 1] = s.best[0][y]; s.best[1][y + 1] = s.best[1][y]; s.best[2][y + 1] = s.best[2][y];
                y--;
            }
            s.best[0][y + 1] = i0; s.best[1][y + 1] = i1; s.best[2][y + 1] = i2;
        }
        s.n_best = cnt_;
    }
    __syncthreads();
    const int cnt_ = s.n_best;
    for (int x = tid; x < cnt_; x += AP3_THREADS) {
        s.vals[x] = w[s.best[0][x] * st[0] + s.best[1][x] * st[1] + s.best[2][x]];
        sol[0 * Dmin + x] = s.best[0][x];
        sol[1 * Dmin + x] = s.best[1][x];
        sol[2 * Dmin + x] = s.best[2][x];
    }
    __syncthreads();
    if (tid == 0) {
        double val = 0.0;
        for (int x = 0; x < cnt_; x++) val += s.vals[x];
        a.n_sol[b] = cnt_;
        a.value[b] = val;
        a.bound[b] = status == AP3_OK ? fmax(val, fmin(s.ub, val + eps)) : s.ub;
        a.status[b] = status;
        if (a.stats) {
            a.stats[b * 4 + 0] = steps;
            a.stats[b * 4 + 1] = (int32_t)(nodes > 0x7fffffffLL ? 0x7fffffff : nodes);
            a.stats[b * 4 + 2] = total_cand;
            a.stats[b * 4 + 3] = proven_at;
        }
    }
}

// ------------------------------------------------------------------------------------- K11 / K12
// matchND.match_frame around the solver, batched over the colours of one frame with every size read
// from device memory, so that a whole frame sequence (matchND.assign, matchND.py:318-360) is enqueued
// without a host round trip:
//   k_nd_weights_raw / _finish : create_weights (matchND.py:141-217) from the previous frame's rows
//   k_nd_rows                  : index bookkeeping and rows (matchND.py:568-632)
constexpr int ND_OK = 0, ND_INVALID = 1, ND_EMPTY = 3, ND_UNBALANCED = 6, ND_GAP = 7, ND_ABORTED = 8;
__device__ __forceinline__ bool nd_usable(int st) { return st == ND_OK || st == ND_GAP; }

struct NdArgs {
    int D;                      // padded rods per problem (Nmax <= 64)
    const double* prev_rows;    // [C, D, 18] rows of frame f-1 (x1..z2 = columns 0..5)
    const int32_t* prev_n;      // [C]
    const int32_t* prev_status; // [C]
    const int32_t* n1;          // [C] frame f
    const int32_t* n2;
    const double* cam1;         // [C, D, 2, 2] frame f (distorted detections, as read)
    const double* cam2;
    const double* p_world;      // [C, D, D, 4, 3]
    const double* rep_errs;     // [C, D, D, 4, 2]
    const double* cost;         // [C, D, D]
    double* w;                  // [C, D, D, D]
    uint8_t* choices;           // [C, D, D, D]
    unsigned long long* wkey;   // [C] ordered key of the running maximum (zeroed by the caller)
    const int32_t* sol;         // [C, 3, D]
    const int32_t* n_sol;       // [C]
    const int32_t* ap3_status;  // [C]
    double* rows;               // [C, D, 18] frame f
    double* costs;              // [C, D]
    int32_t* n_rows;            // [C]
    int32_t* status;            // [C]
};

__device__ __forceinline__ double nd_weight(const double* q1, const double* q2, const double* p, const double* r, int& am)
{
    const double *c1p1 = p, *c2p1 = p + 3, *c2p2 = p + 6, *c1p2 = p + 9;  // matchND.py:184-187
    const double c1_re = ((r[0] + r[1]) + r[6]) + r[7];                   // combos 0 and 3
    const double c2_re = r[2] + r[3];                                     // combo 1 only: the 1:2 slice
    const double s0 = (dist3(c1p1, q1) + dist3(c1p2, q2)) * c1_re;
    const double s1 = (dist3(c2p1, q1) + dist3(c2p2, q2)) * c2_re;
    const double s2 = (dist3(c2p1, q2) + dist3(c2p2, q1)) * c2_re;
    const double s3 = (dist3(c1p1, q2) + dist3(c1p2, q1)) * c1_re;
    am = 0;
    double best = s0;
    if (s1 < best) { best = s1; am = 1; }
    if (s2 < best) { best = s2; am = 2; }
    if (s3 < best) { best = s3; am = 3; }
    return best;
}

__global__ void __launch_bounds__(256) k_nd_weights_raw(const __grid_constant__ NdArgs a)
{
    __shared__ double red[8];
    const int c = blockIdx.y, D = a.D;
    const int np = nd_usable(a.prev_status[c]) ? a.prev_n[c] : 0, n1 = a.n1[c], n2 = a.n2[c];
    const int tot = np * n1 * n2;
    double mine = -INFINITY;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += gridDim.x * blockDim.x) {
        const int k = e % n2, j = (e / n2) % n1, i = e / (n2 * n1);
        const double* q1 = a.prev_rows + ((size_t)c * D + i) * 18;
        const size_t jk = ((size_t)c * D + j) * D + k;
        int am;
        const double v = nd_weight(q1, q1 + 3, a.p_world + jk * 12, a.rep_errs + jk * 8, am);
        const size_t o = (((size_t)c * D + i) * D + j) * D + k;
        a.w[o] = v;
        a.choices[o] = (uint8_t)am;
        mine = fmax(mine, v);  // NaN-ignoring here; NaN weights are caught by the solver's validation
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine = fmax(mine, __shfl_xor_sync(kFull, mine, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mine;
    __syncthreads();
    if (threadIdx.x == 0 && tot > 0) {
        double m = red[0];
        for (int k = 1; k < 8; k++) m = fmax(m, red[k]);
        if (m > -INFINITY) atomicMax(a.wkey + c, ordered_key(m));
    }
}

__global__ void __launch_bounds__(256) k_nd_weights_finish(const __grid_constant__ NdArgs a)
{
    const int c = blockIdx.y, D = a.D;
    const int np = nd_usable(a.prev_status[c]) ? a.prev_n[c] : 0, n1 = a.n1[c], n2 = a.n2[c];
    const int tot = np * n1 * n2;
    const double wmax = ordered_value(a.wkey[c]);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += gridDim.x * blockDim.x) {
        const int k = e % n2, j = (e / n2) % n1, i = e / (n2 * n1);
        const size_t o = (((size_t)c * D + i) * D + j) * D + k;
        a.w[o] = wmax - a.w[o];  // matchND.py:216
    }
}

__global__ void __launch_bounds__(64) k_nd_rows(const __grid_constant__ NdArgs a)
{
    __shared__ int idx[3][AP3_MAXN];
    __shared__ int st_s;
    const int c = blockIdx.x, D = a.D, tid = threadIdx.x;
    const int np = a.prev_n[c], n1 = a.n1[c], n2 = a.n2[c];
    double* rows = a.rows + (size_t)c * D * 18;
    double* costs = a.costs + (size_t)c * D;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int t = tid; t < D * 18; t += blockDim.x) rows[t] = nan;
    for (int t = tid; t < D; t += blockDim.x) costs[t] = nan;
    if (tid == 0) {
        int st = ND_OK;
        if (!nd_usable(a.prev_status[c])) st = ND_ABORTED;
        else if (np <= 0 || n1 <= 0 || n2 <= 0) st = ND_EMPTY;  // match_frame returns None (matchND.py:465-467)
        else if (a.ap3_status[c] == AP3_INVALID) st = ND_INVALID;
        else {
            const int ns = a.n_sol[c];
            const int32_t* sol = a.sol + (size_t)c * 3 * D;
            // matchND.py:570-583: idx_out[:, rod] = triples; rods without a triple take the indices
            // (< nprev) that no triple uses, in increasing order
            unsigned long long u1 = 0ull, u2 = 0ull;
            for (int r = 0; r < np; r++) { idx[0][r] = r; idx[1][r] = -1; idx[2][r] = -1; }
            for (int x = 0; x < ns; x++) {
                const int r = sol[x];
                idx[1][r] = sol[D + x];
                idx[2][r] = sol[2 * D + x];
                u1 |= 1ull << sol[D + x];
                u2 |= 1ull << sol[2 * D + x];
            }
            if (ns < np) {
                int f1 = 0, f2 = 0, miss = 0, free1 = 0, free2 = 0;
                for (int r = 0; r < np; r++) { free1 += !((u1 >> r) & 1ull); free2 += !((u2 >> r) & 1ull); }
                miss = np - ns;
                if (free1 != miss || free2 != miss) st = ND_UNBALANCED;  // numpy: shape mismatch in the fill-in
                else
                    for (int r = 0; r < np; r++) {
                        if (idx[1][r] >= 0) continue;
                        while ((u1 >> f1) & 1ull) f1++;
                        while ((u2 >> f2) & 1ull) f2++;
                        idx[1][r] = f1++;
                        idx[2][r] = f2++;
                    }
            }
            if (st == ND_OK)
                for (int r = 0; r < np; r++)
                    if (idx[1][r] >= n1 || idx[2][r] >= n2) st = ND_UNBALANCED;  // IndexError -> ValueError (matchND.py:592-598)
            if (st == ND_OK && a.ap3_status[c] == AP3_GAP) st = ND_GAP;
        }
        st_s = st;
        a.status[c] = st;
        a.n_rows[c] = (st == ND_OK || st == ND_GAP) ? np : 0;
    }
    __syncthreads();
    if (st_s != ND_OK && st_s != ND_GAP) return;
    for (int r = tid; r < np; r += blockDim.x) {
        const int i1 = idx[1][r], i2 = idx[2][r];
        const int i3 = a.choices[(((size_t)c * D + r) * D + i1) * D + i2];
        const double* pw = a.p_world + (((size_t)c * D + i1) * D + i2) * 12;
        const double* e1 = pw + 3 * i3;
        const double* e2 = pw + 3 * (3 - i3);
        double* o = rows + (size_t)r * 18;
        for (int q = 0; q < 3; q++) {
            o[q] = e1[q];
            o[3 + q] = e2[q];
            o[6 + q] = (e1[q] + e2[q]) / 2.0;
        }
        o[9] = dist3(e1, e2);
        const double* r1 = a.cam1 + ((size_t)c * D + i1) * 4;
        const double* r2 = a.cam2 + ((size_t)c * D + i2) * 4;
        for (int q = 0; q < 4; q++) { o[10 + q] = r1[q]; o[14 + q] = r2[q]; }
        costs[r] = a.cost[((size_t)c * D + i1) * D + i2];
    }
}

// running totals of the solver statistics over the frames of a sequence: steps, nodes, candidates,
// and the number of problems that needed the search
__global__ void k_nd_stats(int C, const int32_t* __restrict__ stats, int32_t* __restrict__ total)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    total[c * 4 + 0] += stats[c * 4 + 0];
    total[c * 4 + 1] = (int32_t)min(0x7fffffffLL, (long long)total[c * 4 + 1] + stats[c * 4 + 1]);
    total[c * 4 + 2] = (int32_t)min(0x7fffffffLL, (long long)total[c * 4 + 2] + stats[c * 4 + 2]);
    total[c * 4 + 3] += stats[c * 4 + 3] < 0 ? 1 : 0;
}

}  // namespace ptk
