// ptk_kernels.cuh -- the kernels of the stereo hot path (sm_100a), K0..K6.
//
//  K0 k_undistort     cam[P,Nmax,2,2] -> und[P,Nmax,2,2]           (match2D.py:846-863, quirk-exact)
//  K0b k_prep2       canonical rigs: the camera-2 share of every DLT, once per camera-2 end point (records)
//  K1 k_cost_rows     (canonical rigs) / k_cost (general rigs): one thread per end-point combination, 4 lanes
//                     per rod pair: DLT null vector by inverse power iteration -> 2x reproject -> quad
//                     shuffle -> cost[P,Nmax,Nmax] (+choice, +per-combo extras)  (match2D.py:865-939)
//  K2 k_lsap          one warp per problem, SciPy-exact rectangular LSAP     (match2D.py:933, tracking.py:122)
//  K3 k_rows          re-evaluates the n matched pairs and writes the 18-column rows (match2D.py:963-983)
//  K4 k_track_cost    cost[Q,N,N] over all consecutive frame pairs, numpy-exact (tracking.py:97-121)
//  K5 k_chain_*       segmented permutation composition -> track ids (extension)
//  K6 k_weights_*     matchND.create_weights (matchND.py:141-217)
#pragma once
#include <limits.h>

#include "ptk_device.cuh"

namespace ptk {

constexpr unsigned kFull = 0xffffffffu;

// ------------------------------------------------------------------------------------- K0
// grid = (ceil(2*Nmax/128), P, 2 cameras); thread j handles pseudo-point j of problem p.
__global__ void __launch_bounds__(128) k_undistort(const __grid_constant__ DevRig rig, int Nmax,
                                                   const double* __restrict__ cam1, const double* __restrict__ cam2,
                                                   const int32_t* __restrict__ n1, const int32_t* __restrict__ n2,
                                                   double* __restrict__ und1, double* __restrict__ und2)
{
    const int p = blockIdx.y;
    const int camsel = blockIdx.z;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = camsel ? n2[p] : n1[p];
    if (n <= 0 || n > Nmax || j >= 2 * n) return;  // n > Nmax: the problem is reported INVALID by K2
    const double* in = (camsel ? cam2 : cam1) + (size_t)p * Nmax * 4;
    double* out = (camsel ? und2 : und1) + (size_t)p * Nmax * 4;
    const double* cm = camsel ? rig.CM2 : rig.CM1;
    const double* k = camsel ? rig.k2 : rig.k1;
    double ou, ov;
    if (n == 1) {
        // 2x2 input is read row-wise by OpenCV (proper points); the write-back transposes.
        undistort_point(cm, k, PTK_AT(in, 2 * j, 4 * Nmax, 1), PTK_AT(in, 2 * j + 1, 4 * Nmax, 2), ou, ov);
        PTK_AT(out, j, 4 * Nmax, 3) = ou;
        PTK_AT(out, j + 2, 4 * Nmax, 4) = ov;
    } else {
        // rods.reshape(2,-1): pseudo-point j = (flat[j], flat[j+2n])
        undistort_point(cm, k, PTK_AT(in, j, 4 * Nmax, 5), PTK_AT(in, j + 2 * n, 4 * Nmax, 6), ou, ov);
        PTK_AT(out, j, 4 * Nmax, 7) = ou;
        PTK_AT(out, j + 2 * n, 4 * Nmax, 8) = ov;
    }
}

// ------------------------------------------------------------------------------------- K1
// One thread per end-point combination; lanes 4q..4q+3 of a warp hold combos
// (e1,e2) = (0,0),(0,1),(1,0),(1,1) of one rod pair.  Two kernels share the arguments below:
//  * k_cost_rows (canonical rigs, the reference's): a CTA owns a block of the cost matrix, stages its inputs
//    (k_prep2's records, the points) in shared memory and loops over the block's pairs -- see its comment;
//  * k_cost (general rigs): a block of 128 threads covers 32 consecutive pairs (row-major i*N2+j) of one
//    problem, grid = (tiles, problems), inputs are 16-byte loads through the read-only path.
// Both are bound by instruction dispatch (an FP64 or integer-ALU instruction holds the scheduler's port for
// two cycles), so the non-FP64 work is counted in single instructions: the problem index is blockIdx.y (no
// division), the pair's row comes from one float reciprocal instead of an integer division (exact for
// q < 2^16, N2 <= 256, see pair_row), the rig class is a template parameter (no rig-constant branches).
struct CostArgs {
    int Nmax;
    int tiles;  // tiles per problem = ceil(Nmax*Nmax/32)
    const double* cam1;
    const double* cam2;
    const double* und1;
    const double* und2;
    const double2* rec2;  // canonical rigs: k_prep2's records of the camera-2 end points (else unused)
    const int32_t* n1;
    const int32_t* n2;
    double* cost;      // [P,Nmax,Nmax]
    uint8_t* choice;   // optional
    double* p_world;   // optional [P,Nmax,Nmax,4,3]
    double* rep_errs;  // optional [P,Nmax,Nmax,4,2]
};

__device__ __forceinline__ double2 ldg2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }

// The camera-2 records of the canonical-rig DLT (ptk_device.cuh: canon_prep2) in memory: per problem,
// kCanon2RecPlanes planes of 2 * Nmax double2, plane k holding fields (2k, 2k + 1) of end point e = 2 * rod + end.
// K1's prologue copies a chunk of every plane into shared memory (coalesced runs), where a thread's fourteen
// 16-byte reads share one address register (the plane is an immediate offset).
__host__ __device__ inline size_t rec_problem_stride(int Nmax) { return (size_t)kCanon2RecPlanes * 2 * Nmax; }  // double2

// K0b (canonical rigs): one thread per camera-2 end point, grid = (ceil(2*Nmax/128), P).
__global__ void __launch_bounds__(128) k_prep2(const __grid_constant__ DevRig rig, int Nmax, const double* __restrict__ und2,
                                               const int32_t* __restrict__ n2, double2* __restrict__ rec2)
{
    const int p = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;  // 2 * rod + end
    const int n = n2[p];
    if (n <= 0 || n > Nmax || e >= 2 * n) return;
    const double2 u = ldg2(und2 + ((size_t)p * Nmax * 2 + e) * 2);
    Canon2Rec c;
    canon_prep2(rig.CM1, rig.fx2, rig.P2o, u.x, u.y, c);
    double2* __restrict__ r = rec2 + (size_t)p * rec_problem_stride(Nmax) + e;
#pragma unroll
    for (int k = 0; k < kCanon2RecPlanes; k++) r[(size_t)k * 2 * Nmax] = make_double2(c.f[2 * k], c.f[2 * k + 1]);
}

// q / n for 0 <= q < 65536, 1 <= n <= 256 without an integer division: (q + 0.5) / n is at least
// 0.5 / n away from every integer, and its float evaluation with the hardware reciprocal (MUFU.RCP,
// relative error 2^-22) is off by less than 65536.5 / n * 2^-20 < 0.07 / n.
__device__ __forceinline__ int pair_row(int q, int n)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((float)n));
    return __float2int_rz(((float)q + 0.5f) * r);
}

// K1 for general rigs.  6 CTAs x 4 warps per SM (80 registers) measured fastest: 2.67 / 2.57 / 2.53 / 2.55 /
// 2.54 ms at 4 / 5 / 6 / 7 / 8 CTAs (Jacobi era, tools/ab_bench.py, profiles/README.md).
#ifndef PTK_KCOST_MINBLOCKS
#define PTK_KCOST_MINBLOCKS 6
#endif
template <bool SIMPLE, int MINB = PTK_KCOST_MINBLOCKS>
__global__ void __launch_bounds__(128, MINB) k_cost(const __grid_constant__ DevRig rig, const __grid_constant__ CostArgs a)
{
    const int p = blockIdx.y;
    const int tile = blockIdx.x;
    const int N1 = a.n1[p], N2 = a.n2[p];
    const int npairs = N1 * N2;
    if (N1 <= 0 || N2 <= 0 || N1 > a.Nmax || N2 > a.Nmax || tile * 32 >= npairs) return;
    int q = tile * 32 + (threadIdx.x >> 2);
    const bool active = q < npairs;
    if (!active) q = npairs - 1;  // keep the quad shuffles convergent; results are discarded
    const int c = threadIdx.x & 3;
    const int e1 = c >> 1, e2 = c & 1;
    const int i = pair_row(q, N2), j = q - i * N2;
    if (!PTK_OK_IDX(i, N1, 10) || !PTK_OK_IDX(j, N2, 11)) return;  // (debug build) the reciprocal-based row index
    // 32-bit element offsets (p < 65536 per launch, Nmax <= 256: below 2^26 rods): one IMAD.WIDE per address
    const unsigned base = (unsigned)p * (unsigned)a.Nmax;
    const unsigned off1 = (base + (unsigned)i) * 4u + 2u * e1, off2 = (base + (unsigned)j) * 4u + 2u * e2;
    const double2 u1 = ldg2(a.und1 + off1);
    const double2 o1 = centre_detection(rig.CM1, ldg2(a.cam1 + off1));
    const double2 o2 = centre_detection(rig.CM2, ldg2(a.cam2 + off2));
    Canon2Rec rec;  // unused by general rigs
    double d1, d2, h[4];
    eval_combo<false, SIMPLE>(rig, u1.x, u1.y, a.und2 + off2, rec, o1.x, o1.y, o2.x, o2.y, d1, d2, h);
    // match2D.py:910 np.mean over the two cameras; matchND.py:525 np.sum
    const double e = (d1 + d2) * rig.agg_scale;
    // e0+e3 (straight) lives in lanes 0,3 of the quad, e1+e2 (crossed) in lanes 1,2
    const double s_mine = e + __shfl_xor_sync(kFull, e, 3);
    const double s_other = __shfl_xor_sync(kFull, s_mine, 1);
    if (active) {
        // (p, i, j) of one launch stays below 65528 * 256 * 256 < 2^32
        const unsigned o = (base + (unsigned)i) * (unsigned)a.Nmax + (unsigned)j;
        if (c == 0) {
            a.cost[o] = np_min(s_mine, s_other);          // match2D.py:923-932
            if (a.choice) a.choice[o] = (s_mine <= s_other);  // match2D.py:935-938
        }
        if (a.p_world) {
            double X, Y, Z, wx, wy, wz;
            dehomogenise(h, X, Y, Z);
            to_world(rig, X, Y, Z, wx, wy, wz);
            double* w = a.p_world + ((size_t)o * 4 + c) * 3;
            w[0] = wx; w[1] = wy; w[2] = wz;
        }
        if (a.rep_errs) {
            double* r = a.rep_errs + ((size_t)o * 4 + c) * 2;
            r[0] = d1; r[1] = d2;
        }
    }
}

// K1 for canonical rigs (round 2): a CTA owns a block of the cost matrix -- `rows_per` camera-1 rods x one
// chunk of WC camera-2 rods of one problem -- instead of one 32-pair tile.  Its prologue stages the chunk's
// camera-2 records (k_prep2's planes), the chunk's original camera-2 points and the rows' camera-1 points in
// shared memory; the loop then walks the block's pairs 32 at a time (flat index, so ragged rod counts leave no
// idle lanes) with every per-pair input a shared-memory read off ONE base register.  The kernel is bound by
// instruction dispatch (FP64 and integer-ALU instructions hold a scheduler's port for two cycles, ncu:
// profiles/README.md), so what the loop does besides its ~250 FP64 instructions is counted in single
// instructions: the loop state is four registers (pair index, this lane's shared-memory base, output base,
// pair count), the shared-memory reads are inline PTX with immediate offsets (the compiler otherwise rebuilds
// every array's window address per iteration), the distorted points are read after the triangulation.
constexpr int kRowsMax = 16;  // camera-1 rods per CTA at most
struct RowsLoopArgs {
    int rows_per;  // camera-1 rods per CTA (<= kRowsMax); gridDim.x = ceil(Nmax / rows_per)
};
template <int OFF>
__device__ __forceinline__ double2 lds2(unsigned addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(addr), "n"(OFF));
    return v;
}
template <int WC, int K>
__device__ __forceinline__ void rec_load_shared(unsigned addr, Canon2Rec& c)
{
    if constexpr (K < kCanon2RecPlanes) {
        const double2 t = lds2<K * 2 * WC * 16>(addr);
        c.f[2 * K] = t.x;
        c.f[2 * K + 1] = t.y;
        rec_load_shared<WC, K + 1>(addr, c);
    }
}

// EXTRAS: the optional outputs (choice, p_world, rep_errs) exist; the reconstruction path wants the cost only.
template <bool SIMPLE, int WC, int MINB, bool EXTRAS>
__global__ void __launch_bounds__(128, MINB) k_cost_rows(const __grid_constant__ DevRig rig, const __grid_constant__ CostArgs a,
                                                         const __grid_constant__ RowsLoopArgs la)
{
    // shared memory, in double2: records [planes][2 WC] | o2 [2 WC] | u1 [2 kRowsMax] | o1 [2 kRowsMax]
    // (o1, o2: the distorted detections minus their camera's principal point, see centre_detection)
    constexpr int kRec = kCanon2RecPlanes * 2 * WC, kO2 = kRec, kU1 = kO2 + 2 * WC, kO1 = kU1 + 2 * kRowsMax;
    __shared__ __align__(16) double2 sm[kO1 + 2 * kRowsMax];
    const int p = blockIdx.y;
    const int N1 = a.n1[p], N2 = a.n2[p];
    if (N1 <= 0 || N2 <= 0 || N1 > a.Nmax || N2 > a.Nmax) return;
    const int j0 = blockIdx.z * WC, i0 = blockIdx.x * la.rows_per;
    if (j0 >= N2 || i0 >= N1) return;
    const int W = min(WC, N2 - j0), R = min(la.rows_per, N1 - i0);
    const unsigned tid = threadIdx.x;
    const unsigned base = (unsigned)p * (unsigned)a.Nmax;
    {
        const unsigned nE = 2u * W;  // camera-2 end points of the chunk
        const double2* __restrict__ src = a.rec2 + (size_t)p * rec_problem_stride(a.Nmax) + 2u * j0;
        for (unsigned e = tid; e < nE; e += 128) {
#pragma unroll
            for (int k = 0; k < kCanon2RecPlanes; k++) sm[k * 2 * WC + e] = __ldg(src + (size_t)k * 2 * a.Nmax + e);
            sm[kO2 + e] = centre_detection(rig.CM2, ldg2(a.cam2 + ((size_t)(base + j0) * 2 + e) * 2));
        }
        if (tid < 2u * R) {
            sm[kU1 + tid] = ldg2(a.und1 + ((size_t)(base + i0) * 2 + tid) * 2);
            sm[kO1 + tid] = centre_detection(rig.CM1, ldg2(a.cam1 + ((size_t)(base + i0) * 2 + tid) * 2));
        }
    }
    __syncthreads();
    const int npairs = R * W;
    float rW;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rW) : "f"((float)W));
    const unsigned c = tid & 3u;
    // this lane's shared-memory base: byte address of sm[e2] (camera-2 end) -- rod jl is 32 bytes further per rod;
    // the camera-1 end e1 is folded into the row offset below
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm) + (c & 1u) * 16u;
    const unsigned s1base = (unsigned)__cvta_generic_to_shared(sm) + (c >> 1) * 16u;
    const unsigned obase = (base + (unsigned)i0) * (unsigned)a.Nmax + (unsigned)j0;  // < 2^32 per launch
    const unsigned und2_base = (base + (unsigned)j0) * 4u + 2u * (c & 1u);
#pragma unroll 1
    for (int q = (int)(tid >> 2); q - (int)(tid >> 2) < npairs; q += 32) {
        const bool active = q < npairs;
        const int qq = active ? q : npairs - 1;  // keep the quad shuffles convergent; results are discarded
        const int r = __float2int_rz(((float)qq + 0.5f) * rW);  // qq / W, see pair_row
        const int jl = qq - r * W;
        if (!PTK_OK_IDX(r, R, 12) || !PTK_OK_IDX(jl, W, 13)) continue;  // (debug build)
        const unsigned s2 = sbase + 32u * (unsigned)jl;                  // record / o2 of (jl, e2)
        const unsigned s1 = s1base + 32u * (unsigned)r;                  // u1 / o1 of (r, e1)
        Canon2Rec rec;
        rec_load_shared<WC, 0>(s2, rec);
        const double2 u1 = lds2<kU1 * 16>(s1);
        double h[4], d1, d2;
        bool good = eval_combo_dlt<true>(rig, u1.x, u1.y, nullptr, rec, h);
        const double2 o1 = lds2<kO1 * 16>(s1), o2 = lds2<kO2 * 16>(s2);
        eval_combo_dist<true, SIMPLE>(rig, h, o1.x, o1.y, o2.x, o2.y, good, d1, d2);
        if (__builtin_expect(!good, 0)) {
            const double2 v1 = lds2<kU1 * 16>(s1);
            eval_combo_redo<true, SIMPLE>(rig, v1.x, v1.y, a.und2 + (und2_base + 4u * (unsigned)jl), o1.x, o1.y, o2.x, o2.y,
                                          d1, d2, h);
        }
        const double e = (d1 + d2) * rig.agg_scale;  // match2D.py:910 np.mean / matchND.py:525 np.sum
        const double s_mine = e + __shfl_xor_sync(kFull, e, 3);
        const double s_other = __shfl_xor_sync(kFull, s_mine, 1);
        if (active) {
            const unsigned o = obase + (unsigned)r * (unsigned)a.Nmax + (unsigned)jl;
            if (c == 0) {
                a.cost[o] = np_min(s_mine, s_other);                         // match2D.py:923-932
                if (EXTRAS && a.choice) a.choice[o] = (s_mine <= s_other);  // match2D.py:935-938
            }
            if (EXTRAS && a.p_world) {
                double X, Y, Z, wx, wy, wz;
                dehomogenise(h, X, Y, Z);
                to_world(rig, X, Y, Z, wx, wy, wz);
                double* w = a.p_world + ((size_t)o * 4 + c) * 3;
                w[0] = wx; w[1] = wy; w[2] = wz;
            }
            if (EXTRAS && a.rep_errs) {
                double* re = a.rep_errs + ((size_t)o * 4 + c) * 2;
                re[0] = d1; re[1] = d2;
            }
        }
    }
}

// ------------------------------------------------------------------------------------- K2
// scipy.optimize.linear_sum_assignment, one warp per problem (block = 1 warp, so "one CTA per
// (frame, colour) problem").  Crouse's shortest augmenting path with SciPy's exact scan
// order and tie rule: the `remaining` list starts reversed, a visited
// column is removed by swapping in the last one, and among columns of minimal reduced cost
// the LAST unassigned one in list order wins, else the FIRST one.  Lanes own columns
// j = lane, lane+32, ...; all solver state is register resident; the cost matrix is staged in
// shared memory when small (else read from L2); warp-wide min / argmin are REDUX ops.  The
// reduced cost is evaluated unfused, left to right, as SciPy's C++ does, so exact ties tie here too.
struct LsapArgs {
    int ld_rows, ld_cols;    // padded problem shape; cost row stride = ld_cols
    size_t cost_stride;      // doubles between problems
    const double* cost;
    const int32_t* nr;       // optional
    const int32_t* nc;       // optional
    int out_ld;              // stride of the per-problem outputs
    int32_t* row_ind;        // optional
    int32_t* col_ind;
    int32_t* n_out;          // optional
    int32_t* status;         // optional
    double* assigned;        // optional [P,out_ld]: cost[row_ind[m], col_ind[m]]
    double* total;           // optional [P]: sum of assigned
    uint8_t* keep;           // optional [P,out_ld]: assigned <= max_err
    double max_err;
    int stage;               // set by the launcher: cost matrix fits the shared-memory staging budget
};

__device__ __forceinline__ double warp_min(double x)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmin(x, __shfl_xor_sync(kFull, x, o));
    return x;
}

__device__ __forceinline__ double warp_sum(double x)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
    return x;
}

// order-preserving map double -> unsigned 64 (for REDUX-based min / equality)
__device__ __forceinline__ unsigned long long ordered_key(double x)
{
    const unsigned long long k = (unsigned long long)__double_as_longlong(x);
    return (k >> 63) ? ~k : (k | 0x8000000000000000ull);
}
__device__ __forceinline__ double ordered_value(unsigned long long k)
{
    k = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)k);
}

// value of arr[idx >> 5] held by lane (idx & 31), broadcast to the warp (idx is warp-uniform): every lane first
// picks its own arr[idx >> 5] (a chain of selects on a warp-uniform predicate), then ONE shuffle fetches the
// owner's -- instead of one shuffle per slot
template <int CPL, typename T>
__device__ __forceinline__ T lane_array_get(const T (&arr)[CPL], int idx)
{
    T mine = arr[0];
#pragma unroll
    for (int t = 1; t < CPL; t++)
        if (t == (idx >> 5)) mine = arr[t];
    return __shfl_sync(kFull, mine, idx & 31);
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

constexpr int kLsapStageBytes = 16 * 1024;  // stage the cost matrix in shared memory up to this size

__host__ __device__ inline size_t lsap_row_state_bytes(int M) { return (size_t)M * (sizeof(double) + sizeof(int32_t)) + 16; }

// CPL = columns per lane: the problem must satisfy max(nr, nc) <= 32*CPL.
// Lane l owns columns l, l+32, ...: v, shortest-path cost, path, row4col and the column's position
// in SciPy's `remaining` list live in registers (the list itself is never stored: only positions,
// which is all the tie rule needs; removal = "the column at the last position takes the freed
// one").  Row duals u and col4row sit in shared memory (one read per Dijkstra step).  A step is
// ~60 warp instructions + ~25 per owned column; the kernel is issue bound, not latency bound.
// (The tracking assignment has its own kernel without a cost matrix: ptk_lsap_trk.cuh.)
template <int CPL>
__global__ void __launch_bounds__(32) k_lsap(const __grid_constant__ LsapArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int M = a.ld_rows > a.ld_cols ? a.ld_rows : a.ld_cols;
    double* u = reinterpret_cast<double*>(smem_raw);
    int32_t* c4r = reinterpret_cast<int32_t*>(u + M);
    double* sm = reinterpret_cast<double*>(smem_raw + ((lsap_row_state_bytes(M) + 15) / 16) * 16);
    const int lane = threadIdx.x;
    const long long p = blockIdx.x;
    const int nr0 = a.nr ? a.nr[p] : a.ld_rows;
    const int nc0 = a.nc ? a.nc[p] : a.ld_cols;
    const double* C = a.cost + (size_t)p * a.cost_stride;
    const int ldc = a.ld_cols;
    int32_t* orow = a.row_ind ? a.row_ind + (size_t)p * a.out_ld : nullptr;
    int32_t* ocol = a.col_ind + (size_t)p * a.out_ld;
    double* oasg = a.assigned ? a.assigned + (size_t)p * a.out_ld : nullptr;
    uint8_t* okeep = a.keep ? a.keep + (size_t)p * a.out_ld : nullptr;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);

    int st = PTK_PROBLEM_OK;
    int n = 0;
    if (nr0 <= 0 || nc0 <= 0) st = PTK_PROBLEM_EMPTY;
    else if (nr0 > a.ld_rows || nc0 > a.ld_cols) st = PTK_PROBLEM_INVALID;  // counts beyond the padded shape
    const bool tr = nc0 < nr0;  // tall matrices are solved transposed, as SciPy does
    const int nr = tr ? nc0 : nr0, nc = tr ? nr0 : nc0;
    const bool staged = a.stage != 0;

    if (st == PTK_PROBLEM_OK) {
        // validate (SciPy: "matrix contains invalid numeric entries") and stage in solver orientation
        bool bad = false;
        {
            // one flat sweep (independent loads: the row-by-row form measured 50 % slower); the row index comes
            // from a float reciprocal instead of an integer division per element (exact here, see pair_row)
            const int tot = nr0 * nc0;
            float rnc;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rnc) : "f"((float)nc0));
            for (int idx = lane; idx < tot; idx += 32) {
                const int i = __float2int_rz(((float)idx + 0.5f) * rnc), j = idx - i * nc0;
                const double c = C[(size_t)i * ldc + j];
                bad |= (c != c) || (c == -INFINITY);
                if (staged) PTK_AT(sm, tr ? (j * nc + i) : (i * nc + j), nr * nc, 20) = c;
            }
        }
        if (__any_sync(kFull, bad)) st = PTK_PROBLEM_INVALID;
        for (int r = lane; r < nr; r += 32) { u[r] = 0.0; c4r[r] = -1; }
        __syncwarp();
    }

    // The Dijkstra step is bound by instruction dispatch (integer / select / compare instructions hold the
    // scheduler's port for two cycles), so it is written like the tracking solver's (ptk_lsap_trk.cuh):
    //  * keys are the RAW bits of the shortest-path cost -- for non-negative doubles the signed integer order
    //    of the bits is the value order; one unsigned compare on the first REDUX result catches "every
    //    remaining column at +inf" and "a negative candidate", and only the latter is redone with the
    //    order-preserving transform (skey);
    //  * a visited (or absent) column's high word becomes +NaN's (its real one moves to seen_hi): `r < NaN`,
    //    `NaN == min` are false and NaN's bits are above +inf's -- no "live" predicate in the loop;
    //  * the tie-rule candidate is ONE multiply-add from two per-column constants that only change on
    //    augmentation; it carries row4col + 1 and the column, so the winner needs no further broadcast.
    const int kDeadHi = 0x7ff80000;
    double v[CPL];
    int spc_hi[CPL], spc_lo[CPL], seen_hi[CPL];
    int path[CPL], r4c[CPL], pos[CPL], jc[CPL];
    unsigned ka[CPL], kb[CPL];
#pragma unroll
    for (int s = 0; s < CPL; s++) {
        const int j = lane + 32 * s;
        v[s] = 0.0; spc_hi[s] = kDeadHi; spc_lo[s] = 0; seen_hi[s] = kDeadHi; path[s] = -1; r4c[s] = -1; pos[s] = -1;
        jc[s] = j < nc ? j : (nc > 0 ? nc - 1 : 0);  // a column index that may be read (absent columns are never used)
        ka[s] = 0u - (1u << 18);                      // unassigned column: rank = 511 - pos (largest position wins)
        kb[s] = (511u << 18) | (unsigned)j;
    }

    if (st == PTK_PROBLEM_OK) {
        for (int cur = 0; cur < nr; cur++) {
#pragma unroll
            for (int s = 0; s < CPL; s++) {
                const int j = lane + 32 * s;
                spc_hi[s] = (j < nc) ? 0x7ff00000 : kDeadHi;  // +inf
                spc_lo[s] = 0;
                seen_hi[s] = kDeadHi;                 // NaN = not visited in this row
                pos[s] = (j < nc) ? (nc - 1 - j) : -1;  // remaining[it] = nc-1-it  <=>  pos[j] = nc-1-j
            }
            double minVal = 0.0;
            int i = cur, num_rem = nc, sink = -1;
            while (sink == -1) {
                const double ui = PTK_AT(u, i, nr, 21);
                int mhi;
                unsigned mlo;
#pragma unroll
                for (int s = 0; s < CPL; s++) {
                    const int j = jc[s];
                    const double c = staged ? PTK_AT(sm, i * nc + j, nr * nc, 22)
                                            : (tr ? C[(size_t)j * ldc + i] : C[(size_t)i * ldc + j]);
                    const double r = ((minVal + c) - ui) - v[s];
                    if (r < __hiloint2double(spc_hi[s], spc_lo[s])) {  // false for a visited / absent column (NaN)
                        spc_hi[s] = __double2hiint(r);
                        spc_lo[s] = __double2loint(r);
                        path[s] = i;
                    }
                    const int h = spc_hi[s];
                    const unsigned l = (unsigned)spc_lo[s];
                    if (s == 0 || h < mhi || (h == mhi && l < mlo)) { mhi = h; mlo = l; }
                }
                // warp argmin with SciPy's tie rule in three REDUX ops: (1) high and (2) low word of the cost's
                // bits, then (3) one packed integer whose order is the tie rule -- unassigned columns first
                // (largest list position wins), then assigned ones (smallest position wins)
                const int mh = __reduce_min_sync(kFull, mhi);
                if ((unsigned)mh >= 0x7ff00000u) {  // rare: +inf everywhere, or a negative candidate
                    if (mh >= 0) { st = PTK_PROBLEM_INFEASIBLE; break; }
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
                const int r4 = (int)((key3 >> 9) & 511u) - 1;
                const unsigned rank = key3 >> 18;
                const int index = rank < 512u ? (int)(511u - rank) : (int)(rank - 512u);
                if (r4 == -1) sink = jsel; else i = r4;
                // remaining[index] = remaining[--num_remaining]: the column at the last position takes the freed
                // one; the selected column becomes visited (SC): its high word moves to seen_hi
                num_rem--;
#pragma unroll
                for (int s = 0; s < CPL; s++) {
                    if (lane + 32 * s == jsel) { seen_hi[s] = spc_hi[s]; spc_hi[s] = kDeadHi; }
                    else if (pos[s] == num_rem) pos[s] = index;
                }
            }
            if (st != PTK_PROBLEM_OK) break;
            // dual update.  SciPy: u[cur] += minVal; u[i] += minVal - spc[col4row[i]] for the other
            // visited rows; v[j] -= minVal - spc[j] for visited columns.  Every visited row other
            // than cur was entered through its assigned (visited) column, so the row update is
            // done from the column side with the same operands.
            if (lane == 0) u[cur] += minVal;
#pragma unroll
            for (int s = 0; s < CPL; s++) {
                if (seen_hi[s] != kDeadHi) {  // visited in this row
                    const double d = minVal - __hiloint2double(seen_hi[s], spc_lo[s]);
                    if (r4c[s] >= 0) u[r4c[s]] += d;
                    v[s] -= d;
                }
            }
            __syncwarp();
            // augment along the path
            int j = sink;
            while (true) {
                const int r = lane_array_get<CPL>(path, j);
                if (!PTK_OK_IDX(r, nr, 23) || !PTK_OK_IDX(j, nc, 24)) break;  // (debug build) a corrupt augmenting path
#pragma unroll
                for (int s = 0; s < CPL; s++)
                    if (lane + 32 * s == j) {
                        r4c[s] = r;
                        ka[s] = 1u << 18;  // assigned column: rank = 512 + pos (smallest position wins)
                        kb[s] = (512u << 18) | ((unsigned)(r + 1) << 9) | (unsigned)j;
                    }
                const int t = c4r[r];
                __syncwarp();
                if (lane == 0) c4r[r] = j;
                j = t;
                if (r == cur) break;
            }
            __syncwarp();
        }
    }

    // ---- outputs
    if (st == PTK_PROBLEM_OK) {
        n = nr;
        double tsum = 0.0;
        if (!tr) {
            for (int r = lane; r < nr; r += 32) {
                const int cj = c4r[r];
                const double cst = PTK_OK_IDX(cj, nc, 25) ? C[(size_t)r * ldc + cj] : qnan;
                if (orow) orow[r] = r;
                ocol[r] = cj;
                if (oasg) oasg[r] = cst;
                if (okeep) okeep[r] = (cst <= a.max_err) ? 1 : 0;
                tsum += cst;
            }
        } else {
            // rows of the transposed problem are original columns; emit sorted by original row
            int off = 0;
#pragma unroll
            for (int s = 0; s < CPL; s++) {
                const int r0 = lane + 32 * s;
                const bool has = r0 < nc && r4c[s] != -1;
                const unsigned m = __ballot_sync(kFull, has);
                if (has) {
                    const int k = off + __popc(m & ((1u << lane) - 1u));
                    const int cj = r4c[s];
                    const double cst = C[(size_t)r0 * ldc + cj];
                    if (orow) orow[k] = r0;
                    ocol[k] = cj;
                    if (oasg) oasg[k] = cst;
                    if (okeep) okeep[k] = (cst <= a.max_err) ? 1 : 0;
                    tsum += cst;
                }
                off += __popc(m);
            }
        }
        if (a.total) {
            tsum = warp_sum(tsum);
            if (lane == 0) a.total[p] = tsum;
        }
    } else if (a.total && lane == 0) {
        a.total[p] = qnan;
    }
    for (int m = n + lane; m < a.out_ld; m += 32) {
        if (orow) orow[m] = -1;
        ocol[m] = -1;
        if (oasg) oasg[m] = qnan;
        if (okeep) okeep[m] = 0;
    }
    if (lane == 0) {
        if (a.n_out) a.n_out[p] = n;
        if (a.status) a.status[p] = st;
    }
}

// ------------------------------------------------------------------------------------- K3
// Row gather (match2D.py:963-983).  4 lanes per output row m re-evaluate the four combos of
// the rod pair (m, col_ind[m]) -- the loop index m, not row_ind[m], selects the cam1 rod,
// which reproduces the reference when N1 > N2 -- with the very same
// arithmetic K1 used, pick straight/crossed, and write the 18 columns.  Rows m >= n_out are
// filled with NaN so the padded output is fully defined.
struct RowsArgs {
    int Nmax;
    const double* cam1;
    const double* cam2;
    const double* und1;
    const double* und2;
    const int32_t* col_ind;  // [P,Nmax]
    const int32_t* n_out;    // [P]
    double* rows;            // optional [P,Nmax,18]
    double* pair_cost;       // optional [P,Nmax]: min(e0+e3, e1+e2) of the evaluated pair (m, col_ind[m])
    // compact output (ptk_reconstruct_host_compact): the six computed coordinates and the straight /
    // crossed verdict; the other 12 columns are functions of these, of col_ind and of the caller's own
    // input arrays (ptk_expand_rows_host)
    double* ends6;           // optional [P,Nmax,6]
    uint8_t* straight;       // optional [P,Nmax]
    // tracking layout (tracking.py:97-98), written here instead of by a separate gather: the six computed
    // coordinates of row m < trk_N of problem g = trk_p0 + p = frame * C + colour go to ends_trk[colour, frame, m, 2, 3]
    double* ends_trk;        // optional [C,F,trk_N,6]
    int trk_F, trk_C, trk_N;
    long long trk_p0;
    double trk_rC;           // 1 / trk_C: the frame of a problem from a multiply instead of a 64-bit division
};
// frame = g / C for 0 <= g < 2^40: (g + 0.5) / C is at least 0.5 / C away from every integer, far more than the
// rounding of the product
__device__ __forceinline__ long long trk_frame(long long g, double rC) { return (long long)(((double)g + 0.5) * rC); }

template <bool CANON, bool SIMPLE>
__global__ void __launch_bounds__(128) k_rows(const __grid_constant__ DevRig rig, const __grid_constant__ RowsArgs a)
{
    const int p = blockIdx.y;
    const int rpb = blockDim.x >> 2;  // rows per block: 32, fewer for small problems (no idle quads)
    const int m = blockIdx.x * rpb + (threadIdx.x >> 2);
    const int c = threadIdx.x & 3;
    const int n = a.n_out[p];
    const size_t base = (size_t)p * a.Nmax;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    if (blockIdx.x * rpb >= n) {  // whole block past the matches: NaN fill
        if (m < a.Nmax) {
            if (a.rows) {
                double* o = a.rows + (base + m) * PTK_ROW_COLS;
                for (int k = c; k < PTK_ROW_COLS; k += 4) o[k] = qnan;
            }
            if (a.ends6) {
                double* o = a.ends6 + (base + m) * 6;
                for (int k = c; k < 6; k += 4) o[k] = qnan;
            }
            if (a.ends_trk && m < a.trk_N) {
                const long long g = a.trk_p0 + p, f = trk_frame(g, a.trk_rC), col = g - f * a.trk_C;
                double* o = a.ends_trk + (((size_t)col * a.trk_F + f) * a.trk_N + m) * 6;
                for (int k = c; k < 6; k += 4) o[k] = qnan;
            }
            if (c == 0) {
                if (a.pair_cost) a.pair_cost[base + m] = qnan;
                if (a.straight) a.straight[base + m] = 0;
            }
        }
        return;
    }
    const bool active = m < n;
    const int mm = active ? m : n - 1;
    int i2 = a.col_ind[base + mm];
    if (!PTK_OK_IDX(i2, a.Nmax, 30)) i2 = 0;  // (debug build) an assignment outside the padded problem
    const int e1 = c >> 1, e2 = c & 1;
    const double2 u1 = ldg2(a.und1 + (base + mm) * 4 + 2 * e1);
    const double* u2p = a.und2 + (base + i2) * 4 + 2 * e2;
    const double2 o1 = centre_detection(rig.CM1, ldg2(a.cam1 + (base + mm) * 4 + 2 * e1));
    const double2 o2 = centre_detection(rig.CM2, ldg2(a.cam2 + (base + i2) * 4 + 2 * e2));
    Canon2Rec rec;
    if (CANON) {  // the record K1 read from k_prep2's planes, recomputed by the same function: the same bits
        const double2 u2 = ldg2(u2p);
        canon_prep2(rig.CM1, rig.fx2, rig.P2o, u2.x, u2.y, rec);
    }
    double d1, d2, h[4], X, Y, Z;
    eval_combo<CANON, SIMPLE>(rig, u1.x, u1.y, u2p, rec, o1.x, o1.y, o2.x, o2.y, d1, d2, h);
    dehomogenise(h, X, Y, Z);
    const double e = (d1 + d2) * rig.agg_scale;
    const double s_mine = e + __shfl_xor_sync(kFull, e, 3);
    const double s_other = __shfl_xor_sync(kFull, s_mine, 1);
    // lane 0 of the quad holds (s03, s12); broadcast its verdict
    const int straight0 = (s_mine <= s_other);
    const int straight = __shfl_sync(kFull, straight0, threadIdx.x & 28, 32);
    double wx, wy, wz;
    to_world(rig, X, Y, Z, wx, wy, wz);
    // end 1 = combo 0 (straight) or 1 (crossed); end 2 = combo 3 or 2
    const int q0 = threadIdx.x & 28;
    const int s1 = q0 + (straight ? 0 : 1), s2 = q0 + (straight ? 3 : 2);
    const double ax = __shfl_sync(kFull, wx, s1), ay = __shfl_sync(kFull, wy, s1), az = __shfl_sync(kFull, wz, s1);
    const double bx = __shfl_sync(kFull, wx, s2), by = __shfl_sync(kFull, wy, s2), bz = __shfl_sync(kFull, wz, s2);
    if (c == 0) {
        if (a.pair_cost && m < a.Nmax) a.pair_cost[base + m] = active ? np_min(s_mine, s_other) : qnan;
        if (a.straight && m < a.Nmax) a.straight[base + m] = (active && straight) ? 1 : 0;
        if (a.ends_trk && m < a.trk_N) {  // rows >= n_out are NaN in every output form
            const long long g = a.trk_p0 + p, f = trk_frame(g, a.trk_rC), col = g - f * a.trk_C;
            double* o = a.ends_trk + (((size_t)col * a.trk_F + f) * a.trk_N + m) * 6;
            o[0] = active ? ax : qnan; o[1] = active ? ay : qnan; o[2] = active ? az : qnan;
            o[3] = active ? bx : qnan; o[4] = active ? by : qnan; o[5] = active ? bz : qnan;
        }
        if (active) {
            if (a.ends6) {
                double* o = a.ends6 + (base + m) * 6;
                o[0] = ax; o[1] = ay; o[2] = az;
                o[3] = bx; o[4] = by; o[5] = bz;
            }
            if (a.rows) {
                double* o = a.rows + (base + m) * PTK_ROW_COLS;
                o[0] = ax; o[1] = ay; o[2] = az;
                o[3] = bx; o[4] = by; o[5] = bz;
                o[6] = (ax + bx) / 2; o[7] = (ay + by) / 2; o[8] = (az + bz) / 2;  // match2D.py:969
                const double dx = bx - ax, dy = by - ay, dz = bz - az;
                o[9] = sqrt(dx * dx + dy * dy + dz * dz);                           // match2D.py:970-972
                const double* r1 = a.cam1 + (base + m) * 4;
                const double* r2 = a.cam2 + (base + i2) * 4;
                if (straight) {  // match2D.py:973-974
                    o[10] = r1[0]; o[11] = r1[1]; o[12] = r1[2]; o[13] = r1[3];
                    o[14] = r2[0]; o[15] = r2[1]; o[16] = r2[2]; o[17] = r2[3];
                } else {         // match2D.py:982-983: both cameras' end points reversed
                    o[10] = r1[2]; o[11] = r1[3]; o[12] = r1[0]; o[13] = r1[1];
                    o[14] = r2[2]; o[15] = r2[3]; o[16] = r2[0]; o[17] = r2[1];
                }
            }
        } else if (m < a.Nmax) {
            if (a.rows) {
                double* o = a.rows + (base + m) * PTK_ROW_COLS;
                for (int k = 0; k < PTK_ROW_COLS; k++) o[k] = qnan;
            }
            if (a.ends6) {
                double* o = a.ends6 + (base + m) * 6;
                for (int k = 0; k < 6; k++) o[k] = qnan;
            }
        }
    }
}

// gather (x1..z2) of every output row into the tracking layout ends[C,F,N,2,3]
// (what tracking.py:97-98 extracts from the DataFrame), problem index = f*C + c.
__global__ void k_rows_to_ends(int F0, int Fn, int F, int C, int N, int Nmax, int row_cols, const double* __restrict__ rows,
                               double* __restrict__ ends)
{
    // rows: [Fn*C, Nmax, row_cols] (18-column rows, or the compact 6-column ends) for frames F0..F0+Fn-1; ends: [C, F, N, 6]
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long tot = (long long)Fn * C * N * 6;
    if (t >= tot) return;
    const int k = (int)(t % 6);
    long long r = t / 6;
    const int m = (int)(r % N); r /= N;
    const int c = (int)(r % C);
    const int f = (int)(r / C);
    ends[(((size_t)c * F + (F0 + f)) * N + m) * 6 + k] = rows[(((size_t)f * C + c) * Nmax + m) * row_cols + k];
}

// ------------------------------------------------------------------------------------- K4
// tracking.py:100-121.  cost[q,a,b] for q = (colour, frame pair); unfused arithmetic in
// numpy's order: diff, squares, ((x^2+y^2)+z^2), sqrt, sum of the two norms, np.min.
__device__ __forceinline__ double norm3(double ax, double ay, double az, double bx, double by, double bz)
{
    const double dx = ax - bx, dy = ay - by, dz = az - bz;
    return sqrt((dx * dx + dy * dy) + dz * dz);
}

// One CTA per frame pair.  The first candidate d0 = |q1[a]-p1[a]| + |q2[b]-p2[b]| is separable, so
// its 2N norms are computed once into shared memory (same arithmetic, same bits) and only the
// cross term costs two norms per (a,b); the 2 x N x 6 doubles of the two frames are staged in
// shared memory as well.
__global__ void __launch_bounds__(256) k_track_cost(int N, int F, long long q0, const double* __restrict__ ends,
                                                    double* __restrict__ cost)
{
    extern __shared__ double tsm[];  // cur[6N] | nxt[6N] | n1[N] | n2[N]
    double* cur = tsm;
    double* nxt = tsm + 6 * N;
    double* n1 = tsm + 12 * N;
    double* n2 = n1 + N;
    const long long ql = blockIdx.x;  // problem within this launch
    const long long q = q0 + ql;
    const int c = (int)(q / (F - 1)), f = (int)(q - (long long)c * (F - 1));
    const double* g = ends + ((size_t)c * F + f) * N * 6;
    for (int k = threadIdx.x; k < 12 * N; k += blockDim.x) tsm[k] = g[k];  // frames f and f+1 are contiguous
    __syncthreads();
    for (int k = threadIdx.x; k < 2 * N; k += blockDim.x) {
        const int r = k >> 1, e = (k & 1) * 3;  // rod r, end point 1 or 2
        const double v = norm3(nxt[6 * r + e], nxt[6 * r + e + 1], nxt[6 * r + e + 2],
                               cur[6 * r + e], cur[6 * r + e + 1], cur[6 * r + e + 2]);
        if (k & 1) n2[r] = v; else n1[r] = v;
    }
    __syncthreads();
    double* out = cost + (size_t)ql * N * N;
    for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
        const int aa = e / N, bb = e - aa * N;
        const double* p1a = cur + 6 * aa;
        const double* p2b = cur + 6 * bb + 3;
        const double* q1a = nxt + 6 * aa;
        const double* q2b = nxt + 6 * bb + 3;
        const double d0 = n1[aa] + n2[bb];
        const double d1 = norm3(p1a[0], p1a[1], p1a[2], q2b[0], q2b[1], q2b[2]) +
                          norm3(p2b[0], p2b[1], p2b[2], q1a[0], q1a[1], q1a[2]);
        out[e] = np_min(d0, d1);
    }
}

// ------------------------------------------------------------------------------------- K5
// Chain pass.  ids[f+1][col[f][a]] = ids[f][a] is a composition of permutations, which is
// associative, so frames are cut into segments of kChainSeg frames:
//  pass A (one block per (colour, segment)): local ids L[f][x] relative to the segment's
//         first frame (L[f0] = identity) are written to track_id; continuing the recursion
//         through the pair that leaves the segment gives the segment map (position in the
//         next segment's first frame -> local id), stored in seg_map.
//  pass B (one block per colour): start[0] = ids_in | identity, start[s+1][x] = start[s][map_s[x]].
//  pass C (elementwise): track_id[f][x] = start[f / kChainSeg][ L[f][x] ].
// The same three steps give the multi-GPU version: a rank is a super-segment and pass B runs
// over the gathered per-rank maps (particletracking_b200/distributed.py).
//
// Broken chains.  A frame pair whose LSAP failed (status INVALID / INFEASIBLE: NaN end points of a
// ragged frame) has col = -1 for every rod; a caller may also pass arbitrary garbage.  Any entry
// outside 0..N-1 is ignored, positions no valid entry maps to get id -1, and -1 propagates through
// every later composition (all gathers below are guarded) -- the same rule as the oracle's
// chain_tracks.  No index derived from col / ids is ever used unchecked.
constexpr int kChainSeg = 32;

__device__ __forceinline__ int32_t chain_pick(const int32_t* __restrict__ tab, int32_t idx, int N)
{
    return ((unsigned)idx < (unsigned)N) ? tab[idx] : -1;
}

__global__ void k_chain_local(int F, int N, int nseg, const int32_t* __restrict__ col_ind,
                              int32_t* __restrict__ track_id, int32_t* __restrict__ seg_map)
{
    extern __shared__ int32_t sh[];  // cur[N] | nxt[N] | the segment's assignments [<=kChainSeg][N]
    int32_t* cur = sh;
    int32_t* nxt = sh + N;
    int32_t* cols = sh + 2 * N;
    const int c = blockIdx.y, s = blockIdx.x;
    const int f0 = s * kChainSeg;
    const int fl = min(f0 + kChainSeg, F) - 1;  // last frame of the segment
    // pairs f0..fl-1 stay inside the segment; pair fl (if any) leaves it
    const int f_end = (fl < F - 1) ? fl + 1 : fl;
    int32_t* tid = track_id + (size_t)c * F * N;
    // all assignments of the segment in one coalesced sweep (independent loads), then the
    // dependent chain runs out of shared memory
    const int32_t* g = col_ind + ((size_t)c * (F - 1) + f0) * N;
    for (int k = threadIdx.x; k < (f_end - f0) * N; k += blockDim.x) cols[k] = g[k];
    for (int a = threadIdx.x; a < N; a += blockDim.x) {
        cur[a] = a;
        nxt[a] = -1;
        tid[(size_t)f0 * N + a] = a;
    }
    __syncthreads();
    for (int f = f0; f < f_end; f++) {
        const int32_t* col = cols + (size_t)(f - f0) * N;
        for (int a = threadIdx.x; a < N; a += blockDim.x) {
            const int32_t j = col[a];
            if ((unsigned)j < (unsigned)N) nxt[j] = cur[a];
        }
        __syncthreads();
        if (f < fl) {
            for (int a = threadIdx.x; a < N; a += blockDim.x) tid[(size_t)(f + 1) * N + a] = nxt[a];
        } else {
            for (int a = threadIdx.x; a < N; a += blockDim.x) seg_map[((size_t)c * nseg + s) * N + a] = nxt[a];
        }
        for (int a = threadIdx.x; a < N; a += blockDim.x) cur[a] = -1;  // becomes the next frame's target
        int32_t* t = cur; cur = nxt; nxt = t;
        __syncthreads();
    }
}

__global__ void k_chain_scan(int N, int nseg, const int32_t* __restrict__ ids_in, const int32_t* __restrict__ seg_map,
                             int32_t* __restrict__ seg_start)
{
    extern __shared__ int32_t sh[];  // cur[N] | nxt[N] | a block of kChainSeg segment maps
    int32_t* cur = sh;
    int32_t* nxt = sh + N;
    int32_t* maps = sh + 2 * N;
    const int c = blockIdx.x;
    for (int a = threadIdx.x; a < N; a += blockDim.x) cur[a] = ids_in ? ids_in[(size_t)c * N + a] : a;
    __syncthreads();
    for (int s0 = 0; s0 < nseg; s0 += kChainSeg) {
        const int ns = min(kChainSeg, nseg - s0);
        const int32_t* g = seg_map + ((size_t)c * nseg + s0) * N;
        // the last segment of all has no map (nothing follows it): do not read it
        const int nload = (s0 + ns == nseg) ? ns - 1 : ns;
        for (int k = threadIdx.x; k < nload * N; k += blockDim.x) maps[k] = g[k];
        __syncthreads();
        for (int s = 0; s < ns; s++) {
            for (int a = threadIdx.x; a < N; a += blockDim.x) seg_start[((size_t)c * nseg + s0 + s) * N + a] = cur[a];
            if (s0 + s + 1 == nseg) break;
            const int32_t* mp = maps + (size_t)s * N;
            for (int a = threadIdx.x; a < N; a += blockDim.x) nxt[a] = chain_pick(cur, mp[a], N);
            __syncthreads();
            int32_t* t = cur; cur = nxt; nxt = t;
        }
        __syncthreads();
    }
}

// loc[C,F,N]: local ids of pass A.  Frames f < F_out go to track_id[C,F_out,N]; frame F_out (the halo
// frame of a shard, if F_out < F) goes to halo_map[C,N].
__global__ void k_chain_apply(int F, int N, int nseg, const int32_t* __restrict__ seg_start, const int32_t* __restrict__ loc,
                              int F_out, int32_t* __restrict__ track_id, int32_t* __restrict__ halo_map, long long total)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int a = (int)(t % N);
    const long long r = t / N;
    const int f = (int)(r % F);
    const int c = (int)(r / F);
    const int s = f / kChainSeg;
    const int32_t id = chain_pick(seg_start + ((size_t)c * nseg + s) * N, loc[t], N);
    if (f < F_out) track_id[((size_t)c * F_out + f) * N + a] = id;
    else if (f == F_out && halo_map) halo_map[(size_t)c * N + a] = id;
}

// Multi-GPU chain pass: compose the gathered per-rank boundary maps up to `rank`
// (start[c][x] = global track id of local id x on this rank) and re-base the local ids.
// maps[W,C,N] (rank stride map_stride ints): position x in rank r+1's first frame -> local id on rank r.
__global__ void k_chain_compose(int W, int C, int N, int rank, const int32_t* __restrict__ maps, size_t map_stride,
                                int32_t* __restrict__ start)
{
    extern __shared__ int32_t sh[];
    int32_t* cur = sh;
    int32_t* nxt = sh + N;
    const int c = blockIdx.x;
    for (int a = threadIdx.x; a < N; a += blockDim.x) cur[a] = a;
    __syncthreads();
    for (int r = 0; r < rank && r < W; r++) {
        const int32_t* mp = maps + (size_t)r * map_stride + (size_t)c * N;
        for (int a = threadIdx.x; a < N; a += blockDim.x) nxt[a] = chain_pick(cur, mp[a], N);
        __syncthreads();
        int32_t* t = cur; cur = nxt; nxt = t;
        __syncthreads();
    }
    for (int a = threadIdx.x; a < N; a += blockDim.x) start[(size_t)c * N + a] = cur[a];
}

// The root side of the sharded chain pass in one launch: block (c, r - 1) composes the boundary maps of
// ranks 0..r-1 for colour c (start = global id of every local id of rank r) and re-bases rank r's ids in
// place.  blocks: the gathered result blocks, rank stride `stride` ints; ids at tid_off, maps at map_off (ints).
__global__ void k_chain_rebase_gathered(int C, int N, const int32_t* __restrict__ frames, int32_t* __restrict__ blocks,
                                        size_t stride, size_t tid_off, size_t map_off)
{
    extern __shared__ int32_t sh[];
    int32_t* cur = sh;
    int32_t* nxt = sh + N;
    const int c = blockIdx.x, r = blockIdx.y + 1;
    for (int a = threadIdx.x; a < N; a += blockDim.x) cur[a] = a;
    __syncthreads();
    for (int q = 0; q < r; q++) {
        const int32_t* mp = blocks + (size_t)q * stride + map_off + (size_t)c * N;
        for (int a = threadIdx.x; a < N; a += blockDim.x) nxt[a] = chain_pick(cur, mp[a], N);
        __syncthreads();
        int32_t* t = cur; cur = nxt; nxt = t;
        __syncthreads();
    }
    const int F = frames[r];
    int32_t* tid = blocks + (size_t)r * stride + tid_off + (size_t)c * F * N;
    for (int e = threadIdx.x; e < F * N; e += blockDim.x) tid[e] = chain_pick(cur, tid[e], N);
}

__global__ void k_chain_rebase(int F, int N, const int32_t* __restrict__ start, int32_t* __restrict__ track_id,
                               long long total)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int c = (int)(t / ((long long)F * N));
    track_id[t] = chain_pick(start + (size_t)c * N, track_id[t], N);
}

// ------------------------------------------------------------------------------------- K6
// matchND.create_weights (matchND.py:141-217), one thread per (prev rod i, cam1 rod j, cam2 rod k).
__device__ __forceinline__ double dist3(const double* a, const double* b)
{
    const double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    return sqrt((dx * dx + dy * dy) + dz * dz);
}

__global__ void __launch_bounds__(256) k_weights_raw(int Np, int N1, int N2, const double* __restrict__ p3,
                                                     const double* __restrict__ prev, const double* __restrict__ re,
                                                     double* __restrict__ w, double* __restrict__ choices,
                                                     double* __restrict__ block_max)
{
    __shared__ double red[8];
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long tot = (long long)Np * N1 * N2;
    double mine = -INFINITY;
    if (t < tot) {
        const int jk = (int)(t % ((long long)N1 * N2));
        const int i = (int)(t / ((long long)N1 * N2));
        const double* q1 = prev + 6 * (size_t)i;
        const double* q2 = q1 + 3;
        const double* p = p3 + (size_t)jk * 12;
        const double* r = re + (size_t)jk * 8;
        const double *c1p1 = p, *c2p1 = p + 3, *c2p2 = p + 6, *c1p2 = p + 9;  // matchND.py:184-187
        const double c1_re = ((r[0] + r[1]) + r[6]) + r[7];                   // combos 0 and 3
        const double c2_re = r[2] + r[3];                                     // combo 1 only: the 1:2 slice
        double s0 = (dist3(c1p1, q1) + dist3(c1p2, q2)) * c1_re;
        double s1 = (dist3(c2p1, q1) + dist3(c2p2, q2)) * c2_re;
        double s2 = (dist3(c2p1, q2) + dist3(c2p2, q1)) * c2_re;
        double s3 = (dist3(c1p1, q2) + dist3(c1p2, q1)) * c1_re;
        int am = 0; double best = s0;
        if (s1 < best) { best = s1; am = 1; }
        if (s2 < best) { best = s2; am = 2; }
        if (s3 < best) { best = s3; am = 3; }
        w[t] = best;
        choices[t] = (double)am;
        mine = best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine = fmax(mine, __shfl_xor_sync(kFull, mine, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = red[0];
        for (int k = 1; k < 8; k++) m = fmax(m, red[k]);
        block_max[blockIdx.x] = m;
    }
}

__global__ void __launch_bounds__(256) k_weights_finish(long long tot, int nblocks, const double* __restrict__ block_max,
                                                        double* __restrict__ w)
{
    __shared__ double red[8];
    __shared__ double wmax;
    double mine = -INFINITY;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) mine = fmax(mine, block_max[b]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine = fmax(mine, __shfl_xor_sync(kFull, mine, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = red[0];
        for (int k = 1; k < 8; k++) m = fmax(m, red[k]);
        wmax = m;
    }
    __syncthreads();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < tot) w[t] = wmax - w[t];  // matchND.py:216
}

// ------------------------------------------------------------------------------------- K7
// match2D.reorder_endpoints_csv (match2D.py:1001-1146): per particle, frame after frame, swap the
// two end points when the swapped displacement to the previous, already re-ordered frame is
// strictly smaller.  Whether frame f ends up swapped depends on the state s of frame f-1:
//   s = 0: swap iff comb2 < comb1;   s = 1: swap iff comb1 < comb2      (raw, un-swapped costs)
// i.e. each frame is one of four maps {0,1} -> {0,1}; composing them along the frames is a scan.
// pass A (parallel over frames x particles): the two raw costs, the map bits and min cost;
// pass B (thread per particle): compose the maps along the frames;
// pass C (parallel): apply the swaps to the 3D ends and the 2D end points of both cameras.
// ends[C,F,N,2,3]; pts2d[C,F,N,2(cam),2(end),2(xy)] optional.
__global__ void __launch_bounds__(256) k_reorder_costs(int F, int N, long long total, const double* __restrict__ ends,
                                                       uint8_t* __restrict__ bits, double* __restrict__ mincost)
{
    // t indexes (c, f>=1, particle); bits[c,f,n] = (comb2<comb1) | (comb1<comb2)<<1
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int n = (int)(t % N);
    const long long r = t / N;
    const int f = (int)(r % (F - 1)) + 1;
    const int c = (int)(r / (F - 1));
    const double* e0 = ends + (((size_t)c * F + (f - 1)) * N + n) * 6;
    const double* e1 = ends + (((size_t)c * F + f) * N + n) * 6;
    // np.linalg.norm of a 3-vector (dot product, then sqrt) -- match2D.py:1093-1098
    const double c1 = dist3(e0, e1) + dist3(e0 + 3, e1 + 3);
    const double c2 = dist3(e0 + 3, e1) + dist3(e0, e1 + 3);
    bits[((size_t)c * F + f) * N + n] = (uint8_t)((c2 < c1 ? 1 : 0) | (c1 < c2 ? 2 : 0));
    mincost[((size_t)c * (F - 1) + (f - 1)) * N + n] = (c2 < c1) ? c2 : c1;  // Python's min(comb1, comb2)
}

__global__ void k_reorder_scan(int F, int N, int CN, uint8_t* __restrict__ bits)
{
    // thread per (colour, particle): bits[c,f,n] <- swap state of frame f (0/1); frame 0 is never swapped
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= CN) return;
    const int c = t / N, n = t - c * N;
    uint8_t* b = bits + (size_t)c * F * N + n;
    int s = 0;
    b[0] = 0;
    for (int f = 1; f < F; f++) {
        const int m = b[(size_t)f * N];
        s = s ? ((m >> 1) & 1) : (m & 1);
        b[(size_t)f * N] = (uint8_t)s;
    }
}

__global__ void __launch_bounds__(256) k_reorder_apply(long long total, const uint8_t* __restrict__ state,
                                                       const double* __restrict__ ends, const double* __restrict__ pts2d,
                                                       double* __restrict__ out_ends, double* __restrict__ out_2d)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // (c, f, n)
    if (t >= total) return;
    const int s = state[t];
    const double* e = ends + (size_t)t * 6;
    double* o = out_ends + (size_t)t * 6;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        o[k] = s ? e[3 + k] : e[k];
        o[3 + k] = s ? e[k] : e[3 + k];
    }
    if (pts2d) {
        const double* q = pts2d + (size_t)t * 8;
        double* w = out_2d + (size_t)t * 8;
#pragma unroll
        for (int cam = 0; cam < 2; cam++) {
            w[cam * 4 + 0] = s ? q[cam * 4 + 2] : q[cam * 4 + 0];
            w[cam * 4 + 1] = s ? q[cam * 4 + 3] : q[cam * 4 + 1];
            w[cam * 4 + 2] = s ? q[cam * 4 + 0] : q[cam * 4 + 2];
            w[cam * 4 + 3] = s ? q[cam * 4 + 1] : q[cam * 4 + 3];
        }
    }
}

// ------------------------------------------------------------------------------------- K8 / K9
// calibrate_cameras.project_points (calibrate_cameras.py:137-199): cv2.triangulatePoints on raw
// (not undistorted) points, optional world transform.  p1, p2 [n,2] -> out [n,3].
__global__ void __launch_bounds__(128) k_triangulate_points(const __grid_constant__ DevRig rig, long long n, int to_world_,
                                                            const double* __restrict__ p1, const double* __restrict__ p2,
                                                            double* __restrict__ out)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double X, Y, Z;
    dlt_triangulate(rig, p1[2 * t], p1[2 * t + 1], p2[2 * t], p2[2 * t + 1], X, Y, Z);
    if (to_world_) {
        double wx, wy, wz;
        to_world(rig, X, Y, Z, wx, wy, wz);
        X = wx; Y = wy; Z = wz;
    }
    out[3 * t] = X; out[3 * t + 1] = Y; out[3 * t + 2] = Z;
}

// calibrate_cameras.reproject_points (calibrate_cameras.py:202-262): optional inverse world
// transform rot_inv.apply(X - trans) = Rw^T (X - tw), then cv2.projectPoints into both cameras.
__global__ void __launch_bounds__(128) k_project_points(const __grid_constant__ DevRig rig, long long n, int from_world,
                                                        const double* __restrict__ X3, double* __restrict__ uv1,
                                                        double* __restrict__ uv2)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double X = X3[3 * t], Y = X3[3 * t + 1], Z = X3[3 * t + 2];
    if (from_world) {
        const double dx = X - rig.tw[0], dy = Y - rig.tw[1], dz = Z - rig.tw[2];
        X = fma(rig.Rw[6], dz, fma(rig.Rw[3], dy, rig.Rw[0] * dx));
        Y = fma(rig.Rw[7], dz, fma(rig.Rw[4], dy, rig.Rw[1] * dx));
        Z = fma(rig.Rw[8], dz, fma(rig.Rw[5], dy, rig.Rw[2] * dx));
    }
    double u, v;
    project_point(rig.R1, rig.t1, rig.CM1, rig.k1, X, Y, Z, u, v);
    uv1[2 * t] = u; uv1[2 * t + 1] = v;
    project_point(rig.R2, rig.t2, rig.CM2, rig.k2, X, Y, Z, u, v);
    uv2[2 * t] = u; uv2[2 * t + 1] = v;
}

// ------------------------------------------------------------------------- FP64 peak probe
// Dependency-free DFMA chains: 8 accumulators per thread, every SM saturated.
__global__ void __launch_bounds__(256) k_dfma_peak(int iters, double* __restrict__ out)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 1.2345e300) out[0] = s;  // never true; keeps the chains alive
}

}  // namespace ptk
