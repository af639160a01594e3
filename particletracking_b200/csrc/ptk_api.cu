// ptk_api.cu -- C ABI of libptk.so (see include/ptk.h): argument checks, workspace carving,
// kernel launches, and the host-buffer pipeline.  No torch, no C++ types across the boundary.
//
// Build (see particletracking_b200/csrc/Makefile):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <new>
#include <utility>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX 3: ranges cost nothing unless a profiler injects itself

#include "ptk_csv.cuh"
#include "ptk_csv_reader.h"
#include "ptk_kernels.cuh"
#include "ptk_lsap_trk.cuh"
#include "ptk_ap3.cuh"

using namespace ptk;

#define PTK_API extern "C" __attribute__((visibility("default")))

namespace {

thread_local char g_cuda_err[256] = "";
std::atomic<long long> g_launches{0};

// Optional per-kernel timing (ptk_timing_enable): CUDA event pairs recorded on the launching
// stream around every kernel launch, by kernel kind, so bench.py can report the dominant
// kernel's duration (and the step's breakdown) measured inside the timed region.  Off by default
// (no events are recorded).
enum Kind { K_UNDISTORT = 0, K_COST, K_LSAP_MATCH, K_ROWS, K_ROWS_TO_ENDS, K_TRACK_COST, K_LSAP_TRACK, K_CHAIN, K_WEIGHTS, K_AP3, K_N };
std::atomic<int> g_timing{0};
std::mutex g_timing_mu;
struct TimedLaunch { int kind; cudaEvent_t t0, t1; };
std::vector<TimedLaunch> g_timed;
std::vector<cudaEvent_t> g_event_pool;

cudaEvent_t timing_event()
{
    if (!g_event_pool.empty()) {
        cudaEvent_t e = g_event_pool.back();
        g_event_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

// RAII: construct right before a launch, destroy right after it
struct LaunchTimer {
    int kind;
    cudaStream_t st;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    LaunchTimer(int k, cudaStream_t s) : kind(k), st(s)
    {
        const int mode = g_timing.load(std::memory_order_relaxed);
        if (!mode || (mode == 2 && k != K_COST)) return;  // mode 2: the dominant kernel (K1) only
        {
            std::lock_guard<std::mutex> lk(g_timing_mu);
            t0 = timing_event();
            t1 = timing_event();
        }
        cudaEventRecord(t0, st);
    }
    ~LaunchTimer()
    {
        if (!t0) return;
        cudaEventRecord(t1, st);
        std::lock_guard<std::mutex> lk(g_timing_mu);
        g_timed.push_back({kind, t0, t1});
    }
};

// NVTX range around a stage of the pipeline (visible in Nsight Systems / ncu --nvtx)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

inline int cuda_fail(cudaError_t e, const char* what)
{
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", what, cudaGetErrorString(e));
    return PTK_ERR_CUDA;
}

#define CK(call)                                              \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

#define LAUNCHED(name)                                           \
    do {                                                         \
        g_launches.fetch_add(1, std::memory_order_relaxed);      \
        cudaError_t e__ = cudaPeekAtLastError();                 \
        if (e__ != cudaSuccess) return cuda_fail(e__, name);     \
    } while (0)

DevRig make_devrig(const ptk_rig* r)
{
    DevRig d;
    memcpy(d.P1o, r->P1, sizeof(d.P1o));
    memcpy(d.P2o, r->P2, sizeof(d.P2o));
    // Static de Rijk ordering: norms of the DLT matrix's columns at the two principal points decide
    // the column order the Jacobi starts from; the columns of P1/P2 are permuted accordingly (the
    // kernel verifies the order per point and re-sorts the rare exception).
    {
        const double x1 = r->CM1[2], y1 = r->CM1[5], x2 = r->CM2[2], y2 = r->CM2[5];
        double w[4];
        for (int k = 0; k < 4; k++) {
            const double a0 = x1 * r->P1[8 + k] - r->P1[k], a1 = y1 * r->P1[8 + k] - r->P1[4 + k];
            const double a2 = x2 * r->P2[8 + k] - r->P2[k], a3 = y2 * r->P2[8 + k] - r->P2[4 + k];
            w[k] = a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
            d.perm[k] = k;
        }
        for (int i = 0; i < 3; i++) {
            int j = i;
            for (int k = i + 1; k < 4; k++)
                if (w[d.perm[k]] > w[d.perm[j]]) j = k;
            const int t = d.perm[i]; d.perm[i] = d.perm[j]; d.perm[j] = t;
        }
        for (int k = 0; k < 4; k++)
            for (int row = 0; row < 3; row++) {
                d.P1[row * 4 + k] = r->P1[row * 4 + d.perm[k]];
                d.P2[row * 4 + k] = r->P2[row * 4 + d.perm[k]];
            }
    }
    memcpy(d.R1, r->R1, sizeof(d.R1));
    memcpy(d.t1, r->t1, sizeof(d.t1));
    memcpy(d.R2, r->R2, sizeof(d.R2));
    memcpy(d.t2, r->t2, sizeof(d.t2));
    d.CM1[0] = r->CM1[0]; d.CM1[1] = r->CM1[4]; d.CM1[2] = r->CM1[2]; d.CM1[3] = r->CM1[5];
    d.CM2[0] = r->CM2[0]; d.CM2[1] = r->CM2[4]; d.CM2[2] = r->CM2[2]; d.CM2[3] = r->CM2[5];
    memcpy(d.k1, r->dist1, sizeof(d.k1));
    memcpy(d.k2, r->dist2, sizeof(d.k2));
    memcpy(d.Rw, r->Rw, sizeof(d.Rw));
    memcpy(d.tw, r->tw, sizeof(d.tw));
    d.agg_sum = r->agg_sum;
    d.agg_scale = r->agg_sum ? 1.0 : 0.5;
    const double eye[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    const double* Rs[2] = {r->R1, r->R2};
    const double* ts[2] = {r->t1, r->t2};
    const double* ks[2] = {r->dist1, r->dist2};
    for (int c = 0; c < 2; c++) {
        bool id = true;
        for (int k = 0; k < 9; k++) id = id && (Rs[c][k] == eye[k]);
        for (int k = 0; k < 3; k++) id = id && (ts[c][k] == 0.0);
        d.ext_identity[c] = id ? 1 : 0;
        bool tg = false;
        for (int k : {2, 3, 8, 9, 10, 11}) tg = tg || (ks[c][k] != 0.0);
        d.tangential[c] = tg ? 1 : 0;
    }
    {
        // rig class: what the reference's drivers build (match2D.py:584-592) is canonical and, with the
        // calibrations it ships, simple; anything else runs the general instantiation
        const double fx = r->CM1[0], fy = r->CM1[4], cx = r->CM1[2], cy = r->CM1[5];
        const double canonP1[12] = {fx, 0, cx, 0, 0, fy, cy, 0, 0, 0, 1, 0};
        bool canon = d.ext_identity[0] != 0 && r->CM1[1] == 0.0 && r->CM1[3] == 0.0 && r->CM1[6] == 0.0 && r->CM1[7] == 0.0 &&
                     r->CM1[8] == 1.0 && fx > 0.0 && fy > 0.0;
        for (int k = 0; k < 12; k++) canon = canon && (r->P1[k] == canonP1[k]);
        bool simple = !d.tangential[0] && !d.tangential[1];
        for (int c = 0; c < 2; c++)
            for (int k : {5, 6, 7}) simple = simple && (ks[c][k] == 0.0);
        d.canon = canon ? 1 : 0;
        d.simple = simple ? 1 : 0;
        d.fx2 = fx * fx;
        if (getenv("PTK_RIG_GENERAL")) d.canon = d.simple = 0;  // A/B switch: force the general instantiation
    }
    return d;
}

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Problems per internal pass of ptk_match_batch / ptk_track_batch.  Two constraints:
//  - small problems: keep the cost-matrix chunk around 64 MB so that it is still L2 resident
//    (126 MB) when K2 reads what K1 / K4 wrote;
//  - large problems: K2 is one warp per problem, so a pass needs thousands of problems to fill
//    148 SMs x 32 warps -- at least 4736 per pass, as long as the chunk stays below 2 GB
//    (a 296-problem pass of 256x256 matrices left 146 SMs' worth of warp slots empty and made
//    the N=256 tracking LSAP 10x slower; profiles/README.md).
long long match_chunk(long long P, int Nmax)
{
    const size_t per = (size_t)Nmax * Nmax * sizeof(double);
    long long c = (long long)((64ull << 20) / (per ? per : 1));
    if (c < 4736) c = 4736;
    const long long cap = (long long)((2048ull << 20) / (per ? per : 1));
    if (c > cap) c = cap;
    if (c < 296) c = 296;
    if (c > P) c = P;
    if (c < 1) c = 1;
    return c;
}

struct MatchWs {
    double *und1, *und2, *cost;
    double2* rec2;  // k_prep2's planes (canonical rigs): kCanon2RecPlanes x [chunk * Nmax * 2] double2
    size_t bytes;
};

inline size_t rec2_bytes(long long chunk, int Nmax)
{
    return (size_t)chunk * rec_problem_stride(Nmax) * sizeof(double2);
}

MatchWs carve_match(void* ws, long long chunk, int Nmax, bool need_cost)
{
    MatchWs m;
    char* b = static_cast<char*>(ws);
    size_t o = 0;
    const size_t und = align_up((size_t)chunk * Nmax * 4 * sizeof(double));
    m.und1 = reinterpret_cast<double*>(b + o); o += und;
    m.und2 = reinterpret_cast<double*>(b + o); o += und;
    m.rec2 = reinterpret_cast<double2*>(b + o); o += align_up(rec2_bytes(chunk, Nmax));
    m.cost = reinterpret_cast<double*>(b + o);
    if (need_cost) o += align_up((size_t)chunk * Nmax * Nmax * sizeof(double));
    m.bytes = o;
    return m;
}

int launch_undistort(const DevRig& rig, long long P, int Nmax, const double* cam1, const double* cam2,
                     const int32_t* n1, const int32_t* n2, double* und1, double* und2, cudaStream_t st)
{
    // gridDim.y is limited to 65535: slice the problem axis
    for (long long p0 = 0; p0 < P; p0 += 65535) {
        const int np = (int)((P - p0) < 65535 ? (P - p0) : 65535);
        // 2 * Nmax pseudo-points per camera and problem: no idle warps in the block for small problems
        const int thr = 2 * Nmax >= 128 ? 128 : ((2 * Nmax + 31) / 32) * 32;
        dim3 grid((2 * Nmax + thr - 1) / thr, np, 2);
        {
            LaunchTimer lt(K_UNDISTORT, st);
            k_undistort<<<grid, thr, 0, st>>>(rig, Nmax, cam1 + (size_t)p0 * Nmax * 4, cam2 + (size_t)p0 * Nmax * 4, n1 + p0,
                                              n2 + p0, und1 + (size_t)p0 * Nmax * 4, und2 + (size_t)p0 * Nmax * 4);
        }
        LAUNCHED("k_undistort");
    }
    return PTK_OK;
}

// rec2: workspace for k_prep2's planes, rec2_bytes(P, Nmax) (used by canonical rigs only)
int launch_cost(const DevRig& rig, long long P, int Nmax, const double* cam1, const double* cam2, const double* und1,
                const double* und2, double2* rec2, const int32_t* n1, const int32_t* n2, double* cost, uint8_t* choice,
                double* p_world, double* rep_errs, cudaStream_t st)
{
    CostArgs a;
    a.Nmax = Nmax;
    a.tiles = (Nmax * Nmax + 31) / 32;
    if (rig.canon && !rec2) return PTK_ERR_INVALID_ARG;
    // grid = (tiles or row blocks, problems): gridDim.y holds 65535
    const long long max_p = 65528;
    for (long long p0 = 0; p0 < P; p0 += max_p) {
        const long long np = (P - p0) < max_p ? (P - p0) : max_p;
        const size_t o4 = (size_t)p0 * Nmax * 4, o2 = (size_t)p0 * Nmax * Nmax;
        a.cam1 = cam1 + o4; a.cam2 = cam2 + o4; a.und1 = und1 + o4; a.und2 = und2 + o4;
        a.rec2 = rec2 ? rec2 + (size_t)p0 * rec_problem_stride(Nmax) : nullptr;
        a.n1 = n1 + p0; a.n2 = n2 + p0;
        a.cost = cost + o2;
        if (rig.canon) {  // K0b: the camera-2 share of the DLT, once per end point (ptk_device.cuh: canon_prep2)
            const int pthr = 2 * Nmax >= 128 ? 128 : ((2 * Nmax + 31) / 32) * 32;
            const dim3 pgrid((2 * Nmax + pthr - 1) / pthr, (unsigned)np);
            {
                LaunchTimer lt(K_UNDISTORT, st);
                k_prep2<<<pgrid, pthr, 0, st>>>(rig, Nmax, a.und2, a.n2, const_cast<double2*>(a.rec2));
            }
            LAUNCHED("k_prep2");
        }
        a.choice = choice ? choice + o2 : nullptr;
        a.p_world = p_world ? p_world + o2 * 12 : nullptr;
        a.rep_errs = rep_errs ? rep_errs + o2 * 8 : nullptr;
        const dim3 grid((unsigned)a.tiles, (unsigned)np);
        {
            LaunchTimer lt(K_COST, st);
            // the rig class picks the instantiation (same bits from every one of them, see eval_combo)
            static const int mb = getenv("PTK_KCOST_CFG") ? atoi(getenv("PTK_KCOST_CFG")) : 0;  // tuning knob
            static const int rows_env = getenv("PTK_KCOST_ROWS") ? atoi(getenv("PTK_KCOST_ROWS")) : 16;
            if (rig.canon) {
                RowsLoopArgs la;
                la.rows_per = rows_env > kRowsMax ? kRowsMax : (rows_env < 1 ? 1 : rows_env);
                const unsigned gx = (unsigned)((Nmax + la.rows_per - 1) / la.rows_per);
                // 5 CTAs per SM (96 registers: loop state and constants stay in registers) measured fastest:
                // 0.600 / 0.580 / 0.586 ms at 6 / 5 / 4 (profiles/README.md)
                const bool extras = choice || p_world || rep_errs;
                if (Nmax <= 32) {
                    const dim3 g(gx, (unsigned)np, 1);
                    if (extras && rig.simple) k_cost_rows<true, 32, 5, true><<<g, 128, 0, st>>>(rig, a, la);
                    else if (extras) k_cost_rows<false, 32, 5, true><<<g, 128, 0, st>>>(rig, a, la);
                    else if (rig.simple && mb == 6) k_cost_rows<true, 32, 6, false><<<g, 128, 0, st>>>(rig, a, la);
                    else if (rig.simple && mb == 4) k_cost_rows<true, 32, 4, false><<<g, 128, 0, st>>>(rig, a, la);
                    else if (rig.simple) k_cost_rows<true, 32, 5, false><<<g, 128, 0, st>>>(rig, a, la);
                    else k_cost_rows<false, 32, 5, false><<<g, 128, 0, st>>>(rig, a, la);
                } else {
                    const dim3 g(gx, (unsigned)np, (unsigned)((Nmax + 63) / 64));
                    if (extras && rig.simple) k_cost_rows<true, 64, 5, true><<<g, 128, 0, st>>>(rig, a, la);
                    else if (extras) k_cost_rows<false, 64, 5, true><<<g, 128, 0, st>>>(rig, a, la);
                    else if (rig.simple && mb == 6) k_cost_rows<true, 64, 6, false><<<g, 128, 0, st>>>(rig, a, la);
                    else if (rig.simple) k_cost_rows<true, 64, 5, false><<<g, 128, 0, st>>>(rig, a, la);
                    else k_cost_rows<false, 64, 5, false><<<g, 128, 0, st>>>(rig, a, la);
                }
            }
            else if (rig.simple) k_cost<true><<<grid, 128, 0, st>>>(rig, a);
            else k_cost<false><<<grid, 128, 0, st>>>(rig, a);
        }
        LAUNCHED("k_cost");
    }
    return PTK_OK;
}

int launch_lsap(long long P, LsapArgs a, cudaStream_t st)
{
    const int M = a.ld_rows > a.ld_cols ? a.ld_rows : a.ld_cols;
    const size_t bytes = (size_t)a.ld_rows * a.ld_cols * sizeof(double);
    a.stage = (bytes <= (size_t)kLsapStageBytes) ? 1 : 0;
    const size_t smem = ((lsap_row_state_bytes(M) + 15) / 16) * 16 + (a.stage ? bytes : 0);
    const long long max_p = 1ll << 30;
    for (long long p0 = 0; p0 < P; p0 += max_p) {
        const long long np = (P - p0) < max_p ? (P - p0) : max_p;
        LsapArgs b = a;
        if (a.cost) b.cost = a.cost + (size_t)p0 * a.cost_stride;
        if (a.nr) b.nr = a.nr + p0;
        if (a.nc) b.nc = a.nc + p0;
        if (a.row_ind) b.row_ind = a.row_ind + (size_t)p0 * a.out_ld;
        b.col_ind = a.col_ind + (size_t)p0 * a.out_ld;
        if (a.n_out) b.n_out = a.n_out + p0;
        if (a.status) b.status = a.status + p0;
        if (a.assigned) b.assigned = a.assigned + (size_t)p0 * a.out_ld;
        if (a.total) b.total = a.total + p0;
        if (a.keep) b.keep = a.keep + (size_t)p0 * a.out_ld;
        {
            LaunchTimer lt(a.total ? K_LSAP_TRACK : K_LSAP_MATCH, st);
            if (M <= 32) k_lsap<1><<<(unsigned)np, 32, smem, st>>>(b);
            else if (M <= 64) k_lsap<2><<<(unsigned)np, 32, smem, st>>>(b);
            else if (M <= 128) k_lsap<4><<<(unsigned)np, 32, smem, st>>>(b);
            else k_lsap<8><<<(unsigned)np, 32, smem, st>>>(b);
        }
        LAUNCHED("k_lsap");
    }
    return PTK_OK;
}

// K2T: the tracking assignment straight from the end points (ptk_lsap_trk.cuh): k_trk_prepare writes the
// separable norms and the exception list of every problem into `scratch`, k_lsap_trk solves, WARPS
// problems per CTA.
template <int CPL, int WARPS, int MINB>
int launch_lsap_trk_t(const TrkArgs& a, void* scratch, cudaStream_t st)
{
    const size_t smem = (size_t)WARPS * trk_warp_bytes<CPL>(a.N);
    if (smem > 48 * 1024) {
        static std::once_flag once;  // per instantiation
        static cudaError_t attr_rc = cudaSuccess;
        std::call_once(once, [&] {
            attr_rc = cudaFuncSetAttribute(k_lsap_trk<CPL, WARPS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)((size_t)WARPS * trk_warp_bytes<CPL>(32 * CPL)));
        });
        if (attr_rc != cudaSuccess) return cuda_fail(attr_rc, "cudaFuncSetAttribute(k_lsap_trk)");
    }
    // first half of the grid: problems with exceptions (k_trk_prepare's flag), second half: the rest
    const long long blocks = 2 * ((a.Q + WARPS - 1) / WARPS);
    if (blocks >= (1ll << 31)) return PTK_ERR_INVALID_ARG;
    {
        LaunchTimer lt(K_LSAP_TRACK, st);
        k_lsap_trk<CPL, WARPS, MINB><<<(unsigned)blocks, 32 * WARPS, smem, st>>>(a, scratch);
    }
    LAUNCHED("k_lsap_trk");
    return PTK_OK;
}

int launch_lsap_trk(const TrkArgs& a, void* scratch, cudaStream_t st)
{
    if (a.Q >= (1ll << 30)) return PTK_ERR_INVALID_ARG;
    {
        const size_t smem = (size_t)kTrkPrepWarps * 14 * a.N * sizeof(double);  // <= 112 KB at N = 256
        if (smem > 48 * 1024) {
            static std::once_flag once;
            static cudaError_t attr_rc = cudaSuccess;
            std::call_once(once, [&] {
                attr_rc = cudaFuncSetAttribute(k_trk_prepare, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)((size_t)kTrkPrepWarps * 14 * PTK_MAX_RODS * sizeof(double)));
            });
            if (attr_rc != cudaSuccess) return cuda_fail(attr_rc, "cudaFuncSetAttribute(k_trk_prepare)");
        }
        LaunchTimer lt(K_TRACK_COST, st);
        k_trk_prepare<<<(unsigned)((a.Q + kTrkPrepWarps - 1) / kTrkPrepWarps), 32 * kTrkPrepWarps, smem, st>>>(a, scratch);
    }
    LAUNCHED("k_trk_prepare");
    // occupancy of the solver is set by registers: 40 / 24 / 20 / 12 warps per SM
    if (a.N <= 32) {
        static const int cfg = getenv("PTK_TRK_CFG") ? atoi(getenv("PTK_TRK_CFG")) : 0;  // tuning knob (tools/trk_ab.py)
        if (cfg == 1) return launch_lsap_trk_t<1, 4, 7>(a, scratch, st);
        if (cfg == 2) return launch_lsap_trk_t<1, 2, 16>(a, scratch, st);
        if (cfg == 3) return launch_lsap_trk_t<1, 1, 32>(a, scratch, st);
        if (cfg == 4) return launch_lsap_trk_t<1, 4, 12>(a, scratch, st);
        if (cfg == 5) return launch_lsap_trk_t<1, 8, 5>(a, scratch, st);
        return launch_lsap_trk_t<1, 4, 10>(a, scratch, st);
    }
    if (a.N <= 64) return launch_lsap_trk_t<2, 4, 6>(a, scratch, st);
    if (a.N <= 128) return launch_lsap_trk_t<4, 4, 5>(a, scratch, st);
    return launch_lsap_trk_t<8, 4, 3>(a, scratch, st);
}

// Where K3 also writes the tracking layout ends[C,F,N,2,3] (replaces a separate gather kernel): the call's
// problem 0 is global problem p0 = frame * C + colour.
struct TrkTarget {
    double* ends;
    int F, C, N;
    long long p0;
};

int launch_rows(const DevRig& rig, long long P, int Nmax, const double* cam1, const double* cam2, const double* und1,
                const double* und2, const int32_t* col_ind, const int32_t* n_out, double* rows, cudaStream_t st,
                double* pair_cost = nullptr, double* ends6 = nullptr, uint8_t* straight = nullptr,
                const TrkTarget* trk = nullptr)
{
    for (long long p0 = 0; p0 < P; p0 += 65535) {
        const int np = (int)((P - p0) < 65535 ? (P - p0) : 65535);
        RowsArgs a;
        a.Nmax = Nmax;
        const size_t o4 = (size_t)p0 * Nmax * 4;
        a.cam1 = cam1 + o4; a.cam2 = cam2 + o4; a.und1 = und1 + o4; a.und2 = und2 + o4;
        a.col_ind = col_ind + (size_t)p0 * Nmax;
        a.n_out = n_out + p0;
        a.rows = rows ? rows + (size_t)p0 * Nmax * PTK_ROW_COLS : nullptr;
        a.pair_cost = pair_cost ? pair_cost + (size_t)p0 * Nmax : nullptr;
        a.ends6 = ends6 ? ends6 + (size_t)p0 * Nmax * 6 : nullptr;
        a.straight = straight ? straight + (size_t)p0 * Nmax : nullptr;
        a.ends_trk = trk ? trk->ends : nullptr;
        a.trk_F = trk ? trk->F : 0; a.trk_C = trk ? trk->C : 1; a.trk_N = trk ? trk->N : 0;
        a.trk_p0 = (trk ? trk->p0 : 0) + p0;
        a.trk_rC = 1.0 / (double)a.trk_C;
        const int thr = Nmax >= 32 ? 128 : ((4 * Nmax + 31) / 32) * 32;  // 4 lanes per row, up to 32 rows per block
        dim3 grid((4 * Nmax + thr - 1) / thr, np);
        {
            LaunchTimer lt(K_ROWS, st);
            if (rig.canon && rig.simple) k_rows<true, true><<<grid, thr, 0, st>>>(rig, a);
            else if (rig.canon) k_rows<true, false><<<grid, thr, 0, st>>>(rig, a);
            else if (rig.simple) k_rows<false, true><<<grid, thr, 0, st>>>(rig, a);
            else k_rows<false, false><<<grid, thr, 0, st>>>(rig, a);
        }
        LAUNCHED("k_rows");
    }
    return PTK_OK;
}

bool bad_n(int Nmax) { return Nmax <= 0 || Nmax > PTK_MAX_RODS; }

// NULL, or a camera model the kernels do not implement: cv2's tilted-sensor coefficients tauX, tauY
// (dist[12], dist[13]) -- rejected rather than silently computed with a different model.
bool bad_rig(const ptk_rig* r)
{
    return !r || r->dist1[12] != 0.0 || r->dist1[13] != 0.0 || r->dist2[12] != 0.0 || r->dist2[13] != 0.0;
}

}  // namespace

// ------------------------------------------------------------------------------------ library
PTK_API int ptk_version(void) { return PTK_VERSION; }

PTK_API const char* ptk_strerror(int status)
{
    switch (status) {
        case PTK_OK: return "ok";
        case PTK_ERR_INVALID_ARG: return "invalid argument";
        case PTK_ERR_CUDA: return "CUDA runtime error (see ptk_last_cuda_error)";
        case PTK_ERR_WORKSPACE: return "workspace too small (see ptk_workspace_bytes)";
        case PTK_ERR_NO_DEVICE: return "no usable CUDA device (need compute capability 10.x)";
        case PTK_ERR_NUMERIC: return "cost matrix contains invalid numeric entries or is infeasible";
        default: return "unknown status";
    }
}

PTK_API const char* ptk_last_cuda_error(void) { return g_cuda_err; }

PTK_API int ptk_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int d = 0; d < n; d++) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ok++;
    }
    return ok;
}

// Debug build (-DPTK_BOUNDS_CHECK, libptk_dbg.so): the device-side log of index violations.  Release: enabled = 0.
PTK_API int ptk_bounds_report(int* enabled, long long* count, int* first_site, long long* first_idx, long long* first_bound,
                              int reset)
{
    if (enabled) *enabled = 0;
    if (count) *count = 0;
    if (first_site) *first_site = 0;
    if (first_idx) *first_idx = 0;
    if (first_bound) *first_bound = 0;
#ifdef PTK_BOUNDS_CHECK
    BoundsLog h;
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpyFromSymbol(&h, g_bounds_log, sizeof(h)));
    if (enabled) *enabled = 1;
    if (count) *count = (long long)h.count;
    if (first_site) *first_site = h.first_site;
    if (first_idx) *first_idx = h.first_idx;
    if (first_bound) *first_bound = h.first_bound;
    if (reset) {
        BoundsLog z = {0ull, 0ll, 0ll, 0, 1};
        CK(cudaMemcpyToSymbol(g_bounds_log, &z, sizeof(z)));
    }
#else
    (void)reset;
#endif
    return PTK_OK;
}

PTK_API long long ptk_launch_count(int reset)
{
    return reset ? g_launches.exchange(0) : g_launches.load();
}

PTK_API void ptk_timing_enable(int on) { g_timing.store(on == 2 ? 2 : (on ? 1 : 0)); }

PTK_API int ptk_timing_read(double* ms_by_kind, long long* launches_by_kind, int reset)
{
    // caller must have synchronised the streams the work ran on; arrays hold PTK_TIMING_KINDS entries
    std::lock_guard<std::mutex> lk(g_timing_mu);
    for (int k = 0; k < K_N; k++) {
        if (ms_by_kind) ms_by_kind[k] = 0;
        if (launches_by_kind) launches_by_kind[k] = 0;
    }
    for (auto& tl : g_timed) {
        float t = 0;
        CK(cudaEventElapsedTime(&t, tl.t0, tl.t1));
        if (ms_by_kind) ms_by_kind[tl.kind] += t;
        if (launches_by_kind) launches_by_kind[tl.kind] += 1;
    }
    if (reset) {
        for (auto& tl : g_timed) { g_event_pool.push_back(tl.t0); g_event_pool.push_back(tl.t1); }
        g_timed.clear();
    }
    return PTK_OK;
}

PTK_API size_t ptk_workspace_bytes(int kind, long long n_problems, int Nmax)
{
    if (n_problems < 0) return 0;
    switch (kind) {
        case PTK_WS_COST: {
            if (bad_n(Nmax)) return 0;
            return carve_match(nullptr, n_problems, Nmax, false).bytes + 256;
        }
        case PTK_WS_MATCH: {
            if (bad_n(Nmax)) return 0;
            return carve_match(nullptr, match_chunk(n_problems, Nmax), Nmax, true).bytes + 256;
        }
        case PTK_WS_TRACK: {
            if (bad_n(Nmax)) return 0;
            return align_up((size_t)n_problems * trk_prep_bytes(Nmax)) + 256;  // k_trk_prepare's records
        }
        case PTK_WS_WEIGHTS:
            return align_up((size_t)((n_problems + 255) / 256) * sizeof(double)) + 256;
        case PTK_WS_AP3:
            return align_up((size_t)n_problems * 2 * 32768 * sizeof(Ap3Cand)) + 256;
        default: return 0;
    }
}

// ------------------------------------------------------------------------------------ K0
PTK_API int ptk_undistort_batch(const ptk_rig* rig, int P, int Nmax, const double* cam1, const double* cam2,
                                const int32_t* n1, const int32_t* n2, double* und1, double* und2, void* stream)
{
    if (bad_rig(rig) || P < 0 || bad_n(Nmax) || !cam1 || !cam2 || !n1 || !n2 || !und1 || !und2) return PTK_ERR_INVALID_ARG;
    if (P == 0) return PTK_OK;
    return launch_undistort(make_devrig(rig), P, Nmax, cam1, cam2, n1, n2, und1, und2, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------ K0+K1
PTK_API int ptk_cost_batch(const ptk_rig* rig, int P, int Nmax, const double* cam1, const double* cam2,
                           const int32_t* n1, const int32_t* n2, double* cost, uint8_t* choice, double* p_world,
                           double* rep_errs, void* workspace, size_t workspace_bytes, void* stream)
{
    if (bad_rig(rig) || P < 0 || bad_n(Nmax) || !cam1 || !cam2 || !n1 || !n2 || !cost) return PTK_ERR_INVALID_ARG;
    if (P == 0) return PTK_OK;
    if (!workspace || workspace_bytes < ptk_workspace_bytes(PTK_WS_COST, P, Nmax)) return PTK_ERR_WORKSPACE;
    const DevRig d = make_devrig(rig);
    cudaStream_t st = (cudaStream_t)stream;
    MatchWs w = carve_match((void*)align_up((size_t)workspace), P, Nmax, false);
    int rc = launch_undistort(d, P, Nmax, cam1, cam2, n1, n2, w.und1, w.und2, st);
    if (rc) return rc;
    return launch_cost(d, P, Nmax, cam1, cam2, w.und1, w.und2, w.rec2, n1, n2, cost, choice, p_world, rep_errs, st);
}

// ------------------------------------------------------------------------------------ K2
PTK_API int ptk_lsap_batch(int P, int ld_rows, int ld_cols, const double* cost, const int32_t* nr, const int32_t* nc,
                           int32_t* row_ind, int32_t* col_ind, int32_t* n_out, int32_t* status, void* stream)
{
    if (P < 0 || bad_n(ld_rows) || bad_n(ld_cols) || !cost || !row_ind || !col_ind) return PTK_ERR_INVALID_ARG;
    if (P == 0) return PTK_OK;
    LsapArgs a{};
    a.ld_rows = ld_rows; a.ld_cols = ld_cols;
    a.cost_stride = (size_t)ld_rows * ld_cols;
    a.cost = cost; a.nr = nr; a.nc = nc;
    a.out_ld = ld_rows < ld_cols ? ld_rows : ld_cols;
    a.row_ind = row_ind; a.col_ind = col_ind; a.n_out = n_out; a.status = status;
    a.max_err = INFINITY;
    return launch_lsap(P, a, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------ K0..K3
namespace {
// K0..K3 with either output form: the 18-column rows, or the compact pair (ends6, straight), or both
int match_batch_impl(const ptk_rig* rig, int P, int Nmax, const double* cam1, const double* cam2,
                     const int32_t* n1, const int32_t* n2, double* cost, uint8_t* choice, int32_t* row_ind,
                     int32_t* col_ind, int32_t* n_out, int32_t* status, double* rows, double* ends6, uint8_t* straight,
                     double* match_cost, double max_repr_err, uint8_t* keep_mask, void* workspace, size_t workspace_bytes,
                     void* stream, const TrkTarget* trk = nullptr)
{
    if (bad_rig(rig) || P < 0 || bad_n(Nmax) || !cam1 || !cam2 || !n1 || !n2 || !row_ind || !col_ind || !n_out || !status ||
        (!rows && !(ends6 && straight)) || !match_cost)
        return PTK_ERR_INVALID_ARG;
    if (P == 0) return PTK_OK;
    if (!workspace || workspace_bytes < ptk_workspace_bytes(PTK_WS_MATCH, P, Nmax)) return PTK_ERR_WORKSPACE;
    const DevRig d = make_devrig(rig);
    cudaStream_t st = (cudaStream_t)stream;
    const long long chunk = match_chunk(P, Nmax);
    MatchWs w = carve_match((void*)align_up((size_t)workspace), chunk, Nmax, true);
    const size_t NN = (size_t)Nmax * Nmax;
    TrkTarget tchunk{};
    for (long long p0 = 0; p0 < P; p0 += chunk) {
        const long long np = (P - p0) < chunk ? (P - p0) : chunk;
        const size_t o4 = (size_t)p0 * Nmax * 4;
        double* cchunk = cost ? cost + (size_t)p0 * NN : w.cost;
        int rc = launch_undistort(d, np, Nmax, cam1 + o4, cam2 + o4, n1 + p0, n2 + p0, w.und1, w.und2, st);
        if (rc) return rc;
        rc = launch_cost(d, np, Nmax, cam1 + o4, cam2 + o4, w.und1, w.und2, w.rec2, n1 + p0, n2 + p0, cchunk,
                         choice ? choice + (size_t)p0 * NN : nullptr, nullptr, nullptr, st);
        if (rc) return rc;
        LsapArgs a{};
        a.ld_rows = Nmax; a.ld_cols = Nmax; a.cost_stride = NN; a.cost = cchunk;
        a.nr = n1 + p0; a.nc = n2 + p0; a.out_ld = Nmax;
        a.row_ind = row_ind + (size_t)p0 * Nmax; a.col_ind = col_ind + (size_t)p0 * Nmax;
        a.n_out = n_out + p0; a.status = status + p0;
        a.assigned = match_cost + (size_t)p0 * Nmax;
        a.keep = keep_mask ? keep_mask + (size_t)p0 * Nmax : nullptr;
        a.max_err = max_repr_err;
        rc = launch_lsap(np, a, st);
        if (rc) return rc;
        rc = launch_rows(d, np, Nmax, cam1 + o4, cam2 + o4, w.und1, w.und2, col_ind + (size_t)p0 * Nmax, n_out + p0,
                         rows ? rows + (size_t)p0 * Nmax * PTK_ROW_COLS : nullptr, st, nullptr,
                         ends6 ? ends6 + (size_t)p0 * Nmax * 6 : nullptr, straight ? straight + (size_t)p0 * Nmax : nullptr,
                         trk ? &(tchunk = TrkTarget{trk->ends, trk->F, trk->C, trk->N, trk->p0 + p0}) : nullptr);
        if (rc) return rc;
    }
    return PTK_OK;
}
}  // namespace

PTK_API int ptk_match_batch(const ptk_rig* rig, int P, int Nmax, const double* cam1, const double* cam2,
                            const int32_t* n1, const int32_t* n2, double* cost, uint8_t* choice, int32_t* row_ind,
                            int32_t* col_ind, int32_t* n_out, int32_t* status, double* rows, double* match_cost,
                            double max_repr_err, uint8_t* keep_mask, void* workspace, size_t workspace_bytes,
                            void* stream)
{
    if (!rows) return PTK_ERR_INVALID_ARG;
    return match_batch_impl(rig, P, Nmax, cam1, cam2, n1, n2, cost, choice, row_ind, col_ind, n_out, status, rows, nullptr,
                            nullptr, match_cost, max_repr_err, keep_mask, workspace, workspace_bytes, stream);
}


// ------------------------------------------------------------------------------------ K0+K3
PTK_API int ptk_rows_batch(const ptk_rig* rig, int P, int Nmax, const double* cam1, const double* cam2,
                           const int32_t* n1, const int32_t* n2, const int32_t* col_ind, const int32_t* n_out,
                           double* rows, double* pair_cost, void* workspace, size_t workspace_bytes, void* stream)
{
    if (bad_rig(rig) || P < 0 || bad_n(Nmax) || !cam1 || !cam2 || !n1 || !n2 || !col_ind || !n_out || !rows)
        return PTK_ERR_INVALID_ARG;
    if (P == 0) return PTK_OK;
    if (!workspace || workspace_bytes < ptk_workspace_bytes(PTK_WS_COST, P, Nmax)) return PTK_ERR_WORKSPACE;
    const DevRig d = make_devrig(rig);
    cudaStream_t st = (cudaStream_t)stream;
    MatchWs w = carve_match((void*)align_up((size_t)workspace), P, Nmax, false);
    int rc = launch_undistort(d, P, Nmax, cam1, cam2, n1, n2, w.und1, w.und2, st);
    if (rc) return rc;
    return launch_rows(d, P, Nmax, cam1, cam2, w.und1, w.und2, col_ind, n_out, rows, st, pair_cost);
}

// rows[F,C,Nmax,18] -> ends[C,F,N,2,3] (tracking.py:97-98 extracts the same six columns)
PTK_API int ptk_rows_to_ends(int F, int C, int N, int Nmax, const double* rows, double* ends, void* stream)
{
    if (F < 0 || C < 0 || bad_n(Nmax) || N < 0 || N > Nmax || !rows || !ends) return PTK_ERR_INVALID_ARG;
    const long long tot = (long long)F * C * N * 6;
    if (tot == 0) return PTK_OK;
    {
        LaunchTimer lt(K_ROWS_TO_ENDS, (cudaStream_t)stream);
        k_rows_to_ends<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(0, F, F, C, N, Nmax, PTK_ROW_COLS, rows, ends);
    }
    LAUNCHED("k_rows_to_ends");
    return PTK_OK;
}

// ------------------------------------------------------------------------------------ K4 (+K2)
PTK_API int ptk_track_batch(int C, int F, int N, const double* ends, double* cost, int32_t* col_ind,
                            double* total_cost, int32_t* status, void* workspace, size_t workspace_bytes, void* stream)
{
    if (C < 0 || F < 1 || bad_n(N) || !ends || !col_ind || !total_cost) return PTK_ERR_INVALID_ARG;
    const long long Q = (long long)C * (F - 1);
    if (Q == 0) return PTK_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t NN = (size_t)N * N;
    if (!cost) {
        // no matrix wanted: solve straight from the end points (k_trk_prepare + k_lsap_trk) -- no
        // N^2 matrix in HBM or shared memory for any N; the workspace holds 32..60 N bytes per problem
        if (!workspace || workspace_bytes < ptk_workspace_bytes(PTK_WS_TRACK, Q, N)) return PTK_ERR_WORKSPACE;
        TrkArgs a{};
        a.N = N; a.F = F; a.Q = Q;
        a.ends = ends; a.col_ind = col_ind; a.total = total_cost; a.status = status;
        return launch_lsap_trk(a, (void*)align_up((size_t)workspace), st);
    }
    // the caller wants the matrix (tests, diagnostics): K4 writes it, the generic K2 solves it
    (void)workspace; (void)workspace_bytes;
    const long long chunk = Q;
    for (long long q0 = 0; q0 < Q; q0 += chunk) {
        const long long nq = (Q - q0) < chunk ? (Q - q0) : chunk;
        double* cchunk = cost + (size_t)q0 * NN;
        {
            LaunchTimer lt(K_TRACK_COST, st);
            k_track_cost<<<(unsigned)nq, 256, (size_t)14 * N * sizeof(double), st>>>(N, F, q0, ends, cchunk);
        }
        LAUNCHED("k_track_cost");
        LsapArgs a{};
        a.ld_rows = N; a.ld_cols = N; a.cost_stride = NN; a.cost = cchunk;
        a.out_ld = N;
        a.col_ind = col_ind + (size_t)q0 * N;
        a.status = status ? status + q0 : nullptr;
        a.total = total_cost + q0;
        a.max_err = INFINITY;
        int rc = launch_lsap(nq, a, st);
        if (rc) return rc;
    }
    return PTK_OK;
}

// ------------------------------------------------------------------------------------ K5
namespace {
// F frames are chained; the first F_out of them are stored in track_id[C,F_out,N], frame F_out (if
// F_out < F: the halo frame of a shard) in halo_map[C,N].  scratch: chain_scratch_ints(C, F, N).
int chain_impl(int C, int F, int N, const int32_t* col_ind, const int32_t* ids_in, int F_out, int32_t* track_id,
               int32_t* halo_map, int32_t* scratch, cudaStream_t st)
{
    const int nseg = (F + kChainSeg - 1) / kChainSeg;
    int32_t* seg_map = scratch;
    int32_t* seg_start = scratch + (size_t)C * nseg * N;
    int32_t* loc = seg_start + (size_t)C * nseg * N;
    const int threads = N < 128 ? 128 : ((N + 31) / 32 * 32);
    const size_t smem = (2 + (size_t)kChainSeg) * N * sizeof(int32_t);
    const long long total = (long long)C * F * N;
    {
        LaunchTimer lt(K_CHAIN, st);
        k_chain_local<<<dim3(nseg, C), threads, smem, st>>>(F, N, nseg, col_ind, loc, seg_map);
        LAUNCHED("k_chain_local");
        k_chain_scan<<<C, threads, smem, st>>>(N, nseg, ids_in, seg_map, seg_start);
        LAUNCHED("k_chain_scan");
        k_chain_apply<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(F, N, nseg, seg_start, loc, F_out, track_id, halo_map, total);
        LAUNCHED("k_chain_apply");
    }
    return PTK_OK;
}
size_t chain_scratch_ints(int C, int F, int N)
{
    const int nseg = (F + kChainSeg - 1) / kChainSeg;
    return 2 * (size_t)C * nseg * N + (size_t)C * F * N;
}
}  // namespace

PTK_API int ptk_chain_tracks(int C, int F, int N, const int32_t* col_ind, const int32_t* ids_in, int32_t* track_id,
                             void* stream)
{
    if (C < 0 || F < 1 || bad_n(N) || !track_id || (F > 1 && !col_ind)) return PTK_ERR_INVALID_ARG;
    if (C == 0) return PTK_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // segment maps are tiny (2*C*ceil(F/32)*N ints); stream-ordered allocation keeps the ABI simple
    int32_t* scratch = nullptr;
    CK(cudaMallocAsync((void**)&scratch, chain_scratch_ints(C, F, N) * sizeof(int32_t) + 16, st));
    int rc = chain_impl(C, F, N, col_ind, ids_in, F, track_id, nullptr, scratch, st);
    cudaFreeAsync(scratch, st);
    return rc;
}

PTK_API int ptk_chain_rebase(int W, int C, int F, int N, int rank, const int32_t* maps, int32_t* start,
                             int32_t* track_id, void* stream)
{
    if (W < 1 || C < 0 || F < 0 || bad_n(N) || rank < 0 || rank >= W || !maps || !start || !track_id)
        return PTK_ERR_INVALID_ARG;
    if (C == 0) return PTK_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int threads = N < 32 ? 32 : ((N + 31) / 32 * 32);
    {
        LaunchTimer lt(K_CHAIN, st);
        k_chain_compose<<<C, threads, 2 * (size_t)N * sizeof(int32_t), st>>>(W, C, N, rank, maps, (size_t)C * N, start);
        LAUNCHED("k_chain_compose");
        const long long total = (long long)C * F * N;
        if (total > 0) {
            k_chain_rebase<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(F, N, start, track_id, total);
            LAUNCHED("k_chain_rebase");
        }
    }
    return PTK_OK;
}

// ------------------------------------------------------------------------------------ whole path, device buffers
namespace {
struct DevWs { size_t match, ends, track, chain, total; };
DevWs device_ws(int F, int C, int Nmax, int do_track)
{
    DevWs w{};
    const long long P = (long long)F * C, Q = (long long)C * (F > 0 ? F - 1 : 0);
    w.match = align_up(ptk_workspace_bytes(PTK_WS_MATCH, P, Nmax));
    if (do_track) {
        w.ends = align_up((size_t)C * F * Nmax * 6 * sizeof(double));
        w.track = align_up(ptk_workspace_bytes(PTK_WS_TRACK, Q > 0 ? Q : 1, Nmax));
        w.chain = align_up(chain_scratch_ints(C, F, Nmax) * sizeof(int32_t) + 16);
    }
    w.total = w.match + w.ends + w.track + w.chain + 256;
    return w;
}
}  // namespace

PTK_API size_t ptk_reconstruct_device_workspace_bytes(int F, int C, int Nmax, int do_track)
{
    if (F < 0 || C < 0 || bad_n(Nmax)) return 0;
    return device_ws(F, C, Nmax, do_track).total;
}

PTK_API size_t ptk_reconstruct_shard_workspace_bytes(int F, int n_halo, int C, int Nmax)
{
    if (F < 0 || n_halo < 0 || n_halo > 1 || C < 0 || bad_n(Nmax)) return 0;
    return device_ws(F + n_halo, C, Nmax, 1).total;
}

// One rank's frame range of a sharded run: F own frames followed by n_halo (0 or 1) frames that belong
// to the next rank and are re-computed here, so that the boundary pair is tracked locally -- the
// matching is a pure function of its inputs, so the re-computed frame has the bits the owner gets.
PTK_API int ptk_reconstruct_shard(const ptk_rig* rig, int F, int n_halo, int C, int Nmax, const double* cam1,
                                  const double* cam2, const int32_t* n1, const int32_t* n2, double* rows, double* match_cost,
                                  int32_t* row_ind, int32_t* col_ind, int32_t* n_out, int32_t* status, double max_repr_err,
                                  uint8_t* keep_mask, int32_t* track_col, double* track_total, int32_t* track_status,
                                  int32_t* track_id, int32_t* halo_map, void* workspace, size_t workspace_bytes, void* stream)
{
    if (bad_rig(rig) || F < 1 || n_halo < 0 || n_halo > 1 || C < 0 || bad_n(Nmax)) return PTK_ERR_INVALID_ARG;
    if (!track_col || !track_total || !track_id || (n_halo && !halo_map)) return PTK_ERR_INVALID_ARG;
    const int Ft = F + n_halo;
    const long long P = (long long)Ft * C;
    if (P == 0) return PTK_OK;
    const DevWs w = device_ws(Ft, C, Nmax, 1);
    if (!workspace || workspace_bytes < w.total) return PTK_ERR_WORKSPACE;
    char* base = (char*)align_up((size_t)workspace);
    NvtxRange nv("ptk_reconstruct_shard");
    int rc;
    {
        NvtxRange nv1("match");
        if (!rows) return PTK_ERR_INVALID_ARG;
        // K3 writes the tracking layout beside the rows (no separate gather)
        const TrkTarget trk{reinterpret_cast<double*>(base + w.match), Ft, C, Nmax, 0};
        rc = match_batch_impl(rig, (int)P, Nmax, cam1, cam2, n1, n2, nullptr, nullptr, row_ind, col_ind, n_out, status, rows,
                              nullptr, nullptr, match_cost, max_repr_err, keep_mask, base, w.match, stream, &trk);
        if (rc) return rc;
    }
    double* ends = reinterpret_cast<double*>(base + w.match);
    {
        NvtxRange nv2("track");
        if (Ft > 1) {
            rc = ptk_track_batch(C, Ft, Nmax, ends, nullptr, track_col, track_total, track_status, base + w.match + w.ends, w.track,
                                 stream);
            if (rc) return rc;
        }
    }
    NvtxRange nv3("chain");
    return chain_impl(C, Ft, Nmax, track_col, nullptr, F, track_id, halo_map,
                      reinterpret_cast<int32_t*>(base + w.match + w.ends + w.track), (cudaStream_t)stream);
}

PTK_API int ptk_reconstruct_device(const ptk_rig* rig, int F, int C, int Nmax, const double* cam1, const double* cam2,
                                   const int32_t* n1, const int32_t* n2, double* rows, double* match_cost,
                                   int32_t* row_ind, int32_t* col_ind, int32_t* n_out, int32_t* status,
                                   double max_repr_err, uint8_t* keep_mask, int do_track, int32_t* track_col,
                                   double* track_total, int32_t* track_id, int32_t* track_status, void* workspace,
                                   size_t workspace_bytes, void* stream)
{
    if (bad_rig(rig) || F < 0 || C < 0 || bad_n(Nmax)) return PTK_ERR_INVALID_ARG;
    if (do_track && (!track_col || !track_total || F < 1)) return PTK_ERR_INVALID_ARG;
    const long long P = (long long)F * C;
    if (P == 0) return PTK_OK;
    const DevWs w = device_ws(F, C, Nmax, do_track);
    if (!workspace || workspace_bytes < w.total) return PTK_ERR_WORKSPACE;
    char* base = (char*)align_up((size_t)workspace);
    NvtxRange nv("ptk_reconstruct_device");
    int rc;
    {
        NvtxRange nv1("match");
        if (!rows) return PTK_ERR_INVALID_ARG;
        // with tracking, K3 writes the tracking layout beside the rows (no separate gather)
        const TrkTarget trk{reinterpret_cast<double*>(base + w.match), F, C, Nmax, 0};
        rc = match_batch_impl(rig, (int)P, Nmax, cam1, cam2, n1, n2, nullptr, nullptr, row_ind, col_ind, n_out, status, rows,
                              nullptr, nullptr, match_cost, max_repr_err, keep_mask, base, w.match, stream,
                              do_track ? &trk : nullptr);
    }
    if (rc || !do_track) return rc;
    double* ends = reinterpret_cast<double*>(base + w.match);
    {
        NvtxRange nv2("track");
        if (F > 1) {
            rc = ptk_track_batch(C, F, Nmax, ends, nullptr, track_col, track_total, track_status, base + w.match + w.ends, w.track,
                                 stream);
            if (rc) return rc;
        }
    }
    if (track_id) {
        NvtxRange nv3("chain");
        rc = chain_impl(C, F, Nmax, track_col, nullptr, F, track_id, nullptr,
                        reinterpret_cast<int32_t*>(base + w.match + w.ends + w.track), (cudaStream_t)stream);
    }
    return rc;
}

// Root side of the sharded chain pass: blocks[W] are the gathered per-rank result blocks (block r at
// blocks + r * block_stride bytes) holding local ids track_id[C,F_r,N] at tid_off and the rank's halo map
// [C,N] at map_off.  Composes the maps up to every rank and re-bases that rank's ids in place.
PTK_API int ptk_chain_rebase_gathered(int W, int C, int N, const int32_t* frames_per_rank, void* blocks, size_t block_stride,
                                      size_t tid_off, size_t map_off, int32_t* frames_dev, void* stream)
{
    if (W < 1 || W > 1024 || C < 0 || bad_n(N) || !frames_per_rank || !blocks || !frames_dev || (block_stride % 4) ||
        (tid_off % 4) || (map_off % 4))
        return PTK_ERR_INVALID_ARG;
    if (C == 0 || W == 1) return PTK_OK;
    for (int r = 0; r < W; r++)
        if (frames_per_rank[r] < 0) return PTK_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    // the shard sizes are launch parameters of one kernel for all ranks: 4 W bytes through frames_dev
    // (a pageable-to-device copy of a few bytes is staged by the driver before the call returns)
    CK(cudaMemcpyAsync(frames_dev, frames_per_rank, (size_t)W * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    {
        LaunchTimer lt(K_CHAIN, st);
        k_chain_rebase_gathered<<<dim3(C, W - 1), 256, 2 * (size_t)N * sizeof(int32_t), st>>>(
            C, N, frames_dev, static_cast<int32_t*>(blocks), block_stride / 4, tid_off / 4, map_off / 4);
    }
    LAUNCHED("k_chain_rebase_gathered");
    return PTK_OK;
}

// ------------------------------------------------------------------------------------ K6
PTK_API int ptk_create_weights(int Np, int N1, int N2, const double* p_3D, const double* p_prev, const double* rep_errs,
                               double* weights, double* choices, void* workspace, size_t workspace_bytes, void* stream)
{
    if (Np < 0 || N1 < 0 || N2 < 0 || !p_3D || !p_prev || !rep_errs || !weights || !choices) return PTK_ERR_INVALID_ARG;
    const long long tot = (long long)Np * N1 * N2;
    if (tot == 0) return PTK_OK;
    if (!workspace || workspace_bytes < ptk_workspace_bytes(PTK_WS_WEIGHTS, tot, 0)) return PTK_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    double* bm = reinterpret_cast<double*>(align_up((size_t)workspace));
    const int nb = (int)((tot + 255) / 256);
    k_weights_raw<<<nb, 256, 0, st>>>(Np, N1, N2, p_3D, p_prev, rep_errs, weights, choices, bm);
    LAUNCHED("k_weights_raw");
    k_weights_finish<<<nb, 256, 0, st>>>(tot, nb, bm, weights);
    LAUNCHED("k_weights_finish");
    return PTK_OK;
}

// ------------------------------------------------------------------------------------ K10
static const int kAp3CandCap = 32768, kAp3Steps = 60, kAp3Subgradient = 1200;
static const long long kAp3Nodes = 2000000LL;
PTK_API int ptk_npartite3_batch(int B, int D0, int D1, int D2, const double* weights, const int32_t* n0, const int32_t* n1,
                                const int32_t* n2, int32_t* sol, int32_t* n_sol, double* value, double* bound,
                                int32_t* status, int32_t* stats, int max_steps, long long node_limit, void* workspace,
                                size_t workspace_bytes, void* stream)
{
    if (B < 0 || D0 < 0 || D1 < 0 || D2 < 0 || D0 > AP3_MAXN || D1 > AP3_MAXN || D2 > AP3_MAXN) return PTK_ERR_INVALID_ARG;
    if (!n_sol || !value || !bound || !status) return PTK_ERR_INVALID_ARG;
    if (B == 0) return PTK_OK;
    const long long tot = (long long)D0 * D1 * D2;
    if (tot > 0 && (!weights || !sol)) return PTK_ERR_INVALID_ARG;
    if (!workspace || workspace_bytes < ptk_workspace_bytes(PTK_WS_AP3, B, 0)) return PTK_ERR_WORKSPACE;
    Ap3Args a;
    a.w = weights;
    a.D0 = D0; a.D1 = D1; a.D2 = D2;
    a.n0 = n0; a.n1 = n1; a.n2 = n2;
    a.sol = sol; a.n_sol = n_sol; a.value = value; a.bound = bound; a.status = status; a.stats = stats;
    a.cand = reinterpret_cast<Ap3Cand*>(align_up((size_t)workspace));
    a.cand_cap = kAp3CandCap;
    a.max_steps = max_steps > 0 ? max_steps : kAp3Steps;
    a.max_sg = kAp3Subgradient;
    a.node_limit = node_limit > 0 ? node_limit : kAp3Nodes;
    static_assert(sizeof(Ap3Smem) <= 48 * 1024, "k_ap3 shared memory");
    cudaStream_t st = (cudaStream_t)stream;
    LaunchTimer tm(K_AP3, st);
    k_ap3<<<B, AP3_THREADS, sizeof(Ap3Smem), st>>>(a);
    LAUNCHED("k_ap3");
    return PTK_OK;
}

// ------------------------------------------------------------------------------------ matchND.assign
namespace {
struct NdWs {
    double *und1, *und2, *cost, *pw, *re, *w, *value, *bound;
    double2* rec2;
    uint8_t* choices;
    unsigned long long* wkey;
    int32_t *sol, *n_sol, *ap3_status, *stats, *row_ind, *col_ind;
    Ap3Cand* cand;
    void* match_ws;
    size_t match_ws_bytes;
    size_t bytes;
};
NdWs carve_nd(void* base, int F, int C, int D)
{
    NdWs w;
    char* b = static_cast<char*>(base);
    size_t o = 0;
    auto take = [&](size_t n) { char* p = b + o; o += align_up(n); return p; };
    const size_t CD = (size_t)C * D, CDD = CD * D, CDDD = CDD * D;
    w.und1 = (double*)take(CD * 4 * sizeof(double));
    w.und2 = (double*)take(CD * 4 * sizeof(double));
    w.rec2 = (double2*)take(rec2_bytes(C, D));
    w.cost = (double*)take(CDD * sizeof(double));
    w.pw = (double*)take(CDD * 12 * sizeof(double));
    w.re = (double*)take(CDD * 8 * sizeof(double));
    w.w = (double*)take(CDDD * sizeof(double));
    w.choices = (uint8_t*)take(CDDD);
    w.wkey = (unsigned long long*)take((size_t)F * C * sizeof(unsigned long long));
    w.sol = (int32_t*)take(CD * 3 * sizeof(int32_t));
    w.n_sol = (int32_t*)take((size_t)C * sizeof(int32_t));
    w.ap3_status = (int32_t*)take((size_t)C * sizeof(int32_t));
    w.stats = (int32_t*)take((size_t)C * 4 * sizeof(int32_t));
    w.value = (double*)take((size_t)C * sizeof(double));
    w.bound = (double*)take((size_t)C * sizeof(double));
    w.row_ind = (int32_t*)take(CD * sizeof(int32_t));
    w.col_ind = (int32_t*)take(CD * sizeof(int32_t));
    w.cand = (Ap3Cand*)take((size_t)C * 2 * kAp3CandCap * sizeof(Ap3Cand));
    w.match_ws_bytes = ptk_workspace_bytes(PTK_WS_MATCH, C, D);
    w.match_ws = take(w.match_ws_bytes);
    w.bytes = o;
    return w;
}
}  // namespace

PTK_API size_t ptk_matchnd_sequence_workspace_bytes(int F, int C, int Nmax)
{
    if (F <= 0 || C <= 0 || Nmax <= 0 || Nmax > AP3_MAXN) return 0;
    return carve_nd(nullptr, F, C, Nmax).bytes + 256;
}

PTK_API int ptk_matchnd_sequence(const ptk_rig* rig, int F, int C, int Nmax, const double* cam1, const double* cam2,
                                 const int32_t* n1, const int32_t* n2, double* rows, double* costs, int32_t* n_rows,
                                 int32_t* status, int32_t* solver_stats, void* workspace, size_t workspace_bytes,
                                 void* stream)
{
    if (bad_rig(rig) || F < 0 || C < 0 || Nmax <= 0 || Nmax > AP3_MAXN || !cam1 || !cam2 || !n1 || !n2 || !rows || !costs ||
        !n_rows || !status)
        return PTK_ERR_INVALID_ARG;
    if (F == 0 || C == 0) return PTK_OK;
    if (!workspace || workspace_bytes < ptk_matchnd_sequence_workspace_bytes(F, C, Nmax)) return PTK_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int D = Nmax;
    NdWs w = carve_nd((void*)align_up((size_t)workspace), F, C, D);
    ptk_rig rmean = *rig, rsum = *rig;
    rmean.agg_sum = 0;  // frame 0: match2D.match_frame (matchND.py:322-340)
    rsum.agg_sum = 1;   // later frames: matchND.match_frame sums the two cameras' errors (matchND.py:525)
    const DevRig dsum = make_devrig(&rsum);
    const size_t fr4 = (size_t)C * D * 4, frR = (size_t)C * D * PTK_ROW_COLS, frD = (size_t)C * D;
    int rc = ptk_match_batch(&rmean, C, D, cam1, cam2, n1, n2, nullptr, nullptr, w.row_ind, w.col_ind, n_rows, status, rows,
                             costs, INFINITY, nullptr, w.match_ws, w.match_ws_bytes, stream);
    if (rc) return rc;
    CK(cudaMemsetAsync(w.wkey, 0, (size_t)F * C * sizeof(unsigned long long), st));
    if (solver_stats) CK(cudaMemsetAsync(solver_stats, 0, (size_t)C * 4 * sizeof(int32_t), st));
    for (int f = 1; f < F; f++) {
        const double* c1 = cam1 + f * fr4;
        const double* c2 = cam2 + f * fr4;
        const int32_t* m1 = n1 + (size_t)f * C;
        const int32_t* m2 = n2 + (size_t)f * C;
        rc = launch_undistort(dsum, C, D, c1, c2, m1, m2, w.und1, w.und2, st);
        if (rc) return rc;
        rc = launch_cost(dsum, C, D, c1, c2, w.und1, w.und2, w.rec2, m1, m2, w.cost, nullptr, w.pw, w.re, st);
        if (rc) return rc;
        NdArgs a;
        a.D = D;
        a.prev_rows = rows + (size_t)(f - 1) * frR;
        a.prev_n = n_rows + (size_t)(f - 1) * C;
        a.prev_status = status + (size_t)(f - 1) * C;
        a.n1 = m1; a.n2 = m2; a.cam1 = c1; a.cam2 = c2;
        a.p_world = w.pw; a.rep_errs = w.re; a.cost = w.cost; a.w = w.w; a.choices = w.choices;
        a.wkey = w.wkey + (size_t)f * C;
        a.sol = w.sol; a.n_sol = w.n_sol; a.ap3_status = w.ap3_status;
        a.rows = rows + (size_t)f * frR;
        a.costs = costs + (size_t)f * frD;
        a.n_rows = n_rows + (size_t)f * C;
        a.status = status + (size_t)f * C;
        const int gx = std::max(1, std::min(64, (D * D * D + 2047) / 2048));
        {
            LaunchTimer lt(K_WEIGHTS, st);
            k_nd_weights_raw<<<dim3(gx, C), 256, 0, st>>>(a);
        }
        LAUNCHED("k_nd_weights_raw");
        {
            LaunchTimer lt(K_WEIGHTS, st);
            k_nd_weights_finish<<<dim3(gx, C), 256, 0, st>>>(a);
        }
        LAUNCHED("k_nd_weights_finish");
        Ap3Args s;
        s.w = w.w; s.D0 = D; s.D1 = D; s.D2 = D;
        s.n0 = a.prev_n; s.n1 = m1; s.n2 = m2;
        s.sol = w.sol; s.n_sol = w.n_sol; s.value = w.value; s.bound = w.bound; s.status = w.ap3_status;
        s.stats = w.stats;
        s.cand = w.cand; s.cand_cap = kAp3CandCap; s.max_steps = kAp3Steps; s.max_sg = kAp3Subgradient; s.node_limit = kAp3Nodes;
        {
            LaunchTimer lt(K_AP3, st);
            k_ap3<<<C, AP3_THREADS, sizeof(Ap3Smem), st>>>(s);
        }
        LAUNCHED("k_ap3");
        k_nd_rows<<<C, 64, 0, st>>>(a);
        LAUNCHED("k_nd_rows");
        if (solver_stats) {
            k_nd_stats<<<(C + 63) / 64, 64, 0, st>>>(C, w.stats, solver_stats);
            LAUNCHED("k_nd_stats");
        }
    }
    return PTK_OK;
}

// ------------------------------------------------------------------------------------ K7
PTK_API int ptk_reorder_endpoints(int C, int F, int N, const double* ends, const double* pts2d, double* out_ends,
                                  double* out_2d, uint8_t* swapped, double* mincost, void* stream)
{
    if (C < 0 || F < 1 || N < 0 || !ends || !out_ends || !swapped || (pts2d && !out_2d) || (F > 1 && !mincost))
        return PTK_ERR_INVALID_ARG;
    const long long total = (long long)C * F * N;
    if (total == 0) return PTK_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (F > 1) {
        const long long t1 = (long long)C * (F - 1) * N;
        k_reorder_costs<<<(unsigned)((t1 + 255) / 256), 256, 0, st>>>(F, N, t1, ends, swapped, mincost);
        LAUNCHED("k_reorder_costs");
    }
    const int CN = C * N;
    k_reorder_scan<<<(CN + 127) / 128, 128, 0, st>>>(F, N, CN, swapped);
    LAUNCHED("k_reorder_scan");
    k_reorder_apply<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(total, swapped, ends, pts2d, out_ends, out_2d);
    LAUNCHED("k_reorder_apply");
    return PTK_OK;
}

// ------------------------------------------------------------------------------------ K8 / K9
PTK_API int ptk_triangulate_points(const ptk_rig* rig, long long n, const double* p1, const double* p2, int to_world,
                                   double* out, void* stream)
{
    if (bad_rig(rig) || n < 0 || !p1 || !p2 || !out) return PTK_ERR_INVALID_ARG;
    if (n == 0) return PTK_OK;
    k_triangulate_points<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(make_devrig(rig), n, to_world, p1, p2, out);
    LAUNCHED("k_triangulate_points");
    return PTK_OK;
}

PTK_API int ptk_project_points(const ptk_rig* rig, long long n, const double* X, int from_world, double* uv1, double* uv2,
                               void* stream)
{
    if (bad_rig(rig) || n < 0 || !X || !uv1 || !uv2) return PTK_ERR_INVALID_ARG;
    if (n == 0) return PTK_OK;
    k_project_points<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(make_devrig(rig), n, from_world, X, uv1, uv2);
    LAUNCHED("k_project_points");
    return PTK_OK;
}

// ------------------------------------------------------------------------------------ CSV writer (host)
PTK_API int ptk_csv_write_rods(const char* path, const char* header, long long n_rows, int n_float_cols,
                               const double* values, const int64_t* frame, const char* color, const int64_t* particle,
                               int n_seen)
{
    if (!path || !header || n_rows < 0 || n_float_cols < 0 || (n_rows && (!values || !frame || !particle)) || !color ||
        n_seen < 0)
        return PTK_ERR_INVALID_ARG;
    FILE* f = fopen(path, "wb");
    if (!f) return PTK_ERR_INVALID_ARG;
    const size_t clen = strlen(color);
    fwrite(header, 1, strlen(header), f);
    fputc('\n', f);
    // rows are formatted in parallel over contiguous ranges, then written in order
    const size_t worst = (size_t)n_float_cols * 26 + clen + 3 * 24 + (size_t)n_seen * 2 + 16;
    int T = (int)std::min<long long>(std::max(1u, std::thread::hardware_concurrency()), std::max<long long>(1, n_rows / 2048));
    T = std::min(T, 32);
    std::vector<std::vector<char>> bufs(T);
    std::vector<size_t> used(T, 0);
    std::atomic<int> oom{0};
    auto work = [&](int w) {
        const long long lo = n_rows * w / T, hi = n_rows * (w + 1) / T;
        try {
            bufs[w].resize((size_t)(hi - lo) * worst + 16);
        } catch (...) {
            oom = 1;
            return;
        }
        char* o = bufs[w].data();
        for (long long r = lo; r < hi; r++) {
            o += format_int(r, o);
            const double* v = values + (size_t)r * n_float_cols;
            for (int c = 0; c < n_float_cols; c++) {
                *o++ = ',';
                o += format_double_repr(v[c], o);
            }
            *o++ = ',';
            o += format_int(frame[r], o);
            *o++ = ',';
            memcpy(o, color, clen); o += clen;
            *o++ = ',';
            o += format_int(particle[r], o);
            for (int k = 0; k < n_seen; k++) { *o++ = ','; *o++ = '1'; }
            *o++ = '\n';
        }
        used[w] = (size_t)(o - bufs[w].data());
    };
    if (T == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int w = 0; w < T; w++) th.emplace_back(work, w);
        for (auto& x : th) x.join();
    }
    if (!oom)
        for (int w = 0; w < T; w++)
            if (used[w]) fwrite(bufs[w].data(), 1, used[w], f);
    const int bad = ferror(f) || oom;
    fclose(f);
    return bad ? PTK_ERR_INVALID_ARG : PTK_OK;
}

// test hook: Python repr() of one double (returns the length; buf needs 32 bytes)
PTK_API int ptk_format_double(double x, char* buf) { return format_double_repr(x, buf); }

// ------------------------------------------------------------------------------------ CSV reader (host)
struct ptk_csv {
    ptk::CsvTable t;
};

PTK_API int ptk_csv_open(const char* path, int n_threads, ptk_csv** out)
{
    if (!path || !out) return PTK_ERR_INVALID_ARG;
    ptk_csv* h = new (std::nothrow) ptk_csv();
    if (!h) return PTK_ERR_INVALID_ARG;
    int rc = -1;
    try {
        rc = ptk::csv_read_file(path, n_threads, &h->t);
    } catch (...) {
        h->t.error = "out of memory";
    }
    *out = h;  // kept on failure so that ptk_csv_error() can be read; the caller closes it
    return rc == 0 ? PTK_OK : PTK_ERR_INVALID_ARG;
}
PTK_API void ptk_csv_close(ptk_csv* h) { delete h; }
PTK_API const char* ptk_csv_error(const ptk_csv* h) { return h ? h->t.error.c_str() : "null handle"; }
PTK_API long long ptk_csv_n_rows(const ptk_csv* h) { return h ? h->t.n_rows : -1; }
PTK_API int ptk_csv_n_cols(const ptk_csv* h) { return h ? (int)h->t.cols.size() : -1; }
static const ptk::CsvColumn* csv_col(const ptk_csv* h, int c)
{
    return (h && c >= 0 && c < (int)h->t.cols.size()) ? &h->t.cols[c] : nullptr;
}
PTK_API const char* ptk_csv_col_name(const ptk_csv* h, int c)
{
    const ptk::CsvColumn* col = csv_col(h, c);
    return col ? col->name.c_str() : nullptr;
}
PTK_API int ptk_csv_col_kind(const ptk_csv* h, int c)
{
    const ptk::CsvColumn* col = csv_col(h, c);
    return col ? col->kind : PTK_ERR_INVALID_ARG;
}
PTK_API int ptk_csv_col_copy(const ptk_csv* h, int c, void* dst)
{
    const ptk::CsvColumn* col = csv_col(h, c);
    if (!col || (!dst && h->t.n_rows)) return PTK_ERR_INVALID_ARG;
    const size_t n = (size_t)h->t.n_rows;
    if (n == 0) return PTK_OK;
    switch (col->kind) {
    case ptk::CSV_INT64:
    case ptk::CSV_BOOL: memcpy(dst, col->i64.data(), n * sizeof(int64_t)); break;
    case ptk::CSV_FLOAT64: memcpy(dst, col->f64.data(), n * sizeof(double)); break;
    default: memcpy(dst, col->codes.data(), n * sizeof(int32_t)); break;
    }
    return PTK_OK;
}
PTK_API int ptk_csv_col_dict_size(const ptk_csv* h, int c)
{
    const ptk::CsvColumn* col = csv_col(h, c);
    return col ? (int)col->dict.size() : PTK_ERR_INVALID_ARG;
}
PTK_API const char* ptk_csv_col_dict(const ptk_csv* h, int c, int k, int* len)
{
    const ptk::CsvColumn* col = csv_col(h, c);
    if (!col || k < 0 || k >= (int)col->dict.size()) return nullptr;
    if (len) *len = (int)col->dict[k].size();
    return col->dict[k].data();
}
// test hook: pandas' float parser on one field (returns 1 if the field is a number)
PTK_API int ptk_parse_double(const char* s, int len, double* out)
{
    if (!s || len < 0 || !out) return PTK_ERR_INVALID_ARG;
    return ptk::csv_parse_double(s, s + len, out) ? 1 : 0;
}

// ------------------------------------------------------------------------------------ FP64 peak
PTK_API int ptk_measure_fp64_peak(int iters, double* flops, void* stream)
{
    if (iters <= 0 || !flops) return PTK_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double* out = nullptr;
    CK(cudaMalloc((void**)&out, sizeof(double)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int blocks = sms * 8, threads = 256;
    k_dfma_peak<<<blocks, threads, 0, st>>>(iters / 8 + 1, out);  // warm-up
    LAUNCHED("k_dfma_peak");
    CK(cudaEventRecord(e0, st));
    k_dfma_peak<<<blocks, threads, 0, st>>>(iters, out);
    LAUNCHED("k_dfma_peak");
    CK(cudaEventRecord(e1, st));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    *flops = 2.0 * 64.0 * (double)iters * blocks * threads / (ms * 1e-3);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return PTK_OK;
}

// ------------------------------------------------------------------------------------ host pipeline
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return PTK_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        CK(cudaMalloc(&p, bytes));
        cap = bytes;
        return PTK_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

struct ptk_ctx {
    int device = 0;
    cudaStream_t s_in = nullptr, s_comp = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_comp;
    cudaEvent_t ev_track = nullptr;
    DevBuf cam1, cam2, n1, n2, rows, straight, mc, row, col, nout, status, keep, ws;
    DevBuf ends, tcol, ttot, tid, tstatus, tws, tscratch;
};

PTK_API int ptk_ctx_create(int device, ptk_ctx** out)
{
    if (!out) return PTK_ERR_INVALID_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return PTK_ERR_NO_DEVICE; }
    if (device < 0 || device >= n) return PTK_ERR_INVALID_ARG;
    int major = 0;
    CK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (major != 10) return PTK_ERR_NO_DEVICE;
    CK(cudaSetDevice(device));
    ptk_ctx* c = new (std::nothrow) ptk_ctx();
    if (!c) return PTK_ERR_INVALID_ARG;
    c->device = device;
    CK(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->s_comp, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->ev_track, cudaEventDisableTiming));
    *out = c;
    return PTK_OK;
}

PTK_API void ptk_ctx_destroy(ptk_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    DevBuf* bufs[] = {&c->cam1, &c->cam2, &c->n1, &c->n2, &c->rows, &c->mc, &c->row, &c->col, &c->nout, &c->status,
                      &c->keep, &c->ws, &c->straight, &c->ends, &c->tcol, &c->ttot, &c->tid, &c->tstatus, &c->tws, &c->tscratch};
    for (DevBuf* b : bufs) b->release();
    for (cudaEvent_t e : c->ev_in) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_comp) cudaEventDestroy(e);
    if (c->ev_track) cudaEventDestroy(c->ev_track);
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_comp) cudaStreamDestroy(c->s_comp);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    delete c;
}

namespace {
// rows != NULL: the 18-column rows travel back (36.9 MB at 8000 x 32); otherwise ends6 / straight do
// (12.5 MB): the device never forms the other 12 columns, ptk_expand_rows_host can on the host.
int reconstruct_host_impl(ptk_ctx* c, const ptk_rig* rig, int F, int C, int Nmax, const double* cam1,
                          const double* cam2, const int32_t* n1, const int32_t* n2, double* rows, double* ends6,
                          uint8_t* straight, double* match_cost, int32_t* row_ind, int32_t* col_ind, int32_t* n_out,
                          int32_t* status, double max_repr_err, uint8_t* keep_mask, int do_track,
                          int32_t* track_col, double* track_total, int32_t* track_id, int32_t* track_status)
{
    const bool compact = rows == nullptr;
    if (!c || bad_rig(rig) || F < 0 || C < 0 || bad_n(Nmax) || !cam1 || !cam2 || !n1 || !n2 || !n_out || !status ||
        (compact && (!ends6 || !straight)))
        return PTK_ERR_INVALID_ARG;
    if (do_track && (!track_col || !track_total || F < 1)) return PTK_ERR_INVALID_ARG;
    const long long P = (long long)F * C;
    if (P == 0) return PTK_OK;
    CK(cudaSetDevice(c->device));
    NvtxRange nv("ptk_reconstruct_host");
    const DevRig d = make_devrig(rig);
    const size_t n4 = (size_t)Nmax * 4 * sizeof(double);
    const size_t n18 = (size_t)Nmax * (compact ? 6 : PTK_ROW_COLS) * sizeof(double);  // bytes of the big per-problem output
    const size_t nd = (size_t)Nmax * sizeof(double), ni = (size_t)Nmax * sizeof(int32_t);

    // frames per chunk: ~2048 problems, at most 32 full chunks.  PTK_HOST_RAMP=1 makes the first
    // chunks smaller (1/8, 1/4, 1/2 of a full one) so that the first host-to-device copy -- the one
    // transfer no kernel can hide -- is short.  Measured (profiles/README.md): slower (527 k vs 546 k
    // frames/s): the small chunks' launches fill the GPU badly and every extra chunk adds kernel
    // tails, which costs more than the 60 us of copy it hides.  Off by default.
    const char* ce = getenv("PTK_HOST_CHUNK");  // tuning knob (problems per chunk)
    const int want = ce ? atoi(ce) : 2048;
    int fchunk = ((want > 0 ? want : 2048) + C - 1) / C;
    if ((F + fchunk - 1) / fchunk > 32) fchunk = (F + 31) / 32;
    const long long pchunk = (long long)fchunk * C;
    int chunk_f0[40], nchunks = 0;
    {
        const char* re = getenv("PTK_HOST_RAMP");
        int f = 0, step = (re && atoi(re) == 1) ? (fchunk + 7) / 8 : fchunk;
        while (f < F) {
            chunk_f0[nchunks++] = f;
            f += step;
            step = step * 2 < fchunk ? step * 2 : fchunk;
        }
        chunk_f0[nchunks] = F;
    }

    int rc;
    if ((rc = c->cam1.ensure(P * n4))) return rc;
    if ((rc = c->cam2.ensure(P * n4))) return rc;
    if ((rc = c->n1.ensure(P * sizeof(int32_t)))) return rc;
    if ((rc = c->n2.ensure(P * sizeof(int32_t)))) return rc;
    if ((rc = c->rows.ensure(P * n18))) return rc;  // [P,Nmax,18] rows, or [P,Nmax,6] in compact mode
    if (compact && (rc = c->straight.ensure((size_t)P * Nmax))) return rc;
    if ((rc = c->mc.ensure(P * nd))) return rc;
    if ((rc = c->row.ensure(P * ni))) return rc;
    if ((rc = c->col.ensure(P * ni))) return rc;
    if ((rc = c->nout.ensure(P * sizeof(int32_t)))) return rc;
    if ((rc = c->status.ensure(P * sizeof(int32_t)))) return rc;
    if (keep_mask && (rc = c->keep.ensure(P * Nmax))) return rc;
    const size_t wsb = ptk_workspace_bytes(PTK_WS_MATCH, pchunk, Nmax);
    if ((rc = c->ws.ensure(wsb))) return rc;
    while ((int)c->ev_in.size() < nchunks) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->ev_in.push_back(e);
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->ev_comp.push_back(e);
    }
    const int N = Nmax;
    const long long Q = (long long)C * (F - 1);
    if (do_track) {
        if ((rc = c->ends.ensure((size_t)C * F * N * 6 * sizeof(double)))) return rc;
        if ((rc = c->tcol.ensure((size_t)(Q > 0 ? Q : 1) * N * sizeof(int32_t)))) return rc;
        if ((rc = c->ttot.ensure((size_t)(Q > 0 ? Q : 1) * sizeof(double)))) return rc;
        if ((rc = c->tstatus.ensure((size_t)(Q > 0 ? Q : 1) * sizeof(int32_t)))) return rc;
        if ((rc = c->tws.ensure(ptk_workspace_bytes(PTK_WS_TRACK, Q > 0 ? Q : 1, N)))) return rc;
        if (track_id) {
            if ((rc = c->tid.ensure((size_t)C * F * N * sizeof(int32_t)))) return rc;
            if ((rc = c->tscratch.ensure(chain_scratch_ints(C, F, N) * sizeof(int32_t) + 16))) return rc;
        }
    }

    // The host issues ~10 API calls per chunk and a small chunk's kernels are shorter than that, so
    // only the two large streams are chunked: detections in, rows out.  The rod counts go in once,
    // up front (4 bytes per problem); the small per-problem results come back once, after the last
    // chunk, under the tracking kernels.
    CK(cudaMemcpyAsync(c->n1.as<int32_t>(), n1, P * sizeof(int32_t), cudaMemcpyHostToDevice, c->s_in));
    CK(cudaMemcpyAsync(c->n2.as<int32_t>(), n2, P * sizeof(int32_t), cudaMemcpyHostToDevice, c->s_in));
    TrkTarget trk{};
    for (int k = 0; k < nchunks; k++) {
        NvtxRange nvc("chunk: H2D + match + D2H");
        const int f0 = chunk_f0[k];
        const int fn = chunk_f0[k + 1] - f0;
        const long long p0 = (long long)f0 * C, np = (long long)fn * C;
        // H2D
        CK(cudaMemcpyAsync(c->cam1.as<char>() + p0 * n4, (const char*)cam1 + p0 * n4, np * n4, cudaMemcpyHostToDevice, c->s_in));
        CK(cudaMemcpyAsync(c->cam2.as<char>() + p0 * n4, (const char*)cam2 + p0 * n4, np * n4, cudaMemcpyHostToDevice, c->s_in));
        CK(cudaEventRecord(c->ev_in[k], c->s_in));
        // compute
        CK(cudaStreamWaitEvent(c->s_comp, c->ev_in[k], 0));
        const int rcols = compact ? 6 : PTK_ROW_COLS;
        double* drows = c->rows.as<double>() + p0 * Nmax * rcols;
        rc = match_batch_impl(rig, (int)np, Nmax, c->cam1.as<double>() + p0 * Nmax * 4, c->cam2.as<double>() + p0 * Nmax * 4,
                              c->n1.as<int32_t>() + p0, c->n2.as<int32_t>() + p0, nullptr, nullptr,
                              c->row.as<int32_t>() + p0 * Nmax, c->col.as<int32_t>() + p0 * Nmax,
                              c->nout.as<int32_t>() + p0, c->status.as<int32_t>() + p0, compact ? nullptr : drows,
                              compact ? drows : nullptr, compact ? c->straight.as<uint8_t>() + p0 * Nmax : nullptr,
                              c->mc.as<double>() + p0 * Nmax, max_repr_err,
                              keep_mask ? c->keep.as<uint8_t>() + p0 * Nmax : nullptr, c->ws.p, c->ws.cap, c->s_comp,
                              do_track ? &(trk = TrkTarget{c->ends.as<double>(), F, C, N, p0}) : nullptr);
        if (rc) return rc;
        CK(cudaEventRecord(c->ev_comp[k], c->s_comp));
        // D2H
        CK(cudaStreamWaitEvent(c->s_out, c->ev_comp[k], 0));
        CK(cudaMemcpyAsync((char*)(compact ? ends6 : rows) + p0 * n18, c->rows.as<char>() + p0 * n18, np * n18,
                           cudaMemcpyDeviceToHost, c->s_out));
    }
    if (compact) CK(cudaMemcpyAsync(straight, c->straight.p, (size_t)P * Nmax, cudaMemcpyDeviceToHost, c->s_out));
    // s_out has waited for the last chunk's kernels: everything of the matching stage is final
    if (match_cost) CK(cudaMemcpyAsync(match_cost, c->mc.p, P * nd, cudaMemcpyDeviceToHost, c->s_out));
    if (row_ind) CK(cudaMemcpyAsync(row_ind, c->row.p, P * ni, cudaMemcpyDeviceToHost, c->s_out));
    if (col_ind) CK(cudaMemcpyAsync(col_ind, c->col.p, P * ni, cudaMemcpyDeviceToHost, c->s_out));
    CK(cudaMemcpyAsync(n_out, c->nout.p, P * sizeof(int32_t), cudaMemcpyDeviceToHost, c->s_out));
    CK(cudaMemcpyAsync(status, c->status.p, P * sizeof(int32_t), cudaMemcpyDeviceToHost, c->s_out));
    if (keep_mask) CK(cudaMemcpyAsync(keep_mask, c->keep.p, (size_t)P * Nmax, cudaMemcpyDeviceToHost, c->s_out));
    if (do_track && Q > 0) {
        NvtxRange nvt("track + chain + D2H");
        rc = ptk_track_batch(C, F, N, c->ends.as<double>(), nullptr, c->tcol.as<int32_t>(), c->ttot.as<double>(),
                             c->tstatus.as<int32_t>(), c->tws.p, c->tws.cap, c->s_comp);
        if (rc) return rc;
        if (track_id) {
            rc = chain_impl(C, F, N, c->tcol.as<int32_t>(), nullptr, F, c->tid.as<int32_t>(), nullptr, c->tscratch.as<int32_t>(), c->s_comp);
            if (rc) return rc;
        }
        CK(cudaEventRecord(c->ev_track, c->s_comp));
        CK(cudaStreamWaitEvent(c->s_out, c->ev_track, 0));
        CK(cudaMemcpyAsync(track_col, c->tcol.p, (size_t)Q * N * sizeof(int32_t), cudaMemcpyDeviceToHost, c->s_out));
        CK(cudaMemcpyAsync(track_total, c->ttot.p, (size_t)Q * sizeof(double), cudaMemcpyDeviceToHost, c->s_out));
        if (track_status) CK(cudaMemcpyAsync(track_status, c->tstatus.p, (size_t)Q * sizeof(int32_t), cudaMemcpyDeviceToHost, c->s_out));
        if (track_id) CK(cudaMemcpyAsync(track_id, c->tid.p, (size_t)C * F * N * sizeof(int32_t), cudaMemcpyDeviceToHost, c->s_out));
    } else if (do_track && track_id && F == 1) {
        for (int cc = 0; cc < C; cc++)
            for (int a = 0; a < N; a++) track_id[(size_t)cc * N + a] = a;
    }
    CK(cudaStreamSynchronize(c->s_out));
    CK(cudaStreamSynchronize(c->s_comp));
    for (long long p = 0; p < P; p++)
        if (status[p] != PTK_PROBLEM_OK && status[p] != PTK_PROBLEM_EMPTY) return PTK_ERR_NUMERIC;
    if (do_track && track_status)
        for (long long q = 0; q < Q; q++)
            if (track_status[q] != PTK_PROBLEM_OK) return PTK_ERR_NUMERIC;
    return PTK_OK;
}
}  // namespace

PTK_API int ptk_reconstruct_host(ptk_ctx* c, const ptk_rig* rig, int F, int C, int Nmax, const double* cam1,
                                 const double* cam2, const int32_t* n1, const int32_t* n2, double* rows,
                                 double* match_cost, int32_t* row_ind, int32_t* col_ind, int32_t* n_out,
                                 int32_t* status, double max_repr_err, uint8_t* keep_mask, int do_track,
                                 int32_t* track_col, double* track_total, int32_t* track_id, int32_t* track_status)
{
    if (!rows) return PTK_ERR_INVALID_ARG;
    return reconstruct_host_impl(c, rig, F, C, Nmax, cam1, cam2, n1, n2, rows, nullptr, nullptr, match_cost, row_ind, col_ind,
                                 n_out, status, max_repr_err, keep_mask, do_track, track_col, track_total, track_id,
                                 track_status);
}

PTK_API int ptk_reconstruct_host_compact(ptk_ctx* c, const ptk_rig* rig, int F, int C, int Nmax, const double* cam1,
                                         const double* cam2, const int32_t* n1, const int32_t* n2, double* ends6,
                                         uint8_t* straight, double* match_cost, int32_t* row_ind, int32_t* col_ind,
                                         int32_t* n_out, int32_t* status, double max_repr_err, uint8_t* keep_mask,
                                         int do_track, int32_t* track_col, double* track_total, int32_t* track_id,
                                         int32_t* track_status)
{
    if (!ends6 || !straight || !col_ind) return PTK_ERR_INVALID_ARG;
    return reconstruct_host_impl(c, rig, F, C, Nmax, cam1, cam2, n1, n2, nullptr, ends6, straight, match_cost, row_ind,
                                 col_ind, n_out, status, max_repr_err, keep_mask, do_track, track_col, track_total, track_id,
                                 track_status);
}

// The 18 columns from the compact result (match2D.py:963-983): x1..z2 as computed, centre = (end1 + end2) / 2,
// length = |end2 - end1| -- the same IEEE operations in the same order as K3, hence the same bits -- and the two
// cameras' 2D end points of the matched rods (m, col_ind[m]) from the CALLER'S OWN input arrays, reversed where
// the crossed pairing won.  Rows beyond n_out are NaN.  Host code, n_threads workers over the problems.
PTK_API int ptk_expand_rows_host(long long P, int Nmax, const double* cam1, const double* cam2, const double* ends6,
                                 const uint8_t* straight, const int32_t* col_ind, const int32_t* n_out, double* rows,
                                 int n_threads)
{
    if (P < 0 || bad_n(Nmax) || (P && (!cam1 || !cam2 || !ends6 || !straight || !col_ind || !n_out || !rows)))
        return PTK_ERR_INVALID_ARG;
    if (P == 0) return PTK_OK;
    int T = n_threads > 0 ? n_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    T = (int)std::min<long long>(std::min(T, 64), std::max<long long>(1, P / 64));
    std::atomic<int> bad{0};
    auto work = [&](int w) {
        const long long lo = P * w / T, hi = P * (w + 1) / T;
        const double qnan = std::numeric_limits<double>::quiet_NaN();
        for (long long p = lo; p < hi; p++) {
            const int n = n_out[p];
            if (n < 0 || n > Nmax) { bad = 1; continue; }
            const double* c1 = cam1 + (size_t)p * Nmax * 4;
            const double* c2 = cam2 + (size_t)p * Nmax * 4;
            for (int m = 0; m < Nmax; m++) {
                double* o = rows + ((size_t)p * Nmax + m) * PTK_ROW_COLS;
                if (m >= n) {
                    for (int k = 0; k < PTK_ROW_COLS; k++) o[k] = qnan;
                    continue;
                }
                const double* e = ends6 + ((size_t)p * Nmax + m) * 6;
                const int j = col_ind[(size_t)p * Nmax + m];
                if (j < 0 || j >= Nmax) { bad = 1; continue; }
                const double ax = e[0], ay = e[1], az = e[2], bx = e[3], by = e[4], bz = e[5];
                o[0] = ax; o[1] = ay; o[2] = az; o[3] = bx; o[4] = by; o[5] = bz;
                o[6] = (ax + bx) / 2; o[7] = (ay + by) / 2; o[8] = (az + bz) / 2;
                const double dx = bx - ax, dy = by - ay, dz = bz - az;
                volatile double sq = dx * dx;  // keep the three products and two sums separately rounded (no contraction)
                volatile double t2 = dy * dy;
                volatile double t3 = dz * dz;
                volatile double s1 = sq + t2;
                o[9] = sqrt(s1 + t3);
                const double* r1 = c1 + (size_t)m * 4;
                const double* r2 = c2 + (size_t)j * 4;
                if (straight[(size_t)p * Nmax + m]) {
                    o[10] = r1[0]; o[11] = r1[1]; o[12] = r1[2]; o[13] = r1[3];
                    o[14] = r2[0]; o[15] = r2[1]; o[16] = r2[2]; o[17] = r2[3];
                } else {
                    o[10] = r1[2]; o[11] = r1[3]; o[12] = r1[0]; o[13] = r1[1];
                    o[14] = r2[2]; o[15] = r2[3]; o[16] = r2[0]; o[17] = r2[1];
                }
            }
        }
    };
    if (T == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int w = 0; w < T; w++) th.emplace_back(work, w);
        for (auto& x : th) x.join();
    }
    return bad ? PTK_ERR_INVALID_ARG : PTK_OK;
}
