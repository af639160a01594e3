// ptk_bounds.cuh -- the bounds-checked debug build (make debug: -DPTK_BOUNDS_CHECK -> libptk_dbg.so).
//
// compute-sanitizer is closed on this GPU pool, so memory safety cannot be checked from outside.  In the
// debug build every index that is derived from data -- the rod counts n1 / n2, an assignment read back
// from memory (col_ind, col4row), a position decoded from a REDUX result -- goes through PTK_AT / PTK_OK
// before it is used: a violation is COUNTED (device-global log: count, first site, index, bound) and the
// access is redirected to a sink / skipped, never trapped -- a trap would poison the context, and a
// faulting process under this driver has left boxes unusable.  tests/conftest.py fails the session if
// the log is not empty at the end (PTK_LIB=.../libptk_dbg.so python -m pytest tests -m gpu).
// In the release build both macros vanish.
#pragma once
#include <cuda_runtime.h>

namespace ptk {

struct BoundsLog {
    unsigned long long count;
    long long first_idx, first_bound;
    int first_site, enabled;
};

#ifdef PTK_BOUNDS_CHECK
__device__ BoundsLog g_bounds_log = {0ull, 0ll, 0ll, 0, 1};
__device__ double g_bounds_sink[4];

__device__ __forceinline__ bool bounds_ok(long long i, long long n, int site)
{
    if (i >= 0 && i < n) return true;
    if (atomicAdd(&g_bounds_log.count, 1ull) == 0ull) {
        g_bounds_log.first_idx = i;
        g_bounds_log.first_bound = n;
        g_bounds_log.first_site = site;
    }
    return false;
}
// element i of p (valid range [0, n)), or the sink
template <typename T>
__device__ __forceinline__ T& bounds_at(T* p, long long i, long long n, int site)
{
    return bounds_ok(i, n, site) ? p[i] : *reinterpret_cast<T*>(const_cast<double*>(g_bounds_sink));
}
#define PTK_OK_IDX(i, n, site) (::ptk::bounds_ok((long long)(i), (long long)(n), (site)))
#define PTK_AT(p, i, n, site) (::ptk::bounds_at((p), (long long)(i), (long long)(n), (site)))
#else
#define PTK_OK_IDX(i, n, site) (true)
#define PTK_AT(p, i, n, site) ((p)[(i)])
#endif

}  // namespace ptk
