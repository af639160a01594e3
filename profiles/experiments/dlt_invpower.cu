// Host harness for dlt_null_invpower (csrc/ptk_device.cuh): the function is __host__ __device__, so
// its arithmetic can be checked against cv2.triangulatePoints on the CPU (dlt_invpower.py) before
// any GPU time is spent.  nvcc -O2 -fmad=false -shared -Xcompiler -fPIC -o dlt_invpower.so dlt_invpower.cu
#include "../../particletracking_b200/csrc/ptk_device.cuh"

extern "C" void tri_invpower_host(const double* P1, const double* P2, long long M, const double* pts /*[M,4]: x1 y1 x2 y2*/,
                                  double* X /*[M,3]*/, int* converged /*[M]*/)
{
    for (long long m = 0; m < M; m++) {
        double v[4];
        converged[m] = ptk::dlt_null_invpower(P1, P2, pts[4 * m], pts[4 * m + 1], pts[4 * m + 2], pts[4 * m + 3], v) ? 1 : 0;
        X[3 * m] = v[0] / v[3];
        X[3 * m + 1] = v[1] / v[3];
        X[3 * m + 2] = v[2] / v[3];
    }
}
