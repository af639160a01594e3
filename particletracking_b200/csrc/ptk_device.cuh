// ptk_device.cuh -- FP64 device math of the stereo hot path (sm_100a).
//
// Everything here is plain CUDA-core FP64: the path is ~1300 flop/B, has no
// dense contraction, so neither tensor cores nor TMA have anything to do; the design rules
// that matter are (i) keep the 4x4 state in registers, (ii) count instructions of every kind --
// FP64 and integer-ALU instructions hold a scheduler's dispatch port for two cycles, and that,
// not latency, bounds K1 --, (iii) keep div/sqrt sequences and rare cases off the main path.
//
// The translation unit is compiled with -fmad=false: the compiler never contracts on its
// own, every fused multiply-add below is an explicit fma().  That makes the arithmetic a
// pure function of the inputs -- K1 (cost matrix) and K3 (row gather) re-evaluate the same
// rod pair and must get the same bits -- and lets K0/K4 follow OpenCV's / numpy's unfused
// operation order exactly where bit-identity with the CPU path is wanted.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/ptk.h"
#include "ptk_bounds.cuh"

namespace ptk {

// Device-side copy of the rig, pre-digested (reciprocal focal lengths etc.).
struct DevRig {
    double P1[12], P2[12];    // columns permuted by `perm` (host-side static de Rijk ordering, see make_devrig)
    double P1o[12], P2o[12];  // as passed by the caller (original column order)
    int perm[4];              // column k of the permuted matrices is original column perm[k]
    double R1[9], t1[3], R2[9], t2[3];
    double CM1[4], CM2[4];  // fx, fy, cx, cy
    double k1[12], k2[12];  // k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4
    double Rw[9], tw[3];
    int agg_sum;
    double agg_scale;     // 1 (agg_sum: matchND.py:525 np.sum) or 0.5 (match2D.py:910 np.mean; x * 0.5 == x / 2.0 exactly)
    // rig-constant shortcuts for the projection (set by make_devrig; warp-uniform branches):
    int ext_identity[2];  // camera k has R = I, t = 0 (camera 1 of a stereo rig)
    int tangential[2];    // camera k has any of p1, p2, s1..s4 non-zero
    // rig class, selects the kernel instantiation (make_devrig):
    int canon;            // P1 == CM1 [I|0] exactly, zero skew, and camera 1 has identity extrinsics
    int simple;           // both cameras: radial k1, k2, k3 only (no rational, tangential, thin-prism terms)
    double fx2;           // CM1 fx * fx (canon)
};

// ---------------------------------------------------------------------------------------
// cv2.undistortImagePoints for one point (reference call sites match2D.py:847,856; SURVEY
// App. B.1): 5 fixed-point iterations, then re-projection with P = CM.  Unfused arithmetic
// in OpenCV's operation order: bit-identical to cv2 4.13 on the CPU (tests/test_gpu_*).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void undistort_point(const double* __restrict__ cm, const double* __restrict__ k,
                                                double u, double v, double& ou, double& ov)
{
    const double fx = cm[0], fy = cm[1], cx = cm[2], cy = cm[3];
    const double ifx = 1.0 / fx, ify = 1.0 / fy;
    double x = (u - cx) * ifx, y = (v - cy) * ify;
    const double x0 = x, y0 = y;
#pragma unroll 1
    for (int j = 0; j < 5; j++) {
        double r2 = x * x + y * y;
        double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
        if (icdist < 0) {
            x = (u - cx) * ifx;
            y = (v - cy) * ify;
            break;
        }
        double dx = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
        double dy = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
        x = (x0 - dx) * icdist;
        y = (y0 - dy) * icdist;
    }
    double xx = fx * x + 0.0 * y + cx;
    double yy = 0.0 * x + fy * y + cy;
    double ww = 1.0 / (0.0 * x + 0.0 * y + 1.0);
    ou = xx * ww;
    ov = yy * ww;
}

// ---------------------------------------------------------------------------------------
// One Jacobi rotation of rows i, j of At (columns of A) and of Vt, as OpenCV's
// JacobiSVDImpl_<double> does it, restructured for the FP64 pipe:
//  - the skip test |p| <= eps*sqrt(a*b) is evaluated squared (no sqrt);
//  - (c, s) come from two reciprocal square roots instead of hypot + 2 sqrt + 2 div:
//        g = 1/hypot(2p, beta),  t = (1 + |beta| g)/2,  big = sqrt(t) = t*rsqrt(t),
//        small = p g rsqrt(t)        (p here is the un-doubled dot product)
//    which is the same rotation to a few ulp; Jacobi is self-correcting, the column norms
//    are recomputed from the rotated data exactly as OpenCV does.
// ---------------------------------------------------------------------------------------
#define PTK_JACOBI_EPS2 (4.930380657631324e-30) /* (10*DBL_EPSILON)^2 */

// rsqrt in two pieces: the hardware seed (MUFU.RSQ64H, ~2^-22 relative) and one third-order
// refinement y0 (1 + e/2 + 3e^2/8), e = 1 - x y0^2, which lands within an ulp or two.  Splitting
// lets the second seed of a rotation be issued from an *approximate* argument while the first
// result is still being refined (the seed only needs 7 digits), shortening the dependent chain;
// it also drops the library version's special-case branch (arguments here are finite, positive).
__device__ __forceinline__ double rsqrt_seed(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}
// 1/x to within an ulp or two: hardware seed (MUFU.RCP64H, ~2^-22) and one third-order step
// y0 (1 + e + e^2), e = 1 - x y0.  Four instructions instead of the ~12 of an IEEE division;
// used where 1e-16 does not matter (de-homogenisation, projection), never in K0 / K4 whose
// results are bit-compared with OpenCV / numpy.
__device__ __forceinline__ double fast_rcp(double x)
{
    double y0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    const double e = fma(-x, y0, 1.0);
    return fma(y0, fma(e, e, e), y0);
}
__device__ __forceinline__ double rsqrt_refine(double x, double y0)
{
    const double h = x * y0;
    const double e = fma(-h, y0, 1.0);
    const double q = fma(0.375, e, 0.5);
    const double ye = y0 * e;
    return fma(q, ye, y0);
}

__device__ __forceinline__ bool jacobi_pair(double (&Ai)[4], double (&Aj)[4], double& Wi, double& Wj)
{
    double p = Ai[0] * Aj[0];
    p = fma(Ai[1], Aj[1], p);
    p = fma(Ai[2], Aj[2], p);
    p = fma(Ai[3], Aj[3], p);
    const double a = Wi, b = Wj;
    if (!(p * p > PTK_JACOBI_EPS2 * (a * b))) return false;  // also false for p == 0
    const double beta = a - b;
    const double p2 = p + p;
#ifndef PTK_LIBRARY_RSQRT
    // 3 % faster than two rsqrt() calls (tools/ab_bench.py); results differ by ~3e-15
    const double x1 = fma(p2, p2, beta * beta);
    const double hb = 0.5 * fabs(beta);
    const double g0 = rsqrt_seed(x1);
    const double r0 = rsqrt_seed(fma(hb, g0, 0.5));  // seed for rsqrt(t) from the unrefined g
    const double g = rsqrt_refine(x1, g0);
    const double t = fma(hb, g, 0.5);
    const double rt = rsqrt_refine(t, r0);
#else
    const double g = rsqrt(fma(p2, p2, beta * beta));
    const double t = fma(fabs(beta), 0.5 * g, 0.5);
    const double rt = rsqrt(t);
#endif
    const double big = t * rt;
    const double small_ = p * g * rt;
    const double c = beta < 0 ? small_ : big;
    const double s = beta < 0 ? big : small_;
    double na = 0, nb = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const double x = Ai[k], y = Aj[k];
        const double t0 = fma(c, x, s * y);
        const double t1 = fma(c, y, -(s * x));
        Ai[k] = t0;
        Aj[k] = t1;
        na = fma(t0, t0, na);
        nb = fma(t1, t1, nb);
    }
    Wi = na;
    Wj = nb;
    return true;
}

// The same rotation without the skip test, for the first two sweeps: on DLT matrices every pair
// needs its rotation there anyway (measured: identical rotation counts with and without the test,
// profiles/experiments), and straight-line code lets the compiler keep the 36 doubles of state in
// place instead of shuffling registers at every if-merge (16 moves per rotation) -- this kernel
// is co-limited by instruction issue, not only by the FP64 pipe.  A rotation of an already
// orthogonal pair is the identity (s = 0); the only singular case, p == 0 with equal norms, is
// mapped to the identity explicitly.
__device__ __forceinline__ void jacobi_pair_always(double (&Ai)[4], double (&Aj)[4], double& Wi, double& Wj)
{
    double p = Ai[0] * Aj[0];
    p = fma(Ai[1], Aj[1], p);
    p = fma(Ai[2], Aj[2], p);
    p = fma(Ai[3], Aj[3], p);
    const double beta = Wi - Wj;
    const double p2 = p + p;
    const double x1 = fma(p2, p2, beta * beta);
    const double hb = 0.5 * fabs(beta);
    const double g0 = rsqrt_seed(x1);
    const double r0 = rsqrt_seed(fma(hb, g0, 0.5));
    const double g = rsqrt_refine(x1, g0);
    const double t = fma(hb, g, 0.5);
    const double rt = rsqrt_refine(t, r0);
    const double big = t * rt;
    const double small_ = p * g * rt;
    const bool ok = x1 > 0.0;
    const double c = ok ? (beta < 0 ? small_ : big) : 1.0;
    const double s = ok ? (beta < 0 ? big : small_) : 0.0;
    double na = 0, nb = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const double x = Ai[k], y = Aj[k];
        const double t0 = fma(c, x, s * y);
        const double t1 = fma(c, y, -(s * x));
        Ai[k] = t0;
        Aj[k] = t1;
        na = fma(t0, t0, na);
        nb = fma(t1, t1, nb);
    }
    Wi = na;
    Wj = nb;
}

// cv2.triangulatePoints for one point (match2D.py:889) + the reference's p[0:3]/p[3]
// (match2D.py:895).  Returns the point in camera-1 coordinates.
//
// One-sided Jacobi on the columns of the 4x4 DLT matrix A, but WITHOUT accumulating V: after
// convergence the columns b_k = A v_k = sigma_k u_k are orthogonal, so w_k = A^T b_k = sigma_k^2 v_k
// gives the right singular vectors of the three LARGE singular values accurately, and the wanted
// vector of the smallest one is the direction orthogonal to all three -- their generalised
// cross product.  Against accumulating V (what OpenCV does) this is 88 FP64 instructions once
// instead of 16 per rotation (~220), and 32 fewer live registers; measured agreement with the
// V-accumulating Jacobi on 327 680 combinations incl. dummy rods: max 1.0e-15 relative on the
// triangulated point (profiles/experiments/jacobi_no_v.c).
__device__ __forceinline__ void dlt_triangulate_jacobi(const DevRig& rig, double x1, double y1, double x2, double y2,
                                                       double& X, double& Y, double& Z, int* n_rot = nullptr)
{
    const double* __restrict__ P1 = rig.P1;
    const double* __restrict__ P2 = rig.P2;
    double A0[4], A1[4], A2[4], A3[4];  // rows of At = columns of the (column-permuted) DLT matrix
    // At[k][r]: r = 0: x1*P1[2,k]-P1[0,k]; 1: y1*P1[2,k]-P1[1,k]; 2,3: same for view 2
#define PTK_COL(Ak, k)                       \
    Ak[0] = fma(x1, P1[8 + k], -P1[0 + k]);  \
    Ak[1] = fma(y1, P1[8 + k], -P1[4 + k]);  \
    Ak[2] = fma(x2, P2[8 + k], -P2[0 + k]);  \
    Ak[3] = fma(y2, P2[8 + k], -P2[4 + k]);
    PTK_COL(A0, 0) PTK_COL(A1, 1) PTK_COL(A2, 2) PTK_COL(A3, 3)
#undef PTK_COL
    double W0, W1, W2, W3;
#define PTK_NRM(Ak) fma(Ak[3], Ak[3], fma(Ak[2], Ak[2], fma(Ak[1], Ak[1], Ak[0] * Ak[0])))
    W0 = PTK_NRM(A0); W1 = PTK_NRM(A1); W2 = PTK_NRM(A2); W3 = PTK_NRM(A3);
#undef PTK_NRM
    // de Rijk ordering: the columns are taken by decreasing norm before the first sweep.  OpenCV
    // does not, but the singular vectors do not depend on the column order, and on DLT matrices
    // (one column ~ f*T dominates) it cuts the rotations from 19.6 to 13.9 and the sweeps from 4.8
    // to 3.9 per point at the same 1e-15 accuracy (profiles/experiments/jacobi_variants.c).
    // The order is fixed by the rig, so the host already permuted the columns of P1/P2
    // (make_devrig); here it is only verified, and the rare point that disagrees (e.g. a dummy
    // detection far off the image) goes through a branch-free 5-comparator network.
#define PTK_CSWAP(a, b)                                                      \
    {                                                                        \
        const bool sw = W##b > W##a;                                         \
        _Pragma("unroll") for (int k = 0; k < 4; k++) {                      \
            const double lo = sw ? A##a[k] : A##b[k];                        \
            A##a[k] = sw ? A##b[k] : A##a[k];                                \
            A##b[k] = lo;                                                    \
        }                                                                    \
        const double wl = sw ? W##a : W##b;                                  \
        W##a = sw ? W##b : W##a;                                             \
        W##b = wl;                                                           \
    }
    if (!(W0 >= W1 && W1 >= W2 && W2 >= W3)) {
        PTK_CSWAP(0, 1) PTK_CSWAP(2, 3) PTK_CSWAP(0, 2) PTK_CSWAP(1, 3) PTK_CSWAP(1, 2)
    }
    // sweeps 1 and 2: every pair, no skip test (see jacobi_pair_always)
#pragma unroll 1
    for (int iter = 0; iter < 2; iter++) {
        jacobi_pair_always(A0, A1, W0, W1);
        jacobi_pair_always(A0, A2, W0, W2);
        jacobi_pair_always(A0, A3, W0, W3);
        jacobi_pair_always(A1, A2, W1, W2);
        jacobi_pair_always(A1, A3, W1, W3);
        jacobi_pair_always(A2, A3, W2, W3);
    }
    int rot = 12;
#pragma unroll 1
    for (int iter = 2; iter < 30; iter++) {
        bool ch = false;
        bool r;
        r = jacobi_pair(A0, A1, W0, W1); ch |= r; rot += r;
        r = jacobi_pair(A0, A2, W0, W2); ch |= r; rot += r;
        r = jacobi_pair(A0, A3, W0, W3); ch |= r; rot += r;
        r = jacobi_pair(A1, A2, W1, W2); ch |= r; rot += r;
        r = jacobi_pair(A1, A3, W1, W3); ch |= r; rot += r;
        r = jacobi_pair(A2, A3, W2, W3); ch |= r; rot += r;
        if (!ch) break;
    }
    if (n_rot) *n_rot = rot;
    // W holds the squared column norms of the final A V (recomputed by every rotation, exactly
    // what OpenCV's closing pass computes before its sqrt).  The column of the smallest one is
    // moved to the last position (it is there already unless the point is unusual; ties -> the
    // later column, as OpenCV's descending selection sort leaves it).
    if (!(W3 <= W0 && W3 <= W1 && W3 <= W2)) {
#define PTK_MINLAST(a) { const bool sw = W##a < W3; _Pragma("unroll") for (int k = 0; k < 4; k++) { \
            const double lo = sw ? A##a[k] : A3[k]; A##a[k] = sw ? A3[k] : A##a[k]; A3[k] = lo; }      \
        const double wl = sw ? W##a : W3; W##a = sw ? W3 : W##a; W3 = wl; }
        PTK_MINLAST(0) PTK_MINLAST(1) PTK_MINLAST(2)
#undef PTK_MINLAST
    }
#undef PTK_CSWAP
    // w_j = A^T b_j for the three large columns, with A in the ORIGINAL column order (so the
    // result needs no un-permuting): A[r][c] = coord_r * P[2][c] - P[row_r][c]
    const double* __restrict__ Q1 = rig.P1o;
    const double* __restrict__ Q2 = rig.P2o;
    double w0[4], w1[4], w2[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const double r0 = fma(x1, Q1[8 + c], -Q1[c]);
        const double r1 = fma(y1, Q1[8 + c], -Q1[4 + c]);
        const double r2 = fma(x2, Q2[8 + c], -Q2[c]);
        const double r3 = fma(y2, Q2[8 + c], -Q2[4 + c]);
        w0[c] = fma(r3, A0[3], fma(r2, A0[2], fma(r1, A0[1], r0 * A0[0])));
        w1[c] = fma(r3, A1[3], fma(r2, A1[2], fma(r1, A1[1], r0 * A1[0])));
        w2[c] = fma(r3, A2[3], fma(r2, A2[2], fma(r1, A2[1], r0 * A2[0])));
    }
    // generalised cross product of w0, w1, w2 (the 4-vector orthogonal to all three)
#define PTK_M2(a, b) fma(w1[a], w2[b], -(w1[b] * w2[a]))
    const double m01 = PTK_M2(0, 1), m02 = PTK_M2(0, 2), m03 = PTK_M2(0, 3);
    const double m12 = PTK_M2(1, 2), m13 = PTK_M2(1, 3), m23 = PTK_M2(2, 3);
#undef PTK_M2
    const double vx = fma(w0[3], m12, fma(-w0[2], m13, w0[1] * m23));
    const double vy = -fma(w0[3], m02, fma(-w0[2], m03, w0[0] * m23));
    const double vz = fma(w0[3], m01, fma(-w0[1], m03, w0[0] * m13));
    const double vw = -fma(w0[2], m01, fma(-w0[1], m02, w0[0] * m12));
    const double rw = fast_rcp(vw);  // match2D.py:895: p[0:3] / p[3]
    X = vx * rw;
    Y = vy * rw;
    Z = vz * rw;
}

// ---------------------------------------------------------------------------------------
// The same null vector without any SVD iteration: QR + inverse power by repeated squaring.
//
//   A = Q R (modified Gram-Schmidt on the 4 columns),  S = R^-1 (upper triangular),
//   B = S S^T = (A^T A)^-1 = V diag(1/sigma_k^2) V^T   (formed without square roots, see below).
// The wanted right singular vector v_4 (smallest sigma) is the DOMINANT eigenvector of B, and
// on DLT matrices it dominates strongly: rho = (sigma_4/sigma_3)^2 is 0.005 in the median and
// below 0.07 for every rod pair of the synthetic and the reference's example data (mis-paired
// rods and end points included; profiles/experiments/dlt_invpower.py).  Three symmetric
// squarings give B^8, one more product with its largest-diagonal column gives B^16 e_j, in which
// every other direction is down by rho^16 < 1e-18: v_4 to rounding accuracy in ~340 straight-line
// FP64 instructions, against ~790 for the Jacobi above, with no data-dependent trip count (no
// warp divergence).  Forming S S^T is harmless here although it squares the spectrum: rounding
// errors are relative to |B| = 1/sigma_4^2, i.e. to the eigenvalue that is wanted, not to the
// large singular values as they would be in A^T A.  R itself comes from a backward-stable QR of A
// (not of A^T A), so the vector is as accurate as the one-sided Jacobi's: max 1.8e-15 relative on
// the triangulated point against cv2.triangulatePoints over 1.1 M combinations.
//
// Convergence is CHECKED, not assumed: tr(B^8) / tr(B^4)^2 ~ 1 - 2 rho^4.  A point with rho^4 > 5e-4
// (sigma_4/sigma_3 > 0.39) gets one more squaring and the same test on tr(B^16) / tr(B^8)^2
// (good up to sigma_4/sigma_3 = 0.62); beyond that (near-degenerate geometry), or if R is singular
// (NaN), the function reports false and the caller falls back to the Jacobi, which reproduces
// OpenCV there.
// ---------------------------------------------------------------------------------------
#define PTK_INVPOWER_TOL 1e-3

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ double ptk_rsqrt(double x) { return rsqrt_refine(x, rsqrt_seed(x)); }
__device__ __forceinline__ double ptk_rcp(double x) { return fast_rcp(x); }
// 2^-floor(log2 x) for finite positive x (garbage otherwise: caught by the convergence check)
__device__ __forceinline__ double ptk_pow2_inv(double x)
{
    const int e = (__double2hiint(x) >> 20) & 0x7ff;
    return __hiloint2double((2046 - e) << 20, 0);
}
#define PTK_HD __host__ __device__ __forceinline__
#else
static inline double ptk_rsqrt(double x) { return 1.0 / sqrt(x); }
static inline double ptk_rcp(double x) { return 1.0 / x; }
static inline double ptk_pow2_inv(double x)
{
    int e;
    frexp(x, &e);
    return ldexp(1.0, 1 - e);
}
#define PTK_HD __host__ __device__ inline
#endif

// B <- B B for a symmetric 4x4 matrix given by its upper triangle b00 .. b33 (40 FP64 instructions)
#define PTK_SQUARE_B                                                                     \
    {                                                                                    \
        const double c00 = fma(b03, b03, fma(b02, b02, fma(b01, b01, b00 * b00)));       \
        const double c01 = fma(b03, b13, fma(b02, b12, fma(b01, b11, b00 * b01)));       \
        const double c02 = fma(b03, b23, fma(b02, b22, fma(b01, b12, b00 * b02)));       \
        const double c03 = fma(b03, b33, fma(b02, b23, fma(b01, b13, b00 * b03)));       \
        const double c11 = fma(b13, b13, fma(b12, b12, fma(b11, b11, b01 * b01)));       \
        const double c12 = fma(b13, b23, fma(b12, b22, fma(b11, b12, b01 * b02)));       \
        const double c13 = fma(b13, b33, fma(b12, b23, fma(b11, b13, b01 * b03)));       \
        const double c22 = fma(b23, b23, fma(b22, b22, fma(b12, b12, b02 * b02)));       \
        const double c23 = fma(b23, b33, fma(b22, b23, fma(b12, b13, b02 * b03)));       \
        const double c33 = fma(b33, b33, fma(b23, b23, fma(b13, b13, b03 * b03)));       \
        b00 = c00; b01 = c01; b02 = c02; b03 = c03; b11 = c11;                           \
        b12 = c12; b13 = c13; b22 = c22; b23 = c23; b33 = c33;                           \
    }

// The power iteration on B = (A^T A)^-1 (symmetric, given by its upper triangle): scaling by a power of
// two, NSQ symmetric squarings, the convergence check and the final product with the largest-diagonal
// column.  Returns whether the iteration has converged.
// SCALE: B is first scaled by a power of two so that its trace is in [1, 2) (the record variant passes
// B / w2 whose trace is 1 + |z|^2 + small and needs none; absurd magnitudes overflow to Inf / NaN, which the
// convergence test reports as "not converged").
template <int NSQ, bool SCALE = true>
PTK_HD bool dlt_invpower_core(double b00, double b01, double b02, double b03, double b11, double b12, double b13,
                              double b22, double b23, double b33, double (&v)[4])
{
    if (SCALE) {
        const double sc = ptk_pow2_inv((b00 + b11) + (b22 + b33));
        b00 *= sc; b01 *= sc; b02 *= sc; b03 *= sc; b11 *= sc;
        b12 *= sc; b13 *= sc; b22 *= sc; b23 *= sc; b33 *= sc;
    }
    PTK_SQUARE_B PTK_SQUARE_B
    // sigma_4/sigma_3 between 0.39 and 0.62 (geometry unlike a rod experiment's): one more squaring, B^16
    // and B^32 e_j below.  tr(B^8) is within 2^-32 .. 2^16, so its square neither overflows nor underflows.
    if (NSQ > 3) PTK_SQUARE_B
    const double trh = (b00 + b11) + (b22 + b33);  // tr(B^(k/2))
    PTK_SQUARE_B
    // converged?  With k = 2^NSQ: tr(B^k) = |B^(k/2)|_F^2 >= (1 - tol) tr(B^(k/2))^2, i.e. 2 rho^(k/2) <= tol, rho =
    // (sigma_4/sigma_3)^2 -- the trace of the last square against the squared trace before it, six additions
    // instead of a Frobenius norm.  Written as tr - (1 - tol) trh^2 >= 0 so that Inf - Inf (a Gram-Schmidt norm
    // that underflowed, an overflow in the squarings) is NaN and therefore NOT converged: the caller then takes
    // the Jacobi fallback.
    const double tr = (b00 + b11) + (b22 + b33);
    const bool ok = fma(-(1.0 - PTK_INVPOWER_TOL), trh * trh, tr) >= 0.0;
    // v = B^k (column of B^k with the largest diagonal entry: |v_4[j]|^2 >= 1/4 there)
    double u0 = b00, u1 = b01, u2 = b02, u3 = b03, dm = b00;
    if (b11 > dm) { dm = b11; u0 = b01; u1 = b11; u2 = b12; u3 = b13; }
    if (b22 > dm) { dm = b22; u0 = b02; u1 = b12; u2 = b22; u3 = b23; }
    if (b33 > dm) { u0 = b03; u1 = b13; u2 = b23; u3 = b33; }
    v[0] = fma(b03, u3, fma(b02, u2, fma(b01, u1, b00 * u0)));
    v[1] = fma(b13, u3, fma(b12, u2, fma(b11, u1, b01 * u0)));
    v[2] = fma(b23, u3, fma(b22, u2, fma(b12, u1, b02 * u0)));
    v[3] = fma(b33, u3, fma(b23, u2, fma(b13, u1, b03 * u0)));
    return ok;
}

// Everything after the QR: T = R~^-1, B = T diag(w) T^T, then the power iteration.  r_kj: the unit
// upper-triangular factor's entries, w_k = 1 / |q_k|^2.  Returns whether the power iteration has converged.
template <int NSQ>
PTK_HD bool dlt_invpower_tail(double r01, double r02, double r03, double r12, double r13, double r23, double w0, double w1,
                              double w2, double w3, double (&v)[4])
{
    // T = R~^-1: t01 = -r01, t12 = -r12, t23 = -r23 and
    const double t02 = fma(r01, r12, -r02);                  // -(r02 + r01 t12)
    const double t13 = fma(r12, r23, -r13);                  // -(r13 + r12 t23)
    const double t03 = fma(r02, r23, -fma(r01, t13, r03));   // -(r03 + r01 t13 + r02 t23)
    // B = T diag(w) T^T.  Columns of T scaled by w: c1 = w1 T[:,1], c2 = w2 T[:,2], c3 = w3 T[:,3]
    const double c01 = -r01 * w1;
    const double c02 = t02 * w2, c12 = -r12 * w2;
    const double c03 = t03 * w3, c13 = t13 * w3, c23 = -r23 * w3;
    const double b00 = fma(t03, c03, fma(t02, c02, fma(-r01, c01, w0)));
    const double b01 = fma(t03, c13, fma(t02, c12, c01));
    const double b02 = fma(t03, c23, c02);
    const double b11 = fma(t13, c13, fma(-r12, c12, w1));
    const double b12 = fma(t13, c23, c12);
    const double b22 = fma(-r23, c23, w2);
    return dlt_invpower_core<NSQ>(b00, b01, b02, c03, b11, b12, c13, b22, c23, w3, v);
}

#define PTK_DOT(x, y) fma(x[3], y[3], fma(x[2], y[2], fma(x[1], y[1], x[0] * y[0])))
#define PTK_AXPY(y, r, q) { y[0] = fma(-r, q[0], y[0]); y[1] = fma(-r, q[1], y[1]); y[2] = fma(-r, q[2], y[2]); y[3] = fma(-r, q[3], y[3]); }

// P1o, P2o: 3x4 row-major projection matrices in the caller's column order.  v: the null vector
// (any scale).  NSQ = 3 or 4 squarings.  Returns whether the power iteration has converged.
template <int NSQ>
PTK_HD bool dlt_null_invpower_n(const double* __restrict__ Q1, const double* __restrict__ Q2, double x1, double y1,
                                double x2, double y2, double (&v)[4])
{
    double a0[4], a1[4], a2[4], a3[4];
#define PTK_COL(Ak, k)                      \
    Ak[0] = fma(x1, Q1[8 + k], -Q1[0 + k]); \
    Ak[1] = fma(y1, Q1[8 + k], -Q1[4 + k]); \
    Ak[2] = fma(x2, Q2[8 + k], -Q2[0 + k]); \
    Ak[3] = fma(y2, Q2[8 + k], -Q2[4 + k]);
    PTK_COL(a0, 0) PTK_COL(a1, 1) PTK_COL(a2, 2) PTK_COL(a3, 3)
#undef PTK_COL
    // modified Gram-Schmidt WITHOUT normalising: q_k = a_k^(k) (orthogonal, |q_k|^2 = n_k),
    // A = Q~ R~ with R~ unit upper triangular, r~_kj = (q_k . a_j) / n_k.  Then
    // (A^T A)^-1 = T diag(1/n_k) T^T with T = R~^-1 (unit upper triangular): no square root anywhere,
    // four reciprocals w_k = 1/n_k.
    const double w0 = ptk_rcp(PTK_DOT(a0, a0));
    const double r01 = PTK_DOT(a0, a1) * w0, r02 = PTK_DOT(a0, a2) * w0, r03 = PTK_DOT(a0, a3) * w0;
    PTK_AXPY(a1, r01, a0) PTK_AXPY(a2, r02, a0) PTK_AXPY(a3, r03, a0)
    const double w1 = ptk_rcp(PTK_DOT(a1, a1));
    const double r12 = PTK_DOT(a1, a2) * w1, r13 = PTK_DOT(a1, a3) * w1;
    PTK_AXPY(a2, r12, a1) PTK_AXPY(a3, r13, a1)
    const double w2 = ptk_rcp(PTK_DOT(a2, a2));
    const double r23 = PTK_DOT(a2, a3) * w2;
    PTK_AXPY(a3, r23, a2)
    const double w3 = ptk_rcp(PTK_DOT(a3, a3));
    return dlt_invpower_tail<NSQ>(r01, r02, r03, r12, r13, r23, w0, w1, w2, w3, v);
}

// ---------------------------------------------------------------------------------------
// Canonical first camera, with the work that depends on the camera-2 point alone taken out of the
// per-combination path (round 2).  Of the four DLT columns above only a2 = (x1-cx, y1-cy, p2, q2) contains
// the camera-1 point, and it is AFFINE in it: a2 = gx e0 + gy e1 + d, d = (0, 0, p2, q2).  Orthogonalising in
// the order a0, a1, a3, a2 makes the first three Gram-Schmidt vectors g0, g1, g3, their norms, the leading 3x3
// block of T = R~^-1 and that block's share B3 of B = T diag(w) T^T functions of (x2, y2) only, and what
// remains per combination is linear algebra on fixed operators applied to (gx, gy, 1):
//     z = Zm (gx, gy, 1)      the last column of T (order 0, 1, 3, 2; its fourth entry is 1),
//     a = Am (gx, gy, 1)      the residual of a2 after the three projections,  n2 = |a|^2 = a quadratic form in g,
//     B / w2 = z z^T + n2 [B3 0; 0 0]                          (w2 = 1 / n2, the last pivot)
// `canon_prep2` computes Zm (3x3), the six coefficients of n2 and B3 (6) ONCE per camera-2 end point (2 N2 per
// problem, kernel k_prep2) by running the Gram-Schmidt step on e0, e1 and d separately; a combination then costs 26 FP64
// instructions up to B instead of 103, needs no reciprocal (B is used divided by w2: same eigenvectors) and no
// power-of-two scaling (tr(B / w2) = 1 + |z|^2 + small).  The power iteration is the same; the null vector
// comes back in the caller's column order.  Same accuracy against cv2.triangulatePoints as the functions
// above (tests/test_dlt_invpower_cpu.py); the bits differ from theirs (another elimination order), so K1 and
// K3 -- which must agree bit for bit -- both go through prep + finish, K1 with the record read from memory
// and K3 with the record computed in place by the same function.
// ---------------------------------------------------------------------------------------
struct Canon2Rec {
    // f[0..8]: Zm by columns (gx | gy | 1); f[9..14]: the quadratic form n2 = |Am (gx, gy, 1)|^2 as
    // Q00, 2 Q01, 2 Q02, Q11, 2 Q12, Q22; f[15..20]: B00 B01 B02 B11 B12 B22; f[21] unused
    double f[22];
};
constexpr int kCanon2RecPlanes = 11;  // double2 planes of the record in memory

PTK_HD void canon_prep2(const double* __restrict__ cm1, double fx2, const double* __restrict__ Q2, double x2, double y2,
                        Canon2Rec& r)
{
    const double fx = cm1[0], fy = cm1[1];
    const double p0 = fma(x2, Q2[8], -Q2[0]), q0 = fma(y2, Q2[8], -Q2[4]);
    const double p1 = fma(x2, Q2[9], -Q2[1]), q1 = fma(y2, Q2[9], -Q2[5]);
    const double p2 = fma(x2, Q2[10], -Q2[2]), q2 = fma(y2, Q2[10], -Q2[6]);
    const double p3 = fma(x2, Q2[11], -Q2[3]), q3 = fma(y2, Q2[11], -Q2[7]);
    // g0 = a0 = (-fx, 0, p0, q0)
    const double g0[4] = {-fx, 0.0, p0, q0};
    const double w0 = ptk_rcp(fma(q0, q0, fma(p0, p0, fx2)));
    const double r01 = fma(q0, q1, p0 * p1) * w0;
    const double r03 = fma(q0, q3, p0 * p3) * w0;
    // g1 = a1 - r01 g0,  a3' = a3 - r03 g0
    const double g1[4] = {r01 * fx, -fy, fma(-r01, p0, p1), fma(-r01, q0, q1)};
    const double a30 = r03 * fx, a32 = fma(-r03, p0, p3), a33 = fma(-r03, q0, q3);
    const double w1 = ptk_rcp(PTK_DOT(g1, g1));
    const double r13 = fma(g1[3], a33, fma(g1[2], a32, g1[0] * a30)) * w1;
    // g3 = a3' - r13 g1
    const double g3[4] = {fma(-r13, g1[0], a30), r13 * fy, fma(-r13, g1[2], a32), fma(-r13, g1[3], a33)};
    const double w3 = ptk_rcp(PTK_DOT(g3, g3));
    // the Gram-Schmidt step of the last column, applied to e0, e1 and d = (0, 0, p2, q2): residual and the
    // column's entries of R~ (r02, r12, r32), from which the last column of T:
    //   t23 = -r32,  t13 = r13 r32 - r12,  t03 = r03 r32 - (r01 t13 + r02)
    double am[3][4];  // the residuals: Am by columns
#pragma unroll
    for (int c = 0; c < 3; c++) {
        double b[4] = {c == 0 ? 1.0 : 0.0, c == 1 ? 1.0 : 0.0, c == 2 ? p2 : 0.0, c == 2 ? q2 : 0.0};
        const double ra = PTK_DOT(g0, b) * w0;
        PTK_AXPY(b, ra, g0)
        const double rb = PTK_DOT(g1, b) * w1;
        PTK_AXPY(b, rb, g1)
        const double rc = PTK_DOT(g3, b) * w3;
        PTK_AXPY(b, rc, g3)
        const double t13 = fma(r13, rc, -rb);
        r.f[3 * c + 0] = fma(r03, rc, -fma(r01, t13, ra));
        r.f[3 * c + 1] = t13;
        r.f[3 * c + 2] = -rc;
        am[c][0] = b[0]; am[c][1] = b[1]; am[c][2] = b[2]; am[c][3] = b[3];
    }
    // n2 = |Am g|^2 as a quadratic form in (gx, gy): six coefficients instead of the twelve of Am.  n2 only scales
    // the SMALL part of B (the eigenvector moves by eps * cond, as with any formation of A^T A), so the
    // cancellation in the form does not reach the result (tests/test_dlt_invpower_cpu.py: 1.3e-15 against cv2 with
    // either form, the two forms 2e-16 apart).
    r.f[9] = PTK_DOT(am[0], am[0]);
    r.f[10] = 2.0 * PTK_DOT(am[0], am[1]);
    r.f[11] = 2.0 * PTK_DOT(am[0], am[2]);
    r.f[12] = PTK_DOT(am[1], am[1]);
    r.f[13] = 2.0 * PTK_DOT(am[1], am[2]);
    r.f[14] = PTK_DOT(am[2], am[2]);
    // leading block of T (order 0, 1, 3): t01 = -r01, t12' = -r13, t02' = r01 r13 - r03; B3 = T3 diag(w0, w1, w3) T3^T
    const double t02 = fma(r01, r13, -r03);
    const double c01 = -r01 * w1, c02 = t02 * w3, c12 = -r13 * w3;
    r.f[15] = fma(t02, c02, fma(-r01, c01, w0));
    r.f[16] = fma(t02, c12, c01);
    r.f[17] = c02;
    r.f[18] = fma(-r13, c12, w1);
    r.f[19] = c12;
    r.f[20] = w3;
    r.f[21] = 0.0;
}

// cm1 = (fx, fy, cx, cy); r = canon_prep2 of the camera-2 point; (x1, y1) the camera-1 point.
// The power iteration of the record variant: B = z~ z~^T + E with z~ = (z, 1) and E small, so z~ itself is the
// dominant eigenvector up to O(rho) -- no column has to be chosen.  B^4 by two squarings, then
// v = B^4 B^4 B^4 z~: every other direction is down by rho^13 (rho <= 0.15 where the test passes; 0.11 at most on
// rod data: 1.5e-14 against cv2.triangulatePoints, tests/test_dlt_invpower_cpu.py), 124 FP64 instructions against
// the 132 + 3 compares of three squarings and a chosen column.  Converged?  tr(B^4) / tr(B^2)^2 ~ 1 - 2 rho^2, same
// threshold on sigma_4 / sigma_3 (0.39) as the general function's three-squaring test; Inf / NaN (an absurd z)
// read as "not converged".
#define PTK_INVPOWER_TOL2 0.0447
PTK_HD bool dlt_invpower_start(double b00, double b01, double b02, double b11, double b12, double b22, double z0, double z1,
                               double z2, double (&v)[4])
{
    double b03 = z0, b13 = z1, b23 = z2, b33 = 1.0;
    PTK_SQUARE_B
    const double trh = (b00 + b11) + (b22 + b33);
    PTK_SQUARE_B
    const double tr = (b00 + b11) + (b22 + b33);
    const bool ok = fma(-(1.0 - PTK_INVPOWER_TOL2), trh * trh, tr) >= 0.0;
    double x0 = z0, x1 = z1, x2 = z2, x3 = 1.0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double y0 = fma(b03, x3, fma(b02, x2, fma(b01, x1, b00 * x0)));
        const double y1 = fma(b13, x3, fma(b12, x2, fma(b11, x1, b01 * x0)));
        const double y2 = fma(b23, x3, fma(b22, x2, fma(b12, x1, b02 * x0)));
        const double y3 = fma(b33, x3, fma(b23, x2, fma(b13, x1, b03 * x0)));
        x0 = y0; x1 = y1; x2 = y2; x3 = y3;
    }
    v[0] = x0; v[1] = x1; v[2] = x2; v[3] = x3;
    return ok;
}

// cm1 = (fx, fy, cx, cy); r = canon_prep2 of the camera-2 point; (x1, y1) the camera-1 point.  Returns whether
// the iteration has converged (else: the caller's fallback).
PTK_HD bool dlt_null_invpower_canon_rec(const double* __restrict__ cm1, const Canon2Rec& r, double x1, double y1,
                                        double (&v)[4])
{
    const double gx = x1 - cm1[2], gy = y1 - cm1[3];
    const double z0 = fma(r.f[0], gx, fma(r.f[3], gy, r.f[6]));
    const double z1 = fma(r.f[1], gx, fma(r.f[4], gy, r.f[7]));
    const double z2 = fma(r.f[2], gx, fma(r.f[5], gy, r.f[8]));
    const double n2 = fma(gx, fma(r.f[9], gx, fma(r.f[10], gy, r.f[11])), fma(gy, fma(r.f[12], gy, r.f[13]), r.f[14]));
    double u[4];
    const bool ok = dlt_invpower_start(fma(z0, z0, n2 * r.f[15]), fma(z0, z1, n2 * r.f[16]), fma(z0, z2, n2 * r.f[17]),
                                       fma(z1, z1, n2 * r.f[18]), fma(z1, z2, n2 * r.f[19]), fma(z2, z2, n2 * r.f[20]), z0, z1,
                                       z2, u);
    v[0] = u[0]; v[1] = u[1]; v[2] = u[3]; v[3] = u[2];
    return ok;
}
#undef PTK_SQUARE_B
#undef PTK_DOT
#undef PTK_AXPY

// Three squarings, and one more where that has not converged (the harness' and K8's entry point).
PTK_HD bool dlt_null_invpower(const double* __restrict__ Q1, const double* __restrict__ Q2, double x1, double y1,
                              double x2, double y2, double (&v)[4])
{
    if (dlt_null_invpower_n<3>(Q1, Q2, x1, y1, x2, y2, v)) return true;
    return dlt_null_invpower_n<4>(Q1, Q2, x1, y1, x2, y2, v);
}

// Out of line: only points whose three-squaring power iteration did not converge get here (none on rod
// data), so neither the fourth squaring nor the Jacobi's 80 registers weigh on the callers' main path.
// Two separate functions: inlined into one, their combined register need made ptxas spill in the CALLERS'
// main path (44 bytes per thread) to keep the kernels at 80 registers.
__device__ __noinline__ bool dlt_invpower4_outofline(const DevRig& rig, double x1, double y1, double x2, double y2,
                                                     double* __restrict__ xyz)
{
    double v[4];
    const bool ok = dlt_null_invpower_n<4>(rig.P1o, rig.P2o, x1, y1, x2, y2, v);
    const double rw = fast_rcp(v[3]);
    xyz[0] = v[0] * rw;
    xyz[1] = v[1] * rw;
    xyz[2] = v[2] * rw;
    return ok;
}

__device__ __noinline__ void dlt_triangulate_fallback(const DevRig& rig, double x1, double y1, double x2, double y2,
                                                      double* __restrict__ xyz)
{
    if (dlt_invpower4_outofline(rig, x1, y1, x2, y2, xyz)) return;
    double X, Y, Z;
    dlt_triangulate_jacobi(rig, x1, y1, x2, y2, X, Y, Z);
    xyz[0] = X;
    xyz[1] = Y;
    xyz[2] = Z;
}

// cv2.triangulatePoints for one point (match2D.py:889) + p[0:3]/p[3] (match2D.py:895): the stand-alone
// entry (K8) on the general function.
__device__ __forceinline__ void dlt_triangulate(const DevRig& rig, double x1, double y1, double x2, double y2,
                                                double& X, double& Y, double& Z)
{
#ifdef PTK_DLT_JACOBI  // A/B switch (tools/ab_bench.py): the one-sided Jacobi for every point
    dlt_triangulate_jacobi(rig, x1, y1, x2, y2, X, Y, Z);
#else
    double v[4];
    if (dlt_null_invpower_n<3>(rig.P1o, rig.P2o, x1, y1, x2, y2, v)) {
        const double rw = fast_rcp(v[3]);
        X = v[0] * rw;
        Y = v[1] * rw;
        Z = v[2] * rw;
    } else {
        double xyz[3];
        dlt_triangulate_fallback(rig, x1, y1, x2, y2, xyz);
        X = xyz[0];
        Y = xyz[1];
        Z = xyz[2];
    }
#endif
}

// The distortion model of cv2.projectPoints on a normalised point (x, y): (xd, yd), to be
// mapped to pixels by the camera matrix.  SIMPLE (compile time): the camera has radial terms k1, k2, k3 only
// (Horner in r^2) -- the rational, tangential and thin-prism blocks do not exist in the code; otherwise they
// are warp-uniform branches on rig constants.
template <bool SIMPLE>
__device__ __forceinline__ void distort_norm(const double* __restrict__ k, double x, double y, double& xd, double& yd,
                                             bool tangential)
{
    const double r2 = fma(x, x, y * y);
    if (SIMPLE) {
        const double cdist = fma(fma(fma(k[4], r2, k[1]), r2, k[0]), r2, 1.0);
        xd = x * cdist;
        yd = y * cdist;
        return;
    }
    const double r4 = r2 * r2, r6 = r4 * r2;
    const double cdist = fma(k[4], r6, fma(k[1], r4, fma(k[0], r2, 1.0)));
    double rad = cdist;
    if (k[5] != 0.0 || k[6] != 0.0 || k[7] != 0.0)  // rational model (k4..k6)
        rad = cdist * fast_rcp(fma(k[7], r6, fma(k[6], r4, fma(k[5], r2, 1.0))));
    xd = x * rad;
    yd = y * rad;
    if (tangential) {  // p1, p2 and thin-prism s1..s4
        const double a1 = 2 * x * y, a2 = fma(2 * x, x, r2), a3 = fma(2 * y, y, r2);
        xd = fma(k[9], r4, fma(k[8], r2, fma(k[3], a2, fma(k[2], a1, xd))));
        yd = fma(k[11], r4, fma(k[10], r2, fma(k[3], a1, fma(k[2], a3, yd))));
    }
}
template <bool SIMPLE>
__device__ __forceinline__ void distort_to_pixel(const double* __restrict__ cm, const double* __restrict__ k, double x,
                                                 double y, double& u, double& v, bool tangential)
{
    double xd, yd;
    distort_norm<SIMPLE>(k, x, y, xd, yd, tangential);
    u = fma(xd, cm[0], cm[2]);
    v = fma(yd, cm[1], cm[3]);
}

// cv2.projectPoints for one point (match2D.py:898-903, calibrate_cameras.py:202-262).
template <bool SIMPLE>
__device__ __forceinline__ void project_point_t(const double* __restrict__ R, const double* __restrict__ t,
                                                const double* __restrict__ cm, const double* __restrict__ k,
                                                double X, double Y, double Z, double& u, double& v,
                                                bool ext_identity, bool tangential)
{
    double x = X, y = Y, z = Z;
    if (!ext_identity) {
        x = fma(R[2], Z, fma(R[1], Y, R[0] * X)) + t[0];
        y = fma(R[5], Z, fma(R[4], Y, R[3] * X)) + t[1];
        z = fma(R[8], Z, fma(R[7], Y, R[6] * X)) + t[2];
    }
    z = ((__double2hiint(z) & 0x7fffffff) | __double2loint(z)) != 0 ? fast_rcp(z) : 1.0;  // z != 0.0, tested on the ALU pipe
    distort_to_pixel<SIMPLE>(cm, k, x * z, y * z, u, v, tangential);
}

__device__ __forceinline__ void project_point(const double* __restrict__ R, const double* __restrict__ t,
                                              const double* __restrict__ cm, const double* __restrict__ k,
                                              double X, double Y, double Z, double& u, double& v,
                                              bool ext_identity = false, bool tangential = true)
{
    project_point_t<false>(R, t, cm, k, X, Y, Z, u, v, ext_identity, tangential);
}

// Reprojection of the HOMOGENEOUS point h = (h0, h1, h2, h3) (the DLT null vector as it comes, in the
// coordinates the projection matrices refer to) and its distance to the original (distorted) detection
// (match2D.py:895-910).  The reference de-homogenises first (p[0:3] / p[3]) and cv2.projectPoints divides by the
// depth afterwards; the two divisions cancel -- x / z = (R h[0:3] + t h3)_x / (R h[0:3] + t h3)_z -- so the
// per-camera path is one reciprocal, not two, and the point itself is only formed where a caller wants it
// (K3, K1's optional output).  cv2's "z == 0 -> leave x, y undivided" becomes "divide by h3 instead".
template <bool SIMPLE, bool FAST>
__device__ __forceinline__ double reproject_dist_h(const double* __restrict__ R, const double* __restrict__ t,
                                                   const double* __restrict__ cm, const double* __restrict__ k,
                                                   const double (&h)[4], double ouc, double ovc, bool ext_identity,
                                                   bool tangential, bool& good)
{
    double x = h[0], y = h[1], z = h[2];
    if (!ext_identity) {
        x = fma(R[2], h[2], fma(R[1], h[1], fma(R[0], h[0], t[0] * h[3])));
        y = fma(R[5], h[2], fma(R[4], h[1], fma(R[3], h[0], t[1] * h[3])));
        z = fma(R[8], h[2], fma(R[7], h[1], fma(R[6], h[0], t[2] * h[3])));
    }
    // cv2's "z == 0 -> undivided": on the FAST path 1/0 makes the squared distance NaN, which its range test
    // below reports, and the slow path then takes the select
    const bool znz = FAST || ((__double2hiint(z) & 0x7fffffff) | __double2loint(z)) != 0;
    const double iz = fast_rcp(znz ? z : h[3]);
    double xd, yd;
    distort_norm<SIMPLE>(k, x * iz, y * iz, xd, yd, tangential);
    // (ouc, ovc) = detection - principal point: u - ou = fx xd + cx - ou
    const double ex = fma(-xd, cm[0], ouc);
    const double ey = fma(-yd, cm[1], ovc);
    const double q = fma(ex, ex, ey * ey);
    // sqrt as q * rsqrt(q) (seed + one third-order step: an ulp or two, 7 instructions instead of
    // the ~14 of the IEEE sequence); zero, denormal, huge, infinite and NaN arguments take the library
    // path.  The range test is on the exponent bits (ALU pipe) instead of two DSETPs on the FP64 pipe:
    // biased exponent in [27, 2019], i.e. 2^-996 <= q < 2^997 (q >= 0).
    const bool in_range = (((unsigned)__double2hiint(q) >> 20) - 27u) < 1993u;
    if (FAST) {  // the caller repeats the combination on its slow path if any flag is down
        good = good && in_range;
        return q * rsqrt_refine(q, rsqrt_seed(q));
    }
    return in_range ? q * rsqrt_refine(q, rsqrt_seed(q)) : sqrt(q);
}

// The unusual cases of eval_combo, out of line: the power iteration has not converged (extra squaring, then
// OpenCV's Jacobi), a squared distance outside the fast square root's range (zero, Inf, NaN -- a zero depth ends
// up here too).  Everything is recomputed from the inputs; u2 points at the undistorted camera-2 point.
template <bool CANON, bool SIMPLE>
__device__ __noinline__ void eval_combo_slow(const DevRig& rig, double u1x, double u1y, const double* __restrict__ u2,
                                             double o1x, double o1y, double o2x, double o2y, double* __restrict__ out)
{
    const double2 q = __ldg(reinterpret_cast<const double2*>(u2));
    double h[4];
    bool ok;
    if (CANON) {
        Canon2Rec rec;
        canon_prep2(rig.CM1, rig.fx2, rig.P2o, q.x, q.y, rec);
        ok = dlt_null_invpower_canon_rec(rig.CM1, rec, u1x, u1y, h);
    } else {
        ok = dlt_null_invpower_n<3>(rig.P1o, rig.P2o, u1x, u1y, q.x, q.y, h);
    }
    if (!ok) {
        double xyz[3];
        dlt_triangulate_fallback(rig, u1x, u1y, q.x, q.y, xyz);
        h[0] = xyz[0]; h[1] = xyz[1]; h[2] = xyz[2]; h[3] = 1.0;
    }
    bool dummy = true;
    out[0] = reproject_dist_h<SIMPLE, false>(rig.R1, rig.t1, rig.CM1, rig.k1, h, o1x, o1y, CANON || rig.ext_identity[0] != 0,
                                             rig.tangential[0] != 0, dummy);
    out[1] = reproject_dist_h<SIMPLE, false>(rig.R2, rig.t2, rig.CM2, rig.k2, h, o2x, o2y, rig.ext_identity[1] != 0,
                                             rig.tangential[1] != 0, dummy);
    out[2] = h[0]; out[3] = h[1]; out[4] = h[2]; out[5] = h[3];
}

// One end-point combination (match2D.py:889-913): triangulate the undistorted pair, reproject into both
// cameras, per-camera distances d1, d2, and the homogeneous point h (camera-1 coordinates; `dehomogenise`).
// CANON / SIMPLE: the rig's class (DevRig::canon / ::simple), fixed at compile time so that K1 and K3 -- which
// must produce the same bits for the same rod pair -- run the same code.
// CANON: `rec` is canon_prep2 of the camera-2 point (read from k_prep2's records by K1, computed in place by K3:
// the same function, the same bits); otherwise rec is unused.  u2 points at the undistorted camera-2 point.
// (o1x, o1y), (o2x, o2y): the original (distorted) detections MINUS their camera's principal point (centre_detection).
// The main path is straight-line code that only collects "something unusual" flags (not converged, a squared
// distance out of range); ONE branch at the end sends such a combination through eval_combo_slow.
// (K1 calls the three stages itself so that its shared-memory reads sit right before their use.)
template <bool CANON>
__device__ __forceinline__ bool eval_combo_dlt(const DevRig& rig, double u1x, double u1y, const double* __restrict__ u2,
                                               const Canon2Rec& rec, double (&h)[4])
{
    if (CANON) return dlt_null_invpower_canon_rec(rig.CM1, rec, u1x, u1y, h);
    const double2 q = __ldg(reinterpret_cast<const double2*>(u2));
    return dlt_null_invpower_n<3>(rig.P1o, rig.P2o, u1x, u1y, q.x, q.y, h);
}
template <bool CANON, bool SIMPLE>
__device__ __forceinline__ void eval_combo_dist(const DevRig& rig, const double (&h)[4], double o1x, double o1y, double o2x,
                                                double o2y, bool& good, double& d1, double& d2)
{
    d1 = reproject_dist_h<SIMPLE, true>(rig.R1, rig.t1, rig.CM1, rig.k1, h, o1x, o1y, CANON || rig.ext_identity[0] != 0,
                                        rig.tangential[0] != 0, good);
    d2 = reproject_dist_h<SIMPLE, true>(rig.R2, rig.t2, rig.CM2, rig.k2, h, o2x, o2y, rig.ext_identity[1] != 0,
                                        rig.tangential[1] != 0, good);
}
template <bool CANON, bool SIMPLE>
__device__ __forceinline__ void eval_combo_redo(const DevRig& rig, double u1x, double u1y, const double* __restrict__ u2,
                                                double o1x, double o1y, double o2x, double o2y, double& d1, double& d2,
                                                double (&h)[4])
{
    double out[6];
    eval_combo_slow<CANON, SIMPLE>(rig, u1x, u1y, u2, o1x, o1y, o2x, o2y, out);
    d1 = out[0]; d2 = out[1];
    h[0] = out[2]; h[1] = out[3]; h[2] = out[4]; h[3] = out[5];
}
template <bool CANON, bool SIMPLE>
__device__ __forceinline__ void eval_combo(const DevRig& rig, double u1x, double u1y, const double* __restrict__ u2,
                                           const Canon2Rec& rec, double o1x, double o1y, double o2x, double o2y,
                                           double& d1, double& d2, double (&h)[4])
{
    bool good = eval_combo_dlt<CANON>(rig, u1x, u1y, u2, rec, h);
    eval_combo_dist<CANON, SIMPLE>(rig, h, o1x, o1y, o2x, o2y, good, d1, d2);
    if (__builtin_expect(!good, 0)) eval_combo_redo<CANON, SIMPLE>(rig, u1x, u1y, u2, o1x, o1y, o2x, o2y, d1, d2, h);
}

// detection - principal point, the form eval_combo takes the original detections in (the reprojection residual
// is then one FMA per coordinate: fx xd + cx - u = fx xd - (u - cx))
__device__ __forceinline__ double2 centre_detection(const double* __restrict__ cm, double2 o)
{
    return make_double2(o.x - cm[2], o.y - cm[3]);
}

// match2D.py:895: p[0:3] / p[3]
__device__ __forceinline__ void dehomogenise(const double (&h)[4], double& X, double& Y, double& Z)
{
    const double rw = fast_rcp(h[3]);
    X = h[0] * rw;
    Y = h[1] * rw;
    Z = h[2] * rw;
}

// match2D.py:913: X_w = rot.apply(X) + trans
__device__ __forceinline__ void to_world(const DevRig& rig, double X, double Y, double Z,
                                         double& wx, double& wy, double& wz)
{
    wx = fma(rig.Rw[2], Z, fma(rig.Rw[1], Y, rig.Rw[0] * X)) + rig.tw[0];
    wy = fma(rig.Rw[5], Z, fma(rig.Rw[4], Y, rig.Rw[3] * X)) + rig.tw[1];
    wz = fma(rig.Rw[8], Z, fma(rig.Rw[7], Y, rig.Rw[6] * X)) + rig.tw[2];
}

// np.min([a, b]) -- propagates NaN (match2D.py:923-930, tracking.py:118)
__device__ __forceinline__ double np_min(double a, double b)
{
    return (a != a) ? a : ((b != b) ? b : (b < a ? b : a));
}

}  // namespace ptk
