/*
 * ptk_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the ARITHMETIC on the reference's stereo-reconstruction hot
 * path, i.e. of what the third-party binaries the reference calls compute:
 *   cv2.undistortImagePoints / triangulatePoints / projectPoints  (OpenCV 4.13, pinned
 *     `opencv-python-headless >=4.8.1.78`, ParticleDetection/pyproject.toml:43)
 *   scipy.optimize.linear_sum_assignment  (SciPy 1.18.1, pinned `scipy >=1.9.3`)
 *   numpy.linalg.norm / diff / minimum as used by tracking.py
 * None of that source is under /root/reference; the published algorithms are restated
 * here and anchored on the reference's call sites, cited per
 * function (paths relative to ParticleDetection/src/ParticleDetection/reconstruct_3D/).
 *
 * Parity status: PINNED -- tests/test_c_oracle.py checks every function here against
 * the installed cv2 / scipy / numpy binaries, against oracle/port.py, against the
 * reference itself when /root/reference exists, and against tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product never does.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 * -ffp-contract=off matters: the restatement uses separate multiplies and adds where
 * the originals do.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/ptk.h"

#define PTKO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------
 * B.1  cv2.undistortImagePoints(src, CM, dist): 5 fixed-point iterations (Python default
 * criteria = COUNT 5), R = I, P = CM.  Called at match2D.py:847,856.
 * ---------------------------------------------------------------------------------- */
static void undistort_point(const double* CM, const double* k, double u, double v,
                            double* ou, double* ov)
{
    const double fx = CM[0], fy = CM[4], cx = CM[2], cy = CM[5];
    const double ifx = 1.0 / fx, ify = 1.0 / fy;
    double x = (u - cx) * ifx, y = (v - cy) * ify;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; j++) {
        double r2 = x * x + y * y;
        double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) /
                        (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
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
    /* re-projection with P = CM: [fx 0 cx; 0 fy cy; 0 0 1] * (x, y, 1) */
    double xx = fx * x + 0.0 * y + cx;
    double yy = 0.0 * x + fy * y + cy;
    double ww = 1.0 / (0.0 * x + 0.0 * y + 1.0);
    *ou = xx * ww;
    *ov = yy * ww;
}

/* A.2: the reference's reshape quirk (match2D.py:847-854).  rods[N][2][2] flat = 4N
 * doubles; pseudo-point j = (flat[j], flat[j+2N]); results go back to the same flat
 * positions.  N == 1: proper points (flat[0],flat[1]), (flat[2],flat[3]); output flat
 * = [x1', x2', y1', y2'] (the write-back transposes). */
PTKO_API void ptko_undistort_rods(const double* CM, const double* dist14, int N,
                                  const double* rods, double* out)
{
    if (N <= 0) return;
    if (N == 1) {
        double a, b, c, d;
        undistort_point(CM, dist14, rods[0], rods[1], &a, &b);
        undistort_point(CM, dist14, rods[2], rods[3], &c, &d);
        out[0] = a; out[1] = c; out[2] = b; out[3] = d;
        return;
    }
    for (int j = 0; j < 2 * N; j++)
        undistort_point(CM, dist14, rods[j], rods[j + 2 * N], &out[j], &out[j + 2 * N]);
}

/* ------------------------------------------------------------------------------------
 * B.2  cv2.triangulatePoints for one point (match2D.py:889-895): build the 4x4 DLT
 * matrix, one-sided (Hestenes) Jacobi SVD as OpenCV's JacobiSVDImpl_<double> does for
 * small matrices, take the right singular vector of the smallest singular value, then
 * the reference's p[0:3]/p[3].  stats (optional): [0] += rotations, [1] += sweeps.
 * ---------------------------------------------------------------------------------- */
PTKO_API void ptko_triangulate(const double* P1, const double* P2,
                               double x1, double y1, double x2, double y2,
                               double* X3, double* Xh4, long long* stats)
{
    /* At[i][k] = A[k][i]: rows of At are the columns of A */
    double At[4][4], Vt[4][4], W[4];
    for (int k = 0; k < 4; k++) {
        At[k][0] = x1 * P1[8 + k] - P1[0 + k];
        At[k][1] = y1 * P1[8 + k] - P1[4 + k];
        At[k][2] = x2 * P2[8 + k] - P2[0 + k];
        At[k][3] = y2 * P2[8 + k] - P2[4 + k];
    }
    const double eps = DBL_EPSILON * 10;
    for (int i = 0; i < 4; i++) {
        double sd = 0;
        for (int k = 0; k < 4; k++) sd += At[i][k] * At[i][k];
        W[i] = sd;
        for (int k = 0; k < 4; k++) Vt[i][k] = (i == k);
    }
    for (int iter = 0; iter < 30; iter++) {
        int changed = 0;
        for (int i = 0; i < 3; i++)
            for (int j = i + 1; j < 4; j++) {
                double a = W[i], p = 0, b = W[j];
                for (int k = 0; k < 4; k++) p += At[i][k] * At[j][k];
                if (fabs(p) <= eps * sqrt(a * b)) continue;
                p *= 2;
                double beta = a - b, gamma = hypot(p, beta), c, s;
                if (beta < 0) {
                    double delta = (gamma - beta) * 0.5;
                    s = sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
                a = b = 0;
                for (int k = 0; k < 4; k++) {
                    double t0 = c * At[i][k] + s * At[j][k];
                    double t1 = -s * At[i][k] + c * At[j][k];
                    At[i][k] = t0; At[j][k] = t1;
                    a += t0 * t0; b += t1 * t1;
                }
                W[i] = a; W[j] = b;
                changed = 1;
                if (stats) stats[0]++;
                for (int k = 0; k < 4; k++) {
                    double t0 = c * Vt[i][k] + s * Vt[j][k];
                    double t1 = -s * Vt[i][k] + c * Vt[j][k];
                    Vt[i][k] = t0; Vt[j][k] = t1;
                }
            }
        if (stats) stats[1]++;
        if (!changed) break;
    }
    for (int i = 0; i < 4; i++) {
        double sd = 0;
        for (int k = 0; k < 4; k++) sd += At[i][k] * At[i][k];
        W[i] = sqrt(sd);
    }
    /* selection sort, descending; only the permutation of Vt rows matters here */
    int idx[4] = {0, 1, 2, 3};
    for (int i = 0; i < 3; i++) {
        int j = i;
        for (int k = i + 1; k < 4; k++)
            if (W[j] < W[k]) j = k;
        if (i != j) {
            double tw = W[i]; W[i] = W[j]; W[j] = tw;
            int ti = idx[i]; idx[i] = idx[j]; idx[j] = ti;
        }
    }
    const double* v = Vt[idx[3]];
    if (Xh4) { Xh4[0] = v[0]; Xh4[1] = v[1]; Xh4[2] = v[2]; Xh4[3] = v[3]; }
    X3[0] = v[0] / v[3];
    X3[1] = v[1] / v[3];
    X3[2] = v[2] / v[3];
}

/* ------------------------------------------------------------------------------------
 * B.3  cv2.projectPoints(X, Rmat, t, CM, dist) for one point (match2D.py:898-903; the
 * reference passes 3x3 rotation MATRICES as rvec, which OpenCV uses as matrices).
 * ---------------------------------------------------------------------------------- */
PTKO_API void ptko_project(const double* R, const double* t, const double* CM,
                           const double* k, const double* X, double* uv)
{
    double x = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
    double y = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
    double z = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
    z = z ? 1. / z : 1;
    x *= z; y *= z;
    double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
    double a1 = 2 * x * y, a2 = r2 + 2 * x * x, a3 = r2 + 2 * y * y;
    double cdist = 1 + k[0] * r2 + k[1] * r4 + k[4] * r6;
    double icdist2 = 1. / (1 + k[5] * r2 + k[6] * r4 + k[7] * r6);
    double xd = x * cdist * icdist2 + k[2] * a1 + k[3] * a2 + k[8] * r2 + k[9] * r4;
    double yd = y * cdist * icdist2 + k[2] * a3 + k[3] * a1 + k[10] * r2 + k[11] * r4;
    uv[0] = xd * CM[0] + CM[2];
    uv[1] = yd * CM[4] + CM[5];
}

/* one end-point combination: match2D.py:889-913.  d[2] = per-camera reprojection
 * distance, Xw[3] = world-transformed point. */
static void combo(const ptk_rig* rig, const double* u1, const double* u2,
                  const double* o1, const double* o2, double* d, double* Xw, long long* stats)
{
    double X[3], q[2];
    ptko_triangulate(rig->P1, rig->P2, u1[0], u1[1], u2[0], u2[1], X, NULL, stats);
    ptko_project(rig->R1, rig->t1, rig->CM1, rig->dist1, X, q);
    double ex = o1[0] - q[0], ey = o1[1] - q[1];
    d[0] = sqrt(ex * ex + ey * ey);
    ptko_project(rig->R2, rig->t2, rig->CM2, rig->dist2, X, q);
    ex = o2[0] - q[0]; ey = o2[1] - q[1];
    d[1] = sqrt(ex * ex + ey * ey);
    const double* M = rig->Rw;
    Xw[0] = M[0] * X[0] + M[1] * X[1] + M[2] * X[2] + rig->tw[0];
    Xw[1] = M[3] * X[0] + M[4] * X[1] + M[5] * X[2] + rig->tw[1];
    Xw[2] = M[6] * X[0] + M[7] * X[1] + M[8] * X[2] + rig->tw[2];
}

/* all four combos of one rod pair, combo c -> (e1,e2) = (0,0),(0,1),(1,0),(1,1)
 * (match2D.py:865-887).  Returns cost (A.6) and fills choice, Xw[4][3], d[4][2]. */
static double pair_cost(const ptk_rig* rig, const double* und1, const double* und2,
                        const double* rod1, const double* rod2,
                        int* choice, double Xw[4][3], double d[4][2], long long* stats)
{
    double e[4];
    for (int c = 0; c < 4; c++) {
        int e1 = c >> 1, e2 = c & 1;
        combo(rig, und1 + 2 * e1, und2 + 2 * e2, rod1 + 2 * e1, rod2 + 2 * e2, d[c], Xw[c], stats);
        e[c] = rig->agg_sum ? (d[c][0] + d[c][1]) : (d[c][0] + d[c][1]) / 2.0;
    }
    double s03 = e[0] + e[3], s12 = e[1] + e[2];
    *choice = (s03 <= s12);
    return s03 < s12 ? s03 : (s12 < s03 ? s12 : (s03 != s03 ? s03 : s12)); /* np.min propagates NaN */
}

/* match2D.py:846-939 for one problem.  cost[N1*ld], choice[N1*ld] (ld = row stride);
 * optional p_world[N1][N2][4][3], rep_errs[N1][N2][4][2] (dense, no padding). */
PTKO_API void ptko_cost_problem(const ptk_rig* rig, int N1, int N2,
                                const double* cam1, const double* cam2, int ld,
                                double* cost, uint8_t* choice,
                                double* p_world, double* rep_errs, long long* stats)
{
    double* u1 = (double*)malloc(sizeof(double) * 4 * (size_t)(N1 > 0 ? N1 : 1));
    double* u2 = (double*)malloc(sizeof(double) * 4 * (size_t)(N2 > 0 ? N2 : 1));
    ptko_undistort_rods(rig->CM1, rig->dist1, N1, cam1, u1);
    ptko_undistort_rods(rig->CM2, rig->dist2, N2, cam2, u2);
    for (int i = 0; i < N1; i++)
        for (int j = 0; j < N2; j++) {
            double Xw[4][3], d[4][2];
            int ch;
            double c = pair_cost(rig, u1 + 4 * i, u2 + 4 * j, cam1 + 4 * i, cam2 + 4 * j, &ch, Xw, d, stats);
            cost[(size_t)i * ld + j] = c;
            if (choice) choice[(size_t)i * ld + j] = (uint8_t)ch;
            if (p_world) memcpy(p_world + ((size_t)i * N2 + j) * 12, Xw, sizeof(double) * 12);
            if (rep_errs) memcpy(rep_errs + ((size_t)i * N2 + j) * 8, d, sizeof(double) * 8);
        }
    free(u1); free(u2);
}

/* ------------------------------------------------------------------------------------
 * C  scipy.optimize.linear_sum_assignment (rectangular shortest augmenting path, Crouse
 * 2016) with SciPy's scan order and tie rule; called at match2D.py:933, tracking.py:122.
 * cost[nr][ld].  Writes min(nr,nc) pairs, rows ascending.  Returns PTK_PROBLEM_*.
 * ---------------------------------------------------------------------------------- */
PTKO_API int ptko_lsap(int nr0, int nc0, const double* cost0, int ld,
                       int32_t* row_ind, int32_t* col_ind)
{
    if (nr0 <= 0 || nc0 <= 0) return PTK_PROBLEM_EMPTY;
    int transpose = nc0 < nr0;
    int nr = transpose ? nc0 : nr0, nc = transpose ? nr0 : nc0;
    double* cost = (double*)malloc(sizeof(double) * (size_t)nr * nc);
    for (int i = 0; i < nr; i++)
        for (int j = 0; j < nc; j++) {
            double c = transpose ? cost0[(size_t)j * ld + i] : cost0[(size_t)i * ld + j];
            if (c != c || c == -INFINITY) { free(cost); return PTK_PROBLEM_INVALID; }
            cost[(size_t)i * nc + j] = c;
        }
    double* u = (double*)calloc(nr, sizeof(double));
    double* v = (double*)calloc(nc, sizeof(double));
    double* spc = (double*)malloc(sizeof(double) * nc);
    int* path = (int*)malloc(sizeof(int) * nc);
    int* col4row = (int*)malloc(sizeof(int) * nr);
    int* row4col = (int*)malloc(sizeof(int) * nc);
    int* remaining = (int*)malloc(sizeof(int) * nc);
    uint8_t* SR = (uint8_t*)malloc(nr);
    uint8_t* SC = (uint8_t*)malloc(nc);
    for (int i = 0; i < nr; i++) col4row[i] = -1;
    for (int j = 0; j < nc; j++) { row4col[j] = -1; path[j] = -1; }
    int status = PTK_PROBLEM_OK;

    for (int cur = 0; cur < nr && status == PTK_PROBLEM_OK; cur++) {
        double minVal = 0;
        int num_remaining = nc, i = cur, sink = -1;
        for (int it = 0; it < nc; it++) remaining[it] = nc - it - 1;
        memset(SR, 0, nr); memset(SC, 0, nc);
        for (int j = 0; j < nc; j++) spc[j] = INFINITY;
        while (sink == -1) {
            int index = -1;
            double lowest = INFINITY;
            SR[i] = 1;
            for (int it = 0; it < num_remaining; it++) {
                int j = remaining[it];
                double r = minVal + cost[(size_t)i * nc + j] - u[i] - v[j];
                if (r < spc[j]) { path[j] = i; spc[j] = r; }
                if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) {
                    lowest = spc[j];
                    index = it;
                }
            }
            minVal = lowest;
            if (minVal == INFINITY) { status = PTK_PROBLEM_INFEASIBLE; break; }
            int j = remaining[index];
            if (row4col[j] == -1) sink = j; else i = row4col[j];
            SC[j] = 1;
            remaining[index] = remaining[--num_remaining];
        }
        if (status != PTK_PROBLEM_OK) break;
        u[cur] += minVal;
        for (int r = 0; r < nr; r++)
            if (SR[r] && r != cur) u[r] += minVal - spc[col4row[r]];
        for (int j = 0; j < nc; j++)
            if (SC[j]) v[j] -= minVal - spc[j];
        int j = sink;
        while (1) {
            int r = path[j];
            row4col[j] = r;
            int t = col4row[r]; col4row[r] = j; j = t;
            if (r == cur) break;
        }
    }
    if (status == PTK_PROBLEM_OK) {
        if (transpose) {
            /* rows of the transposed problem are original columns: sort by original row */
            int k = 0;
            for (int r0 = 0; r0 < nc; r0++) /* original rows = transposed columns */
                if (row4col[r0] != -1) { row_ind[k] = r0; col_ind[k] = row4col[r0]; k++; }
        } else {
            for (int r = 0; r < nr; r++) { row_ind[r] = r; col_ind[r] = col4row[r]; }
        }
    }
    free(cost); free(u); free(v); free(spc); free(path); free(col4row); free(row4col);
    free(remaining); free(SR); free(SC);
    return status;
}

/* ------------------------------------------------------------------------------------
 * match2D.py:843-983 (renumber=True) for one problem -> rows[n][18], costs[n], indices.
 * Includes the loop-index quirk (row m pairs cam1 rod m with cam2 rod col_ind[m] also when
 * N1 > N2) and the crossed-case reversal of the 2D columns.
 * ---------------------------------------------------------------------------------- */
PTKO_API int ptko_match_problem(const ptk_rig* rig, int N1, int N2,
                                const double* cam1, const double* cam2,
                                double* cost /* [N1*N2] scratch or output */,
                                int32_t* row_ind, int32_t* col_ind,
                                double* rows, double* costs, long long* stats)
{
    if (N1 <= 0 || N2 <= 0) return PTK_PROBLEM_EMPTY;
    uint8_t* choice = (uint8_t*)malloc((size_t)N1 * N2);
    ptko_cost_problem(rig, N1, N2, cam1, cam2, N2, cost, choice, NULL, NULL, stats);
    int st = ptko_lsap(N1, N2, cost, N2, row_ind, col_ind);
    free(choice);
    if (st != PTK_PROBLEM_OK) return st;
    int n = N1 < N2 ? N1 : N2;
    double* u1 = (double*)malloc(sizeof(double) * 4 * (size_t)N1);
    double* u2 = (double*)malloc(sizeof(double) * 4 * (size_t)N2);
    ptko_undistort_rods(rig->CM1, rig->dist1, N1, cam1, u1);
    ptko_undistort_rods(rig->CM2, rig->dist2, N2, cam2, u2);
    for (int m = 0; m < n; m++) {
        int i2 = col_ind[m], ch;
        double Xw[4][3], d[4][2];
        pair_cost(rig, u1 + 4 * m, u2 + 4 * i2, cam1 + 4 * m, cam2 + 4 * i2, &ch, Xw, d, NULL);
        const double* e1 = ch ? Xw[0] : Xw[1];
        const double* e2 = ch ? Xw[3] : Xw[2];
        double* o = rows + (size_t)m * 18;
        for (int k = 0; k < 3; k++) {
            o[k] = e1[k]; o[3 + k] = e2[k];
            o[6 + k] = (e1[k] + e2[k]) / 2;
        }
        double dx = e2[0] - e1[0], dy = e2[1] - e1[1], dz = e2[2] - e1[2];
        o[9] = sqrt(dx * dx + dy * dy + dz * dz);
        const double* c1 = cam1 + 4 * m; const double* c2 = cam2 + 4 * i2;
        if (ch) {
            for (int k = 0; k < 4; k++) { o[10 + k] = c1[k]; o[14 + k] = c2[k]; }
        } else {
            o[10] = c1[2]; o[11] = c1[3]; o[12] = c1[0]; o[13] = c1[1];
            o[14] = c2[2]; o[15] = c2[3]; o[16] = c2[0]; o[17] = c2[1];
        }
        costs[m] = cost[(size_t)row_ind[m] * N2 + col_ind[m]];
    }
    free(u1); free(u2);
    return PTK_PROBLEM_OK;
}

/* batch over P problems in the padded tensor layout of include/ptk.h, OpenMP over problems.
 * Used by bench.py as the "C port" CPU baseline and by the GPU parity tests at sizes where
 * the Python port would take minutes. */
PTKO_API int ptko_match_batch(const ptk_rig* rig, int P, int Nmax,
                              const double* cam1, const double* cam2,
                              const int32_t* n1, const int32_t* n2,
                              double* cost /* [P,Nmax,Nmax] or NULL */,
                              int32_t* row_ind, int32_t* col_ind, int32_t* n_out, int32_t* status,
                              double* rows, double* match_cost, int nthreads, long long* stats)
{
    long long st_rot = 0, st_swp = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : st_rot, st_swp)
    for (int p = 0; p < P; p++) {
        int N1 = n1[p], N2 = n2[p];
        int n = N1 < N2 ? N1 : N2;
        size_t o = (size_t)p * Nmax;
        for (int m = 0; m < Nmax; m++) { row_ind[o + m] = -1; col_ind[o + m] = -1; }
        n_out[p] = 0;
        if (n <= 0) { status[p] = PTK_PROBLEM_EMPTY; continue; }
        double* c = (double*)malloc(sizeof(double) * (size_t)N1 * N2);
        double* r = (double*)malloc(sizeof(double) * (size_t)n * 18);
        double* mc = (double*)malloc(sizeof(double) * (size_t)n);
        long long s[2] = {0, 0};
        int stt = ptko_match_problem(rig, N1, N2, cam1 + o * 4, cam2 + o * 4, c,
                                     row_ind + o, col_ind + o, r, mc, s);
        status[p] = stt;
        st_rot += s[0]; st_swp += s[1];
        if (cost)
            for (int i = 0; i < N1; i++)
                memcpy(cost + ((size_t)p * Nmax + i) * Nmax, c + (size_t)i * N2, sizeof(double) * N2);
        if (stt == PTK_PROBLEM_OK) {
            n_out[p] = n;
            memcpy(rows + o * 18, r, sizeof(double) * (size_t)n * 18);
            memcpy(match_cost + o, mc, sizeof(double) * (size_t)n);
        }
        free(c); free(r); free(mc);
    }
    if (stats) { stats[0] = st_rot; stats[1] = st_swp; }
    return 0;
}

/* ------------------------------------------------------------------------------------
 * D.1  tracking.py:100-121 cost for one frame pair, numpy's operation order:
 * diff = next - cur, squares, ((x^2 + y^2) + z^2), sqrt, sum of the two norms, min.
 * ends[f] = [N][2][3].
 * ---------------------------------------------------------------------------------- */
static double norm3(double ax, double ay, double az, double bx, double by, double bz)
{
    double dx = ax - bx, dy = ay - by, dz = az - bz;
    return sqrt((dx * dx + dy * dy) + dz * dz);
}

PTKO_API void ptko_track_cost(int N, const double* cur, const double* nxt, double* cost)
{
    for (int a = 0; a < N; a++)
        for (int b = 0; b < N; b++) {
            const double* p1a = cur + 6 * a;       /* p1[f,a]   */
            const double* p2b = cur + 6 * b + 3;   /* p2[f,b]   */
            const double* q1a = nxt + 6 * a;       /* p1[f+1,a] */
            const double* q2b = nxt + 6 * b + 3;   /* p2[f+1,b] */
            double d0 = norm3(q1a[0], q1a[1], q1a[2], p1a[0], p1a[1], p1a[2]) +
                        norm3(q2b[0], q2b[1], q2b[2], p2b[0], p2b[1], p2b[2]);
            double d1 = norm3(p1a[0], p1a[1], p1a[2], q2b[0], q2b[1], q2b[2]) +
                        norm3(p2b[0], p2b[1], p2b[2], q1a[0], q1a[1], q1a[2]);
            /* np.min over the stacked pair propagates NaN */
            cost[(size_t)a * N + b] = (d0 != d0) ? d0 : ((d1 != d1) ? d1 : (d1 < d0 ? d1 : d0));
        }
}

/* tracking.py:97-123,136-137 for C colours: ends[C,F,N,2,3] -> col_ind[C,F-1,N],
 * total_cost[C,F-1], status[C,F-1]; cost[C,F-1,N,N] optional. */
PTKO_API int ptko_track_batch(int C, int F, int N, const double* ends, double* cost,
                              int32_t* col_ind, double* total_cost, int32_t* status, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    long long Q = (long long)C * (F - 1);
#pragma omp parallel for schedule(dynamic, 16)
    for (long long q = 0; q < Q; q++) {
        int c = (int)(q / (F - 1)), f = (int)(q % (F - 1));
        const double* cur = ends + ((size_t)c * F + f) * N * 6;
        double* cm = (double*)malloc(sizeof(double) * (size_t)N * N);
        int32_t* ri = (int32_t*)malloc(sizeof(int32_t) * N);
        ptko_track_cost(N, cur, cur + (size_t)N * 6, cm);
        int st = ptko_lsap(N, N, cm, N, ri, col_ind + q * N);
        status[q] = st;
        double tot = 0;
        if (st == PTK_PROBLEM_OK)
            for (int k = 0; k < N; k++) tot += cm[(size_t)k * N + col_ind[q * N + k]];
        total_cost[q] = tot;
        if (cost) memcpy(cost + (size_t)q * N * N, cm, sizeof(double) * (size_t)N * N);
        free(cm); free(ri);
    }
    return 0;
}

/* D.3 chain pass (extension): ids[0] = ids_in or arange, ids[f+1][col[f][a]] = ids[f][a].
 * A frame pair whose assignment failed (col < 0 or >= N: the LSAP reported INVALID / INFEASIBLE,
 * e.g. NaN end points of a ragged frame) breaks the chain: positions that no valid entry maps
 * to get id -1, and -1 propagates along the frames. */
PTKO_API void ptko_chain_tracks(int C, int F, int N, const int32_t* col_ind,
                                const int32_t* ids_in, int32_t* track_id)
{
    for (int c = 0; c < C; c++) {
        int32_t* ids = track_id + (size_t)c * F * N;
        for (int a = 0; a < N; a++) ids[a] = ids_in ? ids_in[(size_t)c * N + a] : a;
        for (int f = 0; f + 1 < F; f++) {
            const int32_t* col = col_ind + ((size_t)c * (F - 1) + f) * N;
            for (int a = 0; a < N; a++) ids[(size_t)(f + 1) * N + a] = -1;
            for (int a = 0; a < N; a++)
                if (col[a] >= 0 && col[a] < N) ids[(size_t)(f + 1) * N + col[a]] = ids[(size_t)f * N + a];
        }
    }
}

/* ------------------------------------------------------------------------------------
 * E  matchND.create_weights (matchND.py:141-217): the (1:2) slice quirk is kept.
 * ---------------------------------------------------------------------------------- */
static double nrm3(const double* a, const double* b)
{
    double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    return sqrt(dx * dx + dy * dy + dz * dz);
}

PTKO_API void ptko_create_weights(int Np, int N1, int N2, const double* p3, const double* prev,
                                  const double* re, double* weights, double* choices)
{
    double wmax = -INFINITY;
    for (int i = 0; i < Np; i++) {
        const double* q1 = prev + 6 * i;
        const double* q2 = q1 + 3;
        for (int j = 0; j < N1; j++)
            for (int k = 0; k < N2; k++) {
                const double* p = p3 + ((size_t)j * N2 + k) * 12;
                const double* r = re + ((size_t)j * N2 + k) * 8;
                const double *c1p1 = p, *c2p1 = p + 3, *c2p2 = p + 6, *c1p2 = p + 9;
                double c1_re = ((r[0] + r[1]) + r[6]) + r[7];
                double c2_re = r[2] + r[3];
                double s[4];
                s[0] = (nrm3(c1p1, q1) + nrm3(c1p2, q2)) * c1_re;
                s[1] = (nrm3(c2p1, q1) + nrm3(c2p2, q2)) * c2_re;
                s[2] = (nrm3(c2p1, q2) + nrm3(c2p2, q1)) * c2_re;
                s[3] = (nrm3(c1p1, q2) + nrm3(c1p2, q1)) * c1_re;
                int am = 0;
                for (int t = 1; t < 4; t++)
                    if (s[t] < s[am]) am = t;
                size_t o = ((size_t)i * N1 + j) * N2 + k;
                weights[o] = s[am];
                choices[o] = (double)am;
                if (s[am] > wmax) wmax = s[am];
            }
    }
    size_t tot = (size_t)Np * N1 * N2;
    for (size_t o = 0; o < tot; o++) weights[o] = wmax - weights[o];
}

PTKO_API int ptko_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
