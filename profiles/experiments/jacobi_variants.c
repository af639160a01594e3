// Experiment (CPU, not product code): how do Jacobi variants change rotations/sweeps and the
// triangulated point, relative to the OpenCV-order baseline?  Build:
//   gcc -O2 -ffp-contract=off -o /tmp/jv profiles/experiments/jacobi_variants.c -lm
// Input: binary file of doubles [M][4] = (x1,y1,x2,y2) undistorted combos + P1[12] + P2[12] header.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <float.h>
#include <string.h>

typedef struct { double eps; int sort_first; int small_angle; int fixed_sweeps; int rr; } Var;

static void tri(const double* P1, const double* P2, const double* q, Var v, double* X, long* rot, long* swp, long* chk)
{
    double At[4][4], Vt[4][4], W[4];
    for (int k = 0; k < 4; k++) {
        At[k][0] = q[0] * P1[8 + k] - P1[0 + k];
        At[k][1] = q[1] * P1[8 + k] - P1[4 + k];
        At[k][2] = q[2] * P2[8 + k] - P2[0 + k];
        At[k][3] = q[3] * P2[8 + k] - P2[4 + k];
    }
    for (int i = 0; i < 4; i++) {
        double sd = 0;
        for (int k = 0; k < 4; k++) sd += At[i][k] * At[i][k];
        W[i] = sd;
        for (int k = 0; k < 4; k++) Vt[i][k] = (i == k);
    }
    if (v.sort_first) {  // sort columns by decreasing norm (de Rijk)
        for (int i = 0; i < 3; i++) {
            int j = i;
            for (int k = i + 1; k < 4; k++) if (W[k] > W[j]) j = k;
            if (j != i) {
                double t;
                t = W[i]; W[i] = W[j]; W[j] = t;
                for (int k = 0; k < 4; k++) { t = At[i][k]; At[i][k] = At[j][k]; At[j][k] = t; t = Vt[i][k]; Vt[i][k] = Vt[j][k]; Vt[j][k] = t; }
            }
        }
    }
    double eps2 = v.eps * v.eps;
    for (int iter = 0; iter < 30; iter++) {
        int changed = 0;
        static const int ORD[2][6][2] = {{{0,1},{0,2},{0,3},{1,2},{1,3},{2,3}}, {{0,1},{2,3},{0,2},{1,3},{0,3},{1,2}}};
        for (int pi = 0; pi < 6; pi++) {
            {
                int i = ORD[v.rr][pi][0], j = ORD[v.rr][pi][1];
                double a = W[i], b = W[j], p = 0;
                for (int k = 0; k < 4; k++) p += At[i][k] * At[j][k];
                (*chk)++;
                if (!(p * p > eps2 * (a * b))) continue;
                double beta = a - b, p2 = p + p;
                double g = 1.0 / sqrt(p2 * p2 + beta * beta);
                double t = 0.5 + 0.5 * fabs(beta) * g;
                double big = sqrt(t), small = p * g / big;
                double c, s;
                if (v.small_angle) { c = big; s = beta < 0 ? -small : small; /* rotation by the small angle; no implicit swap */ }
                else { c = beta < 0 ? small : big; s = beta < 0 ? big : small; }
                a = b = 0;
                for (int k = 0; k < 4; k++) {
                    double t0 = c * At[i][k] + s * At[j][k], t1 = -s * At[i][k] + c * At[j][k];
                    At[i][k] = t0; At[j][k] = t1; a += t0 * t0; b += t1 * t1;
                    t0 = c * Vt[i][k] + s * Vt[j][k]; t1 = -s * Vt[i][k] + c * Vt[j][k];
                    Vt[i][k] = t0; Vt[j][k] = t1;
                }
                W[i] = a; W[j] = b;
                changed = 1; (*rot)++;
            }
        }
        (*swp)++;
        if (v.fixed_sweeps ? (iter + 1 >= v.fixed_sweeps) : !changed) break;
    }
    int m = 0;
    for (int i = 1; i < 4; i++) if (W[i] <= W[m]) m = i;
    X[0] = Vt[m][0] / Vt[m][3]; X[1] = Vt[m][1] / Vt[m][3]; X[2] = Vt[m][2] / Vt[m][3];
}

int main(int argc, char** argv)
{
    FILE* f = fopen(argv[1], "rb");
    double P1[12], P2[12];
    long M;
    if (fread(&M, sizeof(long), 1, f) != 1) return 1;
    if (fread(P1, 8, 12, f) != 12 || fread(P2, 8, 12, f) != 12) return 1;
    double* q = malloc(M * 4 * 8);
    if (fread(q, 8, M * 4, f) != (size_t)(M * 4)) return 1;
    Var vars[] = {{10 * DBL_EPSILON, 0, 0, 0}, {1e-13, 0, 0, 0}, {1e-12, 0, 0, 0}, {1e-11, 0, 0, 0}, {1e-10, 0, 0, 0},
                  {10 * DBL_EPSILON, 1, 0, 0}, {1e-12, 1, 0, 0}, {10 * DBL_EPSILON, 0, 1, 0}, {1e-12, 0, 1, 0},
                  {10 * DBL_EPSILON, 1, 1, 0}, {1e-12, 1, 1, 0}, {10 * DBL_EPSILON, 1, 0, 0, 1}, {10 * DBL_EPSILON, 0, 0, 0, 1}};
    int nv = sizeof(vars) / sizeof(vars[0]);
    double* base = malloc(M * 3 * 8);
    for (int v = 0; v < nv; v++) {
        long rot = 0, swp = 0, chk = 0;
        double worst = 0, worst99 = 0;
        int maxs = 0;
        double* errs = malloc(M * 8);
        for (long m = 0; m < M; m++) {
            double X[3];
            long s0 = swp;
            tri(P1, P2, q + 4 * m, vars[v], X, &rot, &swp, &chk);
            if (swp - s0 > maxs) maxs = swp - s0;
            if (v == 0) memcpy(base + 3 * m, X, 24);
            double sc = fmax(fmax(fabs(base[3 * m]), fabs(base[3 * m + 1])), fmax(fabs(base[3 * m + 2]), 1.0));
            double e = fmax(fmax(fabs(X[0] - base[3 * m]), fabs(X[1] - base[3 * m + 1])), fabs(X[2] - base[3 * m + 2])) / sc;
            errs[m] = e;
            if (e > worst) worst = e;
        }
        long big = 0;
        for (long m = 0; m < M; m++) if (errs[m] > 1e-10) big++;
        printf("eps=%.1e sort=%d small=%d fixed=%d rr=%d : rot/combo %.2f sweeps %.2f (max %d) checks %.1f  max_rel_err %.2e  n(err>1e-10)=%ld of %ld\n",
               vars[v].eps, vars[v].sort_first, vars[v].small_angle, vars[v].fixed_sweeps, vars[v].rr, (double)rot / M, (double)swp / M, maxs,
               (double)chk / M, worst, big, M);
        free(errs);
    }
    return 0;
}
