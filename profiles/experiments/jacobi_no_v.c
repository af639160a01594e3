// Experiment (CPU, not product code): one-sided Jacobi WITHOUT accumulating V.  After
// convergence B = A V has orthogonal columns b_k = sigma_k u_k; the right singular vectors of
// the three large singular values are v_k ~ A^T b_k, and the wanted smallest one is the
// generalised cross product of those three.  Compares the triangulated point against the
// V-accumulating Jacobi (OpenCV order).   gcc -O2 -ffp-contract=off -o /tmp/jnv jacobi_no_v.c -lm
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <float.h>
#include <string.h>
static void jac(const double A[4][4] /*rows r, cols k*/, int withV, double* X)
{
    double At[4][4], Vt[4][4], W[4]; int perm[4] = {0,1,2,3};
    for (int k = 0; k < 4; k++) for (int r = 0; r < 4; r++) At[k][r] = A[r][k];
    for (int i = 0; i < 4; i++) { double sd = 0; for (int k = 0; k < 4; k++) { sd += At[i][k]*At[i][k]; Vt[i][k] = (i==k); } W[i] = sd; }
    for (int i = 0; i < 3; i++) { int j = i; for (int k = i+1; k < 4; k++) if (W[k] > W[j]) j = k;
        if (j != i) { double t = W[i]; W[i] = W[j]; W[j] = t; int tp = perm[i]; perm[i] = perm[j]; perm[j] = tp;
            for (int k = 0; k < 4; k++) { t = At[i][k]; At[i][k] = At[j][k]; At[j][k] = t; } } }
    double eps2 = 100*DBL_EPSILON*DBL_EPSILON;
    for (int iter = 0; iter < 30; iter++) { int changed = 0;
        for (int i = 0; i < 3; i++) for (int j = i+1; j < 4; j++) { double a = W[i], b = W[j], p = 0; for (int k = 0; k < 4; k++) p += At[i][k]*At[j][k];
            if (iter >= 2 && !(p*p > eps2*(a*b))) continue;
            double beta = a-b, p2 = p+p, x1 = p2*p2+beta*beta; if (!(x1 > 0)) continue;
            double g = 1.0/sqrt(x1), t = 0.5+0.5*fabs(beta)*g, big = sqrt(t), sm = p*g/big, c = beta<0?sm:big, s = beta<0?big:sm; a = b = 0;
            for (int k = 0; k < 4; k++) { double t0 = c*At[i][k]+s*At[j][k], t1 = -s*At[i][k]+c*At[j][k]; At[i][k] = t0; At[j][k] = t1; a += t0*t0; b += t1*t1;
                if (withV) { t0 = c*Vt[i][k]+s*Vt[j][k]; t1 = -s*Vt[i][k]+c*Vt[j][k]; Vt[i][k] = t0; Vt[j][k] = t1; } }
            W[i] = a; W[j] = b; changed = 1; }
        if (iter >= 2 && !changed) break; }
    int m = 0; for (int i = 1; i < 4; i++) if (W[i] <= W[m]) m = i;
    double v[4];
    if (withV) { for (int k = 0; k < 4; k++) v[perm[k]] = Vt[m][k]; }
    else {
        double w[3][4]; int n = 0;
        for (int k = 0; k < 4; k++) { if (k == m) continue; for (int c = 0; c < 4; c++) { double s = 0; for (int r = 0; r < 4; r++) s += A[r][c]*At[k][r]; w[n][c] = s; } n++; }
        // generalised cross product of w[0], w[1], w[2]
        #define M2(a,b) (w[1][a]*w[2][b] - w[1][b]*w[2][a])
        v[0] =  w[0][1]*M2(2,3) - w[0][2]*M2(1,3) + w[0][3]*M2(1,2);
        v[1] = -(w[0][0]*M2(2,3) - w[0][2]*M2(0,3) + w[0][3]*M2(0,2));
        v[2] =  w[0][0]*M2(1,3) - w[0][1]*M2(0,3) + w[0][3]*M2(0,1);
        v[3] = -(w[0][0]*M2(1,2) - w[0][1]*M2(0,2) + w[0][2]*M2(0,1));
    }
    X[0] = v[0]/v[3]; X[1] = v[1]/v[3]; X[2] = v[2]/v[3];
}
int main(int argc, char** argv)
{
    FILE* f = fopen(argv[1], "rb"); double P1[12], P2[12]; long M;
    if (fread(&M, 8, 1, f) != 1 || fread(P1, 8, 12, f) != 12 || fread(P2, 8, 12, f) != 12) return 1;
    double* q = malloc(M*32); if (fread(q, 8, M*4, f) != (size_t)M*4) return 1;
    double worst = 0; long n10 = 0, n12 = 0; double sum = 0;
    for (long m = 0; m < M; m++) {
        double A[4][4];
        for (int k = 0; k < 4; k++) { A[0][k] = q[4*m]*P1[8+k]-P1[k]; A[1][k] = q[4*m+1]*P1[8+k]-P1[4+k]; A[2][k] = q[4*m+2]*P2[8+k]-P2[k]; A[3][k] = q[4*m+3]*P2[8+k]-P2[4+k]; }
        double X0[3], X1[3]; jac(A, 1, X0); jac(A, 0, X1);
        double sc = fmax(fmax(fabs(X0[0]), fabs(X0[1])), fmax(fabs(X0[2]), 1.0));
        double e = fmax(fmax(fabs(X1[0]-X0[0]), fabs(X1[1]-X0[1])), fabs(X1[2]-X0[2]))/sc;
        if (e > worst) worst = e; if (e > 1e-10) n10++; if (e > 1e-12) n12++; sum += e;
    }
    printf("no-V vs V-accumulating Jacobi over %ld combos: max rel err %.2e, mean %.2e, n(>1e-12)=%ld, n(>1e-10)=%ld\n", M, worst, sum/M, n12, n10);
    return 0;
}
