// Latency probe (sm_100a): dependent chains of the instructions on k_lsap_trk's critical path.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_probe lat_probe.cu && ./lat_probe
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
__global__ void k(double* out, long long* cyc, double seed)
{
    __shared__ double sm[64];
    sm[threadIdx.x] = seed + threadIdx.x;
    __syncthreads();
    long long t0, t1;
    // DADD chain
    double x = seed;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = x + seed;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // DFMA chain
    double y = seed;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) y = fma(y, seed, seed);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // REDUX min chain (u32)
    unsigned u = threadIdx.x + (unsigned)seed;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) u = __reduce_min_sync(0xffffffffu, u + threadIdx.x) + 1;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // SHFL chain
    int s = threadIdx.x;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) s = __shfl_sync(0xffffffffu, s, (s + 1) & 31);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // LDS pointer chase
    int idx = threadIdx.x & 31;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) idx = (int)sm[idx & 63] & 31;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // DSETP + FSEL chain
    double z = seed, w = seed * 0.5;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) { z = (z < w) ? w + 1.0 : z; }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // integer ALU chain (LOP3/IADD)
    unsigned a = threadIdx.x;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) a = (a ^ 0x5bd1e995u) + (a >> 3);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = t1 - t0;
    out[threadIdx.x] = x + y + u + s + idx + z + a;
}
int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, 64 * 8); cudaMalloc(&cyc, 8 * 8);
    k<<<1, 32>>>(out, cyc, 1.000001);
    k<<<1, 32>>>(out, cyc, 1.000001);
    long long h[8];
    cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    const char* names[] = {"DADD", "DFMA", "REDUX.MIN+IADD", "SHFL(+IADD,LOP)", "LDS(+F2I,LOP)", "DSETP+FSEL(+DADD)", "2xALU"};
    for (int i = 0; i < 7; i++) printf("%-20s %.1f cycles per dependent op\n", names[i], (double)h[i] / N);
    return 0;
}
