// Latency microbenchmarks for the fp64 scalar chains of the sweep kernels (one warp, dependent ops).
#include <cstdio>
#include <cuda_runtime.h>
#define N 256
__global__ void k(double *out, long long *cyc, double a, double b) {
    double x = a; long long t0, t1; int lane = threadIdx.x & 31;
    // DADD
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = x + b;
    t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = x * b;
    t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = fma(x, b, a);
    t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) x = a / (x + b);
    t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) x = exp(x * 1e-3);
    t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) {
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) x += __shfl_xor_sync(0xffffffffu, x, m);
    }
    t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // shfl alone (64-bit)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = __shfl_xor_sync(0xffffffffu, x, 1);
    t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
    // logistic dloss
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) { double z = x * b; x = -b / (exp(z) + 1.0); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[7] = t1 - t0;
    // float ops for comparison
    float f = (float)a, g = (float)b;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) f = fmaf(f, g, f);
    t1 = clock64(); if (threadIdx.x == 0) cyc[8] = t1 - t0;
    // LDS dependent chain
    __shared__ int sm[64];
    sm[lane] = (lane + 1) & 31; __syncwarp();
    int p = lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) p = ((volatile int *)sm)[p];
    t1 = clock64(); if (threadIdx.x == 0) cyc[9] = t1 - t0;
    // threadfence_block
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) { sm[lane] = i; __threadfence_block(); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[10] = t1 - t0;
    out[threadIdx.x] = x + f + p;
}
// ping-pong between two warps of one CTA through a volatile smem flag
__global__ void pingpong(long long *cyc) {
    __shared__ volatile int flag;
    if (threadIdx.x == 0) flag = 0;
    __syncthreads();
    int w = threadIdx.x >> 5;
    long long t0 = clock64();
    for (int i = 0; i < N; i++) {
        if (w == 0) { while (flag != 2 * i) {} if ((threadIdx.x & 31) == 0) flag = 2 * i + 1; __syncwarp(); }
        else if (w == 1) { while (flag != 2 * i + 1) {} if ((threadIdx.x & 31) == 0) flag = 2 * i + 2; __syncwarp(); }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[11] = t1 - t0;
}
int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 16 * 8); cudaMemset(cyc, 0, 128);
    for (int rep = 0; rep < 2; rep++) { k<<<1, 32>>>(out, cyc, 1.000001, 0.999999); pingpong<<<1, 64>>>(cyc); }
    long long h[16]; cudaMemcpy(h, cyc, 128, cudaMemcpyDeviceToHost);
    const char *nm[] = {"dadd", "dmul", "dfma", "ddiv(+add)", "exp(+mul)", "warp_allsum(5x shfl+dadd)", "shfl64",
                        "logistic dloss", "ffma", "lds chain", "sts+fence_block", "smem pingpong (2 hops)"};
    for (int i = 0; i < 12; i++) printf("%-28s %8.1f cycles/op\n", nm[i], (double)h[i] / N);
    // multi-warp contention: same kernel with 16 warps
    k<<<1, 512>>>(out, cyc, 1.000001, 0.999999); cudaMemcpy(h, cyc, 128, cudaMemcpyDeviceToHost);
    printf("with 16 warps on the SM:\n");
    for (int i = 0; i < 11; i++) printf("%-28s %8.1f cycles/op\n", nm[i], (double)h[i] / N);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
