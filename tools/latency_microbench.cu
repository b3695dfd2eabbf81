#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *t, double a, double b, const double *tab, double *sm_out)
{
    __shared__ double sm[64];
    sm[threadIdx.x] = a; __syncthreads();
    double x = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) x = fma(x, b, a);
    }
    long long t1 = clock64();
    double y = a * 3.0 + threadIdx.x;
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) y = y * b;
    }
    long long t2 = clock64();
    double z = a * 5.0;
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) { sm[threadIdx.x] = z; __syncwarp(); z = sm[threadIdx.x ^ 1] + 1.0; __syncwarp(); }
    }
    long long t3 = clock64();
    double w = a * 7.0;
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) w = __shfl_sync(0xffffffffu, w, (threadIdx.x + 1) & 31) + 1.0;
    }
    long long t4 = clock64();
    double v = a * 9.0;
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v = __ldg(tab + ((int)v & 63)) + v;
    }
    long long t5 = clock64();
    double q = a * 11.0;
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) { __syncthreads(); q = q + 1.0; }
    }
    long long t6 = clock64();
    float fq = (float)a;
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) fq = fmaf(fq, 1.0001f, 0.5f);
    }
    long long t7 = clock64();
    out[threadIdx.x] = q + fq + x + y + z + w + v;
    if (threadIdx.x == 0) { t[0] = t1 - t0; t[1] = t2 - t1; t[2] = t3 - t2; t[3] = t4 - t3; t[4] = t5 - t4; t[5] = t6 - t5; t[6] = t7 - t6; }
}
int main()
{
    double *out, *tab; long long *t;
    cudaMalloc(&out, 1024); cudaMalloc(&t, 64); cudaMalloc(&tab, 512); cudaMemset(tab, 0, 512);
    for (int threads : {32, 64}) {
        k<<<1, threads>>>(out, t, 0.5, 0.999, tab, nullptr); cudaDeviceSynchronize();
        k<<<1, threads>>>(out, t, 0.5, 0.999, tab, nullptr); cudaDeviceSynchronize();
        long long h[7]; cudaMemcpy(h, t, 56, cudaMemcpyDeviceToHost);
        printf("threads %d: per-op cycles  DFMA %.1f  DMUL %.1f  STS+LDS+DADD %.1f  SHFL64+DADD %.1f  LDG(L1)+cvt+DADD %.1f  BAR+DADD %.1f  FFMA %.1f\n", threads,
               h[0] / 1024.0, h[1] / 1024.0, h[2] / 1024.0, h[3] / 1024.0, h[4] / 1024.0, h[5] / 1024.0, h[6] / 1024.0);
    }
    return 0;
}
