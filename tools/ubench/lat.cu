#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* t, double x0) {
    double x = x0 + threadIdx.x * 1e-9, y = 1.0000001;
    long long c[8];
    c[0] = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) { x = fma(x, y, 1e-9); x = fma(x, y, 1e-9); x = fma(x, y, 1e-9); x = fma(x, y, 1e-9); }
    c[1] = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) { x = rsqrt(x + 2.0); x = rsqrt(x + 2.0); x = rsqrt(x + 2.0); x = rsqrt(x + 2.0); }
    c[2] = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) { x = __shfl_sync(0xffffffffu, x, (i + 1) & 31); x = __shfl_sync(0xffffffffu, x, (i + 3) & 31); x = __shfl_sync(0xffffffffu, x, (i + 5) & 31); x = __shfl_sync(0xffffffffu, x, (i + 7) & 31); }
    c[3] = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) { x = 1.0 / (x + 2.0); x = 1.0 / (x + 2.0); x = 1.0 / (x + 2.0); x = 1.0 / (x + 2.0); }
    c[4] = clock64();
    float f = (float)x;
#pragma unroll 1
    for (int i = 0; i < 256; ++i) { f = rsqrtf(f + 2.0f); f = rsqrtf(f + 2.0f); f = rsqrtf(f + 2.0f); f = rsqrtf(f + 2.0f); }
    c[5] = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) { x = (double)(float)(x + 1.0); x = (double)(float)(x + 1.0); x = (double)(float)(x + 1.0); x = (double)(float)(x + 1.0); }
    c[6] = clock64();
    __shared__ double sm[64];
    sm[threadIdx.x] = x; __syncwarp();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) { x = sm[((int)x) & 31]; x = sm[((int)x + 1) & 31]; x = sm[((int)x + 2) & 31]; x = sm[((int)x + 3) & 31]; }
    c[7] = clock64();
    out[threadIdx.x] = x + f;
    if (threadIdx.x == 0) for (int i = 0; i < 7; ++i) t[i] = c[i + 1] - c[i];
}
int main() {
    double* o; long long* t; cudaMalloc(&o, 256); cudaMalloc(&t, 64);
    for (int rep = 0; rep < 2; ++rep) { k<<<1, 32>>>(o, t, 1.5); cudaDeviceSynchronize(); }
    long long h[7]; cudaMemcpy(h, t, 56, cudaMemcpyDeviceToHost);
    const char* n[7] = {"DFMA", "rsqrt(double)", "SHFL double", "1.0/x double", "rsqrtf", "double->float->double (+DADD)", "LDS double (+cvt)"};
    for (int i = 0; i < 7; ++i) printf("%-32s %.1f cycles per dependent op\n", n[i], h[i] / 1024.0);
    return 0;
}
