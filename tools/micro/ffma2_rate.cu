// Microbenchmark: sustained issue rate of FFMA2 (fma.rn.f32x2) vs scalar FFMA on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu && ./ffma2_rate
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
    float2 acc[16];
    float2 w[4];
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, seed);
    for (int i = 0; i < 4; ++i) w[i] = make_float2(seed + i, 0.5f * seed - i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) acc[i] = fma2(w[i & 3], acc[(i + 5) & 15], acc[i]);        // distinct a, b, c
            else if (MODE == 1) acc[i] = fma2(w[0], w[1], acc[i]);                       // shared a, b
            else { acc[i].x = fmaf(w[i & 3].x, acc[(i + 5) & 15].y, acc[i].x); acc[i].y = fmaf(w[i & 3].y, acc[(i + 5) & 15].x, acc[i].y); }
        }
    }
    float s = 0.f;
    for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int iters = 20000;
    for (int mode = 0; mode < 3; ++mode)
        for (int threads : {128, 256, 512, 1024}) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<sms, threads>>>(out, iters, 1.0001f);
                if (mode == 1) k<1><<<sms, threads>>>(out, iters, 1.0001f);
                if (mode == 2) k<2><<<sms, threads>>>(out, iters, 1.0001f);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double fma = (double)iters * 16 * 2 * threads * sms;  // lane-FMAs
            printf("mode %d (%s) threads/SM %4d: %.3f ms  %.1f TFMA/s  %.1f FMA/clk/SM at %d MHz nominal\n", mode,
                   mode == 0 ? "FFMA2 distinct regs" : mode == 1 ? "FFMA2 shared a,b" : "scalar FFMA", threads, ms,
                   fma / ms / 1e9, fma / (ms * 1e-3) / sms / (clk_khz * 1e3), clk_khz / 1000);
        }
    return 0;
}
