// FFMA vs FFMA2 (fma.rn.f32x2, sm_100) issue / pipe throughput: 8 independent chains per thread, 1024 threads per SM.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_ffma(float* out, float a, float b, int iters) {
    float x[16];
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
    float s = 0;
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, float a, float b, int iters) {
    unsigned long long x[8], aa, bb;
    float2 t = make_float2(a, a), u = make_float2(b, b);
    aa = *reinterpret_cast<unsigned long long*>(&t);
    bb = *reinterpret_cast<unsigned long long*>(&u);
    for (int i = 0; i < 8; ++i) { float2 v = make_float2(threadIdx.x + i, threadIdx.x - i); x[i] = *reinterpret_cast<unsigned long long*>(&v); }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(aa), "l"(bb));
    float s = 0;
    for (int i = 0; i < 8; ++i) { float2 v = *reinterpret_cast<float2*>(&x[i]); s += v.x + v.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: FFMA2 + MUFU to see whether the packed op frees issue slots
__global__ void k_mix(float* out, float a, float b, int iters) {
    unsigned long long x[8], aa, bb;
    float2 t = make_float2(a, a), u = make_float2(b, b);
    aa = *reinterpret_cast<unsigned long long*>(&t);
    bb = *reinterpret_cast<unsigned long long*>(&u);
    float m[4] = {1.f, 2.f, 3.f, 4.f};
    for (int i = 0; i < 8; ++i) { float2 v = make_float2(threadIdx.x + i, threadIdx.x - i); x[i] = *reinterpret_cast<unsigned long long*>(&v); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(aa), "l"(bb));
#pragma unroll
        for (int i = 0; i < 2; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[i]));
    }
    float s = m[0] + m[1];
    for (int i = 0; i < 8; ++i) { float2 v = *reinterpret_cast<float2*>(&x[i]); s += v.x + v.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out;
    cudaMalloc(&out, 148 * 2 * 1024 * 4);
    int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int which = 0; which < 3; ++which) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (which == 0) k_ffma<<<148 * 2, 1024>>>(out, 1.0001f, 0.5f, iters);
            else if (which == 1) k_ffma2<<<148 * 2, 1024>>>(out, 1.0001f, 0.5f, iters);
            else k_mix<<<148 * 2, 1024>>>(out, 1.0001f, 0.5f, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fma = 148.0 * 2 * 1024 * 16.0 * iters;           // scalar FMAs
        double clk = ms * 1e-3 * 1.965e9;
        printf("%s: %.3f ms, %.1f scalar FMA / clk / SM\n", which == 0 ? "FFMA" : which == 1 ? "FFMA2" : "FFMA2+2 MUFU per 8", ms, fma / clk / 148.0);
    }
    return 0;
}
