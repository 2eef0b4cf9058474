// Pipe micro-benchmarks that decide the liquid-cell kernel design (B200, sm_100a).
//   1. FFMA issue rate: register operands, smem-broadcast weights (LDS.128 : 4 FFMA, : 8 FFMA),
//      constant-bank weights over working sets of 1..16 KB.
//   2. MUFU rate: ex2.approx, rcp.approx.
//   3. legacy mma.sync m16n8k8 TF32 rate (for the 3xTF32 question).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
// Prints lane-ops/clk/SM for each case. Not part of the product.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int NCONST = 4096;                 // 16 KB of constants
__constant__ float c_w[NCONST];

// ---------------------------------------------------------------- FFMA, register operands
__global__ void k_ffma_reg(float* out, int iters, float a, float b) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// volatile so that the loop-invariant weight loads are not hoisted out of the time loop (and spilled)
__device__ __forceinline__ float4 lds128(const float4* p) {
    float4 v;
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}

// ---------------------------------------------------------------- FFMA, weights by LDS.128 broadcast
// NS samples per thread: one LDS.128 feeds 4*NS FFMA.
template <int NS>
__global__ void k_ffma_lds(float* out, int iters, const float* w_g, int nw) {
    extern __shared__ float4 sw[];
    for (int i = threadIdx.x; i < nw / 4; i += blockDim.x) sw[i] = reinterpret_cast<const float4*>(w_g)[i];
    __syncthreads();
    float acc[NS][32];
    float z[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        z[s] = threadIdx.x * 1e-3f + s;
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[s][i] = 0.f;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 20; ++k) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float4 w = lds128(&sw[k * 8 + q]);
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    acc[s][4 * q + 0] = fmaf(z[s], w.x, acc[s][4 * q + 0]);
                    acc[s][4 * q + 1] = fmaf(z[s], w.y, acc[s][4 * q + 1]);
                    acc[s][4 * q + 2] = fmaf(z[s], w.z, acc[s][4 * q + 2]);
                    acc[s][4 * q + 3] = fmaf(z[s], w.w, acc[s][4 * q + 3]);
                }
            }
#pragma unroll
            for (int s = 0; s < NS; ++s) z[s] = acc[s][k] * 0.5f;   // recurrence-like dependence
        }
    }
    float r = 0.f;
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int i = 0; i < 32; ++i) r += acc[s][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// ---------------------------------------------------------------- FFMA, weights from the constant bank
// NW = number of distinct constants walked per iteration (working set = 4*NW bytes).
template <int NW>
__global__ void k_ffma_const(float* out, int iters) {
    float acc[32];
    float z = threadIdx.x * 1e-3f;
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < NW / 32; ++k) {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = fmaf(z, c_w[k * 32 + i], acc[i]);
            z = acc[k & 31] * 0.5f;
        }
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) r += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// ---------------------------------------------------------------- MUFU
template <int OP>
__global__ void k_mufu(float* out, int iters) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.5f + threadIdx.x * 1e-4f + i * 0.01f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float y;
                if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v[i]));
                else if (OP == 1) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v[i]));
                else asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(v[i]));
                v[i] = y;
            }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------------------------------------------------------- mixed FFMA + MUFU (does MUFU steal issue slots only?)
__global__ void k_mix(float* out, int iters, float a, float b) {
    float acc[16];
    float v[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3f + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = 0.5f + i * 0.01f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
            // 2 MUFU per 16 FFMA (1:8, the liquid-cell forward ratio)
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(v[r & 3]) : "f"(v[r & 3]));
            asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(v[(r + 1) & 3]) : "f"(v[(r + 1) & 3]));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------------------------------------------------------- legacy mma.sync TF32 m16n8k8
__global__ void k_mma_tf32(float* out, int iters) {
    float c[8][4];
    uint32_t a[4], b[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + threadIdx.x * 1e-3f + i);
    b[0] = __float_as_uint(0.5f);
    b[1] = __float_as_uint(0.25f);
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) c[j][i] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int j = 0; j < 8; ++j)
                asm volatile(
                    "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                    : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                    : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) s += c[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m16n8k16 bf16
__global__ void k_mma_bf16(float* out, int iters) {
    float c[8][4];
    uint32_t a[4], b[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = 0x3f803f80u + threadIdx.x;
    b[0] = 0x3f003f00u;
    b[1] = 0x3e803e80u;
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) c[j][i] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int j = 0; j < 8; ++j)
                asm volatile(
                    "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                    : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                    : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) s += c[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch();
    launch();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("device %s, %d SMs, max clock %d MHz\n", p.name, sms, clk_khz / 1000);
    float* out;
    CK(cudaMalloc(&out, sizeof(float) * sms * 8 * 1024));
    float* w_g;
    CK(cudaMalloc(&w_g, sizeof(float) * NCONST));
    {
        static float h[NCONST];
        for (int i = 0; i < NCONST; ++i) h[i] = 1e-3f * (i % 17) - 4e-3f;
        CK(cudaMemcpy(w_g, h, sizeof(h), cudaMemcpyHostToDevice));
        CK(cudaMemcpyToSymbol(c_w, h, sizeof(h)));
    }
    // Rates are reported as lane-ops per SM per *max* clock; the real clock under load is usually lower,
    // so numbers slightly below the nominal (128 FFMA, 16 MUFU) are expected.
    const double hz = clk_khz * 1e3;
    auto report = [&](const char* name, double lane_ops, float ms) {
        double per_clk_sm = lane_ops / (ms * 1e-3) / hz / sms;
        printf("%-34s %8.3f ms  %8.2f lane-ops/clk/SM (at max clock)\n", name, ms, per_clk_sm);
    };
    for (int nt : {256, 512, 1024}) {
        int blocks = sms * (1024 / nt) * 2;
        int iters = 2000;
        char name[64];
        float ms = time_ms([&] { k_ffma_reg<<<blocks, nt>>>(out, iters, 0.999f, 1e-3f); });
        snprintf(name, 64, "ffma_reg nt=%d", nt);
        report(name, (double)blocks * nt * iters * 128, ms);
    }
    for (int nt : {128, 256, 384}) {
        int blocks = sms * 4;
        int iters = 400;
        char name[64];
        float ms = time_ms([&] { k_ffma_lds<1><<<blocks, nt, 4096>>>(out, iters, w_g, 640); });
        snprintf(name, 64, "ffma_lds 1:4 nt=%d", nt);
        report(name, (double)blocks * nt * iters * 640, ms);
        ms = time_ms([&] { k_ffma_lds<2><<<blocks, nt, 4096>>>(out, iters, w_g, 640); });
        snprintf(name, 64, "ffma_lds 1:8 nt=%d", nt);
        report(name, (double)blocks * nt * iters * 1280, ms);
    }
    {
        int nt = 256, blocks = sms * 8, iters = 200;
        float ms;
        ms = time_ms([&] { k_ffma_const<256><<<blocks, nt>>>(out, iters); });
        report("ffma_const 1KB", (double)blocks * nt * iters * 256, ms);
        ms = time_ms([&] { k_ffma_const<512><<<blocks, nt>>>(out, iters); });
        report("ffma_const 2KB", (double)blocks * nt * iters * 512, ms);
        ms = time_ms([&] { k_ffma_const<768><<<blocks, nt>>>(out, iters); });
        report("ffma_const 3KB", (double)blocks * nt * iters * 768, ms);
        ms = time_ms([&] { k_ffma_const<1024><<<blocks, nt>>>(out, iters); });
        report("ffma_const 4KB", (double)blocks * nt * iters * 1024, ms);
        ms = time_ms([&] { k_ffma_const<2048><<<blocks, nt>>>(out, iters); });
        report("ffma_const 8KB", (double)blocks * nt * iters * 2048, ms);
        ms = time_ms([&] { k_ffma_const<4096><<<blocks, nt>>>(out, iters); });
        report("ffma_const 16KB", (double)blocks * nt * iters * 4096, ms);
    }
    {
        int nt = 512, blocks = sms * 4, iters = 2000;
        float ms;
        ms = time_ms([&] { k_mufu<0><<<blocks, nt>>>(out, iters); });
        report("mufu ex2", (double)blocks * nt * iters * 64, ms);
        ms = time_ms([&] { k_mufu<1><<<blocks, nt>>>(out, iters); });
        report("mufu rcp", (double)blocks * nt * iters * 64, ms);
        ms = time_ms([&] { k_mufu<2><<<blocks, nt>>>(out, iters); });
        report("mufu tanh", (double)blocks * nt * iters * 64, ms);
        ms = time_ms([&] { k_mix<<<blocks, nt>>>(out, iters, 0.999f, 1e-3f); });
        report("mix ffma(128)+mufu(16) [ffma]", (double)blocks * nt * iters * 128, ms);
    }
    {
        int nt = 256, blocks = sms * 8, iters = 1000;
        float ms = time_ms([&] { k_mma_tf32<<<blocks, nt>>>(out, iters); });
        // per warp-instruction: 16*8*8 = 1024 MAC
        double macs = (double)blocks * (nt / 32) * iters * 32 * 1024;
        printf("%-34s %8.3f ms  %8.2f MAC/clk/SM (at max clock)\n", "mma.sync m16n8k8 tf32", ms, macs / (ms * 1e-3) / hz / sms);
        ms = time_ms([&] { k_mma_bf16<<<blocks, nt>>>(out, iters); });
        macs = (double)blocks * (nt / 32) * iters * 32 * 2048;
        printf("%-34s %8.3f ms  %8.2f MAC/clk/SM (at max clock)\n", "mma.sync m16n8k16 bf16", ms, macs / (ms * 1e-3) / hz / sms);
    }
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
