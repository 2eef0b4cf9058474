// Does tcgen05.mma accumulate into the fp32 TMEM accumulator with round-to-nearest or truncation?
// MMA #0 writes 1.0 (or -1.0) to D[0][0]; each of the next n MMAs adds +-0.75 * 2^-23 (0.75 ulp of 1.0).
// RN: every add rounds up to the next float (1 + n * 2^-23).  RZ: the accumulator never moves.
// Also: within ONE MMA (K = 16), 16 such addends = 12 * 2^-23 exactly representable -> checks internal precision.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | (uint64_t)((lbo >> 4) & 0x3fff) << 16 | (uint64_t)((sbo >> 4) & 0x3fff) << 32 | (uint64_t)1 << 46;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void mma_bf16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
constexpr int ROWS = 128, PLANE = ROWS * 16;
// A0: row r, k=0 -> sign; A1: row r: k=0 -> 0.75*sign (test a: one addend per MMA), A2: all 16 k -> 0.75*sign (test b)
// B0: n=0: k=0 -> 1.0;  B1: n=0: all k -> 2^-23
__global__ void __launch_bounds__(128) k(float* out, int n_adds) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __nv_bfloat16* A0 = reinterpret_cast<__nv_bfloat16*>(smem);          // 2 planes each
    __nv_bfloat16* A1 = A0 + 2 * ROWS * 8;
    __nv_bfloat16* A2 = A1 + 2 * ROWS * 8;
    __nv_bfloat16* B0 = A2 + 2 * ROWS * 8;                               // 16 rows (N=16), 2 planes
    __nv_bfloat16* B1 = B0 + 2 * 16 * 8;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const float sign = (tid & 1) ? -1.f : 1.f;    // odd rows test the negative side
    for (int p = 0; p < 2; ++p)
        for (int e = 0; e < 8; ++e) {
            int kk = p * 8 + e;
            A0[(p * ROWS + tid) * 8 + e] = __float2bfloat16(kk == 0 ? sign : 0.f);
            A1[(p * ROWS + tid) * 8 + e] = __float2bfloat16(kk == 0 ? 0.75f * sign : 0.f);
            A2[(p * ROWS + tid) * 8 + e] = __float2bfloat16(0.75f * sign);
            if (tid < 16) {
                B0[(p * 16 + tid) * 8 + e] = __float2bfloat16((tid == 0 && kk == 0) ? 1.f : 0.f);
                B1[(p * 16 + tid) * 8 + e] = __float2bfloat16(tid == 0 ? 1.1920928955078125e-07f : 0.f);   // 2^-23
            }
        }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t id = make_idesc(128, 16);
    if (tid == 0) {
        auto dA = [&](const void* p) { return make_desc(smem_u32(p), PLANE, 128); };
        auto dB = [&](const void* p) { return make_desc(smem_u32(p), 16 * 16, 128); };
        // test a -> col 0..15: 1.0 then n_adds single addends
        mma_bf16(tmem + 0, dA(A0), dB(B0), id, 0);
        for (int i = 0; i < n_adds; ++i) mma_bf16(tmem + 0, dA(A1), dB(B1), id, 1);
        // test b -> col 16..31: 1.0 then ONE MMA with 16 addends (sum = 12 * 2^-23)
        mma_bf16(tmem + 16, dA(A0), dB(B0), id, 0);
        mma_bf16(tmem + 16, dA(A2), dB(B1), id, 1);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t a, b;
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(a) : "r"(taddr + 0));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(b) : "r"(taddr + 16));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    out[tid * 2] = __uint_as_float(a);
    out[tid * 2 + 1] = __uint_as_float(b);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}
int main() {
    float* d; static float h[256];
    CK(cudaMalloc(&d, sizeof(h)));
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    for (int n : {1, 8}) {
        k<<<1, 128, 32768>>>(d, n);
        CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
        const double u = ldexp(1.0, -23);
        printf("n_adds=%d  row0 (+): across-MMA (acc-1)/ulp = %.3f   [RN -> %d, RZ -> 0, exact %.2f]\n", n, (h[0] - 1.0) / u, n, 0.75 * n);
        printf("          row1 (-): across-MMA (acc+1)/ulp = %.3f\n", (h[2] + 1.0) / u);
        printf("          one MMA with 16 addends (exact 12 ulp): row0 %.3f, row1 %.3f\n", (h[1] - 1.0) / u, (h[3] + 1.0) / u);
    }
    return 0;
}
