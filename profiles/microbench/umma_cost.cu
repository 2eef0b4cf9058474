// Cost model of small tcgen05.mma (bf16, cta_group::1) on B200: cycles per MMA for back-to-back accumulating MMAs of
// various shapes / operand sources, one CTA per SM (and 4 CTAs per SM to see contention).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | (uint64_t)((lbo >> 4) & 0x3fff) << 16 | (uint64_t)((sbo >> 4) & 0x3fff) << 32 | (uint64_t)1 << 46;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int amn, int bmn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)amn << 15) | ((uint32_t)bmn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
constexpr int PLANE = 2048;
// mode: 0 K-major SS M128 N32 | 1 MN-major SS M128 N32 | 2 MN SS M64 N32 | 3 MN SS M128 N96 | 4 TS (A in TMEM) M128 N32 K-major B
//       5 K-major SS M128 N64 | 6 K-major SS M128 N128 | 7 MN SS M128 N16 | 8 same D for all vs (9) alternating two accumulators (mode 0 shape)
__global__ void __launch_bounds__(128) k(int mode, int n_mma, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 100 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t sa = smem_u32(smem), sb = sa + 48 * 1024;
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        // descriptors precomputed; the timed loop is 8 unrolled MMAs per iteration (issue overhead ~2 instructions per MMA)
        uint64_t ad[8], bd[8];
        uint32_t idesc = 0, dcol[8];
        for (int i = 0; i < 8; ++i) {
            const uint32_t off = i * 256;
            dcol[i] = tmem;
            switch (mode) {
                case 0: ad[i] = make_desc(sa + (i & 1) * 2 * PLANE, PLANE, 128); bd[i] = make_desc(sb, 512, 128); idesc = make_idesc(128, 32, 0, 0); break;
                case 1: ad[i] = make_desc(sa + off, 128, PLANE); bd[i] = make_desc(sb + off, 128, PLANE); idesc = make_idesc(128, 32, 1, 1); break;
                case 2: ad[i] = make_desc(sa + off, 128, PLANE); bd[i] = make_desc(sb + off, 128, PLANE); idesc = make_idesc(64, 32, 1, 1); break;
                case 3: ad[i] = make_desc(sa + off, 128, PLANE); bd[i] = make_desc(sb + off, 128, PLANE); idesc = make_idesc(128, 96, 1, 1); break;
                case 4: ad[i] = 0; bd[i] = make_desc(sb, 512, 128); idesc = make_idesc(128, 32, 0, 0); break;
                case 5: ad[i] = make_desc(sa + (i & 1) * 2 * PLANE, PLANE, 128); bd[i] = make_desc(sb, 1024, 128); idesc = make_idesc(128, 64, 0, 0); break;
                case 6: ad[i] = make_desc(sa + (i & 1) * 2 * PLANE, PLANE, 128); bd[i] = make_desc(sb, 2048, 128); idesc = make_idesc(128, 128, 0, 0); break;
                case 7: ad[i] = make_desc(sa + off, 128, PLANE); bd[i] = make_desc(sb + off, 128, PLANE); idesc = make_idesc(128, 16, 1, 1); break;
                case 8: ad[i] = make_desc(sa + (i & 1) * 2 * PLANE, PLANE, 128); bd[i] = make_desc(sb, 512, 128); idesc = make_idesc(128, 32, 0, 0); dcol[i] = tmem + (i & 1) * 32; break;
                case 9: ad[i] = make_desc(sa + off, 128, PLANE); bd[i] = make_desc(sb + off, 128, PLANE); idesc = make_idesc(64, 96, 1, 1); break;
            }
        }
        t0 = clock64();
        for (int it = 0; it < n_mma / 8; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (mode == 4) mma_ts(dcol[i], tmem + 64, bd[i], idesc, 1);
                else mma_ss(dcol[i], ad[i], bd[i], idesc, 1);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    if (tid == 0) { t1 = clock64(); out[blockIdx.x] = t1 - t0; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}
int main() {
    long long* d; static long long h[1024];
    CK(cudaMalloc(&d, sizeof(h)));
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    const char* names[] = {"K-major SS  M128 N32", "MN-major SS M128 N32", "MN-major SS M64  N32", "MN-major SS M128 N96", "TS (A in TMEM) M128 N32",
                           "K-major SS  M128 N64", "K-major SS  M128 N128", "MN-major SS M128 N16", "K-major SS M128 N32, two accumulators", "MN-major SS M64 N96"};
    for (int ctas : {1, 2}) {
        printf("--- %d CTA(s) per SM, cycles per MMA (n=256 back-to-back, one issuing thread per CTA) ---\n", ctas);
        for (int mode = 0; mode < 10; ++mode) {
            int grid = 148 * ctas;
            k<<<grid, 128, 100 * 1024>>>(mode, 256, d);
            CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
            k<<<grid, 128, 100 * 1024>>>(mode, 256, d);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h, d, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
            double s = 0; for (int i = 0; i < grid; ++i) s += h[i];
            printf("%-40s %7.1f cyc/MMA per CTA  -> %6.1f cyc/MMA of SM time\n", names[mode], s / grid / 256.0, s / grid / 256.0 / ctas);
        }
    }
    return 0;
}
