// bf16 variant of umma_test.cu: does ONE shared-memory copy of per-row vectors, stored as 16-byte (8 x bf16) chunks in
// "planes" [chunk][row][16 B], serve BOTH as a K-major operand (rows = M, features = K) and as an MN-major operand
// (features = M/N, rows = K)?  No swizzle in both cases.
//   T1: D1[128 x 32]  = A[128 x 32] * B[32 x 32]^T            (K-major A planes, K-major B planes; K = 32 = 2 MMAs)
//   T2: D2[128 x 32]  = sum_rows P[row, 0:128]^T * Q[row, 0:32]  (MN-major both, M = 128, K = 128 rows = 8 MMAs)
//   T3: D3[64 x 24]   = same with M = 64, N = 24
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16 | (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32 | (uint64_t)1 << 46;
}
// c = f32, a = b = bf16 (1)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
constexpr int ROWS = 128, PLANE = ROWS * 16;
constexpr int KA = 32, NB = 32, PW = 128, QW = 32;   // widths in bf16 elements (multiples of 8)

__global__ void __launch_bounds__(128) k(const __nv_bfloat16* A, const __nv_bfloat16* B, const __nv_bfloat16* P, const __nv_bfloat16* Q, float* DUMP) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint4* sA = reinterpret_cast<uint4*>(smem);            // KA/8 planes of ROWS x 16 B
    uint4* sB = sA + (KA / 8) * ROWS;                      // KA/8 planes of NB x 16 B
    uint4* sP = sB + (KA / 8) * NB;
    uint4* sQ = sP + (PW / 8) * ROWS;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int c = 0; c < KA / 8; ++c) sA[c * ROWS + tid] = reinterpret_cast<const uint4*>(A)[tid * (KA / 8) + c];
    for (int c = 0; c < PW / 8; ++c) sP[c * ROWS + tid] = reinterpret_cast<const uint4*>(P)[tid * (PW / 8) + c];
    for (int c = 0; c < QW / 8; ++c) sQ[c * ROWS + tid] = reinterpret_cast<const uint4*>(Q)[tid * (QW / 8) + c];
    if (tid < NB) for (int c = 0; c < KA / 8; ++c) sB[c * NB + tid] = reinterpret_cast<const uint4*>(B)[tid * (KA / 8) + c];
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
    if (tid == 0) {
        for (int ks = 0; ks < KA / 16; ++ks)   // T1 -> cols 0..31
            mma_bf16(tmem + 0, make_desc(smem_u32(sA) + ks * 2 * PLANE, PLANE, 128), make_desc(smem_u32(sB) + ks * 2 * NB * 16, NB * 16, 128), make_idesc(128, NB, 0, 0), ks > 0);
        for (int ks = 0; ks < ROWS / 16; ++ks) // T2 -> cols 32..63: MN-major, K = 16 rows per MMA = 2 groups of 8 rows (LBO = 128 B), SBO = next 8 MN elements = plane
            mma_bf16(tmem + 32, make_desc(smem_u32(sP) + ks * 256, 128, PLANE), make_desc(smem_u32(sQ) + ks * 256, 128, PLANE), make_idesc(128, 32, 1, 1), ks > 0);
        for (int ks = 0; ks < ROWS / 16; ++ks) // T3 -> cols 64..87: M = 64, N = 24
            mma_bf16(tmem + 64, make_desc(smem_u32(sP) + ks * 256, 128, PLANE), make_desc(smem_u32(sQ) + ks * 256, 128, PLANE), make_idesc(64, 24, 1, 1), ks > 0);
        for (int ks = 0; ks < ROWS / 16; ++ks) // T4 -> cols 96..127: swapped LBO/SBO
            mma_bf16(tmem + 96, make_desc(smem_u32(sP) + ks * 256, PLANE, 128), make_desc(smem_u32(sQ) + ks * 256, PLANE, 128), make_idesc(128, 32, 1, 1), ks > 0);
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t w[8];
    for (int c0 = 0; c0 < 128; c0 += 8) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "r"(taddr + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int n = 0; n < 8; ++n) DUMP[tid * 128 + c0 + n] = __uint_as_float(w[n]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

int main() {
    static float A[ROWS * KA], B[NB * KA], P[ROWS * PW], Q[ROWS * QW], DUMP[128 * 128];
    static __nv_bfloat16 hA[ROWS * KA], hB[NB * KA], hP[ROWS * PW], hQ[ROWS * QW];
    for (int r = 0; r < ROWS; ++r) {
        for (int c = 0; c < KA; ++c) A[r * KA + c] = ((r * 7 + c * 3) % 11 - 5) * 0.25f;
        for (int c = 0; c < PW; ++c) P[r * PW + c] = ((r * 3 + c * 5) % 13 - 6) * 0.25f;
        for (int c = 0; c < QW; ++c) Q[r * QW + c] = ((r * 5 + c * 7) % 9 - 4) * 0.5f;
    }
    for (int n = 0; n < NB; ++n) for (int c = 0; c < KA; ++c) B[n * KA + c] = ((n * 5 + c) % 7 - 3) * 0.5f;
    for (int i = 0; i < ROWS * KA; ++i) hA[i] = __float2bfloat16(A[i]);
    for (int i = 0; i < NB * KA; ++i) hB[i] = __float2bfloat16(B[i]);
    for (int i = 0; i < ROWS * PW; ++i) hP[i] = __float2bfloat16(P[i]);
    for (int i = 0; i < ROWS * QW; ++i) hQ[i] = __float2bfloat16(Q[i]);
    __nv_bfloat16 *dA, *dB, *dP, *dQ; float* dD;
    CK(cudaMalloc(&dA, sizeof(hA))); CK(cudaMalloc(&dB, sizeof(hB))); CK(cudaMalloc(&dP, sizeof(hP))); CK(cudaMalloc(&dQ, sizeof(hQ))); CK(cudaMalloc(&dD, sizeof(DUMP)));
    CK(cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dP, hP, sizeof(hP), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dQ, hQ, sizeof(hQ), cudaMemcpyHostToDevice));
    size_t smem = ((KA / 8) * ROWS + (KA / 8) * NB + (PW / 8) * ROWS + (QW / 8) * ROWS) * 16 + 4096;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<1, 128, smem>>>(dA, dB, dP, dQ, dD);
    CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(DUMP, dD, sizeof(DUMP), cudaMemcpyDeviceToHost));
    int b1 = 0, b2 = 0, b3 = 0, b4 = 0;
    for (int r = 0; r < ROWS; ++r) for (int n = 0; n < NB; ++n) { float ref = 0; for (int c = 0; c < KA; ++c) ref += A[r * KA + c] * B[n * KA + c]; b1 += DUMP[r * 128 + n] != ref; }
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) { float ref = 0; for (int r = 0; r < ROWS; ++r) ref += P[r * PW + m] * Q[r * QW + n]; b2 += DUMP[m * 128 + 32 + n] != ref; b4 += DUMP[m * 128 + 96 + n] != ref; }
    for (int m = 0; m < 64; ++m) for (int n = 0; n < 24; ++n) { float ref = 0; for (int r = 0; r < ROWS; ++r) ref += P[r * PW + m] * Q[r * QW + n]; int lane = (m % 16) + 32 * (m / 16); b3 += DUMP[lane * 128 + 64 + n] != ref; }
    printf("T1 K-major bf16 128x32x32: %d mismatches\nT2 MN-major bf16 M=128 N=32 K=128: %d mismatches\nT3 MN-major bf16 M=64 N=24: %d mismatches\nT4 (swapped LBO/SBO): %d mismatches\n", b1, b2, b3, b4);
    return 0;
}
