// Unit test of the tcgen05 building blocks used by the tensor-core liquid kernels (B200, sm_100a).
//   1. K-major, no-swizzle TF32 MMA:  D[128 x 32] = A[128 x 24] * B[32 x 24]^T   (heads of one step for 128 samples)
//      A lives in "planes": plane c holds floats 4c..4c+3 of every row, 16 B per row, rows contiguous (2 KB/plane).
//   2. The SAME planes read MN-major:   D3[64 x 24] = sum_rows P[row, 0:64]^T * Q[row, 0:24]   (weight-gradient form:
//      reduction over the 128 rows = samples).  M=64 accumulator layout: row m -> lane (m%16) + 32*(m/16).
// Inputs are small multiples of 2^-2, exactly representable in TF32, so results must match the CPU exactly.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_test.bin umma_test.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// no-swizzle shared-memory matrix descriptor: start>>4 | LBO>>4 << 16 | SBO>>4 << 32 | version(1) << 46
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor: c=f32 (1<<4), a=b=tf32 (2<<7, 2<<10), majors, N>>3 << 17, M>>4 << 24
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}

constexpr int ROWS = 128;
constexpr int PLANE = ROWS * 16;          // bytes per plane
constexpr int KA = 24, NB = 32;           // test 1
constexpr int PW = 128, QW = 24;           // test 2: P has 64 floats per row (16 planes), Q has 24 (6 planes)

__global__ void __launch_bounds__(128) umma_test_kernel(const float* A, const float* B, const float* P, const float* Q,
                                                        float* D1, float* D3, float* DUMP) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sA = reinterpret_cast<float*>(smem);                       // KA/4 planes
    float* sB = sA + (KA / 4) * ROWS * 4;                             // K-major B: [k/4][n][4]  (NB rows)
    float* sP = sB + (KA / 4) * NB * 4;
    float* sQ = sP + (PW / 4) * ROWS * 4;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;

    // row `tid` of A / P / Q into the planes: 16 B per (plane, row)
    for (int c = 0; c < KA / 4; ++c) reinterpret_cast<float4*>(sA)[c * ROWS + tid] = reinterpret_cast<const float4*>(A)[tid * (KA / 4) + c];
    for (int c = 0; c < PW / 4; ++c) reinterpret_cast<float4*>(sP)[c * ROWS + tid] = reinterpret_cast<const float4*>(P)[tid * (PW / 4) + c];
    for (int c = 0; c < QW / 4; ++c) reinterpret_cast<float4*>(sQ)[c * ROWS + tid] = reinterpret_cast<const float4*>(Q)[tid * (QW / 4) + c];
    if (tid < NB)
        for (int c = 0; c < KA / 4; ++c) reinterpret_cast<float4*>(sB)[c * NB + tid] = reinterpret_cast<const float4*>(B)[tid * (KA / 4) + c];

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy smem writes -> async proxy (MMA)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    if (tid == 0) {
        // ---- test 1: K-major A (rows = M), K-major B (rows = N); one MMA per 8 k (= 2 planes)
        const uint32_t idesc1 = make_idesc(128, NB, 0, 0);
        for (int ks = 0; ks < KA / 8; ++ks) {
            uint64_t ad = make_desc(smem_u32(sA) + ks * 2 * PLANE, PLANE, 128);          // LBO = next k-plane, SBO = next 8 rows
            uint64_t bd = make_desc(smem_u32(sB) + ks * 2 * NB * 16, NB * 16, 128);
            mma_tf32(tmem + 0, ad, bd, idesc1, ks > 0);
        }
        // ---- test 2: MN-major A = planes of P (M = 64 floats of each row), MN-major B = planes of Q (N = 24);
        //      K = rows: one MMA per 8 rows.  MN-major canonical: 16 B along MN, 8 k-rows at 16 B stride,
        //      LBO = next 8 k-rows (128 B), SBO = next 4 MN elements (one plane).
        const uint32_t idesc2 = make_idesc(64, QW, 1, 1);
        for (int ks = 0; ks < ROWS / 8; ++ks) {
            uint64_t ad = make_desc(smem_u32(sP) + ks * 128, 128, PLANE);
            uint64_t bd = make_desc(smem_u32(sQ) + ks * 128, 128, PLANE);
            mma_tf32(tmem + 32, ad, bd, idesc2, ks > 0);
        }
        // ---- variants for isolation
        {   // T2a: M=64, K-major A (rows 0..63 of A), K-major B -> cols 64..95
            const uint32_t id = make_idesc(64, NB, 0, 0);
            for (int ks = 0; ks < KA / 8; ++ks)
                mma_tf32(tmem + 64, make_desc(smem_u32(sA) + ks * 2 * PLANE, PLANE, 128), make_desc(smem_u32(sB) + ks * 2 * NB * 16, NB * 16, 128), id, ks > 0);
        }
        {   // T2b: M=128 (all 128 floats of P), N=32 (Q has 24 -> reads 2 planes past Q: garbage cols 24..31 ignored), MN-major both -> cols 96..127
            const uint32_t id = make_idesc(128, 32, 1, 1);
            for (int ks = 0; ks < ROWS / 8; ++ks)
                mma_tf32(tmem + 96, make_desc(smem_u32(sP) + ks * 128, 128, PLANE), make_desc(smem_u32(sQ) + ks * 128, 128, PLANE), id, ks > 0);
        }
        {   // T2c: M=128, A MN-major (P), B K-major: B2[n][k=row]: use sA planes as B?  B K-major needs [k/4][n][4]: sA is [k/4][row][4] with
            // row as the N index: D[m][n] = sum_k P^T... not meaningful; instead A K-major (sA, K=24) x B MN-major: B[n][k]: Q^T? skip
        }
        {   // T2d: swapped offsets (LBO = plane, SBO = 128) for MN-major, M=64, N=24 -> cols 128..151
            const uint32_t id = make_idesc(64, QW, 1, 1);
            for (int ks = 0; ks < ROWS / 8; ++ks)
                mma_tf32(tmem + 128, make_desc(smem_u32(sP) + ks * 128, PLANE, 128), make_desc(smem_u32(sQ) + ks * 128, PLANE, 128), id, ks > 0);
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // D1: lane = row; 32 columns
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
          "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
          "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int n = 0; n < 32; ++n) D1[tid * 32 + n] = __uint_as_float(v[n]);

    // D3 (M = 64): row m lives in lane (m % 16) + 32 * (m / 16); 24 columns starting at column 32
    uint32_t w[8];
    for (int c0 = 0; c0 < QW; c0 += 8) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                     : "r"(taddr + 32 + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int lane = tid & 31;
        if (lane < 16)
            for (int n = 0; n < 8; ++n) D3[(warp * 16 + lane) * QW + c0 + n] = __uint_as_float(w[n]);
    }
    for (int c0 = 0; c0 < 128; c0 += 8) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                     : "r"(taddr + 32 + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int n = 0; n < 8; ++n) DUMP[tid * 128 + c0 + n] = __uint_as_float(w[n]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

int main() {
    static float A[ROWS * KA], B[NB * KA], P[ROWS * PW], Q[ROWS * QW], D1[ROWS * NB], D3[64 * QW];
    for (int r = 0; r < ROWS; ++r) {
        for (int k = 0; k < KA; ++k) A[r * KA + k] = ((r * 7 + k * 3) % 11 - 5) * 0.25f;
        for (int k = 0; k < PW; ++k) P[r * PW + k] = ((r * 3 + k * 5) % 13 - 6) * 0.25f;
        for (int k = 0; k < QW; ++k) Q[r * QW + k] = ((r * 5 + k * 7) % 9 - 4) * 0.5f;
    }
    for (int n = 0; n < NB; ++n)
        for (int k = 0; k < KA; ++k) B[n * KA + k] = ((n * 5 + k) % 7 - 3) * 0.5f;
    float *dA, *dB, *dP, *dQ, *dD1, *dD3, *dDump; static float DUMP[128*128]; CK(cudaMalloc(&dDump, sizeof(DUMP)));
    CK(cudaMalloc(&dA, sizeof(A))); CK(cudaMalloc(&dB, sizeof(B))); CK(cudaMalloc(&dP, sizeof(P))); CK(cudaMalloc(&dQ, sizeof(Q)));
    CK(cudaMalloc(&dD1, sizeof(D1))); CK(cudaMalloc(&dD3, sizeof(D3)));
    CK(cudaMemcpy(dA, A, sizeof(A), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B, sizeof(B), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dP, P, sizeof(P), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dQ, Q, sizeof(Q), cudaMemcpyHostToDevice));
    CK(cudaMemset(dD1, 0xff, sizeof(D1))); CK(cudaMemset(dD3, 0xff, sizeof(D3)));
    size_t smem = ((KA / 4) * ROWS + (KA / 4) * NB + (PW / 4) * ROWS + (QW / 4) * ROWS) * 16 + 1024;
    CK(cudaFuncSetAttribute(umma_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_test_kernel<<<1, 128, smem>>>(dA, dB, dP, dQ, dD1, dD3, dDump);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D1, dD1, sizeof(D1), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(D3, dD3, sizeof(D3), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(DUMP, dDump, sizeof(DUMP), cudaMemcpyDeviceToHost));
    {
        auto cnt = [&](int c_lo, int c_hi) { int n = 0; for (int l = 0; l < 128; ++l) for (int c = c_lo; c < c_hi; ++c) n += DUMP[l * 128 + c] != 0.f; return n; };
        printf("nonzeros: T2(cols32-63) %d, T2a(64-95) %d, T2b(96-127) %d, T2d(128-159) %d\n", cnt(0, 32), cnt(32, 64), cnt(64, 96), cnt(96, 128));
        // T2a check: rows 0..63 of D1 reference at lanes (m%16)+32*(m/16)
        int bad = 0;
        for (int m = 0; m < 64; ++m) for (int n = 0; n < NB; ++n) { float ref = 0.f; for (int k = 0; k < KA; ++k) ref += A[m * KA + k] * B[n * KA + k]; int lane = (m % 16) + 32 * (m / 16); bad += DUMP[lane * 128 + 32 + n] != ref; }
        printf("T2a (M=64 K-major, 16-lane layout): %d mismatches\n", bad);
        bad = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < QW; ++n) { float ref = 0.f; for (int r = 0; r < ROWS; ++r) ref += P[r * PW + m] * Q[r * QW + n]; bad += DUMP[m * 128 + 64 + n] != ref; }
        printf("T2b (M=128 N=32 MN-major both): %d mismatches\n", bad);
        bad = 0;
        for (int m = 0; m < 64; ++m) for (int n = 0; n < QW; ++n) { float ref = 0.f; for (int r = 0; r < ROWS; ++r) ref += P[r * PW + m] * Q[r * QW + n]; int lane = (m % 16) + 32 * (m / 16); bad += DUMP[lane * 128 + 96 + n] != ref; }
        printf("T2d (M=64 N=24 MN-major, swapped LBO/SBO): %d mismatches\n", bad);
    }
    int bad1 = 0, bad3 = 0;
    for (int r = 0; r < ROWS; ++r)
        for (int n = 0; n < NB; ++n) {
            float ref = 0.f;
            for (int k = 0; k < KA; ++k) ref += A[r * KA + k] * B[n * KA + k];
            if (D1[r * NB + n] != ref) { if (bad1 < 5) printf("D1[%d][%d] = %g, want %g\n", r, n, D1[r * NB + n], ref); ++bad1; }
        }
    for (int m = 0; m < 64; ++m)
        for (int n = 0; n < QW; ++n) {
            float ref = 0.f;
            for (int r = 0; r < ROWS; ++r) ref += P[r * PW + m] * Q[r * QW + n];
            if (D3[m * QW + n] != ref) { if (bad3 < 5) printf("D3[%d][%d] = %g, want %g\n", m, n, D3[m * QW + n], ref); ++bad3; }
        }
    printf("test1 (K-major 128x32x24): %s (%d mismatches)\n", bad1 ? "FAIL" : "PASS", bad1);
    printf("test2 (MN-major 64x24x128, M=64 lanes): %s (%d mismatches)\n", bad3 ? "FAIL" : "PASS", bad3);
    return (bad1 || bad3) ? 1 : 0;
}
