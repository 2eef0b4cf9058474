// Tile routines shared by the generic (any-size) liquid and NCP kernels.
//
// A CTA of kTileThreads threads owns BM batch rows ("samples") at a time.  Activations live in shared memory as
// [feature][BM + 4] (sample fastest; the +4 pad keeps 128-bit row reads conflict-free).  Thread t works on sample
// s = t % BM for output blocks g = t / BM, g + G, ...   Weights are read through the read-only path with
// warp-uniform 128-bit loads (one L1 transaction per warp); when the packed block fits it has been staged into
// shared memory with one bulk (TMA) copy and the same code reads it from there.
#pragma once
#include "common.cuh"

namespace vb {
namespace tile {

constexpr int kTileThreads = 256;

template <int BM>
struct Geo {
    static constexpr int S = BM + 4;                 // row stride in floats
    static constexpr int RS = BM >= 64 ? 4 : (BM == 32 ? 2 : 1);   // samples per thread in dense(): RS x 8 register blocks
    static constexpr int LN = BM / RS;               // threads along the sample axis
    static constexpr int G = kTileThreads / LN;      // output groups
};

// out[j][s] = epi(bias[j] + sum_i Wt[i][j] * in[i][s]),  j < n_out (Wt rows are ldw floats, ldw % 4 == 0, zero padded)
template <int BM, class Epi>
__device__ __forceinline__ void dense(const float* __restrict__ in, int n_in, const float* __restrict__ Wt, int ldw,
                                      const float* __restrict__ bias, int n_out, Epi epi) {
    constexpr int S = Geo<BM>::S, G = Geo<BM>::G, RS = Geo<BM>::RS, LN = Geo<BM>::LN;
    const int sl = threadIdx.x % LN, g = threadIdx.x / LN;
    const int n_blk = (n_out + 3) >> 2;
    if constexpr (RS > 1) {
        // RS samples x 8 outputs per thread: one vector activation read + two warp-uniform 128-bit weight reads feed 8 RS FMAs
        // (the 1 x 8 variant below issues 3 shared-memory reads per 8 FMAs and is bound by the LSU, not the FMA pipe)
        const float* ip = in + RS * sl;
        for (int jb = 2 * g; jb < n_blk; jb += 2 * G) {
            const bool two = (jb + 1) < n_blk;
            const float4 b0 = bias ? *reinterpret_cast<const float4*>(bias + 4 * jb) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 b1 = (bias && two) ? *reinterpret_cast<const float4*>(bias + 4 * jb + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            float acc[RS][8];
#pragma unroll
            for (int r = 0; r < RS; ++r) {
                acc[r][0] = b0.x; acc[r][1] = b0.y; acc[r][2] = b0.z; acc[r][3] = b0.w;
                acc[r][4] = b1.x; acc[r][5] = b1.y; acc[r][6] = b1.z; acc[r][7] = b1.w;
            }
            const float* w = Wt + 4 * jb;
#pragma unroll 2
            for (int i = 0; i < n_in; ++i) {
                float v[RS];
                if constexpr (RS == 4) {
                    const float4 v4 = *reinterpret_cast<const float4*>(ip + i * S);
                    v[0] = v4.x; v[1] = v4.y; v[2] = v4.z; v[3] = v4.w;
                } else {
                    const float2 v2 = *reinterpret_cast<const float2*>(ip + i * S);
                    v[0] = v2.x; v[1] = v2.y;
                }
                const float4 w0 = *reinterpret_cast<const float4*>(w + (int64_t)i * ldw);
                const float4 w1 = two ? *reinterpret_cast<const float4*>(w + (int64_t)i * ldw + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int r = 0; r < RS; ++r) {
                    acc[r][0] = fmaf(v[r], w0.x, acc[r][0]); acc[r][1] = fmaf(v[r], w0.y, acc[r][1]);
                    acc[r][2] = fmaf(v[r], w0.z, acc[r][2]); acc[r][3] = fmaf(v[r], w0.w, acc[r][3]);
                    acc[r][4] = fmaf(v[r], w1.x, acc[r][4]); acc[r][5] = fmaf(v[r], w1.y, acc[r][5]);
                    acc[r][6] = fmaf(v[r], w1.z, acc[r][6]); acc[r][7] = fmaf(v[r], w1.w, acc[r][7]);
                }
            }
            const int j = 4 * jb;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (j + k < n_out && (k < 4 || two)) {
#pragma unroll
                    for (int r = 0; r < RS; ++r) epi(j + k, RS * sl + r, acc[r][k]);
                }
        }
        return;
    }
    const int s = sl;
    for (int jb = 2 * g; jb < n_blk; jb += 2 * G) {
        const bool two = (jb + 1) < n_blk;
        float4 a0 = bias ? *reinterpret_cast<const float4*>(bias + 4 * jb) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 a1 = (bias && two) ? *reinterpret_cast<const float4*>(bias + 4 * jb + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float* w = Wt + 4 * jb;
        const float* ip = in + s;
#pragma unroll 4
        for (int i = 0; i < n_in; ++i) {
            float v = ip[i * S];
            float4 w0 = *reinterpret_cast<const float4*>(w + (int64_t)i * ldw);
            a0.x = fmaf(v, w0.x, a0.x); a0.y = fmaf(v, w0.y, a0.y); a0.z = fmaf(v, w0.z, a0.z); a0.w = fmaf(v, w0.w, a0.w);
            if (two) {
                float4 w1 = *reinterpret_cast<const float4*>(w + (int64_t)i * ldw + 4);
                a1.x = fmaf(v, w1.x, a1.x); a1.y = fmaf(v, w1.y, a1.y); a1.z = fmaf(v, w1.z, a1.z); a1.w = fmaf(v, w1.w, a1.w);
            }
        }
        const int j = 4 * jb;
        if (j + 0 < n_out) epi(j + 0, s, a0.x);
        if (j + 1 < n_out) epi(j + 1, s, a0.y);
        if (j + 2 < n_out) epi(j + 2, s, a0.z);
        if (j + 3 < n_out) epi(j + 3, s, a0.w);
        if (two) {
            if (j + 4 < n_out) epi(j + 4, s, a1.x);
            if (j + 5 < n_out) epi(j + 5, s, a1.y);
            if (j + 6 < n_out) epi(j + 6, s, a1.z);
            if (j + 7 < n_out) epi(j + 7, s, a1.w);
        }
    }
}

template <bool ATOMIC>
__device__ __forceinline__ void acc_add(float* p, float v) {
    if (ATOMIC) atomicAdd(p, v);
    else *p += v;   // the entry is owned by exactly one thread of the CTA
}

// dWt[i][j] += sum_s in[i][s] * dout[j][s]   (i < n_in, j < n_out; dWt rows are ldw floats)
// db[j]     += sum_s dout[j][s]              (if db != nullptr)
template <int BM, bool ATOMIC>
__device__ __forceinline__ void wgrad(float* dWt, int ldw, float* db, const float* __restrict__ in, int n_in,
                                      const float* __restrict__ dout, int n_out) {
    constexpr int S = Geo<BM>::S;
    // dWt[i][j] += sum_s in[i][s] * dout[j][s].  One thread owns 2 inputs x 4 outputs: six 128-bit reads feed 32 FMAs per four
    // samples.  The four outputs of a thread are J = ceil(n_out / 4) apart and consecutive threads own consecutive outputs, so
    // that a warp's reads of dout rows (row stride = 4 mod 32 banks) and its read-modify-writes of dWt are conflict-free
    // (with 4 adjacent outputs per thread ncu showed 17 wavefronts per LDS.128 instead of 4).
    const int J = (n_out + 3) >> 2, n_ip = (n_in + 1) >> 1;
    const int total = n_ip * J;
    for (int e = threadIdx.x; e < total; e += kTileThreads) {
        const int ipair = e / J, j0 = e - ipair * J;
        const int i0 = 2 * ipair, i1 = min(i0 + 1, n_in - 1);
        const float* ip0 = in + i0 * S;
        const float* ip1 = in + i1 * S;
        const float* dp[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) dp[k] = dout + min(j0 + k * J, n_out - 1) * S;
        float a0[4], a1[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { a0[k] = 0.f; a1[k] = 0.f; }
#pragma unroll 4
        for (int q = 0; q < BM / 4; ++q) {
            const float4 v0 = *reinterpret_cast<const float4*>(ip0 + 4 * q);
            const float4 v1 = *reinterpret_cast<const float4*>(ip1 + 4 * q);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float4 d = *reinterpret_cast<const float4*>(dp[k] + 4 * q);
                a0[k] = fmaf(v0.x, d.x, fmaf(v0.y, d.y, fmaf(v0.z, d.z, fmaf(v0.w, d.w, a0[k]))));
                a1[k] = fmaf(v1.x, d.x, fmaf(v1.y, d.y, fmaf(v1.z, d.z, fmaf(v1.w, d.w, a1[k]))));
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int j = j0 + k * J;
            if (j < n_out) {
                acc_add<ATOMIC>(dWt + (int64_t)i0 * ldw + j, a0[k]);
                if (i0 + 1 < n_in) acc_add<ATOMIC>(dWt + (int64_t)(i0 + 1) * ldw + j, a1[k]);
            }
        }
    }
    if (db) {
        for (int j = threadIdx.x; j < n_out; j += kTileThreads) {
            const float* dp = dout + j * S;
            float a = 0.f;
#pragma unroll
            for (int q = 0; q < BM / 4; ++q) {
                float4 d = *reinterpret_cast<const float4*>(dp + 4 * q);
                a += (d.x + d.y) + (d.z + d.w);
            }
            acc_add<ATOMIC>(db + j, a);
        }
    }
}

// [BM samples][n] rows of a global [B, pitch] array  ->  smem [n][S]; rows >= nvalid are zero filled
template <int BM>
__device__ __forceinline__ void load_rows(float* dst, const float* __restrict__ src, int64_t pitch, int n, int nvalid) {
    constexpr int S = Geo<BM>::S;
    for (int idx = threadIdx.x; idx < BM * n; idx += kTileThreads) {
        int s = idx / n, c = idx - s * n;
        dst[c * S + s] = (src && s < nvalid) ? __ldg(src + s * pitch + c) : 0.f;
    }
}
template <int BM>
__device__ __forceinline__ void store_rows(float* dst, int64_t pitch, const float* src, int n, int nvalid) {
    constexpr int S = Geo<BM>::S;
    if (!dst) return;
    for (int idx = threadIdx.x; idx < BM * n; idx += kTileThreads) {
        int s = idx / n, c = idx - s * n;
        if (s < nvalid) dst[s * pitch + c] = src[c * S + s];
    }
}

// One bulk (TMA, UBLKCP) copy of `bytes` (multiple of 16) from global to shared memory, completion on an mbarrier.
__device__ __forceinline__ void bulk_load_and_wait(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* mbar) {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(mbar);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        // chunks of <= 64 KB keep each bulk request comfortably inside the engine's limits
        uint32_t off = 0;
        while (off < bytes) {
            uint32_t n = min(bytes - off, 65536u);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst + off), "l"(reinterpret_cast<const char*>(gsrc) + off), "r"(n), "r"(bar) : "memory");
            off += n;
        }
    }
    // all threads wait for phase 0
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(0u) : "memory");
    }
}

}  // namespace tile
}  // namespace vb
