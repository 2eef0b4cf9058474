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
    static constexpr int G = kTileThreads / BM;      // output groups
};

// out[j][s] = epi(bias[j] + sum_i Wt[i][j] * in[i][s]),  j < n_out (Wt rows are ldw floats, ldw % 4 == 0, zero padded)
template <int BM, class Epi>
__device__ __forceinline__ void dense(const float* __restrict__ in, int n_in, const float* __restrict__ Wt, int ldw,
                                      const float* __restrict__ bias, int n_out, Epi epi) {
    constexpr int S = Geo<BM>::S, G = Geo<BM>::G;
    const int s = threadIdx.x % BM, g = threadIdx.x / BM;
    const int n_blk = (n_out + 3) >> 2;
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
    const int n_blk = (n_out + 3) >> 2;
    const int total = n_in * n_blk;
    for (int e = threadIdx.x; e < total; e += kTileThreads) {
        const int i = e / n_blk, jb = e - i * n_blk;
        const float* ip = in + i * S;
        const float* dp = dout + (4 * jb) * S;
        const int nj = min(4, n_out - 4 * jb);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int q = 0; q < BM / 4; ++q) {
            float4 v = *reinterpret_cast<const float4*>(ip + 4 * q);
            float4 d0 = *reinterpret_cast<const float4*>(dp + 4 * q);
            a0 = fmaf(v.x, d0.x, a0); a0 = fmaf(v.y, d0.y, a0); a0 = fmaf(v.z, d0.z, a0); a0 = fmaf(v.w, d0.w, a0);
            if (nj > 1) {
                float4 d1 = *reinterpret_cast<const float4*>(dp + S + 4 * q);
                a1 = fmaf(v.x, d1.x, a1); a1 = fmaf(v.y, d1.y, a1); a1 = fmaf(v.z, d1.z, a1); a1 = fmaf(v.w, d1.w, a1);
            }
            if (nj > 2) {
                float4 d2 = *reinterpret_cast<const float4*>(dp + 2 * S + 4 * q);
                a2 = fmaf(v.x, d2.x, a2); a2 = fmaf(v.y, d2.y, a2); a2 = fmaf(v.z, d2.z, a2); a2 = fmaf(v.w, d2.w, a2);
            }
            if (nj > 3) {
                float4 d3 = *reinterpret_cast<const float4*>(dp + 3 * S + 4 * q);
                a3 = fmaf(v.x, d3.x, a3); a3 = fmaf(v.y, d3.y, a3); a3 = fmaf(v.z, d3.z, a3); a3 = fmaf(v.w, d3.w, a3);
            }
        }
        float* o = dWt + (int64_t)i * ldw + 4 * jb;
        acc_add<ATOMIC>(o + 0, a0);
        if (nj > 1) acc_add<ATOMIC>(o + 1, a1);
        if (nj > 2) acc_add<ATOMIC>(o + 2, a2);
        if (nj > 3) acc_add<ATOMIC>(o + 3, a3);
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
