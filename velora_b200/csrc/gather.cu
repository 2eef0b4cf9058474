// Fused replay-buffer minibatch gather: the six `tensor[indices]` launches of velora/buffer/replay.py:81-88 in one
// kernel.  HBM-bound byte movement: per transition 76 B of random row reads (six arrays: 16+4+4+16+4+32 B), 76 B
// of coalesced writes and 8 B of index.  Every thread owns one sampled transition, reads its int64 index once and
// issues all row loads before the first store (six independent requests in flight per thread); rows are moved with
// the widest aligned vector type (128-bit for states / hiddens).
#include "common.cuh"

namespace vb {

constexpr int kMaxGatherArrays = 8;

struct GatherArgs {
    const float* src[kMaxGatherArrays];
    float* dst[kMaxGatherArrays];
    int row_floats[kMaxGatherArrays];
    int vec[kMaxGatherArrays];   // 4, 2 or 1 floats per access
    int n;
};

template <int V> struct VecT;
template <> struct VecT<4> { using type = float4; };
template <> struct VecT<2> { using type = float2; };
template <> struct VecT<1> { using type = float; };

template <int V>
__device__ __forceinline__ void copy_row(const float* __restrict__ s, float* __restrict__ d, int n, bool ok) {
    using T = typename VecT<V>::type;
    const T* sp = reinterpret_cast<const T*>(s);
    T* dp = reinterpret_cast<T*>(d);
    const int cnt = n / V;
    if (ok) {
        // rows are short (<= 8 vectors for every buffer the reference builds): batch the loads, then the stores
        int c = 0;
        for (; c + 4 <= cnt; c += 4) {
            T v0 = __ldg(sp + c), v1 = __ldg(sp + c + 1), v2 = __ldg(sp + c + 2), v3 = __ldg(sp + c + 3);
            __stcs(dp + c, v0); __stcs(dp + c + 1, v1); __stcs(dp + c + 2, v2); __stcs(dp + c + 3, v3);
        }
        for (; c < cnt; ++c) __stcs(dp + c, __ldg(sp + c));
    } else {
        const float nanv = __int_as_float(0x7fc00000);
        for (int c = 0; c < n; ++c) d[c] = nanv;
    }
}

__global__ void __launch_bounds__(256) gather_rows_kernel(GatherArgs a, const int64_t* __restrict__ idx, int64_t B, int64_t n_rows) {
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = __ldg(idx + b);
        const bool ok = r >= 0 && r < n_rows;
        const int64_t rs = ok ? r : 0;
#pragma unroll
        for (int k = 0; k < kMaxGatherArrays; ++k) {
            if (k < a.n) {
                const int n = a.row_floats[k];
                const float* s = a.src[k] + rs * n;
                float* d = a.dst[k] + b * n;
                if (a.vec[k] == 4) copy_row<4>(s, d, n, ok);
                else if (a.vec[k] == 2) copy_row<2>(s, d, n, ok);
                else copy_row<1>(s, d, n, ok);
            }
        }
    }
}

}  // namespace vb

using namespace vb;

extern "C" int vb_gather_rows(const float* const* src, float* const* dst, const int* row_floats, int n_arrays,
                              const int64_t* idx, int64_t B, int64_t n_rows, void* stream) {
    VB_CHECK_ARG(src && dst && row_floats && idx, VB_ERR_NULL_POINTER, "vb_gather_rows: null pointer");
    VB_CHECK_ARG(n_arrays >= 1 && n_arrays <= kMaxGatherArrays, VB_ERR_BAD_DIMS, "vb_gather_rows: n_arrays=%d not in [1,%d]", n_arrays, kMaxGatherArrays);
    VB_CHECK_ARG(B >= 0 && n_rows >= 0, VB_ERR_BAD_DIMS, "vb_gather_rows: negative size");
    if (B == 0) return 0;
    GatherArgs a;
    a.n = n_arrays;
    for (int k = 0; k < kMaxGatherArrays; ++k) { a.src[k] = nullptr; a.dst[k] = nullptr; a.row_floats[k] = 0; a.vec[k] = 1; }
    for (int k = 0; k < n_arrays; ++k) {
        VB_CHECK_ARG(src[k] && dst[k] && row_floats[k] > 0, VB_ERR_NULL_POINTER, "vb_gather_rows: bad array %d", k);
        a.src[k] = src[k]; a.dst[k] = dst[k]; a.row_floats[k] = row_floats[k];
        uintptr_t al = reinterpret_cast<uintptr_t>(src[k]) | reinterpret_cast<uintptr_t>(dst[k]);
        if ((row_floats[k] & 3) == 0 && (al & 15) == 0) a.vec[k] = 4;
        else if ((row_floats[k] & 1) == 0 && (al & 7) == 0) a.vec[k] = 2;
    }
    int sms = device_sms();
    if (sms <= 0) sms = 148;
    int64_t blocks = (B + 255) / 256;
    int64_t cap = (int64_t)sms * 8;   // 8 resident CTAs of 256 threads per SM; grid-stride beyond that
    int grid = (int)(blocks < cap ? blocks : cap);
    gather_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, idx, B, n_rows);
    VB_LAUNCH_CHECK();
    return 0;
}
