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
    int src_stride[kMaxGatherArrays];   // floats between source rows (= row_floats for the six separate arrays, = the packed row
                                        // length when the sources are columns of one packed shadow table)
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
                const float* s = a.src[k] + rs * a.src_stride[k];
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

static int gather_impl(const float* const* src, const int* src_strides, float* const* dst, const int* row_floats, int n_arrays,
                       const int64_t* idx, int64_t B, int64_t n_rows, void* stream, const char* who) {
    VB_CHECK_ARG(B >= 0 && n_rows >= 0, VB_ERR_BAD_DIMS, "%s: negative size", who);
    if (B == 0) return 0;      // empty index vector: nothing to move
    VB_CHECK_ARG(src && dst && row_floats && idx, VB_ERR_NULL_POINTER, "%s: null pointer", who);
    VB_CHECK_ARG(n_arrays >= 1 && n_arrays <= kMaxGatherArrays, VB_ERR_BAD_DIMS, "%s: n_arrays=%d not in [1,%d]", who, n_arrays, kMaxGatherArrays);
    GatherArgs a;
    a.n = n_arrays;
    for (int k = 0; k < kMaxGatherArrays; ++k) { a.src[k] = nullptr; a.dst[k] = nullptr; a.row_floats[k] = 0; a.src_stride[k] = 0; a.vec[k] = 1; }
    for (int k = 0; k < n_arrays; ++k) {
        VB_CHECK_ARG(src[k] && dst[k] && row_floats[k] > 0, VB_ERR_NULL_POINTER, "%s: bad array %d", who, k);
        a.src[k] = src[k]; a.dst[k] = dst[k]; a.row_floats[k] = row_floats[k];
        a.src_stride[k] = src_strides ? src_strides[k] : row_floats[k];
        VB_CHECK_ARG(a.src_stride[k] >= row_floats[k], VB_ERR_BAD_DIMS, "%s: source stride of array %d shorter than its rows", who, k);
        uintptr_t al = reinterpret_cast<uintptr_t>(src[k]) | reinterpret_cast<uintptr_t>(dst[k]);
        if ((row_floats[k] & 3) == 0 && (a.src_stride[k] & 3) == 0 && (al & 15) == 0) a.vec[k] = 4;
        else if ((row_floats[k] & 1) == 0 && (a.src_stride[k] & 1) == 0 && (al & 7) == 0) a.vec[k] = 2;
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

extern "C" int vb_gather_rows(const float* const* src, float* const* dst, const int* row_floats, int n_arrays,
                              const int64_t* idx, int64_t B, int64_t n_rows, void* stream) {
    return gather_impl(src, nullptr, dst, row_floats, n_arrays, idx, B, n_rows, stream, "vb_gather_rows");
}

// ------------------------------------------------------------------------------------------------ packed shadow table
// One row per transition holding all arrays' rows back to back: a sampled transition is then ONE run of consecutive sectors
// (80 B = 3 sectors for the reference's 19 floats) instead of one 32-byte sector in each of six arrays (192 B of DRAM traffic
// for 76 useful bytes).  The six public arrays stay the truth; the shadow is refreshed for the rows that were written.
namespace vb {
struct PackRowsArgs {
    const float* src[kMaxGatherArrays];
    int row_floats[kMaxGatherArrays];
    int offset[kMaxGatherArrays];
    int n;
};
__global__ void __launch_bounds__(256) pack_rows_kernel(PackRowsArgs a, float* __restrict__ packed, int row_stride,
                                                        const int64_t* __restrict__ rows, int64_t row0, int64_t n_rows_to_write,
                                                        int64_t capacity, int total_floats) {
    const int64_t total = n_rows_to_write * total_floats;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / total_floats;
        int c = (int)(e - i * total_floats);
        int64_t r = rows ? __ldg(rows + i) : row0 + i;
        if (r < 0 || r >= capacity) continue;
        int k = 0;
        while (k + 1 < a.n && c >= a.row_floats[k]) { c -= a.row_floats[k]; ++k; }
        packed[r * row_stride + a.offset[k] + c] = __ldg(a.src[k] + r * a.row_floats[k] + c);
    }
}
}  // namespace vb

extern "C" int vb_pack_rows(const float* const* src, const int* row_floats, const int* offsets, int n_arrays, float* packed,
                            int row_stride, const int64_t* rows, int64_t row0, int64_t n, int64_t capacity, void* stream) {
    VB_CHECK_ARG(src && row_floats && offsets && packed, VB_ERR_NULL_POINTER, "vb_pack_rows: null pointer");
    VB_CHECK_ARG(n_arrays >= 1 && n_arrays <= kMaxGatherArrays && row_stride > 0 && n >= 0 && capacity >= 0, VB_ERR_BAD_DIMS, "vb_pack_rows: bad sizes");
    if (n == 0) return 0;
    PackRowsArgs a;
    a.n = n_arrays;
    int total = 0;
    for (int k = 0; k < kMaxGatherArrays; ++k) { a.src[k] = nullptr; a.row_floats[k] = 0; a.offset[k] = 0; }
    for (int k = 0; k < n_arrays; ++k) {
        VB_CHECK_ARG(src[k] && row_floats[k] > 0 && offsets[k] >= 0 && offsets[k] + row_floats[k] <= row_stride, VB_ERR_BAD_DIMS,
                     "vb_pack_rows: array %d does not fit the packed row", k);
        a.src[k] = src[k]; a.row_floats[k] = row_floats[k]; a.offset[k] = offsets[k];
        total += row_floats[k];
    }
    int sms = device_sms();
    if (sms <= 0) sms = 148;
    int64_t blocks = (n * total + 255) / 256;
    int64_t cap = (int64_t)sms * 8;
    pack_rows_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(a, packed, row_stride, rows, row0, n, capacity, total);
    VB_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------ one-launch add
// BufferBase.add (velora/buffer/base.py:73-101) writes ONE transition with six indexing assignments (six launches, two of them scalar
// fills) and the shadow table needs a seventh: here the row goes into the six arrays and the packed row in one launch.
namespace vb {
struct AddRowArgs {
    float* dst[kMaxGatherArrays];
    const float* src[kMaxGatherArrays];     // device row of row_floats values, or NULL: broadcast `scalar`
    float scalar[kMaxGatherArrays];
    int row_floats[kMaxGatherArrays];
    int offset[kMaxGatherArrays];
    int n;
};
__global__ void __launch_bounds__(64) buffer_add_row_kernel(AddRowArgs a, int64_t row, float* __restrict__ shadow, int row_stride) {
#pragma unroll
    for (int k = 0; k < kMaxGatherArrays; ++k) {
        if (k < a.n) {
            for (int c = threadIdx.x; c < a.row_floats[k]; c += blockDim.x) {
                const float v = a.src[k] ? __ldg(a.src[k] + c) : a.scalar[k];
                a.dst[k][row * a.row_floats[k] + c] = v;
                if (shadow) shadow[row * row_stride + a.offset[k] + c] = v;
            }
        }
    }
}
}  // namespace vb

extern "C" int vb_buffer_add_row(float* const* dst, const int* row_floats, const float* const* src, const float* scalars, int n_arrays,
                                 int64_t row, int64_t capacity, float* shadow, int row_stride, const int* offsets, void* stream) {
    VB_CHECK_ARG(dst && row_floats && src && scalars, VB_ERR_NULL_POINTER, "vb_buffer_add_row: null pointer");
    VB_CHECK_ARG(n_arrays >= 1 && n_arrays <= kMaxGatherArrays && row >= 0 && row < capacity, VB_ERR_BAD_DIMS,
                 "vb_buffer_add_row: bad sizes (row %lld of %lld)", (long long)row, (long long)capacity);
    VB_CHECK_ARG(!shadow || (offsets && row_stride > 0), VB_ERR_NULL_POINTER, "vb_buffer_add_row: shadow table without a layout");
    AddRowArgs a;
    a.n = n_arrays;
    for (int k = 0; k < kMaxGatherArrays; ++k) { a.dst[k] = nullptr; a.src[k] = nullptr; a.scalar[k] = 0.f; a.row_floats[k] = 0; a.offset[k] = 0; }
    for (int k = 0; k < n_arrays; ++k) {
        VB_CHECK_ARG(dst[k] && row_floats[k] > 0, VB_ERR_BAD_DIMS, "vb_buffer_add_row: array %d", k);
        VB_CHECK_ARG(!shadow || (offsets[k] >= 0 && offsets[k] + row_floats[k] <= row_stride), VB_ERR_BAD_DIMS,
                     "vb_buffer_add_row: array %d does not fit the packed row", k);
        a.dst[k] = dst[k]; a.src[k] = src[k]; a.scalar[k] = scalars[k]; a.row_floats[k] = row_floats[k];
        a.offset[k] = shadow ? offsets[k] : 0;
    }
    buffer_add_row_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(a, row, shadow, row_stride);
    VB_LAUNCH_CHECK();
    return 0;
}

extern "C" int vb_gather_packed(const float* packed, int row_stride, const int* offsets, float* const* dst, const int* row_floats,
                                int n_arrays, const int64_t* idx, int64_t B, int64_t n_rows, void* stream) {
    VB_CHECK_ARG(packed && offsets, VB_ERR_NULL_POINTER, "vb_gather_packed: null pointer");
    VB_CHECK_ARG(n_arrays >= 1 && n_arrays <= kMaxGatherArrays && row_stride > 0, VB_ERR_BAD_DIMS, "vb_gather_packed: bad sizes");
    const float* src[kMaxGatherArrays];
    int strides[kMaxGatherArrays];
    for (int k = 0; k < n_arrays; ++k) { src[k] = packed + offsets[k]; strides[k] = row_stride; }
    return gather_impl(src, strides, dst, row_floats, n_arrays, idx, B, n_rows, stream, "vb_gather_packed");
}

// ------------------------------------------------------------------------------------------------ [B,T,F] <-> [T,B,F]
// The liquid tensor-core kernels read one row per thread: with time-major storage a warp's 32 rows are contiguous, with batch-major
// storage they are T rows apart (32 lines per access instead of 4: measured 0.43 against 0.29 ms on config 3).  A batch-major sequence
// tensor is therefore re-laid once per call by this tiled transposition of F-float rows (both sides coalesced, HBM-bound).
namespace vb {
// V = float, float2 or float4: a row is F elements of V
template <int TT, class V>
__global__ void __launch_bounds__(256) transpose_rows_kernel(const V* __restrict__ src, V* __restrict__ dst, int64_t B, int T, int F,
                                                             int to_time_major) {
    extern __shared__ __align__(16) unsigned char tile_raw[];
    V* tile = reinterpret_cast<V*>(tile_raw);        // [32 samples][TT * F + 1]
    const int pitch = TT * F + 1;
    const int64_t b0 = (int64_t)blockIdx.y * 32;
    const int t0 = blockIdx.x * TT;
    const int nt = min(TT, T - t0);
    const int run = nt * F;                          // elements of one sample inside the tile (contiguous in the batch-major tensor)
    const int nb = (int)min((int64_t)32, B - b0);
    // batch-major side: sample s, elements [t0 * F, t0 * F + run) of its T * F row;  time-major side: step t, elements [b0 * F, (b0 + nb) * F)
    if (to_time_major) {
        for (int e = threadIdx.x; e < nb * run; e += 256) {
            const int s = e / run, j = e - s * run;
            tile[s * pitch + j] = __ldg(src + ((b0 + s) * T + t0) * F + j);
        }
        __syncthreads();
        for (int e = threadIdx.x; e < nt * nb * F; e += 256) {
            const int t = e / (nb * F), r = e - t * (nb * F), s = r / F, f = r - s * F;
            dst[((int64_t)(t0 + t) * B + b0) * F + r] = tile[s * pitch + t * F + f];
        }
    } else {
        for (int e = threadIdx.x; e < nt * nb * F; e += 256) {
            const int t = e / (nb * F), r = e - t * (nb * F), s = r / F, f = r - s * F;
            tile[s * pitch + t * F + f] = __ldg(src + ((int64_t)(t0 + t) * B + b0) * F + r);
        }
        __syncthreads();
        for (int e = threadIdx.x; e < nb * run; e += 256) {
            const int s = e / run, j = e - s * run;
            dst[((b0 + s) * T + t0) * F + j] = tile[s * pitch + j];
        }
    }
}

template <class V>
static int launch_transpose(const float* src, float* dst, int64_t B, int T, int F, int to_time_major, cudaStream_t stream) {
    const int row_bytes = F * (int)sizeof(V);
    const int TT = row_bytes <= 32 ? 32 : (row_bytes <= 64 ? 16 : 4);
    const dim3 grid((T + TT - 1) / TT, (unsigned)((B + 31) / 32));
    const size_t smem = (size_t)32 * (TT * F + 1) * sizeof(V);
    const V* s = reinterpret_cast<const V*>(src);
    V* d = reinterpret_cast<V*>(dst);
    if (TT == 32) transpose_rows_kernel<32, V><<<grid, 256, smem, stream>>>(s, d, B, T, F, to_time_major);
    else if (TT == 16) transpose_rows_kernel<16, V><<<grid, 256, smem, stream>>>(s, d, B, T, F, to_time_major);
    else transpose_rows_kernel<4, V><<<grid, 256, smem, stream>>>(s, d, B, T, F, to_time_major);
    VB_LAUNCH_CHECK();
    return 0;
}
}  // namespace vb

extern "C" int vb_transpose_rows(const float* src, float* dst, int64_t B, int T, int F, int to_time_major, void* stream) {
    VB_CHECK_ARG(B >= 0 && T >= 1 && F >= 1 && F <= 64, VB_ERR_BAD_DIMS, "vb_transpose_rows: bad sizes B=%lld T=%d F=%d", (long long)B, T, F);
    if (B == 0) return 0;
    VB_CHECK_ARG(src && dst && src != dst, VB_ERR_NULL_POINTER, "vb_transpose_rows: null or aliased pointer");
    VB_CHECK_ARG((B + 31) / 32 <= 65535, VB_ERR_TOO_LARGE, "vb_transpose_rows: batch too large for one launch");
    const uintptr_t al = reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst);
    if ((F & 3) == 0 && (al & 15) == 0) return launch_transpose<float4>(src, dst, B, T, F / 4, to_time_major, (cudaStream_t)stream);
    if ((F & 1) == 0 && (al & 7) == 0) return launch_transpose<float2>(src, dst, B, T, F / 2, to_time_major, (cudaStream_t)stream);
    return launch_transpose<float>(src, dst, B, T, F, to_time_major, (cudaStream_t)stream);
}
