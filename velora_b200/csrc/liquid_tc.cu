// Tensor-core (tcgen05 / TMEM) kernels for the 20-neuron class of LiquidNCPNetwork.
//
// Measured notes (profiles/r01): one tcgen05.mma costs the issuing CTA ~60-95 cycles of serial latency whatever its shape
// (N = 16..96, M = 64/128, SS or TS operands) and ~47-50 cycles of SM time when two CTAs interleave - it is bounded by the
// fetch of the 4 KB A operand - so the design minimises the NUMBER of MMAs per step rather than their size: the forward
// stacks parts along K and N (5 MMAs of N = 96), the backward issues 12 heads + 12 dz + 8 dW (its 128 TMEM columns per CTA
// leave no room for the stacked forms); building the two 64-bit descriptors at issue time is free (it hides inside that
// latency; a precomputed table was slower).
// cudaOccupancyMaxActiveBlocksPerMultiprocessor reports 1 CTA/SM for any kernel that allocates TMEM; the real limits
// (shared memory, registers, 512 TMEM columns) allow 4 CTAs per SM for both kernels here.
//
// Why: the FFMA kernels (liquid_small.cu) spend ~2200 of their 4240 warp-instructions per step on three small
// contractions (head pre-activations, dz = ds W, dW = ds^T z) plus ~850 uniform constant loads feeding them
// (profiles/r01).  Here those contractions run on the 5th-generation tensor cores and the CUDA cores keep only
// the element-wise work (Mish, tanh, sigmoid, gating derivatives, the 48-MAC inter layer and the 16-MAC motor layer).
//
// How fp32 accuracy is kept: every operand is split into three bf16 parts (x = b1 + b2 + b3, 24 mantissa bits) and
// the six products of weight <= 2^-16 are accumulated in fp32 TMEM (bf16x3; the same MMA count as 3xTF32, but a
// 16-bit layout).  Measured on B200 (profiles/microbench/umma_round_test.cu): the TMEM accumulator TRUNCATES when
// one MMA is added onto the previous one (inside one MMA the 16 products are summed exactly enough), so per-step
// results issue the small products first and the b1*b1 products last, and the long-running weight-gradient
// accumulator is flushed to an fp32 buffer every kFlush steps.
//
// Layout: a CTA owns 128 samples (thread r = sample r = TMEM lane r).  Per-sample operand vectors live in shared memory
// as "planes": plane c holds features 8c..8c+7 (8 x bf16 = 16 B) of all 128 samples, 2 KB, rows contiguous.  Read K-major
// (rows = samples = M, features = K) the planes are the canonical no-swizzle UMMA layout for  s = z W^T  and
// dz = ds W; read MN-major (features = M / N, samples = K) THE SAME BYTES are the canonical layout for
// dW = ds^T z (verified: profiles/microbench/umma_bf16_test.cu; the 32-bit TF32 types do not accept that
// interleaved MN-major form, umma_test.cu).  The bias is folded in as a constant-one feature.
//
// Replaces the same reference lines as liquid_small.cu (cell.py:147-169, ncp.py:168-171, sparse.py:81 and autograd).
#include <mutex>
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace vb {
namespace tc {
using namespace tcx;   // PTX helpers, split3, store_chunk3, TM, PLANE (tc_common.cuh)

// Row index of (sample b, step t) in a [B,T,F] tensor stored batch-major (row b*T + t) or time-major (row t*B + b).
// One thread owns one sample, so time-major storage makes every warp access one contiguous run of 32 rows.
// MODE 0: every sequence tensor batch-major, 1: every one time-major (strides come straight from the kernel parameters, no
// registers), 2: per-tensor at run time.
template <int MODE>
struct RowMap;
template <>
struct RowMap<0> {
    int T;
    __device__ __forceinline__ RowMap(bool, int64_t, int T_) : T(T_) {}
    __device__ __forceinline__ int64_t operator()(int64_t b, int t) const { return b * T + t; }
};
template <>
struct RowMap<1> {
    int64_t B;
    __device__ __forceinline__ RowMap(bool, int64_t B_, int) : B(B_) {}
    __device__ __forceinline__ int64_t operator()(int64_t b, int t) const { return (int64_t)t * B + b; }
};
template <>
struct RowMap<2> {
    int64_t sb, st;
    __device__ __forceinline__ RowMap(bool time_major, int64_t B, int T) : sb(time_major ? 1 : T), st(time_major ? B : 1) {}
    __device__ __forceinline__ int64_t operator()(int64_t b, int t) const { return b * sb + (int64_t)t * st; }
};
constexpr int LAY_X = 16, LAY_Y = 32, LAY_H = 64, LAY_DX = 128, LAY_DH = 256;   // = VB_FLAG_*_TIME_MAJOR

// constant-bank copy of the small layers' weights (inter / motor and their data-gradients): F section + transposed
// copies, same layout as liquid_small.cu.  A __constant__ symbol is private to its translation unit.
constexpr int kConstFloats = 2048;
__constant__ __align__(16) float c_p[kConstFloats];
static std::mutex g_mu;
static cudaEvent_t g_ev[64] = {nullptr};

static int upload_constants(const float* packed, int n_floats, cudaStream_t stream, int* dev_out, bool* capturing_out) {
    int dev = 0;
    VB_CUDA(cudaGetDevice(&dev));
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    VB_CUDA(cudaStreamIsCapturing(stream, &cs));
    bool capturing = cs != cudaStreamCaptureStatusNone;
    if (!capturing) {
        if (!g_ev[dev & 63]) VB_CUDA(cudaEventCreateWithFlags(&g_ev[dev & 63], cudaEventDisableTiming));
        else VB_CUDA(cudaStreamWaitEvent(stream, g_ev[dev & 63], 0));
    }
    VB_CUDA(cudaMemcpyToSymbolAsync(c_p, packed, (size_t)n_floats * sizeof(float), 0, cudaMemcpyDeviceToDevice, stream));
    *dev_out = dev;
    *capturing_out = capturing;
    return 0;
}

constexpr int kFlush = 11;               // steps between flushes of the truncating dW accumulator (T = 32: three flushes)

template <int I_, int NI_, int NC_, int O_>
struct Cfg {
    static constexpr int I = I_, NI = NI_, NC = NC_, O = O_;
    static constexpr int NIp = ceil4(NI), NCp = ceil4(NC), Op = ceil4(O), K = NI + NC, H4 = 4 * NCp;
    static constexpr int Ip = ceil4(I), Kp = ceil4(K);
    static constexpr int WIN_T = 0, B_IN = WIN_T + I * NIp, WH_T = B_IN + NIp, B_H = WH_T + K * H4, WOUT = B_H + H4,
                         B_OUT = WOUT + O * NCp, F_TOTAL = B_OUT + Op, WIN = F_TOTAL, WH = WIN + NI * Ip, C_TOTAL = WH + H4 * Kp;
    // z~ = [a (NI) | h (NC) | 1 at ONE | 0..] : 24 features = three 8-feature planes per bf16 part
    static constexpr int ONE = K;
    static_assert(H4 == 32, "tensor-core path: 4 * Nc must be 32 (Nc = 8)");
    static_assert(K + 1 <= 24, "tensor-core path: Ni + Nc + 1 must fit three 8-feature planes");
    static_assert(C_TOTAL <= kConstFloats, "constant-bank budget");
    // small-layer gradients reduced on the CUDA cores: dW_in, dW_out, db_in, db_out
    static constexpr int N_SMALL = I * NI + O * NC + NI + O;
    static constexpr int SROWS = NI + I + O + NC;      // fp32 rows du | x | dy | m of the transposition buffer
    static_assert(SROWS * 132 * 4 <= 12 * PLANE, "transposition buffer must fit the ds planes");
    static_assert((16 * PLANE + NC * 128 * 4) / 4 < 0xfffe, "row offsets are packed in 16 bits");
};


// product order for one per-step contraction: small terms first, b1*b1 last (the accumulator truncates)
// heads: D[128 x 32] = z~ W^T.  z~ planes: 3 chunks per part (24 features) followed by ONE shared all-zero plane (index 9)
// that stands in for features 24..31 of every part: the second K-step's descriptor reaches it through its own LBO.
// 6 part-products x 2 K-steps = 12 MMAs, issued by one thread.
__device__ __forceinline__ void issue_heads(uint32_t d_tmem, uint32_t z_planes, uint32_t w_planes) {
    constexpr uint32_t idesc = make_idesc(128, 32, 0, 0);
#pragma unroll
    for (int p = 0; p < 6; ++p) {
        const int pa = (p == 0) ? 1 : (p == 1) ? 0 : (p == 2) ? 2 : (p == 3) ? 0 : (p == 4) ? 1 : 0;
        const int pb = (p == 0) ? 1 : (p == 1) ? 2 : (p == 2) ? 0 : (p == 3) ? 1 : (p == 4) ? 0 : 0;
        uint64_t a0 = make_desc(z_planes + (pa * 3 + 0) * PLANE, PLANE, 128);                   // chunks 0, 1
        uint64_t a1 = make_desc(z_planes + (pa * 3 + 2) * PLANE, (9 - (pa * 3 + 2)) * PLANE, 128);   // chunk 2, zero plane
        uint64_t b0 = make_desc(w_planes + (pb * 4 + 0) * (32 * 16), 32 * 16, 128);
        uint64_t b1 = make_desc(w_planes + (pb * 4 + 2) * (32 * 16), 32 * 16, 128);
        mma_bf16(d_tmem, a0, b0, idesc, p != 0);
        mma_bf16(d_tmem, a1, b1, idesc, 1);
    }
}

// weight operand planes, built once per CTA from the packed fp32 block: out[part][kchunk][n (32 rows)][8 bf16]
//   value(n, k) supplied by f(n, k)
template <class F>
__device__ __forceinline__ void build_weight_planes(unsigned char* out, F f) {
    const int n = threadIdx.x & 31, kc = threadIdx.x >> 5;   // 32 rows x 4 chunks = 128 threads
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = f(n, kc * 8 + e);
    uint4 a, b, c;
    split3(v[0], v[1], a.x, b.x, c.x);
    split3(v[2], v[3], a.y, b.y, c.y);
    split3(v[4], v[5], a.z, b.z, c.z);
    split3(v[6], v[7], a.w, b.w, c.w);
    uint4* p = reinterpret_cast<uint4*>(out);
    p[(0 * 4 + kc) * 32 + n] = a;
    p[(1 * 4 + kc) * 32 + n] = b;
    p[(2 * 4 + kc) * 32 + n] = c;
}

template <int N>
__device__ __forceinline__ void ldg_row(float (&dst)[N], const float* p, bool vec_ok) {
    if ((N & 3) == 0 && vec_ok) {
#pragma unroll
        for (int q = 0; q < N / 4; ++q) {
            float4 v = __ldg(reinterpret_cast<const float4*>(p) + q);
            dst[4 * q] = v.x; dst[4 * q + 1] = v.y; dst[4 * q + 2] = v.z; dst[4 * q + 3] = v.w;
        }
    } else if ((N & 1) == 0 && vec_ok) {
#pragma unroll
        for (int q = 0; q < N / 2; ++q) {
            float2 v = __ldg(reinterpret_cast<const float2*>(p) + q);
            dst[2 * q] = v.x; dst[2 * q + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int q = 0; q < N; ++q) dst[q] = __ldg(p + q);
    }
}
template <int N>
__device__ __forceinline__ void stg_row(float* p, const float (&src)[N], bool vec_ok) {
    if ((N & 3) == 0 && vec_ok) {
#pragma unroll
        for (int q = 0; q < N / 4; ++q)
            reinterpret_cast<float4*>(p)[q] = make_float4(src[4 * q], src[4 * q + 1], src[4 * q + 2], src[4 * q + 3]);
    } else if ((N & 1) == 0 && vec_ok) {
#pragma unroll
        for (int q = 0; q < N / 2; ++q) reinterpret_cast<float2*>(p)[q] = make_float2(src[2 * q], src[2 * q + 1]);
    } else {
#pragma unroll
        for (int q = 0; q < N; ++q) p[q] = src[q];
    }
}

template <class C>
__device__ __forceinline__ void inter_layer(const float (&x)[C::I], float (&u)[C::NI]) {
#pragma unroll
    for (int j = 0; j < C::NI; ++j) u[j] = c_p[C::B_IN + j];
#pragma unroll
    for (int i = 0; i < C::I; ++i)
#pragma unroll
        for (int j = 0; j < C::NI; ++j) u[j] = fmaf(x[i], c_p[C::WIN_T + i * C::NIp + j], u[j]);
}

// z~ = [a (NI), h (NC), 1, 0..] -> chunks 0..2 of the three part planes
template <class C>
__device__ __forceinline__ void store_z(unsigned char* zP, int row, const float (&a)[C::NI], const float (&h)[C::NC]) {
    float zt[24];
#pragma unroll
    for (int i = 0; i < 24; ++i) zt[i] = 0.f;
#pragma unroll
    for (int j = 0; j < C::NI; ++j) zt[j] = a[j];
#pragma unroll
    for (int j = 0; j < C::NC; ++j) zt[C::NI + j] = h[j];
    zt[C::ONE] = 1.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = zt[8 * c + e];
        store_chunk3<3>(zP, c, row, f);
    }
}

// ================================================================================================ forward
template <class C>
struct FwdSmem {
    static constexpr int Z = 0;                                  // 9 z~ planes + the shared zero plane
    static constexpr int WB = Z + 10 * PLANE;                    // stacked heads operand: 10 K-chunks x 96 rows x 16 B
    static constexpr int BAR = WB + 10 * 96 * 16;
    static constexpr int BYTES = BAR + 64;
};

// Forward heads in FIVE MMAs: the three bf16 parts of z~ are stacked along K (the 9 part planes + the zero plane are
// contiguous: K = 80) and the three parts of the weights along N (96 rows), D[128 x 96] = [z1|z2|z3|0] B^T with
// B[row(pb, n')][pa * 24 + f] = W_pb[n'][f].  Column group pb then holds z W_pb (all nine part products); the epilogue adds
// the three groups, smallest first.  An M = 128, K = 16 MMA costs ~50 cycles of shared-memory operand fetch whatever its N
// (A is 4 KB), so 5 wide MMAs replace 12 narrow ones.  Row order: row = (n' / 8) * 24 + pb * 8 + n' % 8, i.e. the three groups
// of one 8-column chunk (= two neurons) are adjacent TMEM columns.
__device__ __forceinline__ void issue_heads_stacked(uint32_t d_tmem, uint32_t z_planes, uint32_t w_planes) {
    constexpr uint32_t idesc = make_idesc(128, 96, 0, 0);
#pragma unroll
    for (int ks = 0; ks < 5; ++ks)
        mma_bf16(d_tmem, make_desc(z_planes + ks * 2 * PLANE, PLANE, 128), make_desc(w_planes + ks * 2 * (96 * 16), 96 * 16, 128), idesc, ks != 0);
}

template <class C, bool VEC, int LAY>
__global__ void __launch_bounds__(TM) liquid_fwd_tc_kernel(const float* __restrict__ packed, const float* __restrict__ x,
                                                           const float* __restrict__ h0, float* __restrict__ y,
                                                           float* __restrict__ h_traj, int64_t B, int T, int vec_flags) {
    using S = FwdSmem<C>;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* zP = smem + S::Z;
    unsigned char* wB = smem + S::WB;
    uint64_t* bar1 = reinterpret_cast<uint64_t*>(smem + S::BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::BAR + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // VEC: every row pointer is aligned for its widest vector access (decided on the host); the scalar variant serves odd views
    constexpr bool vx = VEC, vy = VEC, vh = VEC;
    const RowMap<LAY> rx(vec_flags & LAY_X, B, T), ry(vec_flags & LAY_Y, B, T), rh(vec_flags & LAY_H, B, T), rdx(vec_flags & LAY_DX, B, T),
        rdh(vec_flags & LAY_DH, B, T);

    // stacked heads operand (see issue_heads_stacked); head outputs in NEURON-MAJOR order n' = 4 j + hs, so that one 8-column
    // chunk holds two complete neurons
    if (warp < 3) {
        const int n = lane, c = warp;     // features 8c .. 8c+7 of output n'
        const int col = (n & 3) * C::NCp + (n >> 2);
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = 8 * c + e;
            v[e] = k < C::K ? __ldg(packed + C::WH_T + k * C::H4 + col) : (k == C::ONE ? __ldg(packed + C::B_H + col) : 0.f);
        }
        uint4 part[3];
        split3(v[0], v[1], part[0].x, part[1].x, part[2].x);
        split3(v[2], v[3], part[0].y, part[1].y, part[2].y);
        split3(v[4], v[5], part[0].z, part[1].z, part[2].z);
        split3(v[6], v[7], part[0].w, part[1].w, part[2].w);
        uint4* q = reinterpret_cast<uint4*>(wB);
#pragma unroll
        for (int pa = 0; pa < 3; ++pa)
#pragma unroll
            for (int pb = 0; pb < 3; ++pb) q[(pa * 3 + c) * 96 + (n >> 3) * 24 + pb * 8 + (n & 7)] = part[pb];
    } else {
        for (int r = lane; r < 96; r += 32) reinterpret_cast<uint4*>(wB)[9 * 96 + r] = make_uint4(0, 0, 0, 0);
    }
    reinterpret_cast<uint4*>(zP)[9 * TM + tid] = make_uint4(0, 0, 0, 0);   // the shared zero plane
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(bar1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t z_addr = smem_u32(zP), w_addr = smem_u32(wB);
    uint32_t ph1 = 0;

    const int64_t n_tiles = (B + TM - 1) / TM;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t b = tile * TM + tid;
        const bool valid = b < B;
        float h[C::NC];
#pragma unroll
        for (int j = 0; j < C::NC; ++j) h[j] = 0.f;
        if (valid && h0) ldg_row<C::NC>(h, h0 + b * C::NC, vh);
        float xn[C::I];
#pragma unroll
        for (int i = 0; i < C::I; ++i) xn[i] = 0.f;
        if (valid) ldg_row<C::I>(xn, x + rx(b, 0) * C::I, vx);
        for (int t = 0; t < T; ++t) {
            float xr[C::I];
#pragma unroll
            for (int i = 0; i < C::I; ++i) xr[i] = xn[i];
            if (valid && t + 1 < T) ldg_row<C::I>(xn, x + rx(b, t + 1) * C::I, vx);   // prefetch the next step's input
            float a[C::NI];
            {
                float u[C::NI];
                inter_layer<C>(xr, u);
                static_assert(C::NI % 4 == 0, "inter neurons are activated four at a time");
#pragma unroll
                for (int j = 0; j < C::NI; j += 4) {
                    float u4[4] = {u[j], u[j + 1], u[j + 2], u[j + 3]}, a4[4];
                    mish4(u4, a4);
                    a[j] = a4[0]; a[j + 1] = a4[1]; a[j + 2] = a4[2]; a[j + 3] = a4[3];
                }
            }
            store_z<C>(zP, tid, a, h);
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                issue_heads_stacked(tmem, z_addr, w_addr);
                mma_commit(bar1);
            }
            mbar_wait(bar1, ph1);
            ph1 ^= 1;
            tc_fence_after();
            float m[C::NC];
#pragma unroll
            for (int c = 0; c < C::NC / 2; ++c) {
                float s8[8];
                {
                    float hi[8], mid[8], lo[8];
                    tmem_ld8(taddr + 24 * c, hi);
                    tmem_ld8(taddr + 24 * c + 8, mid);
                    tmem_ld8(taddr + 24 * c + 16, lo);
#pragma unroll
                    for (int e = 0; e < 8; ++e) s8[e] = (lo[e] + mid[e]) + hi[e];
                }
                float cc[2];
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    float tg, th, gate;
                    tanh2_sigmoid(s8[4 * jj + 0], s8[4 * jj + 1], s8[4 * jj + 2], tg, th, gate);
                    float hn = fmaf(gate, th - tg, tg);
                    h[2 * c + jj] = hn;
                    cc[jj] = s8[4 * jj + 3] + hn;
                }
                mish2(cc[0], cc[1], m[2 * c], m[2 * c + 1]);
            }
            float yo[C::O];
#pragma unroll
            for (int o = 0; o < C::O; ++o) {
                float acc = c_p[C::B_OUT + o];
#pragma unroll
                for (int j = 0; j < C::NC; ++j) acc = fmaf(m[j], c_p[C::WOUT + o * C::NCp + j], acc);
                yo[o] = acc;
            }
            if (valid) {
                stg_row<C::O>(y + ry(b, t) * C::O, yo, vy);
                stg_row<C::NC>(h_traj + rh(b, t) * C::NC, h, vh);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

// ================================================================================================ backward
template <class C>
struct BwdSmem {
    static constexpr int Z = 0;                                  // 9 z~ planes + shared zero plane
    static constexpr int DS = Z + 10 * PLANE;                    // ds parts: 12 planes; the MN-major M = 128 stack reads 4 planes past the end
    static constexpr int WB = DS + 12 * PLANE;                   // heads operand (6 KB), also read MN-major for dz = ds W
    static constexpr int BAR = WB + 3 * 4 * 32 * 16;
    static constexpr int TAIL = BAR + 64;
    static constexpr int MSTRIP = DS + 16 * PLANE > TAIL ? DS + 16 * PLANE : TAIL;  // keep the over-read inside the CTA's window
    static constexpr int BYTES = MSTRIP + C::NC * TM * 4;                           // m[Nc][128] fp32, parked between gating and the small-gradient pass
};

constexpr int D3_COLS = 96;                 // flush buffer columns per thread (80 used)
constexpr int D3_FLOATS = TM * D3_COLS;     // per-CTA flush buffer
constexpr int SBUF_STRIDE = 132;            // fp32 transposition buffer row stride (floats): 128 samples + 4 pad

template <class C, bool VEC, int LAY>
__global__ void __launch_bounds__(TM, 4) liquid_bwd_tc_kernel(
    const float* __restrict__ packed, const float* __restrict__ x, const float* __restrict__ h0, const float* __restrict__ h_traj,
    const float* __restrict__ dy, const float* __restrict__ dh_traj, const float* __restrict__ dh_last,
    float* __restrict__ dx, float* __restrict__ dh0, float* __restrict__ partials, float* __restrict__ d3buf,
    int64_t B, int T, int vec_flags, long long* __restrict__ trace) {
    using S = BwdSmem<C>;
    // phase timing for profiles/microbench/trace_run.py: compiled in only with -DVB_ENABLE_TRACE (VB_TRACE_BUILD=1 build.sh)
#ifdef VB_ENABLE_TRACE
#define VB_TRACE(slot) do { if (trace && blockIdx.x == 0 && tid == 0 && T - 1 - t < 40) trace[(T - 1 - t) * 16 + (slot)] = clock64(); } while (0)
#define VB_MARK(slot) do { if (trace && blockIdx.x == 0 && tid == 0) trace[40 * 16 + (slot)] = clock64(); } while (0)
#else
#define VB_TRACE(slot) do { } while (0)
#define VB_MARK(slot) do { } while (0)
#endif
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    VB_MARK(0);
#ifdef VB_ENABLE_TRACE
    if (trace && tid == 0 && blockIdx.x < 1024) {
        unsigned long long gt; unsigned smid;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        trace[1024 + blockIdx.x * 4] = (long long)gt; trace[1024 + blockIdx.x * 4 + 2] = smid;
    }
#endif
    unsigned char* zP = smem + S::Z;
    unsigned char* dsP = smem + S::DS;
    unsigned char* wB = smem + S::WB;
    float* sbuf = reinterpret_cast<float*>(dsP);                   // fp32 [SROWS][SBUF_STRIDE], aliases the ds planes between MMAs
    float* mstrip = reinterpret_cast<float*>(smem + S::MSTRIP);
    uint64_t* bar1 = reinterpret_cast<uint64_t*>(smem + S::BAR);   // heads ready
    uint64_t* bar2 = bar1 + 1;                                     // dz ready
    uint64_t* bar3 = bar1 + 2;                                     // dW batch done (z~ and ds planes may be overwritten)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::BAR + 32);
    // VEC: every row pointer is aligned for its widest vector access (decided on the host); the scalar variant serves odd views
    constexpr bool vx = VEC, vy = VEC, vh = VEC;
    const RowMap<LAY> rx(vec_flags & LAY_X, B, T), ry(vec_flags & LAY_Y, B, T), rh(vec_flags & LAY_H, B, T), rdx(vec_flags & LAY_DX, B, T),
        rdh(vec_flags & LAY_DH, B, T);

    // heads operand B[n][k]: W_heads[k][n] for k < K, bias for k == ONE, zero elsewhere
    // rows in NEURON-MAJOR head order n' = 4 j + hs, so that one 8-column TMEM load holds two complete neurons
    build_weight_planes(wB, [&](int n, int k) {
        const int col = (n & 3) * C::NCp + (n >> 2);
        return k < C::K ? __ldg(packed + C::WH_T + k * C::H4 + col) : (k == C::ONE ? __ldg(packed + C::B_H + col) : 0.f);
    });
    {   // zero plane, the ds area and the over-read tail (so the first M = 128 stack read is finite)
        const uint4 z4 = make_uint4(0, 0, 0, 0);
        reinterpret_cast<uint4*>(zP)[9 * TM + tid] = z4;
        uint4* q = reinterpret_cast<uint4*>(dsP);
        for (int i = tid; i < (S::BYTES - S::DS) / 16; i += TM)
            if (i < 12 * TM || i >= (S::TAIL - S::DS + 15) / 16) q[i] = z4;
    }
    // fp32 flush buffer of the dW accumulator: element (stack row = thread, column c) at ((c / 4) * TM + row) * 4 + c % 4, so that the
    // 32 lanes of one vector access touch 512 contiguous bytes
    float* d3row = d3buf + (int64_t)blockIdx.x * D3_FLOATS + tid * 4;
#pragma unroll
    for (int q = 0; q < D3_COLS / 4; ++q) *reinterpret_cast<float4*>(d3row + q * (TM * 4)) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(bar1, 1);
        mbar_init(bar2, 1);
        mbar_init(bar3, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t z_addr = smem_u32(zP), ds_addr = smem_u32(dsP), wb_addr = smem_u32(wB);
    // TMEM columns: [0,32) heads, then dz (reused);  [32,112) dW = stack(ds)^T x [z~_b1 | z~_b2 | z~_b3 | 0]
    uint32_t ph = 0;              // all three barriers complete exactly once per step: one shared phase bit
    int since_flush = 0;          // dW batches accumulated in TMEM since the last flush (0 => next batch overwrites)

    // small-layer gradients: register-blocked reductions over the fp32 transposition buffer (rows = features, columns = the
    // tile's 128 samples), 5x less shared-memory traffic than one (entry, sample-quarter) unit per thread:
    //   warps 0-1  dW_in[i][j], db_in[j] for the three inter neurons j = 3 blk + r (4 blocks) over samples {4 sg .. 4 sg + 3} and
    //              {64 + 4 sg ..} (16 sample groups): 7 rows x 2 LDS.128 feed 120 FMA/ADD
    //   warps 2-3  dW_out[o][j], db_out[o] for the four command neurons j = 4 blk + r (2 blocks) over samples {4 sg ..} (32 groups)
    constexpr int R_DU = 0, R_X = C::NI, R_DY = C::NI + C::I;
    static_assert(C::NI == 12 && C::NC == 8, "small-gradient blocking is written for 12 inter / 8 command neurons");
    constexpr int JA = 3, JB = 4;
    constexpr int NACC = (C::I * JA + JA > C::O * JB + C::O) ? (C::I * JA + JA) : (C::O * JB + C::O);
    float acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.f;
    const int sgA = tid & 15, blkA = (tid >> 4) & 3, sgB = tid & 31, blkB = (tid >> 5) & 1;
    const float* rowsA = sbuf + 4 * sgA;                                      // + row * SBUF_STRIDE (+ 64 for the second half)
    const float* rowsB = sbuf + 4 * sgB;
    const float* duA = rowsA + (R_DU + JA * blkA) * SBUF_STRIDE;
    const float* mB = mstrip + (JB * blkB) * TM + 4 * sgB;

    // Reduction of the rows parked in sbuf / mstrip by the previous step (see the step tail).  It only needs shared memory, so it
    // runs while the tensor core works on the next step's heads, before the gating pass overwrites the ds planes.
    bool pending = false;
    auto reduce_small = [&]() {
        if (warp < 2) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float4 xv[C::I], dv[JA];
#pragma unroll
                for (int i = 0; i < C::I; ++i) xv[i] = *reinterpret_cast<const float4*>(rowsA + (R_X + i) * SBUF_STRIDE + 64 * half);
#pragma unroll
                for (int r = 0; r < JA; ++r) dv[r] = *reinterpret_cast<const float4*>(duA + r * SBUF_STRIDE + 64 * half);
#pragma unroll
                for (int r = 0; r < JA; ++r) {
#pragma unroll
                    for (int i = 0; i < C::I; ++i)
                        acc[i * JA + r] = fmaf(xv[i].x, dv[r].x, fmaf(xv[i].y, dv[r].y, fmaf(xv[i].z, dv[r].z, fmaf(xv[i].w, dv[r].w, acc[i * JA + r]))));
                    acc[C::I * JA + r] += (dv[r].x + dv[r].y) + (dv[r].z + dv[r].w);
                }
            }
        } else {
            float4 yv[C::O], mv[JB];
#pragma unroll
            for (int o = 0; o < C::O; ++o) yv[o] = *reinterpret_cast<const float4*>(rowsB + (R_DY + o) * SBUF_STRIDE);
#pragma unroll
            for (int r = 0; r < JB; ++r) mv[r] = *reinterpret_cast<const float4*>(mB + r * TM);
#pragma unroll
            for (int o = 0; o < C::O; ++o) {
#pragma unroll
                for (int r = 0; r < JB; ++r)
                    acc[o * JB + r] = fmaf(yv[o].x, mv[r].x, fmaf(yv[o].y, mv[r].y, fmaf(yv[o].z, mv[r].z, fmaf(yv[o].w, mv[r].w, acc[o * JB + r]))));
                acc[C::O * JB + o] += (yv[o].x + yv[o].y) + (yv[o].z + yv[o].w);
            }
        }
    };

    auto flush_dw = [&]() {   // all threads; requires the dW batch to be complete
        tc_fence_after();
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            float v[32];
            tmem_ld32(taddr + 32 + 32 * g, v);
            // fire-and-forget vector reductions into this thread's private row: no load, nothing to wait for
#pragma unroll
            for (int q = 0; q < 8; ++q)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d3row + (8 * g + q) * (TM * 4)), "f"(v[4 * q]), "f"(v[4 * q + 1]),
                             "f"(v[4 * q + 2]), "f"(v[4 * q + 3]) : "memory");
        }
        since_flush = 0;
    };

    VB_MARK(1);
    const int64_t n_tiles = (B + TM - 1) / TM;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t b = tile * TM + tid;
        const bool valid = b < B;
        float dhc[C::NC];
#pragma unroll
        for (int j = 0; j < C::NC; ++j) dhc[j] = 0.f;
        if (valid && dh_last) ldg_row<C::NC>(dhc, dh_last + b * C::NC, vh);
        // x_t is prefetched one step ahead (it is needed first thing); h_{t-1} and dy_t are requested at the top of the
        // step and consumed after the inter layer / Mish, which covers their latency without extra live registers
        float xn[C::I], dyn[C::O];
#pragma unroll
        for (int i = 0; i < C::I; ++i) xn[i] = 0.f;
#pragma unroll
        for (int o = 0; o < C::O; ++o) dyn[o] = 0.f;
        if (valid) {
            ldg_row<C::I>(xn, x + rx(b, T - 1) * C::I, vx);
            ldg_row<C::O>(dyn, dy + ry(b, T - 1) * C::O, vy);
        }
        for (int t = T - 1; t >= 0; --t) {
            float xr[C::I], dyr[C::O], hp[C::NC];
#pragma unroll
            for (int i = 0; i < C::I; ++i) xr[i] = xn[i];
#pragma unroll
            for (int o = 0; o < C::O; ++o) dyr[o] = dyn[o];
#pragma unroll
            for (int j = 0; j < C::NC; ++j) hp[j] = 0.f;
            if (valid) {
                if (t > 0) ldg_row<C::NC>(hp, h_traj + rh(b, t - 1) * C::NC, vh);
                else if (h0) ldg_row<C::NC>(hp, h0 + b * C::NC, vh);
                if (t > 0) {
                    ldg_row<C::I>(xn, x + rx(b, t - 1) * C::I, vx);
                    ldg_row<C::O>(dyn, dy + ry(b, t - 1) * C::O, vy);
                }
            }
            VB_TRACE(0);
            // ---------------- recompute: inter layer on CUDA cores, heads on the tensor cores
            {
                float a[C::NI], u[C::NI], park[16];
                inter_layer<C>(xr, u);
#pragma unroll
                for (int j = 0; j < C::NI; ++j) a[j] = mish_fwd_grad(u[j], park[j]);
                // Mish'(u) and x_t are next needed after the dz MMA: they wait in this thread's TMEM lane (the 16 columns the
                // accumulators leave free) instead of holding 16 registers through the gating pass
                static_assert(C::NI + C::I <= 16, "parking area is 16 TMEM columns");
#pragma unroll
                for (int i = 0; i < 16 - C::NI; ++i) park[C::NI + i] = i < C::I ? xr[i < C::I ? i : 0] : 0.f;
                tmem_st16(taddr + 112, park);
                store_z<C>(zP, tid, a, hp);     // the previous dW batch was waited for at the end of the previous step
            }
            VB_TRACE(1);
            fence_async_smem();
            tc_fence_before();
            __syncthreads();                    // also: everybody is done reading the fp32 buffer of the previous step
            VB_TRACE(2);
            if (tid == 96) {                  // a lane of the lighter (dW_out) reduction warps issues
                tc_fence_after();
                issue_heads(tmem + 0, z_addr, wb_addr);
                mma_commit(bar1);
            }
            VB_TRACE(3);
            if (pending) {
                reduce_small();
                __syncthreads();              // everybody has read sbuf / mstrip: the gating pass may overwrite them
            }
            mbar_wait(bar1, ph);
            tc_fence_after();
            VB_TRACE(4);
            // ---------------- gating derivatives, two neurons (= one 8-feature chunk of ds) at a time
            {
                const float* dhe_p = (valid && dh_traj) ? dh_traj + rdh(b, t) * C::NC : nullptr;
#pragma unroll
                for (int c = 0; c < C::NC / 2; ++c) {
                    float s8[8];
                    tmem_ld8(taddr + 8 * c, s8);
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) {
                        const int j = 2 * c + jj;
                        float tg = tanh_f(s8[4 * jj + 0]), th = tanh_f(s8[4 * jj + 1]), gate = sigmoid_f(s8[4 * jj + 2]);
                        float hn = fmaf(gate, th - tg, tg);
                        float dmc;
                        mstrip[j * TM + tid] = mish_fwd_grad(s8[4 * jj + 3] + hn, dmc);
                        float dm = 0.f;
#pragma unroll
                        for (int o = 0; o < C::O; ++o) dm = fmaf(dyr[o], c_p[C::WOUT + o * C::NCp + j], dm);
                        float dc = dm * dmc;
                        float dhn = dc + dhc[j] + (dhe_p ? __ldg(dhe_p + j) : 0.f);
                        float omg = 1.0f - gate;
                        s8[4 * jj + 0] = dhn * omg * (1.0f - tg * tg);       // ds_g
                        s8[4 * jj + 1] = dhn * gate * (1.0f - th * th);      // ds_h
                        s8[4 * jj + 2] = dhn * (th - tg) * gate * omg;       // ds_f (both f heads)
                        s8[4 * jj + 3] = dc;                                 // ds_proj
                    }
                    store_chunk3<4>(dsP, c, tid, s8);
                }
            }
            VB_TRACE(5);
            fence_async_smem();
            tc_fence_before();     // also orders this thread's TMEM read of the heads before dz overwrites those columns
            __syncthreads();
            VB_TRACE(6);
            if (lane == 0) {
                if (warp == 1) {            // dz = ds W
                    tc_fence_after();
                    constexpr uint32_t idesc2 = make_idesc(128, 32, 0, 1);
#pragma unroll
                    for (int p = 0; p < 6; ++p) {
                        const int pa = (p == 0) ? 1 : (p == 1) ? 0 : (p == 2) ? 2 : (p == 3) ? 0 : (p == 4) ? 1 : 0;
                        const int pb = (p == 0) ? 1 : (p == 1) ? 2 : (p == 2) ? 0 : (p == 3) ? 1 : (p == 4) ? 0 : 0;
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks)
                            mma_bf16(tmem + 0, make_desc(ds_addr + (pa * 4 + 2 * ks) * PLANE, PLANE, 128),
                                     make_desc(wb_addr + pb * (4 * 32 * 16) + ks * 256, 128, 32 * 16), idesc2, (p | ks) != 0);
                    }
                    mma_commit(bar2);
                } else if (warp == 2) {     // dW (+ bias column) += stack(ds)^T x [z~_b1 | z~_b2 | z~_b3 | 0] : all nine part-products at once
                    tc_fence_after();
                    constexpr uint32_t idesc3 = make_idesc(128, 80, 1, 1);
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        mma_bf16(tmem + 32, make_desc(ds_addr + ks * 256, 128, PLANE), make_desc(z_addr + ks * 256, 128, PLANE), idesc3,
                                 !(since_flush == 0 && ks == 0));
                    mma_commit(bar3);
                }
            }
            since_flush += 1;
            VB_TRACE(7);
            mbar_wait(bar2, ph);
            tc_fence_after();
            VB_TRACE(8);
            float du[C::NI], xr2[C::I];
            {
                float dmu[16];
                tmem_ld16(taddr + 112, dmu);
#pragma unroll
                for (int i = 0; i < C::I; ++i) xr2[i] = dmu[C::NI + i];
                float dz[16];
                tmem_ld16(taddr + 0, dz);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (i < C::NI) du[i < C::NI ? i : 0] = dz[i] * dmu[i < C::NI ? i : 0];
                    else if (i < C::K) dhc[(i - C::NI) >= 0 ? (i - C::NI) : 0] = dz[i];
                }
                float dz2[8];
                tmem_ld8(taddr + 16, dz2);
#pragma unroll
                for (int i = 16; i < 24; ++i) {
                    if (i < C::NI) du[i < C::NI ? i : 0] = dz2[i - 16] * dmu[i < C::NI ? i : 0];
                    else if (i < C::K) dhc[(i - C::NI) < C::NC ? (i - C::NI) : 0] = dz2[i - 16];
                }
            }
            {
                float dxr[C::I];
#pragma unroll
                for (int i = 0; i < C::I; ++i) dxr[i] = 0.f;
#pragma unroll
                for (int j = 0; j < C::NI; ++j)
#pragma unroll
                    for (int i = 0; i < C::I; ++i) dxr[i] = fmaf(du[j], c_p[C::WIN + j * C::Ip + i], dxr[i]);
                if (valid && dx) stg_row<C::I>(dx + rdx(b, t) * C::I, dxr, vx);
            }
            // ---------------- small-layer gradients: fp32 transposition through the (now idle) ds planes
            VB_TRACE(9);
            mbar_wait(bar3, ph);       // the dW batch has finished reading the z~ and ds planes
            ph ^= 1;
            VB_TRACE(10);
            if (since_flush >= kFlush) flush_dw();
#pragma unroll
            for (int j = 0; j < C::NI; ++j) sbuf[(R_DU + j) * SBUF_STRIDE + tid] = du[j];
#pragma unroll
            for (int i = 0; i < C::I; ++i) sbuf[(R_X + i) * SBUF_STRIDE + tid] = xr2[i];
#pragma unroll
            for (int o = 0; o < C::O; ++o) sbuf[(R_DY + o) * SBUF_STRIDE + tid] = dyr[o];
            pending = true;      // reduced in the shadow of the next step's heads MMAs (or after the loop)
            VB_TRACE(11);
        }
        if (valid && dh0) stg_row<C::NC>(dh0 + b * C::NC, dhc, vh);
    }
    if (pending) {
        __syncthreads();
        reduce_small();
    }
    VB_MARK(2);
    if (since_flush > 0) flush_dw();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");

    // ---------------- per-CTA partial gradient row (F layout)
    float* prow = partials + (int64_t)blockIdx.x * C::F_TOTAL;
    for (int idx = tid; idx < C::F_TOTAL; idx += TM) prow[idx] = 0.f;
    // per-thread small-gradient accumulators -> red[entry][sample group] -> one sum per entry
    constexpr int RS = 33;
    float* red = reinterpret_cast<float*>(smem);   // [N_SMALL][RS]; the operand planes are dead
    if (warp < 2) {
#pragma unroll
        for (int r = 0; r < JA; ++r) {
#pragma unroll
            for (int i = 0; i < C::I; ++i) {
                const int e = i * C::NI + JA * blkA + r;
                red[e * RS + sgA] = acc[i * JA + r];
                red[e * RS + 16 + sgA] = 0.f;
            }
            const int e = C::I * C::NI + C::O * C::NC + JA * blkA + r;
            red[e * RS + sgA] = acc[C::I * JA + r];
            red[e * RS + 16 + sgA] = 0.f;
        }
    } else {
#pragma unroll
        for (int o = 0; o < C::O; ++o) {
#pragma unroll
            for (int r = 0; r < JB; ++r) red[(C::I * C::NI + o * C::NC + JB * blkB + r) * RS + sgB] = acc[o * JB + r];
            if (blkB == 0) red[(C::I * C::NI + C::O * C::NC + C::NI + o) * RS + sgB] = acc[C::O * JB + o];
        }
    }
    __syncthreads();
    {
        const float* d3 = d3buf + (int64_t)blockIdx.x * D3_FLOATS;
        // stack row (pa, r) x column (pb, i): dW_heads[i][r] (i < K), db_heads[r] (i == ONE)
        for (int e = tid; e < 32 * 24; e += TM) {
            const int r = e / 24, i = e % 24;           // r = 4 j + hs (neuron-major)
            const int col = (r & 3) * C::NCp + (r >> 2);
            if (i < C::K || i == C::ONE) {
                float acc = 0.f;
#pragma unroll
                for (int pa = 0; pa < 3; ++pa)
#pragma unroll
                    for (int pb = 2; pb >= 0; --pb) acc += d3[(((pb * 24 + i) >> 2) * TM + pa * 32 + r) * 4 + ((pb * 24 + i) & 3)];
                if (i < C::K) prow[C::WH_T + i * C::H4 + col] = acc;
                else prow[C::B_H + col] = acc;
            }
        }
        for (int e = tid; e < C::N_SMALL; e += TM) {
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int g = 0; g < 32; g += 2) { a0 += red[e * RS + g]; a1 += red[e * RS + g + 1]; }
            const float tot = a0 + a1;
            int idx;
            if (e < C::I * C::NI) idx = C::WIN_T + (e / C::NI) * C::NIp + (e % C::NI);
            else if (e < C::I * C::NI + C::O * C::NC) { const int f = e - C::I * C::NI; idx = C::WOUT + (f / C::NC) * C::NCp + (f % C::NC); }
            else if (e < C::I * C::NI + C::O * C::NC + C::NI) idx = C::B_IN + (e - C::I * C::NI - C::O * C::NC);
            else idx = C::B_OUT + (e - C::I * C::NI - C::O * C::NC - C::NI);
            prow[idx] = tot;
        }
    }
    VB_MARK(3);
#ifdef VB_ENABLE_TRACE
    if (trace && tid == 0 && blockIdx.x < 1024) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
        trace[1024 + blockIdx.x * 4 + 1] = (long long)gt;
    }
#else
    (void)trace;
#endif
}

// ------------------------------------------------------------------------------------------------ host side
__global__ void __launch_bounds__(1024) reduce_partials_tc_kernel(const float* __restrict__ partials, int n_rows, int n, float* __restrict__ out) {
    __shared__ float sm[32][33];
    const int cx = threadIdx.x & 31, rg = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + cx;
    float a = 0.f;
    if (col < n)
        for (int r = rg; r < n_rows; r += 32) a += partials[(int64_t)r * n + col];
    sm[rg][cx] = a;
    __syncthreads();
    if (rg == 0 && col < n) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) t += sm[k][cx];
        out[col] = t;
    }
}

static int vec_flags_for(const void* x, const void* y, const void* h_a, const void* h_b, const void* h_c, const void* h_d,
                         const void* x2, const void* h_e, int T, int I, int O, int Nc) {
    auto al = [](const void* p, int bytes) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & (bytes - 1)) == 0; };
    auto width = [](int n) { return (n & 3) == 0 ? 16 : ((n & 1) == 0 ? 8 : 4); };
    (void)T;
    int f = 0;
    if (al(x, width(I)) && al(x2, width(I))) f |= 1;
    if (al(y, width(O))) f |= 2;
    if (al(h_a, width(Nc)) && al(h_b, width(Nc)) && al(h_c, width(Nc)) && al(h_d, width(Nc)) && al(h_e, width(Nc))) f |= 4;
    return f;
}

// CTAs per SM.  cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for every kernel that allocates tensor memory
// (measured: 30 KB smem, 53 registers -> "1"), so the limits are evaluated here: shared memory (+1 KB reserved per
// CTA), registers, TMEM columns (512 per SM) and the 32-CTA / 2048-thread caps.
template <class K>
static int grid_cap(K kernel, int smem, int tmem_cols, int* per_sm) {
    LaunchCfg cfg;
    int rc = cached_launch_cfg(kernel, TM, smem, &cfg);
    if (rc) return rc;
    static int smem_sm[64], regs_sm[64];
    int dev = 0;
    VB_CUDA(cudaGetDevice(&dev));
    if (!regs_sm[dev & 63]) {
        VB_CUDA(cudaDeviceGetAttribute(&smem_sm[dev & 63], cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
        VB_CUDA(cudaDeviceGetAttribute(&regs_sm[dev & 63], cudaDevAttrMaxRegistersPerMultiprocessor, dev));
    }
    const int regs_cta = ((cfg.regs + 7) & ~7) * TM;
    int n = smem_sm[dev & 63] / (smem + cfg.static_smem + 1024);
    if (regs_cta > 0 && regs_sm[dev & 63] / regs_cta < n) n = regs_sm[dev & 63] / regs_cta;
    if (512 / tmem_cols < n) n = 512 / tmem_cols;
    if (n > 2048 / TM) n = 2048 / TM;
    if (n > 32) n = 32;
    *per_sm = n < 1 ? 1 : n;
    return 0;
}

template <class C>
static int launch_fwd(const float* packed, const float* x, const float* h0, float* y, float* h_traj, int64_t B, int T, int layout, cudaStream_t stream) {
    std::lock_guard<std::mutex> lock(g_mu);
    int dev;
    bool capturing;
    int rc = upload_constants(packed, C::C_TOTAL, stream, &dev, &capturing);
    if (rc) return rc;
    int vf = vec_flags_for(x, y, h0, h_traj, nullptr, nullptr, nullptr, nullptr, T, C::I, C::O, C::NC) | layout;
    // fast variants: aligned rows and one storage order for every sequence tensor; anything else takes the generic variant
    const int all = LAY_X | LAY_Y | LAY_H;
    const bool vec = (vf & 7) == 7;
    auto kern = (vec && (layout & all) == all) ? liquid_fwd_tc_kernel<C, true, 1>
              : (vec && (layout & all) == 0)   ? liquid_fwd_tc_kernel<C, true, 0> : liquid_fwd_tc_kernel<C, false, 2>;
    int per_sm = 1;
    rc = grid_cap(kern, FwdSmem<C>::BYTES, 128, &per_sm);
    if (rc) return rc;
    int64_t n_tiles = (B + TM - 1) / TM;
    int64_t cap = (int64_t)device_sms() * per_sm;
    int grid = (int)(n_tiles < cap ? n_tiles : cap);
    if (grid < 1) grid = 1;
    kern<<<grid, TM, FwdSmem<C>::BYTES, stream>>>(packed, x, h0, y, h_traj, B, T, vf);
    VB_LAUNCH_CHECK();
    if (!capturing) VB_CUDA(cudaEventRecord(g_ev[dev & 63], stream));
    return 0;
}

template <class C, class K>
static int bwd_grid(K kern, int64_t B, int* grid_out) {
    int per_sm = 1;
    int rc = grid_cap(kern, BwdSmem<C>::BYTES, 128, &per_sm);
    if (rc) return rc;
    int64_t n_tiles = (B + TM - 1) / TM;
    int64_t cap = (int64_t)device_sms() * per_sm;
    int grid = (int)(n_tiles < cap ? n_tiles : cap);
    *grid_out = grid < 1 ? 1 : grid;
    return 0;
}

template <class C>
static int launch_bwd(const float* packed, const float* x, const float* h0, const float* h_traj, const float* dy,
                      const float* dh_traj, const float* dh_last, float* dx, float* dh0, float* dpacked,
                      int64_t B, int T, int layout, void* ws, int64_t ws_bytes, cudaStream_t stream) {
    int vf = vec_flags_for(x, dy, h0, h_traj, dh_traj, dh_last, dx, dh0, T, C::I, C::O, C::NC) | layout;
    const int all = LAY_X | LAY_Y | LAY_H | (dx ? LAY_DX : 0) | (dh_traj ? LAY_DH : 0);
    const bool vec = (vf & 7) == 7;
    auto kern = (vec && (layout & all) == all) ? liquid_bwd_tc_kernel<C, true, 1>
              : (vec && (layout & all) == 0)   ? liquid_bwd_tc_kernel<C, true, 0> : liquid_bwd_tc_kernel<C, false, 2>;
    int grid = 1;
    int rc = bwd_grid<C>(kern, B, &grid);
    if (rc) return rc;
    const int64_t need = (int64_t)grid * (C::F_TOTAL + D3_FLOATS) * (int64_t)sizeof(float);
    VB_CHECK_ARG(ws && ws_bytes >= need, VB_ERR_WORKSPACE, "vb_liquid_bwd: workspace too small (%lld bytes, need %lld)",
                 (long long)ws_bytes, (long long)need);
    std::lock_guard<std::mutex> lock(g_mu);
    int dev;
    bool capturing;
    rc = upload_constants(packed, C::C_TOTAL, stream, &dev, &capturing);
    if (rc) return rc;
    float* partials = reinterpret_cast<float*>(ws);
    float* d3buf = partials + (int64_t)grid * C::F_TOTAL;
    long long* trace = nullptr;
#ifdef VB_ENABLE_TRACE
    if (getenv("VB_TRACE")) { cudaMalloc(&trace, (1024 + 4096) * sizeof(long long)); cudaMemset(trace, 0, (1024 + 4096) * sizeof(long long)); }
#endif
    kern<<<grid, TM, BwdSmem<C>::BYTES, stream>>>(packed, x, h0, h_traj, dy, dh_traj, dh_last, dx, dh0, partials, d3buf, B, T, vf, trace);
    VB_LAUNCH_CHECK();
    if (trace) {
        long long h[64 * 16];
        cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost);
        const char* names[12] = {"top(loads issued)", "inter+mish+store_z", "sync1", "issue heads", "wait heads", "gating+store ds", "sync2", "issue dz/dW",
                                 "wait dz", "du,dx", "wait dW", "small grads"};
        for (int st = 0; st < 4; ++st) {
            fprintf(stderr, "[vb trace] step %d:", st);
            for (int k = 1; k < 12; ++k) fprintf(stderr, " %s=%lld", names[k], h[st * 16 + k] - h[st * 16 + k - 1]);
            if (st < 3) fprintf(stderr, " | to next top=%lld", h[(st + 1) * 16] - h[st * 16 + 11]);
            fprintf(stderr, "\n");
        }
        fprintf(stderr, "[vb trace] step totals:");
        for (int st = 0; st + 1 < T && st < 39; ++st) fprintf(stderr, " %lld", h[(st + 1) * 16] - h[st * 16]);
        fprintf(stderr, "\n[vb trace] prologue=%lld loop=%lld epilogue=%lld\n", h[40 * 16 + 1] - h[40 * 16], h[40 * 16 + 2] - h[40 * 16 + 1],
                h[40 * 16 + 3] - h[40 * 16 + 2]);
        {
            static long long hb[4096];
            cudaMemcpy(hb, trace + 1024, sizeof(hb), cudaMemcpyDeviceToHost);
            long long t0 = hb[0], t1 = 0;
            for (int c = 0; c < grid && c < 1024; ++c) { if (hb[c * 4] < t0) t0 = hb[c * 4]; if (hb[c * 4 + 1] > t1) t1 = hb[c * 4 + 1]; }
            int per_sm[256] = {0};
            long long sm_end[256] = {0};
            for (int c = 0; c < grid && c < 1024; ++c) { int sm = (int)hb[c * 4 + 2] & 255; per_sm[sm]++; if (hb[c * 4 + 1] - t0 > sm_end[sm]) sm_end[sm] = hb[c * 4 + 1] - t0; }
            long long e3 = 0, e4 = 0; int n3 = 0, n4 = 0; long long latest_start = 0;
            for (int c = 0; c < grid && c < 1024; ++c) if (hb[c * 4] - t0 > latest_start) latest_start = hb[c * 4] - t0;
            for (int sm = 0; sm < 256; ++sm) { if (per_sm[sm] == 3) { e3 += sm_end[sm]; n3++; } if (per_sm[sm] >= 4) { e4 += sm_end[sm]; n4++; } }
            fprintf(stderr, "[vb trace] grid %d span %lld ns, latest CTA start %lld ns; SMs with 3 CTAs: %d (avg end %lld ns), with >=4: %d (avg end %lld ns)\n",
                    grid, t1 - t0, latest_start, n3, n3 ? e3 / n3 : 0, n4, n4 ? e4 / n4 : 0);
        }
        cudaFree(trace);
    }
    if (!capturing) VB_CUDA(cudaEventRecord(g_ev[dev & 63], stream));
    reduce_partials_tc_kernel<<<(C::F_TOTAL + 31) / 32, 1024, 0, stream>>>(partials, grid, C::F_TOTAL, dpacked);
    VB_LAUNCH_CHECK();
    return 0;
}

#define VB_TC_SHAPES(X) X(4, 12, 8, 2) X(3, 12, 8, 2)

bool has_tc_path(int I, int Ni, int Nc, int O) {
#define X(a, b, c, d) if (I == a && Ni == b && Nc == c && O == d) return true;
    VB_TC_SHAPES(X)
#undef X
    return false;
}

int fwd(const float* packed, const float* x, const float* h0, float* y, float* h_traj, int64_t B, int T,
        int I, int Ni, int Nc, int O, int layout, cudaStream_t stream) {
#define X(a, b, c, d) if (I == a && Ni == b && Nc == c && O == d) return launch_fwd<Cfg<a, b, c, d>>(packed, x, h0, y, h_traj, B, T, layout, stream);
    VB_TC_SHAPES(X)
#undef X
    set_error("no tensor-core kernel for (%d,%d,%d,%d)", I, Ni, Nc, O);
    return VB_ERR_BAD_DIMS;
}

int bwd(const float* packed, const float* x, const float* h0, const float* h_traj, const float* dy, const float* dh_traj,
        const float* dh_last, float* dx, float* dh0, float* dpacked, int64_t B, int T, int I, int Ni, int Nc, int O,
        int layout, void* ws, int64_t ws_bytes, cudaStream_t stream) {
#define X(a, b, c, d)                                            \
    if (I == a && Ni == b && Nc == c && O == d)                  \
        return launch_bwd<Cfg<a, b, c, d>>(packed, x, h0, h_traj, dy, dh_traj, dh_last, dx, dh0, dpacked, B, T, layout, ws, ws_bytes, stream);
    VB_TC_SHAPES(X)
#undef X
    set_error("no tensor-core kernel for (%d,%d,%d,%d)", I, Ni, Nc, O);
    return VB_ERR_BAD_DIMS;
}

int64_t bwd_workspace_bytes(int I, int Ni, int Nc, int O) {
    LiquidLayout L(I, Ni, Nc, O);
    int sms = device_sms();
    if (sms <= 0) sms = 148;
    return (int64_t)sms * 4 * (L.f_total + D3_FLOATS) * (int64_t)sizeof(float);
}

}  // namespace tc
}  // namespace vb
