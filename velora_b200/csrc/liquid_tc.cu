// Tensor-core (tcgen05 / TMEM) kernels for the 20-neuron class of LiquidNCPNetwork.
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

namespace vb {
namespace tc {

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

constexpr int TM = 128;                  // samples per CTA = threads per CTA
constexpr int PLANE = TM * 16;           // bytes per plane
constexpr int KP = 32;                   // padded feature count of the z / ds operand vectors (4 planes per part)
constexpr int kFlush = 8;                // steps between flushes of the truncating dW accumulator

template <int I_, int NI_, int NC_, int O_>
struct Cfg {
    static constexpr int I = I_, NI = NI_, NC = NC_, O = O_;
    static constexpr int NIp = ceil4(NI), NCp = ceil4(NC), Op = ceil4(O), K = NI + NC, H4 = 4 * NCp;
    static constexpr int Ip = ceil4(I), Kp = ceil4(K);
    static constexpr int WIN_T = 0, B_IN = WIN_T + I * NIp, WH_T = B_IN + NIp, B_H = WH_T + K * H4, WOUT = B_H + H4,
                         B_OUT = WOUT + O * NCp, F_TOTAL = B_OUT + Op, WIN = F_TOTAL, WH = WIN + NI * Ip, C_TOTAL = WH + H4 * Kp;
    static constexpr int ONE = K;        // index of the constant-one feature (bias row)
    static_assert(H4 == 32, "tensor-core path: 4 * Nc must be 32 (Nc = 8)");
    static_assert(K + 1 <= 24, "tensor-core path: Ni + Nc + 1 must fit three 8-feature planes");
    static constexpr int N_SMALL = I * NI + O * NC + NI + O;   // register-accumulated small gradients
    static_assert(N_SMALL <= 96, "tensor-core backward keeps the small layers' gradients in registers: I*Ni + O*Nc + Ni + O <= 96");
    static_assert(C_TOTAL <= kConstFloats, "constant-bank budget");
};

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16 |
           (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32 | (uint64_t)1 << 46;
}
// c = f32, a = b = bf16
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
          "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// x0, x1 -> three packed bf16x2 words (low half = x0): x = p1 + p2 + p3 to 24 bits
__device__ __forceinline__ void split3(float x0, float x1, uint32_t& p1, uint32_t& p2, uint32_t& p3) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p1) : "f"(x1), "f"(x0));
    float r0 = x0 - __uint_as_float(p1 << 16), r1 = x1 - __uint_as_float(p1 & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p2) : "f"(r1), "f"(r0));
    float q0 = r0 - __uint_as_float(p2 << 16), q1 = r1 - __uint_as_float(p2 & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p3) : "f"(q1), "f"(q0));
}

// eight consecutive features of this thread's sample -> one 16-byte chunk in each of the three part planes
__device__ __forceinline__ void store_chunk3(unsigned char* planes, int chunk, int row, const float (&f)[8]) {
    uint4 a, b, c;
    split3(f[0], f[1], a.x, b.x, c.x);
    split3(f[2], f[3], a.y, b.y, c.y);
    split3(f[4], f[5], a.z, b.z, c.z);
    split3(f[6], f[7], a.w, b.w, c.w);
    uint4* p = reinterpret_cast<uint4*>(planes);
    p[(0 * 4 + chunk) * TM + row] = a;
    p[(1 * 4 + chunk) * TM + row] = b;
    p[(2 * 4 + chunk) * TM + row] = c;
}

// product order for one per-step contraction: small terms first, b1*b1 last (the accumulator truncates)
// D[128 x 32] = A(parts, K-major planes, K = 32) * B(parts)^T : 12 MMAs, issued by one thread
__device__ __forceinline__ void issue_contraction(uint32_t d_tmem, uint32_t a_planes, uint32_t b_planes) {
    constexpr uint32_t idesc = make_idesc(128, 32, 0, 0);
#pragma unroll
    for (int p = 0; p < 6; ++p) {
        const int pa = (p == 0) ? 1 : (p == 1) ? 0 : (p == 2) ? 2 : (p == 3) ? 0 : (p == 4) ? 1 : 0;
        const int pb = (p == 0) ? 1 : (p == 1) ? 2 : (p == 2) ? 0 : (p == 3) ? 1 : (p == 4) ? 0 : 0;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            uint64_t ad = make_desc(a_planes + (pa * 4 + 2 * ks) * PLANE, PLANE, 128);
            uint64_t bd = make_desc(b_planes + (pb * 4 + 2 * ks) * (32 * 16), 32 * 16, 128);
            mma_bf16(d_tmem, ad, bd, idesc, (p | ks) != 0);
        }
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

// z~ = [a (NI), h (NC), 1, 0...] -> chunks 0..2 of the three part planes
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
        store_chunk3(zP, c, row, f);
    }
}

// ================================================================================================ forward
template <class C>
struct FwdSmem {
    static constexpr int Z = 0;                                  // 12 planes
    static constexpr int WB = Z + 12 * PLANE;                    // 3 parts x 4 chunks x 32 rows x 16 B
    static constexpr int BAR = WB + 3 * 4 * 32 * 16;
    static constexpr int BYTES = BAR + 64;
};

template <class C>
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
    const bool vx = vec_flags & 1, vy = vec_flags & 2, vh = vec_flags & 4;

    // heads operand: B[n][k] = W_heads[k][n] for k < K, bias for k == ONE
    build_weight_planes(wB, [&](int n, int k) {
        return k < C::K ? __ldg(packed + C::WH_T + k * C::H4 + n) : (k == C::ONE ? __ldg(packed + C::B_H + n) : 0.f);
    });
    {   // zero the never-written fourth chunk of the z planes
        uint4* p = reinterpret_cast<uint4*>(zP);
        const uint4 z4 = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int part = 0; part < 3; ++part) p[(part * 4 + 3) * TM + tid] = z4;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(tmem_slot)) : "memory");
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
        if (valid) ldg_row<C::I>(xn, x + (b * T) * C::I, vx);
        for (int t = 0; t < T; ++t) {
            float xr[C::I];
#pragma unroll
            for (int i = 0; i < C::I; ++i) xr[i] = xn[i];
            if (valid && t + 1 < T) ldg_row<C::I>(xn, x + (b * T + t + 1) * C::I, vx);   // prefetch the next step's input
            float a[C::NI];
            {
                float u[C::NI];
                inter_layer<C>(xr, u);
#pragma unroll
                for (int j = 0; j < C::NI; ++j) a[j] = mish_f(u[j]);
            }
            store_z<C>(zP, tid, a, h);
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                issue_contraction(tmem, z_addr, w_addr);
                mma_commit(bar1);
            }
            mbar_wait(bar1, ph1);
            ph1 ^= 1;
            tc_fence_after();
            float s[32];
            tmem_ld32(taddr, s);
            float m[C::NC];
#pragma unroll
            for (int j = 0; j < C::NC; ++j) {
                float tg = tanh_f(s[0 * C::NCp + j]), th = tanh_f(s[1 * C::NCp + j]), gate = sigmoid_f(s[2 * C::NCp + j]);
                float hn = fmaf(gate, th - tg, tg);
                h[j] = hn;
                m[j] = mish_f(s[3 * C::NCp + j] + hn);
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
                stg_row<C::O>(y + (b * T + t) * C::O, yo, vy);
                stg_row<C::NC>(h_traj + (b * T + t) * C::NC, h, vh);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

// ================================================================================================ backward
template <class C>
struct BwdSmem {
    static constexpr int Z = 0;                                  // z~ parts: 12 planes
    static constexpr int DS = Z + 12 * PLANE;                    // ds parts: 12 planes (MN-major M = 128 reads 4 planes past: WB/WT)
    static constexpr int WB = DS + 12 * PLANE;                   // heads operand  (6 KB)
    static constexpr int WT = WB + 3 * 4 * 32 * 16;              // dz operand     (6 KB): B[i][r] = W_heads[i][r]
    static constexpr int BAR = WT + 3 * 4 * 32 * 16;
    static constexpr int BYTES = BAR + 64;
    static_assert(WT + 3 * 4 * 32 * 16 - DS >= 16 * PLANE, "M = 128 stack must stay inside the CTA's shared memory");
};

constexpr int D3_FLOATS = TM * 64;   // per-CTA flush buffer: two 128 x 32 accumulators

template <class C>
__global__ void __launch_bounds__(TM, 3) liquid_bwd_tc_kernel(
    const float* __restrict__ packed, const float* __restrict__ x, const float* __restrict__ h0, const float* __restrict__ h_traj,
    const float* __restrict__ dy, const float* __restrict__ dh_traj, const float* __restrict__ dh_last,
    float* __restrict__ dx, float* __restrict__ dh0, float* __restrict__ partials, float* __restrict__ d3buf,
    int64_t B, int T, int vec_flags) {
    using S = BwdSmem<C>;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* zP = smem + S::Z;
    unsigned char* dsP = smem + S::DS;
    unsigned char* wB = smem + S::WB;
    unsigned char* wT = smem + S::WT;
    uint64_t* bar1 = reinterpret_cast<uint64_t*>(smem + S::BAR);
    uint64_t* bar2 = bar1 + 1;
    uint64_t* bar3 = bar1 + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::BAR + 32);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool vx = vec_flags & 1, vy = vec_flags & 2, vh = vec_flags & 4;

    build_weight_planes(wB, [&](int n, int k) {
        return k < C::K ? __ldg(packed + C::WH_T + k * C::H4 + n) : (k == C::ONE ? __ldg(packed + C::B_H + n) : 0.f);
    });
    // dz operand: B[i][r] = W_heads[i][r] (weight of input i to (head, neuron) r); rows i >= K are zero
    build_weight_planes(wT, [&](int i, int r) { return i < C::K ? __ldg(packed + C::WH_T + i * C::H4 + r) : 0.f; });
    {
        uint4* p = reinterpret_cast<uint4*>(zP);
        const uint4 z4 = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int part = 0; part < 3; ++part) p[(part * 4 + 3) * TM + tid] = z4;
    }
    float* d3row = d3buf + (int64_t)blockIdx.x * D3_FLOATS + tid * 64;   // this thread's rows of the two dW accumulators
#pragma unroll
    for (int q = 0; q < 16; ++q) reinterpret_cast<float4*>(d3row)[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(bar1, 1);
        mbar_init(bar2, 1);
        mbar_init(bar3, 2);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t z_addr = smem_u32(zP), ds_addr = smem_u32(dsP), wb_addr = smem_u32(wB), wt_addr = smem_u32(wT);
    uint32_t ph1 = 0, ph2 = 0, ph3 = 0;
    bool wgrad_pending = false;   // a dW batch whose completion has not been waited for yet
    int since_flush = 0;          // dW batches accumulated in TMEM since the last flush (0 => next batch overwrites)

    // gradients of the small layers: per-sample register accumulators, reduced across the CTA once at the end
    float g_win[C::I][C::NI], g_wout[C::O][C::NC], g_bin[C::NI], g_bout[C::O];
#pragma unroll
    for (int i = 0; i < C::I; ++i)
#pragma unroll
        for (int j = 0; j < C::NI; ++j) g_win[i][j] = 0.f;
#pragma unroll
    for (int o = 0; o < C::O; ++o) {
        g_bout[o] = 0.f;
#pragma unroll
        for (int j = 0; j < C::NC; ++j) g_wout[o][j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < C::NI; ++j) g_bin[j] = 0.f;

    auto flush_dw = [&]() {   // all threads; requires the last dW batch to be complete
        tc_fence_after();
        float v[32];
        tmem_ld32(taddr + 64, v);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float4 a = reinterpret_cast<float4*>(d3row)[q];
            a.x += v[4 * q]; a.y += v[4 * q + 1]; a.z += v[4 * q + 2]; a.w += v[4 * q + 3];
            reinterpret_cast<float4*>(d3row)[q] = a;
        }
        tmem_ld32(taddr + 96, v);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float4 a = reinterpret_cast<float4*>(d3row)[8 + q];
            a.x += v[4 * q]; a.y += v[4 * q + 1]; a.z += v[4 * q + 2]; a.w += v[4 * q + 3];
            reinterpret_cast<float4*>(d3row)[8 + q] = a;
        }
        since_flush = 0;
    };

    const int64_t n_tiles = (B + TM - 1) / TM;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t b = tile * TM + tid;
        const bool valid = b < B;
        float dhc[C::NC];
#pragma unroll
        for (int j = 0; j < C::NC; ++j) dhc[j] = 0.f;
        if (valid && dh_last) ldg_row<C::NC>(dhc, dh_last + b * C::NC, vh);
        for (int t = T - 1; t >= 0; --t) {
            // ---------------- loads of this step (one to two full sectors per thread)
            float xr[C::I], dyr[C::O], hp[C::NC], dhe[C::NC];
#pragma unroll
            for (int i = 0; i < C::I; ++i) xr[i] = 0.f;
#pragma unroll
            for (int o = 0; o < C::O; ++o) dyr[o] = 0.f;
#pragma unroll
            for (int j = 0; j < C::NC; ++j) { hp[j] = 0.f; dhe[j] = 0.f; }
            if (valid) {
                ldg_row<C::I>(xr, x + (b * T + t) * C::I, vx);
                ldg_row<C::O>(dyr, dy + (b * T + t) * C::O, vy);
                if (t > 0) ldg_row<C::NC>(hp, h_traj + (b * T + t - 1) * C::NC, vh);
                else if (h0) ldg_row<C::NC>(hp, h0 + b * C::NC, vh);
                if (dh_traj) ldg_row<C::NC>(dhe, dh_traj + (b * T + t) * C::NC, vh);
            }
            // ---------------- recompute: inter layer on CUDA cores, heads on the tensor cores
            float a[C::NI], dmu[C::NI];
            {
                float u[C::NI];
                inter_layer<C>(xr, u);
#pragma unroll
                for (int j = 0; j < C::NI; ++j) a[j] = mish_fwd_grad(u[j], dmu[j]);
            }
            if (wgrad_pending) {   // the previous step's dW MMAs still read the z / ds planes
                mbar_wait(bar3, ph3);
                ph3 ^= 1;
                wgrad_pending = false;
                if (since_flush >= kFlush) flush_dw();
            }
            store_z<C>(zP, tid, a, hp);
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                issue_contraction(tmem + 0, z_addr, wb_addr);
                mma_commit(bar1);
            }
            mbar_wait(bar1, ph1);
            ph1 ^= 1;
            tc_fence_after();
            float s[32];
            tmem_ld32(taddr + 0, s);
            // ---------------- gating derivatives; ds -> planes
            float m[C::NC];
#pragma unroll
            for (int j = 0; j < C::NC; ++j) {
                float tg = tanh_f(s[0 * C::NCp + j]), th = tanh_f(s[1 * C::NCp + j]), gate = sigmoid_f(s[2 * C::NCp + j]);
                float hn = fmaf(gate, th - tg, tg);
                float dmc;
                m[j] = mish_fwd_grad(s[3 * C::NCp + j] + hn, dmc);
                float dm = 0.f;
#pragma unroll
                for (int o = 0; o < C::O; ++o) dm = fmaf(dyr[o], c_p[C::WOUT + o * C::NCp + j], dm);
                float dc = dm * dmc;
                float dhn = dc + dhc[j] + dhe[j];
                float omg = 1.0f - gate;
                s[0 * C::NCp + j] = dhn * omg * (1.0f - tg * tg);
                s[1 * C::NCp + j] = dhn * gate * (1.0f - th * th);
                s[2 * C::NCp + j] = dhn * (th - tg) * gate * omg;
                s[3 * C::NCp + j] = dc;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = s[8 * c + e];
                store_chunk3(dsP, c, tid, f);
            }
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
            if (lane == 0) {
                tc_fence_after();
                if (warp == 1) {            // dz = ds W : K-major ds planes x dz operand
                    issue_contraction(tmem + 32, ds_addr, wt_addr);
                    mma_commit(bar2);
                } else if (warp >= 2) {     // dW += ds^T z~ : MN-major reads of the same planes, K = 16 samples per MMA
                    constexpr uint32_t idesc3 = make_idesc(128, 32, 1, 1);
                    const uint32_t dcol = tmem + (warp == 2 ? 64 : 96);
                    const int ks0 = (warp == 2) ? 0 : 4;
#pragma unroll
                    for (int zp = 2; zp >= 0; --zp)          // small parts first
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) {
                            const int ks = ks0 + k4;
                            uint64_t ad = make_desc(ds_addr + ks * 256, 128, PLANE);                    // M = [b1|b2|b3|x] stack
                            uint64_t bd = make_desc(z_addr + zp * 4 * PLANE + ks * 256, 128, PLANE);    // N = 32 features of part zp
                            mma_bf16(dcol, ad, bd, idesc3, (since_flush != 0) || zp != 2 || k4 != 0);
                        }
                    mma_commit(bar3);
                }
            }
            wgrad_pending = true;
            since_flush += 1;
            // small-layer gradients that do not need dz (overlap with the tensor cores)
#pragma unroll
            for (int o = 0; o < C::O; ++o) {
                g_bout[o] += dyr[o];
#pragma unroll
                for (int j = 0; j < C::NC; ++j) g_wout[o][j] = fmaf(dyr[o], m[j], g_wout[o][j]);
            }
            mbar_wait(bar2, ph2);
            ph2 ^= 1;
            tc_fence_after();
            float dz[32];
            tmem_ld32(taddr + 32, dz);
            float du[C::NI];
#pragma unroll
            for (int j = 0; j < C::NI; ++j) {
                du[j] = dz[j] * dmu[j];
                g_bin[j] += du[j];
            }
#pragma unroll
            for (int j = 0; j < C::NC; ++j) dhc[j] = dz[C::NI + j];
            float dxr[C::I];
#pragma unroll
            for (int i = 0; i < C::I; ++i) dxr[i] = 0.f;
#pragma unroll
            for (int j = 0; j < C::NI; ++j)
#pragma unroll
                for (int i = 0; i < C::I; ++i) {
                    dxr[i] = fmaf(du[j], c_p[C::WIN + j * C::Ip + i], dxr[i]);
                    g_win[i][j] = fmaf(xr[i], du[j], g_win[i][j]);
                }
            if (valid && dx) stg_row<C::I>(dx + (b * T + t) * C::I, dxr, vx);
        }
        if (valid && dh0) stg_row<C::NC>(dh0 + b * C::NC, dhc, vh);
    }
    if (wgrad_pending) {
        mbar_wait(bar3, ph3);
        ph3 ^= 1;
    }
    if (since_flush > 0) flush_dw();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");

    // ---------------- per-CTA partial gradient row (F layout)
    float* prow = partials + (int64_t)blockIdx.x * C::F_TOTAL;
    // heads: dW[i][r] = sum over the three stacked part rows of both accumulators; column ONE is the bias gradient
    __threadfence_block();
    __syncthreads();
    {
        const float* d3 = d3buf + (int64_t)blockIdx.x * D3_FLOATS;
        for (int e = tid; e < 32 * 32; e += TM) {
            const int r = e >> 5, i = e & 31;
            float acc = 0.f;
#pragma unroll
            for (int part = 0; part < 3; ++part) {
                const float* row = d3 + (part * 32 + r) * 64;
                acc += row[i] + row[32 + i];
            }
            if (i < C::K) prow[C::WH_T + i * C::H4 + r] = acc;
            else if (i == C::ONE) prow[C::B_H + r] = acc;
        }
    }
    // small layers: reduce the per-sample registers over the CTA through shared memory (planes are free now)
    float* red = reinterpret_cast<float*>(smem);   // [N_SMALL][TM + 1]
    {
        int k = 0;
#pragma unroll
        for (int i = 0; i < C::I; ++i)
#pragma unroll
            for (int j = 0; j < C::NI; ++j) red[(k++) * (TM + 1) + tid] = g_win[i][j];
#pragma unroll
        for (int o = 0; o < C::O; ++o)
#pragma unroll
            for (int j = 0; j < C::NC; ++j) red[(k++) * (TM + 1) + tid] = g_wout[o][j];
#pragma unroll
        for (int j = 0; j < C::NI; ++j) red[(k++) * (TM + 1) + tid] = g_bin[j];
#pragma unroll
        for (int o = 0; o < C::O; ++o) red[(k++) * (TM + 1) + tid] = g_bout[o];
    }
    __syncthreads();
    for (int k = tid; k < C::N_SMALL; k += TM) {
        float acc = 0.f;
        for (int s2 = 0; s2 < TM; ++s2) acc += red[k * (TM + 1) + s2];
        int idx;
        if (k < C::I * C::NI) idx = C::WIN_T + (k / C::NI) * C::NIp + (k % C::NI);
        else if (k < C::I * C::NI + C::O * C::NC) { int q = k - C::I * C::NI; idx = C::WOUT + (q / C::NC) * C::NCp + (q % C::NC); }
        else if (k < C::I * C::NI + C::O * C::NC + C::NI) idx = C::B_IN + (k - C::I * C::NI - C::O * C::NC);
        else idx = C::B_OUT + (k - C::I * C::NI - C::O * C::NC - C::NI);
        prow[idx] = acc;
    }
    // padding entries of the F layout
    for (int idx = tid; idx < C::F_TOTAL; idx += TM) {
        bool pad = false;
        if (idx >= C::B_IN && idx < C::WH_T) pad = (idx - C::B_IN) >= C::NI;
        else if (idx < C::B_IN) pad = (idx % C::NIp) >= C::NI;
        else if (idx >= C::WOUT && idx < C::B_OUT) pad = ((idx - C::WOUT) % C::NCp) >= C::NC;
        else if (idx >= C::B_OUT) pad = (idx - C::B_OUT) >= C::O;
        if (pad) prow[idx] = 0.f;
    }
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
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    VB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaFuncAttributes fa;
    VB_CUDA(cudaFuncGetAttributes(&fa, kernel));
    int dev = 0, smem_sm = 0, regs_sm = 0;
    VB_CUDA(cudaGetDevice(&dev));
    VB_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
    VB_CUDA(cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev));
    const int regs_cta = ((fa.numRegs + 7) & ~7) * TM;
    int n = smem_sm / (smem + (int)fa.sharedSizeBytes + 1024);
    if (regs_cta > 0 && regs_sm / regs_cta < n) n = regs_sm / regs_cta;
    if (512 / tmem_cols < n) n = 512 / tmem_cols;
    if (n > 2048 / TM) n = 2048 / TM;
    if (n > 32) n = 32;
    if (getenv("VB_DEBUG"))
        fprintf(stderr, "[vb] tc kernel: smem=%d regs=%d tmem=%d -> %d CTAs/SM\n", smem, fa.numRegs, tmem_cols, n);
    *per_sm = n < 1 ? 1 : n;
    return 0;
}

template <class C>
static int launch_fwd(const float* packed, const float* x, const float* h0, float* y, float* h_traj, int64_t B, int T, cudaStream_t stream) {
    std::lock_guard<std::mutex> lock(g_mu);
    int dev;
    bool capturing;
    int rc = upload_constants(packed, C::C_TOTAL, stream, &dev, &capturing);
    if (rc) return rc;
    int per_sm = 1;
    rc = grid_cap(liquid_fwd_tc_kernel<C>, FwdSmem<C>::BYTES, 32, &per_sm);
    if (rc) return rc;
    int64_t n_tiles = (B + TM - 1) / TM;
    int64_t cap = (int64_t)device_sms() * per_sm;
    int grid = (int)(n_tiles < cap ? n_tiles : cap);
    if (grid < 1) grid = 1;
    int vf = vec_flags_for(x, y, h0, h_traj, nullptr, nullptr, nullptr, nullptr, T, C::I, C::O, C::NC);
    liquid_fwd_tc_kernel<C><<<grid, TM, FwdSmem<C>::BYTES, stream>>>(packed, x, h0, y, h_traj, B, T, vf);
    VB_LAUNCH_CHECK();
    if (!capturing) VB_CUDA(cudaEventRecord(g_ev[dev & 63], stream));
    return 0;
}

template <class C>
static int bwd_grid(int64_t B, int* grid_out) {
    int per_sm = 1;
    int rc = grid_cap(liquid_bwd_tc_kernel<C>, BwdSmem<C>::BYTES, 128, &per_sm);
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
                      int64_t B, int T, void* ws, int64_t ws_bytes, cudaStream_t stream) {
    int grid = 1;
    int rc = bwd_grid<C>(B, &grid);
    if (rc) return rc;
    const int64_t need = (int64_t)grid * (C::F_TOTAL + D3_FLOATS) * (int64_t)sizeof(float);
    VB_CHECK_ARG(ws && ws_bytes >= need, VB_ERR_WORKSPACE, "vb_liquid_bwd: workspace too small (%lld bytes, need %lld)",
                 (long long)ws_bytes, (long long)need);
    std::lock_guard<std::mutex> lock(g_mu);
    int dev;
    bool capturing;
    rc = upload_constants(packed, C::C_TOTAL, stream, &dev, &capturing);
    if (rc) return rc;
    int vf = vec_flags_for(x, dy, h0, h_traj, dh_traj, dh_last, dx, dh0, T, C::I, C::O, C::NC);
    float* partials = reinterpret_cast<float*>(ws);
    float* d3buf = partials + (int64_t)grid * C::F_TOTAL;
    liquid_bwd_tc_kernel<C><<<grid, TM, BwdSmem<C>::BYTES, stream>>>(packed, x, h0, h_traj, dy, dh_traj, dh_last, dx, dh0, partials,
                                                                     d3buf, B, T, vf);
    VB_LAUNCH_CHECK();
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
        int I, int Ni, int Nc, int O, cudaStream_t stream) {
#define X(a, b, c, d) if (I == a && Ni == b && Nc == c && O == d) return launch_fwd<Cfg<a, b, c, d>>(packed, x, h0, y, h_traj, B, T, stream);
    VB_TC_SHAPES(X)
#undef X
    set_error("no tensor-core kernel for (%d,%d,%d,%d)", I, Ni, Nc, O);
    return VB_ERR_BAD_DIMS;
}

int bwd(const float* packed, const float* x, const float* h0, const float* h_traj, const float* dy, const float* dh_traj,
        const float* dh_last, float* dx, float* dh0, float* dpacked, int64_t B, int T, int I, int Ni, int Nc, int O,
        void* ws, int64_t ws_bytes, cudaStream_t stream) {
#define X(a, b, c, d)                                            \
    if (I == a && Ni == b && Nc == c && O == d)                  \
        return launch_bwd<Cfg<a, b, c, d>>(packed, x, h0, h_traj, dy, dh_traj, dh_last, dx, dh0, dpacked, B, T, ws, ws_bytes, stream);
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
