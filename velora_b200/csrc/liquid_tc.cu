// Tensor-core (tcgen05 / TMEM) kernels for the 20-neuron class of LiquidNCPNetwork (12 inter + 8 command neurons).
//
// Round-2 design (profiles/r02): the round-1 backward recomputed the five heads on the tensor cores, ran four block-wide
// phases per step with ONE thread per sample (16 warps per SM, 128 registers, spills) and was bound by its own dependency
// chain (issue slots 44 % busy, no eligible warp 56 % of the cycles) while HBM sat 87 % idle.  Now
//   * the FORWARD saves, per sample and step, the four gate factors of every command neuron
//         A = (1 - gate)(1 - tanh(s_g)^2),  Bc = gate (1 - tanh(s_h)^2),  Cc = (tanh(s_h) - tanh(s_g)) gate (1 - gate),  mish'(c)
//     and m = mish(c): 40 floats = 160 B per sample-step, written as per-tile quad arrays so that every warp store is 512
//     contiguous bytes.  ds = dh' * (A, Bc, Cc) and dc = dm * mish'(c) are then a handful of FMAs: the backward has no heads
//     MMAs (12 of 32 gone), no tanh / sigmoid / Mish of the command layer and one block-wide barrier per step;
//   * the BACKWARD runs TWO threads per sample (256-thread CTAs, warps w and w + 4 address the same TMEM lanes): role A
//     recomputes the inter layer (a = mish(W_in x), mish'), builds the a-part of z~, and after the dz MMA does du, dx and the
//     dW_in reduction; role B loads the saved factors, builds ds and the h-part of z~ and carries dh.  32 warps per SM at
//     <= 64 registers, nothing spilled;
//   * the inter / motor weights are staged per CTA in SHARED memory by bulk (TMA) copies on an mbarrier and read as broadcast
//     LDS.128 - the process-global __constant__ bank, its mutex and its event are gone, so calls on different streams (or
//     graphs replayed concurrently) are independent.
//
// What stays from round 1 (measured there, profiles/microbench): an M = 128, K = 16 MMA is bounded by the shared-memory fetch
// of its operands (~4 KB A + B), so the number / size of operand fetches per step is what matters; fp32 accuracy comes from
// the bf16x3 split (x = b1 + b2 + b3 to 24 bits, six significant part products, smallest first because the TMEM accumulator
// truncates between MMAs; the long-running dW accumulator is flushed to fp32 every kFlush steps); operand "planes"
// [8-feature chunk][128 rows][16 B] are the canonical no-swizzle layout both K-major (dz = ds W) and MN-major (dW = ds^T z~).
//
// Replaces cell.py:147-169, ncp.py:168-171, sparse.py:81 of the reference and their autograd backward.
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace vb {
namespace tc {
using namespace tcx;   // PTX helpers, split3, store_chunk3, TM, PLANE (tc_common.cuh)

// Row index of (sample b, step t) in a [B,T,F] tensor stored batch-major (row b*T + t) or time-major (row t*B + b).
// MODE 0: every sequence tensor batch-major, 1: every one time-major, 2: per-tensor at run time.
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

constexpr int kFlush = 11;               // steps between flushes of the truncating dW accumulator (T = 32: three flushes)

template <int I_, int NI_, int NC_, int O_>
struct Cfg {
    static constexpr int I = I_, NI = NI_, NC = NC_, O = O_;
    static constexpr int NIp = ceil4(NI), NCp = ceil4(NC), Op = ceil4(O), K = NI + NC, H4 = 4 * NCp;
    static constexpr int Ip = ceil4(I), Kp = ceil4(K);
    // packed block (include/velora_b200.h)
    static constexpr int WIN_T = 0, B_IN = WIN_T + I * NIp, WH_T = B_IN + NIp, B_H = WH_T + K * H4, WOUT = B_H + H4,
                         B_OUT = WOUT + O * NCp, F_TOTAL = B_OUT + Op, WIN = F_TOTAL, WH = WIN + NI * Ip;
    // z~ = [a (NI) | h (NC) | 1 at ONE | 0..] : 24 features = three 8-feature planes per bf16 part
    static constexpr int ONE = K;
    static_assert(NI == 12 && NC == 8, "tensor-core path: 12 inter and 8 command neurons (4 * Nc = 32 head outputs, z~ in three planes)");
    // small-layer weights staged in shared memory (floats): three contiguous pieces of the packed block
    static constexpr int S_IN = 0, S_IN_N = I * NIp + NIp;                    // win_t [I][NIp] + b_in [NIp]
    static constexpr int S_OUT = S_IN + S_IN_N, S_OUT_N = O * NCp + Op;       // wout [O][NCp] + b_out [Op]
    static constexpr int S_WIN = S_OUT + S_OUT_N, S_WIN_N = NI * Ip;          // win [NI][Ip]
    static constexpr int S_TOTAL = S_WIN + S_WIN_N;
    // saved forward factors: one float4 (A, Bc, Cc, mish'(c)) per command neuron + m in NC / 4 float4
    // saved forward factors per sample-step: per pair of command neurons (2c, 2c + 1) the float4s (A0, A1, Bc0, Bc1) and (Cc0, Cc1, G0, G1)
    // [G = mish'(c)], then m = mish(c) in NC / 4 float4s.  (Saving the raw tanh / gate values instead - 128 B - and forming the factors in the
    // backward was measured in round 2: forward 0.104 -> 0.094 ms but backward 0.203 -> 0.226 ms, profiles/r02/README.md.)
    static constexpr int NQ = NC + NC / 4;
    static constexpr int N_SMALL = I * NI + O * NC + NI + O;                  // dW_in, dW_out, db_in, db_out entries
};

// ------------------------------------------------------------------------------------------------ small helpers
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

// streaming 128-bit load that does not allocate in L1 (the saved block is read exactly once)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// Stage the three small-layer pieces of the packed block in shared memory: bulk (TMA) copies completing on an mbarrier when the
// block is 16-byte aligned (torch allocations are), plain loads otherwise.  Called by all threads; returns after the data is visible.
template <class C>
__device__ __forceinline__ void stage_small_weights(float* sw, const float* __restrict__ packed, uint64_t* bar) {
    const bool aligned = (reinterpret_cast<uintptr_t>(packed) & 15) == 0;
    if (aligned) {
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            constexpr uint32_t bytes = C::S_TOTAL * 4;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(sw + C::S_IN)), "l"(packed + C::WIN_T), "r"((uint32_t)C::S_IN_N * 4), "r"(smem_u32(bar)) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(sw + C::S_OUT)), "l"(packed + C::WOUT), "r"((uint32_t)C::S_OUT_N * 4), "r"(smem_u32(bar)) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(sw + C::S_WIN)), "l"(packed + C::WIN), "r"((uint32_t)C::S_WIN_N * 4), "r"(smem_u32(bar)) : "memory");
        }
        __syncthreads();          // the barrier is initialised before anybody polls it
        mbar_wait(bar, 0);
    } else {
        for (int k = threadIdx.x; k < C::S_TOTAL; k += blockDim.x) {
            const int src = k < C::S_OUT ? C::WIN_T + k : (k < C::S_WIN ? C::WOUT + (k - C::S_OUT) : C::WIN + (k - C::S_WIN));
            sw[k] = __ldg(packed + src);
        }
        __syncthreads();
    }
}

// u = W_in x + b_in with the weights read as broadcast LDS.128 (all lanes read the same address)
template <class C>
__device__ __forceinline__ void inter_layer(const float* sw, const float (&x)[C::I], float (&u)[C::NI]) {
    const float4* w4 = reinterpret_cast<const float4*>(sw + C::S_IN);
#pragma unroll
    for (int q = 0; q < C::NI / 4; ++q) {
        float4 b = w4[C::I * (C::NIp / 4) + q];
        u[4 * q] = b.x; u[4 * q + 1] = b.y; u[4 * q + 2] = b.z; u[4 * q + 3] = b.w;
    }
#pragma unroll
    for (int i = 0; i < C::I; ++i)
#pragma unroll
        for (int q = 0; q < C::NI / 4; ++q) {
            float4 w = w4[i * (C::NIp / 4) + q];
            u[4 * q] = fmaf(x[i], w.x, u[4 * q]);
            u[4 * q + 1] = fmaf(x[i], w.y, u[4 * q + 1]);
            u[4 * q + 2] = fmaf(x[i], w.z, u[4 * q + 2]);
            u[4 * q + 3] = fmaf(x[i], w.w, u[4 * q + 3]);
        }
}
// the same on packed pairs: u2[k] = (u[2k], u[2k + 1]); one LDS.128 brings two weight pairs, one FFMA2 does two neurons
template <class C>
__device__ __forceinline__ void inter_layer_p(const float* sw, const float (&x)[C::I], f2 (&u)[C::NI / 2]) {
    const ulonglong2* w2 = reinterpret_cast<const ulonglong2*>(sw + C::S_IN);
#pragma unroll
    for (int q = 0; q < C::NI / 4; ++q) {
        const ulonglong2 b = w2[C::I * (C::NIp / 4) + q];
        u[2 * q] = b.x; u[2 * q + 1] = b.y;
    }
#pragma unroll
    for (int i = 0; i < C::I; ++i) {
        const f2 xx = bc2(x[i]);
#pragma unroll
        for (int q = 0; q < C::NI / 4; ++q) {
            const ulonglong2 w = w2[i * (C::NIp / 4) + q];
            u[2 * q] = fma2(xx, w.x, u[2 * q]);
            u[2 * q + 1] = fma2(xx, w.y, u[2 * q + 1]);
        }
    }
}

// z~ = [a (NI), h (NC), 1, 0..] -> chunks 0..2 of the three part planes (forward: one thread owns the whole row)
template <class C>
__device__ __forceinline__ void store_z_p(unsigned char* zP, int row, const f2 (&a)[C::NI / 2], const f2 (&h)[C::NC / 2]) {
    static_assert(C::NI == 12 && C::NC == 8, "chunk 0 = a[0..8), chunk 1 = a[8..12) | h[0..4), chunk 2 = h[4..8) | 1 | 0");
    store_chunk3_p<3>(zP, 0, row, a[0], a[1], a[2], a[3]);
    store_chunk3_p<3>(zP, 1, row, a[4], a[5], h[0], h[1]);
    {   // chunk 2: the parts of the constant tail (1, 0, 0, 0) are known: bf16(1) = 0x3f80, residuals 0
        uint2 pa, pb, pc;
        split3_p(h[2], pa.x, pb.x, pc.x);
        split3_p(h[3], pa.y, pb.y, pc.y);
        uint4* p = reinterpret_cast<uint4*>(zP);
        p[(0 * 3 + 2) * TM + row] = make_uint4(pa.x, pa.y, 0x00003f80u, 0u);
        p[(1 * 3 + 2) * TM + row] = make_uint4(pb.x, pb.y, 0u, 0u);
        p[(2 * 3 + 2) * TM + row] = make_uint4(pc.x, pc.y, 0u, 0u);
    }
}

// four consecutive features -> the lower (half = 0) or upper (half = 1) 8 bytes of one chunk in each of the three part planes
template <int CPP>
__device__ __forceinline__ void store_half_chunk3(unsigned char* planes, int chunk, int row, int half, float f0, float f1, float f2, float f3) {
    uint2 a, b, c;
    split3(f0, f1, a.x, b.x, c.x);
    split3(f2, f3, a.y, b.y, c.y);
    uint2* p = reinterpret_cast<uint2*>(planes);
    p[((0 * CPP + chunk) * TM + row) * 2 + half] = a;
    p[((1 * CPP + chunk) * TM + row) * 2 + half] = b;
    p[((2 * CPP + chunk) * TM + row) * 2 + half] = c;
}

// Experiment switches of the forward kernel (side builds: VB_OUT=../lib_exp VB_EXTRA_FLAGS="-DVB_FWD_PIPELINED=1" build.sh; measured on
// config 3, same box, profiles/r02/README.md):
//   VB_FWD_PIPELINED = 1: the inter layer of step t + 1 is computed in the shadow of step t's MMAs.  One CTA per SM: 2700 -> 2330 cycles
//     per step (B <= 19k: 5-8 % faster); four CTAs per SM (the benchmarked size): 0.099 -> 0.105 ms, so it stays off.
//   VB_FWD_LD24 = 0: three separate 8-column TMEM loads per chunk, each with its own wait, instead of three in flight: no difference.
#ifndef VB_FWD_PIPELINED
#define VB_FWD_PIPELINED 0
#endif
constexpr bool kFwdPipelined = VB_FWD_PIPELINED != 0;
#ifndef VB_FWD_LD24
#define VB_FWD_LD24 1
#endif


// clock64 phase traces of both kernels (VB_TRACE_BUILD=1 build.sh; profiles/microbench/trace_fwd.py, trace_bwd.py)
#ifdef VB_ENABLE_TRACE
#define VB_TR(k) do { if (tr_on) tr[tr_n][k] = clock64(); } while (0)
#else
#define VB_TR(k) do { } while (0)
#endif

// ================================================================================================ forward
template <class C>
struct FwdSmem {
    static constexpr int Z = 0;                                  // 9 z~ planes + the shared zero plane
    static constexpr int WB = Z + 10 * PLANE;                    // stacked heads operand: 10 K-chunks x 96 rows x 16 B
    static constexpr int SW = WB + 10 * 96 * 16;                 // small-layer weights (fp32)
    static constexpr int BAR = SW + C::S_TOTAL * 4;
    static constexpr int BYTES = BAR + 64;
    static_assert(BAR % 16 == 0, "barrier alignment");
};

// Forward heads in FIVE MMAs: the three bf16 parts of z~ are stacked along K (the 9 part planes + the zero plane are
// contiguous: K = 80) and the three parts of the weights along N (96 rows), D[128 x 96] = [z1|z2|z3|0] B^T with
// B[row(pb, n')][pa * 24 + f] = W_pb[n'][f].  Column group pb then holds z W_pb (all nine part products); the epilogue adds
// the three groups, smallest first.  Row order: row = (n' / 8) * 24 + pb * 8 + n' % 8, i.e. the three groups of one 8-column
// chunk (= two neurons) are adjacent TMEM columns.
__device__ __forceinline__ void issue_heads_stacked(uint32_t d_tmem, uint32_t z_planes, uint32_t w_planes) {
    constexpr uint32_t idesc = make_idesc(128, 96, 0, 0);
#pragma unroll
    for (int ks = 0; ks < 5; ++ks)
        mma_bf16(d_tmem, make_desc(z_planes + ks * 2 * PLANE, PLANE, 128), make_desc(w_planes + ks * 2 * (96 * 16), 96 * 16, 128), idesc, ks != 0);
}

// SAVE: also write the gate factors the backward needs (see the file header) into `saved`, laid out
//   float4 saved[t][tile][q][128 rows],  q < NC: (A, Bc, Cc, mish'(c)) of command neuron q;  q >= NC: m[4 (q - NC) ..]
// y / h_traj may be NULL (the backward's internal re-run when the caller kept no saved block).
template <class C, bool VEC, int LAY, bool SAVE>
__global__ void __launch_bounds__(TM) liquid_fwd_tc_kernel(const float* __restrict__ packed, const float* __restrict__ x,
                                                           const float* __restrict__ h0, float* __restrict__ y,
                                                           float* __restrict__ h_traj, float4* __restrict__ saved, int64_t B, int T,
                                                           int vec_flags) {
    using S = FwdSmem<C>;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* zP = smem + S::Z;
    unsigned char* wB = smem + S::WB;
    float* sw = reinterpret_cast<float*>(smem + S::SW);
    uint64_t* bar1 = reinterpret_cast<uint64_t*>(smem + S::BAR);
    uint64_t* barw = bar1 + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::BAR + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr bool vx = VEC, vy = VEC, vh = VEC;
    const RowMap<LAY> rx(vec_flags & LAY_X, B, T), ry(vec_flags & LAY_Y, B, T), rh(vec_flags & LAY_H, B, T);

    // stacked heads operand (see issue_heads_stacked); head outputs in PAIR order n' = 8 (j / 2) + 2 hs + j % 2: one 8-column chunk holds
    // two complete neurons and adjacent columns hold the same head of both, i.e. the packed pairs of the epilogue
    if (warp < 3) {
        const int n = lane, c = warp;     // features 8c .. 8c+7 of output n'
        const int col = ((n & 7) >> 1) * C::NCp + 2 * (n >> 3) + (n & 1);
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
    stage_small_weights<C>(sw, packed, barw);
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
        f2 h2[C::NC / 2];          // hidden state as packed pairs (h[2k], h[2k + 1])
        {
            float h[C::NC];
#pragma unroll
            for (int j = 0; j < C::NC; ++j) h[j] = 0.f;
            if (valid && h0) ldg_row<C::NC>(h, h0 + b * C::NC, vh);
#pragma unroll
            for (int k = 0; k < C::NC / 2; ++k) h2[k] = pk(h[2 * k], h[2 * k + 1]);
        }
        // a2 = mish(W_in x_t + b_in); xn = the next input row, prefetched (with kFwdPipelined a2 is carried from the previous step)
        f2 a2[C::NI / 2];
        float xn[C::I];
#pragma unroll
        for (int i = 0; i < C::I; ++i) xn[i] = 0.f;
        if constexpr (kFwdPipelined) {
            float x0[C::I];
#pragma unroll
            for (int i = 0; i < C::I; ++i) x0[i] = 0.f;
            if (valid) ldg_row<C::I>(x0, x + rx(b, 0) * C::I, vx);
            if (valid && T > 1) ldg_row<C::I>(xn, x + rx(b, 1) * C::I, vx);
            f2 u[C::NI / 2];
            inter_layer_p<C>(sw, x0, u);
#pragma unroll
            for (int k = 0; k < C::NI / 2; k += 2) mish4_p(u[k], u[k + 1], a2[k], a2[k + 1]);
        } else {
            if (valid) ldg_row<C::I>(xn, x + rx(b, 0) * C::I, vx);
        }
#ifdef VB_ENABLE_TRACE
        long long tr[3][8];
        int tr_n = 0;
        const bool tr_lane = lane == 0 && (blockIdx.x == 0 || blockIdx.x == 300) && (warp == 0 || warp == 2);
#endif
        // The element-wise work runs on packed fp32 pairs (FFMA2 / FMUL2 / FADD2: two neurons per issue slot): 963 -> 698 instructions per
        // thread-step, 0.104 -> 0.099 ms on config 3 (profiles/r02/README.md).
        for (int t = 0; t < T; ++t) {
#ifdef VB_ENABLE_TRACE
            const bool tr_on = tr_lane && t >= 16 && t < 19;
            if (tr_on) tr_n = t - 16;
#endif
            VB_TR(0);
            if constexpr (!kFwdPipelined) {
                {
                    float xr[C::I];
#pragma unroll
                    for (int i = 0; i < C::I; ++i) xr[i] = xn[i];
                    if (valid && t + 1 < T) ldg_row<C::I>(xn, x + rx(b, t + 1) * C::I, vx);
                    f2 u[C::NI / 2];
                    inter_layer_p<C>(sw, xr, u);
#pragma unroll
                    for (int k = 0; k < C::NI / 2; k += 2) mish4_p(u[k], u[k + 1], a2[k], a2[k + 1]);
                }
            }
            store_z_p<C>(zP, tid, a2, h2);
            fence_async_smem();
            tc_fence_before();
            VB_TR(1);
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                VB_TR(3);
                issue_heads_stacked(tmem, z_addr, w_addr);
                mma_commit(bar1);
            }
            VB_TR(2);
            // in the shadow of the MMAs: the inter layer of the NEXT step (it does not depend on this step's heads)
            if (kFwdPipelined && t + 1 < T) {
                float xr[C::I];
#pragma unroll
                for (int i = 0; i < C::I; ++i) xr[i] = xn[i];
                if (valid && t + 2 < T) ldg_row<C::I>(xn, x + rx(b, t + 2) * C::I, vx);   // prefetch one more step ahead
                f2 u[C::NI / 2];
                inter_layer_p<C>(sw, xr, u);
#pragma unroll
                for (int k = 0; k < C::NI / 2; k += 2) mish4_p(u[k], u[k + 1], a2[k], a2[k + 1]);
            }
            VB_TR(4);
            mbar_wait(bar1, ph1);
            ph1 ^= 1;
            VB_TR(5);
            tc_fence_after();
            f2 m2[C::NC / 2];
            ulonglong2* sv = SAVE ? reinterpret_cast<ulonglong2*>(saved + ((int64_t)t * n_tiles + tile) * (C::NQ * TM) + tid) : nullptr;
            const f2 one = bc2(1.0f);
#pragma unroll
            for (int c = 0; c < C::NC / 2; ++c) {
                // 24 columns = one 8-column chunk x (W_hi, W_mid, W_lo): column 24 c + 8 g + 2 hs + jj is z W_g of head hs, neuron 2c + jj
                f2 s4[4];
                {
                    f2 v[12];
#if VB_FWD_LD24
                    tmem_ld24_p(taddr + 24 * c, v);
#else
                    {
                        f2 q[4];
                        tmem_ld8_p(taddr + 24 * c, q);
                        v[0] = q[0]; v[1] = q[1]; v[2] = q[2]; v[3] = q[3];
                        tmem_ld8_p(taddr + 24 * c + 8, q);
                        v[4] = q[0]; v[5] = q[1]; v[6] = q[2]; v[7] = q[3];
                        tmem_ld8_p(taddr + 24 * c + 16, q);
                        v[8] = q[0]; v[9] = q[1]; v[10] = q[2]; v[11] = q[3];
                    }
#endif
#pragma unroll
                    for (int e = 0; e < 4; ++e) s4[e] = add2(add2(v[8 + e], v[4 + e]), v[e]);
                }
                f2 tg, th, gate;
                tanh2_sigmoid_p(s4[0], s4[1], s4[2], tg, th, gate);
                const f2 d = sub2(th, tg);
                const f2 hn = fma2(gate, d, tg);
                h2[c] = hn;
                const f2 cc2 = add2(s4[3], hn);
                f2 g2;
                mish_grad_p(cc2, m2[c], g2);
                if (SAVE) {
                    const f2 omg = sub2(one, gate);
                    const f2 fa = mul2(omg, sub2(one, mul2(tg, tg)));
                    const f2 fb = mul2(gate, sub2(one, mul2(th, th)));
                    const f2 fc = mul2(mul2(d, gate), omg);
                    sv[(2 * c) * TM] = make_ulonglong2(fa, fb);
                    sv[(2 * c + 1) * TM] = make_ulonglong2(fc, g2);
                }
            }
            if (SAVE) {
#pragma unroll
                for (int q = 0; q < C::NC / 4; ++q) sv[(C::NC + q) * TM] = make_ulonglong2(m2[2 * q], m2[2 * q + 1]);
            }
            VB_TR(6);
            float yo[C::O];
            {
                const float* wo = sw + C::S_OUT;
#pragma unroll
                for (int o = 0; o < C::O; ++o) {
                    f2 acc = pk(wo[C::O * C::NCp + o], 0.f);
#pragma unroll
                    for (int q = 0; q < C::NC / 4; ++q) {
                        const ulonglong2 w = *reinterpret_cast<const ulonglong2*>(wo + o * C::NCp + 4 * q);
                        acc = fma2(m2[2 * q], w.x, acc);
                        acc = fma2(m2[2 * q + 1], w.y, acc);
                    }
                    yo[o] = lo(acc) + hi(acc);
                }
            }
            if (valid) {
                if (y) stg_row<C::O>(y + ry(b, t) * C::O, yo, vy);
                if (h_traj) {
                    float h[C::NC];
#pragma unroll
                    for (int k = 0; k < C::NC / 2; ++k) { h[2 * k] = lo(h2[k]); h[2 * k + 1] = hi(h2[k]); }
                    stg_row<C::NC>(h_traj + rh(b, t) * C::NC, h, vh);
                }
            }
            VB_TR(7);
        }
#ifdef VB_ENABLE_TRACE
        if (tr_lane && T >= 19)
            for (int n = 0; n < 3; ++n)
                printf("FWDTRACE cta %d warp %d step %d : top %lld | storez %lld | sync+issue %lld (issue %lld) | next-inter %lld | mmawait %lld | heads-epi %lld | motor+stores %lld | next-top %lld\n",
                       (int)blockIdx.x, warp, 16 + n, tr[n][0] - tr[0][0], tr[n][1] - tr[n][0], tr[n][2] - tr[n][1],
                       warp == 0 ? tr[n][2] - tr[n][3] : 0LL, tr[n][4] - tr[n][2], tr[n][5] - tr[n][4], tr[n][6] - tr[n][5], tr[n][7] - tr[n][6],
                       n < 2 ? tr[n + 1][0] - tr[n][7] : 0LL);
#endif
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

// ================================================================================================ backward
constexpr int BT = 2 * TM;                  // backward CTA: two threads per sample
constexpr int D3_COLS = 80;                 // dW accumulator columns: three z~ parts x 24 features (+ 8 unused)
constexpr int D3_FLOATS = TM * D3_COLS;     // per-CTA fp32 flush buffer

template <class C>
struct BwdSmem {
    static constexpr int Z = 0;                                  // 9 z~ part planes; the 10th chunk of the N = 80 operand reads into DS (unused columns)
    static constexpr int DS = Z + 9 * PLANE;                     // 12 ds part planes; the MN-major M = 128 stack reads 4 planes further (unused rows)
    static constexpr int SBUF = DS + 12 * PLANE;                 // fp32 rows du (NI) | x (I), 128 samples each: the dW_in transposition buffer
    static constexpr int SROWS = C::NI + C::I;
    static constexpr int SBUF_BYTES = SROWS * TM * 4 > 4 * PLANE ? SROWS * TM * 4 : 4 * PLANE;   // also the over-read tail of the ds stack
    static constexpr int WB = SBUF + SBUF_BYTES;                 // heads weights, three parts x three 8-feature chunks x 32 rows (dz = ds W)
    static constexpr int WB_BYTES = 3 * 3 * 32 * 16 + 32 * 16;   // + the 4th chunk of the last part (unused columns, must be readable)
    static constexpr int SW = WB + WB_BYTES;
    static constexpr int BAR = SW + C::S_TOTAL * 4;
    static constexpr int BYTES = BAR + 64;
    static_assert(BAR % 16 == 0, "barrier alignment");
    // the two deliberate operand over-reads stay inside this CTA's window, in memory the kernel initialises:
    static_assert(DS + 16 * PLANE <= WB, "the MN-major M = 128 stack (16 planes from DS) must end before the weight planes");
    static_assert(Z + 10 * PLANE <= SBUF, "the N = 80 operand (10 planes from Z) must end inside the ds planes");
};

template <class C, bool VEC, int LAY>
__global__ void __launch_bounds__(BT, (BwdSmem<C>::BYTES + 1024 > 58368 ? 3 : 4)) liquid_bwd_tc_kernel(
    const float* __restrict__ packed, const float* __restrict__ x, const float* __restrict__ h0, const float* __restrict__ h_traj,
    const float* __restrict__ dy, const float* __restrict__ dh_traj, const float* __restrict__ dh_last, const float4* __restrict__ saved,
    float* __restrict__ dx, float* __restrict__ dh0, float* __restrict__ partials, float* __restrict__ d3buf,
    int64_t B, int T, int vec_flags) {
    using S = BwdSmem<C>;
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = tid & (TM - 1);          // sample of the tile = TMEM lane
    const bool roleB = tid >= TM;            // warps 4-7: gate factors -> ds, h-part of z~, dh carry; warps 0-3: inter layer, du, dx, dW_in
    const int wq = warp & 3;
    unsigned char* zP = smem + S::Z;
    unsigned char* dsP = smem + S::DS;
    float* sbuf = reinterpret_cast<float*>(smem + S::SBUF);
    unsigned char* wB = smem + S::WB;
    float* sw = reinterpret_cast<float*>(smem + S::SW);
    uint64_t* bar_dz = reinterpret_cast<uint64_t*>(smem + S::BAR);
    uint64_t* bar_dw = bar_dz + 1;
    uint64_t* barw = bar_dz + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::BAR + 32);
    constexpr bool vx = VEC, vy = VEC, vh = VEC;
    const RowMap<LAY> rx(vec_flags & LAY_X, B, T), ry(vec_flags & LAY_Y, B, T), rh(vec_flags & LAY_H, B, T), rdx(vec_flags & LAY_DX, B, T),
        rdh(vec_flags & LAY_DH, B, T);

    // dz operand B[n = ds feature (neuron-major: 4 j + hs)][k = z~ feature], three parts x three chunks; chunk stride 512 B
    if (tid < 96) {
        const int n = tid & 31, kc = tid >> 5;
        const int col = (n & 3) * C::NCp + (n >> 2);
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = 8 * kc + e;
            v[e] = k < C::K ? __ldg(packed + C::WH_T + k * C::H4 + col) : 0.f;
        }
        uint4 a, b, c;
        split3(v[0], v[1], a.x, b.x, c.x);
        split3(v[2], v[3], a.y, b.y, c.y);
        split3(v[4], v[5], a.z, b.z, c.z);
        split3(v[6], v[7], a.w, b.w, c.w);
        uint4* p = reinterpret_cast<uint4*>(wB);
        p[(0 * 3 + kc) * 32 + n] = a;
        p[(1 * 3 + kc) * 32 + n] = b;
        p[(2 * 3 + kc) * 32 + n] = c;
    } else if (tid < 128) {
        reinterpret_cast<uint4*>(wB)[9 * 32 + (tid - 96)] = make_uint4(0, 0, 0, 0);
    }
    {   // every operand byte an MMA can touch is initialised (unused rows / columns only ever meet finite garbage)
        const uint4 z4 = make_uint4(0, 0, 0, 0);
        uint4* q = reinterpret_cast<uint4*>(smem);
        for (int i = tid; i < S::WB / 16; i += BT) q[i] = z4;
    }
    // fp32 flush buffer of the dW accumulator: element (stack row, column c) at ((c / 4) * TM + row) * 4 + c % 4, so that the 32 lanes
    // of one vector access touch 512 contiguous bytes.  Role A owns columns [0, 40), role B [40, 80).
    float* d3row = d3buf + (int64_t)blockIdx.x * D3_FLOATS + row * 4 + (roleB ? (D3_COLS / 8) * (TM * 4) : 0);
#pragma unroll
    for (int q = 0; q < D3_COLS / 8; ++q) *reinterpret_cast<float4*>(d3row + q * (TM * 4)) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        mbar_init(bar_dz, 1);
        mbar_init(bar_dw, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    stage_small_weights<C>(sw, packed, barw);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16);
    const uint32_t z_addr = smem_u32(zP), ds_addr = smem_u32(dsP), wb_addr = smem_u32(wB);
    // TMEM columns: [0,32) dz;  [32,112) dW = stack(ds)^T x [z~_b1 | z~_b2 | z~_b3 | -];  [112,128) role A's parking (mish'(u), x)
    uint32_t ph_dz = 0, ph_dw = 0;
    bool dw_in_flight = false;    // a dW batch has been committed and not yet waited for
    int since_flush = 0;          // dW batches accumulated in TMEM since the last flush (0 => next batch overwrites)

    // role A: dW_in / db_in accumulators of warp wq's three inter neurons (j = 3 wq + r) over this lane's four samples
    // role B: dW_out / db_out accumulators of this thread's own sample (reduced over the CTA at the end)
    constexpr int JA = C::NI / 4;
    static_assert(C::NI % 4 == 0, "dW_in blocking: NI / 4 inter neurons per warp");
    constexpr int NACC_A = C::I * JA + JA, NACC_B = C::O * C::NC + C::O;
    constexpr int NACC = NACC_A > NACC_B ? NACC_A : NACC_B;
    float acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.f;

    auto flush_dw = [&]() {   // all threads; requires the dW batch to be complete.  40 columns per thread.
        tc_fence_after();
        const uint32_t c0 = 32 + (roleB ? 40 : 0);
#pragma unroll
        for (int g = 0; g < 5; ++g) {
            float v[8];
            tmem_ld8(taddr + c0 + 8 * g, v);
            // fire-and-forget vector reductions into this thread's private columns: no load, nothing to wait for
#pragma unroll
            for (int q = 0; q < 2; ++q)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d3row + (2 * g + q) * (TM * 4)), "f"(v[4 * q]), "f"(v[4 * q + 1]),
                             "f"(v[4 * q + 2]), "f"(v[4 * q + 3]) : "memory");
        }
        tc_fence_before();
        since_flush = 0;
    };

    const int64_t n_tiles = (B + TM - 1) / TM;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t b = tile * TM + row;
        const bool valid = b < B;
        // the step-carried state of the two roles shares ONE register array (the allocator sees both roles' live ranges in the one loop):
        // role B: dhc = dL/dh_t carried backwards;  role A: xn = x_t prefetched one step ahead
        constexpr int NCARRY = C::NC > C::I ? C::NC : C::I;
        float carry[NCARRY];
#pragma unroll
        for (int j = 0; j < NCARRY; ++j) carry[j] = 0.f;
        float (&dhc)[C::NC] = reinterpret_cast<float (&)[C::NC]>(carry);
        float (&xn)[C::I] = reinterpret_cast<float (&)[C::I]>(carry);
        if (roleB && valid && dh_last) ldg_row<C::NC>(dhc, dh_last + b * C::NC, vh);
        if (!roleB && valid) ldg_row<C::I>(xn, x + rx(b, T - 1) * C::I, vx);

#ifdef VB_ENABLE_TRACE
        long long tr[3][8];
        int tr_n = 0;
        const bool tr_lane = lane == 0 && (blockIdx.x == 0 || blockIdx.x == 300) && (warp == 0 || warp == 2 || warp == 4 || warp == 6);
#endif
        for (int t = T - 1; t >= 0; --t) {
#ifdef VB_ENABLE_TRACE
            const bool tr_on = tr_lane && t <= 16 && t > 13;
            if (tr_on) tr_n = 16 - t;
#endif
            VB_TR(0);
            // ------------------------------------------------------------------ phase 1: operands of this step
            // the previous dW batch has finished reading the z~ / ds planes (it was queued right behind the dz this CTA already
            // waited for and has had all of phase 2 to complete: this wait is normally free)
            if (dw_in_flight) {
                mbar_wait(bar_dw, ph_dw);
                ph_dw ^= 1;
                dw_in_flight = false;
            }
            if (since_flush >= kFlush) flush_dw();
            VB_TR(1);
            if (!roleB) {
                float xr[C::I];
#pragma unroll
                for (int i = 0; i < C::I; ++i) xr[i] = xn[i];
                if (valid && t > 0) ldg_row<C::I>(xn, x + rx(b, t - 1) * C::I, vx);
                // role A has slack before the barrier: it pulls the NEXT step's saved factors of its warp's 32 samples (NQ x 512 B, read by
                // role B's warp wq + 4) from DRAM into L2, one 128-byte line per lane.  Role B's loads are issued chunk by chunk (its register
                // budget keeps only one neuron pair's factors live), so each of them used to wait a full DRAM latency: 0.200 -> 0.192 ms.
                if (t > 0) {
                    const char* nx = reinterpret_cast<const char*>(saved + ((int64_t)(t - 1) * n_tiles + tile) * (C::NQ * TM) + wq * 32);
#pragma unroll
                    for (int l = lane; l < C::NQ * 4; l += 32)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + (l >> 2) * (TM * 16) + (l & 3) * 128));
                }
                float u[C::NI];
                inter_layer<C>(sw, xr, u);
                // Mish'(u) and x_t wait in this thread's TMEM lane (16 spare columns) for the dz MMA; a goes to the z~ planes
                // (x_t rides along when it fits the four spare columns; wider inputs are re-read from global memory after the MMA)
                static_assert(C::NI == 12, "parking area: 8 + 8 TMEM columns = Mish'(u[0..8)) | Mish'(u[8..12)), x (I <= 4)");
                {
                    float a8[8], g8[8];
#pragma unroll
                    for (int j = 0; j < 8; j += 2) mish2_grad(u[j], u[j + 1], a8[j], a8[j + 1], g8[j], g8[j + 1]);
                    tmem_st8(taddr + 112, g8);
                    store_chunk3<3>(zP, 0, row, a8);
                }
                {
                    float a4[4], g8[8];
                    mish2_grad(u[8], u[9], a4[0], a4[1], g8[0], g8[1]);
                    mish2_grad(u[10], u[11], a4[2], a4[3], g8[2], g8[3]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) g8[4 + i] = i < C::I ? xr[i < C::I ? i : 0] : 0.f;
                    tmem_st8(taddr + 120, g8);
                    store_half_chunk3<3>(zP, 1, row, 0, a4[0], a4[1], a4[2], a4[3]);
                }
            } else {
                float hp[C::NC], dyr[C::O];
#pragma unroll
                for (int j = 0; j < C::NC; ++j) hp[j] = 0.f;
#pragma unroll
                for (int o = 0; o < C::O; ++o) dyr[o] = 0.f;
                if (valid) {
                    if (t > 0) ldg_row<C::NC>(hp, h_traj + rh(b, t - 1) * C::NC, vh);
                    else if (h0) ldg_row<C::NC>(hp, h0 + b * C::NC, vh);
                    ldg_row<C::O>(dyr, dy + ry(b, t) * C::O, vy);
                    if (dh_traj) {
                        float e[C::NC];
                        ldg_row<C::NC>(e, dh_traj + rdh(b, t) * C::NC, vh);
#pragma unroll
                        for (int j = 0; j < C::NC; ++j) dhc[j] += e[j];
                    }
                }
                const float4* sv = saved + ((int64_t)t * n_tiles + tile) * (C::NQ * TM) + row;
                store_half_chunk3<3>(zP, 1, row, 1, hp[0], hp[1], hp[2], hp[3]);
                {
                    float f[8] = {hp[4], hp[5], hp[6], hp[7], 1.0f, 0.f, 0.f, 0.f};
                    store_chunk3<3>(zP, 2, row, f);
                }
                // gating derivatives from the saved values, two neurons (= one 8-feature chunk of ds) at a time; dW_out / db_out of this
                // thread's own sample along the way
                const float* wo = sw + C::S_OUT;
#pragma unroll
                for (int o = 0; o < C::O; ++o) acc[C::O * C::NC + o] += dyr[o];
#pragma unroll
                for (int c = 0; c < C::NC / 2; ++c) {
                    float s8[8];
                    float4 g[2];
                    float mm[2];
                    {   // saved per neuron pair: (A0, A1, Bc0, Bc1), (Cc0, Cc1, G0, G1)
                        const float4 qa = ldg_stream(sv + (2 * c) * TM), qb = ldg_stream(sv + (2 * c + 1) * TM);
                        g[0] = make_float4(qa.x, qa.z, qb.x, qb.z);
                        g[1] = make_float4(qa.y, qa.w, qb.y, qb.w);
                    }
                    {
                        const float4 mq = ldg_stream(sv + (C::NC + c / 2) * TM);
                        mm[0] = (c & 1) ? mq.z : mq.x;
                        mm[1] = (c & 1) ? mq.w : mq.y;
                    }
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) {
                        const int j = 2 * c + jj;
                        float dm = 0.f;
#pragma unroll
                        for (int o = 0; o < C::O; ++o) {
                            dm = fmaf(dyr[o], wo[o * C::NCp + j], dm);
                            acc[o * C::NC + j] = fmaf(dyr[o], mm[jj], acc[o * C::NC + j]);
                        }
                        const float dc = dm * g[jj].w;
                        const float dhn = dc + dhc[j];
                        s8[4 * jj + 0] = dhn * g[jj].x;        // ds_g
                        s8[4 * jj + 1] = dhn * g[jj].y;        // ds_h
                        s8[4 * jj + 2] = dhn * g[jj].z;        // ds_f (both f heads)
                        s8[4 * jj + 3] = dc;                   // ds_proj
                    }
                    store_chunk3<4>(dsP, c, row, s8);
                }
            }
            fence_async_smem();
            tc_fence_before();
            VB_TR(2);
            __syncthreads();
            if (tid == TM) {                    // one lane of role B issues: B is idle until dz arrives
                tc_fence_after();
                // dz = ds W: six part products (small first), two K-steps each
                constexpr uint32_t idesc2 = make_idesc(128, 32, 0, 1);
#pragma unroll
                for (int p = 0; p < 6; ++p) {
                    const int pa = (p == 0) ? 1 : (p == 1) ? 0 : (p == 2) ? 2 : (p == 3) ? 0 : (p == 4) ? 1 : 0;
                    const int pb = (p == 0) ? 1 : (p == 1) ? 2 : (p == 2) ? 0 : (p == 3) ? 1 : (p == 4) ? 0 : 0;
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
                        mma_bf16(tmem + 0, make_desc(ds_addr + (pa * 4 + 2 * ks) * PLANE, PLANE, 128),
                                 make_desc(wb_addr + pb * (3 * 32 * 16) + ks * 256, 128, 32 * 16), idesc2, (p | ks) != 0);
                }
                mma_commit(bar_dz);
                // dW (+ bias column) += stack(ds)^T x [z~_b1 | z~_b2 | z~_b3 | -] : all nine part products in one accumulator
                constexpr uint32_t idesc3 = make_idesc(128, 80, 1, 1);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    mma_bf16(tmem + 32, make_desc(ds_addr + ks * 256, 128, PLANE), make_desc(z_addr + ks * 256, 128, PLANE), idesc3,
                             !(since_flush == 0 && ks == 0));
                mma_commit(bar_dw);
            }
            since_flush += 1;
            dw_in_flight = true;
            VB_TR(3);
            // ------------------------------------------------------------------ phase 2: consume dz
            mbar_wait(bar_dz, ph_dz);
            ph_dz ^= 1;
            VB_TR(4);
            tc_fence_after();
            if (!roleB) {
                float du[C::NI], xr2[C::I];
                {
                    float dmu[16], dz8[8], dz4[4];
                    tmem_ld16(taddr + 112, dmu);
                    tmem_ld8(taddr + 0, dz8);
                    tmem_ld4(taddr + 8, dz4);
#pragma unroll
                    for (int j = 0; j < 8; ++j) du[j] = dz8[j] * dmu[j];
#pragma unroll
                    for (int j = 8; j < C::NI; ++j) du[j] = dz4[j - 8] * dmu[j];
                    if constexpr (C::I <= 4) {
#pragma unroll
                        for (int i = 0; i < C::I; ++i) xr2[i] = dmu[C::NI + i];
                    } else {
#pragma unroll
                        for (int i = 0; i < C::I; ++i) xr2[i] = 0.f;
                        if (valid) ldg_row<C::I>(xr2, x + rx(b, t) * C::I, vx);
                    }
                }
                tc_fence_before();
                if (dx) {
                    float dxr[C::I];
#pragma unroll
                    for (int i = 0; i < C::I; ++i) dxr[i] = 0.f;
                    const float* wi = sw + C::S_WIN;
#pragma unroll
                    for (int j = 0; j < C::NI; ++j) {
                        if constexpr (C::Ip == 4) {
                            const float4 w = *reinterpret_cast<const float4*>(wi + j * C::Ip);
                            const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                            for (int i = 0; i < C::I; ++i) dxr[i] = fmaf(du[j], wv[i], dxr[i]);
                        } else {
#pragma unroll
                            for (int i = 0; i < C::I; ++i) dxr[i] = fmaf(du[j], wi[j * C::Ip + i], dxr[i]);
                        }
                    }
                    if (valid) stg_row<C::I>(dx + rdx(b, t) * C::I, dxr, vx);
                }
                // dW_in / db_in: transpose du and x through shared memory, then each warp reduces its three inter neurons over the
                // tile's 128 samples (lane = four samples).  The two named barriers involve role A's four warps only.
                asm volatile("bar.sync 1, 128;" ::: "memory");      // the previous step's reduction has finished reading sbuf
#pragma unroll
                for (int j = 0; j < C::NI; ++j) sbuf[j * TM + row] = du[j];
#pragma unroll
                for (int i = 0; i < C::I; ++i) sbuf[(C::NI + i) * TM + row] = xr2[i];
                asm volatile("bar.sync 1, 128;" ::: "memory");
                {
                    float4 xv[C::I], dv[JA];
#pragma unroll
                    for (int i = 0; i < C::I; ++i) xv[i] = *reinterpret_cast<const float4*>(sbuf + (C::NI + i) * TM + 4 * lane);
#pragma unroll
                    for (int r = 0; r < JA; ++r) dv[r] = *reinterpret_cast<const float4*>(sbuf + (JA * wq + r) * TM + 4 * lane);
#pragma unroll
                    for (int r = 0; r < JA; ++r) {
#pragma unroll
                        for (int i = 0; i < C::I; ++i)
                            acc[i * JA + r] = fmaf(xv[i].x, dv[r].x, fmaf(xv[i].y, dv[r].y, fmaf(xv[i].z, dv[r].z, fmaf(xv[i].w, dv[r].w, acc[i * JA + r]))));
                        acc[C::I * JA + r] += (dv[r].x + dv[r].y) + (dv[r].z + dv[r].w);
                    }
                }
            } else {
                float dz8[8];
                tmem_ld8(taddr + C::NI, dz8);      // dz[:, NI .. NI + NC) = dL/dh_{t-1}
#pragma unroll
                for (int j = 0; j < C::NC; ++j) dhc[j] = dz8[j];
                tc_fence_before();
            }
            VB_TR(5);
        }
#ifdef VB_ENABLE_TRACE
        if (tr_lane && T >= 19)
            for (int n = 0; n < 3; ++n)
                printf("BWDTRACE cta %d warp %d step %d : top %lld | dw-wait/flush %lld | phase1 %lld | sync(+issue) %lld | dz-wait %lld | phase2 %lld | next-top %lld\n",
                       (int)blockIdx.x, warp, 16 - n, tr[n][0] - tr[0][0], tr[n][1] - tr[n][0], tr[n][2] - tr[n][1], tr[n][3] - tr[n][2],
                       tr[n][4] - tr[n][3], tr[n][5] - tr[n][4], n < 2 ? tr[n + 1][0] - tr[n][5] : 0LL);
#endif
        if (roleB && valid && dh0) stg_row<C::NC>(dh0 + b * C::NC, dhc, vh);
    }
    if (dw_in_flight) mbar_wait(bar_dw, ph_dw);
    if (since_flush > 0) flush_dw();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");

    // ---------------- per-CTA partial gradient row (F layout)
    float* prow = partials + (int64_t)blockIdx.x * C::F_TOTAL;
    for (int idx = tid; idx < C::F_TOTAL; idx += BT) prow[idx] = 0.f;
    // role A: red[entry][lane] (32 partial sums per dW_in / db_in entry); role B: redB[entry][128 samples]
    constexpr int RS = 33;
    constexpr int NA = C::I * C::NI + C::NI, NB = C::O * C::NC + C::O;
    float* red = reinterpret_cast<float*>(smem);          // the operand planes are dead
    float* redB = red + NA * RS;
    static_assert((NA * RS + NB * (TM + 1)) * 4 <= S::WB, "end-of-kernel reduction buffers must fit the operand planes");
    if (!roleB) {
#pragma unroll
        for (int r = 0; r < JA; ++r) {
#pragma unroll
            for (int i = 0; i < C::I; ++i) red[(i * C::NI + JA * wq + r) * RS + lane] = acc[i * JA + r];
            red[(C::I * C::NI + JA * wq + r) * RS + lane] = acc[C::I * JA + r];
        }
    } else {
#pragma unroll
        for (int k = 0; k < NB; ++k) redB[k * (TM + 1) + row] = acc[k];
    }
    __syncthreads();
    {
        const float* d3 = d3buf + (int64_t)blockIdx.x * D3_FLOATS;
        // stack row (pa, r) x column (pb, i): dW_heads[i][r] (i < K), db_heads[r] (i == ONE)
        for (int e = tid; e < 32 * 24; e += BT) {
            const int r = e / 24, i = e % 24;           // r = 4 j + hs (neuron-major)
            const int col = (r & 3) * C::NCp + (r >> 2);
            if (i < C::K || i == C::ONE) {
                float s = 0.f;
#pragma unroll
                for (int pa = 0; pa < 3; ++pa)
#pragma unroll
                    for (int pb = 2; pb >= 0; --pb) s += d3[(((pb * 24 + i) >> 2) * TM + pa * 32 + r) * 4 + ((pb * 24 + i) & 3)];
                if (i < C::K) prow[C::WH_T + i * C::H4 + col] = s;
                else prow[C::B_H + col] = s;
            }
        }
        for (int e = tid; e < NA; e += BT) {
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int g = 0; g < 32; g += 2) { a0 += red[e * RS + g]; a1 += red[e * RS + g + 1]; }
            const int idx = e < C::I * C::NI ? C::WIN_T + (e / C::NI) * C::NIp + (e % C::NI) : C::B_IN + (e - C::I * C::NI);
            prow[idx] = a0 + a1;
        }
        for (int e = tid; e < NB; e += BT) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
            for (int g = 0; g < TM; g += 4) {
                a0 += redB[e * (TM + 1) + g]; a1 += redB[e * (TM + 1) + g + 1]; a2 += redB[e * (TM + 1) + g + 2]; a3 += redB[e * (TM + 1) + g + 3];
            }
            const int idx = e < C::O * C::NC ? C::WOUT + (e / C::NC) * C::NCp + (e % C::NC) : C::B_OUT + (e - C::O * C::NC);
            prow[idx] = (a0 + a1) + (a2 + a3);
        }
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
                         const void* x2, const void* h_e, int I, int O, int Nc) {
    auto al = [](const void* p, int bytes) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & (bytes - 1)) == 0; };
    auto width = [](int n) { return (n & 3) == 0 ? 16 : ((n & 1) == 0 ? 8 : 4); };
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
static int grid_cap(K kernel, int threads, int smem, int tmem_cols, int* per_sm) {
    LaunchCfg cfg;
    int rc = cached_launch_cfg(kernel, threads, smem, &cfg);
    if (rc) return rc;
    static int smem_sm[64], regs_sm[64];
    int dev = 0;
    VB_CUDA(cudaGetDevice(&dev));
    if (!regs_sm[dev & 63]) {
        VB_CUDA(cudaDeviceGetAttribute(&smem_sm[dev & 63], cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
        VB_CUDA(cudaDeviceGetAttribute(&regs_sm[dev & 63], cudaDevAttrMaxRegistersPerMultiprocessor, dev));
    }
    const int regs_cta = ((cfg.regs + 7) & ~7) * threads;
    int n = smem_sm[dev & 63] / (smem + cfg.static_smem + 1024);
    if (regs_cta > 0 && regs_sm[dev & 63] / regs_cta < n) n = regs_sm[dev & 63] / regs_cta;
    if (512 / tmem_cols < n) n = 512 / tmem_cols;
    if (n > 2048 / threads) n = 2048 / threads;
    if (n > 32) n = 32;
    *per_sm = n < 1 ? 1 : n;
    return 0;
}

template <class C>
static int64_t saved_bytes_for(int64_t B, int T) {
    const int64_t n_tiles = (B + TM - 1) / TM;
    return (int64_t)T * n_tiles * C::NQ * TM * (int64_t)sizeof(float4);
}

template <class C>
static int launch_fwd(const float* packed, const float* x, const float* h0, float* y, float* h_traj, void* saved, int64_t saved_bytes,
                      int64_t B, int T, int layout, cudaStream_t stream) {
    if (saved) {
        VB_CHECK_ARG((reinterpret_cast<uintptr_t>(saved) & 15) == 0, VB_ERR_ALIGNMENT, "vb_liquid_fwd: the saved block must be 16-byte aligned");
        VB_CHECK_ARG(saved_bytes >= saved_bytes_for<C>(B, T), VB_ERR_WORKSPACE, "vb_liquid_fwd: saved block too small (%lld bytes, need %lld)",
                     (long long)saved_bytes, (long long)saved_bytes_for<C>(B, T));
    }
    int vf = vec_flags_for(x, y, h0, h_traj, nullptr, nullptr, nullptr, nullptr, C::I, C::O, C::NC) | layout;
    // fast variants: aligned rows and one storage order for every sequence tensor; anything else takes the generic variant
    const int all = LAY_X | LAY_Y | LAY_H;
    const bool vec = (vf & 7) == 7;
    const int variant = (vec && (layout & all) == all) ? 1 : (vec && (layout & all) == 0) ? 0 : 2;
    auto kern = saved ? (variant == 1 ? liquid_fwd_tc_kernel<C, true, 1, true> : variant == 0 ? liquid_fwd_tc_kernel<C, true, 0, true>
                                                                                              : liquid_fwd_tc_kernel<C, false, 2, true>)
                      : (variant == 1 ? liquid_fwd_tc_kernel<C, true, 1, false> : variant == 0 ? liquid_fwd_tc_kernel<C, true, 0, false>
                                                                                               : liquid_fwd_tc_kernel<C, false, 2, false>);
    int per_sm = 1;
    int rc = grid_cap(kern, TM, FwdSmem<C>::BYTES, 128, &per_sm);
    if (rc) return rc;
    int64_t n_tiles = (B + TM - 1) / TM;
    int64_t cap = (int64_t)device_sms() * per_sm;
    int grid = (int)(n_tiles < cap ? n_tiles : cap);
    if (grid < 1) grid = 1;
    kern<<<grid, TM, FwdSmem<C>::BYTES, stream>>>(packed, x, h0, y, h_traj, reinterpret_cast<float4*>(saved), B, T, vf);
    VB_LAUNCH_CHECK();
    return 0;
}

template <class C>
static int launch_bwd(const float* packed, const float* x, const float* h0, const float* h_traj, const float* dy,
                      const float* dh_traj, const float* dh_last, const void* saved, int64_t saved_bytes, float* dx, float* dh0, float* dpacked,
                      int64_t B, int T, int layout, void* ws, int64_t ws_bytes, cudaStream_t stream) {
    int vf = vec_flags_for(x, dy, h0, h_traj, dh_traj, dh_last, dx, dh0, C::I, C::O, C::NC) | layout;
    const int all = LAY_X | LAY_Y | LAY_H | (dx ? LAY_DX : 0) | (dh_traj ? LAY_DH : 0);
    const bool vec = (vf & 7) == 7;
    auto kern = (vec && (layout & all) == all) ? liquid_bwd_tc_kernel<C, true, 1>
              : (vec && (layout & all) == 0)   ? liquid_bwd_tc_kernel<C, true, 0> : liquid_bwd_tc_kernel<C, false, 2>;
    int per_sm = 1;
    int rc = grid_cap(kern, BT, BwdSmem<C>::BYTES, 128, &per_sm);
    if (rc) return rc;
    const int64_t n_tiles = (B + TM - 1) / TM;
    const int64_t cap = (int64_t)device_sms() * per_sm;
    int grid = (int)(n_tiles < cap ? n_tiles : cap);
    if (grid < 1) grid = 1;
    int64_t need = (int64_t)grid * (C::F_TOTAL + D3_FLOATS) * (int64_t)sizeof(float);
    need = (need + 255) & ~(int64_t)255;
    const int64_t sbytes = saved_bytes_for<C>(B, T);
    const bool own_saved = saved == nullptr;       // the caller kept no saved block: re-run the forward into the workspace first
    VB_CHECK_ARG(ws && ws_bytes >= need + (own_saved ? sbytes : 0), VB_ERR_WORKSPACE, "vb_liquid_bwd: workspace too small (%lld bytes, need %lld)",
                 (long long)ws_bytes, (long long)(need + (own_saved ? sbytes : 0)));
    VB_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 15) == 0, VB_ERR_ALIGNMENT, "vb_liquid_bwd: the workspace must be 16-byte aligned");
    if (own_saved) {
        void* sv = reinterpret_cast<char*>(ws) + need;
        rc = launch_fwd<C>(packed, x, h0, nullptr, nullptr, sv, sbytes, B, T, (layout & LAY_X) ? (LAY_X | LAY_Y | LAY_H) : 0, stream);
        if (rc) return rc;
        saved = sv;
    } else {
        VB_CHECK_ARG((reinterpret_cast<uintptr_t>(saved) & 15) == 0, VB_ERR_ALIGNMENT, "vb_liquid_bwd: the saved block must be 16-byte aligned");
        VB_CHECK_ARG(saved_bytes >= sbytes, VB_ERR_WORKSPACE, "vb_liquid_bwd: saved block too small (%lld bytes, need %lld)",
                     (long long)saved_bytes, (long long)sbytes);
    }
    float* partials = reinterpret_cast<float*>(ws);
    float* d3buf = partials + (int64_t)grid * C::F_TOTAL;
    kern<<<grid, BT, BwdSmem<C>::BYTES, stream>>>(packed, x, h0, h_traj, dy, dh_traj, dh_last, reinterpret_cast<const float4*>(saved), dx, dh0,
                                                  partials, d3buf, B, T, vf);
    VB_LAUNCH_CHECK();
    reduce_partials_tc_kernel<<<(C::F_TOTAL + 31) / 32, 1024, 0, stream>>>(partials, grid, C::F_TOTAL, dpacked);
    VB_LAUNCH_CHECK();
    return 0;
}

// (4,20,2) / (3,20,2): the SAC actors of BASELINE configs 1-3; (10,20,5): the reference's own fixture (tests/models/test_lnn.py:280-309);
// (8,20,2): an 8-observation actor.  The wide shapes run three CTAs per SM (their dW_in transposition buffer is larger).
#define VB_TC_SHAPES(X) X(4, 12, 8, 2) X(3, 12, 8, 2) X(10, 12, 8, 5) X(8, 12, 8, 2)

bool has_tc_path(int I, int Ni, int Nc, int O) {
#define X(a, b, c, d) if (I == a && Ni == b && Nc == c && O == d) return true;
    VB_TC_SHAPES(X)
#undef X
    return false;
}

int64_t saved_bytes(int I, int Ni, int Nc, int O, int64_t B, int T) {
#define X(a, b, c, d) if (I == a && Ni == b && Nc == c && O == d) return saved_bytes_for<Cfg<a, b, c, d>>(B, T);
    VB_TC_SHAPES(X)
#undef X
    return 0;
}

int fwd(const float* packed, const float* x, const float* h0, float* y, float* h_traj, void* saved, int64_t saved_bytes_, int64_t B, int T,
        int I, int Ni, int Nc, int O, int layout, cudaStream_t stream) {
#define X(a, b, c, d) \
    if (I == a && Ni == b && Nc == c && O == d) return launch_fwd<Cfg<a, b, c, d>>(packed, x, h0, y, h_traj, saved, saved_bytes_, B, T, layout, stream);
    VB_TC_SHAPES(X)
#undef X
    set_error("no tensor-core kernel for (%d,%d,%d,%d)", I, Ni, Nc, O);
    return VB_ERR_BAD_DIMS;
}

int bwd(const float* packed, const float* x, const float* h0, const float* h_traj, const float* dy, const float* dh_traj,
        const float* dh_last, const void* saved, int64_t saved_bytes_, float* dx, float* dh0, float* dpacked, int64_t B, int T, int I, int Ni,
        int Nc, int O, int layout, void* ws, int64_t ws_bytes, cudaStream_t stream) {
#define X(a, b, c, d)                                            \
    if (I == a && Ni == b && Nc == c && O == d)                  \
        return launch_bwd<Cfg<a, b, c, d>>(packed, x, h0, h_traj, dy, dh_traj, dh_last, saved, saved_bytes_, dx, dh0, dpacked, B, T, layout, ws, ws_bytes, stream);
    VB_TC_SHAPES(X)
#undef X
    set_error("no tensor-core kernel for (%d,%d,%d,%d)", I, Ni, Nc, O);
    return VB_ERR_BAD_DIMS;
}

// scratch of the backward: per-CTA partial rows and flush buffers (+ room for a saved block when the caller passes none)
int64_t bwd_workspace_bytes(int I, int Ni, int Nc, int O, int64_t B, int T, bool with_saved) {
    LiquidLayout L(I, Ni, Nc, O);
    int sms = device_sms();
    if (sms <= 0) sms = 148;
    int64_t need = (int64_t)sms * 4 * (L.f_total + D3_FLOATS) * (int64_t)sizeof(float);
    need = (need + 255) & ~(int64_t)255;
    return need + (with_saved ? saved_bytes(I, Ni, Nc, O, B, T) : 0);
}

}  // namespace tc
}  // namespace vb
