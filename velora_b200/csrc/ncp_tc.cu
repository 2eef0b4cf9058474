// NCPNetwork (the SAC critics, velora/models/lnn/ncp.py:299-328) forward on the tensor cores, fp32-equivalent (bf16x3).
//
//   y = W3 mish(W2 mish(W1 x + b1) + b2) + b3          critic shapes: (5 | 4) -> 77 -> 51 -> (1 | 2)
//
// One CTA = 128 threads = 128 samples = the 128 TMEM lanes.  Layer 1 (K = 5) and layer 3 (N = 1 or 2) are a few FMAs per
// thread; layer 2 (77 x 51, 90 % of the MACs) is ONE accumulator: the three bf16 parts of a~1 = [mish(u1) | 1] are issued one
// after the other (K = 80 each) against the three weight parts stacked along N (3 x 64 = 192 rows), so column group pb holds
// a~1 W2_pb^T (+ bias through the constant-1 feature) and 15 MMAs of N = 192 replace 30 of N = 64.  The epilogue adds the
// three groups smallest first.  Used for large batches; the generic tile kernels (generic.cu) keep the small ones and the
// backward.
#include "common.cuh"
#include "tc_common.cuh"

namespace vb {
namespace ncp_tc {
using namespace tcx;

template <int I_, int NI_, int NC_, int O_>
struct Cfg {
    static constexpr int I = I_, NI = NI_, NC = NC_, O = O_;
    static constexpr int NIp = ceil4(NI), NCp = ceil4(NC), Op = ceil4(O);
    static constexpr int KC = (NI + 1 + 7) / 8;         // 8-feature chunks of a~1 = [a1 | 1 | 0..]
    static constexpr int KF = KC * 8;
    static constexpr int NP = (NC + 15) / 16 * 16;      // layer-2 outputs per part, padded for the MMA
    static constexpr int EC = (NC + 7) / 8;             // 8-column chunks the epilogue reads
    static_assert(KC % 2 == 0, "a~1 must be a whole number of K = 16 steps");
    static_assert(3 * NP <= 256, "stacked layer-2 width must fit one MMA");
    // packed block (NcpLayout, F section)
    static constexpr int W1_T = 0, B1 = W1_T + I * NIp, W2_T = B1 + NIp, B2 = W2_T + NI * NCp, W3 = B2 + NCp, B3 = W3 + O * NCp,
                         F_TOTAL = B3 + Op;
    // shared memory
    static constexpr int S_A = 0;                               // 3 parts x KC planes
    static constexpr int S_WB = S_A + 3 * KC * PLANE;           // [KC][3 NP rows][16 B]
    static constexpr int S_W1 = S_WB + KC * 3 * NP * 16;        // fp32 [I + 1][KF]: W1^T rows, then b1
    static constexpr int S_W3 = S_W1 + (I + 1) * KF * 4;        // fp32 [O][NP], then b3 [4]
    static constexpr int S_YS = S_W3 + (O * NP + 4) * 4;        // fp32 [O][TM]
    static constexpr int S_BAR = S_YS + O * TM * 4;
    static constexpr int BYTES = S_BAR + 64;
    static constexpr int TMEM_COLS = 256;
};

template <class C>
__global__ void __launch_bounds__(2 * TM) ncp_fwd_tc_kernel(const float* __restrict__ packed, const float* __restrict__ x,
                                                        float* __restrict__ y, int64_t B) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* aP = smem + C::S_A;
    unsigned char* wB = smem + C::S_WB;
    float* w1s = reinterpret_cast<float*>(smem + C::S_W1);
    float* w3s = reinterpret_cast<float*>(smem + C::S_W3);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + C::S_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::S_BAR + 16);
    // two threads per sample (rows of warps w and w + 4 share TMEM lanes): each takes every other plane chunk / column chunk
    constexpr int NT = 2 * TM;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & (TM - 1), half = tid >> 7;
    float* ys = reinterpret_cast<float*>(smem + C::S_YS);      // [O][TM]: layer-3 partial sums of the second half

    // layer-2 operand: B[pb * NP + n][k] = part pb of (k < NI ? W2[n][k] : k == NI ? b2[n] : 0), zero rows for n >= NC
    for (int item = tid; item < C::NP * C::KC; item += NT) {
        const int n = item % C::NP, kc = item / C::NP;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = 8 * kc + e;
            v[e] = n >= C::NC ? 0.f : (k < C::NI ? __ldg(packed + C::W2_T + k * C::NCp + n) : (k == C::NI ? __ldg(packed + C::B2 + n) : 0.f));
        }
        uint4 part[3];
        split3(v[0], v[1], part[0].x, part[1].x, part[2].x);
        split3(v[2], v[3], part[0].y, part[1].y, part[2].y);
        split3(v[4], v[5], part[0].z, part[1].z, part[2].z);
        split3(v[6], v[7], part[0].w, part[1].w, part[2].w);
        uint4* q = reinterpret_cast<uint4*>(wB);
#pragma unroll
        for (int pb = 0; pb < 3; ++pb) q[kc * (3 * C::NP) + pb * C::NP + n] = part[pb];
    }
    for (int e = tid; e < (C::I + 1) * C::KF; e += NT) {
        const int i = e / C::KF, j = e % C::KF;
        w1s[e] = j < C::NI ? __ldg(packed + (i < C::I ? C::W1_T + i * C::NIp : C::B1) + j) : 0.f;
    }
    for (int e = tid; e < C::O * C::NP + 4; e += NT) {
        const int o = e / C::NP, j = e % C::NP;
        w3s[e] = e < C::O * C::NP ? (j < C::NC ? __ldg(packed + C::W3 + o * C::NCp + j) : 0.f)
                                  : (e - C::O * C::NP < C::O ? __ldg(packed + C::B3 + e - C::O * C::NP) : 0.f);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t a_addr = smem_u32(aP), w_addr = smem_u32(wB);
    uint32_t ph = 0;

    const int64_t n_tiles = (B + TM - 1) / TM;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t b = tile * TM + row;
        const bool valid = b < B;
        float xr[C::I];
#pragma unroll
        for (int i = 0; i < C::I; ++i) xr[i] = valid ? __ldg(x + b * C::I + i) : 0.f;
        // ---------------- layer 1 + Mish on CUDA cores, eight neurons (one plane chunk) at a time
#pragma unroll 1
        for (int c = half; c < C::KC; c += 2) {
            float u[8];
            {
                const float4 b0 = *reinterpret_cast<const float4*>(w1s + C::I * C::KF + 8 * c);
                const float4 b1 = *reinterpret_cast<const float4*>(w1s + C::I * C::KF + 8 * c + 4);
                u[0] = b0.x; u[1] = b0.y; u[2] = b0.z; u[3] = b0.w; u[4] = b1.x; u[5] = b1.y; u[6] = b1.z; u[7] = b1.w;
            }
#pragma unroll
            for (int i = 0; i < C::I; ++i) {
                const float4 w0 = *reinterpret_cast<const float4*>(w1s + i * C::KF + 8 * c);
                const float4 w1 = *reinterpret_cast<const float4*>(w1s + i * C::KF + 8 * c + 4);
                u[0] = fmaf(xr[i], w0.x, u[0]); u[1] = fmaf(xr[i], w0.y, u[1]); u[2] = fmaf(xr[i], w0.z, u[2]); u[3] = fmaf(xr[i], w0.w, u[3]);
                u[4] = fmaf(xr[i], w1.x, u[4]); u[5] = fmaf(xr[i], w1.y, u[5]); u[6] = fmaf(xr[i], w1.z, u[6]); u[7] = fmaf(xr[i], w1.w, u[7]);
            }
            float a[8];
            {
                float u4[4] = {u[0], u[1], u[2], u[3]}, a4[4];
                mish4(u4, a4);
                a[0] = a4[0]; a[1] = a4[1]; a[2] = a4[2]; a[3] = a4[3];
                float v4[4] = {u[4], u[5], u[6], u[7]};
                mish4(v4, a4);
                a[4] = a4[0]; a[5] = a4[1]; a[6] = a4[2]; a[7] = a4[3];
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {      // the constant-1 (bias) feature and the zero padding behind it
                const int k = 8 * c + e;
                if (k == C::NI) a[e] = 1.0f;
                else if (k > C::NI) a[e] = 0.f;
            }
            store_chunk3<C::KC>(aP, c, row, a);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            constexpr uint32_t idesc = make_idesc(128, 3 * C::NP, 0, 0);
#pragma unroll
            for (int pa = 2; pa >= 0; --pa)          // smallest part first: the accumulator truncates
#pragma unroll
                for (int ks = 0; ks < C::KC / 2; ++ks)
                    mma_bf16(tmem, make_desc(a_addr + (pa * C::KC + 2 * ks) * PLANE, PLANE, 128),
                             make_desc(w_addr + 2 * ks * (3 * C::NP * 16), 3 * C::NP * 16, 128), idesc, !(pa == 2 && ks == 0));
            mma_commit(bar);
        }
        mbar_wait(bar, ph);
        ph ^= 1;
        tc_fence_after();
        // ---------------- Mish + layer 3
        float yo[C::O];
#pragma unroll
        for (int o = 0; o < C::O; ++o) yo[o] = half == 0 ? w3s[C::O * C::NP + o] : 0.f;
#pragma unroll 1
        for (int c = half; c < C::EC; c += 2) {
            float s8[8];
            {
                float hi[8], mid[8], lo[8];
                tmem_ld8(taddr + 8 * c, hi);
                tmem_ld8(taddr + C::NP + 8 * c, mid);
                tmem_ld8(taddr + 2 * C::NP + 8 * c, lo);
#pragma unroll
                for (int e = 0; e < 8; ++e) s8[e] = (lo[e] + mid[e]) + hi[e];
            }
            float a2[8];
            {
                float u4[4] = {s8[0], s8[1], s8[2], s8[3]}, a4[4];
                mish4(u4, a4);
                a2[0] = a4[0]; a2[1] = a4[1]; a2[2] = a4[2]; a2[3] = a4[3];
                float v4[4] = {s8[4], s8[5], s8[6], s8[7]};
                mish4(v4, a4);
                a2[4] = a4[0]; a2[5] = a4[1]; a2[6] = a4[2]; a2[7] = a4[3];
            }
#pragma unroll
            for (int o = 0; o < C::O; ++o) {
                const float4 w0 = *reinterpret_cast<const float4*>(w3s + o * C::NP + 8 * c);
                const float4 w1 = *reinterpret_cast<const float4*>(w3s + o * C::NP + 8 * c + 4);
                yo[o] = fmaf(a2[0], w0.x, fmaf(a2[1], w0.y, fmaf(a2[2], w0.z, fmaf(a2[3], w0.w, yo[o]))));
                yo[o] = fmaf(a2[4], w1.x, fmaf(a2[5], w1.y, fmaf(a2[6], w1.z, fmaf(a2[7], w1.w, yo[o]))));
            }
        }
        if (half == 1) {
#pragma unroll
            for (int o = 0; o < C::O; ++o) ys[o * TM + row] = yo[o];
        }
        tc_fence_before();      // this tile's TMEM reads are ordered before the next tile's MMAs (issued after the next barrier)
        __syncthreads();
        if (half == 0 && valid) {
#pragma unroll
            for (int o = 0; o < C::O; ++o) y[b * C::O + o] = yo[o] + ys[o * TM + row];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

// ================================================================================================ backward
// Recompute + all gradients for one 128-sample tile at a time, 1 CTA per SM (the accumulators take all 512 TMEM columns):
//   u2 = a~1 W~2^T                     15 MMAs, N = 192 (as in the forward)            columns [0, 192)
//   da1 = du2 W2                       6 part products x 4 K-steps, N = 80             columns [0, 80) (u2 is dead by then)
//   dW2 += [du2_1; du2_2]^T [a~1_1 | a~1_2 | a~1_3]      8 MMAs, N = 240 (six products)  columns [192, 432)
//   dW2 += [du2_3; *]^T a~1_1                            8 MMAs, N = 80                  columns [432, 512)
// The two weight-gradient accumulators live in TMEM across tiles and are flushed to an fp32 buffer every kFlushTiles tiles
// (the accumulator truncates).  Layers 1 and 3 (K = 5, N <= 2) and their gradients run on the CUDA cores: dW1, db1, dW3, db3
// are row sums over fp32 transposition strips, accumulated in registers across all tiles of the CTA.
constexpr int kFlushTiles = 8;
constexpr int SS = 132;      // strip row stride in floats (128 samples + 4: conflict-free 128-bit row reads)

template <class C>
struct BwdSmem {
    static constexpr int DC = C::NP / 8;                        // du2 chunks per part
    static constexpr int S_A = 0;                               // a~1: 3 parts x KC planes
    static constexpr int S_D = S_A + 3 * C::KC * PLANE;         // du2: 3 parts x DC planes; the [du2_3; *] stack over-reads DC planes
    static constexpr int S_WB = S_D + 3 * DC * PLANE;           // layer-2 operand (finite bytes behind the du2 planes)
    static constexpr int S_STRIP = S_WB + C::KC * 3 * C::NP * 16;   // fp32 strips: a2 [8 EC][SS], x [I][SS], dy [O][SS]
    static constexpr int S_W1 = S_STRIP + (8 * C::EC + 2 * C::I + C::O) * SS * 4;     // + dx partial sums [I][SS]
    static constexpr int S_W3 = S_W1 + (C::I + 1) * C::KF * 4;
    static constexpr int S_BAR = S_W3 + (C::O * C::NP + 4) * 4;
    static constexpr int BYTES = S_BAR + 64;
    static_assert(DC * PLANE <= C::KC * 3 * C::NP * 16, "over-read of the du2 stack must stay inside the weight operand");
    static_assert(C::KF * SS * 4 <= 3 * DC * PLANE, "du1 strip aliases the du2 planes");
    static constexpr int T_DWA = 3 * C::NP, T_DWB = T_DWA + 3 * C::KF;      // TMEM columns
    static_assert(T_DWB + C::KF <= 512, "TMEM budget");
    static constexpr int D3_COLS = 3 * C::KF + C::KF;           // flushed accumulator columns per lane
};

template <class C>
__global__ void __launch_bounds__(2 * TM, 1) ncp_bwd_tc_kernel(const float* __restrict__ packed, const float* __restrict__ x,
                                                           const float* __restrict__ dy, float* __restrict__ dx,
                                                           float* __restrict__ partials, float* __restrict__ d3buf, int64_t B,
                                                           int want_dw) {
    // want_dw == 0: dL/dx only (a critic inside the actor loss): the sixteen dW2 MMAs per tile, the dW1 / dW3 row sums and the accumulator
    // flushes are skipped
    using S = BwdSmem<C>;
    constexpr int DC = S::DC;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* aP = smem + S::S_A;
    unsigned char* dP = smem + S::S_D;
    unsigned char* wB = smem + S::S_WB;
    float* du1s = reinterpret_cast<float*>(dP);                           // [KF][SS], aliases the du2 planes between MMAs
    float* a2s = reinterpret_cast<float*>(smem + S::S_STRIP);             // [8 EC][SS]
    float* xs = a2s + 8 * C::EC * SS;                                     // [I][SS]
    float* dys = xs + C::I * SS;                                          // [O][SS]
    float* dxs = dys + C::O * SS;                                         // [I][SS]: dx partial sums of the second half
    float* w1s = reinterpret_cast<float*>(smem + S::S_W1);
    float* w3s = reinterpret_cast<float*>(smem + S::S_W3);
    uint64_t* bar1 = reinterpret_cast<uint64_t*>(smem + S::S_BAR);        // u2 ready
    uint64_t* bar2 = bar1 + 1;                                            // da1 ready
    uint64_t* bar3 = bar1 + 2;                                            // dW batch done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::S_BAR + 32);
    // TWO threads per sample: thread (row, half) works on the plane chunks c = half, half + 2, ...; both halves of a row sit in
    // warps w and w + 4, which address the same TMEM lanes.  With the accumulators taking all 512 TMEM columns only one CTA
    // fits per SM, so the second half doubles the warps that hide each other's latency.
    constexpr int NT = 2 * TM;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = tid & (TM - 1), half = tid >> 7;

    for (int item = tid; item < C::NP * C::KC; item += NT) {
        const int n = item % C::NP, kc = item / C::NP;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = 8 * kc + e;
            v[e] = n >= C::NC ? 0.f : (k < C::NI ? __ldg(packed + C::W2_T + k * C::NCp + n) : (k == C::NI ? __ldg(packed + C::B2 + n) : 0.f));
        }
        uint4 part[3];
        split3(v[0], v[1], part[0].x, part[1].x, part[2].x);
        split3(v[2], v[3], part[0].y, part[1].y, part[2].y);
        split3(v[4], v[5], part[0].z, part[1].z, part[2].z);
        split3(v[6], v[7], part[0].w, part[1].w, part[2].w);
        uint4* q = reinterpret_cast<uint4*>(wB);
#pragma unroll
        for (int pb = 0; pb < 3; ++pb) q[kc * (3 * C::NP) + pb * C::NP + n] = part[pb];
    }
    for (int e = tid; e < (C::I + 1) * C::KF; e += NT) {
        const int i = e / C::KF, j = e % C::KF;
        w1s[e] = j < C::NI ? __ldg(packed + (i < C::I ? C::W1_T + i * C::NIp : C::B1) + j) : 0.f;
    }
    for (int e = tid; e < C::O * C::NP + 4; e += NT) {
        const int o = e / C::NP, j = e % C::NP;
        w3s[e] = e < C::O * C::NP ? (j < C::NC ? __ldg(packed + C::W3 + o * C::NCp + j) : 0.f)
                                  : (e - C::O * C::NP < C::O ? __ldg(packed + C::B3 + e - C::O * C::NP) : 0.f);
    }
    // fp32 flush buffer of the two dW2 accumulators: element (lane = thread, column c) at ((c / 4) * TM + lane) * 4 + c % 4
    float* d3row = d3buf + (int64_t)blockIdx.x * (TM * S::D3_COLS) + row * 4;
    for (int q = half; q < S::D3_COLS / 4; q += 2) *reinterpret_cast<float4*>(d3row + q * (TM * 4)) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
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
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t a_addr = smem_u32(aP), d_addr = smem_u32(dP), w_addr = smem_u32(wB);
    uint32_t ph = 0;
    int since_flush = 0;

    // row sums on the CUDA cores: thread j < KF owns column j of dW1 / db1; thread KF - NI.. reuse: threads >= TM - 8 EC own dW3 rows
    float acc1[C::I + 1];
#pragma unroll
    for (int i = 0; i <= C::I; ++i) acc1[i] = 0.f;
    float acc3[C::O], accb3[C::O];
#pragma unroll
    for (int o = 0; o < C::O; ++o) { acc3[o] = 0.f; accb3[o] = 0.f; }
    static_assert(C::KF <= TM && 8 * C::EC <= TM, "row-sum ownership");
    const int n3 = (half == 1 && row < 8 * C::EC) ? row : -1;      // dW3 row of this thread (second half of the CTA), else negative

    auto flush_dw = [&]() {
        tc_fence_after();
#pragma unroll 1
        for (int g = half; g < S::D3_COLS / 32; g += 2) {
            float v[32];
            tmem_ld32(taddr + S::T_DWA + 32 * g, v);
#pragma unroll
            for (int q = 0; q < 8; ++q)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d3row + (8 * g + q) * (TM * 4)), "f"(v[4 * q]), "f"(v[4 * q + 1]),
                             "f"(v[4 * q + 2]), "f"(v[4 * q + 3]) : "memory");
        }
        since_flush = 0;
    };
    static_assert(S::D3_COLS % 32 == 0, "flush works on 32-column groups");

    const int64_t n_tiles = (B + TM - 1) / TM;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t b = tile * TM + row;
        const bool valid = b < B;
        float xr[C::I], dyr[C::O];
#pragma unroll
        for (int i = 0; i < C::I; ++i) xr[i] = valid ? __ldg(x + b * C::I + i) : 0.f;
#pragma unroll
        for (int o = 0; o < C::O; ++o) dyr[o] = valid ? __ldg(dy + b * C::O + o) : 0.f;
        if (half == 0) {
#pragma unroll
            for (int i = 0; i < C::I; ++i) xs[i * SS + row] = xr[i];
#pragma unroll
            for (int o = 0; o < C::O; ++o) dys[o * SS + row] = dyr[o];
        }
        // ---------------- layer 1 + Mish -> a~1 planes
#pragma unroll 1
        for (int c = half; c < C::KC; c += 2) {
            float u[8];
            {
                const float4 b0 = *reinterpret_cast<const float4*>(w1s + C::I * C::KF + 8 * c);
                const float4 b1 = *reinterpret_cast<const float4*>(w1s + C::I * C::KF + 8 * c + 4);
                u[0] = b0.x; u[1] = b0.y; u[2] = b0.z; u[3] = b0.w; u[4] = b1.x; u[5] = b1.y; u[6] = b1.z; u[7] = b1.w;
            }
#pragma unroll
            for (int i = 0; i < C::I; ++i) {
                const float4 w0 = *reinterpret_cast<const float4*>(w1s + i * C::KF + 8 * c);
                const float4 w1 = *reinterpret_cast<const float4*>(w1s + i * C::KF + 8 * c + 4);
                u[0] = fmaf(xr[i], w0.x, u[0]); u[1] = fmaf(xr[i], w0.y, u[1]); u[2] = fmaf(xr[i], w0.z, u[2]); u[3] = fmaf(xr[i], w0.w, u[3]);
                u[4] = fmaf(xr[i], w1.x, u[4]); u[5] = fmaf(xr[i], w1.y, u[5]); u[6] = fmaf(xr[i], w1.z, u[6]); u[7] = fmaf(xr[i], w1.w, u[7]);
            }
            float a[8];
            {
                float u4[4] = {u[0], u[1], u[2], u[3]}, a4[4];
                mish4(u4, a4);
                a[0] = a4[0]; a[1] = a4[1]; a[2] = a4[2]; a[3] = a4[3];
                float v4[4] = {u[4], u[5], u[6], u[7]};
                mish4(v4, a4);
                a[4] = a4[0]; a[5] = a4[1]; a[6] = a4[2]; a[7] = a4[3];
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int k = 8 * c + e;
                if (k == C::NI) a[e] = 1.0f;
                else if (k > C::NI) a[e] = 0.f;
            }
            store_chunk3<C::KC>(aP, c, row, a);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            constexpr uint32_t idesc = make_idesc(128, 3 * C::NP, 0, 0);
#pragma unroll
            for (int pa = 2; pa >= 0; --pa)
#pragma unroll
                for (int ks = 0; ks < C::KC / 2; ++ks)
                    mma_bf16(tmem, make_desc(a_addr + (pa * C::KC + 2 * ks) * PLANE, PLANE, 128),
                             make_desc(w_addr + 2 * ks * (3 * C::NP * 16), 3 * C::NP * 16, 128), idesc, !(pa == 2 && ks == 0));
            mma_commit(bar1);
        }
        mbar_wait(bar1, ph);
        tc_fence_after();
        // ---------------- Mish, Mish', du2 = (dy W3) * Mish'(u2) -> du2 planes; a2 -> strip
#pragma unroll 1
        for (int c = half; c < DC; c += 2) {
            float du2[8];
            if (c < C::EC) {
                float s8[8];
                {
                    float hi[8], mid[8], lo[8];
                    tmem_ld8(taddr + 8 * c, hi);
                    tmem_ld8(taddr + C::NP + 8 * c, mid);
                    tmem_ld8(taddr + 2 * C::NP + 8 * c, lo);
#pragma unroll
                    for (int e = 0; e < 8; ++e) s8[e] = (lo[e] + mid[e]) + hi[e];
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    float dm;
                    const float a2 = mish_fwd_grad(s8[e], dm);
                    a2s[(8 * c + e) * SS + row] = a2;
                    float da = 0.f;
#pragma unroll
                    for (int o = 0; o < C::O; ++o) da = fmaf(dyr[o], w3s[o * C::NP + 8 * c + e], da);
                    du2[e] = (8 * c + e < C::NC) ? da * dm : 0.f;
                }
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) du2[e] = 0.f;
            }
            store_chunk3<DC>(dP, c, row, du2);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (lane == 0) {
            if (warp == 1) {            // da1 = du2 W2: six part products, smallest first
                tc_fence_after();
                constexpr uint32_t idesc2 = make_idesc(128, C::KF, 0, 1);
#pragma unroll
                for (int p = 0; p < 6; ++p) {
                    const int pa = (p == 0) ? 1 : (p == 1) ? 0 : (p == 2) ? 2 : (p == 3) ? 0 : (p == 4) ? 1 : 0;
                    const int pb = (p == 0) ? 1 : (p == 1) ? 2 : (p == 2) ? 0 : (p == 3) ? 1 : (p == 4) ? 0 : 0;
#pragma unroll
                    for (int ks = 0; ks < DC / 2; ++ks)
                        mma_bf16(tmem, make_desc(d_addr + (pa * DC + 2 * ks) * PLANE, PLANE, 128),
                                 make_desc(w_addr + (pb * C::NP + ks * 16) * 16, 128, 3 * C::NP * 16), idesc2, (p | ks) != 0);
                }
                mma_commit(bar2);
            } else if (warp == 2 && want_dw) {     // dW2 (+ db2 through the constant-1 feature)
                tc_fence_after();
                constexpr uint32_t idescA = make_idesc(128, 3 * C::KF, 1, 1);
                constexpr uint32_t idescB = make_idesc(128, C::KF, 1, 1);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    mma_bf16(tmem + S::T_DWA, make_desc(d_addr + ks * 256, 128, PLANE), make_desc(a_addr + ks * 256, 128, PLANE), idescA,
                             !(since_flush == 0 && ks == 0));
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    mma_bf16(tmem + S::T_DWB, make_desc(d_addr + 2 * DC * PLANE + ks * 256, 128, PLANE), make_desc(a_addr + ks * 256, 128, PLANE),
                             idescB, !(since_flush == 0 && ks == 0));
                mma_commit(bar3);
            }
        }
        since_flush += 1;
        mbar_wait(bar2, ph);
        if (want_dw) mbar_wait(bar3, ph);        // the du2 planes are free: the du1 strip may overwrite them (without dW2 the da1 MMAs were their only reader)
        ph ^= 1;
        tc_fence_after();
        // ---------------- du1 = da1 * Mish'(u1) (u1 recomputed), dx = du1 W1, du1 -> strip
        float dxr[C::I];
#pragma unroll
        for (int i = 0; i < C::I; ++i) dxr[i] = 0.f;
#pragma unroll 1
        for (int c = half; c < C::KC; c += 2) {
            float u[8], da1[8];
            tmem_ld8(taddr + 8 * c, da1);
            {
                const float4 b0 = *reinterpret_cast<const float4*>(w1s + C::I * C::KF + 8 * c);
                const float4 b1 = *reinterpret_cast<const float4*>(w1s + C::I * C::KF + 8 * c + 4);
                u[0] = b0.x; u[1] = b0.y; u[2] = b0.z; u[3] = b0.w; u[4] = b1.x; u[5] = b1.y; u[6] = b1.z; u[7] = b1.w;
            }
            float w[C::I][8];
#pragma unroll
            for (int i = 0; i < C::I; ++i) {
                const float4 w0 = *reinterpret_cast<const float4*>(w1s + i * C::KF + 8 * c);
                const float4 w1 = *reinterpret_cast<const float4*>(w1s + i * C::KF + 8 * c + 4);
                w[i][0] = w0.x; w[i][1] = w0.y; w[i][2] = w0.z; w[i][3] = w0.w; w[i][4] = w1.x; w[i][5] = w1.y; w[i][6] = w1.z; w[i][7] = w1.w;
#pragma unroll
                for (int e = 0; e < 8; ++e) u[e] = fmaf(xr[i], w[i][e], u[e]);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float dm;
                (void)mish_fwd_grad(u[e], dm);
                const float du1 = (8 * c + e < C::NI) ? da1[e] * dm : 0.f;
                du1s[(8 * c + e) * SS + row] = du1;
#pragma unroll
                for (int i = 0; i < C::I; ++i) dxr[i] = fmaf(du1, w[i][e], dxr[i]);
            }
        }
        if (half == 1) {          // the two halves hold partial sums over their chunks
#pragma unroll
            for (int i = 0; i < C::I; ++i) dxs[i * SS + row] = dxr[i];
        }
        tc_fence_before();
        __syncthreads();
        if (half == 0 && valid && dx) {
#pragma unroll
            for (int i = 0; i < C::I; ++i) dx[b * C::I + i] = dxr[i] + dxs[i * SS + row];
        }
        // ---------------- row sums: dW1 / db1 (thread j < KF), dW3 / db3 (the last 8 EC threads)
        if (want_dw && tid < C::KF) {
            const float4* dr = reinterpret_cast<const float4*>(du1s + tid * SS);
#pragma unroll 4
            for (int q = 0; q < TM / 4; ++q) {
                const float4 d = dr[q];
#pragma unroll
                for (int i = 0; i < C::I; ++i) {
                    const float4 xv = *reinterpret_cast<const float4*>(xs + i * SS + 4 * q);
                    acc1[i] = fmaf(xv.x, d.x, fmaf(xv.y, d.y, fmaf(xv.z, d.z, fmaf(xv.w, d.w, acc1[i]))));
                }
                acc1[C::I] += (d.x + d.y) + (d.z + d.w);
            }
        }
        if (want_dw && n3 >= 0) {
            const float4* ar = reinterpret_cast<const float4*>(a2s + n3 * SS);
#pragma unroll 4
            for (int q = 0; q < TM / 4; ++q) {
                const float4 a = ar[q];
#pragma unroll
                for (int o = 0; o < C::O; ++o) {
                    const float4 dv = *reinterpret_cast<const float4*>(dys + o * SS + 4 * q);
                    acc3[o] = fmaf(dv.x, a.x, fmaf(dv.y, a.y, fmaf(dv.z, a.z, fmaf(dv.w, a.w, acc3[o]))));
                    accb3[o] += (dv.x + dv.y) + (dv.z + dv.w);
                }
            }
        }
        if (want_dw && since_flush >= kFlushTiles) flush_dw();
        __syncthreads();      // strips and planes are rewritten by the next tile
    }
    if (want_dw && since_flush > 0) flush_dw();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");

    // ---------------- per-CTA partial gradient row (F layout of NcpLayout)
    float* prow = partials + (int64_t)blockIdx.x * C::F_TOTAL;
    for (int idx = tid; idx < C::F_TOTAL; idx += NT) prow[idx] = 0.f;
    __syncthreads();
    if (tid < C::NI) {
#pragma unroll
        for (int i = 0; i < C::I; ++i) prow[C::W1_T + i * C::NIp + tid] = acc1[i];
        prow[C::B1 + tid] = acc1[C::I];
    }
    if (n3 >= 0 && n3 < C::NC) {
#pragma unroll
        for (int o = 0; o < C::O; ++o) prow[C::W3 + o * C::NCp + n3] = acc3[o];
    }
    if (n3 == 0) {
#pragma unroll
        for (int o = 0; o < C::O; ++o) prow[C::B3 + o] = accb3[o];
    }
    {
        const float* d3 = d3buf + (int64_t)blockIdx.x * (TM * S::D3_COLS);
        auto at = [&](int row, int col) { return d3[((col >> 2) * TM + row) * 4 + (col & 3)]; };
        // accumulator A: rows [0, NP) = du2 part 1, [NP, 2 NP) = part 2; columns pb * KF + k.  Accumulator B: rows [0, NP) = part 3, columns 3 KF + k.
        for (int e = tid; e < C::NC * (C::NI + 1); e += NT) {
            const int n = e / (C::NI + 1), k = e % (C::NI + 1);
            float acc = at(C::NP + n, 2 * C::KF + k);                 // (2,3)
            acc += at(n, 2 * C::KF + k);                              // (1,3)
            acc += at(n, 3 * C::KF + k);                              // (3,1)
            acc += at(C::NP + n, C::KF + k);                          // (2,2)
            acc += at(C::NP + n, k);                                  // (2,1)
            acc += at(n, C::KF + k);                                  // (1,2)
            acc += at(n, k);                                          // (1,1)
            if (k < C::NI) prow[C::W2_T + k * C::NCp + n] = acc;
            else prow[C::B2 + n] = acc;
        }
    }
}

__global__ void __launch_bounds__(1024) reduce_rows_tc_kernel(const float* __restrict__ partials, int n_rows, int n, float* __restrict__ out) {
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

#define VB_NCP_TC_SHAPES(X) X(5, 77, 51, 1) X(4, 77, 51, 2)

bool has_path(int I, int Ni, int Nc, int O) {
#define X(a, b, c, d) if (I == a && Ni == b && Nc == c && O == d) return true;
    VB_NCP_TC_SHAPES(X)
#undef X
    return false;
}

template <class C>
static int launch_fwd(const float* packed, const float* x, float* y, int64_t B, cudaStream_t stream) {
    LaunchCfg cfg;
    int rc = cached_launch_cfg(ncp_fwd_tc_kernel<C>, 2 * TM, C::BYTES, &cfg);
    if (rc) return rc;
    int smem_sm = 0;
    int dev = 0;
    VB_CUDA(cudaGetDevice(&dev));
    VB_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
    int per_sm = smem_sm / (C::BYTES + cfg.static_smem + 1024);      // the occupancy API answers 1 for kernels that allocate TMEM
    if (per_sm > 512 / C::TMEM_COLS) per_sm = 512 / C::TMEM_COLS;
    if (per_sm < 1) per_sm = 1;
    const int64_t n_tiles = (B + TM - 1) / TM;
    const int64_t cap = (int64_t)device_sms() * per_sm;
    const int grid = (int)(n_tiles < cap ? n_tiles : cap);
    ncp_fwd_tc_kernel<C><<<grid < 1 ? 1 : grid, 2 * TM, C::BYTES, stream>>>(packed, x, y, B);
    VB_LAUNCH_CHECK();
    return 0;
}

int fwd(const float* packed, const float* x, float* y, int64_t B, int I, int Ni, int Nc, int O, cudaStream_t stream) {
#define X(a, b, c, d) if (I == a && Ni == b && Nc == c && O == d) return launch_fwd<Cfg<a, b, c, d>>(packed, x, y, B, stream);
    VB_NCP_TC_SHAPES(X)
#undef X
    set_error("no tensor-core NCP kernel for (%d,%d,%d,%d)", I, Ni, Nc, O);
    return VB_ERR_BAD_DIMS;
}


template <class C>
static int launch_bwd(const float* packed, const float* x, const float* dy, float* dx, float* dpacked, int64_t B,
                      void* ws, int64_t ws_bytes, cudaStream_t stream) {
    using S = BwdSmem<C>;
    LaunchCfg cfg;
    int rc = cached_launch_cfg(ncp_bwd_tc_kernel<C>, 2 * TM, S::BYTES, &cfg);
    if (rc) return rc;
    const int64_t n_tiles = (B + TM - 1) / TM;
    const int64_t cap = device_sms();
    const int grid = (int)(n_tiles < cap ? n_tiles : cap);
    const int64_t need = (int64_t)grid * (C::F_TOTAL + TM * S::D3_COLS) * (int64_t)sizeof(float);
    VB_CHECK_ARG(ws && ws_bytes >= need, VB_ERR_WORKSPACE, "vb_ncp_bwd: workspace too small (%lld bytes, need %lld)", (long long)ws_bytes,
                 (long long)need);
    float* partials = reinterpret_cast<float*>(ws);
    float* d3buf = partials + (int64_t)grid * C::F_TOTAL;
    ncp_bwd_tc_kernel<C><<<grid, 2 * TM, S::BYTES, stream>>>(packed, x, dy, dx, partials, d3buf, B, dpacked != nullptr);
    VB_LAUNCH_CHECK();
    if (dpacked)      // NULL: dL/dx only; the per-CTA partial rows stay in the workspace, unread
        reduce_rows_tc_kernel<<<(C::F_TOTAL + 31) / 32, 1024, 0, stream>>>(partials, grid, C::F_TOTAL, dpacked);
    VB_LAUNCH_CHECK();
    return 0;
}

int bwd(const float* packed, const float* x, const float* dy, float* dx, float* dpacked, int64_t B, int I, int Ni, int Nc, int O,
        void* ws, int64_t ws_bytes, cudaStream_t stream) {
#define X(a, b, c, d) if (I == a && Ni == b && Nc == c && O == d) return launch_bwd<Cfg<a, b, c, d>>(packed, x, dy, dx, dpacked, B, ws, ws_bytes, stream);
    VB_NCP_TC_SHAPES(X)
#undef X
    set_error("no tensor-core NCP kernel for (%d,%d,%d,%d)", I, Ni, Nc, O);
    return VB_ERR_BAD_DIMS;
}

int64_t bwd_workspace_bytes(int I, int Ni, int Nc, int O) {
    int sms = device_sms();
    if (sms <= 0) sms = 148;
#define X(a, b, c, d) if (I == a && Ni == b && Nc == c && O == d) return (int64_t)sms * (Cfg<a, b, c, d>::F_TOTAL + TM * BwdSmem<Cfg<a, b, c, d>>::D3_COLS) * 4;
    VB_NCP_TC_SHAPES(X)
#undef X
    return 0;
}

}  // namespace ncp_tc
}  // namespace vb
