// Backward of the 512-neuron class (BASELINE config 4: LiquidNCPNetwork(8,512,1)) in bf16 on the tensor cores (fp32 accumulation,
// 2e-2 contract).  The forward (liquid_big.cu, training mode) saved the gate factors G, m = mish(c) and z~; with those the backward
// splits into ONE serial piece and three batched contractions over all R = B * T sample-steps:
//
//   1. liquid_big_rec_kernel (hand-written tcgen05, this file) - the recurrence.  Per 128-sample tile and step, in reverse time:
//        ds = dh' * (A, Bc, Cc) | dm * mish'(c)   from G, dy, W_out and the carried dh' (element-wise, three threads per sample)
//        dh_{t-1} = ds W_h                         [128 x 816] x [816 x 208] on the tensor core, accumulator in TMEM (double-buffered)
//      ds is produced in 17 K-stages of 48 features straight into a shared-memory operand ring while the 339 KB W_h block streams
//      from L2 through a bulk-copy ring, so the MMAs of stage s run while the element-wise warps build stage s + 1; ds is also
//      written to HBM (bf16, row-major [R][816]) for the batched pieces.  Only the h-part of dz belongs to the recurrence.
//   2. dz_a = ds W_a [R x 816] x [816 x 312]  (plain GEMM, cuBLAS),  then liquid_big_inter_bwd_kernel: du = dz_a * mish'(W_in x + b),
//      dx = du W_in;  dW_in | db_in = [x | 1]^T du  (plain split-K GEMM, cuBLAS)
//   3. dW_heads | db_heads = ds^T z~  [816 x R] x [R x 528]  (plain split-K GEMM, cuBLAS; the bias rides on z~'s constant-1 feature)
//   4. dW_out | db_out = [m | 1]^T dy  (liquid_big_dwout_kernel)
// and liquid_big_unpack_kernel sums the split-K partials into the packed gradient block (F layout), which vb_liquid_unpack_grads maps
// to the fourteen parameter gradients like every other path.  The three library GEMMs are exactly that - dense, unfused, no epilogue -
// which is what cuBLAS is for; everything with structure (the recurrence, the gating, the saved-activation formats) is this library's.
//
// Replaces the autograd backward of cell.py:113-169 / ncp.py:129-178 at (8,512,1) for bf16 training.
#include <cublas_v2.h>

#include "big_common.cuh"

namespace vb {
namespace big {

// ------------------------------------------------------------------------------------------------ weight re-packs (bf16)
constexpr int R_STAGES = DSW / 48;                         // 17 K-stages of 48 ds features = 12 neurons
static_assert(R_STAGES * 48 == DSW && R_STAGES * 3 == NQUAD, "one quad per element-wise thread and stage");
constexpr size_t kWhpBytes = (size_t)R_STAGES * STAGE_BYTES;          // [stage][chunk (6)][n (208)][8 bf16]
constexpr size_t kWaBytes = (size_t)DSW * AW * 2;                     // row-major [816][312]

// whp: B operand of dh = ds W_h, K-major: value(k = 4 j + hs, n = j') = weight of hidden input j' into head hs of neuron j
// wa : row-major [k = 4 j + hs][i] = weight of inter input i into head hs of neuron j (i < 308, zero padded to 312)
__global__ void __launch_bounds__(256) big_bwd_pack_kernel(const float* __restrict__ packed, __nv_bfloat16* __restrict__ whp,
                                                           __nv_bfloat16* __restrict__ wa) {
    constexpr LiquidLayout L(I, NI, NC, 1);
    const int total_h = R_STAGES * STAGE_CHUNKS * PROWS * 8, total_a = DSW * AW;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total_h + total_a; e += gridDim.x * blockDim.x) {
        if (e < total_h) {
            const int el = e & 7, n = (e >> 3) % PROWS, c = (e >> 3) / PROWS;          // c = K-chunk 0..101
            const int k = 8 * c + el, j = k >> 2, hs = k & 3;
            float v = 0.f;
            if (n < NC) v = packed[L.wh_t + (NI + n) * L.H4 + hs * L.NCp + j];
            whp[e] = __float2bfloat16(v);
        } else {
            const int f = e - total_h, i = f % AW, k = f / AW, j = k >> 2, hs = k & 3;
            wa[f] = __float2bfloat16(i < NI ? packed[L.wh_t + i * L.H4 + hs * L.NCp + j] : 0.f);
        }
    }
}

// ------------------------------------------------------------------------------------------------ 1. the recurrence
constexpr int R_EW = 3 * TM;                               // element-wise threads: three per sample, one neuron quad per thread and stage
constexpr int R_NTHREADS = R_EW + 64;                      // + producer warp + MMA warp
constexpr int A_RING = 6;                                  // ds operand stages in flight (6 x 12 KB)
constexpr int A_STAGE_BYTES = STAGE_CHUNKS * PLANE;        // 48 features x 128 rows x 2 B
constexpr int R_S_A = 0;
constexpr int R_S_B = R_S_A + A_RING * A_STAGE_BYTES;
constexpr int G_RING = 4;                                  // saved-factor stages in flight: 4 x 12 KB per SM covers the DRAM round trip
constexpr int G_STAGE_BYTES = 3 * TM * 32;                 // the three quads of a stage x 128 rows x 16 bf16: contiguous in the saved block
constexpr int R_S_G = R_S_B + RING * STAGE_BYTES;
constexpr int R_S_WOUT = R_S_G + G_RING * G_STAGE_BYTES;   // fp32 [208]
constexpr int R_S_BAR = R_S_WOUT + 208 * 4;
constexpr int R_SMEM_BYTES = R_S_BAR + 512;

__global__ void __launch_bounds__(R_NTHREADS, 1) liquid_big_rec_kernel(const unsigned char* __restrict__ whp, const float* __restrict__ packed,
                                                                     const unsigned char* __restrict__ saved, const float* __restrict__ dy,
                                                                     const float* __restrict__ dh_last, unsigned char* __restrict__ ds_out,
                                                                     float* __restrict__ dh0, int64_t B, int T, int flags) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* aP = smem + R_S_A;
    float* wouts = reinterpret_cast<float*>(smem + R_S_WOUT);
    uint64_t* full_b = reinterpret_cast<uint64_t*>(smem + R_S_BAR);   // [RING]    weight stage landed
    uint64_t* empty_b = full_b + RING;                                // [RING]    its MMAs completed
    uint64_t* a_full = empty_b + RING;                                // [A_RING]  ds stage written by all element-wise threads
    uint64_t* a_empty = a_full + A_RING;                              // [A_RING]  its MMAs completed
    uint64_t* acc_full = a_empty + A_RING;                            // [2]       dh of a step complete
    uint64_t* acc_empty = acc_full + 2;                               // [2]       everybody has read that dh
    uint64_t* g_full = acc_empty + 2;                                 // [G_RING]  saved-factor stage landed
    uint64_t* g_empty = g_full + G_RING;                              // [G_RING]  every element-wise thread has read it
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(g_empty + G_RING);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    constexpr LiquidLayout L(I, NI, NC, 1);
    for (int e = tid; e < 208; e += R_NTHREADS) wouts[e] = e < NC ? __ldg(packed + L.wout + e) : 0.f;
    const bool y_tm = flags & VB_FLAG_Y_TIME_MAJOR;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < RING; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
        for (int s = 0; s < A_RING; ++s) { mbar_init(&a_full[s], R_EW); mbar_init(&a_empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], R_EW); }
        for (int s = 0; s < G_RING; ++s) { mbar_init(&g_full[s], 1); mbar_init(&g_empty[s], R_EW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int64_t n_tiles = (B + TM - 1) / TM;
    const int64_t my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const int64_t n_steps = my_tiles * T;
    const int64_t R = saved_rows(B, T);
    const unsigned char* sGb = saved + saved_g_off(R);

    if (warp == R_EW / 32) {
        // ---------------------------------------------------------------- producer: per stage one W_h block (the same 17 every step, from
        // L2) and the tile-step's saved gate factors of the stage's three quads (12 KB, contiguous, from HBM).  With the factors loaded
        // by the element-wise threads themselves (two LDG.128 each, even two stages ahead) the kernel had ~24 KB in flight per SM and
        // spent 1 ms of its 2.15 ms waiting for them (profiles/r02); the bulk ring keeps 48 KB in flight and costs no registers.
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                for (int t = T - 1; t >= 0; --t) {
                    const size_t ts = (size_t)((int64_t)t * n_tiles + tile);
                    for (int s = 0; s < R_STAGES; ++s, ++it) {
                        {
                            const int slot = it % G_RING;
                            const uint32_t use = it / G_RING;
                            if (use > 0) mbar_wait(&g_empty[slot], (use - 1) & 1);
                            mbar_expect_tx(&g_full[slot], G_STAGE_BYTES);
                            bulk_g2s(smem + R_S_G + slot * G_STAGE_BYTES, sGb + (ts * NQUAD + 3 * s) * (size_t)(TM * 32), G_STAGE_BYTES, &g_full[slot]);
                        }
                        const int slot = it % RING;
                        const uint32_t use = it / RING;
                        if (use > 0) mbar_wait(&empty_b[slot], (use - 1) & 1);
                        mbar_expect_tx(&full_b[slot], STAGE_BYTES);
                        bulk_g2s(smem + R_S_B + slot * STAGE_BYTES, whp + (size_t)s * STAGE_BYTES, STAGE_BYTES, &full_b[slot]);
                    }
                }
        }
    } else if (warp == R_EW / 32 + 1) {
        // ---------------------------------------------------------------- MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(128, PROWS, 0, 0);
            const uint32_t a_addr = smem_u32(aP), b_addr = smem_u32(smem + R_S_B);
            uint32_t it = 0;
            for (int64_t st = 0; st < n_steps; ++st) {
                const int buf = st & 1;
                // The dh this buffer held (step st - 2) was read during step st - 1.  The element-wise threads arrive on
                // acc_empty[(s + 1) & 1] at the end of EVERY step s (also a tile's first step, which reads nothing), so before step st
                // buffer `buf` has seen n = st / 2 (buf 0) or (st + 1) / 2 (buf 1) completions: wait for the n-th.
                {
                    const int64_t n = buf ? (st + 1) >> 1 : st >> 1;
                    if (n > 0) { mbar_wait(&acc_empty[buf], (uint32_t)((n - 1) & 1)); tc_fence_after(); }
                }
                for (int s = 0; s < R_STAGES; ++s, ++it) {
                    const int sa = it % A_RING, sb = it % RING;
                    mbar_wait(&a_full[sa], (it / A_RING) & 1);
                    mbar_wait(&full_b[sb], (it / RING) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < STAGE_CHUNKS / 2; ++kk)
                        mma_bf16(tmem + buf * PROWS, make_desc(a_addr + sa * A_STAGE_BYTES + 2 * kk * PLANE, PLANE, 128),
                                 make_desc(b_addr + sb * STAGE_BYTES + 2 * kk * (PROWS * 16), PROWS * 16, 128), idesc, (s | kk) != 0);
                    mma_commit(&empty_b[sb]);
                    mma_commit(&a_empty[sa]);
                }
                mma_commit(&acc_full[buf]);
            }
        }
    } else {
        // ---------------------------------------------------------------- element-wise warps: three threads per sample
        const int row = tid & (TM - 1), part = tid >> 7;          // part 0..2: quad 3 s + part of stage s
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t ait = 0;
        int64_t st = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t b = tile * TM + row;
            const bool valid = b < B;
            for (int t = T - 1; t >= 0; --t, ++st) {
                const bool first = t == T - 1;
                const int prev = (st + 1) & 1;                    // buffer of step st - 1
                const float dyv = valid ? __ldg(dy + (y_tm ? (int64_t)t * B + b : b * T + t)) : 0.f;
                const size_t ts = (size_t)((int64_t)t * n_tiles + tile);
                unsigned char* ds_row = ds_out + (ts * TM + row) * (size_t)(DSW * 2);
                if (!first) {
                    mbar_wait(&acc_full[prev], ((st - 1) >> 1) & 1);
                    tc_fence_after();
                }
#pragma unroll 1
                for (int s = 0; s < R_STAGES; ++s) {
                    const int q = 3 * s + part;
                    uint4 g0, g1;
                    {
                        const int gs = ait % G_RING;
                        mbar_wait(&g_full[gs], (ait / G_RING) & 1);
                        const uint4* gp = reinterpret_cast<const uint4*>(smem + R_S_G + gs * G_STAGE_BYTES) + (part * TM + row) * 2;
                        g0 = gp[0];
                        g1 = gp[1];
                        mbar_arrive(&g_empty[gs]);
                    }
                    float dhc[4];
                    if (first) {
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (valid && dh_last) v = __ldg(reinterpret_cast<const float4*>(dh_last + b * NC + 4 * q));
                        dhc[0] = v.x; dhc[1] = v.y; dhc[2] = v.z; dhc[3] = v.w;
                    } else {
                        tmem_ld4(taddr + prev * PROWS + 4 * q, dhc);
                    }
                    const float4 wo = *reinterpret_cast<const float4*>(wouts + 4 * q);
                    const float wv[4] = {wo.x, wo.y, wo.z, wo.w};
                    const uint32_t gw[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                    uint32_t dsw[8];
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        // (A, Bc) in gw[2 jj], (Cc, mish'(c)) in gw[2 jj + 1]; a bf16 is the upper half of the fp32 with the same value
                        const float fa = __uint_as_float(gw[2 * jj] << 16), fb = __uint_as_float(gw[2 * jj] & 0xffff0000u);
                        const float fc = __uint_as_float(gw[2 * jj + 1] << 16), fd = __uint_as_float(gw[2 * jj + 1] & 0xffff0000u);
                        const float dc = dyv * wv[jj] * fd;
                        const float dhn = dc + dhc[jj];
                        dsw[2 * jj] = pack_bf16(dhn * fa, dhn * fb);          // ds_g, ds_h
                        dsw[2 * jj + 1] = pack_bf16(dhn * fc, dc);            // ds_f (both f heads), ds_proj
                    }
                    const uint4 c0 = make_uint4(dsw[0], dsw[1], dsw[2], dsw[3]), c1 = make_uint4(dsw[4], dsw[5], dsw[6], dsw[7]);
                    const int sa = ait % A_RING;
                    const uint32_t use = ait / A_RING;
                    if (use > 0) mbar_wait(&a_empty[sa], (use - 1) & 1);
                    uint4* ap = reinterpret_cast<uint4*>(aP + sa * A_STAGE_BYTES);
                    ap[(2 * part) * TM + row] = c0;
                    ap[(2 * part + 1) * TM + row] = c1;
                    fence_async_smem();
                    mbar_arrive(&a_full[sa]);
                    reinterpret_cast<uint4*>(ds_row)[2 * q] = c0;
                    reinterpret_cast<uint4*>(ds_row)[2 * q + 1] = c1;
                    ++ait;
                }
                tc_fence_before();
                mbar_arrive(&acc_empty[prev]);
            }
            // dL/dh0 = the dh of the tile's last step (t = 0)
            if (dh0) {
                const int64_t sl = st - 1;
                mbar_wait(&acc_full[sl & 1], (sl >> 1) & 1);
                tc_fence_after();
                for (int s = 0; s < R_STAGES; ++s) {
                    const int q = 3 * s + part;
                    float v[4];
                    tmem_ld4(taddr + (int)(sl & 1) * PROWS + 4 * q, v);
                    if (valid) *reinterpret_cast<float4*>(dh0 + b * NC + 4 * q) = make_float4(v[0], v[1], v[2], v[3]);
                }
                tc_fence_before();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// ------------------------------------------------------------------------------------------------ 2. inter layer backward
// One thread per row r: du = dz_a * mish'(W_in x + b_in) (written over dz_a, bf16), dx = du W_in, xr = [x | 1 | 0..] (bf16, the split-K
// GEMM's first operand).  Weights are read as broadcast LDS.128 (every lane of a warp is at the same inter neuron).
constexpr int IB_THREADS = 128;
constexpr int IB_SMEM = (I * AW + AW + AW * I) * 4;        // win_t [8][312] | b_in [312] | win [312][8]

template <bool WITH_DX>
__global__ void __launch_bounds__(IB_THREADS) liquid_big_inter_bwd_kernel(const float* __restrict__ packed, const float* __restrict__ x,
                                                                          __nv_bfloat16* __restrict__ dza, __nv_bfloat16* __restrict__ xr,
                                                                          float* __restrict__ dx, int64_t B, int T, int flags) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* wt = reinterpret_cast<float*>(smem);            // [I][AW]
    float* bi = wt + I * AW;                               // [AW]
    float* wi = bi + AW;                                   // [AW][I]
    constexpr LiquidLayout L(I, NI, NC, 1);
    for (int e = threadIdx.x; e < I * AW; e += IB_THREADS) wt[e] = (e % AW) < NI ? __ldg(packed + L.win_t + (e / AW) * L.NIp + e % AW) : 0.f;
    for (int e = threadIdx.x; e < AW; e += IB_THREADS) bi[e] = e < NI ? __ldg(packed + L.b_in + e) : 0.f;
    for (int e = threadIdx.x; e < AW * I; e += IB_THREADS) wi[e] = (e / I) < NI ? __ldg(packed + L.win + (e / I) * L.Ip + e % I) : 0.f;
    __syncthreads();
    const bool x_tm = flags & VB_FLAG_X_TIME_MAJOR, dx_tm = flags & VB_FLAG_DX_TIME_MAJOR;
    const int64_t n_tiles = (B + TM - 1) / TM, R = saved_rows(B, T);
    for (int64_t r = (int64_t)blockIdx.x * IB_THREADS + threadIdx.x; r < R; r += (int64_t)gridDim.x * IB_THREADS) {
        const int row = (int)(r % TM);
        const int64_t ts = r / TM, tile = ts % n_tiles;
        const int t = (int)(ts / n_tiles);
        const int64_t b = tile * TM + row;
        const bool valid = b < B;
        float xv[I];
        {
            float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
            if (valid) {
                const float4* xp = reinterpret_cast<const float4*>(x + (x_tm ? (int64_t)t * B + b : b * T + t) * I);
                v0 = __ldg(xp); v1 = __ldg(xp + 1);
            }
            xv[0] = v0.x; xv[1] = v0.y; xv[2] = v0.z; xv[3] = v0.w; xv[4] = v1.x; xv[5] = v1.y; xv[6] = v1.z; xv[7] = v1.w;
        }
        {
            uint4* xo = reinterpret_cast<uint4*>(xr + r * 16);
            xo[0] = make_uint4(pack_bf16(xv[0], xv[1]), pack_bf16(xv[2], xv[3]), pack_bf16(xv[4], xv[5]), pack_bf16(xv[6], xv[7]));
            xo[1] = make_uint4(pack_bf16(valid ? 1.0f : 0.f, 0.f), 0, 0, 0);
        }
        float dxa[I];
#pragma unroll
        for (int i = 0; i < I; ++i) dxa[i] = 0.f;
        uint4* drow = reinterpret_cast<uint4*>(dza + r * AW);
#pragma unroll 1
        for (int c = 0; c < A_CHUNKS; ++c) {
            const uint4 dz = drow[c];
            const uint32_t dzw[4] = {dz.x, dz.y, dz.z, dz.w};
            float u[8];
            {
                const float4 b0 = *reinterpret_cast<const float4*>(bi + 8 * c), b1 = *reinterpret_cast<const float4*>(bi + 8 * c + 4);
                u[0] = b0.x; u[1] = b0.y; u[2] = b0.z; u[3] = b0.w; u[4] = b1.x; u[5] = b1.y; u[6] = b1.z; u[7] = b1.w;
            }
#pragma unroll
            for (int i = 0; i < I; ++i) {
                const float4 w0 = *reinterpret_cast<const float4*>(wt + i * AW + 8 * c), w1 = *reinterpret_cast<const float4*>(wt + i * AW + 8 * c + 4);
                u[0] = fmaf(xv[i], w0.x, u[0]); u[1] = fmaf(xv[i], w0.y, u[1]); u[2] = fmaf(xv[i], w0.z, u[2]); u[3] = fmaf(xv[i], w0.w, u[3]);
                u[4] = fmaf(xv[i], w1.x, u[4]); u[5] = fmaf(xv[i], w1.y, u[5]); u[6] = fmaf(xv[i], w1.z, u[6]); u[7] = fmaf(xv[i], w1.w, u[7]);
            }
            float du[8];
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
                float m0, m1, g0, g1;
                mish2_grad(u[e], u[e + 1], m0, m1, g0, g1);
                du[e] = __uint_as_float(dzw[e / 2] << 16) * g0;
                du[e + 1] = __uint_as_float(dzw[e / 2] & 0xffff0000u) * g1;
            }
            drow[c] = make_uint4(pack_bf16(du[0], du[1]), pack_bf16(du[2], du[3]), pack_bf16(du[4], du[5]), pack_bf16(du[6], du[7]));
            if (WITH_DX) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float4 w0 = *reinterpret_cast<const float4*>(wi + (8 * c + e) * I), w1 = *reinterpret_cast<const float4*>(wi + (8 * c + e) * I + 4);
                    dxa[0] = fmaf(du[e], w0.x, dxa[0]); dxa[1] = fmaf(du[e], w0.y, dxa[1]); dxa[2] = fmaf(du[e], w0.z, dxa[2]); dxa[3] = fmaf(du[e], w0.w, dxa[3]);
                    dxa[4] = fmaf(du[e], w1.x, dxa[4]); dxa[5] = fmaf(du[e], w1.y, dxa[5]); dxa[6] = fmaf(du[e], w1.z, dxa[6]); dxa[7] = fmaf(du[e], w1.w, dxa[7]);
                }
            }
        }
        if (WITH_DX && valid) {
            float4* dp = reinterpret_cast<float4*>(dx + (dx_tm ? (int64_t)t * B + b : b * T + t) * I);
            dp[0] = make_float4(dxa[0], dxa[1], dxa[2], dxa[3]);
            dp[1] = make_float4(dxa[4], dxa[5], dxa[6], dxa[7]);
        }
    }
}

// The same without dx (the usual case: the network's input needs no gradient): du = dz_a * mish'(u) with mish'(u) saved by the forward -
// one streaming pass, one thread per 8-feature chunk, every access 16 B and coalesced (the per-row kernel above re-reads its rows'
// 128-byte lines eight times through a thrashing L1: 1.46 ms against 0.6 ms of HBM time).  The first chunk's thread also writes xr.
__global__ void __launch_bounds__(256) liquid_big_du_kernel(const unsigned char* __restrict__ saved, const float* __restrict__ x,
                                                            __nv_bfloat16* __restrict__ dza, __nv_bfloat16* __restrict__ xr, int64_t B, int T,
                                                            int flags) {
    const bool x_tm = flags & VB_FLAG_X_TIME_MAJOR;
    const int64_t n_tiles = (B + TM - 1) / TM, R = saved_rows(B, T);
    const uint4* gU = reinterpret_cast<const uint4*>(saved + saved_u_off(R));
    uint4* d4 = reinterpret_cast<uint4*>(dza);
    const int64_t total = R * A_CHUNKS;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
        const uint4 dz = d4[e], g = __ldg(gU + e);
        const uint32_t dw[4] = {dz.x, dz.y, dz.z, dz.w}, gw[4] = {g.x, g.y, g.z, g.w};
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            o[k] = pack_bf16(__uint_as_float(dw[k] << 16) * __uint_as_float(gw[k] << 16),
                             __uint_as_float(dw[k] & 0xffff0000u) * __uint_as_float(gw[k] & 0xffff0000u));
        d4[e] = make_uint4(o[0], o[1], o[2], o[3]);
        const int64_t r = e / A_CHUNKS;
        if (e - r * A_CHUNKS == 0) {
            const int row = (int)(r % TM);
            const int64_t ts = r / TM, tile = ts % n_tiles;
            const int t = (int)(ts / n_tiles);
            const int64_t b = tile * TM + row;
            float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
            if (b < B) {
                const float4* xp = reinterpret_cast<const float4*>(x + (x_tm ? (int64_t)t * B + b : b * T + t) * I);
                v0 = __ldg(xp); v1 = __ldg(xp + 1);
            }
            uint4* xo = reinterpret_cast<uint4*>(xr + r * 16);
            xo[0] = make_uint4(pack_bf16(v0.x, v0.y), pack_bf16(v0.z, v0.w), pack_bf16(v1.x, v1.y), pack_bf16(v1.z, v1.w));
            xo[1] = make_uint4(pack_bf16(b < B ? 1.0f : 0.f, 0.f), 0, 0, 0);
        }
    }
}

// ------------------------------------------------------------------------------------------------ 4. dW_out | db_out = [m | 1]^T dy
// M is [tile-step][quad (51)][row (128)][4 bf16]: one thread per (quad, row), CTAs stride over the tile-steps, per-CTA partial sums
// reduced over the rows in shared memory: out[cta][0..203] = dW_out, out[cta][204] = db_out.
constexpr int DWO_COLS = 208;
__global__ void __launch_bounds__(1024) liquid_big_dwout_kernel(const unsigned char* __restrict__ saved, const float* __restrict__ dy,
                                                                float* __restrict__ out, int64_t B, int T, int flags) {
    __shared__ float red[8][DWO_COLS + 1];
    const bool y_tm = flags & VB_FLAG_Y_TIME_MAJOR;
    const int64_t n_tiles = (B + TM - 1) / TM, R = saved_rows(B, T), n_ts = (int64_t)T * n_tiles;
    const uint2* sM = reinterpret_cast<const uint2*>(saved + saved_m_off(R));
    const int row = threadIdx.x & (TM - 1), qg = threadIdx.x >> 7;         // 8 quad groups: quads qg, qg + 8, ... (7 per thread, 51 in all)
    float acc[7][4], accb = 0.f;
#pragma unroll
    for (int k = 0; k < 7; ++k) acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f;
    for (int64_t ts = blockIdx.x; ts < n_ts; ts += gridDim.x) {
        const int64_t tile = ts % n_tiles;
        const int t = (int)(ts / n_tiles);
        const int64_t b = tile * TM + row;
        const float dyv = b < B ? __ldg(dy + (y_tm ? (int64_t)t * B + b : b * T + t)) : 0.f;
        if (qg == 0) accb += dyv;
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const int q = qg + 8 * k;
            if (q < NQUAD) {
                const uint2 m = __ldg(sM + ((size_t)ts * NQUAD + q) * TM + row);
                acc[k][0] = fmaf(dyv, __uint_as_float(m.x << 16), acc[k][0]);
                acc[k][1] = fmaf(dyv, __uint_as_float(m.x & 0xffff0000u), acc[k][1]);
                acc[k][2] = fmaf(dyv, __uint_as_float(m.y << 16), acc[k][2]);
                acc[k][3] = fmaf(dyv, __uint_as_float(m.y & 0xffff0000u), acc[k][3]);
            }
        }
    }
    // reduce over the 128 rows: warp shuffles, then the four warps of a quad group through shared memory
    for (int i = threadIdx.x; i < 8 * (DWO_COLS + 1); i += 1024) (&red[0][0])[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 7; ++k)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float v = acc[k][e];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            const int q = qg + 8 * k;
            if (lane == 0 && q < NQUAD) atomicAdd(&red[(threadIdx.x >> 5) & 3][4 * q + e], v);
        }
    {
        float v = accb;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && qg == 0) atomicAdd(&red[(threadIdx.x >> 5) & 3][NC], v);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < DWO_COLS; c += 1024) out[(size_t)blockIdx.x * DWO_COLS + c] = (red[0][c] + red[1][c]) + (red[2][c] + red[3][c]);
}

// ------------------------------------------------------------------------------------------------ partial sums -> packed gradient block
// dwt [S][816][528] (f = 4 j + hs, k = z~ feature), dwin [S][16][312] ([x | 1]^T du), dwo [n_dwo][208]
__global__ void __launch_bounds__(256) liquid_big_unpack_kernel(const float* __restrict__ dwt, int S, const float* __restrict__ dwin,
                                                                const float* __restrict__ dwo, int n_dwo, float* __restrict__ dpacked) {
    constexpr LiquidLayout L(I, NI, NC, 1);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < L.f_total; e += gridDim.x * blockDim.x) {
        float v = 0.f;
        if (e >= L.wh_t && e < L.wout) {          // head weights wh_t [K][4 NCp] and biases b_h [4 NCp]
            const int o = e - L.wh_t, col = o % L.H4, i = o / L.H4;           // i == K: the bias row
            const int hs = col / L.NCp, j = col % L.NCp;
            if (j < NC) {
                const int f = 4 * j + hs, k = i < NI ? i : (i < NI + NC ? 312 + (i - NI) : 312 + NC);
                for (int s = 0; s < S; ++s) v += dwt[((size_t)s * DSW + f) * ZW + k];
            }
        } else if (e < L.wh_t) {                  // win_t [I][NIp] and b_in [NIp]
            const int i = e / L.NIp, j = e % L.NIp;                            // i == I: the bias row
            if (j < NI)
                for (int s = 0; s < S; ++s) v += dwin[((size_t)s * 16 + i) * AW + j];
        } else {                                  // wout [1][NCp], b_out [Op]
            const int o = e - L.wout;
            const int c = o < L.NCp ? o : (o == L.NCp ? NC : -1);              // column NC of dwo = db_out
            if (c >= 0 && (c < NC || c == NC))
                for (int s = 0; s < n_dwo; ++s) v += dwo[(size_t)s * DWO_COLS + c];
        }
        dpacked[e] = v;
    }
}

// ------------------------------------------------------------------------------------------------ host side
static std::mutex g_blas_mu;
static cublasHandle_t g_blas[64] = {nullptr};

static int blas_handle(cublasHandle_t* out) {
    int dev = 0;
    VB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_blas_mu);
    if (!g_blas[dev & 63]) {
        cublasStatus_t st = cublasCreate(&g_blas[dev & 63]);
        VB_CHECK_ARG(st == CUBLAS_STATUS_SUCCESS, 999, "cublasCreate failed (status %d)", (int)st);
    }
    *out = g_blas[dev & 63];
    return 0;
}

static int split_count(int64_t n_ts) {           // split-K batches: a divisor of the tile-step count, at most 32
    for (int s = 32; s > 1; --s)
        if (n_ts % s == 0) return s;
    return 1;
}

struct BwdLayout {
    int64_t R, n_ts;
    int S, n_dwo;
    size_t whp, wa, ds, dza, xr, dwt, dwin, dwo, total;
    BwdLayout(int64_t B, int T) {
        const int64_t n_tiles = (B + TM - 1) / TM;
        R = saved_rows(B, T);
        n_ts = (int64_t)T * n_tiles;
        S = split_count(n_ts);
        n_dwo = (int)(n_ts < 148 ? n_ts : 148);
        auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
        size_t off = 0;
        whp = off; off = up(off + kWhpBytes);
        wa = off; off = up(off + kWaBytes);
        ds = off; off = up(off + (size_t)R * DSW * 2);
        dza = off; off = up(off + (size_t)R * AW * 2);
        xr = off; off = up(off + (size_t)R * 16 * 2);
        dwt = off; off = up(off + (size_t)S * DSW * ZW * 4);
        dwin = off; off = up(off + (size_t)S * 16 * AW * 4);
        dwo = off; off = up(off + (size_t)n_dwo * DWO_COLS * 4);
        total = off;
    }
};

int64_t bwd_workspace_bytes(int64_t B, int T) { return (int64_t)BwdLayout(B, T).total; }

int bwd(const float* packed, const float* x, const float* dy, const float* dh_last, const void* saved, int64_t saved_bytes_, float* dx, float* dh0,
        float* dpacked, int64_t B, int T, int flags, void* ws, int64_t ws_bytes, cudaStream_t stream) {
    const BwdLayout W(B, T);
    VB_CHECK_ARG(ws && ws_bytes >= (int64_t)W.total, VB_ERR_WORKSPACE, "vb_liquid_bf16_bwd: workspace too small (%lld bytes, need %lld)",
                 (long long)ws_bytes, (long long)W.total);
    VB_CHECK_ARG(saved && saved_bytes_ >= (int64_t)saved_total_bytes(W.R), VB_ERR_WORKSPACE, "vb_liquid_bf16_bwd: saved block missing or too small");
    VB_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 127) == 0 && (reinterpret_cast<uintptr_t>(saved) & 127) == 0 &&
                     (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (!dx || (reinterpret_cast<uintptr_t>(dx) & 15) == 0) &&
                     (!dh_last || (reinterpret_cast<uintptr_t>(dh_last) & 15) == 0) && (!dh0 || (reinterpret_cast<uintptr_t>(dh0) & 15) == 0),
                 VB_ERR_ALIGNMENT, "vb_liquid_bf16_bwd: workspace / saved block must be 128-byte, x / dx / dh tensors 16-byte aligned");
    unsigned char* w8 = reinterpret_cast<unsigned char*>(ws);
    const unsigned char* sv = reinterpret_cast<const unsigned char*>(saved);
    __nv_bfloat16* whp = reinterpret_cast<__nv_bfloat16*>(w8 + W.whp);
    __nv_bfloat16* wa = reinterpret_cast<__nv_bfloat16*>(w8 + W.wa);
    __nv_bfloat16* ds = reinterpret_cast<__nv_bfloat16*>(w8 + W.ds);
    __nv_bfloat16* dza = reinterpret_cast<__nv_bfloat16*>(w8 + W.dza);
    __nv_bfloat16* xr = reinterpret_cast<__nv_bfloat16*>(w8 + W.xr);
    float* dwt = reinterpret_cast<float*>(w8 + W.dwt);
    float* dwin = reinterpret_cast<float*>(w8 + W.dwin);
    float* dwo = reinterpret_cast<float*>(w8 + W.dwo);
    const __nv_bfloat16* zsave = reinterpret_cast<const __nv_bfloat16*>(sv + saved_z_off(W.R));
    const int sms = device_sms() > 0 ? device_sms() : 148;
    const int64_t n_tiles = (B + TM - 1) / TM;

    big_bwd_pack_kernel<<<296, 256, 0, stream>>>(packed, whp, wa);
    VB_LAUNCH_CHECK();
    // 1. recurrence
    {
        LaunchCfg cfg;
        int rc = cached_launch_cfg(liquid_big_rec_kernel, R_NTHREADS, R_SMEM_BYTES, &cfg);
        if (rc) return rc;
        const int grid = (int)(n_tiles < sms ? n_tiles : sms);
        liquid_big_rec_kernel<<<grid, R_NTHREADS, R_SMEM_BYTES, stream>>>(reinterpret_cast<const unsigned char*>(whp), packed, sv, dy, dh_last,
                                                                        reinterpret_cast<unsigned char*>(ds), dh0, B, T, flags);
        VB_LAUNCH_CHECK();
    }
    // 4. dW_out (independent of the recurrence)
    liquid_big_dwout_kernel<<<W.n_dwo, 1024, 0, stream>>>(sv, dy, dwo, B, T, flags);
    VB_LAUNCH_CHECK();
    cublasHandle_t h;
    int rc = blas_handle(&h);
    if (rc) return rc;
    {
        std::lock_guard<std::mutex> lock(g_blas_mu);        // the handle's stream is per-call state
        const float one = 1.0f, zero = 0.0f;
        cublasStatus_t st = cublasSetStream(h, stream);
        VB_CHECK_ARG(st == CUBLAS_STATUS_SUCCESS, 999, "cublasSetStream failed (status %d)", (int)st);
        // 2a. dz_a [R][312] = ds [R][816] wa [816][312]   (row-major; as column-major: C[312 x R] = A[312 x 816] B[816 x R])
        st = cublasGemmEx(h, CUBLAS_OP_N, CUBLAS_OP_N, AW, (int)W.R, DSW, &one, wa, CUDA_R_16BF, AW, ds, CUDA_R_16BF, DSW, &zero, dza, CUDA_R_16BF, AW,
                          CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT);
        VB_CHECK_ARG(st == CUBLAS_STATUS_SUCCESS, 999, "cublasGemmEx (dz_a) failed (status %d)", (int)st);
        // 2b. du (and dx), xr
        if (dx) {
            auto kern = liquid_big_inter_bwd_kernel<true>;
            LaunchCfg cfg;
            rc = cached_launch_cfg(kern, IB_THREADS, IB_SMEM, &cfg);
            if (rc) return rc;
            const int64_t blocks = (W.R + IB_THREADS - 1) / IB_THREADS;
            const int grid = (int)(blocks < (int64_t)sms * 8 ? blocks : (int64_t)sms * 8);
            kern<<<grid, IB_THREADS, IB_SMEM, stream>>>(packed, x, dza, xr, dx, B, T, flags);
        } else {
            liquid_big_du_kernel<<<sms * 16, 256, 0, stream>>>(sv, x, dza, xr, B, T, flags);
        }
        VB_LAUNCH_CHECK();
        const int64_t Rs = W.R / W.S;
        // 2c. dW_in | db_in: P_s [16][312] = xr_s^T du_s   (column-major C[312 x 16] = du_s[312 x Rs] xr_s[16 x Rs]^T)
        st = cublasGemmStridedBatchedEx(h, CUBLAS_OP_N, CUBLAS_OP_T, AW, 16, (int)Rs, &one, dza, CUDA_R_16BF, AW, Rs * AW, xr, CUDA_R_16BF, 16, Rs * 16,
                                        &zero, dwin, CUDA_R_32F, AW, (long long)16 * AW, W.S, CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT);
        VB_CHECK_ARG(st == CUBLAS_STATUS_SUCCESS, 999, "cublasGemmStridedBatchedEx (dW_in) failed (status %d)", (int)st);
        // 3. dW_heads | db_heads: D_s [816][528] = ds_s^T z_s   (column-major C[528 x 816] = z_s[528 x Rs] ds_s[816 x Rs]^T)
        st = cublasGemmStridedBatchedEx(h, CUBLAS_OP_N, CUBLAS_OP_T, ZW, DSW, (int)Rs, &one, zsave, CUDA_R_16BF, ZW, Rs * ZW, ds, CUDA_R_16BF, DSW,
                                        Rs * DSW, &zero, dwt, CUDA_R_32F, ZW, (long long)DSW * ZW, W.S, CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT);
        VB_CHECK_ARG(st == CUBLAS_STATUS_SUCCESS, 999, "cublasGemmStridedBatchedEx (dW_heads) failed (status %d)", (int)st);
    }
    liquid_big_unpack_kernel<<<148, 256, 0, stream>>>(dwt, W.S, dwin, dwo, W.n_dwo, dpacked);
    VB_LAUNCH_CHECK();
    return 0;
}

}  // namespace big
}  // namespace vb
