// LiquidNCPNetwork(8,512,1) - the 512-neuron class of BASELINE config 4 - forward on the tensor cores in bf16 (fp32 accumulate,
// fp32 hidden state), opt-in through vb_liquid_bf16_fwd (the fp32 contract of vb_liquid_fwd stays on the tile kernels).
//
// Per sample and step the heads are a 517 x 816 contraction: one CTA owns a 128-sample tile and keeps z~ = [mish(W_in x) | h | 1]
// as 66 bf16 operand planes in shared memory (132 KB); the 816 x 528 bf16 weight block (858 KB, pre-arranged in plane order by
// big_pack_kernel) is streamed from L2 by one bulk-copy producer lane through a 3-stage ring; one lane issues the MMAs (four
// passes of 52 neurons = 208 accumulator columns, double-buffered in TMEM); 12 element-wise warps (three threads per sample)
// run the gating epilogue of pass p while the MMAs of pass p+1 execute and, between the epilogues, compute the inter layer of
// step t+1 into registers, so that the rebuild of z~ at the step boundary is only stores plus the conversion of h.  h travels between
// steps through a two-slot scratch buffer in a tile-interleaved order ([tile][Nc/4][128 rows][4]) so that every warp access is
// 512 contiguous bytes.  Measured (profiles/microbench/big_fwd_proto_r01.log): 0.82 G sample-steps/s = 690 TFLOP/s; the
// MMA + streaming pipeline alone sustains 1.33 PFLOP/s, what is exposed is the rebuild of z~.
// With a saved block (training) the kernel also writes the gate factors, mish(c) and z~ for the backward (liquid_big_bwd.cu).
#include "big_common.cuh"

namespace vb {
namespace big {

// weight block in plane order: [pass][stage][chunk (6)][row (208)][8 bf16]; row r = 4 jj + hs of neuron j = pass * 52 + jj,
// hs in (g, h, f, proj); feature k: a at [0,308), zeros, h at [312,516), the bias (constant-1 feature) at 516, zeros to 528
__global__ void __launch_bounds__(256) big_pack_kernel(const float* __restrict__ packed, __nv_bfloat16* __restrict__ wpack) {
    constexpr LiquidLayout L(I, NI, NC, 1);
    const int total = NPASS * KC * PROWS * 8;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int el = e & 7, r = (e >> 3) % PROWS, c = ((e >> 3) / PROWS) % KC, pass = (e >> 3) / (PROWS * KC);
        const int k = 8 * c + el, j = pass * PN + r / 4, hs = r % 4;
        float v = 0.f;
        if (j < NC) {
            const int col = hs * L.NCp + j;
            if (k < NI) v = packed[L.wh_t + k * L.H4 + col];
            else if (k >= 312 && k < 312 + NC) v = packed[L.wh_t + (NI + k - 312) * L.H4 + col];
            else if (k == 312 + NC) v = packed[L.b_h + col];
        }
        const int s = c / STAGE_CHUNKS, cc = c % STAGE_CHUNKS;
        wpack[((((size_t)pass * NSTAGE_PER_PASS + s) * STAGE_CHUNKS + cc) * PROWS + r) * 8 + el] = __float2bfloat16(v);
    }
}

// TRAIN: also write the activations the backward needs (big_common.cuh: G, M, Z) into `saved`.
template <bool TRAIN>
__global__ void __launch_bounds__(NTHREADS, 1) liquid_big_fwd_kernel(const unsigned char* __restrict__ wpack, const float* __restrict__ packed,
                                                                     const float* __restrict__ x, const float* __restrict__ h0,
                                                                     float* __restrict__ y, float* __restrict__ h_last,
                                                                     float* __restrict__ h_out, float* __restrict__ hbuf,
                                                                     unsigned char* __restrict__ saved, int64_t B, int T, int flags) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* zP = smem + S_Z;
    float* wins = reinterpret_cast<float*>(smem + S_WIN);
    float* bins = reinterpret_cast<float*>(smem + S_BIN);
    float* wouts = reinterpret_cast<float*>(smem + S_WOUT);
    float* ys = reinterpret_cast<float*>(smem + S_YS);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S_BAR);       // [RING]
    uint64_t* empty = full + RING;                                    // [RING]
    uint64_t* acc_full = empty + RING;                                // [2]
    uint64_t* acc_empty = acc_full + 2;                               // [2]
    uint64_t* zready = acc_empty + 2;                                 // all of z~ (a-part and h-part) is in place
    uint64_t* aready = zready + 1;                                    // the a-part planes of the coming step are in place
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aready + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    constexpr LiquidLayout L(I, NI, NC, 1);
    for (int e = tid; e < I * 320; e += NTHREADS) wins[e] = (e % 320) < NI ? __ldg(packed + L.win_t + (e / 320) * L.NIp + e % 320) : 0.f;
    for (int e = tid; e < 320; e += NTHREADS) bins[e] = e < NI ? __ldg(packed + L.b_in + e) : 0.f;
    for (int e = tid; e < 208; e += NTHREADS) wouts[e] = e < NC ? __ldg(packed + L.wout + e) : 0.f;
    const float b_out = __ldg(packed + L.b_out);
    const bool x_tm = flags & VB_FLAG_X_TIME_MAJOR, y_tm = flags & VB_FLAG_Y_TIME_MAJOR, h_tm = flags & VB_FLAG_H_TIME_MAJOR;
    for (int e = tid; e < TM; e += NTHREADS) reinterpret_cast<uint4*>(zP)[(KC - 1) * TM + e] = make_uint4(0, 0, 0, 0);   // chunk 65: zeros
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < RING; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], EW); }
        mbar_init(zready, 1);
        mbar_init(aready, EW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int64_t n_tiles = (B + TM - 1) / TM;
    const int64_t R = saved_rows(B, T);
    uint4* const sG = reinterpret_cast<uint4*>(saved + saved_g_off(R));
    uint2* const sM = reinterpret_cast<uint2*>(saved + saved_m_off(R));
    unsigned char* const sZ = saved + saved_z_off(R);
    unsigned char* const sU = saved + saved_u_off(R);
    const int64_t my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const int64_t n_steps = my_tiles * T;                   // steps this CTA walks (tiles are processed one after the other)

    if (warp == EW / 32) {
        // ---------------------------------------------------------------- producer: weight stages, the same 44 per step
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t st = 0; st < n_steps; ++st)
                for (int ps = 0; ps < NPASS * NSTAGE_PER_PASS; ++ps, ++it) {
                    const int slot = it % RING;
                    const uint32_t use = it / RING;
                    if (use > 0) mbar_wait(&empty[slot], (use - 1) & 1);
                    mbar_expect_tx(&full[slot], STAGE_BYTES);
                    bulk_g2s(smem + S_STAGE + slot * STAGE_BYTES, wpack + (size_t)ps * STAGE_BYTES, STAGE_BYTES, &full[slot]);
                }
        }
    } else if (warp == EW / 32 + 1) {
        // ---------------------------------------------------------------- MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(128, PROWS, 0, 0);
            const uint32_t z_addr = smem_u32(zP), st_addr = smem_u32(smem + S_STAGE);
            uint32_t it = 0, acc_use[2] = {0, 0};
            for (int64_t st = 0; st < n_steps; ++st) {
                // the a-part of z~ (chunks 0..38) is stored while the previous step's last epilogue still runs: the first six stages of
                // pass 0 only contract over those chunks and start before the hidden state has been converted
                mbar_wait(aready, st & 1);
                tc_fence_after();
                for (int pass = 0; pass < NPASS; ++pass) {
                    const int buf = pass & 1;
                    if (acc_use[buf] > 0) { mbar_wait(&acc_empty[buf], (acc_use[buf] - 1) & 1); tc_fence_after(); }
                    for (int s = 0; s < NSTAGE_PER_PASS; ++s, ++it) {
                        const int slot = it % RING;
                        if (pass == 0 && s == A_STAGES) mbar_wait(zready, st & 1);
                        mbar_wait(&full[slot], (it / RING) & 1);
                        tc_fence_after();
#pragma unroll
                        for (int kk = 0; kk < STAGE_CHUNKS / 2; ++kk)
                            mma_bf16(tmem + buf * PROWS, make_desc(z_addr + (STAGE_CHUNKS * s + 2 * kk) * PLANE, PLANE, 128),
                                     make_desc(st_addr + slot * STAGE_BYTES + 2 * kk * (PROWS * 16), PROWS * 16, 128), idesc, (s | kk) != 0);
                        mma_commit(&empty[slot]);
                    }
                    mma_commit(&acc_full[buf]);
                    acc_use[buf]++;
                }
            }
        }
    } else {
        // ---------------------------------------------------------------- element-wise warps: EWT threads per sample
        const int row = tid & (TM - 1), part = tid >> 7;          // part 0 .. EWT - 1
        const int a_off = part_off(A_CHUNKS, EWT, part), a_cnt = part_cnt(A_CHUNKS, EWT, part);
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t acc_use[2] = {0, 0};
        int64_t st = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t b = tile * TM + row;
            const bool valid = b < B;
            uint4 areg[AMAX];                     // this thread's chunks of a = mish(W_in x + b_in), bf16
            auto compute_a = [&](auto lo_c, auto hi_c, int t) {
                constexpr int LO = decltype(lo_c)::value, HI = decltype(hi_c)::value;
                float xr[I];
#pragma unroll
                for (int i = 0; i < I; ++i) xr[i] = valid ? __ldg(x + (x_tm ? (int64_t)t * B + b : b * T + t) * I + i) : 0.f;
#pragma unroll
                for (int k = LO; k < HI; ++k) {
                    if (k >= a_cnt) break;
                    const int c = a_off + k;
                    float u[8];
                    const float4 b0 = *reinterpret_cast<const float4*>(bins + 8 * c), b1 = *reinterpret_cast<const float4*>(bins + 8 * c + 4);
                    u[0] = b0.x; u[1] = b0.y; u[2] = b0.z; u[3] = b0.w; u[4] = b1.x; u[5] = b1.y; u[6] = b1.z; u[7] = b1.w;
#pragma unroll
                    for (int i = 0; i < I; ++i) {
                        const float4 w0 = *reinterpret_cast<const float4*>(wins + i * 320 + 8 * c);
                        const float4 w1 = *reinterpret_cast<const float4*>(wins + i * 320 + 8 * c + 4);
                        u[0] = fmaf(xr[i], w0.x, u[0]); u[1] = fmaf(xr[i], w0.y, u[1]); u[2] = fmaf(xr[i], w0.z, u[2]); u[3] = fmaf(xr[i], w0.w, u[3]);
                        u[4] = fmaf(xr[i], w1.x, u[4]); u[5] = fmaf(xr[i], w1.y, u[5]); u[6] = fmaf(xr[i], w1.z, u[6]); u[7] = fmaf(xr[i], w1.w, u[7]);
                    }
                    float a[8];
                    if (TRAIN) {
                        float g[8];
#pragma unroll
                        for (int e = 0; e < 8; e += 2) mish2_grad(u[e], u[e + 1], a[e], a[e + 1], g[e], g[e + 1]);
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            if (8 * c + e >= NI) g[e] = 0.f;
                        reinterpret_cast<uint4*>(sU + (size_t)(((int64_t)t * n_tiles + tile) * TM + row) * U_ROW_BYTES)[c] =
                            make_uint4(pack_bf16(g[0], g[1]), pack_bf16(g[2], g[3]), pack_bf16(g[4], g[5]), pack_bf16(g[6], g[7]));
                    } else {
                        float u4[4] = {u[0], u[1], u[2], u[3]}, a4[4];
                        mish4(u4, a4);
                        a[0] = a4[0]; a[1] = a4[1]; a[2] = a4[2]; a[3] = a4[3];
                        float v4[4] = {u[4], u[5], u[6], u[7]};
                        mish4(v4, a4);
                        a[4] = a4[0]; a[5] = a4[1]; a[6] = a4[2]; a[7] = a4[3];
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        if (8 * c + e >= NI) a[e] = 0.f;
                    areg[k] = make_uint4(pack_bf16(a[0], a[1]), pack_bf16(a[2], a[3]), pack_bf16(a[4], a[5]), pack_bf16(a[6], a[7]));
                }
            };
            for (int t = 0; t < T; ++t, ++st) {
                // ---- z~ = [mish(W_in x + b_in) | 0 | h_{t-1} | 1 | 0] as bf16 planes; the a-part was computed into registers during the
                //      previous step's MMAs (or right here for the first step of a tile)
                if (t == 0) {
                    compute_a(IC<0>{}, IC<AMAX>{}, 0);
#pragma unroll
                    for (int k = 0; k < AMAX; ++k)
                        if (k < a_cnt) reinterpret_cast<uint4*>(zP)[(a_off + k) * TM + row] = areg[k];
                    if (TRAIN) {
                        uint4* zr = reinterpret_cast<uint4*>(sZ + (size_t)(tile * TM + row) * Z_ROW_BYTES);      // r = (0 * n_tiles + tile) * 128 + row
#pragma unroll
                        for (int k = 0; k < AMAX; ++k)
                            if (k < a_cnt) zr[a_off + k] = areg[k];
                    }
                    fence_async_smem();
                    mbar_arrive(aready);
                }
                {
                    const int hc0 = part_off(H_CHUNKS, EWT, part), hc1 = hc0 + part_cnt(H_CHUNKS, EWT, part);
                    for (int c = hc0; c < hc1; ++c) {
                        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
                        if (t > 0) {
                            if (8 * c < NC) v0 = __ldcg(reinterpret_cast<const float4*>(hbuf + h_index((t - 1) & 1, n_tiles, tile, row, 8 * c)));
                            if (8 * c + 4 < NC) v1 = __ldcg(reinterpret_cast<const float4*>(hbuf + h_index((t - 1) & 1, n_tiles, tile, row, 8 * c + 4)));
                        } else if (h0 && valid) {
                            if (8 * c < NC) v0 = __ldg(reinterpret_cast<const float4*>(h0 + b * NC + 8 * c));
                            if (8 * c + 4 < NC) v1 = __ldg(reinterpret_cast<const float4*>(h0 + b * NC + 8 * c + 4));
                        }
                        if (8 * c + 4 == NC) v1.x = 1.0f;        // the constant-1 (bias) feature follows the last h
                        const uint4 hz = make_uint4(pack_bf16(v0.x, v0.y), pack_bf16(v0.z, v0.w), pack_bf16(v1.x, v1.y), pack_bf16(v1.z, v1.w));
                        reinterpret_cast<uint4*>(zP)[(H_CHUNK0 + c) * TM + row] = hz;
                        if (TRAIN) reinterpret_cast<uint4*>(sZ + (size_t)(((int64_t)t * n_tiles + tile) * TM + row) * Z_ROW_BYTES)[H_CHUNK0 + c] = hz;
                    }
                    if (TRAIN && part == 0)     // chunk 65 (features 520..527) is the all-zero plane
                        reinterpret_cast<uint4*>(sZ + (size_t)(((int64_t)t * n_tiles + tile) * TM + row) * Z_ROW_BYTES)[KC - 1] = make_uint4(0, 0, 0, 0);
                }
                fence_async_smem();
                ew_sync();
                if (tid == 0) mbar_arrive(zready);
                constexpr int AQ = (AMAX + 3) / 4;      // the next step's inter layer in four pieces, in the shadow of the MMAs
                constexpr int A1 = AQ < AMAX ? AQ : AMAX, A2 = 2 * AQ < AMAX ? 2 * AQ : AMAX, A3 = 3 * AQ < AMAX ? 3 * AQ : AMAX;
                if (t + 1 < T) compute_a(IC<0>{}, IC<A1>{}, t + 1);
                // ---- gating epilogue per pass
                float yacc = 0.f;
                for (int pass = 0; pass < NPASS; ++pass) {
                    const int buf = pass & 1;
                    mbar_wait(&acc_full[buf], acc_use[buf] & 1);
                    acc_use[buf]++;
                    tc_fence_after();
                    if (pass == NPASS - 1 && t + 1 < T) {
                        // every MMA of this step has completed: the a-planes are free, the next step's a-part goes in now
#pragma unroll
                        for (int k = 0; k < AMAX; ++k)
                            if (k < a_cnt) reinterpret_cast<uint4*>(zP)[(a_off + k) * TM + row] = areg[k];
                        fence_async_smem();
                        mbar_arrive(aready);
                        if (TRAIN) {
                            uint4* zr = reinterpret_cast<uint4*>(sZ + (size_t)(((int64_t)(t + 1) * n_tiles + tile) * TM + row) * Z_ROW_BYTES);
#pragma unroll
                            for (int k = 0; k < AMAX; ++k)
                                if (k < a_cnt) zr[a_off + k] = areg[k];
                        }
                    }
                    // the EWT threads of a row split the pass's 13 neuron quads, four neurons per iteration
                    const int n0 = 4 * part_off(PN / 4, EWT, part), nq = part_cnt(PN / 4, EWT, part);
#pragma unroll 1
                    for (int q = 0; q < nq; ++q) {
                        float s16[16];
                        tmem_ld16(taddr + buf * PROWS + 4 * n0 + 16 * q, s16);
                        float hn[4], cc[4], m[4];
                        float fa[4], fb[4], fc[4], fd[4];       // TRAIN: the gate factors of big_common.cuh
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            float tg, th, gate;
                            tanh2_sigmoid(s16[4 * jj + 0], s16[4 * jj + 1], s16[4 * jj + 2], tg, th, gate);
                            const float d = th - tg;
                            hn[jj] = fmaf(gate, d, tg);
                            cc[jj] = s16[4 * jj + 3] + hn[jj];
                            if (TRAIN) {
                                const float omg = 1.0f - gate;
                                fa[jj] = omg * fmaf(-tg, tg, 1.0f);
                                fb[jj] = gate * fmaf(-th, th, 1.0f);
                                fc[jj] = d * gate * omg;
                            }
                        }
                        if (TRAIN) {
                            mish2_grad(cc[0], cc[1], m[0], m[1], fd[0], fd[1]);
                            mish2_grad(cc[2], cc[3], m[2], m[3], fd[2], fd[3]);
                        } else {
                            mish4(cc, m);
                        }
                        const int j = pass * PN + n0 + 4 * q;
                        if (TRAIN && j < NC) {
                            const size_t slot = ((size_t)((int64_t)t * n_tiles + tile) * NQUAD + j / 4) * TM + row;
                            sG[2 * slot] = make_uint4(pack_bf16(fa[0], fb[0]), pack_bf16(fc[0], fd[0]), pack_bf16(fa[1], fb[1]), pack_bf16(fc[1], fd[1]));
                            sG[2 * slot + 1] = make_uint4(pack_bf16(fa[2], fb[2]), pack_bf16(fc[2], fd[2]), pack_bf16(fa[3], fb[3]), pack_bf16(fc[3], fd[3]));
                            sM[slot] = make_uint2(pack_bf16(m[0], m[1]), pack_bf16(m[2], m[3]));
                        }
                        if (j < NC) {
                            const float4 w4 = *reinterpret_cast<const float4*>(wouts + j);
                            yacc = fmaf(m[0], w4.x, fmaf(m[1], w4.y, fmaf(m[2], w4.z, fmaf(m[3], w4.w, yacc))));
                            const float4 h4 = make_float4(hn[0], hn[1], hn[2], hn[3]);
                            *reinterpret_cast<float4*>(hbuf + h_index(t & 1, n_tiles, tile, row, j)) = h4;
                            if (valid) {
                                if (t == T - 1 && h_last) *reinterpret_cast<float4*>(h_last + b * NC + j) = h4;
                                if (h_out) *reinterpret_cast<float4*>(h_out + (h_tm ? (int64_t)t * B + b : b * T + t) * NC + j) = h4;
                            }
                        }
                    }
                    tc_fence_before();
                    mbar_arrive(&acc_empty[buf]);
                    if (t + 1 < T) {
                        if (pass == 0) compute_a(IC<A1>{}, IC<A2>{}, t + 1);
                        else if (pass == 1) compute_a(IC<A2>{}, IC<A3>{}, t + 1);
                        else if (pass == 2) compute_a(IC<A3>{}, IC<AMAX>{}, t + 1);
                    }
                }
                if (part > 0) ys[(part - 1) * TM + row] = yacc;
                ew_sync();
                if (part == 0 && valid) {
                    float ysum = yacc;
#pragma unroll
                    for (int k = 0; k < EWT - 1; ++k) ysum += ys[k * TM + row];
                    y[y_tm ? (int64_t)t * B + b : b * T + t] = ysum + b_out;
                }
                ew_sync();          // h_t is visible to every thread of the tile before the next step reads it; ys is free again
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}


constexpr size_t kWpackBytes = (size_t)NPASS * KC * PROWS * 16;

bool has_path(int I_, int Ni, int Nc, int O) { return I_ == I && Ni == NI && Nc == NC && O == 1; }

int64_t fwd_workspace_bytes(int64_t B) {
    const int64_t n_tiles = (B + TM - 1) / TM;
    return (int64_t)kWpackBytes + 2 * n_tiles * TM * NC * (int64_t)sizeof(float);
}

int64_t saved_bytes(int64_t B, int T) { return (int64_t)saved_total_bytes(saved_rows(B, T)); }

int fwd(const float* packed, const float* x, const float* h0, float* y, float* h_last, float* h_traj, void* saved, int64_t saved_bytes_, int64_t B,
        int T, int flags, void* ws, int64_t ws_bytes, cudaStream_t stream) {
    if (saved) {
        VB_CHECK_ARG(saved_bytes_ >= saved_bytes(B, T), VB_ERR_WORKSPACE, "vb_liquid_bf16_fwd: saved block too small (%lld bytes, need %lld)",
                     (long long)saved_bytes_, (long long)saved_bytes(B, T));
        VB_CHECK_ARG((reinterpret_cast<uintptr_t>(saved) & 127) == 0, VB_ERR_ALIGNMENT, "vb_liquid_bf16_fwd: the saved block must be 128-byte aligned");
    }
    VB_CHECK_ARG(ws && ws_bytes >= fwd_workspace_bytes(B), VB_ERR_WORKSPACE, "vb_liquid_bf16_fwd: workspace too small (%lld bytes, need %lld)",
                 (long long)ws_bytes, (long long)fwd_workspace_bytes(B));
    VB_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 127) == 0 && (!h0 || (reinterpret_cast<uintptr_t>(h0) & 15) == 0) &&
                     (!h_last || (reinterpret_cast<uintptr_t>(h_last) & 15) == 0) && (!h_traj || (reinterpret_cast<uintptr_t>(h_traj) & 15) == 0),
                 VB_ERR_ALIGNMENT, "vb_liquid_bf16_fwd: workspace must be 128-byte, h tensors 16-byte aligned");
    __nv_bfloat16* wpack = reinterpret_cast<__nv_bfloat16*>(ws);
    float* hbuf = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(ws) + kWpackBytes);
    big_pack_kernel<<<296, 256, 0, stream>>>(packed, wpack);
    VB_LAUNCH_CHECK();
    LaunchCfg cfg;
    auto kern = saved ? liquid_big_fwd_kernel<true> : liquid_big_fwd_kernel<false>;
    int rc = cached_launch_cfg(kern, NTHREADS, SMEM_BYTES, &cfg);
    if (rc) return rc;
    const int64_t n_tiles = (B + TM - 1) / TM;
    const int sms = device_sms() > 0 ? device_sms() : 148;
    const int grid = (int)(n_tiles < sms ? n_tiles : sms);
    kern<<<grid, NTHREADS, SMEM_BYTES, stream>>>(reinterpret_cast<const unsigned char*>(wpack), packed, x, h0, y, h_last, h_traj, hbuf,
                                                 reinterpret_cast<unsigned char*>(saved), B, T, flags);
    VB_LAUNCH_CHECK();
    return 0;
}

}  // namespace big
}  // namespace vb
