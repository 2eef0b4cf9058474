// Generic (any-size) kernels: LiquidNCPNetwork forward/backward over T steps and NCPNetwork forward/backward.
//
// One CTA owns BM batch rows and keeps every activation of the step in shared memory; the layers are the tile
// routines of tile.cuh.  The packed, pre-masked parameter block is staged into shared memory with one bulk (TMA)
// copy when it fits next to the activations (the 128-neuron critic: 36 KB), otherwise it is read through L1/L2.
// Weight gradients are accumulated per CTA in shared memory across all tiles and time steps and reduced in a fixed
// order (deterministic); networks whose gradient block does not fit fall back to fp32 atomics on the output.
//
// Replaces: ncp.py:129-178 (+ cell.py:147-169, sparse.py:81) for sizes outside liquid_small.cu, and
// ncp.py:299-328 (the SAC critics: continuous.py:219-232, discrete.py:165-184), forward and autograd backward.
#include "tile.cuh"

namespace vb {
namespace generic {

using namespace tile;

constexpr int kModeAMaxFloats = 16384;   // gradient block kept in smem up to 64 KB

struct Plan {
    int BM;            // 32 / 16 / 8, 0 = does not fit
    bool stage_w;      // packed block staged in smem
    bool smem_acc;     // gradient block accumulated in smem (deterministic); else atomics on dpacked
    int smem_bytes;
    int rows;
};

// big_B > 0: the caller also compiled the register-blocked 128 / 64-row tiles; they are used when the batch fills the GPU with
// them and the activations (plus, backward, the shared-memory gradient accumulator) still fit.
static Plan make_plan(int rows, int packed_floats, int grad_floats, bool backward, int64_t big_B = 0) {
    Plan p{0, false, false, 0, rows};
    const int budget = device_max_smem_optin() > 0 ? device_max_smem_optin() : 232448;
    const int fixed = 64;   // mbarrier + alignment slack
    const int sms = device_sms() > 0 ? device_sms() : 148;
    for (int bm : {128, 64, 32, 16, 8}) {
        int act = rows * (bm + 4) * 4;
        // small batches: prefer small tiles so that the few rows spread over many SMs (batch 64 on 32-row tiles is 2 CTAs, each
        // walking the whole network serially: 12 / 24 us per critic forward / backward)
        if (big_B > 0 && bm > 8 && (big_B + bm - 1) / bm < sms / 2) continue;
        if (bm > 32) {   // only while two CTAs (16 warps) still fit per SM: the phases of a tile are separated by CTA barriers
            if (big_B < (int64_t)bm * sms) continue;
            const int g = (backward && grad_floats <= kModeAMaxFloats) ? grad_floats * 4 : 0;
            if (act + fixed + g > budget / 2) continue;
        }
        if (act + fixed <= budget) { p.BM = bm; p.smem_bytes = act + fixed; break; }
    }
    if (!p.BM) return p;
    if (backward && grad_floats <= kModeAMaxFloats && p.smem_bytes + grad_floats * 4 <= budget) {
        p.smem_acc = true;
        p.smem_bytes += grad_floats * 4;
    }
    // stage the weights only while at least two CTAs still fit per SM (latency hiding matters more than LDS vs L1)
    if (p.smem_bytes + packed_floats * 4 <= budget / 2) {
        p.stage_w = true;
        p.smem_bytes += packed_floats * 4;
    }
    return p;
}

// ================================================================================================ liquid forward
// CELL: the standalone NCPLiquidCell (cell.py:147-169): no inter layer (x [B, Ni] IS the a-part of z), no Mish / motor layer
// (y = proj + h', Nc wide).  The layout is LiquidLayout(Ni, Ni, Nc, Nc); its inter / motor sections are unused.
template <int BM, bool CELL>
__global__ void __launch_bounds__(kTileThreads) liquid_fwd_generic_kernel(const float* __restrict__ packed, LiquidLayout L,
                                                                           const float* __restrict__ x, const float* __restrict__ h0,
                                                                           float* __restrict__ y, float* __restrict__ h_traj,
                                                                           int64_t B, int T, int stage_w) {
    constexpr int S = Geo<BM>::S;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);
    float* base = reinterpret_cast<float*>(smem_raw + 64);
    float* xs = base;
    float* z = xs + L.I * S;
    float* sb = z + L.K * S;
    float* mb = sb + L.H4 * S;
    float* yb = mb + L.NCp * S;
    const float* P = packed;
    if (stage_w) {
        float* wsm = yb + L.O * S;
        bulk_load_and_wait(wsm, packed, (uint32_t)L.total * 4u, mbar);
        P = wsm;
    }
    const int64_t n_tiles = (B + BM - 1) / BM;
    for (int64_t tl = blockIdx.x; tl < n_tiles; tl += gridDim.x) {
        const int64_t b0 = tl * BM;
        const int nvalid = (int)min((int64_t)BM, B - b0);
        load_rows<BM>(z + L.Ni * S, h0 ? h0 + b0 * L.Nc : nullptr, L.Nc, L.Nc, nvalid);
        for (int t = 0; t < T; ++t) {
            if (CELL) {
                load_rows<BM>(z, x + (b0 * T + t) * L.Ni, (int64_t)T * L.Ni, L.Ni, nvalid);
            } else {
                load_rows<BM>(xs, x + (b0 * T + t) * L.I, (int64_t)T * L.I, L.I, nvalid);
                __syncthreads();
                dense<BM>(xs, L.I, P + L.win_t, L.NIp, P + L.b_in, L.Ni, [&](int j, int s, float v) { z[j * S + s] = mish_f(v); });
            }
            __syncthreads();
            dense<BM>(z, L.K, P + L.wh_t, L.H4, P + L.b_h, L.H4, [&](int j, int s, float v) { sb[j * S + s] = v; });
            __syncthreads();
            for (int idx = threadIdx.x; idx < L.Nc * BM; idx += kTileThreads) {
                int j = idx / BM, s = idx - j * BM;
                float tg = tanh_f(sb[j * S + s]), th = tanh_f(sb[(L.NCp + j) * S + s]);
                float gate = sigmoid_f(sb[(2 * L.NCp + j) * S + s]);
                float hn = fmaf(gate, th - tg, tg);
                z[(L.Ni + j) * S + s] = hn;
                const float c = sb[(3 * L.NCp + j) * S + s] + hn;
                mb[j * S + s] = CELL ? c : mish_f(c);
            }
            __syncthreads();
            if (CELL) {
                store_rows<BM>(y + (b0 * T + t) * L.Nc, (int64_t)T * L.Nc, mb, L.Nc, nvalid);
            } else {
                dense<BM>(mb, L.Nc, P + L.wout_t, L.Op, P + L.b_out, L.O, [&](int o, int s, float v) { yb[o * S + s] = v; });
                __syncthreads();
                store_rows<BM>(y + (b0 * T + t) * L.O, (int64_t)T * L.O, yb, L.O, nvalid);
            }
            store_rows<BM>(h_traj + (b0 * T + t) * L.Nc, (int64_t)T * L.Nc, z + L.Ni * S, L.Nc, nvalid);
        }
        __syncthreads();
    }
}

// ================================================================================================ liquid backward
template <int BM, bool ATOMIC, bool CELL>
__global__ void __launch_bounds__(kTileThreads) liquid_bwd_generic_kernel(
    const float* __restrict__ packed, LiquidLayout L, const float* __restrict__ x, const float* __restrict__ h0,
    const float* __restrict__ h_traj, const float* __restrict__ dy, const float* __restrict__ dh_traj,
    const float* __restrict__ dh_last, float* __restrict__ dx, float* __restrict__ dh0,
    float* __restrict__ grad_out /* ATOMIC: dpacked (pre-zeroed); else partials [grid][f_total] */, int64_t B, int T, int stage_w) {
    constexpr int S = Geo<BM>::S;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);
    float* base = reinterpret_cast<float*>(smem_raw + 64);
    float* xs = base;
    float* ub = xs + L.I * S;
    float* z = ub + L.NIp * S;
    float* sb = z + L.K * S;
    float* mb = sb + L.H4 * S;
    float* dmcb = mb + L.NCp * S;
    float* dyb = dmcb + L.NCp * S;
    float* dub = dyb + L.O * S;
    float* dhc = dub + L.NIp * S;
    float* dheb = dhc + L.NCp * S;
    float* next = dheb + L.NCp * S;
    float* acc = ATOMIC ? grad_out : next;
    if (!ATOMIC) {
        for (int i = threadIdx.x; i < L.f_total; i += kTileThreads) acc[i] = 0.f;
        next += L.f_total;
    }
    const float* P = packed;
    if (stage_w) {
        bulk_load_and_wait(next, packed, (uint32_t)L.total * 4u, mbar);
        P = next;
    }
    // zero the padded rows of du once (rows Ni..NIp-1 are read by wgrad's 4-row blocks)
    for (int i = threadIdx.x; i < L.NIp * S; i += kTileThreads) dub[i] = 0.f;
    for (int i = threadIdx.x; i < L.NCp * S; i += kTileThreads) mb[i] = 0.f;
    __syncthreads();
    const int64_t n_tiles = (B + BM - 1) / BM;
    for (int64_t tl = blockIdx.x; tl < n_tiles; tl += gridDim.x) {
        const int64_t b0 = tl * BM;
        const int nvalid = (int)min((int64_t)BM, B - b0);
        load_rows<BM>(dhc, dh_last ? dh_last + b0 * L.Nc : nullptr, L.Nc, L.Nc, nvalid);
        for (int t = T - 1; t >= 0; --t) {
            if (CELL) load_rows<BM>(z, x + (b0 * T + t) * L.Ni, (int64_t)T * L.Ni, L.Ni, nvalid);
            else load_rows<BM>(xs, x + (b0 * T + t) * L.I, (int64_t)T * L.I, L.I, nvalid);
            const float* hp = (t == 0) ? (h0 ? h0 + b0 * L.Nc : nullptr) : h_traj + (b0 * T + t - 1) * L.Nc;
            load_rows<BM>(z + L.Ni * S, hp, (t == 0) ? (int64_t)L.Nc : (int64_t)T * L.Nc, L.Nc, nvalid);
            load_rows<BM>(dyb, dy + (b0 * T + t) * L.O, (int64_t)T * L.O, L.O, nvalid);
            load_rows<BM>(dheb, dh_traj ? dh_traj + (b0 * T + t) * L.Nc : nullptr, (int64_t)T * L.Nc, L.Nc, nvalid);
            __syncthreads();
            // recompute
            if (!CELL) {
                dense<BM>(xs, L.I, P + L.win_t, L.NIp, P + L.b_in, L.Ni, [&](int j, int s, float v) {
                    ub[j * S + s] = v;
                    z[j * S + s] = mish_f(v);
                });
                __syncthreads();
            }
            dense<BM>(z, L.K, P + L.wh_t, L.H4, P + L.b_h, L.H4, [&](int j, int s, float v) { sb[j * S + s] = v; });
            __syncthreads();
            // gating derivatives; sb is overwritten by ds
            for (int idx = threadIdx.x; idx < L.Nc * BM; idx += kTileThreads) {
                int j = idx / BM, s = idx - j * BM;
                float tg = tanh_f(sb[j * S + s]), th = tanh_f(sb[(L.NCp + j) * S + s]);
                float gate = sigmoid_f(sb[(2 * L.NCp + j) * S + s]);
                float hn = fmaf(gate, th - tg, tg);
                float dmc = 1.0f, m = 0.f, dm;
                if (CELL) {
                    dm = dyb[j * S + s];            // y = c: the upstream gradient is dL/dc itself
                } else {
                    m = mish_fwd_grad(sb[(3 * L.NCp + j) * S + s] + hn, dmc);
                    dm = 0.f;
                    for (int o = 0; o < L.O; ++o) dm = fmaf(dyb[o * S + s], P[L.wout + o * L.NCp + j], dm);
                }
                float dc = dm * dmc;
                float dhn = dc + dhc[j * S + s] + dheb[j * S + s];
                float omg = 1.0f - gate;
                sb[j * S + s] = dhn * omg * (1.0f - tg * tg);
                sb[(L.NCp + j) * S + s] = dhn * gate * (1.0f - th * th);
                sb[(2 * L.NCp + j) * S + s] = dhn * (th - tg) * gate * omg;
                sb[(3 * L.NCp + j) * S + s] = dc;
                mb[j * S + s] = m;
            }
            __syncthreads();
            // dz = ds W  ->  du (inter part), carry (hidden part)
            dense<BM>(sb, L.H4, P + L.wh, L.Kp, nullptr, L.K, [&](int i, int s, float v) {
                if (i < L.Ni) {
                    if (CELL) {
                        if (dx && s < nvalid) dx[(b0 * T + t + (int64_t)s * T) * L.Ni + i] = v;      // dL/dx = the a-part of dz
                    } else {
                        float d;
                        (void)mish_fwd_grad(ub[i * S + s], d);
                        dub[i * S + s] = v * d;
                    }
                } else {
                    dhc[(i - L.Ni) * S + s] = v;
                }
            });
            __syncthreads();
            // dx = du W_in (written straight to global) and the weight gradients of this step
            if (CELL) {
                wgrad<BM, ATOMIC>(acc + L.wh_t, L.H4, acc + L.b_h, z, L.K, sb, L.H4);
                __syncthreads();
                continue;
            }
            if (dx) {
                float* dxt = dx + (b0 * T + t) * L.I;
                const int64_t pitch = (int64_t)T * L.I;
                dense<BM>(dub, L.Ni, P + L.win, L.Ip, nullptr, L.I, [&](int i, int s, float v) {
                    if (s < nvalid) dxt[s * pitch + i] = v;
                });
            }
            wgrad<BM, ATOMIC>(acc + L.win_t, L.NIp, acc + L.b_in, xs, L.I, dub, L.Ni);
            wgrad<BM, ATOMIC>(acc + L.wh_t, L.H4, acc + L.b_h, z, L.K, sb, L.H4);
            wgrad<BM, ATOMIC>(acc + L.wout, L.NCp, nullptr, dyb, L.O, mb, L.Nc);
            for (int o = threadIdx.x; o < L.O; o += kTileThreads) {
                float a = 0.f;
                for (int s = 0; s < BM; ++s) a += dyb[o * S + s];
                acc_add<ATOMIC>(acc + L.b_out + o, a);
            }
            __syncthreads();
        }
        store_rows<BM>(dh0 ? dh0 + b0 * L.Nc : nullptr, L.Nc, dhc, L.Nc, nvalid);
        __syncthreads();
    }
    if (!ATOMIC) {
        __syncthreads();
        for (int i = threadIdx.x; i < L.f_total; i += kTileThreads) grad_out[(int64_t)blockIdx.x * L.f_total + i] = acc[i];
    }
}

// ================================================================================================ NCP forward
template <int BM>
__global__ void __launch_bounds__(kTileThreads) ncp_fwd_kernel(const float* __restrict__ packed, NcpLayout L,
                                                                const float* __restrict__ x, float* __restrict__ y, int64_t B, int stage_w) {
    constexpr int S = Geo<BM>::S;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);
    float* base = reinterpret_cast<float*>(smem_raw + 64);
    float* xs = base;
    float* a1 = xs + L.I * S;
    float* a2 = a1 + L.NIp * S;
    float* yb = a2 + L.NCp * S;
    const float* P = packed;
    if (stage_w) {
        float* wsm = yb + L.O * S;
        bulk_load_and_wait(wsm, packed, (uint32_t)L.total * 4u, mbar);
        P = wsm;
    }
    const int64_t n_tiles = (B + BM - 1) / BM;
    for (int64_t tl = blockIdx.x; tl < n_tiles; tl += gridDim.x) {
        const int64_t b0 = tl * BM;
        const int nvalid = (int)min((int64_t)BM, B - b0);
        load_rows<BM>(xs, x + b0 * L.I, L.I, L.I, nvalid);
        __syncthreads();
        dense<BM>(xs, L.I, P + L.w1_t, L.NIp, P + L.b1, L.Ni, [&](int j, int s, float v) { a1[j * S + s] = mish_f(v); });
        __syncthreads();
        dense<BM>(a1, L.Ni, P + L.w2_t, L.NCp, P + L.b2, L.Nc, [&](int j, int s, float v) { a2[j * S + s] = mish_f(v); });
        __syncthreads();
        dense<BM>(a2, L.Nc, P + L.w3_t, L.Op, P + L.b3, L.O, [&](int o, int s, float v) { yb[o * S + s] = v; });
        __syncthreads();
        store_rows<BM>(y + b0 * L.O, L.O, yb, L.O, nvalid);
        __syncthreads();
    }
}

// ================================================================================================ NCP backward
template <int BM, bool ATOMIC>
__global__ void __launch_bounds__(kTileThreads) ncp_bwd_kernel(const float* __restrict__ packed, NcpLayout L,
                                                                const float* __restrict__ x, const float* __restrict__ dy,
                                                                float* __restrict__ dx, float* __restrict__ grad_out, int64_t B, int stage_w) {
    constexpr int S = Geo<BM>::S;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);
    float* base = reinterpret_cast<float*>(smem_raw + 64);
    float* xs = base;
    float* u1 = xs + L.I * S;
    float* a1 = u1 + L.NIp * S;
    float* u2 = a1 + L.NIp * S;
    float* a2 = u2 + L.NCp * S;
    float* dyb = a2 + L.NCp * S;
    float* du2 = dyb + L.O * S;
    float* du1 = du2 + L.NCp * S;
    float* next = du1 + L.NIp * S;
    float* acc = ATOMIC ? grad_out : next;
    if (!ATOMIC) {
        for (int i = threadIdx.x; i < L.f_total; i += kTileThreads) acc[i] = 0.f;
        next += L.f_total;
    }
    const float* P = packed;
    if (stage_w) {
        bulk_load_and_wait(next, packed, (uint32_t)L.total * 4u, mbar);
        P = next;
    }
    for (int i = threadIdx.x; i < L.NCp * S; i += kTileThreads) { du2[i] = 0.f; a2[i] = 0.f; }
    for (int i = threadIdx.x; i < L.NIp * S; i += kTileThreads) { du1[i] = 0.f; a1[i] = 0.f; }
    __syncthreads();
    const int64_t n_tiles = (B + BM - 1) / BM;
    for (int64_t tl = blockIdx.x; tl < n_tiles; tl += gridDim.x) {
        const int64_t b0 = tl * BM;
        const int nvalid = (int)min((int64_t)BM, B - b0);
        load_rows<BM>(xs, x + b0 * L.I, L.I, L.I, nvalid);
        load_rows<BM>(dyb, dy + b0 * L.O, L.O, L.O, nvalid);
        __syncthreads();
        dense<BM>(xs, L.I, P + L.w1_t, L.NIp, P + L.b1, L.Ni, [&](int j, int s, float v) { u1[j * S + s] = v; a1[j * S + s] = mish_f(v); });
        __syncthreads();
        dense<BM>(a1, L.Ni, P + L.w2_t, L.NCp, P + L.b2, L.Nc, [&](int j, int s, float v) { u2[j * S + s] = v; a2[j * S + s] = mish_f(v); });
        __syncthreads();
        // du2 = (dy W3) * mish'(u2):  "Wt" = w3 [O][NCp]
        dense<BM>(dyb, L.O, P + L.w3, L.NCp, nullptr, L.Nc, [&](int j, int s, float v) {
            float d;
            (void)mish_fwd_grad(u2[j * S + s], d);
            du2[j * S + s] = v * d;
        });
        __syncthreads();
        // du1 = (du2 W2) * mish'(u1):  "Wt" = w2 [Nc][NIp]
        dense<BM>(du2, L.Nc, P + L.w2, L.NIp, nullptr, L.Ni, [&](int j, int s, float v) {
            float d;
            (void)mish_fwd_grad(u1[j * S + s], d);
            du1[j * S + s] = v * d;
        });
        __syncthreads();
        if (dx) {
            float* dxt = dx + b0 * L.I;
            dense<BM>(du1, L.Ni, P + L.w1, L.Ip, nullptr, L.I, [&](int i, int s, float v) {
                if (s < nvalid) dxt[(int64_t)s * L.I + i] = v;
            });
        }
        if (grad_out) {       // NULL: the caller wants dL/dx only (a critic evaluated inside the actor loss)
            wgrad<BM, ATOMIC>(acc + L.w1_t, L.NIp, acc + L.b1, xs, L.I, du1, L.Ni);
            wgrad<BM, ATOMIC>(acc + L.w2_t, L.NCp, acc + L.b2, a1, L.Ni, du2, L.Nc);
            wgrad<BM, ATOMIC>(acc + L.w3, L.NCp, nullptr, dyb, L.O, a2, L.Nc);
            for (int o = threadIdx.x; o < L.O; o += kTileThreads) {
                float a = 0.f;
                for (int s = 0; s < BM; ++s) a += dyb[o * S + s];
                acc_add<ATOMIC>(acc + L.b3 + o, a);
            }
        }
        __syncthreads();
    }
    if (!ATOMIC) {
        __syncthreads();
        for (int i = threadIdx.x; i < L.f_total; i += kTileThreads) grad_out[(int64_t)blockIdx.x * L.f_total + i] = acc[i];
    }
}

// fixed-order (deterministic) column sums of the per-CTA partial gradient rows; see liquid_small.cu
__global__ void __launch_bounds__(1024) reduce_rows_kernel(const float* __restrict__ partials, int n_rows, int n, float* __restrict__ out) {
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

// ------------------------------------------------------------------------------------------------ host launchers
static int liquid_fwd_rows(const LiquidLayout& L) { return L.I + L.K + L.H4 + L.NCp + L.O; }
static int liquid_bwd_rows(const LiquidLayout& L) { return L.I + 2 * L.NIp + L.K + L.H4 + 5 * L.NCp + L.O; }
static int ncp_fwd_rows(const NcpLayout& L) { return L.I + L.NIp + L.NCp + L.O; }
static int ncp_bwd_rows(const NcpLayout& L) { return L.I + 3 * L.NIp + 3 * L.NCp + L.O; }

template <class K>
static int grid_for(K kernel, int smem, int64_t n_tiles, int cap_per_sm, int* grid) {
    LaunchCfg cfg;
    int rc = cached_launch_cfg(kernel, kTileThreads, smem, &cfg);
    if (rc) return rc;
    int occ = cfg.blocks_per_sm;
    if (cap_per_sm > 0 && occ > cap_per_sm) occ = cap_per_sm;
    int64_t cap = (int64_t)device_sms() * occ;
    int64_t g = n_tiles < cap ? n_tiles : cap;
    *grid = (int)(g < 1 ? 1 : g);
    return 0;
}

template <bool CELL>
static int liquid_fwd_impl(const float* packed, const float* x, const float* h0, float* y, float* h_traj, int64_t B, int T,
                           int I, int Ni, int Nc, int O, cudaStream_t stream) {
    LiquidLayout L(I, Ni, Nc, O);
    Plan p = make_plan(liquid_fwd_rows(L), L.total, 0, false);
    VB_CHECK_ARG(p.BM, VB_ERR_TOO_LARGE, "liquid forward: activations of (%d,%d,%d,%d) do not fit shared memory", I, Ni, Nc, O);
    int grid = 1, rc = 0;
#define LAUNCH(BM_)                                                                                                    \
    rc = grid_for(liquid_fwd_generic_kernel<BM_, CELL>, p.smem_bytes, (B + BM_ - 1) / BM_, 0, &grid);                  \
    if (rc) return rc;                                                                                                 \
    liquid_fwd_generic_kernel<BM_, CELL><<<grid, kTileThreads, p.smem_bytes, stream>>>(packed, L, x, h0, y, h_traj, B, T, p.stage_w);
    if (p.BM == 32) { LAUNCH(32) } else if (p.BM == 16) { LAUNCH(16) } else { LAUNCH(8) }
#undef LAUNCH
    VB_LAUNCH_CHECK();
    return 0;
}
int liquid_fwd(const float* packed, const float* x, const float* h0, float* y, float* h_traj, int64_t B, int T,
               int I, int Ni, int Nc, int O, cudaStream_t stream) {
    return liquid_fwd_impl<false>(packed, x, h0, y, h_traj, B, T, I, Ni, Nc, O, stream);
}
// standalone cell: x [B, In], h [B, Nc] -> y [B, Nc], h' [B, Nc]
int cell_fwd(const float* packed, const float* x, const float* h, float* y, float* h_new, int64_t B, int In, int Nc, cudaStream_t stream) {
    return liquid_fwd_impl<true>(packed, x, h, y, h_new, B, 1, In, In, Nc, Nc, stream);
}

int64_t liquid_bwd_workspace_bytes(int I, int Ni, int Nc, int O) {
    LiquidLayout L(I, Ni, Nc, O);
    Plan p = make_plan(liquid_bwd_rows(L), L.total, L.f_total, true);
    if (!p.BM || !p.smem_acc) return 16;
    int sms = device_sms() > 0 ? device_sms() : 148;
    return (int64_t)sms * 2 * L.f_total * 4;
}

template <bool CELL>
static int liquid_bwd_impl(const float* packed, const float* x, const float* h0, const float* h_traj, const float* dy,
                           const float* dh_traj, const float* dh_last, float* dx, float* dh0, float* dpacked, int64_t B, int T,
                           int I, int Ni, int Nc, int O, void* ws, int64_t ws_bytes, cudaStream_t stream) {
    LiquidLayout L(I, Ni, Nc, O);
    Plan p = make_plan(liquid_bwd_rows(L), L.total, L.f_total, true);
    VB_CHECK_ARG(p.BM, VB_ERR_TOO_LARGE, "liquid backward: activations of (%d,%d,%d,%d) do not fit shared memory", I, Ni, Nc, O);
    int grid = 1, rc = 0;
    if (p.smem_acc) {
        VB_CHECK_ARG(ws && ws_bytes >= liquid_bwd_workspace_bytes(I, Ni, Nc, O), VB_ERR_WORKSPACE, "vb_liquid_bwd: workspace too small");
#define LAUNCH(BM_)                                                                                                    \
    rc = grid_for(liquid_bwd_generic_kernel<BM_, false, CELL>, p.smem_bytes, (B + BM_ - 1) / BM_, 2, &grid);            \
    if (rc) return rc;                                                                                                 \
    liquid_bwd_generic_kernel<BM_, false, CELL><<<grid, kTileThreads, p.smem_bytes, stream>>>(                          \
        packed, L, x, h0, h_traj, dy, dh_traj, dh_last, dx, dh0, reinterpret_cast<float*>(ws), B, T, p.stage_w);
        if (p.BM == 32) { LAUNCH(32) } else if (p.BM == 16) { LAUNCH(16) } else { LAUNCH(8) }
#undef LAUNCH
        VB_LAUNCH_CHECK();
        reduce_rows_kernel<<<(L.f_total + 31) / 32, 1024, 0, stream>>>(reinterpret_cast<float*>(ws), grid, L.f_total, dpacked);
        VB_LAUNCH_CHECK();
    } else {
        VB_CUDA(cudaMemsetAsync(dpacked, 0, (size_t)L.f_total * 4, stream));
#define LAUNCH(BM_)                                                                                                    \
    rc = grid_for(liquid_bwd_generic_kernel<BM_, true, CELL>, p.smem_bytes, (B + BM_ - 1) / BM_, 0, &grid);             \
    if (rc) return rc;                                                                                                 \
    liquid_bwd_generic_kernel<BM_, true, CELL><<<grid, kTileThreads, p.smem_bytes, stream>>>(                           \
        packed, L, x, h0, h_traj, dy, dh_traj, dh_last, dx, dh0, dpacked, B, T, p.stage_w);
        if (p.BM == 32) { LAUNCH(32) } else if (p.BM == 16) { LAUNCH(16) } else { LAUNCH(8) }
#undef LAUNCH
        VB_LAUNCH_CHECK();
    }
    return 0;
}
int liquid_bwd(const float* packed, const float* x, const float* h0, const float* h_traj, const float* dy,
               const float* dh_traj, const float* dh_last, float* dx, float* dh0, float* dpacked, int64_t B, int T,
               int I, int Ni, int Nc, int O, void* ws, int64_t ws_bytes, cudaStream_t stream) {
    return liquid_bwd_impl<false>(packed, x, h0, h_traj, dy, dh_traj, dh_last, dx, dh0, dpacked, B, T, I, Ni, Nc, O, ws, ws_bytes, stream);
}
// standalone cell: dy [B, Nc] (gradient on y = proj + h'), dh_new [B, Nc] or NULL (gradient on h') -> dx [B, In], dh [B, Nc], dpacked
int cell_bwd(const float* packed, const float* x, const float* h, const float* dy, const float* dh_new, float* dx, float* dh,
             float* dpacked, int64_t B, int In, int Nc, void* ws, int64_t ws_bytes, cudaStream_t stream) {
    return liquid_bwd_impl<true>(packed, x, h, nullptr, dy, dh_new, nullptr, dx, dh, dpacked, B, 1, In, In, Nc, Nc, ws, ws_bytes, stream);
}

int ncp_fwd(const float* packed, const float* x, float* y, int64_t B, int I, int Ni, int Nc, int O, cudaStream_t stream) {
    NcpLayout L(I, Ni, Nc, O);
    Plan p = make_plan(ncp_fwd_rows(L), L.total, 0, false, B);
    VB_CHECK_ARG(p.BM, VB_ERR_TOO_LARGE, "ncp forward: activations of (%d,%d,%d,%d) do not fit shared memory", I, Ni, Nc, O);
    int grid = 1, rc = 0;
#define LAUNCH(BM_)                                                                                  \
    rc = grid_for(ncp_fwd_kernel<BM_>, p.smem_bytes, (B + BM_ - 1) / BM_, 0, &grid);                 \
    if (rc) return rc;                                                                               \
    ncp_fwd_kernel<BM_><<<grid, kTileThreads, p.smem_bytes, stream>>>(packed, L, x, y, B, p.stage_w);
    if (p.BM == 128) { LAUNCH(128) } else if (p.BM == 64) { LAUNCH(64) } else if (p.BM == 32) { LAUNCH(32) } else if (p.BM == 16) { LAUNCH(16) } else { LAUNCH(8) }
#undef LAUNCH
    VB_LAUNCH_CHECK();
    return 0;
}

int64_t ncp_bwd_workspace_bytes(int I, int Ni, int Nc, int O) {
    NcpLayout L(I, Ni, Nc, O);
    Plan p = make_plan(ncp_bwd_rows(L), L.total, L.f_total, true);
    if (!p.BM || !p.smem_acc) return 16;
    int sms = device_sms() > 0 ? device_sms() : 148;
    return (int64_t)sms * 2 * L.f_total * 4;
}

int ncp_bwd(const float* packed, const float* x, const float* dy, float* dx, float* dpacked, int64_t B,
            int I, int Ni, int Nc, int O, void* ws, int64_t ws_bytes, cudaStream_t stream) {
    NcpLayout L(I, Ni, Nc, O);
    Plan p = make_plan(ncp_bwd_rows(L), L.total, L.f_total, true, B);
    VB_CHECK_ARG(p.BM, VB_ERR_TOO_LARGE, "ncp backward: activations of (%d,%d,%d,%d) do not fit shared memory", I, Ni, Nc, O);
    int grid = 1, rc = 0;
    if (p.smem_acc && dpacked) {
        VB_CHECK_ARG(ws && ws_bytes >= ncp_bwd_workspace_bytes(I, Ni, Nc, O), VB_ERR_WORKSPACE, "vb_ncp_bwd: workspace too small");
#define LAUNCH(BM_)                                                                                           \
    rc = grid_for(ncp_bwd_kernel<BM_, false>, p.smem_bytes, (B + BM_ - 1) / BM_, 2, &grid);                    \
    if (rc) return rc;                                                                                        \
    ncp_bwd_kernel<BM_, false><<<grid, kTileThreads, p.smem_bytes, stream>>>(packed, L, x, dy, dx, reinterpret_cast<float*>(ws), B, p.stage_w);
        if (p.BM == 128) { LAUNCH(128) } else if (p.BM == 64) { LAUNCH(64) } else if (p.BM == 32) { LAUNCH(32) } else if (p.BM == 16) { LAUNCH(16) } else { LAUNCH(8) }
#undef LAUNCH
        VB_LAUNCH_CHECK();
        reduce_rows_kernel<<<(L.f_total + 31) / 32, 1024, 0, stream>>>(reinterpret_cast<float*>(ws), grid, L.f_total, dpacked);
        VB_LAUNCH_CHECK();
    } else {
        if (dpacked) VB_CUDA(cudaMemsetAsync(dpacked, 0, (size_t)L.f_total * 4, stream));
#define LAUNCH(BM_)                                                                                           \
    rc = grid_for(ncp_bwd_kernel<BM_, true>, p.smem_bytes, (B + BM_ - 1) / BM_, 0, &grid);                     \
    if (rc) return rc;                                                                                        \
    ncp_bwd_kernel<BM_, true><<<grid, kTileThreads, p.smem_bytes, stream>>>(packed, L, x, dy, dx, dpacked, B, p.stage_w);
        if (p.BM == 128) { LAUNCH(128) } else if (p.BM == 64) { LAUNCH(64) } else if (p.BM == 32) { LAUNCH(32) } else if (p.BM == 16) { LAUNCH(16) } else { LAUNCH(8) }
#undef LAUNCH
        VB_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace generic
}  // namespace vb
