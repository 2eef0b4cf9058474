// Specialised kernels for the 20-neuron class of LiquidNCPNetwork (Ni + Nc <= ~24, F section <= 4 KB).
//
// Design (evidence in profiles/microbench/):
//   * one thread owns one sample (batch row) and walks the whole time loop on chip, h in registers;
//   * the ~750 pre-masked weights live in the CONSTANT BANK, so every FFMA takes its weight as an instruction
//     operand (no load instruction, no register): measured 99-102 FFMA lanes/clk/SM against 59-75 when the same
//     weights are fed by broadcast LDS.128 from shared memory;
//   * x / y / dy / dx are streamed through per-warp shared-memory staging rows with 128-bit, full-sector global
//     accesses; h_traj rows (32 B per sample-step) are written/read as whole sectors directly;
//   * backward recomputes the gates from (x_t, h_{t-1}), then reduces the weight gradients by transposing
//     z / ds / du / x / m / dy through a per-warp shared-memory tile ([feature][32 samples]) so that each lane owns
//     a 4x5 register block of dW across ALL time steps and tiles; one deterministic cross-warp/CTA reduction at
//     the end (no atomics).
//
// Replaces, for T steps: ncp.py:168-171 -> sparse.py:81, cell.py:147-169, nn.Mish (the ~26 ATen launches per step
// counted in SURVEY.md section 2a) and autograd's ~60 backward launches.
#include <mutex>

#include "common.cuh"

namespace vb {
namespace small {

constexpr int kConstFloats = 2048;   // F section + transposed copies for the data-gradient (<= 8 KB)
__constant__ __align__(16) float c_p[kConstFloats];

template <int I_, int NI_, int NC_, int O_>
struct Cfg {
    static constexpr int I = I_, NI = NI_, NC = NC_, O = O_;
    static constexpr int NIp = ceil4(NI), NCp = ceil4(NC), Op = ceil4(O), K = NI + NC, H4 = 4 * NCp;
    static constexpr int WIN_T = 0;
    static constexpr int B_IN = WIN_T + I * NIp;
    static constexpr int WH_T = B_IN + NIp;
    static constexpr int B_H = WH_T + K * H4;
    static constexpr int WOUT = B_H + H4;
    static constexpr int B_OUT = WOUT + O * NCp;
    static constexpr int F_TOTAL = B_OUT + Op;
    // transposed copies (D section of the packed block, contiguous after F): used by the backward so that one
    // LDCU.128 feeds four INDEPENDENT accumulators (dz[i..i+3] / dx[i..i+3]) exactly like the forward does
    static constexpr int Ip = ceil4(I), Kp = ceil4(K);
    static constexpr int WIN = F_TOTAL;                 // [NI][Ip]
    static constexpr int WH = WIN + NI * Ip;            // [H4][Kp]
    static constexpr int C_TOTAL = WH + H4 * Kp;
    static_assert(C_TOTAL <= kConstFloats, "parameter block exceeds the constant-bank budget");
    static_assert(H4 <= 32, "weight-gradient block mapping needs 4*NCp <= 32");
    // weight-gradient register blocking: lane = (jb, ib), 4 ds rows x CB z columns
    static constexpr int CB = (K + 3) / 4;
    static constexpr int ZROWS = 4 * CB;               // z rows padded with zero rows
    static constexpr int R_IN = (I * NIp + 31) / 32;   // entries of dW_in per lane
    static constexpr int R_OUT = (O * NCp + 31) / 32;  // entries of dW_out per lane
    static_assert(NIp + O <= 32, "bias row mapping needs NIp + O <= 32");
};

// shared-memory row stride for a staged row of L floats: multiple of 4 with (stride/4) odd => LDS.128/STS.128 of
// 8 consecutive rows hit 8 distinct 16-byte bank groups
__host__ __device__ constexpr int stage_stride(int L) {
    int l4 = ceil4(L);
    return ((l4 / 4) & 1) ? l4 : l4 + 4;
}

constexpr int TS = 36;   // transposition tile row stride: 32 samples + 4 pad ((36/4) odd)

// ------------------------------------------------------------------------------------------------ staging copies
// rows = 32 samples of this warp; each row is `len` contiguous floats in global memory, `pitch` floats apart.
template <bool IN>
__device__ __forceinline__ void warp_stage(float* smem, int sstride, float* gptr, int64_t pitch, int len, int nvalid,
                                            bool vec_ok, int lane) {
    if (vec_ok && (len & 3) == 0) {
        const int per_row = len >> 2;
        for (int idx = lane; idx < 32 * per_row; idx += 32) {
            int r = idx / per_row, c = idx - r * per_row;
            float4* s4 = reinterpret_cast<float4*>(smem + r * sstride + 4 * c);
            if (IN) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < nvalid) v = __ldcs(reinterpret_cast<const float4*>(gptr + r * pitch + 4 * c));
                *s4 = v;
            } else if (r < nvalid) {
                __stcs(reinterpret_cast<float4*>(gptr + r * pitch + 4 * c), *s4);
            }
        }
    } else {
        for (int idx = lane; idx < 32 * len; idx += 32) {
            int r = idx / len, c = idx - r * len;
            if (IN) {
                smem[r * sstride + c] = (r < nvalid) ? __ldcs(gptr + r * pitch + c) : 0.f;
            } else if (r < nvalid) {
                __stcs(gptr + r * pitch + c, smem[r * sstride + c]);
            }
        }
    }
}

template <int N>
__device__ __forceinline__ void load_row(float (&dst)[N], const float* p, bool vec_ok) {
    if ((N & 3) == 0 && vec_ok) {
#pragma unroll
        for (int q = 0; q < N / 4; ++q) {
            float4 v = *reinterpret_cast<const float4*>(p + 4 * q);
            dst[4 * q] = v.x; dst[4 * q + 1] = v.y; dst[4 * q + 2] = v.z; dst[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int q = 0; q < N; ++q) dst[q] = p[q];
    }
}
template <int N>
__device__ __forceinline__ void store_row(float* p, const float (&src)[N], bool vec_ok) {
    if ((N & 3) == 0 && vec_ok) {
#pragma unroll
        for (int q = 0; q < N / 4; ++q)
            *reinterpret_cast<float4*>(p + 4 * q) = make_float4(src[4 * q], src[4 * q + 1], src[4 * q + 2], src[4 * q + 3]);
    } else {
#pragma unroll
        for (int q = 0; q < N; ++q) p[q] = src[q];
    }
}

// ------------------------------------------------------------------------------------------------ cell math
// u -> a = mish(u) (and mish'(u) when KEEP); the five heads with the two f heads pre-summed.
// opaque(): hides a compile-time index from the front end's value numbering (empty asm), ptxas still sees the constant
__device__ __forceinline__ int opaque(int v) {
    asm volatile("" : "+r"(v));
    return v;
}
#define CQ(idx) c_p[opaque(idx)]
// `zo` is a runtime zero (kernel argument).  The backward reads every weight twice per step (recompute and
// data-gradient); with plain compile-time indices the compiler merges the two reads and keeps ~750 weights in
// vector registers behind per-thread LDC loads (measured: 1186 LDC per warp-step, FFMAs stalled on the short
// scoreboard).  Distinct opaque offsets per section keep each read a uniform LDCU.128 feeding FFMA's UR operand.
template <class C>
__device__ __forceinline__ void inter_layer(const float (&x)[C::I], float (&u)[C::NI], int zo = 0) {
    const float* __restrict__ cp = c_p; (void)zo;
#pragma unroll
    for (int j = 0; j < C::NI; ++j) u[j] = cp[C::B_IN + j];
#pragma unroll
    for (int i = 0; i < C::I; ++i)
#pragma unroll
        for (int j = 0; j < C::NI; ++j) u[j] = fmaf(x[i], cp[C::WIN_T + i * C::NIp + j], u[j]);
}

template <class C>
__device__ __forceinline__ void head_layer(const float (&z)[C::K], float (&s)[4][C::NC], int zo = 0) {
    const float* __restrict__ cp = c_p; (void)zo;
#pragma unroll
    for (int hs = 0; hs < 4; ++hs)
#pragma unroll
        for (int j = 0; j < C::NC; ++j) s[hs][j] = cp[C::B_H + hs * C::NCp + j];
#pragma unroll
    for (int i = 0; i < C::K; ++i)
#pragma unroll
        for (int hs = 0; hs < 4; ++hs)
#pragma unroll
            for (int j = 0; j < C::NC; ++j) s[hs][j] = fmaf(z[i], cp[C::WH_T + i * C::H4 + hs * C::NCp + j], s[hs][j]);
}

// ------------------------------------------------------------------------------------------------ forward
constexpr int kFwdThreads = 128;
constexpr int kFwdTT = 8;   // time steps staged per chunk

template <class C>
struct FwdSmem {
    static constexpr int SX = stage_stride(kFwdTT * C::I);
    static constexpr int SY = stage_stride(kFwdTT * C::O);
    static constexpr int PER_WARP = 32 * (SX + SY);
    static constexpr int BYTES = PER_WARP * (kFwdThreads / 32) * 4;
};

template <class C>
__global__ void __launch_bounds__(kFwdThreads) liquid_fwd_small_kernel(const float* __restrict__ x, const float* __restrict__ h0,
                                                                        float* __restrict__ y, float* __restrict__ h_traj,
                                                                        int64_t B, int T, int vec_flags) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* xs = smem + warp * FwdSmem<C>::PER_WARP;
    float* ys = xs + 32 * FwdSmem<C>::SX;
    const bool vx = vec_flags & 1, vy = vec_flags & 2, vh = vec_flags & 4;
    const int64_t n_groups = (B + 31) / 32;
    const int64_t warps_total = (int64_t)gridDim.x * (kFwdThreads / 32);
    for (int64_t g = (int64_t)blockIdx.x * (kFwdThreads / 32) + warp; g < n_groups; g += warps_total) {
        const int64_t b0 = g * 32;
        const int64_t b = b0 + lane;
        const bool valid = b < B;
        const int nvalid = (int)min((int64_t)32, B - b0);
        float h[C::NC];
        if (valid && h0) load_row<C::NC>(h, h0 + b * C::NC, vh);
        else {
#pragma unroll
            for (int j = 0; j < C::NC; ++j) h[j] = 0.f;
        }
        for (int t0 = 0; t0 < T; t0 += kFwdTT) {
            const int nt = min(kFwdTT, T - t0);
            warp_stage<true>(xs, FwdSmem<C>::SX, const_cast<float*>(x) + (b0 * T + t0) * C::I, (int64_t)T * C::I, nt * C::I, nvalid, vx, lane);
            __syncwarp();
            for (int tt = 0; tt < nt; ++tt) {
                float xr[C::I];
                load_row<C::I>(xr, xs + lane * FwdSmem<C>::SX + tt * C::I, true);
                float u[C::NI];
                inter_layer<C>(xr, u);
                float z[C::K];
#pragma unroll
                for (int j = 0; j < C::NI; ++j) z[j] = mish_f(u[j]);
#pragma unroll
                for (int j = 0; j < C::NC; ++j) z[C::NI + j] = h[j];
                float s[4][C::NC];
                head_layer<C>(z, s);
                float m[C::NC];
#pragma unroll
                for (int j = 0; j < C::NC; ++j) {
                    float tg = tanh_f(s[0][j]), th = tanh_f(s[1][j]), gate = sigmoid_f(s[2][j]);
                    float hn = fmaf(gate, th - tg, tg);          // tg (1-gate) + gate th   (cell.py:134-136)
                    h[j] = hn;
                    m[j] = mish_f(s[3][j] + hn);                 // cell.py:168, ncp.py:171
                }
                float yo[C::O];
#pragma unroll
                for (int o = 0; o < C::O; ++o) {
                    float acc = c_p[C::B_OUT + o];
#pragma unroll
                    for (int j = 0; j < C::NC; ++j) acc = fmaf(m[j], c_p[C::WOUT + o * C::NCp + j], acc);
                    yo[o] = acc;
                }
                store_row<C::O>(ys + lane * FwdSmem<C>::SY + tt * C::O, yo, (C::O & 3) == 0);
                if (valid) store_row<C::NC>(h_traj + (b * T + t0 + tt) * C::NC, h, vh);
            }
            __syncwarp();
            warp_stage<false>(ys, FwdSmem<C>::SY, y + (b0 * T + t0) * C::O, (int64_t)T * C::O, nt * C::O, nvalid, vy, lane);
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------ backward
constexpr int kBwdThreads = 128;
constexpr int kBwdTT = 4;

template <class C>
struct BwdSmem {
    static constexpr int SX = stage_stride(kBwdTT * C::I);
    static constexpr int SY = stage_stride(kBwdTT * C::O);
    // transposition tile rows
    static constexpr int Z0 = 0;
    static constexpr int DS0 = Z0 + C::ZROWS;
    static constexpr int DU0 = DS0 + 32;
    static constexpr int X0 = DU0 + C::NIp;
    static constexpr int M0 = X0 + C::I;
    static constexpr int DY0 = M0 + C::NCp;
    static constexpr int TROWS = DY0 + C::O;
    static constexpr int PER_WARP = 32 * (2 * SX + SY) + TROWS * TS;   // x, dx, dy staging + tile
    static constexpr int BYTES = (PER_WARP * (kBwdThreads / 32) + C::F_TOTAL) * 4;
};

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
    acc = fmaf(a.x, b.x, acc);
    acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc);
    return fmaf(a.w, b.w, acc);
}

template <class C>
__global__ void __launch_bounds__(kBwdThreads) liquid_bwd_small_kernel(
    const float* __restrict__ x, const float* __restrict__ h0, const float* __restrict__ h_traj,
    const float* __restrict__ dy, const float* __restrict__ dh_traj, const float* __restrict__ dh_last,
    float* __restrict__ dx, float* __restrict__ dh0, float* __restrict__ partials, int64_t B, int T, int vec_flags,
    int zo1, int zo2) {
    using S = BwdSmem<C>;
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* xs = smem + warp * S::PER_WARP;
    float* dxs = xs + 32 * S::SX;
    float* dys = dxs + 32 * S::SX;
    float* tile = dys + 32 * S::SY;
    float* cta_acc = smem + (kBwdThreads / 32) * S::PER_WARP;
    const bool vx = vec_flags & 1, vy = vec_flags & 2, vh = vec_flags & 4;

    // zero the padding rows of the tile once (z rows K..ZROWS-1, ds rows H4..31)
    for (int idx = lane; idx < S::TROWS * TS; idx += 32) tile[idx] = 0.f;
    __syncwarp();

    // weight-gradient accumulators, alive across every tile and time step this warp processes
    const int jb = lane >> 2, ib = lane & 3;
    float acc_h[4][C::CB];
    float acc_in[C::R_IN], acc_out[C::R_OUT], acc_bh = 0.f, acc_b2 = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < C::CB; ++c) acc_h[r][c] = 0.f;
#pragma unroll
    for (int r = 0; r < C::R_IN; ++r) acc_in[r] = 0.f;
#pragma unroll
    for (int r = 0; r < C::R_OUT; ++r) acc_out[r] = 0.f;

    const int64_t n_groups = (B + 31) / 32;
    const int64_t warps_total = (int64_t)gridDim.x * (kBwdThreads / 32);
    for (int64_t g = (int64_t)blockIdx.x * (kBwdThreads / 32) + warp; g < n_groups; g += warps_total) {
        const int64_t b0 = g * 32;
        const int64_t b = b0 + lane;
        const bool valid = b < B;
        const int nvalid = (int)min((int64_t)32, B - b0);
        float dhc[C::NC];   // gradient flowing into h_t from step t+1 (and dh_last at t = T-1)
        if (valid && dh_last) load_row<C::NC>(dhc, dh_last + b * C::NC, vh);
        else {
#pragma unroll
            for (int j = 0; j < C::NC; ++j) dhc[j] = 0.f;
        }
        const int n_chunks = (T + kBwdTT - 1) / kBwdTT;
        for (int ch = n_chunks - 1; ch >= 0; --ch) {
            const int t0 = ch * kBwdTT;
            const int nt = min(kBwdTT, T - t0);
            warp_stage<true>(xs, S::SX, const_cast<float*>(x) + (b0 * T + t0) * C::I, (int64_t)T * C::I, nt * C::I, nvalid, vx, lane);
            warp_stage<true>(dys, S::SY, const_cast<float*>(dy) + (b0 * T + t0) * C::O, (int64_t)T * C::O, nt * C::O, nvalid, vy, lane);
            __syncwarp();
            for (int tt = nt - 1; tt >= 0; --tt) {
                const int t = t0 + tt;
                // ---------------- recompute the forward of step t from (x_t, h_{t-1})
                float xr[C::I];
                load_row<C::I>(xr, xs + lane * S::SX + tt * C::I, true);
                float z[C::K];
                float dmu[C::NI];
                {
                    float u[C::NI];
                    inter_layer<C>(xr, u, zo1);
#pragma unroll
                    for (int j = 0; j < C::NI; ++j) z[j] = mish_fwd_grad(u[j], dmu[j]);
                }
                {
                    float hp[C::NC];
                    const float* src = (t == 0) ? h0 : h_traj;
                    if (valid && src) load_row<C::NC>(hp, (t == 0) ? (h0 + b * C::NC) : (h_traj + (b * T + t - 1) * C::NC), vh);
                    else {
#pragma unroll
                        for (int j = 0; j < C::NC; ++j) hp[j] = 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < C::NC; ++j) z[C::NI + j] = hp[j];
                }
                float s[4][C::NC];
                head_layer<C>(z, s, zo1);
                // ---------------- output side: dy -> dm -> dc ; gating derivatives (s is overwritten by ds)
                float dyr[C::O];
                load_row<C::O>(dyr, dys + lane * S::SY + tt * C::O, (C::O & 3) == 0);
                float dhe[C::NC];
                if (valid && dh_traj) load_row<C::NC>(dhe, dh_traj + (b * T + t) * C::NC, vh);
                else {
#pragma unroll
                    for (int j = 0; j < C::NC; ++j) dhe[j] = 0.f;
                }
                float m[C::NC];
#pragma unroll
                for (int j = 0; j < C::NC; ++j) {
                    float tg = tanh_f(s[0][j]), th = tanh_f(s[1][j]), gate = sigmoid_f(s[2][j]);
                    float hn = fmaf(gate, th - tg, tg);
                    float dmc;
                    m[j] = mish_fwd_grad(s[3][j] + hn, dmc);
                    float dm = 0.f;
#pragma unroll
                    for (int o = 0; o < C::O; ++o) dm = fmaf(dyr[o], c_p[C::WOUT + o * C::NCp + j], dm);
                    float dc = dm * dmc;
                    float dhn = dc + dhc[j] + dhe[j];
                    float omg = 1.0f - gate;
                    s[0][j] = dhn * omg * (1.0f - tg * tg);       // ds_g
                    s[1][j] = dhn * gate * (1.0f - th * th);      // ds_h
                    s[2][j] = dhn * (th - tg) * gate * omg;       // ds_f (both f heads)
                    s[3][j] = dc;                                 // ds_proj
                }
                // ---------------- dz = sum_k ds_k W_k ; split into da (-> du) and the carry into h_{t-1}
                float du[C::NI];
                {
                    float dz[C::K];
#pragma unroll
                    for (int i = 0; i < C::K; ++i) dz[i] = 0.f;
#pragma unroll
                    for (int hs = 0; hs < 4; ++hs)
#pragma unroll
                        for (int j = 0; j < C::NC; ++j)
#pragma unroll
                            for (int i = 0; i < C::K; ++i)
                                dz[i] = fmaf(s[hs][j], c_p[C::WH + (hs * C::NCp + j) * C::Kp + i], dz[i]);
#pragma unroll
                    for (int i = 0; i < C::NI; ++i) du[i] = dz[i] * dmu[i];
#pragma unroll
                    for (int j = 0; j < C::NC; ++j) dhc[j] = dz[C::NI + j];
                }
                // ---------------- dx = du W_in
                {
                    float dxr[C::I];
#pragma unroll
                    for (int i = 0; i < C::I; ++i) dxr[i] = 0.f;
#pragma unroll
                    for (int j = 0; j < C::NI; ++j)
#pragma unroll
                        for (int i = 0; i < C::I; ++i) dxr[i] = fmaf(du[j], c_p[C::WIN + j * C::Ip + i], dxr[i]);
                    store_row<C::I>(dxs + lane * S::SX + tt * C::I, dxr, true);
                }
                // ---------------- weight gradients: transpose through the per-warp tile, then register blocks
#pragma unroll
                for (int i = 0; i < C::K; ++i) tile[(S::Z0 + i) * TS + lane] = z[i];
#pragma unroll
                for (int hs = 0; hs < 4; ++hs)
#pragma unroll
                    for (int j = 0; j < C::NC; ++j) tile[(S::DS0 + hs * C::NCp + j) * TS + lane] = s[hs][j];
#pragma unroll
                for (int j = 0; j < C::NI; ++j) tile[(S::DU0 + j) * TS + lane] = du[j];
#pragma unroll
                for (int i = 0; i < C::I; ++i) tile[(S::X0 + i) * TS + lane] = xr[i];
#pragma unroll
                for (int j = 0; j < C::NC; ++j) tile[(S::M0 + j) * TS + lane] = m[j];
#pragma unroll
                for (int o = 0; o < C::O; ++o) tile[(S::DY0 + o) * TS + lane] = dyr[o];
                __syncwarp();
                {
                    const float* ds_rows = tile + (S::DS0 + 4 * jb) * TS;
                    const float* z_rows = tile + (S::Z0 + C::CB * ib) * TS;
#pragma unroll 2
                    for (int q = 0; q < 8; ++q) {
                        float4 d[4];
#pragma unroll
                        for (int r = 0; r < 4; ++r) d[r] = *reinterpret_cast<const float4*>(ds_rows + r * TS + 4 * q);
#pragma unroll
                        for (int c = 0; c < C::CB; ++c) {
                            float4 zz = *reinterpret_cast<const float4*>(z_rows + c * TS + 4 * q);
#pragma unroll
                            for (int r = 0; r < 4; ++r) acc_h[r][c] = dot4(d[r], zz, acc_h[r][c]);
                        }
                    }
                    // inter layer: entry e -> (i = e / NIp, j = e % NIp)
#pragma unroll
                    for (int r = 0; r < C::R_IN; ++r) {
                        int e = min(lane + 32 * r, C::I * C::NIp - 1);
                        const float* xrow = tile + (S::X0 + e / C::NIp) * TS;
                        const float* drow = tile + (S::DU0 + e % C::NIp) * TS;
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            acc_in[r] = dot4(*reinterpret_cast<const float4*>(xrow + 4 * q), *reinterpret_cast<const float4*>(drow + 4 * q), acc_in[r]);
                    }
                    // motor layer: entry e -> (o = e / NCp, j = e % NCp)
#pragma unroll
                    for (int r = 0; r < C::R_OUT; ++r) {
                        int e = min(lane + 32 * r, C::O * C::NCp - 1);
                        const float* yrow = tile + (S::DY0 + e / C::NCp) * TS;
                        const float* mrow = tile + (S::M0 + e % C::NCp) * TS;
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            acc_out[r] = dot4(*reinterpret_cast<const float4*>(yrow + 4 * q), *reinterpret_cast<const float4*>(mrow + 4 * q), acc_out[r]);
                    }
                    // biases: row sums.  lane r < 32: ds row r; lane r < NIp: du row r; NIp <= lane < NIp+O: dy row
                    {
                        const float* row = tile + (S::DS0 + lane) * TS;
                        int r2 = min(lane, C::NIp + C::O - 1);
                        const float* row2 = (r2 < C::NIp) ? tile + (S::DU0 + r2) * TS : tile + (S::DY0 + r2 - C::NIp) * TS;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            float4 v = *reinterpret_cast<const float4*>(row + 4 * q);
                            acc_bh += (v.x + v.y) + (v.z + v.w);
                            float4 w = *reinterpret_cast<const float4*>(row2 + 4 * q);
                            acc_b2 += (w.x + w.y) + (w.z + w.w);
                        }
                    }
                }
                __syncwarp();
            }
            warp_stage<false>(dxs, S::SX, dx ? dx + (b0 * T + t0) * C::I : nullptr, (int64_t)T * C::I, nt * C::I, dx ? nvalid : 0, vx, lane);
            __syncwarp();
        }
        if (valid && dh0) store_row<C::NC>(dh0 + b * C::NC, dhc, vh);
    }

    // ---------------- deterministic reduction: warps add into the CTA accumulator one after another
    for (int idx = threadIdx.x; idx < C::F_TOTAL; idx += kBwdThreads) cta_acc[idx] = 0.f;
    __syncthreads();
    for (int w = 0; w < kBwdThreads / 32; ++w) {
        if (warp == w) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < C::CB; ++c) {
                    int row = 4 * jb + r, col = C::CB * ib + c;
                    if (row < C::H4 && col < C::K) cta_acc[C::WH_T + col * C::H4 + row] += acc_h[r][c];
                }
#pragma unroll
            for (int r = 0; r < C::R_IN; ++r) {
                int e = lane + 32 * r;
                if (e < C::I * C::NIp) cta_acc[C::WIN_T + e] += acc_in[r];
            }
#pragma unroll
            for (int r = 0; r < C::R_OUT; ++r) {
                int e = lane + 32 * r;
                if (e < C::O * C::NCp) cta_acc[C::WOUT + e] += acc_out[r];
            }
            if (lane < C::H4) cta_acc[C::B_H + lane] += acc_bh;
            if (lane < C::NIp) cta_acc[C::B_IN + lane] += acc_b2;
            else if (lane < C::NIp + C::O) cta_acc[C::B_OUT + lane - C::NIp] += acc_b2;
        }
        __syncthreads();
    }
    for (int idx = threadIdx.x; idx < C::F_TOTAL; idx += kBwdThreads) partials[(int64_t)blockIdx.x * C::F_TOTAL + idx] = cta_acc[idx];
}

// sums the per-CTA partial gradient rows in a fixed order (deterministic): a 32x32 thread block owns 32 columns;
// thread (cx, rg) adds rows rg, rg+32, ... of column cx, then the 32 row-group sums are added in index order.
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const float* __restrict__ partials, int n_rows, int n, float* __restrict__ out) {
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

// ------------------------------------------------------------------------------------------------ host side
// The constant bank is a per-device singleton: serialise its update against the previous launch that reads it.
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

static int vec_flags_for(const void* x, const void* y, const void* h_a, const void* h_b, const void* h_c, const void* h_d,
                         int T, int I, int O, int Nc, int tt) {
    auto al = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    int f = 0;
    if (al(x) && ((T * I) & 3) == 0 && ((tt * I) & 3) == 0) f |= 1;
    if (al(y) && ((T * O) & 3) == 0 && ((tt * O) & 3) == 0) f |= 2;
    if (al(h_a) && al(h_b) && al(h_c) && al(h_d) && (Nc & 3) == 0) f |= 4;
    return f;
}

template <class C>
static int launch_fwd(const float* packed, const float* x, const float* h0, float* y, float* h_traj, int64_t B, int T, cudaStream_t stream) {
    LaunchCfg cfg;
    int rc = cached_launch_cfg(liquid_fwd_small_kernel<C>, kFwdThreads, FwdSmem<C>::BYTES, &cfg);
    if (rc) return rc;
    const int occ = cfg.blocks_per_sm;
    std::lock_guard<std::mutex> lock(g_mu);
    int dev;
    bool capturing;
    rc = upload_constants(packed, C::F_TOTAL, stream, &dev, &capturing);
    if (rc) return rc;
    int64_t n_groups = (B + 31) / 32;
    int64_t want = (n_groups + kFwdThreads / 32 - 1) / (kFwdThreads / 32);
    int64_t cap = (int64_t)device_sms() * (occ > 0 ? occ : 1);
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    int vf = vec_flags_for(x, y, h0, h_traj, nullptr, nullptr, T, C::I, C::O, C::NC, kFwdTT);
    liquid_fwd_small_kernel<C><<<grid, kFwdThreads, FwdSmem<C>::BYTES, stream>>>(x, h0, y, h_traj, B, T, vf);
    VB_LAUNCH_CHECK();
    if (!capturing) VB_CUDA(cudaEventRecord(g_ev[dev & 63], stream));
    return 0;
}

template <class C>
static int bwd_grid(int64_t B, int* grid_out) {
    LaunchCfg cfg;
    int rc0 = cached_launch_cfg(liquid_bwd_small_kernel<C>, kBwdThreads, BwdSmem<C>::BYTES, &cfg);
    if (rc0) return rc0;
    const int occ = cfg.blocks_per_sm;
    int64_t n_groups = (B + 31) / 32;
    int64_t want = (n_groups + kBwdThreads / 32 - 1) / (kBwdThreads / 32);
    int64_t cap = (int64_t)device_sms() * (occ > 0 ? occ : 1);
    int grid = (int)(want < cap ? want : cap);
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
    std::lock_guard<std::mutex> lock(g_mu);
    int dev = 0;
    VB_CHECK_ARG(ws && ws_bytes >= (int64_t)grid * C::F_TOTAL * (int64_t)sizeof(float), VB_ERR_WORKSPACE,
                 "vb_liquid_bwd: workspace too small (%lld bytes, need %lld)", (long long)ws_bytes,
                 (long long)((int64_t)grid * C::F_TOTAL * sizeof(float)));
    bool capturing;
    rc = upload_constants(packed, C::C_TOTAL, stream, &dev, &capturing);
    if (rc) return rc;
    int vf = vec_flags_for(x, dy, h0, h_traj, dh_traj, dh_last, T, C::I, C::O, C::NC, kBwdTT);
    if (dx && (reinterpret_cast<uintptr_t>(dx) & 15)) vf &= ~1;
    if (dh0 && (reinterpret_cast<uintptr_t>(dh0) & 15)) vf &= ~4;
    float* partials = reinterpret_cast<float*>(ws);
    liquid_bwd_small_kernel<C><<<grid, kBwdThreads, BwdSmem<C>::BYTES, stream>>>(x, h0, h_traj, dy, dh_traj, dh_last, dx, dh0, partials, B, T, vf, 0, 0);
    VB_LAUNCH_CHECK();
    if (!capturing) VB_CUDA(cudaEventRecord(g_ev[dev & 63], stream));
    reduce_partials_kernel<<<(C::F_TOTAL + 31) / 32, 1024, 0, stream>>>(partials, grid, C::F_TOTAL, dpacked);
    VB_LAUNCH_CHECK();
    return 0;
}

// The shapes compiled in.  (4,12,8,2): NeuroFlow / NeuroFlowCT actor on 4-dim observations (BASELINE configs 1-3);
// (10,12,8,5): the reference's own test fixture LiquidNCPNetwork(10, 20, 5) (tests/models/test_lnn.py:280).
#define VB_SMALL_SHAPES(X) X(4, 12, 8, 2) X(10, 12, 8, 5) X(3, 12, 8, 2) X(8, 12, 8, 2)

bool has_fast_path(int I, int Ni, int Nc, int O) {
#define X(a, b, c, d) if (I == a && Ni == b && Nc == c && O == d) return true;
    VB_SMALL_SHAPES(X)
#undef X
    return false;
}

int fwd(const float* packed, const float* x, const float* h0, float* y, float* h_traj, int64_t B, int T,
        int I, int Ni, int Nc, int O, cudaStream_t stream) {
#define X(a, b, c, d) if (I == a && Ni == b && Nc == c && O == d) return launch_fwd<Cfg<a, b, c, d>>(packed, x, h0, y, h_traj, B, T, stream);
    VB_SMALL_SHAPES(X)
#undef X
    set_error("no specialised kernel for (%d,%d,%d,%d)", I, Ni, Nc, O);
    return VB_ERR_BAD_DIMS;
}

int bwd(const float* packed, const float* x, const float* h0, const float* h_traj, const float* dy, const float* dh_traj,
        const float* dh_last, float* dx, float* dh0, float* dpacked, int64_t B, int T, int I, int Ni, int Nc, int O,
        void* ws, int64_t ws_bytes, cudaStream_t stream) {
#define X(a, b, c, d)                                            \
    if (I == a && Ni == b && Nc == c && O == d)                  \
        return launch_bwd<Cfg<a, b, c, d>>(packed, x, h0, h_traj, dy, dh_traj, dh_last, dx, dh0, dpacked, B, T, ws, ws_bytes, stream);
    VB_SMALL_SHAPES(X)
#undef X
    set_error("no specialised kernel for (%d,%d,%d,%d)", I, Ni, Nc, O);
    return VB_ERR_BAD_DIMS;
}

int64_t bwd_workspace_bytes(int I, int Ni, int Nc, int O, int64_t B) {
    // upper bound that does not need a device query: SMs * 8 resident CTAs
    (void)B;
    LiquidLayout L(I, Ni, Nc, O);
    int sms = device_sms();
    if (sms <= 0) sms = 148;
    return (int64_t)sms * 8 * L.f_total * (int64_t)sizeof(float);
}

}  // namespace small
}  // namespace vb
