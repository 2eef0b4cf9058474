// Shared device/host helpers for libvelora_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <mutex>
#include <stdint.h>
#include <stdio.h>

#include "../../include/velora_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "velora_b200 kernels are written for sm_100a (B200) only"
#endif

namespace vb {

// ------------------------------------------------------------------------------------------------ errors
void set_error(const char* fmt, ...);

#define VB_CHECK_ARG(cond, code, ...)          \
    do {                                       \
        if (!(cond)) {                         \
            vb::set_error(__VA_ARGS__);        \
            return (code);                     \
        }                                      \
    } while (0)

#define VB_CUDA(expr)                                                                             \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            vb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return (int)_e;                                                                       \
        }                                                                                         \
    } while (0)

#define VB_LAUNCH_CHECK()                                                                         \
    do {                                                                                          \
        cudaError_t _e = cudaGetLastError();                                                      \
        if (_e != cudaSuccess) {                                                                  \
            vb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return (int)_e;                                                                       \
        }                                                                                         \
    } while (0)

int device_sms();
int device_max_smem_optin();

// Launch-configuration cache: raising the dynamic shared-memory limit and the occupancy query are driver calls of
// several microseconds each; at batch 64 (BASELINE configs 1/2) they cost more than the kernels.  One entry per
// (device, kernel, shared-memory size).
struct LaunchCfg { int blocks_per_sm; int regs; int static_smem; };
template <class K>
inline int cached_launch_cfg(K kernel, int threads, int smem, LaunchCfg* out) {
    struct Entry { int dev; const void* k; int smem; LaunchCfg cfg; };
    static Entry entries[64];
    static int n_entries = 0;
    static std::mutex mu;
    int dev = 0;
    VB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    const void* kp = reinterpret_cast<const void*>(kernel);
    int max_set = -1;   // the attribute is per-kernel state: only ever raise it
    for (int i = 0; i < n_entries; ++i) {
        if (entries[i].dev == dev && entries[i].k == kp) {
            if (entries[i].smem == smem) { *out = entries[i].cfg; return 0; }
            if (entries[i].smem > max_set) max_set = entries[i].smem;
        }
    }
    if (smem > max_set) VB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int occ = 1;
    VB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem));
    cudaFuncAttributes fa;
    VB_CUDA(cudaFuncGetAttributes(&fa, kernel));
    LaunchCfg cfg{occ < 1 ? 1 : occ, fa.numRegs, (int)fa.sharedSizeBytes};
    if (n_entries < 64) entries[n_entries++] = Entry{dev, kp, smem, cfg};
    *out = cfg;
    return 0;
}

__host__ __device__ constexpr int ceil4(int v) { return (v + 3) & ~3; }

// ------------------------------------------------------------------------------------------------ packed layouts
// Offsets (in floats) of the packed liquid parameter block; see include/velora_b200.h.
struct LiquidLayout {
    int I, Ni, Nc, O;
    int NIp, NCp, Op, Ip, K, Kp, H4;   // H4 = 4*NCp = width of one wh_t row
    // F section
    int win_t, b_in, wh_t, b_h, wout, b_out, f_total;
    // D section
    int win, wh, wout_t, total;
    __host__ __device__ constexpr LiquidLayout(int I_, int Ni_, int Nc_, int O_)
        : I(I_), Ni(Ni_), Nc(Nc_), O(O_),
          NIp(ceil4(Ni_)), NCp(ceil4(Nc_)), Op(ceil4(O_)), Ip(ceil4(I_)), K(Ni_ + Nc_), Kp(ceil4(Ni_ + Nc_)),
          H4(4 * ceil4(Nc_)),
          win_t(0),
          b_in(win_t + I_ * ceil4(Ni_)),
          wh_t(b_in + ceil4(Ni_)),
          b_h(wh_t + (Ni_ + Nc_) * 4 * ceil4(Nc_)),
          wout(b_h + 4 * ceil4(Nc_)),
          b_out(wout + O_ * ceil4(Nc_)),
          f_total(b_out + ceil4(O_)),
          win(f_total),
          wh(win + Ni_ * ceil4(I_)),
          wout_t(wh + 4 * ceil4(Nc_) * ceil4(Ni_ + Nc_)),
          total(wout_t + Nc_ * ceil4(O_)) {}
};

// Three-layer NCPNetwork block.
struct NcpLayout {
    int I, Ni, Nc, O;
    int NIp, NCp, Op, Ip;
    int w1_t, b1, w2_t, b2, w3, b3, f_total;
    int w1, w2, w3_t, total;
    __host__ __device__ constexpr NcpLayout(int I_, int Ni_, int Nc_, int O_)
        : I(I_), Ni(Ni_), Nc(Nc_), O(O_),
          NIp(ceil4(Ni_)), NCp(ceil4(Nc_)), Op(ceil4(O_)), Ip(ceil4(I_)),
          w1_t(0),
          b1(w1_t + I_ * ceil4(Ni_)),
          w2_t(b1 + ceil4(Ni_)),
          b2(w2_t + Ni_ * ceil4(Nc_)),
          w3(b2 + ceil4(Nc_)),
          b3(w3 + O_ * ceil4(Nc_)),
          f_total(b3 + ceil4(O_)),
          w1(f_total),
          w2(w1 + Ni_ * ceil4(I_)),
          w3_t(w2 + Nc_ * ceil4(Ni_)),
          total(w3_t + Nc_ * ceil4(O_)) {}
};

inline bool dims_ok(int I, int Ni, int Nc, int O) { return I > 0 && Ni > 0 && Nc > 0 && O > 0; }

// ------------------------------------------------------------------------------------------------ device math
#ifdef __CUDACC__
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// min(x, c) that keeps a NaN in x (fminf returns the non-NaN operand, which would turn tanh(NaN) into 1 and sigmoid(NaN) into 0)
__device__ __forceinline__ float min_keep_nan(float x, float c) {
    float y;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(y) : "f"(x), "f"(c));
    return y;
}
__device__ __forceinline__ float max_keep_nan(float x, float c) {
    float y;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(y) : "f"(x), "f"(c));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// tanh(x) = 1 - 2 / (e^{2x} + 1).  One ex2 + one rcp; |abs err| ~ 1e-7 (the measure that matters: outputs are
// compared in relative L2 norm).  tanh.approx (2^-11) is far too coarse for the 1e-5 contract.
__device__ __forceinline__ float tanh_f(float x) {
    float e = ex2_approx(min_keep_nan(x, 15.0f) * (2.0f * kLog2e));
    return 1.0f - 2.0f * rcp_approx(e + 1.0f);
}
__device__ __forceinline__ float sigmoid_f(float x) {
    float e = ex2_approx(min_keep_nan(-x, 80.0f) * kLog2e);
    return rcp_approx(1.0f + e);
}
// Mish(x) = x * tanh(softplus(x)) = x * w / (w + 2), w = e^x (e^x + 2).
__device__ __forceinline__ float mish_f(float x) {
    float n = ex2_approx(min_keep_nan(x, 20.0f) * kLog2e);
    float w = n * (n + 2.0f);
    return x * (w * rcp_approx(w + 2.0f));
}
// Mish and its derivative.  With n = e^x, w = n (n + 2), t = tanh(softplus x) = w / (w + 2):
//   dt/dx = (1 - t^2) sigmoid(x) = [4 (w + 1) / (w + 2)^2] [n / (n + 1)] = 4 n (n + 1) / (w + 2)^2     (w + 1 = (n + 1)^2)
// so mish'(x) = t + 4 x n (n + 1) r^2 with r = 1 / (w + 2): one ex2, one rcp, 10 FP32 ops.
__device__ __forceinline__ float mish_fwd_grad(float x, float& dmish) {
    float n = ex2_approx(min_keep_nan(x, 20.0f) * kLog2e);
    float w = n * (n + 2.0f);
    float r = rcp_approx(w + 2.0f);
    float t = w * r;
    float p = fmaf(n, n, n) * r;     // n (n + 1) / (w + 2)
    dmish = fmaf(4.0f * x, p * r, t);
    return x * t;
}

// ---- MUFU-lean variants for kernels bound by the XU pipe (16 lanes/clk/SM): several reciprocals share one MUFU.RCP
// (1/a = b * 1/(ab)).  Operands are clamped where the function is already saturated in fp32 so that the products stay
// finite: e^18 * e^18 * e^30 < 2^128.
// tanh(sg), tanh(sh), sigmoid(sf)
__device__ __forceinline__ void tanh2_sigmoid(float sg, float sh, float sf, float& tg, float& th, float& gate) {
    float a = ex2_approx(min_keep_nan(sg, 9.0f) * (2.0f * kLog2e)) + 1.0f;
    float b = ex2_approx(min_keep_nan(sh, 9.0f) * (2.0f * kLog2e)) + 1.0f;
    float c = ex2_approx(min_keep_nan(-sf, 30.0f) * kLog2e) + 1.0f;
    float ab = a * b;
    float r = rcp_approx(ab * c);
    float rab = c * r;
    gate = ab * r;
    tg = fmaf(-2.0f * b, rab, 1.0f);
    th = fmaf(-2.0f * a, rab, 1.0f);
}
// w = n (n + 2), n = e^min(x, 9): tanh(softplus x) = w / (w + 2) equals 1 to fp32 precision beyond x = 9
__device__ __forceinline__ float mish_w(float x) {
    float n = ex2_approx(min_keep_nan(x, 9.0f) * kLog2e);
    return n * (n + 2.0f);
}
__device__ __forceinline__ void mish2(float x0, float x1, float& y0, float& y1) {
    float w0 = mish_w(x0), w1 = mish_w(x1);
    float d0 = w0 + 2.0f, d1 = w1 + 2.0f;
    float r = rcp_approx(d0 * d1);
    y0 = x0 * (w0 * (d1 * r));
    y1 = x1 * (w1 * (d0 * r));
}
__device__ __forceinline__ void mish4(const float (&x)[4], float (&y)[4]) {
    float w[4], d[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { w[i] = mish_w(x[i]); d[i] = w[i] + 2.0f; }
    float p01 = d[0] * d[1], p23 = d[2] * d[3];
    float r = rcp_approx(p01 * p23);
    float r01 = p23 * r, r23 = p01 * r;
    y[0] = x[0] * (w[0] * (d[1] * r01));
    y[1] = x[1] * (w[1] * (d[0] * r01));
    y[2] = x[2] * (w[2] * (d[3] * r23));
    y[3] = x[3] * (w[3] * (d[2] * r23));
}

// Mish and Mish' of two values sharing one reciprocal (see mish_fwd_grad): n = e^x, w = n (n + 2), t = w / (w + 2),
// mish' = t + 4 x n (n + 1) / (w + 2)^2.  Clamped at 20: (w + 2)^2 stays finite for the pair product.
__device__ __forceinline__ void mish2_grad(float x0, float x1, float& y0, float& y1, float& g0, float& g1) {
    float n0 = ex2_approx(min_keep_nan(x0, 20.0f) * kLog2e), n1 = ex2_approx(min_keep_nan(x1, 20.0f) * kLog2e);
    float w0 = n0 * (n0 + 2.0f), w1 = n1 * (n1 + 2.0f);
    float d0 = w0 + 2.0f, d1 = w1 + 2.0f;
    float r = rcp_approx(d0 * d1);
    float r0 = d1 * r, r1 = d0 * r;
    float t0 = w0 * r0, t1 = w1 * r1;
    float p0 = fmaf(n0, n0, n0) * r0, p1 = fmaf(n1, n1, n1) * r1;
    g0 = fmaf(4.0f * x0, p0 * r0, t0);
    g1 = fmaf(4.0f * x1, p1 * r1, t1);
    y0 = x0 * t0;
    y1 = x1 * t1;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ---- packed fp32 pairs (sm_100: FFMA2 / FMUL2 / FADD2).  One packed instruction does two IEEE fp32 operations in ONE issue slot (the FMA
// pipe is busy two cycles, so the flop rate is unchanged: profiles/microbench/ffma2.cu measures 127 scalar FMA / clk / SM either way) - the
// lever for kernels bound by instruction issue rather than by a pipe.  A pair lives in an aligned 64-bit register; pk / lo / hi are
// register renames, not instructions.  MUFU and min / max stay scalar.
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float lo(f2 a) { float l; asm("{\n\t.reg .b32 t;\n\tmov.b64 {%0, t}, %1;\n\t}" : "=f"(l) : "l"(a)); return l; }
__device__ __forceinline__ float hi(f2 a) { float h; asm("{\n\t.reg .b32 t;\n\tmov.b64 {t, %0}, %1;\n\t}" : "=f"(h) : "l"(a)); return h; }
__device__ __forceinline__ f2 bc2(float c) { return pk(c, c); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
// 2^min(x, clamp) of both halves (the clamp keeps a NaN)
__device__ __forceinline__ f2 ex2_min2(f2 x, float clamp) { return pk(ex2_approx(min_keep_nan(lo(x), clamp)), ex2_approx(min_keep_nan(hi(x), clamp))); }
__device__ __forceinline__ f2 rcp2(f2 x) { return pk(rcp_approx(lo(x)), rcp_approx(hi(x))); }

// tanh(sg), tanh(sh), sigmoid(sf) of two neurons at once: tanh2_sigmoid() above in packed form (per neuron ONE reciprocal serves all three)
__device__ __forceinline__ void tanh2_sigmoid_p(f2 sg, f2 sh, f2 sf, f2& tg, f2& th, f2& gate) {
    const f2 one = bc2(1.0f);
    f2 a = add2(ex2_min2(mul2(sg, bc2(2.0f * kLog2e)), 9.0f * 2.0f * kLog2e), one);
    f2 b = add2(ex2_min2(mul2(sh, bc2(2.0f * kLog2e)), 9.0f * 2.0f * kLog2e), one);
    f2 c = add2(ex2_min2(mul2(sf, bc2(-kLog2e)), 30.0f * kLog2e), one);
    f2 ab = mul2(a, b);
    f2 r = rcp2(mul2(ab, c));
    f2 rab = mul2(c, r);
    gate = mul2(ab, r);
    tg = fma2(mul2(b, bc2(-2.0f)), rab, one);
    th = fma2(mul2(a, bc2(-2.0f)), rab, one);
}
// Mish of four values held as two pairs: 1 / d_i from the two reciprocals of the cross products d_A d_B (no lane swap needed)
__device__ __forceinline__ void mish4_p(f2 xa, f2 xb, f2& ya, f2& yb) {
    const f2 two = bc2(2.0f);
    f2 na = ex2_min2(mul2(xa, bc2(kLog2e)), 9.0f * kLog2e), nb = ex2_min2(mul2(xb, bc2(kLog2e)), 9.0f * kLog2e);
    f2 wa = mul2(na, add2(na, two)), wb = mul2(nb, add2(nb, two));
    f2 da = add2(wa, two), db = add2(wb, two);
    f2 rr = rcp2(mul2(da, db));
    ya = mul2(xa, mul2(wa, mul2(db, rr)));
    yb = mul2(xb, mul2(wb, mul2(da, rr)));
}
// Mish and Mish' of a pair (mish_fwd_grad above in packed form; one reciprocal per value)
__device__ __forceinline__ void mish_grad_p(f2 x, f2& y, f2& g) {
    f2 n = ex2_min2(mul2(x, bc2(kLog2e)), 20.0f * kLog2e);
    f2 w = mul2(n, add2(n, bc2(2.0f)));
    f2 r = rcp2(add2(w, bc2(2.0f)));
    f2 t = mul2(w, r);
    f2 p = mul2(fma2(n, n, n), r);
    g = fma2(mul2(x, bc2(4.0f)), mul2(p, r), t);
    y = mul2(x, t);
}
#endif  // __CUDACC__

}  // namespace vb
