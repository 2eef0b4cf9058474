// Parameter packing / gradient unpacking between the host framework's seven SparseLinear layers and the single
// pre-masked kernel-layout block (see include/velora_b200.h).  Replaces the reference's per-forward `weight * mask`
// (velora/models/lnn/sparse.py:81) and autograd's `grad * mask` on the way back.
#include "common.cuh"

namespace vb {

struct LiquidSrc {
    const float* w_in; const float* b_in; const void* m_in;
    const float* w_h[5]; const float* b_h[5]; const void* m_h[5];
    const float* w_out; const float* b_out; const void* m_out;
    int t_in, t_h, t_out;
};

// element (r, c) of an [R, C] mask; types >= VB_MASK_TRANSPOSED are stored column-major (the reference keeps
// `abs(mask.T)`, a transposed view of a row-major [C, R] tensor: velora/models/lnn/ncp.py:82, cell.py:111)
__device__ __forceinline__ float mask_rc(const void* m, int type, int r, int c, int R, int C) {
    const int64_t idx = (type & VB_MASK_TRANSPOSED) ? (int64_t)c * R + r : (int64_t)r * C + c;
    return (type & 1) == VB_MASK_I32 ? (float)reinterpret_cast<const int*>(m)[idx] : reinterpret_cast<const float*>(m)[idx];
}

// effective (masked) weight of head slot hs in {g,h,f,proj} at [j][i]; the f slot is the sum of both f heads
__device__ __forceinline__ float head_w(const LiquidSrc& s, int hs, int j, int i, int K, int Nc) {
    int64_t e = (int64_t)j * K + i;
    if (hs == 0) return s.w_h[0][e] * mask_rc(s.m_h[0], s.t_h, j, i, Nc, K);
    if (hs == 1) return s.w_h[1][e] * mask_rc(s.m_h[1], s.t_h, j, i, Nc, K);
    if (hs == 2) return s.w_h[2][e] * mask_rc(s.m_h[2], s.t_h, j, i, Nc, K) + s.w_h[3][e] * mask_rc(s.m_h[3], s.t_h, j, i, Nc, K);
    return s.w_h[4][e] * mask_rc(s.m_h[4], s.t_h, j, i, Nc, K);
}
__device__ __forceinline__ float head_b(const LiquidSrc& s, int hs, int j) {
    if (hs == 0) return s.b_h[0][j];
    if (hs == 1) return s.b_h[1][j];
    if (hs == 2) return s.b_h[2][j] + s.b_h[3][j];
    return s.b_h[4][j];
}

__global__ void liquid_pack_kernel(LiquidSrc s, LiquidLayout L, float* __restrict__ packed) {
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < L.total; idx += gridDim.x * blockDim.x) {
        float v = 0.f;
        if (idx < L.b_in) {                       // win_t [I][NIp]
            int i = idx / L.NIp, j = idx % L.NIp;
            if (j < L.Ni && s.w_in) v = s.w_in[j * L.I + i] * mask_rc(s.m_in, s.t_in, j, i, L.Ni, L.I);
        } else if (idx < L.wh_t) {                // b_in
            int j = idx - L.b_in;
            if (j < L.Ni && s.b_in) v = s.b_in[j];
        } else if (idx < L.b_h) {                 // wh_t [K][4][NCp]
            int e = idx - L.wh_t;
            int i = e / L.H4, r = e % L.H4;
            int hs = r / L.NCp, j = r % L.NCp;
            if (j < L.Nc) v = head_w(s, hs, j, i, L.K, L.Nc);
        } else if (idx < L.wout) {                // b_h [4][NCp]
            int r = idx - L.b_h;
            int hs = r / L.NCp, j = r % L.NCp;
            if (j < L.Nc) v = head_b(s, hs, j);
        } else if (idx < L.b_out) {               // wout [O][NCp]
            int e = idx - L.wout;
            int o = e / L.NCp, j = e % L.NCp;
            if (j < L.Nc && s.w_out) v = s.w_out[o * L.Nc + j] * mask_rc(s.m_out, s.t_out, o, j, L.O, L.Nc);
        } else if (idx < L.f_total) {             // b_out
            int o = idx - L.b_out;
            if (o < L.O && s.b_out) v = s.b_out[o];
        } else if (idx < L.wh) {                  // win [Ni][Ip]
            int e = idx - L.win;
            int j = e / L.Ip, i = e % L.Ip;
            if (i < L.I && s.w_in) v = s.w_in[j * L.I + i] * mask_rc(s.m_in, s.t_in, j, i, L.Ni, L.I);
        } else if (idx < L.wout_t) {              // wh [4*NCp][Kp]
            int e = idx - L.wh;
            int r = e / L.Kp, i = e % L.Kp;
            int hs = r / L.NCp, j = r % L.NCp;
            if (j < L.Nc && i < L.K) v = head_w(s, hs, j, i, L.K, L.Nc);
        } else {                                  // wout_t [Nc][Op]
            int e = idx - L.wout_t;
            int j = e / L.Op, o = e % L.Op;
            if (o < L.O && s.w_out) v = s.w_out[o * L.Nc + j] * mask_rc(s.m_out, s.t_out, o, j, L.O, L.Nc);
        }
        packed[idx] = v;
    }
}

struct LiquidGradDst {
    float* g_w_in; float* g_b_in; const void* m_in;
    float* g_w_h[5]; float* g_b_h[5]; const void* m_h[5];
    float* g_w_out; float* g_b_out; const void* m_out;
    int t_in, t_h, t_out;
};

__device__ __forceinline__ void put(float* p, int64_t i, float v, int acc) {
    if (p) p[i] = acc ? p[i] + v : v;
}

// one thread per element of the seven layers' weight/bias tensors
__global__ void liquid_unpack_kernel(const float* __restrict__ dp, LiquidGradDst d, LiquidLayout L, int acc) {
    const int n_in = L.Ni * L.I, n_h = L.Nc * L.K, n_out = L.O * L.Nc;
    const int total = n_in + L.Ni + 5 * (n_h + L.Nc) + n_out + L.O;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        int e = idx;
        if (e < n_in) {
            int j = e / L.I, i = e % L.I;
            if (d.g_w_in) put(d.g_w_in, e, dp[L.win_t + i * L.NIp + j] * mask_rc(d.m_in, d.t_in, j, i, L.Ni, L.I), acc);
            continue;
        }
        e -= n_in;
        if (e < L.Ni) { put(d.g_b_in, e, dp[L.b_in + e], acc); continue; }
        e -= L.Ni;
        if (e < 5 * (n_h + L.Nc)) {
            int hd = e / (n_h + L.Nc);
            int r = e % (n_h + L.Nc);
            int hs = hd < 2 ? hd : (hd < 4 ? 2 : 3);
            if (r < n_h) {
                int j = r / L.K, i = r % L.K;
                put(d.g_w_h[hd], r, dp[L.wh_t + i * L.H4 + hs * L.NCp + j] * mask_rc(d.m_h[hd], d.t_h, j, i, L.Nc, L.K), acc);
            } else {
                int j = r - n_h;
                put(d.g_b_h[hd], j, dp[L.b_h + hs * L.NCp + j], acc);
            }
            continue;
        }
        e -= 5 * (n_h + L.Nc);
        if (e < n_out) {
            int o = e / L.Nc, j = e % L.Nc;
            if (d.g_w_out) put(d.g_w_out, e, dp[L.wout + o * L.NCp + j] * mask_rc(d.m_out, d.t_out, o, j, L.O, L.Nc), acc);
            continue;
        }
        e -= n_out;
        put(d.g_b_out, e, dp[L.b_out + e], acc);
    }
}

// ------------------------------------------------------------------------------------------------ NCP network
struct NcpSrc {
    const float* w[3]; const float* b[3]; const void* m[3];
    int t;
};
// NCP layers share one mask type code; the three masks have the same storage kind (all built by `abs(mask.T)`)

__global__ void ncp_pack_kernel(NcpSrc s, NcpLayout L, float* __restrict__ packed) {
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < L.total; idx += gridDim.x * blockDim.x) {
        float v = 0.f;
        if (idx < L.b1) {                          // w1_t [I][NIp]
            int i = idx / L.NIp, j = idx % L.NIp;
            if (j < L.Ni) v = s.w[0][j * L.I + i] * mask_rc(s.m[0], s.t, j, i, L.Ni, L.I);
        } else if (idx < L.w2_t) {
            int j = idx - L.b1;
            if (j < L.Ni) v = s.b[0][j];
        } else if (idx < L.b2) {                   // w2_t [Ni][NCp]
            int e = idx - L.w2_t;
            int i = e / L.NCp, j = e % L.NCp;
            if (j < L.Nc) v = s.w[1][j * L.Ni + i] * mask_rc(s.m[1], s.t, j, i, L.Nc, L.Ni);
        } else if (idx < L.w3) {
            int j = idx - L.b2;
            if (j < L.Nc) v = s.b[1][j];
        } else if (idx < L.b3) {                   // w3 [O][NCp]
            int e = idx - L.w3;
            int o = e / L.NCp, j = e % L.NCp;
            if (j < L.Nc) v = s.w[2][o * L.Nc + j] * mask_rc(s.m[2], s.t, o, j, L.O, L.Nc);
        } else if (idx < L.f_total) {
            int o = idx - L.b3;
            if (o < L.O) v = s.b[2][o];
        } else if (idx < L.w2) {                   // w1 [Ni][Ip]
            int e = idx - L.w1;
            int j = e / L.Ip, i = e % L.Ip;
            if (i < L.I) v = s.w[0][j * L.I + i] * mask_rc(s.m[0], s.t, j, i, L.Ni, L.I);
        } else if (idx < L.w3_t) {                 // w2 [Nc][NIp]
            int e = idx - L.w2;
            int j = e / L.NIp, i = e % L.NIp;
            if (i < L.Ni) v = s.w[1][j * L.Ni + i] * mask_rc(s.m[1], s.t, j, i, L.Nc, L.Ni);
        } else {                                   // w3_t [Nc][Op]
            int e = idx - L.w3_t;
            int j = e / L.Op, o = e % L.Op;
            if (o < L.O) v = s.w[2][o * L.Nc + j] * mask_rc(s.m[2], s.t, o, j, L.O, L.Nc);
        }
        packed[idx] = v;
    }
}

struct NcpGradDst {
    float* g_w[3]; float* g_b[3]; const void* m[3];
    int t;
};

__global__ void ncp_unpack_kernel(const float* __restrict__ dp, NcpGradDst d, NcpLayout L, int acc) {
    const int n1 = L.Ni * L.I, n2 = L.Nc * L.Ni, n3 = L.O * L.Nc;
    const int total = n1 + L.Ni + n2 + L.Nc + n3 + L.O;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        int e = idx;
        if (e < n1) { int j = e / L.I, i = e % L.I; put(d.g_w[0], e, dp[L.w1_t + i * L.NIp + j] * mask_rc(d.m[0], d.t, j, i, L.Ni, L.I), acc); continue; }
        e -= n1;
        if (e < L.Ni) { put(d.g_b[0], e, dp[L.b1 + e], acc); continue; }
        e -= L.Ni;
        if (e < n2) { int j = e / L.Ni, i = e % L.Ni; put(d.g_w[1], e, dp[L.w2_t + i * L.NCp + j] * mask_rc(d.m[1], d.t, j, i, L.Nc, L.Ni), acc); continue; }
        e -= n2;
        if (e < L.Nc) { put(d.g_b[1], e, dp[L.b2 + e], acc); continue; }
        e -= L.Nc;
        if (e < n3) { int o = e / L.Nc, j = e % L.Nc; put(d.g_w[2], e, dp[L.w3 + o * L.NCp + j] * mask_rc(d.m[2], d.t, o, j, L.O, L.Nc), acc); continue; }
        e -= n3;
        put(d.g_b[2], e, dp[L.b3 + e], acc);
    }
}

static inline int grid_for(int n) {
    int g = (n + 255) / 256;
    return g < 1 ? 1 : (g > 1024 ? 1024 : g);
}

}  // namespace vb

using namespace vb;

extern "C" int64_t vb_liquid_packed_floats(int I, int Ni, int Nc, int O) {
    if (!dims_ok(I, Ni, Nc, O)) return VB_ERR_BAD_DIMS;
    return LiquidLayout(I, Ni, Nc, O).total;
}
extern "C" int64_t vb_liquid_grad_floats(int I, int Ni, int Nc, int O) {
    if (!dims_ok(I, Ni, Nc, O)) return VB_ERR_BAD_DIMS;
    return LiquidLayout(I, Ni, Nc, O).f_total;
}
extern "C" int64_t vb_ncp_packed_floats(int I, int Ni, int Nc, int O) {
    if (!dims_ok(I, Ni, Nc, O)) return VB_ERR_BAD_DIMS;
    return NcpLayout(I, Ni, Nc, O).total;
}
extern "C" int64_t vb_ncp_grad_floats(int I, int Ni, int Nc, int O) {
    if (!dims_ok(I, Ni, Nc, O)) return VB_ERR_BAD_DIMS;
    return NcpLayout(I, Ni, Nc, O).f_total;
}

extern "C" int vb_liquid_pack(const float* w_in, const float* b_in, const void* m_in, int m_in_type,
                              const float* const* w_heads, const float* const* b_heads, const void* const* m_heads, int m_heads_type,
                              const float* w_out, const float* b_out, const void* m_out, int m_out_type,
                              int I, int Ni, int Nc, int O, float* packed, void* stream) {
    VB_CHECK_ARG(dims_ok(I, Ni, Nc, O), VB_ERR_BAD_DIMS, "vb_liquid_pack: bad dims I=%d Ni=%d Nc=%d O=%d", I, Ni, Nc, O);
    // the inter and / or motor layer may be absent (all three of its pointers NULL): a standalone NCPLiquidCell packs its five heads only
    VB_CHECK_ARG(w_heads && b_heads && m_heads && packed && (w_in ? (b_in && m_in) : (!b_in && !m_in)) && (w_out ? (b_out && m_out) : (!b_out && !m_out)),
                 VB_ERR_NULL_POINTER, "vb_liquid_pack: null pointer");
    LiquidSrc s;
    s.w_in = w_in; s.b_in = b_in; s.m_in = m_in;
    for (int k = 0; k < 5; ++k) {
        VB_CHECK_ARG(w_heads[k] && b_heads[k] && m_heads[k], VB_ERR_NULL_POINTER, "vb_liquid_pack: null head pointer %d", k);
        s.w_h[k] = w_heads[k]; s.b_h[k] = b_heads[k]; s.m_h[k] = m_heads[k];
    }
    s.w_out = w_out; s.b_out = b_out; s.m_out = m_out;
    s.t_in = m_in_type; s.t_h = m_heads_type; s.t_out = m_out_type;
    LiquidLayout L(I, Ni, Nc, O);
    liquid_pack_kernel<<<grid_for(L.total), 256, 0, (cudaStream_t)stream>>>(s, L, packed);
    VB_LAUNCH_CHECK();
    return 0;
}

extern "C" int vb_liquid_unpack_grads(const float* dpacked, const void* m_in, int m_in_type,
                                      const void* const* m_heads, int m_heads_type, const void* m_out, int m_out_type,
                                      float* g_w_in, float* g_b_in, float* const* g_w_heads, float* const* g_b_heads,
                                      float* g_w_out, float* g_b_out,
                                      int I, int Ni, int Nc, int O, int accumulate, void* stream) {
    VB_CHECK_ARG(dims_ok(I, Ni, Nc, O), VB_ERR_BAD_DIMS, "vb_liquid_unpack_grads: bad dims");
    VB_CHECK_ARG(dpacked && m_heads && g_w_heads && g_b_heads && (m_in || !g_w_in) && (m_out || !g_w_out), VB_ERR_NULL_POINTER,
                 "vb_liquid_unpack_grads: null pointer");
    LiquidGradDst d;
    d.g_w_in = g_w_in; d.g_b_in = g_b_in; d.m_in = m_in;
    for (int k = 0; k < 5; ++k) {
        VB_CHECK_ARG(m_heads[k], VB_ERR_NULL_POINTER, "vb_liquid_unpack_grads: null head mask %d", k);
        d.g_w_h[k] = g_w_heads[k]; d.g_b_h[k] = g_b_heads[k]; d.m_h[k] = m_heads[k];
    }
    d.g_w_out = g_w_out; d.g_b_out = g_b_out; d.m_out = m_out;
    d.t_in = m_in_type; d.t_h = m_heads_type; d.t_out = m_out_type;
    LiquidLayout L(I, Ni, Nc, O);
    int total = Ni * I + Ni + 5 * (Nc * (Ni + Nc) + Nc) + O * Nc + O;
    liquid_unpack_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(dpacked, d, L, accumulate);
    VB_LAUNCH_CHECK();
    return 0;
}

extern "C" int vb_ncp_pack(const float* const* w, const float* const* b, const void* const* m, int m_type,
                           int I, int Ni, int Nc, int O, float* packed, void* stream) {
    VB_CHECK_ARG(dims_ok(I, Ni, Nc, O), VB_ERR_BAD_DIMS, "vb_ncp_pack: bad dims");
    VB_CHECK_ARG(w && b && m && packed, VB_ERR_NULL_POINTER, "vb_ncp_pack: null pointer");
    NcpSrc s;
    for (int k = 0; k < 3; ++k) {
        VB_CHECK_ARG(w[k] && b[k] && m[k], VB_ERR_NULL_POINTER, "vb_ncp_pack: null layer pointer %d", k);
        s.w[k] = w[k]; s.b[k] = b[k]; s.m[k] = m[k];
    }
    s.t = m_type;
    NcpLayout L(I, Ni, Nc, O);
    ncp_pack_kernel<<<grid_for(L.total), 256, 0, (cudaStream_t)stream>>>(s, L, packed);
    VB_LAUNCH_CHECK();
    return 0;
}

extern "C" int vb_ncp_unpack_grads(const float* dpacked, const void* const* m, int m_type,
                                   float* const* g_w, float* const* g_b,
                                   int I, int Ni, int Nc, int O, int accumulate, void* stream) {
    VB_CHECK_ARG(dims_ok(I, Ni, Nc, O), VB_ERR_BAD_DIMS, "vb_ncp_unpack_grads: bad dims");
    VB_CHECK_ARG(dpacked && m && g_w && g_b, VB_ERR_NULL_POINTER, "vb_ncp_unpack_grads: null pointer");
    NcpGradDst d;
    for (int k = 0; k < 3; ++k) {
        VB_CHECK_ARG(m[k], VB_ERR_NULL_POINTER, "vb_ncp_unpack_grads: null mask %d", k);
        d.g_w[k] = g_w[k]; d.g_b[k] = g_b[k]; d.m[k] = m[k];
    }
    d.t = m_type;
    NcpLayout L(I, Ni, Nc, O);
    int total = Ni * I + Ni + Nc * Ni + Nc + O * Nc + O;
    ncp_unpack_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(dpacked, d, L, accumulate);
    VB_LAUNCH_CHECK();
    return 0;
}
