// C-ABI entry points that dispatch between the specialised and the generic kernels, plus error/query plumbing.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace vb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static int g_sms[64];
static int g_smem[64];
static bool g_props[64];

static void load_props(int dev) {
    if (g_props[dev & 63]) return;
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess) g_sms[dev & 63] = v;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess) g_smem[dev & 63] = v;
    g_props[dev & 63] = true;
}
int device_sms() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    load_props(dev);
    return g_sms[dev & 63];
}
int device_max_smem_optin() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    load_props(dev);
    return g_smem[dev & 63];
}

namespace small {
bool has_fast_path(int I, int Ni, int Nc, int O);
int fwd(const float*, const float*, const float*, float*, float*, int64_t, int, int, int, int, int, cudaStream_t);
int bwd(const float*, const float*, const float*, const float*, const float*, const float*, const float*, float*, float*, float*,
        int64_t, int, int, int, int, int, void*, int64_t, cudaStream_t);
int64_t bwd_workspace_bytes(int, int, int, int, int64_t);
}  // namespace small
namespace tc {
bool has_tc_path(int I, int Ni, int Nc, int O);
int fwd(const float*, const float*, const float*, float*, float*, void*, int64_t, int64_t, int, int, int, int, int, int, cudaStream_t);
int bwd(const float*, const float*, const float*, const float*, const float*, const float*, const float*, const void*, int64_t, float*, float*,
        float*, int64_t, int, int, int, int, int, int, void*, int64_t, cudaStream_t);
int64_t saved_bytes(int, int, int, int, int64_t, int);
int64_t bwd_workspace_bytes(int, int, int, int, int64_t, int, bool);
}  // namespace tc
// the tensor-core critic kernels take over once the batch fills a few 128-sample tiles per SM
static constexpr int64_t kNcpTcMinBatch = 4096;
namespace ncp_tc {
bool has_path(int I, int Ni, int Nc, int O);
int fwd(const float*, const float*, float*, int64_t, int, int, int, int, cudaStream_t);
int bwd(const float*, const float*, const float*, float*, float*, int64_t, int, int, int, int, void*, int64_t, cudaStream_t);
int64_t bwd_workspace_bytes(int, int, int, int);
}  // namespace ncp_tc
namespace big {
bool has_path(int I, int Ni, int Nc, int O);
int64_t fwd_workspace_bytes(int64_t B);
int64_t saved_bytes(int64_t B, int T);
int fwd(const float*, const float*, const float*, float*, float*, float*, void*, int64_t, int64_t, int, int, void*, int64_t, cudaStream_t);
int64_t bwd_workspace_bytes(int64_t B, int T);
int bwd(const float*, const float*, const float*, const float*, const void*, int64_t, float*, float*, float*, int64_t, int, int, void*, int64_t,
        cudaStream_t);
}  // namespace big
namespace generic {
int liquid_fwd(const float*, const float*, const float*, float*, float*, int64_t, int, int, int, int, int, cudaStream_t);
int64_t liquid_bwd_workspace_bytes(int, int, int, int);
int liquid_bwd(const float*, const float*, const float*, const float*, const float*, const float*, const float*, float*, float*, float*,
               int64_t, int, int, int, int, int, void*, int64_t, cudaStream_t);
int cell_fwd(const float*, const float*, const float*, float*, float*, int64_t, int, int, cudaStream_t);
int cell_bwd(const float*, const float*, const float*, const float*, const float*, float*, float*, float*, int64_t, int, int, void*, int64_t,
             cudaStream_t);
int ncp_fwd(const float*, const float*, float*, int64_t, int, int, int, int, cudaStream_t);
int64_t ncp_bwd_workspace_bytes(int, int, int, int);
int ncp_bwd(const float*, const float*, const float*, float*, float*, int64_t, int, int, int, int, void*, int64_t, cudaStream_t);
}  // namespace generic

}  // namespace vb

using namespace vb;

extern "C" const char* vb_last_error(void) { return g_err; }

extern "C" int vb_query(int what) {
    switch (what) {
        case VB_Q_ABI_VERSION: return VB_ABI_VERSION;
        case VB_Q_BUILT_SM: return 100;
        case VB_Q_DEVICE_SMS: return device_sms();
        case VB_Q_DEVICE_CC: {
            int dev = 0, mj = 0, mn = 0;
            if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
            cudaDeviceGetAttribute(&mj, cudaDevAttrComputeCapabilityMajor, dev);
            cudaDeviceGetAttribute(&mn, cudaDevAttrComputeCapabilityMinor, dev);
            return mj * 10 + mn;
        }
        case VB_Q_MAX_SMEM_OPTIN: return device_max_smem_optin();
        default: set_error("vb_query: unknown key %d", what); return VB_ERR_BAD_DIMS;
    }
}

// 0 = generic tile kernels only, 1 = specialised FFMA kernels, 2 = tensor-core (tcgen05) kernels
extern "C" int vb_liquid_has_fast_path(int I, int Ni, int Nc, int O) {
    if (tc::has_tc_path(I, Ni, Nc, O)) return 2;
    return small::has_fast_path(I, Ni, Nc, O) ? 1 : 0;
}

static inline int64_t max64(int64_t a, int64_t b) { return a > b ? a : b; }

extern "C" int64_t vb_liquid_saved_bytes(int I, int Ni, int Nc, int O, int64_t B, int T, int flags) {
    if (!dims_ok(I, Ni, Nc, O) || B < 0 || T < 1) return VB_ERR_BAD_DIMS;
    if ((flags & (VB_FLAG_FORCE_GENERIC | VB_FLAG_FORCE_FFMA)) || !tc::has_tc_path(I, Ni, Nc, O)) return 0;
    return tc::saved_bytes(I, Ni, Nc, O, B, T);
}
extern "C" int vb_liquid_time_major_ok(int I, int Ni, int Nc, int O, int flags) {
    return !(flags & (VB_FLAG_FORCE_GENERIC | VB_FLAG_FORCE_FFMA)) && tc::has_tc_path(I, Ni, Nc, O);
}
extern "C" int64_t vb_liquid_bwd_workspace_bytes(int I, int Ni, int Nc, int O, int64_t B, int T, int have_saved) {
    if (!dims_ok(I, Ni, Nc, O) || B < 0 || T < 1) return VB_ERR_BAD_DIMS;
    int64_t g = generic::liquid_bwd_workspace_bytes(I, Ni, Nc, O);
    int64_t s = small::has_fast_path(I, Ni, Nc, O) ? small::bwd_workspace_bytes(I, Ni, Nc, O, B) : 0;
    int64_t t = tc::has_tc_path(I, Ni, Nc, O) ? tc::bwd_workspace_bytes(I, Ni, Nc, O, B, T, !have_saved) : 0;
    return max64(max64(g, s), t);
}

// ---- the standalone NCPLiquidCell (cell.py:147-169): one step of the five heads without the inter / motor layers.  The packed block is the
// network's with (I, Ni, Nc, O) = (In, In, Nc, Nc) and the inter / motor layers absent (vb_liquid_pack with NULL w_in / w_out).
extern "C" int64_t vb_cell_bwd_workspace_bytes(int In, int Nc) {
    if (!dims_ok(In, In, Nc, Nc)) return VB_ERR_BAD_DIMS;
    return generic::liquid_bwd_workspace_bytes(In, In, Nc, Nc);
}
extern "C" int vb_cell_fwd(const float* packed, const float* x, const float* h, float* y, float* h_new, int64_t B, int In, int Nc,
                           void* stream) {
    VB_CHECK_ARG(dims_ok(In, In, Nc, Nc) && B >= 0, VB_ERR_BAD_DIMS, "vb_cell_fwd: bad dims B=%lld (%d,%d)", (long long)B, In, Nc);
    if (B == 0) return 0;
    VB_CHECK_ARG(packed && x && y && h_new, VB_ERR_NULL_POINTER, "vb_cell_fwd: null pointer");
    return generic::cell_fwd(packed, x, h, y, h_new, B, In, Nc, (cudaStream_t)stream);
}
extern "C" int vb_cell_bwd(const float* packed, const float* x, const float* h, const float* dy, const float* dh_new, float* dx, float* dh,
                           float* dpacked, int64_t B, int In, int Nc, void* workspace, int64_t workspace_bytes, void* stream) {
    VB_CHECK_ARG(dims_ok(In, In, Nc, Nc) && B >= 0, VB_ERR_BAD_DIMS, "vb_cell_bwd: bad dims B=%lld (%d,%d)", (long long)B, In, Nc);
    VB_CHECK_ARG(dpacked, VB_ERR_NULL_POINTER, "vb_cell_bwd: null dpacked");
    if (B == 0) {
        LiquidLayout L(In, In, Nc, Nc);
        return (int)cudaMemsetAsync(dpacked, 0, (size_t)L.f_total * 4, (cudaStream_t)stream);
    }
    VB_CHECK_ARG(packed && x && dy, VB_ERR_NULL_POINTER, "vb_cell_bwd: null pointer");
    return generic::cell_bwd(packed, x, h, dy, dh_new, dx, dh, dpacked, B, In, Nc, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int64_t vb_ncp_bwd_workspace_bytes(int I, int Ni, int Nc, int O, int64_t B) {
    if (!dims_ok(I, Ni, Nc, O)) return VB_ERR_BAD_DIMS;
    int64_t t = (B >= kNcpTcMinBatch && ncp_tc::has_path(I, Ni, Nc, O)) ? ncp_tc::bwd_workspace_bytes(I, Ni, Nc, O) : 0;
    return max64(generic::ncp_bwd_workspace_bytes(I, Ni, Nc, O), t);
}

extern "C" int vb_liquid_fwd(const float* packed, const float* x, const float* h0, float* y, float* h_traj,
                             int64_t B, int T, int I, int Ni, int Nc, int O,
                             void* saved, int64_t saved_bytes, int flags, void* stream) {
    VB_CHECK_ARG(dims_ok(I, Ni, Nc, O) && B >= 0 && T >= 1, VB_ERR_BAD_DIMS, "vb_liquid_fwd: bad dims B=%lld T=%d (%d,%d,%d,%d)", (long long)B, T, I, Ni, Nc, O);
    if (B == 0) return 0;      // empty batch: nothing to launch (tensors of zero rows have no storage to point at)
    VB_CHECK_ARG(packed && x && y && h_traj, VB_ERR_NULL_POINTER, "vb_liquid_fwd: null pointer");
    if (!(flags & (VB_FLAG_FORCE_GENERIC | VB_FLAG_FORCE_FFMA)) && tc::has_tc_path(I, Ni, Nc, O))
        return tc::fwd(packed, x, h0, y, h_traj, saved, saved_bytes, B, T, I, Ni, Nc, O, flags & VB_FLAG_LAYOUT_MASK, (cudaStream_t)stream);
    VB_CHECK_ARG(!(flags & VB_FLAG_LAYOUT_MASK) || T == 1, VB_ERR_LAYOUT,
                 "vb_liquid_fwd: time-major tensors are only supported by the tensor-core kernels (vb_liquid_time_major_ok)");
    if (!(flags & VB_FLAG_FORCE_GENERIC) && small::has_fast_path(I, Ni, Nc, O))
        return small::fwd(packed, x, h0, y, h_traj, B, T, I, Ni, Nc, O, (cudaStream_t)stream);
    return generic::liquid_fwd(packed, x, h0, y, h_traj, B, T, I, Ni, Nc, O, (cudaStream_t)stream);
}

extern "C" int vb_liquid_bwd(const float* packed, const float* x, const float* h0, const float* h_traj,
                             const float* dy, const float* dh_traj, const float* dh_last,
                             const void* saved, int64_t saved_bytes,
                             float* dx, float* dh0, float* dpacked,
                             int64_t B, int T, int I, int Ni, int Nc, int O,
                             void* workspace, int64_t workspace_bytes, int flags, void* stream) {
    VB_CHECK_ARG(dims_ok(I, Ni, Nc, O) && B >= 0 && T >= 1, VB_ERR_BAD_DIMS, "vb_liquid_bwd: bad dims");
    VB_CHECK_ARG(dpacked && (B == 0 || (packed && x && h_traj && dy)), VB_ERR_NULL_POINTER, "vb_liquid_bwd: null pointer");
    if (B == 0) {
        VB_CUDA(cudaMemsetAsync(dpacked, 0, (size_t)LiquidLayout(I, Ni, Nc, O).f_total * 4, (cudaStream_t)stream));
        return 0;
    }
    if (!(flags & (VB_FLAG_FORCE_GENERIC | VB_FLAG_FORCE_FFMA)) && tc::has_tc_path(I, Ni, Nc, O))
        return tc::bwd(packed, x, h0, h_traj, dy, dh_traj, dh_last, saved, saved_bytes, dx, dh0, dpacked, B, T, I, Ni, Nc, O,
                       flags & VB_FLAG_LAYOUT_MASK, workspace, workspace_bytes, (cudaStream_t)stream);
    VB_CHECK_ARG(!(flags & VB_FLAG_LAYOUT_MASK) || T == 1, VB_ERR_LAYOUT,
                 "vb_liquid_bwd: time-major tensors are only supported by the tensor-core kernels (vb_liquid_time_major_ok)");
    if (!(flags & VB_FLAG_FORCE_GENERIC) && small::has_fast_path(I, Ni, Nc, O))
        return small::bwd(packed, x, h0, h_traj, dy, dh_traj, dh_last, dx, dh0, dpacked, B, T, I, Ni, Nc, O, workspace, workspace_bytes, (cudaStream_t)stream);
    return generic::liquid_bwd(packed, x, h0, h_traj, dy, dh_traj, dh_last, dx, dh0, dpacked, B, T, I, Ni, Nc, O, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int vb_ncp_fwd(const float* packed, const float* x, float* y, int64_t B, int I, int Ni, int Nc, int O, int flags, void* stream) {
    VB_CHECK_ARG(dims_ok(I, Ni, Nc, O) && B >= 0, VB_ERR_BAD_DIMS, "vb_ncp_fwd: bad dims");
    if (B == 0) return 0;
    VB_CHECK_ARG(packed && x && y, VB_ERR_NULL_POINTER, "vb_ncp_fwd: null pointer");
    // tensor-core forward for the critic shapes once the batch fills a few tiles per SM; small batches stay on the tile kernels
    if (!(flags & (VB_FLAG_FORCE_GENERIC | VB_FLAG_FORCE_FFMA)) && B >= kNcpTcMinBatch && ncp_tc::has_path(I, Ni, Nc, O))
        return ncp_tc::fwd(packed, x, y, B, I, Ni, Nc, O, (cudaStream_t)stream);
    return generic::ncp_fwd(packed, x, y, B, I, Ni, Nc, O, (cudaStream_t)stream);
}

extern "C" int vb_ncp_bwd(const float* packed, const float* x, const float* dy, float* dx, float* dpacked,
                          int64_t B, int I, int Ni, int Nc, int O, void* workspace, int64_t workspace_bytes, int flags, void* stream) {
    VB_CHECK_ARG(dims_ok(I, Ni, Nc, O) && B >= 0, VB_ERR_BAD_DIMS, "vb_ncp_bwd: bad dims");
    VB_CHECK_ARG((dpacked || dx) && (B == 0 || (packed && x && dy)), VB_ERR_NULL_POINTER, "vb_ncp_bwd: null pointer");
    if (B == 0) {
        if (dpacked) VB_CUDA(cudaMemsetAsync(dpacked, 0, (size_t)NcpLayout(I, Ni, Nc, O).f_total * 4, (cudaStream_t)stream));
        return 0;
    }
    if (!(flags & (VB_FLAG_FORCE_GENERIC | VB_FLAG_FORCE_FFMA)) && B >= kNcpTcMinBatch && ncp_tc::has_path(I, Ni, Nc, O))
        return ncp_tc::bwd(packed, x, dy, dx, dpacked, B, I, Ni, Nc, O, workspace, workspace_bytes, (cudaStream_t)stream);
    return generic::ncp_bwd(packed, x, dy, dx, dpacked, B, I, Ni, Nc, O, workspace, workspace_bytes, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------ bf16 forward (config 4 class)
extern "C" int vb_liquid_bf16_ok(int I, int Ni, int Nc, int O) { return big::has_path(I, Ni, Nc, O) ? 1 : 0; }

extern "C" int64_t vb_liquid_bf16_fwd_workspace_bytes(int I, int Ni, int Nc, int O, int64_t B, int T) {
    (void)T;
    if (!big::has_path(I, Ni, Nc, O) || B < 0) return VB_ERR_BAD_DIMS;
    return big::fwd_workspace_bytes(B);
}

extern "C" int64_t vb_liquid_bf16_saved_bytes(int I, int Ni, int Nc, int O, int64_t B, int T) {
    if (!big::has_path(I, Ni, Nc, O) || B < 0 || T < 1) return VB_ERR_BAD_DIMS;
    return big::saved_bytes(B, T);
}

extern "C" int64_t vb_liquid_bf16_bwd_workspace_bytes(int I, int Ni, int Nc, int O, int64_t B, int T) {
    if (!big::has_path(I, Ni, Nc, O) || B < 0 || T < 1) return VB_ERR_BAD_DIMS;
    return big::bwd_workspace_bytes(B, T);
}

extern "C" int vb_liquid_bf16_bwd(const float* packed, const float* x, const float* dy, const float* dh_last, const void* saved,
                                  int64_t saved_bytes, float* dx, float* dh0, float* dpacked, int64_t B, int T, int I, int Ni, int Nc, int O,
                                  void* workspace, int64_t workspace_bytes, int flags, void* stream) {
    VB_CHECK_ARG(B >= 0 && T >= 1, VB_ERR_BAD_DIMS, "vb_liquid_bf16_bwd: bad dims B=%lld T=%d", (long long)B, T);
    VB_CHECK_ARG(big::has_path(I, Ni, Nc, O), VB_ERR_BAD_DIMS, "vb_liquid_bf16_bwd: no bf16 tensor-core kernel for (%d,%d,%d,%d)", I, Ni, Nc, O);
    VB_CHECK_ARG(dpacked && (B == 0 || (packed && x && dy && saved)), VB_ERR_NULL_POINTER, "vb_liquid_bf16_bwd: null pointer");
    if (B == 0) {
        VB_CUDA(cudaMemsetAsync(dpacked, 0, (size_t)LiquidLayout(I, Ni, Nc, O).f_total * 4, (cudaStream_t)stream));
        return 0;
    }
    return big::bwd(packed, x, dy, dh_last, saved, saved_bytes, dx, dh0, dpacked, B, T, flags, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int vb_liquid_bf16_fwd(const float* packed, const float* x, const float* h0, float* y, float* h_last, float* h_traj,
                                  void* saved, int64_t saved_bytes,
                                  int64_t B, int T, int I, int Ni, int Nc, int O, void* workspace, int64_t workspace_bytes, int flags,
                                  void* stream) {
    VB_CHECK_ARG(B >= 0 && T >= 1, VB_ERR_BAD_DIMS, "vb_liquid_bf16_fwd: bad dims B=%lld T=%d", (long long)B, T);
    VB_CHECK_ARG(big::has_path(I, Ni, Nc, O), VB_ERR_BAD_DIMS, "vb_liquid_bf16_fwd: no bf16 tensor-core kernel for (%d,%d,%d,%d)", I, Ni, Nc, O);
    if (B == 0) return 0;
    VB_CHECK_ARG(packed && x && y, VB_ERR_NULL_POINTER, "vb_liquid_bf16_fwd: null pointer");
    return big::fwd(packed, x, h0, y, h_last, h_traj, saved, saved_bytes, B, T, flags, workspace, workspace_bytes, (cudaStream_t)stream);
}
