// Multi-tensor Adam step and Polyak (soft) target update: one launch over all tensors of a network instead of the
// ~15 foreach launches of torch.optim.Adam and the 3 launches per parameter of velora/utils/torch.py:53-64.
// These are the "next" row 1 of SURVEY.md section 8(f): at batch 64 the train step is a chain of tiny launches.
#include "common.cuh"

namespace vb {
namespace optim {

constexpr int kMaxTensors = 24;   // 14 (liquid actor) and 6 (critic) are the sizes on the path; larger lists are chunked by the host

struct AdamArgs {
    float* p[kMaxTensors];
    const float* g[kMaxTensors];
    float* m[kMaxTensors];
    float* v[kMaxTensors];
    const float* step[kMaxTensors];   // per-tensor device scalar, ALREADY incremented for this step (torch capturable layout)
    int start[kMaxTensors + 1];       // prefix sum of element counts
    int n;
};

struct PolyakArgs {
    float* t[kMaxTensors];
    const float* s[kMaxTensors];
    int start[kMaxTensors + 1];
    int n;
};

__device__ __forceinline__ int find_tensor(const int* start, int n, int e) {
    int k = 0;
#pragma unroll 1
    while (k + 1 < n && e >= start[k + 1]) ++k;
    return k;
}

// torch.optim.Adam (no weight decay, no amsgrad, maximize = False), same operation order as torch's single-tensor path:
//   m = lerp(m, g, 1 - b1); v = b2 v + (1 - b2) g g; p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
__global__ void __launch_bounds__(256) adam_multi_kernel(const __grid_constant__ AdamArgs a, float lr, float b1, float b2, float eps) {
    const int total = a.start[a.n];
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int k = find_tensor(a.start, a.n, e);
        const int i = e - a.start[k];
        const float g = a.g[k][i];
        const float step = *a.step[k];
        const float bc1 = 1.0f - powf(b1, step);
        const float bc2 = 1.0f - powf(b2, step);
        float m = a.m[k][i], v = a.v[k][i];
        m = fmaf(1.0f - b1, g - m, m);
        v = fmaf(1.0f - b2, g * g, b2 * v);
        a.m[k][i] = m;
        a.v[k][i] = v;
        const float denom = sqrtf(v) / sqrtf(bc2) + eps;
        a.p[k][i] = a.p[k][i] - (lr / bc1) * (m / denom);
    }
}

__global__ void __launch_bounds__(256) polyak_multi_kernel(const __grid_constant__ PolyakArgs a, float tau) {
    const int total = a.start[a.n];
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int k = find_tensor(a.start, a.n, e);
        const int i = e - a.start[k];
        a.t[k][i] = tau * a.s[k][i] + (1.0f - tau) * a.t[k][i];
    }
}

}  // namespace optim
}  // namespace vb

using namespace vb;

extern "C" int vb_adam_multi(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                             const float* const* steps, const int64_t* sizes, int n_tensors,
                             float lr, float beta1, float beta2, float eps, void* stream) {
    VB_CHECK_ARG(n_tensors >= 0 && (n_tensors == 0 || (params && grads && exp_avg && exp_avg_sq && steps && sizes)), VB_ERR_NULL_POINTER,
                 "vb_adam_multi: null pointer");
    for (int base = 0; base < n_tensors; base += optim::kMaxTensors) {
        optim::AdamArgs a;
        a.n = n_tensors - base < optim::kMaxTensors ? n_tensors - base : optim::kMaxTensors;
        int64_t tot = 0;
        for (int k = 0; k < a.n; ++k) {
            VB_CHECK_ARG(params[base + k] && grads[base + k] && exp_avg[base + k] && exp_avg_sq[base + k] && steps[base + k] && sizes[base + k] >= 0,
                         VB_ERR_NULL_POINTER, "vb_adam_multi: tensor %d has a null pointer", base + k);
            a.p[k] = params[base + k]; a.g[k] = grads[base + k]; a.m[k] = exp_avg[base + k]; a.v[k] = exp_avg_sq[base + k];
            a.step[k] = steps[base + k];
            a.start[k] = (int)tot;
            tot += sizes[base + k];
            VB_CHECK_ARG(tot < (int64_t)1 << 30, VB_ERR_BAD_DIMS, "vb_adam_multi: more than 2^30 elements in one call");
        }
        a.start[a.n] = (int)tot;
        if (tot == 0) continue;
        int grid = (int)((tot + 255) / 256);
        if (grid > 4 * 148) grid = 4 * 148;
        optim::adam_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, lr, beta1, beta2, eps);
        VB_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int vb_polyak_multi(float* const* target, const float* const* source, const int64_t* sizes, int n_tensors, float tau, void* stream) {
    VB_CHECK_ARG(n_tensors >= 0 && (n_tensors == 0 || (target && source && sizes)), VB_ERR_NULL_POINTER, "vb_polyak_multi: null pointer");
    for (int base = 0; base < n_tensors; base += optim::kMaxTensors) {
        optim::PolyakArgs a;
        a.n = n_tensors - base < optim::kMaxTensors ? n_tensors - base : optim::kMaxTensors;
        int64_t tot = 0;
        for (int k = 0; k < a.n; ++k) {
            VB_CHECK_ARG(target[base + k] && source[base + k] && sizes[base + k] >= 0, VB_ERR_NULL_POINTER, "vb_polyak_multi: tensor %d has a null pointer", base + k);
            a.t[k] = target[base + k]; a.s[k] = source[base + k];
            a.start[k] = (int)tot;
            tot += sizes[base + k];
            VB_CHECK_ARG(tot < (int64_t)1 << 30, VB_ERR_BAD_DIMS, "vb_polyak_multi: more than 2^30 elements in one call");
        }
        a.start[a.n] = (int)tot;
        if (tot == 0) continue;
        int grid = (int)((tot + 255) / 256);
        if (grid > 4 * 148) grid = 4 * 148;
        optim::polyak_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, tau);
        VB_LAUNCH_CHECK();
    }
    return 0;
}
