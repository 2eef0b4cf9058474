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

// ------------------------------------------------------------------------------------------------ SAC policy head
// The tail of SACActor.forward (velora/models/sac/continuous.py:121-143) as one kernel forward and one backward instead of
// ~20 + ~30 elementwise launches:
//   mean, log_std = chunk(x);  log_std = clamp(log_std, lo, hi);  std = exp(log_std);  x_t = mean + std * eps
//   a = tanh(x_t);  actions = a * scale + bias
//   log_prob = sum_a [ Normal(mean, std).log_prob(x_t) - log(scale * (1 - a^2) + 1e-6) ]
// eps is drawn by the caller with torch's own generator (Normal.rsample), so seeded runs stay identical to the reference.
namespace vb {
namespace optim {

constexpr float kHalfLog2Pi = 0.91893853320467274178f;

__global__ void __launch_bounds__(256) sac_head_fwd_kernel(const float* __restrict__ x, const float* __restrict__ eps,
                                                           const float* __restrict__ scale, const float* __restrict__ bias,
                                                           float lo, float hi, int64_t B, int A, float* __restrict__ actions,
                                                           float* __restrict__ log_prob, float* __restrict__ a_norm,
                                                           float* __restrict__ std_out) {
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
        float lp = 0.f;
        for (int j = 0; j < A; ++j) {
            const float mean = x[b * 2 * A + j];
            const float ls = min_keep_nan(max_keep_nan(x[b * 2 * A + A + j], lo), hi)   /* torch.clamp keeps NaN */;
            const float sd = expf(ls);
            const float e = eps[b * A + j];
            const float xt = fmaf(sd, e, mean);
            const float an = tanhf(xt);
            const float sc = __ldg(scale + j);
            actions[b * A + j] = fmaf(an, sc, __ldg(bias + j));
            const float diff = xt - mean;                       // the reference evaluates the density at x_t, not at eps
            lp += -(diff * diff) / (2.0f * sd * sd) - ls - kHalfLog2Pi - logf(sc * (1.0f - an * an) + 1e-6f);
            a_norm[b * A + j] = an;
            std_out[b * A + j] = sd;
        }
        log_prob[b] = lp;
    }
}

// d_actions [B,A] or NULL, d_log_prob [B] or NULL -> dx [B, 2A]
__global__ void __launch_bounds__(256) sac_head_bwd_kernel(const float* __restrict__ x, const float* __restrict__ eps,
                                                           const float* __restrict__ scale, float lo, float hi, int64_t B, int A,
                                                           const float* __restrict__ a_norm, const float* __restrict__ std_in,
                                                           const float* __restrict__ d_actions, const float* __restrict__ d_log_prob,
                                                           float* __restrict__ dx) {
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
        const float glp = d_log_prob ? d_log_prob[b] : 0.f;
        for (int j = 0; j < A; ++j) {
            const float an = a_norm[b * A + j], sd = std_in[b * A + j], e = eps[b * A + j], sc = __ldg(scale + j);
            const float one_m = 1.0f - an * an;
            float d_an = d_actions ? d_actions[b * A + j] * sc : 0.f;
            d_an += glp * (2.0f * sc * an) / (sc * one_m + 1e-6f);
            const float d_xt = d_an * one_m;
            // Normal.log_prob's own partials w.r.t. x_t and mean cancel (+-eps/std); w.r.t. std: (eps^2 - 1)/std - eps^2/std = -1/std
            const float raw = x[b * 2 * A + A + j];
            const float pass = (raw >= lo && raw <= hi) ? 1.0f : 0.f;      // clamp passes the gradient on the closed interval
            dx[b * 2 * A + j] = d_xt;
            dx[b * 2 * A + A + j] = (d_xt * e * sd - glp) * pass;
        }
    }
}

}  // namespace optim
}  // namespace vb

extern "C" int vb_sac_head_fwd(const float* x, const float* eps, const float* scale, const float* bias, float log_std_min,
                               float log_std_max, int64_t B, int A, float* actions, float* log_prob, float* a_norm, float* std_out,
                               void* stream) {
    VB_CHECK_ARG(B >= 0 && A >= 1, VB_ERR_BAD_DIMS, "vb_sac_head_fwd: bad dims");
    if (B == 0) return 0;
    VB_CHECK_ARG(x && eps && scale && bias && actions && log_prob && a_norm && std_out, VB_ERR_NULL_POINTER, "vb_sac_head_fwd: null pointer");
    int64_t blocks = (B + 255) / 256;
    int grid = (int)(blocks < 8 * 148 ? blocks : 8 * 148);
    optim::sac_head_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, eps, scale, bias, log_std_min, log_std_max, B, A, actions, log_prob,
                                                                        a_norm, std_out);
    VB_LAUNCH_CHECK();
    return 0;
}

extern "C" int vb_sac_head_bwd(const float* x, const float* eps, const float* scale, float log_std_min, float log_std_max, int64_t B, int A,
                               const float* a_norm, const float* std_in, const float* d_actions, const float* d_log_prob, float* dx,
                               void* stream) {
    VB_CHECK_ARG(B >= 0 && A >= 1, VB_ERR_BAD_DIMS, "vb_sac_head_bwd: bad dims");
    if (B == 0) return 0;
    VB_CHECK_ARG(x && eps && scale && a_norm && std_in && dx, VB_ERR_NULL_POINTER, "vb_sac_head_bwd: null pointer");
    int64_t blocks = (B + 255) / 256;
    int grid = (int)(blocks < 8 * 148 ? blocks : 8 * 148);
    optim::sac_head_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, eps, scale, log_std_min, log_std_max, B, A, a_norm, std_in, d_actions,
                                                                        d_log_prob, dx);
    VB_LAUNCH_CHECK();
    return 0;
}
