// The loss arithmetic of NeuroFlowCT._update_critics / _train_step around the networks (velora/models/nf/agent.py:194-209, 236-245,
// modules.py:700-706) as one kernel forward and one backward per loss instead of ~40 framework launches per train step:
//
//   critic:   target = r + (1 - done) * gamma * (min(q1_t, q2_t) - exp(log_alpha) * logp_next)           (no gradient: agent.py:194-202)
//             c1 = mean((q1 - target)^2), c2 = mean((q2 - target)^2)                                     (nn.MSELoss, agent.py:204-207)
//   actor:    mean(exp(log_alpha) * logp - min(q1, q2))                                                  (agent.py:238-241)
//   entropy:  -mean(log_alpha * (logp + target_entropy))                                                 (modules.py:700-706)
//
// Every forward is a grid-stride pass with per-CTA partial sums and a LAST-BLOCK finalisation in a fixed order (ticket counter in the
// workspace): one launch, deterministic.  Batch means divide by B (the local batch; the data-parallel path averages gradients afterwards).
#include "common.cuh"

namespace vb {
namespace sacl {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = 296;
constexpr int kMaxSums = 3;

// block-reduce `v[0..N)` and, in the last block to arrive, sum all blocks' partials in block order into out[0..N) * scale
struct NoPost { __device__ void operator()(float*) const {} };
template <int N, class Post = NoPost>
__device__ __forceinline__ void finalize(float (&v)[N], float* __restrict__ ws, float* __restrict__ out, float scale, Post post = Post()) {
    __shared__ float sm[N][kThreads / 32];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        float x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) sm[k][warp] = x;
    }
    __syncthreads();
    unsigned int* ticket = reinterpret_cast<unsigned int*>(ws);
    float* partials = ws + 4;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) t += sm[k][w];
            partials[blockIdx.x * kMaxSums + k] = t;
        }
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x < N) {
        __threadfence();
        float t = 0.f;
        for (unsigned int b = 0; b < gridDim.x; ++b) t += __ldcg(partials + b * kMaxSums + threadIdx.x);
        out[threadIdx.x] = t * scale;
        __syncwarp((1u << N) - 1u);
        if (threadIdx.x == 0) {
            post(out);
            *ticket = 0u;          // ready for the next launch on this stream
        }
    }
}

__global__ void __launch_bounds__(kThreads) critic_loss_fwd_kernel(const float* __restrict__ q1, const float* __restrict__ q2,
                                                                   const float* __restrict__ q1t, const float* __restrict__ q2t,
                                                                   const float* __restrict__ logp, const float* __restrict__ rew,
                                                                   const float* __restrict__ done, const float* __restrict__ log_alpha, float gamma,
                                                                   int64_t B, float* __restrict__ target, float* __restrict__ losses,
                                                                   float* __restrict__ ws) {
    const float alpha = expf(__ldg(log_alpha));
    float s[2] = {0.f, 0.f};
    for (int64_t b = (int64_t)blockIdx.x * kThreads + threadIdx.x; b < B; b += (int64_t)gridDim.x * kThreads) {
        const float nq = fminf(q1t[b], q2t[b]) - alpha * logp[b];
        const float tq = rew[b] + (1.0f - done[b]) * gamma * nq;
        target[b] = tq;
        const float d1 = q1[b] - tq, d2 = q2[b] - tq;
        s[0] = fmaf(d1, d1, s[0]);
        s[1] = fmaf(d2, d2, s[1]);
    }
    finalize<2>(s, ws, losses, 1.0f / (float)B);
}

__global__ void __launch_bounds__(kThreads) critic_loss_bwd_kernel(const float* __restrict__ q1, const float* __restrict__ q2,
                                                                   const float* __restrict__ target, const float* __restrict__ g1,
                                                                   const float* __restrict__ g2, int64_t B, float* __restrict__ dq1,
                                                                   float* __restrict__ dq2) {
    const float c1 = g1 ? 2.0f * __ldg(g1) / (float)B : 0.f, c2 = g2 ? 2.0f * __ldg(g2) / (float)B : 0.f;
    for (int64_t b = (int64_t)blockIdx.x * kThreads + threadIdx.x; b < B; b += (int64_t)gridDim.x * kThreads) {
        const float tq = target[b];
        if (dq1) dq1[b] = c1 * (q1[b] - tq);
        if (dq2) dq2[b] = c2 * (q2[b] - tq);
    }
}

// out[0] = actor loss, out[1] = mean(logp) (for d/d log_alpha)
__global__ void __launch_bounds__(kThreads) actor_loss_fwd_kernel(const float* __restrict__ logp, const float* __restrict__ q1,
                                                                  const float* __restrict__ q2, const float* __restrict__ log_alpha, int64_t B,
                                                                  float* __restrict__ out, float* __restrict__ ws) {
    const float alpha = expf(__ldg(log_alpha));
    float s[2] = {0.f, 0.f};
    for (int64_t b = (int64_t)blockIdx.x * kThreads + threadIdx.x; b < B; b += (int64_t)gridDim.x * kThreads) {
        const float lp = logp[b];
        s[0] += alpha * lp - fminf(q1[b], q2[b]);
        s[1] += lp;
    }
    finalize<2>(s, ws, out, 1.0f / (float)B);
}

// torch.min(a, b) splits the gradient evenly on ties
__global__ void __launch_bounds__(kThreads) actor_loss_bwd_kernel(const float* __restrict__ q1, const float* __restrict__ q2,
                                                                  const float* __restrict__ log_alpha, const float* __restrict__ g, int64_t B,
                                                                  float* __restrict__ dlogp, float* __restrict__ dq1, float* __restrict__ dq2) {
    const float c = __ldg(g) / (float)B, ca = c * expf(__ldg(log_alpha));
    for (int64_t b = (int64_t)blockIdx.x * kThreads + threadIdx.x; b < B; b += (int64_t)gridDim.x * kThreads) {
        const float a = q1[b], bq = q2[b];
        const float w1 = a < bq ? 1.0f : (a == bq ? 0.5f : 0.f);
        dlogp[b] = ca;
        dq1[b] = -c * w1;
        dq2[b] = -c * (1.0f - w1);
    }
}

// out[0] = entropy loss = -log_alpha * mean(logp + target_entropy), out[1] = mean(logp + target_entropy)
struct EntropyPost {
    float la;
    __device__ void operator()(float* o) const { o[-1] = -la * o[0]; }
};
__global__ void __launch_bounds__(kThreads) entropy_loss_fwd_kernel(const float* __restrict__ logp, const float* __restrict__ log_alpha,
                                                                    float target_entropy, int64_t B, float* __restrict__ out,
                                                                    float* __restrict__ ws) {
    float s[1] = {0.f};
    for (int64_t b = (int64_t)blockIdx.x * kThreads + threadIdx.x; b < B; b += (int64_t)gridDim.x * kThreads) s[0] += logp[b] + target_entropy;
    finalize<1>(s, ws, out + 1, 1.0f / (float)B, EntropyPost{__ldg(log_alpha)});
}

// ------------------------------------------------------------------------------------------------ discrete agent (NeuroFlow)
//   critic:   next_q = sum_a p'(a) (min(q1_t, q2_t)(a) - alpha log(p'(a) + 1e-8));  target = r + (1 - done) gamma next_q          (agent.py:549-566)
//             c_k = mean((q_k(s)[action] - target)^2)                                                                           (agent.py:568-576)
//   actor:    mean_b sum_a p(a) (alpha logp_b - min(q1, q2)(a))   [logp_b: log-probability of the SAMPLED action, broadcast]      (agent.py:605-612)
//   entropy:  log_alpha (H - target),  H = -mean_b sum_a p(a) log(p(a) + 1e-8)   (no gradient into p)                             (modules.py:820-838)
__global__ void __launch_bounds__(kThreads) dcritic_loss_fwd_kernel(const float* __restrict__ q1, const float* __restrict__ q2,
                                                                    const float* __restrict__ act, const float* __restrict__ nprobs,
                                                                    const float* __restrict__ q1t, const float* __restrict__ q2t,
                                                                    const float* __restrict__ rew, const float* __restrict__ done,
                                                                    const float* __restrict__ log_alpha, float gamma, int64_t B, int A,
                                                                    float* __restrict__ target, float* __restrict__ losses, float* __restrict__ ws) {
    const float alpha = expf(__ldg(log_alpha));
    float s[2] = {0.f, 0.f};
    for (int64_t b = (int64_t)blockIdx.x * kThreads + threadIdx.x; b < B; b += (int64_t)gridDim.x * kThreads) {
        float nq = 0.f;
        for (int a = 0; a < A; ++a) {
            const float p = nprobs[b * A + a];
            nq += p * (fminf(q1t[b * A + a], q2t[b * A + a]) - alpha * logf(p + 1e-8f));
        }
        const float tq = rew[b] + (1.0f - done[b]) * gamma * nq;
        target[b] = tq;
        const int ai = (int)(long long)act[b];
        const float d1 = q1[b * A + ai] - tq, d2 = q2[b * A + ai] - tq;
        s[0] = fmaf(d1, d1, s[0]);
        s[1] = fmaf(d2, d2, s[1]);
    }
    finalize<2>(s, ws, losses, 1.0f / (float)B);
}

__global__ void __launch_bounds__(kThreads) dcritic_loss_bwd_kernel(const float* __restrict__ q1, const float* __restrict__ q2,
                                                                    const float* __restrict__ act, const float* __restrict__ target,
                                                                    const float* __restrict__ g1, const float* __restrict__ g2, int64_t B, int A,
                                                                    float* __restrict__ dq1, float* __restrict__ dq2) {
    const float c1 = g1 ? 2.0f * __ldg(g1) / (float)B : 0.f, c2 = g2 ? 2.0f * __ldg(g2) / (float)B : 0.f;
    for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < B * A; e += (int64_t)gridDim.x * kThreads) {
        const int64_t b = e / A;
        const bool sel = (int)(e - b * A) == (int)(long long)act[b];
        const float tq = target[b];
        if (dq1) dq1[e] = sel ? c1 * (q1[e] - tq) : 0.f;
        if (dq2) dq2[e] = sel ? c2 * (q2[e] - tq) : 0.f;
    }
}

// out[0] = actor loss, out[1] = mean_b (logp_b sum_a p(a))  (for d/d log_alpha)
__global__ void __launch_bounds__(kThreads) dactor_loss_fwd_kernel(const float* __restrict__ probs, const float* __restrict__ logp,
                                                                   const float* __restrict__ q1, const float* __restrict__ q2,
                                                                   const float* __restrict__ log_alpha, int64_t B, int A, float* __restrict__ out,
                                                                   float* __restrict__ ws) {
    const float alpha = expf(__ldg(log_alpha));
    float s[2] = {0.f, 0.f};
    for (int64_t b = (int64_t)blockIdx.x * kThreads + threadIdx.x; b < B; b += (int64_t)gridDim.x * kThreads) {
        const float al = alpha * logp[b];
        float row = 0.f, ps = 0.f;
        for (int a = 0; a < A; ++a) {
            const float p = probs[b * A + a];
            row += p * (al - fminf(q1[b * A + a], q2[b * A + a]));
            ps += p;
        }
        s[0] += row;
        s[1] += logp[b] * ps;
    }
    finalize<2>(s, ws, out, 1.0f / (float)B);
}

// gradients to the probabilities and to the sampled action's log-probability (the critics are constants here)
__global__ void __launch_bounds__(kThreads) dactor_loss_bwd_kernel(const float* __restrict__ probs, const float* __restrict__ logp,
                                                                   const float* __restrict__ q1, const float* __restrict__ q2,
                                                                   const float* __restrict__ log_alpha, const float* __restrict__ g, int64_t B, int A,
                                                                   float* __restrict__ dprobs, float* __restrict__ dlogp) {
    const float c = __ldg(g) / (float)B, alpha = expf(__ldg(log_alpha));
    for (int64_t b = (int64_t)blockIdx.x * kThreads + threadIdx.x; b < B; b += (int64_t)gridDim.x * kThreads) {
        const float al = alpha * logp[b];
        float ps = 0.f;
        for (int a = 0; a < A; ++a) {
            dprobs[b * A + a] = c * (al - fminf(q1[b * A + a], q2[b * A + a]));
            ps += probs[b * A + a];
        }
        dlogp[b] = c * alpha * ps;
    }
}

// out[0] = entropy loss = log_alpha (H - target), out[1] = H - target
struct DEntropyPost {
    float la, target;
    __device__ void operator()(float* o) const {
        o[0] = -o[0] - target;          // the kernel summed +p log p
        o[-1] = la * o[0];
    }
};
__global__ void __launch_bounds__(kThreads) dentropy_loss_fwd_kernel(const float* __restrict__ probs, const float* __restrict__ log_alpha,
                                                                     float target_entropy, int64_t B, int A, float* __restrict__ out,
                                                                     float* __restrict__ ws) {
    float s[1] = {0.f};
    for (int64_t b = (int64_t)blockIdx.x * kThreads + threadIdx.x; b < B; b += (int64_t)gridDim.x * kThreads) {
        float row = 0.f;
        for (int a = 0; a < A; ++a) {
            const float p = probs[b * A + a];
            row += p * logf(p + 1e-8f);
        }
        s[0] += row;
    }
    finalize<1>(s, ws, out + 1, 1.0f / (float)B, DEntropyPost{__ldg(log_alpha), target_entropy});
}

// The tail of SACActorDiscrete.forward after the network (velora/models/sac/discrete.py:41-57, 80-101): softmax, Categorical(probs=p)
// [normalises p again], one multinomial draw [argmax(p_n / q), q ~ Exp(1) drawn by the caller with torch's generator: the fast path of
// torch.multinomial for one sample] and the log-probability of the draw [log(clamp(p_n, eps, 1 - eps))] - ~20 framework launches - in one.
constexpr float kProbEps = 1.1920928955078125e-07f;        // torch.finfo(torch.float32).eps (probs_to_logits' clamp)
constexpr int kMaxActions = 64;
__global__ void __launch_bounds__(kThreads) categorical_head_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ q, int64_t B,
                                                                        int A, float* __restrict__ probs, long long* __restrict__ actions,
                                                                        float* __restrict__ logp) {
    for (int64_t b = (int64_t)blockIdx.x * kThreads + threadIdx.x; b < B; b += (int64_t)gridDim.x * kThreads) {
        const float* x = logits + b * A;
        float m = x[0];
        for (int a = 1; a < A; ++a) m = fmaxf(m, x[a]);
        float s = 0.f;
        for (int a = 0; a < A; ++a) s += expf(x[a] - m);
        float S = 0.f;
        for (int a = 0; a < A; ++a) {
            const float p = expf(x[a] - m) / s;
            probs[b * A + a] = p;
            S += p;
        }
        int best = 0;
        float bestv = -1.0f, pbest = 0.f;
        for (int a = 0; a < A; ++a) {
            const float pn = probs[b * A + a] / S;
            const float v = pn / q[b * A + a];
            if (v > bestv) { bestv = v; best = a; pbest = pn; }
        }
        actions[b] = best;
        logp[b] = logf(fminf(fmaxf(pbest, kProbEps), 1.0f - kProbEps));
    }
}

// d_probs [B,A] or NULL, d_logp [B] or NULL -> d_logits [B,A]
__global__ void __launch_bounds__(kThreads) categorical_head_bwd_kernel(const float* __restrict__ probs, const long long* __restrict__ actions,
                                                                        const float* __restrict__ d_probs, const float* __restrict__ d_logp,
                                                                        int64_t B, int A, float* __restrict__ d_logits) {
    for (int64_t b = (int64_t)blockIdx.x * kThreads + threadIdx.x; b < B; b += (int64_t)gridDim.x * kThreads) {
        const float* p = probs + b * A;
        float S = 0.f;
        for (int a = 0; a < A; ++a) S += p[a];
        const int act = (int)actions[b];
        // log-prob branch: through the clamp (gradient only inside [eps, 1 - eps]) and the normalisation p_n = p / S
        float g = 0.f;
        if (d_logp) {
            const float pn = p[act] / S;
            if (pn >= kProbEps && pn <= 1.0f - kProbEps) g = d_logp[b] / pn;
        }
        const float gn = g * (p[act] / S);                    // sum_k dpn_k p_n,k
        float dot = 0.f;
        float dp[kMaxActions];
        for (int a = 0; a < A; ++a) {
            float v = d_probs ? d_probs[b * A + a] : 0.f;
            v += ((a == act ? g : 0.f) - gn) / S;
            dp[a] = v;
            dot = fmaf(v, p[a], dot);
        }
        for (int a = 0; a < A; ++a) d_logits[b * A + a] = p[a] * (dp[a] - dot);
    }
}

static int grid_for(int64_t B) {
    int64_t blocks = (B + kThreads - 1) / kThreads;
    return (int)(blocks < kMaxBlocks ? (blocks < 1 ? 1 : blocks) : kMaxBlocks);
}

}  // namespace sacl
}  // namespace vb

using namespace vb;

extern "C" int64_t vb_sac_loss_workspace_bytes(void) { return (4 + sacl::kMaxBlocks * sacl::kMaxSums) * (int64_t)sizeof(float); }

extern "C" int vb_sac_critic_loss_fwd(const float* q1, const float* q2, const float* q1_target, const float* q2_target, const float* logp_next,
                                      const float* rewards, const float* dones, const float* log_alpha, float gamma, int64_t B, float* target,
                                      float* losses, void* workspace, void* stream) {
    VB_CHECK_ARG(B >= 1, VB_ERR_BAD_DIMS, "vb_sac_critic_loss_fwd: empty batch");
    VB_CHECK_ARG(q1 && q2 && q1_target && q2_target && logp_next && rewards && dones && log_alpha && target && losses && workspace, VB_ERR_NULL_POINTER,
                 "vb_sac_critic_loss_fwd: null pointer");
    sacl::critic_loss_fwd_kernel<<<sacl::grid_for(B), sacl::kThreads, 0, (cudaStream_t)stream>>>(q1, q2, q1_target, q2_target, logp_next, rewards, dones,
                                                                                                 log_alpha, gamma, B, target, losses,
                                                                                                 reinterpret_cast<float*>(workspace));
    VB_LAUNCH_CHECK();
    return 0;
}

extern "C" int vb_sac_critic_loss_bwd(const float* q1, const float* q2, const float* target, const float* d_loss1, const float* d_loss2, int64_t B,
                                      float* dq1, float* dq2, void* stream) {
    VB_CHECK_ARG(B >= 1, VB_ERR_BAD_DIMS, "vb_sac_critic_loss_bwd: empty batch");
    VB_CHECK_ARG(q1 && q2 && target, VB_ERR_NULL_POINTER, "vb_sac_critic_loss_bwd: null pointer");
    sacl::critic_loss_bwd_kernel<<<sacl::grid_for(B), sacl::kThreads, 0, (cudaStream_t)stream>>>(q1, q2, target, d_loss1, d_loss2, B, dq1, dq2);
    VB_LAUNCH_CHECK();
    return 0;
}

extern "C" int vb_sac_actor_loss_fwd(const float* logp, const float* q1, const float* q2, const float* log_alpha, int64_t B, float* out,
                                     void* workspace, void* stream) {
    VB_CHECK_ARG(B >= 1, VB_ERR_BAD_DIMS, "vb_sac_actor_loss_fwd: empty batch");
    VB_CHECK_ARG(logp && q1 && q2 && log_alpha && out && workspace, VB_ERR_NULL_POINTER, "vb_sac_actor_loss_fwd: null pointer");
    sacl::actor_loss_fwd_kernel<<<sacl::grid_for(B), sacl::kThreads, 0, (cudaStream_t)stream>>>(logp, q1, q2, log_alpha, B, out,
                                                                                                reinterpret_cast<float*>(workspace));
    VB_LAUNCH_CHECK();
    return 0;
}

extern "C" int vb_sac_actor_loss_bwd(const float* q1, const float* q2, const float* log_alpha, const float* d_loss, int64_t B, float* dlogp,
                                     float* dq1, float* dq2, void* stream) {
    VB_CHECK_ARG(B >= 1, VB_ERR_BAD_DIMS, "vb_sac_actor_loss_bwd: empty batch");
    VB_CHECK_ARG(q1 && q2 && log_alpha && d_loss && dlogp && dq1 && dq2, VB_ERR_NULL_POINTER, "vb_sac_actor_loss_bwd: null pointer");
    sacl::actor_loss_bwd_kernel<<<sacl::grid_for(B), sacl::kThreads, 0, (cudaStream_t)stream>>>(q1, q2, log_alpha, d_loss, B, dlogp, dq1, dq2);
    VB_LAUNCH_CHECK();
    return 0;
}

extern "C" int vb_sac_entropy_loss_fwd(const float* logp, const float* log_alpha, float target_entropy, int64_t B, float* out, void* workspace,
                                       void* stream) {
    VB_CHECK_ARG(B >= 1, VB_ERR_BAD_DIMS, "vb_sac_entropy_loss_fwd: empty batch");
    VB_CHECK_ARG(logp && log_alpha && out && workspace, VB_ERR_NULL_POINTER, "vb_sac_entropy_loss_fwd: null pointer");
    sacl::entropy_loss_fwd_kernel<<<sacl::grid_for(B), sacl::kThreads, 0, (cudaStream_t)stream>>>(logp, log_alpha, target_entropy, B, out,
                                                                                                  reinterpret_cast<float*>(workspace));
    VB_LAUNCH_CHECK();
    return 0;
}

// ---- discrete agent
extern "C" int vb_sacd_critic_loss_fwd(const float* q1, const float* q2, const float* actions, const float* next_probs, const float* q1_target,
                                       const float* q2_target, const float* rewards, const float* dones, const float* log_alpha, float gamma,
                                       int64_t B, int A, float* target, float* losses, void* workspace, void* stream) {
    VB_CHECK_ARG(B >= 1 && A >= 1, VB_ERR_BAD_DIMS, "vb_sacd_critic_loss_fwd: empty batch");
    VB_CHECK_ARG(q1 && q2 && actions && next_probs && q1_target && q2_target && rewards && dones && log_alpha && target && losses && workspace,
                 VB_ERR_NULL_POINTER, "vb_sacd_critic_loss_fwd: null pointer");
    sacl::dcritic_loss_fwd_kernel<<<sacl::grid_for(B), sacl::kThreads, 0, (cudaStream_t)stream>>>(q1, q2, actions, next_probs, q1_target, q2_target,
                                                                                                  rewards, dones, log_alpha, gamma, B, A, target,
                                                                                                  losses, reinterpret_cast<float*>(workspace));
    VB_LAUNCH_CHECK();
    return 0;
}

extern "C" int vb_sacd_critic_loss_bwd(const float* q1, const float* q2, const float* actions, const float* target, const float* d_loss1,
                                       const float* d_loss2, int64_t B, int A, float* dq1, float* dq2, void* stream) {
    VB_CHECK_ARG(B >= 1 && A >= 1, VB_ERR_BAD_DIMS, "vb_sacd_critic_loss_bwd: empty batch");
    VB_CHECK_ARG(q1 && q2 && actions && target, VB_ERR_NULL_POINTER, "vb_sacd_critic_loss_bwd: null pointer");
    sacl::dcritic_loss_bwd_kernel<<<sacl::grid_for(B * A), sacl::kThreads, 0, (cudaStream_t)stream>>>(q1, q2, actions, target, d_loss1, d_loss2, B, A,
                                                                                                      dq1, dq2);
    VB_LAUNCH_CHECK();
    return 0;
}

extern "C" int vb_sacd_actor_loss_fwd(const float* probs, const float* logp, const float* q1, const float* q2, const float* log_alpha, int64_t B,
                                      int A, float* out, void* workspace, void* stream) {
    VB_CHECK_ARG(B >= 1 && A >= 1, VB_ERR_BAD_DIMS, "vb_sacd_actor_loss_fwd: empty batch");
    VB_CHECK_ARG(probs && logp && q1 && q2 && log_alpha && out && workspace, VB_ERR_NULL_POINTER, "vb_sacd_actor_loss_fwd: null pointer");
    sacl::dactor_loss_fwd_kernel<<<sacl::grid_for(B), sacl::kThreads, 0, (cudaStream_t)stream>>>(probs, logp, q1, q2, log_alpha, B, A, out,
                                                                                                 reinterpret_cast<float*>(workspace));
    VB_LAUNCH_CHECK();
    return 0;
}

extern "C" int vb_sacd_actor_loss_bwd(const float* probs, const float* logp, const float* q1, const float* q2, const float* log_alpha,
                                      const float* d_loss, int64_t B, int A, float* dprobs, float* dlogp, void* stream) {
    VB_CHECK_ARG(B >= 1 && A >= 1, VB_ERR_BAD_DIMS, "vb_sacd_actor_loss_bwd: empty batch");
    VB_CHECK_ARG(probs && logp && q1 && q2 && log_alpha && d_loss && dprobs && dlogp, VB_ERR_NULL_POINTER, "vb_sacd_actor_loss_bwd: null pointer");
    sacl::dactor_loss_bwd_kernel<<<sacl::grid_for(B), sacl::kThreads, 0, (cudaStream_t)stream>>>(probs, logp, q1, q2, log_alpha, d_loss, B, A, dprobs,
                                                                                                 dlogp);
    VB_LAUNCH_CHECK();
    return 0;
}

extern "C" int vb_sacd_entropy_loss_fwd(const float* probs, const float* log_alpha, float target_entropy, int64_t B, int A, float* out,
                                        void* workspace, void* stream) {
    VB_CHECK_ARG(B >= 1 && A >= 1, VB_ERR_BAD_DIMS, "vb_sacd_entropy_loss_fwd: empty batch");
    VB_CHECK_ARG(probs && log_alpha && out && workspace, VB_ERR_NULL_POINTER, "vb_sacd_entropy_loss_fwd: null pointer");
    sacl::dentropy_loss_fwd_kernel<<<sacl::grid_for(B), sacl::kThreads, 0, (cudaStream_t)stream>>>(probs, log_alpha, target_entropy, B, A, out,
                                                                                                   reinterpret_cast<float*>(workspace));
    VB_LAUNCH_CHECK();
    return 0;
}

extern "C" int vb_categorical_head_fwd(const float* logits, const float* noise, int64_t B, int A, float* probs, int64_t* actions, float* log_prob,
                                       void* stream) {
    VB_CHECK_ARG(B >= 0 && A >= 1 && A <= sacl::kMaxActions, VB_ERR_BAD_DIMS, "vb_categorical_head_fwd: bad sizes B=%lld A=%d", (long long)B, A);
    if (B == 0) return 0;
    VB_CHECK_ARG(logits && noise && probs && actions && log_prob, VB_ERR_NULL_POINTER, "vb_categorical_head_fwd: null pointer");
    sacl::categorical_head_fwd_kernel<<<sacl::grid_for(B), sacl::kThreads, 0, (cudaStream_t)stream>>>(logits, noise, B, A, probs,
                                                                                                      reinterpret_cast<long long*>(actions), log_prob);
    VB_LAUNCH_CHECK();
    return 0;
}

extern "C" int vb_categorical_head_bwd(const float* probs, const int64_t* actions, const float* d_probs, const float* d_log_prob, int64_t B, int A,
                                       float* d_logits, void* stream) {
    VB_CHECK_ARG(B >= 0 && A >= 1 && A <= sacl::kMaxActions, VB_ERR_BAD_DIMS, "vb_categorical_head_bwd: bad sizes B=%lld A=%d", (long long)B, A);
    if (B == 0) return 0;
    VB_CHECK_ARG(probs && actions && d_logits, VB_ERR_NULL_POINTER, "vb_categorical_head_bwd: null pointer");
    sacl::categorical_head_bwd_kernel<<<sacl::grid_for(B), sacl::kThreads, 0, (cudaStream_t)stream>>>(probs, reinterpret_cast<const long long*>(actions),
                                                                                                      d_probs, d_log_prob, B, A, d_logits);
    VB_LAUNCH_CHECK();
    return 0;
}
