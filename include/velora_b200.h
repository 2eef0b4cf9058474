/* velora_b200 — C ABI of the B200-native Liquid-NN hot path (libvelora_b200.so).
 *
 * This is the drop-in boundary for the path named in BASELINE.json: everything the reference computes with ATen
 * call chains inside
 *     velora/models/lnn/sparse.py:71-81     SparseLinear.forward      (masked linear)
 *     velora/models/lnn/cell.py:113-169     NCPLiquidCell.forward     (five heads + closed-form gating)
 *     velora/models/lnn/ncp.py:129-178      LiquidNCPNetwork.forward  (inter -> Mish -> cell -> Mish -> motor)
 *     velora/models/lnn/ncp.py:299-328      NCPNetwork.forward        (three masked linears, Mish between)
 *     velora/buffer/replay.py:79-88         ReplayBuffer.sample       (six row gathers)
 * and the autograd backward of each, is one entry point here.  The reference has no FFI of its own (it is pure
 * Python over torch); INTEGRATION.md shows the ctypes binding a maintainer adds at those call sites.
 *
 * Conventions
 *   - extern "C"; plain pointers and sizes; no torch types.
 *   - every pointer is a DEVICE pointer owned by the caller (contiguous, row-major, float32 unless stated),
 *     except arrays-of-pointers arguments (`const float* const*`), which are HOST arrays of device pointers.
 *   - nothing is allocated, nothing is synchronised: each call only enqueues work on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream).  Scratch memory is passed in by the caller;
 *     its size comes from the matching *_workspace_bytes() query.
 *   - return value: 0 = ok; negative = argument error (VB_ERR_*); positive = cudaError_t.
 *     vb_last_error() returns a thread-local description of the last failure on the calling thread.
 *   - re-entrant: may be called concurrently from several host threads (autograd's backward runs on its own
 *     thread) and on several streams; the tensor-core and tile kernels keep no process-global device state (weights are
 *     staged per CTA in shared memory).  Exception: the VB_FLAG_FORCE_FFMA tier keeps its weights in a __constant__ bank and
 *     serialises its launches on one mutex + event per device.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with a cudaError_t.
 *
 * Packed parameter block ("packed", float32).  The seven SparseLinear layers of a LiquidNCPNetwork stay the
 * optimizer-visible truth in the host framework; the kernels read one pre-masked, transposed, padded copy
 * produced by vb_liquid_pack().  With NIp = ceil4(Ni), NCp = ceil4(Nc), Op = ceil4(O), Ip = ceil4(I), K = Ni+Nc,
 * head order {g, h, f (= f_head_to_g + f_head_to_h, only ever used summed: cell.py:130-133), proj}:
 *
 *   F section (forward orientation; also the layout of every gradient block "dpacked")
 *     win_t [I ][NIp]      W_in^T (*) mask          b_in [NIp]
 *     wh_t  [K ][4*NCp]    row i = weights of input i to (head, neuron)      b_h [4*NCp]
 *     wout  [O ][NCp]                                b_out[Op]
 *   D section (transposed copies used by the data-gradient of the generic kernels)
 *     win   [Ni][Ip]       wh [4*NCp][Kp] (Kp = ceil4(K))      wout_t [Nc][Op]
 *
 * NCPNetwork uses the same scheme with three layers: w1_t [I][NIp] b1 [NIp] | w2_t [Ni][NCp] b2 [NCp] |
 * w3 [O][NCp] b3 [Op]  and D section  w1 [Ni][Ip]  w2 [Nc][NIp]  w3_t [Nc][Op].
 */
#ifndef VELORA_B200_H
#define VELORA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VB_ABI_VERSION 5      /* 5: additive - vb_transpose_rows, vb_buffer_add_row, vb_sacd_*, vb_categorical_head_*; vb_ncp_bwd accepts dpacked = NULL;
                               *    4: standalone NCPLiquidCell entry points (vb_cell_fwd / vb_cell_bwd), absent layers in vb_liquid_pack;
                               *    3: the liquid forward saves gate factors for the backward (vb_liquid_saved_bytes, `saved` arguments);
                               *    2: layout flags, optimizer / policy-head / packed-gather / bf16-forward entry points added */

/* argument errors */
#define VB_ERR_NULL_POINTER (-1)
#define VB_ERR_BAD_DIMS (-2)
#define VB_ERR_WORKSPACE (-3)
#define VB_ERR_TOO_LARGE (-4)   /* network does not fit the on-chip budget of any compiled kernel */
#define VB_ERR_ALIGNMENT (-5)
#define VB_ERR_LAYOUT (-6)      /* a *_TIME_MAJOR flag on a kernel path that only reads batch-major tensors */

/* mask element types accepted by the pack/unpack entry points (wiring masks are int32, cell masks float32:
 * velora/wiring.py:132-149, velora/models/lnn/cell.py:109-111) */
#define VB_MASK_I32 0
#define VB_MASK_F32 1
#define VB_MASK_TRANSPOSED 2   /* OR-ed in: the [out, in] mask is stored column-major (a transposed view of a row-major
                                * [in, out] tensor, which is how the reference keeps `abs(mask.T)`) */

/* flags */
#define VB_FLAG_FORCE_GENERIC 1   /* use the generic tile kernels even where a specialised kernel exists (parity tests) */
#define VB_FLAG_FORCE_FFMA 2      /* skip the tensor-core (tcgen05) kernels, use the specialised FFMA kernels */
/* Storage order of the [B,T,F] sequence tensors of vb_liquid_fwd / vb_liquid_bwd.  Default: batch-major (row b*T + t).
 * With a flag set the tensor is stored time-major (row t*B + b), i.e. it is the stack of the per-step [B,F] batches the
 * reference's forward consumes one call at a time (ncp.py:129-178).  One thread owns one sample, so time-major rows
 * make every warp access contiguous (measured: -10% backward, -19% forward kernel time).  Only the tensor-core
 * kernels read time-major tensors: ask vb_liquid_time_major_ok() first; other paths answer VB_ERR_LAYOUT. */
#define VB_FLAG_X_TIME_MAJOR 16
#define VB_FLAG_Y_TIME_MAJOR 32    /* y (forward) / dy (backward) */
#define VB_FLAG_H_TIME_MAJOR 64    /* h_traj */
#define VB_FLAG_DX_TIME_MAJOR 128
#define VB_FLAG_DH_TIME_MAJOR 256  /* dh_traj */
#define VB_FLAG_LAYOUT_MASK (16 | 32 | 64 | 128 | 256)

/* vb_query() keys */
#define VB_Q_ABI_VERSION 0
#define VB_Q_BUILT_SM 1           /* 100 => compiled for sm_100a */
#define VB_Q_DEVICE_SMS 2
#define VB_Q_DEVICE_CC 3          /* major*10+minor of the current device */
#define VB_Q_MAX_SMEM_OPTIN 4

int vb_query(int what);
const char* vb_last_error(void);

/* ------------------------------------------------------------------------------------------------ liquid network */
int64_t vb_liquid_packed_floats(int I, int Ni, int Nc, int O);   /* F + D sections */
int64_t vb_liquid_grad_floats(int I, int Ni, int Nc, int O);     /* F section only */
/* 2 = tensor-core (tcgen05) kernels, 1 = specialised FFMA kernels, 0 = generic tile kernels only */
int vb_liquid_has_fast_path(int I, int Ni, int Nc, int O);
/* 1 when vb_liquid_fwd / vb_liquid_bwd called with these dims and `flags` accept the *_TIME_MAJOR layout flags */
int vb_liquid_time_major_ok(int I, int Ni, int Nc, int O, int flags);

/* Replaces the per-forward `weight * mask` of sparse.py:81 for all seven layers of ncp.py:79-102.
 * w_heads/b_heads/m_heads: host arrays of 5 device pointers in the order
 * {g_head, h_head, f_head_to_g, f_head_to_h, proj} (cell.py:61-69). */
int vb_liquid_pack(const float* w_in, const float* b_in, const void* m_in, int m_in_type,
                   const float* const* w_heads, const float* const* b_heads, const void* const* m_heads, int m_heads_type,
                   const float* w_out, const float* b_out, const void* m_out, int m_out_type,
                   int I, int Ni, int Nc, int O, float* packed, void* stream);

/* Activations saved for the backward ("save, don't recompute": the path is far from HBM-bound).  On the tensor-core path the forward
 * writes, per sample and step, the four gate factors of every command neuron and mish(c) - 160 B - into a caller-owned block of
 * vb_liquid_saved_bytes() bytes (16-byte aligned; 0 = this path saves nothing and `saved` is ignored); the backward reads it back instead
 * of re-running the five heads.  A backward called with saved = NULL re-runs the forward into its workspace first
 * (vb_liquid_bwd_workspace_bytes(..., have_saved = 0) includes that room). */
int64_t vb_liquid_saved_bytes(int I, int Ni, int Nc, int O, int64_t B, int T, int flags);
int64_t vb_liquid_bwd_workspace_bytes(int I, int Ni, int Nc, int O, int64_t B, int T, int have_saved);

/* T-fold LiquidNCPNetwork.forward (ncp.py:129-178) with the time loop on chip.
 *   x [B,T,I]; h0 [B,Nc] or NULL (= zeros, ncp.py:162-166); y [B,T,O]; h_traj [B,T,Nc] (h_traj[:,t] is the h'
 *   returned by step t).  T = 1 is exactly one reference forward call. */
int vb_liquid_fwd(const float* packed, const float* x, const float* h0, float* y, float* h_traj,
                  int64_t B, int T, int I, int Ni, int Nc, int O,
                  void* saved, int64_t saved_bytes, int flags, void* stream);      /* saved: NULL for inference */

/* Backward of the above (what autograd does over the reference's op graph); gates from `saved` or recomputed from x and h_traj.
 *   dy [B,T,O]; dh_traj [B,T,Nc] or NULL; dh_last [B,Nc] or NULL (extra gradient on h_traj[:,T-1]);
 *   dx [B,T,I] or NULL; dh0 [B,Nc] or NULL; dpacked: F-section layout, overwritten with the parameter gradients
 *   of the PACKED (pre-masked) weights — vb_liquid_unpack_grads() maps them back to the seven layers. */
int vb_liquid_bwd(const float* packed, const float* x, const float* h0, const float* h_traj,
                  const float* dy, const float* dh_traj, const float* dh_last,
                  const void* saved, int64_t saved_bytes,
                  float* dx, float* dh0, float* dpacked,
                  int64_t B, int T, int I, int Ni, int Nc, int O,
                  void* workspace, int64_t workspace_bytes, int flags, void* stream);

/* dW_layer = dW_packed (*) mask_layer (masked entries exactly 0: reference tests/models/test_lnn.py:570-576);
 * both f heads receive the same gradient.  accumulate != 0 adds into the outputs (torch .grad semantics). */
int vb_liquid_unpack_grads(const float* dpacked, const void* m_in, int m_in_type,
                           const void* const* m_heads, int m_heads_type, const void* m_out, int m_out_type,
                           float* g_w_in, float* g_b_in, float* const* g_w_heads, float* const* g_b_heads,
                           float* g_w_out, float* g_b_out,
                           int I, int Ni, int Nc, int O, int accumulate, void* stream);

/* The standalone NCPLiquidCell (reference velora/models/lnn/cell.py:147-169, one call = one step of the five heads, no inter / motor layer):
 *     z = [x, h];  gate = sigmoid(f_g(z) + f_h(z));  h' = tanh(g(z)) (1 - gate) + gate tanh(h(z));  y = proj(z) + h'.
 * `packed` is the liquid block for (I, Ni, Nc, O) = (In, In, Nc, Nc) written by vb_liquid_pack() with w_in = b_in = m_in = NULL and
 * w_out = b_out = m_out = NULL (absent layers pack as zeros); vb_liquid_unpack_grads() with NULL g_w_in / g_w_out maps dpacked back to the heads.
 *   x [B,In], h [B,Nc] or NULL (zeros) -> y [B,Nc], h_new [B,Nc];
 *   dy [B,Nc], dh_new [B,Nc] or NULL -> dx [B,In] or NULL, dh [B,Nc] or NULL, dpacked (F layout of those dims, overwritten). */
int64_t vb_cell_bwd_workspace_bytes(int In, int Nc);
int vb_cell_fwd(const float* packed, const float* x, const float* h, float* y, float* h_new, int64_t B, int In, int Nc, void* stream);
int vb_cell_bwd(const float* packed, const float* x, const float* h, const float* dy, const float* dh_new,
                float* dx, float* dh, float* dpacked, int64_t B, int In, int Nc,
                void* workspace, int64_t workspace_bytes, void* stream);

/* The 512-neuron class (BASELINE config 4: LiquidNCPNetwork(8,512,1)) with the heads on tcgen05 in bf16 (bf16 operands, fp32 accumulation,
 * fp32 hidden state; 2e-2 relative contract) - opt-in; the fp32 contract of vb_liquid_fwd / vb_liquid_bwd stays on the tile kernels.
 * Forward: x [B,T,I], y [B,T,O] (batch- or time-major: VB_FLAG_X/Y_TIME_MAJOR), h0 [B,Nc] or NULL, h_last [B,Nc] or NULL, h_traj [B,T,Nc]
 * or NULL (VB_FLAG_H_TIME_MAJOR).  With a `saved` block (vb_liquid_bf16_saved_bytes() bytes, 128-byte aligned; NULL for inference) the
 * forward also writes what the backward needs: the gate factors, mish(c), z~ and mish'(u) in bf16, 3720 B per sample-step.
 * Backward: dy [B,T,O] (VB_FLAG_Y_TIME_MAJOR), dh_last [B,Nc] or NULL (gradient on the last hidden state) -> dx [B,T,I] or NULL
 * (VB_FLAG_DX_TIME_MAJOR), dh0 [B,Nc] or NULL, dpacked (F layout, overwritten; vb_liquid_unpack_grads maps it to the layers).  The
 * recurrence runs on a hand-written tcgen05 kernel; three dense GEMMs over all sample-steps (dz_a, dW_in, dW_heads) are cuBLAS calls on
 * `stream` (the library links libcublas; this entry point cannot be captured in a CUDA graph).  Workspaces must be 128-byte aligned. */
int vb_liquid_bf16_ok(int I, int Ni, int Nc, int O);
int64_t vb_liquid_bf16_fwd_workspace_bytes(int I, int Ni, int Nc, int O, int64_t B, int T);
int64_t vb_liquid_bf16_saved_bytes(int I, int Ni, int Nc, int O, int64_t B, int T);
int vb_liquid_bf16_fwd(const float* packed, const float* x, const float* h0, float* y, float* h_last, float* h_traj,
                       void* saved, int64_t saved_bytes,
                       int64_t B, int T, int I, int Ni, int Nc, int O, void* workspace, int64_t workspace_bytes, int flags,
                       void* stream);
int64_t vb_liquid_bf16_bwd_workspace_bytes(int I, int Ni, int Nc, int O, int64_t B, int T);
int vb_liquid_bf16_bwd(const float* packed, const float* x, const float* dy, const float* dh_last, const void* saved, int64_t saved_bytes,
                       float* dx, float* dh0, float* dpacked, int64_t B, int T, int I, int Ni, int Nc, int O,
                       void* workspace, int64_t workspace_bytes, int flags, void* stream);

/* ------------------------------------------------------------------------------------------------ NCP network (critic) */
int64_t vb_ncp_packed_floats(int I, int Ni, int Nc, int O);
int64_t vb_ncp_grad_floats(int I, int Ni, int Nc, int O);
/* w/b/m: host arrays of 3 device pointers, layers ncp.0, ncp.2, ncp.4 (ncp.py:250-274) */
int vb_ncp_pack(const float* const* w, const float* const* b, const void* const* m, int m_type,
                int I, int Ni, int Nc, int O, float* packed, void* stream);
int64_t vb_ncp_bwd_workspace_bytes(int I, int Ni, int Nc, int O, int64_t B);
/* NCPNetwork.forward (ncp.py:299-328): x [B,I] -> y [B,O] */
int vb_ncp_fwd(const float* packed, const float* x, float* y, int64_t B, int I, int Ni, int Nc, int O,
               int flags, void* stream);
/* backward with recompute: dy [B,O] -> dx [B,I] (or NULL), dpacked (F layout, overwritten; NULL = dL/dx only: a critic evaluated inside the
 * actor loss, agent.py:236-245, whose parameter gradients the reference computes and then discards at the next zero_grad) */
int vb_ncp_bwd(const float* packed, const float* x, const float* dy, float* dx, float* dpacked,
               int64_t B, int I, int Ni, int Nc, int O,
               void* workspace, int64_t workspace_bytes, int flags, void* stream);
int vb_ncp_unpack_grads(const float* dpacked, const void* const* m, int m_type,
                        float* const* g_w, float* const* g_b,
                        int I, int Ni, int Nc, int O, int accumulate, void* stream);

/* ------------------------------------------------------------------------------------------------ replay buffer */
/* ReplayBuffer.sample's six gathers (replay.py:81-88) in one launch: dst[k][b,:] = src[k][idx[b],:].
 * src/dst: host arrays of n_arrays (<= 8) device pointers; row_floats[k] = floats per row of array k;
 * idx: int64 [B] (drawn by the caller with torch.randint so that indices stay bit-exact with the reference);
 * n_rows = valid rows in src (indices are checked against it: out-of-range rows are written as NaN and the call
 * still returns 0 — the host wrapper validates the range the way torch indexing would). */
int vb_gather_rows(const float* const* src, float* const* dst, const int* row_floats, int n_arrays,
                   const int64_t* idx, int64_t B, int64_t n_rows, void* stream);

/* Packed shadow table of the replay buffer: row r = the rows r of all arrays back to back at the given offsets
 * (floats), row_stride floats apart.  vb_pack_rows refreshes `n` rows - row0 .. row0+n-1, or the int64 device vector `rows`
 * when it is not NULL - from the arrays after BufferBase.add / add_multi (velora/buffer/base.py:73-143) wrote them;
 * vb_gather_packed is vb_gather_rows reading that table: one run of consecutive sectors per sampled transition instead of
 * one sector in each of six arrays. */
int vb_pack_rows(const float* const* src, const int* row_floats, const int* offsets, int n_arrays, float* packed,
                 int row_stride, const int64_t* rows, int64_t row0, int64_t n, int64_t capacity, void* stream);
/* [B,T,F] (batch-major) -> [T,B,F] (time-major) for to_time_major != 0, the reverse otherwise; F <= 64 floats per row.  Used by the host
 * side to hand batch-major sequence tensors to the tensor-core liquid kernels in the layout they read at full speed. */
int vb_transpose_rows(const float* src, float* dst, int64_t B, int T, int F, int to_time_major, void* stream);

/* BufferBase.add (velora/buffer/base.py:73-101): one transition into the (<= 8) arrays - row `row` of dst[k] gets the row_floats[k] values at
 * src[k] (device) or, where src[k] is NULL, the host value scalars[k] (reward, done) - and, when `shadow` is given, into the packed row. */
int vb_buffer_add_row(float* const* dst, const int* row_floats, const float* const* src, const float* scalars, int n_arrays,
                      int64_t row, int64_t capacity, float* shadow, int row_stride, const int* offsets, void* stream);
int vb_gather_packed(const float* packed, int row_stride, const int* offsets, float* const* dst, const int* row_floats,
                     int n_arrays, const int64_t* idx, int64_t B, int64_t n_rows, void* stream);

/* ------------------------------------------------------------------------------------------------ optimizer */
/* torch.optim.Adam step (no weight decay / amsgrad; modules.py:116-118, 409-415, 715-717 construct it with defaults) over
 * n_tensors fp32 tensors in one launch.  Host arrays of device pointers; steps[k] is the tensor's device-resident step
 * counter (torch's capturable layout), already incremented for this step; sizes in elements. */
int vb_adam_multi(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                  const float* const* steps, const int64_t* sizes, int n_tensors,
                  float lr, float beta1, float beta2, float eps, void* stream);
/* soft_update (velora/utils/torch.py:53-64): target = tau * source + (1 - tau) * target over all tensors in one launch */
int vb_polyak_multi(float* const* target, const float* const* source, const int64_t* sizes, int n_tensors, float tau, void* stream);

/* The tail of SACActor.forward (velora/models/sac/continuous.py:121-143) after the network: x [B, 2A] = (mean | log_std),
 * eps [B, A] drawn by the caller (Normal.rsample's noise), scale / bias [A].  Forward writes actions [B, A], log_prob [B] and the
 * two tensors the backward needs (tanh output a_norm, std).  Backward: d_actions [B, A] or NULL, d_log_prob [B] or NULL -> dx [B, 2A]. */
int vb_sac_head_fwd(const float* x, const float* eps, const float* scale, const float* bias, float log_std_min, float log_std_max,
                    int64_t B, int A, float* actions, float* log_prob, float* a_norm, float* std_out, void* stream);
int vb_sac_head_bwd(const float* x, const float* eps, const float* scale, float log_std_min, float log_std_max, int64_t B, int A,
                    const float* a_norm, const float* std_in, const float* d_actions, const float* d_log_prob, float* dx, void* stream);

/* The loss arithmetic around the networks in NeuroFlowCT's train step (velora/models/nf/agent.py:194-209, 236-245; modules.py:700-706),
 * one launch forward and one backward per loss; all tensors [B] (= [B,1] contiguous), scalars are device pointers; `workspace`:
 * vb_sac_loss_workspace_bytes() bytes, zero-initialised once, one per stream.
 *   critic:  target = r + (1 - done) * gamma * (min(q1_t, q2_t) - exp(log_alpha) * logp_next);  losses[k] = mean((q_k - target)^2)
 *   actor:   out[0] = mean(exp(log_alpha) * logp - min(q1, q2)), out[1] = mean(logp)
 *   entropy: out[0] = -log_alpha * mean(logp + target_entropy), out[1] = that mean */
int64_t vb_sac_loss_workspace_bytes(void);
int vb_sac_critic_loss_fwd(const float* q1, const float* q2, const float* q1_target, const float* q2_target, const float* logp_next,
                           const float* rewards, const float* dones, const float* log_alpha, float gamma, int64_t B, float* target,
                           float* losses, void* workspace, void* stream);
int vb_sac_critic_loss_bwd(const float* q1, const float* q2, const float* target, const float* d_loss1, const float* d_loss2, int64_t B,
                           float* dq1, float* dq2, void* stream);
int vb_sac_actor_loss_fwd(const float* logp, const float* q1, const float* q2, const float* log_alpha, int64_t B, float* out,
                          void* workspace, void* stream);
int vb_sac_actor_loss_bwd(const float* q1, const float* q2, const float* log_alpha, const float* d_loss, int64_t B, float* dlogp,
                          float* dq1, float* dq2, void* stream);
int vb_sac_entropy_loss_fwd(const float* logp, const float* log_alpha, float target_entropy, int64_t B, float* out, void* workspace,
                            void* stream);

/* The same for the discrete agent (NeuroFlow, agent.py:549-590, 592-636; modules.py:820-838): q / probs are [B,A] row-major, actions [B] hold the
 * action index as a float (the replay buffer's storage), logp [B] is the log-probability of the sampled action.
 *   critic:  target = r + (1 - done) gamma sum_a p'(a) (min(q1_t, q2_t)(a) - alpha log(p'(a) + 1e-8));  losses[k] = mean((q_k[b, action_b] - target)^2)
 *   actor:   out[0] = mean_b sum_a p(a) (alpha logp_b - min(q1, q2)(a)),  out[1] = mean_b (logp_b sum_a p(a));  gradients to p and logp
 *   entropy: out[0] = log_alpha (H - target_entropy), out[1] = H - target_entropy,  H = -mean_b sum_a p(a) log(p(a) + 1e-8) */
int vb_sacd_critic_loss_fwd(const float* q1, const float* q2, const float* actions, const float* next_probs, const float* q1_target,
                            const float* q2_target, const float* rewards, const float* dones, const float* log_alpha, float gamma,
                            int64_t B, int A, float* target, float* losses, void* workspace, void* stream);
int vb_sacd_critic_loss_bwd(const float* q1, const float* q2, const float* actions, const float* target, const float* d_loss1,
                            const float* d_loss2, int64_t B, int A, float* dq1, float* dq2, void* stream);
int vb_sacd_actor_loss_fwd(const float* probs, const float* logp, const float* q1, const float* q2, const float* log_alpha, int64_t B,
                           int A, float* out, void* workspace, void* stream);
int vb_sacd_actor_loss_bwd(const float* probs, const float* logp, const float* q1, const float* q2, const float* log_alpha,
                           const float* d_loss, int64_t B, int A, float* dprobs, float* dlogp, void* stream);
int vb_sacd_entropy_loss_fwd(const float* probs, const float* log_alpha, float target_entropy, int64_t B, int A, float* out,
                             void* workspace, void* stream);

/* The tail of SACActorDiscrete.forward after the network (velora/models/sac/discrete.py:41-57, 80-101): probs = softmax(logits) [B,A];
 * Categorical(probs) normalises once more; actions [B] (int64) = argmax(p_n / noise) with noise [B,A] ~ Exp(1) drawn by the caller from torch's
 * generator (what torch.multinomial does for one sample); log_prob [B] = log(clamp(p_n[action], eps, 1 - eps)).  A <= 64.
 * Backward: d_probs [B,A] or NULL, d_log_prob [B] or NULL -> d_logits [B,A]. */
int vb_categorical_head_fwd(const float* logits, const float* noise, int64_t B, int A, float* probs, int64_t* actions, float* log_prob,
                            void* stream);
int vb_categorical_head_bwd(const float* probs, const int64_t* actions, const float* d_probs, const float* d_log_prob, int64_t B, int A,
                            float* d_logits, void* stream);

/* ------------------------------------------------------------------------------------------------ data-parallel exchange */
/* One-shot all-reduce (sum, then * scale) of a small fp32 block over NVLink peer memory: the ONE exchange of the data-parallel path, the
 * parameter-gradient block of a backward.  peer_buffers: host array of `world` device pointers, entry r = rank r's
 * symmetric buffer of vb_allreduce_oneshot_buffer_bytes(cap_floats) bytes, zero-initialised once, mapped into this process (the host
 * framework allocates and exchanges them: torch.distributed._symmetric_memory); counter: a device uint32, zero-initialised, private to
 * this rank.  Every rank must make the same sequence of calls.  In place on `data`; one single-CTA kernel; graph-capturable (the call
 * number lives in `counter`).  A rank that waits ~2 s for a peer traps the kernel instead of spinning forever. */
int64_t vb_allreduce_oneshot_buffer_bytes(int64_t cap_floats);
int vb_allreduce_oneshot(float* data, int64_t n, void* const* peer_buffers, int rank, int world, int64_t cap_floats, float scale,
                         void* counter, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VELORA_B200_H */
