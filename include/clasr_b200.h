/*
 * clasr_b200.h — C ABI of libclasr_sm100.so: the B200 (sm_100a) implementation of the
 * RNNT-joint / transducer-loss / CTC-loss / EWC-MAS regulariser training hot path of
 * FrozenWolf-Cyber/Indic-CL-ASR (patched NeMo 1.23).
 *
 * Conventions (mirroring the reference's numba entry points):
 *   - plain C, raw DEVICE pointers + sizes, no torch types;
 *   - every compute entry returns an int status: 0 success, 1 invalid value, 2 CUDA error
 *     (0/1 mirror RNNTStatus, reference .../rnnt_loss/utils/global_constants.py:66-68);
 *     clasr_last_error() returns a static message for the calling thread;
 *   - every compute entry takes the CUDA stream to enqueue on (the reference uses torch's
 *     current stream, .../rnnt_loss/rnnt.py:173-176); `stream` is a cudaStream_t passed as void*;
 *   - the library NEVER allocates device memory and NEVER synchronises (the reference
 *     synchronises per loss call, gpu_rnnt.py:229 — not needed here); the caller owns all buffers;
 *   - all float buffers are fp32, all label/length buffers are int64 (certify_inputs,
 *     rnnt_pytorch.py:599-632), row-major contiguous unless a stride argument is given.
 *
 * "Reference paths" below are relative to /root/reference/NeMo/nemo/collections/asr/ unless they
 * start with cl_baseline / utils.py (repository root of the reference).
 */
#ifndef CLASR_B200_H_
#define CLASR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CLASR_API __attribute__((visibility("default")))
#else
#define CLASR_API
#endif

#define CLASR_STATUS_SUCCESS 0
#define CLASR_STATUS_INVALID_VALUE 1
#define CLASR_STATUS_CUDA_ERROR 2

#define CLASR_ACT_RELU 0
#define CLASR_ACT_SIGMOID 1
#define CLASR_ACT_TANH 2

/* Precision of the fused joint's tensor-core operands. */
#define CLASR_PREC_BF16 0   /* single bf16 pass (config 5, "bf16 joint GEMM") */
#define CLASR_PREC_BF16X3 1 /* hi/lo bf16 split, 3 MMAs per product: fp32-grade accuracy (config 2, "fp32") */
#define CLASR_PREC_FP16X3 2 /* the same split in fp16 (2^-22 operands instead of 2^-17, less tensor energy).  The fused
                             joint brings W_out, ReLU hidden values and dZ to O(1) with exact power-of-two scales it
                             undoes itself.  Plain GEMM / linear entries apply no scaling: their operands must fit
                             fp16's range. */
#define CLASR_PREC_FP16M8 3 /* fp16 hi.hi + TWO e4m3 correction terms (hi8.lo8 + lo8.hi8 as dense kind::f8f6f4 MMAs, K = 32
                             per instruction): 2 issue-equivalents per product instead of 3 in the two backward GEMMs of the
                             fused joint.  Every operand x is brought to max|x| in [2^13, 2^14) by an exact power of two;
                             x16 = fp16(x), hi8 = e4m3(x 2^-6), lo8 = e4m3((x - x16) 2^6), so that the three MMAs accumulate
                             at ONE common scale with no scale factors.  ~1e-5 of max|result| per GEMM (tools/
                             fp8_const_scale_study.py).  Plain GEMM entries: operands must already be scaled like that. */

CLASR_API int clasr_version(void);
CLASR_API const char* clasr_last_error(void);
/* Number of kernels this library has launched in this process (bench.py "gpu_launches"). */
CLASR_API int64_t clasr_launch_count(void);
/* Optional per-kernel timing for the roofline numbers: when on, the library brackets its major kernels with CUDA
 * events on the launch stream ("joint_fwd", "joint_bwd_dz", "gemm_dhid", "gemm_dw", "rnnt_lattice", "rnnt_lse",
 * "rnnt_grad", "ctc_lattice", "ctc_grad", "cl_penalty_grad", ...).  clasr_profile_ms synchronises on the events. */
CLASR_API void clasr_set_profiling(int on);
CLASR_API void clasr_profile_reset(void);
CLASR_API float clasr_profile_ms(const char* name, int* count);

/* ------------------------------------------------------------------------------------------
 * 1. Continual-learning regulariser sweeps over FLAT fp32 parameter buffers.
 *    The flat buffer is the concatenation, in named_parameters() order filtered by
 *    requires_grad (utils.py:273-321), of every trainable tensor, each padded to 4 floats.
 *
 *    `items` is a device array of clasr_sweep_item describing contiguous work chunks that never
 *    straddle a parameter tensor (so per-tensor statistics can be formed in the same pass).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int64_t start; /* first float of the chunk (multiple of 4)                 */
  int32_t len;   /* floats in the chunk (<= CLASR_SWEEP_CHUNK; the <4-float alignment pad is masked) */
  int32_t seg;   /* index of the parameter tensor the chunk belongs to        */
} clasr_sweep_item;

#define CLASR_SWEEP_CHUNK 8192

/* EWC penalty gradient — replaces get_penalty_grads (cl_baseline_ewc.py:69-81):
 *   g[i] = coef * F[i] * (theta[i] - theta_star[i]),  coef = 2 * e_lambda   (:74)
 *   seg_abs_sum[k] += sum_{i in tensor k} |g[i]|   (double; caller zero-fills)  (:76)
 * accumulate == 0: grad_out[i]  = g[i]   (reference: result dict then set_grads, utils.py:316-321)
 * accumulate != 0: grad_out[i] += g[i]   (fused "add to existing .grad" variant)
 * seg_abs_sum may be NULL (statistics skipped). */
CLASR_API int clasr_cl_penalty_grad(const float* theta, const float* theta_star, const float* fisher, float* grad_out,
                          const clasr_sweep_item* items, int64_t n_items, float coef, int accumulate,
                          double* seg_abs_sum, void* stream);

/* penalty_avg = (1/n_seg) * sum_k seg_abs_sum[k] / seg_numel[k]  (cl_baseline_ewc.py:76-81) -> out_avg[0] (device). */
CLASR_API int clasr_cl_penalty_avg(const double* seg_abs_sum, const int64_t* seg_numel, int64_t n_seg, float* out_avg,
                         void* stream);

/* Fisher accumulation — replaces cl_baseline_ewc.py:245-255:
 *   F[i] += w * grad[i]^2,  w = *weight_dev (the batch loss value, read on the device: no host sync). */
CLASR_API int clasr_cl_fisher_accum(float* fisher, const float* grad, int64_t n, const float* weight_dev, void* stream);

/* MAS importance accumulation — replaces cl_baseline_mas.py:267-270:  Omega[i] += |grad[i]|. */
CLASR_API int clasr_cl_mas_accum(float* omega, const float* grad, int64_t n, void* stream);

/* Finalise + merge — replaces cl_baseline_ewc.py:267-280 and cl_baseline_mas.py:283-287:
 *   src[i] *= (1.0f / count)                  (F /= total_ds ; Omega /= len(dataloader); as ATen's scalar div)
 *   dst[i]  = first ? src[i] : gamma * dst[i] + src[i]
 * dst may equal src (then only the scaling is applied). */
CLASR_API int clasr_cl_scale_merge(float* dst, float* src, int64_t n, float count, float gamma, int first, void* stream);

/* MAS penalty value and gradient — replaces penalty() (cl_baseline_mas.py:70-75) and the autograd
 * pass through it (:231-234):
 *   value[0] += sum_i Omega[i] * (theta[i]-theta_star[i])^2          (double; caller zero-fills)
 *   grad_accum[i] += grad_scale * 2 * Omega[i] * (theta[i]-theta_star[i])   (skipped if NULL)
 * grad_scale = mas_lambda * upstream gradient. */
CLASR_API int clasr_cl_penalty_value_grad(const float* theta, const float* theta_star, const float* omega, int64_t n,
                                float grad_scale, double* value, float* grad_accum, void* stream);

/* theta_star <- theta snapshot (get_params_clone, utils.py:284-293), one pass. */
CLASR_API int clasr_cl_snapshot(float* dst, const float* src, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------
 * 2. Transducer (RNNT) loss on materialised logits — replaces rnnt_loss_gpu + GPURNNT
 *    (parts/numba/rnnt_loss/rnnt.py:138-236, utils/cuda_utils/gpu_rnnt.py:125-231, kernels
 *    gpu_rnnt_kernel.py:73-407, reduce.py:121-248, rnnt_helper.py:106-116).
 *    logits [B,T,U1,Vp] raw joint outputs; labels [B,U1-1]; U1 = max(label_lens)+1.
 * ------------------------------------------------------------------------------------------ */

/* Bytes of caller-provided workspace (reference: B*(3*T*U1+2) floats, rnnt_helper.py:119-144;
 * ours: denominators + compact (blank,label) log-probs, alpha, beta in diagonal-major layout + 2 ll vectors). */
CLASR_API size_t clasr_rnnt_workspace_bytes(int B, int T, int U1);

/* Forward: costs[b] = -llForward[b] * (1 + fastemit_lambda).  Fills the workspace (needed by bwd). */
CLASR_API int clasr_rnnt_loss_fwd(const float* logits, const int64_t* labels, const int64_t* act_lens,
                        const int64_t* label_lens, int B, int T, int U1, int Vp, int blank, float fastemit_lambda,
                        float* costs, void* workspace, size_t workspace_bytes, void* stream);

/* Backward: grads[b,t,u,v] = grad_out[b] * clamp(dcost_b/dlogit)  (softmax-fused logits gradient,
 * gpu_rnnt_kernel.py:351-396; padded cells written as exact zeros, :343).  grad_out may be NULL (== 1).
 * Every element of grads is written (no zero-fill needed, unlike rnnt_pytorch.py:58 / gpu_rnnt.py:156). */
CLASR_API int clasr_rnnt_loss_bwd(const float* logits, const int64_t* labels, const int64_t* act_lens,
                        const int64_t* label_lens, int B, int T, int U1, int Vp, int blank, float fastemit_lambda,
                        float clamp, const float* grad_out, float* grads, const void* workspace,
                        size_t workspace_bytes, void* stream);

/* Debug/test access to the lattice held in the workspace: copies alpha, beta as dense [B,T,U1]
 * (padded cells = 0) and llForward/llBackward [B] (test parity with test_gpu_rnnt_kernel.py). */
CLASR_API int clasr_rnnt_export_lattice(const void* workspace, size_t workspace_bytes, const int64_t* act_lens,
                              const int64_t* label_lens, int B, int T, int U1, float* alphas, float* betas,
                              float* ll_fwd, float* ll_bwd, void* stream);

/* ------------------------------------------------------------------------------------------
 * 3. CTC loss — replaces torch.nn.CTCLoss as called by losses/ctc.py:45-81
 *    (blank = num_classes, zero_infinity, reduction='none'; NeMo-side reduction stays in Python).
 *    log_probs [B,T,Vp] BATCH-major (NeMo layout, the reference's transpose(1,0) is not needed);
 *    targets [B, target_stride].
 * ------------------------------------------------------------------------------------------ */
CLASR_API size_t clasr_ctc_workspace_bytes(int B, int T, int max_target_len);

CLASR_API int clasr_ctc_loss_fwd(const float* log_probs, const int64_t* targets, int64_t target_stride,
                       const int64_t* input_lens, const int64_t* target_lens, int B, int T, int Vp,
                       int max_target_len, int blank, int zero_infinity, int need_beta, float* nll,
                       void* workspace, size_t workspace_bytes, void* stream);

/* grad[b,t,c] = grad_out[b] * (exp(lp) - exp(logsumexp_{s:l'_s=c}(alpha+beta) + nll - lp))  — ATen's
 * convention (a true gradient only after log_softmax backward); rows t >= input_len and infeasible
 * samples under zero_infinity are exact zeros.  Every element of grad is written. */
CLASR_API int clasr_ctc_loss_bwd(const float* log_probs, const int64_t* targets, int64_t target_stride,
                       const int64_t* input_lens, const int64_t* target_lens, int B, int T, int Vp,
                       int max_target_len, int blank, int zero_infinity, const float* grad_out, float* grad,
                       const void* workspace, size_t workspace_bytes, void* stream);

/* Row-wise log_softmax over the last dimension (ConvASRDecoder.forward, modules/conv_asr.py:490) and its
 * backward fused with an incoming gradient: dx = dy - exp(y) * sum(dy). */
CLASR_API int clasr_log_softmax_fwd(const float* x, float* y, int64_t rows, int cols, void* stream);
CLASR_API int clasr_log_softmax_bwd(const float* y, const float* dy, float* dx, int64_t rows, int cols, void* stream);

/* ------------------------------------------------------------------------------------------
 * 3b. Tensor-core GEMM building block:  C[M,N] = A[M,K] . B[N,K]^T  (fp32 in/out, row-major, tcgen05 + TMEM,
 *     operands converted to bf16 or split bf16 hi/lo in `workspace`).  Replaces the cuBLAS calls behind
 *     nn.Linear backward for the joint's output layer (modules/rnnt.py:1634-1641 under autograd).
 * ------------------------------------------------------------------------------------------ */
CLASR_API size_t clasr_gemm_workspace_bytes(int M, int N, int K, int precision);
CLASR_API int clasr_gemm_nt(const float* A, const float* B, float* C, int M, int N, int K, int precision,
                            void* workspace, size_t workspace_bytes, void* stream);
/* General form: a_trans != 0 means A is given as [K,M] (M contiguous, consumed as an MN-major UMMA operand with no
 * transpose copy), likewise b_trans for B given as [K,N]; k_splits > 1 cuts K into independent slices that are
 * accumulated with fp32 atomics (C is zero-filled first) so that short-and-wide outputs still fill 148 SMs. */
CLASR_API int clasr_gemm_ex(const float* A, const float* B, float* C, int M, int N, int K, int a_trans, int b_trans,
                            int k_splits, int precision, void* workspace, size_t workspace_bytes, void* stream);

/* Linear layer on the tcgen05 GEMM: y[M,N] = x[M,K] . w[N,K]^T + bias[N]  (bias may be NULL), and its backward
 *   dx[M,K] = dy . w,  dw[N,K] = dy^T . x,  db[N] = column sums of dy   (each output may be NULL).
 * Replaces nn.Linear for the joint's enc / pred projections (modules/rnnt.py:1563-1585, built :1679-1680) and the
 * kernel-size-1 Conv1d of the CTC head (modules/conv_asr.py:444-446, 467-469).  The backward call reuses the
 * operand splits the forward call left in the same workspace. */
CLASR_API size_t clasr_linear_workspace_bytes(int M, int N, int K, int precision);
CLASR_API int clasr_linear_fwd(const float* x, const float* w, const float* bias, float* y, int M, int N, int K,
                               int precision, void* workspace, size_t workspace_bytes, void* stream);
CLASR_API int clasr_linear_bwd(const float* dy, float* dx, float* dw, float* db, int M, int N, int K, int precision,
                               void* workspace, size_t workspace_bytes, void* stream);

/* y[b, c, r] = x[b, r, c] for contiguous fp32 x [B, R, C]: the NeMo layouts [B, D, T'] / [B, D, U+1] of the encoder and
 * prediction-network outputs (modules/rnnt.py:1457-1459, conv_asr.py:467) to the row-major [B*T, D] operand of the linear
 * layers and back for their input gradients (32 x 32 tiles through shared memory, both sides coalesced). */
CLASR_API int clasr_transpose_last2(const float* x, float* y, int64_t B, int R, int C, void* stream);

/* Design tool (not on the product path): cycles per tcgen05.mma M=128 x N x K=16 issued back to back on an otherwise
 * idle SM.  pattern 0: SS, 1: TS (A from TMEM), 2: SS,SS,TS.  out_dev[0] = cycles, out_dev[1] = MMA count. */
CLASR_API int clasr_debug_mma_rate(int N, int pattern, int iters, long long* out_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * 4. Fused joint + transducer loss — replaces the fused branch of RNNTJoint.forward
 *    (modules/rnnt.py:1403-1561): joint_after_projection (:1587-1665) + RNNTLoss, without the padded
 *    [B,T,U1,Vp] logits tensor (a forward-only call stores no logits at all; see `stash` below).
 *      f [B,T,H] = enc projection, g [B,U1,H] = pred projection (fp32)
 *      z[b,t,u,:] = W_out . act(f[b,t,:] + g[b,u,:]) + b_out,  W_out [Vp,H] (nn.Linear layout)
 *    Pass 1 (tcgen05 GEMM, epilogue = online log-sum-exp + gather) fills the same lattice workspace
 *    as clasr_rnnt_loss_fwd and runs the alpha/beta wavefront; pass 2 forms the softmax-fused gradient (from the
 *    stashed logits, or by recomputing them tile-wise) and contracts it into d_f, d_g, dW_out, db_out.
 *    dropout_p > 0 applies the joint's Dropout (modules/rnnt.py:1699-1709: act -> Dropout(p) -> Linear) inside the
 *    kernels: a counter-based mask keyed by (dropout_seed, compact cell row, feature pair) that the recompute pass and
 *    the d_f / d_g reduction regenerate; kept activations are scaled by 1 / (1 - p).  The backward calls must be given
 *    the same (dropout_p, dropout_seed) as the forward call.  (Not bit-compatible with torch's Philox stream.)
 * ------------------------------------------------------------------------------------------ */
CLASR_API size_t clasr_joint_workspace_bytes(int B, int T, int U1, int H, int Vp, int precision);

CLASR_API int clasr_joint_rnnt_fwd(const float* f, const float* g, const float* w_out, const float* b_out,
                         const int64_t* labels, const int64_t* act_lens, const int64_t* label_lens, int B, int T,
                         int U1, int H, int Vp, int blank, int activation, int precision, float dropout_p,
                         uint64_t dropout_seed, float fastemit_lambda, float* costs,
                         float* sumsq /* [B,T,U1] sum_v z^2 for MAS, or NULL */, void* workspace,
                         size_t workspace_bytes, void* stash /* or NULL */, size_t stash_bytes, void* stream);

/* Optional `stash` (clasr_joint_stash_bytes): when given, pass 1 also keeps z (fp32, compact tile-row order) and the
 * bf16 hi/lo hidden activations, and a backward call given the SAME stash computes dZ with one streaming sweep over z
 * instead of a second joint GEMM (about 2 ms less per step at B32/T250/U100/V1024, for 4 (Vp + H) bytes per lattice
 * cell of HBM held between the two calls).  With stash == NULL the logits are never stored and the backward pass
 * recomputes them tile-wise — the memory-lean mode for lattices that would not fit. */
CLASR_API size_t clasr_joint_stash_bytes(int B, int T, int U1, int H, int Vp, int precision);

/* Backward.  `workspace` is the one the forward call filled (lattice, W split, tile table).  `scratch` holds the
 * GEMM operands of the backward pass (dZ and the hidden activations as bf16 hi/lo in compact tile-row order, dHid,
 * dW accumulator); it is caller-owned and reusable across steps.
 * Pass 2a produces the softmax-fused gradient dZ as bf16 hi/lo in `scratch` — from the stashed logits when `stash`
 * is given, else by recomputing the logits tile-wise on the tensor cores — and the two tcgen05 GEMMs
 * (dHid = dZ.W, dW = dZ^T.Hid) consume it. */
CLASR_API size_t clasr_joint_bwd_scratch_bytes(int B, int T, int U1, int H, int Vp, int precision);
CLASR_API int clasr_joint_rnnt_bwd(const float* f, const float* g, const float* w_out, const float* b_out,
                         const int64_t* labels, const int64_t* act_lens, const int64_t* label_lens, int B, int T,
                         int U1, int H, int Vp, int blank, int activation, int precision, float dropout_p,
                         uint64_t dropout_seed, float fastemit_lambda, float clamp, const float* grad_out /* [B] */,
                         float* d_f, float* d_g, float* d_w_out,
                         float* d_b_out, void* workspace, size_t workspace_bytes, void* scratch, size_t scratch_bytes,
                         void* stash /* the forward call's, or NULL */, size_t stash_bytes, void* stream);

/* Backward of the per-cell sum_v z^2 that clasr_joint_rnnt_fwd writes into `sumsq` — the joint half of the MAS
 * importance objective (cl_baseline_mas.py:258-265: mean over the stored sub-batch logits of sum_v z^2): dZ = 2 z *
 * grad_cells[b,t,u], then the same two GEMMs / reductions as the loss backward.  Replaces autograd through the
 * materialised `store_list` logits (modules/rnnt.py:1480-1496, 1649-1650).  Cells are those selected by act_lens /
 * label_lens of the forward call (pass the sub-batch maxima to include the reference's padded cells). */
CLASR_API int clasr_joint_sumsq_bwd(const float* f, const float* g, const float* w_out, const float* b_out,
                                    const int64_t* labels, const int64_t* act_lens, const int64_t* label_lens, int B,
                                    int T, int U1, int H, int Vp, int blank, int activation, int precision,
                                    float dropout_p, uint64_t dropout_seed, const float* grad_cells, float* d_f,
                                    float* d_g, float* d_w_out, float* d_b_out,
                                    void* workspace, size_t workspace_bytes, void* scratch, size_t scratch_bytes,
                                    void* stash /* the forward call's, or NULL */, size_t stash_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CLASR_B200_H_ */
