// ctc.cu — CTC forward-backward over the blank-interleaved label lattice + row-wise log-softmax.
//
// Replaces torch.nn.CTCLoss as called by the reference (NeMo/nemo/collections/asr/losses/ctc.py:45-81:
// blank = num_classes, zero_infinity=True, reduction='none') and the log_softmax at the end of
// ConvASRDecoder.forward (modules/conv_asr.py:490).  Gradient convention = ATen's (see include/clasr_b200.h).
//
// Structure: alpha and beta recursions run concurrently (grid = (B,2)), one thread per lattice state,
// previous time-step kept in shared memory (double-buffered), the per-step gather of log_probs[t, l'_s]
// software-prefetched kDepth steps ahead (it does not depend on the recursion).  log_probs are consumed
// in NeMo's batch-major [B,T,Vp] layout directly: a sample's rows are contiguous, so no transpose copy.
// The gradient kernel is fully parallel over (b,t): class occupancies are scattered into a shared-memory
// row with fp32 atomics in LINEAR space (each term alpha*beta/(y*P) is a posterior <= 1, so no max-shift
// is needed), then one coalesced vectorised pass writes the Vp-wide gradient row.
#include "common.cuh"

namespace clasr {

// alpha / beta are carried and stored in fp64 while all transcendental work is fp32 on differences (see
// LatticeWs in common.cuh for the rationale: |alpha| ~ 2e3 at the named sizes defeats an fp32 recursion).
struct CtcWs {
  double* alpha;  // [B,T,S]
  double* beta;   // [B,T,S]
  double* nll;    // [B] raw negative log-likelihood (inf when infeasible)
  int S;
};

inline size_t ctc_ws_bytes(int B, int T, int maxU) {
  size_t S = 2 * (size_t)maxU + 1;
  size_t bytes = 2 * (size_t)B * T * S * sizeof(double) + (size_t)B * sizeof(double);
  return (bytes + 255) / 256 * 256;
}
inline CtcWs ctc_ws_carve(void* ws, int B, int T, int maxU) {
  CtcWs w;
  w.S = 2 * maxU + 1;
  size_t n = (size_t)B * T * w.S;
  w.alpha = (double*)ws;
  w.beta = w.alpha + n;
  w.nll = w.beta + n;
  return w;
}

__device__ __forceinline__ double lse3_d(double a, double b, double c) {
  const double m = fmax(a, fmax(b, c));
  if (m == -INFINITY) return -INFINITY;
  const float s = expf((float)(a - m)) + expf((float)(b - m)) + expf((float)(c - m));
  return m + (double)logf(s);
}

constexpr int kCtcDepth = 4;  // prefetch distance (time steps) of the log_probs gather

// One step of the recursion for lattice state s (shared by the two code paths below).
__device__ __forceinline__ double ctc_step(const double* prev, int pad, int s, int Sb, bool backward, bool first,
                                           bool skip, float y) {
  if (first) {
    if (!backward) return (s <= 1) ? (double)y : -INFINITY;      // alpha_0(0), alpha_0(1)
    return (s >= Sb - 2) ? (double)y : -INFINITY;                // beta_{T-1}(S-1), beta_{T-1}(S-2)
  }
  if (!backward) return lse3_d(prev[pad + s], prev[pad + s - 1], skip ? prev[pad + s - 2] : -INFINITY) + (double)y;
  return lse3_d(prev[pad + s], prev[pad + s + 1], skip ? prev[pad + s + 2] : -INFINITY) + (double)y;
}

__global__ void __launch_bounds__(1024) ctc_lattice_kernel(const float* __restrict__ log_probs,
                                                           const int64_t* __restrict__ targets,
                                                           int64_t target_stride,
                                                           const int64_t* __restrict__ input_lens,
                                                           const int64_t* __restrict__ target_lens, int T, int Vp,
                                                           int blank, int zero_infinity, CtcWs w,
                                                           float* __restrict__ nll_out) {
  extern __shared__ double smd[];  // [2][S_b + 4]: 2 pad slots (= -inf) on each side for the s-1,s-2 / s+1,s+2 reads
  const int b = blockIdx.x;
  const bool backward = blockIdx.y == 1;
  const int Tb = (int)input_lens[b];
  const int Ub = (int)target_lens[b];
  const int Sb = 2 * Ub + 1;
  const int S = w.S;
  const float* __restrict__ lp = log_probs + (int64_t)b * T * Vp;
  double* __restrict__ out = (backward ? w.beta : w.alpha) + (int64_t)b * T * S;
  const int64_t* __restrict__ tg = targets + (int64_t)b * target_stride;
  const int pad = 2;
  const int W = Sb + 2 * pad;
  double* prev = smd;
  double* nxt = smd + W;
  for (int i = threadIdx.x; i < 2 * W; i += blockDim.x) smd[i] = -INFINITY;
  if (Tb <= 0) {
    if (threadIdx.x == 0 && !backward) {
      // torch: zero-length input with a non-empty target is infeasible
      const float v = (Ub == 0) ? 0.f : INFINITY;
      w.nll[b] = (double)v;
      nll_out[b] = (zero_infinity && v == INFINITY) ? 0.f : v;
    }
    return;
  }
  __syncthreads();

  const bool one_state = Sb <= (int)blockDim.x;
  // per-thread constants of the one-state path
  const int s1 = threadIdx.x;
  const bool active = one_state && s1 < Sb;
  int cls1 = blank;
  bool skip1 = false;  // forward: may come from s-2 ; backward: may go to s+2
  if (active && (s1 & 1)) {
    cls1 = (int)tg[s1 >> 1];
    if (!backward) skip1 = (s1 >= 3) && ((int)tg[(s1 >> 1) - 1] != cls1);
    else skip1 = (s1 + 2 < Sb) && ((int)tg[(s1 >> 1) + 1] != cls1);
  }
  float q[kCtcDepth];
#pragma unroll
  for (int dd = 0; dd < kCtcDepth; ++dd) {
    const int t = backward ? Tb - 1 - dd : dd;
    q[dd] = (active && dd < Tb) ? __ldg(lp + (int64_t)t * Vp + cls1) : 0.f;
  }

  for (int step0 = 0; step0 < Tb; step0 += kCtcDepth) {
#pragma unroll
    for (int dd = 0; dd < kCtcDepth; ++dd) {
      const int step = step0 + dd;
      if (step < Tb) {  // uniform across the block
        const int t = backward ? Tb - 1 - step : step;
        if (one_state) {
          const float y = q[dd];
          {  // refill this slot for step + kCtcDepth
            const int ns = step + kCtcDepth;
            const int nt = backward ? Tb - 1 - ns : ns;
            q[dd] = (active && ns < Tb) ? __ldg(lp + (int64_t)nt * Vp + cls1) : 0.f;
          }
          if (active) {
            const double v = ctc_step(prev, pad, s1, Sb, backward, step == 0, skip1, y);
            nxt[pad + s1] = v;
            out[(int64_t)t * S + s1] = v;
          }
        } else {
          // very long targets (2U+1 > 1024): several states per thread, no prefetch queue
          for (int s = threadIdx.x; s < Sb; s += blockDim.x) {
            int cls = blank;
            bool skip = false;
            if (s & 1) {
              cls = (int)tg[s >> 1];
              if (!backward) skip = (s >= 3) && ((int)tg[(s >> 1) - 1] != cls);
              else skip = (s + 2 < Sb) && ((int)tg[(s >> 1) + 1] != cls);
            }
            const float y = __ldg(lp + (int64_t)t * Vp + cls);
            const double v = ctc_step(prev, pad, s, Sb, backward, step == 0, skip, y);
            nxt[pad + s] = v;
            out[(int64_t)t * S + s] = v;
          }
        }
        __syncthreads();
        double* tmp = prev; prev = nxt; nxt = tmp;
      }
    }
  }
  if (!backward && threadIdx.x == 0) {
    double l = prev[pad + Sb - 1];
    if (Sb > 1) l = log_sum_exp_d(l, prev[pad + Sb - 2]);
    const double nll = -l;
    w.nll[b] = nll;
    nll_out[b] = (zero_infinity && nll == (double)INFINITY) ? 0.f : (float)nll;
  }
}

// grid = (T, B); one CTA per (t,b) row.  smem: occ[Vp]
__global__ void __launch_bounds__(256) ctc_grad_kernel(const float* __restrict__ log_probs,
                                                       const int64_t* __restrict__ targets, int64_t target_stride,
                                                       const int64_t* __restrict__ input_lens,
                                                       const int64_t* __restrict__ target_lens, int T, int Vp,
                                                       int blank, int zero_infinity,
                                                       const float* __restrict__ grad_out, float* __restrict__ grad,
                                                       CtcWs w) {
  extern __shared__ float occ[];  // [Vp]
  const int t = blockIdx.x, b = blockIdx.y;
  const int Tb = (int)input_lens[b];
  const int Ub = (int)target_lens[b];
  const int Sb = 2 * Ub + 1;
  const int S = w.S;
  const double nll = w.nll[b];
  float* __restrict__ g = grad + ((int64_t)b * T + t) * Vp;
  const bool infeasible = (nll == (double)INFINITY);
  if (t >= Tb || (infeasible && zero_infinity)) {
    for (int c = threadIdx.x; c < Vp; c += blockDim.x) g[c] = 0.f;
    return;
  }
  const float* __restrict__ lp = log_probs + ((int64_t)b * T + t) * Vp;
  for (int c = threadIdx.x; c < Vp; c += blockDim.x) occ[c] = 0.f;
  __syncthreads();
  const double* __restrict__ al = w.alpha + ((int64_t)b * T + t) * S;
  const double* __restrict__ be = w.beta + ((int64_t)b * T + t) * S;
  const int64_t* __restrict__ tg = targets + (int64_t)b * target_stride;
  for (int s = threadIdx.x; s < Sb; s += blockDim.x) {
    const int cls = (s & 1) ? (int)tg[s >> 1] : blank;
    // alpha_t(s) * beta_t(s) / (y_t(cls) * P(l|x)) : posterior mass of state s at time t, in [0,1]
    const float e = (float)(al[s] + be[s] + nll - (double)__ldg(lp + cls));
    if (e > -INFINITY) atomicAdd(occ + cls, expf(e));
  }
  __syncthreads();
  const float go = grad_out ? grad_out[b] : 1.f;
  for (int c = threadIdx.x; c < Vp; c += blockDim.x) {
    const float y = lp[c];
    g[c] = (expf(y) - occ[c]) * go;
  }
}

// ------------------------------------------------------------------------------------------------
// log_softmax rows (warp per row)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) log_softmax_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                              int64_t rows, int cols) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* __restrict__ xr = x + row * cols;
  float* __restrict__ yr = y + row * cols;
  float m = -INFINITY;
  for (int c = lane; c < cols; c += 32) m = fmaxf(m, xr[c]);
  m = warp_max(m);
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s += expf(xr[c] - m);
  s = warp_sum(s);
  const float lse = m + logf(s);
  for (int c = lane; c < cols; c += 32) yr[c] = xr[c] - lse;
}

__global__ void __launch_bounds__(256) log_softmax_bwd_kernel(const float* __restrict__ y,
                                                              const float* __restrict__ dy, float* __restrict__ dx,
                                                              int64_t rows, int cols) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* __restrict__ yr = y + row * cols;
  const float* __restrict__ gr = dy + row * cols;
  float* __restrict__ o = dx + row * cols;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s += gr[c];
  s = warp_sum(s);
  for (int c = lane; c < cols; c += 32) o[c] = gr[c] - expf(yr[c]) * s;
}

}  // namespace clasr

using namespace clasr;

extern "C" size_t clasr_ctc_workspace_bytes(int B, int T, int max_target_len) {
  if (B <= 0 || T <= 0 || max_target_len < 0) return 0;
  return ctc_ws_bytes(B, T, max_target_len);
}

static int check_ctc_args(const char* who, const void* lp, const void* targets, const void* il, const void* tl, int B,
                          int T, int Vp, int maxU, int blank, const void* ws, size_t ws_bytes) {
  CLASR_CHECK_ARG(lp && il && tl && ws, "%s: null pointer", who);
  CLASR_CHECK_ARG(targets || maxU == 0, "%s: null targets", who);
  CLASR_CHECK_ARG(B > 0 && T > 0 && Vp > 0 && maxU >= 0, "%s: bad dimension (B=%d T=%d Vp=%d U=%d)", who, B, T, Vp, maxU);
  CLASR_CHECK_ARG(blank >= 0 && blank < Vp, "%s: blank %d outside [0,%d)", who, blank, Vp);
  CLASR_CHECK_ARG(ws_bytes >= ctc_ws_bytes(B, T, maxU), "%s: workspace too small (%zu < %zu)", who, ws_bytes,
                  ctc_ws_bytes(B, T, maxU));
  return CLASR_STATUS_SUCCESS;
}

extern "C" int clasr_ctc_loss_fwd(const float* log_probs, const int64_t* targets, int64_t target_stride,
                                  const int64_t* input_lens, const int64_t* target_lens, int B, int T, int Vp,
                                  int max_target_len, int blank, int zero_infinity, int need_beta, float* nll,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_ctc_args("ctc_loss_fwd", log_probs, targets, input_lens, target_lens, B, T, Vp, max_target_len, blank,
                          workspace, workspace_bytes);
  if (rc) return rc;
  CLASR_CHECK_ARG(nll, "ctc_loss_fwd: null nll");
  CtcWs w = ctc_ws_carve(workspace, B, T, max_target_len);
  int threads = ((w.S + 31) / 32) * 32;
  if (threads > 1024) threads = 1024;
  size_t smem = (size_t)2 * (w.S + 4) * sizeof(double);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(ctc_lattice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    CLASR_CHECK_ARG(e == cudaSuccess, "ctc_loss_fwd: target too long for shared memory (%zu bytes)", smem);
  }
  prof_begin("ctc_lattice", (cudaStream_t)stream);
  ctc_lattice_kernel<<<dim3(B, need_beta ? 2 : 1), threads, smem, (cudaStream_t)stream>>>(
      log_probs, targets, target_stride, input_lens, target_lens, T, Vp, blank, zero_infinity, w, nll);
  prof_end("ctc_lattice", (cudaStream_t)stream);
  CLASR_CHECK_LAUNCH("ctc_lattice");
  return CLASR_STATUS_SUCCESS;
}

extern "C" int clasr_ctc_loss_bwd(const float* log_probs, const int64_t* targets, int64_t target_stride,
                                  const int64_t* input_lens, const int64_t* target_lens, int B, int T, int Vp,
                                  int max_target_len, int blank, int zero_infinity, const float* grad_out, float* grad,
                                  const void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_ctc_args("ctc_loss_bwd", log_probs, targets, input_lens, target_lens, B, T, Vp, max_target_len, blank,
                          workspace, workspace_bytes);
  if (rc) return rc;
  CLASR_CHECK_ARG(grad, "ctc_loss_bwd: null grad");
  CLASR_CHECK_ARG(B <= 65535, "ctc_loss_bwd: B > 65535 not supported");
  CtcWs w = ctc_ws_carve(const_cast<void*>(workspace), B, T, max_target_len);
  size_t smem = (size_t)Vp * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(ctc_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    CLASR_CHECK_ARG(e == cudaSuccess, "ctc_loss_bwd: vocabulary too large for shared memory (%zu bytes)", smem);
  }
  prof_begin("ctc_grad", (cudaStream_t)stream);
  ctc_grad_kernel<<<dim3(T, B), 256, smem, (cudaStream_t)stream>>>(log_probs, targets, target_stride, input_lens,
                                                                  target_lens, T, Vp, blank, zero_infinity, grad_out,
                                                                  grad, w);
  prof_end("ctc_grad", (cudaStream_t)stream);
  CLASR_CHECK_LAUNCH("ctc_grad");
  return CLASR_STATUS_SUCCESS;
}

extern "C" int clasr_log_softmax_fwd(const float* x, float* y, int64_t rows, int cols, void* stream) {
  CLASR_CHECK_ARG(x && y && rows >= 0 && cols > 0, "log_softmax_fwd: bad arguments");
  if (rows == 0) return CLASR_STATUS_SUCCESS;
  log_softmax_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, y, rows, cols);
  CLASR_CHECK_LAUNCH("log_softmax_fwd");
  return CLASR_STATUS_SUCCESS;
}

extern "C" int clasr_log_softmax_bwd(const float* y, const float* dy, float* dx, int64_t rows, int cols,
                                     void* stream) {
  CLASR_CHECK_ARG(y && dy && dx && rows >= 0 && cols > 0, "log_softmax_bwd: bad arguments");
  if (rows == 0) return CLASR_STATUS_SUCCESS;
  log_softmax_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(y, dy, dx, rows, cols);
  CLASR_CHECK_LAUNCH("log_softmax_bwd");
  return CLASR_STATUS_SUCCESS;
}
