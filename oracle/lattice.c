/* lattice.c — plain-C restatement of the reference's CPU transducer loss.  TEST INFRASTRUCTURE ONLY
 * (checker in tests/, and the timed "port" CPU baseline of bench.py; never on the product path).
 *
 * Follows /root/reference/NeMo/nemo/collections/asr/parts/numba/rnnt_loss/utils/cpu_utils/cpu_rnnt.py:
 *   CpuRNNT_metadata.setup_probs   :128-138   (gather blank / label log-probs)
 *   CPURNNT.compute_alphas         :246-276
 *   CPURNNT.compute_betas_and_grads:278-345   (gradient w.r.t. LOG-PROBS, FastEmit via log1p(lambda))
 *   CPURNNT.cost_and_grad_kernel   :214-244   (ll *= 1 + fastemit_lambda ; cost = -ll)
 *   CPURNNT.cost_and_grad          :347-382   (per-sample loop; here OpenMP over samples)
 * Input is log-softmaxed, batch-first [B,T,U1,Vp] exactly as rnnt_pytorch.py:411-437 feeds the CPU path.
 * The reference runs this as un-jitted Python (~0.5 ms per lattice cell, single thread); this C port is a
 * deliberately generous stand-in for "the reference's CPU path" when used as a timing baseline.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float log_sum_exp(float a, float b) {
  if (a == -INFINITY) return b;
  if (b == -INFINITY) return a;
  if (a > b) return log1pf(expf(b - a)) + a;
  return log1pf(expf(a - b)) + b;
}

/* returns 0 on success */
int oracle_rnnt_cpu(const float* log_probs, const int64_t* labels, const int64_t* act_lens,
                    const int64_t* label_lens, int B, int maxT, int maxU1, int Vp, int blank, float fastemit_lambda,
                    float* costs, float* grads /* may be NULL; else zero-filled here */) {
  int status = 0;
#pragma omp parallel for schedule(dynamic)
  for (int mb = 0; mb < B; ++mb) {
    const int T = (int)act_lens[mb];
    const int U = (int)label_lens[mb] + 1;
    const float* lp = log_probs + (size_t)mb * maxT * maxU1 * Vp;
    float* g = grads ? grads + (size_t)mb * maxT * maxU1 * Vp : NULL;
    const int64_t* lab = labels + (size_t)mb * (maxU1 - 1);
    float* alphas = (float*)malloc(sizeof(float) * T * U);
    float* betas = (float*)malloc(sizeof(float) * T * U);
    float* lp2 = (float*)malloc(sizeof(float) * T * U * 2);
    if (!alphas || !betas || !lp2) { status = 1; free(alphas); free(betas); free(lp2); continue; }
#define IDX3(t, u, v) (((size_t)(t) * maxU1 + (u)) * Vp + (v))
#define IDX(t, u) ((t) * U + (u))
    if (g) memset(g, 0, sizeof(float) * (size_t)maxT * maxU1 * Vp);
    for (int t = 0; t < T; ++t)
      for (int u = 0; u < U; ++u) {
        lp2[IDX(t, u) * 2] = lp[IDX3(t, u, blank)];
        if (u < U - 1) lp2[IDX(t, u) * 2 + 1] = lp[IDX3(t, u, lab[u])];
      }
    alphas[0] = 0.f;
    for (int t = 0; t < T; ++t)
      for (int u = 0; u < U; ++u) {
        if (u == 0 && t > 0) alphas[IDX(t, 0)] = alphas[IDX(t - 1, 0)] + lp2[IDX(t - 1, 0) * 2];
        if (t == 0 && u > 0) alphas[IDX(0, u)] = alphas[IDX(0, u - 1)] + lp2[IDX(0, u - 1) * 2 + 1];
        if (t > 0 && u > 0) {
          float no_emit = alphas[IDX(t - 1, u)] + lp2[IDX(t - 1, u) * 2];
          float emit = alphas[IDX(t, u - 1)] + lp2[IDX(t, u - 1) * 2 + 1];
          alphas[IDX(t, u)] = log_sum_exp(emit, no_emit);
        }
      }
    float ll_fwd = alphas[IDX(T - 1, U - 1)] + lp2[IDX(T - 1, U - 1) * 2];
    if (g) {
      betas[IDX(T - 1, U - 1)] = lp2[IDX(T - 1, U - 1) * 2];
      for (int t = T - 1; t >= 0; --t)
        for (int u = U - 1; u >= 0; --u) {
          if (u == U - 1 && t < T - 1) betas[IDX(t, U - 1)] = betas[IDX(t + 1, U - 1)] + lp2[IDX(t, U - 1) * 2];
          if (t == T - 1 && u < U - 1) betas[IDX(T - 1, u)] = betas[IDX(T - 1, u + 1)] + lp2[IDX(T - 1, u) * 2 + 1];
          if (t < T - 1 && u < U - 1) {
            float no_emit = betas[IDX(t + 1, u)] + lp2[IDX(t, u) * 2];
            float emit = betas[IDX(t, u + 1)] + lp2[IDX(t, u) * 2 + 1];
            betas[IDX(t, u)] = log_sum_exp(emit, no_emit);
          }
        }
      const float loglike = betas[0];
      for (int t = 0; t < T; ++t)
        for (int u = 0; u < U; ++u) {
          if (t < T - 1) {
            float gg = alphas[IDX(t, u)] + betas[IDX(t + 1, u)];
            g[IDX3(t, u, blank)] = -expf(lp2[IDX(t, u) * 2] + gg - loglike);
          }
          if (u < U - 1) {
            float gg = alphas[IDX(t, u)] + betas[IDX(t, u + 1)];
            g[IDX3(t, u, lab[u])] = -expf(log1pf(fastemit_lambda) + lp2[IDX(t, u) * 2 + 1] + gg - loglike);
          }
        }
      g[IDX3(T - 1, U - 1, blank)] = -expf(lp2[IDX(T - 1, U - 1) * 2] + alphas[IDX(T - 1, U - 1)] - loglike);
    }
    ll_fwd += ll_fwd * fastemit_lambda;
    costs[mb] = -ll_fwd;
    free(alphas); free(betas); free(lp2);
#undef IDX
#undef IDX3
  }
  return status;
}
