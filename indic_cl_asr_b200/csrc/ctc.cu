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
#include <stdlib.h>
#include <string.h>

namespace clasr {

// alpha / beta are stored in the scaled linear domain (LatNum, common.cuh: fp32 mantissa + int32 exponent — |alpha| ~ 2e3
// in log space at the named sizes defeats an fp32 log-space recursion, which is what ATen's kernel runs).
struct CtcWs {
  LatNum* alpha;  // [B,T,S]
  LatNum* beta;   // [B,T,S]
  LatNum* prob;   // [B,T,S] y_t(l'_s) = exp(log_probs[t, l'_s]) split into mantissa / exponent (wavefront kernel's pre-pass)
  double* nll;    // [B] raw negative log-likelihood (inf when infeasible)
  int S;
};

inline size_t ctc_ws_bytes(int B, int T, int maxU) {
  size_t S = 2 * (size_t)maxU + 1;
  size_t bytes = 3 * (size_t)B * T * S * sizeof(LatNum) + (size_t)B * sizeof(double);
  return (bytes + 255) / 256 * 256;
}
inline CtcWs ctc_ws_carve(void* ws, int B, int T, int maxU) {
  CtcWs w;
  w.S = 2 * maxU + 1;
  size_t n = (size_t)B * T * w.S;
  w.alpha = (LatNum*)ws;
  w.beta = w.alpha + n;
  w.prob = w.beta + n;
  w.nll = (double*)(w.prob + n);
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

// Generic variant (block-synchronous, fp64 log space): any target length; also the A/B arm (CLASR_CTC_LATTICE=generic).
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
  LatNum* __restrict__ out = (backward ? w.beta : w.alpha) + (int64_t)b * T * S;
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
            out[(int64_t)t * S + s1] = lat_from_log(v);
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
            out[(int64_t)t * S + s] = lat_from_log(v);
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

// ------------------------------------------------------------------------------------------------
// Warp-shuffle wavefront (the default).  Same machinery as the transducer lattice (rnnt_loss.cu): thread j owns lattice
// state j of the alpha recursion AND state S-1-j of the beta recursion (two independent dependent chains per thread),
// values live in the scaled linear domain, the s-1 / s-2 neighbours of the previous time step come from the
// neighbouring LANES (two shuffles), warps are chained through a shared-memory seam (lanes 30 / 31 publish their
// column, one release per kCtcBlock steps), and the step body is branch-free.  A pre-pass gathers y_t(l'_s) for the
// whole sample and splits it into mantissa / exponent with all threads in parallel (the gather is the only
// uncoalesced access: one 4-byte load per lattice state and time step, as in ATen's kernel).
//   alpha_t(s) = (alpha_{t-1}(s) + alpha_{t-1}(s-1) + [skip] alpha_{t-1}(s-2)) y_t(l'_s)      (beta mirrored, ATen's
//   convention: beta_t includes y_t).
// ------------------------------------------------------------------------------------------------
constexpr int kCtcBlock = 8;   // steps per seam hand-shake = probabilities in flight per lane = normalisation period

__device__ __forceinline__ float ctc_pow2(int d) { return __int_as_float(max(d + 127, 0) << 23); }
__device__ __forceinline__ void ctc_normalise(float& m, int& e) {
  const int bits = __float_as_int(m);
  const int k = (bits >> 23) - 127;
  const bool pos = m > 0.f;
  m = pos ? __int_as_float(bits - (k << 23)) : m;
  e = pos ? e + k : kLatZeroExp;
}

template <bool kBeta>
__global__ void __launch_bounds__(512) ctc_lattice_shfl_kernel(const float* __restrict__ log_probs,
                                                                const int64_t* __restrict__ targets,
                                                                int64_t target_stride,
                                                                const int64_t* __restrict__ input_lens,
                                                                const int64_t* __restrict__ target_lens, int T, int Vp,
                                                                int blank, int zero_infinity, CtcWs w,
                                                                float* __restrict__ nll_out) {
  constexpr int kDirs = kBeta ? 2 : 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x;
  const int Tb = (int)input_lens[b];
  const int Ub = (int)target_lens[b];
  const int Sb = 2 * Ub + 1;
  const int S = w.S;
  int* prog = reinterpret_cast<int*>(smem_raw);               // [2][32] time steps published per warp and direction
  uint4* seam = reinterpret_cast<uint4*>(prog + 64);          // [2][warps][T] lanes 30 / 31 as {m30, e30, m31, e31}
  if (threadIdx.x < 64) prog[threadIdx.x] = 0;
  if (Tb <= 0) {
    if (threadIdx.x == 0) {
      const float v = (Ub == 0) ? 0.f : INFINITY;             // torch: zero-length input with a non-empty target is infeasible
      w.nll[b] = (double)v;
      nll_out[b] = (zero_infinity && v == INFINITY) ? 0.f : v;
    }
    return;
  }
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int nw = (Sb + 31) >> 5;
  const int nwc = (int)(blockDim.x >> 5);
  const int j = threadIdx.x;
  const bool col_ok = j < Sb;
  const int64_t* __restrict__ tg = targets + (int64_t)b * target_stride;
  const float* __restrict__ lp = log_probs + (int64_t)b * T * Vp;
  const int64_t sbase = (int64_t)b * T * S;

  // ---- per-direction constants of this thread's lattice state
  int st[2];          // lattice state
  bool skip[2];       // may take the two-state jump (different neighbouring labels)
#pragma unroll
  for (int d = 0; d < kDirs; ++d) {
    const int s = d ? Sb - 1 - j : j;
    st[d] = s;
    skip[d] = false;
    if (col_ok && (s & 1)) {
      const int cls = (int)tg[s >> 1];
      if (!d) skip[d] = (s >= 3) && ((int)tg[(s >> 1) - 1] != cls);
      else skip[d] = (s + 2 < Sb) && ((int)tg[(s >> 1) + 1] != cls);
    }
  }
  // ---- pre-pass: y_t(l'_s) for every (t, s) of this sample, split into mantissa / exponent (independent iterations:
  // full ILP; stores coalesced over s).  The thread that produces prob[t][s] is not the one that consumes it (beta
  // runs mirrored), hence the barrier.
  if (col_ok) {
    const int cls = (j & 1) ? (int)tg[j >> 1] : blank;
    LatNum* __restrict__ pr = w.prob + sbase + j;
    const float* __restrict__ src = lp + cls;
#pragma unroll 16
    for (int t = 0; t < Tb; ++t) {   // 16 independent gathers in flight per thread
      LatNum o;
      lat_prob_split(__ldg(src + (int64_t)t * Vp), o.m, o.e);
      pr[(int64_t)t * S] = o;
    }
  }
  __syncthreads();                                            // the only block-wide barrier before the wavefront
  if (warp < nw) {
    const unsigned t_lim = col_ok ? (unsigned)Tb : 0u;
    const bool has_left = warp > 0;
    const bool publishes = warp + 1 < nw;
    const LatNum* pin[2];
    LatNum* pout[2];
    int64_t dstep[2];
    const uint4* seam_in[2];
    uint4* seam_out[2];
    unsigned prog_prev[2], prog_mine[2];
#pragma unroll
    for (int d = 0; d < kDirs; ++d) {
      const int64_t i0 = sbase + (int64_t)(d ? Tb - 1 : 0) * S + st[d];      // step 0: t = 0 (alpha) / Tb-1 (beta)
      dstep[d] = d ? -(int64_t)S : (int64_t)S;
      pin[d] = w.prob + i0;
      pout[d] = (d ? w.beta : w.alpha) + i0;
      seam_in[d] = seam + (size_t)(d * nwc + (has_left ? warp - 1 : 0)) * T;
      seam_out[d] = seam + (size_t)(d * nwc + warp) * T;
      prog_prev[d] = (unsigned)__cvta_generic_to_shared(prog + d * 32 + (has_left ? warp - 1 : 0));
      prog_mine[d] = (unsigned)__cvta_generic_to_shared(prog + d * 32 + warp);
    }
    uint2 q[2][kCtcBlock];
    auto load_prob = [](uint2& r, const LatNum* p, bool ok) {
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p ld.global.v2.u32 {%0,%1}, [%3];\n\t}"
                   : "+r"(r.x), "+r"(r.y)
                   : "r"((int)ok), "l"(p));
    };
#pragma unroll
    for (int i = 0; i < kCtcBlock; ++i) {
      const bool ok = (unsigned)i < t_lim;
#pragma unroll
      for (int d = 0; d < kDirs; ++d) {
        q[d][i] = make_uint2(0u, (unsigned)kLatZeroExp);
        load_prob(q[d][i], pin[d] + dstep[d] * i, ok);
      }
    }
#pragma unroll
    for (int d = 0; d < kDirs; ++d) pin[d] += dstep[d] * kCtcBlock;
    // state 0 of each direction starts from 1: alpha_0(0) = y_0(blank), alpha_0(1) = (0 + 1) y_0(l_1), the rest 0
    float cm[2];
    int ce[2];
#pragma unroll
    for (int d = 0; d < kDirs; ++d) {
      cm[d] = j == 0 ? 1.f : 0.f;
      ce[d] = j == 0 ? 0 : kLatZeroExp;
    }
    int avail = 0;
    for (int s0 = 0; s0 < Tb; s0 += kCtcBlock) {
      if (has_left) {   // the left-hand warp must have finished every time index < s0 + kCtcBlock - 1 (read at step+1)
        const int need = min(s0 + kCtcBlock - 1, Tb - 1);
        if (avail < need) {
          int v = 0;
          for (uint32_t it = 0;; ++it) {
            if (lane == 0) {
              int v0, v1 = 0x7fffffff;
              asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v0) : "r"(prog_prev[0]) : "memory");
              if (kBeta) asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v1) : "r"(prog_prev[1]) : "memory");
              v = min(v0, v1);
            }
            v = __shfl_sync(0xffffffffu, v, 0);
            if (v >= need) break;
            if (it > (1u << 24)) __trap();
          }
          avail = v;
        }
      }
#pragma unroll
      for (int i = 0; i < kCtcBlock; ++i) {
        const int step = s0 + i;
        const bool valid = (unsigned)step < t_lim;
        const bool ld_ok = (unsigned)(step + kCtcBlock) < t_lim;
#pragma unroll
        for (int d = 0; d < kDirs; ++d) {
          const uint2 cp = q[d][i];
          load_prob(q[d][i], pin[d], ld_ok);
          pin[d] += dstep[d];
          const float pm = __uint_as_float(cp.x);
          const int pe = (int)cp.y;
          // previous time step's values of the two lower neighbours
          float m1 = __shfl_up_sync(0xffffffffu, cm[d], 1);
          int e1 = __shfl_up_sync(0xffffffffu, ce[d], 1);
          float m2 = __shfl_up_sync(0xffffffffu, cm[d], 2);
          int e2 = __shfl_up_sync(0xffffffffu, ce[d], 2);
          if (has_left) {   // warp-uniform; lanes 0 / 1 take them from the left-hand warp's lanes 30 / 31 (time step - 1)
            const uint4 sv = seam_in[d][max(step - 1, 0)];
            const bool any = step > 0;
            const float a30 = any ? __uint_as_float(sv.x) : 0.f, a31 = any ? __uint_as_float(sv.z) : 0.f;
            const int x30 = any ? (int)sv.y : kLatZeroExp, x31 = any ? (int)sv.w : kLatZeroExp;
            m1 = lane == 0 ? a31 : m1;
            e1 = lane == 0 ? x31 : e1;
            m2 = lane == 0 ? a30 : (lane == 1 ? a31 : m2);
            e2 = lane == 0 ? x30 : (lane == 1 ? x31 : e2);
          } else {
            m1 = lane == 0 ? 0.f : m1;
            e1 = lane == 0 ? kLatZeroExp : e1;
            m2 = lane < 2 ? 0.f : m2;
            e2 = lane < 2 ? kLatZeroExp : e2;
          }
          m2 = skip[d] ? m2 : 0.f;
          e2 = skip[d] ? e2 : kLatZeroExp;
          const int E = max(ce[d], max(e1, e2));
          float x = cm[d] * ctc_pow2(ce[d] - E);
          x = fmaf(m1, ctc_pow2(e1 - E), x);
          x = fmaf(m2, ctc_pow2(e2 - E), x);
          float nm = valid ? x * pm : 0.f;
          int ne = valid ? E + pe : kLatZeroExp;
          if (i == kCtcBlock - 1) ctc_normalise(nm, ne);   // growth < 3 x 1.42 per step: 2^17 at most in between
          ne = max(ne, 2 * kLatZeroExp);
          if (valid) {
            LatNum o;
            o.m = nm;
            o.e = ne;
            *pout[d] = o;
          }
          pout[d] += dstep[d];
          cm[d] = nm;
          ce[d] = ne;
          if (publishes) {   // warp-uniform: lanes 30 / 31 publish their (normalised) values of this time step
            float sm = nm;
            int se = ne;
            ctc_normalise(sm, se);
            if (lane >= 30 && step < Tb)   // one predicated 8-byte store per lane: {m30, e30} | {m31, e31}
              reinterpret_cast<uint2*>(seam_out[d] + step)[lane - 30] = make_uint2(__float_as_uint(sm), (unsigned)se);
          }
        }
      }
      if (publishes) {
        const int done = min(s0 + kCtcBlock, Tb);   // time indices < done are in the seam
        __syncwarp();   // lane 30's seam stores are ordered before lane 31's release
        if (lane == 31) {
          asm volatile("fence.acq_rel.cta;" ::: "memory");
          asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(prog_mine[0]), "r"(done) : "memory");
          if (kBeta) asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(prog_mine[1]), "r"(done) : "memory");
        }
      }
    }
  }
  __syncthreads();   // every alpha_{T-1}(s) of this sample is in global memory
  if (threadIdx.x == 0) {
    const LatNum* last = w.alpha + sbase + (int64_t)(Tb - 1) * S;
    double l = lat_log(last[Sb - 1]);
    if (Sb > 1) l = log_sum_exp_d(l, lat_log(last[Sb - 2]));
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
  const LatNum* __restrict__ al = w.alpha + ((int64_t)b * T + t) * S;
  const LatNum* __restrict__ be = w.beta + ((int64_t)b * T + t) * S;
  const int64_t* __restrict__ tg = targets + (int64_t)b * target_stride;
  for (int s = threadIdx.x; s < Sb; s += blockDim.x) {
    const int cls = (s & 1) ? (int)tg[s >> 1] : blank;
    // alpha_t(s) * beta_t(s) / (y_t(cls) * P(l|x)) : posterior mass of state s at time t, in [0,1]
    const float e = (float)(lat_log(al[s]) + lat_log(be[s]) + nll - (double)__ldg(lp + cls));
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
  cudaStream_t s = (cudaStream_t)stream;
  // wavefront kernel when a thread per lattice state (2U+1 <= 512: 128 registers each) and the [2][warps][T] seam fit; CLASR_CTC_LATTICE=generic: A/B arm
  const int warps = (w.S + 31) / 32;
  const size_t seam_smem = 64 * sizeof(int) + (size_t)2 * warps * T * sizeof(uint4);
  const char* sel = getenv("CLASR_CTC_LATTICE");
  if (w.S <= 512 && seam_smem <= 200 * 1024 && !(sel && !strcmp(sel, "generic"))) {
    auto kern = need_beta ? ctc_lattice_shfl_kernel<true> : ctc_lattice_shfl_kernel<false>;
    if (seam_smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seam_smem);
      CLASR_CHECK_ARG(e == cudaSuccess, "ctc_loss_fwd: seam does not fit in shared memory (%zu bytes)", seam_smem);
    }
    prof_begin("ctc_lattice", s);
    kern<<<B, warps * 32, seam_smem, s>>>(log_probs, targets, target_stride, input_lens, target_lens, T, Vp, blank,
                                          zero_infinity, w, nll);
    prof_end("ctc_lattice", s);
    CLASR_CHECK_LAUNCH("ctc_lattice");
    return CLASR_STATUS_SUCCESS;
  }
  int threads = ((w.S + 31) / 32) * 32;
  if (threads > 1024) threads = 1024;
  size_t smem = (size_t)2 * (w.S + 4) * sizeof(double);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(ctc_lattice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    CLASR_CHECK_ARG(e == cudaSuccess, "ctc_loss_fwd: target too long for shared memory (%zu bytes)", smem);
  }
  prof_begin("ctc_lattice", s);
  ctc_lattice_kernel<<<dim3(B, need_beta ? 2 : 1), threads, smem, s>>>(
      log_probs, targets, target_stride, input_lens, target_lens, T, Vp, blank, zero_infinity, w, nll);
  prof_end("ctc_lattice", s);
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
