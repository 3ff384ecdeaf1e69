// rnnt_loss.cu — transducer loss on materialised logits (drop-in for the reference's numba path) and the
// alpha/beta anti-diagonal wavefront shared with the fused joint.
//
// Reference (relative to /root/reference/NeMo/nemo/collections/asr/parts/numba/rnnt_loss/):
//   reduce.py:121-248 (K1/K2 denominators), utils/cuda_utils/gpu_rnnt_kernel.py:73-172 (alpha),
//   :175-269 (beta), :272-407 (grad), utils/rnnt_helper.py:106-116 (costs), gpu_rnnt.py:125-231 (driver).
//
// What differs structurally from the reference (and why):
//   * one online max/sum pass over each logits row (the reference reads the tensor twice: max, then exp-sum)
//     that also gathers the two log-probs the lattice needs -> the lattice never touches [B,T,U,V] again;
//   * lattice data in a diagonal-major layout, staged through shared memory with cp.async, alpha and beta
//     recursions running concurrently in different CTAs (the reference keeps alpha/beta in global memory
//     and runs the two kernels back to back);
//   * gradient kernel writes every element (zeros for padding) and folds grad_output in, so neither the
//     zero-fill passes (rnnt_pytorch.py:58, gpu_rnnt.py:156) nor the in-place mul_ pass of backward
//     (rnnt_pytorch.py:87-91) exist.  Algorithmic traffic: 2 reads + 1 write of the logits tensor.
#include "common.cuh"
#include <stdlib.h>
#include <string.h>

namespace clasr {

constexpr int kRowWarps = 8;  // warps (rows) per CTA in the row-wise kernels

// ------------------------------------------------------------------------------------------------
// K1+K2 fused: denominators + gather.  One warp per (b,t,u) row; padded rows are skipped entirely.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRowWarps * 32) rnnt_lse_gather_kernel(
    const float* __restrict__ logits, const int64_t* __restrict__ labels, const int64_t* __restrict__ act_lens,
    const int64_t* __restrict__ label_lens, int B, int T, int U1, int Vp, int blank, LatticeWs w) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  const int64_t rows = (int64_t)B * T * U1;
  if (row >= rows) return;
  const int b = (int)(row / ((int64_t)T * U1));
  const int rem = (int)(row - (int64_t)b * T * U1);
  const int t = rem / U1, u = rem - t * U1;
  const int Tb = (int)act_lens[b], Ub1 = (int)label_lens[b] + 1;
  if (t >= Tb || u >= Ub1) return;

  const float* __restrict__ z = logits + row * Vp;
  float m = -INFINITY, s = 0.f;
  constexpr int kU = 8;
  for (int v0 = lane; v0 < Vp; v0 += 32 * kU) {
    float x[kU];
#pragma unroll
    for (int j = 0; j < kU; ++j) {
      int v = v0 + 32 * j;
      x[j] = v < Vp ? ld_stream1(z + v) : -INFINITY;
    }
    float cm = x[0];
#pragma unroll
    for (int j = 1; j < kU; ++j) cm = fmaxf(cm, x[j]);
    if (cm > m) {
      s *= expf(m - cm);  // m == -inf -> s == 0 stays 0 (expf(-inf) = 0)
      m = cm;
    }
    if (m > -INFINITY) {
#pragma unroll
      for (int j = 0; j < kU; ++j) s += expf(x[j] - m);
    }
  }
  const float M = warp_max(m);
  s = (m > -INFINITY) ? s * expf(m - M) : 0.f;
  const float S = warp_sum(s);
  const float denom = -M - logf(S);  // reduce.py:243-246: -max - log(sum)
  if (lane == 0) {
    const int64_t idx = ((int64_t)b * w.ND + t + u) * U1 + u;
    float zb = z[blank];
    float zl = -INFINITY;
    if (u < Ub1 - 1) zl = z[labels[(int64_t)b * (U1 - 1) + u]] + denom;
    w.denom[idx] = denom;
    w.lp[idx] = make_float2(zb + denom, zl);
    w.pp[idx] = lat_make_prob(zb + denom, zl);
  }
}

// ------------------------------------------------------------------------------------------------
// Vectorised variants of the two row kernels (used when logits / grads are 16-byte aligned).  A row of V+1 = 1025
// floats starts at a 4-byte-aligned address that is 16-byte aligned only every 4th row, so each row is walked as
// head (0-3 scalars up to the next 16-byte boundary) + aligned float4 body + tail: 512 bytes per warp instruction
// instead of 128, and 8 independent 16-byte loads in flight per lane (a whole 4 KB row per warp).  Warps are
// persistent over rows (grid-stride), exponentials run in the log2 domain (one FFMA + MUFU.EX2 per element).
// ------------------------------------------------------------------------------------------------
constexpr float kLog2eF = 1.4426950408889634f;
__device__ __forceinline__ float ex2f_(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

struct RowSplit {
  int head, n4, tail;        // scalars before the aligned body, float4 groups, scalars after
  const float4* body;
};
__device__ __forceinline__ RowSplit split_row(const float* z, int Vp) {
  RowSplit r;
  const int mis = (int)((reinterpret_cast<uintptr_t>(z) >> 2) & 3);
  r.head = min(Vp, (4 - mis) & 3);
  r.n4 = (Vp - r.head) >> 2;
  r.tail = Vp - r.head - 4 * r.n4;
  r.body = reinterpret_cast<const float4*>(z + r.head);
  return r;
}

__global__ void __launch_bounds__(kRowWarps * 32) rnnt_lse_gather_vec_kernel(
    const float* __restrict__ logits, const int64_t* __restrict__ labels, const int64_t* __restrict__ act_lens,
    const int64_t* __restrict__ label_lens, int B, int T, int U1, int Vp, int blank, LatticeWs w) {
  const int lane = threadIdx.x & 31;
  const int64_t rows = (int64_t)B * T * U1;
  for (int64_t row = (int64_t)blockIdx.x * kRowWarps + (threadIdx.x >> 5); row < rows;
       row += (int64_t)gridDim.x * kRowWarps) {
    const int b = (int)(row / ((int64_t)T * U1));
    const int rem = (int)(row - (int64_t)b * T * U1);
    const int t = rem / U1, u = rem - t * U1;
    const int Tb = (int)act_lens[b], Ub1 = (int)label_lens[b] + 1;
    if (t >= Tb || u >= Ub1) continue;
    const float* __restrict__ z = logits + row * Vp;
    const RowSplit rs = split_row(z, Vp);
    float m = -INFINITY, s = 0.f;
    auto absorb = [&](float cm, const float* x, int n) {   // n values with maximum cm
      if (cm > m) {
        s *= ex2f_((m - cm) * kLog2eF);   // m == -inf -> 0
        m = cm;
      }
      if (m > -INFINITY) {
        const float nm2 = -m * kLog2eF;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < n) s += ex2f_(fmaf(x[j], kLog2eF, nm2));
      }
    };
    // head / tail scalars: one per lane (head + tail <= 6)
    if (lane < rs.head + rs.tail) {
      const int v = lane < rs.head ? lane : rs.head + 4 * rs.n4 + (lane - rs.head);
      float x1[4] = {ld_stream1(z + v), 0.f, 0.f, 0.f};
      absorb(x1[0], x1, 1);
    }
    constexpr int kU = 8;
    for (int g0 = lane; g0 < rs.n4; g0 += 32 * kU) {
      float4 x[kU];
#pragma unroll
      for (int j = 0; j < kU; ++j) {
        const int gi = g0 + 32 * j;
        x[j] = gi < rs.n4 ? ld_stream(rs.body + gi) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      }
      float cm = -INFINITY;
#pragma unroll
      for (int j = 0; j < kU; ++j) cm = fmaxf(cm, fmaxf(fmaxf(x[j].x, x[j].y), fmaxf(x[j].z, x[j].w)));
      if (cm > m) {
        s *= ex2f_((m - cm) * kLog2eF);
        m = cm;
      }
      if (m > -INFINITY) {
        const float nm2 = -m * kLog2eF;
#pragma unroll
        for (int j = 0; j < kU; ++j) {   // -inf padding contributes 2^-inf = 0
          s += ex2f_(fmaf(x[j].x, kLog2eF, nm2)) + ex2f_(fmaf(x[j].y, kLog2eF, nm2));
          s += ex2f_(fmaf(x[j].z, kLog2eF, nm2)) + ex2f_(fmaf(x[j].w, kLog2eF, nm2));
        }
      }
    }
    const float M = warp_max(m);
    s = (m > -INFINITY) ? s * ex2f_((m - M) * kLog2eF) : 0.f;
    const float S = warp_sum(s);
    const float denom = -M - logf(S);  // reduce.py:243-246: -max - log(sum)
    if (lane == 0) {
      const int64_t idx = ((int64_t)b * w.ND + t + u) * U1 + u;
      const float zb = z[blank];
      float zl = -INFINITY;
      if (u < Ub1 - 1) zl = z[labels[(int64_t)b * (U1 - 1) + u]] + denom;
      w.denom[idx] = denom;
      w.lp[idx] = make_float2(zb + denom, zl);
      w.pp[idx] = lat_make_prob(zb + denom, zl);
    }
  }
}

__global__ void __launch_bounds__(kRowWarps * 32) rnnt_grad_vec_kernel(
    const float* __restrict__ logits, const int64_t* __restrict__ labels, const int64_t* __restrict__ act_lens,
    const int64_t* __restrict__ label_lens, int B, int T, int U1, int Vp, int blank, float fastemit_lambda,
    float clamp, const float* __restrict__ grad_out, float* __restrict__ grads, LatticeWs w) {
  const int lane = threadIdx.x & 31;
  const int64_t rows = (int64_t)B * T * U1;
  for (int64_t row = (int64_t)blockIdx.x * kRowWarps + (threadIdx.x >> 5); row < rows;
       row += (int64_t)gridDim.x * kRowWarps) {
    const int b = (int)(row / ((int64_t)T * U1));
    const int rem = (int)(row - (int64_t)b * T * U1);
    const int t = rem / U1, u = rem - t * U1;
    const int Tb = (int)act_lens[b], Ub1 = (int)label_lens[b] + 1;
    float* __restrict__ g = grads + row * Vp;
    const float* __restrict__ z = logits + row * Vp;
    const RowSplit rs = split_row(z, Vp);
    float4* __restrict__ gbody = reinterpret_cast<float4*>(g + rs.head);
    const int tail0 = rs.head + 4 * rs.n4;
    if (t >= Tb || u >= Ub1) {  // gpu_rnnt_kernel.py:343: padded cells keep zero gradient
      if (lane < rs.head) g[lane] = 0.f;
      else if (lane - rs.head < rs.tail) g[tail0 + lane - rs.head] = 0.f;
      for (int gi = lane; gi < rs.n4; gi += 32) st_stream(gbody + gi, make_float4(0.f, 0.f, 0.f, 0.f));
      continue;
    }
    const int64_t idx = ((int64_t)b * w.ND + t + u) * U1 + u;
    const double a = lat_log(w.alpha[idx]), bt = lat_log(w.beta[idx]), ll = w.ll_fwd[b];
    const float dn = w.denom[idx];
    const float2 lpair = w.lp[idx];
    const float go = grad_out ? grad_out[b] : 1.f;
    const bool has_label = u < Ub1 - 1;
    const int label = has_label ? (int)labels[(int64_t)b * (U1 - 1) + u] : -1;
    const double beta_t1 = (t < Tb - 1) ? lat_log(w.beta[idx + U1]) : 0.0;      // beta[t+1,u]
    const double beta_u1 = has_label ? lat_log(w.beta[idx + U1 + 1]) : 0.0;     // beta[t,u+1]
    const float base2 = ((float)(a + bt - ll) + dn) * kLog2eF;  // grad = exp(alpha + beta + logpk - ll), logpk = dn + z
    const bool fe = fastemit_lambda > 0.f && has_label;
    const float fe_base2 = fe ? ((float)(a + beta_u1 - ll + (double)lpair.y) + dn) * kLog2eF : -INFINITY;
    const float fe_coef = fe ? fastemit_lambda : 0.f;
    float blank_sub = 0.f;
    if (t == Tb - 1 && u == Ub1 - 1) blank_sub += expf((float)(a - ll + (double)lpair.x));
    if (t < Tb - 1) blank_sub += expf((float)(a + beta_t1 - ll + (double)lpair.x));
    const float label_sub =
        has_label ? expf(log1pf(fastemit_lambda) + (float)(a + beta_u1 - ll + (double)lpair.y)) : 0.f;
    auto one = [&](float x, int v) -> float {
      float gr = ex2f_(fmaf(x, kLog2eF, base2));
      if (fastemit_lambda > 0.f) gr = fmaf(fe_coef, ex2f_(fmaf(x, kLog2eF, fe_base2)), gr);
      if (v == blank) gr -= blank_sub;
      if (v == label) gr -= label_sub;
      if (clamp > 0.f) gr = fmaxf(fminf(gr, clamp), -clamp);
      return gr * go;
    };
    if (lane < rs.head) g[lane] = one(ld_stream1(z + lane), lane);
    else if (lane - rs.head < rs.tail) {
      const int v = tail0 + lane - rs.head;
      g[v] = one(ld_stream1(z + v), v);
    }
    constexpr int kU = 8;
    for (int g0 = lane; g0 < rs.n4; g0 += 32 * kU) {
      float4 x[kU];
#pragma unroll
      for (int j = 0; j < kU; ++j) {
        const int gi = g0 + 32 * j;
        x[j] = gi < rs.n4 ? ld_stream(rs.body + gi) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < kU; ++j) {
        const int gi = g0 + 32 * j;
        if (gi < rs.n4) {
          const int v = rs.head + 4 * gi;
          float4 o;
          o.x = one(x[j].x, v); o.y = one(x[j].y, v + 1); o.z = one(x[j].z, v + 2); o.w = one(x[j].w, v + 3);
          st_stream(gbody + gi, o);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K3/K4, generic variant (block-synchronous, fp64 log space): grid = (B, 2): blockIdx.y == 0 -> alpha, 1 -> beta.
// Used for U+1 > 1024 or when the seam buffers of the warp-shuffle kernel below would not fit in shared memory.
// Thread u owns lattice column u; diagonal n holds cells t = n - u.  The (blank,label) log-probs of
// `dch` diagonals at a time are staged into shared memory with cp.async, double-buffered.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__global__ void __launch_bounds__(1024) rnnt_lattice_kernel(LatticeWs w, const int64_t* __restrict__ act_lens,
                                                            const int64_t* __restrict__ label_lens, int T, int U1,
                                                            int dch, float fastemit_lambda,
                                                            float* __restrict__ costs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x;
  const bool backward = blockIdx.y == 1;
  const int Tb = (int)act_lens[b], Ub1 = (int)label_lens[b] + 1;
  if (Tb <= 0) {
    if (threadIdx.x == 0) {
      if (!backward) { w.ll_fwd[b] = 0.0; if (costs) costs[b] = 0.f; } else { w.ll_bwd[b] = 0.0; }
    }
    return;
  }
  double* vals = reinterpret_cast<double*>(smem_raw);            // [2][U1] previous/current diagonal (fp64)
  float2* stage = reinterpret_cast<float2*>(vals + 2 * U1);      // [2][dch*U1]
  const int nd = Tb + Ub1 - 1;     // diagonals of this utterance
  const int nrows = nd - 1;        // recursion steps
  const int64_t base = (int64_t)b * w.ND * U1;
  const float2* __restrict__ lp = w.lp + base;
  LatNum* __restrict__ out = (backward ? w.beta : w.alpha) + base;

  // step r (0..nrows-1) consumes lp row:  forward r  (produces diagonal r+1)
  //                                       backward nd-2-r (produces that same diagonal)
  auto issue_chunk = [&](int c) {
    const int r0 = c * dch;
    if (r0 >= nrows) return;
    const int cnt = min(dch, nrows - r0);
    const int row_lo = backward ? (nd - 2 - (r0 + cnt - 1)) : r0;
    const float2* src = lp + (int64_t)row_lo * U1;
    float2* dst = stage + (size_t)(c & 1) * dch * U1;
    for (int i = threadIdx.x; i < cnt * U1; i += blockDim.x) cp_async8(dst + i, src + i);
  };

  // initial diagonal
  if (threadIdx.x == 0) {
    if (!backward) {
      vals[0] = 0.0; out[0] = lat_from_log(0.0);
    } else {
      const int64_t li = (int64_t)(nd - 1) * U1 + (Ub1 - 1);
      const double v = (double)lp[li].x;
      vals[Ub1 - 1] = v;
      out[li] = lat_from_log(v);
    }
  }
  issue_chunk(0);
  cp_async_commit();
  int cur = 0;  // vals[cur] holds the previous diagonal
  const int nchunks = (nrows + dch - 1) / dch;
  for (int c = 0; c < nchunks; ++c) {
    issue_chunk(c + 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const int r0 = c * dch;
    const int cnt = min(dch, nrows - r0);
    const float2* st = stage + (size_t)(c & 1) * dch * U1;
    for (int k = 0; k < cnt; ++k) {
      const int r = r0 + k;
      const double* prev = vals + cur * U1;
      double* nxt = vals + (cur ^ 1) * U1;
      if (!backward) {
        const int n = r + 1;  // diagonal being produced; lp row n-1 staged at chunk-local row k
        const float2* lrow = st + (size_t)k * U1;
        for (int u = threadIdx.x; u < Ub1; u += blockDim.x) {
          const int t = n - u;
          if (t >= 0 && t < Tb) {
            double no_emit = -INFINITY, emit = -INFINITY;
            if (t > 0) no_emit = prev[u] + (double)lrow[u].x;          // alpha[t-1,u] + logp(blank | t-1,u)
            if (u > 0) emit = prev[u - 1] + (double)lrow[u - 1].y;     // alpha[t,u-1] + logp(label_{u-1} | t,u-1)
            const double v = log_sum_exp_d(emit, no_emit);
            nxt[u] = v;
            out[(int64_t)n * U1 + u] = lat_from_log(v);
          }
        }
      } else {
        const int n = nd - 2 - r;  // diagonal being produced; it consumes its own lp row n
        const int row_lo = nd - 2 - (r0 + cnt - 1);
        const float2* lrow = st + (size_t)(n - row_lo) * U1;
        for (int u = threadIdx.x; u < Ub1; u += blockDim.x) {
          const int t = n - u;
          if (t >= 0 && t < Tb) {
            double no_emit = -INFINITY, emit = -INFINITY;
            const float2 l = lrow[u];
            if (t < Tb - 1) no_emit = prev[u] + (double)l.x;           // beta[t+1,u] + logp(blank | t,u)
            if (u < Ub1 - 1) emit = prev[u + 1] + (double)l.y;         // beta[t,u+1] + logp(label_u | t,u)
            const double v = log_sum_exp_d(emit, no_emit);
            nxt[u] = v;
            out[(int64_t)n * U1 + u] = lat_from_log(v);
          }
        }
      }
      cur ^= 1;
      __syncthreads();
    }
  }
  cp_async_wait<0>();
  if (threadIdx.x == 0) {
    if (!backward) {
      // gpu_rnnt_kernel.py:167-172: ll = alpha[T-1,U-1] + logp(blank | T-1,U-1)
      const double ll = vals[cur * U1 + (Ub1 - 1)] + (double)lp[(int64_t)(nd - 1) * U1 + (Ub1 - 1)].x;
      w.ll_fwd[b] = ll;
      if (costs) costs[b] = (float)(-ll * (1.0 + (double)fastemit_lambda));  // rnnt_helper.py:106-116
    } else {
      w.ll_bwd[b] = vals[cur * U1 + 0];  // beta[0,0]
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K3/K4, warp-shuffle wavefront (the default).  Reference: gpu_rnnt_kernel.py:73-172 (alpha), :175-269 (beta) — there
// one block-wide barrier and a global-memory round trip per anti-diagonal.  Here:
//   * thread j owns lattice column j (alpha: u = j; beta runs the mirrored lattice, u = U_b - j) and walks it in time;
//     inside a warp lane l is l steps behind lane l-1 (the anti-diagonal skew), so the value a cell needs from its
//     left neighbour is what the neighbouring LANE produced one step earlier: ONE __shfl_up per step, no barrier;
//   * warps are chained through a shared-memory seam: lane 31 of warp w publishes its column (indexed by time step)
//     with a release store of a progress counter every kSeamPublish steps, lane 0 of warp w+1 acquires it.  Warp w+1
//     therefore runs ~40 steps behind warp w and the block never synchronises after start-up;
//   * arithmetic in a SCALED LINEAR domain instead of log space: a value is m * 2^e (m fp32, e int32), the two incoming
//     terms are aligned with an exact power of two, added, and re-normalised with integer operations on the exponent
//     field.  The dependent chain per step is shuffle + ~8 integer / fp32 operations instead of fp64 adds around
//     expf / log1pf, and the relative rounding error per step is 2^-24 wherever alpha sits (in log space the ABSOLUTE
//     error of an fp32 alpha ~ 2e3 would be 2.4e-4, which is why the generic kernel below carries fp64);
//   * the transition probabilities exp(logp) are split into mantissa and integer exponent off the dependent chain
//     (compensated product with log2(e), MUFU.EX2 on the fraction), prefetched kLatPrefetch steps ahead;
//   * results leave as fp64 natural logs in the same diagonal-major workspace: a warp's 32 stores of one step are
//     256 contiguous bytes.
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ float lat_pow2(int d) {   // 2^d for d <= 0; 0 below 2^-126
  return __int_as_float(max(d + 127, 0) << 23);
}
// x * 2^E = ma 2^ea + mb 2^eb, E = max(ea, eb): no normalisation (the caller does that every kLatPrefetch steps)
__device__ __forceinline__ void lat_add_lazy(float ma, int ea, float mb, int eb, float& x, int& E) {
  E = max(ea, eb);
  x = fmaf(ma, lat_pow2(ea - E), mb * lat_pow2(eb - E));
}
// bring m into [1, 2) (0 and NaN keep their mantissa; 0 gets the canonical exponent)
__device__ __forceinline__ void lat_normalise(float& m, int& e) {
  const int bits = __float_as_int(m);
  const int k = (bits >> 23) - 127;
  const bool pos = m > 0.f;
  m = pos ? __int_as_float(bits - (k << 23)) : m;
  e = pos ? e + k : kLatZeroExp;
}

// One thread runs column j of the alpha lattice AND column j of the mirrored beta lattice: two independent dependent
// chains in one instruction stream (a lone warp per scheduler issues a dependent instruction only every ~5 cycles, so
// the second chain is nearly free), and the step body is branch-free apart from the (rare, warp-uniform) seam wait.
// kPF: cells in flight per lane and direction (4 registers each), also the normalisation period of the chains.
template <int kPF, bool kHasLeft>
__device__ __forceinline__ void rnnt_lattice_run(const LatticeWs& w, int b, int Tb, int Ub1, int T, int U1, int warp,
                                                 int lane, int nw, int* prog, LatNum* seam, float fastemit_lambda,
                                                 float* __restrict__ costs) {
  const int j = warp * 32 + lane;                           // column in processing order
  const bool col_ok = j < Ub1;
  const unsigned t_lim = col_ok ? (unsigned)Tb : 0u;        // valid <=> (unsigned)tt < t_lim
  const int nlanes = min(32, Ub1 - 32 * warp);
  const int nsteps = Tb + nlanes - 1;
  const int64_t base = (int64_t)b * w.ND * U1;
  const bool publishes = warp + 1 < nw;                     // a full warp with a right-hand neighbour
  const int nwc = (int)(blockDim.x >> 5);                   // seam layout: [direction][warp][T]
  const bool last_col = j == Ub1 - 1;

  // This lane's cells lie on one line of the diagonal-major arrays: time index tt = s - lane in processing order is
  // lattice row t = tt (alpha) or Tb-1-tt (beta), array index (t + u) U1 + u -> a constant stride of +-U1 per step.
  const uint4* pin[2];
  LatNum* pout[2];
  int64_t dstep[2];
  const LatNum* seam_in[2];
  LatNum* seam_out[2];
  unsigned prog_prev[2], prog_mine[2];
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const int u = d ? Ub1 - 1 - j : j;
    const int64_t idx0 = base + (int64_t)((d ? Tb - 1 + lane : -lane) + u) * U1 + u;   // index at step s = 0
    dstep[d] = d ? -(int64_t)U1 : (int64_t)U1;
    pin[d] = reinterpret_cast<const uint4*>(w.pp) + idx0;
    pout[d] = (d ? w.beta : w.alpha) + idx0;
    seam_in[d] = seam + (size_t)(d * nwc + (kHasLeft ? warp - 1 : 0)) * T;
    seam_out[d] = seam + (size_t)(d * nwc + warp) * T - lane;   // indexed by step s: time index tt = s - 31 for lane 31
    prog_prev[d] = (unsigned)__cvta_generic_to_shared(prog + d * 32 + (kHasLeft ? warp - 1 : 0));
    prog_mine[d] = (unsigned)__cvta_generic_to_shared(prog + d * 32 + warp);
  }
  // cells in flight: q[d][i] is the cell of step s0 + i.  Loads are predicated and leave the registers untouched for
  // cells outside the lattice (whatever they hold is discarded by the `valid` selects below).
  uint4 q[2][kPF];
  auto load_cell = [](uint4& r, const uint4* p, bool ok) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t@p ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%5];\n\t}"
                 : "+r"(r.x), "+r"(r.y), "+r"(r.z), "+r"(r.w)
                 : "r"((int)ok), "l"(p));
  };
#pragma unroll
  for (int i = 0; i < kPF; ++i) {
    const bool ok = (unsigned)(i - lane) < t_lim;
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      q[d][i] = make_uint4(0u, 0u, 0u, 0u);
      load_cell(q[d][i], pin[d] + dstep[d] * i, ok);
    }
  }
#pragma unroll
  for (int d = 0; d < 2; ++d) pin[d] += dstep[d] * kPF;

  // own chain (alpha: alpha[t-1,u] p_blank[t-1,u]; beta: beta[t+1,u]) / value handed to the right-hand lane next step.
  // Column 0 starts from 1: alpha[0,0] = 1, and beta[T-1,U] = 1 * p_blank[T-1,U].
  float cm[2], rm[2];
  int ce[2], re[2];
  float fin_m[2] = {0.f, 0.f};
  int fin_e[2] = {kLatZeroExp, kLatZeroExp};
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    cm[d] = j == 0 ? 1.f : 0.f;
    ce[d] = j == 0 ? 0 : kLatZeroExp;
    rm[d] = 0.f;
    re[d] = kLatZeroExp;
  }
  int avail = 0;                         // time steps BOTH left-hand chains have published
  int tt = -lane;
  for (int s0 = 0; s0 < nsteps; s0 += kPF) {
    // ---- seam hand-shake, once per kPF steps (keeps the step body free of branches): the left-hand warp must have
    // published every time index this block of steps reads
    if (kHasLeft) {
      const int need = min(s0 + kPF, Tb);
      if (avail < need) {                // warp-uniform; rare once the pipeline is full (the producer runs ahead)
        int v = 0;
        for (uint32_t it = 0;; ++it) {
          if (lane == 0) {
            int v0, v1;
            asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v0) : "r"(prog_prev[0]) : "memory");
            asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v1) : "r"(prog_prev[1]) : "memory");
            v = min(v0, v1);
          }
          v = __shfl_sync(0xffffffffu, v, 0);
          if (v >= need) break;
          if (it > (1u << 24)) __trap();   // a protocol bug must not hang the GPU
        }
        avail = v;
      }
    }
#pragma unroll
    for (int i = 0; i < kPF; ++i, ++tt) {
      const int s = s0 + i;              // steps past nsteps touch invalid cells only (no loads, no stores)
      const bool valid = (unsigned)tt < t_lim;
      const bool ld_ok = (unsigned)(tt + kPF) < t_lim;
      const bool fin = last_col && tt == Tb - 1;   // alpha: the chained product alpha[T-1,U] p_blank; beta: beta[0,0]
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        const uint4 cp = q[d][i];
        load_cell(q[d][i], pin[d], ld_ok);
        pin[d] += dstep[d];
        const float pbm = __uint_as_float(cp.x), plm = __uint_as_float(cp.y);
        const int pbe = (int)cp.z, ple = (int)cp.w;
        // ---- value from the left-hand column (same time index): the neighbouring lane produced it one step ago
        float lm = __shfl_up_sync(0xffffffffu, rm[d], 1);
        int le = __shfl_up_sync(0xffffffffu, re[d], 1);
        if (kHasLeft) {                  // lane 0 takes it from the seam at time index s (all lanes read: broadcast)
          const LatNum sv = seam_in[d][min(s, Tb - 1)];
          const bool take = s < Tb;
          lm = lane == 0 ? (take ? sv.m : 0.f) : lm;
          le = lane == 0 ? (take ? sv.e : kLatZeroExp) : le;
        } else {
          lm = lane == 0 ? 0.f : lm;
          le = lane == 0 ? kLatZeroExp : le;
        }
        float nm;
        int ne;
        if (d == 0) {
          // alpha[t,u] = alpha[t-1,u] p_blank[t-1,u] + alpha[t,u-1] p_label[t,u-1]: both products were formed by their
          // source cells (cm by this lane one step ago, lm by the left-hand lane)
          lat_add_lazy(cm[d], ce[d], lm, le, nm, ne);
        } else {
          // beta[t,u] = beta[t+1,u] p_blank[t,u] + beta[t,u+1] p_label[t,u]   (mirrored: own chain / left-hand lane)
          lat_add_lazy(cm[d] * pbm, ce[d] + pbe, lm * plm, le + ple, nm, ne);
        }
        nm = valid ? nm : 0.f;
        ne = valid ? ne : kLatZeroExp;
        if (i == kPF - 1) lat_normalise(nm, ne);   // mantissas grow by < 2.83x per step: 2^13 at most in between
        ne = max(ne, 2 * kLatZeroExp);             // chains of zeros must not wrap the exponent
        if (valid) {
          LatNum o;
          o.m = nm;
          o.e = ne;
          *pout[d] = o;
        }
        pout[d] += dstep[d];
        if (d == 0) {
          cm[d] = nm * pbm; ce[d] = ne + pbe;
          rm[d] = nm * plm; re[d] = ne + ple;
        } else {
          cm[d] = rm[d] = nm;
          ce[d] = re[d] = ne;
        }
        fin_m[d] = fin ? cm[d] : fin_m[d];
        fin_e[d] = fin ? ce[d] : fin_e[d];
        if (publishes) {                 // warp-uniform; the store itself is a single predicated STS.64
          LatNum o;
          o.m = rm[d];
          o.e = re[d];
          lat_normalise(o.m, o.e);       // the consumer's mantissa bound must not compound across warps
          if (lane == 31 && valid) seam_out[d][s] = o;
        }
      }
    }
    if (publishes) {                     // one release per kPF steps: lane 31 has finished time indices < s0 + kPF - 31
      const int done = min(max(s0 + kPF - 31, 0), Tb);
      if (lane == 31 && done > 0) {
        asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(prog_mine[0]), "r"(done) : "memory");
        asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(prog_mine[1]), "r"(done) : "memory");
      }
    }
  }
  if (col_ok && last_col) {
    // gpu_rnnt_kernel.py:167-172: ll = alpha[T-1,U] + logp(blank | T-1,U);  :266-269: llBackward = beta[0,0]
    const double ll = lat_log(fin_m[0], fin_e[0]);
    w.ll_fwd[b] = ll;
    if (costs) costs[b] = (float)(-ll * (1.0 + (double)fastemit_lambda));  // rnnt_helper.py:106-116
    w.ll_bwd[b] = lat_log(fin_m[1], fin_e[1]);
  }
}

template <int kMaxThreads, int kPF>
__global__ void __launch_bounds__(kMaxThreads) rnnt_lattice_shfl_kernel(LatticeWs w, const int64_t* __restrict__ act_lens,
                                                                        const int64_t* __restrict__ label_lens, int T,
                                                                        int U1, float fastemit_lambda,
                                                                        float* __restrict__ costs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x;
  const int Tb = (int)act_lens[b], Ub1 = (int)label_lens[b] + 1;
  int* prog = reinterpret_cast<int*>(smem_raw);             // [2][32] time steps published by each warp's last lane
  LatNum* seam = reinterpret_cast<LatNum*>(prog + 64);      // [2][warps][T]
  if (threadIdx.x < 64) prog[threadIdx.x] = 0;
  __syncthreads();                                          // the only block-wide barrier
  if (Tb <= 0) {
    if (threadIdx.x == 0) {
      w.ll_fwd[b] = 0.0;
      w.ll_bwd[b] = 0.0;
      if (costs) costs[b] = 0.f;
    }
    return;
  }
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const int nw = (Ub1 + 31) >> 5;
  if (warp >= nw) return;
  if (warp == 0) rnnt_lattice_run<kPF, false>(w, b, Tb, Ub1, T, U1, warp, lane, nw, prog, seam, fastemit_lambda, costs);
  else rnnt_lattice_run<kPF, true>(w, b, Tb, Ub1, T, U1, warp, lane, nw, prog, seam, fastemit_lambda, costs);
}

static int launch_rnnt_lattice_generic(const LatticeWs& w, const int64_t* act_lens, const int64_t* label_lens, int B,
                                       int T, int U1, float fastemit_lambda, float* costs, cudaStream_t stream);

int launch_rnnt_lattice(const LatticeWs& w, const int64_t* act_lens, const int64_t* label_lens, int B, int T, int U1,
                        float fastemit_lambda, float* costs, cudaStream_t stream) {
  // CLASR_LATTICE=generic selects the block-synchronous fp64 log-space kernel (kept for U+1 > 1024 and as an A/B arm)
  const char* sel = getenv("CLASR_LATTICE");
  const int warps = (U1 + 31) / 32;
  const size_t smem = 64 * sizeof(int) + (size_t)2 * warps * T * sizeof(LatNum);
  if (U1 > 1024 || smem > 160 * 1024 || (sel && !strcmp(sel, "generic")))
    return launch_rnnt_lattice_generic(w, act_lens, label_lens, B, T, U1, fastemit_lambda, costs, stream);
  auto kern = warps <= 8 ? rnnt_lattice_shfl_kernel<256, 8> : rnnt_lattice_shfl_kernel<1024, 2>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("rnnt_lattice: T=%d U1=%d needs %zu bytes of shared memory: %s", T, U1, smem, cudaGetErrorString(e));
      return CLASR_STATUS_INVALID_VALUE;
    }
  }
  prof_begin("rnnt_lattice", stream);
  kern<<<B, warps * 32, smem, stream>>>(w, act_lens, label_lens, T, U1, fastemit_lambda, costs);
  prof_end("rnnt_lattice", stream);
  CLASR_CHECK_LAUNCH("rnnt_lattice");
  return CLASR_STATUS_SUCCESS;
}

static int launch_rnnt_lattice_generic(const LatticeWs& w, const int64_t* act_lens, const int64_t* label_lens, int B,
                                       int T, int U1, float fastemit_lambda, float* costs, cudaStream_t stream) {
  int threads = ((U1 + 31) / 32) * 32;
  if (threads > 1024) threads = 1024;
  if (threads < 32) threads = 32;
  int dch = (40 * 1024) / (U1 * 8 * 2);
  if (dch > 16) dch = 16;
  if (dch < 1) dch = 1;
  size_t smem = (size_t)2 * U1 * sizeof(double) + (size_t)2 * dch * U1 * sizeof(float2);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(rnnt_lattice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("rnnt_lattice: U1=%d needs %zu bytes of shared memory: %s", U1, smem, cudaGetErrorString(e));
      return CLASR_STATUS_INVALID_VALUE;
    }
  }
  prof_begin("rnnt_lattice", stream);
  rnnt_lattice_kernel<<<dim3(B, 2), threads, smem, stream>>>(w, act_lens, label_lens, T, U1, dch, fastemit_lambda,
                                                            costs);
  prof_end("rnnt_lattice", stream);
  CLASR_CHECK_LAUNCH("rnnt_lattice");
  return CLASR_STATUS_SUCCESS;
}

// ------------------------------------------------------------------------------------------------
// K5: softmax-fused gradient w.r.t. logits.  One warp per row; every element written.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRowWarps * 32) rnnt_grad_kernel(
    const float* __restrict__ logits, const int64_t* __restrict__ labels, const int64_t* __restrict__ act_lens,
    const int64_t* __restrict__ label_lens, int B, int T, int U1, int Vp, int blank, float fastemit_lambda,
    float clamp, const float* __restrict__ grad_out, float* __restrict__ grads, LatticeWs w) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  const int64_t rows = (int64_t)B * T * U1;
  if (row >= rows) return;
  const int b = (int)(row / ((int64_t)T * U1));
  const int rem = (int)(row - (int64_t)b * T * U1);
  const int t = rem / U1, u = rem - t * U1;
  const int Tb = (int)act_lens[b], Ub1 = (int)label_lens[b] + 1;
  float* __restrict__ g = grads + row * Vp;
  if (t >= Tb || u >= Ub1) {  // gpu_rnnt_kernel.py:343: padded cells keep zero gradient
    for (int v = lane; v < Vp; v += 32) g[v] = 0.f;
    return;
  }
  const float* __restrict__ z = logits + row * Vp;
  const int64_t idx = ((int64_t)b * w.ND + t + u) * U1 + u;
  // fp64 lattice values; every exp() argument is formed in fp64 and rounded once to fp32
  const double a = lat_log(w.alpha[idx]), bt = lat_log(w.beta[idx]), ll = w.ll_fwd[b];
  const float dn = w.denom[idx];
  const float2 lpair = w.lp[idx];
  const float go = grad_out ? grad_out[b] : 1.f;
  const bool has_label = u < Ub1 - 1;
  const int label = has_label ? (int)labels[(int64_t)b * (U1 - 1) + u] : -1;
  const double beta_t1 = (t < Tb - 1) ? lat_log(w.beta[idx + U1]) : 0.0;      // beta[t+1,u]
  const double beta_u1 = has_label ? lat_log(w.beta[idx + U1 + 1]) : 0.0;     // beta[t,u+1]
  const float base = (float)(a + bt - ll) + dn;  // grad = exp(alpha + beta + logpk - ll), logpk = dn + z
  const bool fe = fastemit_lambda > 0.f && has_label;
  const float fe_base = fe ? (float)(a + beta_u1 - ll + (double)lpair.y) + dn : 0.f;
  float blank_sub = 0.f;
  if (t == Tb - 1 && u == Ub1 - 1) blank_sub += expf((float)(a - ll + (double)lpair.x));
  if (t < Tb - 1) blank_sub += expf((float)(a + beta_t1 - ll + (double)lpair.x));
  const float label_sub =
      has_label ? expf(log1pf(fastemit_lambda) + (float)(a + beta_u1 - ll + (double)lpair.y)) : 0.f;

  constexpr int kU = 8;
  for (int v0 = lane; v0 < Vp; v0 += 32 * kU) {
    float x[kU];
#pragma unroll
    for (int j = 0; j < kU; ++j) {
      int v = v0 + 32 * j;
      x[j] = v < Vp ? ld_stream1(z + v) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < kU; ++j) {
      int v = v0 + 32 * j;
      if (v < Vp) {
        float gr = expf(x[j] + base);
        if (fe) gr += fastemit_lambda * expf(x[j] + fe_base);
        if (v == blank) gr -= blank_sub;
        if (v == label) gr -= label_sub;
        if (clamp > 0.f) gr = fmaxf(fminf(gr, clamp), -clamp);
        g[v] = gr * go;
      }
    }
  }
}

__global__ void rnnt_export_lattice_kernel(LatticeWs w, const int64_t* __restrict__ act_lens,
                                           const int64_t* __restrict__ label_lens, int B, int T, int U1,
                                           float* __restrict__ alphas, float* __restrict__ betas,
                                           float* __restrict__ ll_fwd, float* __restrict__ ll_bwd) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t cells = (int64_t)B * T * U1;
  if (i < B) {
    ll_fwd[i] = (float)w.ll_fwd[i];
    ll_bwd[i] = (float)w.ll_bwd[i];
  }
  if (i >= cells) return;
  const int b = (int)(i / ((int64_t)T * U1));
  const int rem = (int)(i - (int64_t)b * T * U1);
  const int t = rem / U1, u = rem - t * U1;
  const bool valid = t < (int)act_lens[b] && u <= (int)label_lens[b];
  const int64_t idx = ((int64_t)b * w.ND + t + u) * U1 + u;
  alphas[i] = valid ? (float)lat_log(w.alpha[idx]) : 0.f;
  betas[i] = valid ? (float)lat_log(w.beta[idx]) : 0.f;
}

}  // namespace clasr

using namespace clasr;

extern "C" size_t clasr_rnnt_workspace_bytes(int B, int T, int U1) {
  if (B <= 0 || T <= 0 || U1 <= 0) return 0;
  return lattice_ws_bytes(B, T, U1);
}

static int check_rnnt_args(const char* who, const void* logits, const void* labels, const void* act_lens,
                           const void* label_lens, int B, int T, int U1, int Vp, int blank, const void* ws,
                           size_t ws_bytes) {
  CLASR_CHECK_ARG(logits && act_lens && label_lens && ws, "%s: null pointer", who);
  CLASR_CHECK_ARG(labels || U1 == 1, "%s: null labels", who);
  CLASR_CHECK_ARG(B > 0 && T > 0 && U1 > 0 && Vp > 0, "%s: non-positive dimension (B=%d T=%d U1=%d Vp=%d)", who, B, T,
                  U1, Vp);
  CLASR_CHECK_ARG(blank >= 0 && blank < Vp, "%s: blank %d outside [0,%d)", who, blank, Vp);
  CLASR_CHECK_ARG(ws_bytes >= lattice_ws_bytes(B, T, U1), "%s: workspace too small (%zu < %zu)", who, ws_bytes,
                  lattice_ws_bytes(B, T, U1));
  CLASR_CHECK_ARG((((uintptr_t)ws) & 15) == 0, "%s: workspace must be 16-byte aligned", who);
  return CLASR_STATUS_SUCCESS;
}

extern "C" int clasr_rnnt_loss_fwd(const float* logits, const int64_t* labels, const int64_t* act_lens,
                                   const int64_t* label_lens, int B, int T, int U1, int Vp, int blank,
                                   float fastemit_lambda, float* costs, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  int rc = check_rnnt_args("rnnt_loss_fwd", logits, labels, act_lens, label_lens, B, T, U1, Vp, blank, workspace,
                           workspace_bytes);
  if (rc) return rc;
  CLASR_CHECK_ARG(costs, "rnnt_loss_fwd: null costs");
  cudaStream_t s = (cudaStream_t)stream;
  LatticeWs w = lattice_ws_carve(workspace, B, T, U1);
  const int64_t rows = (int64_t)B * T * U1;
  const int64_t grid = (rows + kRowWarps - 1) / kRowWarps;
  CLASR_CHECK_ARG(grid < 2147483647LL, "rnnt_loss_fwd: too many rows");
  prof_begin("rnnt_lse", s);
  if ((((uintptr_t)logits) & 15) == 0) {
    const int64_t pgrid = grid < (int64_t)kNumSMs * 16 ? grid : (int64_t)kNumSMs * 16;
    rnnt_lse_gather_vec_kernel<<<(unsigned)pgrid, kRowWarps * 32, 0, s>>>(logits, labels, act_lens, label_lens, B, T, U1,
                                                                         Vp, blank, w);
  } else {
    rnnt_lse_gather_kernel<<<(unsigned)grid, kRowWarps * 32, 0, s>>>(logits, labels, act_lens, label_lens, B, T, U1, Vp,
                                                                    blank, w);
  }
  prof_end("rnnt_lse", s);
  CLASR_CHECK_LAUNCH("rnnt_lse_gather");
  return launch_rnnt_lattice(w, act_lens, label_lens, B, T, U1, fastemit_lambda, costs, s);
}

extern "C" int clasr_rnnt_loss_bwd(const float* logits, const int64_t* labels, const int64_t* act_lens,
                                   const int64_t* label_lens, int B, int T, int U1, int Vp, int blank,
                                   float fastemit_lambda, float clamp, const float* grad_out, float* grads,
                                   const void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_rnnt_args("rnnt_loss_bwd", logits, labels, act_lens, label_lens, B, T, U1, Vp, blank, workspace,
                           workspace_bytes);
  if (rc) return rc;
  CLASR_CHECK_ARG(grads, "rnnt_loss_bwd: null grads");
  CLASR_CHECK_ARG(clamp >= 0.f, "rnnt_loss_bwd: `clamp` must be 0.0 or positive");
  LatticeWs w = lattice_ws_carve(const_cast<void*>(workspace), B, T, U1);
  const int64_t rows = (int64_t)B * T * U1;
  const int64_t grid = (rows + kRowWarps - 1) / kRowWarps;
  CLASR_CHECK_ARG(grid < 2147483647LL, "rnnt_loss_bwd: too many rows");
  prof_begin("rnnt_grad", (cudaStream_t)stream);
  if ((((uintptr_t)logits) & 15) == 0 && (((uintptr_t)grads) & 15) == 0) {
    const int64_t pgrid = grid < (int64_t)kNumSMs * 16 ? grid : (int64_t)kNumSMs * 16;
    rnnt_grad_vec_kernel<<<(unsigned)pgrid, kRowWarps * 32, 0, (cudaStream_t)stream>>>(
        logits, labels, act_lens, label_lens, B, T, U1, Vp, blank, fastemit_lambda, clamp, grad_out, grads, w);
  } else {
    rnnt_grad_kernel<<<(unsigned)grid, kRowWarps * 32, 0, (cudaStream_t)stream>>>(
        logits, labels, act_lens, label_lens, B, T, U1, Vp, blank, fastemit_lambda, clamp, grad_out, grads, w);
  }
  prof_end("rnnt_grad", (cudaStream_t)stream);
  CLASR_CHECK_LAUNCH("rnnt_grad");
  return CLASR_STATUS_SUCCESS;
}

extern "C" int clasr_rnnt_export_lattice(const void* workspace, size_t workspace_bytes, const int64_t* act_lens,
                                         const int64_t* label_lens, int B, int T, int U1, float* alphas, float* betas,
                                         float* ll_fwd, float* ll_bwd, void* stream) {
  CLASR_CHECK_ARG(workspace && act_lens && label_lens && alphas && betas && ll_fwd && ll_bwd,
                  "rnnt_export_lattice: null pointer");
  CLASR_CHECK_ARG(B > 0 && T > 0 && U1 > 0, "rnnt_export_lattice: non-positive dimension");
  CLASR_CHECK_ARG(workspace_bytes >= lattice_ws_bytes(B, T, U1), "rnnt_export_lattice: workspace too small");
  LatticeWs w = lattice_ws_carve(const_cast<void*>(workspace), B, T, U1);
  const int64_t cells = (int64_t)B * T * U1;
  rnnt_export_lattice_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      w, act_lens, label_lens, B, T, U1, alphas, betas, ll_fwd, ll_bwd);
  CLASR_CHECK_LAUNCH("rnnt_export_lattice");
  return CLASR_STATUS_SUCCESS;
}
