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
    const double a = w.alpha[idx], bt = w.beta[idx], ll = w.ll_fwd[b];
    const float dn = w.denom[idx];
    const float2 lpair = w.lp[idx];
    const float go = grad_out ? grad_out[b] : 1.f;
    const bool has_label = u < Ub1 - 1;
    const int label = has_label ? (int)labels[(int64_t)b * (U1 - 1) + u] : -1;
    const double beta_t1 = (t < Tb - 1) ? w.beta[idx + U1] : 0.0;      // beta[t+1,u]
    const double beta_u1 = has_label ? w.beta[idx + U1 + 1] : 0.0;     // beta[t,u+1]
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
// K3/K4: alpha and beta wavefronts.  grid = (B, 2): blockIdx.y == 0 -> alpha, 1 -> beta.
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
  double* __restrict__ out = (backward ? w.beta : w.alpha) + base;

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
      vals[0] = 0.0; out[0] = 0.0;
    } else {
      const int64_t li = (int64_t)(nd - 1) * U1 + (Ub1 - 1);
      const double v = (double)lp[li].x;
      vals[Ub1 - 1] = v;
      out[li] = v;
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
            out[(int64_t)n * U1 + u] = v;
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
            out[(int64_t)n * U1 + u] = v;
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

int launch_rnnt_lattice(const LatticeWs& w, const int64_t* act_lens, const int64_t* label_lens, int B, int T, int U1,
                        float fastemit_lambda, float* costs, cudaStream_t stream) {
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
  const double a = w.alpha[idx], bt = w.beta[idx], ll = w.ll_fwd[b];
  const float dn = w.denom[idx];
  const float2 lpair = w.lp[idx];
  const float go = grad_out ? grad_out[b] : 1.f;
  const bool has_label = u < Ub1 - 1;
  const int label = has_label ? (int)labels[(int64_t)b * (U1 - 1) + u] : -1;
  const double beta_t1 = (t < Tb - 1) ? w.beta[idx + U1] : 0.0;      // beta[t+1,u]
  const double beta_u1 = has_label ? w.beta[idx + U1 + 1] : 0.0;     // beta[t,u+1]
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
  alphas[i] = valid ? (float)w.alpha[idx] : 0.f;
  betas[i] = valid ? (float)w.beta[idx] : 0.f;
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
