// regulariser.cu — EWC / MAS sweeps over flat fp32 parameter buffers (HBM-bandwidth-bound).
//
// Reference semantics (paths relative to /root/reference): cl_baseline_ewc.py:69-81,245-282;
// cl_baseline_mas.py:70-75,257-288; utils.py:273-321.  The reference walks a dict of hundreds of
// tensors with ~4 ATen launches each plus a host sync; here each hook is ONE streaming pass over
// the flat buffer: 128-bit loads/stores that bypass L1, a persistent grid sized from the SM count,
// and (for the EWC penalty) the per-tensor |g| statistics formed in the same pass.
//
// Algorithmic bytes per parameter: penalty_grad 16 B (3 reads + 1 write; 20 B when
// accumulating), fisher/mas accum 12 B, scale_merge 12 B (first: 8 B), penalty_value_grad 12 B (+8 B grad
// RMW), snapshot 8 B.
#include "common.cuh"

namespace clasr {

constexpr int kSweepThreads = 256;
constexpr int kSweepCtasPerSm = 8;  // 2048 threads/SM: full occupancy for a latency-bound stream

// ---------------------------------------------------------------- EWC penalty gradient (+ per-tensor |g| sums)
template <bool kAccumulate, bool kStats>
__global__ void __launch_bounds__(kSweepThreads, 2) penalty_grad_kernel(
    const float* __restrict__ theta, const float* __restrict__ theta_star, const float* __restrict__ fisher,
    float* __restrict__ grad_out, const clasr_sweep_item* __restrict__ items, int64_t n_items, float coef,
    double* __restrict__ seg_abs_sum) {
  __shared__ float warp_part[kSweepThreads / 32];
  for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
    const clasr_sweep_item item = items[it];
    const int nvec = (item.len + 3) >> 2;  // item.len = true floats; the <4-float pad of a tensor is masked to 0
    const float4* th = reinterpret_cast<const float4*>(theta + item.start);
    const float4* ts = reinterpret_cast<const float4*>(theta_star + item.start);
    const float4* fi = reinterpret_cast<const float4*>(fisher + item.start);
    float4* go = reinterpret_cast<float4*>(grad_out + item.start);
    float asum = 0.f;
    // 4 independent 128-bit loads per stream per thread are issued before any use (12-16 loads in flight)
    constexpr int kUnroll = 4;
    for (int base = 0; base < nvec; base += kUnroll * kSweepThreads) {
      float4 a[kUnroll], b[kUnroll], f[kUnroll], g0[kUnroll];
#pragma unroll
      for (int j = 0; j < kUnroll; ++j) {
        int i = base + threadIdx.x + j * kSweepThreads;
        if (i < nvec) {
          a[j] = ld_stream(th + i);
          b[j] = ld_stream(ts + i);
          f[j] = ld_stream(fi + i);
          if (kAccumulate) g0[j] = ld_stream_rw(go + i);
        }
      }
#pragma unroll
      for (int j = 0; j < kUnroll; ++j) {
        int i = base + threadIdx.x + j * kSweepThreads;
        if (i < nvec) {
          float4 g;
          // same association as the reference: ((e_lambda*2) * F) * (theta - theta*)
          g.x = __fmul_rn(__fmul_rn(coef, f[j].x), __fsub_rn(a[j].x, b[j].x));
          g.y = __fmul_rn(__fmul_rn(coef, f[j].y), __fsub_rn(a[j].y, b[j].y));
          g.z = __fmul_rn(__fmul_rn(coef, f[j].z), __fsub_rn(a[j].z, b[j].z));
          g.w = __fmul_rn(__fmul_rn(coef, f[j].w), __fsub_rn(a[j].w, b[j].w));
          if (i == nvec - 1) {  // mask the alignment pad behind the tensor's last element
            const int rem = item.len - (i << 2);
            if (rem < 4) g.w = 0.f;
            if (rem < 3) g.z = 0.f;
            if (rem < 2) g.y = 0.f;
          }
          if (kStats) asum += (fabsf(g.x) + fabsf(g.y)) + (fabsf(g.z) + fabsf(g.w));
          if (kAccumulate) {
            g.x = __fadd_rn(g.x, g0[j].x); g.y = __fadd_rn(g.y, g0[j].y); g.z = __fadd_rn(g.z, g0[j].z); g.w = __fadd_rn(g.w, g0[j].w);
          }
          st_stream(go + i, g);
        }
      }
    }
    if (kStats) {
      asum = warp_sum(asum);
      if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = asum;
      __syncthreads();
      if (threadIdx.x < 32) {
        float v = threadIdx.x < kSweepThreads / 32 ? warp_part[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) atomicAdd(seg_abs_sum + item.seg, (double)v);
      }
      __syncthreads();
    }
  }
}

__global__ void penalty_avg_kernel(const double* __restrict__ seg_abs_sum, const int64_t* __restrict__ seg_numel,
                                   int64_t n_seg, float* __restrict__ out_avg) {
  // penalty_avg = (sum_k mean|pen_k|) / n   — cl_baseline_ewc.py:76-81
  double acc = 0.0;
  for (int64_t k = threadIdx.x; k < n_seg; k += blockDim.x) acc += seg_abs_sum[k] / (double)seg_numel[k];
  __shared__ double part[32];
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.0;
    v = warp_sum_d(v);
    if (threadIdx.x == 0) out_avg[0] = (float)(v / (double)n_seg);
  }
}

// ---------------------------------------------------------------- generic elementwise sweeps
// One body per hook; all share the persistent, 4x-unrolled float4 grid-stride skeleton.
struct FisherOp {  // F += w * g^2       (cl_baseline_ewc.py:247-255)
  const float* w;
  __device__ __forceinline__ void prep(float& ww) const { ww = *w; }
  __device__ __forceinline__ float apply(float F, float g, float ww) const { return __fadd_rn(F, __fmul_rn(ww, __fmul_rn(g, g))); }  // separate roundings, like torch's mul/pow/add_
};
struct MasOp {  // Omega += |g|          (cl_baseline_mas.py:267-270)
  __device__ __forceinline__ void prep(float&) const {}
  __device__ __forceinline__ float apply(float O, float g, float) const { return O + fabsf(g); }
};

template <typename Op>
__global__ void __launch_bounds__(kSweepThreads) accum_kernel(float* __restrict__ acc, const float* __restrict__ grad,
                                                              int64_t nvec, int64_t n, Op op) {
  float ww = 0.f;
  op.prep(ww);
  float4* a4 = reinterpret_cast<float4*>(acc);
  const float4* g4 = reinterpret_cast<const float4*>(grad);
  const int64_t stride = (int64_t)gridDim.x * kSweepThreads;
  int64_t i = (int64_t)blockIdx.x * kSweepThreads + threadIdx.x;
  for (; i + 3 * stride < nvec; i += 4 * stride) {
    float4 a[4], g[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      a[j] = ld_stream_rw(a4 + i + j * stride);
      g[j] = ld_stream(g4 + i + j * stride);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      a[j].x = op.apply(a[j].x, g[j].x, ww);
      a[j].y = op.apply(a[j].y, g[j].y, ww);
      a[j].z = op.apply(a[j].z, g[j].z, ww);
      a[j].w = op.apply(a[j].w, g[j].w, ww);
      st_stream(a4 + i + j * stride, a[j]);
    }
  }
  for (; i < nvec; i += stride) {
    float4 a = ld_stream_rw(a4 + i), g = ld_stream(g4 + i);
    a.x = op.apply(a.x, g.x, ww); a.y = op.apply(a.y, g.y, ww);
    a.z = op.apply(a.z, g.z, ww); a.w = op.apply(a.w, g.w, ww);
    st_stream(a4 + i, a);
  }
  // scalar tail (n not a multiple of 4)
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    int64_t k = (nvec << 2) + threadIdx.x;
    acc[k] = op.apply(acc[k], grad[k], ww);
  }
}

__global__ void __launch_bounds__(kSweepThreads) scale_merge_kernel(float* __restrict__ dst, float* __restrict__ src,
                                                                    int64_t nvec, int64_t n, float count,
                                                                    float gamma, int first, int same) {
  const float inv = 1.0f / count;
  float4* d4 = reinterpret_cast<float4*>(dst);
  float4* s4 = reinterpret_cast<float4*>(src);
  const int64_t stride = (int64_t)gridDim.x * kSweepThreads;
  for (int64_t i = (int64_t)blockIdx.x * kSweepThreads + threadIdx.x; i < nvec; i += stride) {
    float4 s = ld_stream_rw(s4 + i);
    // fish[key] /= total_ds with a Python scalar: ATen's CUDA kernel multiplies by the fp32 reciprocal
    // (div_true_kernel_cuda, scalar fast path) - done the same way for bit parity with the reference on GPU
    s.x = __fmul_rn(s.x, inv); s.y = __fmul_rn(s.y, inv); s.z = __fmul_rn(s.z, inv); s.w = __fmul_rn(s.w, inv);
    st_stream(s4 + i, s);
    if (!same) {
      if (!first) {
        float4 d = ld_stream_rw(d4 + i);
        s.x = __fadd_rn(__fmul_rn(gamma, d.x), s.x); s.y = __fadd_rn(__fmul_rn(gamma, d.y), s.y);
        s.z = __fadd_rn(__fmul_rn(gamma, d.z), s.z); s.w = __fadd_rn(__fmul_rn(gamma, d.w), s.w);
      }
      st_stream(d4 + i, s);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    int64_t k = (nvec << 2) + threadIdx.x;
    float s = __fmul_rn(src[k], inv);
    src[k] = s;
    if (!same) dst[k] = first ? s : __fadd_rn(__fmul_rn(gamma, dst[k]), s);
  }
}

template <bool kGrad>
__global__ void __launch_bounds__(kSweepThreads) penalty_value_grad_kernel(
    const float* __restrict__ theta, const float* __restrict__ theta_star, const float* __restrict__ omega,
    int64_t nvec, int64_t n, float grad_scale, double* __restrict__ value, float* __restrict__ grad_accum) {
  const float4* t4 = reinterpret_cast<const float4*>(theta);
  const float4* s4 = reinterpret_cast<const float4*>(theta_star);
  const float4* o4 = reinterpret_cast<const float4*>(omega);
  float4* g4 = reinterpret_cast<float4*>(grad_accum);
  const int64_t stride = (int64_t)gridDim.x * kSweepThreads;
  double acc = 0.0;
  const float gs2 = 2.f * grad_scale;
  for (int64_t i = (int64_t)blockIdx.x * kSweepThreads + threadIdx.x; i < nvec; i += stride) {
    float4 t = ld_stream(t4 + i), s = ld_stream(s4 + i), o = ld_stream(o4 + i);
    float dx = t.x - s.x, dy = t.y - s.y, dz = t.z - s.z, dw = t.w - s.w;
    float part = (o.x * (dx * dx) + o.y * (dy * dy)) + (o.z * (dz * dz) + o.w * (dw * dw));
    acc += (double)part;
    if (kGrad) {
      float4 g = ld_stream_rw(g4 + i);
      g.x += gs2 * (o.x * dx); g.y += gs2 * (o.y * dy); g.z += gs2 * (o.z * dz); g.w += gs2 * (o.w * dw);
      st_stream(g4 + i, g);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    int64_t k = (nvec << 2) + threadIdx.x;
    float d = theta[k] - theta_star[k];
    acc += (double)(omega[k] * (d * d));
    if (kGrad) grad_accum[k] += gs2 * (omega[k] * d);
  }
  __shared__ double part_s[kSweepThreads / 32];
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) part_s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < kSweepThreads / 32 ? part_s[threadIdx.x] : 0.0;
    v = warp_sum_d(v);
    if (threadIdx.x == 0) atomicAdd(value, v);
  }
}

__global__ void __launch_bounds__(kSweepThreads) snapshot_kernel(float* __restrict__ dst,
                                                                 const float* __restrict__ src, int64_t nvec,
                                                                 int64_t n) {
  float4* d4 = reinterpret_cast<float4*>(dst);
  const float4* s4 = reinterpret_cast<const float4*>(src);
  const int64_t stride = (int64_t)gridDim.x * kSweepThreads;
  int64_t i = (int64_t)blockIdx.x * kSweepThreads + threadIdx.x;
  for (; i + 3 * stride < nvec; i += 4 * stride) {
    float4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = ld_stream(s4 + i + j * stride);
#pragma unroll
    for (int j = 0; j < 4; ++j) st_stream(d4 + i + j * stride, v[j]);
  }
  for (; i < nvec; i += stride) st_stream(d4 + i, ld_stream(s4 + i));
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) dst[(nvec << 2) + threadIdx.x] = src[(nvec << 2) + threadIdx.x];
}

static inline int sweep_grid(int64_t work_units) {
  int64_t g = (int64_t)kNumSMs * kSweepCtasPerSm;
  if (work_units < g) g = work_units > 0 ? work_units : 1;
  return (int)g;
}

static inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

}  // namespace clasr

using namespace clasr;

extern "C" int clasr_cl_penalty_grad(const float* theta, const float* theta_star, const float* fisher, float* grad_out,
                                     const clasr_sweep_item* items, int64_t n_items, float coef, int accumulate,
                                     double* seg_abs_sum, void* stream) {
  CLASR_CHECK_ARG(theta && theta_star && fisher && grad_out && items, "cl_penalty_grad: null pointer");
  CLASR_CHECK_ARG(n_items >= 0, "cl_penalty_grad: n_items < 0");
  CLASR_CHECK_ARG(aligned16(theta) && aligned16(theta_star) && aligned16(fisher) && aligned16(grad_out),
                  "cl_penalty_grad: buffers must be 16-byte aligned");
  if (n_items == 0) return CLASR_STATUS_SUCCESS;
  cudaStream_t s = (cudaStream_t)stream;
  prof_begin("cl_penalty_grad", s);
  int grid = (int)(n_items < (int64_t)kNumSMs * 4 ? n_items : (int64_t)kNumSMs * 4);  // >= 2 resident CTAs per SM, 2 rounds
  if (accumulate) {
    if (seg_abs_sum)
      penalty_grad_kernel<true, true><<<grid, kSweepThreads, 0, s>>>(theta, theta_star, fisher, grad_out, items, n_items, coef, seg_abs_sum);
    else
      penalty_grad_kernel<true, false><<<grid, kSweepThreads, 0, s>>>(theta, theta_star, fisher, grad_out, items, n_items, coef, nullptr);
  } else {
    if (seg_abs_sum)
      penalty_grad_kernel<false, true><<<grid, kSweepThreads, 0, s>>>(theta, theta_star, fisher, grad_out, items, n_items, coef, seg_abs_sum);
    else
      penalty_grad_kernel<false, false><<<grid, kSweepThreads, 0, s>>>(theta, theta_star, fisher, grad_out, items, n_items, coef, nullptr);
  }
  prof_end("cl_penalty_grad", s);
  CLASR_CHECK_LAUNCH("cl_penalty_grad");
  return CLASR_STATUS_SUCCESS;
}

extern "C" int clasr_cl_penalty_avg(const double* seg_abs_sum, const int64_t* seg_numel, int64_t n_seg, float* out_avg,
                                    void* stream) {
  CLASR_CHECK_ARG(seg_abs_sum && seg_numel && out_avg && n_seg > 0, "cl_penalty_avg: bad arguments");
  penalty_avg_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(seg_abs_sum, seg_numel, n_seg, out_avg);
  CLASR_CHECK_LAUNCH("cl_penalty_avg");
  return CLASR_STATUS_SUCCESS;
}

extern "C" int clasr_cl_fisher_accum(float* fisher, const float* grad, int64_t n, const float* weight_dev,
                                     void* stream) {
  CLASR_CHECK_ARG(fisher && grad && weight_dev && n >= 0, "cl_fisher_accum: bad arguments");
  CLASR_CHECK_ARG(aligned16(fisher) && aligned16(grad), "cl_fisher_accum: buffers must be 16-byte aligned");
  if (n == 0) return CLASR_STATUS_SUCCESS;
  int64_t nvec = n >> 2;
  FisherOp op{weight_dev};
  prof_begin("cl_fisher_accum", (cudaStream_t)stream);
  accum_kernel<FisherOp><<<sweep_grid((nvec + kSweepThreads * 4 - 1) / (kSweepThreads * 4)), kSweepThreads, 0,
                           (cudaStream_t)stream>>>(fisher, grad, nvec, n, op);
  prof_end("cl_fisher_accum", (cudaStream_t)stream);
  CLASR_CHECK_LAUNCH("cl_fisher_accum");
  return CLASR_STATUS_SUCCESS;
}

extern "C" int clasr_cl_mas_accum(float* omega, const float* grad, int64_t n, void* stream) {
  CLASR_CHECK_ARG(omega && grad && n >= 0, "cl_mas_accum: bad arguments");
  CLASR_CHECK_ARG(aligned16(omega) && aligned16(grad), "cl_mas_accum: buffers must be 16-byte aligned");
  if (n == 0) return CLASR_STATUS_SUCCESS;
  int64_t nvec = n >> 2;
  accum_kernel<MasOp><<<sweep_grid((nvec + kSweepThreads * 4 - 1) / (kSweepThreads * 4)), kSweepThreads, 0,
                        (cudaStream_t)stream>>>(omega, grad, nvec, n, MasOp{});
  CLASR_CHECK_LAUNCH("cl_mas_accum");
  return CLASR_STATUS_SUCCESS;
}

extern "C" int clasr_cl_scale_merge(float* dst, float* src, int64_t n, float count, float gamma, int first,
                                    void* stream) {
  CLASR_CHECK_ARG(dst && src && n >= 0, "cl_scale_merge: bad arguments");
  CLASR_CHECK_ARG(count != 0.f, "cl_scale_merge: zero divisor");
  CLASR_CHECK_ARG(aligned16(dst) && aligned16(src), "cl_scale_merge: buffers must be 16-byte aligned");
  if (n == 0) return CLASR_STATUS_SUCCESS;
  int64_t nvec = n >> 2;
  scale_merge_kernel<<<sweep_grid((nvec + kSweepThreads - 1) / kSweepThreads), kSweepThreads, 0,
                       (cudaStream_t)stream>>>(dst, src, nvec, n, count, gamma, first, dst == src);
  CLASR_CHECK_LAUNCH("cl_scale_merge");
  return CLASR_STATUS_SUCCESS;
}

extern "C" int clasr_cl_penalty_value_grad(const float* theta, const float* theta_star, const float* omega, int64_t n,
                                           float grad_scale, double* value, float* grad_accum, void* stream) {
  CLASR_CHECK_ARG(theta && theta_star && omega && value && n >= 0, "cl_penalty_value_grad: bad arguments");
  CLASR_CHECK_ARG(aligned16(theta) && aligned16(theta_star) && aligned16(omega) && aligned16(grad_accum),
                  "cl_penalty_value_grad: buffers must be 16-byte aligned");
  if (n == 0) return CLASR_STATUS_SUCCESS;
  int64_t nvec = n >> 2;
  int grid = sweep_grid((nvec + kSweepThreads - 1) / kSweepThreads);
  if (grad_accum)
    penalty_value_grad_kernel<true><<<grid, kSweepThreads, 0, (cudaStream_t)stream>>>(theta, theta_star, omega, nvec, n, grad_scale, value, grad_accum);
  else
    penalty_value_grad_kernel<false><<<grid, kSweepThreads, 0, (cudaStream_t)stream>>>(theta, theta_star, omega, nvec, n, grad_scale, value, nullptr);
  CLASR_CHECK_LAUNCH("cl_penalty_value_grad");
  return CLASR_STATUS_SUCCESS;
}

extern "C" int clasr_cl_snapshot(float* dst, const float* src, int64_t n, void* stream) {
  CLASR_CHECK_ARG(dst && src && n >= 0, "cl_snapshot: bad arguments");
  CLASR_CHECK_ARG(aligned16(dst) && aligned16(src), "cl_snapshot: buffers must be 16-byte aligned");
  if (n == 0) return CLASR_STATUS_SUCCESS;
  int64_t nvec = n >> 2;
  snapshot_kernel<<<sweep_grid((nvec + kSweepThreads * 4 - 1) / (kSweepThreads * 4)), kSweepThreads, 0,
                    (cudaStream_t)stream>>>(dst, src, nvec, n);
  CLASR_CHECK_LAUNCH("cl_snapshot");
  return CLASR_STATUS_SUCCESS;
}
