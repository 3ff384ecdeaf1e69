// common.cuh — shared helpers for libclasr_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/clasr_b200.h"

namespace clasr {

constexpr int kNumSMs = 148;  // B200

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
void prof_begin(const char* name, cudaStream_t s);
void prof_end(const char* name, cudaStream_t s);
static inline bool prec_ok(int p) { return p >= CLASR_PREC_BF16 && p <= CLASR_PREC_FP16M8; }
// hi/lo split (fp32-grade): the joint GEMM itself issues 3 MMAs per product in all of these
static inline bool prec_x3(int p) { return p == CLASR_PREC_BF16X3 || p == CLASR_PREC_FP16X3 || p == CLASR_PREC_FP16M8; }
static inline bool prec_f16(int p) { return p == CLASR_PREC_FP16X3 || p == CLASR_PREC_FP16M8; }   // 16-bit operands are fp16
static inline bool prec_m8(int p) { return p == CLASR_PREC_FP16M8; }   // backward GEMMs: fp16 + 2 e4m3 correction terms

#define CLASR_CHECK_ARG(cond, ...)              \
  do {                                          \
    if (!(cond)) {                              \
      clasr::set_error(__VA_ARGS__);            \
      return CLASR_STATUS_INVALID_VALUE;        \
    }                                           \
  } while (0)

#define CLASR_CHECK_LAUNCH(name)                                              \
  do {                                                                        \
    cudaError_t e__ = cudaGetLastError();                                     \
    if (e__ != cudaSuccess) {                                                 \
      clasr::set_error("%s: %s", name, cudaGetErrorString(e__));              \
      return CLASR_STATUS_CUDA_ERROR;                                         \
    }                                                                         \
    clasr::count_launch();                                                    \
  } while (0)

__device__ __forceinline__ float neg_inf() { return -INFINITY; }

// log(exp(a)+exp(b)) as rnnt_helper.log_sum_exp (reference rnnt_helper.py:32-44): -inf aware.
__device__ __forceinline__ float log_sum_exp(float a, float b) {
  if (a == -INFINITY) return b;
  if (b == -INFINITY) return a;
  float mx = fmaxf(a, b), mn = fminf(a, b);
  return mx + log1pf(expf(mn - mx));
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming 128-bit accesses (read-once / write-once data: keep it out of L1)
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
// same, for buffers the kernel also writes (no .nc: the read-only contract does not hold)
__device__ __forceinline__ float4 ld_stream_rw(const float4* p) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
// 256-bit global store (sm_100: STG.E.ENL2.256): one full 32-byte sector per thread per instruction; p 32-byte aligned
__device__ __forceinline__ void st_global_256(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
// 32 bytes per thread of read-once data (two 128-bit loads; p 32-byte aligned)
__device__ __forceinline__ void ld_global_256(const void* p, uint32_t (&v)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "l"(p));
  asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4+16];"
               : "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(p));
}
// 128-bit vector reduction (sm_90+): four fp32 adds to consecutive, 16-byte-aligned addresses in one instruction
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float ld_stream1(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

// Lattice workspace layout shared by the materialised and fused transducer paths.
// Diagonal-major ("skewed"): cell (t,u) of utterance b lives at ((b*ND + t+u)*U1 + u), ND = T+U1-1, so the
// anti-diagonal wavefront reads/writes contiguous memory.
//
// Precision: at the named sizes the forward/backward variables reach magnitudes of ~2.4e3, where one fp32 ulp
// (2.4e-4) already exceeds the 1e-4 gradient tolerance once exponentiated, and an fp32 log-space recursion
// accumulates that rounding over T+U steps (the reference's numba kernels and ATen's CTC do exactly that).  Here the
// lattice variables live in a scaled linear domain (LatNum below): fp32 mantissa + int32 exponent, ~1e-7 RELATIVE
// error per step wherever alpha sits.
// A lattice value in the SCALED LINEAR domain: value = m * 2^e (m >= 0 fp32, not necessarily normalised; 0 <=> m == 0).
// alpha / beta are stored like this (8 bytes per cell, like the fp64 log they replace) and decoded where they are
// consumed (lat_log): the wavefront's dependent chain then carries no transcendental and no fp64 at all, and the
// relative rounding error per step is 2^-24 wherever alpha sits.
struct __align__(8) LatNum {
  float m;
  int e;
};
// The two transition probabilities of a lattice cell, split into mantissa and integer exponent by whoever produces the
// log-probs (joint epilogue / rnnt_lse_gather), off the wavefront's dependent chain: exp(logp) = m * 2^e.
struct __align__(16) LatProb {
  float bm, lm;   // blank / label mantissa in [2^-0.5, 2^0.5] (0 for probability 0)
  int be, le;     // exponents (kLatZeroExp for probability 0)
};
constexpr int kLatZeroExp = -(1 << 28);   // exponent of the representation of 0

__device__ __forceinline__ float lat_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// exp(x) = pm * 2^pe with pm in [2^-0.5, 2^0.5]; x = -inf -> (0, kLatZeroExp); NaN propagates through pm.
// log2(e) is applied as a compensated product so that the result is good to ~2^-22 relative for any |x|.
__device__ __forceinline__ void lat_prob_split(float x, float& pm, int& pe) {
  if (x < -1e30f) { pm = 0.f; pe = kLatZeroExp; return; }
  const float kH = 1.44269502162933349609f, kL = 1.92596299112661746e-8f;   // log2(e) = kH + kL
  const float yh = x * kH;
  const float yl = fmaf(x, kH, -yh) + x * kL;
  const float n = rintf(yh);
  pm = lat_ex2((yh - n) + yl);
  pe = (int)n;
}
__device__ __forceinline__ LatProb lat_make_prob(float logp_blank, float logp_label) {
  LatProb r;
  lat_prob_split(logp_blank, r.bm, r.be);
  lat_prob_split(logp_label, r.lm, r.le);
  return r;
}
// natural log of m * 2^e as fp64 (-inf for 0, NaN stays NaN)
__device__ __forceinline__ double lat_log(float m, int e) {
  if (!(m > 0.f)) return m == 0.f ? -(double)INFINITY : (double)m;
  return ((double)e + (double)__log2f(m)) * 0.6931471805599453094;
}
__device__ __forceinline__ double lat_log(const LatNum& v) { return lat_log(v.m, v.e); }
// inverse (generic fp64 log-space kernel -> stored form)
__device__ __forceinline__ LatNum lat_from_log(double v) {
  LatNum r;
  if (!(v > -1e300)) { r.m = v != v ? (float)v : 0.f; r.e = kLatZeroExp; return r; }
  const double y = v * 1.4426950408889634074;
  const double n = floor(y);
  r.m = (float)exp2(y - n);
  r.e = (int)n;
  return r;
}

struct LatticeWs {
  float2* lp;      // [B,ND,U1] (log P(blank|t,u), log P(label_u|t,u))
  LatProb* pp;     // [B,ND,U1] the same two probabilities, split for the wavefront
  LatNum* alpha;   // [B,ND,U1]
  LatNum* beta;    // [B,ND,U1]
  float* denom;    // [B,ND,U1] negative log-sum-exp of the logits row (reduce.py:186-248)
  double* ll_fwd;  // [B]
  double* ll_bwd;  // [B]
  int ND;
};

inline size_t lattice_ws_bytes(int B, int T, int U1) {
  size_t nd = (size_t)(T + U1 - 1);
  size_t cells = (size_t)B * nd * (size_t)U1;
  size_t bytes = cells * (sizeof(LatProb) + sizeof(float2) + 2 * sizeof(LatNum) + sizeof(float));
  bytes = (bytes + 15) / 16 * 16;
  bytes += 2 * (size_t)B * sizeof(double);
  return (bytes + 255) / 256 * 256;
}

inline LatticeWs lattice_ws_carve(void* ws, int B, int T, int U1) {
  LatticeWs w;
  size_t nd = (size_t)(T + U1 - 1);
  size_t cells = (size_t)B * nd * (size_t)U1;
  char* p = (char*)ws;
  w.pp = (LatProb*)p;    p += cells * sizeof(LatProb);   // 16-byte elements first: the workspace is 16-byte aligned
  w.lp = (float2*)p;     p += cells * sizeof(float2);
  w.alpha = (LatNum*)p;  p += cells * sizeof(LatNum);
  w.beta = (LatNum*)p;   p += cells * sizeof(LatNum);
  w.denom = (float*)p;   p += cells * sizeof(float);
  p = (char*)ws + ((size_t)(p - (char*)ws) + 15) / 16 * 16;
  w.ll_fwd = (double*)p; p += (size_t)B * sizeof(double);
  w.ll_bwd = (double*)p;
  w.ND = (int)nd;
  return w;
}

// log(exp(a)+exp(b)) with fp64 carriers and fp32 transcendentals on the (non-positive) difference.
__device__ __forceinline__ double log_sum_exp_d(double a, double b) {
  if (a == -INFINITY) return b;
  if (b == -INFINITY) return a;
  const double mx = fmax(a, b), mn = fmin(a, b);
  return mx + (double)log1pf(expf((float)(mn - mx)));
}

// launched from rnnt_loss.cu; reused by the fused joint
int launch_rnnt_lattice(const LatticeWs& w, const int64_t* act_lens, const int64_t* label_lens, int B, int T, int U1,
                        float fastemit_lambda, float* costs, cudaStream_t stream);

}  // namespace clasr
