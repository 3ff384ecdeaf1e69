// joint_fused.cu — fused RNNT joint + log-softmax statistics on the tcgen05 tensor cores (pass 1), and the
// C-ABI drivers of the fused joint/transducer-loss forward and backward.
//
// Replaces the fused branch of the reference's RNNTJoint.forward (NeMo/nemo/collections/asr/modules/rnnt.py
// :1403-1561) = joint_after_projection (:1587-1665) + RNNTLoss, WITHOUT ever writing the [B,T,U+1,V+1] logits:
//
//   z[b,t,u,:] = W_out . act(f[b,t,:] + g[b,u,:]) + b_out          (GEMM: M = lattice cells, K = H, N = V+1)
//
// One persistent CTA per SM walks 128-row tiles of VALID lattice cells (ragged utterances are compacted: a tile
// never contains padding cells except at an utterance's tail), and for each tile
//   producer warps   synthesise the A operand on the fly: act(f + g) -> 16-bit hi halves (fp16 or bf16) straight into
//                    the 128-byte-swizzled K-major UMMA layout in shared memory (A is not TMA-loadable: it does not
//                    exist in memory); the correction operands (e4m3 hi8 / lo8 in FP16M8, 16-bit lo halves in the
//                    three-term modes) go to TENSOR MEMORY (tcgen05.st) and feed the MMA as TMEM A-operands,
//                    because more than the hi halves of a 128 x 640 tile does not fit in 227 KB of shared memory
//   TMA warp         streams W_out (the matching operand set, K-major) through an mbarrier ring
//   MMA warp         one thread issues tcgen05.mma M=128 x N=BN x K=16 into double-buffered TMEM accumulators
//   epilogue warps   tcgen05.ld the accumulators, add the bias and keep a running (max, sum-exp) per row across the
//                    N tiles (online log-sum-exp), pick out logit[blank] and logit[label_u] (and sum z^2 for MAS);
//                    per cell only {-lse, log p(blank), log p(label)} reach HBM (20 B instead of 4.1 KB).
// The alpha/beta wavefront (rnnt_loss.cu) then runs on those compact buffers.
#include "common.cuh"
#include "tc_common.cuh"
#include <stdlib.h>

namespace clasr {

int launch_split_bf16(const float* src, int64_t rows, int cols, int64_t src_ld, void* hi, void* lo, int cols_pad,
                      cudaStream_t s, int f16, const float* scale_dev, void* hi8 = nullptr, void* lo8 = nullptr);

constexpr int kJM = 128;            // rows (lattice cells) per tile
constexpr int kJK = 64;             // K block (one 128-byte swizzle span of bf16)
constexpr int kJMaxH = 640;         // A tile (128 x H bf16) must stay resident in shared memory
// ROLE indices (the kernel maps hardware warps to roles so that the four control roles are the LAST warpgroup):
constexpr int kJThreads = 512;      // roles: 0 TMA, 1 MMA, 2 TMEM alloc, 3 Hid store, 4-7 epilogue, 8-15 A producers
constexpr int kJThreadsWide = 640;  // kWide: 4-7 epilogue of accumulator 0, 8-11 of accumulator 1, 12-19 A producers
constexpr int kJProducerWarps = 8;
constexpr int kJStages = 2;

// kPair = 1: two CTAs of a cluster (one TPC) run each tcgen05.mma together (cta_group::2, M = 256 = two 128-row tiles):
// every CTA keeps its own A tile (smem hi / TMEM lo) and loads only HALF of each W tile, so the W ring holds twice as
// many stages in the same bytes, the L2->smem traffic per CTA halves, and an SS MMA reads 4 KB (A) + 1.5 KB (B half)
// of shared memory per 48 clocks instead of 4 + 3 KB (the 1-CTA N = 96 MMA is smem-bandwidth-bound: r01b profile).
template <int kTerms, int kPair, int kStash = 0>
struct JointCfg {
  // kTerms: 1 = one 16-bit MMA per product; 3 = hi/lo split, three 16-bit MMAs; 4 = FP16M8: one fp16 MMA (K = 16) plus
  // two dense e4m3 correction MMAs (K = 32): A_hi8 . W_lo8 + A_lo8 . W_hi8 — 8 instead of 12 MMAs per 64-wide K block
  static constexpr int kBN = kTerms == 1 ? 256 : 96;          // accumulator tile width (TMEM columns)
  static constexpr int kParts = kTerms == 1 ? 1 : 2;          // W bytes streamed per stage = kParts 16-bit tiles (kTerms 4:
                                                              // fp16 | e4m3 hi8 | e4m3 lo8 = the bytes of two 16-bit tiles)
  static constexpr int kAccCols = 2 * kBN;                    // two accumulator stages
  static constexpr int kAloCol = kAccCols;                    // kTerms 3: A_lo in TMEM columns [192, 192 + H/2);
                                                              // kTerms 4: A_hi8 in [192, 192 + H/4), A_lo8 in [352, 352 + H/4)
  static constexpr int kAlo8Col = kAccCols + kJMaxH / 4;
  static constexpr int kABlockBytes = kJM * kJK * 2;          // 16 KB per K block of A
  static constexpr int kBRows = kBN / (kPair ? 2 : 1);        // W rows this CTA loads per N tile
  static constexpr int kBStageBytes = kParts * kBRows * kJK * 2; // W ring stage (per CTA)
  // kStash (kMode 3, pairs): one ring stage less (measured: no slowdown) pays for the z staging tiles
  static constexpr int kStages = kPair ? (kStash ? 3 : 2 * kJStages) : kJStages;
  static constexpr int kZStageBytes = (kStash && kPair) ? 4 * 2048 : 0;  // per epilogue warp: 2 tiles of 32 rows x 8 fp32, SW32
  static constexpr int kStagingBytes = kTerms == 1 ? 0 : kJProducerWarps * 2048;  // warp-private 32x32 bf16 lo tiles
  static constexpr int kRowTabBytes = 4 * 32 * 8;             // per lane quarter: (f offset, g offset) of its 32 rows
  static constexpr int smem_bytes(int H) {
    return (H / kJK) * kABlockBytes + kStagingBytes + kStages * kBStageBytes + kZStageBytes + kRowTabBytes + 384 + 1024;
  }
};

struct JointFwdParams {
  const float* ef;     // [B,T,H]   scaled pre-activation of f (relu: f itself; tanh: 2 log2e f; sigmoid: -log2e f)
  const float* eg;     // [B,U1,H]  same for g
  const float* bias;   // [Vp]
  const float* bias_pad;  // [round_up(Vp, 32) + 32] zero-padded copy of bias (vector loads in the pass-2 epilogue)
  const int64_t* labels;
  const int64_t* act_lens;
  const int64_t* label_lens;
  const int* tile_offsets;  // [B+1] prefix sums of per-utterance tile counts
  int B, T, U1, H, Vp, blank;
  LatticeWs w;
  // joint dropout (modules/rnnt.py:1699-1709): drop element when its 16-bit hash < drop_thresh; kept ones * drop_scale
  uint32_t drop_thresh, drop_seed_a, drop_seed_b;
  float drop_scale;
  float* sumsq;        // [B,T,U1] sum_v z^2 (MAS), or nullptr
  // ---- pass 2 (kMode == 1): recompute the logits tile and emit the softmax-fused gradient as GEMM operands
  const float* grad_out;      // [B] upstream gradient of each cost (may be nullptr == 1)
  const float* grad_cells;    // kMode 2: [B,T,U1] upstream gradient of sum_v z^2 per cell (MAS objective)
  float fastemit_lambda, clamp;
  __nv_bfloat16* dz_hi;       // [rows_pad, ldz]  dZ split hi/lo, compact tile-row order
  __nv_bfloat16* dz_lo;
  int ldz;                    // multiple of 16, >= Vp (32-byte sectors per 16-column epilogue piece)
  __nv_bfloat16* hid_hi;      // [rows_pad, ldh]  act(f+g) split hi/lo
  __nv_bfloat16* hid_lo;
  int ldh;                    // H
  float* dzb;                 // [rows_pad] dZ[., blank] per compact row (feeds the blank row of dW in joint_dfg), or null
  float* db_acc;              // [Vp] bias gradient: column sums of dZ, accumulated by the pass-2 epilogue (zeroed first)
  int* rows_pad_dev;          // [1] total_tiles * 128 (written by the tile-offset kernel)
  // ---- kMode 3 (forward that keeps the logits for the backward): z = logits + bias, fp32, compact tile-row order
  int f16;                    // 16-bit operands are fp16 instead of bf16 (CLASR_PREC_FP16X3 / FP16M8)
  // CLASR_PREC_FP16M8: the operands of the two backward GEMMs leave as {fp16, e4m3 hi8, e4m3 lo8} (tc::pack_m8) instead of
  // {fp16 hi, fp16 lo}: dz_h8 / dz_l8 [rows_pad, ldz] and hid_h8 / hid_l8 [rows_pad, ldh], one byte per element, occupy
  // the memory of dz_lo / hid_lo.  Every operand is scaled to max < 2^14 (gscale / wscale / ascale).
  int m8;
  int w8_rows;                // rows of the e4m3 hi8 part of the W tensor map (the lo8 part follows)
  uint8_t* dz_h8;
  uint8_t* dz_l8;
  uint8_t* hid_h8;
  uint8_t* hid_l8;
  const float* gscale;        // fp16 only: [2] = {S, 1/S}, the power-of-two pre-scale of dZ (joint_gscale_kernel), or null
  const float* wscale;        // fp16 only: {Sw, 1/Sw, Sa, 1/Sa, 1/(Sw Sa)}: pre-scales of W_out (max|W| Sw ~ 1) and of the
                              // hidden activations (ReLU only, else 1); logits = acc / (Sw Sa) + b
  const float* ascale;        // fp16 + ReLU: &Sa for the A producers, else null
  float* zbuf;                // [rows_pad, ldzf]
  int ldzf;                   // round_up(Vp, 32): whole 32-column epilogue pieces
};

// Developer instrumentation (built only with -DCLASR_TRACE, never into the product library): cycles each role spends
// in its barrier waits, per CTA.  Slots: 0 MMA<-tmem_empty, 1 MMA<-a_ready, 2 MMA<-full, 3 epilogue<-tmem_full (warp 4),
// 4 producer<-a_free (first producer warp), 5 TMA<-empty, 6 total kernel cycles, 7 epilogue z-store waits,
// 8 producer load+activation phase, 9 producer lo-transpose + tcgen05.st phase, 10 producer fence + arrive.
#ifdef CLASR_TRACE
__device__ unsigned long long g_joint_trace[160 * 12];
#define CLASR_TRACE_DECL unsigned long long tr_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; const long long tr_t0 = clock64();
#define CLASR_TRACE_WAIT(slot, stmt) do { const long long t_ = clock64(); stmt; tr_acc[slot] += (unsigned long long)(clock64() - t_); } while (0)
#define CLASR_TRACE_FLUSH(slot) do { if (lane == 0) g_joint_trace[blockIdx.x * 12 + (slot)] = tr_acc[slot]; } while (0)
#define CLASR_TRACE_TOTAL() do { if (threadIdx.x == 0) g_joint_trace[blockIdx.x * 12 + 6] = (unsigned long long)(clock64() - tr_t0); } while (0)
#else
#define CLASR_TRACE_DECL
#define CLASR_TRACE_WAIT(slot, stmt) stmt
#define CLASR_TRACE_FLUSH(slot)
#define CLASR_TRACE_TOTAL()
#endif

// Per-row scalars of the softmax-fused gradient (same algebra as rnnt_grad_kernel, gpu_rnnt_kernel.py:351-396), with
// the exponents pre-scaled by log2(e) so that each logit costs one FFMA + one MUFU.EX2:
//   dZ[v] = go * clamp( 2^(z log2e + base2) + fe_coef 2^(z log2e + fe_base2) - [v==blank] blank_sub - [v==label] label_sub )
// Padding rows: base2 = -inf, go = 0 -> dZ = 0.   kMode 2 (MAS): dZ = go * z with go = 2 * upstream.
struct RowGrad {
  float base2, fe_base2, fe_coef, blank_sub, label_sub, go;
};

template <int kMode>
__device__ __forceinline__ RowGrad joint_row_grad(const JointFwdParams& p, int b, int t, int u, int Tb, int Ub1,
                                                  bool valid) {
  constexpr float kLog2e = 1.4426950408889634f;
  RowGrad rg;
  rg.base2 = -INFINITY; rg.fe_base2 = -INFINITY; rg.fe_coef = 0.f; rg.blank_sub = 0.f; rg.label_sub = 0.f; rg.go = 0.f;
  if (kMode == 2 && valid) rg.go = 2.f * p.grad_cells[((int64_t)b * p.T + t) * p.U1 + u];  // d(sum z^2)/dz = 2z
  if (kMode == 1 && valid) {
    const int64_t idx = ((int64_t)b * p.w.ND + t + u) * p.U1 + u;
    const double a = lat_log(p.w.alpha[idx]), bt = lat_log(p.w.beta[idx]), ll = p.w.ll_fwd[b];
    const float dn = p.w.denom[idx];
    const float2 lpair = p.w.lp[idx];
    rg.go = p.grad_out ? p.grad_out[b] : 1.f;
    const bool has_label = u < Ub1 - 1;
    const double beta_t1 = (t < Tb - 1) ? lat_log(p.w.beta[idx + p.U1]) : 0.0;
    const double beta_u1 = has_label ? lat_log(p.w.beta[idx + p.U1 + 1]) : 0.0;
    rg.base2 = ((float)(a + bt - ll) + dn) * kLog2e;
    if (p.fastemit_lambda > 0.f && has_label) {
      rg.fe_coef = p.fastemit_lambda;
      rg.fe_base2 = ((float)(a + beta_u1 - ll + (double)lpair.y) + dn) * kLog2e;
    }
    if (t == Tb - 1 && u == Ub1 - 1) rg.blank_sub += expf((float)(a - ll + (double)lpair.x));
    if (t < Tb - 1) rg.blank_sub += expf((float)(a + beta_t1 - ll + (double)lpair.x));
    rg.label_sub = has_label ? expf(log1pf(p.fastemit_lambda) + (float)(a + beta_u1 - ll + (double)lpair.y)) : 0.f;
  }
  if (p.gscale) rg.go *= __ldg(p.gscale);   // dZ leaves pre-scaled by S (fp16 operands); every consumer multiplies by 1/S
  return rg;
}

__device__ __forceinline__ float joint_act(float x, int act) {
  if (act == CLASR_ACT_RELU) return fmaxf(x, 0.f);
  if (act == CLASR_ACT_SIGMOID) return __fdividef(1.f, 1.f + __expf(-x));
  // tanh = sign(x) * (1 - 2 / (exp(2|x|) + 1)) : absolute error ~2e-7 (what matters for sum_k h_k W_kv)
  const float ax = fabsf(x);
  const float e = __expf(2.f * ax);
  const float r = 1.f - __fdividef(2.f, e + 1.f);
  return copysignf(r, x);
}

// The A operand act(f[t,k] + g[u,k]) is needed for every lattice cell (t,u): 81 920 activations per 128-row tile.
// The argument scaling of the exponential is hoisted out of the T x U product space: joint_prep_kernel stores
//   sf = c f,  sg = c g   with c = 2 log2(e) for tanh, -log2(e) for sigmoid   (relu reads f, g themselves)
// so that per element   tanh(a+b) = 1 - 2 / (1 + 2^(sf+sg)),   sigmoid(a+b) = 1 / (1 + 2^(sf+sg))
// costs FADD + MUFU.EX2 + FADD + MUFU.RCP (+ FFMA), with IEEE saturation for any input (2^x -> inf / 0, 1/inf = 0):
// no clamping, no special cases.  (An earlier version multiplied pre-computed exp factors e^{2f} e^{2g}; that saves one
// MUFU but needs |f|, |g| clamped to stay finite, which is wrong when a large f meets a large -g.)
__global__ void joint_prep_kernel(const float* __restrict__ x, float* __restrict__ e, int64_t n4, int act) {
  const float c = act == CLASR_ACT_TANH ? 2.885390081777927f : -1.4426950408889634f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(x)[i];
    v.x *= c; v.y *= c; v.z *= c; v.w *= c;
    reinterpret_cast<float4*>(e)[i] = v;
  }
}

template <int kAct>
__device__ __forceinline__ float joint_combine(float a, float b) {
  if (kAct == CLASR_ACT_RELU) return fmaxf(a + b, 0.f);
  const float r = tc::rcp_approx(1.f + tc::ex2_approx(a + b));  // 1 + 2^x >= 1; +inf -> 0
  if (kAct == CLASR_ACT_SIGMOID) return r;
  return fmaf(-2.f, r, 1.f);
}

// Counter-based dropout mask: one 32-bit hash ("lowbias32" finaliser) per PAIR of features (k even, k+1) of one
// lattice cell, keyed by the cell's compact tile-row index (identical in pass 1, pass 2a and joint_dfg) — its two
// 16-bit halves are the uniform numbers of the two features.  tests/test_gpu_dropout.py re-derives the mask on the host.
__host__ __device__ __forceinline__ uint32_t drop_hash(uint32_t pair_index, uint32_t seed_a, uint32_t seed_b) {
  uint32_t x = (pair_index + seed_a) ^ seed_b;
  x ^= x >> 16; x *= 0x7feb352du;
  x ^= x >> 15; x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}

__global__ void joint_tile_offsets_kernel(const int64_t* __restrict__ act_lens, const int64_t* __restrict__ label_lens,
                                          int B, int* __restrict__ tile_offsets, int* __restrict__ rows_pad_dev,
                                          const float* __restrict__ bias, int Vp, float* __restrict__ bias_pad) {
  const int n_pad = (Vp + 31) / 32 * 32 + 32;
  for (int i = threadIdx.x; i < n_pad; i += blockDim.x) bias_pad[i] = i < Vp ? bias[i] : 0.f;
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int acc = 0;
    tile_offsets[0] = 0;
    for (int b = 0; b < B; ++b) {
      const int64_t Tb = act_lens[b] > 0 ? act_lens[b] : 0;
      const int64_t cells = Tb * (label_lens[b] + 1);
      acc += (int)((cells + kJM - 1) / kJM);
      tile_offsets[b + 1] = acc;
    }
    rows_pad_dev[0] = acc * kJM;
  }
}

// fp16 operands (CLASR_PREC_FP16X3): dZ = upstream x (softmax terms) is split into fp16 halves, whose relative accuracy
// degrades below 6e-5 — exactly where small upstream gradients (mean reductions, loss weights) put it.  dZ is therefore
// produced pre-scaled by the power of two S that brings max|upstream| * headroom to ~1, and dHid / dW / db / the blank
// terms are multiplied by 1/S where they are consumed (exact: powers of two).  out[0] = S, out[1] = 1/S.
// Headroom 1 keeps the `lo` halves (2^-11 of the value) in fp16's NORMAL range for every element within 2^-3 of the
// maximum.  Larger headroom (smaller operands) measured faster — 10.07 / 9.84 / 9.76 ms per step at 1 / 32 / 1024 — but
// only because the lo halves then fall into the subnormal range and lose their bits (tensor power is data-dependent).
#ifndef CLASR_SCALE_HEADROOM
#define CLASR_SCALE_HEADROOM 1.f
#endif
constexpr float kScaleHeadroom = CLASR_SCALE_HEADROOM;
// scratch[0] = running maximum (bit pattern of a non-negative float: ordered like an unsigned), scratch[1] = blocks done;
// both zeroed by the launcher.  The last block to finish converts the maximum into the scale.
__global__ void joint_gscale_kernel(const float* __restrict__ g, int64_t n, float headroom, float* __restrict__ out,
                                    unsigned int* __restrict__ scratch) {
  __shared__ float red[32];
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(g[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    m = warp_max(m);
    if (threadIdx.x == 0) {
      if (isfinite(m)) atomicMax(scratch, __float_as_uint(m));
      __threadfence();
      if (atomicAdd(scratch + 1, 1u) == gridDim.x - 1) {
        m = __uint_as_float(atomicMax(scratch, 0u));
        int e = 0;
        if (m > 0.f) {
          frexpf(m * headroom, &e);                 // m * headroom = f * 2^e, f in [0.5, 1)
          e = e > 100 ? 100 : (e < -100 ? -100 : e);
        }
        out[0] = ldexpf(1.f, -e);
        out[1] = ldexpf(1.f, e);
      }
    }
  }
}

static int launch_joint_gscale(const float* g, int64_t n, float headroom, float* out, cudaStream_t s) {
  unsigned int* scratch = reinterpret_cast<unsigned int*>(out) + 8;   // same 256-byte workspace slot
  if (cudaMemsetAsync(scratch, 0, 2 * sizeof(unsigned int), s) != cudaSuccess) {
    set_error("joint_gscale: memset failed");
    return CLASR_STATUS_CUDA_ERROR;
  }
  int64_t blocks = (n + 4095) / 4096;
  if (blocks > kNumSMs) blocks = kNumSMs;
  if (blocks < 1) blocks = 1;
  joint_gscale_kernel<<<(unsigned)blocks, 256, 0, s>>>(g, n, headroom, out, scratch);
  CLASR_CHECK_LAUNCH("joint_gscale");
  return CLASR_STATUS_SUCCESS;
}

// fp16 operands: gs[2..3] = {Sw, 1/Sw} (written before); ReLU: gs[16..17], gs[18..19] = scales of max|f|, max|g| ->
// hidden values relu(f + g) <= max|f| + max|g| <= 2 max(1/Sf, 1/Sg): Sa = min(Sf, Sg) / 2.  Writes gs[4..6].
// `extra` (a power of two): 1, or for CLASR_PREC_FP16M8 the factor that brings the bounded hidden activations (times the
// dropout rescale 1 / (1 - p)) to just below 2^14.
__global__ void joint_ascale_kernel(float* __restrict__ gs, int relu, float extra) {
  const float sa = (relu ? 0.5f * fminf(gs[16], gs[18]) : 1.f) * extra;
  gs[4] = sa;
  gs[5] = 1.f / sa;          // exact: power of two
  gs[6] = gs[3] * gs[5];
}

__device__ __forceinline__ int find_utterance(const int* __restrict__ offs, int B, int tile) {
  int lo = 0, hi = B - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (offs[mid] <= tile) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// kWide = 1 (pass 2): the gradient epilogue is the bottleneck of pass 2a (r01g profile: epilogue warps 82 % busy, MMA
// warp 26 % waiting on tmem_empty), so it gets EIGHT warps — one warpgroup per TMEM accumulator buffer, alternating N
// tiles — in a 640-thread CTA whose register file is re-split with setmaxnreg (control warps 32, epilogue 112,
// producers 112 registers per thread).
template <int kTerms, int kMode, int kAct, int kPair, int kWide>
__global__ void __launch_bounds__(kWide ? kJThreadsWide : kJThreads, 1)
joint_fwd_kernel(const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                 const __grid_constant__ CUtensorMap tmHid, const __grid_constant__ CUtensorMap tmZ, JointFwdParams p) {
  using C = JointCfg<kTerms, kPair, kMode == 3>;
  constexpr bool kZStage = C::kZStageBytes > 0;   // z leaves through swizzled smem tiles + TMA stores
  constexpr int kStages = C::kStages;
  constexpr int kEpiWarps = kWide ? 8 : 4;
  constexpr int kProdWarp0 = 4 + kEpiWarps;          // first producer warp (a multiple of 4: warp % 4 = TMEM quarter)
  constexpr int kBatch = kWide ? 4 : 8;              // row pairs per producer load batch (register budget)
  constexpr int kNB = 16 / kBatch;
  // kMode 0: forward statistics;  3: the same + keep z and Hid for the backward;  4: the same + keep Hid only;
  // 1 / 2: gradient passes that rebuild the hidden activations (transducer loss / MAS sum of squares);
  // 5 / 6: the same gradient passes with the A operands LOADED from the hidden activations a kMode-4 forward kept
  constexpr bool kStats = kMode == 0 || kMode == 3 || kMode == 4;
  constexpr bool kGrad = kMode == 1 || kMode == 2 || kMode == 5 || kMode == 6;
  constexpr int kGMode = (kMode == 1 || kMode == 5) ? 1 : ((kMode == 2 || kMode == 6) ? 2 : 0);   // which gradient
  constexpr bool kEmitHid = kMode >= 1 && kMode <= 4;   // this pass writes the hidden activations as GEMM operands
  constexpr bool kLoadA = kMode >= 5;                   // ... or reads them back instead of evaluating act(f + g)
  constexpr bool kHelp = kWide && kGrad;   // the A producers share the gradient epilogue of the N loop (see below)
  constexpr int kTailTiles = kLoadA ? 0 : 2;   // ... except for the last N tiles of a row tile (nothing to produce: all)
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
  const int kblocks = p.H / kJK;
  uint8_t* a_smem = smem;                                          // [kblocks][128 x 64 bf16], SW128 K-major
  uint8_t* staging = a_smem + kblocks * C::kABlockBytes;            // BF16X3 only: 8 warps x [32 rows x 32 k] bf16
  uint8_t* b_ring = staging + C::kStagingBytes;                     // [stages][parts][BN x 64 bf16]
  uint8_t* zstage = b_ring + kStages * C::kBStageBytes;             // kZStage: [4 epilogue warps][2 tiles][32 rows x 32 B]
  uint8_t* rowtab = zstage + C::kZStageBytes;                       // int2[4 quarters][32 rows]
  uint64_t* bars = (uint64_t*)(rowtab + C::kRowTabBytes);
  uint64_t* full = bars;                 // [kStages]  W stage landed            (pair: the leader's copy)
  uint64_t* empty = full + kStages;      // [kStages]  W stage consumed
  uint64_t* tmem_full = empty + kStages;    // [2]
  uint64_t* tmem_empty = tmem_full + 2;     // [2]                                 (pair: the leader's copy)
  uint64_t* a_ready = tmem_empty + 2;       // [kblocks <= 10] A K-block written   (pair: the leader's copy)
  uint64_t* a_free = a_ready + 10;          // [kblocks <= 10] last MMA reading A K-block retired (+ pass 2: its
                                            //                 TMA store to Hid_hi has read the block)
  uint64_t* hid_ready = a_free + 10;        // [kblocks <= 10] pass 2: A K-block written (local copy for the store warp)
  uint32_t* tmem_base_slot = (uint32_t*)(hid_ready + 10);

  // Roles: 0 TMA, 1 MMA, 2 TMEM alloc, 3 Hid store / load, 4.. epilogue warps, then the A producers.  The four control
  // roles sit in the LAST warpgroup of the CTA: the sub-partition arbiter favours the highest warp id, and the MMA / TMA
  // warps must issue the moment their barrier flips — as warps 0-3 they lost the issue slot to the 4 busy worker warps
  // of their sub-partition (role trace of pass 2a: the MMA warp spent 3.6 M cycles issuing what takes 1.9 M in pass 1).
  // A role keeps (warp index mod 4), i.e. the tensor-memory lane quarter it may touch.
  const int hw_warp = tc::warp_idx_uniform();
  constexpr int kNumWarps = (kWide ? kJThreadsWide : kJThreads) / 32;
#ifndef CLASR_CTRL_LAST
#define CLASR_CTRL_LAST 1
#endif
#if CLASR_CTRL_LAST
  const int warp = hw_warp >= kNumWarps - 4 ? hw_warp - (kNumWarps - 4) : hw_warp + 4;   // role index
#else
  const int warp = hw_warp;
#endif
  const int lane = threadIdx.x & 31;
  CLASR_TRACE_DECL
  const int total_tiles = (int)tc::uniform_u32((uint32_t)p.tile_offsets[p.B]);
  // pair mode: the cluster walks PAIRS of consecutive row tiles; this CTA owns tile 2*step + rank (possibly a null
  // tile past the end, which still takes part in every barrier hand-shake)
  const uint32_t cta_rank = kPair ? tc::cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int tile_first = kPair ? 2 * (int)(blockIdx.x >> 1) + (int)cta_rank : (int)blockIdx.x;
  const int tile_stride = kPair ? (int)(gridDim.x & ~1u) : (int)gridDim.x;
  const int tile_end = kPair ? (total_tiles + 1) / 2 * 2 : total_tiles;  // both CTAs run the same number of steps
  const int n_tiles = (p.Vp + C::kBN - 1) / C::kBN;
  const int n_last = ((p.Vp - (n_tiles - 1) * C::kBN) + 15) / 16 * 16;  // width of the last N tile (multiple of 16)

  if (warp == 0 && tc::elect_one()) {
    tc::prefetch_tmap(&tmW_hi);
    if (kTerms > 1) tc::prefetch_tmap(&tmW_lo);
    if (kMode >= 1) tc::prefetch_tmap(&tmHid);
    if (kZStage) tc::prefetch_tmap(&tmZ);
  }
  if (warp == 1 && tc::elect_one()) {
    constexpr int kCtas = kPair ? 2 : 1;
    for (int i = 0; i < kStages; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&tmem_full[i], 1); tc::mbar_init(&tmem_empty[i], (kHelp ? 12 : 4) * kCtas); }
    for (int i = 0; i < 10; ++i) {
      tc::mbar_init(&a_ready[i], kJProducerWarps * kCtas + (kLoadA ? 1 : 0));   // kLoadA: + the loader's expect_tx
      tc::mbar_init(&a_free[i], kEmitHid ? 2 : 1);
      tc::mbar_init(&hid_ready[i], kJProducerWarps);
    }
    tc::fence_barrier_init();
  }
  if (kPair) tc::cluster_sync_all();  // barrier inits visible to the peer before any remote arrive / multicast commit
  if (warp == 2) {
    if (kPair) tc::tmem_alloc_2sm(tmem_base_slot, 512);
    else tc::tmem_alloc(tmem_base_slot, 512);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tc::uniform_u32(*tmem_base_slot);

  // kWide: the register file is re-split between the warpgroups; every setmaxnreg sits at the top of the branch whose
  // code it governs (ptxas allocates registers per such region), executed by all four warps of the warpgroup.  The
  // pool is what the CTA got at launch (640 x 96 = 61 440 registers, NOT the SM's 65 536): 128 x 32 + 512 x 112 = 61 440.
  // (A first version asked for 64 512 and its last setmaxnreg.inc blocked forever.)
  if (warp < 4) {
  if (kWide) asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
  if (warp == 0) {
    // ============================ TMA producer: W ring (whole warp loops, one elected lane issues) ===========
    int stage = 0;
    uint32_t phase = 0;
    const uint64_t pol_keep = tc::l2_policy_evict_last();
    for (int tile = tile_first; tile < tile_end; tile += tile_stride) {
      for (int nt = 0; nt < n_tiles; ++nt) {
        // pair: this CTA loads rows [n0 + rank * width/2, +kBRows) of the W tile (the MMA reads width/2 of them)
        const int n_cur = (nt == n_tiles - 1) ? n_last : C::kBN;
        const int n0 = nt * C::kBN + (kPair ? (int)cta_rank * (n_cur / 2) : 0);
        for (int kb = 0; kb < kblocks; ++kb) {
          CLASR_TRACE_WAIT(5, tc::mbar_wait(&empty[stage], phase ^ 1));
          if (tc::elect_one()) {
            uint8_t* st = b_ring + stage * C::kBStageBytes;
            if (kTerms == 4) {
              // fp16 tile [kBRows][64 k] (128-byte swizzle) | e4m3 hi8 [kBRows][64 B] | e4m3 lo8 (64-byte swizzle); the two
              // e4m3 arrays are one tensor map: lo8 rows follow w8_rows below the hi8 rows
              uint8_t* st8 = st + C::kBRows * kJK * 2;
              if (kPair) {
                if (leader) tc::mbar_expect_tx(&full[stage], 2 * C::kBStageBytes);
                tc::tma_load_2d_2sm_hint(st, &tmW_hi, &full[stage], kb * kJK, n0, pol_keep);
                tc::tma_load_2d_2sm_hint(st8, &tmW_lo, &full[stage], kb * kJK, n0, pol_keep);
                tc::tma_load_2d_2sm_hint(st8 + C::kBRows * kJK, &tmW_lo, &full[stage], kb * kJK, n0 + p.w8_rows, pol_keep);
              } else {
                tc::mbar_expect_tx(&full[stage], C::kBStageBytes);
                tc::tma_load_2d(st, &tmW_hi, &full[stage], kb * kJK, n0);
                tc::tma_load_2d(st8, &tmW_lo, &full[stage], kb * kJK, n0);
                tc::tma_load_2d(st8 + C::kBRows * kJK, &tmW_lo, &full[stage], kb * kJK, n0 + p.w8_rows);
              }
            } else if (kPair) {
              if (leader) tc::mbar_expect_tx(&full[stage], 2 * C::kBStageBytes);
              if (kMode >= 1) {  // the kernel streams GBs of stores through L2: keep the 2.6 MB of W resident
                tc::tma_load_2d_2sm_hint(st, &tmW_hi, &full[stage], kb * kJK, n0, pol_keep);
                if (kTerms > 1)
                  tc::tma_load_2d_2sm_hint(st + C::kBRows * kJK * 2, &tmW_lo, &full[stage], kb * kJK, n0, pol_keep);
              } else {
                tc::tma_load_2d_2sm(st, &tmW_hi, &full[stage], kb * kJK, n0);
                if (kTerms > 1) tc::tma_load_2d_2sm(st + C::kBRows * kJK * 2, &tmW_lo, &full[stage], kb * kJK, n0);
              }
            } else {
              tc::mbar_expect_tx(&full[stage], C::kBStageBytes);
              tc::tma_load_2d(st, &tmW_hi, &full[stage], kb * kJK, n0);
              if (kTerms > 1) tc::tma_load_2d(st + C::kBRows * kJK * 2, &tmW_lo, &full[stage], kb * kJK, n0);
            }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ============================ MMA issuer (whole warp loops, one elected lane issues; pair: leader CTA) ====
    const uint32_t idesc_full = tc::make_idesc_16(kPair ? 2 * kJM : kJM, C::kBN, 0, 0, p.f16);
    const uint32_t idesc_last = tc::make_idesc_16(kPair ? 2 * kJM : kJM, n_last, 0, 0, p.f16);
    // descriptor words of the A tile / W ring bases (see the issue loop)
    const uint32_t a_dlo = (uint32_t)tc::make_desc_kmajor_sw128(tc::smem_u32(a_smem));
    const uint32_t b_dlo = (uint32_t)tc::make_desc_kmajor_sw128(tc::smem_u32(b_ring));
    const uint32_t d_hi128 = (uint32_t)(tc::make_desc_kmajor_sw128(0) >> 32);
    const uint32_t d_hi64 = (uint32_t)(tc::make_desc_kmajor_sw64(0) >> 32);
    int stage = 0;
    uint32_t phase = 0;
    int acc_it = 0;
    int tile_it = 0;
    for (int tile = tile_first; tile < tile_end; tile += tile_stride, ++tile_it) {
      const uint32_t tile_phase = tile_it & 1;
      for (int nt = 0; nt < n_tiles; ++nt, ++acc_it) {
        const int acc = acc_it & 1;
        const uint32_t acc_phase = (acc_it >> 1) & 1;
        const bool last_nt = nt == n_tiles - 1;
        const uint32_t idesc = last_nt ? idesc_last : idesc_full;
        CLASR_TRACE_WAIT(0, tc::mbar_wait(&tmem_empty[acc], acc_phase ^ 1));
        tc::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * C::kBN;
        for (int kb = 0; kb < kblocks; ++kb) {
          if (nt == 0) CLASR_TRACE_WAIT(1, tc::mbar_wait(&a_ready[kb], tile_phase));
          CLASR_TRACE_WAIT(2, tc::mbar_wait(&full[stage], phase));
          tc::tc_fence_after();
          if (tc::elect_one()) {
            // Descriptor LOW words by plain additions to two per-kernel bases: the start-address field is (byte address
            // >> 4) and shared memory ends below 2^18 bytes, so an offset never carries into the next field.  (Building
            // each of the 12 descriptors of a K block from its address cost ~90 uniform-datapath instructions per block
            // — with 16 busy worker warps on the SM the issue loop, not the tensor pipe, then set the pace of pass 2a.)
            const uint32_t da = a_dlo + kb * (C::kABlockBytes >> 4);
            const uint32_t dbh = b_dlo + stage * (C::kBStageBytes >> 4);
            const uint32_t dbl = dbh + ((C::kBRows * kJK * 2) >> 4);
#pragma unroll
            for (int kk = 0; kk < kJK / 16; ++kk) {
              const uint32_t accum = (kb == 0 && kk == 0) ? 0u : 1u;
              const uint64_t d_a = tc::desc_from_words(d_hi128, da + kk * 2);
              const uint64_t d_bh = tc::desc_from_words(d_hi128, dbh + kk * 2);
              const uint64_t d_bl = tc::desc_from_words(d_hi128, dbl + kk * 2);
              // A_lo from tensor memory: 16 bf16 of K = 8 packed 32-bit columns
              const uint32_t a_lo_t = tmem_base + C::kAloCol + kb * (kJK / 2) + kk * 8;
              if (kPair) {
                tc::umma_ss_2sm(d_tmem, d_a, d_bh, idesc, accum);
                if (kTerms == 3) {
                  tc::umma_ss_2sm(d_tmem, d_a, d_bl, idesc, 1u);
                  tc::umma_ts_2sm(d_tmem, a_lo_t, d_bh, idesc, 1u);
                }
              } else {
                tc::umma_ss(d_tmem, d_a, d_bh, idesc, accum);
                if (kTerms == 3) {
                  tc::umma_ss(d_tmem, d_a, d_bl, idesc, 1u);
                  tc::umma_ts(d_tmem, a_lo_t, d_bh, idesc, 1u);
                }
              }
            }
            if (kTerms == 4) {
              // correction terms: A_hi8 . W_lo8 + A_lo8 . W_hi8, e4m3 A operands from tensor memory (4 k per 32-bit column:
              // 8 columns per K = 32 MMA), W tiles K-major with 64-byte rows (hi8 tile, then lo8 tile, after the fp16 tile)
              const uint32_t dh8 = dbl, dl8 = dbl + ((C::kBRows * kJK) >> 4);
              const uint32_t a8 = tmem_base + kb * (kJK / 4);
#pragma unroll
              for (int k8 = 0; k8 < kJK / 32; ++k8) {
                const uint64_t d_h8 = tc::desc_from_words(d_hi64, dh8 + k8 * 2);
                const uint64_t d_l8 = tc::desc_from_words(d_hi64, dl8 + k8 * 2);
                if (kPair) {
                  tc::umma_f8_ts_2sm(d_tmem, a8 + C::kAloCol + k8 * 8, d_l8, idesc, 1u);
                  tc::umma_f8_ts_2sm(d_tmem, a8 + C::kAlo8Col + k8 * 8, d_h8, idesc, 1u);
                } else {
                  tc::umma_f8_ts(d_tmem, a8 + C::kAloCol + k8 * 8, d_l8, idesc, 1u);
                  tc::umma_f8_ts(d_tmem, a8 + C::kAlo8Col + k8 * 8, d_h8, idesc, 1u);
                }
              }
            }
            // last N tile: this K block of A (smem hi + TMEM lo) is dead once these MMAs retire -> the producers
            // may already write the next row tile's block while the remaining K blocks are still being consumed
            if (kPair) {
              tc::umma_commit_2sm(&empty[stage], 0b11);
              if (last_nt) tc::umma_commit_2sm(&a_free[kb], 0b11);
            } else {
              tc::umma_commit(&empty[stage]);
              if (last_nt) tc::umma_commit(&a_free[kb]);
            }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (tc::elect_one()) {
          if (kPair) tc::umma_commit_2sm(&tmem_full[acc], 0b11);
          else tc::umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 3 && kLoadA) {
    // ============================ kept hidden activations: A_hi load warp ============================
    // The 16-bit A blocks are exactly the boxes the kMode-4 forward stored ([128 rows x 64 k], 128-byte swizzle): one
    // TMA load per K block as soon as the previous row tile's last MMAs on that block have retired.  The bytes are
    // counted on a_ready[kb] (pair: the leader's copy, which also collects both CTAs' correction-operand arrivals).
    int tile_it = 0;
    for (int tile = tile_first; tile < tile_end; tile += tile_stride, ++tile_it) {
      for (int kb = 0; kb < kblocks; ++kb) {
        if (tile_it > 0) tc::mbar_wait(&a_free[kb], (tile_it - 1) & 1);
        if (tc::elect_one()) {
          // (a null tile of a pair lies past the tensor's last row: the TMA zero-fills it and still counts the bytes)
          if (kPair) {
            if (leader) tc::mbar_expect_tx(&a_ready[kb], 2 * C::kABlockBytes);
            tc::tma_load_2d_2sm(a_smem + kb * C::kABlockBytes, &tmHid, &a_ready[kb], kb * kJK, tile * kJM);
          } else {
            tc::mbar_expect_tx(&a_ready[kb], C::kABlockBytes);
            tc::tma_load_2d(a_smem + kb * C::kABlockBytes, &tmHid, &a_ready[kb], kb * kJK, tile * kJM);
          }
          if (tile + tile_stride < total_tiles)   // the next row tile's block: towards L2 a whole tile time ahead
            asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(&tmHid), "r"(kb * kJK),
                         "r"((tile + tile_stride) * kJM)
                         : "memory");
        }
        __syncwarp();
      }
    }
  } else if (warp == 3 && kEmitHid) {
    // ============================ pass 2: Hid_hi store warp ============================
    // The bf16 hi halves of act(f+g) that the dW GEMM consumes ARE the A tile in shared memory (128-byte-swizzled
    // [128 rows x 64 k] blocks = exactly a TMA box): one elected lane stores each block with a TMA store as soon as
    // the producers have written it, and releases the block (second arrival on a_free) once the store has read it.
    int tile_it = 0;
    const uint64_t pol_stream = tc::l2_policy_evict_first();
    for (int tile = tile_first; tile < tile_end; tile += tile_stride, ++tile_it) {
      const bool tile_ok = tile < total_tiles;
      for (int kb = 0; kb < kblocks; ++kb) {
        tc::mbar_wait(&hid_ready[kb], tile_it & 1);
        if (tile_ok && tc::elect_one()) {
          tc::tma_store_2d_hint(&tmHid, a_smem + kb * C::kABlockBytes, kb * kJK, tile * kJM, pol_stream);
          tc::bulk_commit_group();
        }
        __syncwarp();
      }
      if (tc::elect_one()) {
        tc::bulk_wait_group_read0();
        for (int kb = 0; kb < kblocks; ++kb) tc::mbar_arrive(&a_free[kb]);
      }
      __syncwarp();
    }
    if (tc::elect_one()) tc::bulk_wait_group0();  // all stores complete before the CTA exits
    __syncwarp();
  }
  } else {
    if (kWide) asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    // Warps 4 .. kProdWarp0-1 are the epilogue warps, the rest the A producers.  kHelp (pass 2, 640-thread layout): the
    // gradient epilogue is the bottleneck of the N loop (r02e profile: an epilogue warp needs ~16 K cycles per N tile
    // against ~4.6 K of MMA time, while the producers idle 60 % of the kernel on a_free), so after writing a row tile's A
    // blocks the two producer warps of a lane quarter take a third each of the pieces of that tile's N tiles — all but
    // the last kTailTiles, during which they already produce the next row tile.  Both roles run the SAME code below.
    const bool is_prod = warp >= kProdWarp0;
    // ============================ epilogue ============================
    // kMode 0: online log-sum-exp + gather of logit[blank] / logit[label]  (pass 1)
    // kMode 1: softmax-fused gradient dZ = clamp(exp(logp + occupancy) - blank/label terms) * grad_out, split into
    //          bf16 hi/lo and written in tile-row order as the operand of the two backward GEMMs (pass 2)
    const int q = warp & 3;
    const int egrp = (warp - 4) >> 2;   // kWide: this warpgroup serves accumulator buffer `egrp` only
    const uint64_t pol_stream = tc::l2_policy_evict_first();
    (void)pol_stream;
    const float inv_w = p.wscale ? __ldg(p.wscale + 4) : 1.f;   // 1 / (Sw Sa); 1.0: fmaf(acc, 1, b) == acc + b exactly
    const float inv_gs = (kGrad && p.gscale) ? __ldg(p.gscale + 1) : 1.f;   // 1 / S of the pre-scaled dZ (exact power of two)
    const int row = q * 32 + lane;
    const int ehalf = (warp - kProdWarp0) >> 2;   // helper: which third of the pieces (producer warp's K half)
    auto epilogue_tile = [&](const int tile, const int tile_it, const bool helper) {
      int acc_it = tile_it * n_tiles;
      const bool tile_ok = tile < total_tiles;  // pair mode: a null tile past the end only does the hand-shakes
      const int b = find_utterance(p.tile_offsets, p.B, tile_ok ? tile : 0);
      const int Tb = (int)p.act_lens[b], Ub1 = (int)p.label_lens[b] + 1;
      const int r = (tile - p.tile_offsets[b]) * kJM + row;  // cell index inside the utterance (t-major)
      const bool valid = tile_ok && r < Tb * Ub1;
      const int t = valid ? r / Ub1 : 0;
      const int u = valid ? r - t * Ub1 : 0;
      const int label = (valid && u < Ub1 - 1) ? (int)p.labels[(int64_t)b * (p.U1 - 1) + u] : -1;
      const int64_t idx = ((int64_t)b * p.w.ND + t + u) * p.U1 + u;
      float m = -INFINITY, s = 0.f, zb = 0.f, zl = 0.f, ssq = 0.f;
      const int64_t grow = (int64_t)tile * kJM + row;  // compact tile-row index of this thread's row
      // pass-2 per-row scalars (joint_row_grad)
      constexpr float kLog2e = 1.4426950408889634f;
      const RowGrad rg = joint_row_grad<kGMode>(p, b, t, u, Tb, Ub1, valid);
      const float base2 = rg.base2, fe_base2 = rg.fe_base2, fe_coef = rg.fe_coef, blank_sub = rg.blank_sub,
                  label_sub = rg.label_sub, go = rg.go;
      for (int nt = 0; nt < n_tiles; ++nt, ++acc_it) {
        const int acc = acc_it & 1;
        const uint32_t acc_phase = (acc_it >> 1) & 1;
        // Who works on this N tile (warp-uniform).  Its 16-column pieces are cut into three ranges of piece PAIRS
        // (FP16M8 stores the e4m3 operands of a pair as one sector); tmem_empty counts three arrivals per lane quarter.
        //   no helpers:             the epilogue warp of this accumulator buffer takes all three ranges
        //   helpers that produce:   on helped N tiles the producer warp bound to this accumulator buffer takes range 0 —
        //                           with the ~33 K cycles of A production per row tile that evens out the four warps' loads
        //   helpers that only load: the four warps of a lane quarter are equals, three of them serve an N tile (rotating)
        // A warp must never skip a phase of a barrier it waits on later (a parity wait cannot tell phase n from n + 2):
        // a helper that produces only ever touches ITS buffer's barrier (the N tiles it leaves out at the end of a row tile
        // have completed long before it is back: its own A blocks gate the next row tile's MMAs), and a warp whose turn it
        // is to sit out an N tile in the rotation still observes that tile's tmem_full phase.
        int k_lo = 0, k_hi = 3;
        bool mine;
        if (!kHelp) {
          mine = !kWide || acc == egrp;
        } else if (kLoadA) {
          const int k = ((helper ? 2 + ehalf : egrp) - acc_it) & 3;
          mine = k < 3;
          k_lo = k; k_hi = k + 1;
          if (!mine) tc::mbar_wait(&tmem_full[acc], acc_phase);
        } else {
          const bool helped = nt < n_tiles - kTailTiles;
          if (helper) { mine = helped && acc == ehalf; k_hi = 1; }
          else { mine = acc == egrp; k_lo = helped ? 1 : 0; }
        }
        if (!mine) continue;
        // columns this N tile must cover: up to Vp (stats) or up to the padded operand width ldz (gradient)
        const int width = (kGrad ? p.ldz : p.Vp);
        const int ncols = (nt == n_tiles - 1) ? (width - nt * C::kBN) : C::kBN;
        // pass 2: this warp's range of 16-column pieces
        const int npieces_all = tile_ok ? (ncols + 15) >> 4 : 0;
        const int units = (npieces_all + 1) >> 1;
        const int c_begin = 2 * (k_lo * units / 3);
        const int c_end = min(npieces_all, 2 * (k_hi * units / 3));
        (void)c_end;
        // pass 2: the bias of a piece is in flight one piece ahead (with ~28 KB of L1 left beside 226 KB of shared memory
        // these loads come from L2: ~280 cycles per piece were exposed in front of the first FFMA, r02e profile);
        // the first piece's before the accumulator wait
        float4 bv[4];
        if (kGrad) {
          const float4* bias4 = reinterpret_cast<const float4*>(p.bias_pad + nt * C::kBN + c_begin * 16);  // zero-padded copy
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) bv[j4] = __ldg(bias4 + j4);
        }
        CLASR_TRACE_WAIT(3, tc::mbar_wait(&tmem_full[acc], acc_phase));
        tc::tc_fence_after();
        bool acc_released = false;   // pass 2 hands the accumulator back from inside its piece loop
        // arrivals this warp owes tmem_empty: kHelp counts three per lane quarter and N tile
        const int n_arrive = kHelp ? k_hi - k_lo : 1;
        if (kStats) {
          // (TMEM reads run one piece ahead and the accumulator is released after the last one: see pass 2 below)
          const int npieces = tile_ok ? (ncols + 31) >> 5 : 0;
          const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + acc * C::kBN;
          uint32_t rr[32];
          if (npieces > 0) tc::tmem_ld32(tacc, rr);
          acc_released = npieces > 0;
#pragma unroll 1
          for (int c = 0; c < npieces; ++c) {
            const int col0 = nt * C::kBN + c * 32;
            // branch-free per element; every special case (tail columns, blank, label, sum of squares) is a
            // warp-uniform branch around the whole piece.  bias_pad is zero-padded to whole pieces.
            const float4* bias4 = reinterpret_cast<const float4*>(p.bias_pad + col0);
            float4 bv[8];
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) bv[j4] = __ldg(bias4 + j4);
            tc::tmem_ld_wait();
            float z[32];
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              z[4 * j4 + 0] = fmaf(__uint_as_float(rr[4 * j4 + 0]), inv_w, bv[j4].x);
              z[4 * j4 + 1] = fmaf(__uint_as_float(rr[4 * j4 + 1]), inv_w, bv[j4].y);
              z[4 * j4 + 2] = fmaf(__uint_as_float(rr[4 * j4 + 2]), inv_w, bv[j4].z);
              z[4 * j4 + 3] = fmaf(__uint_as_float(rr[4 * j4 + 3]), inv_w, bv[j4].w);
            }
            if (c + 1 < npieces) {
              tc::tmem_ld32(tacc + (c + 1) * 32, rr);
            } else {
              tc::tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                for (int a = 0; a < n_arrive; ++a) {
                  if (kPair) tc::mbar_arrive_cluster(&tmem_empty[acc], 0);
                  else tc::mbar_arrive(&tmem_empty[acc]);
                }
              }
            }
            const bool tail = col0 + 32 > p.Vp;
            if (p.sumsq) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float zz = (tail && col0 + j >= p.Vp) ? 0.f : z[j];
                ssq = fmaf(zz, zz, ssq);
              }
            }
            if (tail) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j >= p.Vp) z[j] = -INFINITY;
            }
            if (p.blank >= col0 && p.blank < col0 + 32) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j == p.blank) zb = z[j];
            }
            if (__any_sync(0xffffffffu, label >= col0 && label < col0 + 32)) {
              const int jl = label - col0;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j == jl) zl = z[j];
            }
            float cm = z[0];
#pragma unroll
            for (int j = 1; j < 32; ++j) cm = fmaxf(cm, z[j]);
            // running maximum in the log2 domain: every term is 2^(z log2e - m) with the SAME rounded m, so the
            // rounding of m cancels in lse = (m + log2 s) ln 2
            const float cm2 = cm * kLog2e;
            if (cm2 > m) {
              s *= tc::ex2_approx(m - cm2);
              m = cm2;
            }
            const float nm2 = -m;   // finite: the first piece of a row always holds valid columns
            if (kMode == 3 && kZStage) {
              // keep the logits: the warp's 32 rows x 32 columns leave as four [32 x 8] fp32 boxes through two
              // alternating 32-byte-swizzled smem tiles and TMA stores.  (Row-per-lane global stores cost the LSU a
              // wavefront per 32 bytes, and a single tile exposes the TMA's read latency twice per piece.)
              const uint32_t zs = tc::smem_u32(zstage) + q * 2048;
#pragma unroll
              for (int sb = 0; sb < 4; ++sb) {
                CLASR_TRACE_WAIT(7, if (lane == 0) tc::bulk_wait_group_read1(); __syncwarp());  // the box stored from THIS tile two steps ago is out
                const uint32_t zt = zs + (sb & 1) * 1024 + lane * 32;
                const int sw = (lane >> 2) & 1;
                tc::st_shared_v4(zt + ((0 ^ sw) * 16), __float_as_uint(z[8 * sb + 0]), __float_as_uint(z[8 * sb + 1]),
                                 __float_as_uint(z[8 * sb + 2]), __float_as_uint(z[8 * sb + 3]));
                tc::st_shared_v4(zt + ((1 ^ sw) * 16), __float_as_uint(z[8 * sb + 4]), __float_as_uint(z[8 * sb + 5]),
                                 __float_as_uint(z[8 * sb + 6]), __float_as_uint(z[8 * sb + 7]));
                tc::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                  tc::tma_store_2d_hint(&tmZ, zstage + q * 2048 + (sb & 1) * 1024, col0 + 8 * sb, tile * kJM + q * 32,
                                        pol_stream);
                  tc::bulk_commit_group();
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) s += tc::ex2_approx(fmaf(z[8 * sb + j], kLog2e, nm2));
              }
            } else {
              if (kMode == 3) {  // 1-CTA variant: row-per-lane 256-bit stores
                float* zrow = p.zbuf + grow * p.ldzf + col0;
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                  uint32_t pk[8];
#pragma unroll
                  for (int j = 0; j < 8; ++j) pk[j] = __float_as_uint(z[8 * j8 + j]);
                  st_global_256(zrow + 8 * j8, pk);
                }
              }
#pragma unroll
              for (int j = 0; j < 32; ++j) s += tc::ex2_approx(fmaf(z[j], kLog2e, nm2));
            }
          }
        } else {
          // pass 2: 16-column pieces (register budget: 96 per thread in the 640-thread layout), branch-free per element;
          // every special case is a warp-uniform branch around a whole piece
          uint32_t m8_h8[4] = {0u, 0u, 0u, 0u}, m8_l8[4] = {0u, 0u, 0u, 0u};   // FP16M8: the even piece of a pair
          // the TMEM read of piece c + 1 is issued as soon as piece c has left its registers, so its latency hides
          // behind the exponentials / packing / stores of piece c; the accumulator goes back to the MMA warp after the
          // LAST read, one piece of epilogue work earlier than the end of the tile
          const int npieces = c_end;
          const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + acc * C::kBN;
          uint32_t rr[16];
          if (npieces > c_begin) tc::tmem_ld16(tacc + c_begin * 16, rr);
          acc_released = npieces > c_begin;
#pragma unroll 1
          for (int c = c_begin; c < npieces; ++c) {
            const int col0 = nt * C::kBN + c * 16;
            float gr[16];
            tc::tmem_ld_wait();
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              gr[4 * j4 + 0] = fmaf(__uint_as_float(rr[4 * j4 + 0]), inv_w, bv[j4].x);
              gr[4 * j4 + 1] = fmaf(__uint_as_float(rr[4 * j4 + 1]), inv_w, bv[j4].y);
              gr[4 * j4 + 2] = fmaf(__uint_as_float(rr[4 * j4 + 2]), inv_w, bv[j4].z);
              gr[4 * j4 + 3] = fmaf(__uint_as_float(rr[4 * j4 + 3]), inv_w, bv[j4].w);
            }
            if (c + 1 < npieces) {
              tc::tmem_ld16(tacc + (c + 1) * 16, rr);
            } else {
              tc::tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                for (int a = 0; a < n_arrive; ++a) {
                  if (kPair) tc::mbar_arrive_cluster(&tmem_empty[acc], 0);
                  else tc::mbar_arrive(&tmem_empty[acc]);
                }
              }
            }
            if (kGMode == 2) {
              // MAS importance objective: dZ = 2 z * upstream (go); nothing else to do per logit
            } else if (p.fastemit_lambda > 0.f) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                gr[j] = fmaf(fe_coef, tc::ex2_approx(fmaf(gr[j], kLog2e, fe_base2)),
                             tc::ex2_approx(fmaf(gr[j], kLog2e, base2)));
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) gr[j] = tc::ex2_approx(fmaf(gr[j], kLog2e, base2));
            }
            if (col0 + 16 > p.Vp) {  // padded tail columns of the operand
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (col0 + j >= p.Vp) gr[j] = 0.f;
            }
            if (kGMode == 1 && p.blank >= col0 && p.blank < col0 + 16) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (col0 + j == p.blank) gr[j] -= blank_sub;
            }
            if (kGMode == 1 && __any_sync(0xffffffffu, label >= col0 && label < col0 + 16)) {
              const int jl = label - col0;
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (j == jl) gr[j] -= label_sub;
            }
            if (kGMode == 1 && p.clamp > 0.f) {
#pragma unroll
              for (int j = 0; j < 16; ++j) gr[j] = fmaxf(fminf(gr[j], p.clamp), -p.clamp);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) gr[j] *= go;
            if (p.dzb && p.blank >= col0 && p.blank < col0 + 16) {
              float zb_ = 0.f;
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (col0 + j == p.blank) zb_ = gr[j];
              p.dzb[grow] = zb_;
            }
            if (p.m8) {
              // FP16M8 operands: fp16 (one 32-byte sector) + e4m3 hi8 / lo8 (16 bytes each per piece: two consecutive
              // pieces are stored together as one sector; ldz is a multiple of 32 in this mode)
              uint32_t ph[8];
              uint32_t h8w[4], l8w[4];
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                uint16_t ha, la, hb, lb;
                tc::pack_m8(gr[4 * j4], gr[4 * j4 + 1], ph[2 * j4], ha, la);
                tc::pack_m8(gr[4 * j4 + 2], gr[4 * j4 + 3], ph[2 * j4 + 1], hb, lb);
                h8w[j4] = (uint32_t)ha | ((uint32_t)hb << 16);
                l8w[j4] = (uint32_t)la | ((uint32_t)lb << 16);
              }
              st_global_256(p.dz_hi + grow * p.ldz + col0, ph);
              if (c & 1) {
                const uint32_t vh[8] = {m8_h8[0], m8_h8[1], m8_h8[2], m8_h8[3], h8w[0], h8w[1], h8w[2], h8w[3]};
                const uint32_t vl[8] = {m8_l8[0], m8_l8[1], m8_l8[2], m8_l8[3], l8w[0], l8w[1], l8w[2], l8w[3]};
                st_global_256(p.dz_h8 + grow * p.ldz + col0 - 16, vh);
                st_global_256(p.dz_l8 + grow * p.ldz + col0 - 16, vl);
              } else {
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) { m8_h8[j4] = h8w[j4]; m8_l8[j4] = l8w[j4]; }
              }
            } else {  // ldz is a multiple of 16: the piece is one whole 32-byte sector of the hi (and lo) operand row
              uint32_t ph[8], pl[8];
#pragma unroll
              for (int j2 = 0; j2 < 8; ++j2) {
                tc::pack_hi_lo(gr[2 * j2], gr[2 * j2 + 1], p.f16, ph[j2], pl[j2]);
              }
              st_global_256(p.dz_hi + grow * p.ldz + col0, ph);
              if (kTerms > 1) st_global_256(p.dz_lo + grow * p.ldz + col0, pl);
            }
            if (c + 1 < npieces) {   // next piece's bias: its registers are dead from the FFMAs above until here
              const float4* bias4 = reinterpret_cast<const float4*>(p.bias_pad + col0 + 16);
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) bv[j4] = __ldg(bias4 + j4);
            }
            // bias gradient d_b[v] = sum over cells of dZ[., v]: transpose-reduce the warp's 32 rows x 16 columns with
            // 16 shuffles (lanes 2c and 2c+1 end up with the column-(col0 + c) sum), one coalesced fp32 RED per warp
#pragma unroll
            for (int off = 16; off >= 2; off >>= 1) {
              const bool up = (lane & off) != 0;
              const int half_n = off >> 1;  // values kept after this step
#pragma unroll
              for (int j = 0; j < half_n; ++j) {
                const float send = up ? gr[j] : gr[j + half_n];
                const float recv = __shfl_xor_sync(0xffffffffu, send, off);
                gr[j] = (up ? gr[j + half_n] : gr[j]) + recv;
              }
            }
            gr[0] += __shfl_xor_sync(0xffffffffu, gr[0], 1);
            if ((lane & 1) == 0 && col0 + (lane >> 1) < p.Vp)
              atomicAdd(p.db_acc + col0 + (lane >> 1), gr[0] * inv_gs);
          }
        }
        if (!acc_released) {
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            for (int a = 0; a < n_arrive; ++a) {
              if (kPair) tc::mbar_arrive_cluster(&tmem_empty[acc], 0);  // the leader's MMA warp owns the hand-off
              else tc::mbar_arrive(&tmem_empty[acc]);
            }
          }
        }
      }
      if (kStats && valid) {
        const float lse = (m + log2f(s)) * 0.6931471805599453f;
        p.w.denom[idx] = -lse;
        const float lpb = zb - lse, lpl = label >= 0 ? zl - lse : -INFINITY;
        p.w.lp[idx] = make_float2(lpb, lpl);
        p.w.pp[idx] = lat_make_prob(lpb, lpl);
        if (p.sumsq) p.sumsq[((int64_t)b * p.T + t) * p.U1 + u] = ssq;
      }
    };
    // ============================ A producers: act(f + g) -> bf16 UMMA tiles ============================
    // Warp pw owns the 32 rows [32q, 32q+32) of the tile (q = pw & 3: the TMEM lane quarter it may write) and the
    // 32-wide K half `half = pw >> 2` of every 64-wide K block.  Compute mapping: half-warp hs handles one row at a
    // time (16 lanes x 2 consecutive k = 128 contiguous bytes of ef / eg per row), 16 row pairs per K block.  The lo
    // halves are transposed through a warp-private smem tile (no block-level barrier) into the warp's TMEM lanes.
    auto produce_tile = [&](const int tile, const int tile_it) {
      // (per-warp constants are rebuilt per tile: a helper warp must not carry them through the epilogue code)
      const int pw = (warp - kProdWarp0) & 7;   // (epilogue warps never use the producer state below)
      const int half = pw >> 2;                  // q = pw & 3 = warp & 3: the same lane quarter as in the epilogue code
      const int hs = lane >> 4, c = lane & 15;
      const uint32_t a_base = tc::smem_u32(a_smem);
      // one table per lane quarter, shared by its two K-half warps.  They drift apart in time (only the a_free waits
      // gate them), so the table of the next tile must not be written while the sibling still reads this tile's: a
      // 64-thread named barrier per quarter at every tile start.  (Without it: occasional corrupted dZ rows in pass 2a,
      // found by repeating the config-3 parity test — tests/test_gpu_repeatability.py.)
      const uint32_t tab = tc::smem_u32(rowtab) + q * 256;
      const uint32_t stg = tc::smem_u32(staging) + pw * 2048;
      // hi: byte offset of (row = 32q + rl, k = 32*half + 2c) in the SW128 K-major block, rl = (i&3) + 8(i>>2) + 4hs
      uint32_t aoff[4], soff[4];
  #pragma unroll
      for (int mth = 0; mth < 4; ++mth) {
        const int r7 = mth + 4 * hs;  // (row & 7)
        aoff[mth] = (uint32_t)((q * 32 + 4 * hs) * 128 + (((half * 4 + (c >> 2)) ^ r7) * 16) + (c & 3) * 4);
        // staging: physical row = rl ^ hs (rows rl, rl+4 of the two half-warps land in different bank halves),
        // 16-byte chunk XORed with (rl >> 1) & 3 so that the row-per-lane read-back is conflict-free
        soff[mth] = (uint32_t)((((mth ^ hs) + 4 * hs) * 64) + ((((c >> 2) ^ ((mth >> 1) | (hs << 1))) & 3) * 16) +
                               (c & 3) * 4);
      }
      const bool tile_ok = tile < total_tiles;
      if (kLoadA) {
        // Kept hidden activations: the 16-bit A blocks arrive by TMA (role 3); this warp moves the correction operands of
        // its 32 rows and K half — 32 B of e4m3 hi8 + 32 B of lo8 per row and K block (64 B of the 16-bit lo halves in the
        // three-term modes), as the kMode-4 forward stored them — from global memory into tensor memory.
        const int64_t erow = ((int64_t)tile * kJM + q * 32 + lane) * p.ldh + half * 32;
        constexpr int kPf = 3;                 // K blocks of correction operands in flight per lane (16 registers each)
        constexpr int kMaxKb = kJMaxH / kJK;   // the loops are unrolled over the largest H: register arrays, no indexing
        uint32_t v0[kPf][8], v1[kPf][8];
        auto load_kb = [&](int kb, uint32_t (&a0)[8], uint32_t (&a1)[8]) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { a0[j] = 0u; a1[j] = 0u; }
          if (kTerms == 4 && tile_ok) {
            ld_global_256(p.hid_h8 + erow + kb * kJK, a0);
            ld_global_256(p.hid_l8 + erow + kb * kJK, a1);
          } else if (kTerms == 3 && tile_ok) {
            ld_global_256(p.hid_lo + erow + kb * kJK, a0);
            ld_global_256(p.hid_lo + erow + kb * kJK + 16, a1);
          }
        };
#pragma unroll
        for (int j = 0; j < kPf; ++j)
          if (j < kblocks) load_kb(j, v0[j], v1[j]);
#pragma unroll
        for (int kb = 0; kb < kMaxKb; ++kb) {
          if (kb < kblocks) {
            if (tile_it > 0) {   // (also in the one-term mode: an arrival must not land in the previous tile's phase)
              CLASR_TRACE_WAIT(4, tc::mbar_wait(&a_free[kb], (tile_it - 1) & 1));
              tc::tc_fence_after();
            }
#ifdef CLASR_TRACE
            const long long tl0 = clock64();
#endif
            if (kTerms == 4) {
              const uint32_t t8 = tmem_base + ((uint32_t)(q * 32) << 16) + kb * (kJK / 4) + half * 8;
              tc::tmem_st8(t8 + C::kAloCol, v0[kb % kPf]);
              tc::tmem_st8(t8 + C::kAlo8Col, v1[kb % kPf]);
            } else if (kTerms == 3) {
              const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + C::kAloCol + kb * (kJK / 2) + half * 16;
              tc::tmem_st8(taddr, v0[kb % kPf]);
              tc::tmem_st8(taddr + 8, v1[kb % kPf]);
            }
            if (kTerms > 1) {
              tc::tmem_st_wait();
              tc::tc_fence_before();
            }
            __syncwarp();
            if (lane == 0) {
              if (kPair) tc::mbar_arrive_cluster(&a_ready[kb], 0);
              else tc::mbar_arrive(&a_ready[kb]);
            }
            if (kb + kPf < kblocks) load_kb(kb + kPf, v0[kb % kPf], v1[kb % kPf]);
#ifdef CLASR_TRACE
            tr_acc[9] += (unsigned long long)(clock64() - tl0);   // slot 9: operand arrival + tensor-memory store + hand-off
#endif
          }
        }
        // the next row tile's correction operands: towards L2 now, while this warp works on this tile's epilogue share
        // (half 0 asks for the hi8 / first 64 bytes of its rows' blocks, half 1 for the lo8 / second 64 bytes)
        if (kTerms > 1 && tile + tile_stride < total_tiles) {
          const int64_t nrow = ((int64_t)(tile + tile_stride) * kJM + q * 32 + lane) * p.ldh;
          const char* base = kTerms == 4 ? (const char*)(half ? p.hid_l8 : p.hid_h8) + nrow
                                         : (const char*)p.hid_lo + 2 * nrow + half * 128;
          const int nlines = kTerms == 4 ? (p.H + 127) / 128 : (2 * p.H + 255) / 256;
          const int step = kTerms == 4 ? 128 : 256;
          for (int j = 0; j < nlines; ++j) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + j * step));
        }
        return;
      }
      const int b = find_utterance(p.tile_offsets, p.B, tile_ok ? tile : 0);
      const int Tb = (int)p.act_lens[b], Ub1 = (int)p.label_lens[b] + 1;
      const int r0 = (tile - p.tile_offsets[b]) * kJM;
      {  // row table: lane l <-> row 32q + l (both K-half warps of the quarter write identical values).
        // Padding rows (tail of an utterance's last tile, null tiles of a pair) alias the utterance's last cell: their
        // A rows only have to be FINITE — every consumer of a padding row multiplies it by an exact zero (go = 0 /
        // dZ = 0) or skips it (`valid`), so the producers carry no per-row predicates at all.
        const int r = min(r0 + q * 32 + lane, Tb * Ub1 - 1);
        const int t = r / Ub1, u = r - t * Ub1;
        const uint32_t fo = (uint32_t)((b * p.T + t) * p.H);
        const uint32_t go_ = (uint32_t)((b * p.U1 + u) * p.H);
        asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");  // the sibling warp is done reading the old table
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(tab + lane * 8), "r"(fo), "r"(go_) : "memory");
      }
      __syncwarp();
      float2 fa[kBatch], ga[kBatch], fb[kBatch], gb[kBatch];
      const float s_a = p.ascale ? __ldg(p.ascale) : 1.f;
      const float* __restrict__ ef_lane = p.ef + half * 32 + 2 * c;   // this lane's two features of a row
      const float* __restrict__ eg_lane = p.eg + half * 32 + 2 * c;
      auto load_batch = [&](int kb, int batch, float2 (&fo)[kBatch], float2 (&go_)[kBatch]) {
        const uint64_t ef_kb = (uint64_t)(ef_lane + kb * kJK), eg_kb = (uint64_t)(eg_lane + kb * kJK);
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
          const int i = batch * kBatch + j;
          const int rl = (i & 3) + 8 * (i >> 2) + 4 * hs;
          const int2 o = tc::ld_shared_i2(tab + rl * 8);
          // unsigned 32-bit element offsets: ONE IMAD.WIDE.U32 per address (nvcc otherwise spends four)
          uint64_t pf, pg;
          asm("mad.wide.u32 %0, %1, 4, %2;" : "=l"(pf) : "r"(o.x), "l"(ef_kb));
          asm("mad.wide.u32 %0, %1, 4, %2;" : "=l"(pg) : "r"(o.y), "l"(eg_kb));
          fo[j] = __ldg(reinterpret_cast<const float2*>(pf));
          go_[j] = __ldg(reinterpret_cast<const float2*>(pg));
        }
      };
      auto compute_batch = [&](int kb, int batch, const float2 (&fi)[kBatch], const float2 (&gi)[kBatch]) {
        const uint32_t ablk = a_base + kb * C::kABlockBytes;
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
          const int i = batch * kBatch + j;
          float h0 = joint_combine<kAct>(fi[j].x, gi[j].x);
          float h1 = joint_combine<kAct>(fi[j].y, gi[j].y);
          if (kAct == CLASR_ACT_RELU || p.m8) {  // fp16 operands: unbounded ReLU values are brought to <= 1; FP16M8: every
            h0 *= s_a;                           // activation to < 2^14 (s_a = 1 otherwise)
            h1 *= s_a;
          }
          if (p.drop_thresh) {  // warp-uniform
            const uint32_t crow = (uint32_t)tile * kJM + q * 32 + ((i & 3) + 8 * (i >> 2)) + 4 * hs;
            const uint32_t x = drop_hash(crow * (uint32_t)(p.H >> 1) + (uint32_t)((kb * kJK + half * 32) >> 1) + c,
                                         p.drop_seed_a, p.drop_seed_b);
            h0 = (x & 0xffffu) >= p.drop_thresh ? h0 * p.drop_scale : 0.f;
            h1 = (x >> 16) >= p.drop_thresh ? h1 * p.drop_scale : 0.f;
          }
          uint32_t hw, lw;
          tc::pack_hi_lo(h0, h1, p.f16, hw, lw);
          const int rimm = (i & 3) + 8 * (i >> 2);  // compile-time part of the row index
          tc::st_shared_u32(ablk + aoff[i & 3] + rimm * 128, hw);
          if (kTerms == 4) {
            // FP16M8: the staging word of a feature pair is (e4m3 hi8 pair) | (e4m3 lo8 pair) << 16 (tc::pack_m8)
            uint32_t h16;
            uint16_t h8, l8;
            tc::pack_m8(h0, h1, h16, h8, l8);
            lw = (uint32_t)h8 | ((uint32_t)l8 << 16);
          }
          if (kTerms > 1) tc::st_shared_u32(stg + soff[i & 3] + 8 * (i >> 2) * 64, lw);
        }
      };
      load_batch(0, 0, fa, ga);
      for (int kb = 0; kb < kblocks; ++kb) {
        // software pipeline over the kNB batches of the block: the loads of batch i+1 are in flight while batch i is
        // computed (two register buffers, ping-pong; kNB is even, so a K block always starts on buffer a)
#ifdef CLASR_TRACE
        long long tp0 = clock64();
#endif
#pragma unroll
        for (int bi = 0; bi < kNB; bi += 2) {
          load_batch(kb, bi + 1, fb, gb);
          if (bi == 0 && tile_it > 0) {  // the previous row tile's last MMAs on this K block have retired
            CLASR_TRACE_WAIT(4, tc::mbar_wait(&a_free[kb], (tile_it - 1) & 1));
            tc::tc_fence_after();
          }
          compute_batch(kb, bi, fa, ga);
          if (bi + 2 < kNB) load_batch(kb, bi + 2, fa, ga);
          else if (kb + 1 < kblocks) load_batch(kb + 1, 0, fa, ga);
          compute_batch(kb, bi + 1, fb, gb);
        }
#ifdef CLASR_TRACE
        { const long long t_ = clock64(); tr_acc[8] += (unsigned long long)(t_ - tp0); tp0 = t_; }
#endif
        if (kTerms > 1) {
          // warp-private staging (row-major, swizzled) -> tensor memory: lane l owns row 32q + l of the tile
          __syncwarp();
          const uint32_t rbase = stg + (uint32_t)((lane ^ ((lane >> 2) & 1)) * 64);
          const int sw = (lane >> 1) & 3;
          uint32_t v0[8], v1[8];
          {
            const uint4 x0 = tc::ld_shared_v4(rbase + ((0 ^ sw) * 16));
            const uint4 x1 = tc::ld_shared_v4(rbase + ((1 ^ sw) * 16));
            const uint4 x2 = tc::ld_shared_v4(rbase + ((2 ^ sw) * 16));
            const uint4 x3 = tc::ld_shared_v4(rbase + ((3 ^ sw) * 16));
            v0[0] = x0.x; v0[1] = x0.y; v0[2] = x0.z; v0[3] = x0.w;
            v0[4] = x1.x; v0[5] = x1.y; v0[6] = x1.z; v0[7] = x1.w;
            v1[0] = x2.x; v1[1] = x2.y; v1[2] = x2.z; v1[3] = x2.w;
            v1[4] = x3.x; v1[5] = x3.y; v1[6] = x3.z; v1[7] = x3.w;
          }
          if (kTerms == 4) {
            // de-interleave the 16 (hi8 pair | lo8 pair) words of this lane's row into 32 hi8 bytes and 32 lo8 bytes
            // (k ascending from the least significant byte: 4 k per 32-bit TMEM column)
            uint32_t vh[8], vl[8];
#pragma unroll
            for (int i2 = 0; i2 < 4; ++i2) {
              vh[i2] = __byte_perm(v0[2 * i2], v0[2 * i2 + 1], 0x5410);
              vl[i2] = __byte_perm(v0[2 * i2], v0[2 * i2 + 1], 0x7632);
              vh[4 + i2] = __byte_perm(v1[2 * i2], v1[2 * i2 + 1], 0x5410);
              vl[4 + i2] = __byte_perm(v1[2 * i2], v1[2 * i2 + 1], 0x7632);
            }
            const uint32_t t8 = tmem_base + ((uint32_t)(q * 32) << 16) + kb * (kJK / 4) + half * 8;
            tc::tmem_st8(t8 + C::kAloCol, vh);
            tc::tmem_st8(t8 + C::kAlo8Col, vl);
            if (kEmitHid && tile_ok) {  // the e4m3 operands of the dW GEMM: 32 contiguous bytes of this lane's row each
              const int64_t e = ((int64_t)tile * kJM + q * 32 + lane) * p.ldh + kb * kJK + half * 32;
              st_global_256(p.hid_h8 + e, vh);
              st_global_256(p.hid_l8 + e, vl);
            }
          } else {
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + C::kAloCol + kb * (kJK / 2) + half * 16;
          tc::tmem_st8(taddr, v0);
          tc::tmem_st8(taddr + 8, v1);
          if (kEmitHid && tile_ok) {  // pass 2: the lo halves of this lane's row (32 k = 64 contiguous bytes) for dW
            __nv_bfloat16* dst = p.hid_lo + ((int64_t)tile * kJM + q * 32 + lane) * p.ldh + kb * kJK + half * 32;
            st_global_256(dst, v0);
            st_global_256(dst + 16, v1);
          }
          }
          tc::tmem_st_wait();
          tc::tc_fence_before();
        }
#ifdef CLASR_TRACE
        { const long long t_ = clock64(); tr_acc[9] += (unsigned long long)(t_ - tp0); tp0 = t_; }
#endif
        tc::fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the UMMA (async proxy) reads
        __syncwarp();                  // also: staging reads done before the next K block's writes
        if (lane == 0) {
          if (kPair) tc::mbar_arrive_cluster(&a_ready[kb], 0);  // the leader's MMA warp waits for both CTAs' A blocks
          else tc::mbar_arrive(&a_ready[kb]);
          if (kEmitHid) tc::mbar_arrive(&hid_ready[kb]);      // the local store warp
        }
#ifdef CLASR_TRACE
        { const long long t_ = clock64(); tr_acc[10] += (unsigned long long)(t_ - tp0); }
#endif
      }
    };
    // ============================ the tile loop of both roles ============================
    {
      int tile_it = 0;
      for (int tile = tile_first; tile < tile_end; tile += tile_stride, ++tile_it) {
        if (is_prod) produce_tile(tile, tile_it);
        if (!is_prod || kHelp) epilogue_tile(tile, tile_it, is_prod);
      }
    }
    if (kZStage && !is_prod) {  // all z boxes written before the CTA (and its shared memory) goes away
      if (lane == 0) tc::bulk_wait_group0();
      __syncwarp();
    }
  }
#ifdef CLASR_TRACE
  if (warp == 0) CLASR_TRACE_FLUSH(5);
  if (warp == 1 && leader) { CLASR_TRACE_FLUSH(0); CLASR_TRACE_FLUSH(1); CLASR_TRACE_FLUSH(2); }
  if (warp == 4) { CLASR_TRACE_FLUSH(3); CLASR_TRACE_FLUSH(7); }
  if (warp == kProdWarp0) { CLASR_TRACE_FLUSH(4); CLASR_TRACE_FLUSH(8); CLASR_TRACE_FLUSH(9); CLASR_TRACE_FLUSH(10); }
#endif
  tc::tc_fence_before();
  if (kPair) tc::cluster_sync_all();  // nobody leaves while the peer may still read its smem / signal its barriers
  else __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    if (kPair) tc::tmem_dealloc_2sm(tmem_base, 512);
    else tc::tmem_dealloc(tmem_base, 512);
  }
  CLASR_TRACE_TOTAL();
}

// ------------------------------------------------------------------------------------------------
// Backward pass 2a when the forward kept the logits (kMode 3): dZ from z in ONE streaming sweep instead of a second
// joint GEMM.  Reads z [rows_pad, ldzf] fp32, writes dZ as bf16 hi/lo GEMM operands [rows_pad, ldz], dZ[., blank]
// and the bias gradient.  A thread owns one 8-column group for the whole kernel (column sums stay in registers,
// one RED per column per block at the end); a block walks chunks of kDzRows rows whose per-row scalars
// (joint_row_grad) are computed once into shared memory.  Algorithmic bytes per lattice cell: 4 ldzf + 4 ldz.
// ------------------------------------------------------------------------------------------------
constexpr int kDzRows = 32;        // rows per chunk (one scalar set per row, computed by one warp)
// measured at B32/T250/U100/V1024 (ms in the step): (rows in flight, register cap) = (2,48) 1.33, (1,40) 1.40, (4,64) 1.41,
// (1,32) 1.42, (8,128) 1.44, (2,40) 1.46, (4,80) 1.41-1.5, (8,96) 1.51 — occupancy beats per-thread batching
#ifndef CLASR_DZ_BATCH
#define CLASR_DZ_BATCH 2
#endif
#ifndef CLASR_DZ_REGS
#define CLASR_DZ_REGS 48
#endif
constexpr int kDzBatch = CLASR_DZ_BATCH;  // rows whose loads are in flight per thread (2 x LDG.128 each)
constexpr int kDzMaxGroups = 512;  // 8-column groups per block (beyond that: blockIdx.y chunks)
constexpr int kDzMaxThreads = kDzMaxGroups + 32;

struct DzRowScalars {
  float base2[kDzRows], fe_base2[kDzRows], fe_coef[kDzRows], blank_sub[kDzRows], label_sub[kDzRows], go[kDzRows];
  int label[kDzRows];
};

// `groups` = column groups per block (a multiple of 32 is NOT required).  The LAST warp of the block computes the
// next chunk's row scalars into the other half of a double buffer while the block streams the current chunk; the
// launcher adds a warp for it when the last column warp would be more than half busy.
template <int kTerms, int kMode>
__global__ void __maxnreg__(CLASR_DZ_REGS) joint_dz_kernel(JointFwdParams p, int groups) {
  __shared__ DzRowScalars sc[2];
  constexpr float kLog2e = 1.4426950408889634f;
  const int ncg = p.ldz >> 3;
  const int cg = blockIdx.y * groups + threadIdx.x;
  const bool col_ok = (int)threadIdx.x < groups && cg < ncg;
  const int col0 = cg * 8;
  const int rows_pad = p.rows_pad_dev[0];
  const int nchunks = rows_pad / kDzRows;  // rows_pad is a multiple of 128
  const bool has_blank = p.blank >= col0 && p.blank < col0 + 8;
  const int scalar_lane = (int)threadIdx.x - ((int)blockDim.x - 32);  // >= 0: this thread belongs to the scalar warp
  float db[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) db[j] = 0.f;

  auto fill_scalars = [&](int chunk, DzRowScalars& o) {
    const int i = scalar_lane;
    const int64_t grow = (int64_t)chunk * kDzRows + i;
    const int tile = (int)(grow / kJM);
    const int b = find_utterance(p.tile_offsets, p.B, tile);
    const int Tb = (int)p.act_lens[b], Ub1 = (int)p.label_lens[b] + 1;
    const int r = (tile - p.tile_offsets[b]) * kJM + (int)(grow % kJM);
    const bool valid = r < Tb * Ub1;
    const int t = valid ? r / Ub1 : 0;
    const int u = valid ? r - t * Ub1 : 0;
    const RowGrad rg = joint_row_grad<kMode>(p, b, t, u, Tb, Ub1, valid);
    o.base2[i] = rg.base2; o.fe_base2[i] = rg.fe_base2; o.fe_coef[i] = rg.fe_coef;
    o.blank_sub[i] = rg.blank_sub; o.label_sub[i] = rg.label_sub; o.go[i] = rg.go;
    o.label[i] = (valid && u < Ub1 - 1) ? (int)p.labels[(int64_t)b * (p.U1 - 1) + u] : -1;
  };

  if (scalar_lane >= 0 && (int)blockIdx.x < nchunks) fill_scalars(blockIdx.x, sc[0]);
  __syncthreads();
  int it = 0;
  for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x, ++it) {
    const DzRowScalars& cur = sc[it & 1];
    if (scalar_lane >= 0 && chunk + (int)gridDim.x < nchunks) fill_scalars(chunk + gridDim.x, sc[(it + 1) & 1]);
    if (col_ok) {
      const int64_t grow0 = (int64_t)chunk * kDzRows;
#pragma unroll 1
      for (int r0 = 0; r0 < kDzRows; r0 += kDzBatch) {
        float4 za[kDzBatch], zb[kDzBatch];
#pragma unroll
        for (int j = 0; j < kDzBatch; ++j) {  // all loads of the batch first: they do not move past the stores below
          const float4* zp = reinterpret_cast<const float4*>(p.zbuf + (grow0 + r0 + j) * p.ldzf + col0);
          za[j] = ld_stream(zp);
          zb[j] = ld_stream(zp + 1);
        }
#pragma unroll
        for (int jr = 0; jr < kDzBatch; ++jr) {
          const int r = r0 + jr;
          const int64_t grow = grow0 + r;
          float gr[8] = {za[jr].x, za[jr].y, za[jr].z, za[jr].w, zb[jr].x, zb[jr].y, zb[jr].z, zb[jr].w};
          const float go = cur.go[r];
          if (kMode == 1) {
            const float base2 = cur.base2[r];
            if (p.fastemit_lambda > 0.f) {
              const float fe_coef = cur.fe_coef[r], fe_base2 = cur.fe_base2[r];
#pragma unroll
              for (int j = 0; j < 8; ++j)
                gr[j] = fmaf(fe_coef, tc::ex2_approx(fmaf(gr[j], kLog2e, fe_base2)),
                             tc::ex2_approx(fmaf(gr[j], kLog2e, base2)));
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) gr[j] = tc::ex2_approx(fmaf(gr[j], kLog2e, base2));
            }
          }
          if (col0 + 8 > p.Vp) {  // padded tail columns of the operand
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (col0 + j >= p.Vp) gr[j] = 0.f;
          }
          if (kMode == 1) {
            if (has_blank) {
              const float bs = cur.blank_sub[r];
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (col0 + j == p.blank) gr[j] -= bs;
            }
            const int jl = cur.label[r] - col0;
            if (jl >= 0 && jl < 8) {
              const float ls = cur.label_sub[r];
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (j == jl) gr[j] -= ls;
            }
            if (p.clamp > 0.f) {
#pragma unroll
              for (int j = 0; j < 8; ++j) gr[j] = fmaxf(fminf(gr[j], p.clamp), -p.clamp);
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            gr[j] *= go;
            db[j] += gr[j];
          }
          if (p.dzb && has_blank) {
            float zb_ = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (col0 + j == p.blank) zb_ = gr[j];
            p.dzb[grow] = zb_;
          }
          uint4 ph, pl;
          uint32_t* phw = reinterpret_cast<uint32_t*>(&ph);
          uint32_t* plw = reinterpret_cast<uint32_t*>(&pl);
          if (p.m8) {   // FP16M8: fp16 + e4m3 hi8 / lo8 (8 bytes each; a warp's 32 column groups are 256 contiguous bytes)
            uint16_t h8[4], l8[4];
#pragma unroll
            for (int j2 = 0; j2 < 4; ++j2) tc::pack_m8(gr[2 * j2], gr[2 * j2 + 1], phw[j2], h8[j2], l8[j2]);
            *reinterpret_cast<uint4*>(p.dz_hi + grow * p.ldz + col0) = ph;
            *reinterpret_cast<uint2*>(p.dz_h8 + grow * p.ldz + col0) =
                make_uint2((uint32_t)h8[0] | ((uint32_t)h8[1] << 16), (uint32_t)h8[2] | ((uint32_t)h8[3] << 16));
            *reinterpret_cast<uint2*>(p.dz_l8 + grow * p.ldz + col0) =
                make_uint2((uint32_t)l8[0] | ((uint32_t)l8[1] << 16), (uint32_t)l8[2] | ((uint32_t)l8[3] << 16));
          } else {
#pragma unroll
            for (int j2 = 0; j2 < 4; ++j2) {
              tc::pack_hi_lo(gr[2 * j2], gr[2 * j2 + 1], p.f16, phw[j2], plw[j2]);
            }
            *reinterpret_cast<uint4*>(p.dz_hi + grow * p.ldz + col0) = ph;
            if (kTerms > 1) *reinterpret_cast<uint4*>(p.dz_lo + grow * p.ldz + col0) = pl;
          }
        }
      }
    }
    __syncthreads();  // the next chunk's scalars are complete; this chunk's are no longer read
  }
  if (col_ok) {
    const float inv_s = p.gscale ? __ldg(p.gscale + 1) : 1.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (col0 + j < p.Vp) atomicAdd(p.db_acc + col0 + j, db[j] * inv_s);
  }
}

template <int kTerms, int kMode>
static int launch_joint_dz(const JointFwdParams& p, cudaStream_t s) {
  const int ncg = p.ldz / 8;
  const int ychunks = (ncg + kDzMaxGroups - 1) / kDzMaxGroups;
  const int groups = (ncg + ychunks - 1) / ychunks;
  int threads = (groups + 31) / 32 * 32;
  const int last = groups - (threads - 32);            // column lanes of the last warp
  if (last > 16) threads += 32;                        // dedicated scalar warp
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, joint_dz_kernel<kTerms, kMode>, threads, 0);
  if (per_sm < 1) per_sm = 1;
  int gx = kNumSMs * per_sm / ychunks;
  if (gx < 1) gx = 1;
  joint_dz_kernel<kTerms, kMode><<<dim3(gx, ychunks), threads, 0, s>>>(p, groups);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("joint_dz launch: %s", cudaGetErrorString(e));
    return CLASR_STATUS_CUDA_ERROR;
  }
  return CLASR_STATUS_SUCCESS;
}

// ------------------------------------------------------------------------------------------------
// Pass-2 tail: d_pre = dHid * act'(f+g), reduced over u (-> d_f) and over t (-> d_g).
// dHid is in compact tile-row order: row(b,t,u) = tile_offsets[b]*128 + t*U_b1 + u.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float joint_act_grad(float pre, int act) {
  if (act == CLASR_ACT_RELU) return pre > 0.f ? 1.f : 0.f;
  if (act == CLASR_ACT_SIGMOID) {
    const float sg = __fdividef(1.f, 1.f + __expf(-pre));
    return sg * (1.f - sg);
  }
  const float h = joint_act(pre, CLASR_ACT_TANH);
  return 1.f - h * h;
}

// mode 0: grid (T, B) -> d_f[b,t,:] = sum_u ;  mode 1: grid (U1, B) -> d_g[b,u,:] = sum_t
__global__ void __launch_bounds__(256) joint_dfg_kernel(const float* __restrict__ dhid, const float* __restrict__ f,
                                                        const float* __restrict__ g,
                                                        const int64_t* __restrict__ act_lens,
                                                        const int64_t* __restrict__ label_lens,
                                                        const int* __restrict__ tile_offsets, int T, int U1, int H,
                                                        int activation, int mode, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int i = blockIdx.x;  // t (mode 0) or u (mode 1)
  const int Tb = (int)act_lens[b], Ub1 = (int)label_lens[b] + 1;
  const int64_t row0 = (int64_t)tile_offsets[b] * kJM;
  float* __restrict__ o = out + ((int64_t)b * (mode == 0 ? T : U1) + i) * H;
  const bool live = mode == 0 ? (i < Tb) : (i < Ub1 && Tb > 0);
  for (int k = threadIdx.x * 4; k < H; k += blockDim.x * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
      if (mode == 0) {
        const float4 fv = *reinterpret_cast<const float4*>(f + ((int64_t)b * T + i) * H + k);
        for (int u = 0; u < Ub1; ++u) {
          const float4 gv = __ldg(reinterpret_cast<const float4*>(g + ((int64_t)b * U1 + u) * H + k));
          const float4 d = __ldg(reinterpret_cast<const float4*>(dhid + (row0 + (int64_t)i * Ub1 + u) * H + k));
          acc.x = fmaf(d.x, joint_act_grad(fv.x + gv.x, activation), acc.x);
          acc.y = fmaf(d.y, joint_act_grad(fv.y + gv.y, activation), acc.y);
          acc.z = fmaf(d.z, joint_act_grad(fv.z + gv.z, activation), acc.z);
          acc.w = fmaf(d.w, joint_act_grad(fv.w + gv.w, activation), acc.w);
        }
      } else {
        const float4 gv = *reinterpret_cast<const float4*>(g + ((int64_t)b * U1 + i) * H + k);
        for (int t = 0; t < Tb; ++t) {
          const float4 fv = __ldg(reinterpret_cast<const float4*>(f + ((int64_t)b * T + t) * H + k));
          const float4 d = __ldg(reinterpret_cast<const float4*>(dhid + (row0 + (int64_t)t * Ub1 + i) * H + k));
          acc.x = fmaf(d.x, joint_act_grad(fv.x + gv.x, activation), acc.x);
          acc.y = fmaf(d.y, joint_act_grad(fv.y + gv.y, activation), acc.y);
          acc.z = fmaf(d.z, joint_act_grad(fv.z + gv.z, activation), acc.z);
          acc.w = fmaf(d.w, joint_act_grad(fv.w + gv.w, activation), acc.w);
        }
      }
    }
    *reinterpret_cast<float4*>(o + k) = acc;
  }
}

// Single-pass variant: a block owns (utterance b, 32-wide slice of H, chunk of 32 time steps) and reads each dHid
// element ONCE (128-byte row segments, 8 loads in flight per warp, 8 blocks per SM): d_f[t] accumulates in registers
// over the consecutive u rows, d_g[u] in a shared-memory tile via shared atomics (8 warps = 8 different t), flushed
// with one global RED per element per block (d_g is zeroed first).  act' comes from the scaled pre-activations:
// tanh' = 4 r (1 - r), sigmoid' = r (1 - r) with r = 1 / (1 + 2^(sf+sg)); relu' = [f + g > 0].
constexpr int kDfgTChunk = 32;
template <int kAct>
__global__ void __launch_bounds__(256) joint_dfg_fused_kernel(const float* __restrict__ dhid,
                                                              const float* __restrict__ ef,
                                                              const float* __restrict__ eg,
                                                              const int64_t* __restrict__ act_lens,
                                                              const int64_t* __restrict__ label_lens,
                                                              const int* __restrict__ tile_offsets, int T, int U1, int H,
                                                              float* __restrict__ d_f, float* __restrict__ d_g,
                                                              const float* __restrict__ dzb,
                                                              float* __restrict__ d_w_blank,
                                                              const float* __restrict__ w_blank,
                                                              const float* __restrict__ gscale,
                                                              const float* __restrict__ wscale, uint32_t drop_thresh,
                                                              uint32_t drop_seed_a, uint32_t drop_seed_b,
                                                              float drop_scale) {
  extern __shared__ float sm_dfg[];
  const int b = blockIdx.y, k0 = blockIdx.x * 32, t_begin = blockIdx.z * kDfgTChunk;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int Tb = (int)act_lens[b], Ub1 = (int)label_lens[b] + 1;
  const int t_end = min(T, t_begin + kDfgTChunk);
  const int64_t row0 = (int64_t)tile_offsets[b] * kJM;
  float* eg_s = sm_dfg;            // [U1][32]
  float* dg_s = sm_dfg + U1 * 32;  // [U1][32]
  const bool live = t_begin < Tb;  // block-uniform
  if (live) {
    for (int i = threadIdx.x; i < U1 * 32; i += blockDim.x) {
      const int u = i >> 5;
      eg_s[i] = u < Ub1 ? __ldg(eg + ((int64_t)b * U1 + u) * H + k0 + (i & 31)) : 0.f;
      dg_s[i] = 0.f;
    }
    __syncthreads();
  }
  // (hidden activation, its derivative) from the scaled pre-activations
  auto act_pair = [](float a, float bb, float& h, float& dh) {
    if (kAct == CLASR_ACT_RELU) {
      const float x = a + bb;
      h = fmaxf(x, 0.f);
      dh = x > 0.f ? 1.f : 0.f;
      return;
    }
    const float r = tc::rcp_approx(1.f + tc::ex2_approx(a + bb));
    const float t1 = fmaf(-r, r, r);
    if (kAct == CLASR_ACT_TANH) { h = fmaf(-2.f, r, 1.f); dh = 4.f * t1; }
    else { h = r; dh = t1; }
  };
  // dropout between the activation and the output layer: hid = mask * act / (1-p), d hid / d pre = mask * act' / (1-p)
  const uint32_t kk = (uint32_t)(k0 + lane);
  auto drop = [&](int64_t row, float& h, float& dh) {
    const uint32_t x = drop_hash((uint32_t)row * (uint32_t)(H >> 1) + (kk >> 1), drop_seed_a, drop_seed_b);
    const float m = ((kk & 1u) ? (x >> 16) : (x & 0xffffu)) >= drop_thresh ? drop_scale : 0.f;
    h *= m;
    dh *= m;
  };
  // dW[blank, k] = sum over cells of dZ[cell, blank] * hid[cell, k]: the blank is the (V+1)-th class, a 1025th GEMM row
  // that would cost a whole extra 256-row tile in the dW GEMM; here it is one FMA per element on data already in flight
  float wb = 0.f;
  // ... and the blank column of the dHid GEMM (K = V instead of V+1): dHid[cell, k] += dZ[cell, blank] * W[blank, k]
  // fp16 operands: dZ (hence dzb) arrives pre-scaled by S, dHid by S * Sw (W_out is split pre-scaled by Sw)
  const float s_w = wscale ? __ldg(wscale) : 1.f;
  const float wbk = (dzb && live) ? __ldg(w_blank + k0 + lane) * s_w : 0.f;
  const float inv_s = gscale ? __ldg(gscale + 1) : 1.f;
  const float inv_sw = inv_s * (wscale ? __ldg(wscale + 1) : 1.f);
  for (int t = t_begin + warp; t < t_end; t += nw) {
    float df = 0.f;
    if (t < Tb) {
      const float fv = __ldg(ef + ((int64_t)b * T + t) * H + k0 + lane);
      const int64_t r0 = row0 + (int64_t)t * Ub1;
      const float* dp = dhid + r0 * H + k0 + lane;
      int u = 0;
      for (; u + 8 <= Ub1; u += 8) {
        float d[8], zb[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = ld_stream1(dp + (int64_t)(u + j) * H);
        if (dzb) {
#pragma unroll
          for (int j = 0; j < 8; ++j) zb[j] = __ldg(dzb + r0 + u + j);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float h, dh;
          act_pair(fv, eg_s[(u + j) * 32 + lane], h, dh);
          if (drop_thresh) drop(r0 + u + j, h, dh);
          const float pv = (dzb ? fmaf(zb[j], wbk, d[j]) : d[j]) * (dh * inv_sw);
          df += pv;
          atomicAdd(dg_s + (u + j) * 32 + lane, pv);
          if (dzb) wb = fmaf(zb[j], h, wb);
        }
      }
      for (; u < Ub1; ++u) {
        float h, dh;
        act_pair(fv, eg_s[u * 32 + lane], h, dh);
        if (drop_thresh) drop(r0 + u, h, dh);
        const float zb1 = dzb ? __ldg(dzb + r0 + u) : 0.f;
        const float pv = fmaf(zb1, wbk, ld_stream1(dp + (int64_t)u * H)) * (dh * inv_sw);
        df += pv;
        atomicAdd(dg_s + u * 32 + lane, pv);
        if (dzb) wb = fmaf(zb1, h, wb);
      }
    }
    d_f[((int64_t)b * T + t) * H + k0 + lane] = df;
  }
  if (dzb && live) atomicAdd(d_w_blank + k0 + lane, wb * inv_s);
  if (live) {
    __syncthreads();
    for (int i = threadIdx.x; i < Ub1 * 32; i += blockDim.x)
      atomicAdd(d_g + ((int64_t)b * U1 + (i >> 5)) * H + k0 + (i & 31), dg_s[i]);
  }
}

// TMA variant of the single-pass kernel (the default for U+1 <= 104).  The kernel above is issue-bound on its shared-memory
// float atomics (compare-and-swap loops in SASS, 36 instructions per element, ncu r02g) and keeps only 8 loads of 128 B per
// warp in flight.  Here a block is still (utterance b, 32-wide slice of H, chunk of 32 time steps), but
//   * one elected lane of a ninth warp streams the block's dHid rows of time step t — (U_b+1) consecutive compact rows x 128
//     bytes, ONE TMA box — through a ring of shared-memory stages, several time steps ahead of the consumers (full / empty
//     mbarriers, no block barrier in the loop): the memory parallelism no longer costs registers or issue slots;
//   * consumer warp w OWNS the prediction rows u in [13 w, 13 w + 13): their eg values and d_g partial sums live in registers
//     for the whole chunk, d_f[t] is the sum of the warps' partials (plain shared-memory stores, added up once at the end);
//   * the element loop is straight-line (dropout is a template flag) so that the 13 EX2 -> RCP chains of a time step overlap.
// Measured at B32/T250/U100/H640: 0.86 -> 0.58 ms (3.6 TB/s, 0.54 of the HBM peak); neither the ring depth nor a third block
// per SM changes that, so what is left is the access pattern itself (128-byte row segments at a 2.5 KB stride).
constexpr int kDfgUW = 13;
constexpr int kDfgConsWarps = 8;
constexpr int kDfgStages = 3;   // measured: 2, 3, 4 and 6 stages within noise of each other (0.57-0.62 ms)
template <int kAct, bool kDrop>
__global__ void __launch_bounds__(32 * (kDfgConsWarps + 1), 2)
joint_dfg_tma_kernel(const __grid_constant__ CUtensorMap tmD, const float* __restrict__ ef, const float* __restrict__ eg,
                     const int64_t* __restrict__ act_lens, const int64_t* __restrict__ label_lens,
                     const int* __restrict__ tile_offsets, int T, int U1, int H, float* __restrict__ d_f,
                     float* __restrict__ d_g, const float* __restrict__ dzb, float* __restrict__ d_w_blank,
                     const float* __restrict__ w_blank, const float* __restrict__ gscale, const float* __restrict__ wscale,
                     uint32_t drop_thresh, uint32_t drop_seed_a, uint32_t drop_seed_b, float drop_scale, int box_rows) {
  extern __shared__ uint8_t sm_dfg_raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)sm_dfg_raw + 127) & ~(uintptr_t)127);
  const int stage_floats = box_rows * 32;
  float* ring = (float*)sm;                                              // [kDfgStages][box_rows][32]
  float* df_s = ring + kDfgStages * stage_floats;                        // [kDfgTChunk][kDfgConsWarps][32]
  uint64_t* full = (uint64_t*)(df_s + kDfgTChunk * kDfgConsWarps * 32);  // [kDfgStages]
  uint64_t* empty = full + kDfgStages;                                   // [kDfgStages]
  const int b = blockIdx.y, k0 = blockIdx.x * 32, t_begin = blockIdx.z * kDfgTChunk;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Tb = (int)act_lens[b], Ub1 = (int)label_lens[b] + 1;
  const int t_stop = min(T, t_begin + kDfgTChunk);     // rows of d_f this block writes (zeros beyond Tb)
  const int t_end = min(Tb, t_stop);                   // time steps with lattice cells
  const int64_t row0 = (int64_t)tile_offsets[b] * kJM;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kDfgStages; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], kDfgConsWarps); }
    tc::fence_barrier_init();
  }
  __syncthreads();
  float wb = 0.f;
  const uint32_t kk = (uint32_t)(k0 + lane);
  const float inv_s = gscale ? __ldg(gscale + 1) : 1.f;
  if (warp == kDfgConsWarps) {
    // ---- TMA producer: one box per time step
    if (tc::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        tc::mbar_wait(&empty[stage], phase ^ 1);
        tc::mbar_expect_tx(&full[stage], (uint32_t)stage_floats * 4u);
        tc::tma_load_2d(ring + stage * stage_floats, &tmD, &full[stage], k0, (int)(row0 + (int64_t)t * Ub1));
        if (++stage == kDfgStages) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ---- consumers
    const int u0 = warp * kDfgUW;
    const int nu = min(kDfgUW, Ub1 - u0);                // prediction rows this warp owns (<= 0: none)
    float egr[kDfgUW], dg[kDfgUW];
#pragma unroll
    for (int j = 0; j < kDfgUW; ++j) {
      egr[j] = j < nu ? __ldg(eg + ((int64_t)b * U1 + u0 + j) * H + kk) : 0.f;
      dg[j] = 0.f;
    }
    const float s_w = wscale ? __ldg(wscale) : 1.f;
    const float wbk = dzb ? __ldg(w_blank + kk) * s_w : 0.f;
    const float inv_sw = inv_s * (wscale ? __ldg(wscale + 1) : 1.f);
    // per time step: f value of this lane's feature, dZ[., blank] of the warp's rows (lane j <-> row u0 + j) — one step ahead
    auto side_loads = [&](int t, float& fv, float& zz) {
      fv = __ldg(ef + ((int64_t)b * T + t) * H + kk);
      zz = (dzb && lane < nu) ? __ldg(dzb + row0 + (int64_t)t * Ub1 + u0 + lane) : 0.f;
    };
    float fv_n = 0.f, zz_n = 0.f;
    if (t_begin < t_end) side_loads(t_begin, fv_n, zz_n);
    int stage = 0;
    uint32_t phase = 0;
    for (int t = t_begin; t < t_end; ++t) {
      const float fv = fv_n, zzv = zz_n;
      if (t + 1 < t_end) side_loads(t + 1, fv_n, zz_n);
      const int64_t r0 = row0 + (int64_t)t * Ub1 + u0;
      tc::mbar_wait(&full[stage], phase);
      const float* st = ring + stage * stage_floats + u0 * 32 + lane;
      float dd[kDfgUW];
#pragma unroll
      for (int j = 0; j < kDfgUW; ++j) dd[j] = j < nu ? st[j * 32] : 0.f;
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&empty[stage]);     // the stage is in registers: hand it back
      float df = 0.f;
#pragma unroll
      for (int j = 0; j < kDfgUW; ++j) {
        float h, dh;
        if (kAct == CLASR_ACT_RELU) {
          const float x = fv + egr[j];
          h = fmaxf(x, 0.f);
          dh = x > 0.f ? 1.f : 0.f;
        } else {
          const float r = tc::rcp_approx(1.f + tc::ex2_approx(fv + egr[j]));
          const float t1 = fmaf(-r, r, r);
          if (kAct == CLASR_ACT_TANH) { h = fmaf(-2.f, r, 1.f); dh = 4.f * t1; }
          else { h = r; dh = t1; }
        }
        if (kDrop) {
          const uint32_t x = drop_hash((uint32_t)(r0 + j) * (uint32_t)(H >> 1) + (kk >> 1), drop_seed_a, drop_seed_b);
          const float m = ((kk & 1u) ? (x >> 16) : (x & 0xffffu)) >= drop_thresh ? drop_scale : 0.f;
          h *= m;
          dh *= m;
        }
        // rows this warp does not own carry d = zb = 0 -> pv = 0 (dh is finite for every input)
        const float zj = __shfl_sync(0xffffffffu, zzv, j);
        const float pv = fmaf(zj, wbk, dd[j]) * (dh * inv_sw);
        df += pv;
        dg[j] += pv;
        wb = fmaf(zj, h, wb);
      }
      df_s[((t - t_begin) * kDfgConsWarps + warp) * 32 + lane] = df;
      if (++stage == kDfgStages) { stage = 0; phase ^= 1; }
    }
    if (t_begin < t_end) {
#pragma unroll
      for (int j = 0; j < kDfgUW; ++j)
        if (j < nu) atomicAdd(d_g + ((int64_t)b * U1 + u0 + j) * H + kk, dg[j]);
    }
  }
  __syncthreads();
  // d_f[b, t, k0 .. k0+32) = sum over the warps' partials; frames beyond T_b get exact zeros
  for (int i = threadIdx.x; i < (t_stop - t_begin) * 32; i += blockDim.x) {
    const int tl = i >> 5, l = i & 31;
    float acc = 0.f;
    if (t_begin + tl < t_end)
      for (int w = 0; w < kDfgConsWarps; ++w) acc += df_s[(tl * kDfgConsWarps + w) * 32 + l];
    d_f[((int64_t)b * T + t_begin + tl) * H + k0 + l] = acc;
  }
  if (dzb && t_begin < t_end) {   // blank row of dW: one RED per column and block
    __syncthreads();
    if (warp < kDfgConsWarps) df_s[warp * 32 + lane] = wb;
    __syncthreads();
    if (warp == 0) {
      float acc = 0.f;
      for (int w = 0; w < kDfgConsWarps; ++w) acc += df_s[w * 32 + lane];
      atomicAdd(d_w_blank + kk, acc * inv_s);
    }
  }
}

int launch_gemm_tc(const void* A_hi, const void* A_lo, int64_t lda, int a_mn, const void* B_hi, const void* B_lo,
                   int64_t ldb, int b_mn, int M, int N, int K, float* C, int64_t ldc, int precision, int atomic_add,
                   int k_splits, cudaStream_t s, const int* m_dev, const int* k_dev, const float* bias = nullptr,
                   const float* alpha_dev = nullptr, const float* alpha_dev2 = nullptr, const void* A_l8 = nullptr,
                   const void* B_l8 = nullptr);

// ------------------------------------------------------------------------------------------------
// workspace layout of the fused path
// ------------------------------------------------------------------------------------------------
struct JointWs {
  void* lattice;
  void* w_hi;
  void* w_lo;
  void* w_h8;          // FP16M8 only
  void* w_l8;
  int* tile_offsets;   // [B+1], then [1] rows_pad
  float* gscale;       // [2] fp16 operands: power-of-two pre-scale of dZ and its inverse
  float* bias_pad;     // [round_up(Vp,32)+32]
  float* ef;           // [B,T,H]  scaled pre-activation c*f (tanh / sigmoid; relu reads f directly)
  float* eg;           // [B,U1,H]
  int vp_pad;
  size_t total;
};

// Scratch of the backward pass (caller-allocated, reusable across steps).  Sized for the worst case (no ragged
// savings); the kernels only touch the first rows_pad rows, a count that stays on the device.
struct JointBwdScratch {
  void* dz_hi; void* dz_lo;     // [rows_cap, ldz] bf16
  void* hid_hi; void* hid_lo;   // [rows_cap, ldh] bf16
  float* dhid;                  // [rows_cap, H]
  float* dzb;                   // [rows_cap]
  int64_t rows_cap;
  int ldz, ldh;
  size_t total;
};

static inline JointBwdScratch joint_bwd_scratch_carve(void* base, int B, int T, int U1, int H, int Vp, int precision) {
  JointBwdScratch sc;
  const bool x3 = prec_x3(precision);
  sc.rows_cap = (int64_t)B * ((((int64_t)T * U1) + kJM - 1) / kJM) * kJM;
  // FP16M8: the one-byte operand rows must start on 32-byte boundaries too (256-bit stores of two 16-column pieces)
  sc.ldz = prec_m8(precision) ? (Vp + 31) / 32 * 32 : (Vp + 15) / 16 * 16;
  sc.ldh = H;
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) { void* r = p + off; off += (bytes + 255) / 256 * 256; return r; };
  sc.dz_hi = take((size_t)sc.rows_cap * sc.ldz * 2);
  sc.dz_lo = x3 ? take((size_t)sc.rows_cap * sc.ldz * 2) : sc.dz_hi;
  sc.hid_hi = take((size_t)sc.rows_cap * sc.ldh * 2);
  sc.hid_lo = x3 ? take((size_t)sc.rows_cap * sc.ldh * 2) : sc.hid_hi;
  sc.dhid = (float*)take((size_t)sc.rows_cap * H * 4);
  sc.dzb = (float*)take((size_t)sc.rows_cap * 4);
  sc.total = off;
  return sc;
}

// What a kMode-3 forward leaves behind for the backward: the logits and the bf16 hi/lo hidden activations (caller-
// allocated, lives from the forward call to the backward call).
struct JointStash {
  float* z;                     // [rows_cap, ldzf]
  void* hid_hi; void* hid_lo;   // [rows_cap, H] bf16
  int64_t rows_cap;
  int ldzf;
  size_t total;
};

static inline JointStash joint_stash_carve(void* base, int B, int T, int U1, int H, int Vp, int precision,
                                           bool hid_only = false) {
  JointStash st;
  const bool x3 = prec_x3(precision);
  st.rows_cap = (int64_t)B * ((((int64_t)T * U1) + kJM - 1) / kJM) * kJM;
  st.ldzf = (Vp + 31) / 32 * 32;
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) { void* r = p + off; off += (bytes + 255) / 256 * 256; return r; };
  st.z = hid_only ? nullptr : (float*)take((size_t)st.rows_cap * st.ldzf * 4);
  st.hid_hi = take((size_t)st.rows_cap * H * 2);
  st.hid_lo = x3 ? take((size_t)st.rows_cap * H * 2) : st.hid_hi;
  st.total = off;
  return st;
}

static inline JointWs joint_ws_carve(void* base, int B, int T, int U1, int H, int Vp, int precision) {
  JointWs j;
  char* p = (char*)base;
  size_t off = 0;
  j.lattice = p + off;
  off += lattice_ws_bytes(B, T, U1);
  j.vp_pad = (Vp + 15) / 16 * 16;
  const size_t wbytes = ((size_t)j.vp_pad * H * 2 + 255) / 256 * 256;
  j.w_hi = p + off; off += wbytes;
  j.w_lo = p + off; off += prec_x3(precision) ? wbytes : 0;
  j.w_h8 = p + off;                       // FP16M8: e4m3 hi8 | lo8 of W_out, [2 vp_pad, H] bytes (ONE tensor map)
  j.w_l8 = p + off + (size_t)j.vp_pad * H;
  off += prec_m8(precision) ? wbytes : 0;
  j.tile_offsets = (int*)(p + off);
  off += ((size_t)(B + 2) * sizeof(int) + 255) / 256 * 256;
  j.gscale = (float*)(p + off);
  off += 256;
  j.bias_pad = (float*)(p + off);
  off += ((size_t)((Vp + 31) / 32 * 32 + 32) * sizeof(float) + 255) / 256 * 256;
  j.ef = (float*)(p + off);
  off += ((size_t)B * T * H * sizeof(float) + 255) / 256 * 256;
  j.eg = (float*)(p + off);
  off += ((size_t)B * U1 * H * sizeof(float) + 255) / 256 * 256;
  j.total = off;
  return j;
}

static int set_dropout(JointFwdParams& p, float dropout_p, uint64_t seed) {
  CLASR_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "joint: dropout_p must be in [0, 1)");
  p.drop_thresh = (uint32_t)lrintf(dropout_p * 65536.f);   // p quantised to 1/65536
  if (p.drop_thresh > 65535u) p.drop_thresh = 65535u;
  p.drop_scale = p.drop_thresh ? 65536.f / (65536.f - (float)p.drop_thresh) : 1.f;
  p.drop_seed_a = (uint32_t)seed;
  p.drop_seed_b = (uint32_t)(seed >> 32);
  return CLASR_STATUS_SUCCESS;
}

// CLASR_PREC_FP16M8 operand scales: headroom factor that moves a "max -> ~1" scale to "max -> just below 2^14", and the
// activation scale (bounded activations x dropout rescale 1 / (1 - p) -> below 2^14)
static inline float m8_headroom(bool m8) { return m8 ? ldexpf(1.f, -14) : 1.f; }
static inline float m8_act_scale(bool m8, float dropout_p) {
  if (!m8) return 1.f;
  int e = 0;
  if (dropout_p > 0.f) frexpf(1.f / (1.f - dropout_p), &e);   // 1/(1-p) = f 2^e, f in [0.5, 1)
  return ldexpf(1.f, 14 - e);
}

// CLASR_JOINT_PAIR=0/1 selects the 1-CTA / CTA-pair variant (default: pairs)
static bool joint_use_pair() {
  const char* e = getenv("CLASR_JOINT_PAIR");  // read per call: the tests toggle it
  return e ? atoi(e) != 0 : true;
}

template <int kTerms, int kMode, int kAct, int kPair, int kWide>
static int launch_joint_variant(int H, const CUtensorMap& tw_hi, const CUtensorMap& tw_lo, const CUtensorMap& t_hid,
                                const CUtensorMap& t_z, const JointFwdParams& p, cudaStream_t s) {
  const int smem = JointCfg<kTerms, kPair, kMode == 3>::smem_bytes(H);
  auto kern = joint_fwd_kernel<kTerms, kMode, kAct, kPair, kWide>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kNumSMs & ~1);
  cfg.blockDim = dim3(kWide ? kJThreadsWide : kJThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kPair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tw_hi, tw_lo, t_hid, t_z, p);
  if (e != cudaSuccess) {
    set_error("joint kernel launch: %s", cudaGetErrorString(e));
    return CLASR_STATUS_CUDA_ERROR;
  }
  return CLASR_STATUS_SUCCESS;
}

template <int kTerms, int kMode>
static int launch_joint_kernel(int activation, int H, const CUtensorMap& tw_hi, const CUtensorMap& tw_lo,
                               const CUtensorMap& t_hid, const CUtensorMap& t_z, const JointFwdParams& p,
                               cudaStream_t s) {
  // pass 2 (kMode 1 / 2) runs the 640-thread layout with two epilogue warpgroups unless CLASR_JOINT_WIDE=0
  constexpr int kW = (kMode == 1 || kMode == 2 || kMode >= 5) ? 1 : 0;
#ifdef CLASR_FAST_BUILD
  // developer builds (20 s instead of 3.5 min of ptxas): FP16M8, tanh, CTA pairs, the default thread layouts only
  if constexpr (kTerms == 4 && (kMode == 0 || kMode == 1 || kMode == 4 || kMode == 5)) {
    if (activation == CLASR_ACT_TANH && joint_use_pair())
      return launch_joint_variant<kTerms, kMode, CLASR_ACT_TANH, 1, kW>(H, tw_hi, tw_lo, t_hid, t_z, p, s);
  }
  set_error("joint kernel: this variant is not part of a CLASR_FAST_BUILD library");
  return CLASR_STATUS_INVALID_VALUE;
#else
  if constexpr (kMode >= 4) {
    // kept-hidden-activation modes: built as CTA pairs (and, for the gradient passes, the 640-thread layout) only
    if (activation == CLASR_ACT_RELU) return launch_joint_variant<kTerms, kMode, CLASR_ACT_RELU, 1, kW>(H, tw_hi, tw_lo, t_hid, t_z, p, s);
    if (activation == CLASR_ACT_SIGMOID) return launch_joint_variant<kTerms, kMode, CLASR_ACT_SIGMOID, 1, kW>(H, tw_hi, tw_lo, t_hid, t_z, p, s);
    return launch_joint_variant<kTerms, kMode, CLASR_ACT_TANH, 1, kW>(H, tw_hi, tw_lo, t_hid, t_z, p, s);
  }
  const char* we = getenv("CLASR_JOINT_WIDE");
  const bool wide = kW && (we ? atoi(we) != 0 : true);
#define CLASR_LAUNCH_JOINT(ACT)                                                                             \
  (joint_use_pair()                                                                                         \
       ? (wide ? launch_joint_variant<kTerms, kMode, ACT, 1, kW>(H, tw_hi, tw_lo, t_hid, t_z, p, s)              \
               : launch_joint_variant<kTerms, kMode, ACT, 1, 0>(H, tw_hi, tw_lo, t_hid, t_z, p, s))              \
       : (wide ? launch_joint_variant<kTerms, kMode, ACT, 0, kW>(H, tw_hi, tw_lo, t_hid, t_z, p, s)              \
               : launch_joint_variant<kTerms, kMode, ACT, 0, 0>(H, tw_hi, tw_lo, t_hid, t_z, p, s)))
  if (activation == CLASR_ACT_RELU) return CLASR_LAUNCH_JOINT(CLASR_ACT_RELU);
  if (activation == CLASR_ACT_SIGMOID) return CLASR_LAUNCH_JOINT(CLASR_ACT_SIGMOID);
  return CLASR_LAUNCH_JOINT(CLASR_ACT_TANH);
#undef CLASR_LAUNCH_JOINT
#endif
}

// scaled pre-activations (see joint_prep_kernel); relu needs none
static int launch_joint_prep(const float* f, const float* g, int B, int T, int U1, int H, int activation,
                             const JointWs& jw, const float** ef, const float** eg, cudaStream_t s) {
  if (activation == CLASR_ACT_RELU) {
    *ef = f;
    *eg = g;
    return CLASR_STATUS_SUCCESS;
  }
  const int64_t nf4 = (int64_t)B * T * H / 4, ng4 = (int64_t)B * U1 * H / 4;
  auto grid = [](int64_t n4) { int64_t gsz = (n4 + 255) / 256; return (unsigned)(gsz < 1 ? 1 : (gsz > kNumSMs * 8 ? kNumSMs * 8 : gsz)); };
  joint_prep_kernel<<<grid(nf4), 256, 0, s>>>(f, jw.ef, nf4, activation);
  CLASR_CHECK_LAUNCH("joint_prep_f");
  joint_prep_kernel<<<grid(ng4), 256, 0, s>>>(g, jw.eg, ng4, activation);
  CLASR_CHECK_LAUNCH("joint_prep_g");
  *ef = jw.ef;
  *eg = jw.eg;
  return CLASR_STATUS_SUCCESS;
}

}  // namespace clasr

using namespace clasr;

extern "C" size_t clasr_joint_workspace_bytes(int B, int T, int U1, int H, int Vp, int precision) {
  if (B <= 0 || T <= 0 || U1 <= 0 || H <= 0 || Vp <= 0) return 0;
  return joint_ws_carve(nullptr, B, T, U1, H, Vp, precision).total;
}

static int check_joint_args(const char* who, const void* f, const void* g, const void* w_out, const void* b_out,
                            const void* labels, const void* act_lens, const void* label_lens, int B, int T, int U1,
                            int H, int Vp, int blank, int activation, int precision, const void* ws, size_t ws_bytes) {
  CLASR_CHECK_ARG(f && g && w_out && b_out && act_lens && label_lens && ws, "%s: null pointer", who);
  CLASR_CHECK_ARG(labels || U1 == 1, "%s: null labels", who);
  CLASR_CHECK_ARG(B > 0 && T > 0 && U1 > 0 && H > 0 && Vp > 0, "%s: non-positive dimension", who);
  CLASR_CHECK_ARG(H % kJK == 0 && H <= kJMaxH, "%s: joint_hidden must be a multiple of %d and <= %d (got %d)", who, kJK,
                  kJMaxH, H);
  CLASR_CHECK_ARG(blank >= 0 && blank < Vp, "%s: blank %d outside [0,%d)", who, blank, Vp);
  CLASR_CHECK_ARG(activation >= CLASR_ACT_RELU && activation <= CLASR_ACT_TANH, "%s: unknown activation %d", who,
                  activation);
  CLASR_CHECK_ARG(prec_ok(precision), "%s: unknown precision %d", who, precision);
  CLASR_CHECK_ARG(ws_bytes >= clasr_joint_workspace_bytes(B, T, U1, H, Vp, precision), "%s: workspace too small", who);
  CLASR_CHECK_ARG((((uintptr_t)ws) & 255) == 0, "%s: workspace must be 256-byte aligned", who);
  CLASR_CHECK_ARG((((uintptr_t)f) & 15) == 0 && (((uintptr_t)g) & 15) == 0, "%s: f/g must be 16-byte aligned", who);
  return CLASR_STATUS_SUCCESS;
}

extern "C" size_t clasr_joint_stash_bytes(int B, int T, int U1, int H, int Vp, int precision);
// EXPERIMENT, not built by default (-DCLASR_KEEP_HIDDEN): a stash that holds only the hidden activations (kMode 4 forward,
// kMode 5 / 6 backward passes that LOAD their A operands instead of evaluating act(f + g)).  Parity-tested against the
// recompute mode at config-2 size, but measured SLOWER on B200 (pass 2a 4.9 vs 2.8 ms, profiles/r02f): with all 16
// worker warps busy for the whole N loop the MMA warp's issue loop, not the tensor pipe, sets the pace.  Loose end: its last
// builds passed the parity tests but bench.py at B = 32 ended in a CUDA error that was not investigated.
#ifdef CLASR_KEEP_HIDDEN
extern "C" __attribute__((visibility("default"))) size_t clasr_joint_hidden_bytes(int B, int T, int U1, int H, int Vp,
                                                                                  int precision) {
  if (B <= 0 || T <= 0 || U1 <= 0 || H <= 0 || Vp <= 0) return 0;
  return joint_stash_carve(nullptr, B, T, U1, H, Vp, precision, true).total;
}
#endif
// What a `stash` of `bytes` bytes holds: 1 = logits + hidden activations (>= clasr_joint_stash_bytes), 2 = hidden
// activations only (CLASR_KEEP_HIDDEN builds), 0 = too small.
static int joint_stash_kind(size_t bytes, int B, int T, int U1, int H, int Vp, int precision) {
  if (bytes >= clasr_joint_stash_bytes(B, T, U1, H, Vp, precision)) return 1;
#ifdef CLASR_KEEP_HIDDEN
  if (bytes >= clasr_joint_hidden_bytes(B, T, U1, H, Vp, precision)) return 2;
#endif
  return 0;
}

extern "C" int clasr_joint_rnnt_fwd(const float* f, const float* g, const float* w_out, const float* b_out,
                                    const int64_t* labels, const int64_t* act_lens, const int64_t* label_lens, int B,
                                    int T, int U1, int H, int Vp, int blank, int activation, int precision,
                                    float dropout_p, uint64_t dropout_seed, float fastemit_lambda, float* costs,
                                    float* sumsq, void* workspace, size_t workspace_bytes, void* stash,
                                    size_t stash_bytes, void* stream) {
  int rc = check_joint_args("joint_rnnt_fwd", f, g, w_out, b_out, labels, act_lens, label_lens, B, T, U1, H, Vp, blank,
                            activation, precision, workspace, workspace_bytes);
  if (rc) return rc;
  CLASR_CHECK_ARG(costs, "joint_rnnt_fwd: null costs");
  const int stash_kind = stash ? joint_stash_kind(stash_bytes, B, T, U1, H, Vp, precision) : 0;
  if (stash) {
    CLASR_CHECK_ARG(stash_kind != 0, "joint_rnnt_fwd: stash too small");
    CLASR_CHECK_ARG((((uintptr_t)stash) & 255) == 0, "joint_rnnt_fwd: stash must be 256-byte aligned");
  }
  cudaStream_t s = (cudaStream_t)stream;
  JointWs jw = joint_ws_carve(workspace, B, T, U1, H, Vp, precision);
  const bool x3 = prec_x3(precision);
  // W_out [Vp,H] fp32 -> bf16 hi[,lo] (rows beyond Vp are never read: TMA zero-fills out-of-bounds rows)
  const float* wscale = nullptr;
  const bool m8 = prec_m8(precision);
  if (prec_f16(precision)) {   // fp16 operands: W_out is split after a power-of-two scale that brings max|W| to ~1
    // (FP16M8: every operand to just below 2^14, see include/clasr_b200.h)
    if ((rc = launch_joint_gscale(w_out, (int64_t)Vp * H, kScaleHeadroom * m8_headroom(m8), jw.gscale + 2, s))) return rc;
    const int relu = activation == CLASR_ACT_RELU;
    if (relu) {
      if ((rc = launch_joint_gscale(f, (int64_t)B * T * H, 1.f, jw.gscale + 16, s))) return rc;
      if ((rc = launch_joint_gscale(g, (int64_t)B * U1 * H, 1.f, jw.gscale + 18, s))) return rc;
    }
    joint_ascale_kernel<<<1, 1, 0, s>>>(jw.gscale, relu, m8_act_scale(m8, dropout_p));
    CLASR_CHECK_LAUNCH("joint_ascale");
    wscale = jw.gscale + 2;
  }
  if (m8) {   // the pad rows [Vp, vp_pad) of the hi8 part sit INSIDE the combined hi8 | lo8 tensor map: keep them zero
    cudaError_t e0 = cudaMemsetAsync(jw.w_h8, 0, (size_t)2 * jw.vp_pad * H, s);
    CLASR_CHECK_ARG(e0 == cudaSuccess, "joint_rnnt_fwd: memset failed");
  }
  if ((rc = launch_split_bf16(w_out, Vp, H, H, jw.w_hi, x3 ? jw.w_lo : nullptr, H, s, prec_f16(precision), wscale,
                              m8 ? jw.w_h8 : nullptr, m8 ? jw.w_l8 : nullptr)))
    return rc;
  joint_tile_offsets_kernel<<<1, 256, 0, s>>>(act_lens, label_lens, B, jw.tile_offsets, jw.tile_offsets + B + 1,
                                              b_out, Vp, jw.bias_pad);
  CLASR_CHECK_LAUNCH("joint_tile_offsets");

  JointFwdParams p = {};
  if ((rc = launch_joint_prep(f, g, B, T, U1, H, activation, jw, &p.ef, &p.eg, s))) return rc;
  p.bias = b_out; p.bias_pad = jw.bias_pad; p.labels = labels; p.act_lens = act_lens; p.label_lens = label_lens;
  p.tile_offsets = jw.tile_offsets;
  p.B = B; p.T = T; p.U1 = U1; p.H = H; p.Vp = Vp; p.blank = blank;
  p.f16 = prec_f16(precision) ? 1 : 0;
  p.m8 = m8 ? 1 : 0;
  p.wscale = p.f16 ? jw.gscale + 2 : nullptr;   // written by the forward call
  p.ascale = (p.f16 && (activation == CLASR_ACT_RELU || m8)) ? jw.gscale + 4 : nullptr;
  p.w = lattice_ws_carve(jw.lattice, B, T, U1);
  p.sumsq = sumsq;
  if ((rc = set_dropout(p, dropout_p, dropout_seed))) return rc;
  CUtensorMap tw_hi, tw_lo;
  const int bn = (joint_use_pair() || stash_kind == 2) ? (x3 ? JointCfg<3, 1>::kBRows : JointCfg<1, 1>::kBRows)
                                  : (x3 ? JointCfg<3, 0>::kBRows : JointCfg<1, 0>::kBRows);
  if ((rc = make_tmap_bf16_2d(&tw_hi, jw.w_hi, Vp, H, H, bn, kJK))) return rc;
  if (m8) {   // e4m3 hi8 | lo8 of W as ONE [2 vp_pad, H]-byte tensor, boxes [bn rows][64 B], 64-byte swizzle
    p.w8_rows = jw.vp_pad;
    if ((rc = make_tmap_2d(&tw_lo, jw.w_h8, 2 * (uint64_t)jw.vp_pad, H, H, bn, kJK, 1, 64))) return rc;
  } else if (x3) {
    if ((rc = make_tmap_bf16_2d(&tw_lo, jw.w_lo, Vp, H, H, bn, kJK))) return rc;
  } else {
    tw_lo = tw_hi;
  }
  prof_begin("joint_fwd", s);
#ifdef CLASR_KEEP_HIDDEN
  if (stash_kind == 2) {  // keep the hidden activations (as GEMM operands) for the backward (kMode 4); no logits
    JointStash st = joint_stash_carve(stash, B, T, U1, H, Vp, precision, true);
    CLASR_CHECK_ARG(st.rows_cap < 2147483647LL, "joint_rnnt_fwd: too many lattice cells");
    p.hid_hi = (__nv_bfloat16*)st.hid_hi; p.hid_lo = (__nv_bfloat16*)st.hid_lo; p.ldh = H;
    p.hid_h8 = (uint8_t*)st.hid_lo; p.hid_l8 = (uint8_t*)st.hid_lo + (size_t)st.rows_cap * H;   // FP16M8: in hid_lo's place
    CUtensorMap t_hid;
    if ((rc = make_tmap_bf16_2d(&t_hid, st.hid_hi, (uint64_t)st.rows_cap, H, H, kJM, kJK))) return rc;
    rc = m8 ? launch_joint_kernel<4, 4>(activation, H, tw_hi, tw_lo, t_hid, tw_hi, p, s)
         : x3 ? launch_joint_kernel<3, 4>(activation, H, tw_hi, tw_lo, t_hid, tw_hi, p, s)
              : launch_joint_kernel<1, 4>(activation, H, tw_hi, tw_lo, t_hid, tw_hi, p, s);
  } else
#endif
  if (stash) {  // keep z and Hid for the backward (kMode 3)
    JointStash st = joint_stash_carve(stash, B, T, U1, H, Vp, precision);
    CLASR_CHECK_ARG(st.rows_cap < 2147483647LL, "joint_rnnt_fwd: too many lattice cells");
    p.zbuf = st.z; p.ldzf = st.ldzf;
    p.hid_hi = (__nv_bfloat16*)st.hid_hi; p.hid_lo = (__nv_bfloat16*)st.hid_lo; p.ldh = H;
    p.hid_h8 = (uint8_t*)st.hid_lo; p.hid_l8 = (uint8_t*)st.hid_lo + (size_t)st.rows_cap * H;   // FP16M8: in hid_lo's place
    CUtensorMap t_hid, t_z;
    if ((rc = make_tmap_bf16_2d(&t_hid, st.hid_hi, (uint64_t)st.rows_cap, H, H, kJM, kJK))) return rc;
    // TMA-store view of z: [rows_cap, ldzf] fp32, box = 32 rows x 8 columns (32-byte rows, 32-byte swizzle)
    if ((rc = make_tmap_2d(&t_z, st.z, (uint64_t)st.rows_cap, st.ldzf, st.ldzf, 32, 8, 4, 32))) return rc;
    rc = m8 ? launch_joint_kernel<4, 3>(activation, H, tw_hi, tw_lo, t_hid, t_z, p, s)
         : x3 ? launch_joint_kernel<3, 3>(activation, H, tw_hi, tw_lo, t_hid, t_z, p, s)
              : launch_joint_kernel<1, 3>(activation, H, tw_hi, tw_lo, t_hid, t_z, p, s);
  } else {
    rc = m8 ? launch_joint_kernel<4, 0>(activation, H, tw_hi, tw_lo, tw_hi /*unused in pass 1*/, tw_hi, p, s)
         : x3 ? launch_joint_kernel<3, 0>(activation, H, tw_hi, tw_lo, tw_hi, tw_hi, p, s)
              : launch_joint_kernel<1, 0>(activation, H, tw_hi, tw_lo, tw_hi, tw_hi, p, s);
  }
  if (rc) return rc;
  prof_end("joint_fwd", s);
  CLASR_CHECK_LAUNCH("joint_fwd");
  return launch_rnnt_lattice(p.w, act_lens, label_lens, B, T, U1, fastemit_lambda, costs, s);
}

size_t clasr_joint_stash_bytes(int B, int T, int U1, int H, int Vp, int precision) {
  if (B <= 0 || T <= 0 || U1 <= 0 || H <= 0 || Vp <= 0) return 0;
  return joint_stash_carve(nullptr, B, T, U1, H, Vp, precision).total;
}

extern "C" size_t clasr_joint_bwd_scratch_bytes(int B, int T, int U1, int H, int Vp, int precision) {
  if (B <= 0 || T <= 0 || U1 <= 0 || H <= 0 || Vp <= 0) return 0;
  return joint_bwd_scratch_carve(nullptr, B, T, U1, H, Vp, precision).total;
}

// mode 1: transducer-loss backward (grad_out [B]);  mode 2: backward of sum_v z^2 per cell (grad_cells [B,T,U1])
static int joint_bwd_impl(int mode, const float* f, const float* g, const float* w_out, const float* b_out,
                          const int64_t* labels, const int64_t* act_lens, const int64_t* label_lens, int B, int T, int U1,
                          int H, int Vp, int blank, int activation, int precision, float dropout_p,
                          uint64_t dropout_seed, float fastemit_lambda, float clamp, const float* grad_out,
                          const float* grad_cells, float* d_f, float* d_g, float* d_w_out,
                          float* d_b_out, void* workspace, size_t workspace_bytes, void* scratch, size_t scratch_bytes,
                          void* stash, size_t stash_bytes, void* stream) {
  int rc = check_joint_args("joint_rnnt_bwd", f, g, w_out, b_out, labels, act_lens, label_lens, B, T, U1, H, Vp, blank,
                            activation, precision, workspace, workspace_bytes);
  if (rc) return rc;
  CLASR_CHECK_ARG(d_f && d_g && d_w_out && d_b_out && scratch, "joint_rnnt_bwd: null output / scratch");
  CLASR_CHECK_ARG(clamp >= 0.f, "joint_rnnt_bwd: `clamp` must be 0.0 or positive");
  CLASR_CHECK_ARG(scratch_bytes >= clasr_joint_bwd_scratch_bytes(B, T, U1, H, Vp, precision),
                  "joint_rnnt_bwd: scratch too small");
  CLASR_CHECK_ARG((((uintptr_t)scratch) & 255) == 0, "joint_rnnt_bwd: scratch must be 256-byte aligned");
  CLASR_CHECK_ARG((H & 3) == 0, "joint_rnnt_bwd: H must be a multiple of 4");
  const int stash_kind = stash ? joint_stash_kind(stash_bytes, B, T, U1, H, Vp, precision) : 0;
  if (stash) {
    CLASR_CHECK_ARG(stash_kind != 0, "joint_rnnt_bwd: stash too small");
    CLASR_CHECK_ARG((((uintptr_t)stash) & 255) == 0, "joint_rnnt_bwd: stash must be 256-byte aligned");
  }
  cudaStream_t s = (cudaStream_t)stream;
  const bool x3 = prec_x3(precision);
  JointWs jw = joint_ws_carve(workspace, B, T, U1, H, Vp, precision);   // filled by the forward call
  JointBwdScratch sc = joint_bwd_scratch_carve(scratch, B, T, U1, H, Vp, precision);
  CLASR_CHECK_ARG(sc.rows_cap < 2147483647LL, "joint_rnnt_bwd: too many lattice cells");
  int* rows_pad_dev = jw.tile_offsets + B + 1;

  // ---- pass 2a: recompute logits tile-wise, emit dZ (bf16 hi/lo) and the hidden activations as GEMM operands
  JointFwdParams p = {};
  // the scaled pre-activations were written into the workspace by the forward call (relu: f / g themselves)
  p.ef = activation == CLASR_ACT_RELU ? f : jw.ef;
  p.eg = activation == CLASR_ACT_RELU ? g : jw.eg;
  p.bias = b_out; p.bias_pad = jw.bias_pad; p.labels = labels; p.act_lens = act_lens; p.label_lens = label_lens;
  p.tile_offsets = jw.tile_offsets;
  p.B = B; p.T = T; p.U1 = U1; p.H = H; p.Vp = Vp; p.blank = blank;
  const bool m8 = prec_m8(precision);
  p.f16 = prec_f16(precision) ? 1 : 0;
  p.m8 = m8 ? 1 : 0;
  p.wscale = p.f16 ? jw.gscale + 2 : nullptr;   // written by the forward call
  p.ascale = (p.f16 && (activation == CLASR_ACT_RELU || m8)) ? jw.gscale + 4 : nullptr;
  p.w = lattice_ws_carve(jw.lattice, B, T, U1);
  p.grad_out = grad_out; p.grad_cells = grad_cells; p.fastemit_lambda = fastemit_lambda; p.clamp = clamp;
  if ((rc = set_dropout(p, dropout_p, dropout_seed))) return rc;
  p.dz_hi = (__nv_bfloat16*)sc.dz_hi; p.dz_lo = (__nv_bfloat16*)sc.dz_lo; p.ldz = sc.ldz;
  p.hid_hi = (__nv_bfloat16*)sc.hid_hi; p.hid_lo = (__nv_bfloat16*)sc.hid_lo; p.ldh = sc.ldh;
  // FP16M8: e4m3 hi8 | lo8 (one byte per element each) in the memory of the 16-bit lo arrays
  p.dz_h8 = (uint8_t*)sc.dz_lo; p.dz_l8 = (uint8_t*)sc.dz_lo + (size_t)sc.rows_cap * sc.ldz;
  p.hid_h8 = (uint8_t*)sc.hid_lo; p.hid_l8 = (uint8_t*)sc.hid_lo + (size_t)sc.rows_cap * sc.ldh;
  p.rows_pad_dev = rows_pad_dev;
  // blank == last class (NeMo: RNNTLoss._blank = num_classes): its dW row comes from joint_dfg, the GEMM covers V rows
  const bool blank_split = blank == Vp - 1 && Vp > 1 && (H % 32) == 0 &&
                           (size_t)2 * U1 * 32 * sizeof(float) <= 200 * 1024;
  p.dzb = blank_split ? sc.dzb : nullptr;
  p.db_acc = d_b_out;
  if (p.f16 && (mode == 2 || grad_out)) {   // fp16 operands: pre-scale dZ to O(1) (see joint_gscale_kernel)
    rc = mode == 1 ? launch_joint_gscale(grad_out, B, (1.f + fastemit_lambda) * kScaleHeadroom * m8_headroom(m8), jw.gscale, s)
                   // dZ = 2 z upstream: headroom for |z| up to 64 (fp16 operands ~1) / up to 512 in FP16M8, whose operands
                   // sit 2^14 higher and must stay below fp16's 65504
                   : launch_joint_gscale(grad_cells, (int64_t)B * T * U1, (m8 ? 1024.f : 128.f) * m8_headroom(m8), jw.gscale, s);
    if (rc) return rc;
    p.gscale = jw.gscale;
  }
  {  // d_b accumulates in the pass-2a epilogue, d_W in the split-K GEMM: both start from zero
    cudaError_t e1 = cudaMemsetAsync(d_b_out, 0, (size_t)Vp * sizeof(float), s);
    cudaError_t e2 = cudaMemsetAsync(d_w_out, 0, (size_t)Vp * H * sizeof(float), s);
    CLASR_CHECK_ARG(e1 == cudaSuccess && e2 == cudaSuccess, "joint_rnnt_bwd: memset failed");
  }
  CUtensorMap tw_hi, tw_lo;
  const int bn = (joint_use_pair() || stash_kind == 2) ? (x3 ? JointCfg<3, 1>::kBRows : JointCfg<1, 1>::kBRows)
                                  : (x3 ? JointCfg<3, 0>::kBRows : JointCfg<1, 0>::kBRows);
  if ((rc = make_tmap_bf16_2d(&tw_hi, jw.w_hi, Vp, H, H, bn, kJK))) return rc;
  if (m8) {
    p.w8_rows = jw.vp_pad;
    if ((rc = make_tmap_2d(&tw_lo, jw.w_h8, 2 * (uint64_t)jw.vp_pad, H, H, bn, kJK, 1, 64))) return rc;
  } else if (x3) {
    if ((rc = make_tmap_bf16_2d(&tw_lo, jw.w_lo, Vp, H, H, bn, kJK))) return rc;
  } else {
    tw_lo = tw_hi;
  }
  const void* hid_hi = sc.hid_hi;
  const void* hid_lo = sc.hid_lo;            // FP16M8: the e4m3 hi8 array, lo8 follows at rows_cap * H bytes
  const void* hid_l8 = p.hid_l8;
#ifdef CLASR_KEEP_HIDDEN
  if (stash_kind == 2) {
    // the forward call kept the hidden activations (kMode 4): the logits are recomputed tile-wise with the A operands
    // LOADED (TMA + tensor-memory stores) instead of rebuilt from f and g (kMode 5 / 6)
    JointStash st = joint_stash_carve(stash, B, T, U1, H, Vp, precision, true);
    hid_hi = st.hid_hi; hid_lo = st.hid_lo;
    hid_l8 = (const uint8_t*)st.hid_lo + (size_t)st.rows_cap * H;
    p.hid_hi = (__nv_bfloat16*)st.hid_hi; p.hid_lo = (__nv_bfloat16*)st.hid_lo; p.ldh = H;
    p.hid_h8 = (uint8_t*)st.hid_lo; p.hid_l8 = (uint8_t*)st.hid_lo + (size_t)st.rows_cap * H;
    prof_begin("joint_bwd_dz", s);
    CUtensorMap t_hid;  // TMA-load view of Hid_hi: [rows_cap, H] 16-bit, box = one 128 x 64 A block
    if ((rc = make_tmap_bf16_2d(&t_hid, st.hid_hi, (uint64_t)st.rows_cap, H, H, kJM, kJK))) return rc;
    if (mode == 1)
      rc = m8 ? launch_joint_kernel<4, 5>(activation, H, tw_hi, tw_lo, t_hid, tw_hi, p, s)
           : x3 ? launch_joint_kernel<3, 5>(activation, H, tw_hi, tw_lo, t_hid, tw_hi, p, s)
                : launch_joint_kernel<1, 5>(activation, H, tw_hi, tw_lo, t_hid, tw_hi, p, s);
    else
      rc = m8 ? launch_joint_kernel<4, 6>(activation, H, tw_hi, tw_lo, t_hid, tw_hi, p, s)
           : x3 ? launch_joint_kernel<3, 6>(activation, H, tw_hi, tw_lo, t_hid, tw_hi, p, s)
                : launch_joint_kernel<1, 6>(activation, H, tw_hi, tw_lo, t_hid, tw_hi, p, s);
    if (rc) return rc;
    prof_end("joint_bwd_dz", s);
    CLASR_CHECK_LAUNCH("joint_bwd_dz");
  } else
#endif
  if (stash) {
    // the forward call kept z and Hid (kMode 3): dZ is one streaming sweep over z
    JointStash st = joint_stash_carve(stash, B, T, U1, H, Vp, precision);
    p.zbuf = st.z; p.ldzf = st.ldzf;
    hid_hi = st.hid_hi; hid_lo = st.hid_lo;
    hid_l8 = (const uint8_t*)st.hid_lo + (size_t)st.rows_cap * H;
    prof_begin("joint_dz_sweep", s);
    if (mode == 1) rc = x3 ? launch_joint_dz<3, 1>(p, s) : launch_joint_dz<1, 1>(p, s);
    else rc = x3 ? launch_joint_dz<3, 2>(p, s) : launch_joint_dz<1, 2>(p, s);
    if (rc) return rc;
    prof_end("joint_dz_sweep", s);
  } else {
    prof_begin("joint_bwd_dz", s);
    CUtensorMap t_hid;  // TMA-store view of Hid_hi: [rows_cap, H] bf16, box = one 128 x 64 A block
    if ((rc = make_tmap_bf16_2d(&t_hid, sc.hid_hi, (uint64_t)sc.rows_cap, H, sc.ldh, kJM, kJK))) return rc;
    if (mode == 1)
      rc = m8 ? launch_joint_kernel<4, 1>(activation, H, tw_hi, tw_lo, t_hid, tw_hi, p, s)
           : x3 ? launch_joint_kernel<3, 1>(activation, H, tw_hi, tw_lo, t_hid, tw_hi, p, s)
                : launch_joint_kernel<1, 1>(activation, H, tw_hi, tw_lo, t_hid, tw_hi, p, s);
    else
      rc = m8 ? launch_joint_kernel<4, 2>(activation, H, tw_hi, tw_lo, t_hid, tw_hi, p, s)
           : x3 ? launch_joint_kernel<3, 2>(activation, H, tw_hi, tw_lo, t_hid, tw_hi, p, s)
                : launch_joint_kernel<1, 2>(activation, H, tw_hi, tw_lo, t_hid, tw_hi, p, s);
    if (rc) return rc;
    prof_end("joint_bwd_dz", s);
    CLASR_CHECK_LAUNCH("joint_bwd_dz");
  }

  // ---- pass 2b: dHid[rows, H] = dZ[rows, Vp] . W[Vp, H]      (A K-major, B = W consumed MN-major: no transpose)
  prof_begin("gemm_dhid", s);
  // blank_split: the blank column (class V, the only one past a multiple of 64 in K when V = 1024) would cost a whole
  // extra K block; its rank-1 contribution dZ[., blank] x W[blank, :] is added by joint_dfg from dzb instead
  if ((rc = launch_gemm_tc(sc.dz_hi, sc.dz_lo, sc.ldz, 0, jw.w_hi, m8 ? jw.w_h8 : jw.w_lo, H, 1, (int)sc.rows_cap, H,
                           blank_split ? Vp - 1 : Vp, sc.dhid, H, precision, 0, 1, s, rows_pad_dev, nullptr, nullptr,
                           nullptr, nullptr, m8 ? p.dz_l8 : nullptr, m8 ? jw.w_l8 : nullptr)))
    return rc;
  prof_end("gemm_dhid", s);
  // ---- pass 2c: dW[Vp, H] = dZ^T . Hid       (both operands MN-major, split-K over the rows, fp32 atomics)
  prof_begin("gemm_dw", s);
  if ((rc = launch_gemm_tc(sc.dz_hi, sc.dz_lo, sc.ldz, 1, hid_hi, hid_lo, sc.ldh, 1, blank_split ? Vp - 1 : Vp, H,
                           (int)sc.rows_cap, d_w_out, H, precision, 1, /*auto split-K*/ 0, s, nullptr, rows_pad_dev,
                           nullptr, p.gscale ? p.gscale + 1 : nullptr, p.ascale ? p.ascale + 1 : nullptr,
                           m8 ? p.dz_l8 : nullptr, m8 ? hid_l8 : nullptr)))
    return rc;
  prof_end("gemm_dw", s);
  // ---- pass 2d: through the activation and the broadcast add: d_f = sum_u, d_g = sum_t of dHid * act'(f+g)
  prof_begin("joint_dfg", s);
  {
    const size_t smem = (size_t)2 * U1 * 32 * sizeof(float);
    if (smem <= 200 * 1024 && (H % 32) == 0) {
      const float* ef_ = activation == CLASR_ACT_RELU ? f : jw.ef;
      const float* eg_ = activation == CLASR_ACT_RELU ? g : jw.eg;
      cudaError_t e0 = cudaMemsetAsync(d_g, 0, (size_t)B * U1 * H * sizeof(float), s);
      CLASR_CHECK_ARG(e0 == cudaSuccess, "joint_rnnt_bwd: memset failed");
      const dim3 grid(H / 32, B, (T + kDfgTChunk - 1) / kDfgTChunk);
      // TMA variant: U+1 <= 13 x 8 prediction rows (CLASR_DFG_TMA=0 selects the shared-atomics kernel)
      const char* dte = getenv("CLASR_DFG_TMA");
      const bool use_tma = U1 <= kDfgUW * kDfgConsWarps && (dte ? atoi(dte) != 0 : true);
      CUtensorMap t_dhid;
      const size_t smem_tma = (size_t)kDfgStages * U1 * 32 * 4 + (size_t)kDfgTChunk * kDfgConsWarps * 32 * 4 + 2 * kDfgStages * 8 + 128;
      if (use_tma) {   // dHid [rows_cap, H] fp32, box = (U+1) rows x 32 columns, no swizzle (rows of 128 B: lane = feature)
        if ((rc = make_tmap_2d(&t_dhid, sc.dhid, (uint64_t)sc.rows_cap, H, H, (uint32_t)U1, 32, 4, 0))) return rc;
      }
#define CLASR_LAUNCH_DFG_TMA(ACT, DROP)                                                                            \
  do {                                                                                                             \
    cudaFuncSetAttribute(joint_dfg_tma_kernel<ACT, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tma); \
    joint_dfg_tma_kernel<ACT, DROP><<<grid, 32 * (kDfgConsWarps + 1), smem_tma, s>>>(                              \
        t_dhid, ef_, eg_, act_lens, label_lens, jw.tile_offsets, T, U1, H, d_f, d_g, blank_split ? sc.dzb : nullptr, \
        d_w_out + (size_t)blank * H, w_out + (size_t)blank * H, p.gscale, p.wscale, p.drop_thresh, p.drop_seed_a,  \
        p.drop_seed_b, p.drop_scale, U1);                                                                          \
  } while (0)
#define CLASR_LAUNCH_DFG(ACT)                                                                                      \
  do {                                                                                                             \
    cudaFuncSetAttribute(joint_dfg_fused_kernel<ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
    joint_dfg_fused_kernel<ACT><<<grid, 256, smem, s>>>(sc.dhid, ef_, eg_, act_lens, label_lens, jw.tile_offsets,  \
                                                        T, U1, H, d_f, d_g, blank_split ? sc.dzb : nullptr,        \
                                                        d_w_out + (size_t)blank * H, w_out + (size_t)blank * H,    \
                                                        p.gscale, p.wscale, p.drop_thresh, p.drop_seed_a,          \
                                                        p.drop_seed_b, p.drop_scale);                              \
  } while (0)
      if (use_tma) {
        if (p.drop_thresh) {
          if (activation == CLASR_ACT_RELU) CLASR_LAUNCH_DFG_TMA(CLASR_ACT_RELU, true);
          else if (activation == CLASR_ACT_SIGMOID) CLASR_LAUNCH_DFG_TMA(CLASR_ACT_SIGMOID, true);
          else CLASR_LAUNCH_DFG_TMA(CLASR_ACT_TANH, true);
        } else {
          if (activation == CLASR_ACT_RELU) CLASR_LAUNCH_DFG_TMA(CLASR_ACT_RELU, false);
          else if (activation == CLASR_ACT_SIGMOID) CLASR_LAUNCH_DFG_TMA(CLASR_ACT_SIGMOID, false);
          else CLASR_LAUNCH_DFG_TMA(CLASR_ACT_TANH, false);
        }
      } else if (activation == CLASR_ACT_RELU) CLASR_LAUNCH_DFG(CLASR_ACT_RELU);
      else if (activation == CLASR_ACT_SIGMOID) CLASR_LAUNCH_DFG(CLASR_ACT_SIGMOID);
      else CLASR_LAUNCH_DFG(CLASR_ACT_TANH);
#undef CLASR_LAUNCH_DFG
#undef CLASR_LAUNCH_DFG_TMA
      CLASR_CHECK_LAUNCH("joint_dfg_fused");
    } else {  // very long label sequences: the two-pass kernel (reads dHid twice, no shared-memory partials)
      CLASR_CHECK_ARG(p.drop_thresh == 0, "joint_rnnt_bwd: in-kernel dropout needs U+1 <= 800 (single-pass d_f/d_g kernel)");
      CLASR_CHECK_ARG(!p.f16, "joint_rnnt_bwd: fp16x3 needs U+1 <= 800 (single-pass d_f/d_g kernel)");
      joint_dfg_kernel<<<dim3(T, B), 160, 0, s>>>(sc.dhid, f, g, act_lens, label_lens, jw.tile_offsets, T, U1, H,
                                                 activation, 0, d_f);
      CLASR_CHECK_LAUNCH("joint_df");
      joint_dfg_kernel<<<dim3(U1, B), 160, 0, s>>>(sc.dhid, f, g, act_lens, label_lens, jw.tile_offsets, T, U1, H,
                                                  activation, 1, d_g);
      CLASR_CHECK_LAUNCH("joint_dg");
    }
  }
  prof_end("joint_dfg", s);
  return CLASR_STATUS_SUCCESS;
}

extern "C" int clasr_joint_rnnt_bwd(const float* f, const float* g, const float* w_out, const float* b_out,
                                    const int64_t* labels, const int64_t* act_lens, const int64_t* label_lens, int B,
                                    int T, int U1, int H, int Vp, int blank, int activation, int precision,
                                    float dropout_p, uint64_t dropout_seed, float fastemit_lambda, float clamp,
                                    const float* grad_out, float* d_f, float* d_g, float* d_w_out, float* d_b_out,
                                    void* workspace, size_t workspace_bytes, void* scratch, size_t scratch_bytes,
                                    void* stash, size_t stash_bytes, void* stream) {
  return joint_bwd_impl(1, f, g, w_out, b_out, labels, act_lens, label_lens, B, T, U1, H, Vp, blank, activation,
                        precision, dropout_p, dropout_seed, fastemit_lambda, clamp, grad_out, nullptr, d_f, d_g, d_w_out, d_b_out, workspace,
                        workspace_bytes, scratch, scratch_bytes, stash, stash_bytes, stream);
}

extern "C" int clasr_joint_sumsq_bwd(const float* f, const float* g, const float* w_out, const float* b_out,
                                     const int64_t* labels, const int64_t* act_lens, const int64_t* label_lens, int B,
                                     int T, int U1, int H, int Vp, int blank, int activation, int precision,
                                     float dropout_p, uint64_t dropout_seed, const float* grad_cells, float* d_f,
                                     float* d_g, float* d_w_out, float* d_b_out,
                                     void* workspace, size_t workspace_bytes, void* scratch, size_t scratch_bytes,
                                     void* stash, size_t stash_bytes, void* stream) {
  CLASR_CHECK_ARG(grad_cells, "joint_sumsq_bwd: null grad_cells");
  return joint_bwd_impl(2, f, g, w_out, b_out, labels, act_lens, label_lens, B, T, U1, H, Vp, blank, activation,
                        precision, dropout_p, dropout_seed, 0.f, 0.f, nullptr, grad_cells, d_f, d_g, d_w_out, d_b_out,
                        workspace, workspace_bytes,
                        scratch, scratch_bytes, stash, stash_bytes, stream);
}

#ifdef CLASR_TRACE
// developer builds only: copies the per-CTA wait-cycle counters of the LAST joint kernel launch to host memory
extern "C" __attribute__((visibility("default"))) int clasr_debug_joint_trace(unsigned long long* out_host, int n) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(out_host, clasr::g_joint_trace, sizeof(unsigned long long) * (size_t)n);
}
#endif
