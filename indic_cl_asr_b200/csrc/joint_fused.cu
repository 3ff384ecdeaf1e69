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
//   producer warps   synthesise the A operand on the fly: act(f + g) -> bf16 (hi[,lo]) straight into the
//                    128-byte-swizzled K-major UMMA layout in shared memory (A is not TMA-loadable: it does not
//                    exist in memory); in BF16X3 mode the lo halves go to TENSOR MEMORY (tcgen05.st) and feed
//                    the MMA as a TMEM A-operand, because hi+lo of a 128 x 640 tile does not fit in 227 KB smem
//   TMA warp         streams W_out (bf16 hi[,lo], K-major) through an mbarrier ring
//   MMA warp         one thread issues tcgen05.mma M=128 x N=BN x K=16 into double-buffered TMEM accumulators
//   epilogue warps   tcgen05.ld the accumulators, add the bias and keep a running (max, sum-exp) per row across the
//                    N tiles (online log-sum-exp), pick out logit[blank] and logit[label_u] (and sum z^2 for MAS);
//                    per cell only {-lse, log p(blank), log p(label)} reach HBM (20 B instead of 4.1 KB).
// The alpha/beta wavefront (rnnt_loss.cu) then runs on those compact buffers.
#include "common.cuh"
#include "tc_common.cuh"

namespace clasr {

int launch_split_bf16(const float* src, int64_t rows, int cols, int64_t src_ld, void* hi, void* lo, int cols_pad,
                      cudaStream_t s);

constexpr int kJM = 128;            // rows (lattice cells) per tile
constexpr int kJK = 64;             // K block (one 128-byte swizzle span of bf16)
constexpr int kJMaxH = 640;         // A tile (128 x H bf16) must stay resident in shared memory
constexpr int kJThreads = 512;      // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-7 epilogue, 8-15 A producers
constexpr int kJProducerWarps = 8;
constexpr int kJStages = 2;

template <int kTerms>
struct JointCfg {
  static constexpr int kBN = kTerms == 1 ? 256 : 96;          // accumulator tile width (TMEM columns)
  static constexpr int kParts = kTerms == 1 ? 1 : 2;          // W parts streamed per stage (hi[,lo])
  static constexpr int kAccCols = 2 * kBN;                    // two accumulator stages
  static constexpr int kAloCol = kAccCols;                    // BF16X3: A_lo lives in TMEM columns [192, 192+H/2)
  static constexpr int kABlockBytes = kJM * kJK * 2;          // 16 KB per K block of A
  static constexpr int kBStageBytes = kParts * kBN * kJK * 2; // W ring stage
  static constexpr int kStagingBytes = kTerms == 1 ? 0 : kABlockBytes;  // A_lo staging (row-major -> lane-major)
  static constexpr int smem_bytes(int H) {
    return (H / kJK) * kABlockBytes + kStagingBytes + kJStages * kBStageBytes + 1024 + 512;
  }
};

struct JointFwdParams {
  const float* f;      // [B,T,H]
  const float* g;      // [B,U1,H]
  const float* bias;   // [Vp]
  const int64_t* labels;
  const int64_t* act_lens;
  const int64_t* label_lens;
  const int* tile_offsets;  // [B+1] prefix sums of per-utterance tile counts
  int B, T, U1, H, Vp, blank, activation;
  LatticeWs w;
  float* sumsq;        // [B,T,U1] sum_v z^2 (MAS), or nullptr
};

__device__ __forceinline__ float joint_act(float x, int act) {
  if (act == CLASR_ACT_RELU) return fmaxf(x, 0.f);
  if (act == CLASR_ACT_SIGMOID) return __fdividef(1.f, 1.f + __expf(-x));
  // tanh = sign(x) * (1 - 2 / (exp(2|x|) + 1)) : absolute error ~2e-7 (what matters for sum_k h_k W_kv)
  const float ax = fabsf(x);
  const float e = __expf(2.f * ax);
  const float r = 1.f - __fdividef(2.f, e + 1.f);
  return copysignf(r, x);
}

__global__ void joint_tile_offsets_kernel(const int64_t* __restrict__ act_lens, const int64_t* __restrict__ label_lens,
                                          int B, int* __restrict__ tile_offsets) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int acc = 0;
    tile_offsets[0] = 0;
    for (int b = 0; b < B; ++b) {
      const int64_t Tb = act_lens[b] > 0 ? act_lens[b] : 0;
      const int64_t cells = Tb * (label_lens[b] + 1);
      acc += (int)((cells + kJM - 1) / kJM);
      tile_offsets[b + 1] = acc;
    }
  }
}

__device__ __forceinline__ int find_utterance(const int* __restrict__ offs, int B, int tile) {
  int lo = 0, hi = B - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (offs[mid] <= tile) lo = mid; else hi = mid - 1;
  }
  return lo;
}

template <int kTerms>
__global__ void __launch_bounds__(kJThreads, 1)
joint_fwd_kernel(const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                 JointFwdParams p) {
  using C = JointCfg<kTerms>;
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
  const int kblocks = p.H / kJK;
  uint8_t* a_smem = smem;                                          // [kblocks][128 x 64 bf16], SW128 K-major
  uint8_t* staging = a_smem + kblocks * C::kABlockBytes;            // BF16X3 only
  uint8_t* b_ring = staging + C::kStagingBytes;                     // [stages][parts][BN x 64 bf16]
  uint64_t* bars = (uint64_t*)(b_ring + kJStages * C::kBStageBytes);
  uint64_t* full = bars;                 // [kJStages]  W stage landed
  uint64_t* empty = full + kJStages;     // [kJStages]  W stage consumed
  uint64_t* tmem_full = empty + kJStages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;     // [2]
  uint64_t* a_ready = tmem_empty + 2;       // [kblocks <= 10] A K-block written
  uint64_t* a_free = a_ready + 10;          // [1] all MMAs of the tile retired -> A may be overwritten
  uint32_t* tmem_base_slot = (uint32_t*)(a_free + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.tile_offsets[p.B];
  const int n_tiles = (p.Vp + C::kBN - 1) / C::kBN;
  const int n_last = ((p.Vp - (n_tiles - 1) * C::kBN) + 15) / 16 * 16;  // width of the last N tile (multiple of 16)

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmW_hi);
    if (kTerms > 1) tc::prefetch_tmap(&tmW_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kJStages; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&tmem_full[i], 1); tc::mbar_init(&tmem_empty[i], 4); }
    for (int i = 0; i < 10; ++i) tc::mbar_init(&a_ready[i], kJProducerWarps);
    tc::mbar_init(a_free, 1);
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc(tmem_base_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ============================ TMA producer: W ring ============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        for (int nt = 0; nt < n_tiles; ++nt) {
          for (int kb = 0; kb < kblocks; ++kb) {
            tc::mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* st = b_ring + stage * C::kBStageBytes;
            tc::mbar_expect_tx(&full[stage], C::kBStageBytes);
            tc::tma_load_2d(st, &tmW_hi, &full[stage], kb * kJK, nt * C::kBN);
            if (kTerms > 1) tc::tma_load_2d(st + C::kBN * kJK * 2, &tmW_lo, &full[stage], kb * kJK, nt * C::kBN);
            if (++stage == kJStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      const uint32_t idesc_full = tc::make_idesc_bf16(kJM, C::kBN);
      const uint32_t idesc_last = tc::make_idesc_bf16(kJM, n_last);
      int stage = 0;
      uint32_t phase = 0;
      int acc_it = 0;
      int tile_it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_it) {
        const uint32_t tile_phase = tile_it & 1;
        for (int nt = 0; nt < n_tiles; ++nt, ++acc_it) {
          const int acc = acc_it & 1;
          const uint32_t acc_phase = (acc_it >> 1) & 1;
          const uint32_t idesc = (nt == n_tiles - 1) ? idesc_last : idesc_full;
          tc::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
          tc::tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * C::kBN;
          for (int kb = 0; kb < kblocks; ++kb) {
            if (nt == 0) tc::mbar_wait(&a_ready[kb], tile_phase);
            tc::mbar_wait(&full[stage], phase);
            tc::tc_fence_after();
            const uint32_t a_hi = tc::smem_u32(a_smem + kb * C::kABlockBytes);
            const uint32_t b_hi = tc::smem_u32(b_ring + stage * C::kBStageBytes);
            const uint32_t b_lo = b_hi + C::kBN * kJK * 2;
#pragma unroll
            for (int kk = 0; kk < kJK / 16; ++kk) {
              const uint32_t koff = kk * 32;
              const uint32_t accum = (kb == 0 && kk == 0) ? 0u : 1u;
              tc::umma_ss(d_tmem, tc::make_desc_kmajor_sw128(a_hi + koff), tc::make_desc_kmajor_sw128(b_hi + koff),
                          idesc, accum);
              if (kTerms > 1) {
                tc::umma_ss(d_tmem, tc::make_desc_kmajor_sw128(a_hi + koff), tc::make_desc_kmajor_sw128(b_lo + koff),
                            idesc, 1u);
                // A_lo from tensor memory: 16 bf16 of K = 8 packed 32-bit columns
                tc::umma_ts(d_tmem, tmem_base + C::kAloCol + kb * (kJK / 2) + kk * 8,
                            tc::make_desc_kmajor_sw128(b_hi + koff), idesc, 1u);
              }
            }
            tc::umma_commit(&empty[stage]);
            if (++stage == kJStages) { stage = 0; phase ^= 1; }
          }
          tc::umma_commit(&tmem_full[acc]);
        }
        tc::umma_commit(a_free);  // every MMA reading this tile's A has retired
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ============================ epilogue: online log-sum-exp + gather ============================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int acc_it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int b = find_utterance(p.tile_offsets, p.B, tile);
      const int Tb = (int)p.act_lens[b], Ub1 = (int)p.label_lens[b] + 1;
      const int r = (tile - p.tile_offsets[b]) * kJM + row;  // cell index inside the utterance (t-major)
      const bool valid = r < Tb * Ub1;
      const int t = valid ? r / Ub1 : 0;
      const int u = valid ? r - t * Ub1 : 0;
      const int label = (valid && u < Ub1 - 1) ? (int)p.labels[(int64_t)b * (p.U1 - 1) + u] : -1;
      float m = -INFINITY, s = 0.f, zb = 0.f, zl = 0.f, ssq = 0.f;
      for (int nt = 0; nt < n_tiles; ++nt, ++acc_it) {
        const int acc = acc_it & 1;
        const uint32_t acc_phase = (acc_it >> 1) & 1;
        tc::mbar_wait(&tmem_full[acc], acc_phase);
        tc::tc_fence_after();
        const int ncols = (nt == n_tiles - 1) ? (p.Vp - nt * C::kBN) : C::kBN;
#pragma unroll 1
        for (int c = 0; c * 32 < ncols; ++c) {
          uint32_t rr[32];
          tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * C::kBN + c * 32, rr);
          tc::tmem_ld_wait();
          const int col0 = nt * C::kBN + c * 32;
          float z[32];
          float cm = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = col0 + j;
            const bool in = col < p.Vp;
            z[j] = in ? __uint_as_float(rr[j]) + __ldg(p.bias + (in ? col : 0)) : -INFINITY;
            cm = fmaxf(cm, z[j]);
            if (col == p.blank) zb = z[j];
            if (col == label) zl = z[j];
            if (in) ssq = fmaf(z[j], z[j], ssq);
          }
          if (cm > m) {
            s *= __expf(m - cm);
            m = cm;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) s += __expf(z[j] - m);
        }
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&tmem_empty[acc]);
      }
      if (valid) {
        const float lse = m + logf(s);
        const int64_t idx = ((int64_t)b * p.w.ND + t + u) * p.U1 + u;
        p.w.denom[idx] = -lse;
        p.w.lp[idx] = make_float2(zb - lse, label >= 0 ? zl - lse : -INFINITY);
        if (p.sumsq) p.sumsq[((int64_t)b * p.T + t) * p.U1 + u] = ssq;
      }
    }
  } else if (warp >= 8) {
    // ============================ A producers: act(f + g) -> bf16 UMMA tiles ============================
    const int pw = warp - 8;
    int tile_it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_it) {
      const int b = find_utterance(p.tile_offsets, p.B, tile);
      const int Tb = (int)p.act_lens[b], Ub1 = (int)p.label_lens[b] + 1;
      const int r0 = (tile - p.tile_offsets[b]) * kJM;
      const int cells = Tb * Ub1;
      // this warp's 16 rows: row = pw + 8*i ; (t,u) computed once per tile
      const float* frow[kJM / kJProducerWarps];
      const float* grow[kJM / kJProducerWarps];
#pragma unroll
      for (int i = 0; i < kJM / kJProducerWarps; ++i) {
        const int r = r0 + pw + kJProducerWarps * i;
        if (r < cells) {
          const int t = r / Ub1, u = r - t * Ub1;
          frow[i] = p.f + ((int64_t)b * p.T + t) * p.H;
          grow[i] = p.g + ((int64_t)b * p.U1 + u) * p.H;
        } else {
          frow[i] = nullptr;
          grow[i] = nullptr;
        }
      }
      if (tile_it > 0) tc::mbar_wait(a_free, (tile_it - 1) & 1);  // previous tile's MMAs are done with A
      tc::tc_fence_after();
      for (int kb = 0; kb < kblocks; ++kb) {
        uint8_t* ablk = a_smem + kb * C::kABlockBytes;
        const int k = kb * kJK + 2 * lane;
#pragma unroll
        for (int i = 0; i < kJM / kJProducerWarps; ++i) {
          const int row = pw + kJProducerWarps * i;
          float h0 = 0.f, h1 = 0.f;
          if (frow[i]) {
            const float2 fv = __ldg(reinterpret_cast<const float2*>(frow[i] + k));
            const float2 gv = __ldg(reinterpret_cast<const float2*>(grow[i] + k));
            h0 = joint_act(fv.x + gv.x, p.activation);
            h1 = joint_act(fv.y + gv.y, p.activation);
          }
          __nv_bfloat16 hi0, lo0, hi1, lo1;
          tc::split_bf16(h0, hi0, lo0);
          tc::split_bf16(h1, hi1, lo1);
          const uint32_t off = tc::sw128_offset(row, 2 * lane);
          *reinterpret_cast<__nv_bfloat162*>(ablk + off) = __halves2bfloat162(hi0, hi1);
          if (kTerms > 1) *reinterpret_cast<__nv_bfloat162*>(staging + off) = __halves2bfloat162(lo0, lo1);
        }
        if (kTerms > 1) {
          // staging (row-major, swizzled) -> tensor memory (lane = row): thread owns row q*32+lane, half of the block
          asm volatile("bar.sync 1, 256;" ::: "memory");
          const int q = pw & 3, half = pw >> 2;
          const int row = q * 32 + lane;
          uint32_t v[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int chunk = (half * 4 + j) ^ (row & 7);
            const uint4 x = *reinterpret_cast<const uint4*>(staging + row * 128 + chunk * 16);
            v[4 * j + 0] = x.x; v[4 * j + 1] = x.y; v[4 * j + 2] = x.z; v[4 * j + 3] = x.w;
          }
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + C::kAloCol + kb * (kJK / 2) + half * 16;
          uint32_t v0[8], v1[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { v0[j] = v[j]; v1[j] = v[8 + j]; }
          tc::tmem_st8(taddr, v0);
          tc::tmem_st8(taddr + 8, v1);
          tc::tmem_st_wait();
          tc::tc_fence_before();
          asm volatile("bar.sync 1, 256;" ::: "memory");  // staging may be overwritten by the next K block
        }
        tc::fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the UMMA (async proxy) reads
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&a_ready[kb]);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// workspace layout of the fused path
// ------------------------------------------------------------------------------------------------
struct JointWs {
  void* lattice;
  void* w_hi;
  void* w_lo;
  int* tile_offsets;
  int vp_pad;
  size_t total;
};

static inline JointWs joint_ws_carve(void* base, int B, int T, int U1, int H, int Vp, int precision) {
  JointWs j;
  char* p = (char*)base;
  size_t off = 0;
  j.lattice = p + off;
  off += lattice_ws_bytes(B, T, U1);
  j.vp_pad = (Vp + 15) / 16 * 16;
  const size_t wbytes = ((size_t)j.vp_pad * H * 2 + 255) / 256 * 256;
  j.w_hi = p + off; off += wbytes;
  j.w_lo = p + off; off += (precision == CLASR_PREC_BF16X3) ? wbytes : 0;
  j.tile_offsets = (int*)(p + off);
  off += ((size_t)(B + 1) * sizeof(int) + 255) / 256 * 256;
  j.total = off;
  return j;
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
  CLASR_CHECK_ARG(precision == CLASR_PREC_BF16 || precision == CLASR_PREC_BF16X3, "%s: unknown precision %d", who,
                  precision);
  CLASR_CHECK_ARG(ws_bytes >= clasr_joint_workspace_bytes(B, T, U1, H, Vp, precision), "%s: workspace too small", who);
  CLASR_CHECK_ARG((((uintptr_t)ws) & 255) == 0, "%s: workspace must be 256-byte aligned", who);
  CLASR_CHECK_ARG((((uintptr_t)f) & 7) == 0 && (((uintptr_t)g) & 7) == 0, "%s: f/g must be 8-byte aligned", who);
  return CLASR_STATUS_SUCCESS;
}

extern "C" int clasr_joint_rnnt_fwd(const float* f, const float* g, const float* w_out, const float* b_out,
                                    const int64_t* labels, const int64_t* act_lens, const int64_t* label_lens, int B,
                                    int T, int U1, int H, int Vp, int blank, int activation, int precision,
                                    float fastemit_lambda, float* costs, float* sumsq, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  int rc = check_joint_args("joint_rnnt_fwd", f, g, w_out, b_out, labels, act_lens, label_lens, B, T, U1, H, Vp, blank,
                            activation, precision, workspace, workspace_bytes);
  if (rc) return rc;
  CLASR_CHECK_ARG(costs, "joint_rnnt_fwd: null costs");
  cudaStream_t s = (cudaStream_t)stream;
  JointWs jw = joint_ws_carve(workspace, B, T, U1, H, Vp, precision);
  const bool x3 = precision == CLASR_PREC_BF16X3;
  // W_out [Vp,H] fp32 -> bf16 hi[,lo] (rows beyond Vp are never read: TMA zero-fills out-of-bounds rows)
  if ((rc = launch_split_bf16(w_out, Vp, H, H, jw.w_hi, x3 ? jw.w_lo : nullptr, H, s))) return rc;
  joint_tile_offsets_kernel<<<1, 32, 0, s>>>(act_lens, label_lens, B, jw.tile_offsets);
  CLASR_CHECK_LAUNCH("joint_tile_offsets");

  JointFwdParams p;
  p.f = f; p.g = g; p.bias = b_out; p.labels = labels; p.act_lens = act_lens; p.label_lens = label_lens;
  p.tile_offsets = jw.tile_offsets;
  p.B = B; p.T = T; p.U1 = U1; p.H = H; p.Vp = Vp; p.blank = blank; p.activation = activation;
  p.w = lattice_ws_carve(jw.lattice, B, T, U1);
  p.sumsq = sumsq;
  CUtensorMap tw_hi, tw_lo;
  const int bn = x3 ? JointCfg<3>::kBN : JointCfg<1>::kBN;
  if ((rc = make_tmap_bf16_2d(&tw_hi, jw.w_hi, Vp, H, H, bn, kJK))) return rc;
  if (x3) {
    if ((rc = make_tmap_bf16_2d(&tw_lo, jw.w_lo, Vp, H, H, bn, kJK))) return rc;
  } else {
    tw_lo = tw_hi;
  }
  if (x3) {
    const int smem = JointCfg<3>::smem_bytes(H);
    cudaFuncSetAttribute(joint_fwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    joint_fwd_kernel<3><<<kNumSMs, kJThreads, smem, s>>>(tw_hi, tw_lo, p);
  } else {
    const int smem = JointCfg<1>::smem_bytes(H);
    cudaFuncSetAttribute(joint_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    joint_fwd_kernel<1><<<kNumSMs, kJThreads, smem, s>>>(tw_hi, tw_lo, p);
  }
  CLASR_CHECK_LAUNCH("joint_fwd");
  return launch_rnnt_lattice(p.w, act_lens, label_lens, B, T, U1, fastemit_lambda, costs, s);
}

extern "C" int clasr_joint_rnnt_bwd(const float*, const float*, const float*, const float*, const int64_t*,
                                    const int64_t*, const int64_t*, int, int, int, int, int, int, int, int, float,
                                    float, const float*, float*, float*, float*, float*, void*, size_t, void*) {
  set_error("joint_rnnt_bwd: not implemented yet");
  return CLASR_STATUS_INVALID_VALUE;
}
