// gemm_tc.cu — warp-specialised tcgen05 GEMM  C[M,N] (+)= A[M,K] . B[N,K]^T  (bf16 operands, fp32 accumulate in
// TMEM), the building block of the joint backward (dHid = dZ.W, dW = dZ^T.Hid) and the bring-up vehicle for the
// TMA / UMMA-descriptor / TMEM plumbing the fused joint kernels share (tc_common.cuh).
//
// Precision modes (CLASR_PREC_*):
//   BF16    one MMA per product, operands rounded to bf16                  (config 5, "bf16 joint GEMM")
//   BF16X3  operands split x = hi + lo (bf16 each); hi.hi + hi.lo + lo.hi accumulate into the same TMEM tile:
//           ~2^-17 relative per product, i.e. fp32-grade results from the bf16 tensor pipe at 3 MMAs/product
//           (config 2 "fp32": rel 1e-5 loss / 1e-4 grads cannot be met by a single bf16 or tf32 pass).
//
// Structure (one CTA per SM, persistent over output tiles):
//   warp 0  TMA producer   cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier expect_tx
//   warp 1  MMA issuer     one thread issues tcgen05.mma (M=128, N=BN, K=16) and tcgen05.commit
//   warp 2  TMEM allocator 512 columns = 2 accumulator stages x BN(256) fp32 columns
//   warps 4-7 epilogue     tcgen05.ld (32 lanes x 32 columns) -> 128-byte row segments -> global
#include "common.cuh"
#include "tc_common.cuh"
#include <stdlib.h>

namespace clasr {

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                      uint32_t box_rows, uint32_t box_cols) {
  return make_tmap_2d(out, base, rows, cols, row_stride_elems, box_rows, box_cols, 2, 128);
}

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                 uint32_t box_rows, uint32_t box_cols, int elem_bytes, int swizzle_bytes) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
      set_error("cuTensorMapEncodeTiled unavailable: %s", cudaGetErrorString(e));
      return CLASR_STATUS_CUDA_ERROR;
    }
    encode = (EncodeFn)fn;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_elems * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(out, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                           : elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                      const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      swizzle_bytes == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE
                      : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                      : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu stride=%llu box=%ux%u", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)row_stride_elems, box_rows,
              box_cols);
    return CLASR_STATUS_CUDA_ERROR;
  }
  return CLASR_STATUS_SUCCESS;
}

// ------------------------------------------------------------------------------------------------
// fp32 [rows, cols] -> bf16 hi (and lo) [rows, cols_pad], cols_pad multiple of 8, pad zero-filled
// ------------------------------------------------------------------------------------------------
// One thread converts 8 consecutive columns of a row (cols_pad is a multiple of 8): two 128-bit loads when the source row
// is 16-byte aligned, 128-bit stores of the 16-bit halves, 64-bit stores of the e4m3 bytes; one division per 8 elements.
// (The first version handled one element per thread and iteration with a 64-bit division each: 3-4x off the HBM roofline,
// ten launches = 0.16 ms per step.)  Same conversions element for element: the results are bit-identical.
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ src, int64_t rows, int cols,
                                                         int64_t src_ld, __nv_bfloat16* __restrict__ hi,
                                                         __nv_bfloat16* __restrict__ lo, int cols_pad, int f16,
                                                         const float* __restrict__ scale_dev, uint8_t* __restrict__ hi8,
                                                         uint8_t* __restrict__ lo8) {
  const float scale = scale_dev ? __ldg(scale_dev) : 1.f;   // power of two (exact)
  const int groups = cols_pad >> 3;
  const int64_t ng = rows * (int64_t)groups;
  const bool vec_src = ((src_ld & 3) == 0) && ((((uintptr_t)src) & 15) == 0);
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ng; g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = g / groups;
    const int c = (int)(g - r * groups) << 3;
    const float* sp = src + r * src_ld + c;
    float x[8];
    if (vec_src && c + 8 <= cols) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(sp)), b = __ldg(reinterpret_cast<const float4*>(sp) + 1);
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = c + j < cols ? __ldg(sp + j) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] *= scale;
    const int64_t i = r * cols_pad + c;   // multiple of 8 elements: 16-byte aligned in the 2-byte arrays
    uint32_t h[4], l[4];
    if (f16) {   // the same 2-byte slots hold fp16 (CLASR_PREC_FP16X3 / FP16M8)
      float res[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __half h0 = __float2half_rn(x[2 * j]), h1 = __float2half_rn(x[2 * j + 1]);
        res[2 * j] = x[2 * j] - __half2float(h0);
        res[2 * j + 1] = x[2 * j + 1] - __half2float(h1);
        const __half l0 = __float2half_rn(res[2 * j]), l1 = __float2half_rn(res[2 * j + 1]);
        h[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
        l[j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
      }
      if (hi8) {   // FP16M8: the e4m3 correction operands (tc::pack_m8)
        uint32_t h8[2] = {0u, 0u}, l8[2] = {0u, 0u};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t a8 = (uint32_t)(uint8_t)__nv_cvt_float_to_fp8(x[j] * 0.015625f, __NV_SATFINITE, __NV_E4M3);
          const uint32_t b8 = (uint32_t)(uint8_t)__nv_cvt_float_to_fp8(res[j] * 64.f, __NV_SATFINITE, __NV_E4M3);
          h8[j >> 2] |= a8 << (8 * (j & 3));
          l8[j >> 2] |= b8 << (8 * (j & 3));
        }
        *reinterpret_cast<uint2*>(hi8 + i) = make_uint2(h8[0], h8[1]);
        *reinterpret_cast<uint2*>(lo8 + i) = make_uint2(l8[0], l8[1]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat16 h0, l0, h1, l1;
        tc::split_bf16(x[2 * j], h0, l0);
        tc::split_bf16(x[2 * j + 1], h1, l1);
        h[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        l[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
      }
    }
    *reinterpret_cast<uint4*>(hi + i) = make_uint4(h[0], h[1], h[2], h[3]);
    if (lo) *reinterpret_cast<uint4*>(lo + i) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

int launch_split_bf16(const float* src, int64_t rows, int cols, int64_t src_ld, void* hi, void* lo, int cols_pad,
                      cudaStream_t s, int f16, const float* scale_dev, void* hi8 = nullptr, void* lo8 = nullptr) {
  const int64_t n = rows * (int64_t)(cols_pad >> 3);   // one thread per 8 columns
  int grid = (int)((n + 255) / 256);
  if (grid > kNumSMs * 16) grid = kNumSMs * 16;
  if (grid < 1) grid = 1;
  split_bf16_kernel<<<grid, 256, 0, s>>>(src, rows, cols, src_ld, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, cols_pad, f16,
                                         scale_dev, (uint8_t*)hi8, (uint8_t*)lo8);
  CLASR_CHECK_LAUNCH("split_bf16");
  return CLASR_STATUS_SUCCESS;
}

// ------------------------------------------------------------------------------------------------
constexpr int kBM = 128;
constexpr int kBN = 256;
constexpr int kBK = 64;
constexpr int kGemmThreads = 256;

template <int kTerms>
struct GemmSmem {
  static constexpr int kParts = kTerms == 1 ? 1 : 2;
  static constexpr int kABytes = kBM * kBK * 2;  // 16 KB
  static constexpr int kBBytes = kBN * kBK * 2;  // 32 KB
  static constexpr int kStageBytes = kParts * (kABytes + kBBytes);
  static constexpr int kStages = kTerms == 1 ? 4 : 2;
  static constexpr int kRingBytes = kStages * kStageBytes;
  static constexpr int kTotalBytes = kRingBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct GemmParams {
  int M, N, K;
  float* C;
  int64_t ldc;
  int atomic_add;  // 0: C = result ; 1: atomicAdd into C (required when k_splits > 1; C must be pre-zeroed)
  int k_splits;    // the K range is cut into this many slices, each an independent tile
  int a_mn;        // 0: A given as [M,K] (K contiguous) ; 1: A given as [K,M] (M contiguous)
  int b_mn;        // 0: B given as [N,K] (K contiguous) ; 1: B given as [K,N] (N contiguous)
  const int* m_dev;  // optional device-side M (<= M): lets a ragged row count stay on the device (no host sync)
  const int* k_dev;  // optional device-side K (<= K)
  const float* bias; // optional [N]: C = acc + bias[n] (non-atomic epilogue only)
  int group_n;       // pair kernel, no split-K: walk all N tiles of a row block back to back
  int b_keep;        // pair kernel: B is small and re-read by every row block -> TMA loads carry L2 evict_last
  int f16;           // operands are fp16 instead of bf16 (CLASR_PREC_FP16X3)
  const float* alpha_dev;  // optional device scalar: the atomic epilogue adds alpha * acc (undoes an operand pre-scale)
  const float* alpha_dev2; // optional second factor of alpha
};

template <int kTerms>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, GemmParams p) {
  using S = GemmSmem<kTerms>;
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + S::kRingBytes);
  uint64_t* full = bars;                      // [kStages]
  uint64_t* empty = bars + S::kStages;        // [kStages]
  uint64_t* tmem_full = bars + 2 * S::kStages;      // [2]
  uint64_t* tmem_empty = bars + 2 * S::kStages + 2; // [2]
  uint32_t* tmem_base_slot = (uint32_t*)(bars + 2 * S::kStages + 4);

  const int warp = tc::warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  const int Mdyn = (int)tc::uniform_u32((uint32_t)(p.m_dev ? min(p.M, *p.m_dev) : p.M));
  const int Kdyn = (int)tc::uniform_u32((uint32_t)(p.k_dev ? min(p.K, *p.k_dev) : p.K));
  const int m_tiles = (Mdyn + kBM - 1) / kBM;
  const int n_tiles = (p.N + kBN - 1) / kBN;
  const int mn_tiles = m_tiles * n_tiles;
  const int kb_total = (Kdyn + kBK - 1) / kBK;
  const int kb_per_split = max(1, (kb_total + p.k_splits - 1) / p.k_splits);
  const int k_splits = max(1, (kb_total + kb_per_split - 1) / kb_per_split);  // no empty slice
  const int num_tiles = mn_tiles * k_splits;

  if (warp == 0 && tc::elect_one()) {
    tc::prefetch_tmap(&tmA_hi);
    tc::prefetch_tmap(&tmB_hi);
    if (kTerms > 1) { tc::prefetch_tmap(&tmA_lo); tc::prefetch_tmap(&tmB_lo); }
  }
  if (warp == 1 && tc::elect_one()) {
    for (int i = 0; i < S::kStages; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&tmem_full[i], 1); tc::mbar_init(&tmem_empty[i], 4); }
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc(tmem_base_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tc::uniform_u32(*tmem_base_slot);

  if (warp == 0) {
    // ================= TMA producer (whole warp loops, one elected lane issues) =================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int ks = tile / mn_tiles, mn = tile - ks * mn_tiles;
      const int m0 = (mn / n_tiles) * kBM;
      const int n0 = (mn % n_tiles) * kBN;
      const int kb_begin = ks * kb_per_split, kb_end = min(kb_total, kb_begin + kb_per_split);
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        tc::mbar_wait(&empty[stage], phase ^ 1);
        if (tc::elect_one()) {
          uint8_t* st = smem + stage * S::kStageBytes;
          tc::mbar_expect_tx(&full[stage], S::kStageBytes);
#pragma unroll
          for (int part = 0; part < S::kParts; ++part) {
            const CUtensorMap* ta = part == 0 ? &tmA_hi : &tmA_lo;
            const CUtensorMap* tb = part == 0 ? &tmB_hi : &tmB_lo;
            uint8_t* sa = st + part * S::kABytes;
            uint8_t* sb = st + S::kParts * S::kABytes + part * S::kBBytes;
            if (!p.a_mn) {
              tc::tma_load_2d(sa, ta, &full[stage], kb * kBK, m0);
            } else {  // [K rows][64 M] boxes, one per 64-wide M block
#pragma unroll
              for (int j = 0; j < kBM / 64; ++j) tc::tma_load_2d(sa + j * 8192, ta, &full[stage], m0 + 64 * j, kb * kBK);
            }
            if (!p.b_mn) {
              tc::tma_load_2d(sb, tb, &full[stage], kb * kBK, n0);
            } else {
#pragma unroll
              for (int j = 0; j < kBN / 64; ++j) tc::tma_load_2d(sb + j * 8192, tb, &full[stage], n0 + 64 * j, kb * kBK);
            }
          }
        }
        __syncwarp();
        if (++stage == S::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (whole warp loops, one elected lane issues) =================
    const uint32_t idesc = tc::make_idesc_16(kBM, kBN, p.a_mn, p.b_mn, p.f16);
    const uint32_t smem_base = tc::smem_u32(smem);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int ks = tile / mn_tiles;
      const int kb_begin = ks * kb_per_split, kb_end = min(kb_total, kb_begin + kb_per_split);
      tc::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kBN;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        tc::mbar_wait(&full[stage], phase);
        tc::tc_fence_after();
        if (tc::elect_one()) {
          const uint32_t a_hi = smem_base + stage * S::kStageBytes;
          const uint32_t a_lo = a_hi + S::kABytes;
          const uint32_t b_hi = a_hi + S::kParts * S::kABytes;
          const uint32_t b_lo = b_hi + S::kBBytes;
#pragma unroll
          for (int kk = 0; kk < kBK / 16; ++kk) {
            // one MMA consumes 16 K: 32 bytes inside the 128-byte swizzle span (K-major) or two 8-row
            // swizzle atoms = 2048 bytes (MN-major)
            const uint32_t aoff = p.a_mn ? kk * 2048 : kk * 32;
            const uint32_t boff = p.b_mn ? kk * 2048 : kk * 32;
            auto adesc = [&](uint32_t base) {
              return p.a_mn ? tc::make_desc_mnmajor_sw128(base + aoff, 8192) : tc::make_desc_kmajor_sw128(base + aoff);
            };
            auto bdesc = [&](uint32_t base) {
              return p.b_mn ? tc::make_desc_mnmajor_sw128(base + boff, 8192) : tc::make_desc_kmajor_sw128(base + boff);
            };
            const uint32_t first = (kb == kb_begin && kk == 0) ? 0u : 1u;
            tc::umma_ss(d_tmem, adesc(a_hi), bdesc(b_hi), idesc, first);
            if (kTerms > 1) {
              tc::umma_ss(d_tmem, adesc(a_hi), bdesc(b_lo), idesc, 1u);
              tc::umma_ss(d_tmem, adesc(a_lo), bdesc(b_hi), idesc, 1u);
            }
          }
          tc::umma_commit(&empty[stage]);  // smem slot reusable once these MMAs retire
        }
        __syncwarp();
        if (++stage == S::kStages) { stage = 0; phase ^= 1; }
      }
      if (tc::elect_one()) tc::umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ================= epilogue: TMEM -> registers -> global =================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int mn = tile % mn_tiles;
      const int m0 = (mn / n_tiles) * kBM;
      const int n0 = (mn % n_tiles) * kBN;
      tc::mbar_wait(&tmem_full[acc], acc_phase);
      tc::tc_fence_after();
      const int row = m0 + q * 32 + lane;
      float* crow = p.C + (int64_t)row * p.ldc;
      const bool vec_ok = ((p.ldc & 3) == 0) && ((((uintptr_t)p.C) & 15) == 0);
      const float alpha = p.atomic_add ? (p.alpha_dev ? __ldg(p.alpha_dev) : 1.f) * (p.alpha_dev2 ? __ldg(p.alpha_dev2) : 1.f) : 1.f;
      const bool vec32_ok = ((p.ldc & 7) == 0) && ((((uintptr_t)p.C) & 31) == 0);
#pragma unroll 1
      for (int c = 0; c < kBN / 32; ++c) {
        uint32_t r[32];
        tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * kBN + c * 32, r);
        tc::tmem_ld_wait();
        const int col0 = n0 + c * 32;
        if (row < Mdyn && col0 < p.N) {
          if (p.atomic_add) {
            if (vec_ok && col0 + 32 <= p.N) {   // 128-bit vector reductions: a quarter of the atomic instructions
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                red_add_v4(crow + col0 + j, alpha * __uint_as_float(r[j]), alpha * __uint_as_float(r[j + 1]),
                           alpha * __uint_as_float(r[j + 2]), alpha * __uint_as_float(r[j + 3]));
            } else {
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) atomicAdd(crow + col0 + j, alpha * __uint_as_float(r[j]));
            }
          } else if (vec_ok && col0 + 32 <= p.N) {
            if (p.bias) {
#pragma unroll
              for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __ldg(p.bias + col0 + j));
            }
            if (vec32_ok) {  // 32-byte-aligned rows: 256-bit stores, one full sector each
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                const uint32_t v[8] = {r[j], r[j + 1], r[j + 2], r[j + 3], r[j + 4], r[j + 5], r[j + 6], r[j + 7]};
                st_global_256(crow + col0 + j, v);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(crow + col0 + j) =
                    make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                __uint_as_float(r[j + 3]));
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) crow[col0 + j] = __uint_as_float(r[j]) + (p.bias ? __ldg(p.bias + col0 + j) : 0.f);
          }
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tmem_empty[acc]);
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
// CTA-pair variant (cta_group::2): a 2-CTA cluster computes a 256 x 256 tile; each CTA loads its own 128 rows of A
// and HALF of the B tile (128 of the 256 N) — half the L2->smem traffic for B per CTA and half its smem footprint,
// i.e. the arithmetic intensity of a 256 x 256 tile at the smem cost of 128 x 128.  The 1-CTA kernel above moves
// 64 B/clk/SM at full MMA rate in BF16X3 (measured: L2-bound at ~10.8 TB/s); this one 42.7 B/clk/SM.
// ------------------------------------------------------------------------------------------------
// kTerms == 4 (CLASR_PREC_FP16M8): a stage holds {fp16, e4m3 hi8, e4m3 lo8} of each operand — 2 + 1 + 1 bytes per element,
// the same 64 KB as the {hi16, lo16} stage of kTerms == 3:  A16 16 KB | A_h8 8 KB | A_l8 8 KB | B16 16 KB | B_h8 8 KB | B_l8 8 KB
template <int kTerms>
struct Gemm2Smem {
  static constexpr int kParts = kTerms == 1 ? 1 : 2;
  static constexpr int kABytes = kBM * kBK * 2;          // own 128 rows of A: 16 KB
  static constexpr int kBBytes = (kBN / 2) * kBK * 2;    // half of B: 16 KB
  static constexpr int kStageBytes = kParts * (kABytes + kBBytes);
  static constexpr int kStages = kTerms == 1 ? 6 : 3;
  static constexpr int kRingBytes = kStages * kStageBytes;
  static constexpr int kTotalBytes = kRingBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int kTerms>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm2_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                const __grid_constant__ CUtensorMap tmA_l8, const __grid_constant__ CUtensorMap tmB_l8, GemmParams p) {
  // kTerms == 4: tmA_lo / tmB_lo are the e4m3 hi8 maps, tmA_l8 / tmB_l8 the e4m3 lo8 maps
  constexpr bool kM8 = kTerms == 4;
  using S = Gemm2Smem<kTerms>;
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + S::kRingBytes);
  uint64_t* full = bars;                            // [kStages]  (the leader's copy is the one in use)
  uint64_t* empty = bars + S::kStages;              // [kStages]  per CTA: both ring slots of the pair consumed
  uint64_t* tmem_full = bars + 2 * S::kStages;      // [2]  per CTA
  uint64_t* tmem_empty = bars + 2 * S::kStages + 2; // [2]  leader's copy: 8 epilogue warps of the pair
  uint32_t* tmem_base_slot = (uint32_t*)(bars + 2 * S::kStages + 4);

  const int warp = tc::warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = tc::cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int Mdyn = (int)tc::uniform_u32((uint32_t)(p.m_dev ? min(p.M, *p.m_dev) : p.M));
  const int Kdyn = (int)tc::uniform_u32((uint32_t)(p.k_dev ? min(p.K, *p.k_dev) : p.K));
  constexpr int kPM = 2 * kBM;  // rows per pair tile
  const int m_tiles = (Mdyn + kPM - 1) / kPM;
  const int n_tiles = (p.N + kBN - 1) / kBN;
  const int mn_tiles = m_tiles * n_tiles;
  const int kb_total = (Kdyn + kBK - 1) / kBK;
  const int kb_per_split = max(1, (kb_total + p.k_splits - 1) / p.k_splits);
  const int k_splits = max(1, (kb_total + kb_per_split - 1) / kb_per_split);
  const int num_tiles = mn_tiles * k_splits;
  // Work order.  Default: tile ids round-robin over the pairs, so the N tiles of one 256-row block run CONCURRENTLY on
  // neighbouring pairs (the A rows are fetched from HBM once, the other pairs hit under the miss).  group_n = 1 walks
  // them back to back on ONE pair instead; measured slower (dHid 2.66 vs 2.55 ms: one tile in three pays the full HBM
  // miss latency that the 3-stage ring cannot hide), kept as a switch (CLASR_GEMM_GROUP=1).
  const int group = (k_splits == 1 && p.group_n) ? n_tiles : 1;
  const int num_groups = (num_tiles + group - 1) / group;
  auto tile_at = [&](int it) -> int {
    const int g = pair + (it / group) * num_pairs;
    const int t = g * group + it % group;
    return (g < num_groups && t < num_tiles) ? t : -1;
  };

  if (warp == 0 && tc::elect_one()) {
    tc::prefetch_tmap(&tmA_hi);
    tc::prefetch_tmap(&tmB_hi);
    if (kTerms > 1) { tc::prefetch_tmap(&tmA_lo); tc::prefetch_tmap(&tmB_lo); }
    if (kM8) { tc::prefetch_tmap(&tmA_l8); tc::prefetch_tmap(&tmB_l8); }
  }
  if (warp == 1 && tc::elect_one()) {
    for (int i = 0; i < S::kStages; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&tmem_full[i], 1); tc::mbar_init(&tmem_empty[i], 8); }
    tc::fence_barrier_init();
  }
  tc::cluster_sync_all();  // barrier inits visible to the peer before any remote arrive / multicast commit
  if (warp == 2) tc::tmem_alloc_2sm(tmem_base_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tc::uniform_u32(*tmem_base_slot);

  // width of N tile `nt` (multiple of 128 so that each CTA's half is a whole number of 64-wide swizzle atoms)
  auto n_width = [&](int n0) { return min(kBN, (p.N - n0 + 127) / 128 * 128); };

  if (warp == 0) {
    // ================= TMA producer (both CTAs: own A rows, own half of B; bytes land on the leader's barrier) ====
    int stage = 0;
    uint32_t phase = 0;
    const uint64_t pol_keep = tc::l2_policy_evict_last();
    for (int wi = 0, tile; (tile = tile_at(wi)) >= 0; ++wi) {
      const int ks = tile / mn_tiles, mn = tile - ks * mn_tiles;
      const int m0 = (mn / n_tiles) * kPM + (int)cta_rank * kBM;
      const int nbase = (mn % n_tiles) * kBN;
      const int n0 = nbase + (int)cta_rank * (n_width(nbase) / 2);
      const int kb_begin = ks * kb_per_split, kb_end = min(kb_total, kb_begin + kb_per_split);
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        tc::mbar_wait(&empty[stage], phase ^ 1);
        if (tc::elect_one()) {
          uint8_t* st = smem + stage * S::kStageBytes;
          if (leader) tc::mbar_expect_tx(&full[stage], 2 * S::kStageBytes);
          if (kM8) {
            // 8-bit correction operands: K-major = [128 rows][64 B] boxes (64-byte swizzle); MN-major = ONE [64 K rows][128 B]
            // box (128-byte swizzle) for this CTA's 128 rows of A / 128 columns of B
            uint8_t* sa8 = st + S::kABytes;                    // A_h8 | A_l8
            uint8_t* sb8 = st + 2 * S::kABytes + S::kBBytes;   // B_h8 | B_l8
            if (!p.a_mn) {
              tc::tma_load_2d_2sm(sa8, &tmA_lo, &full[stage], kb * kBK, m0);
              tc::tma_load_2d_2sm(sa8 + 8192, &tmA_l8, &full[stage], kb * kBK, m0);
            } else {
              tc::tma_load_2d_2sm(sa8, &tmA_lo, &full[stage], m0, kb * kBK);
              tc::tma_load_2d_2sm(sa8 + 8192, &tmA_l8, &full[stage], m0, kb * kBK);
            }
            if (!p.b_mn) {
              tc::tma_load_2d_2sm(sb8, &tmB_lo, &full[stage], kb * kBK, n0);
              tc::tma_load_2d_2sm(sb8 + 8192, &tmB_l8, &full[stage], kb * kBK, n0);
            } else {
              tc::tma_load_2d_2sm(sb8, &tmB_lo, &full[stage], n0, kb * kBK);
              tc::tma_load_2d_2sm(sb8 + 8192, &tmB_l8, &full[stage], n0, kb * kBK);
            }
          }
#pragma unroll
          for (int part = 0; part < (kM8 ? 1 : S::kParts); ++part) {
            const CUtensorMap* ta = part == 0 ? &tmA_hi : &tmA_lo;
            const CUtensorMap* tb = part == 0 ? &tmB_hi : &tmB_lo;
            uint8_t* sa = st + part * S::kABytes;
            uint8_t* sb = st + S::kParts * S::kABytes + part * S::kBBytes;
            if (!p.a_mn) {
              tc::tma_load_2d_2sm(sa, ta, &full[stage], kb * kBK, m0);
            } else {
#pragma unroll
              for (int j = 0; j < kBM / 64; ++j)
                tc::tma_load_2d_2sm(sa + j * 8192, ta, &full[stage], m0 + 64 * j, kb * kBK);
            }
            if (p.b_keep) {  // (the A operand and the output stream through L2 and would otherwise push B out)
              if (!p.b_mn) {
                tc::tma_load_2d_2sm_hint(sb, tb, &full[stage], kb * kBK, n0, pol_keep);
              } else {
#pragma unroll
                for (int j = 0; j < kBN / 128; ++j)
                  tc::tma_load_2d_2sm_hint(sb + j * 8192, tb, &full[stage], n0 + 64 * j, kb * kBK, pol_keep);
              }
            } else if (!p.b_mn) {
              tc::tma_load_2d_2sm(sb, tb, &full[stage], kb * kBK, n0);
            } else {
#pragma unroll
              for (int j = 0; j < kBN / 128; ++j)
                tc::tma_load_2d_2sm(sb + j * 8192, tb, &full[stage], n0 + 64 * j, kb * kBK);
            }
          }
        }
        __syncwarp();
        if (++stage == S::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && leader) {
    // ================= MMA issuer (leader CTA; one elected lane issues for the pair) =================
    const uint32_t smem_base = tc::smem_u32(smem);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile; (tile = tile_at(it)) >= 0; ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int ks = tile / mn_tiles, mn = tile - ks * mn_tiles;
      const int nw = n_width((mn % n_tiles) * kBN);
      const uint32_t idesc = tc::make_idesc_16(kPM, nw, p.a_mn, p.b_mn, p.f16);
      const int kb_begin = ks * kb_per_split, kb_end = min(kb_total, kb_begin + kb_per_split);
      tc::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kBN;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        tc::mbar_wait(&full[stage], phase);
        tc::tc_fence_after();
        if (tc::elect_one()) {
          const uint32_t a_hi = smem_base + stage * S::kStageBytes;
          const uint32_t a_lo = a_hi + S::kABytes;
          const uint32_t b_hi = a_hi + S::kParts * S::kABytes;
          const uint32_t b_lo = b_hi + S::kBBytes;
#pragma unroll
          for (int kk = 0; kk < kBK / 16; ++kk) {
            const uint32_t aoff = p.a_mn ? kk * 2048 : kk * 32;
            const uint32_t boff = p.b_mn ? kk * 2048 : kk * 32;
            auto adesc = [&](uint32_t base) {
              return p.a_mn ? tc::make_desc_mnmajor_sw128(base + aoff, 8192) : tc::make_desc_kmajor_sw128(base + aoff);
            };
            auto bdesc = [&](uint32_t base) {
              return p.b_mn ? tc::make_desc_mnmajor_sw128(base + boff, 8192) : tc::make_desc_kmajor_sw128(base + boff);
            };
            const uint32_t first = (kb == kb_begin && kk == 0) ? 0u : 1u;
            tc::umma_ss_2sm(d_tmem, adesc(a_hi), bdesc(b_hi), idesc, first);
            if (kTerms == 3) {
              tc::umma_ss_2sm(d_tmem, adesc(a_hi), bdesc(b_lo), idesc, 1u);
              tc::umma_ss_2sm(d_tmem, adesc(a_lo), bdesc(b_hi), idesc, 1u);
            }
          }
          if (kM8) {
            // the two correction terms as dense e4m3 MMAs, K = 32 each: A_h8 . B_l8 + A_l8 . B_h8.  K advance: 32 B inside
            // the 64-byte swizzle span (K-major) or four 8-row swizzle atoms = 4096 B (MN-major, 128-byte rows)
            const uint32_t a_h8 = a_hi + S::kABytes, a_l8 = a_h8 + 8192;
            const uint32_t b_h8 = a_hi + 2 * S::kABytes + S::kBBytes, b_l8 = b_h8 + 8192;
#pragma unroll
            for (int k8 = 0; k8 < kBK / 32; ++k8) {
              auto a8 = [&](uint32_t base) {
                return p.a_mn ? tc::make_desc_mnmajor_sw128(base + k8 * 4096, 8192) : tc::make_desc_kmajor_sw64(base + k8 * 32);
              };
              auto b8 = [&](uint32_t base) {
                return p.b_mn ? tc::make_desc_mnmajor_sw128(base + k8 * 4096, 8192) : tc::make_desc_kmajor_sw64(base + k8 * 32);
              };
              tc::umma_f8_ss_2sm(d_tmem, a8(a_h8), b8(b_l8), idesc, 1u);
              tc::umma_f8_ss_2sm(d_tmem, a8(a_l8), b8(b_h8), idesc, 1u);
            }
          }
          tc::umma_commit_2sm(&empty[stage], 0b11);  // both CTAs' ring slots are reusable once these MMAs retire
        }
        __syncwarp();
        if (++stage == S::kStages) { stage = 0; phase ^= 1; }
      }
      if (tc::elect_one()) tc::umma_commit_2sm(&tmem_full[acc], 0b11);  // accumulators complete in both CTAs
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ================= epilogue (both CTAs: own 128 accumulator rows) =================
    const int q = warp & 3;
    int it = 0;
    for (int tile; (tile = tile_at(it)) >= 0; ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int mn = tile % mn_tiles;
      const int m0 = (mn / n_tiles) * kPM + (int)cta_rank * kBM;
      const int n0 = (mn % n_tiles) * kBN;
      const int nw = n_width(n0);
      tc::mbar_wait(&tmem_full[acc], acc_phase);
      tc::tc_fence_after();
      const int row = m0 + q * 32 + lane;
      float* crow = p.C + (int64_t)row * p.ldc;
      const bool vec_ok = ((p.ldc & 3) == 0) && ((((uintptr_t)p.C) & 15) == 0);
      const float alpha = p.atomic_add ? (p.alpha_dev ? __ldg(p.alpha_dev) : 1.f) * (p.alpha_dev2 ? __ldg(p.alpha_dev2) : 1.f) : 1.f;
      const bool vec32_ok = ((p.ldc & 7) == 0) && ((((uintptr_t)p.C) & 31) == 0);
#pragma unroll 1
      for (int c = 0; c < nw / 32; ++c) {
        uint32_t r[32];
        tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * kBN + c * 32, r);
        tc::tmem_ld_wait();
        const int col0 = n0 + c * 32;
        if (row < Mdyn && col0 < p.N) {
          if (p.atomic_add) {
            if (vec_ok && col0 + 32 <= p.N) {   // 128-bit vector reductions: a quarter of the atomic instructions
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                red_add_v4(crow + col0 + j, alpha * __uint_as_float(r[j]), alpha * __uint_as_float(r[j + 1]),
                           alpha * __uint_as_float(r[j + 2]), alpha * __uint_as_float(r[j + 3]));
            } else {
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) atomicAdd(crow + col0 + j, alpha * __uint_as_float(r[j]));
            }
          } else if (vec_ok && col0 + 32 <= p.N) {
            if (p.bias) {
#pragma unroll
              for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __ldg(p.bias + col0 + j));
            }
            if (vec32_ok) {  // 32-byte-aligned rows: 256-bit stores, one full sector each
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                const uint32_t v[8] = {r[j], r[j + 1], r[j + 2], r[j + 3], r[j + 4], r[j + 5], r[j + 6], r[j + 7]};
                st_global_256(crow + col0 + j, v);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(crow + col0 + j) =
                    make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                __uint_as_float(r[j + 3]));
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) crow[col0 + j] = __uint_as_float(r[j]) + (p.bias ? __ldg(p.bias + col0 + j) : 0.f);
          }
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive_cluster(&tmem_empty[acc], 0);  // the leader's MMA warp owns the accumulator hand-off
    }
  }
  tc::tc_fence_before();
  tc::cluster_sync_all();  // nobody leaves while the peer may still read its smem / signal its barriers
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc_2sm(tmem_base, 512);
  }
}

// C[M,N] (+)= op(A) . op(B)^T from pre-split bf16 operands.
//   a_mn == 0: A is [M, lda] with K contiguous      a_mn == 1: A is [K, lda] with M contiguous
//   b_mn == 0: B is [N, ldb] with K contiguous      b_mn == 1: B is [K, ldb] with N contiguous
// lda / ldb in elements (multiples of 8).  k_splits > 1 requires atomic_add and a zero-initialised C.
// CLASR_PREC_FP16M8: A_lo / B_lo are the e4m3 hi8 arrays and A_l8 / B_l8 the e4m3 lo8 arrays (one BYTE per element, the
// same leading dimensions in elements); CTA-pair kernel only.
int launch_gemm_tc(const void* A_hi, const void* A_lo, int64_t lda, int a_mn, const void* B_hi, const void* B_lo,
                   int64_t ldb, int b_mn, int M, int N, int K, float* C, int64_t ldc, int precision, int atomic_add,
                   int k_splits, cudaStream_t s, const int* m_dev, const int* k_dev, const float* bias,
                   const float* alpha_dev, const float* alpha_dev2, const void* A_l8 = nullptr,
                   const void* B_l8 = nullptr) {
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo, ta_l8, tb_l8;
  int rc;
  const bool m8 = prec_m8(precision);
  const bool x3 = prec_x3(precision) && !m8;
  if (m8 && (!A_l8 || !B_l8)) {
    set_error("gemm_tc: fp16m8 needs the e4m3 operand arrays");
    return CLASR_STATUS_INVALID_VALUE;
  }
  // CTA pairs (256-row tiles) pay off once there are enough rows; CLASR_GEMM_PAIR=0/1 forces the choice (tests)
  static const int force_pair = [] { const char* e = getenv("CLASR_GEMM_PAIR"); return e ? atoi(e) : -1; }();
  const bool use_pair = m8 ? true : (force_pair >= 0 ? force_pair != 0 : (M >= 4 * kBM));
  const int b_box = use_pair ? kBN / 2 : kBN;
  auto mk = [&](CUtensorMap* t, const void* base, int64_t ld, int mn, int rows_mn, int box_mn) -> int {
    if (!mn) return make_tmap_bf16_2d(t, base, rows_mn, K, ld, box_mn, kBK);       // [MN rows, K cols]
    return make_tmap_bf16_2d(t, base, K, rows_mn, ld, kBK, 64);                    // [K rows, MN cols], 64x64 boxes
  };
  if ((rc = mk(&ta_hi, A_hi, lda, a_mn, M, kBM))) return rc;
  if ((rc = mk(&tb_hi, B_hi, ldb, b_mn, N, b_box))) return rc;
  if (x3) {
    if ((rc = mk(&ta_lo, A_lo, lda, a_mn, M, kBM))) return rc;
    if ((rc = mk(&tb_lo, B_lo, ldb, b_mn, N, b_box))) return rc;
  } else {
    ta_lo = ta_hi;
    tb_lo = tb_hi;
  }
  ta_l8 = ta_hi;
  tb_l8 = tb_hi;
  if (m8) {
    // one byte per element.  K-major: [MN rows, K bytes], box [128 rows][64 B], 64-byte swizzle.  MN-major: [K rows, MN
    // bytes], box [64 K rows][128 B], 128-byte swizzle (this CTA's 128 rows / 128 columns are one swizzle span).
    auto mk8 = [&](CUtensorMap* t, const void* base, int64_t ld, int mn, int rows_mn) -> int {
      if (!mn) return make_tmap_2d(t, base, rows_mn, K, ld, 128, kBK, 1, 64);
      return make_tmap_2d(t, base, K, rows_mn, ld, kBK, 128, 1, 128);
    };
    if ((rc = mk8(&ta_lo, A_lo, lda, a_mn, M))) return rc;
    if ((rc = mk8(&ta_l8, A_l8, lda, a_mn, M))) return rc;
    if ((rc = mk8(&tb_lo, B_lo, ldb, b_mn, N))) return rc;
    if ((rc = mk8(&tb_l8, B_l8, ldb, b_mn, N))) return rc;
  }
  const int kb_total = (K + kBK - 1) / kBK;
  if (k_splits <= 0) {
    // auto split-K: fill the persistent grid (74 CTA pairs or 148 CTAs) with as little wave-quantisation loss as
    // possible — 15 output tiles x 11 slices on 74 pairs run 3 rounds at 74 % occupancy, x 24 slices 5 rounds at 97 %
    const int units = use_pair ? kNumSMs / 2 : kNumSMs;
    const int mn = use_pair ? ((M + 2 * kBM - 1) / (2 * kBM)) * ((N + kBN - 1) / kBN)
                            : ((M + kBM - 1) / kBM) * ((N + kBN - 1) / kBN);
    const int s_max = kb_total / 16 > 1 ? (kb_total / 16 < 64 ? kb_total / 16 : 64) : 1;  // >= 16 K blocks per slice
    int best = 1;
    double best_u = 0.0;
    for (int sp = 1; sp <= s_max; ++sp) {
      const int items = mn * sp;
      const double u = (double)items / ((double)((items + units - 1) / units) * units);
      if (u > best_u + 0.02) { best_u = u; best = sp; }  // prefer fewer slices unless clearly better
    }
    k_splits = best;
  }
  if (k_splits > kb_total) k_splits = kb_total;
  const int per = (kb_total + k_splits - 1) / k_splits;
  k_splits = (kb_total + per - 1) / per;  // no empty slice
  if (k_splits > 1 && !atomic_add) {
    set_error("gemm_tc: split-K needs atomic accumulation");
    return CLASR_STATUS_INVALID_VALUE;
  }
  const char* ge = getenv("CLASR_GEMM_GROUP");
  GemmParams p{M, N, K, C, ldc, atomic_add, k_splits, a_mn, b_mn, m_dev, k_dev, atomic_add ? nullptr : bias,
               ge ? atoi(ge) : 0,
               /*b_keep*/ ((size_t)N * K * 2 * (x3 ? 2 : 1) <= ((size_t)16 << 20) && (int64_t)M >= 8 * (int64_t)N) ? 1 : 0,
               prec_f16(precision) ? 1 : 0, alpha_dev, alpha_dev2};
  if (use_pair) {
    const int tiles = ((M + 2 * kBM - 1) / (2 * kBM)) * ((N + kBN - 1) / kBN) * k_splits;
    int pairs = tiles < kNumSMs / 2 ? tiles : kNumSMs / 2;
    if (pairs < 1) pairs = 1;
    if (m8) {
      cudaFuncSetAttribute(gemm2_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, Gemm2Smem<4>::kTotalBytes);
      gemm2_tc_kernel<4><<<2 * pairs, kGemmThreads, Gemm2Smem<4>::kTotalBytes, s>>>(ta_hi, ta_lo, tb_hi, tb_lo, ta_l8,
                                                                                   tb_l8, p);
    } else if (x3) {
      cudaFuncSetAttribute(gemm2_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Gemm2Smem<3>::kTotalBytes);
      gemm2_tc_kernel<3><<<2 * pairs, kGemmThreads, Gemm2Smem<3>::kTotalBytes, s>>>(ta_hi, ta_lo, tb_hi, tb_lo, ta_l8,
                                                                                   tb_l8, p);
    } else {
      cudaFuncSetAttribute(gemm2_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Gemm2Smem<1>::kTotalBytes);
      gemm2_tc_kernel<1><<<2 * pairs, kGemmThreads, Gemm2Smem<1>::kTotalBytes, s>>>(ta_hi, ta_lo, tb_hi, tb_lo, ta_l8,
                                                                                   tb_l8, p);
    }
    CLASR_CHECK_LAUNCH("gemm2_tc");
    return CLASR_STATUS_SUCCESS;
  }
  const int tiles = ((M + kBM - 1) / kBM) * ((N + kBN - 1) / kBN) * k_splits;
  const int grid = tiles < kNumSMs ? tiles : kNumSMs;
  if (x3) {
    cudaFuncSetAttribute(gemm_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmSmem<3>::kTotalBytes);
    gemm_tc_kernel<3><<<grid, kGemmThreads, GemmSmem<3>::kTotalBytes, s>>>(ta_hi, ta_lo, tb_hi, tb_lo, p);
  } else {
    cudaFuncSetAttribute(gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmSmem<1>::kTotalBytes);
    gemm_tc_kernel<1><<<grid, kGemmThreads, GemmSmem<1>::kTotalBytes, s>>>(ta_hi, ta_lo, tb_hi, tb_lo, p);
  }
  CLASR_CHECK_LAUNCH("gemm_tc");
  return CLASR_STATUS_SUCCESS;
}

}  // namespace clasr

using namespace clasr;

static inline int pad8(int k) { return (k + 7) / 8 * 8; }
static inline int pad16(int k) { return (k + 15) / 16 * 16; }

extern "C" size_t clasr_gemm_workspace_bytes(int M, int N, int K, int precision) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const size_t parts = prec_x3(precision) ? 2 : 1;   // fp16m8: fp16 + two e4m3 arrays = the bytes of two 16-bit arrays
  // either orientation of either operand fits: rows x pad16(cols)
  size_t a = ((size_t)pad8(M) * pad16(K) * 2 + 255) / 256 * 256, b = ((size_t)pad8(N) * pad16(K) * 2 + 255) / 256 * 256;
  a = ((size_t)pad16(M) * pad16(K) * 2 + 255) / 256 * 256;
  b = ((size_t)pad16(N) * pad16(K) * 2 + 255) / 256 * 256;
  return parts * (a + b);
}

extern "C" int clasr_gemm_ex(const float* A, const float* B, float* C, int M, int N, int K, int a_trans, int b_trans,
                             int k_splits, int precision, void* workspace, size_t workspace_bytes, void* stream) {
  CLASR_CHECK_ARG(A && B && C && workspace, "gemm: null pointer");
  CLASR_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: non-positive dimension");
  CLASR_CHECK_ARG(prec_ok(precision), "gemm: bad precision");
  CLASR_CHECK_ARG(workspace_bytes >= clasr_gemm_workspace_bytes(M, N, K, precision), "gemm: workspace too small");
  CLASR_CHECK_ARG((((uintptr_t)workspace) & 255) == 0, "gemm: workspace must be 256-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  const bool m8 = prec_m8(precision);
  const bool x3 = prec_x3(precision) && !m8;
  // row pitch: 8 elements (16 bytes) for the 16-bit arrays; fp16m8 pads to 16 so that the 1-byte arrays' rows are 16-byte
  // multiples too (TMA global strides)
  auto padk = [&](int k) { return m8 ? pad16(k) : pad8(k); };
  const size_t a_sz = ((size_t)pad16(M) * pad16(K) * 2 + 255) / 256 * 256;
  const size_t b_sz = ((size_t)pad16(N) * pad16(K) * 2 + 255) / 256 * 256;
  char* w = (char*)workspace;
  void* a_hi = w; w += a_sz;
  void* a_lo = (x3 || m8) ? w : nullptr; if (x3 || m8) w += a_sz;      // fp16m8: hi8 in the first half, lo8 in the second
  void* b_hi = w; w += b_sz;
  void* b_lo = (x3 || m8) ? w : nullptr;
  void* a_l8 = m8 ? (char*)a_lo + a_sz / 2 : nullptr;
  void* b_l8 = m8 ? (char*)b_lo + b_sz / 2 : nullptr;
  // a_trans: A is given as [K, M] row-major; else [M, K].  Same for B with N.
  const int a_rows = a_trans ? K : M, a_cols = a_trans ? M : K;
  const int b_rows = b_trans ? K : N, b_cols = b_trans ? N : K;
  int rc;
  if ((rc = launch_split_bf16(A, a_rows, a_cols, a_cols, a_hi, m8 ? nullptr : a_lo, padk(a_cols), s, prec_f16(precision),
                              nullptr, m8 ? a_lo : nullptr, a_l8)))
    return rc;
  if ((rc = launch_split_bf16(B, b_rows, b_cols, b_cols, b_hi, m8 ? nullptr : b_lo, padk(b_cols), s, prec_f16(precision),
                              nullptr, m8 ? b_lo : nullptr, b_l8)))
    return rc;
  if (k_splits > 1) {
    cudaError_t e = cudaMemsetAsync(C, 0, (size_t)M * N * sizeof(float), s);
    CLASR_CHECK_ARG(e == cudaSuccess, "gemm: memset failed");
  }
  return launch_gemm_tc(a_hi, a_lo, padk(a_cols), a_trans, b_hi, b_lo, padk(b_cols), b_trans, M, N, K, C, N, precision,
                        k_splits > 1, k_splits, s, nullptr, nullptr, nullptr, nullptr, nullptr, a_l8, b_l8);
}

extern "C" int clasr_gemm_nt(const float* A, const float* B, float* C, int M, int N, int K, int precision,
                             void* workspace, size_t workspace_bytes, void* stream) {
  return clasr_gemm_ex(A, B, C, M, N, K, 0, 0, 1, precision, workspace, workspace_bytes, stream);
}

// ------------------------------------------------------------------------------------------------
// Linear layer y = x . W^T + b and its backward on the tcgen05 GEMM (bf16 hi/lo split, fp32-grade results).
// Replaces nn.Linear for the joint's enc / pred projections (reference modules/rnnt.py:1563-1585, 1679-1680) and
// the k=1 Conv1d of the CTC head (modules/conv_asr.py:444-446, 467-469), which torch runs as SIMT sgemm / cuDNN
// convolutions when TF32 is off.
// ------------------------------------------------------------------------------------------------
namespace clasr {
__global__ void colsum_kernel(const float* __restrict__ dy, int64_t M, int N, int64_t rows_per_block,
                              float* __restrict__ db) {
  // block = 32 columns x 8 row lanes; grid = (ceil(N/32), row slices)
  const int col = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ry = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
  float acc = 0.f;
  if (col < N)
    for (int64_t r = r0 + ry; r < r1; r += 8) acc += dy[r * N + col];
  __shared__ float sm[8][33];
  sm[ry][threadIdx.x & 31] = acc;
  __syncthreads();
  if (ry == 0 && col < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x & 31];
    atomicAdd(db + col, t);
  }
}
}  // namespace clasr

static size_t linear_ws_bytes(int M, int N, int K, int precision) {
  // operands of the three GEMMs are split one pair at a time: max over (x,w), (dy,w), (dy,x)
  const size_t parts = prec_x3(precision) ? 2 : 1;
  auto sz = [&](int64_t r, int64_t c) { return ((size_t)pad8((int)r) * pad8((int)c) * 2 + 255) / 256 * 256; };
  const size_t x = sz(M, K), w = sz(N, K), dy = sz(M, N);
  return parts * (x + w + dy);
}

extern "C" size_t clasr_linear_workspace_bytes(int M, int N, int K, int precision) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  return linear_ws_bytes(M, N, K, precision);
}

struct LinearWs { void *x_hi, *x_lo, *w_hi, *w_lo, *dy_hi, *dy_lo; };
static LinearWs linear_ws_carve(void* base, int M, int N, int K, int precision) {
  const bool x3 = prec_x3(precision);
  auto sz = [&](int64_t r, int64_t c) { return ((size_t)pad8((int)r) * pad8((int)c) * 2 + 255) / 256 * 256; };
  char* p = (char*)base;
  LinearWs w;
  w.x_hi = p; p += sz(M, K); w.x_lo = x3 ? p : nullptr; if (x3) p += sz(M, K);
  w.w_hi = p; p += sz(N, K); w.w_lo = x3 ? p : nullptr; if (x3) p += sz(N, K);
  w.dy_hi = p; p += sz(M, N); w.dy_lo = x3 ? p : nullptr;
  return w;
}

extern "C" int clasr_linear_fwd(const float* x, const float* w, const float* bias, float* y, int M, int N, int K,
                                int precision, void* workspace, size_t workspace_bytes, void* stream) {
  CLASR_CHECK_ARG(x && w && y && workspace, "linear_fwd: null pointer");
  CLASR_CHECK_ARG(M > 0 && N > 0 && K > 0, "linear_fwd: non-positive dimension");
  CLASR_CHECK_ARG(prec_ok(precision) && !prec_m8(precision), "linear_fwd: bad precision (fp16m8 is a joint-backward mode)");
  CLASR_CHECK_ARG(workspace_bytes >= linear_ws_bytes(M, N, K, precision), "linear_fwd: workspace too small");
  CLASR_CHECK_ARG((((uintptr_t)workspace) & 255) == 0, "linear_fwd: workspace must be 256-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  LinearWs ws = linear_ws_carve(workspace, M, N, K, precision);
  int rc;
  if ((rc = launch_split_bf16(x, M, K, K, ws.x_hi, ws.x_lo, pad8(K), s, prec_f16(precision), nullptr))) return rc;
  if ((rc = launch_split_bf16(w, N, K, K, ws.w_hi, ws.w_lo, pad8(K), s, prec_f16(precision), nullptr))) return rc;
  prof_begin("linear_fwd", s);
  rc = launch_gemm_tc(ws.x_hi, ws.x_lo, pad8(K), 0, ws.w_hi, ws.w_lo, pad8(K), 0, M, N, K, y, N, precision, 0, 1, s,
                      nullptr, nullptr, bias, nullptr, nullptr);
  prof_end("linear_fwd", s);
  return rc;
}

// dx [M,K] = dy . W ; dw [N,K] = dy^T . x ; db [N] = column sums of dy.  Any of dx / dw / db may be NULL.
// The workspace must be the one clasr_linear_fwd filled (x and w splits are reused).
extern "C" int clasr_linear_bwd(const float* dy, float* dx, float* dw, float* db, int M, int N, int K, int precision,
                                void* workspace, size_t workspace_bytes, void* stream) {
  CLASR_CHECK_ARG(dy && workspace, "linear_bwd: null pointer");
  CLASR_CHECK_ARG(M > 0 && N > 0 && K > 0, "linear_bwd: non-positive dimension");
  CLASR_CHECK_ARG(prec_ok(precision), "linear_bwd: bad precision");
  CLASR_CHECK_ARG(workspace_bytes >= linear_ws_bytes(M, N, K, precision), "linear_bwd: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  LinearWs ws = linear_ws_carve(workspace, M, N, K, precision);
  int rc;
  if (dx || dw)
    if ((rc = launch_split_bf16(dy, M, N, N, ws.dy_hi, ws.dy_lo, pad8(N), s, prec_f16(precision), nullptr))) return rc;
  prof_begin("linear_bwd", s);
  if (dx) {  // A = dy [M, N] K-major (K_gemm = N); B = W given as [K_gemm = N rows, N_gemm = K cols] (MN-major)
    if ((rc = launch_gemm_tc(ws.dy_hi, ws.dy_lo, pad8(N), 0, ws.w_hi, ws.w_lo, pad8(K), 1, M, K, N, dx, K, precision, 0, 1,
                             s, nullptr, nullptr, nullptr, nullptr, nullptr)))
      return rc;
  }
  if (dw) {  // A = dy given as [K_gemm = M rows, M_gemm = N cols]; B = x given as [K_gemm = M rows, N_gemm = K cols]
    // split-K over the rows with launch_gemm_tc's own heuristic (>= 16 K blocks per slice, best fill of the persistent
    // grid): a slice pays a whole 256 x 256 tile of fp32 reductions, so short slices (B_local = 4: 16 K blocks in all)
    // cost more in atomics than they gain in parallelism.  The accumulator starts from zero either way.
    cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)N * K * sizeof(float), s);
    CLASR_CHECK_ARG(e == cudaSuccess, "linear_bwd: memset failed");
    if ((rc = launch_gemm_tc(ws.dy_hi, ws.dy_lo, pad8(N), 1, ws.x_hi, ws.x_lo, pad8(K), 1, N, K, M, dw, K, precision,
                             1, /*auto split-K*/ 0, s, nullptr, nullptr, nullptr, nullptr, nullptr)))
      return rc;
  }
  if (db) {
    cudaError_t e = cudaMemsetAsync(db, 0, (size_t)N * sizeof(float), s);
    CLASR_CHECK_ARG(e == cudaSuccess, "linear_bwd: memset failed");
    const int64_t rows_per_block = 256;
    dim3 grid((N + 31) / 32, (unsigned)((M + rows_per_block - 1) / rows_per_block));
    colsum_kernel<<<grid, 256, 0, s>>>(dy, M, N, rows_per_block, db);
    CLASR_CHECK_LAUNCH("linear_colsum");
  }
  prof_end("linear_bwd", s);
  return CLASR_STATUS_SUCCESS;
}

// ------------------------------------------------------------------------------------------------
// y[b, c, r] = x[b, r, c]: 32 x 32 tiles through shared memory, 128-byte coalesced on both sides
// ------------------------------------------------------------------------------------------------
namespace clasr {
__global__ void __launch_bounds__(256) transpose_last2_kernel(const float* __restrict__ x, float* __restrict__ y, int R,
                                                              int C) {
  __shared__ float tile[32][33];
  const int64_t b = blockIdx.z;
  const float* xb = x + b * (int64_t)R * C;
  float* yb = y + b * (int64_t)R * C;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8 threads
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int r = r0 + ty + i, c = c0 + tx;
    if (r < R && c < C) tile[ty + i][tx] = xb[(int64_t)r * C + c];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int c = c0 + ty + i, r = r0 + tx;
    if (r < R && c < C) yb[(int64_t)c * R + r] = tile[tx][ty + i];
  }
}
}  // namespace clasr

extern "C" int clasr_transpose_last2(const float* x, float* y, int64_t B, int R, int C, void* stream) {
  CLASR_CHECK_ARG(x && y, "transpose_last2: null pointer");
  CLASR_CHECK_ARG(B > 0 && R > 0 && C > 0 && B <= 65535, "transpose_last2: bad dimensions");
  const dim3 grid((C + 31) / 32, (R + 31) / 32, (unsigned)B);
  clasr::transpose_last2_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, R, C);
  CLASR_CHECK_LAUNCH("transpose_last2");
  return CLASR_STATUS_SUCCESS;
}

// ------------------------------------------------------------------------------------------------
// Micro-benchmark (design tool, not on the product path): cycles per tcgen05.mma M=128 x N x K=16 for a given
// operand pattern, issued back to back by one thread with nothing else running on the SM.
//   pattern 0: SS only   1: TS only (A from TMEM)   2: SS,SS,TS (the BF16X3 joint sequence)
// ------------------------------------------------------------------------------------------------
namespace clasr {
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int pattern, int iters, long long* out) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = tc::warp_idx_uniform();
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (warp == 0 && tc::elect_one()) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  if (warp == 1) tc::tmem_alloc(&slot, 512);
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tc::uniform_u32(slot);
  if (warp == 0) {
    const uint32_t idesc = tc::make_idesc_bf16(128, N);
    const uint32_t a = tc::smem_u32(smem), b = a + 16384;
    long long t0 = 0, t1 = 0;
    if (tc::elect_one()) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t da = tc::make_desc_kmajor_sw128(a + kk * 32), db = tc::make_desc_kmajor_sw128(b + kk * 32);
          if (pattern == 0) {
            tc::umma_ss(tmem_base, da, db, idesc, 1u);
          } else if (pattern == 1) {
            tc::umma_ts(tmem_base, tmem_base + 256 + kk * 8, db, idesc, 1u);
          } else {
            tc::umma_ss(tmem_base, da, db, idesc, 1u);
            tc::umma_ss(tmem_base, da, db, idesc, 1u);
            tc::umma_ts(tmem_base, tmem_base + 256 + kk * 8, db, idesc, 1u);
          }
        }
      }
      tc::umma_commit(&bar);
    }
    __syncwarp();
    tc::mbar_wait(&bar, 0);
    if (tc::elect_one()) {
      t1 = clock64();
      out[0] = t1 - t0;
      out[1] = (long long)iters * 4 * (pattern == 2 ? 3 : 1);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, 512); }
}
}  // namespace clasr

extern "C" int clasr_debug_mma_rate(int N, int pattern, int iters, long long* out_dev, void* stream) {
  CLASR_CHECK_ARG(N >= 16 && N <= 256 && N % 16 == 0 && out_dev, "debug_mma_rate: bad arguments");
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  mma_rate_kernel<<<1, 128, 64 * 1024, (cudaStream_t)stream>>>(N, pattern, iters, out_dev);
  CLASR_CHECK_LAUNCH("mma_rate");
  return CLASR_STATUS_SUCCESS;
}
