// Probe (NOT part of the product library): one tcgen05 block-scaled MMA tile, kind::mxf8f6f4, e4m3 x e4m3 with UE8M0
// scales per 32 elements of K, checked against a host evaluation of the same quantised operands.  Purpose: pin down
// the scale-factor path (canonical 512-byte SF atom in smem -> tcgen05.cp.32x128b.warpx4 -> TMEM word per row, sf_id
// selecting the byte) before the correction terms of the BF16X3 split are moved to fp8.
//
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I indic_cl_asr_b200/csrc \
//        tools/experimental/mxf8_probe.cu -o tools/experimental/mxf8_probe -lcuda && tools/experimental/mxf8_probe
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "tc_common.cuh"

namespace tc = clasr::tc;

constexpr int kM = 128, kN = 128, kK = 128;   // one 128-byte swizzle row of fp8 = the whole K

__device__ __forceinline__ uint64_t make_desc_sf(uint32_t smem_addr) {  // no swizzle, 8 x 16 B core matrices, SBO 128 B
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(128 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// InstrDescriptorBlockScaled: b_sf_id [4,6), a_format [7,10) (0 = e4m3), b_format [10,13), a_major [15], b_major [16],
// n_dim [17,23) = N >> 3, scale_format [23] (1 = UE8M0), m_dim [24,29) = M >> 4, a_sf_id [29,31), k_size [31] = 0 (K32)
__device__ __forceinline__ uint32_t make_idesc_mx(int M, int N, int sf_id, int b_mn_major = 0) {
  return ((uint32_t)sf_id << 4) | ((uint32_t)(b_mn_major & 1) << 16) | ((uint32_t)(N >> 3) << 17) | (1u << 23) |
         ((uint32_t)(M >> 4) << 24) | ((uint32_t)sf_id << 29);
}

// mixed != 0: D = A16 . B16^T (kind::f16, bf16 operands, K = 64) FIRST, then the four block-scaled MMAs accumulate on
// top of it in the same TMEM accumulator.  timing[0..1] = cycles for 256 back-to-back kind::f16 / mxf8f6f4 MMAs.
__global__ void __launch_bounds__(128) probe_kernel(const uint8_t* __restrict__ a8, const uint8_t* __restrict__ b8,
                                                    const uint8_t* __restrict__ sfa, const uint8_t* __restrict__ sfb,
                                                    const __nv_bfloat16* __restrict__ a16,
                                                    const __nv_bfloat16* __restrict__ b16, int mixed,
                                                    float* __restrict__ c, long long* __restrict__ timing) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
  uint8_t* a_s = smem;                 // [128 rows][128 B], SW128 K-major
  uint8_t* b_s = smem + 16384;
  uint8_t* sfa_s = smem + 32768;       // canonical SF atom: (row % 32) * 16 + (row / 32) * 4 + kblock
  uint8_t* sfb_s = sfa_s + 512;
  uint64_t* bar = (uint64_t*)(sfb_s + 512);
  uint32_t* slot = (uint32_t*)(bar + 4);
  uint8_t* a16_s = smem + 34816;       // [128 rows][64 bf16 = 128 B], SW128 K-major (1024-aligned)
  uint8_t* b16_s = a16_s + 16384;
  uint8_t* bt_s = b16_s + 16384;       // mixed == 2: B as an MN-major e4m3 operand: [K rows][128 N bytes], 8-row SW128 atoms
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < kM * 8; i += 128) {   // 16-byte chunks: chunk index XOR (row & 7)
    const int row = i >> 3, ch = i & 7;
    *reinterpret_cast<uint4*>(a_s + row * 128 + ((ch ^ (row & 7)) * 16)) =
        *reinterpret_cast<const uint4*>(a8 + row * kK + ch * 16);
    *reinterpret_cast<uint4*>(b_s + row * 128 + ((ch ^ (row & 7)) * 16)) =
        *reinterpret_cast<const uint4*>(b8 + row * kK + ch * 16);
  }
  for (int i = tid; i < kM * 8; i += 128) {
    const int row = i >> 3, ch = i & 7;
    *reinterpret_cast<uint4*>(a16_s + row * 128 + ((ch ^ (row & 7)) * 16)) =
        *reinterpret_cast<const uint4*>(a16 + row * 64 + ch * 8);
    *reinterpret_cast<uint4*>(b16_s + row * 128 + ((ch ^ (row & 7)) * 16)) =
        *reinterpret_cast<const uint4*>(b16 + row * 64 + ch * 8);
  }
  for (int i = tid; i < kK * 8; i += 128) {   // element (k, n) = b8[n][k]; row k, 16-byte chunk of n XOR (k & 7)
    const int k = i >> 3, ch = i & 7;
    uint8_t tmp[16];
    for (int j = 0; j < 16; ++j) tmp[j] = b8[(ch * 16 + j) * kK + k];
    *reinterpret_cast<uint4*>(bt_s + k * 128 + ((ch ^ (k & 7)) * 16)) = *reinterpret_cast<uint4*>(tmp);
  }
  for (int i = tid; i < 512; i += 128) {
    const int row = i >> 2, kb = i & 3;
    sfa_s[(row & 31) * 16 + (row >> 5) * 4 + kb] = sfa[row * 4 + kb];
    sfb_s[(row & 31) * 16 + (row >> 5) * 4 + kb] = sfb[row * 4 + kb];
  }
  if (tid == 0) { tc::mbar_init(bar, 1); tc::mbar_init(bar + 1, 1); tc::mbar_init(bar + 2, 1); tc::fence_barrier_init(); }
  if (warp == 0) tc::tmem_alloc(slot, 512);
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *slot;
  const uint32_t t_sfa = tmem + 256, t_sfb = tmem + 260;
  if (warp == 0) {
    if (tc::elect_one()) {
      asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(t_sfa), "l"(make_desc_sf(tc::smem_u32(sfa_s))));
      asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(t_sfb), "l"(make_desc_sf(tc::smem_u32(sfb_s))));
      if (mixed) {
        // mixed == 3: A holds fp16 bits, B bf16 (a_format [7,10) = 0, b_format [10,13) = 1), nothing else accumulated
        const uint32_t id16 = mixed == 3 ? (tc::make_idesc_bf16(kM, kN) & ~(7u << 7)) : tc::make_idesc_bf16(kM, kN);
        for (int kk = 0; kk < 4; ++kk)
          tc::umma_ss(tmem, tc::make_desc_kmajor_sw128(tc::smem_u32(a16_s) + kk * 32),
                      tc::make_desc_kmajor_sw128(tc::smem_u32(b16_s) + kk * 32), id16, kk > 0);
      }
      for (int kb = 0; kb < (mixed == 3 ? 0 : kK / 32); ++kb) {
        const uint64_t da = tc::make_desc_kmajor_sw128(tc::smem_u32(a_s) + kb * 32);
        // MN-major: 32 K rows per MMA = four 1024-byte atoms; one 128-element MN block, so LBO is never used
        const uint64_t db = mixed == 2 ? tc::make_desc_mnmajor_sw128(tc::smem_u32(bt_s) + kb * 4096, 16384)
                                       : tc::make_desc_kmajor_sw128(tc::smem_u32(b_s) + kb * 32);
        const uint32_t idesc = make_idesc_mx(kM, kN, kb, mixed == 2);
        const uint32_t acc = (kb > 0) || mixed;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n\t}\n" ::"r"(tmem),
            "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(t_sfa), "r"(t_sfb)
            : "memory");
      }
      tc::umma_commit(bar);
    }
    __syncwarp();
  }
  tc::mbar_wait(bar, 0);
  tc::tc_fence_after();
  const int row = warp * 32 + (tid & 31);
  for (int cb = 0; cb < kN / 32; ++cb) {
    uint32_t r[32];
    tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + cb * 32, r);
    tc::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) c[row * kN + cb * 32 + j] = __uint_as_float(r[j]);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0 && timing) {   // issue rate: 256 MMAs of each kind into a scratch accumulator (columns 128..255)
    tc::tc_fence_after();
    for (int kind = 0; kind < 2; ++kind) {
      const long long t0 = clock64();
      if (tc::elect_one()) {
        for (int it = 0; it < 256; ++it) {
          if (kind == 0) {
            tc::umma_ss(tmem + 128, tc::make_desc_kmajor_sw128(tc::smem_u32(a16_s) + (it & 3) * 32),
                        tc::make_desc_kmajor_sw128(tc::smem_u32(b16_s) + (it & 3) * 32), tc::make_idesc_bf16(kM, kN), 1u);
          } else {
            const uint32_t idesc = make_idesc_mx(kM, kN, it & 3);
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n\t}\n" ::"r"(tmem + 128),
                "l"(tc::make_desc_kmajor_sw128(tc::smem_u32(a_s) + (it & 3) * 32)),
                "l"(tc::make_desc_kmajor_sw128(tc::smem_u32(b_s) + (it & 3) * 32)), "r"(idesc), "r"(1u), "r"(t_sfa), "r"(t_sfb)
                : "memory");
          }
        }
        tc::umma_commit(bar + 1 + kind);
      }
      __syncwarp();
      tc::mbar_wait(bar + 1 + kind, 0);
      tc::tc_fence_after();
      if (tid == 0) timing[kind] = clock64() - t0;
    }
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}


// ---- CTA-pair variant: cluster of 2, tcgen05 cta_group::2, M = 256 (128 rows per CTA), N = 256 (each CTA holds half of
// B), K = 128.  Hypothesis under test (from CUTLASS' TileShape_SF): every CTA keeps SFA for ITS 128 rows and SFB for
// ALL 256 columns in its own shared memory at the same offsets, and the leader's cta_group::2 tcgen05.cp moves both
// CTAs' copies into their own tensor memories.
constexpr int kN2 = 256;
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
probe2_kernel(const uint8_t* __restrict__ a8, const uint8_t* __restrict__ b8, const uint8_t* __restrict__ sfa,
              const uint8_t* __restrict__ sfb, float* __restrict__ c) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
  uint8_t* a_s = smem;                 // this CTA's 128 rows of A
  uint8_t* b_s = smem + 16384;         // this CTA's 128 columns (rows of B) = half of N
  uint8_t* sfa_s = smem + 32768;       // 1 atom: this CTA's rows
  uint8_t* sfb_s = sfa_s + 512;        // 2 atoms: all 256 columns
  uint64_t* bar = (uint64_t*)(sfb_s + 1024);
  uint32_t* slot = (uint32_t*)(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = tc::cluster_ctarank();
  const uint8_t* a_src = a8 + (size_t)rank * 128 * kK;
  const uint8_t* b_src = b8 + (size_t)rank * 128 * kK;
  for (int i = tid; i < 128 * 8; i += 128) {
    const int row = i >> 3, ch = i & 7;
    *reinterpret_cast<uint4*>(a_s + row * 128 + ((ch ^ (row & 7)) * 16)) = *reinterpret_cast<const uint4*>(a_src + row * kK + ch * 16);
    *reinterpret_cast<uint4*>(b_s + row * 128 + ((ch ^ (row & 7)) * 16)) = *reinterpret_cast<const uint4*>(b_src + row * kK + ch * 16);
  }
  for (int i = tid; i < 512; i += 128) {
    const int row = i >> 2, kb = i & 3;
    sfa_s[(row & 31) * 16 + (row >> 5) * 4 + kb] = sfa[(rank * 128 + row) * 4 + kb];
  }
  for (int i = tid; i < 1024; i += 128) {
    const int col = i >> 2, kb = i & 3;   // column 0..255 -> atom col / 128
    sfb_s[(col >> 7) * 512 + (col & 31) * 16 + ((col >> 5) & 3) * 4 + kb] = sfb[col * 4 + kb];
  }
  if (tid == 0) { tc::mbar_init(bar, 1); tc::fence_barrier_init(); }
  tc::cluster_sync_all();
  if (warp == 0) tc::tmem_alloc_2sm(slot, 512);
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync_all();             // both CTAs' operands and scale atoms are in place
  tc::tc_fence_after();
  const uint32_t tmem = *slot;
  const uint32_t t_sfa = tmem + 256, t_sfb = tmem + 264;
  if (rank == 0 && warp == 0) {
    if (tc::elect_one()) {
      asm volatile("tcgen05.cp.cta_group::2.32x128b.warpx4 [%0], %1;" ::"r"(t_sfa), "l"(make_desc_sf(tc::smem_u32(sfa_s))));
      asm volatile("tcgen05.cp.cta_group::2.32x128b.warpx4 [%0], %1;" ::"r"(t_sfb), "l"(make_desc_sf(tc::smem_u32(sfb_s))));
      asm volatile("tcgen05.cp.cta_group::2.32x128b.warpx4 [%0], %1;" ::"r"(t_sfb + 4), "l"(make_desc_sf(tc::smem_u32(sfb_s) + 512)));
      for (int kb = 0; kb < kK / 32; ++kb) {
        const uint64_t da = tc::make_desc_kmajor_sw128(tc::smem_u32(a_s) + kb * 32);
        const uint64_t db = tc::make_desc_kmajor_sw128(tc::smem_u32(b_s) + kb * 32);
        const uint32_t idesc = make_idesc_mx(256, kN2, kb);
        const uint32_t acc = kb > 0;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n\t}\n" ::"r"(tmem),
            "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(t_sfa), "r"(t_sfb)
            : "memory");
      }
      tc::umma_commit_2sm(bar, 0b11);   // arrives on both CTAs' barriers
    }
    __syncwarp();
  }
  tc::mbar_wait(bar, 0);
  tc::tc_fence_after();
  const int row = rank * 128 + warp * 32 + (tid & 31);
  for (int cb = 0; cb < kN2 / 32; ++cb) {
    uint32_t r[32];
    tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + cb * 32, r);
    tc::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) c[row * kN2 + cb * 32 + j] = __uint_as_float(r[j]);
  }
  tc::tc_fence_before();
  tc::cluster_sync_all();
  if (warp == 0) tc::tmem_dealloc_2sm(tmem, 512);
}

// ---- host: e4m3 encode / decode, block quantisation
static float e4m3_decode(uint8_t v) {
  const int s = v >> 7, e = (v >> 3) & 15, m = v & 7;
  float x = e == 0 ? std::ldexp((float)m, -9) : std::ldexp(1.f + m / 8.f, e - 7);
  return s ? -x : x;
}
static uint8_t e4m3_encode(float x) {   // nearest by search (the probe only needs correctness)
  uint8_t best = 0;
  float bd = 1e30f;
  for (int v = 0; v < 256; ++v) {
    if ((v & 0x7f) == 0x7f) continue;   // NaN
    const float d = std::fabs(e4m3_decode((uint8_t)v) - x);
    if (d < bd) { bd = d; best = (uint8_t)v; }
  }
  return best;
}
static void quantise(const std::vector<float>& x, int rows, std::vector<uint8_t>& q, std::vector<uint8_t>& sf) {
  q.resize((size_t)rows * kK);
  sf.resize((size_t)rows * 4);
  for (int r = 0; r < rows; ++r)
    for (int kb = 0; kb < 4; ++kb) {
      float amax = 0.f;
      for (int k = 0; k < 32; ++k) amax = std::fmax(amax, std::fabs(x[(size_t)r * kK + kb * 32 + k]));
      int e = amax > 0.f ? (int)std::ceil(std::log2(amax / 448.f)) : 0;
      sf[r * 4 + kb] = (uint8_t)(e + 127);
      for (int k = 0; k < 32; ++k) q[(size_t)r * kK + kb * 32 + k] = e4m3_encode(std::ldexp(x[(size_t)r * kK + kb * 32 + k], -e));
    }
}

int main(int argc, char** argv) {
  const bool try_mixed_formats = argc > 1 && std::string(argv[1]) == "--mixed-formats";
  std::vector<float> A((size_t)kM * kK), B((size_t)kN * kK);
  srand(1);
  for (auto& v : A) v = (rand() / (float)RAND_MAX * 2 - 1) * std::ldexp(1.f, rand() % 24 - 20);   // wide dynamic range
  for (auto& v : B) v = (rand() / (float)RAND_MAX * 2 - 1) * std::ldexp(1.f, rand() % 12 - 10);
  std::vector<uint8_t> a8, b8, sfa, sfb;
  quantise(A, kM, a8, sfa);
  quantise(B, kN, b8, sfb);
  std::vector<double> ref((size_t)kM * kN);
  for (int m = 0; m < kM; ++m)
    for (int n = 0; n < kN; ++n) {
      double acc = 0;
      for (int kb = 0; kb < 4; ++kb) {
        double part = 0;
        for (int k = 0; k < 32; ++k) part += (double)e4m3_decode(a8[m * kK + kb * 32 + k]) * e4m3_decode(b8[n * kK + kb * 32 + k]);
        acc += part * std::ldexp(1.0, sfa[m * 4 + kb] - 127) * std::ldexp(1.0, sfb[n * 4 + kb] - 127);
      }
      ref[(size_t)m * kN + n] = acc;
    }
  std::vector<__nv_bfloat16> A16((size_t)kM * 64), B16((size_t)kN * 64);
  for (auto& v : A16) v = __float2bfloat16(rand() / (float)RAND_MAX * 2 - 1);
  for (auto& v : B16) v = __float2bfloat16(rand() / (float)RAND_MAX * 2 - 1);
  std::vector<double> ref16((size_t)kM * kN);
  for (int m = 0; m < kM; ++m)
    for (int n = 0; n < kN; ++n) {
      double acc = 0;
      for (int k = 0; k < 64; ++k) acc += (double)__bfloat162float(A16[m * 64 + k]) * __bfloat162float(B16[n * 64 + k]);
      ref16[(size_t)m * kN + n] = acc;
    }
  __nv_bfloat16 *da16, *db16;
  long long* dt;
  cudaMalloc(&da16, A16.size() * 2); cudaMalloc(&db16, B16.size() * 2); cudaMalloc(&dt, 16);
  cudaMemcpy(da16, A16.data(), A16.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db16, B16.data(), B16.size() * 2, cudaMemcpyHostToDevice);
  uint8_t *da, *db, *dsa, *dsb;
  float* dc;
  cudaMalloc(&da, a8.size()); cudaMalloc(&db, b8.size()); cudaMalloc(&dsa, sfa.size()); cudaMalloc(&dsb, sfb.size());
  cudaMalloc(&dc, ref.size() * 4);
  cudaMemcpy(da, a8.data(), a8.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(db, b8.data(), b8.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dsa, sfa.data(), sfa.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dsb, sfb.data(), sfb.size(), cudaMemcpyHostToDevice);
  const int smem = 34816 + 32768 + 16384 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int rc_total = 0;
  for (int mixed = 0; mixed < 3; ++mixed) {
  printf("---- %s\n", mixed == 0 ? "4 x kind::mxf8f6f4 alone"
                      : mixed == 1 ? "kind::f16 (bf16, K=64) then 4 x kind::mxf8f6f4 into the SAME accumulator"
                                   : "the same with B consumed as an MN-major e4m3 operand");
  if (mixed == 1) for (size_t i = 0; i < ref.size(); ++i) ref[i] += ref16[i];
  probe_kernel<<<1, 128, smem>>>(da, db, dsa, dsb, da16, db16, mixed, dc, mixed == 1 ? dt : nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> got(ref.size());
  cudaMemcpy(got.data(), dc, got.size() * 4, cudaMemcpyDeviceToHost);
  double maxref = 0, maxerr = 0;
  int bad_m = -1, bad_n = -1;
  for (size_t i = 0; i < ref.size(); ++i) {
    maxref = std::fmax(maxref, std::fabs(ref[i]));
    const double er = std::fabs(got[i] - ref[i]);
    if (er > maxerr) { maxerr = er; bad_m = (int)(i / kN); bad_n = (int)(i % kN); }
  }
  printf("max|ref| = %.6e  max|err| = %.6e  rel = %.3e  (worst at m=%d n=%d: got %.6e ref %.6e)\n", maxref, maxerr,
         maxerr / maxref, bad_m, bad_n, got[(size_t)bad_m * kN + bad_n], ref[(size_t)bad_m * kN + bad_n]);
  // per-row relative error: a wrong SF row mapping shows up as whole rows / columns off by powers of two
  int rows_off = 0;
  for (int m = 0; m < kM; ++m) {
    double rm = 0, em = 0;
    for (int n = 0; n < kN; ++n) { rm = std::fmax(rm, std::fabs(ref[(size_t)m * kN + n])); em = std::fmax(em, std::fabs(got[(size_t)m * kN + n] - ref[(size_t)m * kN + n])); }
    if (em > 1e-5 * rm) { if (rows_off < 8) printf("  row %d: max err %.3e vs max ref %.3e\n", m, em, rm); ++rows_off; }
  }
  printf("%d of %d rows off; %s\n", rows_off, kM, rows_off == 0 ? "PROBE OK" : "PROBE MISMATCH");
  rc_total |= rows_off != 0;
  }

  {  // ---- CTA pair: M = 256, N = 256
    printf("---- cta_group::2: M=256, N=256, K=128 block-scaled\n");
    std::vector<float> A2((size_t)256 * kK), B2((size_t)256 * kK);
    for (auto& v : A2) v = (rand() / (float)RAND_MAX * 2 - 1) * std::ldexp(1.f, rand() % 24 - 20);
    for (auto& v : B2) v = (rand() / (float)RAND_MAX * 2 - 1) * std::ldexp(1.f, rand() % 12 - 10);
    std::vector<uint8_t> qa, qb, sa, sb;
    quantise(A2, 256, qa, sa);
    quantise(B2, 256, qb, sb);
    std::vector<double> ref2((size_t)256 * 256);
    for (int m = 0; m < 256; ++m)
      for (int n = 0; n < 256; ++n) {
        double acc = 0;
        for (int kb = 0; kb < 4; ++kb) {
          double part = 0;
          for (int k = 0; k < 32; ++k) part += (double)e4m3_decode(qa[m * kK + kb * 32 + k]) * e4m3_decode(qb[n * kK + kb * 32 + k]);
          acc += part * std::ldexp(1.0, sa[m * 4 + kb] - 127) * std::ldexp(1.0, sb[n * 4 + kb] - 127);
        }
        ref2[(size_t)m * 256 + n] = acc;
      }
    uint8_t *pa, *pb, *psa, *psb;
    float* pc;
    cudaMalloc(&pa, qa.size()); cudaMalloc(&pb, qb.size()); cudaMalloc(&psa, sa.size()); cudaMalloc(&psb, sb.size());
    cudaMalloc(&pc, ref2.size() * 4);
    cudaMemset(pc, 0, ref2.size() * 4);
    cudaMemcpy(pa, qa.data(), qa.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(pb, qb.data(), qb.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(psa, sa.data(), sa.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(psb, sb.data(), sb.size(), cudaMemcpyHostToDevice);
    const int smem2 = 32768 + 512 + 1024 + 64 + 1024;
    cudaFuncSetAttribute(probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
    probe2_kernel<<<2, 128, smem2>>>(pa, pb, psa, psb, pc);
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e2 != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e2)); return 1; }
    std::vector<float> got2(ref2.size());
    cudaMemcpy(got2.data(), pc, got2.size() * 4, cudaMemcpyDeviceToHost);
    int rows_off = 0, cols_off = 0;
    double maxref = 0, maxerr = 0;
    for (int m = 0; m < 256; ++m) {
      double rm = 0, em = 0;
      for (int n = 0; n < 256; ++n) { rm = std::fmax(rm, std::fabs(ref2[(size_t)m * 256 + n])); em = std::fmax(em, std::fabs(got2[(size_t)m * 256 + n] - ref2[(size_t)m * 256 + n])); }
      maxref = std::fmax(maxref, rm); maxerr = std::fmax(maxerr, em);
      if (em > 1e-5 * rm) { if (rows_off < 6) printf("  row %d: max err %.3e vs max ref %.3e\n", m, em, rm); ++rows_off; }
    }
    for (int n = 0; n < 256; ++n) {
      double rm = 0, em = 0;
      for (int m = 0; m < 256; ++m) { rm = std::fmax(rm, std::fabs(ref2[(size_t)m * 256 + n])); em = std::fmax(em, std::fabs(got2[(size_t)m * 256 + n] - ref2[(size_t)m * 256 + n])); }
      if (em > 1e-5 * rm) ++cols_off;
    }
    printf("max|ref| = %.6e max|err| = %.6e; %d of 256 rows, %d of 256 columns off; %s\n", maxref, maxerr, rows_off, cols_off,
           rows_off == 0 ? "PROBE OK" : "PROBE MISMATCH");
    rc_total |= rows_off != 0;
  }
  // ---- kind::f16 with an fp16 A operand and a bf16 B operand.  MEASURED on B200: "an illegal instruction was
  // encountered" — the two 16-bit formats of one kind::f16 MMA must match.  Run with --mixed-formats to reproduce (it
  // kills the context, so it is last and optional).
  if (try_mixed_formats) {
    printf("---- kind::f16, A = fp16, B = bf16, K = 64 (expected: illegal instruction)\n");
    std::vector<__half> Ah((size_t)kM * 64);
    for (auto& v : Ah) v = __float2half(rand() / (float)RAND_MAX * 2 - 1);
    cudaMemcpy(da16, Ah.data(), Ah.size() * 2, cudaMemcpyHostToDevice);
    probe_kernel<<<1, 128, smem>>>(da, db, dsa, dsb, da16, db16, 3, dc, nullptr);
    cudaError_t e3 = cudaDeviceSynchronize();
    if (e3 != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e3)); return 1; }
    std::vector<float> got3((size_t)kM * kN);
    cudaMemcpy(got3.data(), dc, got3.size() * 4, cudaMemcpyDeviceToHost);
    double maxref = 0, maxerr = 0;
    for (int m = 0; m < kM; ++m)
      for (int n = 0; n < kN; ++n) {
        double acc = 0;
        for (int k = 0; k < 64; ++k) acc += (double)__half2float(Ah[m * 64 + k]) * __bfloat162float(B16[n * 64 + k]);
        maxref = std::fmax(maxref, std::fabs(acc));
        maxerr = std::fmax(maxerr, std::fabs(acc - got3[(size_t)m * kN + n]));
      }
    printf("max|ref| = %.6e max|err| = %.6e; %s\n", maxref, maxerr, maxerr < 1e-5 * maxref ? "PROBE OK" : "PROBE MISMATCH");
    rc_total |= !(maxerr < 1e-5 * maxref);
  }
  long long t[2] = {0, 0};
  if (!try_mixed_formats) cudaMemcpy(t, dt, 16, cudaMemcpyDeviceToHost);
  printf("issue rate, M=128 N=128, 256 MMAs each: kind::f16 (K=16) %.1f cycles/MMA, kind::mxf8f6f4 (K=32) %.1f cycles/MMA\n",
         t[0] / 256.0, t[1] / 256.0);
  return rc_total;
}
