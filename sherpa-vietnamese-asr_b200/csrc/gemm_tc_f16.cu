// 16-bit tensor-core operands for the Linear layers (tcgen05 kind::f16, FP32 accumulation in TMEM), two modes:
//
// F16X3 - the FP32 (token-exact) mode on fp16 operands. Every operand is pre-scaled by a power of two and split
//     x' = x * 2^s = hi + lo,   hi = fp16(x'), lo = fp16(x' - hi)
// and each K step issues  A_lo*W_hi + A_hi*W_lo + A_hi*W_hi  into one FP32 accumulator, which the epilogue scales back by
// 2^-(sA+sW). With hi and lo both fp16 the pair carries x' to 2^-22 while lo is a normal fp16 number (|x'| >= 0.25) and to
// 2^-25 absolute below that; the scales (activations x 64, weights x 1024) put this network's operands there with a range of
// +-1023 / +-63 before fp16 would overflow (hi is clamped, so larger values degrade to fp16 precision instead of producing
// infinities). Products of fp16 numbers are exact in FP32, so the result is fp32-grade: measured on a B200 against a float64
// product 4.7e-7 .. 2.2e-6 of the output scale for K = 64 .. 512 and 8.8e-6 at K = 2432 - the same as the 3xTF32 kernel
// (gemm_tc.cu), at half the MMA issue time and half the weight-tile bytes. (An fp16-hi / bf16-lo split needs no scaling, but
// tcgen05 traps with "illegal instruction" on kind::f16 MMAs whose a and b formats differ - measured, tools/f16_probe.py.)
//
// BF16 - the reduced-precision mode the north star names: activations rounded to bf16 as they enter a GEMM, bf16 weights, one
// MMA per K step, FP32 accumulation; the residual stream between layers stays fp32.
//
// Same structure as the A-in-TMEM 3xTF32 kernel: persistent CTAs over 128 x BN tiles claimed through the tile scheduler;
// warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 epilogue (gemm_tc_epilogue.cuh), warps 10..13 read the landed fp32 A
// tile once, convert it and store packed 16-bit rows to tensor memory (two K elements per 32-bit column, lane = row; element
// 2j in the low half-word - measured); the MMAs take A from there and the pre-converted W tiles from shared memory. One stage
// = 64 K elements: two 128 x 32 fp32 boxes of A (32 KB) and BN x 64 16-bit weights per operand part (128-byte swizzled rows).
// A ragged last K block is zero-filled by TMA.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "gemm_tc_epilogue.cuh"
#include "tc_common.cuh"

namespace b200asr {

using namespace tc;

namespace tc {
// gemm_tc.cu: 16-bit row-major [rows, K] tensor map, box = 64 x box_rows (128 bytes), 128-byte swizzle, OOB -> 0
void make_map_16(CUtensorMap *map, const void *ptr, bool bf16, int rows, int K, int ld, int box_rows);
}  // namespace tc

namespace {

constexpr int FBK = 64;          // K elements per stage
constexpr int UMMA_K16 = 16;     // kind::f16: 32 bytes per instruction
constexpr int kF16Threads = 448;
constexpr float kAScale = kF16AScale, kWScale = kF16WScale;

__host__ __device__ constexpr int f16_stages(int BN, bool bf16) { return bf16 ? 4 : (BN == 64 ? 4 : 3); }

// c_format F32 at [4,6); a_format / b_format at [7,10) / [10,13): 0 = F16, 1 = BF16 (must be equal); K-major both; N >> 3 at
// [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc16(int M, int N, int bf16) {
  return (1u << 4) | ((uint32_t)bf16 << 7) | ((uint32_t)bf16 << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// A from tensor memory: 128 lanes = rows, two 16-bit K elements per 32-bit column (16 elements = 8 columns per instruction)
__device__ __forceinline__ void umma_f16_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float x0, float x1) {
  const __nv_bfloat162 b = __floats2bfloat162_rn(x0, x1);
  return *reinterpret_cast<const uint32_t *>(&b);
}

template <int BN, int EPI, bool BF16>
__global__ void __launch_bounds__(kF16Threads, 1)
gemm_f16_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_whi,
                        const __grid_constant__ CUtensorMap map_wlo, TcParams p) {
  constexpr int NS = f16_stages(BN, BF16);
  constexpr uint32_t kTmemACol = 256;                         // A stages at columns 256 + 64 s (hi, 32 columns) / + 32 (lo)
  constexpr uint32_t kTmemCols = 512;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kABox = TBM * TBK * 4;     // one 128 x 32 fp32 box, 16 KB
  constexpr int kABytes = 2 * kABox;       // 64 K elements of A per stage
  constexpr int kWBytes = BN * FBK * 2;    // BN x 64 16-bit elements: 16 or 8 KB
  constexpr int kWParts = BF16 ? 1 : 2;
  uint8_t *sA = smem;
  uint8_t *sWhi = sA + NS * kABytes;
  uint8_t *sWlo = sWhi + NS * kWBytes;     // not carved in the BF16 mode
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(sWhi + kWParts * NS * kWBytes);
  uint64_t *empty_bar = full_bar + NS;
  uint64_t *ready_bar = empty_bar + NS;
  uint64_t *tmem_full_bar = ready_bar + NS;                // [2]
  uint64_t *tmem_empty_bar = tmem_full_bar + 2;            // [2]
  uint64_t *sched_full = tmem_empty_bar + 2;               // [kSchedSlots]
  uint64_t *sched_empty = sched_full + kSchedSlots;
  uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(sched_empty + kSchedSlots);
  int *sched_tile = reinterpret_cast<int *>(tmem_ptr_smem + 4);
  float *epi_stage = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(sched_tile + kSchedSlots) + 15) & ~uintptr_t(15));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) tc_trace_mark<EPI>(p, 0);
  const int nk = (p.K + FBK - 1) / FBK;                     // a ragged last block is zero-filled by TMA (A: columns >= K; W: padded to ld, then OOB)
  const int tiles_n = (p.N + BN - 1) / BN;
  const int tiles_m = (p.M + TBM - 1) / TBM;
  const int n_tiles = tiles_m * tiles_n;
  const TileSched sched{sched_tile, sched_full, sched_empty, p.tile_counter, n_tiles};

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_whi)) : "memory");
    if constexpr (!BF16) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_wlo)) : "memory");
    for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); mbar_init(&ready_bar[s], 128); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], 8); }
    sched_init(sched, 1 + 8 + 4);                            // MMA issuer, 8 epilogue warps, 4 converter warps
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();
  pdl_trigger();
  if (threadIdx.x == 0) tc_trace_mark<EPI>(p, 2);

  if (warp == 0) {
    // ===== TMA producer
    if (lane == 0) {
      int it = 0;
      for (int ti = 0;; ++ti) {
        const int tile = sched_produce(sched, ti);
        if (tile < 0) break;
        const int m0 = (tile / tiles_n) * TBM, n0 = (tile % tiles_n) * BN;
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % NS;
          const uint32_t ph = (it / NS) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          const bool no_wlo = BF16 || (p.dbg & 1), no_a2 = (p.dbg & 2) != 0;
          mbar_expect_tx(&full_bar[s], (no_a2 ? kABox : kABytes) + (no_wlo ? 1 : 2) * kWBytes);
          tma_load_2d(&map_a, &full_bar[s], sA + s * kABytes, kb * FBK, m0);
          if (!no_a2) tma_load_2d(&map_a, &full_bar[s], sA + s * kABytes + kABox, kb * FBK + TBK, m0);
          tma_load_2d(&map_whi, &full_bar[s], sWhi + s * kWBytes, kb * FBK, n0);
          if (!no_wlo) tma_load_2d(&map_wlo, &full_bar[s], sWlo + s * kWBytes, kb * FBK, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer. F16X3: per K = 16 step  A_lo * W_hi + A_hi * W_lo + A_hi * W_hi, small terms first; BF16: one MMA
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc16(TBM, BN, BF16 ? 1 : 0);
      int it = 0;
      for (int ti = 0;; ++ti) {
        if (sched_consume_thread(sched, ti) < 0) break;
        const int acc = ti & 1;
        mbar_wait(&tmem_empty_bar[acc], ((ti >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % NS;
          const uint32_t ph = (it / NS) & 1;
          mbar_wait(&ready_bar[s], ph);
          if (it == 0) tc_trace_mark<EPI>(p, 3);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t dwh = make_smem_desc(smem_u32(sWhi + s * kWBytes));
          const uint64_t dwl = make_smem_desc(smem_u32(sWlo + s * kWBytes));
          const uint32_t ta_hi = tmem_base + kTmemACol + (uint32_t)(s * 64), ta_lo = ta_hi + 32;
#pragma unroll
          for (int k = 0; k < FBK / UMMA_K16; ++k) {
            const uint64_t o = (uint64_t)(k * 2);             // +32 bytes inside the 128-byte swizzle row, in 16-byte units
            const uint32_t c = (uint32_t)(k * (UMMA_K16 / 2));   // 16 packed elements = 8 columns
            if constexpr (BF16) {
              umma_f16_ta(tmem_d, ta_hi + c, dwh + o, idesc, (kb | k) ? 1u : 0u);
            } else {
              umma_f16_ta(tmem_d, ta_lo + c, dwh + o, idesc, (kb | k) ? 1u : 0u);
              umma_f16_ta(tmem_d, ta_hi + c, dwl + o, idesc, 1u);
              umma_f16_ta(tmem_d, ta_hi + c, dwh + o, idesc, 1u);
            }
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full_bar[acc]);
      }
    }
  } else if (warp < 10) {
    tc_epilogue_warps<BN, EPI>(p, tmem_base, tmem_full_bar, tmem_empty_bar, epi_stage, sched, tiles_n, warp, lane);
  } else {
    // ===== operand converter warps 10..13: thread = one row of the landed A tile (two 128-byte swizzled box rows)
    int it = 0;
    const int row = (warp & 3) * 32 + lane;
    for (int ti = 0;; ++ti) {
      if (sched_consume_warp(sched, ti, lane) < 0) break;
      for (int kb = 0; kb < nk; ++kb, ++it) {
        const int s = it % NS;
        const uint32_t ph = (it / NS) & 1;
        mbar_wait(&full_bar[s], ph);
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const uint32_t rowp = smem_u32(sA + s * kABytes + b * kABox + row * 128);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 v = lds128(rowp + ((c ^ (row & 7)) << 4));
            if constexpr (BF16) {
              hi[b * 16 + 2 * c] = pack_bf16(v.x, v.y);
              hi[b * 16 + 2 * c + 1] = pack_bf16(v.z, v.w);
            } else {
              split_pair_f16(v.x * kAScale, v.y * kAScale, hi[b * 16 + 2 * c], lo[b * 16 + 2 * c]);
              split_pair_f16(v.z * kAScale, v.w * kAScale, hi[b * 16 + 2 * c + 1], lo[b * 16 + 2 * c + 1]);
            }
          }
        }
        const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kTmemACol + (uint32_t)(s * 64);
        tmem_st_32x32(ta, hi);
        if constexpr (!BF16) tmem_st_32x32(ta + 32, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&ready_bar[s])) : "memory");
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
  if (threadIdx.x == 0) tc_trace_mark<EPI>(p, 1);
}

// ---------------------------------------------------------------------------------------------------------------------
// Joiner step of the search with the activation already split: X arrives as scaled fp16 hi / lo planes (the selection kernel
// writes them), so both operands go from TMA straight into the MMAs (shared-memory descriptors on both sides) - no converter
// warps and no tensor-memory staging between the load and the first MMA. The frame step is a latency chain, and the conversion
// stage was 2 of its ~12 us. Same three products per K = 16 step in the same order as the kernel above, so the records are
// bit-identical to it. Records only (EPI > 0, no logits stored).
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

constexpr int kSsThreads = 320;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
__host__ __device__ constexpr int ss_stages(int BN) { return BN == 64 ? 4 : 3; }
constexpr size_t ss_smem_bytes(int BN) {
  return 1024 + (size_t)ss_stages(BN) * (2 * TBM * FBK * 2 + 2 * BN * FBK * 2) + (2 * 4 + 4 + 2 * kSchedSlots) * 8 + 16 + 16 + 16;
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kSsThreads, 1)
joiner_f16ss_tcgen05_kernel(const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_alo,
                            const __grid_constant__ CUtensorMap map_whi, const __grid_constant__ CUtensorMap map_wlo, TcParams p) {
  static_assert(EPI > 0, "record epilogue only");
  constexpr int NS = ss_stages(BN);
  constexpr uint32_t kTmemCols = 2 * BN;                     // two accumulators: 128 or 256 columns
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kABytes = TBM * FBK * 2;   // 128 x 64 fp16: 16 KB per operand part
  constexpr int kWBytes = BN * FBK * 2;
  uint8_t *sAhi = smem;
  uint8_t *sAlo = sAhi + NS * kABytes;
  uint8_t *sWhi = sAlo + NS * kABytes;
  uint8_t *sWlo = sWhi + NS * kWBytes;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(sWlo + NS * kWBytes);
  uint64_t *empty_bar = full_bar + NS;
  uint64_t *tmem_full_bar = empty_bar + NS;                // [2]
  uint64_t *tmem_empty_bar = tmem_full_bar + 2;            // [2]
  uint64_t *sched_full = tmem_empty_bar + 2;               // [kSchedSlots]
  uint64_t *sched_empty = sched_full + kSchedSlots;
  uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(sched_empty + kSchedSlots);
  int *sched_tile = reinterpret_cast<int *>(tmem_ptr_smem + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) tc_trace_mark<EPI>(p, 0);
  const int nk = (p.K + FBK - 1) / FBK;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int tiles_m = (p.M + TBM - 1) / TBM;
  const int n_tiles = tiles_m * tiles_n;
  const TileSched sched{sched_tile, sched_full, sched_empty, p.tile_counter, n_tiles};

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_ahi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_alo)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_whi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_wlo)) : "memory");
    for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], 8); }
    sched_init(sched, 1 + 8);                                // MMA issuer, 8 epilogue warps
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();
  pdl_trigger();
  if (threadIdx.x == 0) tc_trace_mark<EPI>(p, 2);

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int ti = 0;; ++ti) {
        const int tile = sched_produce(sched, ti);
        if (tile < 0) break;
        const int m0 = (tile / tiles_n) * TBM, n0 = (tile % tiles_n) * BN;
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % NS;
          const uint32_t ph = (it / NS) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_expect_tx(&full_bar[s], 2 * kABytes + 2 * kWBytes);
          tma_load_2d(&map_ahi, &full_bar[s], sAhi + s * kABytes, kb * FBK, m0);
          tma_load_2d(&map_whi, &full_bar[s], sWhi + s * kWBytes, kb * FBK, n0);
          tma_load_2d(&map_alo, &full_bar[s], sAlo + s * kABytes, kb * FBK, m0);
          tma_load_2d(&map_wlo, &full_bar[s], sWlo + s * kWBytes, kb * FBK, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc16(TBM, BN, 0);
      int it = 0;
      for (int ti = 0;; ++ti) {
        if (sched_consume_thread(sched, ti) < 0) break;
        const int acc = ti & 1;
        mbar_wait(&tmem_empty_bar[acc], ((ti >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % NS;
          const uint32_t ph = (it / NS) & 1;
          mbar_wait(&full_bar[s], ph);
          if (it == 0) tc_trace_mark<EPI>(p, 3);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t dah = make_smem_desc(smem_u32(sAhi + s * kABytes)), dal = make_smem_desc(smem_u32(sAlo + s * kABytes));
          const uint64_t dwh = make_smem_desc(smem_u32(sWhi + s * kWBytes)), dwl = make_smem_desc(smem_u32(sWlo + s * kWBytes));
#pragma unroll
          for (int k = 0; k < FBK / UMMA_K16; ++k) {
            const uint64_t o = (uint64_t)(k * 2);             // +32 bytes inside the 128-byte swizzle row, in 16-byte units
            umma_f16_ss(tmem_d, dal + o, dwh + o, idesc, (kb | k) ? 1u : 0u);
            umma_f16_ss(tmem_d, dah + o, dwl + o, idesc, 1u);
            umma_f16_ss(tmem_d, dah + o, dwh + o, idesc, 1u);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full_bar[acc]);
      }
    }
  } else {
    tc_epilogue_warps<BN, EPI>(p, tmem_base, tmem_full_bar, tmem_empty_bar, nullptr, sched, tiles_n, warp, lane);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
  if (threadIdx.x == 0) tc_trace_mark<EPI>(p, 1);
}

constexpr size_t f16_smem_bytes(int BN, bool bf16) {
  return 1024 + (size_t)f16_stages(BN, bf16) * (2 * TBM * TBK * 4 + (bf16 ? 1 : 2) * BN * FBK * 2) + (3 * 4 + 4 + 2 * kSchedSlots) * 8 + 16 + 16 +
         8 * 32 * 32 * 4 + 16;
}

// 16-bit copies of a weight matrix [N, K] -> [N, ld] (ld = K rounded up to 8, zero padded)
__global__ void split_w16_kernel(const float *__restrict__ w, uint16_t *__restrict__ hi, uint16_t *__restrict__ lo, int N, int K, int ld,
                                 int bf16) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * ld) return;
  const int n = (int)(i / ld), k = (int)(i % ld);
  const float x = k < K ? w[(long long)n * K + k] : 0.f;
  if (bf16) {
    const __nv_bfloat16 b = __float2bfloat16_rn(x);
    hi[i] = *reinterpret_cast<const uint16_t *>(&b);
  } else {
    const float xs = x * kWScale;
    const __half h = __float2half_rn(clamp_f16(xs));
    const __half l = __float2half_rn(xs - __half2float(h));
    hi[i] = *reinterpret_cast<const uint16_t *>(&h);
    lo[i] = *reinterpret_cast<const uint16_t *>(&l);
  }
}

}  // namespace

// Makes the 16-bit operand copies of W[N, K] the kernels below read (device memory owned by the caller: cudaFree both).
// bf16: one bf16 copy (*lo = null); else the scaled fp16 hi / lo pair.
void split_weights_16(const float *W, int N, int K, bool bf16, void **hi, void **lo, int *ld, cudaStream_t st) {
  const int l = (K + 7) & ~7;
  const size_t n = (size_t)N * l;
  *hi = nullptr; *lo = nullptr; *ld = l;
  CUDA_CHECK(cudaMalloc(hi, n * 2));
  if (!bf16) CUDA_CHECK(cudaMalloc(lo, n * 2));
  split_w16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(W, reinterpret_cast<uint16_t *>(*hi), reinterpret_cast<uint16_t *>(*lo), N, K, l, bf16 ? 1 : 0);
  count_launch();
  KERNEL_CHECK();
}

// C = act(A W^T + bias) (+ R) on the 16-bit operand copies in g.W16hi / g.W16lo. Returns false when the shape is outside what
// the kernel takes (the caller then runs a TF32 kernel).
bool launch_gemm_16(const GemmArgs &g, bool bf16, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return true;
  if (!tc_init()) return false;
  if (!g.W16hi || (!bf16 && !g.W16lo)) return false;
  const bool joiner = g.act == ACT_JOINER;
  const bool pre_split = g.A16hi != nullptr;
  if (pre_split && (bf16 || !joiner || g.C || !g.A16lo || (g.a16_ld & 7) || (g.K & 7) ||
                    ((reinterpret_cast<uintptr_t>(g.A16hi) | reinterpret_cast<uintptr_t>(g.A16lo)) & 15)))
    throw CudaError("pre-split activations are taken by the joiner record GEMM of the fp16-split mode only");
  if (!pre_split && ((g.K & 3) || (g.lda & 3) || (reinterpret_cast<uintptr_t>(g.A) & 15))) return false;
  if (joiner) {
    if (!g.partials || !g.bias || (g.part_kb != 4 && g.part_kb != 8 && g.part_kb != 16) ||
        ((reinterpret_cast<uintptr_t>(g.partials) | reinterpret_cast<uintptr_t>(g.bias)) & 15)) {
      if (pre_split) throw CudaError("joiner record GEMM: unaligned records / bias");
      return false;
    }
  } else if (g.act != ACT_NONE && g.act != ACT_SWOOSH_L && g.act != ACT_SWOOSH_R) {
    return false;
  }
  static int n_sms = 0;
  if (n_sms == 0) {
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    CUDA_CHECK(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  int BN = g.N > 64 ? 128 : 64;
  if (joiner && (long long)((g.M + TBM - 1) / TBM) * ((g.N + 63) / 64) <= n_sms) BN = 64;
  CUtensorMap ma, mwh, mwl;
  if (pre_split) {
    CUtensorMap mah, mal;
    const int a_rows = g.a16_rows > g.M ? g.a16_rows : g.M;   // a constant row count keeps the tensor map cached across frame steps
    make_map_16(&mah, g.A16hi, false, a_rows, g.K, g.a16_ld, TBM);
    make_map_16(&mal, g.A16lo, false, a_rows, g.K, g.a16_ld, TBM);
    make_map_16(&mwh, g.W16hi, false, g.N, g.w16_ld, g.w16_ld, BN);
    make_map_16(&mwl, g.W16lo, false, g.N, g.w16_ld, g.w16_ld, BN);
    TcParams p{g.bias, nullptr, 0, nullptr, g.ldc, g.M, g.N, g.K, g.act, g.partials, g.trace, g.tile_counter, 0, 1.0f / (kAScale * kWScale)};
    const long long n_tiles = (long long)((g.M + TBM - 1) / TBM) * ((g.N + BN - 1) / BN);
    const unsigned grid = (unsigned)std::min<long long>(n_tiles, persistent_grid_limit(n_sms));
#define B200_SS_LAUNCH(BN_, EPI_)                                                                                          \
  do {                                                                                                                     \
    set_max_dynamic_smem(joiner_f16ss_tcgen05_kernel<BN_, EPI_>, ss_smem_bytes(BN_));                                      \
    launch_pdl(joiner_f16ss_tcgen05_kernel<BN_, EPI_>, dim3(grid), dim3(kSsThreads), ss_smem_bytes(BN_), st, g.pdl != 0, mah, mal, \
               mwh, mwl, p);                                                                                               \
  } while (0)
#define B200_SS_BN(EPI_) do { if (BN == 128) B200_SS_LAUNCH(128, EPI_); else B200_SS_LAUNCH(64, EPI_); } while (0)
    if (g.part_kb == 4) B200_SS_BN(4);
    else if (g.part_kb == 8) B200_SS_BN(8);
    else B200_SS_BN(16);
#undef B200_SS_BN
#undef B200_SS_LAUNCH
    count_launch();
    KERNEL_CHECK();
    return true;
  }
  make_map(&ma, g.A, g.M, g.K, g.lda, TBM);
  make_map_16(&mwh, g.W16hi, bf16, g.N, g.w16_ld, g.w16_ld, BN);
  make_map_16(&mwl, bf16 ? g.W16hi : g.W16lo, bf16, g.N, g.w16_ld, g.w16_ld, BN);
  TcParams p{g.bias, g.R, g.ldr, g.C, g.ldc, g.M, g.N, g.K, g.act, g.partials, g.trace, g.tile_counter, 0, bf16 ? 1.0f : 1.0f / (kAScale * kWScale)};
  static const int dbg = getenv("B200ASR_DBG_GEMM") ? atoi(getenv("B200ASR_DBG_GEMM")) : 0;
  p.dbg = dbg;
  const long long n_tiles = (long long)((g.M + TBM - 1) / TBM) * ((g.N + BN - 1) / BN);
  const unsigned grid = (unsigned)std::min<long long>(n_tiles, persistent_grid_limit(n_sms));
#define B200_F16_LAUNCH(BN_, EPI_, BF_)                                                                                   \
  do {                                                                                                                    \
    set_max_dynamic_smem(gemm_f16_tcgen05_kernel<BN_, EPI_, BF_>, f16_smem_bytes(BN_, BF_));                              \
    launch_pdl(gemm_f16_tcgen05_kernel<BN_, EPI_, BF_>, dim3(grid), dim3(kF16Threads), f16_smem_bytes(BN_, BF_), st, g.pdl != 0, ma, \
               mwh, mwl, p);                                                                                              \
  } while (0)
#define B200_F16_BN(EPI_)                                                                          \
  do {                                                                                             \
    if (bf16) { if (BN == 128) B200_F16_LAUNCH(128, EPI_, true); else B200_F16_LAUNCH(64, EPI_, true); }    \
    else { if (BN == 128) B200_F16_LAUNCH(128, EPI_, false); else B200_F16_LAUNCH(64, EPI_, false); }       \
  } while (0)
  if (joiner) {
    if (g.part_kb == 4) B200_F16_BN(4);
    else if (g.part_kb == 8) B200_F16_BN(8);
    else B200_F16_BN(16);
  } else if (g.act == ACT_SWOOSH_L) {
    B200_F16_BN(-1);
  } else if (g.act == ACT_SWOOSH_R) {
    B200_F16_BN(-2);
  } else {
    B200_F16_BN(0);
  }
#undef B200_F16_BN
#undef B200_F16_LAUNCH
  count_launch();
  KERNEL_CHECK();
  return true;
}

}  // namespace b200asr
