// EXPERIMENTAL, off by default (B200ASR_GEMM_F16SPLIT=1 turns it on), NOT YET RUN ON HARDWARE when this was committed: the
// FP32 (token-exact) GEMM mode on 16-bit tensor-core operands.
//
// Today's FP32 mode is error-compensated 3xTF32 (gemm_tc.cu): three kind::tf32 MMAs per K step. tools/fp16_split_study.py
// replays every Linear of the oracle encoder and shows that the split  x = hi + lo,  hi = fp16(x), lo = bf16(x - hi)  with
//     A*W ~ A_lo*W_hi + A_hi*W_lo + A_hi*W_hi        (FP32 accumulation in TMEM, mixed a/b formats per MMA)
// has the same error as 3xTF32 (9e-7 vs 7e-7 relative, native fp32 1.2e-6) and keeps the decoded token ids exact, while
// kind::f16 runs at twice the TF32 rate and the weight tiles are half the bytes. Range: inside fp16's normal range
// (6.1e-5 .. 65504) the two terms carry x to 2^-20; lo is bf16 (fp32's exponent), so under that range the pair degrades
// smoothly - |x - (hi + lo)| <= max(2^-20 |x|, 6e-11), i.e. down to bf16 precision for |x| < 6e-8 - which is harmless next
// to O(1) terms of the same dot product but would show if a whole operand were tiny (a power-of-two pre-scale of W would
// remove that; not done); hi is clamped to +-65504, so values over the range also degrade to bf16 precision instead of
// producing infinities. tests/test_kernel_math.py::test_f16_split_product_is_fp32_grade pins these statements on the CPU.
//
// Same structure as the A-in-TMEM 3xTF32 kernel: persistent CTAs over 128 x BN tiles; warp 0 TMA producer, warp 1 MMA
// issuer, warps 2..9 epilogue (gemm_tc_epilogue.cuh), warps 10..13 read the landed fp32 A tile once, convert it and store
// packed hi / lo rows to tensor memory (two 16-bit K elements per 32-bit column, lane = row); the MMAs take A from there
// and the pre-split 16-bit W tiles from shared memory. One stage = 64 K elements: two 128 x 32 fp32 boxes of A (32 KB) and
// BN x 64 fp16 + BN x 64 bf16 of W (128-byte swizzled rows), 12 MMAs of K = 16.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>

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
// all-fp16 variant (16): x' = x * kAScale, w' = w * kWScale before the split, accumulator * 1 / (kAScale * kWScale) after. With
// lo = fp16(x' - fp16(x')) the pair carries x' to 2^-22 while lo is a normal fp16 (|x'| >= 0.25) and to 2^-25 absolute below.
constexpr float kAScale = 64.0f, kWScale = 1024.0f;

__host__ __device__ constexpr int f16_stages(int BN) { return BN == 64 ? 4 : 3; }

// c_format F32 at [4,6); a_format / b_format at [7,10) / [10,13): 0 = F16, 1 = BF16; K-major both; N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc16(int M, int N, int a_bf16, int b_bf16) {
  return (1u << 4) | ((uint32_t)a_bf16 << 7) | ((uint32_t)b_bf16 << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
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

__device__ __forceinline__ float clamp_f16(float x) { return fminf(fmaxf(x, -65504.0f), 65504.0f); }   // NaN stays NaN

// (x0, x1) -> packed fp16 hi pair and packed bf16 lo pair; element 0 in the low half-word
__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t &hi, uint32_t &lo) {
  const __half2 h = __floats2half2_rn(clamp_f16(x0), clamp_f16(x1));
  const float2 hf = __half22float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t *>(&h);
  lo = *reinterpret_cast<const uint32_t *>(&l);
}

// all-fp16 variant: both halves fp16
__device__ __forceinline__ void split_pair_f16(float x0, float x1, uint32_t &hi, uint32_t &lo) {
  const __half2 h = __floats2half2_rn(clamp_f16(x0), clamp_f16(x1));
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t *>(&h);
  lo = *reinterpret_cast<const uint32_t *>(&l);
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kF16Threads, 1)
gemm_f16split_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_whi,
                             const __grid_constant__ CUtensorMap map_wlo, TcParams p, int variant) {
  // `variant` (B200ASR_F16_VARIANT, hardware bring-up only): 1 / 2 / 4 drop the lo*hi / hi*lo / hi*hi term, 8 swaps the two
  // half-words of a packed A column, 16 makes lo fp16 too (all three MMAs f16 x f16)
  constexpr int NS = f16_stages(BN);
  constexpr uint32_t kTmemACol = 256;                         // A stages at columns 256 + 64 s (hi, 32 columns) / + 32 (lo)
  constexpr uint32_t kTmemCols = 512;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kABox = TBM * TBK * 4;     // one 128 x 32 fp32 box, 16 KB
  constexpr int kABytes = 2 * kABox;       // 64 K elements of A per stage
  constexpr int kWBytes = BN * FBK * 2;    // BN x 64 16-bit elements: 16 or 8 KB
  uint8_t *sA = smem;
  uint8_t *sWhi = sA + NS * kABytes;
  uint8_t *sWlo = sWhi + NS * kWBytes;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(sWlo + NS * kWBytes);
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
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_wlo)) : "memory");
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
          mbar_expect_tx(&full_bar[s], kABytes + 2 * kWBytes);
          tma_load_2d(&map_a, &full_bar[s], sA + s * kABytes, kb * FBK, m0);
          tma_load_2d(&map_a, &full_bar[s], sA + s * kABytes + kABox, kb * FBK + TBK, m0);
          tma_load_2d(&map_whi, &full_bar[s], sWhi + s * kWBytes, kb * FBK, n0);
          tma_load_2d(&map_wlo, &full_bar[s], sWlo + s * kWBytes, kb * FBK, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: per K = 16 step  A_lo(bf16) * W_hi(fp16)  +  A_hi(fp16) * W_lo(bf16)  +  A_hi * W_hi, small terms first
    if (lane == 0) {
      const int lo_bf16 = (variant & 16) ? 0 : 1;
      const uint32_t idesc_lh = make_idesc16(TBM, BN, lo_bf16, 0);
      const uint32_t idesc_hl = make_idesc16(TBM, BN, 0, lo_bf16);
      constexpr uint32_t idesc_hh = make_idesc16(TBM, BN, 0, 0);
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
            uint32_t accum = (kb | k) ? 1u : 0u;
            if (!(variant & 1)) { umma_f16_ta(tmem_d, ta_lo + c, dwh + o, idesc_lh, accum); accum = 1u; }   // small terms first
            if (!(variant & 2)) { umma_f16_ta(tmem_d, ta_hi + c, dwl + o, idesc_hl, accum); accum = 1u; }
            if (!(variant & 4)) { umma_f16_ta(tmem_d, ta_hi + c, dwh + o, idesc_hh, accum); }
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
            if (variant & 16) {
              split_pair_f16(v.x * kAScale, v.y * kAScale, hi[b * 16 + 2 * c], lo[b * 16 + 2 * c]);
              split_pair_f16(v.z * kAScale, v.w * kAScale, hi[b * 16 + 2 * c + 1], lo[b * 16 + 2 * c + 1]);
            } else {
              split_pair(v.x, v.y, hi[b * 16 + 2 * c], lo[b * 16 + 2 * c]);
              split_pair(v.z, v.w, hi[b * 16 + 2 * c + 1], lo[b * 16 + 2 * c + 1]);
            }
          }
        }
        if (variant & 8) {
#pragma unroll
          for (int j = 0; j < 32; ++j) { hi[j] = (hi[j] >> 16) | (hi[j] << 16); lo[j] = (lo[j] >> 16) | (lo[j] << 16); }
        }
        const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kTmemACol + (uint32_t)(s * 64);
        tmem_st_32x32(ta, hi);
        tmem_st_32x32(ta + 32, lo);
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

constexpr size_t f16_smem_bytes(int BN) {
  return 1024 + (size_t)f16_stages(BN) * (2 * TBM * TBK * 4 + 2 * BN * FBK * 2) + (3 * 4 + 4 + 2 * kSchedSlots) * 8 + 16 + 16 + 8 * 32 * 32 * 4 + 16;
}

// ---- pre-split 16-bit copies of a weight matrix, made on first use and kept for the life of the process
__global__ void split_w16_kernel(const float *__restrict__ w, __half *__restrict__ hi, __nv_bfloat16 *__restrict__ lo, int N, int K, int ld,
                                 int lo_f16) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * ld) return;
  const int n = (int)(i / ld), k = (int)(i % ld);
  const float x = (k < K ? w[(long long)n * K + k] : 0.f) * (lo_f16 ? kWScale : 1.0f);
  const __half h = __float2half_rn(clamp_f16(x));
  hi[i] = h;
  if (lo_f16) reinterpret_cast<__half *>(lo)[i] = __float2half_rn(x - __half2float(h));   // bring-up variant 16
  else lo[i] = __float2bfloat16_rn(x - __half2float(h));
}

struct W16 { __half *hi; __nv_bfloat16 *lo; int ld; };
std::mutex g_w16_mu;
std::map<std::tuple<const float *, int, int, int>, W16> g_w16;     // (pointer, N, K, device + 64 * lo_f16)

W16 w16_for(const float *W, int N, int K, cudaStream_t st, bool lo_f16 = false) {
  int dev = 0;
  CUDA_CHECK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_w16_mu);
  const auto key = std::make_tuple(W, N, K, dev + (lo_f16 ? 64 : 0));
  auto it = g_w16.find(key);
  if (it != g_w16.end()) return it->second;
  W16 w{nullptr, nullptr, (K + 7) & ~7};
  const size_t n = (size_t)N * w.ld;
  CUDA_CHECK(cudaMalloc(&w.hi, n * sizeof(__half)));
  CUDA_CHECK(cudaMalloc(&w.lo, n * sizeof(__nv_bfloat16)));
  split_w16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(W, w.hi, w.lo, N, K, w.ld, lo_f16 ? 1 : 0);
  count_launch();
  KERNEL_CHECK();
  g_w16.emplace(key, w);
  return w;
}

}  // namespace

// Drops the cached 16-bit copies of a weight matrix (callers that reuse a scratch pointer for different weights: B200AsrGemm).
// The stream that used them must have been synchronised.
void gemm_f16split_forget(const float *W) {
  std::lock_guard<std::mutex> lk(g_w16_mu);
  for (auto it = g_w16.begin(); it != g_w16.end();) {
    if (std::get<0>(it->first) == W) { cudaFree(it->second.hi); cudaFree(it->second.lo); it = g_w16.erase(it); }
    else ++it;
  }
}

// Returns false when the shape is outside what this variant takes (the caller then runs the 3xTF32 kernel).
bool launch_gemm_f16split(const GemmArgs &g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return true;
  if (!tc_init()) return false;
  if ((g.K & 3) || (g.lda & 3) || (reinterpret_cast<uintptr_t>(g.A) & 15)) return false;
  const bool joiner = g.act == ACT_JOINER;
  if (joiner) {
    if (!g.partials || !g.bias || (g.part_kb != 4 && g.part_kb != 8 && g.part_kb != 16) ||
        ((reinterpret_cast<uintptr_t>(g.partials) | reinterpret_cast<uintptr_t>(g.bias)) & 15))
      return false;
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
  const int variant = getenv("B200ASR_F16_VARIANT") ? atoi(getenv("B200ASR_F16_VARIANT")) : 16;
  const W16 w = w16_for(g.W, g.N, g.K, st, (variant & 16) != 0);
  CUtensorMap ma, mwh, mwl;
  make_map(&ma, g.A, g.M, g.K, g.lda, TBM);
  make_map_16(&mwh, w.hi, false, g.N, w.ld, w.ld, BN);
  make_map_16(&mwl, w.lo, true, g.N, w.ld, w.ld, BN);
  TcParams p{g.bias, g.R, g.ldr, g.C, g.ldc, g.M, g.N, g.K, g.act, g.partials, g.trace, g.tile_counter, (variant & 16) ? 1.0f / (kAScale * kWScale) : 1.0f};
  const long long n_tiles = (long long)((g.M + TBM - 1) / TBM) * ((g.N + BN - 1) / BN);
  const unsigned grid = (unsigned)std::min<long long>(n_tiles, persistent_grid_limit(n_sms));
#define B200_F16_LAUNCH(BN_, EPI_)                                                                                        \
  do {                                                                                                                    \
    set_max_dynamic_smem(gemm_f16split_tcgen05_kernel<BN_, EPI_>, f16_smem_bytes(BN_));                                   \
    launch_pdl(gemm_f16split_tcgen05_kernel<BN_, EPI_>, dim3(grid), dim3(kF16Threads), f16_smem_bytes(BN_), st, g.pdl != 0, ma, mwh, \
               mwl, p, variant);                                                                                          \
  } while (0)
#define B200_F16_BN(EPI_)                                                                          \
  do {                                                                                             \
    if (BN == 128) B200_F16_LAUNCH(128, EPI_); else B200_F16_LAUNCH(64, EPI_);                     \
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
