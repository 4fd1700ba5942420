// Tensor-core GEMM for sm_100a: tcgen05.mma (kind::tf32, FP32 accumulate in TMEM), operands staged by TMA
// (cp.async.bulk.tensor, 128-byte swizzle) through a 3-stage mbarrier ring, epilogue read back with tcgen05.ld.
//   C[M,N] = act(A[M,K] * W[N,K]^T + bias[N]) (+ R[M,N])
// This is the tensor-core mode (precision=1) of every dense Linear the reference runs inside onnxruntime
// (/root/reference core/asr_engine.py:1047 encoder graph: attention in_proj/out_proj, feed-forward, pointwise
// convolutions, encoder_proj; :1092 joiner output_linear).
//
// Why kind::tf32 and not kind::f16 on bf16 copies: activations stay fp32 in HBM (residual stream, BiasNorm),
// and at the path's shapes (K = 192..512, N = 48..1920) the GEMMs are HBM-bound, so reading the fp32 tensors
// straight through TMA avoids a conversion pass over every activation while keeping a 10-bit mantissa.
//
// Persistent CTAs (one per SM) walk 128 x BN output tiles (BN = 128 or 64). Warp roles:
//   warp 0      TMA producer (one elected lane): A tile 128x32 fp32 and W tile BNx32 fp32 per stage
//   warp 1      TMEM allocator + MMA issuer (one elected lane): 4 (x3) tcgen05.mma (K=8 each) per stage,
//               tcgen05.commit releases the stage and, after the last K block, hands the accumulator to the epilogue
//   warps 2..9  epilogue (lane quarter x column half): tcgen05.ld 32 lanes x 32 columns, transpose through a
//               per-warp smem tile so every global access is a coalesced 128-byte row, 32 residual loads in
//               flight per thread, bias / Swoosh / residual
//   warps 10..13 (3xTF32 mode) split each landed stage into hi/lo operands
// Two TMEM accumulators (2 x BN columns) overlap the epilogue of one tile with the main loop of the next.
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>
#include <mutex>

#include "common.cuh"
#include "tc_common.cuh"
#include "gemm_tc_epilogue.cuh"

namespace b200asr {

using namespace tc;

namespace {

// smem ring depth: 3xTF32 keeps hi and lo copies of both operands, so 3 stages at BN = 128 and 4 at BN = 64
__host__ __device__ constexpr int stages_for(int BN, bool split3) { return split3 ? (BN == 64 ? 4 : kStages3) : kStages1; }

// SPLIT3 = error-compensated "3xTF32": every fp32 operand x is split into hi = x with the 13 low mantissa bits
// cleared (what the tensor core reads from a raw fp32 anyway) and lo = x - hi (exact in fp32; activations are split
// in shared memory by dedicated warps, weights are pre-split once at load), and each K step issues
// A_lo*W_hi + A_hi*W_lo + A_hi*W_hi into the same FP32 TMEM accumulator. The dropped lo*lo term and the TF32
// rounding of lo are ~2^-21 relative, i.e. the product is fp32-grade while still running on the tensor pipe.
// This is the FP32 (token-exact) mode.
//
// Persistent kernel: grid = min(#tiles, #SMs); each CTA walks tiles (n fastest, so concurrently running CTAs share
// an A row block in L2). Roles: warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc), warps 2..9 epilogue,
// warps 10..13 operand splitter (SPLIT3 only). Two TMEM accumulators (2 x BN columns) let the epilogue of tile i
// overlap the main loop of tile i+1; the smem stage ring runs continuously across tiles.
//
// EPI > 0 selects the joiner epilogue (ACT_JOINER): out = acc + bias with the blank penalty applied, and each
// epilogue thread - it owns 32 consecutive columns of one output row straight out of TMEM - also emits the
// softmax/top-EPI partial record of those columns, so modified_beam_search's per-frame selection (search.cu) reads
// 3 KB of records per hypothesis instead of the 8 KB logits row (/root/reference core/asr_engine.py:1096-1106).
//
// ATMEM (3xTF32 only): the activation operand goes through tensor memory. In SS mode the kernel is bound by shared
// memory bandwidth, not by the tensor pipe: per 32-wide K block the three MMAs read A and W tiles 12 times (96 KB),
// TMA writes 48 KB and the splitter moves 32 KB, against 98 KB the SM can move in the 768 cycles the MMAs need
// (ncu: tensor pipe 56 % active). With ATMEM the splitter warps read the landed A tile once, and store hi and lo
// rows to TMEM (tcgen05.st, lane = row); the MMAs take A from there, so shared memory only carries the TMA writes,
// the W reads and one A read (112 KB), and the freed A_lo buffers pay for a fourth stage.
template <int BN, bool SPLIT3, int EPI, bool ATMEM>
__global__ void __launch_bounds__(SPLIT3 ? 448 : 320, 1)
gemm_tf32_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                         const __grid_constant__ CUtensorMap map_wlo, TcParams p) {
  static_assert(!ATMEM || SPLIT3, "A-in-TMEM is the 3xTF32 path");
  // EPI <= 0 also carries the activation as a compile-time constant (0 none, -1 SwooshL, -2 SwooshR): the epilogue of
  // the K = 192..256 layers is as long as their main loop, and a run-time switch per element showed up in it
  constexpr int NS = ATMEM ? 4 : stages_for(BN, SPLIT3);
  constexpr uint32_t kAccCols = 2 * BN;                       // two accumulators
  constexpr uint32_t kTmemACol = 256;                         // ATMEM: A stages at columns 256 + 64 s (hi) / + 32 (lo)
  constexpr uint32_t kTmemCols = ATMEM ? 512u : kAccCols;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kABytes = TBM * TBK * 4;   // 16 KB
  constexpr int kWBytes = BN * TBK * 4;    // 16 or 8 KB
  uint8_t *sA = smem;
  uint8_t *sW = smem + NS * kABytes;
  uint8_t *sAlo = sW + NS * kWBytes;                       // only carved when SPLIT3 without ATMEM
  uint8_t *sWlo = sAlo + ((SPLIT3 && !ATMEM) ? NS * kABytes : 0);
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(sWlo + (SPLIT3 ? NS * kWBytes : 0));
  uint64_t *empty_bar = full_bar + NS;
  uint64_t *ready_bar = empty_bar + NS;                    // split done (SPLIT3)
  uint64_t *tmem_full_bar = ready_bar + NS;                // [2]
  uint64_t *tmem_empty_bar = tmem_full_bar + 2;            // [2]
  uint64_t *sched_full = tmem_empty_bar + 2;               // [kSchedSlots]
  uint64_t *sched_empty = sched_full + kSchedSlots;        // [kSchedSlots]
  uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(sched_empty + kSchedSlots);
  int *sched_tile = reinterpret_cast<int *>(tmem_ptr_smem + 4);
  float *epi_stage = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(sched_tile + kSchedSlots) + 15) & ~uintptr_t(15));   // 8 warps x 32 x 32 floats, 16-byte aligned

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (EPI > 0 && p.trace && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    p.trace[0] = now;
  }
  const int nk = (p.K + TBK - 1) / TBK;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int tiles_m = (p.M + TBM - 1) / TBM;
  const int n_tiles = tiles_m * tiles_n;
  const TileSched sched{sched_tile, sched_full, sched_empty, p.tile_counter, n_tiles};

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); mbar_init(&ready_bar[s], 128); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], 8); }
    sched_init(sched, 1 + 8 + (SPLIT3 ? 4 : 0));            // MMA issuer, 8 epilogue warps, 4 splitter warps
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
  pdl_wait();      // (PDL launches) everything above overlapped the previous kernel; its results are visible from here
  pdl_trigger();   // and the next kernel in the stream may start its own prologue
  auto mark = [&](int slot) {
    if (EPI > 0 && p.trace && blockIdx.x == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      p.trace[slot] = now;
    }
  };
  if (threadIdx.x == 0) mark(2);

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
          mbar_expect_tx(&full_bar[s], kABytes + kWBytes * (SPLIT3 ? 2 : 1));
          tma_load_2d(&map_a, &full_bar[s], sA + s * kABytes, kb * TBK, m0);
          tma_load_2d(&map_w, &full_bar[s], sW + s * kWBytes, kb * TBK, n0);
          if constexpr (SPLIT3) tma_load_2d(&map_wlo, &full_bar[s], sWlo + s * kWBytes, kb * TBK, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TBM, BN);
      int it = 0;
      for (int ti = 0;; ++ti) {
        if (sched_consume_thread(sched, ti) < 0) break;
        const int acc = ti & 1;
        mbar_wait(&tmem_empty_bar[acc], ((ti >> 1) & 1) ^ 1);   // epilogue drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % NS;
          const uint32_t ph = (it / NS) & 1;
          mbar_wait(SPLIT3 ? &ready_bar[s] : &full_bar[s], ph);
          if (it == 0) mark(3);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = make_smem_desc(smem_u32(sA + s * kABytes));
          const uint64_t dw = make_smem_desc(smem_u32(sW + s * kWBytes));
          if constexpr (ATMEM) {
            const uint64_t dwl = make_smem_desc(smem_u32(sWlo + s * kWBytes));
            const uint32_t ta_hi = tmem_base + kTmemACol + (uint32_t)(s * 64), ta_lo = ta_hi + 32;
#pragma unroll
            for (int k = 0; k < TBK / UMMA_K; ++k) {
              const uint64_t o = (uint64_t)(k * 2);
              umma_tf32_ta(tmem_d, ta_lo + k * UMMA_K, dw + o, idesc, (kb | k) ? 1u : 0u);   // small terms first
              umma_tf32_ta(tmem_d, ta_hi + k * UMMA_K, dwl + o, idesc, 1u);
              umma_tf32_ta(tmem_d, ta_hi + k * UMMA_K, dw + o, idesc, 1u);
            }
          } else if constexpr (SPLIT3) {
            const uint64_t dal = make_smem_desc(smem_u32(sAlo + s * kABytes));
            const uint64_t dwl = make_smem_desc(smem_u32(sWlo + s * kWBytes));
#pragma unroll
            for (int k = 0; k < TBK / UMMA_K; ++k) {
              const uint64_t o = (uint64_t)(k * 2);   // +32 bytes inside the 128-byte swizzle row, in 16-byte units
              umma_tf32(tmem_d, dal + o, dw + o, idesc, (kb | k) ? 1u : 0u);   // small terms first
              umma_tf32(tmem_d, da + o, dwl + o, idesc, 1u);
              umma_tf32(tmem_d, da + o, dw + o, idesc, 1u);
            }
          } else {
#pragma unroll
            for (int k = 0; k < TBK / UMMA_K; ++k)
              umma_tf32(tmem_d, da + (uint64_t)(k * 2), dw + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);        // frees the stage when these MMAs retire
        }
        umma_commit(&tmem_full_bar[acc]);    // accumulator complete
      }
    }
  } else if (warp < 10) {
    // ===== epilogue warps 2..9 (gemm_tc_epilogue.cuh)
    tc_epilogue_warps<BN, EPI>(p, tmem_base, tmem_full_bar, tmem_empty_bar, epi_stage, sched, tiles_n, warp, lane);
  } else {
    // ===== operand splitter warps 10..13 (SPLIT3): hi in place, lo into the shadow stage (same swizzled layout)
    if constexpr (SPLIT3) {
      const int t = threadIdx.x - 320;   // 0..127
      int it = 0;
      for (int ti = 0;; ++ti) {
        if (sched_consume_warp(sched, ti, lane) < 0) break;
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % NS;
          const uint32_t ph = (it / NS) & 1;
          mbar_wait(&full_bar[s], ph);
          // The tensor core truncates fp32 operands to TF32 (tools/tf32_rounding_probe.py), so the raw tile IS the
          // hi operand; only lo = x - trunc(x) has to be produced. W_lo comes pre-split through TMA.
          if constexpr (ATMEM) {
            // thread = one row of the tile: its 128 bytes sit in 16-byte chunks XOR-swizzled by (row & 7), so lanes
            // reading the same logical chunk hit different banks. hi = the raw value (the MMA truncates), lo = x - trunc.
            const int row = (warp & 3) * 32 + lane;
            const uint32_t rowp = smem_u32(sA + s * kABytes + row * 128);
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 v = lds128(rowp + ((c ^ (row & 7)) << 4));
              hi[4 * c] = __float_as_uint(v.x); hi[4 * c + 1] = __float_as_uint(v.y);
              hi[4 * c + 2] = __float_as_uint(v.z); hi[4 * c + 3] = __float_as_uint(v.w);
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) lo[j] = __float_as_uint(__uint_as_float(hi[j]) - __uint_as_float(hi[j] & 0xFFFFE000u));
            const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kTmemACol + (uint32_t)(s * 64);
            tmem_st_32x32(ta, hi);
            tmem_st_32x32(ta + 32, lo);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          } else {
          const uint32_t a4 = smem_u32(sA + s * kABytes), al4 = smem_u32(sAlo + s * kABytes);
#pragma unroll 8
          for (int i = t; i < kABytes / 16; i += 128) {
            const float4 v = lds128(a4 + i * 16);
            float4 l;
            l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
            l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
            l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
            l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
            sts128(al4 + i * 16, l);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05.mma
          }
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&ready_bar[s])) : "memory");
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
  if (EPI > 0 && p.trace && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    p.trace[1] = now;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn g_encode = nullptr;
std::once_flag g_once;
bool g_ok = false;

void init_once() {
  std::call_once(g_once, [] {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn &&
        qres == cudaDriverEntryPointSuccess) {
      g_encode = reinterpret_cast<EncodeFn>(fn);
      g_ok = true;
    }
  });
}

constexpr size_t smem_bytes(int BN, bool split3, bool atmem = false) {
  if (atmem) return 1024 + (size_t)4 * (TBM * TBK * 4 + 2 * BN * TBK * 4) + (3 * 5 + 4 + 2 * kSchedSlots) * 8 + 16 + 16 + 8 * 32 * 32 * 4 + 16;
  return 1024 + (size_t)stages_for(BN, split3) * (split3 ? 2 : 1) * (TBM * TBK * 4 + BN * TBK * 4) + (3 * 5 + 4 + 2 * kSchedSlots) * 8 + 16 + 16 + 8 * 32 * 32 * 4 + 16;
}

// 2-D fp32 row-major [rows, K] with row stride ld (elements); box = 32 x box_rows, 128-byte swizzle, OOB -> 0
void make_map_uncached_impl(CUtensorMap *map, const float *ptr, int rows, int K, int ld, int box_rows);
// Encoding a tensor map costs ~1 us on the host; the per-frame search GEMMs reuse the same few (pointer, shape)
// pairs 2 x T' times per batch, so keep a small per-thread cache.
void make_map_impl(CUtensorMap *map, const float *ptr, int rows, int K, int ld, int box_rows) {
  struct Key { const float *p; int rows, K, ld, box; };
  struct Ent { Key k; CUtensorMap m; };
  static thread_local Ent cache[16];
  static thread_local int n_cached = 0, next = 0;
  for (int i = 0; i < n_cached; ++i) {
    const Key &k = cache[i].k;
    if (k.p == ptr && k.rows == rows && k.K == K && k.ld == ld && k.box == box_rows) { *map = cache[i].m; return; }
  }
  make_map_uncached_impl(map, ptr, rows, K, ld, box_rows);
  Ent &e = cache[next];
  e.k = Key{ptr, rows, K, ld, box_rows};
  e.m = *map;
  next = (next + 1) % 16;
  if (n_cached < 16) ++n_cached;
}
void make_map_uncached_impl(CUtensorMap *map, const float *ptr, int rows, int K, int ld, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)TBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw CudaError("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
}

}  // namespace

namespace tc {
void make_map(CUtensorMap *map, const float *ptr, int rows, int K, int ld, int box_rows) {
  init_once();
  if (!g_ok) throw CudaError("cuTensorMapEncodeTiled entry point unavailable");
  make_map_impl(map, ptr, rows, K, ld, box_rows);
}
void make_map_plain(CUtensorMap *map, const float *ptr, int rows, int cols, int ld, int box_cols, int box_rows) {
  init_once();
  if (!g_ok) throw CudaError("cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw CudaError("cuTensorMapEncodeTiled (plain) failed (" + std::to_string((int)r) + ")");
}
void make_map_uncached(CUtensorMap *map, const float *ptr, int rows, int K, int ld, int box_rows) {
  init_once();
  if (!g_ok) throw CudaError("cuTensorMapEncodeTiled entry point unavailable");
  make_map_uncached_impl(map, ptr, rows, K, ld, box_rows);
}
bool tc_init() { init_once(); return g_ok; }
// 16-bit row-major [rows, K] (fp16 or bf16), box = 64 x box_rows = 128-byte rows, 128-byte swizzle, OOB -> 0 (gemm_tc_f16.cu)
void make_map_16(CUtensorMap *map, const void *ptr, bool bf16, int rows, int K, int ld, int box_rows) {
  init_once();
  if (!g_ok) throw CudaError("cuTensorMapEncodeTiled entry point unavailable");
  // same small per-thread cache as make_map: a frame step of the search uses four of these (pointer, shape) pairs T' times
  struct Key { const void *p; int bf16, rows, K, ld, box; };
  struct Ent { Key k; CUtensorMap m; };
  static thread_local Ent cache[16];
  static thread_local int n_cached = 0, next = 0;
  for (int i = 0; i < n_cached; ++i) {
    const Key &k = cache[i].k;
    if (k.p == ptr && k.bf16 == (int)bf16 && k.rows == rows && k.K == K && k.ld == ld && k.box == box_rows) { *map = cache[i].m; return; }
  }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = g_encode(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(ptr),
                              dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw CudaError("cuTensorMapEncodeTiled (16-bit) failed (" + std::to_string((int)r) + ")");
  Ent &e = cache[next];
  e.k = Key{ptr, (int)bf16, rows, K, ld, box_rows};
  e.m = *map;
  next = (next + 1) % 16;
  if (n_cached < 16) ++n_cached;
}
}  // namespace tc

bool gemm_tc_available() {
  init_once();
  return g_ok;
}

static void launch_tc_impl(const GemmArgs &g, cudaStream_t st, bool split3) {
  if (g.M <= 0 || g.N <= 0) return;
  init_once();
  if (!g_ok) throw CudaError("tcgen05 GEMM: cuTensorMapEncodeTiled entry point unavailable");
  // TMA needs 16-byte aligned base and row strides; anything else goes to the CUDA-core kernel
  if ((g.lda & 3) || (g.K & 3) || (reinterpret_cast<uintptr_t>(g.A) & 15) || (reinterpret_cast<uintptr_t>(g.W) & 15)) {
    if (g.act == ACT_JOINER) throw CudaError("joiner GEMM: operands must be 16-byte aligned with K % 4 == 0");
    launch_gemm_fp32(g, st);
    return;
  }
  const bool joiner = g.act == ACT_JOINER;
  if (joiner) {
    if (!g.partials || !g.bias || (g.part_kb != 4 && g.part_kb != 8 && g.part_kb != 16) ||
        ((reinterpret_cast<uintptr_t>(g.partials) | reinterpret_cast<uintptr_t>(g.bias)) & 15))
      throw CudaError("joiner GEMM: partials buffer and bias (16-byte aligned) and part_kb in {4,8,16} are required");
  }
  static int n_sms = 0;
  if (n_sms == 0) {
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    CUDA_CHECK(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  // The per-frame joiner GEMM is latency-bound (one tile per CTA): halve the tile while the grid still fits one wave
  int BN = g.N > 64 ? 128 : 64;
  if (joiner && (long long)((g.M + TBM - 1) / TBM) * ((g.N + 63) / 64) <= n_sms) BN = 64;
  if (split3 && (!g.Wlo || (reinterpret_cast<uintptr_t>(g.Wlo) & 15)))
    throw CudaError("3xTF32 GEMM needs the pre-split low part of the weights (GemmArgs::Wlo)");
  CUtensorMap ma, mw, mwl;
  make_map_impl(&ma, g.A, g.M, g.K, g.lda, TBM);
  make_map_impl(&mw, g.W, g.N, g.K, g.K, BN);
  make_map_impl(&mwl, split3 ? g.Wlo : g.W, g.N, g.K, g.K, BN);
  TcParams p{g.bias, g.R, g.ldr, g.C, g.ldc, g.M, g.N, g.K, g.act, g.partials, g.trace, g.tile_counter, 0, 1.0f};
  const long long n_tiles = (long long)((g.M + TBM - 1) / TBM) * ((g.N + BN - 1) / BN);
  const unsigned grid = (unsigned)std::min<long long>(n_tiles, persistent_grid_limit(n_sms));   // persistent: one CTA per SM
  // one launcher per instantiation; the opt-in shared-memory attribute is set on first use
#define B200_TC_LAUNCH_(BN_, S3_, EPI_, ATM_)                                                                              \
  do {                                                                                                                    \
    set_max_dynamic_smem(gemm_tf32_tcgen05_kernel<BN_, S3_, EPI_, ATM_>, smem_bytes(BN_, S3_, ATM_));                     \
    launch_pdl(gemm_tf32_tcgen05_kernel<BN_, S3_, EPI_, ATM_>, dim3(grid), dim3((S3_) ? 448 : 320),                       \
               smem_bytes(BN_, S3_, ATM_), st, g.pdl != 0, ma, mw, mwl, p);                                               \
  } while (0)
  // 3xTF32 takes the A-in-TMEM kernel unless B200ASR_GEMM_SS is set (the shared-memory-operand variant, kept for A/B runs)
  static const bool use_atmem = getenv("B200ASR_GEMM_SS") == nullptr;
#define B200_TC_LAUNCH(BN_, S3_, EPI_)                                                                                     \
  do {                                                                                                                    \
    if ((S3_) && use_atmem) B200_TC_LAUNCH_(BN_, S3_, EPI_, (S3_));                                                       \
    else B200_TC_LAUNCH_(BN_, S3_, EPI_, false);                                                                          \
  } while (0)
#define B200_TC_JOINER(BN_, S3_)                                                                   \
  do {                                                                                             \
    if (g.part_kb == 4) B200_TC_LAUNCH(BN_, S3_, 4);                                               \
    else if (g.part_kb == 8) B200_TC_LAUNCH(BN_, S3_, 8);                                          \
    else B200_TC_LAUNCH(BN_, S3_, 16);                                                             \
  } while (0)
  if (joiner) {
    if (split3) { if (BN == 128) B200_TC_JOINER(128, true); else B200_TC_JOINER(64, true); }
    else { if (BN == 128) B200_TC_JOINER(128, false); else B200_TC_JOINER(64, false); }
  } else {
#define B200_TC_ACT(BN_, S3_)                                                                      \
  do {                                                                                             \
    if (g.act == ACT_SWOOSH_L) B200_TC_LAUNCH(BN_, S3_, -1);                                       \
    else if (g.act == ACT_SWOOSH_R) B200_TC_LAUNCH(BN_, S3_, -2);                                  \
    else B200_TC_LAUNCH(BN_, S3_, 0);                                                              \
  } while (0)
    if (g.act != ACT_NONE && g.act != ACT_SWOOSH_L && g.act != ACT_SWOOSH_R) throw CudaError("tcgen05 GEMM: unknown activation");
    if (split3) { if (BN == 128) B200_TC_ACT(128, true); else B200_TC_ACT(64, true); }
    else { if (BN == 128) B200_TC_ACT(128, false); else B200_TC_ACT(64, false); }
#undef B200_TC_ACT
  }
#undef B200_TC_JOINER
#undef B200_TC_LAUNCH
#undef B200_TC_LAUNCH_
  count_launch();
  KERNEL_CHECK();
}

namespace {
__global__ void split_lo_kernel(const float *__restrict__ w, float *__restrict__ lo, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) lo[i] = w[i] - __uint_as_float(__float_as_uint(w[i]) & 0xFFFFE000u);
}
}  // namespace
// lo = w - trunc_tf32(w): the pre-split low part the 3xTF32 kernel needs for a weight matrix
void launch_split_lo(const float *w, float *lo, long long n, cudaStream_t st) {
  if (n <= 0) return;
  split_lo_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w, lo, n);
  count_launch();
  KERNEL_CHECK();
}

void launch_gemm_tc(const GemmArgs &g, cudaStream_t st) { launch_tc_impl(g, st, false); }
// FP32-grade product on the tensor pipe: the fp16 operand split (gemm_tc_f16.cu) when the caller supplies the 16-bit weight
// copies, else (and for shapes that kernel does not take) error-compensated 3xTF32
void launch_gemm_tc3(const GemmArgs &g, cudaStream_t st) {
  if (g.W16hi && g.W16lo && launch_gemm_16(g, false, st)) return;
  launch_tc_impl(g, st, true);
}
// BF16 operands, FP32 accumulate; shapes the 16-bit kernel does not take run as single-pass TF32
void launch_gemm_bf16(const GemmArgs &g, cudaStream_t st) {
  if (g.W16hi && launch_gemm_16(g, true, st)) return;
  launch_tc_impl(g, st, false);
}

}  // namespace b200asr
