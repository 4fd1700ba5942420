// Attention application on the tensor pipe: out[i, c] = (sum_j A_h[i, j] * V[j, c]) (* Y[i, c]) as a GROUPED GEMM.
// Inside the encoder graph the reference runs these as MatMul nodes (self_attn1/2: softmax weights x 12-dim value
// heads; nonlin_attention: head-0 weights x (x * tanh(s)), then * y) - /root/reference core/asr_engine.py:1047,
// architecture per SURVEY.md App. B.3.
//
// One group per utterance (ragged batch): the attention weights A[u] = [H][Tk][Tk4] (Tk4 = Tk rounded up to 4 so a
// row pitch is TMA-legal) are the K-major "A operand"; V is transposed once per call into VT[u] = [C][Tk4] so it is
// a K-major "B operand" (the transpose kernel also applies x * tanh(s) for nonlin-attention and emits the low part
// of the 3xTF32 split). Per-utterance tensor maps live in global memory; a persistent CTA walks tiles
// (utterance, head | column tile, 128 query rows), the K loop runs over the keys in blocks of 32.
// Same warp roles / 3xTF32 scheme as gemm_tc.cu: TMA producer, MMA issuer, epilogue x4, A-tile splitter x4.
// The kernel streams A exactly once from HBM: 4*H*Tk^2 bytes per utterance per call is its roofline.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "tc_common.cuh"

namespace b200asr {

using namespace tc;

namespace {

// smem stages: the 12-wide value tiles are tiny, so the kernel is bound by how many bytes of A it keeps in flight
constexpr int stages_of(int BN) { return BN <= 16 ? 6 : 4; }   // 6 x 36 KB (A, A_lo, V, V_lo) / 4 x 48 KB of the 227 KB

struct AttnTcParams {
  const CUtensorMap *mapsA, *mapsV, *mapsVlo;
  const int *tile_off;     // [n_utt + 1]
  const int *len, *off;
  int n_utt, n_tiles;
  int single_head;         // 1: nonlin attention (head 0, column tiles of BN); 0: self attention (tile = head)
  int C, dv;
  const float *Y; int ldy;
  float *out; int ldo;
  int *tile_counter;       // dynamic tile scheduling (tc_common.cuh); null = static
  const float *Ls; int H;  // unnormalised weights (single-pass attn_weights): output row i of head h is divided by Ls[row, h]
};

struct TileInfo { int u, a_row0, v_row0, m_row0, col0, ncols, nk, Tk; };

template <int BN>
__device__ __forceinline__ TileInfo decode_tile(const AttnTcParams &p, int tile) {
  int lo = 0, hi = p.n_utt - 1;
  while (lo < hi) {   // last u with tile_off[u] <= tile
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(p.tile_off + mid) <= tile) lo = mid; else hi = mid - 1;
  }
  TileInfo t;
  t.u = lo;
  t.Tk = __ldg(p.len + lo);
  const int mt = (t.Tk + TBM - 1) / TBM;
  const int lt = tile - __ldg(p.tile_off + lo);
  if (!p.single_head) {
    const int h = lt / mt, mi = lt - h * mt;
    t.m_row0 = mi * TBM; t.a_row0 = h * t.Tk + t.m_row0; t.v_row0 = h * p.dv; t.col0 = h * p.dv; t.ncols = p.dv;
  } else {
    const int nt = (p.C + BN - 1) / BN;
    const int mi = lt / nt, ni = lt - mi * nt;
    t.m_row0 = mi * TBM; t.a_row0 = t.m_row0; t.v_row0 = ni * BN; t.col0 = ni * BN; t.ncols = min(BN, p.C - t.col0);
  }
  t.nk = (t.Tk + TBK - 1) / TBK;
  return t;
}

template <int NC>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[NC]);
template <>
__device__ __forceinline__ void tmem_ld_cols<16>(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int BN, bool SPLIT3>
__global__ void __launch_bounds__(SPLIT3 ? 320 : 192, 1) attn_apply_tcgen05_kernel(AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kNS = stages_of(BN);
  constexpr int kABytes = TBM * TBK * 4;               // 16 KB
  constexpr int kVBytes = BN * TBK * 4;                // 2 / 8 KB
  constexpr int kVStride = (kVBytes + 1023) & ~1023;   // keep every stage buffer 1024-byte aligned
  uint8_t *sA = smem;
  uint8_t *sV = sA + kNS * kABytes;
  uint8_t *sAlo = sV + kNS * kVStride;
  uint8_t *sVlo = sAlo + (SPLIT3 ? kNS * kABytes : 0);
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(sVlo + (SPLIT3 ? kNS * kVStride : 0));
  uint64_t *empty_bar = full_bar + kNS;
  uint64_t *ready_bar = empty_bar + kNS;
  uint64_t *tmem_full_bar = ready_bar + kNS;    // [2]
  uint64_t *tmem_empty_bar = tmem_full_bar + 2; // [2]
  uint64_t *sched_full = tmem_empty_bar + 2, *sched_empty = sched_full + kSchedSlots;
  uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(sched_empty + kSchedSlots);
  int *sched_tile = reinterpret_cast<int *>(tmem_ptr_smem + 4);
  constexpr uint32_t kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;
  const TileSched sched{sched_tile, sched_full, sched_empty, p.tile_counter, p.n_tiles};

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kNS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); mbar_init(&ready_bar[s], 128); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], 4); }
    sched_init(sched, 1 + 4 + (SPLIT3 ? 4 : 0));            // MMA issuer, 4 epilogue warps, 4 splitter warps
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

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int ti = 0;; ++ti) {
        const int tile = sched_produce(sched, ti);
        if (tile < 0) break;
        const TileInfo t = decode_tile<BN>(p, tile);
        const CUtensorMap *ma = p.mapsA + t.u, *mv = p.mapsV + t.u, *mvl = p.mapsVlo + t.u;
        for (int kb = 0; kb < t.nk; ++kb, ++it) {
          const int s = it % kNS;
          const uint32_t ph = (it / kNS) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_expect_tx(&full_bar[s], kABytes + kVBytes * (SPLIT3 ? 2 : 1));
          tma_load_2d(ma, &full_bar[s], sA + s * kABytes, kb * TBK, t.a_row0);
          tma_load_2d(mv, &full_bar[s], sV + s * kVStride, kb * TBK, t.v_row0);
          if constexpr (SPLIT3) tma_load_2d(mvl, &full_bar[s], sVlo + s * kVStride, kb * TBK, t.v_row0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TBM, BN);
      int it = 0;
      for (int ti = 0;; ++ti) {
        const int tile = sched_consume_thread(sched, ti);
        if (tile < 0) break;
        const TileInfo t = decode_tile<BN>(p, tile);
        const int acc = ti & 1;
        mbar_wait(&tmem_empty_bar[acc], ((ti >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < t.nk; ++kb, ++it) {
          const int s = it % kNS;
          const uint32_t ph = (it / kNS) & 1;
          mbar_wait(SPLIT3 ? &ready_bar[s] : &full_bar[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = make_smem_desc(smem_u32(sA + s * kABytes));
          const uint64_t dv = make_smem_desc(smem_u32(sV + s * kVStride));
          if constexpr (SPLIT3) {
            const uint64_t dal = make_smem_desc(smem_u32(sAlo + s * kABytes));
            const uint64_t dvl = make_smem_desc(smem_u32(sVlo + s * kVStride));
#pragma unroll
            for (int k = 0; k < TBK / UMMA_K; ++k) {
              const uint64_t o = (uint64_t)(k * 2);
              umma_tf32(tmem_d, dal + o, dv + o, idesc, (kb | k) ? 1u : 0u);
              umma_tf32(tmem_d, da + o, dvl + o, idesc, 1u);
              umma_tf32(tmem_d, da + o, dv + o, idesc, 1u);
            }
          } else {
#pragma unroll
            for (int k = 0; k < TBK / UMMA_K; ++k)
              umma_tf32(tmem_d, da + (uint64_t)(k * 2), dv + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full_bar[acc]);
      }
    }
  } else if (warp < 6) {
    // epilogue: thread = one query row; BN accumulator columns -> (* Y) -> out
    const int q = warp & 3;
    for (int ti = 0;; ++ti) {
      const int tile = sched_consume_warp(sched, ti, lane);
      if (tile < 0) break;
      const TileInfo t = decode_tile<BN>(p, tile);
      const int acc = ti & 1;
      mbar_wait(&tmem_full_bar[acc], (ti >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int i = t.m_row0 + q * 32 + lane;
      const long long grow = (long long)__ldg(p.off + t.u) + i;
      float scale = 1.0f;
      if (p.Ls && i < t.Tk) scale = 1.0f / __ldg(p.Ls + grow * p.H + (p.single_head ? 0 : t.col0 / p.dv));
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 16) {
        uint32_t r[16];
        tmem_ld_cols<16>(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), r);
        if (i < t.Tk && c0 < t.ncols) {
          float *orow = p.out + grow * p.ldo + t.col0 + c0;
          const float *yrow = p.Y ? p.Y + grow * p.ldy + t.col0 + c0 : nullptr;
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            if (c0 + j + 3 < t.ncols && ((p.ldo & 3) == 0) && (((t.col0 + c0) & 3) == 0) && (!yrow || (p.ldy & 3) == 0)) {
              float4 v = make_float4(__uint_as_float(r[j]) * scale, __uint_as_float(r[j + 1]) * scale, __uint_as_float(r[j + 2]) * scale,
                                     __uint_as_float(r[j + 3]) * scale);
              if (yrow) {
                const float4 y = *reinterpret_cast<const float4 *>(yrow + j);
                v.x *= y.x; v.y *= y.y; v.z *= y.z; v.w *= y.w;
              }
              *reinterpret_cast<float4 *>(orow + j) = v;
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (c0 + j + e < t.ncols) {
                  float v = __uint_as_float(r[j + e]) * scale;
                  if (yrow) v *= yrow[j + e];
                  orow[j + e] = v;
                }
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[acc])) : "memory");
    }
  } else {
    if constexpr (SPLIT3) {
      const int tix = threadIdx.x - 192;
      int it = 0;
      for (int ti = 0;; ++ti) {
        const int tile = sched_consume_warp(sched, ti, lane);
        if (tile < 0) break;
        const TileInfo t = decode_tile<BN>(p, tile);
        for (int kb = 0; kb < t.nk; ++kb, ++it) {
          const int s = it % kNS;
          const uint32_t ph = (it / kNS) & 1;
          mbar_wait(&full_bar[s], ph);
          const uint32_t a4 = smem_u32(sA + s * kABytes), al4 = smem_u32(sAlo + s * kABytes);
#pragma unroll 8
          for (int i = tix; i < kABytes / 16; i += 128) {
            const float4 v = lds128(a4 + i * 16);
            float4 l;
            l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
            l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
            l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
            l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
            sts128(al4 + i * 16, l);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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
}

// VT[u][c][j] = X[off[u]+j][c] (* tanh(S[...])) with row pitch Tk4; VTlo = VT - trunc_tf32(VT). 32x32 smem transpose.
__global__ void __launch_bounds__(256) transpose_v_kernel(const float *__restrict__ X, int ldx, const float *__restrict__ S, int lds,
                                                          int C, const int *__restrict__ len, const int *__restrict__ off,
                                                          const int *__restrict__ tile_off, int n_utt,
                                                          const long long *__restrict__ vt_off, float *__restrict__ VT,
                                                          float *__restrict__ VTlo) {
  __shared__ float tile[32][33];
  // blockIdx.x = (128-frame tile of the ragged batch) * 4 + 32-frame sub-tile: no CTA lands past a short utterance's end
  const int t128 = blockIdx.x >> 2;
  int lo = 0, hi = n_utt - 1;
  while (lo < hi) {   // last u with tile_off[u] <= t128
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(tile_off + mid) <= t128) lo = mid; else hi = mid - 1;
  }
  const int u = lo;
  const int Tk = len[u];
  const int j0 = (t128 - __ldg(tile_off + u)) * 128 + (blockIdx.x & 3) * 32, c0 = blockIdx.y * 32;
  if (j0 >= Tk) return;
  const int Tk4 = (Tk + 3) & ~3;
  const long long rbase = off[u];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int j = j0 + r, c = c0 + tx;
    float v = 0.f;
    if (j < Tk && c < C) {
      v = X[(rbase + j) * ldx + c];
      if (S) v *= tanh_sfu(S[(rbase + j) * lds + c]);
    }
    tile[r][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, j = j0 + tx;
    if (c < C && j < Tk4) {
      const float v = tile[tx][r];   // zero for j >= Tk (pitch padding)
      const long long o = vt_off[u] + (long long)c * Tk4 + j;
      VT[o] = v;
      if (VTlo) VTlo[o] = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    }
  }
}

constexpr size_t attn_smem(int BN, bool split3) {
  return 1024 + (size_t)stages_of(BN) * (TBM * TBK * 4 + ((BN * TBK * 4 + 1023) & ~1023)) * (split3 ? 2 : 1) + (3 * stages_of(BN) + 4 + 2 * kSchedSlots) * 8 + 16 + 64;
}

}  // namespace

void launch_transpose_v(const float *X, int ldx, const float *S, int lds, int C, const RaggedDesc &r, const int *tile_off, int n_tiles,
                        const long long *vt_off, float *VT, float *VTlo, cudaStream_t st) {
  if (r.total <= 0 || n_tiles <= 0) return;
  dim3 grid(n_tiles * 4, (C + 31) / 32);
  transpose_v_kernel<<<grid, 256, 0, st>>>(X, ldx, S, lds, C, r.len, r.off, tile_off, r.n, vt_off, VT, VTlo);
  count_launch();
  KERNEL_CHECK();
}

void launch_attn_apply_tc(const AttnTcLaunch &a, cudaStream_t st) {
  if (a.n_tiles <= 0) return;
  if (!tc_init()) throw CudaError("tcgen05 attention: cuTensorMapEncodeTiled entry point unavailable");
  set_max_dynamic_smem(attn_apply_tcgen05_kernel<16, false>, attn_smem(16, false));
  set_max_dynamic_smem(attn_apply_tcgen05_kernel<16, true>, attn_smem(16, true));
  set_max_dynamic_smem(attn_apply_tcgen05_kernel<64, false>, attn_smem(64, false));
  set_max_dynamic_smem(attn_apply_tcgen05_kernel<64, true>, attn_smem(64, true));
  static int n_sms = 0;
  if (n_sms == 0) {
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    CUDA_CHECK(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  AttnTcParams p{};
  p.mapsA = reinterpret_cast<const CUtensorMap *>(a.mapsA);
  p.mapsV = reinterpret_cast<const CUtensorMap *>(a.mapsV);
  p.mapsVlo = reinterpret_cast<const CUtensorMap *>(a.split3 ? a.mapsVlo : a.mapsV);
  p.tile_off = a.tile_off; p.len = a.len; p.off = a.off; p.n_utt = a.n_utt; p.n_tiles = a.n_tiles;
  p.single_head = a.single_head; p.C = a.C; p.dv = a.dv; p.Y = a.Y; p.ldy = a.ldy; p.out = a.out; p.ldo = a.ldo;
  p.tile_counter = a.tile_counter; p.Ls = a.Ls; p.H = a.H;
  const unsigned grid = (unsigned)std::min(a.n_tiles, persistent_grid_limit(n_sms));
  if (a.single_head) {
    if (a.split3) attn_apply_tcgen05_kernel<64, true><<<grid, 320, attn_smem(64, true), st>>>(p);
    else attn_apply_tcgen05_kernel<64, false><<<grid, 192, attn_smem(64, false), st>>>(p);
  } else {
    if (a.dv > 16) throw CudaError("tcgen05 attention: value_head_dim > 16 not built");
    if (a.split3) attn_apply_tcgen05_kernel<16, true><<<grid, 320, attn_smem(16, true), st>>>(p);
    else attn_apply_tcgen05_kernel<16, false><<<grid, 192, attn_smem(16, false), st>>>(p);
  }
  count_launch();
  KERNEL_CHECK();
}

// Host-side plan for one stack of one batch: per-utterance tensor maps + tile offsets, uploaded once and reused
// by every layer of the stack (the A / VT buffers do not move between layers).
void attn_tc_encode_maps(void *h_maps, int n, const float *base, const long long *elem_off, const int *len, int rows_mult,
                         int rows_fixed, int box_rows) {
  CUtensorMap *m = reinterpret_cast<CUtensorMap *>(h_maps);
  for (int u = 0; u < n; ++u) {
    const int Tk = std::max(len[u], 1);
    const int Tk4 = (Tk + 3) & ~3;
    const int rows = rows_fixed > 0 ? rows_fixed : rows_mult * Tk;
    make_map_uncached(&m[u], base + elem_off[u], rows, Tk, Tk4, box_rows);
  }
}

}  // namespace b200asr
