// Relative-position multi-head attention weights on the tensor pipe:
//   A[u][h][i][j] = softmax_j( q_i . k_j + p_i . pos[(j - i) + Lmax - 1] )          (per utterance u, head h)
// Inside the encoder graph the reference runs this as MatMul + the relative-shift gather + Softmax
// (/root/reference core/asr_engine.py:1047; architecture per SURVEY.md App. B.3, RelPositionMultiheadAttentionWeights).
//
// Work item = (utterance, head, 128 query rows); persistent CTAs walk the items. The scores of one 128 x 128
// (queries x keys) block are one tcgen05 accumulator: Q and K tiles are 128 x 32 fp32 (query_head_dim = 32 = one
// 128-byte swizzle row) landed by TMA straight from the packed in_proj output, 4 K-steps x 3 MMAs (3xTF32) per block.
// The 4-dim positional term is a Toeplitz rank-4 update that no GEMM shape expresses; it is added on the CUDA cores
// in the epilogue from a 255-row window of the per-head positional projection that TMA stages next to each key tile.
// The softmax needs the whole row, and 128 rows x Tk scores do not fit on chip, so a work item makes two passes over
// its key tiles: pass 1 keeps a running (max, sum) per row, pass 2 recomputes the block (the MMA work is negligible,
// K = 32) and writes exp(s - max) / sum. The kernel is bound by its epilogue (one shared-memory read, one ex2 and a
// handful of FMAs per score) and by the single HBM write of A, 4*H*Tk^2 bytes per utterance.
//
// ONEPASS (the default on the tensor-core path): softmax is invariant to the shift, so the row's DIAGONAL score s_ii - one
// 36-dim dot product per row, known before any block - takes the place of the row maximum, the pass writes the unnormalised
// exp(s_ij - s_ii) and leaves the row sum in Ls[row][head]; the three consumers of A (attn_tc.cu) scale their output rows by
// 1 / Ls in their epilogues, as a fused attention kernel would. Half the epilogue work of the two-pass form. exp(s_ij - s_ii)
// overflows only if some score exceeds the diagonal one by more than ~80; a row sum that is not finite or beyond 1e36 raises a
// device flag and the engine repeats the pass with the exact two-pass kernel (and keeps it for that recognizer).
//
// Warp roles (704 threads): warps 0..15 epilogue (TMEM lane quarter = warp % 4, 32-column group = warp / 4),
// warp 16 TMA producer, warp 17 TMEM allocator + MMA issuer, warps 18..21 hi/lo operand splitter (3xTF32 mode).
#include <algorithm>

#include "common.cuh"
#include "tc_common.cuh"

namespace b200asr {

using namespace tc;

namespace {

constexpr int kWS = 3;              // key-tile stages
constexpr int kEpiWarps = 16;
constexpr int kThreads = (kEpiWarps + 6) * 32;
constexpr int kTileBytes = TBM * TBK * 4;   // 16 KB: 128 rows x 32 fp32
constexpr int kWinRows = 256;               // positional window rows per (query tile, key tile): offsets -127 .. +127 (+1 pad)
constexpr int kWinBytes = kWinRows * 16;

struct AwParams {
  const float *proj; int ldp;      // packed in_proj output [M, H*(2*32+4)]
  const int *len, *off;            // per utterance
  const int *tile_off;             // [n_utt + 1] cumulative H * ceil(Tk / 128)
  const long long *aoff;           // [n_utt] element offset of A[u]
  int n_utt, n_tiles, H, Lmax;
  float *A;
  int *tile_counter;               // dynamic tile scheduling (tc_common.cuh); null = static
  const float *pos;                // ONEPASS: the positional projection [2*Lmax-1, H*4] (row Lmax-1 = offset 0)
  float *Ls;                       // ONEPASS: row sums [M, H]
  int *overflow;                   // ONEPASS: set when a row sum leaves the safe range
};

struct AwTile { int u, h, i0, Tk, nkt; long long row0; };

__device__ __forceinline__ AwTile aw_decode(const AwParams &p, int tile) {
  int lo = 0, hi = p.n_utt - 1;
  while (lo < hi) {   // last u with tile_off[u] <= tile
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(p.tile_off + mid) <= tile) lo = mid; else hi = mid - 1;
  }
  AwTile t;
  t.u = lo;
  t.Tk = __ldg(p.len + lo);
  t.nkt = (t.Tk + TBM - 1) / TBM;
  const int lt = tile - __ldg(p.tile_off + lo);
  t.h = lt / t.nkt;
  t.i0 = (lt - t.h * t.nkt) * TBM;
  t.row0 = __ldg(p.off + lo);
  return t;
}

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool SPLIT3, bool ONEPASS>
__global__ void __launch_bounds__(kThreads, 1)
attn_weights_tcgen05_kernel(const __grid_constant__ CUtensorMap map_proj, const __grid_constant__ CUtensorMap map_pos, AwParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *sQ = smem;                                   // 16 KB
  uint8_t *sQlo = sQ + kTileBytes;                      // 16 KB
  uint8_t *sK = sQlo + kTileBytes;                      // kWS x 16 KB
  uint8_t *sKlo = sK + kWS * kTileBytes;                // kWS x 16 KB
  uint8_t *sW = sKlo + kWS * kTileBytes;                // kWS x 4 KB positional windows [256][4]
  float2 *sML = reinterpret_cast<float2 *>(sW + kWS * kWinBytes);          // [4 column groups][128 rows] (max, sum)
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(sML + 4 * TBM);
  uint64_t *empty_bar = full_bar + kWS;
  uint64_t *ready_bar = empty_bar + kWS;
  uint64_t *q_full = ready_bar + kWS;
  uint64_t *q_ready = q_full + 1;
  uint64_t *q_empty = q_ready + 1;
  uint64_t *tmem_full_bar = q_empty + 1;     // [2]
  uint64_t *tmem_empty_bar = tmem_full_bar + 2;
  uint64_t *sched_full = tmem_empty_bar + 2, *sched_empty = sched_full + kSchedSlots;
  uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(sched_empty + kSchedSlots);
  int *sched_tile = reinterpret_cast<int *>(tmem_ptr_smem + 4);
  const TileSched sched{sched_tile, sched_full, sched_empty, p.tile_counter, p.n_tiles};

  constexpr int kPasses = ONEPASS ? 1 : 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == kEpiWarps && lane == 0) {
    sched_init(sched, 1 + kEpiWarps + (SPLIT3 ? 4 : 0));    // MMA issuer, epilogue warps, splitter warps
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_proj)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_pos)) : "memory");
    for (int s = 0; s < kWS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1 + kEpiWarps); mbar_init(&ready_bar[s], 128); }
    mbar_init(q_full, 1); mbar_init(q_ready, 128); mbar_init(q_empty, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kEpiWarps + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(256u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == kEpiWarps) {
    // ===== TMA producer
    if (lane == 0) {
      int it = 0;
      for (int ti = 0;; ++ti) {
        const int tile = sched_produce(sched, ti);
        if (tile < 0) break;
        const AwTile t = aw_decode(p, tile);
        mbar_wait(q_empty, (ti & 1) ^ 1);                 // the previous item's MMAs no longer read sQ / sQlo
        mbar_expect_tx(q_full, kTileBytes);
        tma_load_2d(&map_proj, q_full, sQ, t.h * 32, (int)t.row0 + t.i0);
        for (int pass = 0; pass < kPasses; ++pass)
          for (int kt = 0; kt < t.nkt; ++kt, ++it) {
            const int s = it % kWS;
            const uint32_t ph = (it / kWS) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            mbar_expect_tx(&full_bar[s], kTileBytes + kWinBytes);
            tma_load_2d(&map_proj, &full_bar[s], sK + s * kTileBytes, p.H * 32 + t.h * 32, (int)t.row0 + kt * TBM);
            // window row w holds pos[(j - i) + Lmax - 1] for (j - i) = (kt*128 - i0 - 127) + w; rows outside the table read 0
            tma_load_2d(&map_pos, &full_bar[s], sW + s * kWinBytes, t.h * 4, kt * TBM - t.i0 - 127 + p.Lmax - 1);
          }
      }
    }
  } else if (warp == kEpiWarps + 1) {
    // ===== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TBM, TBM);
      int it = 0;
      for (int ti = 0;; ++ti) {
        const int tile = sched_consume_thread(sched, ti);
        if (tile < 0) break;
        const AwTile t = aw_decode(p, tile);
        mbar_wait(SPLIT3 ? q_ready : q_full, ti & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t dq = make_smem_desc(smem_u32(sQ)), dql = make_smem_desc(smem_u32(sQlo));
        for (int blk = 0; blk < kPasses * t.nkt; ++blk, ++it) {
          const int s = it % kWS, acc = it & 1;
          mbar_wait(&tmem_empty_bar[acc], ((it >> 1) & 1) ^ 1);
          mbar_wait(SPLIT3 ? &ready_bar[s] : &full_bar[s], (it / kWS) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t tmem_d = tmem_base + (uint32_t)(acc * TBM);
          const uint64_t dk = make_smem_desc(smem_u32(sK + s * kTileBytes));
          const uint64_t dkl = make_smem_desc(smem_u32(sKlo + s * kTileBytes));
#pragma unroll
          for (int k = 0; k < TBK / UMMA_K; ++k) {
            const uint64_t o = (uint64_t)(k * 2);
            if constexpr (SPLIT3) {
              umma_tf32(tmem_d, dql + o, dk + o, idesc, k ? 1u : 0u);   // small terms first
              umma_tf32(tmem_d, dq + o, dkl + o, idesc, 1u);
              umma_tf32(tmem_d, dq + o, dk + o, idesc, 1u);
            } else {
              umma_tf32(tmem_d, dq + o, dk + o, idesc, k ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[s]);          // key stage free when these MMAs retire
          umma_commit(&tmem_full_bar[acc]);    // scores block complete
        }
        umma_commit(q_empty);                  // all MMAs of this item retired -> Q buffers reusable
      }
    }
  } else if (warp < kEpiWarps) {
    // ===== epilogue: thread = one query row x 32 key columns of each block
    const int q = warp & 3, cg = warp >> 2;
    int it = 0;
    for (int ti = 0;; ++ti) {
      const int tile = sched_consume_warp(sched, ti, lane);
      if (tile < 0) break;
      const AwTile t = aw_decode(p, tile);
      const int il = q * 32 + lane;            // row inside the tile
      const int i = t.i0 + il;
      const bool row_ok = i < t.Tk;
      float4 pi = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row_ok) pi = __ldg(reinterpret_cast<const float4 *>(p.proj + (t.row0 + i) * p.ldp + 2 * p.H * 32 + t.h * 4));
      const int Tk4 = (t.Tk + 3) & ~3;
      float *arow = p.A + __ldg(p.aoff + t.u) + ((long long)t.h * t.Tk + i) * Tk4;
      float m = -INFINITY, l = 0.f, inv = 0.f;
      constexpr float kLog2e = 1.4426950408889634f;
      if constexpr (ONEPASS) {
        // the shift: this row's diagonal score q_i . k_i + p_i . pos[0] (the four column-group warps of a row compute the same
        // sequence of operations on the same data, so they agree bit for bit)
        m = 0.f;
        inv = 1.0f;
        if (row_ok) {
          const float4 *qv = reinterpret_cast<const float4 *>(p.proj + (t.row0 + i) * p.ldp + t.h * 32);
          const float4 *kv = reinterpret_cast<const float4 *>(p.proj + (t.row0 + i) * p.ldp + p.H * 32 + t.h * 32);
          float d0 = 0.f, d1 = 0.f;
#pragma unroll
          for (int c = 0; c < 8; c += 2) {
            const float4 qa = __ldg(qv + c), ka = __ldg(kv + c), qb = __ldg(qv + c + 1), kb = __ldg(kv + c + 1);
            d0 = fmaf(qa.x, ka.x, d0); d0 = fmaf(qa.y, ka.y, d0); d0 = fmaf(qa.z, ka.z, d0); d0 = fmaf(qa.w, ka.w, d0);
            d1 = fmaf(qb.x, kb.x, d1); d1 = fmaf(qb.y, kb.y, d1); d1 = fmaf(qb.z, kb.z, d1); d1 = fmaf(qb.w, kb.w, d1);
          }
          const float4 w0 = __ldg(reinterpret_cast<const float4 *>(p.pos + (long long)(p.Lmax - 1) * p.H * 4 + t.h * 4));
          float ps = pi.x * w0.x;
          ps = fmaf(pi.y, w0.y, ps); ps = fmaf(pi.z, w0.z, ps); ps = fmaf(pi.w, w0.w, ps);
          m = (d0 + d1) + ps;
        }
      }
      for (int pass = 0; pass < kPasses; ++pass) {
        for (int kt = 0; kt < t.nkt; ++kt, ++it) {
          const int s = it % kWS, acc = it & 1;
          mbar_wait(&tmem_full_bar[acc], (it >> 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * TBM + cg * 32), r);
          const int j0 = kt * TBM + cg * 32;
          const int nvalid = min(32, t.Tk - j0);             // warp-uniform; <= 0: this column group is past the keys
          if (nvalid > 0) {
            // positional term: window index of (i, j) is (j - kt*128) - il + 127; scores replace the raw accumulators in r[]
            const uint32_t win = smem_u32(sW + s * kWinBytes) + (uint32_t)(cg * 32 - il + 127) * 16u;
            // (the four FMAs accumulate onto the q.k score: one instruction less per element than a separate dot product + add,
            // and full key blocks skip the validity select)
            if (nvalid == 32) {
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) {
                const float4 w = lds128(win + jj * 16);   // explicit ld.shared: a generic load here stalls on the long scoreboard
                float sc = fmaf(pi.x, w.x, __uint_as_float(r[jj]));
                sc = fmaf(pi.y, w.y, sc); sc = fmaf(pi.z, w.z, sc); sc = fmaf(pi.w, w.w, sc);
                r[jj] = __float_as_uint(sc);
              }
            } else {
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) {
                const float4 w = lds128(win + jj * 16);
                float sc = fmaf(pi.x, w.x, __uint_as_float(r[jj]));
                sc = fmaf(pi.y, w.y, sc); sc = fmaf(pi.z, w.z, sc); sc = fmaf(pi.w, w.w, sc);
                r[jj] = __float_as_uint(jj < nvalid ? sc : -INFINITY);
              }
            }
            if constexpr (ONEPASS) {
              const float mb = m * kLog2e;
              float add0 = 0.f, add1 = 0.f;
              float *dst = arow + j0;
              if (nvalid == 32) {
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                  float4 v;
                  v.x = ex2f(fmaf(__uint_as_float(r[4 * j4]), kLog2e, -mb));
                  v.y = ex2f(fmaf(__uint_as_float(r[4 * j4 + 1]), kLog2e, -mb));
                  v.z = ex2f(fmaf(__uint_as_float(r[4 * j4 + 2]), kLog2e, -mb));
                  v.w = ex2f(fmaf(__uint_as_float(r[4 * j4 + 3]), kLog2e, -mb));
                  add0 += v.x + v.z; add1 += v.y + v.w;
                  if (row_ok) *reinterpret_cast<float4 *>(dst + 4 * j4) = v;
                }
              } else {
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) {
                  const float e = ex2f(fmaf(__uint_as_float(r[jj]), kLog2e, -mb));   // -inf past the keys: 0
                  add0 += e;
                  if (row_ok && jj < nvalid) dst[jj] = e;
                }
              }
              l += add0 + add1;
            } else if (pass == 0) {
              float cm = __uint_as_float(r[0]);
#pragma unroll
              for (int jj = 1; jj < 32; ++jj) cm = fmaxf(cm, __uint_as_float(r[jj]));
              const float nm = fmaxf(m, cm);
              const float nmb = nm * kLog2e;
              float add0 = 0.f, add1 = 0.f;
#pragma unroll
              for (int jj = 0; jj < 32; jj += 2) {
                add0 += ex2f(fmaf(__uint_as_float(r[jj]), kLog2e, -nmb));
                add1 += ex2f(fmaf(__uint_as_float(r[jj + 1]), kLog2e, -nmb));
              }
              l = l * ex2f((m - nm) * kLog2e) + (add0 + add1);   // m = -inf on the first block: ex2(-inf) = 0
              m = nm;
            } else if (row_ok) {
              float *dst = arow + j0;
              const float mb = m * kLog2e;
              if (nvalid == 32) {
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                  float4 v;
                  v.x = ex2f(fmaf(__uint_as_float(r[4 * j4]), kLog2e, -mb)) * inv;
                  v.y = ex2f(fmaf(__uint_as_float(r[4 * j4 + 1]), kLog2e, -mb)) * inv;
                  v.z = ex2f(fmaf(__uint_as_float(r[4 * j4 + 2]), kLog2e, -mb)) * inv;
                  v.w = ex2f(fmaf(__uint_as_float(r[4 * j4 + 3]), kLog2e, -mb)) * inv;
                  *reinterpret_cast<float4 *>(dst + 4 * j4) = v;
                }
              } else {
#pragma unroll
                for (int jj = 0; jj < 32; ++jj)
                  if (jj < nvalid) dst[jj] = ex2f(fmaf(__uint_as_float(r[jj]), kLog2e, -mb)) * inv;
              }
            }
          }
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[acc])) : "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty_bar[s])) : "memory");   // window consumed
          }
        }
        if constexpr (ONEPASS) {
          // row sum = the four column groups' partial sums (same shift); the consumers divide by it
          sML[cg * TBM + il] = make_float2(m, l);
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
          if (cg == 0 && row_ok) {
            const float L = (sML[il].y + sML[TBM + il].y) + (sML[2 * TBM + il].y + sML[3 * TBM + il].y);
            p.Ls[(t.row0 + i) * p.H + t.h] = L;
            if (!(L < 1e36f)) atomicOr(p.overflow, 1);          // inf / NaN / close to the fp32 range
          }
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");   // sML may be rewritten by the next item
        } else if (pass == 0) {
          // merge the four column groups' running (max, sum) of each row
          sML[cg * TBM + il] = make_float2(m, l);
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
          float M = -INFINITY;
#pragma unroll
          for (int c = 0; c < 4; ++c) M = fmaxf(M, sML[c * TBM + il].x);
          float L = 0.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float2 ml = sML[c * TBM + il];
            if (ml.x > -INFINITY) L += ml.y * ex2f((ml.x - M) * kLog2e);
          }
          m = M;
          inv = 1.0f / L;
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");   // sML may be rewritten by the next item
        }
      }
    }
  } else {
    // ===== operand splitter (3xTF32): lo = x - trunc_tf32(x) for the Q tile once per item and every key tile
    if constexpr (SPLIT3) {
      const int tix = threadIdx.x - (kEpiWarps + 2) * 32;   // 0..127
      auto split = [&](const uint8_t *src, uint8_t *dst) {
        const uint32_t a4 = smem_u32(src), l4 = smem_u32(dst);
#pragma unroll 8
        for (int i = tix; i < kTileBytes / 16; i += 128) {
          const float4 v = lds128(a4 + i * 16);
          float4 lo;
          lo.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
          lo.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
          lo.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
          lo.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
          sts128(l4 + i * 16, lo);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      };
      int it = 0;
      for (int ti = 0;; ++ti) {
        const int tile = sched_consume_warp(sched, ti, lane);
        if (tile < 0) break;
        const AwTile t = aw_decode(p, tile);
        mbar_wait(q_full, ti & 1);
        split(sQ, sQlo);
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(q_ready)) : "memory");
        for (int blk = 0; blk < kPasses * t.nkt; ++blk, ++it) {
          const int s = it % kWS;
          mbar_wait(&full_bar[s], (it / kWS) & 1);
          split(sK + s * kTileBytes, sKlo + s * kTileBytes);
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&ready_bar[s])) : "memory");
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kEpiWarps + 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u));
  }
}

constexpr size_t kAwSmem = 1024 + 2 * kTileBytes + 2 * kWS * kTileBytes + kWS * kWinBytes + 4 * TBM * sizeof(float2) + (3 * kWS + 7 + 2 * kSchedSlots) * 8 + 16 + 64;

}  // namespace

bool attn_weights_tc_supported(int qd, int pd) { return qd == 32 && pd == 4 && tc_init(); }

void launch_attn_weights_tc(const float *proj, int ldp, int m_total, const float *pos, const RaggedDesc &r, const long long *aoff,
                            const int *tile_off, int n_tiles, int H, float *A, bool split3, cudaStream_t st, int *tile_counter,
                            float *Ls, int *overflow) {
  if (n_tiles <= 0 || r.total <= 0) return;
  if (!tc_init()) throw CudaError("tcgen05 attention weights: cuTensorMapEncodeTiled entry point unavailable");
  set_max_dynamic_smem(attn_weights_tcgen05_kernel<true, false>, kAwSmem);
  set_max_dynamic_smem(attn_weights_tcgen05_kernel<false, false>, kAwSmem);
  set_max_dynamic_smem(attn_weights_tcgen05_kernel<true, true>, kAwSmem);
  set_max_dynamic_smem(attn_weights_tcgen05_kernel<false, true>, kAwSmem);
  static int n_sms = 0;
  if (n_sms == 0) {
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    CUDA_CHECK(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  CUtensorMap mp, mw;
  make_map(&mp, proj, m_total, ldp, ldp, TBM);                                        // 128-byte swizzle, box 32 x 128
  make_map_plain(&mw, pos, 2 * r.max_len - 1, H * 4, H * 4, 4, kWinRows);            // per-head 4-float rows, box 4 x 256
  AwParams p{};
  p.proj = proj; p.ldp = ldp; p.len = r.len; p.off = r.off; p.tile_off = tile_off; p.aoff = aoff; p.n_utt = r.n; p.n_tiles = n_tiles;
  p.H = H; p.Lmax = r.max_len; p.A = A; p.tile_counter = tile_counter;
  p.pos = pos; p.Ls = Ls; p.overflow = overflow;
  const unsigned grid = (unsigned)std::min(n_tiles, persistent_grid_limit(n_sms));
  if (Ls) {   // single pass, unnormalised weights + row sums
    if (!overflow) throw CudaError("single-pass attention weights need the overflow flag");
    if (split3) attn_weights_tcgen05_kernel<true, true><<<grid, kThreads, kAwSmem, st>>>(mp, mw, p);
    else attn_weights_tcgen05_kernel<false, true><<<grid, kThreads, kAwSmem, st>>>(mp, mw, p);
  } else if (split3) attn_weights_tcgen05_kernel<true, false><<<grid, kThreads, kAwSmem, st>>>(mp, mw, p);
  else attn_weights_tcgen05_kernel<false, false><<<grid, kThreads, kAwSmem, st>>>(mp, mw, p);
  count_launch();
  KERNEL_CHECK();
}

}  // namespace b200asr
