// Device-resident modified_beam_search / greedy search for the stateless RNN-T decoder + joiner.
// Replaces the reference's per-frame Python loop `_ort_beam_search` (/root/reference core/asr_engine.py:1023-1153:
// dec_cache lookup :1072-1088, joiner run :1090-1093, float32 log-softmax + score add :1096-1100, global top-k
// :1103-1106, expansion + ContextGraph + log-add dedup :1109-1140, finalize + length-normalised pick :1143-1153)
// and `_compute_token_entropy` (:1159-1181). Semantics that are part of the spec and kept bit-for-bit:
//   * hypothesis scores are float64; they enter the next frame's addition rounded to float32 (:1099-1100);
//   * candidates are ordered (value desc, flat index hyp*V+token asc);
//   * hypotheses with identical token sequences merge with log-add, the first inserted keeps its payload;
//   * the final pick is the first maximum of log_prob / len(ys) with len counting the 2 context slots.
//
// All utterances of a batch advance together; the loop is latency-bound (T' sequential steps), so a step is three
// lean launches:
//   decoder_joinin  decoder_proj only for hypotheses whose 2-token context changed (compacted list, fp32 CUDA-core
//                   tiles), a row copy for blank extensions, and the joiner input X = tanh(enc_t + dec) for every row
//   joiner GEMM     gemm_tc.cu, [active*beam, jd] x [jd, V] on tcgen05; its epilogue also emits, per row and per
//                   32-column part, the softmax/top-k partial record (ACT_JOINER)
//   select          one CTA per utterance merges the records: log-softmax, global top-k, expansion, hotword arcs,
//                   log-add dedup, per-token statistics (only emitted tokens touch their full logits row)
// The CUDA-core GEMM mode (precision 2) and shapes the records cannot cover use select_step_kernel on the full
// logits instead. Hypotheses live in HBM as back-pointer chains in a per-utterance arena; nothing returns to the
// host until the last frame. Utterances are processed longest-first so the active set at step t is a prefix.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "common.cuh"
#include "context_graph.h"

namespace b200asr {

namespace {

constexpr int kMaxBeam = 16;
constexpr int kSelThreads = 256;

struct HypSlot {         // one active hypothesis
  double score;          // log_prob (float64 like the reference's Python float)
  unsigned long long hash;
  int node;              // arena index of the last emitted token, -1 if none
  int len;               // emitted tokens
  int ctx;               // ContextGraph state
  int y0, y1;            // decoder context, already max(0, .)
  int dec_src;           // slot of the previous frame whose decoder output can be reused, -1 = recompute
};

struct ArenaNode {
  int parent, token, frame;
  float tok_lp;
  float stats[4];        // tsallis_norm, margin, entropy_norm, top1
};

struct SearchDev {
  // per batch
  const float *enc;            // packed [sum T', jd]
  const long long *enc_off;    // [n] row offset of sorted utterance s
  const int *lens;             // [n] T' of sorted utterance s
  const long long *arena_off;  // [n]
  HypSlot *hyps;               // [2][n][kMaxBeam] ping-pong
  int *hyp_count;              // [2][n]
  int *node_count;             // [n]
  ArenaNode *arena;
  float *E;                    // [n*beam, dd] relu(conv(embeddings)) of hypotheses whose context changed
  float *dec;                  // [2][n*beam, jd] decoder outputs, ping-pong by frame parity
  float *X;                    // [n*beam, jd] joiner input tanh(enc+dec)
  uint2 *X16hi, *X16lo;        // the same as scaled fp16 hi / lo planes ([n*beam, jd] halves each) for the pre-split joiner GEMM, or
                               // null (then X is written); decoder-table mode only
  float *logits;               // [n*beam, V]
  float *partials;             // [n*beam, ceil(V/32), 4+2*KB] records from the joiner epilogue (or null)
  int2 *chg_list;              // [2][n*beam] {row, encoder_out row of its next frame} of the hypotheses whose decoder
                               // output must be recomputed, by frame parity
  int *chg_count;              // [2]
  int2 *rowdesc;               // [2][n*beam] per joiner row {parent row to copy dec from | -1 recompute | -2 unused,
                               //                              encoder_out row of its next frame}, by frame parity
  int n, beam;
  long long *prof;             // optional per-phase cycle counters of CTA 0 (B200ASR_SEARCH_PROF=1)
  unsigned long long *trace;   // optional [steps][6] %globaltimer at entry/exit of CTA 0 of the three step kernels
};

__device__ __forceinline__ void trace_mark(unsigned long long *trace, int t, int slot) {
  if (trace && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    trace[t * 16 + slot] = now;
  }
}

// ------------------------------------------------------------------ stateless decoder (App. B.4)
// e[o] = relu(sum_{i<4,k<ctx} w[o,i,k] * emb[y_k][4*(o/4)+i]);  out = Wp e + bp.  One CTA per row.
__device__ void decoder_row(const SearchModel &m, int y0, int y1, float *s_e, float *out_row) {
  const int tid = threadIdx.x;
  for (int o = tid; o < m.dd; o += blockDim.x) {
    const int g4 = (o >> 2) << 2;
    const float *w = m.conv_w + (long long)o * 8;  // [4][2]
    const float *e0 = m.emb + (long long)y0 * m.dd + g4;
    const float *e1 = m.emb + (long long)y1 * m.dd + g4;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc = fmaf(__ldg(w + i * 2 + 0), __ldg(e0 + i), acc);
      acc = fmaf(__ldg(w + i * 2 + 1), __ldg(e1 + i), acc);
    }
    s_e[o] = fmaxf(acc, 0.f);
  }
  __syncthreads();
  // decoder_proj GEMV: each warp owns 4 output rows at a time, 128-bit loads, up to 16 independent loads in
  // flight per lane (the weights come from L2; latency, not bandwidth, is what has to be hidden)
  const int warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  const int dd4 = m.dd >> 2;
  const float4 *e4 = reinterpret_cast<const float4 *>(s_e);
  for (int j0 = warp * 4; j0 < m.jd; j0 += nw * 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float4 *w4 = reinterpret_cast<const float4 *>(m.dec_proj_w + (long long)j0 * m.dd);
#pragma unroll 4
    for (int c = lane; c < dd4; c += 32) {
      const float4 ev = e4[c];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float4 wv = __ldg(w4 + (long long)r * dd4 + c);
        acc[r] = fmaf(wv.x, ev.x, acc[r]); acc[r] = fmaf(wv.y, ev.y, acc[r]);
        acc[r] = fmaf(wv.z, ev.z, acc[r]); acc[r] = fmaf(wv.w, ev.w, acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], sft);
    }
    if (lane < 4 && j0 + lane < m.jd) {
      const float v = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
      out_row[j0 + lane] = v + __ldg(m.dec_proj_b + j0 + lane);
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) decoder_rows_kernel(SearchModel m, const long long *__restrict__ y, float *__restrict__ out) {
  extern __shared__ float s_e[];
  const int r = blockIdx.x;
  const int y0 = (int)max(0LL, y[2 * r]), y1 = (int)max(0LL, y[2 * r + 1]);
  decoder_row(m, y0, y1, s_e, out + (long long)r * m.jd);
}

__global__ void tanh_add_kernel(const float *__restrict__ enc, const float *__restrict__ dec, long long total, float *__restrict__ X) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) X[i] = tanhf(enc[i] + dec[i]);
}

// ------------------------------------------------------------------ selection
__device__ __forceinline__ unsigned ord_f32(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unord_f32(unsigned k) {
  const unsigned u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}
__device__ __forceinline__ unsigned long long mix_hash(unsigned long long h, int tok) {
  h ^= (unsigned long long)(unsigned)tok + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
  h *= 0xff51afd7ed558ccdULL;
  h ^= h >> 33;
  return h;
}
__device__ double log_add_d(double a, double b) {
  if (a < b) { const double t = a; a = b; b = t; }
  const double diff = b - a;
  return diff < -36.0 ? a : a + log1p(exp(diff));
}
// are the token chains ending at arena nodes a and b identical? (equal lengths assumed)
__device__ bool same_chain(const ArenaNode *arena, int a, int b) {
  while (a != b) {
    if (a < 0 || b < 0) return false;
    if (arena[a].token != arena[b].token) return false;
    a = arena[a].parent;
    b = arena[b].parent;
  }
  return true;
}

struct NewNode { int arena_idx; int row; };

// joiner input of 4 consecutive columns: fp32, or the scaled fp16 hi / lo planes the pre-split joiner GEMM reads
__device__ __forceinline__ void store_joiner_input(const SearchDev &d, long long row, int jd, int c4, const float4 &x) {
  if (d.X16hi) {
    uint2 hi, lo;
    split_pair_f16(x.x * kF16AScale, x.y * kF16AScale, hi.x, lo.x);
    split_pair_f16(x.z * kF16AScale, x.w * kF16AScale, hi.y, lo.y);
    d.X16hi[row * (jd >> 2) + c4] = hi;
    d.X16lo[row * (jd >> 2) + c4] = lo;
  } else {
    reinterpret_cast<float4 *>(d.X + row * jd)[c4] = x;
  }
}

#define SEL_PROF(k)                                                               \
  do {                                                                            \
    if (d.prof && blockIdx.x == 0 && threadIdx.x == 0) {                          \
      const long long _now = clock64();                                           \
      atomicAdd((unsigned long long *)&d.prof[k], (unsigned long long)(_now - _t0)); \
      _t0 = _now;                                                                 \
    }                                                                             \
  } while (0)

// ------------------------------------------------------------------ decoder update + joiner input
// One wave of CTAs (at most two per SM) walks the 16 x 32 output tiles of dec = E[list] * Wp^T + bp over the compacted
// list of rows whose context changed (the count lives on the device): both operand panels of a tile (16 rows x K
// of E, 32 rows x K of Wp, K <= 512 per pass) are pulled into shared memory with one burst of cp.async so the
// kernel pays a single L2 round trip. The rows are gathered and few (tens to hundreds), which is not a tcgen05
// shape (128-row tiles fed by TMA); the product runs on warp-level mma.sync m16n8k8 TF32 with the same
// error-compensated 3xTF32 split as the big GEMMs (hi = 13 low mantissa bits cleared, lo = x - hi, small terms
// first); a tile is 16 rows x 32 columns, one 16 x 8 mma tile per warp (the legacy tensor path is the bound here,
// so many small tiles spread over all SMs beat fewer large ones). The epilogue writes dec and X = tanh(enc_t + dec). While its first panels are in
// flight every CTA also handles its share of the blank extensions: dec = previous dec of the parent slot.
constexpr int DJ_TM = 16, DJ_TN = 32, DJ_THREADS = 128, DJ_KC = 512, DJ_LD = DJ_KC + 4;
constexpr size_t kDjSmem = (size_t)(DJ_TM + DJ_TN) * DJ_LD * sizeof(float);

__device__ __forceinline__ void cp_async16(void *dst, const void *src, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int sz = valid ? 16 : 0;   // 0 -> the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__global__ void __launch_bounds__(DJ_THREADS) decoder_joinin_kernel(SearchModel m, SearchDev d, int t, int n_rows) {
  const int tid = threadIdx.x, warp_ = tid >> 5, lane_ = tid & 31;
  const int par = t & 1;
  extern __shared__ __align__(16) float dj_smem[];
  float *As = dj_smem, *Bs = dj_smem + DJ_TM * DJ_LD;
  __shared__ int2 s_row[DJ_TM];
  const int ntn = (m.jd + DJ_TN - 1) / DJ_TN;
  const int K = m.dd;
  const int kw0 = min(DJ_KC, K);
  auto load_w_panel = [&](int n0, int kc0, int kw) {
    for (int row = 0; row < DJ_TN; ++row) {
      const int n = n0 + row;
      for (int c = tid; c < (kw >> 2); c += DJ_THREADS)
        cp_async16(Bs + row * DJ_LD + c * 4, n < m.jd ? m.dec_proj_w + (long long)n * K + kc0 + c * 4 : m.dec_proj_w, n < m.jd);
    }
  };
  // the weight panel of this CTA's first tile does not depend on the previous kernel: fetch it under its tail
  load_w_panel(((int)blockIdx.x % ntn) * DJ_TN, 0, kw0);
  pdl_wait();
  pdl_trigger();       // the joiner GEMM may become resident now; it waits for this grid before loading X
  trace_mark(d.trace, t, 0);
  const size_t plane = (size_t)d.n * d.beam * m.jd;
  float *dec_new = d.dec + par * plane;
  const float *dec_old = d.dec + (par ^ 1) * plane;
  const int2 *list = d.chg_list + (size_t)par * d.n * d.beam;
  const int2 *rdesc = d.rowdesc + (size_t)par * d.n * d.beam;
  // count and this CTA's first list entries in one round trip (entries past the count are stale but in range)
  const int r00 = ((int)blockIdx.x / ntn) * DJ_TM;
  int2 first_ent = make_int2(-1, 0);
  if (tid < DJ_TM && r00 + tid < n_rows) first_ent = list[r00 + tid];
  const int n_chg = d.chg_count[par];
  const int n_tiles = ((n_chg + DJ_TM - 1) / DJ_TM) * ntn;
  const int g = lane_ >> 2, tq = lane_ & 3;             // mma fragment coordinates
  const int rb = 0, cb = warp_ * 8;
  bool rows_done = false;
  // blank extensions: flattened over (row, float4) units, four units in flight per thread. CTAs without a tile take
  // them all when there are enough of them; otherwise everybody shares them while the first operand panels land.
  const int free_ctas = (int)gridDim.x - min(n_tiles, (int)gridDim.x);
  const bool free_only = free_ctas >= 32;
  auto copy_rows = [&]() {
    rows_done = true;
    int part = blockIdx.x, parts = gridDim.x;
    if (free_only) {
      if ((int)blockIdx.x < n_tiles) return;
      part = (int)blockIdx.x - n_tiles; parts = free_ctas;
    }
    const int jd4 = m.jd >> 2;
    const long long total = (long long)n_rows * jd4, stride = (long long)parts * DJ_THREADS;
    for (long long u0 = (long long)part * DJ_THREADS + tid; u0 < total; u0 += 4 * stride) {
      int2 ds[4];
      int rr[4], cc[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long u = u0 + i * stride;
        ds[i] = make_int2(-2, 0);
        rr[i] = 0; cc[i] = 0;
        if (u < total) { rr[i] = (int)(u / jd4); cc[i] = (int)(u - (long long)rr[i] * jd4); ds[i] = rdesc[rr[i]]; }
      }
      float4 dv[4], ev[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (ds[i].x >= 0) {
          dv[i] = reinterpret_cast<const float4 *>(dec_old + (long long)ds[i].x * m.jd)[cc[i]];
          ev[i] = __ldg(reinterpret_cast<const float4 *>(d.enc + (long long)ds[i].y * m.jd) + cc[i]);
        }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (ds[i].x >= 0) {
          reinterpret_cast<float4 *>(dec_new + (long long)rr[i] * m.jd)[cc[i]] = dv[i];
          reinterpret_cast<float4 *>(d.X + (long long)rr[i] * m.jd)[cc[i]] =
              make_float4(tanhf(ev[i].x + dv[i].x), tanhf(ev[i].y + dv[i].y), tanhf(ev[i].z + dv[i].z), tanhf(ev[i].w + dv[i].w));
        }
    }
  };
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int r0 = (tile / ntn) * DJ_TM, n0 = (tile % ntn) * DJ_TN;
    const bool first = tile == (int)blockIdx.x;
    if (!first) __syncthreads();                          // previous tile's readers are done with s_row / panels
    if (tid < DJ_TM) {
      int2 e = first ? first_ent : ((r0 + tid < n_chg) ? list[r0 + tid] : make_int2(-1, 0));
      if (r0 + tid >= n_chg) e.x = -1;
      s_row[tid] = e;
    }
    __syncthreads();
    // 3 independent accumulator chains, one per product term (mma.sync latency, not throughput, bounds this)
    float acc[3][4];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[a][i] = 0.f;
    float2 bs = make_float2(0.f, 0.f), ev[2];
    int rowi[2];
    for (int kc0 = 0; kc0 < K; kc0 += DJ_KC) {
      const int kw = min(DJ_KC, K - kc0);
      const int kw4 = kw >> 2;   // float4 per row in this pass (K % 4 == 0)
      if (kc0) __syncthreads();
      for (int row = 0; row < DJ_TM; ++row) {
        const int r = s_row[row].x;
        for (int c = tid; c < kw4; c += DJ_THREADS)
          cp_async16(As + row * DJ_LD + c * 4, r >= 0 ? d.E + (long long)r * K + kc0 + c * 4 : d.E, r >= 0);
      }
      if (!(first && kc0 == 0)) load_w_panel(n0, kc0, kw);
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (first) trace_mark(d.trace, t, 10);
      if (kc0 == 0) {
        // epilogue operands, requested before the product so they are in registers when it ends
        const int n = n0 + cb + 2 * tq;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int2 e = s_row[rb + g + 8 * h];
          rowi[h] = e.x;
          ev[h] = (e.x >= 0 && n < m.jd) ? __ldg(reinterpret_cast<const float2 *>(d.enc + (long long)e.y * m.jd + n)) : make_float2(0.f, 0.f);
        }
        if (n < m.jd) bs = __ldg(reinterpret_cast<const float2 *>(m.dec_proj_b + n));
      }
      if (!rows_done) copy_rows();
      if (first) trace_mark(d.trace, t, 11);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      if (first) trace_mark(d.trace, t, 12);
      // fragment reads are bank-conflict free: row stride 516 words -> bank = 4 * row + k (mod 32)
      const float *ap0 = As + (rb + g) * DJ_LD + tq, *ap1 = ap0 + 8 * DJ_LD;
      const float *bp0 = Bs + (cb + g) * DJ_LD + tq;
#pragma unroll 4
      for (int k0 = 0; k0 + 8 <= kw; k0 += 8) {
        const float av[4] = {ap0[k0], ap1[k0], ap0[k0 + 4], ap1[k0 + 4]};
        const float bv[2] = {bp0[k0], bp0[k0 + 4]};
        unsigned ahi[4], alo[4], bhi[2], blo[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          ahi[i] = __float_as_uint(av[i]) & 0xFFFFE000u;
          alo[i] = __float_as_uint(av[i] - __uint_as_float(ahi[i]));
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          bhi[i] = __float_as_uint(bv[i]) & 0xFFFFE000u;
          blo[i] = __float_as_uint(bv[i] - __uint_as_float(bhi[i]));
        }
        mma_tf32_16x8x8(acc[0], alo, bhi);
        mma_tf32_16x8x8(acc[1], ahi, blo);
        mma_tf32_16x8x8(acc[2], ahi, bhi);
      }
    }
    if (first) trace_mark(d.trace, t, 13);
    // accumulator fragment: c0,c1 -> row g, columns 2 tq, 2 tq + 1; c2,c3 -> row g + 8
    const int n = n0 + cb + 2 * tq;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = rowi[h];
      if (r < 0 || n >= m.jd) continue;          // jd % 4 == 0, n even -> n + 1 < jd
      const float2 dv = make_float2((acc[0][2 * h] + acc[1][2 * h]) + acc[2][2 * h] + bs.x,
                                    (acc[0][2 * h + 1] + acc[1][2 * h + 1]) + acc[2][2 * h + 1] + bs.y);
      *reinterpret_cast<float2 *>(dec_new + (long long)r * m.jd + n) = dv;
      *reinterpret_cast<float2 *>(d.X + (long long)r * m.jd + n) = make_float2(tanhf(ev[h].x + dv.x), tanhf(ev[h].y + dv.y));
    }
    if (first) trace_mark(d.trace, t, 14);
  }
  if (!rows_done) copy_rows();
  asm volatile("cp.async.wait_all;" ::: "memory");   // a CTA without tiles still has its speculative weight panel in flight
  trace_mark(d.trace, t, 1);
}

// ------------------------------------------------------------------ selection: shared tail
// Expansion of the k winners (serial, exactly in top-k order :1109-1140), hotword arcs, log-add dedup, the per-token
// statistics of emitted tokens (_compute_token_entropy :1159-1181) from their logits rows (`rows`, shared or global
// memory; null = already in sh.rstats), write-back of the new beam, and for hypotheses whose context changed the decoder pre-activation
// E = relu(grouped conv over the two context embeddings) plus their entry in the next step's recompute list.
struct SelShared {
  HypSlot old_[kMaxBeam], new_[kMaxBeam];
  float mx[kMaxBeam], lse[kMaxBeam], sum[kMaxBeam], prev[kMaxBeam];
  unsigned long long win[kMaxBeam];
  NewNode nodes[kMaxBeam];
  int chg_pos[kMaxBeam];
  int slot_of[kMaxBeam];        // winner r -> slot in the new beam, -1 if it merged into an earlier one (decoder-table mode)
  // the k winners expanded in parallel (lane r of warp 0): the hypothesis it would become, its arena node, parent slot
  HypSlot cand[kMaxBeam];
  ArenaNode cand_node[kMaxBeam];
  int cand_hi[kMaxBeam];        // parent slot, -1 = no winner r
  int rep[kMaxBeam];            // slot q of the new beam -> the winner that opened it
  int node_of[kMaxBeam];        // winner r -> arena index of its new node (non-blank winners that open a slot)
  float rstats[kMaxBeam][4];   // per-row token statistics when they come from the partial records
  int n_new, n_nodes;
  int node_count0;             // arena fill of this utterance, fetched at kernel start (off the serial section)
};

__device__ void finish_step(const SearchModel &m, const SearchDev &d, const ContextGraphView &g, int has_graph, int s, int t, int cur,
                            int greedy, int k, SelShared &sh, const float *rows, long long row_stride, long long &_t0) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int V = m.V;
  ArenaNode *arena = d.arena + d.arena_off[s];
  const bool more = t + 1 < d.lens[s];      // another frame follows: the new beam needs joiner inputs
  // Decoder-table mode: the joiner input of every new hypothesis is tanh(enc[t + 1] + table[y0, y1]), written here, so a frame
  // step has no decoder kernel. Warps 1.. fetch the table and encoder rows of all k winners while thread 0 runs the serial
  // expansion below (the context of winner r is known from its key; whether it survives the dedup is not, yet).
  const bool table = m.dec_table != nullptr;
  constexpr int kPf = 3;
  constexpr int kPfThreads = kSelThreads - 32;
  const int jd4 = m.jd >> 2;
  const long long enc_next_row = d.enc_off[s] + t + 1;
  float4 pf_t[kPf], pf_e[kPf];
  if (table && more && warp > 0) {
#pragma unroll
    for (int i = 0; i < kPf; ++i) {
      const int u = (tid - 32) + i * kPfThreads;
      const int r = u / jd4, c = u - r * jd4;
      pf_t[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      pf_e[i] = pf_t[i];
      const unsigned long long key = r < k ? sh.win[r] : 0ULL;
      if (key != 0ULL) {
        const int idx = (int)(~(unsigned)(key & 0xffffffffULL));
        const int hi = idx / V, tok = idx - hi * V;
        const bool blank = tok == m.blank_id;
        const int y0 = blank ? sh.old_[hi].y0 : sh.old_[hi].y1, y1 = blank ? sh.old_[hi].y1 : tok;
        pf_t[i] = __ldg(reinterpret_cast<const float4 *>(m.dec_table + ((long long)y0 * V + y1) * m.jd) + c);
        pf_e[i] = __ldg(reinterpret_cast<const float4 *>(d.enc + enc_next_row * m.jd) + c);
      }
    }
  }
  // Expansion in two halves. Everything that depends on one winner only - key decode, the integer division, hash, hotword
  // arc, float64 score, the arena node's fields - is computed by lane r of warp 0 for all k winners at once; what depends on the
  // ORDER of the winners (:1109-1140: dedup against the hypotheses already inserted, log-add merge, slot and arena numbering)
  // stays one thread's serial loop, which is then a few compares and struct copies per winner.
  const long long _tA = (d.prof && blockIdx.x == 0 && tid == 0) ? clock64() : 0;
  if (warp == 0) {
    if (lane < kMaxBeam) {
      int hi = -1;
      const unsigned long long key = lane < k ? sh.win[lane] : 0ULL;
      if (key != 0ULL) {
        const int idx = (int)(~(unsigned)(key & 0xffffffffULL));
        const float lp32 = unord_f32((unsigned)(key >> 32));
        hi = idx / V;
        const int tok = idx - hi * V;
        const HypSlot &p = sh.old_[hi];
        HypSlot c;
        c.score = (double)lp32;
        if (tok == m.blank_id) {
          c.hash = p.hash; c.node = p.node; c.len = p.len; c.ctx = p.ctx; c.y0 = p.y0; c.y1 = p.y1; c.dec_src = hi;
        } else {
          c.hash = mix_hash(p.hash, tok);
          c.len = p.len + 1; c.ctx = p.ctx; c.y0 = p.y1; c.y1 = tok; c.dec_src = -1; c.node = -1;
          if (has_graph && !greedy && tok != m.unk_id) {
            int nxt;
            c.score += cg_forward_one_step(g, p.ctx, tok, &nxt);
            c.ctx = nxt;
          }
        }
        if (greedy) c.score = 0.0;
        ArenaNode nd;
        nd.parent = p.node; nd.token = tok; nd.frame = t;
        nd.tok_lp = (float)((double)lp32 - (greedy ? 0.0 : p.score));   // (:1121)
        if (rows) { nd.stats[0] = nd.stats[1] = nd.stats[2] = nd.stats[3] = 0.f; }
        else { nd.stats[0] = sh.rstats[hi][0]; nd.stats[1] = sh.rstats[hi][1]; nd.stats[2] = sh.rstats[hi][2]; nd.stats[3] = sh.rstats[hi][3]; }
        sh.cand[lane] = c;
        sh.cand_node[lane] = nd;
      }
      sh.cand_hi[lane] = hi;
      sh.slot_of[lane] = -1;
    }
    __syncwarp();
  }
  if (tid == 0) {
    const long long _tB = (d.prof && blockIdx.x == 0) ? clock64() : 0;
    int n_new = 0, n_nodes = 0, n_dup = 0;
    int node_count = sh.node_count0;
    for (int r = 0; r < k; ++r) {
      const int hi = sh.cand_hi[r];
      if (hi < 0) break;
      const int len_r = sh.cand[r].len;
      const unsigned long long hash_r = sh.cand[r].hash;
      const int tok_r = sh.cand_node[r].token, pn_r = sh.cand_node[r].parent;
      const bool blank_r = (tok_r == m.blank_id);
      // dedup against the hypotheses already inserted (same token sequence). A slot's sequence is its opener's: the parent's
      // chain, plus the new token unless the opener was a blank extension - compared without the arena node that is written
      // only after this loop (same comparisons as walking it: core/asr_engine.py:1126-1134 compares the token tuples).
      int dup = -1;
      for (int q = 0; q < n_new && dup < 0; ++q) {
        const int rq = sh.rep[q];
        if (sh.cand[rq].len != len_r || sh.cand[rq].hash != hash_r) continue;
        const int tok_q = sh.cand_node[rq].token, pn_q = sh.cand_node[rq].parent;
        const bool blank_q = (tok_q == m.blank_id);
        bool same;
        if (blank_q == blank_r) same = (blank_r || tok_q == tok_r) && same_chain(arena, pn_q, pn_r);
        else if (blank_r) same = pn_r >= 0 && arena[pn_r].token == tok_q && same_chain(arena, pn_q, arena[pn_r].parent);
        else same = pn_q >= 0 && arena[pn_q].token == tok_r && same_chain(arena, arena[pn_q].parent, pn_r);
        if (same) dup = q;
      }
      if (dup >= 0) {
        const int rq = sh.rep[dup];
        sh.cand[rq].score = log_add_d(sh.cand[rq].score, sh.cand[r].score);
        ++n_dup;
        continue;
      }
      sh.rep[n_new] = r;
      sh.slot_of[r] = n_new;
      if (!blank_r) {
        const int ai = node_count++;
        sh.node_of[r] = ai;
        sh.nodes[n_nodes].arena_idx = ai;
        sh.nodes[n_nodes].row = hi;
        ++n_nodes;
      }
      ++n_new;
    }
    sh.n_new = n_new;
    sh.n_nodes = n_nodes;
    if (d.prof && blockIdx.x == 0) {   // CTA 0: merges per step, cycles of the parallel and of the serial half
      const long long now = clock64();
      atomicAdd((unsigned long long *)&d.prof[5], (unsigned long long)n_dup);
      atomicAdd((unsigned long long *)&d.prof[6], (unsigned long long)(_tB - _tA));
      atomicAdd((unsigned long long *)&d.prof[7], (unsigned long long)(now - _tB));
    }
    d.node_count[s] = node_count;
    d.hyp_count[(cur ^ 1) * d.n + s] = n_new;
    // rows of the new beam whose decoder output has to be recomputed before the next joiner step, and the per-row
    // descriptors of the blank extensions (parent row to copy from); both carry the encoder_out row of frame t + 1
    // (on-demand decoder only: with the table every row is written below)
    if (!table) {
      int n_chg = 0;
      for (int q = 0; q < n_new; ++q) {
        sh.chg_pos[q] = -1;
        if (more && sh.cand[sh.rep[q]].dec_src < 0) sh.chg_pos[q] = n_chg++;   // (new_ is filled after this block)
      }
      const int enc_next = (int)d.enc_off[s] + t + 1;
      int2 *rdesc = d.rowdesc + (size_t)((t + 1) & 1) * d.n * d.beam + (size_t)s * d.beam;
      for (int q = 0; q < d.beam; ++q) {
        int src = -2;
        if (more && q < n_new) src = sh.cand[sh.rep[q]].dec_src < 0 ? -1 : s * d.beam + sh.cand[sh.rep[q]].dec_src;
        rdesc[q] = make_int2(src, enc_next);
      }
      if (n_chg > 0) {
        const int base = atomicAdd(&d.chg_count[(t + 1) & 1], n_chg);
        int2 *list = d.chg_list + (size_t)((t + 1) & 1) * d.n * d.beam;
        for (int q = 0; q < n_new; ++q)
          if (sh.chg_pos[q] >= 0) list[base + sh.chg_pos[q]] = make_int2(s * d.beam + q, enc_next);
      }
    }
  }
  if (warp == 0) {
    // the structs move in parallel: winner r that opened slot q becomes new_[q]; its arena node goes out
    __syncwarp();
    if (lane < k && sh.cand_hi[lane] >= 0 && sh.slot_of[lane] >= 0) {
      HypSlot c = sh.cand[lane];
      if (sh.cand_node[lane].token != m.blank_id) {
        c.node = sh.node_of[lane];
        arena[c.node] = sh.cand_node[lane];
      }
      sh.new_[sh.slot_of[lane]] = c;
    }
  }
  __syncthreads();
  SEL_PROF(3);
  if (table && more) {
    auto tanh4 = [](const float4 &a, const float4 &b) {
      return make_float4(tanhf(a.x + b.x), tanhf(a.y + b.y), tanhf(a.z + b.z), tanhf(a.w + b.w));
    };
    const long long xrow0 = (long long)s * d.beam;
    if (warp > 0) {
#pragma unroll
      for (int i = 0; i < kPf; ++i) {
        const int u = (tid - 32) + i * kPfThreads;
        const int r = u / jd4, c = u - r * jd4;
        const int q = r < k ? sh.slot_of[r] : -1;
        if (q >= 0) store_joiner_input(d, xrow0 + q, m.jd, c, tanh4(pf_e[i], pf_t[i]));
      }
    }
    // winners past the prefetch window (beam > 4 or a wider joiner)
    for (int u = kPf * kPfThreads + tid; u < k * jd4; u += kSelThreads) {
      const int r = u / jd4, c = u - r * jd4;
      const int q = sh.slot_of[r];
      if (q < 0) continue;
      const HypSlot &hs = sh.new_[q];
      const float4 tv = __ldg(reinterpret_cast<const float4 *>(m.dec_table + ((long long)hs.y0 * V + hs.y1) * m.jd) + c);
      const float4 ev = __ldg(reinterpret_cast<const float4 *>(d.enc + enc_next_row * m.jd) + c);
      store_joiner_input(d, xrow0 + q, m.jd, c, tanh4(ev, tv));
    }
  }

  // per emitted token statistics from its logits row, one warp per token (full-logits selection only; the
  // partial-record selection has them per row already)
  for (int e = warp; rows && e < sh.n_nodes; e += kSelThreads / 32) {
    const int b = sh.nodes[e].row;
    const float *row = rows + (long long)b * row_stride;
    const float mx = sh.mx[b], sum = sh.sum[b];
    float ent = 0.f, ts = 0.f, t1 = 0.f, t2 = 0.f;
    for (int v = lane; v < V; v += 32) {
      const float p = expf(row[v] - mx) / sum;
      ent += p * __logf(p + 1e-30f);
      ts += exp2f(__log2f(p) * (1.0f / 3.0f));        // p^(1/3); p = 0 -> log2 = -inf -> 0
      if (p > t1) { t2 = t1; t1 = p; } else if (p > t2) { t2 = p; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ent += __shfl_xor_sync(0xffffffffu, ent, o);
      ts += __shfl_xor_sync(0xffffffffu, ts, o);
      const float o1 = __shfl_xor_sync(0xffffffffu, t1, o), o2 = __shfl_xor_sync(0xffffffffu, t2, o);
      if (o1 > t1) { t2 = fmaxf(t1, o2); t1 = o1; } else { t2 = fmaxf(t2, o1); }
    }
    if (lane == 0) {
      const double a = 1.0 / 3.0;
      const double ts_max = m.ts_max;
      const double tsallis = (1.0 / (a - 1.0)) * (1.0 - (double)ts);
      const double max_ent = m.max_ent;
      ArenaNode &nd = arena[sh.nodes[e].arena_idx];
      nd.stats[0] = (float)(ts_max > 0 ? tsallis / ts_max : 0.0);
      nd.stats[1] = t1 - (V > 1 ? t2 : 1e-10f);
      nd.stats[2] = (float)((-(double)ent) / max_ent);
      nd.stats[3] = t1;
    }
  }
  const int n_new = sh.n_new;
  if (tid < n_new) d.hyps[((long long)(cur ^ 1) * d.n + s) * kMaxBeam + tid] = sh.new_[tid];
  SEL_PROF(4);

  // decoder pre-activation for the hypotheses with a new context: E = relu(P0[y0] + P1[y1]), the grouped k=2
  // convolution split into its two per-token halves (tables built once at load)
  if (more && !table) {
    const int dd4 = m.dd >> 2;
    for (int i = tid; i < n_new * dd4; i += kSelThreads) {
      const int q = i / dd4, c = i - q * dd4;
      if (sh.chg_pos[q] < 0) continue;
      const HypSlot &hs = sh.new_[q];
      const float4 a = __ldg(reinterpret_cast<const float4 *>(m.conv_p0 + (long long)hs.y0 * m.dd) + c);
      const float4 b = __ldg(reinterpret_cast<const float4 *>(m.conv_p1 + (long long)hs.y1 * m.dd) + c);
      reinterpret_cast<float4 *>(d.E + ((long long)s * d.beam + q) * m.dd)[c] =
          make_float4(fmaxf(a.x + b.x, 0.f), fmaxf(a.y + b.y, 0.f), fmaxf(a.z + b.z, 0.f), fmaxf(a.w + b.w, 0.f));
    }
  }
}

// warp-wide maximum of 64-bit keys with two 32-bit hardware reductions (redux.sync) instead of five rounds of 64-bit shuffles:
// first the high words, then the low words of the lanes that hold the maximal high word. Keys are unique (they carry the flat
// index), 0 = empty.
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
  const unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
  const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
  const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
  return ((unsigned long long)mh << 32) | ml;
}

// block-wide top-k over per-thread descending candidate lists loc[KB] (0 = empty): per warp KB reduction rounds, then
// warp 0 merges the 8 x KB survivors. Two barriers in total. Results in sh.win[0..k).
template <int KB>
__device__ void block_topk(unsigned long long (&loc)[KB], int k, SelShared &sh, unsigned long long *s_wk /* [8*KB] */) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int head = 0;
#pragma unroll 1
  for (int round = 0; round < KB; ++round) {
    if (round >= k) {                       // block-uniform: only the top k of each warp can matter
      if (lane == 0) s_wk[warp * KB + round] = 0ULL;
      continue;
    }
    unsigned long long best = 0ULL;
#pragma unroll
    for (int i = 0; i < KB; ++i) if (i == head) best = loc[i];
    const unsigned long long wb = warp_max_u64(best);
    if (best == wb && wb != 0ULL) ++head;
    if (lane == 0) s_wk[warp * KB + round] = wb;
  }
  __syncthreads();
  if (warp == 0) {
    constexpr int NC = (kSelThreads / 32) * KB;   // 32, 64 or 128 survivors
    constexpr int PER = (NC + 31) / 32;
    unsigned long long c[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) c[j] = (lane + 32 * j < NC) ? s_wk[lane + 32 * j] : 0ULL;
    for (int round = 0; round < k; ++round) {
      unsigned long long best = 0ULL;
#pragma unroll
      for (int j = 0; j < PER; ++j) best = c[j] > best ? c[j] : best;
      const unsigned long long wb = warp_max_u64(best);
      if (wb != 0ULL) {
#pragma unroll
        for (int j = 0; j < PER; ++j) if (c[j] == wb) c[j] = 0ULL;   // keys are unique (they carry the flat index)
      }
      if (lane == 0) sh.win[round] = wb;
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------ selection from the joiner's partial records
template <int KB>
__global__ void __launch_bounds__(kSelThreads) select_partials_kernel(SearchModel m, SearchDev d, ContextGraphView g, int has_graph,
                                                                      int t, int cur, int greedy) {
  const int s = blockIdx.x;
  pdl_wait();
  pdl_trigger();       // next step's decoder_joinin may become resident; it waits for this grid before reading anything
  trace_mark(d.trace, t, 8);
  if (s == 0 && threadIdx.x == 0) d.chg_count[t & 1] = 0;   // consumed by this step's decoder_joinin; refilled at t + 1
  if (t >= d.lens[s]) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int V = m.V;
  constexpr int REC = 4 + 2 * KB;
  constexpr int RPT = KB <= 4 ? 1 : (KB <= 8 ? 2 : 4);     // records per thread (host checks beam * parts <= 256 * RPT)
  __shared__ SelShared sh;
  __shared__ unsigned long long s_wk[(kSelThreads / 32) * KB];
  extern __shared__ __align__(16) float s_dyn[];           // [6][count][parts]: part max, sums (3), two best values
  long long _t0 = clock64();
  const int count = d.hyp_count[cur * d.n + s];
  const int P = (V + kPartCols - 1) / kPartCols;
  if (tid < count) sh.old_[tid] = d.hyps[((long long)cur * d.n + s) * kMaxBeam + tid];
  if (tid == kSelThreads - 1) sh.node_count0 = d.node_count[s];
  const int nrec = count * P;
  float *pm = s_dyn, *ps = pm + nrec, *pu = ps + nrec, *pt = pu + nrec, *pv0 = pt + nrec, *pv1 = pv0 + nrec;
  const float4 *recs = reinterpret_cast<const float4 *>(d.partials + (long long)s * d.beam * P * REC);
  float4 rv[RPT][KB / 4], ri[RPT][KB / 4];
#pragma unroll
  for (int it = 0; it < RPT; ++it) {
    const int i = tid + it * kSelThreads;
    if (i < nrec) {
      const float4 *rc = recs + (long long)i * (REC / 4);
      const float4 h = rc[0];
#pragma unroll
      for (int j = 0; j < KB / 4; ++j) { rv[it][j] = rc[1 + j]; ri[it][j] = rc[1 + KB / 4 + j]; }
      pm[i] = h.x; ps[i] = h.y; pu[i] = h.z; pt[i] = h.w;
      pv0[i] = rv[it][0].x; pv1[i] = rv[it][0].y;
    }
  }
  __syncthreads();
  SEL_PROF(0);
  // (a) per row, from its parts: max and log-sum-exp in float32 as the reference (:1096-1098), and the token
  // statistics of _compute_token_entropy (:1159-1181) with p = e^(x-M)/S:
  //   sum p log p = (1/S) sum_parts e^(m-M) (U + (m-M) S_part) - log S,   sum p^(1/3) = S^(-1/3) sum_parts e^((m-M)/3) T
  for (int b = warp; b < count; b += kSelThreads / 32) {
    float mx = -INFINITY;
    for (int q = lane; q < P; q += 32) mx = fmaxf(mx, pm[b * P + q]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f, su = 0.f, st = 0.f, t1 = -INFINITY, t2 = -INFINITY;
    for (int q = lane; q < P; q += 32) {
      const int i = b * P + q;
      const float dm = pm[i] - mx, e = expf(dm);
      sum = fmaf(ps[i], e, sum);
      su = fmaf(e, fmaf(dm, ps[i], pu[i]), su);
      st = fmaf(__expf(dm * (1.0f / 3.0f)), pt[i], st);
      const float a0 = pv0[i], a1 = pv1[i];            // a0 >= a1
      if (a0 > t1) { t2 = fmaxf(t1, a1); t1 = a0; } else { t2 = fmaxf(t2, a0); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sum += __shfl_xor_sync(0xffffffffu, sum, o);
      su += __shfl_xor_sync(0xffffffffu, su, o);
      st += __shfl_xor_sync(0xffffffffu, st, o);
      const float o1 = __shfl_xor_sync(0xffffffffu, t1, o), o2 = __shfl_xor_sync(0xffffffffu, t2, o);
      if (o1 > t1) { t2 = fmaxf(t1, o2); t1 = o1; } else { t2 = fmaxf(t2, o1); }
    }
    if (lane == 0) {
      const float lse = logf(sum);
      sh.mx[b] = mx;
      sh.sum[b] = sum;
      sh.lse[b] = lse;
      sh.prev[b] = greedy ? 0.f : (float)sh.old_[b].score;   // score rounded to f32 before the add (:1099-1100)
      // fp32 is ample for statistics compared at 1e-4 (FP64 is a slow path on this part and sits on the step's critical path)
      const float ts_max = (float)m.ts_max, max_ent = (float)m.max_ent;
      const float tsallis = -1.5f * (1.0f - st * __expf(-lse * (1.0f / 3.0f)));   // 1 / (alpha - 1), alpha = 1/3
      const float p1 = expf(t1 - mx) / sum, p2 = V > 1 ? expf(t2 - mx) / sum : 1e-10f;
      sh.rstats[b][0] = ts_max > 0.f ? tsallis / ts_max : 0.f;
      sh.rstats[b][1] = p1 - p2;
      sh.rstats[b][2] = -(su / sum - lse) / max_ent;
      sh.rstats[b][3] = p1;
    }
  }
  __syncthreads();
  SEL_PROF(1);
  // (b) global top-k; key = (ordered value, ~flat index) so max = value desc, index asc
  const int k = min(d.beam, count * V);
  unsigned long long loc[KB];
#pragma unroll
  for (int i = 0; i < KB; ++i) loc[i] = 0ULL;
#pragma unroll
  for (int it = 0; it < RPT; ++it) {
    const int i = tid + it * kSelThreads;
    if (i < nrec) {
      const int b = i / P;
      const float mxb = sh.mx[b], lseb = sh.lse[b], prevb = sh.prev[b];
#pragma unroll
      for (int j = 0; j < KB; ++j) {
        const float4 v4 = rv[it][j >> 2], i4 = ri[it][j >> 2];
        const float v = (j & 3) == 0 ? v4.x : (j & 3) == 1 ? v4.y : (j & 3) == 2 ? v4.z : v4.w;
        const int col = __float_as_int((j & 3) == 0 ? i4.x : (j & 3) == 1 ? i4.y : (j & 3) == 2 ? i4.z : i4.w);
        if (col < 0) continue;
        const float lp = ((v - mxb) - lseb) + prevb;
        const unsigned long long key = ((unsigned long long)ord_f32(lp) << 32) | (unsigned)(~(unsigned)(b * V + col));
        if (key > loc[KB - 1]) {
          unsigned long long carry = key;
#pragma unroll
          for (int q = 0; q < KB; ++q) {
            if (carry > loc[q]) { const unsigned long long tmp = loc[q]; loc[q] = carry; carry = tmp; }
          }
        }
      }
    }
  }
  block_topk<KB>(loc, k, sh, s_wk);
  SEL_PROF(2);
  finish_step(m, d, g, has_graph, s, t, cur, greedy, k, sh, nullptr, 0, _t0);
  trace_mark(d.trace, t, 9);
}

// ------------------------------------------------------------------ selection from the full logits (CUDA-core GEMM
// mode, and vocabulary/beam combinations the partial records do not cover)
template <int KB>
__global__ void __launch_bounds__(kSelThreads) select_step_kernel(SearchModel m, SearchDev d, ContextGraphView g, int has_graph,
                                                                  int t, int cur, int greedy, float blank_penalty) {
  const int s = blockIdx.x;
  pdl_wait();
  pdl_trigger();
  if (s == 0 && threadIdx.x == 0) d.chg_count[t & 1] = 0;
  if (t >= d.lens[s]) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int V = m.V;
  __shared__ SelShared sh;
  __shared__ unsigned long long s_wk[(kSelThreads / 32) * KB];

  long long _t0 = clock64();
  const int count = d.hyp_count[cur * d.n + s];
  if (tid < count) sh.old_[tid] = d.hyps[((long long)cur * d.n + s) * kMaxBeam + tid];
  if (tid == kSelThreads - 1) sh.node_count0 = d.node_count[s];
  // stage this utterance's logits rows in shared memory with wide, fully pipelined loads: every later pass
  // (max, sum-exp, top-k scan, token statistics) would otherwise pay L2 latency per element
  extern __shared__ __align__(16) float s_dyn[];
  float *lg = s_dyn;                                  // [count][V]
  {
    const float *glg = d.logits + (long long)s * d.beam * V;
    const int n4 = (count * V) >> 2;                  // V % 4 == 0 (checked on the host)
    const float4 *g4 = reinterpret_cast<const float4 *>(glg);
    float4 *l4 = reinterpret_cast<float4 *>(lg);
#pragma unroll 8
    for (int i = tid; i < n4; i += kSelThreads) l4[i] = g4[i];
  }
  __syncthreads();
  SEL_PROF(0);

  // (a) per-row max and log-sum-exp, float32 as the reference (:1096-1098)
  for (int b = warp; b < count; b += kSelThreads / 32) {
    float *row = lg + (long long)b * V;
    if (blank_penalty != 0.f && lane == 0) row[m.blank_id] -= blank_penalty;
    __syncwarp();
    float mx = -INFINITY;
    for (int v = lane; v < V; v += 32) mx = fmaxf(mx, row[v]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int v = lane; v < V; v += 32) sum += expf(row[v] - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) {
      sh.mx[b] = mx;
      sh.sum[b] = sum;
      sh.lse[b] = logf(sum);
      sh.prev[b] = greedy ? 0.f : (float)sh.old_[b].score;   // score rounded to f32 before the add (:1099-1100)
    }
  }
  __syncthreads();

  SEL_PROF(1);
  // (b) global top-k over count*V candidates; key = (ordered value, ~flat index) so max = value desc, index asc
  const int k = min(d.beam, count * V);
  unsigned long long loc[KB];   // KB >= beam: per-thread candidates, descending
#pragma unroll
  for (int i = 0; i < KB; ++i) loc[i] = 0ULL;
  for (int b = 0; b < count; ++b) {
    const float mxb = sh.mx[b], lseb = sh.lse[b], prevb = sh.prev[b];
    const float *row = lg + b * V;
    const int base = b * V;
    for (int v = tid; v < V; v += kSelThreads) {
      const float lp = ((row[v] - mxb) - lseb) + prevb;
      const unsigned long long key = ((unsigned long long)ord_f32(lp) << 32) | (unsigned)(~(unsigned)(base + v));
      if (key > loc[KB - 1]) {
        // insertion into the descending local list (fully unrolled so it stays in registers)
        unsigned long long carry = key;
#pragma unroll
        for (int i = 0; i < KB; ++i) {
          if (carry > loc[i]) { const unsigned long long tmp = loc[i]; loc[i] = carry; carry = tmp; }
        }
      }
    }
  }
  block_topk<KB>(loc, k, sh, s_wk);
  SEL_PROF(2);
  finish_step(m, d, g, has_graph, s, t, cur, greedy, k, sh, lg, V, _t0);
}

__global__ void bias_penalty_kernel(const float *__restrict__ b, int V, int blank_id, float penalty, float *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < V) out[i] = b[i] - (i == blank_id ? penalty : 0.f);
}

__global__ void init_search_kernel(SearchModel m, SearchDev d, int n_pos) {
  const int s = blockIdx.x;
  if (threadIdx.x == 0) {
    HypSlot h;
    h.score = 0.0; h.hash = 0x12345ULL; h.node = -1; h.len = 0; h.ctx = 0; h.y0 = 0; h.y1 = 0; h.dec_src = -1;
    d.hyps[(long long)s * kMaxBeam] = h;     // ys = [-1, 0] -> decoder input [0, 0] (:1051-1052)
    d.hyp_count[s] = 1;
    d.hyp_count[d.n + s] = 0;
    d.node_count[s] = 0;
    // frame 0: the first row of every utterance that has frames needs its decoder output (longest first: those are a
    // prefix; an empty utterance has no encoder row to pair it with)
    if (s < n_pos) d.chg_list[s] = make_int2(s * d.beam, (int)d.enc_off[s]);
    if (s == 0) { d.chg_count[0] = n_pos; d.chg_count[1] = 0; }
  }
  if ((int)threadIdx.x < d.beam) {
    d.rowdesc[(size_t)s * d.beam + threadIdx.x] = make_int2((threadIdx.x == 0 && s < n_pos) ? -1 : -2, (int)d.enc_off[s]);
    d.rowdesc[(size_t)(d.n + s) * d.beam + threadIdx.x] = make_int2(-2, 0);
  }
  if (m.dec_table) {   // joiner input of frame 0: context [0, 0] is row 0 of the decoder table
    if (s < n_pos)
      for (int c = threadIdx.x; c < (m.jd >> 2); c += blockDim.x) {
        const float4 e = __ldg(reinterpret_cast<const float4 *>(d.enc + d.enc_off[s] * m.jd) + c);
        const float4 v = __ldg(reinterpret_cast<const float4 *>(m.dec_table) + c);
        store_joiner_input(d, (long long)s * d.beam, m.jd, c, make_float4(tanhf(e.x + v.x), tanhf(e.y + v.y), tanhf(e.z + v.z), tanhf(e.w + v.w)));
      }
    return;
  }
  for (int o = threadIdx.x; o < m.dd; o += blockDim.x)   // context [0, 0]
    d.E[((long long)s * d.beam) * m.dd + o] = fmaxf(__ldg(m.conv_p0 + o) + __ldg(m.conv_p1 + o), 0.f);
}

// decoder-table build: pre-activations of the contexts ctx0 .. ctx0 + rows
__global__ void context_pre_kernel(SearchModel m, long long ctx0, int rows, float *__restrict__ E) {
  const int dd4 = m.dd >> 2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * dd4) return;
  const int r = (int)(i / dd4), c = (int)(i - (long long)r * dd4);
  const long long ctx = ctx0 + r;
  const int y0 = (int)(ctx / m.V), y1 = (int)(ctx - (long long)y0 * m.V);
  const float4 a = __ldg(reinterpret_cast<const float4 *>(m.conv_p0 + (long long)y0 * m.dd) + c);
  const float4 b = __ldg(reinterpret_cast<const float4 *>(m.conv_p1 + (long long)y1 * m.dd) + c);
  reinterpret_cast<float4 *>(E)[i] = make_float4(fmaxf(a.x + b.x, 0.f), fmaxf(a.y + b.y, 0.f), fmaxf(a.z + b.z, 0.f), fmaxf(a.w + b.w, 0.f));
}

// decoder-table mode of the parity hook: dec = table[y0, y1], X = tanh(enc + dec)
__global__ void table_rows_kernel(SearchModel m, const long long *__restrict__ y, const float *__restrict__ enc, int rows,
                                  float *__restrict__ dec_out, float *__restrict__ x_out) {
  const int i = blockIdx.x;
  if (i >= rows) return;
  const int y0 = (int)max(0LL, y[2 * i]), y1 = (int)max(0LL, y[2 * i + 1]);
  const float *row = m.dec_table + ((long long)y0 * m.V + y1) * m.jd;
  for (int o = threadIdx.x; o < m.jd; o += blockDim.x) {
    const float v = __ldg(row + o);
    dec_out[(long long)i * m.jd + o] = v;
    x_out[(long long)i * m.jd + o] = tanhf((enc ? enc[(long long)i * m.jd + o] : 0.f) + v);
  }
}

// finalize (:1143-1153): subtract unfinished hotword score, pick first max of log_prob/len(ys), unroll the chain.
// Results are packed per utterance at out_off[u] (= sum of T' of the utterances before u in the caller's order; an
// utterance emits at most one token per frame, so T' slots always suffice).
__global__ void finalize_kernel(SearchDev d, ContextGraphView g, int has_graph, const int *__restrict__ final_buf,
                                const int *__restrict__ orig_index, const long long *__restrict__ out_off, int *__restrict__ n_tokens,
                                int *__restrict__ tokens, int *__restrict__ frames, float *__restrict__ tok_lp,
                                float *__restrict__ stats) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= d.n) return;
  const int cur = final_buf[s];
  const int count = d.hyp_count[cur * d.n + s];
  const HypSlot *hy = d.hyps + ((long long)cur * d.n + s) * kMaxBeam;
  int best = -1;
  double best_v = 0.0;
  for (int b = 0; b < count; ++b) {
    double lp = hy[b].score;
    if (has_graph) lp += cg_finalize(g, hy[b].ctx);
    const double v = lp / (double)max(hy[b].len + 2, 1);
    if (best < 0 || v > best_v) { best = b; best_v = v; }
  }
  const int u = orig_index[s];
  if (best < 0) { n_tokens[u] = 0; return; }
  const ArenaNode *arena = d.arena + d.arena_off[s];
  const int len = hy[best].len, cap = d.lens[s];
  n_tokens[u] = len;
  int node = hy[best].node;
  const long long base = out_off[u];
  for (int i = len - 1; i >= 0 && node >= 0; --i) {
    const ArenaNode &nd = arena[node];
    if (i < cap) {
      const long long o = base + i;
      tokens[o] = nd.token; frames[o] = nd.frame; tok_lp[o] = nd.tok_lp;
      reinterpret_cast<float4 *>(stats)[o] = make_float4(nd.stats[0], nd.stats[1], nd.stats[2], nd.stats[3]);
    }
    node = nd.parent;
  }
}

template <typename T>
void ensure(T *&p, size_t &cap, size_t need) {
  if (need <= cap) return;
  if (p) cudaFree(p);
  p = nullptr;
  cap = need + need / 4;
  CUDA_CHECK(cudaMalloc(&p, cap * sizeof(T)));
}

}  // namespace

struct SearchState {
  HypSlot *hyps = nullptr; size_t hyps_cap = 0;
  int *hyp_count = nullptr; size_t hc_cap = 0;
  int *node_count = nullptr; size_t nc_cap = 0;
  ArenaNode *arena = nullptr; size_t arena_cap = 0;
  float *E = nullptr; size_t e_cap = 0;
  float *dec = nullptr; size_t dec_cap = 0;
  float *X = nullptr; size_t x_cap = 0;
  uint2 *X16 = nullptr; size_t x16_cap = 0;   // hi plane then lo plane
  float *logits = nullptr; size_t lg_cap = 0;
  float *partials = nullptr; size_t pt_cap = 0;
  float *bias_pen = nullptr; size_t bp_cap = 0;
  int2 *chg_list = nullptr; size_t cl_cap = 0;
  int2 *rowdesc = nullptr; size_t rd_cap = 0;
  int *chg_count = nullptr; size_t cc_cap = 0;
  long long *enc_off = nullptr; size_t eo_cap = 0;
  long long *arena_off = nullptr; size_t ao_cap = 0;
  int *lens = nullptr; size_t ln_cap = 0;
  int *orig = nullptr; size_t og_cap = 0;
  int *final_buf = nullptr; size_t fb_cap = 0;
  int *o_ntok = nullptr; size_t ont_cap = 0;
  int *o_tok = nullptr; size_t ot_cap = 0;
  int *o_frm = nullptr; size_t of_cap = 0;
  float *o_lp = nullptr; size_t ol_cap = 0;
  float *o_st = nullptr; size_t os_cap = 0;
  long long *o_off = nullptr; size_t oo_cap = 0;
  void (*gemm)(const GemmArgs &, cudaStream_t) = launch_gemm_fp32;
  bool fused_partials = false;
  // results of the last issued search: pinned host copies in the packed layout (h_off[u] = sum of T' before u)
  void *h_pinned = nullptr; size_t h_cap = 0;
  int *h_ntok = nullptr, *h_tok = nullptr, *h_frm = nullptr;
  float *h_lp = nullptr, *h_st = nullptr;
  std::vector<long long> h_off;
  int n_last = 0;
  long long *d_prof = nullptr;
  unsigned long long *d_trace = nullptr;
  int prof_steps = 0;
  bool table_mode = false;   // last issued search ran without decoder_joinin (decoder table)
};

SearchState *search_state_create() { return new SearchState(); }
void search_set_gemm(SearchState *s, void (*fn)(const GemmArgs &, cudaStream_t), bool fused_partials) {
  s->gemm = fn;
  s->fused_partials = fused_partials;
}
void search_state_destroy(SearchState *s) {
  if (!s) return;
  cudaFree(s->hyps); cudaFree(s->hyp_count); cudaFree(s->node_count); cudaFree(s->arena); cudaFree(s->E); cudaFree(s->dec);
  cudaFree(s->X); cudaFree(s->X16); cudaFree(s->logits); cudaFree(s->partials); cudaFree(s->bias_pen); cudaFree(s->chg_list); cudaFree(s->rowdesc); cudaFree(s->chg_count);
  cudaFree(s->enc_off); cudaFree(s->arena_off); cudaFree(s->lens);
  cudaFree(s->orig); cudaFree(s->final_buf); cudaFree(s->o_ntok); cudaFree(s->o_tok); cudaFree(s->o_frm);
  cudaFree(s->o_lp); cudaFree(s->o_st); cudaFree(s->o_off);
  if (s->h_pinned) cudaFreeHost(s->h_pinned);
  delete s;
}

void search_issue(SearchState *S, const SearchModel &m, const ContextGraphDev *g, const float *enc, const int *h_lens, int n,
                  int method, int beam, float blank_penalty, cudaStream_t st) {
  S->n_last = 0;
  if (n <= 0) return;
  const int greedy = (method == 0);
  if (greedy) beam = 1;
  if (beam < 1 || beam > kMaxBeam) throw CudaError("beam search: max_active_paths must be in [1,16]");
  // longest first so the active set is a prefix
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return h_lens[a] > h_lens[b]; });
  std::vector<long long> off_orig(n + 1, 0);
  for (int u = 0; u < n; ++u) off_orig[u + 1] = off_orig[u] + std::max(h_lens[u], 0);
  std::vector<long long> enc_off(n), arena_off(n);
  std::vector<int> lens(n), final_buf(n);
  long long arena_total = 0;
  int max_len = 0;
  for (int s = 0; s < n; ++s) {
    const int u = order[s];
    lens[s] = std::max(h_lens[u], 0);
    enc_off[s] = off_orig[u];
    arena_off[s] = arena_total;
    arena_total += (long long)lens[s] * beam;
    max_len = std::max(max_len, lens[s]);
    final_buf[s] = lens[s] & 1;   // after T' steps the live state sits in buffer (T' mod 2)
  }
  // selection from the joiner's partial records when the GEMM emits them and one CTA pass covers beam x parts
  const int KB = beam <= 4 ? 4 : (beam <= 8 ? 8 : 16);
  const int RPT = KB <= 4 ? 1 : (KB <= 8 ? 2 : 4);
  const int P = (m.V + kPartCols - 1) / kPartCols;
  const bool fused = S->fused_partials && (long long)beam * P <= (long long)kSelThreads * RPT && (m.jd & 3) == 0 && (m.dd & 3) == 0;
  const int REC = part_rec_floats(KB);
  const size_t rows = (size_t)n * beam;
  ensure(S->hyps, S->hyps_cap, 2 * (size_t)n * kMaxBeam);
  ensure(S->hyp_count, S->hc_cap, 2 * (size_t)n);
  ensure(S->node_count, S->nc_cap, (size_t)n);
  ensure(S->arena, S->arena_cap, (size_t)std::max<long long>(arena_total, 1));
  ensure(S->E, S->e_cap, rows * m.dd);
  ensure(S->dec, S->dec_cap, 2 * rows * m.jd);
  ensure(S->X, S->x_cap, rows * m.jd);
  // decoder-table mode on the fp16 operand split: the selection kernel writes the joiner input already split, and the joiner
  // GEMM feeds both operands to the MMAs straight from TMA (B200ASR_JOINER_SS=0: fp32 X through the converting kernel)
  static const bool ss_env = !(getenv("B200ASR_JOINER_SS") && atoi(getenv("B200ASR_JOINER_SS")) == 0);
  const bool x16 = ss_env && fused && m.dec_table && m.join_w16hi && m.join_w16lo && (m.jd & 7) == 0;
  if (x16) ensure(S->X16, S->x16_cap, 2 * rows * (size_t)(m.jd >> 2));
  if (!fused) ensure(S->logits, S->lg_cap, rows * m.V);
  if (fused) ensure(S->partials, S->pt_cap, rows * P * REC);
  ensure(S->chg_list, S->cl_cap, 2 * rows);
  ensure(S->rowdesc, S->rd_cap, 2 * rows);
  ensure(S->chg_count, S->cc_cap, (size_t)2);
  ensure(S->enc_off, S->eo_cap, (size_t)n);
  ensure(S->arena_off, S->ao_cap, (size_t)n);
  ensure(S->lens, S->ln_cap, (size_t)n);
  ensure(S->orig, S->og_cap, (size_t)n);
  ensure(S->final_buf, S->fb_cap, (size_t)n);
  const size_t slots = (size_t)std::max<long long>(off_orig[n], 1);
  ensure(S->o_ntok, S->ont_cap, (size_t)n);
  ensure(S->o_tok, S->ot_cap, slots);
  ensure(S->o_frm, S->of_cap, slots);
  ensure(S->o_lp, S->ol_cap, slots);
  ensure(S->o_st, S->os_cap, slots * 4);
  ensure(S->o_off, S->oo_cap, (size_t)n);
  {
    // pinned host mirror: [n] counts, then tokens / frames / log-probs [slots] each, then stats [slots][4] (16-byte aligned)
    const size_t n4 = ((size_t)n + 3) & ~size_t(3), s4 = (slots + 3) & ~size_t(3);
    const size_t need = (n4 + 3 * s4 + 4 * s4) * 4;
    if (need > S->h_cap) {
      if (S->h_pinned) CUDA_CHECK(cudaFreeHost(S->h_pinned));
      S->h_pinned = nullptr;
      S->h_cap = need + need / 4;
      CUDA_CHECK(cudaHostAlloc(&S->h_pinned, S->h_cap, cudaHostAllocPortable));
    }
    int *base = reinterpret_cast<int *>(S->h_pinned);
    S->h_ntok = base; S->h_tok = base + n4; S->h_frm = S->h_tok + s4;
    S->h_lp = reinterpret_cast<float *>(S->h_frm + s4); S->h_st = S->h_lp + s4;
  }
  S->h_off.assign(off_orig.begin(), off_orig.end());
  CUDA_CHECK(cudaMemcpyAsync(S->o_off, off_orig.data(), n * sizeof(long long), cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(S->enc_off, enc_off.data(), n * sizeof(long long), cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(S->arena_off, arena_off.data(), n * sizeof(long long), cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(S->lens, lens.data(), n * sizeof(int), cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(S->orig, order.data(), n * sizeof(int), cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(S->final_buf, final_buf.data(), n * sizeof(int), cudaMemcpyHostToDevice, st));

  SearchDev d;
  d.enc = enc; d.enc_off = S->enc_off; d.lens = S->lens; d.arena_off = S->arena_off; d.hyps = S->hyps;
  d.hyp_count = S->hyp_count; d.node_count = S->node_count; d.arena = S->arena; d.E = S->E; d.dec = S->dec; d.X = S->X;
  d.X16hi = x16 ? S->X16 : nullptr; d.X16lo = x16 ? S->X16 + rows * (size_t)(m.jd >> 2) : nullptr;
  d.logits = S->logits; d.partials = fused ? S->partials : nullptr; d.chg_list = S->chg_list; d.chg_count = S->chg_count; d.rowdesc = S->rowdesc;
  d.n = n; d.beam = beam;
  d.prof = nullptr;
  static const bool want_prof = getenv("B200ASR_SEARCH_PROF") != nullptr;
  if (S->d_prof) { cudaFree(S->d_prof); S->d_prof = nullptr; }
  if (S->d_trace) { cudaFree(S->d_trace); S->d_trace = nullptr; }
  long long *d_prof = nullptr;
  if (want_prof) {
    CUDA_CHECK(cudaMalloc(&d_prof, 8 * sizeof(long long)));
    CUDA_CHECK(cudaMemsetAsync(d_prof, 0, 8 * sizeof(long long), st));
    d.prof = d_prof;
  }
  d.trace = nullptr;
  unsigned long long *d_trace = nullptr;
  if (want_prof && fused) {
    CUDA_CHECK(cudaMalloc(&d_trace, (size_t)std::max(max_len, 1) * 16 * sizeof(unsigned long long)));
    CUDA_CHECK(cudaMemsetAsync(d_trace, 0, (size_t)std::max(max_len, 1) * 16 * sizeof(unsigned long long), st));
    d.trace = d_trace;
  }
  S->d_prof = d_prof; S->d_trace = d_trace; S->prof_steps = max_len; S->table_mode = m.dec_table != nullptr;
  ContextGraphView gv{};
  const int has_graph = (g && g->n_nodes > 1 && !greedy) ? 1 : 0;
  if (has_graph)
    gv = ContextGraphView{g->n_nodes, g->edge_start, g->edge_token, g->edge_child, g->fail, g->token, g->is_end, g->output,
                          g->token_score, g->node_score, g->output_score};

  if (m.V & 3) throw CudaError("beam search: vocab_size must be a multiple of 4");
  if (!fused) {
    set_max_dynamic_smem(select_step_kernel<4>, 200 * 1024);
    set_max_dynamic_smem(select_step_kernel<8>, 200 * 1024);
    set_max_dynamic_smem(select_step_kernel<16>, 200 * 1024);
    if ((size_t)beam * m.V * sizeof(float) > 200 * 1024) throw CudaError("beam * vocab_size too large for the selection kernel");
  }
  set_max_dynamic_smem(decoder_joinin_kernel, kDjSmem);
  if (!m.conv_p0 || !m.conv_p1) throw CudaError("beam search: decoder convolution tables are missing");
  int n_pos = n;
  while (n_pos > 0 && lens[n_pos - 1] <= 0) --n_pos;
  init_search_kernel<<<n, 128, 0, st>>>(m, d, n_pos);
  count_launch(); KERNEL_CHECK();
  // the joiner epilogue takes the blank penalty folded into its bias
  const float *join_bias = m.join_b;
  if (fused && blank_penalty != 0.f) {
    ensure(S->bias_pen, S->bp_cap, (size_t)m.V);
    bias_penalty_kernel<<<(m.V + 255) / 256, 256, 0, st>>>(m.join_b, m.V, m.blank_id, blank_penalty, S->bias_pen);
    count_launch(); KERNEL_CHECK();
    join_bias = S->bias_pen;
  }
  int n_active = n;
  // programmatic dependent launch: each step kernel becomes resident (and runs its prologue) while its predecessor
  // drains; B200ASR_NO_PDL=1 falls back to plain stream order (debugging)
  static const bool use_pdl = getenv("B200ASR_NO_PDL") == nullptr;
  const int ntn = (m.jd + DJ_TN - 1) / DJ_TN;
  int n_sms = 148;
  {
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    CUDA_CHECK(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  for (int t = 0; t < max_len; ++t) {
    while (n_active > 0 && lens[n_active - 1] <= t) --n_active;
    const int cur = t & 1;
    const int act_rows = n_active * beam;
    // decoder outputs for the changed contexts + joiner input X = tanh(enc_t + dec) for every live row
    // (with the decoder table the previous selection - or init_search_kernel - has already written X)
    if (!m.dec_table) {
      const int dj_grid = std::min(2 * n_sms, std::max(((act_rows + DJ_TM - 1) / DJ_TM) * ntn, 1));
      launch_pdl(decoder_joinin_kernel, dim3(dj_grid), dim3(DJ_THREADS), kDjSmem, st, use_pdl, m, d, t, act_rows);
      count_launch();
    }
    // joiner output_linear: logits = X * Wj^T + bj (+ partial records)
    GemmArgs ga{};
    ga.A = S->X; ga.lda = m.jd; ga.W = m.join_w; ga.Wlo = m.join_w_lo; ga.bias = m.join_b; ga.R = nullptr; ga.ldr = 0; ga.C = S->logits;
    ga.W16hi = m.join_w16hi; ga.W16lo = m.join_w16lo; ga.w16_ld = m.join_w16_ld;
    ga.ldc = m.V; ga.M = act_rows; ga.N = m.V; ga.K = m.jd; ga.act = ACT_NONE; ga.pdl = use_pdl ? 1 : 0;
    if (fused) {   // records only: nothing downstream reads the logits
      ga.act = ACT_JOINER; ga.C = nullptr; ga.partials = S->partials; ga.part_kb = KB; ga.bias = join_bias;
      if (x16) { ga.A = nullptr; ga.A16hi = d.X16hi; ga.A16lo = d.X16lo; ga.a16_ld = m.jd; ga.a16_rows = (int)rows; }
      ga.trace = d_trace ? d_trace + (size_t)t * 16 + 2 : nullptr;
    }
    S->gemm(ga, st);
    if (fused) {
      const size_t smem = (size_t)6 * beam * P * sizeof(float);
      if (KB == 4) launch_pdl(select_partials_kernel<4>, dim3(n_active), dim3(kSelThreads), smem, st, use_pdl, m, d, gv, has_graph, t, cur, greedy);
      else if (KB == 8) launch_pdl(select_partials_kernel<8>, dim3(n_active), dim3(kSelThreads), smem, st, use_pdl, m, d, gv, has_graph, t, cur, greedy);
      else launch_pdl(select_partials_kernel<16>, dim3(n_active), dim3(kSelThreads), smem, st, use_pdl, m, d, gv, has_graph, t, cur, greedy);
    } else {
      const size_t sel_smem = (size_t)beam * m.V * sizeof(float);
      if (KB == 4) launch_pdl(select_step_kernel<4>, dim3(n_active), dim3(kSelThreads), sel_smem, st, use_pdl, m, d, gv, has_graph, t, cur, greedy, blank_penalty);
      else if (KB == 8) launch_pdl(select_step_kernel<8>, dim3(n_active), dim3(kSelThreads), sel_smem, st, use_pdl, m, d, gv, has_graph, t, cur, greedy, blank_penalty);
      else launch_pdl(select_step_kernel<16>, dim3(n_active), dim3(kSelThreads), sel_smem, st, use_pdl, m, d, gv, has_graph, t, cur, greedy, blank_penalty);
    }
    count_launch();
  }
  KERNEL_CHECK();
  // Utterances that stopped at step T' hold their final state in buffer (T' & 1): select writes to cur^1 = (t+1)&1.
  finalize_kernel<<<(n + 127) / 128, 128, 0, st>>>(d, gv, has_graph, S->final_buf, S->orig, S->o_off, S->o_ntok, S->o_tok, S->o_frm,
                                                   S->o_lp, S->o_st);
  count_launch(); KERNEL_CHECK();
  CUDA_CHECK(cudaMemcpyAsync(S->h_ntok, S->o_ntok, n * sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaMemcpyAsync(S->h_tok, S->o_tok, slots * sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaMemcpyAsync(S->h_frm, S->o_frm, slots * sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaMemcpyAsync(S->h_lp, S->o_lp, slots * sizeof(float), cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaMemcpyAsync(S->h_st, S->o_st, slots * 4 * sizeof(float), cudaMemcpyDeviceToHost, st));
  S->n_last = n;
}

long long search_result_bytes(const SearchState *S) {
  if (!S || S->n_last <= 0) return 0;
  const long long slots = S->h_off.empty() ? 0 : S->h_off.back();
  return (long long)S->n_last * 4 + slots * 28;
}

// Valid once the stream search_issue ran on has been synchronised.
SearchView search_view(const SearchState *S) {
  SearchView v{};
  v.n = S->n_last;
  if (v.n <= 0) return v;
  v.n_tokens = S->h_ntok; v.off = S->h_off.data(); v.tokens = S->h_tok; v.frames = S->h_frm; v.tok_lp = S->h_lp; v.stats = S->h_st;
  return v;
}

void search_print_prof(SearchState *S) {
  if (S->d_prof) {
    long long h[8];
    CUDA_CHECK(cudaMemcpy(h, S->d_prof, sizeof h, cudaMemcpyDeviceToHost));
    cudaFree(S->d_prof); S->d_prof = nullptr;
    const int max_len = S->prof_steps;
    const char *names[8] = {"load", "logsumexp", "top-k", "expand/dedup", "stats+writeback", "merges (count)", "expansion: parallel half", "expansion: serial half"};
    fprintf(stderr, "[b200asr search prof] CTA0 cycles over %d steps:", max_len);
    for (int i = 0; i < 8; ++i) fprintf(stderr, " %s=%.1f/step", names[i], (double)h[i] / std::max(max_len, 1));
    fprintf(stderr, "\n");
  }
  if (S->d_trace) {
    const int max_len = S->prof_steps;
    std::vector<unsigned long long> h((size_t)max_len * 16);
    CUDA_CHECK(cudaMemcpy(h.data(), S->d_trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    cudaFree(S->d_trace); S->d_trace = nullptr;
    // average timeline of a step (CTA 0 of each kernel), microseconds relative to decoder_joinin's entry
    const char *names[16] = {"dj.in", "dj.out", "gemm.in", "gemm.out", "gemm.prologue", "gemm.first_operands", "gemm.acc_done",
                             "gemm.epilogue_done", "sel.in", "sel.out", "dj.issued", "dj.rows", "dj.landed", "dj.mma", "dj.tile0", ""};
    double acc[16] = {0};
    double step = 0;
    int cnt = 0;
    // decoder-table mode: a step starts with the joiner GEMM (slot 2) and has no decoder_joinin marks
    const int b0 = S->table_mode ? 2 : 0;
    auto used = [&](int i) { return !S->table_mode || (i >= 2 && i <= 9); };
    for (int t = 1; t + 1 < max_len; ++t) {
      const unsigned long long *r = &h[(size_t)t * 16], *nx = &h[(size_t)(t + 1) * 16];
      bool ok = nx[b0] != 0;
      for (int i = 0; i < 15; ++i) ok = ok && (!used(i) || r[i] != 0);
      if (!ok) continue;
      for (int i = 0; i < 15; ++i) if (used(i)) acc[i] += (double)((long long)(r[i] - r[b0]));
      step += (double)(nx[b0] - r[b0]);
      ++cnt;
    }
    {
      // the same per eighth of the search (the active set shrinks with t: early steps carry every utterance)
      fprintf(stderr, "[b200asr search trace] step us by eighth of the %d steps (gemm, select, whole step):", max_len);
      for (int b = 0; b < 8; ++b) {
        double g = 0, sl = 0, w = 0;
        int c = 0;
        for (int t = std::max(1, b * max_len / 8); t < (b + 1) * max_len / 8 && t + 1 < max_len; ++t) {
          const unsigned long long *r = &h[(size_t)t * 16], *nx = &h[(size_t)(t + 1) * 16];
          if (!r[2] || !r[3] || !r[8] || !r[9] || !nx[b0] || !r[b0]) continue;
          g += (double)((long long)(r[3] - r[2])); sl += (double)((long long)(r[9] - r[8])); w += (double)((long long)(nx[b0] - r[b0]));
          ++c;
        }
        if (c) fprintf(stderr, " [%.1f %.1f %.1f]", g / c / 1e3, sl / c / 1e3, w / c / 1e3);
      }
      fprintf(stderr, "\n");
    }
    if (cnt) {
      fprintf(stderr, "[b200asr search trace] %d steps, mean us from step start:", cnt);
      const int order[15] = {0, 10, 11, 12, 13, 14, 1, 2, 4, 5, 6, 7, 3, 8, 9};
      for (int i : order) if (used(i)) fprintf(stderr, " %s=%.2f", names[i], acc[i] / cnt / 1e3);
      fprintf(stderr, " next_step=%.2f\n", step / cnt / 1e3);
    }
  }
}

// Scatter of the packed results into caller arrays of pitch max_tokens (the raw B200AsrBeamSearch layout).
void search_collect(SearchState *S, SearchResultHost *out) {
  search_print_prof(S);
  const SearchView v = search_view(S);
  for (int u = 0; u < v.n; ++u) {
    out->n_tokens[u] = v.n_tokens[u];
    if (!out->tokens) continue;
    const int cnt = std::min(v.n_tokens[u], out->max_tokens);
    const long long o = v.off[u], q = (long long)u * out->max_tokens;
    for (int j = 0; j < cnt; ++j) {
      out->tokens[q + j] = v.tokens[o + j]; out->frames[q + j] = v.frames[o + j]; out->tok_lp[q + j] = v.tok_lp[o + j];
      for (int c = 0; c < 4; ++c) out->stats[(q + j) * 4 + c] = v.stats[(o + j) * 4 + c];
    }
  }
}

void run_search(SearchState *S, const SearchModel &m, const ContextGraphDev *g, const float *enc, const int *h_lens, int n,
                int method, int beam, float blank_penalty, SearchResultHost *out, cudaStream_t st) {
  if (n <= 0) return;
  search_issue(S, m, g, enc, h_lens, n, method, beam, blank_penalty, st);
  CUDA_CHECK(cudaStreamSynchronize(st));
  search_collect(S, out);
}

// ---- the product's per-step kernels on caller-supplied rows (parity tests call them through the C-ABI) ----
namespace {
__global__ void prep_product_rows_kernel(SearchModel m, SearchDev d, const long long *__restrict__ y, int rows) {
  const int i = blockIdx.x;
  if (i >= rows) return;
  const int y0 = (int)max(0LL, y[2 * i]), y1 = (int)max(0LL, y[2 * i + 1]);
  for (int o = threadIdx.x; o < m.dd; o += blockDim.x)
    d.E[(long long)i * m.dd + o] = fmaxf(__ldg(m.conv_p0 + (long long)y0 * m.dd + o) + __ldg(m.conv_p1 + (long long)y1 * m.dd + o), 0.f);
  if (threadIdx.x == 0) {
    d.chg_list[i] = make_int2(i, i);
    d.rowdesc[i] = make_int2(-1, i);
    if (i == 0) { d.chg_count[0] = rows; d.chg_count[1] = 0; }
  }
}
}  // namespace

// decoder_joinin_kernel exactly as a frame step launches it, with every row in the recompute list: dec = decoder_proj(relu(conv
// (emb[y0], emb[y1]))) and X = tanh(enc + dec). enc may be null (zeros). All pointers device.
void launch_decoder_product_rows(const SearchModel &m, const long long *y, const float *enc, int rows, float *dec_out, float *x_out,
                                 cudaStream_t st) {
  if (rows <= 0) return;
  if (m.dec_table) {   // what a frame step reads in decoder-table mode
    table_rows_kernel<<<rows, 128, 0, st>>>(m, y, enc, rows, dec_out, x_out);
    count_launch(); KERNEL_CHECK();
    CUDA_CHECK(cudaStreamSynchronize(st));
    return;
  }
  if (!m.conv_p0 || !m.conv_p1) throw CudaError("decoder convolution tables are missing");
  float *E = nullptr, *dec = nullptr, *X = nullptr, *zero = nullptr;
  int2 *lists = nullptr;
  int *cnt = nullptr;
  CUDA_CHECK(cudaMalloc(&E, (size_t)rows * m.dd * sizeof(float)));
  CUDA_CHECK(cudaMalloc(&dec, 2 * (size_t)rows * m.jd * sizeof(float)));
  CUDA_CHECK(cudaMalloc(&X, (size_t)rows * m.jd * sizeof(float)));
  CUDA_CHECK(cudaMalloc(&lists, 4 * (size_t)rows * sizeof(int2)));
  CUDA_CHECK(cudaMalloc(&cnt, 2 * sizeof(int)));
  if (!enc) {
    CUDA_CHECK(cudaMalloc(&zero, ((size_t)rows + 1) * m.jd * sizeof(float)));
    CUDA_CHECK(cudaMemsetAsync(zero, 0, ((size_t)rows + 1) * m.jd * sizeof(float), st));
  }
  SearchDev d{};
  d.enc = enc ? enc : zero; d.E = E; d.dec = dec; d.X = X; d.chg_list = lists; d.rowdesc = lists + 2 * (size_t)rows; d.chg_count = cnt;
  d.n = rows; d.beam = 1; d.trace = nullptr; d.prof = nullptr;
  prep_product_rows_kernel<<<rows, 128, 0, st>>>(m, d, y, rows);
  count_launch(); KERNEL_CHECK();
  set_max_dynamic_smem(decoder_joinin_kernel, kDjSmem);
  const int ntn = (m.jd + DJ_TN - 1) / DJ_TN;
  const int grid = std::min(296, std::max(((rows + DJ_TM - 1) / DJ_TM) * ntn, 1));
  launch_pdl(decoder_joinin_kernel, dim3(grid), dim3(DJ_THREADS), kDjSmem, st, false, m, d, 0, rows);
  count_launch(); KERNEL_CHECK();
  CUDA_CHECK(cudaMemcpyAsync(dec_out, dec, (size_t)rows * m.jd * sizeof(float), cudaMemcpyDeviceToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(x_out, X, (size_t)rows * m.jd * sizeof(float), cudaMemcpyDeviceToDevice, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  cudaFree(E); cudaFree(dec); cudaFree(X); cudaFree(lists); cudaFree(cnt); cudaFree(zero);
}

// The joiner GEMM exactly as a frame step launches it (ACT_JOINER: per row and 32-column part a softmax / top-kb record, no
// logits stored). X [rows, jd] -> records [rows, ceil(V/32), 4 + 2 kb]. Device pointers.
namespace {
__global__ void split_x16_kernel(const float *__restrict__ X, long long n4, uint2 *__restrict__ hi, uint2 *__restrict__ lo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 x = reinterpret_cast<const float4 *>(X)[i];
  uint2 h, l;
  split_pair_f16(x.x * kF16AScale, x.y * kF16AScale, h.x, l.x);
  split_pair_f16(x.z * kF16AScale, x.w * kF16AScale, h.y, l.y);
  hi[i] = h; lo[i] = l;
}
}  // namespace

void launch_joiner_records(SearchState *S, const SearchModel &m, const float *X, int rows, int kb, float *records, cudaStream_t st) {
  if (rows <= 0) return;
  if (!S->fused_partials) throw CudaError("this precision mode has no record epilogue (CUDA-core GEMM selects from full logits)");
  // the kernel a frame step launches: pre-split operands in the fp16-split mode (read at call time so a test can run both)
  const bool ss = !(getenv("B200ASR_JOINER_SS") && atoi(getenv("B200ASR_JOINER_SS")) == 0) && m.join_w16hi && m.join_w16lo && (m.jd & 7) == 0;
  if (ss) {
    const size_t n4 = (size_t)rows * (m.jd >> 2);
    ensure(S->X16, S->x16_cap, 2 * n4);
    split_x16_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(X, (long long)n4, S->X16, S->X16 + n4);
    count_launch(); KERNEL_CHECK();
  }
  GemmArgs ga{};
  ga.A = X; ga.lda = m.jd; ga.W = m.join_w; ga.Wlo = m.join_w_lo; ga.bias = m.join_b; ga.R = nullptr; ga.ldr = 0; ga.C = nullptr;
  ga.W16hi = m.join_w16hi; ga.W16lo = m.join_w16lo; ga.w16_ld = m.join_w16_ld;
  ga.ldc = m.V; ga.M = rows; ga.N = m.V; ga.K = m.jd; ga.act = ACT_JOINER; ga.partials = records; ga.part_kb = kb; ga.pdl = 0;
  if (ss) { ga.A = nullptr; ga.A16hi = S->X16; ga.A16lo = S->X16 + (size_t)rows * (m.jd >> 2); ga.a16_ld = m.jd; }
  S->gemm(ga, st);
}

void launch_context_preactivations(const SearchModel &m, long long ctx0, int rows, float *E, cudaStream_t st) {
  if (rows <= 0) return;
  const long long total = (long long)rows * (m.dd >> 2);
  context_pre_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(m, ctx0, rows, E);
  count_launch(); KERNEL_CHECK();
}

void launch_decoder_rows(const SearchModel &m, const long long *y, int rows, float *out, cudaStream_t st) {
  if (rows <= 0) return;
  decoder_rows_kernel<<<rows, 256, (size_t)m.dd * sizeof(float), st>>>(m, y, out);
  count_launch(); KERNEL_CHECK();
}

void launch_joiner_rows(const SearchModel &m, const float *enc, const float *dec, int rows, float *tmp, float *logits,
                        cudaStream_t st) {
  if (rows <= 0) return;
  const long long total = (long long)rows * m.jd;
  tanh_add_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(enc, dec, total, tmp);
  count_launch(); KERNEL_CHECK();
  GemmArgs ga{};
  ga.A = tmp; ga.lda = m.jd; ga.W = m.join_w; ga.bias = m.join_b; ga.C = logits; ga.ldc = m.V; ga.M = rows; ga.N = m.V;
  ga.K = m.jd; ga.act = ACT_NONE;
  launch_gemm_fp32(ga, st);
}

}  // namespace b200asr
