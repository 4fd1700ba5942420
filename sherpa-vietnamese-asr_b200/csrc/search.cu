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
// All utterances of a batch advance together: per frame step three launches
//   decoder_step  (stateless decoder for new contexts / copy for blank extensions, then tanh(enc_t + dec))
//   joiner GEMM   (gemm.cu / gemm_tc.cu: [active*beam, jd] x [jd, V])
//   select_step   (one CTA per utterance: log-softmax, top-k, expansion, hotword arcs, dedup, token statistics)
// Hypotheses live in HBM as back-pointer chains in a per-utterance arena; nothing returns to the host until
// the last frame. Utterances are processed longest-first so the active set at step t is a prefix.
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
  float *E;                    // [n*beam, dd] relu(conv(embeddings)) of each live hypothesis (decoder GEMM input)
  int *rowmap;                 // [n*beam] encoder_out row joined with that hypothesis at the next step
  float *X;                    // [n*beam, jd] joiner input tanh(enc+dec)
  float *logits;               // [n*beam, V]
  int n, beam;
  long long *prof;             // optional per-phase cycle counters of CTA 0 (B200ASR_SEARCH_PROF=1)
};

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

#define SEL_PROF(k)                                                               \
  do {                                                                            \
    if (d.prof && blockIdx.x == 0 && threadIdx.x == 0) {                          \
      const long long _now = clock64();                                           \
      atomicAdd((unsigned long long *)&d.prof[k], (unsigned long long)(_now - _t0)); \
      _t0 = _now;                                                                 \
    }                                                                             \
  } while (0)

template <int KB>
__global__ void __launch_bounds__(kSelThreads) select_step_kernel(SearchModel m, SearchDev d, ContextGraphView g, int has_graph,
                                                                  int t, int cur, int greedy, float blank_penalty) {
  const int s = blockIdx.x;
  if (t >= d.lens[s]) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int V = m.V;
  __shared__ HypSlot s_old[kMaxBeam], s_new[kMaxBeam];
  __shared__ float s_mx[kMaxBeam], s_lse[kMaxBeam], s_sum[kMaxBeam], s_prev[kMaxBeam];
  __shared__ unsigned long long s_warpbest[kSelThreads / 32];
  __shared__ unsigned long long s_win[kMaxBeam];
  __shared__ NewNode s_nodes[kMaxBeam];
  __shared__ int s_n_new, s_n_nodes;

  long long _t0 = clock64();
  const int count = d.hyp_count[cur * d.n + s];
  if (tid < count) s_old[tid] = d.hyps[((long long)cur * d.n + s) * kMaxBeam + tid];
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
      s_mx[b] = mx;
      s_sum[b] = sum;
      s_lse[b] = logf(sum);
      s_prev[b] = greedy ? 0.f : (float)s_old[b].score;   // score rounded to f32 before the add (:1099-1100)
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
    const float mxb = s_mx[b], lseb = s_lse[b], prevb = s_prev[b];
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
  int head = 0;
  for (int round = 0; round < k; ++round) {
    unsigned long long best = 0ULL;
#pragma unroll
    for (int i = 0; i < KB; ++i) if (i == head) best = loc[i];
    unsigned long long wb = best;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, wb, o);
      wb = other > wb ? other : wb;
    }
    if (lane == 0) s_warpbest[warp] = wb;
    __syncthreads();
    unsigned long long win = 0ULL;
#pragma unroll
    for (int w = 0; w < kSelThreads / 32; ++w) win = s_warpbest[w] > win ? s_warpbest[w] : win;
    if (best == win && win != 0ULL) ++head;
    if (tid == 0) s_win[round] = win;
    __syncthreads();
  }

  SEL_PROF(2);
  // (c) expansion, hotword arcs, dedup: serial over <= beam winners, exactly in top-k order (:1109-1140)
  ArenaNode *arena = d.arena + d.arena_off[s];
  if (tid == 0) {
    int n_new = 0, n_nodes = 0;
    int node_count = d.node_count[s];
    for (int r = 0; r < k; ++r) {
      const unsigned long long key = s_win[r];
      const int idx = (int)(~(unsigned)(key & 0xffffffffULL));
      const float lp32 = unord_f32((unsigned)(key >> 32));
      const int hi = idx / V, tok = idx - hi * V;
      const HypSlot &p = s_old[hi];
      HypSlot c;
      c.score = (double)lp32;
      bool is_blank = (tok == m.blank_id);
      if (is_blank) {
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
      // dedup against already inserted hypotheses (same token sequence)
      int dup = -1;
      for (int q = 0; q < n_new && dup < 0; ++q) {
        if (s_new[q].len != c.len || s_new[q].hash != c.hash) continue;
        if (is_blank) {
          if (same_chain(arena, s_new[q].node, p.node)) dup = q;
        } else {
          const int qn = s_new[q].node;
          if (qn >= 0 && arena[qn].token == tok && same_chain(arena, arena[qn].parent, p.node)) dup = q;
        }
      }
      if (dup >= 0) {
        s_new[dup].score = log_add_d(s_new[dup].score, c.score);
        continue;
      }
      if (!is_blank) {
        const int ai = node_count++;
        ArenaNode nd;
        nd.parent = p.node; nd.token = tok; nd.frame = t;
        nd.tok_lp = (float)((double)lp32 - (greedy ? 0.0 : p.score));   // (:1121)
        nd.stats[0] = nd.stats[1] = nd.stats[2] = nd.stats[3] = 0.f;
        arena[ai] = nd;
        c.node = ai;
        s_nodes[n_nodes].arena_idx = ai;
        s_nodes[n_nodes].row = hi;
        ++n_nodes;
      }
      s_new[n_new++] = c;
    }
    s_n_new = n_new;
    s_n_nodes = n_nodes;
    d.node_count[s] = node_count;
    d.hyp_count[(cur ^ 1) * d.n + s] = n_new;
  }
  __syncthreads();

  SEL_PROF(3);
  // (d) per emitted token statistics from its logits row (_compute_token_entropy :1159-1181), one warp per token
  for (int e = warp; e < s_n_nodes; e += kSelThreads / 32) {
    const int b = s_nodes[e].row;
    const float *row = lg + (long long)b * V;
    const float mx = s_mx[b], sum = s_sum[b];
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
      ArenaNode &nd = arena[s_nodes[e].arena_idx];
      nd.stats[0] = (float)(ts_max > 0 ? tsallis / ts_max : 0.0);
      nd.stats[1] = t1 - (V > 1 ? t2 : 1e-10f);
      nd.stats[2] = (float)((-(double)ent) / max_ent);
      nd.stats[3] = t1;
    }
  }
  if (tid < s_n_new) d.hyps[((long long)(cur ^ 1) * d.n + s) * kMaxBeam + tid] = s_new[tid];
  SEL_PROF(4);

  // (e) decoder pre-activation for the new hypotheses: E = relu(grouped conv over the two context embeddings).
  // decoder_proj and the joiner's tanh(enc + dec) then run as ONE tensor-core GEMM over all live hypotheses of
  // the batch (run_search), so the 1 MB projection is streamed once per step instead of once per utterance.
  const int n_new = s_n_new;
  const int next_row = (int)d.enc_off[s] + min(t + 1, d.lens[s] - 1);
  for (int i = tid; i < n_new * m.dd; i += kSelThreads) {
    const int q = i / m.dd, o = i - q * m.dd;
    const HypSlot &hs = s_new[q];
    const int g4 = (o >> 2) << 2;
    const float *w = m.conv_w + (long long)o * 8;
    const float *e0 = m.emb + (long long)hs.y0 * m.dd + g4;
    const float *e1 = m.emb + (long long)hs.y1 * m.dd + g4;
    float a2 = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      a2 = fmaf(__ldg(w + k * 2 + 0), __ldg(e0 + k), a2);
      a2 = fmaf(__ldg(w + k * 2 + 1), __ldg(e1 + k), a2);
    }
    d.E[((long long)s * d.beam + q) * m.dd + o] = fmaxf(a2, 0.f);
  }
  if (tid < d.beam) d.rowmap[s * d.beam + tid] = next_row;
  SEL_PROF(7);
}

__global__ void init_search_kernel(SearchModel m, SearchDev d) {
  const int s = blockIdx.x;
  if (threadIdx.x == 0) {
    HypSlot h;
    h.score = 0.0; h.hash = 0x12345ULL; h.node = -1; h.len = 0; h.ctx = 0; h.y0 = 0; h.y1 = 0; h.dec_src = -1;
    d.hyps[(long long)s * kMaxBeam] = h;     // ys = [-1, 0] -> decoder input [0, 0] (:1051-1052)
    d.hyp_count[s] = 1;
    d.hyp_count[d.n + s] = 0;
    d.node_count[s] = 0;
  }
  if ((int)threadIdx.x < d.beam) d.rowmap[s * d.beam + threadIdx.x] = (int)d.enc_off[s];   // frame 0 (any valid row if T' = 0)
  for (int i = threadIdx.x; i < d.beam * m.dd; i += blockDim.x) {
    const int q = i / m.dd, o = i - q * m.dd;
    float v = 0.f;
    if (q == 0) {
      const int g4 = (o >> 2) << 2;
      const float *w = m.conv_w + (long long)o * 8;
      const float *e0 = m.emb + g4;          // token 0 twice
      float a2 = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        a2 = fmaf(__ldg(w + k * 2 + 0), __ldg(e0 + k), a2);
        a2 = fmaf(__ldg(w + k * 2 + 1), __ldg(e0 + k), a2);
      }
      v = fmaxf(a2, 0.f);
    }
    d.E[((long long)s * d.beam + q) * m.dd + o] = v;
  }
}

// finalize (:1143-1153): subtract unfinished hotword score, pick first max of log_prob/len(ys), unroll the chain
__global__ void finalize_kernel(SearchDev d, ContextGraphView g, int has_graph, const int *__restrict__ final_buf,
                                const int *__restrict__ orig_index, int max_tokens, int *__restrict__ n_tokens,
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
  int len = hy[best].len;
  n_tokens[u] = len;
  int node = hy[best].node;
  for (int i = len - 1; i >= 0 && node >= 0; --i) {
    const ArenaNode &nd = arena[node];
    if (i < max_tokens) {
      const long long o = (long long)u * max_tokens + i;
      tokens[o] = nd.token; frames[o] = nd.frame; tok_lp[o] = nd.tok_lp;
      stats[o * 4 + 0] = nd.stats[0]; stats[o * 4 + 1] = nd.stats[1];
      stats[o * 4 + 2] = nd.stats[2]; stats[o * 4 + 3] = nd.stats[3];
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
  int *rowmap = nullptr; size_t rm_cap = 0;
  float *X = nullptr; size_t x_cap = 0;
  float *logits = nullptr; size_t lg_cap = 0;
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
  void (*gemm)(const GemmArgs &, cudaStream_t) = launch_gemm_fp32;
};

SearchState *search_state_create() { return new SearchState(); }
void search_set_gemm(SearchState *s, void (*fn)(const GemmArgs &, cudaStream_t)) { s->gemm = fn; }
void search_state_destroy(SearchState *s) {
  if (!s) return;
  cudaFree(s->hyps); cudaFree(s->hyp_count); cudaFree(s->node_count); cudaFree(s->arena); cudaFree(s->E); cudaFree(s->rowmap);
  cudaFree(s->X); cudaFree(s->logits); cudaFree(s->enc_off); cudaFree(s->arena_off); cudaFree(s->lens);
  cudaFree(s->orig); cudaFree(s->final_buf); cudaFree(s->o_ntok); cudaFree(s->o_tok); cudaFree(s->o_frm);
  cudaFree(s->o_lp); cudaFree(s->o_st);
  delete s;
}

void run_search(SearchState *S, const SearchModel &m, const ContextGraphDev *g, const float *enc, const int *h_lens, int n,
                int method, int beam, float blank_penalty, SearchResultHost *out, cudaStream_t st) {
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
  const size_t rows = (size_t)n * beam;
  ensure(S->hyps, S->hyps_cap, 2 * (size_t)n * kMaxBeam);
  ensure(S->hyp_count, S->hc_cap, 2 * (size_t)n);
  ensure(S->node_count, S->nc_cap, (size_t)n);
  ensure(S->arena, S->arena_cap, (size_t)std::max<long long>(arena_total, 1));
  ensure(S->E, S->e_cap, rows * m.dd);
  ensure(S->rowmap, S->rm_cap, rows);
  ensure(S->X, S->x_cap, rows * m.jd);
  ensure(S->logits, S->lg_cap, rows * m.V);
  ensure(S->enc_off, S->eo_cap, (size_t)n);
  ensure(S->arena_off, S->ao_cap, (size_t)n);
  ensure(S->lens, S->ln_cap, (size_t)n);
  ensure(S->orig, S->og_cap, (size_t)n);
  ensure(S->final_buf, S->fb_cap, (size_t)n);
  const int max_tokens = out->max_tokens;
  ensure(S->o_ntok, S->ont_cap, (size_t)n);
  ensure(S->o_tok, S->ot_cap, (size_t)n * max_tokens);
  ensure(S->o_frm, S->of_cap, (size_t)n * max_tokens);
  ensure(S->o_lp, S->ol_cap, (size_t)n * max_tokens);
  ensure(S->o_st, S->os_cap, (size_t)n * max_tokens * 4);
  CUDA_CHECK(cudaMemcpyAsync(S->enc_off, enc_off.data(), n * sizeof(long long), cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(S->arena_off, arena_off.data(), n * sizeof(long long), cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(S->lens, lens.data(), n * sizeof(int), cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(S->orig, order.data(), n * sizeof(int), cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(S->final_buf, final_buf.data(), n * sizeof(int), cudaMemcpyHostToDevice, st));

  SearchDev d;
  d.enc = enc; d.enc_off = S->enc_off; d.lens = S->lens; d.arena_off = S->arena_off; d.hyps = S->hyps;
  d.hyp_count = S->hyp_count; d.node_count = S->node_count; d.arena = S->arena; d.E = S->E; d.rowmap = S->rowmap; d.X = S->X;
  d.logits = S->logits; d.n = n; d.beam = beam;
  d.prof = nullptr;
  static const bool want_prof = getenv("B200ASR_SEARCH_PROF") != nullptr;
  long long *d_prof = nullptr;
  if (want_prof) {
    CUDA_CHECK(cudaMalloc(&d_prof, 8 * sizeof(long long)));
    CUDA_CHECK(cudaMemsetAsync(d_prof, 0, 8 * sizeof(long long), st));
    d.prof = d_prof;
  }
  ContextGraphView gv{};
  const int has_graph = (g && g->n_nodes > 1 && !greedy) ? 1 : 0;
  if (has_graph)
    gv = ContextGraphView{g->n_nodes, g->edge_start, g->edge_token, g->edge_child, g->fail, g->token, g->is_end, g->output,
                          g->token_score, g->node_score, g->output_score};

  if (m.V & 3) throw CudaError("beam search: vocab_size must be a multiple of 4");
  {
    static bool attr_set = false;
    if (!attr_set) {
      CUDA_CHECK(cudaFuncSetAttribute(select_step_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      CUDA_CHECK(cudaFuncSetAttribute(select_step_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      CUDA_CHECK(cudaFuncSetAttribute(select_step_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_set = true;
    }
    if ((size_t)beam * m.V * sizeof(float) > 200 * 1024) throw CudaError("beam * vocab_size too large for the selection kernel");
  }
  init_search_kernel<<<n, 128, 0, st>>>(m, d);
  count_launch(); KERNEL_CHECK();
  int n_active = n;
  for (int t = 0; t < max_len; ++t) {
    while (n_active > 0 && lens[n_active - 1] <= t) --n_active;
    const int cur = t & 1;
    // decoder_proj + joiner input in one GEMM: X = tanh(E * Wp^T + bp + enc[rowmap])
    GemmArgs gd{};
    gd.A = S->E; gd.lda = m.dd; gd.W = m.dec_proj_w; gd.Wlo = m.dec_proj_w_lo; gd.bias = m.dec_proj_b; gd.R = enc; gd.ldr = m.jd;
    gd.r_rows = S->rowmap; gd.C = S->X; gd.ldc = m.jd; gd.M = n_active * beam; gd.N = m.jd; gd.K = m.dd; gd.act = ACT_TANH_RES;
    S->gemm(gd, st);
    // joiner output_linear: logits = X * Wj^T + bj
    GemmArgs ga{};
    ga.A = S->X; ga.lda = m.jd; ga.W = m.join_w; ga.Wlo = m.join_w_lo; ga.bias = m.join_b; ga.R = nullptr; ga.ldr = 0; ga.C = S->logits;
    ga.ldc = m.V; ga.M = n_active * beam; ga.N = m.V; ga.K = m.jd; ga.act = ACT_NONE;
    S->gemm(ga, st);
    const size_t sel_smem = (size_t)beam * m.V * sizeof(float);
    if (beam <= 4) select_step_kernel<4><<<n_active, kSelThreads, sel_smem, st>>>(m, d, gv, has_graph, t, cur, greedy, blank_penalty);
    else if (beam <= 8) select_step_kernel<8><<<n_active, kSelThreads, sel_smem, st>>>(m, d, gv, has_graph, t, cur, greedy, blank_penalty);
    else select_step_kernel<16><<<n_active, kSelThreads, sel_smem, st>>>(m, d, gv, has_graph, t, cur, greedy, blank_penalty);
    count_launch();
  }
  KERNEL_CHECK();
  // Utterances that stopped at step T' hold their final state in buffer (T' & 1): select writes to cur^1 = (t+1)&1.
  finalize_kernel<<<(n + 127) / 128, 128, 0, st>>>(d, gv, has_graph, S->final_buf, S->orig, max_tokens, S->o_ntok, S->o_tok,
                                                   S->o_frm, S->o_lp, S->o_st);
  count_launch(); KERNEL_CHECK();
  CUDA_CHECK(cudaMemcpyAsync(out->n_tokens, S->o_ntok, n * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (out->tokens) {
    CUDA_CHECK(cudaMemcpyAsync(out->tokens, S->o_tok, (size_t)n * max_tokens * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(out->frames, S->o_frm, (size_t)n * max_tokens * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(out->tok_lp, S->o_lp, (size_t)n * max_tokens * sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(out->stats, S->o_st, (size_t)n * max_tokens * 4 * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  CUDA_CHECK(cudaStreamSynchronize(st));
  if (d_prof) {
    long long h[8];
    CUDA_CHECK(cudaMemcpy(h, d_prof, sizeof h, cudaMemcpyDeviceToHost));
    cudaFree(d_prof);
    const char *names[8] = {"stage logits", "logsumexp", "top-k", "expand/dedup", "stats+writeback", "-", "-", "decoder pre-activation"};
    fprintf(stderr, "[b200asr search prof] CTA0 cycles over %d steps:", max_len);
    for (int i = 0; i < 8; ++i) fprintf(stderr, " %s=%.1f/step", names[i], (double)h[i] / std::max(max_len, 1));
    fprintf(stderr, "\n");
  }
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
