// Shared declarations for libb200asr.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdexcept>
#include <string>

namespace b200asr {

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define CUDA_CHECK(expr)                                                                              \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess)                                                                            \
      throw ::b200asr::CudaError(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " + \
                                 __FILE__ + ":" + std::to_string(__LINE__));                          \
  } while (0)

#define KERNEL_CHECK() CUDA_CHECK(cudaGetLastError())

// Programmatic dependent launch (sm_90+): a kernel launched with launch_pdl() may become resident while the previous
// kernel in the stream is still running; it must execute pdl_wait() before touching anything that kernel (or an
// earlier one) produces, and it lets its own successor start launching with pdl_trigger(). Both are no-ops for
// normal launches.
// The fp16 operand split of the FP32 mode (gemm_tc_f16.cu): x' = x * scale = hi + lo with hi = fp16(x'), lo = fp16(x' - hi);
// activations are scaled by 64, weights by 1024 (powers of two, undone in the epilogue).
constexpr float kF16AScale = 64.0f, kF16WScale = 1024.0f;
#ifdef __CUDACC__
}  // namespace b200asr
#include <cuda_fp16.h>
namespace b200asr {
__device__ __forceinline__ float clamp_f16(float x) { return fminf(fmaxf(x, -65504.0f), 65504.0f); }   // NaN stays NaN
// (x0, x1), already scaled -> packed fp16 hi pair and packed fp16 lo pair; element 0 in the low half-word
__device__ __forceinline__ void split_pair_f16(float x0, float x1, uint32_t &hi, uint32_t &lo) {
  // one F2FP.SATFINITE per pair (overflow -> +-65504, as clamp_f16 + round-to-nearest gives) instead of four FMNMX and an F2FP:
  // the converter warps of the GEMM are its busiest role, every instruction per element counts
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
  const float2 hf = __half22float2(*reinterpret_cast<const __half2 *>(&hi));
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - hf.y), "f"(x0 - hf.x));
}
// SFU forms of the activations' transcendentals (ex2 / lg2 / rcp .approx.ftz: one MUFU each, no denormal fix-up code around
// them). softplus: the 1 + t rounding bounds the absolute error at ~1e-7, fp32 noise on the O(1) Swoosh output; denormal
// exponentials flush to 0 where 1 + t = 1 anyway. sigmoid: 2 ulp. Measured in the kernels that use them: a third of the
// epilogue instructions of the Swoosh GEMMs and over half of glu_dwconv's were expf / log1pf / division sequences.
__device__ __forceinline__ float sfu_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sfu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sfu_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float softplus_sfu(float x) {
  return fmaf(sfu_lg2(1.0f + sfu_ex2(-fabsf(x) * 1.4426950408889634f)), 0.6931471805599453f, fmaxf(x, 0.f));
}
__device__ __forceinline__ float sigmoid_sfu(float x) { return sfu_rcp(1.0f + sfu_ex2(-x * 1.4426950408889634f)); }
// tanh = 1 - 2 / (1 + e^(2x)): absolute error ~1e-7 (saturates correctly: e^(2x) -> inf gives 1, -> 0 gives -1)
__device__ __forceinline__ float tanh_sfu(float x) { return fmaf(-2.0f, sfu_rcp(1.0f + sfu_ex2(x * 2.8853900817779268f)), 1.0f); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, KArgs(args)...));
}
#endif

// Opt-in dynamic shared memory: the attribute belongs to the (function, device) pair, so remember what has been set per
// device - a process may hold recognizers on several GPUs.
void set_max_dynamic_smem_impl(const void *func, int bytes);
template <typename F>
inline void set_max_dynamic_smem(F *func, size_t bytes) { set_max_dynamic_smem_impl(reinterpret_cast<const void *>(func), (int)bytes); }

// Persistent kernels (tcgen05 GEMM, attention) size their grid to the SM count. While the search of an earlier sub-batch runs
// beside the encoder of the next one (engine.cu, pipelined decode) the encoder's launches leave `reserve` SMs free for it:
// a persistent CTA holds its SM for the whole kernel, so without the reservation the search's small kernels would wait
// for encoder kernel boundaries. Thread-local (one decode call = one host thread).
void set_sm_reserve(int reserve);
int persistent_grid_limit(int n_sms);

// Global launch counter (host side) so bench.py can report `gpu_launches`.
extern long long g_launches;
inline void count_launch(int n = 1) { g_launches += n; }

// out = act(acc + bias) + R
// ACT_JOINER (tensor-core kernels only): out = acc + bias (the caller folds a blank penalty into the bias) and, per row and per
// 32-column part, a partial record {mx = max, S e^(x-mx), S e^(x-mx)(x-mx), S e^((x-mx)/3), top-kb values, top-kb
// columns}: enough for the log-softmax, the global top-k and the per-token entropy / Tsallis / margin statistics,
// so the beam-search selection never reads the logits (search.cu); C may be null to skip storing them at all.
enum Act : int { ACT_NONE = 0, ACT_SWOOSH_L = 1, ACT_SWOOSH_R = 2, ACT_JOINER = 4 };
constexpr int kPartCols = 32;                                        // logits columns per partial record
inline int part_rec_floats(int kb) { return 4 + 2 * kb; }            // floats per record

// ---------------------------------------------------------------- fbank (fbank.cu)
struct FbankTables {
  float *window;     // [400] povey
  float *twiddle;    // [512] interleaved cos,sin for k=0..255 of exp(-2*pi*i*k/512)
  int *mel_start;    // [80] first fft bin with non-zero weight
  int *mel_count;    // [80]
  int *mel_off;      // [80] offset into mel_w
  float *mel_w;      // packed weights
};
void fbank_tables_create(FbankTables *t);
void fbank_tables_destroy(FbankTables *t);
// samples + sample_off[u] = first sample of utterance u; its length is sample_len[u], or sample_off[u+1] - sample_off[u] when
// sample_len is null (packed PCM); frame_off[n+1]; out [total_frames, 80]
void launch_fbank(const FbankTables &t, const float *samples, const long long *sample_off, const long long *sample_len,
                  const long long *frame_off, int n_utts, int max_frames, float *out, cudaStream_t st);

// ---------------------------------------------------------------- GEMM (gemm.cu / gemm_tc.cu)
// C[M,N] = act(A[M,K] * W[N,K]^T + bias[N]) (+ R[M,N] if R). Row strides lda / ldc / ldr in elements.
struct GemmArgs {
  const float *A; int lda;
  const float *W;            // [N,K] row-major (ldw = K)
  const float *Wlo;          // W - trunc_tf32(W), needed by the 3xTF32 tensor-core kernel only (else null)
  const void *W16hi, *W16lo; // 16-bit operand copies made by split_weights_16 (gemm_tc_f16.cu), row pitch w16_ld; else null
  int w16_ld;
  const void *A16hi, *A16lo; // the activation already in the same scaled fp16 hi / lo form ([M, K] each, row pitch a16_ld): the search's
  int a16_ld, a16_rows;      // joiner input, written that way by the selection kernel (A may then be null; a16_rows = rows the planes hold, >= M, 0 = M); ACT_JOINER records only
  const float *bias;         // [N] or null
  const float *R; int ldr;   // residual or null
  float *C; int ldc;
  int M, N, K;
  int act;
  // ACT_JOINER only
  float *partials;           // [M, ceil(N/32), 4 + 2*part_kb]
  int part_kb;               // 4, 8 or 16 candidates kept per part
  unsigned long long *trace; // optional: CTA 0 stores %globaltimer at entry / exit (search step timeline, debug)
  int pdl;                   // launch with programmatic stream serialization (the kernel prologue overlaps the previous kernel)
  int *tile_counter;         // tensor-core kernels: a device int that is zero at launch -> tiles are claimed dynamically (null: static)
};
void launch_gemm_fp32(const GemmArgs &g, cudaStream_t st);
// gemm_tc_f16.cu: 16-bit tensor-core operands. split_weights_16 makes the copies of W[N, K] (caller frees hi / lo with cudaFree);
// launch_gemm_16 returns false for shapes it does not take. bf16 = single-pass BF16 mode, else the fp32-grade fp16 split.
void split_weights_16(const float *W, int N, int K, bool bf16, void **hi, void **lo, int *ld, cudaStream_t st);
bool launch_gemm_16(const GemmArgs &g, bool bf16, cudaStream_t st);
void launch_split_lo(const float *w, float *lo, long long n, cudaStream_t st);

// ---------------------------------------------------------------- encoder kernels (encoder.cu)
struct RaggedDesc {     // per-rate description of the packed batch, all device pointers
  const int *len;       // [n] frames per utterance at this rate
  const int *off;       // [n+1] row offsets
  int n;                // utterances
  int total;            // total rows
  int max_len;
};

// foff / ooff: [n+1] packed row offsets of the features / conv0 outputs; total_rows = ooff[n]
void launch_embed_conv0(const float *feats, const long long *foff, const long long *ooff, int n, long long total_rows, const float *w,
                        const float *b, float *out, cudaStream_t st);
// ioff / ooff: [n+1] packed row offsets of the conv0 / conv1 outputs; total_rows = ooff[n]
void launch_embed_conv1(const float *in, const long long *ioff, const long long *ooff, int n, long long total_rows, const float *w,
                        const float *b, float *out, cudaStream_t st);
// ioff: [n+1] packed row offsets of the conv1 outputs, ooff: [n+1] of the conv2 outputs; total_rows = ooff[n]
void launch_embed_im2col2(const float *in, const long long *ioff, const int *ooff, int n, long long total_rows, float *out,
                          cudaStream_t st);
// tile_off = cumulative ceil(len / 128) per utterance
void launch_embed_dw7(const float *in, const RaggedDesc &r, const int *tile_off, int n_tiles, const float *w, const float *b,
                      float *out, cudaStream_t st);
void launch_biasnorm(const float *x, int M, int D, const float *bias, const float *log_scale, float *out, cudaStream_t st);
// out = orig + (biasnorm(x) - orig) * bypass
void launch_biasnorm_bypass(const float *x, const float *orig, int M, int D, const float *bias, const float *log_scale,
                            const float *bypass, float *out, cudaStream_t st);
void launch_bypass(const float *x, const float *orig, long long M, int D, const float *scale, float *out, cudaStream_t st);
void launch_convert_channels(const float *in, int Cin, float *out, int Cout, long long M, cudaStream_t st);
// row maps between the full frame rate (r1) and a stack's rate (rq): up[rows at r1], down[rows at rq] (encoder.cu)
void launch_build_row_maps(const RaggedDesc &r1, const RaggedDesc &rq, int ds, int *up, int2 *down, cudaStream_t st);
void launch_downsample(const float *in, const int2 *down, int rows_out, int C, int ds, const float *bias, float *out, cudaStream_t st);
void launch_upsample_combine(const float *y, const int *up, const float *orig, int rows_full, int C, const float *scale, float *out,
                             cudaStream_t st);
struct ConcatPiece { const float *src; int ld; int c0; int c1; };
void launch_concat_downsample2(const ConcatPiece *pieces, int n_pieces, const int2 *down, int rows_out, int C, const float *bias,
                               float *out, cudaStream_t st);
void launch_pos_emb(float *pe, int max_len, int pos_dim, cudaStream_t st);  // [2*max_len-1, pos_dim]
// attention weights: proj [M, H*(2*qd+pd)], pos [2*max_len-1, H*pd] -> A packed per utterance: H*len*len at aoff[n]
void launch_attn_weights(const float *proj, int ldp, const float *pos, const RaggedDesc &r, const long long *aoff, int H, int qd,
                         int pd, float *A, cudaStream_t st);
// the same on the tensor pipe (attn_weights_tc.cu; query_head_dim 32, pos_head_dim 4): tile_off = cumulative
// H * ceil(len / 128) per utterance, A row pitch (len + 3) & ~3
bool attn_weights_tc_supported(int qd, int pd);
void launch_attn_weights_tc(const float *proj, int ldp, int m_total, const float *pos, const RaggedDesc &r, const long long *aoff,
                            const int *tile_off, int n_tiles, int H, float *A, bool split3, cudaStream_t st, int *tile_counter = nullptr,
                            float *Ls = nullptr, int *overflow = nullptr);   // Ls != null: single pass, unnormalised A + row sums [M, H]
// out[i, c] = (sum_j A[h(c)][i][j] * V[j,c]) (* Y[i,c]);  V = X (* tanh(S) if S). head = c / dv_per_head (0 if single_head)
void launch_attn_apply(const float *A, const long long *aoff, const RaggedDesc &r, const float *X, int ldx, const float *S, int lds,
                       const float *Y, int ldy, int C, int dv_per_head, int single_head, float *out, int ldo, cudaStream_t st);
// ---- attention application on the tensor pipe (attn_tc.cu); A rows have pitch Tk4 = (Tk + 3) & ~3
struct AttnTcLaunch {
  const void *mapsA, *mapsV, *mapsVlo;   // device arrays of CUtensorMap, one per utterance
  const int *tile_off;                   // [n_utt + 1]
  const int *len, *off;
  int n_utt, n_tiles;
  int single_head, C, dv;
  const float *Y; int ldy;
  float *out; int ldo;
  int split3;
  int *tile_counter;                     // dynamic tile scheduling: a device int that is zero at launch (null: static)
  const float *Ls; int H;                // row sums [M, H] of unnormalised weights (single-pass attn_weights) or null
};
// tile_off = cumulative ceil(len / 128) per utterance
void launch_transpose_v(const float *X, int ldx, const float *S, int lds, int C, const RaggedDesc &r, const int *tile_off, int n_tiles,
                        const long long *vt_off, float *VT, float *VTlo, cudaStream_t st);
void launch_attn_apply_tc(const AttnTcLaunch &a, cudaStream_t st);
void attn_tc_encode_maps(void *h_maps, int n, const float *base, const long long *elem_off, const int *len, int rows_mult,
                         int rows_fixed, int box_rows);
// conv module middle: h [M, 2D] -> out [M, D] = SwooshR(dwconv_k(x * sigmoid(s)) + b); tile_off = cumulative
// ceil(len / 128) per utterance
void launch_glu_dwconv(const float *h, const RaggedDesc &r, const int *tile_off, int n_tiles, int D, int k, const float *w,
                       const float *b, float *out, cudaStream_t st);

// ---------------------------------------------------------------- search kernels (search.cu)
struct ContextGraphDev {   // flattened Aho-Corasick automaton (BFS order; node 0 = root)
  int n_nodes = 0;
  int *edge_start = nullptr;   // [n_nodes+1]
  int *edge_token = nullptr;   // sorted by token within a node
  int *edge_child = nullptr;
  int *fail = nullptr;         // [n_nodes]
  int *token = nullptr;        // [n_nodes] (-1 root)
  int *is_end = nullptr;
  int *output = nullptr;       // node id or -1
  double *token_score = nullptr, *node_score = nullptr, *output_score = nullptr;
};

struct SearchModel {
  const float *emb;        // [V, dd]
  const float *conv_w;     // [dd, 4, ctx]
  const float *conv_p0;    // [V, dd] the grouped k=2 convolution applied to each token in context slot 0 ...
  const float *conv_p1;    // ... and slot 1: relu(conv(emb[y0], emb[y1])) = relu(conv_p0[y0] + conv_p1[y1])
  const float *dec_proj_w; // [jd, dd]
  const float *dec_proj_b;
  const float *dec_table;  // [V * V, jd] decoder output of every 2-token context (y0 * V + y1), built once by the engine; null = the
                           // search computes decoder_proj on demand (decoder_joinin_kernel)
  const float *join_w;     // [V, jd]
  const float *join_w_lo;  // low part for the 3xTF32 joiner GEMM (or null)
  const void *join_w16hi, *join_w16lo;   // 16-bit operand copies for the joiner GEMM (or null)
  int join_w16_ld;
  const float *join_b;
  int V, dd, jd;
  int blank_id, unk_id;
  double ts_max, max_ent;  // Tsallis (alpha = 1/3) and Shannon normalisers for V (core/asr_engine.py:1163-1165)
};

struct SearchState;   // opaque, search.cu
SearchState *search_state_create();
void search_state_destroy(SearchState *s);
// joiner GEMM implementation; fused_partials = it implements ACT_JOINER (the tcgen05 kernels do)
void search_set_gemm(SearchState *s, void (*fn)(const GemmArgs &, cudaStream_t), bool fused_partials);
// Runs the whole search for a batch. enc [sum T', jd] packed with enc_off; results to device arrays then host.
struct SearchResultHost {
  int n_utts;
  int max_tokens;
  int *n_tokens;      // [n]
  int *tokens;        // [n, max_tokens]
  int *frames;
  float *tok_lp;
  float *stats;       // [n, max_tokens, 4]
};
void run_search(SearchState *s, const SearchModel &m, const ContextGraphDev *g, const float *enc, const int *h_lens, int n_utts,
                int method, int beam, float blank_penalty, SearchResultHost *out, cudaStream_t st);
// The same in two halves, for callers that overlap several searches (engine.cu): search_issue enqueues everything on `st`
// (device->host result copies included) without synchronising; once `st` has been synchronised the results are read through
// search_view (packed per utterance: off[u] = sum of T' of the utterances before u) or scattered with search_collect.
void search_issue(SearchState *s, const SearchModel &m, const ContextGraphDev *g, const float *enc, const int *h_lens, int n_utts,
                  int method, int beam, float blank_penalty, cudaStream_t st);
struct SearchView {
  int n;
  const int *n_tokens;      // [n]
  const long long *off;     // [n+1]
  const int *tokens, *frames;
  const float *tok_lp, *stats;   // stats [slots][4]
};
SearchView search_view(const SearchState *s);
void search_collect(SearchState *s, SearchResultHost *out);
void search_print_prof(SearchState *s);   // B200ASR_SEARCH_PROF=1: phase counters / step timeline of the last issued search, to stderr
long long search_result_bytes(const SearchState *s);   // device->host bytes the last issued search copies
void launch_decoder_product_rows(const SearchModel &m, const long long *y, const float *enc, int rows, float *dec_out, float *x_out,
                                 cudaStream_t st);
void launch_joiner_records(SearchState *s, const SearchModel &m, const float *X, int rows, int kb, float *records, cudaStream_t st);
void launch_decoder_rows(const SearchModel &m, const long long *y, int rows, float *out, cudaStream_t st);
// E[r] = relu(conv_p0[y0] + conv_p1[y1]) for the contexts ctx0 + r = y0 * V + y1 (the decoder-table build, engine.cu)
void launch_context_preactivations(const SearchModel &m, long long ctx0, int rows, float *E, cudaStream_t st);
void launch_joiner_rows(const SearchModel &m, const float *enc, const float *dec, int rows, float *tmp, float *logits,
                        cudaStream_t st);

}  // namespace b200asr
