/* b200asr.h — C-ABI of libb200asr.so: the B200-native offline Zipformer RNN-T transcription path.
 *
 * Every entry point replaces one call the reference makes into a third-party CPU library; names mirror
 * sherpa-onnx's C API as bound by the reference's vendored wrapper
 *   /root/reference offline_pwa/static/vendor/sherpa-onnx-wasm/sherpa-onnx-asr.js
 * and the Python surface used at
 *   /root/reference streaming_asr.py:224-243,285,308-312,355-359,408-409
 *   /root/reference core/audio_analyzer.py:333-367, web_service/audio_quality.py:177-209,268-292
 * plus raw stage entry points that stand where core/asr_engine.py calls kaldi-native-fbank and
 * onnxruntime (core/asr_engine.py:698-721,1045-1056,1084-1093) and runs `_ort_beam_search` (:1023-1153).
 *
 * Plain C: pointers and sizes only, no exceptions cross the boundary. Functions returning int return
 * 0 on success, non-zero on failure; B200AsrGetLastError() gives the thread-local message.
 * There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef B200ASR_H_
#define B200ASR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200ASR_API __attribute__((visibility("default")))

/* ---- config PODs (layout follows SherpaOnnxOfflineRecognizerConfig as the wrapper packs it,
 *      sherpa-onnx-asr.js:1678-1782: feat, model, decoding method, max active paths, hotwords) ---- */
typedef struct B200AsrFeatureConfig {
  int32_t sample_rate;   /* 16000 (core/asr_engine.py:706) */
  int32_t feature_dim;   /* 80    (core/asr_engine.py:710) */
} B200AsrFeatureConfig;

typedef struct B200AsrOfflineTransducerModelConfig {
  const char *encoder;   /* `.b200w` container holding encoder.* tensors (stands for encoder-*.onnx) */
  const char *decoder;   /* container with decoder.*  */
  const char *joiner;    /* container with joiner.*   */
} B200AsrOfflineTransducerModelConfig;

typedef struct B200AsrOfflineModelConfig {
  B200AsrOfflineTransducerModelConfig transducer;
  const char *tokens;        /* tokens.txt: "<piece> <id>" per line (core/asr_engine.py:981-987) */
  int32_t num_threads;       /* accepted for source compatibility; ignored */
  int32_t debug;
  const char *provider;      /* must be "cuda" or "" ; anything else is an error (no CPU provider) */
  const char *model_type;
  const char *modeling_unit; /* "bpe" | "cjkchar" | "token_id" */
  const char *bpe_vocab;
} B200AsrOfflineModelConfig;

typedef struct B200AsrOfflineRecognizerConfig {
  B200AsrFeatureConfig feat_config;
  B200AsrOfflineModelConfig model_config;
  const char *decoding_method; /* "greedy_search" | "modified_beam_search" */
  int32_t max_active_paths;    /* beam; reference default 8 (core/asr_engine.py:903), sherpa default 4 */
  const char *hotwords_file;   /* optional; lines of space separated token ids [" :score"] when
                                  modeling_unit == "token_id"; text phrases are tokenised by the host
                                  binding and passed through B200AsrSetHotwordsTokenIds */
  float hotwords_score;        /* default boost 1.5 (core/config.py:405-412) */
  float blank_penalty;
  int32_t device_id;           /* CUDA device ordinal */
  int32_t precision;           /* 0 = FP32 mode (token-exact): tcgen05 with error-compensated operand splits (fp16 hi + lo x 3 MMAs;
                                      3xTF32 for attention and for shapes the fp16 kernel does not take), fp32-grade products;
                                  1 = tcgen05, TF32 operands in one pass, FP32 accumulate;
                                  2 = FP32 on CUDA cores (plain FFMA kernels, the cross-check of mode 0);
                                  3 = BF16 mode: bf16 weights, activations rounded to bf16 at every Linear, FP32 accumulate */
} B200AsrOfflineRecognizerConfig;

typedef struct B200AsrOfflineRecognizer B200AsrOfflineRecognizer;
typedef struct B200AsrOfflineStream B200AsrOfflineStream;

/* Result of one stream (SherpaOnnxOfflineRecognizerResult + the per-token statistics the reference
 * derives from the emitted logits rows, core/asr_engine.py:1159-1181). Library-owned. */
typedef struct B200AsrOfflineRecognizerResult {
  const char *text;            /* pieces joined, U+2581 -> space, stripped */
  const char *json;            /* {"text","tokens","timestamps","ys_log_probs","lang","emotion","event"} */
  const char *const *tokens;   /* count piece strings */
  const int32_t *token_ids;    /* count */
  const float *timestamps;     /* seconds, frame * duration / T' (core/asr_engine.py:1237) */
  const int32_t *frames;       /* encoder frame index per token */
  const float *ys_log_probs;   /* per-token log-prob (core/asr_engine.py:1121) */
  const float *tsallis;        /* normalised Tsallis entropy (alpha = 1/3), unrounded */
  const float *margin;         /* top1 - top2 probability */
  const float *entropy;        /* Shannon entropy / ln V */
  const float *top1;           /* top-1 probability */
  int32_t count;
  int32_t num_frames;          /* T' */
  float duration;              /* seconds of audio in the stream */
} B200AsrOfflineRecognizerResult;

/* ---- recognizer surface ---- */
/* SherpaOnnxCreateOfflineRecognizer (sherpa-onnx-asr.js:1843-1851); OfflineRecognizer.from_transducer
 * (streaming_asr.py:243). Strings are borrowed for the call only. NULL on failure. */
B200ASR_API const B200AsrOfflineRecognizer *B200AsrCreateOfflineRecognizer(const B200AsrOfflineRecognizerConfig *config);
/* SherpaOnnxDestroyOfflineRecognizer (sherpa-onnx-asr.js:1859-1862) */
B200ASR_API void B200AsrDestroyOfflineRecognizer(const B200AsrOfflineRecognizer *r);
/* SherpaOnnxOfflineRecognizerSetConfig (sherpa-onnx-asr.js:1853-1857): decoding method, beam, hotword score, blank penalty.
 * Fields left at "" / 0 / NaN (blank_penalty) keep their current value. */
B200ASR_API int32_t B200AsrOfflineRecognizerSetConfig(const B200AsrOfflineRecognizer *r, const B200AsrOfflineRecognizerConfig *config);
/* Hotword phrases as token ids (what build_context_graph produces after SentencePiece,
 * core/hotword_context.py:222-259): phrase p = tokens[offsets[p] .. offsets[p+1]). n_phrases == 0 clears. */
B200ASR_API int32_t B200AsrSetHotwordsTokenIds(const B200AsrOfflineRecognizer *r, const int32_t *tokens,
                                               const int32_t *offsets, const float *scores, int32_t n_phrases);
/* SherpaOnnxCreateOfflineStream (sherpa-onnx-asr.js:1864-1867; streaming_asr.py:308) */
B200ASR_API const B200AsrOfflineStream *B200AsrCreateOfflineStream(const B200AsrOfflineRecognizer *r);
/* SherpaOnnxCreateOfflineStreamWithHotwords (upstream C API; Python create_stream(hotwords=...)): a stream with its own
 * hotword automaton instead of the recognizer's. `hotwords` = phrases separated by '/', each a list of space-separated token
 * ids with an optional " :score" (text phrases are tokenised by the host binding, as build_context_graph does with
 * SentencePiece, core/hotword_context.py:222-259). Streams with different automata are decoded in separate passes. */
B200ASR_API const B200AsrOfflineStream *B200AsrCreateOfflineStreamWithHotwords(const B200AsrOfflineRecognizer *r, const char *hotwords);
B200ASR_API void B200AsrDestroyOfflineStream(const B200AsrOfflineStream *s);
/* SherpaOnnxAcceptWaveformOffline (sherpa-onnx-asr.js:1798-1806; streaming_asr.py:285,312,355): copies the
 * samples ([-1,1] floats) and appends on repeated calls. The copy lands in pinned host memory and its upload to the
 * recognizer's GPU is queued at once, so a later decode call finds the PCM resident. sherpa-onnx's function returns void;
 * this one returns 0 / -1 (wrong sample rate, host memory exhausted: B200AsrGetLastError) so a failed copy cannot pass
 * silently - a binding written against the void signature can ignore the value. */
B200ASR_API int32_t B200AsrAcceptWaveformOffline(const B200AsrOfflineStream *s, int32_t sample_rate, const float *samples, int32_t n);
/* The same for n streams in one call (stream i takes samples[i][0 .. ns[i])): the accept loop of a batch without n trips
 * through the host binding. Returns 0, or -1 at the first stream that fails. */
B200ASR_API int32_t B200AsrAcceptWaveformsOffline(const B200AsrOfflineStream *const *ss, int32_t sample_rate,
                                                  const float *const *samples, const int32_t *ns, int32_t n);
/* Precomputed features in place of samples: what decode_chunk's `precomputed_features` argument carries when ROVER computes
 * one fbank for both models (core/asr_engine.py:1209-1216,2346-2350). feats[num_frames * 80] as B200AsrFbank returns them for
 * num_samples samples (num_frames == (num_samples + 80) / 160; num_samples also gives the result's duration and timestamps).
 * A stream holds either samples or features. Returns 0 / -1. */
B200ASR_API int32_t B200AsrAcceptFeaturesOffline(const B200AsrOfflineStream *s, const float *feats, int32_t num_frames,
                                                 int32_t feature_dim, int64_t num_samples);
/* SherpaOnnxDecodeOfflineStream (sherpa-onnx-asr.js:1869-1871; streaming_asr.py:358,408) */
B200ASR_API int32_t B200AsrDecodeOfflineStream(const B200AsrOfflineRecognizer *r, const B200AsrOfflineStream *s);
/* SherpaOnnxDecodeMultipleOfflineStreams (upstream C API; Python decode_streams): the batch entry point.
 * Ragged batch, per-utterance unpadded semantics. Returns after results are on the host. */
B200ASR_API int32_t B200AsrDecodeMultipleOfflineStreams(const B200AsrOfflineRecognizer *r, const B200AsrOfflineStream *const *ss, int32_t n);
/* SherpaOnnxGetOfflineStreamResult / ...AsJson (sherpa-onnx-asr.js:1873-1880). The struct stays valid until
 * the stream is decoded again or destroyed; B200AsrDestroyOfflineRecognizerResult is a no-op kept for
 * source compatibility; the JSON string is malloc'ed and must be released with ...ResultJson. */
B200ASR_API const B200AsrOfflineRecognizerResult *B200AsrGetOfflineStreamResult(const B200AsrOfflineStream *s);
B200ASR_API void B200AsrDestroyOfflineRecognizerResult(const B200AsrOfflineRecognizerResult *r);
B200ASR_API const char *B200AsrGetOfflineStreamResultAsJson(const B200AsrOfflineStream *s);
B200ASR_API void B200AsrDestroyOfflineStreamResultJson(const char *s);

B200ASR_API const char *B200AsrGetLastError(void);
B200ASR_API const char *B200AsrVersion(void);
B200ASR_API int32_t B200AsrVocabSize(const B200AsrOfflineRecognizer *r);
B200ASR_API int32_t B200AsrEncoderOutDim(const B200AsrOfflineRecognizer *r);

/* ---- raw stage entry points (host pointers; used by parity tests, ncu and the asr_engine-style sessions) ---- */
/* compute_fbank_ort (core/asr_engine.py:698-721): samples[n] -> out[T*80], T = (n+80)/160. `out` may be NULL
 * to query T. Returns T or <0. */
B200ASR_API int32_t B200AsrFbank(const B200AsrOfflineRecognizer *r, const float *samples, int32_t n, float *out);
/* Ragged batch variant: utterance u = samples[sample_offsets[u] .. sample_offsets[u+1]); features packed,
 * frame_offsets[n_utts+1] written. */
B200ASR_API int32_t B200AsrFbankBatch(const B200AsrOfflineRecognizer *r, const float *samples, const int64_t *sample_offsets,
                                      int32_t n_utts, float *out, int64_t *frame_offsets);
/* Energy scan of find_silent_regions (core/asr_engine.py:521-553): quiet[f] = 1 when the RMS of the 10 ms frame
 * samples[160 f .. 160 f + 160) is below `threshold`, computed in NumPy's float32 arithmetic (same flags as the reference,
 * bit for bit). Needs no recognizer. `quiet` may be NULL to query the frame count n / 160. Returns the frame count or <0. */
B200ASR_API int32_t B200AsrSilentFrames(const float *samples, int64_t n, int32_t sample_rate, float threshold, uint8_t *quiet,
                                        int32_t device_id);
/* ---- audio staging ----
 * preprocess_audio (core/audio_preprocessing.py:251-292) over the uploaded PCM: optional per-segment RMS normalisation
 * (:46-155: gain = median segment RMS / segment RMS clamped to +-20 dB, 5 ms linear fades at segment edges) then the peak
 * limiter (:226-244: peak > 0.95 -> scaled to 0.95); `boost_low_volume` != 0 first applies the load step's boost
 * (core/asr_engine.py:512-516: 0 < peak < 0.5 -> audio / peak * 0.95). Segments are [start, end) sample ranges of the VAD.
 * samples[n] -> out[n] (host pointers; may alias). Bit-equal to the NumPy code except that segment RMS values are
 * accumulated in float64 (gains agree to ~1e-7 relative). Needs no recognizer. Returns 0 / -1. */
B200ASR_API int32_t B200AsrPreprocessAudio(const float *samples, int64_t n, const int64_t *seg_starts, const int64_t *seg_ends,
                                           int32_t n_seg, int32_t enable_rms_normalize, int32_t boost_low_volume, int32_t sample_rate,
                                           float *out, int32_t device_id);
/* ---- voice-activity network (the step before the recognizer) ----
 * Stands where the reference runs the Silero VAD ONNX session once per 512-sample window, carrying a 64-sample context and
 * the LSTM state from call to call (core/vad_utils.py:62-118: `session.run(None, {'input','state','sr'})`, state reset per
 * recording). Here all windows of all recordings of a batch go through the window-parallel part at once and one persistent
 * CTA per recording runs the recurrence. `weights_path` = a `.b200w` container with the vad.* tensors (weights.save_vad).
 * Create returns NULL on failure. Needs no recognizer. */
B200ASR_API void *B200AsrVadCreate(const char *weights_path, int32_t device_id);
B200ASR_API void B200AsrVadDestroy(void *vad);
/* One recording: samples[n] (16 kHz, [-1, 1]) -> probs[n / 512] (a partial last window is dropped, core/vad_utils.py:84).
 * `probs` may be NULL to query the window count. Returns the window count or <0. */
B200ASR_API int32_t B200AsrVadProbs(void *vad, const float *samples, int64_t n, float *probs);
/* A batch of recordings: recording r = samples[sample_offsets[r] .. sample_offsets[r+1]); probabilities packed, window
 * offsets written to prob_offsets[n_rec + 1] (may be NULL). Each recording starts from a zero state. */
B200ASR_API int32_t B200AsrVadProbsBatch(void *vad, const float *samples, const int64_t *sample_offsets, int32_t n_rec, float *probs,
                                         int64_t *prob_offsets);
/* Device times (ms, CUDA events) of the last call: window-parallel front end, recurrence. */
B200ASR_API int32_t B200AsrVadLastTimings(void *vad, float *frontend_ms, float *recurrence_ms);
/* enc_sess.run (core/asr_engine.py:1045-1049): packed features [sum T, 80] with x_lens[n] ->
 * packed encoder_out [sum T', 512], out_lens[n]. `out` may be NULL to query lengths only. */
B200ASR_API int32_t B200AsrEncoder(const B200AsrOfflineRecognizer *r, const float *feats, const int32_t *x_lens,
                                   int32_t n_utts, float *out, int32_t *out_lens);
/* Debug tap: after B200AsrEncoder on ONE utterance, copies an intermediate ("embed", "stack0".."stack5") into
 * out (rows*dim floats); returns rows, writes dim. */
B200ASR_API int32_t B200AsrEncoderTap(const B200AsrOfflineRecognizer *r, const char *name, float *out, int32_t *dim);
/* dec_sess.run (core/asr_engine.py:1055,1085): y[m*2] (already max(0,.)) -> out[m*512] */
B200ASR_API int32_t B200AsrDecoder(const B200AsrOfflineRecognizer *r, const int64_t *y, int32_t m, float *out);
/* joi_sess.run (core/asr_engine.py:1092): enc[m*512], dec[m*512] -> logits[m*V] */
B200ASR_API int32_t B200AsrJoiner(const B200AsrOfflineRecognizer *r, const float *enc, const float *dec, int32_t m, float *logits);
/* The two kernels a frame step of the device search runs, on caller rows (the parity tests pin them directly):
 * B200AsrDecoderJoinerInput = dec_sess.run for contexts y[m*2] followed by the joiner's input activation,
 *   dec_out[m*512] = decoder(y) and x_out[m*512] = tanh(enc + dec) (enc may be NULL = zeros), core/asr_engine.py:1072-1093;
 * B200AsrJoinerRecords = joi_sess.run's output_linear on x[m*512] with the log-softmax / top-k / entropy reduction of
 *   core/asr_engine.py:1096-1106,1159-1181 folded into its epilogue: per row and per 32-column part a record of 4 + 2*kb
 *   floats {max, sum e^(x-max), sum e^(x-max)(x-max), sum e^((x-max)/3), kb best values (descending), kb columns (int bits,
 *   -1 = none)}. Returns floats per row (ceil(V/32) * (4 + 2*kb)); `records` may be NULL to query it. kb in {4, 8, 16}. */
B200ASR_API int32_t B200AsrDecoderJoinerInput(const B200AsrOfflineRecognizer *r, const int64_t *y, const float *enc, int32_t m,
                                              float *dec_out, float *x_out);
B200ASR_API int32_t B200AsrJoinerRecords(const B200AsrOfflineRecognizer *r, const float *x, int32_t m, int32_t kb, float *records);
/* _ort_beam_search from encoder_out on (core/asr_engine.py:1051-1153), batch of n utterances with packed
 * enc_out and lens; method 0 = greedy, 1 = modified beam search. Results are returned per utterance into
 * caller arrays of capacity max_tokens each: tokens, frames, tok_logprobs, stats[4] (tsallis, margin,
 * entropy, top1); n_tokens[n]. */
B200ASR_API int32_t B200AsrBeamSearch(const B200AsrOfflineRecognizer *r, const float *enc_out, const int32_t *lens, int32_t n_utts,
                                      int32_t method, int32_t beam, int32_t max_tokens, int32_t *tokens, int32_t *frames,
                                      float *tok_logprobs, float *stats, int32_t *n_tokens);
/* One dense Linear as the encoder/joiner graphs run it (onnxruntime MatMul+Add inside enc_sess/joi_sess,
 * core/asr_engine.py:1047,1092): C[M,N] = act(A[M,K] W[N,K]^T + bias) (+ R). Host pointers; act 0/1/2 = none/SwooshL/
 * SwooshR; impl 0 = FP32 CUDA-core kernel, 1 = tcgen05 TF32, 2 = tcgen05 3xTF32, 3 = tcgen05 fp16 operand split (fp32-grade),
 * 4 = tcgen05 BF16. reps > 1 repeats the launch and
 * writes the mean device time per launch (ms, CUDA events) to *ms_per_launch (may be NULL). */
B200ASR_API int32_t B200AsrGemm(const B200AsrOfflineRecognizer *r, const float *A, const float *W, const float *bias, const float *R,
                                float *C, int32_t M, int32_t N, int32_t K, int32_t act, int32_t impl, int32_t reps, float *ms_per_launch);
/* ContextGraph.forward_one_step / finalize on the flattened automaton the device uses
 * (core/hotword_context.py:139-184): returns the score delta, writes the next state id. */
B200ASR_API double B200AsrContextForwardOneStep(const B200AsrOfflineRecognizer *r, int32_t state, int32_t token, int32_t *next_state);
B200ASR_API double B200AsrContextFinalize(const B200AsrOfflineRecognizer *r, int32_t state);
B200ASR_API int32_t B200AsrContextNumNodes(const B200AsrOfflineRecognizer *r);
/* The same automaton without a recognizer or a GPU (host only): build from token-id phrases (phrase p =
 * tokens[offsets[p] .. offsets[p+1]), per-phrase scores), step and finalize exactly as the search kernel does. Lets the
 * hotword graph be checked against core/hotword_context.py:17-184 on any machine. Create returns NULL on failure. */
B200ASR_API void *B200AsrHotwordGraphCreate(const int32_t *tokens, const int32_t *offsets, const float *scores, int32_t n_phrases);
B200ASR_API void B200AsrHotwordGraphDestroy(void *graph);
B200ASR_API int32_t B200AsrHotwordGraphNumNodes(const void *graph);
B200ASR_API double B200AsrHotwordGraphStep(const void *graph, int32_t state, int32_t token, int32_t *next_state);
B200ASR_API double B200AsrHotwordGraphFinalize(const void *graph, int32_t state);

/* ---- device-resident benchmarking hooks (inputs staged once, timed region touches HBM only) ---- */
/* Stages a ragged batch of PCM in HBM; returns a handle (>=0). */
B200ASR_API int32_t B200AsrStageBatch(const B200AsrOfflineRecognizer *r, const float *samples, const int64_t *sample_offsets, int32_t n_utts);
/* Runs fbank -> encoder -> search on a staged batch entirely on the device; token counts per utterance are
 * copied back (n_tokens may be NULL). */
B200ASR_API int32_t B200AsrRunStagedBatch(const B200AsrOfflineRecognizer *r, int32_t handle, int32_t *n_tokens);
/* `reps` (1..8) passes over a staged batch issued back to back: the search of pass k runs on its own stream beside the fbank
 * and encoder of pass k + 1, as consecutive batches of one long B200AsrDecodeMultipleOfflineStreams call do. Returns after the
 * last search; total_ms = device time from the first launch to the end of the last search; n_tokens = counts of the last pass. */
B200ASR_API int32_t B200AsrRunStagedBatchChained(const B200AsrOfflineRecognizer *r, int32_t handle, int32_t reps, int32_t *n_tokens,
                                                 float *total_ms);
B200ASR_API int32_t B200AsrReleaseBatch(const B200AsrOfflineRecognizer *r, int32_t handle);
/* Token ids / frames of utterance u of the last staged run or decode pass (bench.py's parity self-check against the
 * oracle's `_ort_beam_search` output, core/asr_engine.py:1143-1153). Returns the token count or <0. */
B200ASR_API int32_t B200AsrLastPassTokens(const B200AsrOfflineRecognizer *r, int32_t u, int32_t *tokens, int32_t *frames, int32_t cap);
/* Shape of the last pass: a batch is split into length-sorted groups whose searches run on their own streams beside the
 * encoder of the following groups (where the reference splits chunks over two worker threads, core/asr_engine.py:2384-2397).
 * n_groups, the sum of the groups' search times and each one's (ms, lane_ms8[8]), and the device->host result bytes. */
B200ASR_API int32_t B200AsrLastPipelineStats(const B200AsrOfflineRecognizer *r, int32_t *n_groups, float *search_busy_ms,
                                             float *lane_ms8, int64_t *d2h_bytes);
/* Timeline of the last pass: for each of its groups (chained batches / length-sorted groups) four device times in ms from the
 * start of the pass - encoder begin, encoder end, search begin, search end (CUDA events on the streams they ran on): the
 * search of group g beside the encoder of group g + 1. t_ms holds 4 * max_groups floats; returns the number of groups. */
B200ASR_API int32_t B200AsrLastPipelineTimeline(const B200AsrOfflineRecognizer *r, float *t_ms, int32_t max_groups);
/* Per-stage device times (ms, CUDA events on the engine's stream) of the last decode / staged run:
 * out[0]=fbank, [1]=encoder, [2]=search, [3]=total, [4]=H2D, [5]=D2H; and kernel launch count. */
B200ASR_API int32_t B200AsrLastTimings(const B200AsrOfflineRecognizer *r, float *out6, int64_t *n_launches);
/* Dominant-kernel timing: accumulated device time (ms) and FLOPs of all GEMM launches in the last run. */
B200ASR_API int32_t B200AsrLastGemmStats(const B200AsrOfflineRecognizer *r, double *ms, double *flops, int64_t *launches);
/* Algorithmic bytes of the same launches: 4 (M K + N K + M N [+ M N residual]) summed. */
B200ASR_API double B200AsrLastGemmBytes(const B200AsrOfflineRecognizer *r);
/* Enables per-GEMM CUDA-event timing (costs a little launch overhead); 0/1. */
B200ASR_API int32_t B200AsrSetProfiling(const B200AsrOfflineRecognizer *r, int32_t on);

#ifdef __cplusplus
}
#endif
#endif /* B200ASR_H_ */
