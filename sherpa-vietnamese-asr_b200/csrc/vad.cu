// Voice-activity network on the GPU: the step before the recognizer.
// The reference runs the Silero VAD ONNX graph one 512-sample window at a time - 31 sequential onnxruntime calls per
// audio-second, the LSTM state carried from call to call (/root/reference core/vad_utils.py:62-118). Architecture as restated
// in oracle/silero_ref.py (the model file is a third-party artefact that is not available offline):
//   [64 context | 512 window] -> reflect-pad 64 -> STFT (strided conv, 258 x 256 basis, hop 128) -> magnitude [129, 4]
//   -> Conv1d 129->128 k3 -> ReLU -> Conv1d 128->64 k3 s2 -> ReLU -> Conv1d 64->64 k3 s2 -> ReLU -> Conv1d 64->128 k3 -> ReLU
//   -> LSTMCell(128, 128) -> ReLU -> Conv1d 128->1 -> sigmoid.
// Everything up to and including the LSTM's input projection W_ih x + b is independent per window, so it runs over ALL windows
// of ALL recordings of a batch at once (vad_frontend_kernel: a CTA takes 16 windows, activations stay in shared memory,
// transposed weights stream from L2 with coalesced loads, every weight is used for 16 windows x frames). Only the recurrence
// W_hh h is sequential: one persistent CTA per recording (vad_lstm_kernel), thread j owns gate row j - half of its 128
// weights in registers, half in shared memory - so a window step costs ~1 us and recordings advance in parallel on different
// SMs. fp32 CUDA-core arithmetic: the whole network is 0.6 MMAC per window (19 MMAC per audio-second).
#include <math.h>
#include <stdio.h>

#include <algorithm>
#include <fstream>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/b200asr.h"
#include "common.cuh"

namespace b200asr {
void set_last_error(const std::string &msg);   // engine.cu

namespace {

constexpr int kWin = 512, kCtx = 64, kPadded = 640;
constexpr int kTW = 16;              // windows per CTA
constexpr int kFeThreads = 288;      // 258 STFT channels rounded up to warps
constexpr int kMagLd = 132;          // 129 bins padded to a multiple of 4

struct VadWeights {   // device pointers; convolution / dense weights transposed to [tap][in (padded)][out]
  const float *basisT;   // [256][258]
  const float *w0T, *b0; // [3][132][128]
  const float *w1T, *b1; // [3][128][64]
  const float *w2T, *b2; // [3][64][64]
  const float *w3T, *b3; // [1][64][128]  (a length-1 sequence with padding 1 only sees the centre tap)
  const float *wihT, *bg; // [128][512], b_ih + b_hh
  const float *whh;      // [512][128]
  const float *wo; float bo;
};

__device__ __forceinline__ float4 lds4(const float *p) { return *reinterpret_cast<const float4 *>(p); }

// acc[w][fo] += sum_{tap, ci} WT[tap][ci][co] * in[w0 + w][fo * STRIDE + tap - PAD][ci]   (frames outside [0, FIN) are zero padding)
template <int NWIN, int FIN, int FOUT, int STRIDE, int TAPS, int PAD, int CIN_PAD, int IN_LD>
__device__ __forceinline__ void conv_acc(float (&acc)[NWIN * FOUT], const float *in, int w0, const float *__restrict__ WT, int cout, int co) {
#pragma unroll
  for (int tap = 0; tap < TAPS; ++tap) {
    const float *wt = WT + (size_t)tap * CIN_PAD * cout + co;
#pragma unroll 2
    for (int c4 = 0; c4 < CIN_PAD; c4 += 4) {
      const float a0 = __ldg(wt + (size_t)(c4 + 0) * cout), a1 = __ldg(wt + (size_t)(c4 + 1) * cout);
      const float a2 = __ldg(wt + (size_t)(c4 + 2) * cout), a3 = __ldg(wt + (size_t)(c4 + 3) * cout);
#pragma unroll
      for (int w = 0; w < NWIN; ++w)
#pragma unroll
        for (int fo = 0; fo < FOUT; ++fo) {
          const int fi = fo * STRIDE + tap - PAD;
          if (fi < 0 || fi >= FIN) continue;                         // compile-time after unrolling
          const float4 x = lds4(in + ((size_t)(w0 + w) * FIN + fi) * IN_LD + c4);
          float &a = acc[w * FOUT + fo];
          a = fmaf(a0, x.x, a); a = fmaf(a1, x.y, a); a = fmaf(a2, x.z, a); a = fmaf(a3, x.w, a);
        }
    }
  }
}

struct VadBatch {
  const float *pcm;            // all recordings, concatenated
  const long long *soff;       // [n_rec] first sample of each recording
  const int *woff;             // [n_rec + 1] cumulative window counts
  int n_rec, n_win;
};

__device__ __forceinline__ int rec_of_window(const VadBatch &b, int g) {
  int lo = 0, hi = b.n_rec - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(b.woff + mid) <= g) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// shared memory of a CTA: padded inputs (dead after the STFT, so conv0's output reuses the space), magnitudes, h1..h3 = 95 KB,
// two CTAs per SM
constexpr size_t kFeSmem = (size_t)(kTW * kPadded + kTW * 4 * kMagLd + kTW * 2 * 64 + kTW * 64 + kTW * 128) * sizeof(float);
static_assert(kTW * kPadded >= kTW * 4 * 128, "h0 must fit in the input staging area");

__global__ void __launch_bounds__(kFeThreads) vad_frontend_kernel(VadWeights W, VadBatch b, float *__restrict__ gx) {
  extern __shared__ __align__(16) float sm[];
  float *xin = sm;                          // [16][640]
  float *h0 = sm;                           // [16][4][128], written after the last read of xin
  float *mag = xin + kTW * kPadded;         // [16][4][132]
  float *h1 = mag + kTW * 4 * kMagLd;       // [16][2][64]
  float *h2 = h1 + kTW * 2 * 64;            // [16][1][64]
  float *h3 = h2 + kTW * 64;                // [16][1][128]
  const int tid = threadIdx.x;
  const int g0 = blockIdx.x * kTW;
  // ---- stage the padded inputs: 64 samples of context (zeros for a recording's first window), the window, 64 reflected
  for (int w = 0; w < kTW; ++w) {
    const int g = g0 + w;
    float *row = xin + w * kPadded;
    if (g >= b.n_win) { for (int i = tid; i < kPadded; i += kFeThreads) row[i] = 0.f; continue; }
    const int r = rec_of_window(b, g);
    const int i_loc = g - __ldg(b.woff + r);
    const float *src = b.pcm + __ldg(b.soff + r) + (long long)i_loc * kWin - kCtx;
    for (int i = tid; i < kCtx + kWin; i += kFeThreads) row[i] = (i_loc == 0 && i < kCtx) ? 0.f : __ldg(src + i);
  }
  __syncthreads();
  for (int i = tid; i < kTW * 64; i += kFeThreads) {       // reflect: padded[576 + j] = x[574 - j]
    const int w = i >> 6, j = i & 63;
    xin[w * kPadded + 576 + j] = xin[w * kPadded + 574 - j];
  }
  __syncthreads();
  // ---- STFT: thread = one of the 258 basis rows, 16 windows x 4 frames accumulators
  {
    float acc[kTW * 4];
#pragma unroll
    for (int i = 0; i < kTW * 4; ++i) acc[i] = 0.f;
    if (tid < 258) {
#pragma unroll 1
      for (int k = 0; k < 256; k += 4) {
        const float a0 = __ldg(W.basisT + (k + 0) * 258 + tid), a1 = __ldg(W.basisT + (k + 1) * 258 + tid);
        const float a2 = __ldg(W.basisT + (k + 2) * 258 + tid), a3 = __ldg(W.basisT + (k + 3) * 258 + tid);
#pragma unroll
        for (int w = 0; w < kTW; ++w)
#pragma unroll
          for (int f = 0; f < 4; ++f) {
            const float4 x = lds4(xin + w * kPadded + f * 128 + k);
            float &a = acc[w * 4 + f];
            a = fmaf(a0, x.x, a); a = fmaf(a1, x.y, a); a = fmaf(a2, x.z, a); a = fmaf(a3, x.w, a);
          }
      }
    }
    // imaginary parts to shared memory, then the owners of the real parts form the magnitudes in place
    if (tid >= 129 && tid < 258) {
#pragma unroll
      for (int i = 0; i < kTW * 4; ++i) mag[i * kMagLd + (tid - 129)] = acc[i];
    }
    __syncthreads();
    if (tid < 129) {
#pragma unroll
      for (int i = 0; i < kTW * 4; ++i) {
        const float im = mag[i * kMagLd + tid];
        mag[i * kMagLd + tid] = sqrtf(acc[i] * acc[i] + im * im);
      }
    } else if (tid < 132) {
      for (int i = 0; i < kTW * 4; ++i) mag[i * kMagLd + tid] = 0.f;      // bins 129..131: padding of the reduction
    }
    __syncthreads();
  }
  // ---- conv0 129 -> 128, k3 p1, 4 frames: thread = (channel, half of the windows)
  if (tid < 256) {
    const int co = tid & 127, w0 = (tid >> 7) * 8;
    float acc[8 * 4];
    const float bias = __ldg(W.b0 + co);
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = bias;
    conv_acc<8, 4, 4, 1, 3, 1, kMagLd, kMagLd>(acc, mag, w0, W.w0T, 128, co);
#pragma unroll
    for (int w = 0; w < 8; ++w)
#pragma unroll
      for (int f = 0; f < 4; ++f) h0[((w0 + w) * 4 + f) * 128 + co] = fmaxf(acc[w * 4 + f], 0.f);
  }
  __syncthreads();
  // ---- conv1 128 -> 64, k3 s2 p1, 4 -> 2 frames: thread = (channel, quarter of the windows)
  if (tid < 256) {
    const int co = tid & 63, w0 = (tid >> 6) * 4;
    float acc[4 * 2];
    const float bias = __ldg(W.b1 + co);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = bias;
    conv_acc<4, 4, 2, 2, 3, 1, 128, 128>(acc, h0, w0, W.w1T, 64, co);
#pragma unroll
    for (int w = 0; w < 4; ++w)
#pragma unroll
      for (int f = 0; f < 2; ++f) h1[((w0 + w) * 2 + f) * 64 + co] = fmaxf(acc[w * 2 + f], 0.f);
  }
  __syncthreads();
  // ---- conv2 64 -> 64, k3 s2 p1, 2 -> 1 frame
  if (tid < 256) {
    const int co = tid & 63, w0 = (tid >> 6) * 4;
    float acc[4];
    const float bias = __ldg(W.b2 + co);
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = bias;
    conv_acc<4, 2, 1, 2, 3, 1, 64, 64>(acc, h1, w0, W.w2T, 64, co);
#pragma unroll
    for (int w = 0; w < 4; ++w) h2[(w0 + w) * 64 + co] = fmaxf(acc[w], 0.f);
  }
  __syncthreads();
  // ---- conv3 64 -> 128, k3 p1 on one frame = its centre tap
  if (tid < 256) {
    const int co = tid & 127, w0 = (tid >> 7) * 8;
    float acc[8];
    const float bias = __ldg(W.b3 + co);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = bias;
    conv_acc<8, 1, 1, 1, 1, 0, 64, 64>(acc, h2, w0, W.w3T, 128, co);
#pragma unroll
    for (int w = 0; w < 8; ++w) h3[(w0 + w) * 128 + co] = fmaxf(acc[w], 0.f);
  }
  __syncthreads();
  // ---- LSTM input projection: gx = W_ih x + b_ih + b_hh, 512 gate rows
  for (int co = tid; co < 512; co += kFeThreads) {
    float acc[kTW];
    const float bias = __ldg(W.bg + co);
#pragma unroll
    for (int i = 0; i < kTW; ++i) acc[i] = bias;
    conv_acc<kTW, 1, 1, 1, 1, 0, 128, 128>(acc, h3, 0, W.wihT, 512, co);
#pragma unroll
    for (int w = 0; w < kTW; ++w)
      if (g0 + w < b.n_win) gx[(size_t)(g0 + w) * 512 + co] = acc[w];
  }
}

// ---- recurrence: one CTA per recording, thread j = gate row j (PyTorch order i, f, g, o)
constexpr int kLstmThreads = 512;
constexpr size_t kLstmSmem = (size_t)(64 * 512 + 512 + 2 * 128 + 4) * sizeof(float);

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(kLstmThreads, 1) vad_lstm_kernel(VadWeights W, VadBatch b, const float *__restrict__ gx,
                                                                   float *__restrict__ probs) {
  extern __shared__ __align__(16) float sm[];
  float *wsm = sm;                 // [64][512]: W_hh[j][64 + k] at wsm[k * 512 + j]
  float *gates = wsm + 64 * 512;   // [512]
  float *hbuf = gates + 512;       // [2][128]
  float *psum = hbuf + 256;        // [4] per-warp partial sums of the output head
  const int j = threadIdx.x, r = blockIdx.x;
  float wreg[64];
#pragma unroll
  for (int k = 0; k < 64; ++k) wreg[k] = __ldg(W.whh + (size_t)j * 128 + k);
  for (int k = 0; k < 64; ++k) wsm[k * 512 + j] = __ldg(W.whh + (size_t)j * 128 + 64 + k);
  if (j < 256) hbuf[j] = 0.f;
  float c = 0.f;                              // cell state of unit j (threads 0..127)
  const float wo = j < 128 ? __ldg(W.wo + j) : 0.f;
  const int w_begin = __ldg(b.woff + r), w_end = __ldg(b.woff + r + 1);
  __syncthreads();
  float g_next = w_begin < w_end ? __ldg(gx + (size_t)w_begin * 512 + j) : 0.f;
  for (int w = w_begin; w < w_end; ++w) {
    const float *h = hbuf + ((w - w_begin) & 1) * 128;
    float *hn = hbuf + (((w - w_begin) & 1) ^ 1) * 128;
    float a0 = g_next, a1 = 0.f;
    if (w + 1 < w_end) g_next = __ldg(gx + (size_t)(w + 1) * 512 + j);     // next step's input term, off the critical path
#pragma unroll
    for (int k = 0; k < 64; k += 4) {
      const float4 hv = lds4(h + k);
      a0 = fmaf(wreg[k], hv.x, a0); a0 = fmaf(wreg[k + 1], hv.y, a0); a0 = fmaf(wreg[k + 2], hv.z, a0); a0 = fmaf(wreg[k + 3], hv.w, a0);
    }
#pragma unroll
    for (int k = 0; k < 64; k += 4) {
      const float4 hv = lds4(h + 64 + k);
      a1 = fmaf(wsm[(k + 0) * 512 + j], hv.x, a1); a1 = fmaf(wsm[(k + 1) * 512 + j], hv.y, a1);
      a1 = fmaf(wsm[(k + 2) * 512 + j], hv.z, a1); a1 = fmaf(wsm[(k + 3) * 512 + j], hv.w, a1);
    }
    gates[j] = a0 + a1;
    __syncthreads();
    if (j < 128) {
      const float ig = sigmoidf_(gates[j]), fg = sigmoidf_(gates[128 + j]), gg = tanhf(gates[256 + j]), og = sigmoidf_(gates[384 + j]);
      c = fmaf(fg, c, ig * gg);
      const float hv = og * tanhf(c);
      hn[j] = hv;
      // output head on this step's h: relu -> 128 -> 1 -> sigmoid (warp partial sums, merged by warp 0 below)
      float part = fmaxf(hv, 0.f) * wo;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if ((j & 31) == 0) psum[j >> 5] = part;
    }
    __syncthreads();     // h of this step and the partial sums are complete; gates[] may be rewritten by the next step
    if (j == 0) probs[w] = sigmoidf_(psum[0] + psum[1] + psum[2] + psum[3] + W.bo);
    // psum is rewritten only after the next step's first barrier, which thread 0 reaches after this read
  }
}

struct Vad {
  int device = 0;
  VadWeights W{};
  std::vector<float *> owned;
  cudaStream_t st = nullptr;
  std::mutex mu;
  // workspaces
  float *d_pcm = nullptr; size_t pcm_cap = 0;
  float *d_gx = nullptr; size_t gx_cap = 0;
  float *d_probs = nullptr; size_t pr_cap = 0;
  long long *d_soff = nullptr; int *d_woff = nullptr; size_t rec_cap = 0;
  float last_ms[2] = {0.f, 0.f};
  cudaEvent_t ev[3]{};
  ~Vad() {
    for (float *p : owned) cudaFree(p);
    cudaFree(d_pcm); cudaFree(d_gx); cudaFree(d_probs); cudaFree(d_soff); cudaFree(d_woff);
    for (auto &e : ev) if (e) cudaEventDestroy(e);
    if (st) cudaStreamDestroy(st);
  }
};

// minimal reader of the B200ASRW container (weights.py): tensors named vad.*
std::map<std::string, std::pair<std::vector<int>, std::vector<float>>> read_container(const std::string &path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open weight container: " + path);
  std::string first;
  std::getline(f, first);
  char magic[32];
  int ver = 0;
  long long hb = 0;
  if (sscanf(first.c_str(), "%31s %d %lld", magic, &ver, &hb) != 3 || std::string(magic) != "B200ASRW")
    throw std::runtime_error(path + ": not a B200ASRW container");
  f.seekg(0);
  std::string header((size_t)hb, '\0');
  f.read(&header[0], hb);
  std::stringstream hs(header);
  std::string line;
  std::getline(hs, line);
  std::map<std::string, std::pair<std::vector<int>, std::vector<float>>> out;
  while (std::getline(hs, line)) {
    std::stringstream ls(line);
    std::string kind;
    ls >> kind;
    if (kind == "end" || kind.empty() || kind[0] == '\0') break;
    if (kind != "tensor") continue;
    std::string name, dt;
    int nd;
    ls >> name >> dt >> nd;
    std::vector<int> shape(nd);
    for (int i = 0; i < nd; ++i) ls >> shape[i];
    long long off, nbytes;
    ls >> off >> nbytes;
    if (dt != "f32") throw std::runtime_error("unsupported dtype in container: " + dt);
    std::vector<float> v((size_t)nbytes / 4);
    const auto pos = f.tellg();
    f.seekg(hb + off);
    f.read(reinterpret_cast<char *>(v.data()), nbytes);
    if (!f) throw std::runtime_error(path + ": truncated tensor " + name);
    f.seekg(pos);
    out[name] = {shape, std::move(v)};
  }
  return out;
}

template <typename T>
void grow(T *&p, size_t &cap, size_t need) {
  if (need <= cap) return;
  if (p) cudaFree(p);
  p = nullptr;
  cap = need + need / 4 + 64;
  CUDA_CHECK(cudaMalloc(&p, cap * sizeof(T)));
}

}  // namespace
}  // namespace b200asr

using namespace b200asr;

extern "C" {

void *B200AsrVadCreate(const char *weights_path, int32_t device_id) {
  Vad *v = nullptr;
  try {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) throw std::runtime_error(std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libb200asr has no CPU fallback)");
    if (device_id < 0 || device_id >= ndev) throw std::runtime_error("device_id out of range");
    CUDA_CHECK(cudaSetDevice(device_id));
    auto T = read_container(weights_path ? weights_path : "");
    auto need = [&](const char *name, std::vector<int> shape) -> const std::vector<float> & {
      auto it = T.find(name);
      if (it == T.end()) throw std::runtime_error(std::string("missing tensor: ") + name);
      if (it->second.first != shape) throw std::runtime_error(std::string("unexpected shape for ") + name);
      return it->second.second;
    };
    v = new Vad();
    v->device = device_id;
    auto up = [&](const std::vector<float> &h) {
      float *d;
      CUDA_CHECK(cudaMalloc(&d, std::max<size_t>(h.size(), 4) * sizeof(float)));
      CUDA_CHECK(cudaMemcpy(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
      v->owned.push_back(d);
      return d;
    };
    // [out][in][k] -> [k][in padded][out]
    auto convT = [&](const std::vector<float> &w, int co, int ci, int k, int ci_pad, int k0, int k1) {
      std::vector<float> t((size_t)(k1 - k0) * ci_pad * co, 0.f);
      for (int o = 0; o < co; ++o) for (int i = 0; i < ci; ++i) for (int kk = k0; kk < k1; ++kk)
        t[((size_t)(kk - k0) * ci_pad + i) * co + o] = w[((size_t)o * ci + i) * k + kk];
      return t;
    };
    const auto &basis = need("vad.stft.basis", {258, 256});
    std::vector<float> bt((size_t)256 * 258);
    for (int c = 0; c < 258; ++c) for (int k = 0; k < 256; ++k) bt[(size_t)k * 258 + c] = basis[(size_t)c * 256 + k];
    v->W.basisT = up(bt);
    v->W.w0T = up(convT(need("vad.enc0.weight", {128, 129, 3}), 128, 129, 3, kMagLd, 0, 3)); v->W.b0 = up(need("vad.enc0.bias", {128}));
    v->W.w1T = up(convT(need("vad.enc1.weight", {64, 128, 3}), 64, 128, 3, 128, 0, 3)); v->W.b1 = up(need("vad.enc1.bias", {64}));
    v->W.w2T = up(convT(need("vad.enc2.weight", {64, 64, 3}), 64, 64, 3, 64, 0, 3)); v->W.b2 = up(need("vad.enc2.bias", {64}));
    v->W.w3T = up(convT(need("vad.enc3.weight", {128, 64, 3}), 128, 64, 3, 64, 1, 2)); v->W.b3 = up(need("vad.enc3.bias", {128}));
    const auto &wih = need("vad.lstm.weight_ih", {512, 128});
    std::vector<float> wt((size_t)128 * 512);
    for (int o = 0; o < 512; ++o) for (int i = 0; i < 128; ++i) wt[(size_t)i * 512 + o] = wih[(size_t)o * 128 + i];
    v->W.wihT = up(wt);
    const auto &bih = need("vad.lstm.bias_ih", {512}), &bhh = need("vad.lstm.bias_hh", {512});
    std::vector<float> bg(512);
    for (int i = 0; i < 512; ++i) bg[i] = bih[i] + bhh[i];
    v->W.bg = up(bg);
    v->W.whh = up(need("vad.lstm.weight_hh", {512, 128}));
    v->W.wo = up(need("vad.out.weight", {1, 128, 1}));
    v->W.bo = need("vad.out.bias", {1})[0];
    CUDA_CHECK(cudaStreamCreateWithFlags(&v->st, cudaStreamNonBlocking));
    for (auto &e2 : v->ev) CUDA_CHECK(cudaEventCreate(&e2));
    set_max_dynamic_smem(vad_frontend_kernel, kFeSmem);
    set_max_dynamic_smem(vad_lstm_kernel, kLstmSmem);
    return v;
  } catch (const std::exception &e) {
    set_last_error(e.what());
    delete v;
    return nullptr;
  }
}

void B200AsrVadDestroy(void *vad) { delete reinterpret_cast<Vad *>(vad); }

int32_t B200AsrVadProbsBatch(void *vad, const float *samples, const int64_t *sample_offsets, int32_t n_rec, float *probs,
                             int64_t *prob_offsets) {
  try {
    Vad *v = reinterpret_cast<Vad *>(vad);
    if (!v || n_rec < 0 || (n_rec > 0 && (!samples || !sample_offsets))) throw std::runtime_error("B200AsrVadProbsBatch: bad arguments");
    std::vector<long long> soff(n_rec);
    std::vector<int> woff(n_rec + 1, 0);
    const long long base = n_rec ? sample_offsets[0] : 0;
    for (int r = 0; r < n_rec; ++r) {
      const long long n = sample_offsets[r + 1] - sample_offsets[r];
      if (n < 0) throw std::runtime_error("B200AsrVadProbsBatch: sample offsets must not decrease");
      soff[r] = sample_offsets[r] - base;
      const long long nw = n / kWin;                                   // a partial last window is dropped (core/vad_utils.py:84)
      if ((long long)woff[r] + nw > INT32_MAX) throw std::runtime_error("B200AsrVadProbsBatch: batch too long");
      woff[r + 1] = woff[r] + (int)nw;
    }
    if (prob_offsets) for (int r = 0; r <= n_rec; ++r) prob_offsets[r] = woff[r];
    const int n_win = woff[n_rec];
    if (!probs || n_win == 0) return n_win;
    std::lock_guard<std::mutex> lk(v->mu);
    CUDA_CHECK(cudaSetDevice(v->device));
    const long long total = sample_offsets[n_rec] - base;
    grow(v->d_pcm, v->pcm_cap, (size_t)total + kCtx);
    grow(v->d_gx, v->gx_cap, (size_t)n_win * 512);
    grow(v->d_probs, v->pr_cap, (size_t)n_win);
    if ((size_t)n_rec + 1 > v->rec_cap) {
      cudaFree(v->d_soff); cudaFree(v->d_woff);
      v->rec_cap = (size_t)n_rec + 16;
      CUDA_CHECK(cudaMalloc(&v->d_soff, v->rec_cap * sizeof(long long)));
      CUDA_CHECK(cudaMalloc(&v->d_woff, v->rec_cap * sizeof(int)));
    }
    // the PCM sits kCtx floats into its buffer so the first window's (unused) context read stays inside the allocation
    CUDA_CHECK(cudaMemcpyAsync(v->d_pcm + kCtx, samples + base, (size_t)total * sizeof(float), cudaMemcpyHostToDevice, v->st));
    CUDA_CHECK(cudaMemcpyAsync(v->d_soff, soff.data(), n_rec * sizeof(long long), cudaMemcpyHostToDevice, v->st));
    CUDA_CHECK(cudaMemcpyAsync(v->d_woff, woff.data(), (n_rec + 1) * sizeof(int), cudaMemcpyHostToDevice, v->st));
    VadBatch b{v->d_pcm + kCtx, v->d_soff, v->d_woff, n_rec, n_win};
    CUDA_CHECK(cudaEventRecord(v->ev[0], v->st));
    vad_frontend_kernel<<<(n_win + kTW - 1) / kTW, kFeThreads, kFeSmem, v->st>>>(v->W, b, v->d_gx);
    count_launch(); KERNEL_CHECK();
    CUDA_CHECK(cudaEventRecord(v->ev[1], v->st));
    vad_lstm_kernel<<<n_rec, kLstmThreads, kLstmSmem, v->st>>>(v->W, b, v->d_gx, v->d_probs);
    count_launch(); KERNEL_CHECK();
    CUDA_CHECK(cudaEventRecord(v->ev[2], v->st));
    CUDA_CHECK(cudaMemcpyAsync(probs, v->d_probs, (size_t)n_win * sizeof(float), cudaMemcpyDeviceToHost, v->st));
    CUDA_CHECK(cudaStreamSynchronize(v->st));
    cudaEventElapsedTime(&v->last_ms[0], v->ev[0], v->ev[1]);
    cudaEventElapsedTime(&v->last_ms[1], v->ev[1], v->ev[2]);
    return n_win;
  } catch (const std::exception &e) {
    set_last_error(e.what());
    return -1;
  }
}

int32_t B200AsrVadProbs(void *vad, const float *samples, int64_t n, float *probs) {
  const int64_t offs[2] = {0, n};
  return B200AsrVadProbsBatch(vad, samples, offs, 1, probs, nullptr);
}

int32_t B200AsrVadLastTimings(void *vad, float *frontend_ms, float *recurrence_ms) {
  Vad *v = reinterpret_cast<Vad *>(vad);
  if (!v) return -1;
  if (frontend_ms) *frontend_ms = v->last_ms[0];
  if (recurrence_ms) *recurrence_ms = v->last_ms[1];
  return 0;
}

}  // extern "C"
