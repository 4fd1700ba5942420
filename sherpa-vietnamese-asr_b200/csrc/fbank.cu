// Kaldi-compatible 80-bin log-mel fbank on sm_100a.
// Replaces kaldi-native-fbank as called by compute_fbank_ort (/root/reference core/asr_engine.py:698-721;
// options :703-713: 16 kHz, 25/10 ms, povey, dither 0, snip_edges False, 80 bins, 20-7600 Hz).
// Exact algorithm statement: SURVEY.md Appendix A.
//
// One CTA = 16 consecutive frames of one utterance. The 400+15*160 = 2800 PCM samples the tile needs are read
// from HBM once, coalesced, into shared memory (reflected at the utterance edges); each warp then owns two
// frames: DC removal (warp reduction), pre-emphasis, povey window, 512-point real FFT done as a 256-point
// complex FFT held in registers (8 points per lane: radix-8 over the register index, one conflict-free exchange
// through shared memory, radix-8 again, then a radix-4 across lane quads with shuffles; twiddles by short
// multiplication chains), power spectrum, sparse mel projection (each bin touches a contiguous run of FFT bins)
// and log. Output rows are written coalesced.
// HBM-bound: algorithmic bytes = 4 B/sample in + 320 B/frame out (96 KB per audio second).
#include <math.h>
#include <vector>

#include "common.cuh"

namespace b200asr {

namespace {
constexpr int kWin = 400, kShift = 160, kFft = 512, kBins = 80, kHalf = 256;
constexpr int kFramesPerCta = 16, kWarps = 8;
constexpr int kTileSamples = kWin + (kFramesPerCta - 1) * kShift;  // 2800
constexpr float kPreemph = 0.97f;
constexpr float kFltEps = 1.1920929e-07f;

__device__ __forceinline__ void cmul(float ar, float ai, float br, float bi, float &cr, float &ci) {
  cr = ar * br - ai * bi;
  ci = ar * bi + ai * br;
}
// 4-point forward DFT of (p0..p3) -> natural order
__device__ __forceinline__ void dft4(float (&r)[4], float (&i)[4]) {
  const float u0r = r[0] + r[2], u0i = i[0] + i[2], u1r = r[0] - r[2], u1i = i[0] - i[2];
  const float v0r = r[1] + r[3], v0i = i[1] + i[3];
  const float dr = r[1] - r[3], di = i[1] - i[3];
  const float v1r = di, v1i = -dr;                      // (p1 - p3) * (-i)
  r[0] = u0r + v0r; i[0] = u0i + v0i;
  r[1] = u1r + v1r; i[1] = u1i + v1i;
  r[2] = u0r - v0r; i[2] = u0i - v0i;
  r[3] = u1r - v1r; i[3] = u1i - v1i;
}
// 8-point forward DFT in registers, natural order in and out
__device__ __forceinline__ void dft8(float (&xr)[8], float (&xi)[8]) {
  constexpr float kS = 0.70710678118654752f;
  float ar[4], ai[4], br[4], bi[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    ar[j] = xr[j] + xr[j + 4]; ai[j] = xi[j] + xi[j + 4];
    br[j] = xr[j] - xr[j + 4]; bi[j] = xi[j] - xi[j + 4];
  }
  { const float t = (br[1] + bi[1]) * kS, u = (bi[1] - br[1]) * kS; br[1] = t; bi[1] = u; }      // * (1 - i)/sqrt2
  { const float t = bi[2], u = -br[2]; br[2] = t; bi[2] = u; }                                   // * (-i)
  { const float t = (bi[3] - br[3]) * kS, u = -(br[3] + bi[3]) * kS; br[3] = t; bi[3] = u; }     // * (-1 - i)/sqrt2
  dft4(ar, ai);
  dft4(br, bi);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    xr[2 * j] = ar[j]; xi[2 * j] = ai[j];
    xr[2 * j + 1] = br[j]; xi[2 * j + 1] = bi[j];
  }
}
// w^0..w^7 from w (multiplication chains of depth <= 3)
__device__ __forceinline__ void powers8(float wr, float wi, float (&pr)[8], float (&pi)[8]) {
  pr[0] = 1.f; pi[0] = 0.f; pr[1] = wr; pi[1] = wi;
  cmul(wr, wi, wr, wi, pr[2], pi[2]);
  cmul(pr[2], pi[2], wr, wi, pr[3], pi[3]);
  cmul(pr[2], pi[2], pr[2], pi[2], pr[4], pi[4]);
  cmul(pr[4], pi[4], wr, wi, pr[5], pi[5]);
  cmul(pr[3], pi[3], pr[3], pi[3], pr[6], pi[6]);
  cmul(pr[4], pi[4], pr[3], pi[3], pr[7], pi[7]);
}
constexpr int kExStride = 36;              // exchange buffer: [8][36] keeps both the stores and the strided loads conflict-free
constexpr int kExSize = 8 * kExStride;     // 288 >= 280 = natural-order spectrum with 8 floats of padding per 64
__device__ __forceinline__ int spec_addr(int k) { return k + 8 * (k >> 6); }

__global__ void __launch_bounds__(kWarps * 32) fbank_kernel(const float *__restrict__ samples, const long long *__restrict__ sample_off,
                                                            const long long *__restrict__ sample_len,
                                                            const long long *__restrict__ frame_off, const float *__restrict__ window,
                                                            const float *__restrict__ twiddle, const int *__restrict__ mel_start,
                                                            const int *__restrict__ mel_count, const int *__restrict__ mel_off,
                                                            const float *__restrict__ mel_w, float *__restrict__ out) {
  const int u = blockIdx.y;
  const long long s_begin = sample_off[u];
  const long long n = sample_len ? sample_len[u] : sample_off[u + 1] - s_begin;   // explicit lengths: utterances need not be packed
  const long long f_begin = frame_off[u];
  const int T = (int)(frame_off[u + 1] - f_begin);
  const int f0 = blockIdx.x * kFramesPerCta;
  if (f0 >= T) return;

  __shared__ float s_pcm[kTileSamples];
  __shared__ float s_re[kWarps][kExSize];
  __shared__ float s_im[kWarps][kExSize];
  __shared__ float s_pow[kWarps][kHalf];
  __shared__ float s_tw[kFft];
  __shared__ float s_win[kWin];

  const int tid = threadIdx.x;
  const float *pcm = samples + s_begin;
  const long long base = (long long)f0 * kShift + (kShift / 2 - kWin / 2);  // first sample index of the tile (may be < 0)
  for (int i = tid; i < kTileSamples; i += blockDim.x) {
    long long j = base + i;
    // Kaldi ExtractWindow reflection, repeated until in range
    while (j < 0 || j >= n) {
      if (j < 0) j = -j - 1;
      if (j >= n) j = 2 * n - 1 - j;
    }
    s_pcm[i] = __ldg(pcm + j);
  }
  for (int i = tid; i < kFft; i += blockDim.x) s_tw[i] = twiddle[i];
  for (int i = tid; i < kWin; i += blockDim.x) s_win[i] = window[i];
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31;
  float *re = s_re[warp], *im = s_im[warp], *pw = s_pow[warp];

  for (int fl = warp; fl < kFramesPerCta; fl += kWarps) {
    const int f = f0 + fl;
    if (f >= T) break;  // warp-uniform
    const float *x = s_pcm + fl * kShift;
    // mean of the 400 samples
    float sum = 0.f;
    for (int i = lane; i < kWin; i += 32) sum += x[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)kWin;
    // pre-emphasis + window, packed as z[n] = w[2n] + i*w[2n+1]; lane holds n = 32 n1 + lane for n1 = 0..7
    float xr[8], xi[8];
#pragma unroll
    for (int n1 = 0; n1 < 8; ++n1) {
      const int nidx = 32 * n1 + lane;
      float v[2];
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const int i = 2 * nidx + p;
        float w = 0.f;
        if (i < kWin) {
          const float a = x[i] - mean;
          const float b = x[i > 0 ? i - 1 : 0] - mean;
          w = (a - kPreemph * b) * s_win[i];
        }
        v[p] = w;
      }
      xr[n1] = v[0];
      xi[n1] = v[1];
    }
    // 256 = 8 x 8 x 4. Step 1: radix-8 over n1 (registers), twiddle W_256^(lane k1)
    dft8(xr, xi);
    {
      float pr[8], pi[8];
      powers8(s_tw[4 * lane], s_tw[4 * lane + 1], pr, pi);     // W_256^lane = W_512^(2 lane)
#pragma unroll
      for (int k1 = 1; k1 < 8; ++k1) cmul(xr[k1], xi[k1], pr[k1], pi[k1], xr[k1], xi[k1]);
    }
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) { re[k1 * kExStride + lane] = xr[k1]; im[k1 * kExStride + lane] = xi[k1]; }
    __syncwarp();
    // Step 2: lane' = (k1, b) takes n2 = 4 a + b, a = 0..7: radix-8 over a, twiddle W_32^(b c)
    {
      const int k1 = lane >> 2, b = lane & 3;
#pragma unroll
      for (int a = 0; a < 8; ++a) { xr[a] = re[k1 * kExStride + 4 * a + b]; xi[a] = im[k1 * kExStride + 4 * a + b]; }
      dft8(xr, xi);
      float pr[8], pi[8];
      powers8(s_tw[2 * (16 * b)], s_tw[2 * (16 * b) + 1], pr, pi);   // W_32^b = W_512^(16 b)
#pragma unroll
      for (int c = 1; c < 8; ++c) cmul(xr[c], xi[c], pr[c], pi[c], xr[c], xi[c]);
      // Step 3: radix-4 over b, across the lane quad, two shuffle butterflies
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float pr2 = __shfl_xor_sync(0xffffffffu, xr[c], 2), pi2 = __shfl_xor_sync(0xffffffffu, xi[c], 2);
        float ur = (b & 2) ? pr2 - xr[c] : xr[c] + pr2;
        float ui = (b & 2) ? pi2 - xi[c] : xi[c] + pi2;
        if (b == 3) { const float t = ui; ui = -ur; ur = t; }        // * (-i)
        const float qr = __shfl_xor_sync(0xffffffffu, ur, 1), qi = __shfl_xor_sync(0xffffffffu, ui, 1);
        xr[c] = (b & 1) ? qr - ur : ur + qr;
        xi[c] = (b & 1) ? qi - ui : ui + qi;
      }
      __syncwarp();        // everybody has read the exchange buffer
      // lane quad position b holds output index d = {0, 2, 1, 3}[b]; spectrum index k = k1 + 8 c + 64 d
      const int d = ((b & 1) << 1) | (b >> 1);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int ad = spec_addr(k1 + 8 * c + 64 * d);
        re[ad] = xr[c];
        im[ad] = xi[c];
      }
    }
    __syncwarp();
    // unpack to the real FFT of 512 points: X[k] = E[k] + W^k O[k]
    for (int k = lane; k < kHalf; k += 32) {
      const int km = (kHalf - k) & (kHalf - 1);
      const float zr = re[spec_addr(k)], zi = im[spec_addr(k)], yr = re[spec_addr(km)], yi = -im[spec_addr(km)];  // conj(Z[N-k])
      const float er = 0.5f * (zr + yr), ei = 0.5f * (zi + yi);
      // O = (Z - conj(Z'))/(2i) = ( (zi - yi) - i (zr - yr) ) / 2
      const float orr = 0.5f * (zi - yi), oi = -0.5f * (zr - yr);
      const float c = s_tw[2 * k], sn = s_tw[2 * k + 1];
      const float xr = er + (orr * c - oi * sn);
      const float xi = ei + (orr * sn + oi * c);
      pw[k] = xr * xr + xi * xi;
    }
    __syncwarp();
    float *orow = out + (f_begin + f) * kBins;
    for (int b = lane; b < kBins; b += 32) {
      const int st = mel_start[b], cnt = mel_count[b];
      const float *w = mel_w + mel_off[b];
      float acc = 0.f;
      for (int i = 0; i < cnt; ++i) acc += w[i] * pw[st + i];
      orow[b] = logf(fmaxf(acc, kFltEps));
    }
    __syncwarp();
  }
}
}  // namespace

void fbank_tables_create(FbankTables *t) {
  std::vector<float> win(kWin), tw(kFft);
  for (int i = 0; i < kWin; ++i) win[i] = (float)pow(0.5 - 0.5 * cos(2.0 * M_PI * i / (kWin - 1)), 0.85);
  for (int k = 0; k < kHalf; ++k) {
    tw[2 * k] = (float)cos(-2.0 * M_PI * k / kFft);
    tw[2 * k + 1] = (float)sin(-2.0 * M_PI * k / kFft);
  }
  auto mel = [](double f) { return 1127.0 * log(1.0 + f / 700.0); };
  const double mlo = mel(20.0), mhi = mel(7600.0), delta = (mhi - mlo) / (kBins + 1);
  std::vector<int> st(kBins), cnt(kBins), off(kBins);
  std::vector<float> w;
  for (int b = 0; b < kBins; ++b) {
    const double l = mlo + b * delta, c = mlo + (b + 1) * delta, r = mlo + (b + 2) * delta;
    int first = -1, last = -1;
    std::vector<float> row(kHalf, 0.f);
    for (int k = 0; k < kHalf; ++k) {
      const double m = mel(k * 16000.0 / kFft);
      double v = 0.0;
      if (m > l && m <= c) v = (m - l) / (c - l);
      else if (m > c && m < r) v = (r - m) / (r - c);
      if (v != 0.0) {
        if (first < 0) first = k;
        last = k;
      }
      row[k] = (float)v;
    }
    off[b] = (int)w.size();
    st[b] = first < 0 ? 0 : first;
    cnt[b] = first < 0 ? 0 : last - first + 1;
    for (int k = 0; k < cnt[b]; ++k) w.push_back(row[st[b] + k]);
  }
  CUDA_CHECK(cudaMalloc(&t->window, kWin * sizeof(float)));
  CUDA_CHECK(cudaMalloc(&t->twiddle, kFft * sizeof(float)));
  CUDA_CHECK(cudaMalloc(&t->mel_start, kBins * sizeof(int)));
  CUDA_CHECK(cudaMalloc(&t->mel_count, kBins * sizeof(int)));
  CUDA_CHECK(cudaMalloc(&t->mel_off, kBins * sizeof(int)));
  CUDA_CHECK(cudaMalloc(&t->mel_w, w.size() * sizeof(float)));
  CUDA_CHECK(cudaMemcpy(t->window, win.data(), kWin * sizeof(float), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(t->twiddle, tw.data(), kFft * sizeof(float), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(t->mel_start, st.data(), kBins * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(t->mel_count, cnt.data(), kBins * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(t->mel_off, off.data(), kBins * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(t->mel_w, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
}

void fbank_tables_destroy(FbankTables *t) {
  cudaFree(t->window); cudaFree(t->twiddle); cudaFree(t->mel_start);
  cudaFree(t->mel_count); cudaFree(t->mel_off); cudaFree(t->mel_w);
  *t = FbankTables{};
}

void launch_fbank(const FbankTables &t, const float *samples, const long long *sample_off, const long long *sample_len,
                  const long long *frame_off, int n_utts, int max_frames, float *out, cudaStream_t st) {
  if (n_utts <= 0 || max_frames <= 0) return;
  // grid.x covers the longest utterance; CTAs past a shorter utterance's end exit immediately.
  dim3 grid((max_frames + kFramesPerCta - 1) / kFramesPerCta, n_utts);
  fbank_kernel<<<grid, kWarps * 32, 0, st>>>(samples, sample_off, sample_len, frame_off, t.window, t.twiddle, t.mel_start, t.mel_count,
                                             t.mel_off, t.mel_w, out);
  count_launch();
  KERNEL_CHECK();
}

}  // namespace b200asr
