// Kaldi-compatible 80-bin log-mel fbank on sm_100a.
// Replaces kaldi-native-fbank as called by compute_fbank_ort (/root/reference core/asr_engine.py:698-721;
// options :703-713: 16 kHz, 25/10 ms, povey, dither 0, snip_edges False, 80 bins, 20-7600 Hz).
// Exact algorithm statement: SURVEY.md Appendix A.
//
// One CTA = 16 consecutive frames of one utterance. The 400+15*160 = 2800 PCM samples the tile needs are read
// from HBM once, coalesced, into shared memory (reflected at the utterance edges); each warp then owns two
// frames: DC removal (warp reduction), pre-emphasis, povey window, 512-point real FFT done as a 256-point
// complex radix-2 FFT in shared memory, power spectrum, sparse mel projection (each bin touches a contiguous
// run of FFT bins) and log. Output rows are written coalesced.
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

__device__ __forceinline__ int bitrev8(int v) { return (int)(__brev((unsigned)v) >> 24); }

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
  __shared__ float s_re[kWarps][kHalf];
  __shared__ float s_im[kWarps][kHalf];
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
    // pre-emphasis + window, packed as z[n] = w[2n] + i*w[2n+1] at bit-reversed positions
    for (int nidx = lane; nidx < kHalf; nidx += 32) {
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
      const int r = bitrev8(nidx);
      re[r] = v[0];
      im[r] = v[1];
    }
    __syncwarp();
    // 256-point complex FFT, radix-2 decimation in time, 8 stages, 128 butterflies per stage
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const int half = 1 << s;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int b = lane + 32 * q;
        const int j = b & (half - 1);
        const int i0 = ((b >> s) << (s + 1)) + j;
        const int i1 = i0 + half;
        const int tw = j * (kHalf >> s);  // exp(-2*pi*i*j/(2*half)) = table512[j*256/half]
        const float c = s_tw[2 * tw], sn = s_tw[2 * tw + 1];
        const float xr = re[i1], xi = im[i1];
        const float tr = xr * c - xi * sn;
        const float ti = xr * sn + xi * c;
        const float ur = re[i0], ui = im[i0];
        re[i0] = ur + tr;
        im[i0] = ui + ti;
        re[i1] = ur - tr;
        im[i1] = ui - ti;
      }
      __syncwarp();
    }
    // unpack to the real FFT of 512 points: X[k] = E[k] + W^k O[k]
    for (int k = lane; k < kHalf; k += 32) {
      const int km = (kHalf - k) & (kHalf - 1);
      const float zr = re[k], zi = im[k], yr = re[km], yi = -im[km];  // conj(Z[N-k])
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
