// Audio staging on the GPU: the sample-touching passes between the decoder and the recognizer -
// low-volume boost (/root/reference core/asr_engine.py:512-516), per-segment RMS normalisation
// (core/audio_preprocessing.py:46-155) and the peak limiter (:226-244), i.e. preprocess_audio (:251-292) - over the uploaded
// PCM instead of three NumPy passes over every sample on the host.
//   1. per-segment sum of squares: one CTA per VAD segment, float64 accumulation (the reference's float32 pairwise mean is
//      reproduced to ~1e-7 relative, so gains agree to the last bits but not always bit for bit);
//   2. host: median target, clamped gains, and the <= 80-sample linear fades at segment edges exactly as the reference's
//      sequential loop makes them (np.linspace in float64, cast to float32) - a few hundred numbers, kept sparse;
//   3. y = x * gain: piecewise-constant gains found by binary search over the breakpoints, then the fade samples;
//   4. |y| maximum (atomic on the float bits), host decides the limiter scale 0.95 / peak in float32 as NumPy does, y *= scale.
// Multiplications are single float32 roundings, so everything except step 1 is bit-equal to the NumPy code.
#include <math.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <vector>

#include "../../include/b200asr.h"
#include "common.cuh"

namespace b200asr {
void set_last_error(const std::string &msg);   // engine.cu

namespace {

__global__ void __launch_bounds__(256) segment_sumsq_kernel(const float *__restrict__ x, const long long *__restrict__ seg_s,
                                                            const long long *__restrict__ seg_e, double *__restrict__ out) {
  const int g = blockIdx.x;
  const long long s = seg_s[g], e = seg_e[g];
  double acc = 0.0;
  for (long long i = s + threadIdx.x; i < e; i += 256) {
    const float v = __ldg(x + i);
    acc += (double)(v * v);          // the square is rounded to float32 first, as `segment ** 2` is
  }
  __shared__ double sh[256];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[g] = sh[0];
}

// y[i] = x[i] * pre * gain(i): gain = value of the last breakpoint <= i (bp[0] = 0)
__global__ void apply_gain_kernel(const float *__restrict__ x, float *__restrict__ y, long long n, const long long *__restrict__ bp,
                                  const float *__restrict__ bg, int n_bp, int has_pre, float pre_div, float pre_mul) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = x[i];
  if (has_pre) { v = v / pre_div; v = v * pre_mul; }       // audio / peak * 0.95, two float32 roundings (core/asr_engine.py:515)
  if (n_bp > 0) {
    int lo = 0, hi = n_bp - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (__ldg(bp + mid) <= i) lo = mid; else hi = mid - 1;
    }
    v = v * __ldg(bg + lo);
  }
  y[i] = v;
}

__global__ void apply_overrides_kernel(const float *__restrict__ x, float *__restrict__ y, const long long *__restrict__ idx,
                                       const float *__restrict__ val, int n_over, int has_pre, float pre_div, float pre_mul) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_over) return;
  float v = x[idx[j]];
  if (has_pre) { v = v / pre_div; v = v * pre_mul; }
  y[idx[j]] = v * val[j];
}

__global__ void absmax_kernel(const float *__restrict__ y, long long n, unsigned int *__restrict__ out) {
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(y[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));      // non-negative floats order like their bit patterns
}

__global__ void scale_kernel(float *__restrict__ y, long long n, float scale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = y[i] * scale;
}

struct StageBuffers {
  std::mutex mu;
  int device = -1;
  float *d_x = nullptr, *d_y = nullptr; size_t cap = 0;
  long long *d_ll = nullptr; size_t ll_cap = 0;
  float *d_f = nullptr; size_t f_cap = 0;
  double *d_sums = nullptr; size_t sums_cap = 0;
  unsigned int *d_peak = nullptr;
  cudaStream_t st = nullptr;
  int n_sms = 148;
  void reserve(int dev, size_t n) {
    if (dev != device) {
      cudaFree(d_x); cudaFree(d_y); cudaFree(d_ll); cudaFree(d_f); cudaFree(d_sums); cudaFree(d_peak);
      d_x = d_y = d_f = nullptr; d_ll = nullptr; d_sums = nullptr; d_peak = nullptr;
      cap = ll_cap = f_cap = sums_cap = 0;
      if (st) { cudaStreamDestroy(st); st = nullptr; }
      device = dev;
    }
    if (!st) {
      CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
      CUDA_CHECK(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
      CUDA_CHECK(cudaMalloc(&d_peak, sizeof(unsigned int)));
    }
    if (n > cap) {
      cudaFree(d_x); cudaFree(d_y);
      cap = n + n / 4 + 1024;
      CUDA_CHECK(cudaMalloc(&d_x, cap * sizeof(float)));
      CUDA_CHECK(cudaMalloc(&d_y, cap * sizeof(float)));
    }
  }
  template <typename T>
  static void grow(T *&p, size_t &c, size_t need) {
    if (need <= c) return;
    cudaFree(p);
    c = need + need / 2 + 64;
    CUDA_CHECK(cudaMalloc(&p, c * sizeof(T)));
  }
};
StageBuffers g_stage;

// np.linspace(a, b, k, dtype=float32): float64 arithmetic, last element = b exactly, cast to float32
void linspace_f32(float a, float b, int k, float *out) {
  if (k <= 0) return;
  if (k == 1) { out[0] = a; return; }
  const double step = ((double)b - (double)a) / (double)(k - 1);
  for (int j = 0; j < k; ++j) {
    volatile double t = (double)j * step;       // mul then add, as NumPy's y * step + start
    out[j] = (float)(t + (double)a);
  }
  out[k - 1] = b;
}

}  // namespace
}  // namespace b200asr

using namespace b200asr;

extern "C" int32_t B200AsrPreprocessAudio(const float *samples, int64_t n, const int64_t *seg_starts, const int64_t *seg_ends, int32_t n_seg,
                                          int32_t enable_rms_normalize, int32_t boost_low_volume, int32_t sample_rate, float *out,
                                          int32_t device_id) {
  try {
    if (n < 0 || (n > 0 && (!samples || !out)) || n_seg < 0 || (n_seg > 0 && (!seg_starts || !seg_ends)))
      throw std::runtime_error("B200AsrPreprocessAudio: bad arguments");
    if (n == 0) return 0;
    for (int i = 0; i < n_seg; ++i)
      if (seg_starts[i] < 0 || seg_ends[i] > n || seg_starts[i] > seg_ends[i]) throw std::runtime_error("B200AsrPreprocessAudio: segment out of range");
    std::lock_guard<std::mutex> lk(g_stage.mu);
    CUDA_CHECK(cudaSetDevice(device_id));
    StageBuffers &S = g_stage;
    S.reserve(device_id, (size_t)n);
    cudaStream_t st = S.st;
    CUDA_CHECK(cudaMemcpyAsync(S.d_x, samples, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, st));
    const unsigned nblk = (unsigned)((n + 255) / 256);
    const int red_grid = std::min<long long>((n + 1023) / 1024, (long long)S.n_sms * 8);
    // ---- low-volume boost of the load step: 0 < peak < 0.5 -> audio / peak * 0.95
    int has_pre = 0;
    float pre_div = 1.f, pre_mul = 1.f;
    if (boost_low_volume) {
      CUDA_CHECK(cudaMemsetAsync(S.d_peak, 0, sizeof(unsigned int), st));
      absmax_kernel<<<red_grid, 256, 0, st>>>(S.d_x, n, S.d_peak);
      KERNEL_CHECK();
      float peak = 0.f;
      CUDA_CHECK(cudaMemcpyAsync(&peak, S.d_peak, sizeof(float), cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaStreamSynchronize(st));
      if (peak > 0.f && peak < 0.5f) { has_pre = 1; pre_div = peak; pre_mul = 0.95f; }
    }
    // ---- per-segment RMS normalisation
    std::vector<long long> bp;     // breakpoints of the piecewise-constant gain
    std::vector<float> bg;
    std::vector<long long> o_idx;  // fade samples
    std::vector<float> o_val;
    if (enable_rms_normalize && n_seg > 0) {
      const long long min_samples = (long long)(100.0 * sample_rate / 1000.0);
      std::vector<long long> ms, me;
      for (int i = 0; i < n_seg; ++i)
        if (seg_ends[i] - seg_starts[i] >= min_samples) { ms.push_back(seg_starts[i]); me.push_back(seg_ends[i]); }
      const int nm = (int)ms.size();
      if (nm > 0) {
        if (has_pre) {   // the RMS is taken on the boosted audio: materialise it first
          apply_gain_kernel<<<nblk, 256, 0, st>>>(S.d_x, S.d_y, n, nullptr, nullptr, 0, 1, pre_div, pre_mul);
          KERNEL_CHECK();
          std::swap(S.d_x, S.d_y);
          has_pre = 0;
        }
        StageBuffers::grow(S.d_ll, S.ll_cap, (size_t)2 * nm + 4096);
        StageBuffers::grow(S.d_sums, S.sums_cap, (size_t)nm);
        CUDA_CHECK(cudaMemcpyAsync(S.d_ll, ms.data(), nm * sizeof(long long), cudaMemcpyHostToDevice, st));
        CUDA_CHECK(cudaMemcpyAsync(S.d_ll + nm, me.data(), nm * sizeof(long long), cudaMemcpyHostToDevice, st));
        segment_sumsq_kernel<<<nm, 256, 0, st>>>(S.d_x, S.d_ll, S.d_ll + nm, S.d_sums);
        KERNEL_CHECK();
        std::vector<double> sums(nm);
        CUDA_CHECK(cudaMemcpyAsync(sums.data(), S.d_sums, nm * sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        // rms = float(np.sqrt(np.mean(segment ** 2))): float32 mean, float32 sqrt
        struct Seg { long long s, e; double rms; float g; };
        std::vector<Seg> segs;
        for (int i = 0; i < nm; ++i) {
          const float mean = (float)(sums[i] / (double)(me[i] - ms[i]));
          const double rms = (double)sqrtf(mean);
          if (rms > 1e-8) segs.push_back(Seg{ms[i], me[i], rms, 1.f});
        }
        if (!segs.empty()) {
          std::vector<double> r;
          for (auto &sg : segs) r.push_back(sg.rms);
          std::sort(r.begin(), r.end());
          const double target = (r.size() & 1) ? r[r.size() / 2] : 0.5 * (r[r.size() / 2 - 1] + r[r.size() / 2]);
          if (target >= 1e-8) {
            const double hi = pow(10.0, 20.0 / 20.0);
            for (auto &sg : segs) sg.g = (float)std::max(std::min(target / sg.rms, hi), 1.0 / hi);   // stored into a float32 map
            // base curve: assignments in order, the last segment covering a sample wins
            std::vector<long long> cuts{0, n};
            for (auto &sg : segs) { cuts.push_back(sg.s); cuts.push_back(sg.e); }
            std::sort(cuts.begin(), cuts.end());
            cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
            for (size_t c = 0; c + 1 < cuts.size(); ++c) {
              float g = 1.0f;
              for (auto &sg : segs) if (sg.s <= cuts[c] && cuts[c] < sg.e) g = sg.g;
              bp.push_back(cuts[c]);
              bg.push_back(g);
            }
            std::map<long long, float> over;
            auto get = [&](long long i) -> float {
              auto it = over.find(i);
              if (it != over.end()) return it->second;
              const size_t k = std::upper_bound(bp.begin(), bp.end(), i) - bp.begin() - 1;
              return bg[k];
            };
            const int fade = (int)(5.0 * sample_rate / 1000.0);
            std::vector<float> ramp;
            for (auto &sg : segs) {
              const int k = (int)std::min<long long>(fade, (sg.e - sg.s) / 4);
              if (k <= 0) continue;
              ramp.resize(k);
              if (sg.s > 0) {
                linspace_f32(get(sg.s - 1), get(sg.s), k, ramp.data());
                for (int j = 0; j < k; ++j) over[sg.s + j] = ramp[j];
              }
              if (sg.e < n) {
                linspace_f32(get(sg.e - 1), get(std::min<long long>(n - 1, sg.e)), k, ramp.data());
                for (int j = 0; j < k; ++j) over[sg.e - k + j] = ramp[j];
              }
            }
            for (auto &kv : over) { o_idx.push_back(kv.first); o_val.push_back(kv.second); }
          } else {
            bp.clear(); bg.clear();
          }
        }
      }
    }
    // ---- y = (boosted) x * gain
    const int n_bp = (int)bp.size(), n_over = (int)o_idx.size();
    if (n_bp > 0 || n_over > 0) {
      StageBuffers::grow(S.d_ll, S.ll_cap, (size_t)n_bp + n_over + 8);
      StageBuffers::grow(S.d_f, S.f_cap, (size_t)n_bp + n_over + 8);
      if (n_bp) {
        CUDA_CHECK(cudaMemcpyAsync(S.d_ll, bp.data(), n_bp * sizeof(long long), cudaMemcpyHostToDevice, st));
        CUDA_CHECK(cudaMemcpyAsync(S.d_f, bg.data(), n_bp * sizeof(float), cudaMemcpyHostToDevice, st));
      }
      if (n_over) {
        CUDA_CHECK(cudaMemcpyAsync(S.d_ll + n_bp, o_idx.data(), n_over * sizeof(long long), cudaMemcpyHostToDevice, st));
        CUDA_CHECK(cudaMemcpyAsync(S.d_f + n_bp, o_val.data(), n_over * sizeof(float), cudaMemcpyHostToDevice, st));
      }
    }
    apply_gain_kernel<<<nblk, 256, 0, st>>>(S.d_x, S.d_y, n, S.d_ll, S.d_f, n_bp, has_pre, pre_div, pre_mul);
    KERNEL_CHECK();
    if (n_over) {
      apply_overrides_kernel<<<(n_over + 255) / 256, 256, 0, st>>>(S.d_x, S.d_y, S.d_ll + n_bp, S.d_f + n_bp, n_over, has_pre, pre_div, pre_mul);
      KERNEL_CHECK();
    }
    // ---- peak limiter: peak > 0.95 -> * (0.95 / peak), the ratio taken in float32
    CUDA_CHECK(cudaMemsetAsync(S.d_peak, 0, sizeof(unsigned int), st));
    absmax_kernel<<<red_grid, 256, 0, st>>>(S.d_y, n, S.d_peak);
    KERNEL_CHECK();
    float peak = 0.f;
    CUDA_CHECK(cudaMemcpyAsync(&peak, S.d_peak, sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    if (peak > 0.95f) {
      scale_kernel<<<nblk, 256, 0, st>>>(S.d_y, n, 0.95f / peak);
      KERNEL_CHECK();
    }
    CUDA_CHECK(cudaMemcpyAsync(out, S.d_y, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    return 0;
  } catch (const std::exception &e) {
    set_last_error(e.what());
    return -1;
  }
}
