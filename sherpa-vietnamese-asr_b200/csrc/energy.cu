// Energy scan for the chunk planner: per 10 ms frame, is the RMS under a threshold?
// Replaces the host NumPy pass of /root/reference core/asr_engine.py:521-553 (`find_silent_regions`, :532-536):
//     energies = np.sqrt(np.mean(frames ** 2, axis=1));  is_silent = energies < threshold        (float32 throughout)
// The comparison decides where a 30 s chunk is cut, so the flags have to be the reference's flags, not "close": the kernel
// reproduces NumPy's float32 arithmetic exactly - squares rounded to fp32, NumPy's pairwise summation of the 160 squares
// (two halves of 80; in each half 8 interleaved accumulators a[j], a[8+j], ... added in index order, then combined as
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7))), fp32 divide by 160, fp32 sqrt, fp32 compare. No FMA contraction anywhere.
// (tests/test_kernel_math.py checks this statement of the order against NumPy itself, bit for bit.)
//
// HBM-bound: 4 bytes read per sample, 1 byte written per 160 samples. One CTA stages 64 frames (40 KB) in shared memory with
// coalesced 128-bit loads; 16 lanes per frame then walk the 16 accumulator chains (frame stride padded to 168 floats so the
// two frames of a warp fall on disjoint banks) and fold them with xor-shuffles - fp32 addition is commutative, so the xor
// tree gives the same bits as NumPy's left-to-right pairing.
#include <cuda_runtime.h>

#include <cstdint>
#include <mutex>
#include <stdexcept>
#include <string>

#include "../../include/b200asr.h"
#include "common.cuh"

namespace b200asr {

void set_last_error(const std::string &msg);   // engine.cu

namespace {

constexpr int kFrame = 160;            // 10 ms at 16 kHz
constexpr int kFramesPerCta = 64;
constexpr int kThreads = 256;
constexpr int kStride = 168;           // padded frame stride in floats (168 mod 32 = 8)

__global__ void __launch_bounds__(kThreads) silent_frames_kernel(const float *__restrict__ pcm, long long n_frames, float threshold,
                                                                 unsigned char *__restrict__ quiet) {
  __shared__ __align__(16) float tile[kFramesPerCta * kStride];
  const long long n_tiles = (n_frames + kFramesPerCta - 1) / kFramesPerCta;
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const long long f0 = t * kFramesPerCta;
    const int nf = (int)min((long long)kFramesPerCta, n_frames - f0);
    const float4 *src = reinterpret_cast<const float4 *>(pcm + f0 * kFrame);    // frame starts are 640-byte multiples
    for (int i = threadIdx.x; i < nf * (kFrame / 4); i += kThreads) {
      const int fr = i / (kFrame / 4), q = i % (kFrame / 4);
      *reinterpret_cast<float4 *>(&tile[fr * kStride + q * 4]) = __ldg(src + i);
    }
    __syncthreads();
    // thread -> (frame, half h, accumulator j): 16 consecutive lanes own one frame
    for (int c = threadIdx.x; c < kFramesPerCta * 16; c += kThreads) {
      const int fr = c >> 4, h = (c >> 3) & 1, j = c & 7;
      float r = 0.f;
      if (fr < nf) {
        const float *a = &tile[fr * kStride + h * 80 + j];
        const float x0 = a[0];
        r = __fmul_rn(x0, x0);
#pragma unroll
        for (int i = 1; i < 10; ++i) {
          const float x = a[8 * i];
          r = __fadd_rn(r, __fmul_rn(x, x));
        }
      }
      r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
      r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
      r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
      r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 8));     // half 0 + half 1
      if ((c & 15) == 0 && fr < nf) {
        const float rms = __fsqrt_rn(__fdiv_rn(r, (float)kFrame));
        quiet[f0 + fr] = rms < threshold ? 1 : 0;
      }
    }
    __syncthreads();
  }
}

struct ScanBuffers {
  std::mutex mu;
  int device = -1;
  float *d_pcm = nullptr;
  unsigned char *d_quiet = nullptr;
  size_t cap_frames = 0;
  cudaStream_t st = nullptr;
  int n_sms = 0;
  void reserve(int dev, size_t frames) {
    if (dev != device || frames > cap_frames) {
      if (d_pcm) cudaFree(d_pcm);
      if (d_quiet) cudaFree(d_quiet);
      d_pcm = nullptr; d_quiet = nullptr; cap_frames = 0;
      if (st && dev != device) { cudaStreamDestroy(st); st = nullptr; }
      device = dev;
      const size_t want = frames + frames / 4 + 1024;
      CUDA_CHECK(cudaMalloc(&d_pcm, want * kFrame * sizeof(float)));
      CUDA_CHECK(cudaMalloc(&d_quiet, want));
      cap_frames = want;
    }
    if (!st) {
      CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
      CUDA_CHECK(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
    }
  }
};
ScanBuffers g_scan;

}  // namespace

// frames of a device-resident recording (the caller's stream); used by the C entry point below
void launch_silent_frames(const float *d_pcm, long long n_frames, float threshold, unsigned char *d_quiet, int n_sms, cudaStream_t st) {
  if (n_frames <= 0) return;
  const long long tiles = (n_frames + kFramesPerCta - 1) / kFramesPerCta;
  const int grid = (int)std::min<long long>(tiles, (long long)n_sms * 4);     // 4 resident CTAs of 43 KB per SM
  silent_frames_kernel<<<grid, kThreads, 0, st>>>(d_pcm, n_frames, threshold, d_quiet);
  KERNEL_CHECK();
}

}  // namespace b200asr

extern "C" int32_t B200AsrSilentFrames(const float *samples, int64_t n, int32_t sample_rate, float threshold, uint8_t *quiet,
                                       int32_t device_id) {
  using namespace b200asr;
  try {
    if (sample_rate != 16000) throw std::runtime_error("B200AsrSilentFrames: only 16000 Hz (160-sample frames) is supported");
    if (n < 0 || (n > 0 && !samples)) throw std::runtime_error("B200AsrSilentFrames: bad arguments");
    const long long n_frames = n / kFrame;
    if (n_frames > INT32_MAX) throw std::runtime_error("B200AsrSilentFrames: recording too long");
    if (n_frames == 0 || !quiet) return (int32_t)n_frames;
    std::lock_guard<std::mutex> lk(g_scan.mu);
    CUDA_CHECK(cudaSetDevice(device_id));
    g_scan.reserve(device_id, (size_t)n_frames);
    CUDA_CHECK(cudaMemcpyAsync(g_scan.d_pcm, samples, (size_t)n_frames * kFrame * sizeof(float), cudaMemcpyHostToDevice, g_scan.st));
    launch_silent_frames(g_scan.d_pcm, n_frames, threshold, g_scan.d_quiet, g_scan.n_sms, g_scan.st);
    CUDA_CHECK(cudaMemcpyAsync(quiet, g_scan.d_quiet, (size_t)n_frames, cudaMemcpyDeviceToHost, g_scan.st));
    CUDA_CHECK(cudaStreamSynchronize(g_scan.st));
    return (int32_t)n_frames;
  } catch (const std::exception &e) {
    set_last_error(e.what());
    return -1;
  } catch (...) {
    set_last_error("unknown error");
    return -1;
  }
}
