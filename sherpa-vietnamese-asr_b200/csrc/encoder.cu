// CUDA-core kernels of the Zipformer2 encoder (everything that is not a dense Linear):
// Conv2dSubsampling front-end, BiasNorm, Bypass, SimpleDownsample/Upsample, relative-position attention
// weights, attention application, and the convolution module's GLU + depthwise conv + SwooshR.
// The reference runs this arithmetic inside the encoder ONNX graph (/root/reference core/asr_engine.py:1045-1049);
// the architecture statement followed here is SURVEY.md Appendix B (icefall Zipformer2, inference only).
//
// Batch layout: ragged, packed. Every activation is [rows, C] row-major fp32 where rows = sum over utterances
// of that utterance's own frame count at the stack's rate; per-utterance lengths/offsets (RaggedDesc) make
// every time-dependent op stop at each utterance's own end (App. B.6), which is what the reference's batch-1
// execution computes.
#include <math.h>

#include <algorithm>

#include "common.cuh"

namespace b200asr {

namespace {

__device__ __forceinline__ float softplus_f(float x) { return softplus_sfu(x); }
__device__ __forceinline__ float swoosh_r(float v) { return softplus_f(v - 1.0f) - 0.08f * v - 0.313261687f; }
__device__ __forceinline__ float sigmoid_f(float x) { return sigmoid_sfu(x); }

// ------------------------------------------------------------------ Conv2dSubsampling (App. B.2)
// last u with off[u] <= row (off ascending, off[n] = total)
template <typename T>
__device__ __forceinline__ int find_utt(const T *__restrict__ off, int n, long long row) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if ((long long)__ldg(off + mid) <= row) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// conv0: [T,80] -> [(T-2),80,8] channels-last, k3, pad (0,1), SwooshR.   w: [3][3][8] (kh,kw,co)
// flat over the packed output pixels (row, f) of the whole ragged batch; thread = one pixel, 8 channels
__global__ void embed_conv0_kernel(const float *__restrict__ feats, const long long *__restrict__ foff,
                                   const long long *__restrict__ ooff, int n_utt, const float *__restrict__ w,
                                   const float *__restrict__ b, float *__restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ooff[n_utt] * 80) return;
  const long long row = idx / 80;
  const int f = (int)(idx - row * 80);
  const int u = find_utt(ooff, n_utt, row);
  const int t = (int)(row - __ldg(ooff + u));
  const float *x = feats + __ldg(foff + u) * 80;
  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = __ldg(b + c);
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int ff = f + kw - 1;
      const float v = (ff >= 0 && ff < 80) ? __ldg(x + (long long)(t + kh) * 80 + ff) : 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = fmaf(v, __ldg(w + (kh * 3 + kw) * 8 + c), acc[c]);
    }
  float4 *o = reinterpret_cast<float4 *>(out + idx * 8);
  o[0] = make_float4(swoosh_r(acc[0]), swoosh_r(acc[1]), swoosh_r(acc[2]), swoosh_r(acc[3]));
  o[1] = make_float4(swoosh_r(acc[4]), swoosh_r(acc[5]), swoosh_r(acc[6]), swoosh_r(acc[7]));
}

// conv1: [(T-2),80,8] -> [t2,39,32], k3 stride 2, SwooshR.   w: [3][3][8][32]
// Persistent CTAs (the 9 KB of weights are staged in shared memory once per CTA) walk the packed output pixels of
// the whole ragged batch; thread -> one output pixel (all 32 channels); the utterance of a pixel is found by a binary
// search over the packed row offsets.
__global__ void __launch_bounds__(256) embed_conv1_kernel(const float *__restrict__ in, const long long *__restrict__ ioff,
                                                          const long long *__restrict__ ooff, int n_utt,
                                                          const float *__restrict__ w, const float *__restrict__ b,
                                                          float *__restrict__ out) {
  __shared__ __align__(16) float sw[72 * 32];
  __shared__ float sb[32];
  for (int i = threadIdx.x; i < 72 * 32; i += blockDim.x) sw[i] = w[i];
  if (threadIdx.x < 32) sb[threadIdx.x] = b[threadIdx.x];
  __syncthreads();
  const long long total = ooff[n_utt] * 39;         // output pixels
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += (long long)gridDim.x * blockDim.x) {
    const long long row = pix / 39;                 // packed output row (utterance, t)
    const int f = (int)(pix - row * 39);
    int lo = 0, hi = n_utt - 1;
    while (lo < hi) {   // last u with ooff[u] <= row
      const int mid = (lo + hi + 1) >> 1;
      if (__ldg(ooff + mid) <= row) lo = mid; else hi = mid - 1;
    }
    const int t = (int)(row - __ldg(ooff + lo));
    const float *x = in + __ldg(ioff + lo) * 80 * 8;
    // one thread = one output pixel, all 32 output channels: each of the 72 input values is loaded once and feeds 32 FMAs
    float acc[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = sb[c];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const float *px = x + ((long long)(2 * t + kh) * 80 + (2 * f + kw)) * 8;
        const float4 v0 = __ldg(reinterpret_cast<const float4 *>(px));
        const float4 v1 = __ldg(reinterpret_cast<const float4 *>(px + 4));
        const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) {
          const float4 *wr = reinterpret_cast<const float4 *>(sw + ((kh * 3 + kw) * 8 + ci) * 32);   // same address in the whole warp
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const float4 w4 = wr[c4];
            acc[4 * c4] = fmaf(v[ci], w4.x, acc[4 * c4]); acc[4 * c4 + 1] = fmaf(v[ci], w4.y, acc[4 * c4 + 1]);
            acc[4 * c4 + 2] = fmaf(v[ci], w4.z, acc[4 * c4 + 2]); acc[4 * c4 + 3] = fmaf(v[ci], w4.w, acc[4 * c4 + 3]);
          }
        }
      }
    float4 *o = reinterpret_cast<float4 *>(out + pix * 32);
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4)
      o[c4] = make_float4(swoosh_r(acc[4 * c4]), swoosh_r(acc[4 * c4 + 1]), swoosh_r(acc[4 * c4 + 2]), swoosh_r(acc[4 * c4 + 3]));
  }
}

// conv2 ([t2,39,32] -> [T1,19,128], k3 stride (1,2), SwooshR) as a GEMM: im2col rows [pixel (t,f)][(kh,kw,ci)] = 9 contiguous 128-byte chunks of the channels-last
// conv1 output, so both the gather reads and the row writes are fully coalesced; the 288 -> 128 contraction then
// runs on the tensor pipe (gemm_tc.cu) with bias + SwooshR in its epilogue.
// Flat over the packed output pixels; thread = (pixel, one float4 of the 32-channel chunk) walks the 9 taps.
__global__ void embed_im2col2_kernel(const float *__restrict__ in, const long long *__restrict__ ioff, const int *__restrict__ ooff,
                                     int n_utt, float *__restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n_pix = (long long)ooff[n_utt] * 19;
  if (idx >= n_pix * 8) return;
  const int q = (int)(idx & 7);            // float4 within the 32-channel chunk
  const long long pix = idx >> 3;          // packed output pixel
  const long long row = pix / 19;
  const int f = (int)(pix - row * 19);
  const int u = find_utt(ooff, n_utt, row);
  const int t = (int)(row - __ldg(ooff + u));
  const float *src = in + ((__ldg(ioff + u) + t) * 39 + 2 * f) * 32;
  float4 v[9];
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) v[kh * 3 + kw] = __ldg(reinterpret_cast<const float4 *>(src + ((long long)kh * 39 + kw) * 32) + q);
  float4 *dst = reinterpret_cast<float4 *>(out + pix * 288) + q;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) dst[tap * 8] = v[tap];
}

// ConvNeXt depthwise 7x7 over (time, freq) of [T1,19,128], zero padded at each utterance's own edges. w: [7][7][128]
// CTA = 8 time rows x 19 freq x 32 channels; the (8+6) x (19+6) x 32 zero-padded patch sits in shared memory;
// thread = (channel, time row) keeps its 49 taps in registers and slides over frequency.
constexpr int kDw7Rows = 8, kDw7Ch = 32;
__global__ void __launch_bounds__(256) embed_dw7_kernel(const float *__restrict__ in, const int *__restrict__ len,
                                                        const int *__restrict__ off, const int *__restrict__ tile_off, int n_utt,
                                                        const float *__restrict__ w, const float *__restrict__ b,
                                                        float *__restrict__ out) {
  __shared__ __align__(16) float patch[(kDw7Rows + 6) * 25 * kDw7Ch];
  // blockIdx.x = (128-frame tile of the ragged batch) * 16 + 8-row sub-tile
  const int tile = blockIdx.x >> 4;
  const int u = find_utt(tile_off, n_utt, tile);
  const int T1 = len[u];
  const int t0 = (tile - __ldg(tile_off + u)) * 128 + (blockIdx.x & 15) * kDw7Rows;
  if (t0 >= T1) return;
  const int c0 = blockIdx.y * kDw7Ch;
  const float *x = in + (long long)off[u] * 19 * 128;
  for (int i = threadIdx.x; i < (kDw7Rows + 6) * 25 * (kDw7Ch / 4); i += 256) {
    const int c4 = (i % (kDw7Ch / 4)) * 4, ff = (i / (kDw7Ch / 4)) % 25 - 3, tt = t0 + i / ((kDw7Ch / 4) * 25) - 3;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tt >= 0 && tt < T1 && ff >= 0 && ff < 19) v = __ldg(reinterpret_cast<const float4 *>(x + ((long long)tt * 19 + ff) * 128 + c0 + c4));
    reinterpret_cast<float4 *>(patch)[i] = v;
  }
  __syncthreads();
  const int c = threadIdx.x % kDw7Ch, tr = threadIdx.x / kDw7Ch;
  if (t0 + tr >= T1) return;
  float wr[49];
#pragma unroll
  for (int k = 0; k < 49; ++k) wr[k] = __ldg(w + k * 128 + c0 + c);
  const float bias = __ldg(b + c0 + c);
  float *o = out + ((long long)off[u] + t0 + tr) * 19 * 128 + c0 + c;
  // per kernel row a register window slides over the frequency axis (two halves of 10 + 9 outputs keep it small), so a
  // patch value is read from shared memory once per kernel row instead of once per tap
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int fb = half * 10;
    float acc[10];
#pragma unroll
    for (int f = 0; f < 10; ++f) acc[f] = bias;
#pragma unroll
    for (int dt = 0; dt < 7; ++dt) {
      float row[16];
#pragma unroll
      for (int ff = 0; ff < 16; ++ff) row[ff] = (fb + ff < 25) ? patch[((tr + dt) * 25 + fb + ff) * kDw7Ch + c] : 0.f;
#pragma unroll
      for (int f = 0; f < 10; ++f)
#pragma unroll
        for (int df = 0; df < 7; ++df) acc[f] = fmaf(wr[dt * 7 + df], row[f + df], acc[f]);
    }
#pragma unroll
    for (int f = 0; f < 10; ++f)
      if (fb + f < 19) o[(fb + f) * 128] = acc[f];
  }
}

// ------------------------------------------------------------------ BiasNorm / Bypass
// one warp per row: x * (mean((x-bias)^2))^-0.5 * exp(log_scale); optional bypass against `orig`
__global__ void biasnorm_kernel(const float *__restrict__ x, const float *__restrict__ orig, int M, int D,
                                const float *__restrict__ bias, const float *__restrict__ log_scale,
                                const float *__restrict__ bypass, float *__restrict__ out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float *xr = x + (long long)row * D;
  float ss = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float d = xr[c] - __ldg(bias + c);
    ss = fmaf(d, d, ss);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float scale = (1.0f / sqrtf(ss / (float)D)) * expf(__ldg(log_scale));
  float *orow = out + (long long)row * D;
  if (orig) {
    const float *og = orig + (long long)row * D;
    for (int c = lane; c < D; c += 32) {
      const float o0 = og[c];
      orow[c] = o0 + (xr[c] * scale - o0) * __ldg(bypass + c);
    }
  } else {
    for (int c = lane; c < D; c += 32) orow[c] = xr[c] * scale;
  }
}

// The elementwise kernels below run flat over the packed rows (no per-utterance grid dimension: a ragged batch would
// launch mostly empty CTAs) with 128-bit accesses; every channel count on the path is a multiple of 4.
__global__ void bypass_kernel(const float4 *__restrict__ x, const float4 *__restrict__ orig, long long total4, int D4,
                              const float4 *__restrict__ scale, float4 *__restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const float4 o0 = orig[i], v = x[i], sc = __ldg(scale + (int)(i % D4));
  out[i] = make_float4(o0.x + (v.x - o0.x) * sc.x, o0.y + (v.y - o0.y) * sc.y, o0.z + (v.z - o0.z) * sc.z, o0.w + (v.w - o0.w) * sc.w);
}

__global__ void convert_channels_kernel(const float4 *__restrict__ in, int Cin4, float4 *__restrict__ out, int Cout4, long long M) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * Cout4) return;
  const long long m = i / Cout4;
  const int c = (int)(i - m * Cout4);
  out[i] = c < Cin4 ? in[m * Cin4 + c] : make_float4(0.f, 0.f, 0.f, 0.f);
}

// ------------------------------------------------------------------ SimpleDownsample / SimpleUpsample (App. B.3)
// Row maps between the full frame rate and a stack's rate, built once per batch: up[R] = low-rate row of full-rate
// row R; down[Rq] = {first full-rate row feeding low-rate row Rq, last row of that utterance} (the right padding
// repeats the utterance's own last frame).
__global__ void build_row_maps_kernel(const int *__restrict__ len1, const int *__restrict__ off1, const int *__restrict__ lenq,
                                      const int *__restrict__ offq, int ds, int *__restrict__ up, int2 *__restrict__ down) {
  const int u = blockIdx.x;
  const int L1 = len1[u], o1 = off1[u], Lq = lenq[u], oq = offq[u];
  for (int t = threadIdx.x; t < L1; t += blockDim.x) up[o1 + t] = oq + t / ds;
  for (int t = threadIdx.x; t < Lq; t += blockDim.x) down[oq + t] = make_int2(o1 + t * ds, o1 + L1 - 1);
}

template <int DS>
__global__ void downsample_kernel(const float4 *__restrict__ in, const int2 *__restrict__ down, long long total4, int C4,
                                  const float *__restrict__ bias, float4 *__restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total4) return;
  const int r = (int)(idx / C4), c = (int)(idx - (long long)r * C4);
  // softmax over the ds bias entries
  float wk[DS], mx = -INFINITY, den = 0.f;
#pragma unroll
  for (int k = 0; k < DS; ++k) mx = fmaxf(mx, __ldg(bias + k));
#pragma unroll
  for (int k = 0; k < DS; ++k) { wk[k] = expf(__ldg(bias + k) - mx); den += wk[k]; }
  const int2 fl = __ldg(down + r);
  float4 v[DS];
#pragma unroll
  for (int k = 0; k < DS; ++k) v[k] = in[(long long)min(fl.x + k, fl.y) * C4 + c];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < DS; ++k) {
    const float w = wk[k] / den;
    acc.x += v[k].x * w; acc.y += v[k].y * w; acc.z += v[k].z * w; acc.w += v[k].w * w;
  }
  out[idx] = acc;
}

__global__ void upsample_combine_kernel(const float4 *__restrict__ y, const int *__restrict__ up, const float4 *__restrict__ orig,
                                        long long total4, int C4, const float4 *__restrict__ scale, float4 *__restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total4) return;
  const int r = (int)(idx / C4), c = (int)(idx - (long long)r * C4);
  const float4 o0 = orig[idx], v = y[(long long)__ldg(up + r) * C4 + c], sc = __ldg(scale + c);
  out[idx] = make_float4(o0.x + (v.x - o0.x) * sc.x, o0.y + (v.y - o0.y) * sc.y, o0.z + (v.z - o0.z) * sc.z, o0.w + (v.w - o0.w) * sc.w);
}

struct ConcatArgs {
  const float *src[4];
  int ld[4], c0[4], c1[4];
  int n;
};
__global__ void concat_downsample2_kernel(ConcatArgs a, const int2 *__restrict__ down, long long total4, int C4,
                                          const float *__restrict__ bias, float4 *__restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total4) return;
  const int r = (int)(idx / C4), c = (int)(idx - (long long)r * C4) * 4;
  const float b0 = __ldg(bias), b1 = __ldg(bias + 1);
  const float mx = fmaxf(b0, b1);
  const float e0 = expf(b0 - mx), e1 = expf(b1 - mx);
  const float den = e0 + e1;
  const float *src = nullptr;
  int ld = 0;
  for (int p = 0; p < a.n; ++p)
    if (c >= a.c0[p] && c < a.c1[p]) { src = a.src[p]; ld = a.ld[p]; }   // piece boundaries are multiples of 4
  const int2 fl = __ldg(down + r);
  const float4 v0 = *reinterpret_cast<const float4 *>(src + (long long)fl.x * ld + c);
  const float4 v1 = *reinterpret_cast<const float4 *>(src + (long long)min(fl.x + 1, fl.y) * ld + c);
  const float w0 = e0 / den, w1 = e1 / den;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  acc.x += v0.x * w0; acc.y += v0.y * w0; acc.z += v0.z * w0; acc.w += v0.w * w0;
  acc.x += v1.x * w1; acc.y += v1.y * w1; acc.z += v1.z * w1; acc.w += v1.w * w1;
  out[idx] = acc;
}

// ------------------------------------------------------------------ CompactRelPositionalEncoding
__global__ void pos_emb_kernel(float *__restrict__ pe, int L, int pos_dim) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = pos_dim / 2;
  if (idx >= (2 * L - 1) * half) return;
  const int r = idx / half, j = idx % half;
  const float x = (float)(r - (L - 1));
  const float cl = (float)sqrt((double)pos_dim);
  const float sgn = (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f);
  const float xc = cl * sgn * (logf(fabsf(x) + cl) - (float)log(sqrt((double)pos_dim)));
  const float length_scale = (float)((double)pos_dim / (2.0 * M_PI));
  const float th = atanf(xc / length_scale);
  const float fr = (float)(j + 1);
  pe[(long long)r * pos_dim + 2 * j] = cosf(th * fr);
  pe[(long long)r * pos_dim + 2 * j + 1] = (2 * j + 1 == pos_dim - 1) ? 1.0f : sinf(th * fr);
}

// ------------------------------------------------------------------ attention weights (scores + rel-pos + softmax)
// CTA = (16 query rows, head, utterance), 256 threads. Keys are processed in tiles of 256: the K tile sits
// transposed in shared memory ([d][key]) so each thread computes a 4 (rows) x 4 (keys) register tile from two
// 128-bit shared loads per d; the relative-position term needs only the 7 pos rows (j-i) in [-3, 3] of that tile.
// All 16 x Tk scores stay in shared memory for the softmax; the normalised weights are written once: A[u][h][i][j].
constexpr int kAttRows = 16, kAttKeys = 256, kAttKStride = kAttKeys + 4;
template <int QD, int PD>
__global__ void __launch_bounds__(256) attn_weights_kernel(const float *__restrict__ proj, int ldp, const float *__restrict__ pos,
                                                           int ldpos, const int *__restrict__ len, const int *__restrict__ off,
                                                           const long long *__restrict__ aoff, int H, int Lmax,
                                                           float *__restrict__ A) {
  static_assert(QD == 32 && PD == 4, "tile mapping assumes 32/4");
  extern __shared__ __align__(16) float smem[];
  const int u = blockIdx.z, h = blockIdx.y;
  const int Tk = len[u];
  const int i0 = blockIdx.x * kAttRows;
  if (i0 >= Tk) return;
  float *sQT = smem;                          // [QD][16]   q transposed
  float *sP = sQT + QD * kAttRows;            // [16][PD]
  float *sKT = sP + kAttRows * PD;            // [QD][kAttKStride]
  float *sc = sKT + QD * kAttKStride;         // [16][Tk]
  const float *base = proj + (long long)off[u] * ldp;
  const int nrow = min(kAttRows, Tk - i0);
  const int tid = threadIdx.x;
  for (int i = tid; i < kAttRows * QD; i += 256) {
    const int r = i / QD, d = i % QD;
    sQT[d * kAttRows + r] = r < nrow ? base[(long long)(i0 + r) * ldp + h * QD + d] : 0.f;
  }
  for (int i = tid; i < kAttRows * PD; i += 256) {
    const int r = i / PD, d = i % PD;
    sP[i] = r < nrow ? base[(long long)(i0 + r) * ldp + 2 * H * QD + h * PD + d] : 0.f;
  }
  const int kg = tid & 63, rg = tid >> 6;     // 4 keys kg*4.., 4 rows rg*4..
  for (int j0 = 0; j0 < Tk; j0 += kAttKeys) {
    __syncthreads();                          // previous tile consumed (and sQT/sP visible on the first pass)
    // stage K tile transposed: thread -> (key, 4 consecutive d)
    for (int i = tid; i < kAttKeys * (QD / 4); i += 256) {
      const int key = i >> 3, dq = (i & 7) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j0 + key < Tk) v = __ldg(reinterpret_cast<const float4 *>(base + (long long)(j0 + key) * ldp + H * QD + h * QD + dq));
      sKT[(dq + 0) * kAttKStride + key] = v.x; sKT[(dq + 1) * kAttKStride + key] = v.y;
      sKT[(dq + 2) * kAttKStride + key] = v.z; sKT[(dq + 3) * kAttKStride + key] = v.w;
    }
    __syncthreads();
    if (j0 + kg * 4 < Tk) {
      float acc[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
#pragma unroll 8
      for (int d = 0; d < QD; ++d) {
        const float4 q4 = *reinterpret_cast<const float4 *>(sQT + d * kAttRows + rg * 4);
        const float4 k4 = *reinterpret_cast<const float4 *>(sKT + d * kAttKStride + kg * 4);
        const float q[4] = {q4.x, q4.y, q4.z, q4.w}, k[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(q[r], k[c], acc[r][c]);
      }
      // relative position: pos row for offset (j - i) is (j - i) + (Lmax - 1); (c - r) spans [-3, 3]
      const int pbase = (j0 + kg * 4) - (i0 + rg * 4) + (Lmax - 1);
      float4 pr[7];
#pragma unroll
      for (int o = 0; o < 7; ++o) {
        const int row = min(max(pbase + o - 3, 0), 2 * Lmax - 2);
        pr[o] = __ldg(reinterpret_cast<const float4 *>(pos + (long long)row * ldpos + h * PD));
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float4 p4 = *reinterpret_cast<const float4 *>(sP + (rg * 4 + r) * PD);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 w = pr[c - r + 3];
          float ps = p4.x * w.x;
          ps = fmaf(p4.y, w.y, ps); ps = fmaf(p4.z, w.z, ps); ps = fmaf(p4.w, w.w, ps);
          const int j = j0 + kg * 4 + c;
          if (j < Tk) sc[(rg * 4 + r) * Tk + j] = acc[r][c] + ps;
        }
      }
    }
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  const int Tk4 = (Tk + 3) & ~3;                 // row pitch (TMA-legal for the tensor-core consumer)
  float *Abase = A + aoff[u] + (long long)h * Tk * Tk4;
  for (int r = warp; r < nrow; r += 8) {
    float *row = sc + r * Tk;
    float mx = -INFINITY;
    for (int j = lane; j < Tk; j += 32) mx = fmaxf(mx, row[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < Tk; j += 32) {
      const float e = expf(row[j] - mx);
      row[j] = e;
      sum += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    float *orow = Abase + (long long)(i0 + r) * Tk4;
    for (int j = lane; j < Tk; j += 32) orow[j] = row[j] / sum;
  }
}

// ------------------------------------------------------------------ attention application: out = (A_h * V) (* Y)
// CTA = 32*RPT query rows x 4*CG columns of one utterance; 32*CG threads, warp = column group, lane = row group.
// Thread tile = RPT rows (lane + 32 r) x 4 columns; keys go through shared memory in chunks of 32
// (A tile [rows][33]: conflict-free scalar reads; V tile [32][4*CG]: one 128-bit broadcast read per key).
template <int CG, int RPT>
__global__ void __launch_bounds__(32 * CG) attn_apply_kernel(const float *__restrict__ A, const long long *__restrict__ aoff,
                                                             const int *__restrict__ len, const int *__restrict__ off,
                                                             const float *__restrict__ X, int ldx, const float *__restrict__ S, int lds,
                                                             const float *__restrict__ Y, int ldy, int C, int single_head,
                                                             float *__restrict__ out, int ldo) {
  constexpr int KC = 32, ROWS = 32 * RPT, CT = 4 * CG, NT = 32 * CG;
  extern __shared__ __align__(16) float smem[];
  float *As = smem;                    // [ROWS][KC + 1]
  float *Vs = As + ROWS * (KC + 1);    // [KC][CT]
  const int u = blockIdx.z;
  const int Tk = len[u];
  const int i0 = blockIdx.x * ROWS;
  if (i0 >= Tk) return;
  const int ctile = blockIdx.y;
  const int c0 = ctile * CT;
  const int head = single_head ? 0 : ctile;
  const int Tk4 = (Tk + 3) & ~3;
  const float *Ah = A + aoff[u] + (long long)head * Tk * Tk4;
  const long long rbase = off[u];
  const int cg = threadIdx.x >> 5, rg = threadIdx.x & 31;
  float acc[RPT][4];
#pragma unroll
  for (int r = 0; r < RPT; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
  for (int j0 = 0; j0 < Tk; j0 += KC) {
    for (int i = threadIdx.x; i < ROWS * KC; i += NT) {
      const int r = i / KC, jj = i % KC;
      const int gi = i0 + r, gj = j0 + jj;
      As[r * (KC + 1) + jj] = (gi < Tk && gj < Tk) ? Ah[(long long)gi * Tk4 + gj] : 0.f;
    }
    for (int i = threadIdx.x; i < KC * CT; i += NT) {
      const int jj = i / CT, c = i % CT;
      const int gj = j0 + jj, gc = c0 + c;
      float v = 0.f;
      if (gj < Tk && gc < C) {
        v = X[(rbase + gj) * ldx + gc];
        if (S) v *= tanhf(S[(rbase + gj) * lds + gc]);
      }
      Vs[jj * CT + c] = v;
    }
    __syncthreads();
#pragma unroll 8
    for (int jj = 0; jj < KC; ++jj) {
      const float4 v = *reinterpret_cast<const float4 *>(Vs + jj * CT + cg * 4);
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const float a = As[(rg + 32 * r) * (KC + 1) + jj];
        acc[r][0] = fmaf(a, v.x, acc[r][0]); acc[r][1] = fmaf(a, v.y, acc[r][1]);
        acc[r][2] = fmaf(a, v.z, acc[r][2]); acc[r][3] = fmaf(a, v.w, acc[r][3]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    const int gi = i0 + rg + 32 * r;
    if (gi >= Tk) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int gc = c0 + cg * 4 + c;
      if (gc >= C) continue;
      float v = acc[r][c];
      if (Y) v *= Y[(rbase + gi) * ldy + gc];
      out[(rbase + gi) * ldo + gc] = v;
    }
  }
}

// ------------------------------------------------------------------ conv module: GLU -> depthwise conv -> SwooshR
// CTA = one 128-frame tile of one utterance (tile list over the ragged batch, no empty CTAs) x 64 channels. The GLU
// values x * sigmoid(gate) of the tile plus its (k-1)-frame halo (zero outside the utterance) are staged once in
// shared memory with 128-bit loads; each thread then owns one channel and 32 consecutive frames and slides a register
// window over them, so a staged value is read from shared memory once per 8 outputs instead of once per tap.
constexpr int kDwT = 128, kDwC = 64, kDwMaxK = 31;
// KMAX = 15 or 31: the register window and tap array are sized for the stack's kernel, so the k = 15 stacks run at
// twice the occupancy of the k = 31 ones
template <int KMAX>
__global__ void __launch_bounds__(256, KMAX <= 15 ? 4 : 3) glu_dwconv_kernel(const float *__restrict__ h, const int *__restrict__ len,
                                                         const int *__restrict__ off, const int *__restrict__ tile_off, int n_utt,
                                                         int D, int k, const float *__restrict__ w, const float *__restrict__ b,
                                                         float *__restrict__ out) {
  __shared__ __align__(16) float g[(kDwT + kDwMaxK - 1)][kDwC];
  int lo = 0, hi = n_utt - 1;
  const int tile = blockIdx.x;
  while (lo < hi) {   // last u with tile_off[u] <= tile
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(tile_off + mid) <= tile) lo = mid; else hi = mid - 1;
  }
  const int u = lo;
  const int L = len[u];
  const int t0 = (tile - __ldg(tile_off + u)) * kDwT;
  const int c0 = blockIdx.y * kDwC;
  const int pad = k / 2;
  const long long rbase = off[u];
  const int rows = min(kDwT, L - t0) + k - 1;
  for (int i = threadIdx.x; i < rows * (kDwC / 4); i += blockDim.x) {
    const int r = i / (kDwC / 4), c4 = (i % (kDwC / 4)) * 4;
    const int t = t0 - pad + r, gc = c0 + c4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t >= 0 && t < L && gc < D) {     // D % 4 == 0
      const float *row = h + (rbase + t) * (2LL * D);
      const float4 x = *reinterpret_cast<const float4 *>(row + gc), s4 = *reinterpret_cast<const float4 *>(row + D + gc);
      v = make_float4(x.x * sigmoid_f(s4.x), x.y * sigmoid_f(s4.y), x.z * sigmoid_f(s4.z), x.w * sigmoid_f(s4.w));
    }
    *reinterpret_cast<float4 *>(&g[r][c4]) = v;
  }
  __syncthreads();
  const int c = threadIdx.x % kDwC, tq = threadIdx.x / kDwC;  // 4 groups x 32 frames
  const int gc = c0 + c;
  if (gc >= D) return;
  float wr[KMAX];
#pragma unroll
  for (int j = 0; j < KMAX; ++j) wr[j] = j < k ? __ldg(w + j * D + gc) : 0.f;
  const float bias = __ldg(b + gc);
  const int tend = min(kDwT, L - t0);
#pragma unroll 1
  for (int f0 = tq * 32; f0 < tq * 32 + 32 && f0 < tend; f0 += 8) {
    // register window: outputs f0..f0+7 read staged rows f0..f0+7+k-1 (rows past the staged range belong to frames
    // that are not written)
    float win[8 + KMAX - 1];
#pragma unroll
    for (int j = 0; j < 8 + KMAX - 1; ++j) win[j] = (j < 8 + k - 1 && f0 + j < rows) ? g[f0 + j][c] : 0.f;
#pragma unroll
    for (int f = 0; f < 8; ++f) {
      float acc = bias;
#pragma unroll
      for (int j = 0; j < KMAX; ++j) acc = fmaf(wr[j], win[f + j], acc);   // wr[j] = 0 for j >= k
      if (f0 + f < tend) out[(rbase + t0 + f0 + f) * D + gc] = swoosh_r(acc);
    }
  }
}

inline unsigned cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

}  // namespace

// ====================================================================== launchers
void launch_embed_conv0(const float *feats, const long long *foff, const long long *ooff, int n, long long total_rows, const float *w,
                        const float *b, float *out, cudaStream_t st) {
  if (total_rows <= 0) return;
  embed_conv0_kernel<<<cdiv(total_rows * 80, 256), 256, 0, st>>>(feats, foff, ooff, n, w, b, out);
  count_launch(); KERNEL_CHECK();
}
void launch_embed_conv1(const float *in, const long long *ioff, const long long *ooff, int n, long long total_rows, const float *w,
                        const float *b, float *out, cudaStream_t st) {
  if (total_rows <= 0) return;
  static int n_sms = 0;
  if (n_sms == 0) {
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    CUDA_CHECK(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const long long items = total_rows * 39;
  const unsigned grid = (unsigned)std::min<long long>((items + 255) / 256, (long long)n_sms * 8);
  embed_conv1_kernel<<<grid, 256, 0, st>>>(in, ioff, ooff, n, w, b, out);
  count_launch(); KERNEL_CHECK();
}
void launch_embed_im2col2(const float *in, const long long *ioff, const int *ooff, int n, long long total_rows, float *out,
                          cudaStream_t st) {
  if (total_rows <= 0) return;
  embed_im2col2_kernel<<<cdiv(total_rows * 19 * 8, 256), 256, 0, st>>>(in, ioff, ooff, n, out);
  count_launch(); KERNEL_CHECK();
}
void launch_embed_dw7(const float *in, const RaggedDesc &r, const int *tile_off, int n_tiles, const float *w, const float *b,
                      float *out, cudaStream_t st) {
  if (r.total <= 0 || n_tiles <= 0) return;
  dim3 grid(n_tiles * 16, 128 / kDw7Ch);
  embed_dw7_kernel<<<grid, 256, 0, st>>>(in, r.len, r.off, tile_off, r.n, w, b, out);
  count_launch(); KERNEL_CHECK();
}
void launch_biasnorm(const float *x, int M, int D, const float *bias, const float *log_scale, float *out, cudaStream_t st) {
  if (M <= 0) return;
  biasnorm_kernel<<<cdiv(M, 8), 256, 0, st>>>(x, nullptr, M, D, bias, log_scale, nullptr, out);
  count_launch(); KERNEL_CHECK();
}
void launch_biasnorm_bypass(const float *x, const float *orig, int M, int D, const float *bias, const float *log_scale,
                            const float *bypass, float *out, cudaStream_t st) {
  if (M <= 0) return;
  biasnorm_kernel<<<cdiv(M, 8), 256, 0, st>>>(x, orig, M, D, bias, log_scale, bypass, out);
  count_launch(); KERNEL_CHECK();
}
void launch_bypass(const float *x, const float *orig, long long M, int D, const float *scale, float *out, cudaStream_t st) {
  if (M <= 0) return;
  if (D & 3) throw CudaError("bypass: channel count must be a multiple of 4");
  const long long total4 = M * (D / 4);
  bypass_kernel<<<cdiv(total4, 256), 256, 0, st>>>(reinterpret_cast<const float4 *>(x), reinterpret_cast<const float4 *>(orig), total4,
                                                   D / 4, reinterpret_cast<const float4 *>(scale), reinterpret_cast<float4 *>(out));
  count_launch(); KERNEL_CHECK();
}
void launch_convert_channels(const float *in, int Cin, float *out, int Cout, long long M, cudaStream_t st) {
  if (M <= 0) return;
  if ((Cin | Cout) & 3) throw CudaError("convert_channels: channel counts must be multiples of 4");
  convert_channels_kernel<<<cdiv(M * (Cout / 4), 256), 256, 0, st>>>(reinterpret_cast<const float4 *>(in), Cin / 4,
                                                                    reinterpret_cast<float4 *>(out), Cout / 4, M);
  count_launch(); KERNEL_CHECK();
}
void launch_build_row_maps(const RaggedDesc &r1, const RaggedDesc &rq, int ds, int *up, int2 *down, cudaStream_t st) {
  if (r1.n <= 0) return;
  build_row_maps_kernel<<<r1.n, 256, 0, st>>>(r1.len, r1.off, rq.len, rq.off, ds, up, down);
  count_launch(); KERNEL_CHECK();
}
void launch_downsample(const float *in, const int2 *down, int rows_out, int C, int ds, const float *bias, float *out, cudaStream_t st) {
  if (rows_out <= 0) return;
  if (C & 3) throw CudaError("downsample: channel count must be a multiple of 4");
  const long long total4 = (long long)rows_out * (C / 4);
  const float4 *i4 = reinterpret_cast<const float4 *>(in);
  float4 *o4 = reinterpret_cast<float4 *>(out);
  if (ds == 2) downsample_kernel<2><<<cdiv(total4, 256), 256, 0, st>>>(i4, down, total4, C / 4, bias, o4);
  else if (ds == 4) downsample_kernel<4><<<cdiv(total4, 256), 256, 0, st>>>(i4, down, total4, C / 4, bias, o4);
  else if (ds == 8) downsample_kernel<8><<<cdiv(total4, 256), 256, 0, st>>>(i4, down, total4, C / 4, bias, o4);
  else throw CudaError("downsample: factor must be 2, 4 or 8");
  count_launch(); KERNEL_CHECK();
}
void launch_upsample_combine(const float *y, const int *up, const float *orig, int rows_full, int C, const float *scale, float *out,
                             cudaStream_t st) {
  if (rows_full <= 0) return;
  if (C & 3) throw CudaError("upsample: channel count must be a multiple of 4");
  const long long total4 = (long long)rows_full * (C / 4);
  upsample_combine_kernel<<<cdiv(total4, 256), 256, 0, st>>>(reinterpret_cast<const float4 *>(y), up, reinterpret_cast<const float4 *>(orig),
                                                             total4, C / 4, reinterpret_cast<const float4 *>(scale),
                                                             reinterpret_cast<float4 *>(out));
  count_launch(); KERNEL_CHECK();
}
void launch_concat_downsample2(const ConcatPiece *pieces, int n_pieces, const int2 *down, int rows_out, int C, const float *bias,
                               float *out, cudaStream_t st) {
  if (rows_out <= 0) return;
  ConcatArgs a{};
  a.n = n_pieces;
  for (int i = 0; i < n_pieces && i < 4; ++i) {
    a.src[i] = pieces[i].src; a.ld[i] = pieces[i].ld; a.c0[i] = pieces[i].c0; a.c1[i] = pieces[i].c1;
    if ((pieces[i].ld | pieces[i].c0 | pieces[i].c1) & 3) throw CudaError("concat: piece boundaries must be multiples of 4");
  }
  if (C & 3) throw CudaError("concat: channel count must be a multiple of 4");
  const long long total4 = (long long)rows_out * (C / 4);
  concat_downsample2_kernel<<<cdiv(total4, 256), 256, 0, st>>>(a, down, total4, C / 4, bias, reinterpret_cast<float4 *>(out));
  count_launch(); KERNEL_CHECK();
}
void launch_pos_emb(float *pe, int max_len, int pos_dim, cudaStream_t st) {
  if (max_len <= 0) return;
  pos_emb_kernel<<<cdiv((long long)(2 * max_len - 1) * (pos_dim / 2), 256), 256, 0, st>>>(pe, max_len, pos_dim);
  count_launch(); KERNEL_CHECK();
}

void launch_attn_weights(const float *proj, int ldp, const float *pos, const RaggedDesc &r, const long long *aoff, int H, int qd,
                         int pd, float *A, cudaStream_t st) {
  if (r.total <= 0) return;
  if (qd != 32 || pd != 4) throw CudaError("attn_weights: only query_head_dim=32, pos_head_dim=4 are built");
  const size_t smem = (size_t)(kAttRows * (qd + pd) + qd * kAttKStride + (size_t)kAttRows * r.max_len) * sizeof(float);
  if (smem > 227 * 1024) throw CudaError("attn_weights: segment too long for one pass (max ~2900 frames at the stack rate)");
  set_max_dynamic_smem(attn_weights_kernel<32, 4>, 227 * 1024);
  dim3 grid(cdiv(r.max_len, kAttRows), H, r.n);
  attn_weights_kernel<32, 4><<<grid, 256, smem, st>>>(proj, ldp, pos, H * pd, r.len, r.off, aoff, H, r.max_len, A);
  count_launch(); KERNEL_CHECK();
}

void launch_attn_apply(const float *A, const long long *aoff, const RaggedDesc &r, const float *X, int ldx, const float *S, int lds,
                       const float *Y, int ldy, int C, int dv_per_head, int single_head, float *out, int ldo, cudaStream_t st) {
  if (r.total <= 0) return;
  constexpr int RPT = 8, ROWS = 32 * RPT;
  set_max_dynamic_smem(attn_apply_kernel<8, RPT>, 64 * 1024);
  set_max_dynamic_smem(attn_apply_kernel<3, RPT>, 64 * 1024);
  if (single_head) {
    const size_t smem = (size_t)(ROWS * 33 + 32 * 32) * sizeof(float);
    dim3 grid(cdiv(r.max_len, ROWS), cdiv(C, 32), r.n);
    attn_apply_kernel<8, RPT><<<grid, 256, smem, st>>>(A, aoff, r.len, r.off, X, ldx, S, lds, Y, ldy, C, 1, out, ldo);
  } else {
    if (dv_per_head != 12) throw CudaError("attn_apply: only value_head_dim=12 is built");
    const size_t smem = (size_t)(ROWS * 33 + 32 * 12) * sizeof(float);
    dim3 grid(cdiv(r.max_len, ROWS), C / 12, r.n);
    attn_apply_kernel<3, RPT><<<grid, 96, smem, st>>>(A, aoff, r.len, r.off, X, ldx, S, lds, Y, ldy, C, 0, out, ldo);
  }
  count_launch(); KERNEL_CHECK();
}

void launch_glu_dwconv(const float *h, const RaggedDesc &r, const int *tile_off, int n_tiles, int D, int k, const float *w,
                       const float *b, float *out, cudaStream_t st) {
  if (r.total <= 0 || n_tiles <= 0) return;
  if (k > kDwMaxK) throw CudaError("glu_dwconv: kernel size > 31 not built");
  if (D & 3) throw CudaError("glu_dwconv: channel count must be a multiple of 4");
  dim3 grid(n_tiles, cdiv(D, kDwC));
  if (k <= 15) glu_dwconv_kernel<15><<<grid, 256, 0, st>>>(h, r.len, r.off, tile_off, r.n, D, k, w, b, out);
  else glu_dwconv_kernel<31><<<grid, 256, 0, st>>>(h, r.len, r.off, tile_off, r.n, D, k, w, b, out);
  count_launch(); KERNEL_CHECK();
}

}  // namespace b200asr
