// FP32 CUDA-core GEMM with fused epilogue:  C = act(A * W^T + bias) (+ R).
// This is the FP32 (bit-exact token) mode of every Linear on the path (encoder in_proj/out_proj,
// feed-forward, pointwise convolutions, encoder_proj, joiner): the reference runs them in fp32 through
// onnxruntime (/root/reference core/asr_engine.py:1047,1055,1092). True fp32 FMA accumulation keeps the
// logits inside the margin the token-exact gate needs; the BF16 tensor-core (tcgen05) twin lives in gemm_tc.cu.
//
// Tiling: 128 x BN x 16 per CTA, 256 threads, 8 x (BN/16) outputs per thread, register-staged double buffering,
// operands transposed into shared memory so the inner product reads are conflict-free 128-bit loads.
#include <mutex>
#include <set>
#include <utility>

#include "common.cuh"

namespace b200asr {

long long g_launches = 0;

namespace { thread_local int tl_sm_reserve = 0; }
void set_sm_reserve(int reserve) { tl_sm_reserve = reserve > 0 ? reserve : 0; }
int persistent_grid_limit(int n_sms) { return n_sms - tl_sm_reserve > 8 ? n_sms - tl_sm_reserve : (n_sms < 8 ? n_sms : 8); }

void set_max_dynamic_smem_impl(const void *func, int bytes) {
  static std::mutex mu;
  static std::set<std::pair<const void *, int>> done;   // (function, device)
  int dev = 0;
  CUDA_CHECK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  if (done.count({func, dev})) return;
  CUDA_CHECK(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done.insert({func, dev});
}

namespace {

__device__ __forceinline__ float softplus_f(float x) {
  // log(1 + exp(x)) computed as torch.logaddexp(0, x) does: max(x,0) + log1p(exp(-|x|))
  return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
}
__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_SWOOSH_L) return softplus_f(v - 4.0f) - 0.08f * v - 0.035f;
  if (act == ACT_SWOOSH_R) return softplus_f(v - 1.0f) - 0.08f * v - 0.313261687f;
  return v;
}

constexpr int BM = 128, BK = 16, NT = 256;

template <int BN>
__global__ void __launch_bounds__(NT) gemm_fp32_kernel(GemmArgs g) {
  constexpr int TN = BN / 16;        // 8 or 4 columns per thread
  constexpr int TNV = TN / 4;        // float4 groups along N
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  // global -> register staging: A tile 128x16 = 512 float4, 2 per thread; W tile BNx16 = BN*4 float4
  constexpr int A_LD = (BM * BK / 4) / NT;  // 2
  constexpr int B_LD = (BN * BK / 4) / NT;  // 2 or 1
  float4 ra[A_LD], rb[B_LD];

  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
      const int idx = tid + i * NT;
      const int r = idx >> 2, kq = (idx & 3) * 4;
      const int m = m0 + r, k = k0 + kq;
      ra[i] = (m < g.M && k < g.K) ? __ldg(reinterpret_cast<const float4 *>(g.A + (long long)m * g.lda + k))
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < B_LD; ++i) {
      const int idx = tid + i * NT;
      const int r = idx >> 2, kq = (idx & 3) * 4;
      const int n = n0 + r, k = k0 + kq;
      rb[i] = (n < g.N && k < g.K) ? __ldg(reinterpret_cast<const float4 *>(g.W + (long long)n * g.K + k))
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
      const int idx = tid + i * NT;
      const int r = idx >> 2, kq = (idx & 3) * 4;
      As[buf][kq + 0][r] = ra[i].x; As[buf][kq + 1][r] = ra[i].y;
      As[buf][kq + 2][r] = ra[i].z; As[buf][kq + 3][r] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < B_LD; ++i) {
      const int idx = tid + i * NT;
      const int r = idx >> 2, kq = (idx & 3) * 4;
      Bs[buf][kq + 0][r] = rb[i].x; Bs[buf][kq + 1][r] = rb[i].y;
      Bs[buf][kq + 2][r] = rb[i].z; Bs[buf][kq + 3][r] = rb[i].w;
    }
  };

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nk = (g.K + BK - 1) / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kb = 0; kb < nk; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nk) load_tiles((kb + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[8], b[TN];
      const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][k][64 + ty * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
      for (int v = 0; v < TNV; ++v) {
        const float4 bv = *reinterpret_cast<const float4 *>(&Bs[buf][k][v * (BN / TNV) + tx * 4]);
        b[v * 4 + 0] = bv.x; b[v * 4 + 1] = bv.y; b[v * 4 + 2] = bv.z; b[v * 4 + 3] = bv.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kb + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
#pragma unroll
    for (int v = 0; v < TNV; ++v) {
      const int nb = n0 + v * (BN / TNV) + tx * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = nb + j;
        if (n >= g.N) continue;
        float val = acc[i][v * 4 + j];
        if (g.bias) val += __ldg(g.bias + n);
        val = apply_act(val, g.act);
        if (g.R) val += g.R[(long long)m * g.ldr + n];
        g.C[(long long)m * g.ldc + n] = val;
      }
    }
  }
}

}  // namespace

void launch_gemm_fp32(const GemmArgs &g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return;
  if ((g.K & 3) || (g.lda & 3) || (reinterpret_cast<uintptr_t>(g.A) & 15) || (reinterpret_cast<uintptr_t>(g.W) & 15))
    throw CudaError("gemm_fp32: K, lda must be multiples of 4 and A/W 16-byte aligned");
  if (g.N > 64) {
    dim3 grid((g.M + BM - 1) / BM, (g.N + 127) / 128);
    gemm_fp32_kernel<128><<<grid, NT, 0, st>>>(g);
  } else {
    dim3 grid((g.M + BM - 1) / BM, (g.N + 63) / 64);
    gemm_fp32_kernel<64><<<grid, NT, 0, st>>>(g);
  }
  count_launch();
  KERNEL_CHECK();
}

}  // namespace b200asr
