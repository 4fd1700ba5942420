// Epilogue of the tcgen05 GEMM kernels (gemm_tc.cu and its operand-format variants): warps 2..9 drain a finished TMEM
// accumulator - bias / Swoosh / residual with coalesced 128-bit stores, or (EPI > 0) the joiner's softmax / top-k partial
// records - and hand the accumulator back to the MMA issuer.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace b200asr {
namespace tc {

__device__ __forceinline__ float softplus_f(float x) { return softplus_sfu(x); }
__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_SWOOSH_L) return softplus_f(v - 4.0f) - 0.08f * v - 0.035f;
  if (act == ACT_SWOOSH_R) return softplus_f(v - 1.0f) - 0.08f * v - 0.313261687f;
  return v;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct TcParams {
  const float *bias;
  const float *R; int ldr;
  float *C; int ldc;
  int M, N, K, act;
  float *partials;   // joiner epilogue (EPI > 0)
  unsigned long long *trace;
  int *tile_counter; // dynamic tile scheduling: a global counter that is zero at launch (null = static round-robin)
  int dbg;           // timing experiments only (B200ASR_DBG_GEMM): bit 0 = the W_lo tile is not loaded, bit 1 = half of the A tile is not
                     // loaded (results are wrong; shows how the kernel's time depends on operand bytes per stage)
  float acc_scale;   // the accumulator is multiplied by this (a power of two) before the epilogue; 0 or 1 = none. The all-fp16
                     // operand split pre-scales both operands so their low parts stay in fp16's normal range.
};

// step-trace slot (joiner GEMM of the search only): %globaltimer of CTA 0
template <int EPI>
__device__ __forceinline__ void tc_trace_mark(const TcParams &p, int slot) {
  if (EPI > 0 && p.trace && blockIdx.x == 0) {
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    p.trace[slot] = now;
  }
}

// EPI > 0: joiner record epilogue with top-EPI; EPI <= 0 carries the activation as a compile-time constant
// (0 none, -1 SwooshL, -2 SwooshR). Called by warps 2..9 of the CTA; accumulator `ti & 1` of tile number ti sits at TMEM
// columns [acc * BN, acc * BN + BN).
template <int BN, int EPI>
__device__ __forceinline__ void tc_epilogue_warps(const TcParams &p, uint32_t tmem_base, uint64_t *tmem_full_bar, uint64_t *tmem_empty_bar,
                                                  float *epi_stage, const TileSched &sched, int tiles_n, int warp, int lane) {
  constexpr int kAct = EPI == -1 ? (int)ACT_SWOOSH_L : (EPI == -2 ? (int)ACT_SWOOSH_R : (int)ACT_NONE);
  // ===== epilogue warps 2..9: TMEM lane quarter = warp % 4, column half = (warp - 2) / 4
  const int q = warp & 3;
  const int chalf = (warp - 2) >> 2;
  for (int ti = 0;; ++ti) {
    const int tile = sched_consume_warp(sched, ti, lane);
    if (tile < 0) break;
    const int m0 = (tile / tiles_n) * TBM, n0 = (tile % tiles_n) * BN;
    const int acc = ti & 1;
    // joiner records: the bias of this warp's first 32 columns is fetched while the main loop runs (the step is latency-bound
    // and the bias would otherwise cost an L2 round trip after the accumulator is complete)
    float4 bias_pf[EPI > 0 ? 8 : 1];
    if constexpr (EPI > 0) {
      const int nb = n0 + chalf * (BN / 2);
      if (nb + 32 <= p.N) {
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) bias_pf[j4] = __ldg(reinterpret_cast<const float4 *>(p.bias + nb) + j4);
      }
    }
    mbar_wait(&tmem_full_bar[acc], (ti >> 1) & 1);
    if (ti == 0 && warp == 2 && lane == 0) tc_trace_mark<EPI>(p, 4);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t stg = smem_u32(epi_stage + (warp - 2) * (32 * 32));   // per-warp 32x32 transpose tile, 16-byte chunks XOR-swizzled by row
    const bool vec_ok = ((p.ldc & 3) == 0) && ((p.N & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                        (!p.R || (((p.ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.R) & 15) == 0))) &&
                        (!p.bias || EPI > 0 || ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0));
#pragma unroll 1
    for (int c0 = chalf * (BN / 2); c0 < (chalf + 1) * (BN / 2); c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), r);
      if (n0 + c0 >= p.N) continue;                      // warp-uniform
      // the accumulator scale (a power of two, so the product is exact) rides in the FMA that adds the bias: same bits as
      // scaling first, one instruction less per element
      const float asc = (p.acc_scale != 0.f) ? p.acc_scale : 1.0f;
      const int mrow0 = m0 + q * 32;
      if constexpr (EPI > 0) {
        const int nbase = n0 + c0;
        float mx = -INFINITY;
        if (nbase + 32 <= p.N) {                          // warp-uniform; only the last tile of a row is ragged
          const float4 *b4 = reinterpret_cast<const float4 *>(p.bias + nbase);   // bias is 16-byte aligned (checked on the host)
          const bool first_chunk = c0 == chalf * (BN / 2);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 bq = first_chunk ? bias_pf[j4] : __ldg(b4 + j4);
            const float x0 = fmaf(__uint_as_float(r[4 * j4]), asc, bq.x), x1 = fmaf(__uint_as_float(r[4 * j4 + 1]), asc, bq.y);
            const float x2 = fmaf(__uint_as_float(r[4 * j4 + 2]), asc, bq.z), x3 = fmaf(__uint_as_float(r[4 * j4 + 3]), asc, bq.w);
            r[4 * j4] = __float_as_uint(x0); r[4 * j4 + 1] = __float_as_uint(x1);
            r[4 * j4 + 2] = __float_as_uint(x2); r[4 * j4 + 3] = __float_as_uint(x3);
            mx = fmaxf(mx, fmaxf(fmaxf(x0, x1), fmaxf(x2, x3)));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = nbase + j;
            const float x = n < p.N ? fmaf(__uint_as_float(r[j]), asc, __ldg(p.bias + n)) : -INFINITY;
            r[j] = __float_as_uint(x);
            mx = fmaxf(mx, x);
          }
        }
        // sum = S e^(x-mx) feeds the log-softmax, su = S e^(x-mx) (x-mx) and st = S e^((x-mx)/3) the per-token entropy /
        // Tsallis statistics. ex2.approx on (x-mx) log2(e): relative error <= 2^-22 on the terms near the maximum
        // that carry the sum, far inside the fp32 noise of the logits themselves. Two interleaved accumulator sets
        // halve the dependent chains (the epilogue is latency-bound, two warps per scheduler).
        float sum[2] = {0.f, 0.f}, su[2] = {0.f, 0.f}, st[2] = {0.f, 0.f}, tv[EPI];
        int ti[EPI];
#pragma unroll
        for (int i = 0; i < EPI; ++i) { tv[i] = -INFINITY; ti[i] = -1; }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float x = __uint_as_float(r[j]);
          const float dx = x - mx;                        // -inf past N: e = 0, and the product below is skipped
          const float tl = dx * 1.4426950408889634f;
          const float e = ex2_approx(tl);
          sum[j & 1] += e;
          su[j & 1] = fmaf(e, fmaxf(dx, -3.0e38f), su[j & 1]);   // dx = -inf past N: 0 * finite
          st[j & 1] += ex2_approx(tl * (1.0f / 3.0f));
          if (x > tv[EPI - 1]) {                          // strict: equal values keep the lower column first
            float cv = x;
            int ci = nbase + j;
#pragma unroll
            for (int i = 0; i < EPI; ++i) {
              if (cv > tv[i]) { const float fv = tv[i]; const int fi = ti[i]; tv[i] = cv; ti[i] = ci; cv = fv; ci = fi; }
            }
          }
        }
        const int mrow = mrow0 + lane;
        if (mrow < p.M) {
          const int n_parts = (p.N + 31) >> 5;
          float4 *rec = reinterpret_cast<float4 *>(p.partials + ((long long)mrow * n_parts + (nbase >> 5)) * (4 + 2 * EPI));
          rec[0] = make_float4(mx, sum[0] + sum[1], su[0] + su[1], st[0] + st[1]);
#pragma unroll
          for (int i = 0; i < EPI / 4; ++i) {
            rec[1 + i] = make_float4(tv[4 * i], tv[4 * i + 1], tv[4 * i + 2], tv[4 * i + 3]);
            rec[1 + EPI / 4 + i] = make_float4(__int_as_float(ti[4 * i]), __int_as_float(ti[4 * i + 1]), __int_as_float(ti[4 * i + 2]),
                                               __int_as_float(ti[4 * i + 3]));
          }
        }
      }
      if (EPI > 0 && p.C == nullptr) continue;               // records only: the selection never reads the logits
      const float *bias_late = EPI > 0 ? nullptr : p.bias;   // the joiner epilogue has already added it (and scaled)
      const float sc_late = EPI > 0 ? 1.0f : asc;
      __syncwarp();
#pragma unroll
      for (int k4 = 0; k4 < 8; ++k4)                      // row = lane; 128-bit stores, conflict-free per quarter warp
        sts128(stg + (uint32_t)(lane * 32 + ((k4 ^ (lane & 7)) << 2)) * 4u,
               make_float4(__uint_as_float(r[4 * k4]), __uint_as_float(r[4 * k4 + 1]), __uint_as_float(r[4 * k4 + 2]), __uint_as_float(r[4 * k4 + 3])));
      __syncwarp();
      if (vec_ok) {
        // thread = (row lane/8 + 4*it, 4 columns (lane%8)*4): a warp instruction covers 4 full 128-byte rows
        const int c4 = (lane & 7) * 4, rsub = lane >> 3;
        const int n = n0 + c0 + c4;
        const bool nv = n < p.N;
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias_late && nv) bv = __ldg(reinterpret_cast<const float4 *>(bias_late + n));
        float4 res[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {                   // all residual loads in flight before use
          const int m = mrow0 + rsub + 4 * it;
          float4 rv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.R && nv && m < p.M) rv = *reinterpret_cast<const float4 *>(p.R + (long long)m * p.ldr + n);
          res[it] = rv;
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rr = rsub + 4 * it, m = mrow0 + rr;
          if (nv && m < p.M) {
            float4 v = lds128(stg + (uint32_t)(rr * 32 + (((lane & 7) ^ (rr & 7)) << 2)) * 4u);
            v.x = apply_act(fmaf(v.x, sc_late, bv.x), kAct) + res[it].x; v.y = apply_act(fmaf(v.y, sc_late, bv.y), kAct) + res[it].y;
            v.z = apply_act(fmaf(v.z, sc_late, bv.z), kAct) + res[it].z; v.w = apply_act(fmaf(v.w, sc_late, bv.w), kAct) + res[it].w;
            *reinterpret_cast<float4 *>(p.C + (long long)m * p.ldc + n) = v;
          }
        }
      } else {
        // generic path: lane = column, one row per iteration
        const int n = n0 + c0 + lane;
        const bool nv = n < p.N;
        const float bs = (bias_late && nv) ? __ldg(bias_late + n) : 0.f;
#pragma unroll 4
        for (int rr = 0; rr < 32; ++rr) {
          const int m = mrow0 + rr;
          if (nv && m < p.M) {
            const float4 v4 = lds128(stg + (uint32_t)(rr * 32 + (((lane >> 2) ^ (rr & 7)) << 2)) * 4u);
            float v = fmaf((lane & 3) == 0 ? v4.x : (lane & 3) == 1 ? v4.y : (lane & 3) == 2 ? v4.z : v4.w, sc_late, bs);
            const float rv = p.R ? p.R[(long long)m * p.ldr + n] : 0.f;
            v = apply_act(v, kAct) + rv;
            p.C[(long long)m * p.ldc + n] = v;
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    if (ti == 0 && warp == 2 && lane == 0) tc_trace_mark<EPI>(p, 5);
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[acc])) : "memory");
  }
}

}  // namespace tc
}  // namespace b200asr
