// Shared tcgen05 / TMA / mbarrier helpers for the sm_100a tensor-core kernels (gemm_tc.cu, attn_tc.cu).
#pragma once
#include <cuda.h>
#include <stdint.h>

#include "common.cuh"

namespace b200asr {
namespace tc {

constexpr int TBM = 128;       // tile M (UMMA M)
constexpr int TBK = 32;        // fp32 elements per stage = 128 bytes = one swizzle row
constexpr int UMMA_K = 8;      // tf32: 32 bytes per instruction
constexpr int kStages1 = 5;    // TF32 mode: 5 x (16 + BN/8) KB
constexpr int kStages3 = 3;    // 3xTF32 mode: 3 x 2 x (16 + BN/8) KB

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Explicit shared-space 128-bit accesses. Pointers into the 1024-byte-aligned dynamic shared buffer are computed through
// uintptr_t, which makes them generic for the compiler: it then emits LD.E / ST.E that go through the global-memory
// path of L1TEX (long scoreboard) instead of LDS / STS.
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, const float4 &v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with the A operand read from tensor memory (128 lanes = rows, one 32-bit column per K element)
__device__ __forceinline__ void umma_tf32_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// registers -> tensor memory: lane = row of the warp's 32-lane quarter, 32 consecutive columns
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, "
      "%27, %28, %29, %30, %31, %32};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
      "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, "
      "%26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Tile scheduler of the persistent kernels. One producer thread hands tile numbers to the other warp roles of its CTA through a
// 4-slot ring in shared memory (full / empty mbarriers per slot); -1 ends the kernel. With a global counter the tiles are
// claimed dynamically (atomicAdd), so a CTA that starts late - its SM was busy with a kernel of another stream, e.g. a frame
// step of the search running beside the encoder - simply takes fewer tiles instead of stretching the whole kernel by its
// static share; without a counter (null) tile = blockIdx.x + i * gridDim.x as before.
constexpr int kSchedSlots = 4;
struct TileSched {
  int *tile;              // shared [kSchedSlots]
  uint64_t *full, *empty; // shared [kSchedSlots] each
  int *counter;           // global, zero at launch (or null)
  int n_tiles;
};
__device__ __forceinline__ void sched_init(const TileSched &s, int n_consumer_arrivals) {   // one thread, before the CTA-wide sync
  for (int i = 0; i < kSchedSlots; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], (uint32_t)n_consumer_arrivals); }
}
// producer thread: claims the ti-th tile of this CTA and publishes it
__device__ __forceinline__ int sched_produce(const TileSched &s, int ti) {
  const int slot = ti & (kSchedSlots - 1);
  mbar_wait(&s.empty[slot], (uint32_t)(((ti / kSchedSlots) & 1) ^ 1));
  int t = s.counter ? atomicAdd(s.counter, 1) : (int)(blockIdx.x + (unsigned)ti * gridDim.x);
  if (t >= s.n_tiles) t = -1;
  *reinterpret_cast<volatile int *>(&s.tile[slot]) = t;
  mbar_arrive(&s.full[slot]);          // release: the slot's content is visible to whoever acquires the phase
  return t;
}
// a consumer role that runs as ONE thread
__device__ __forceinline__ int sched_consume_thread(const TileSched &s, int ti) {
  const int slot = ti & (kSchedSlots - 1);
  mbar_wait(&s.full[slot], (uint32_t)((ti / kSchedSlots) & 1));
  const int t = *reinterpret_cast<volatile int *>(&s.tile[slot]);
  mbar_arrive(&s.empty[slot]);
  return t;
}
// a consumer role that runs as a full warp (one arrival per warp)
__device__ __forceinline__ int sched_consume_warp(const TileSched &s, int ti, int lane) {
  const int slot = ti & (kSchedSlots - 1);
  mbar_wait(&s.full[slot], (uint32_t)((ti / kSchedSlots) & 1));
  const int t = *reinterpret_cast<volatile int *>(&s.tile[slot]);
  __syncwarp();
  if (lane == 0) mbar_arrive(&s.empty[slot]);
  return t;
}

// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start address >> 4, [16,30) leading byte offset >> 4 (unused for swizzled K-major), [32,46) stride byte
//   offset >> 4 (1024 B between 8-row groups), [46,48) version = 1, [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// cute::UMMA::InstrDescriptor: c_format F32 (1) at [4,6), a/b format TF32 (2) at [7,10)/[10,13), K-major both,
// N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// softplus with the SFU exp/log: absolute error ~1e-7 (the 1 + t rounding), i.e. fp32 noise on the O(1) Swoosh
// output, at a fifth of the instructions of expf/log1pf - the epilogue is issue-bound, not memory-bound.

// 2-D fp32 row-major [rows, K] with row stride ld (elements); box = 32 x box_rows, 128-byte swizzle, OOB -> 0
void make_map(CUtensorMap *map, const float *ptr, int rows, int K, int ld, int box_rows);            // cached
void make_map_uncached(CUtensorMap *map, const float *ptr, int rows, int K, int ld, int box_rows);
// same tensor without swizzle and with a free box (box_cols * 4 bytes must be a multiple of 16)
void make_map_plain(CUtensorMap *map, const float *ptr, int rows, int cols, int ld, int box_cols, int box_rows);
bool tc_init();   // resolves cuTensorMapEncodeTiled; false if unavailable

}  // namespace tc
}  // namespace b200asr
