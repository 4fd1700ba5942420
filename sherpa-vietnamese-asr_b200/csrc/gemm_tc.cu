// BF16 tensor-core GEMM (tcgen05 + TMEM + TMA) — placeholder until the kernel lands.
#include "common.cuh"
namespace b200asr {
bool gemm_tc_available() { return false; }
void launch_gemm_tc(const GemmArgs &, cudaStream_t) { throw CudaError("tcgen05 GEMM not built"); }
}  // namespace b200asr
