// tcgen05 TF32 GEMM (placeholder until the tensor-core kernel lands; the engine falls back to the exact-fp32 path).
#include "kernels.h"
int d4_gemm_tc_supported(const GemmArgs&) { return 0; }
int d4_gemm_tc(const GemmArgs&, int, cudaStream_t) { return d4_fail("tcgen05 GEMM not built"); }
