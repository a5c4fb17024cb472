// TF32 tensor-core GEMM / implicit-GEMM convolution on tcgen05 (sm_100a): TMA-fed, TMEM accumulators.
#pragma once
#include "common.cuh"

namespace zvx {

// True when `a` can run on the tcgen05 path (alignment / layout constraints of the TMA descriptors).
bool gemm_tc_supported(const GemmArgs& a);
// Same contract as gemm_simt, operands rounded to TF32 by the tensor core, fp32 accumulation.
void gemm_tc(const GemmArgs& a, cudaStream_t st);

}  // namespace zvx
