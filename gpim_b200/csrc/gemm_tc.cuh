// gemm_tc.cuh -- tcgen05 split-fp16 GEMM (placeholder until the tensor-core path lands).
#pragma once
#include "common.cuh"
