// reduction.cuh -- placeholder kept so that `#include "../include/reduction.cuh"` in reference-style sources
// resolves.  The reference's __device__ warp_reduce / dead `reduce` kernel (src/reduction.cu:3-23) have no host API;
// the engine's reductions are fixed-order trees inside libgcn_b200 (csrc/loss.cu, csrc/elementwise.cu).
#ifndef REDUCTION_CUH
#define REDUCTION_CUH
#include "../include/utils.cuh"
#endif
