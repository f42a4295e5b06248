// common.cuh -- shared helpers for the sm_100a kernels of libgcn_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/gcnb.h"

#define GCNB_CHECK(call)                       \
  do {                                         \
    cudaError_t e__ = (call);                  \
    if (e__ != cudaSuccess) return (int)e__;   \
  } while (0)

#define GCNB_LAUNCH_CHECK()                    \
  do {                                         \
    cudaError_t e__ = cudaPeekAtLastError();   \
    if (e__ != cudaSuccess) return (int)cudaGetLastError(); \
  } while (0)

static inline cudaStream_t as_stream(gcnb_stream_t s) { return (cudaStream_t)s; }

namespace gcnb {

constexpr int kWarp = 32;

// device properties cached per process (SM count decides persistent grid sizes)
struct DeviceInfo {
  int sm_count = 0;
  int cc_major = 0;
  int ok = 0;
};
const DeviceInfo &device_info();

// host threads a plan builder may use (window staging, bit tiles, ELL): min(16, hardware threads) unless
// gcnb_set_host_threads / GCNB_HOST_THREADS say otherwise -- N ranks on one host share its cores
int host_threads();

// Kernels whose arguments change from epoch to epoch (the Philox descriptor of the dropout kernels, Adam's step size) are
// registered with the index of that argument, so that a captured epoch (CUDA graph) can be replayed with the node's
// arguments patched in place (gcnb_graph_patch_node, spmm.cu) instead of being captured again.
void register_patchable(const void *kernel, int rng_arg, int step_arg);
bool lookup_patchable(const void *kernel, int *rng_arg, int *step_arg);

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// streaming (read-once) loads: keep them out of L1 so gathered rows stay resident
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_stream_f32(const float *p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

}  // namespace gcnb
