// philox.cuh -- stateless Philox4x32-10 (Salmon et al., SC'11) laid out the way the reference's cuRAND
// states are consumed: draw t of state g = Philox(ctr=(t,0,g,0), key=(seed,0)); uniform = x*2^-32 + 2^-33
// (reference: curand_init(seed, g, 0) src/variable.cu:10, curand_uniform4 src/variable.cu:51, src/module.cu:25).
// No state array: the reference keeps 64 B of Philox state per 4 elements in HBM.
#pragma once
#include "common.cuh"

namespace gcnb {

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                         uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
#ifdef __CUDA_ARCH__
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// draw index of state g given the history table (see gcnb_rng_t)
__device__ __forceinline__ uint32_t rng_draw_index(const gcnb_rng_t &rng, uint32_t g) {
  uint32_t t = 0;
  for (int i = 0; i < rng.n_hist; i++)  // n_hist is 2-4 in practice; the table lives in constant memory
    if (rng.hist_groups[i] > g) t += rng.hist_count[i];
  return t;
}

// four uniforms in (0,1] for element group g; same expression as cuRAND's _curand_uniform so nvcc contracts it
// to the same single FMA.
__device__ __forceinline__ void rng_uniform4(const gcnb_rng_t &rng, uint32_t g_local, float u[4]) {
  const uint32_t g = g_local + rng.group_offset;  // global element group (independent of the row partition)
  uint32_t x[4];
  philox4x32_10(rng_draw_index(rng, g), 0u, g, 0u, rng.seed, 0u, x);
#pragma unroll
  for (int k = 0; k < 4; k++) u[k] = x[k] * 2.3283064e-10f + (2.3283064e-10f / 2.0f);
}

}  // namespace gcnb
