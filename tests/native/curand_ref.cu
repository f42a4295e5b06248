// curand_ref.cu -- TEST-ONLY: the reference's RNG calls (curand_init(seed, i, 0, &state[i]) src/variable.cu:10 and
// curand_uniform4(&state[i]) src/variable.cu:51 / src/module.cu:25) executed with the real cuRAND device API, so
// that the engine's stateless Philox can be checked bit for bit on the GPU.  Also buildable for the host
// (-DCURAND_REF_HOST with g++) where the same headers run on the CPU.
#ifdef CURAND_REF_HOST
#define QUALIFIERS static inline __attribute__((always_inline))
#define __host__
#define __device__
#define __forceinline__ inline
#endif
#include <cuda_runtime.h>
#include <curand_kernel.h>

#ifdef CURAND_REF_HOST
extern "C" __attribute__((visibility("default"))) int curand_ref_host(unsigned seed, int n_states, int n_draws, float *out) {
  for (int i = 0; i < n_states; i++) {
    curandStatePhilox4_32_10_t s;
    curand_init(seed, i, 0, &s);
    for (int t = 0; t < n_draws; t++) {
      const float4 u = curand_uniform4(&s);
      float *o = out + ((size_t)t * n_states + i) * 4;
      o[0] = u.x; o[1] = u.y; o[2] = u.z; o[3] = u.w;
    }
  }
  return 0;
}
#else
__global__ void k(unsigned seed, int n_states, int n_draws, float *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_states) return;
  curandStatePhilox4_32_10_t s;
  curand_init(seed, i, 0, &s);
  for (int t = 0; t < n_draws; t++) {
    const float4 u = curand_uniform4(&s);
    float *o = out + ((size_t)t * n_states + i) * 4;
    o[0] = u.x; o[1] = u.y; o[2] = u.z; o[3] = u.w;
  }
}
// out: host buffer [n_draws][n_states][4]
extern "C" __attribute__((visibility("default"))) int curand_ref_device(unsigned seed, int n_states, int n_draws, float *out) {
  float *d = nullptr;
  const size_t bytes = (size_t)n_states * n_draws * 4 * sizeof(float);
  if (cudaMalloc(&d, bytes) != cudaSuccess) return 1;
  k<<<(n_states + 127) / 128, 128>>>(seed, n_states, n_draws, d);
  const int rc = (int)cudaMemcpy(out, d, bytes, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return rc;
}
#endif
