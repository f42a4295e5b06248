// tma_gather.cu -- microbenchmark (not product code): how fast can an SM gather random 64-byte rows of an
// L2-resident matrix WITHOUT the LSU data pipe, i.e. through the TMA unit (cp.async.bulk per lane, or
// cp.async.bulk.tensor tile::gather4), landing them in shared memory and reading them back with conflict-free LDS.128?
// Decides whether GraphSum's remainder entries (random neighbours, 1.3 LSU wavefronts per entry through L1) can move
// off the pipe the window-staged kernel saturates.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/tma_gather tma_gather.cu
//   ./build/tma_gather [entries] [rows]
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e = (x);                                                             \
    if (e != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                       \
    }                                                                                \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  // bounded: a probe must never hang the box
  for (uint32_t spin = 0; spin < (1u << 24); spin++) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void gather4_g2s(void *dst, const void *tmap, int col, int r0, int r1, int r2, int r3,
                                            uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, "
      "%6}], [%7];" ::"r"(smem_u32(dst)),
      "l"(tmap), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar))
      : "memory");
}

// ---- variant 0: the LSU path (reference): 4 lanes per row, 8 rows per LDG.128, indices broadcast by shuffle -----------
__global__ void __launch_bounds__(256) ldg_kernel(const uint32_t *__restrict__ idx, int64_t n, const float *__restrict__ B,
                                                  float *__restrict__ out) {
  const int lane = threadIdx.x & 31, g = lane >> 2, l = lane & 3;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int64_t base = warp * 32; base < n; base += nwarps * 32) {
    const uint32_t my = base + lane < n ? __ldg(idx + base + lane) : 0;
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const uint32_t c = __shfl_sync(0xffffffffu, my, u * 8 + g);
      const float4 x = __ldg(reinterpret_cast<const float4 *>(B + (size_t)c * 16 + l * 4));
      acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
    }
  }
  out[blockIdx.x * (int64_t)blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

// ---- variant 1: every lane issues its own 64-byte bulk copy; S stages per warp ------------------------------------------
// smem per warp: S * (32 slots * 64 B) + S barriers
template <int S>
__global__ void __launch_bounds__(512) bulk_kernel(const uint32_t *__restrict__ idx, int64_t n, const float *__restrict__ B,
                                                   float *__restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  unsigned char *slots = smem + (size_t)w * S * 2048;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)nw * S * 2048) + w * S;
  if (lane == 0)
    for (int s = 0; s < S; s++) mbar_init(bars + s, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const int64_t warp = blockIdx.x * (int64_t)nw + w, nwarps = (int64_t)gridDim.x * nw;
  const int64_t steps = (n / 32 - warp + nwarps - 1) / nwarps;  // full chunks only
  float4 acc = make_float4(0, 0, 0, 0);
  const uint32_t sw = (uint32_t)((lane >> 1) & 3);
  auto issue = [&](int64_t k) {
    const int s = (int)(k % S);
    const uint32_t c = __ldg(idx + (warp + k * nwarps) * 32 + lane);
    if (lane == 0) mbar_expect_tx(bars + s, 2048);
    __syncwarp();
    bulk_g2s(slots + s * 2048 + lane * 64, B + (size_t)c * 16, 64, bars + s);
  };
  for (int64_t k = 0; k < S - 1 && k < steps; k++) issue(k);
  for (int64_t k = 0; k < steps; k++) {
    if (k + S - 1 < steps) issue(k + S - 1);
    const int s = (int)(k % S);
    mbar_wait(bars + s, (uint32_t)((k / S) & 1));
    const unsigned char *p = slots + s * 2048 + lane * 64;
#pragma unroll
    for (uint32_t c = 0; c < 4; c++) {
      const float4 x = *reinterpret_cast<const float4 *>(p + ((c ^ sw) << 4));
      acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
    }
    __syncwarp();  // every lane has read the stage before it is refilled (generic reads -> async writes)
  }
  out[blockIdx.x * (int64_t)blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

// ---- variant 2: tile::gather4 -- lane issues ONE instruction for 4 rows (256 B); a warp step = 128 entries ---------------
// LANES = lanes of the warp that issue (32: 128 entries per step, 8 KB per stage; 8: 32 entries per step, 2 KB per stage)
template <int S, int LANES>
__global__ void __launch_bounds__(512) gather4_kernel(const uint32_t *__restrict__ idx, int64_t n,
                                                      const __grid_constant__ CUtensorMap tmap, float *__restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int STAGE = LANES * 256;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  unsigned char *slots = smem + (size_t)w * S * STAGE;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)nw * S * STAGE) + w * S;
  if (lane == 0)
    for (int s = 0; s < S; s++) mbar_init(bars + s, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const int64_t warp = blockIdx.x * (int64_t)nw + w, nwarps = (int64_t)gridDim.x * nw;
  constexpr int PER = LANES * 4;
  const int64_t steps = (n / PER - warp + nwarps - 1) / nwarps;
  float4 acc = make_float4(0, 0, 0, 0);
  const uint32_t sw = (uint32_t)((lane >> 1) & 3);
  auto issue = [&](int64_t k) {
    const int s = (int)(k % S);
    if (lane == 0) mbar_expect_tx(bars + s, STAGE);
    __syncwarp();
    if (lane < LANES) {
      const uint4 c = __ldg(reinterpret_cast<const uint4 *>(idx + (warp + k * nwarps) * PER) + lane);
      gather4_g2s(slots + s * STAGE + lane * 256, &tmap, 0, (int)c.x, (int)c.y, (int)c.z, (int)c.w, bars + s);
    }
  };
  for (int64_t k = 0; k < S - 1 && k < steps; k++) issue(k);
  for (int64_t k = 0; k < steps; k++) {
    if (k + S - 1 < steps) issue(k + S - 1);
    const int s = (int)(k % S);
    mbar_wait(bars + s, (uint32_t)((k / S) & 1));
#pragma unroll
    for (int r = 0; r < PER / 32; r++) {
      const unsigned char *p = slots + s * STAGE + (r * 32 + lane) * 64;
#pragma unroll
      for (uint32_t c = 0; c < 4; c++) {
        const float4 x = *reinterpret_cast<const float4 *>(p + ((c ^ sw) << 4));
        acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
      }
    }
    __syncwarp();
  }
  out[blockIdx.x * (int64_t)blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

static double checksum(const std::vector<float> &v) {
  double s = 0;
  for (float x : v) s += x;
  return s;
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
  const int64_t n = argc > 1 ? atoll(argv[1]) : 23 * 1000 * 1000;
  const int64_t rows = argc > 2 ? atoll(argv[2]) : 232965;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, entries %lld, rows %lld (matrix %.1f MB)\n", prop.name, sms, (long long)n, (long long)rows,
         rows * 64 / 1e6);
  std::vector<uint32_t> h_idx((size_t)n);
  uint64_t st = 88172645463325252ull;
  for (auto &x : h_idx) {
    st ^= st << 13; st ^= st >> 7; st ^= st << 17;
    x = (uint32_t)(st % (uint64_t)rows);
  }
  std::vector<float> h_B((size_t)rows * 16);
  for (size_t i = 0; i < h_B.size(); i++) h_B[i] = (float)((i * 2654435761u) >> 20 & 1023) / 1024.f;
  // exact expected sum over all gathered rows (values are multiples of 2^-10: fp32 partial sums differ only by rounding)
  double expect = 0;
  {
    std::vector<double> rs((size_t)rows);
    for (int64_t r = 0; r < rows; r++) {
      double s = 0;
      for (int c = 0; c < 16; c++) s += h_B[(size_t)r * 16 + c];
      rs[r] = s;
    }
    for (int64_t i = 0; i < n / 128 * 128; i++) expect += rs[h_idx[i]];
  }
  uint32_t *d_idx;
  float *d_B, *d_out;
  const size_t out_n = (size_t)sms * 8 * 1024;
  CK(cudaMalloc(&d_idx, n * 4));
  CK(cudaMalloc(&d_B, h_B.size() * 4));
  CK(cudaMalloc(&d_out, out_n * 4));
  CK(cudaMemcpy(d_idx, h_idx.data(), n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_B, h_B.data(), h_B.size() * 4, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  std::vector<float> h_out(out_n);
  const int64_t n128 = n / 128 * 128;

  auto report = [&](const char *name, float ms, size_t threads) {
    CK(cudaMemcpy(h_out.data(), d_out, threads * 4, cudaMemcpyDeviceToHost));
    double s = 0;
    for (size_t i = 0; i < threads; i++) s += h_out[i];
    const double clk = 1.9e9;
    printf("%-34s %8.1f us  %6.3f entries/clk/SM(@1.9GHz)  %7.1f GB/s gathered  sum rel err %.2e\n", name, ms * 1e3,
           n128 / (ms * 1e-3) / clk / sms, n128 * 64.0 / (ms * 1e-3) / 1e9, (s - expect) / expect);
  };
#define TIME(name, threads, launch)                          \
  do {                                                       \
    CK(cudaMemset(d_out, 0, out_n * 4));                     \
    launch;                                                  \
    CK(cudaGetLastError());                                  \
    CK(cudaDeviceSynchronize());                             \
    CK(cudaEventRecord(e0));                                 \
    for (int it = 0; it < 5; it++) launch;                   \
    CK(cudaEventRecord(e1));                                 \
    CK(cudaDeviceSynchronize());                             \
    float ms;                                                \
    CK(cudaEventElapsedTime(&ms, e0, e1));                   \
    report(name, ms / 5, threads);                           \
  } while (0)

  for (int bps : {4, 8}) {
    char nm[64];
    snprintf(nm, 64, "ldg 256thr x%d/SM", bps);
    TIME(nm, (size_t)sms * bps * 256, (ldg_kernel<<<sms * bps, 256>>>(d_idx, n128, d_B, d_out)));
  }
#define BULK(S, NT, BPS)                                                                                         \
  do {                                                                                                           \
    const size_t sm = (size_t)(NT / 32) * S * 2048 + (NT / 32) * S * 8;                                          \
    CK(cudaFuncSetAttribute(bulk_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));              \
    char nm[64];                                                                                                 \
    snprintf(nm, 64, "bulk64 S=%d thr=%d x%d/SM (%zu KB)", S, NT, BPS, sm * BPS / 1024);                          \
    TIME(nm, (size_t)sms * BPS * NT, (bulk_kernel<S><<<sms * BPS, NT, sm>>>(d_idx, n128, d_B, d_out)));          \
  } while (0)
  BULK(2, 256, 1);
  BULK(4, 256, 1);
  BULK(8, 256, 1);
  BULK(4, 512, 1);
  BULK(2, 512, 1);
  BULK(2, 128, 1);
  BULK(4, 128, 1);

  // ---- gather4 ----
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres));
  if (!encode) {
    printf("no cuTensorMapEncodeTiled\n");
    return 0;
  }
  for (int boxrows : {1, 4}) {
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {16, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {64};
    const cuuint32_t box[2] = {16, (cuuint32_t)boxrows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_B, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("tensor map box {16,%d}: encode rc=%d\n", boxrows, (int)r);
    if (r != CUDA_SUCCESS) continue;
#define G4(S, LANES, NT, BPS)                                                                                          \
  do {                                                                                                                 \
    const size_t sm = (size_t)(NT / 32) * S * LANES * 256 + (NT / 32) * S * 8;                                         \
    CK(cudaFuncSetAttribute(gather4_kernel<S, LANES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));          \
    char nm[64];                                                                                                       \
    snprintf(nm, 64, "gather4 S=%d lanes=%d thr=%d x%d (%zu KB)", S, LANES, NT, BPS, sm * BPS / 1024);                  \
    TIME(nm, (size_t)sms * BPS * NT, (gather4_kernel<S, LANES><<<sms * BPS, NT, sm>>>(d_idx, n128, tmap, d_out)));     \
  } while (0)
    G4(2, 8, 256, 1);
    G4(4, 8, 256, 1);
    G4(8, 8, 256, 1);
    G4(4, 8, 512, 1);
    G4(2, 32, 256, 1);
    G4(3, 32, 256, 1);
    G4(2, 8, 128, 1);
    G4(4, 8, 128, 1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("gather4 with box rows %d failed: %s\n", boxrows, cudaGetErrorString(e));
      return 0;
    }
  }
  return 0;
}
