// dense.cu -- tall-skinny fp32 products of the Matmul module (and of SparseMatmul when the feature matrix is
// dense, as Reddit's 602 columns are).  Replaces matmul_kernel_forward / _backward_1 / _backward_2
// (src/module.cu:274-391): 16x16 one-thread-per-element tiles there and atomicAdd for the weight gradient;
// here a register-tiled FFMA kernel (plain TF32 would miss the fp32 parity bar of 1e-5; a 3xTF32 mma.sync variant of
// this tiling was measured for the wide model -- DESIGN.md section 8 -- and is not kept: without a cp.async pipeline it
// is no faster than the FFMA kernel once its accumulation chains are bounded) and a split-K weight-gradient
// kernel whose slab partials are reduced in ascending order by a second pass (deterministic).
//
// One abstract problem  C[M x N] = sum_k A'(i,k) * B'(k,j)  with three operand layouts:
//   NN: A'(i,k)=A[i*K+k]      B'(k,j)=B[k*N+j]   (M=m rows, N=p, K=n)
//   NT: A'(i,k)=dC[i*K+k]     B'(k,j)=B[j*K+k]   (M=m rows, N=n, K=p)
//   TN: A'(i,k)=A[k*M+i]      B'(k,j)=dC[k*N+j]  (M=n, N=p, K=m rows, split over slabs of K)
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

using namespace gcnb;

namespace {

enum { MODE_NN = 0, MODE_NT = 1, MODE_TN = 2 };

// BM x BN block tile, BK k-step, 256 threads as (BM/TM) x (BN/TN) micro-tiles.
template <int BM, int BN, int BK, int TM, int TN, int MODE>
__global__ void __launch_bounds__(256)
sgemm_kernel(const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ C, int64_t M, int N,
             int64_t K, int64_t k_slab) {
  static_assert((BM / TM) * (BN / TN) == 256, "thread tiling");
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int64_t i0 = (int64_t)blockIdx.x * BM;  // row tiles on grid.x (2^31-1 limit; the reference's grid.y caps m at 1M rows)
  const int j0 = blockIdx.y * BN;
  int64_t kb = 0, ke = K;
  if (MODE == MODE_TN) {
    kb = (int64_t)blockIdx.z * k_slab;
    ke = min(K, kb + k_slab);
    C += (size_t)blockIdx.z * M * N;
  }
  float acc[TM][TN];
#pragma unroll
  for (int r = 0; r < TM; r++)
#pragma unroll
    for (int c = 0; c < TN; c++) acc[r][c] = 0.f;

  for (int64_t k0 = kb; k0 < ke; k0 += BK) {
    // ---- A' tile -> As[k][i]
    if (MODE == MODE_TN) {
      // contiguous along i
      for (int e = tid; e < BM * BK; e += 256) {
        const int i = e % BM, k = e / BM;
        const int64_t gi = i0 + i, gk = k0 + k;
        As[k][i] = (gi < M && gk < ke) ? __ldg(A + (size_t)gk * M + gi) : 0.f;
      }
    } else {
      // contiguous along k
      for (int e = tid; e < BM * BK; e += 256) {
        const int k = e % BK, i = e / BK;
        const int64_t gi = i0 + i, gk = k0 + k;
        As[k][i] = (gi < M && gk < ke) ? __ldg(A + (size_t)gi * K + gk) : 0.f;
      }
    }
    // ---- B' tile -> Bs[k][j]
    if (MODE == MODE_NT) {
      for (int e = tid; e < BN * BK; e += 256) {
        const int k = e % BK, j = e / BK;
        const int gj = j0 + j;
        const int64_t gk = k0 + k;
        Bs[k][j] = (gj < N && gk < ke) ? __ldg(B + (size_t)gj * K + gk) : 0.f;
      }
    } else {
      for (int e = tid; e < BN * BK; e += 256) {
        const int j = e % BN, k = e / BN;
        const int gj = j0 + j;
        const int64_t gk = k0 + k;
        Bs[k][j] = (gj < N && gk < ke) ? __ldg(B + (size_t)gk * N + gj) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; k++) {
      float a[TM], b[TN];
#pragma unroll
      for (int r = 0; r < TM; r++) a[r] = As[k][ty * TM + r];
#pragma unroll
      for (int c = 0; c < TN; c++) b[c] = Bs[k][tx * TN + c];
#pragma unroll
      for (int r = 0; r < TM; r++)
#pragma unroll
        for (int c = 0; c < TN; c++) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < TM; r++) {
    const int64_t gi = i0 + ty * TM + r;
    if (gi >= M) continue;
#pragma unroll
    for (int c = 0; c < TN; c++) {
      const int gj = j0 + tx * TN + c;
      if (gj < N) C[(size_t)gi * N + gj] = acc[r][c];
    }
  }
}

// dB = sum over slabs, ascending (fixed order)
__global__ void slab_reduce_kernel(const float *__restrict__ ws, float *__restrict__ out, int64_t elems, int slabs) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < elems; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < slabs; z++) s += __ldg(ws + (size_t)z * elems + i);
    out[i] = s;
  }
}

template <int MODE>
int launch_rows(const float *A, const float *B, float *C, int64_t M, int N, int64_t K, cudaStream_t stream) {
  // M = number of node rows (huge), N small
  if (N <= 16) {
    constexpr int BM = 128, BN = 16, BK = 16, TM = 8, TN = 1;
    dim3 grid((unsigned)((M + BM - 1) / BM), (N + BN - 1) / BN);
    sgemm_kernel<BM, BN, BK, TM, TN, MODE><<<grid, 256, 0, stream>>>(A, B, C, M, N, K, 0);
  } else if (N <= 32) {
    constexpr int BM = 128, BN = 32, BK = 16, TM = 8, TN = 2;
    dim3 grid((unsigned)((M + BM - 1) / BM), (N + BN - 1) / BN);
    sgemm_kernel<BM, BN, BK, TM, TN, MODE><<<grid, 256, 0, stream>>>(A, B, C, M, N, K, 0);
  } else {
    constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN = 4;
    dim3 grid((unsigned)((M + BM - 1) / BM), (N + BN - 1) / BN);
    sgemm_kernel<BM, BN, BK, TM, TN, MODE><<<grid, 256, 0, stream>>>(A, B, C, M, N, K, 0);
  }
  GCNB_LAUNCH_CHECK();
  return 0;
}

struct TnShape {
  int tiles_m, tiles_n, slabs;
  int64_t k_slab;
};
TnShape tn_shape(int64_t m, int n, int p) {
  TnShape s;
  s.tiles_m = (n + 63) / 64;
  s.tiles_n = (p + 63) / 64;
  const int tiles = s.tiles_m * s.tiles_n;
  const int sm = std::max(1, device_info().sm_count);
  int slabs = std::max(1, (4 * sm + tiles - 1) / tiles);
  const int64_t min_slab = 256;
  slabs = (int)std::max<int64_t>(1, std::min<int64_t>(slabs, (m + min_slab - 1) / min_slab));
  s.k_slab = ((m + slabs - 1) / slabs + 15) / 16 * 16;
  s.slabs = (int)((m + s.k_slab - 1) / s.k_slab);
  if (s.slabs < 1) s.slabs = 1;
  return s;
}

}  // namespace

extern "C" {

int gcnb_matmul_nn_f32(const float *d_A, const float *d_B, float *d_C, int64_t m, int n, int p, gcnb_stream_t s) {
  if (!d_A || !d_B || !d_C || m < 0 || n <= 0 || p <= 0) return GCNB_E_BADARG;
  if (m == 0) return 0;
  return launch_rows<MODE_NN>(d_A, d_B, d_C, m, p, n, as_stream(s));
}

int gcnb_matmul_nt_f32(const float *d_dC, const float *d_B, float *d_dA, int64_t m, int n, int p, gcnb_stream_t s) {
  if (!d_dC || !d_B || !d_dA || m < 0 || n <= 0 || p <= 0) return GCNB_E_BADARG;
  if (m == 0) return 0;
  return launch_rows<MODE_NT>(d_dC, d_B, d_dA, m, n, p, as_stream(s));
}

int64_t gcnb_matmul_tn_workspace(int64_t m, int n, int p) {
  const TnShape s = tn_shape(m, n, p);
  return (int64_t)s.slabs * n * p * (int64_t)sizeof(float);
}

int gcnb_matmul_tn_f32(const float *d_A, const float *d_dC, float *d_dB, int64_t m, int n, int p, void *d_ws,
                       int64_t ws_bytes, gcnb_stream_t s) {
  if (!d_A || !d_dC || !d_dB || m < 0 || n <= 0 || p <= 0) return GCNB_E_BADARG;
  cudaStream_t stream = as_stream(s);
  if (m == 0) {
    GCNB_CHECK(cudaMemsetAsync(d_dB, 0, (size_t)n * p * 4, stream));
    return 0;
  }
  const TnShape sh = tn_shape(m, n, p);
  if (!d_ws || ws_bytes < (int64_t)sh.slabs * n * p * 4) return GCNB_E_BADARG;
  constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;
  dim3 grid(sh.tiles_m, sh.tiles_n, sh.slabs);
  sgemm_kernel<BM, BN, BK, TM, TN, MODE_TN><<<grid, 256, 0, stream>>>(d_A, d_dC, (float *)d_ws, n, p, m, sh.k_slab);
  GCNB_LAUNCH_CHECK();
  const int64_t elems = (int64_t)n * p;
  const int blocks = (int)std::min<int64_t>((elems + 255) / 256, 4096);
  slab_reduce_kernel<<<blocks, 256, 0, stream>>>((const float *)d_ws, d_dB, elems, sh.slabs);
  GCNB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
