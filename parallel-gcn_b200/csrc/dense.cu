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

// Weight gradient of a NARROW layer (n <= 32 inputs, p <= 64 outputs: the 16 -> 41 layer of the Part-1 model), where the
// 64 x 64 tiles of the kernel above are 16 % full: dB[n x p] = A[m x n]^T * dC[m x p] is an outer product per row, 53 MB of
// input for 656 outputs.  CTA = 4 groups of 64 threads; thread (grp, j) keeps column j of dB (n sums) in registers and
// walks the rows r = grp (mod 4) of its CTA's row slab: one LDS for dC[r][j], n/4 broadcast LDS.128 for A[r][*], n FMA.
// Groups are added in order through shared memory, CTAs by slab_reduce_kernel: fixed order, deterministic.
template <int NMAX>
__global__ void __launch_bounds__(256)
tn_narrow_kernel(const float *__restrict__ A, const float *__restrict__ dC, float *__restrict__ ws, int64_t m, int n, int p,
                 int64_t rows_per_cta) {
  constexpr int TR = 64;
  __shared__ __align__(16) float As[TR][NMAX];
  __shared__ float Gs[TR][64];
  __shared__ float red[3][NMAX][64];  // groups 1..3; group 0 adds them to its registers
  const int tid = threadIdx.x, grp = tid >> 6, j = tid & 63;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta, r1 = min(m, r0 + rows_per_cta);
  float acc[NMAX];
#pragma unroll
  for (int i = 0; i < NMAX; i++) acc[i] = 0.f;
  for (int64_t base = r0; base < r1; base += TR) {
    const int nr = (int)min((int64_t)TR, r1 - base);
    for (int e = tid; e < TR * NMAX; e += 256) {
      const int r = e / NMAX, i = e % NMAX;
      As[r][i] = (r < nr && i < n) ? __ldg(A + (size_t)(base + r) * n + i) : 0.f;
    }
    for (int e = tid; e < TR * 64; e += 256) {
      const int r = e >> 6, c = e & 63;
      Gs[r][c] = (r < nr && c < p) ? __ldg(dC + (size_t)(base + r) * p + c) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = grp; r < TR; r += 4) {  // rows beyond nr hold zeros
      const float gval = Gs[r][j];
#pragma unroll
      for (int i4 = 0; i4 < NMAX / 4; i4++) {
        const float4 a = *reinterpret_cast<const float4 *>(&As[r][i4 * 4]);
        acc[i4 * 4 + 0] = fmaf(a.x, gval, acc[i4 * 4 + 0]);
        acc[i4 * 4 + 1] = fmaf(a.y, gval, acc[i4 * 4 + 1]);
        acc[i4 * 4 + 2] = fmaf(a.z, gval, acc[i4 * 4 + 2]);
        acc[i4 * 4 + 3] = fmaf(a.w, gval, acc[i4 * 4 + 3]);
      }
    }
    __syncthreads();
  }
  if (grp > 0) {
#pragma unroll
    for (int i = 0; i < NMAX; i++) red[grp - 1][i][j] = acc[i];
  }
  __syncthreads();
  if (grp == 0 && j < p) {
    float *dst = ws + (size_t)blockIdx.x * n * p;
#pragma unroll
    for (int i = 0; i < NMAX; i++)
      if (i < n) dst[(size_t)i * p + j] = ((acc[i] + red[0][i][j]) + red[1][i][j]) + red[2][i][j];
  }
}

// dB = sum over slabs, ascending (fixed order)
__global__ void __launch_bounds__(256) slab_reduce_kernel(const float *__restrict__ ws, float *__restrict__ out, int64_t elems, int slabs) {
  // 32 outputs x 8 groups per CTA: group g adds the partials g, g + 8, ... in ascending order, then the groups are added in
  // order -- a fixed tree (deterministic) with 8x the parallelism and 1/8 of the dependent chain of one thread per output
  // (9632 outputs x ~300 partials took 52 us that way)
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  for (int64_t base = (int64_t)blockIdx.x * 32; base < elems; base += (int64_t)gridDim.x * 32) {
    const int64_t i = base + lane;
    float s = 0.f;
    if (i < elems)
      for (int z = g; z < slabs; z += 8) s += __ldg(ws + (size_t)z * elems + i);
    red[g][lane] = s;
    __syncthreads();
    if (g == 0 && i < elems) {
      float t = red[0][lane];
#pragma unroll
      for (int k = 1; k < 8; k++) t += red[k][lane];
      out[i] = t;
    }
    __syncthreads();
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
  // narrow shapes (tn_narrow_kernel) walk 64 rows per iteration with synchronous tile loads: on a small matrix one iteration
  // per CTA keeps the kernel off the critical path of the epoch (cora, 2708 rows: 11 CTAs x 4 iterations took 22 us)
  const int64_t min_slab = (n <= 32 && p <= 64) ? 64 : 256;
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
  if (n <= 32 && p <= 64 && m >= 512) {  // (the 64 x 64 split-K tiles are 16 % full here: 73 us on cora's 2708 x 16 x 7)
    const int sm = std::max(1, device_info().sm_count);
    const int ctas = (int)std::min<int64_t>(std::min<int64_t>(4 * sm, sh.slabs), (m + 63) / 64);  // 4 per SM: the tile loads are synchronous
    const int64_t rows_per_cta = ((m + ctas - 1) / ctas + 3) / 4 * 4;
    const int used = (int)((m + rows_per_cta - 1) / rows_per_cta);
    if (n <= 16) tn_narrow_kernel<16><<<used, 256, 0, stream>>>(d_A, d_dC, (float *)d_ws, m, n, p, rows_per_cta);
    else tn_narrow_kernel<32><<<used, 256, 0, stream>>>(d_A, d_dC, (float *)d_ws, m, n, p, rows_per_cta);
    GCNB_LAUNCH_CHECK();
    const int64_t elems = (int64_t)n * p;
    slab_reduce_kernel<<<(int)std::min<int64_t>((elems + 31) / 32, 4096), 256, 0, stream>>>((const float *)d_ws, d_dB, elems,
                                                                                              used);
    GCNB_LAUNCH_CHECK();
    return 0;
  }
  constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;
  dim3 grid(sh.tiles_m, sh.tiles_n, sh.slabs);
  sgemm_kernel<BM, BN, BK, TM, TN, MODE_TN><<<grid, 256, 0, stream>>>(d_A, d_dC, (float *)d_ws, n, p, m, sh.k_slab);
  GCNB_LAUNCH_CHECK();
  const int64_t elems = (int64_t)n * p;
  const int blocks = (int)std::min<int64_t>((elems + 31) / 32, 4096);
  slab_reduce_kernel<<<blocks, 256, 0, stream>>>((const float *)d_ws, d_dB, elems, sh.slabs);
  GCNB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
