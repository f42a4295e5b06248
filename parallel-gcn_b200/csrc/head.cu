// head.cu -- the output head of the model in ONE kernel (sm_100a).
//
// With the (A_hat a) W association of the last layer (in_dim < classes, e.g. 16 -> 41) everything after the narrow
// GraphSum is local to a node row:
//     logits z = y W                       Matmul::forward          src/module.cu:274-317
//     masked softmax cross-entropy, counts CrossEntropyLoss::forward src/module.cu:484-541, get_accuracy_kernel src/gcn.cu:264-289
//     dy = dz W^T, dW = y^T dz             Matmul::backward         src/module.cu:319-391, :456-472
// The reference runs these as five kernels with three [N x C] round trips through HBM (logits, gradient, gradient again).
// Here lane l of a warp owns row l of a 32-row tile: y comes straight into registers (one 64-byte row per lane), W and W^T
// sit in shared memory and are read as broadcast LDS.128, the logits tile lives in shared memory (odd row stride:
// conflict-free), leaves once through coalesced 128-bit stores (the CE-shifted logits are API-visible, gcnb_gcn_get_logits)
// and is overwritten in place by the exponentials and then by the gradient, which never reaches HBM unless the caller asks.
// dW: lane j of the warp owns columns j and j + 32 of dW and walks the labelled rows of the tile (y broadcast from shared
// memory); warps keep their sums in registers over all their tiles, the CTA adds its warps in warp order, and
// gcnb_head_reduce_dw_f32 adds the CTAs in ascending order (fixed tile -> warp assignment: deterministic).
// The arithmetic of every element is the one of sgemm_kernel (ascending-k fmaf chains, dense.cu) and of
// softmax_ce_rows_kernel (loss.cu), so the fused and the unfused paths agree bit for bit in logits, loss and dy.
// head_tc_kernel (below; in_dim 16, classes <= 48; gcnb_head_tc_f32, the engine's GCNB_HEAD_TC=1) runs the three products
// on the tensor cores instead (split TF32, fp32-level accuracy but not the same bits): 59 / 37 us against 75 / 45 us on the
// Reddit-shape graph -- both variants are now bound by latency at 16 warps per SM, not by the products.  The FMA kernel
// stays the default because it keeps the engine's results identical to the unfused kernels': any other rounding moves a
// handful of first-layer weights whose first Adam step is ill-conditioned (|g| ~ eps) past the 1e-6 floor of the engine test.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

using namespace gcnb;

namespace {

constexpr int kT = 256;
constexpr int kWarps = kT / 32;
constexpr int kMaxC = 64;
constexpr int kMaxBlocks = 1024;

__host__ __device__ inline int head_warp_floats(int in, int C) {
  const int tile = 32 * (C | 1) + 32 * in;  // logits tile + y tile
  const int fin = in * 64;                   // the warp's dW sums at the end
  return tile > fin ? tile : fin;
}

template <int IN, int CTAS>
__global__ void __launch_bounds__(kT, CTAS)
head_kernel(const float *__restrict__ Y, const float *__restrict__ W, const int32_t *__restrict__ truth, int64_t n, int C,
            uint32_t num_samples, int training, int aligned16, uint32_t div_magic, float *__restrict__ logits,
            float *__restrict__ grad, float *__restrict__ dY, float *__restrict__ dw_part, float *__restrict__ result,
            float *__restrict__ part_loss, uint32_t *__restrict__ part_cnt, unsigned int *__restrict__ ticket) {
  extern __shared__ __align__(16) float smem[];
  constexpr int IN4 = IN / 4;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int CW = (C + 3) & ~3, CP = C | 1;
  float *Ws = smem;            // [IN][CW], columns beyond C are zero
  float *Wt = Ws + IN * CW;    // [CW][IN] (rows beyond C unused): W transposed, training only
  float *zt = Wt + CW * IN + (size_t)wib * head_warp_floats(IN, C);  // the warp's logits tile [32][CP]
  float *yt = zt + 32 * CP;                                           // the warp's y tile [32][IN]
  for (int e = threadIdx.x; e < IN * CW; e += kT) {
    const int k = e / CW, j = e - k * CW;
    Ws[e] = j < C ? __ldg(W + (size_t)k * C + j) : 0.f;
  }
  if (training)
    for (int e = threadIdx.x; e < C * IN; e += kT) {
      const int j = e / IN, k = e - j * IN;
      Wt[e] = __ldg(W + (size_t)k * C + j);
    }
  __syncthreads();

  const int64_t ntiles = (n + 31) / 32;
  const int64_t gw = (int64_t)blockIdx.x * kWarps + wib, nw = (int64_t)gridDim.x * kWarps;
  const float inv_ns = 1.0f / (float)num_samples;
  const bool contiguous = CP == C && aligned16;
  float acc0[IN], acc1[IN];  // dW[k][lane], dW[k][lane + 32]
#pragma unroll
  for (int k = 0; k < IN; k++) acc0[k] = acc1[k] = 0.f;
  float loss = 0.f;
  uint32_t wrong = 0, labelled = 0;

  for (int64_t tile = gw; tile < ntiles; tile += nw) {
    const int64_t r0 = tile * 32;
    const int rows = (int)min((int64_t)32, n - r0);
    const int total = rows * C;
    float y[IN];
    int t = -1;
    {
      float4 yn[IN4];
      if (lane < rows) {
        const float4 *src = reinterpret_cast<const float4 *>(Y + (size_t)(r0 + lane) * IN);
#pragma unroll
        for (int c = 0; c < IN4; c++) yn[c] = __ldg(src + c);
        t = __ldg(truth + r0 + lane);
      } else {
#pragma unroll
        for (int c = 0; c < IN4; c++) yn[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (training) {
        float4 *dst = reinterpret_cast<float4 *>(yt + lane * IN);
#pragma unroll
        for (int c = 0; c < IN4; c++) dst[c] = yn[c];
      }
#pragma unroll
      for (int c = 0; c < IN4; c++) {
        y[4 * c] = yn[c].x;
        y[4 * c + 1] = yn[c].y;
        y[4 * c + 2] = yn[c].z;
        y[4 * c + 3] = yn[c].w;
      }
    }

    // ---- z = y W: ascending-k fmaf chains (the chains of sgemm_kernel)
    float *zr = zt + lane * CP;
    float mx = -INFINITY;  // row maximum, gathered on the way
    int j4 = 0;
    for (; j4 + 8 <= CW; j4 += 8) {  // eight columns at a time: eight independent chains
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < IN; k++) {
        const float4 w = *reinterpret_cast<const float4 *>(Ws + k * CW + j4);
        const float4 v = *reinterpret_cast<const float4 *>(Ws + k * CW + j4 + 4);
        a.x = fmaf(y[k], w.x, a.x);
        a.y = fmaf(y[k], w.y, a.y);
        a.z = fmaf(y[k], w.z, a.z);
        a.w = fmaf(y[k], w.w, a.w);
        b.x = fmaf(y[k], v.x, b.x);
        b.y = fmaf(y[k], v.y, b.y);
        b.z = fmaf(y[k], v.z, b.z);
        b.w = fmaf(y[k], v.w, b.w);
      }
      zr[j4] = a.x;
      zr[j4 + 1] = a.y;
      zr[j4 + 2] = a.z;
      zr[j4 + 3] = a.w;
      mx = fmaxf(fmaxf(fmaxf(mx, a.x), fmaxf(a.y, a.z)), a.w);
      zr[j4 + 4] = b.x;
      mx = fmaxf(mx, b.x);
      if (j4 + 5 < C) zr[j4 + 5] = b.y, mx = fmaxf(mx, b.y);
      if (j4 + 6 < C) zr[j4 + 6] = b.z, mx = fmaxf(mx, b.z);
      if (j4 + 7 < C) zr[j4 + 7] = b.w, mx = fmaxf(mx, b.w);
    }
    if (j4 < CW) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < IN; k++) {
        const float4 w = *reinterpret_cast<const float4 *>(Ws + k * CW + j4);
        a.x = fmaf(y[k], w.x, a.x);
        a.y = fmaf(y[k], w.y, a.y);
        a.z = fmaf(y[k], w.z, a.z);
        a.w = fmaf(y[k], w.w, a.w);
      }
      zr[j4] = a.x;
      mx = fmaxf(mx, a.x);
      if (j4 + 1 < C) zr[j4 + 1] = a.y, mx = fmaxf(mx, a.y);
      if (j4 + 2 < C) zr[j4 + 2] = a.z, mx = fmaxf(mx, a.z);
      if (j4 + 3 < C) zr[j4 + 3] = a.w, mx = fmaxf(mx, a.w);
    }
    // ---- shift by the row maximum (labelled rows; written back: API-visible side effect of the reference)
    float xt = 0.f;
    if (t >= 0) {
#pragma unroll 4
      for (int j = 0; j < C; j++) zr[j] -= mx;
      xt = zr[t];
    }
    __syncwarp();
    {  // coalesced copy-out of the tile
      float *lg = logits + (size_t)r0 * C;
      if (contiguous) {
        const int n4 = total >> 2;
        float4 *d0 = reinterpret_cast<float4 *>(lg);
        const float4 *s0 = reinterpret_cast<const float4 *>(zt);
#pragma unroll 4
        for (int i = lane; i < n4; i += 32) d0[i] = s0[i];
        for (int i = (n4 << 2) + lane; i < total; i += 32) lg[i] = zt[i];
      } else {
#pragma unroll 4
        for (int i = lane; i < total; i += 32) {
          const int row = (int)(((uint32_t)i * div_magic) >> 17);
          lg[i] = zt[row * CP + (i - row * C)];
        }
      }
    }
    __syncwarp();
    // ---- exponentials, loss, counts, gradient: in place (the arithmetic of softmax_ce_rows_kernel)
    float sum = 0.f;
    if (t >= 0) {
#pragma unroll 4
      for (int j = 0; j < C; j++) {
        const float e = expf(zr[j]);
        zr[j] = e;
        sum += e;
      }
      loss += logf(sum) - xt;
      labelled++;
      if (xt < 0.f) wrong++;  // src/gcn.cu:273-276
    }
    if (training) {
      const unsigned lab = __ballot_sync(0xffffffffu, t >= 0);
      // ---- dz in place and dy = dz W^T (ascending-j fmaf chains) in the same walk; rows without a label: zero gradient
      if (lane < rows) {
        float d[IN];
#pragma unroll
        for (int k = 0; k < IN; k++) d[k] = 0.f;
        if (t >= 0) {
          const float inv = 1.0f / sum;
#pragma unroll 2
          for (int j = 0; j < C; j++) {
            float g = (zr[j] * inv) * inv_ns;
            if (j == t) g = (float)((double)g - 1.0 / (double)num_samples);  // double literal in the reference (src/module.cu:517)
            zr[j] = g;
            const float4 *w4 = reinterpret_cast<const float4 *>(Wt + j * IN);
#pragma unroll
            for (int c = 0; c < IN4; c++) {
              const float4 w = w4[c];
              d[4 * c] = fmaf(g, w.x, d[4 * c]);
              d[4 * c + 1] = fmaf(g, w.y, d[4 * c + 1]);
              d[4 * c + 2] = fmaf(g, w.z, d[4 * c + 2]);
              d[4 * c + 3] = fmaf(g, w.w, d[4 * c + 3]);
            }
          }
        }
        float4 *dst = reinterpret_cast<float4 *>(dY + (size_t)(r0 + lane) * IN);
#pragma unroll
        for (int c = 0; c < IN4; c++) dst[c] = make_float4(d[4 * c], d[4 * c + 1], d[4 * c + 2], d[4 * c + 3]);
      }
      if (grad) {  // the caller wants dz itself (module-level API); the engine does not
        if (t < 0)
          for (int j = 0; j < C; j++) zr[j] = 0.f;
        __syncwarp();
        float *gr = grad + (size_t)r0 * C;
        if (contiguous) {
          const int n4 = total >> 2;
          float4 *d1 = reinterpret_cast<float4 *>(gr);
          const float4 *s1 = reinterpret_cast<const float4 *>(zt);
#pragma unroll 4
          for (int i = lane; i < n4; i += 32) d1[i] = s1[i];
          for (int i = (n4 << 2) + lane; i < total; i += 32) gr[i] = zt[i];
        } else {
#pragma unroll 4
          for (int i = lane; i < total; i += 32) {
            const int row = (int)(((uint32_t)i * div_magic) >> 17);
            gr[i] = zt[row * CP + (i - row * C)];
          }
        }
      }
      __syncwarp();
      // ---- dW += y^T dz over the labelled rows of the tile, in ascending row order
      const bool two = C > 32;
      for (unsigned m = lab; m; m &= m - 1) {
        const int r = __ffs(m) - 1;
        const float d0 = lane < C ? zt[r * CP + lane] : 0.f;
        const float d1 = (two && lane + 32 < C) ? zt[r * CP + lane + 32] : 0.f;
        const float4 *yr = reinterpret_cast<const float4 *>(yt + r * IN);
#pragma unroll
        for (int c = 0; c < IN4; c++) {
          const float4 v = yr[c];
          acc0[4 * c] = fmaf(v.x, d0, acc0[4 * c]);
          acc0[4 * c + 1] = fmaf(v.y, d0, acc0[4 * c + 1]);
          acc0[4 * c + 2] = fmaf(v.z, d0, acc0[4 * c + 2]);
          acc0[4 * c + 3] = fmaf(v.w, d0, acc0[4 * c + 3]);
          if (two) {
            acc1[4 * c] = fmaf(v.x, d1, acc1[4 * c]);
            acc1[4 * c + 1] = fmaf(v.y, d1, acc1[4 * c + 1]);
            acc1[4 * c + 2] = fmaf(v.z, d1, acc1[4 * c + 2]);
            acc1[4 * c + 3] = fmaf(v.w, d1, acc1[4 * c + 3]);
          }
        }
      }
    }
    __syncwarp();  // the next tile overwrites zt / yt
  }

  // ---- dW: warps -> CTA in warp order -> one partial block per CTA
  if (training) {
    float *fin = zt;  // [IN][64]
#pragma unroll
    for (int k = 0; k < IN; k++) {
      fin[k * 64 + lane] = acc0[k];
      fin[k * 64 + 32 + lane] = acc1[k];
    }
    __syncthreads();
    const float *base = Wt + CW * IN;
    const int stride = head_warp_floats(IN, C);
    float *dst = dw_part + (size_t)blockIdx.x * IN * C;
    for (int e = threadIdx.x; e < IN * C; e += kT) {
      const int k = e / C, j = e - k * C;
      float s = base[k * 64 + j];
#pragma unroll
      for (int w = 1; w < kWarps; w++) s += base[(size_t)w * stride + k * 64 + j];
      dst[e] = s;
    }
  }
  // ---- loss / counts: lanes -> warp (fixed shuffle tree) -> CTA in warp order -> ascending CTA order by the last CTA
  loss = warp_sum(loss);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    wrong += __shfl_xor_sync(0xffffffffu, wrong, o);
    labelled += __shfl_xor_sync(0xffffffffu, labelled, o);
  }
  __shared__ float s_loss[kWarps];
  __shared__ uint32_t s_wrong[kWarps], s_lab[kWarps];
  __shared__ bool last;
  if (lane == 0) {
    s_loss[wib] = loss;
    s_wrong[wib] = wrong;
    s_lab[wib] = labelled;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float bl = 0.f;
    uint32_t bw = 0, bb = 0;
    for (int i = 0; i < kWarps; i++) {
      bl += s_loss[i];
      bw += s_wrong[i];
      bb += s_lab[i];
    }
    part_loss[blockIdx.x] = bl;
    part_cnt[2 * blockIdx.x] = bw;
    part_cnt[2 * blockIdx.x + 1] = bb;
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    float tl = 0.f;
    uint32_t tw = 0, tb = 0;
    for (unsigned i = 0; i < gridDim.x; i++) {
      tl += ((volatile float *)part_loss)[i];
      tw += ((volatile uint32_t *)part_cnt)[2 * i];
      tb += ((volatile uint32_t *)part_cnt)[2 * i + 1];
    }
    result[0] = tl;
    result[1] = __uint_as_float(tw);
    result[2] = __uint_as_float(tb);
    *ticket = 0;
  }
}

// ---- tensor-core variant (in_dim 16, classes <= 48): the three products as mma.sync m16n8k8 TF32 with the 3xTF32 split
// (x = hi + lo with both pieces rounded to tf32, all four piece products, fp32 accumulate: fp32-level accuracy).  The FMA kernel
// above is bound by shared memory -> register bandwidth (a broadcast LDS.128 of W per four FMAs) and by issue slots (ncu:
// 7200 warp instructions per 32-row tile); here W lives in shared memory as ready-made hi / lo B fragments (one conflict-free
// LDS.32 per fragment register), Y and dZ fragments come from the warp's own tiles, and a tile costs 288 mma instead of
// ~2000 FMA + ~400 LDS.128.  Softmax, loss, counts and the copy-out are the lane-per-row code of the FMA kernel (the row
// maximum is taken from the accumulator fragments, so the logits reach shared memory already shifted).  Accumulators are
// kept short, as in dense_feat.cu: dW of a tile is computed in fresh registers and added to the running sums with a
// rounded FADD.
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo) {
  // both pieces ROUNDED to tf32 (add half an ulp of the 10-bit significand, clear the low 13 bits: two integer ops each):
  // |lo| <= 2^-11 |x| and the rounding of lo costs 2^-23 |x| -- with the lo * lo term kept (mma4) the products are as
  // accurate as fp32 FMAs.  (Truncation and three terms, as dense_feat.cu does for its 602-long sums, left 2^-20 per product
  // here: seven of 9632 first-layer weights moved past the 1e-6 floor of the engine test after three Adam steps.)
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = (__float_as_uint(x - __uint_as_float(hi)) + 0x1000u) & 0xffffe000u;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma4(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0, uint32_t bh1,
                                     uint32_t bl0, uint32_t bl1) {
  mma_tf32(c, al, bl0, bl1);  // small terms first
  mma_tf32(c, al, bh0, bh1);
  mma_tf32(c, ah, bl0, bl1);
  mma_tf32(c, ah, bh0, bh1);
}

constexpr int kYS = 20;  // row stride of the y tile: 80 bytes keep float4 stores aligned and the fragment loads conflict-free

__host__ __device__ inline int head_tc_warp_floats(int C) {
  const int tile = 32 * (C | 1) + 32 * kYS;
  return tile > 16 * 64 ? tile : 16 * 64;
}

template <int NT>  // column tiles of 8: classes <= 8 * NT
__global__ void __launch_bounds__(kT, 2)
head_tc_kernel(const float *__restrict__ Y, const float *__restrict__ W, const int32_t *__restrict__ truth, int64_t n, int C,
               uint32_t num_samples, int training, int aligned16, uint32_t div_magic, float *__restrict__ logits,
               float *__restrict__ grad, float *__restrict__ dY, float *__restrict__ dw_part, float *__restrict__ result,
               float *__restrict__ part_loss, uint32_t *__restrict__ part_cnt, unsigned int *__restrict__ ticket) {
  extern __shared__ __align__(16) float smem[];
  constexpr int IN = 16;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int g = lane >> 2, tig = lane & 3;
  const int CP = C | 1;
  // B fragments of W, pre-split.  z = Y W:      B[k][n] = W[k][n],  k = 8 s + tig + 4 q, n = 8 nt + g   -> wz[((s * NT + nt) * 2 + q) * 32 + lane]
  //                               dy = dZ W^T:  B[j][kk] = W[kk][j], j = 8 s + tig + 4 q, kk = 8 u + g  -> wd[((s * 2 + u) * 2 + q) * 32 + lane]
  uint32_t *wz_hi = reinterpret_cast<uint32_t *>(smem);
  uint32_t *wz_lo = wz_hi + 2 * NT * 2 * 32;
  uint32_t *wd_hi = wz_lo + 2 * NT * 2 * 32;
  uint32_t *wd_lo = wd_hi + NT * 2 * 2 * 32;
  float *zt = reinterpret_cast<float *>(wd_lo + NT * 2 * 2 * 32) + (size_t)wib * head_tc_warp_floats(C);  // logits tile [32][CP]
  float *yt = zt + 32 * CP;                                                                               // y tile [32][kYS]
  for (int e = threadIdx.x; e < 2 * NT * 2 * 32; e += kT) {
    const int l = e & 31, q = (e >> 5) & 1, nt = (e >> 6) % NT, s = (e >> 6) / NT;
    const int k = 8 * s + (l & 3) + 4 * q, col = 8 * nt + (l >> 2);
    uint32_t hi, lo;
    split_tf32(col < C ? __ldg(W + (size_t)k * C + col) : 0.f, hi, lo);
    wz_hi[e] = hi;
    wz_lo[e] = lo;
  }
  if (training)
    for (int e = threadIdx.x; e < NT * 2 * 2 * 32; e += kT) {
      const int l = e & 31, q = (e >> 5) & 1, u = (e >> 6) & 1, s = e >> 7;
      const int j = 8 * s + (l & 3) + 4 * q, kk = 8 * u + (l >> 2);
      uint32_t hi, lo;
      split_tf32(j < C ? __ldg(W + (size_t)kk * C + j) : 0.f, hi, lo);
      wd_hi[e] = hi;
      wd_lo[e] = lo;
    }
  __syncthreads();

  const int64_t ntiles = (n + 31) / 32;
  const int64_t gw = (int64_t)blockIdx.x * kWarps + wib, nw = (int64_t)gridDim.x * kWarps;
  const float inv_ns = 1.0f / (float)num_samples;
  const bool contiguous = CP == C && aligned16;
  float dw[NT][4];  // running dW fragments: (kk = g, g + 8) x (j = 8 nt + 2 tig, + 1)
#pragma unroll
  for (int nt = 0; nt < NT; nt++) dw[nt][0] = dw[nt][1] = dw[nt][2] = dw[nt][3] = 0.f;
  float loss = 0.f;
  uint32_t wrong = 0, labelled = 0;

  for (int64_t tile = gw; tile < ntiles; tile += nw) {
    const int64_t r0 = tile * 32;
    const int rows = (int)min((int64_t)32, n - r0);
    const int total = rows * C;
    int t = -1;
    {
      float4 *dst = reinterpret_cast<float4 *>(yt + lane * kYS);
      if (lane < rows) {
        const float4 *src = reinterpret_cast<const float4 *>(Y + (size_t)(r0 + lane) * IN);
#pragma unroll
        for (int c = 0; c < 4; c++) dst[c] = __ldg(src + c);
        t = __ldg(truth + r0 + lane);
      } else {
#pragma unroll
        for (int c = 0; c < 4; c++) dst[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    const unsigned lab = __ballot_sync(0xffffffffu, t >= 0);
    __syncwarp();
    // ---- z = Y W on the tensor cores: two m16 row tiles x NT column tiles, K = 16 in two steps
    {
      float acc[2][NT][4];
#pragma unroll
      for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
#pragma unroll
      for (int s = 0; s < 2; s++) {
        uint32_t ah[2][4], al[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
          const float *yr = yt + (16 * mt + g) * kYS + 8 * s + tig;
          split_tf32(yr[0], ah[mt][0], al[mt][0]);
          split_tf32(yr[8 * kYS], ah[mt][1], al[mt][1]);
          split_tf32(yr[4], ah[mt][2], al[mt][2]);
          split_tf32(yr[8 * kYS + 4], ah[mt][3], al[mt][3]);
        }
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
          const int o = ((s * NT + nt) * 2) * 32 + lane;
          const uint32_t bh0 = wz_hi[o], bh1 = wz_hi[o + 32], bl0 = wz_lo[o], bl1 = wz_lo[o + 32];
          mma4(acc[0][nt], ah[0], al[0], bh0, bh1, bl0, bl1);
          mma4(acc[1][nt], ah[1], al[1], bh0, bh1, bl0, bl1);
        }
      }
      // shift by the row maximum (labelled rows) while the logits are still in the fragments, then to shared memory
#pragma unroll
      for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int h = 0; h < 2; h++) {  // rows 16 mt + g (fragment registers 0, 1) and 16 mt + g + 8 (2, 3)
          const int row = 16 * mt + g + 8 * h;
          float mx = -INFINITY;
#pragma unroll
          for (int nt = 0; nt < NT; nt++) {
            const int col = 8 * nt + 2 * tig;
            if (col < C) mx = fmaxf(mx, acc[mt][nt][2 * h]);
            if (col + 1 < C) mx = fmaxf(mx, acc[mt][nt][2 * h + 1]);
          }
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
          const float sub = (lab >> row) & 1u ? mx : 0.f;
          float *zr = zt + row * CP + 2 * tig;
#pragma unroll
          for (int nt = 0; nt < NT; nt++) {
            const int col = 8 * nt + 2 * tig;
            if (col < C) zr[8 * nt] = acc[mt][nt][2 * h] - sub;
            if (col + 1 < C) zr[8 * nt + 1] = acc[mt][nt][2 * h + 1] - sub;
          }
        }
    }
    __syncwarp();
    float *zr = zt + lane * CP;
    const float xt = t >= 0 ? zr[t] : 0.f;
    {  // coalesced copy-out of the tile (the CE-shifted logits are API-visible)
      float *lg = logits + (size_t)r0 * C;
      if (contiguous) {
        const int n4 = total >> 2;
        float4 *d0 = reinterpret_cast<float4 *>(lg);
        const float4 *s0 = reinterpret_cast<const float4 *>(zt);
#pragma unroll 4
        for (int i = lane; i < n4; i += 32) d0[i] = s0[i];
        for (int i = (n4 << 2) + lane; i < total; i += 32) lg[i] = zt[i];
      } else {
#pragma unroll 4
        for (int i = lane; i < total; i += 32) {
          const int row = (int)(((uint32_t)i * div_magic) >> 17);
          lg[i] = zt[row * CP + (i - row * C)];
        }
      }
    }
    __syncwarp();
    // ---- exponentials, loss, counts, gradient: lane l owns row l, in place (the arithmetic of softmax_ce_rows_kernel)
    if (t >= 0) {
      float sum = 0.f;
#pragma unroll 4
      for (int j = 0; j < C; j++) {
        const float e = expf(zr[j]);
        zr[j] = e;
        sum += e;
      }
      loss += logf(sum) - xt;
      labelled++;
      if (xt < 0.f) wrong++;  // src/gcn.cu:273-276
      if (training) {
        const float inv = 1.0f / sum;
#pragma unroll 4
        for (int j = 0; j < C; j++) zr[j] = (zr[j] * inv) * inv_ns;
        zr[t] = (float)((double)zr[t] - 1.0 / (double)num_samples);  // double literal in the reference (src/module.cu:517)
      }
    }
    if (training) {
      __syncwarp();
      // ---- dy = dz W^T: K = the classes in NT steps of 8, N = 16 in two tiles; rows without a label contribute zeros
      {
        float acc[2][2][4];
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
          for (int u = 0; u < 2; u++) acc[mt][u][0] = acc[mt][u][1] = acc[mt][u][2] = acc[mt][u][3] = 0.f;
#pragma unroll
        for (int s = 0; s < NT; s++) {
          uint32_t ah[2][4], al[2][4];
          const int j0 = 8 * s + tig, j1 = j0 + 4;
#pragma unroll
          for (int mt = 0; mt < 2; mt++) {
            const int ra = 16 * mt + g, rb = ra + 8;
            const bool la = (lab >> ra) & 1u, lb = (lab >> rb) & 1u;
            split_tf32(la && j0 < C ? zt[ra * CP + j0] : 0.f, ah[mt][0], al[mt][0]);
            split_tf32(lb && j0 < C ? zt[rb * CP + j0] : 0.f, ah[mt][1], al[mt][1]);
            split_tf32(la && j1 < C ? zt[ra * CP + j1] : 0.f, ah[mt][2], al[mt][2]);
            split_tf32(lb && j1 < C ? zt[rb * CP + j1] : 0.f, ah[mt][3], al[mt][3]);
          }
#pragma unroll
          for (int u = 0; u < 2; u++) {
            const int o = ((s * 2 + u) * 2) * 32 + lane;
            const uint32_t bh0 = wd_hi[o], bh1 = wd_hi[o + 32], bl0 = wd_lo[o], bl1 = wd_lo[o + 32];
            mma4(acc[0][u], ah[0], al[0], bh0, bh1, bl0, bl1);
            mma4(acc[1][u], ah[1], al[1], bh0, bh1, bl0, bl1);
          }
        }
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const int row = 16 * mt + g + 8 * h;
            if (row < rows) {
              float2 *dst = reinterpret_cast<float2 *>(dY + (size_t)(r0 + row) * IN + 2 * tig);
              dst[0] = make_float2(acc[mt][0][2 * h], acc[mt][0][2 * h + 1]);
              dst[4] = make_float2(acc[mt][1][2 * h], acc[mt][1][2 * h + 1]);
            }
          }
      }
      // ---- dW += Y^T dz over the tile: M = 16, K = the 32 rows in four steps, N = the classes
      if (lab) {
        float tmp[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; nt++) tmp[nt][0] = tmp[nt][1] = tmp[nt][2] = tmp[nt][3] = 0.f;
#pragma unroll
        for (int s = 0; s < 4; s++) {
          const int ra = 8 * s + tig, rb = ra + 4;
          uint32_t ah[4], al[4];
          split_tf32(yt[ra * kYS + g], ah[0], al[0]);
          split_tf32(yt[ra * kYS + g + 8], ah[1], al[1]);
          split_tf32(yt[rb * kYS + g], ah[2], al[2]);
          split_tf32(yt[rb * kYS + g + 8], ah[3], al[3]);
          const bool la = (lab >> ra) & 1u, lb = (lab >> rb) & 1u;
#pragma unroll
          for (int nt = 0; nt < NT; nt++) {
            const int col = 8 * nt + g;
            uint32_t bh0, bl0, bh1, bl1;
            split_tf32(la && col < C ? zt[ra * CP + col] : 0.f, bh0, bl0);
            split_tf32(lb && col < C ? zt[rb * CP + col] : 0.f, bh1, bl1);
            mma4(tmp[nt], ah, al, bh0, bh1, bl0, bl1);
          }
        }
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
          dw[nt][0] += tmp[nt][0];
          dw[nt][1] += tmp[nt][1];
          dw[nt][2] += tmp[nt][2];
          dw[nt][3] += tmp[nt][3];
        }
      }
      if (grad) {  // the caller wants dz itself (module-level API); the engine does not
        __syncwarp();
        if (t < 0)
          for (int j = 0; j < C; j++) zr[j] = 0.f;
        __syncwarp();
        float *gr = grad + (size_t)r0 * C;
        if (contiguous) {
          const int n4 = total >> 2;
          float4 *d1 = reinterpret_cast<float4 *>(gr);
          const float4 *s1 = reinterpret_cast<const float4 *>(zt);
#pragma unroll 4
          for (int i = lane; i < n4; i += 32) d1[i] = s1[i];
          for (int i = (n4 << 2) + lane; i < total; i += 32) gr[i] = zt[i];
        } else {
#pragma unroll 4
          for (int i = lane; i < total; i += 32) {
            const int row = (int)(((uint32_t)i * div_magic) >> 17);
            gr[i] = zt[row * CP + (i - row * C)];
          }
        }
      }
    }
    __syncwarp();  // the next tile overwrites zt / yt
  }

  // ---- dW: fragments -> [16][64] per warp -> CTA in warp order -> one partial block per CTA
  if (training) {
    float *fin = zt;
#pragma unroll
    for (int nt = 0; nt < NT; nt++) {
      fin[g * 64 + 8 * nt + 2 * tig] = dw[nt][0];
      fin[g * 64 + 8 * nt + 2 * tig + 1] = dw[nt][1];
      fin[(g + 8) * 64 + 8 * nt + 2 * tig] = dw[nt][2];
      fin[(g + 8) * 64 + 8 * nt + 2 * tig + 1] = dw[nt][3];
    }
    __syncthreads();
    const float *base = reinterpret_cast<const float *>(wd_lo + NT * 2 * 2 * 32);
    const int stride = head_tc_warp_floats(C);
    float *dst = dw_part + (size_t)blockIdx.x * IN * C;
    for (int e = threadIdx.x; e < IN * C; e += kT) {
      const int k = e / C, j = e - k * C;
      float sacc = base[k * 64 + j];
#pragma unroll
      for (int w = 1; w < kWarps; w++) sacc += base[(size_t)w * stride + k * 64 + j];
      dst[e] = sacc;
    }
  }
  // ---- loss / counts: lanes -> warp (fixed shuffle tree) -> CTA in warp order -> ascending CTA order by the last CTA
  loss = warp_sum(loss);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    wrong += __shfl_xor_sync(0xffffffffu, wrong, o);
    labelled += __shfl_xor_sync(0xffffffffu, labelled, o);
  }
  __shared__ float s_loss[kWarps];
  __shared__ uint32_t s_wrong[kWarps], s_lab[kWarps];
  __shared__ bool last;
  if (lane == 0) {
    s_loss[wib] = loss;
    s_wrong[wib] = wrong;
    s_lab[wib] = labelled;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float bl = 0.f;
    uint32_t bw = 0, bb = 0;
    for (int i = 0; i < kWarps; i++) {
      bl += s_loss[i];
      bw += s_wrong[i];
      bb += s_lab[i];
    }
    part_loss[blockIdx.x] = bl;
    part_cnt[2 * blockIdx.x] = bw;
    part_cnt[2 * blockIdx.x + 1] = bb;
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    float tl = 0.f;
    uint32_t tw = 0, tb = 0;
    for (unsigned i = 0; i < gridDim.x; i++) {
      tl += ((volatile float *)part_loss)[i];
      tw += ((volatile uint32_t *)part_cnt)[2 * i];
      tb += ((volatile uint32_t *)part_cnt)[2 * i + 1];
    }
    result[0] = tl;
    result[1] = __uint_as_float(tw);
    result[2] = __uint_as_float(tb);
    *ticket = 0;
  }
}

size_t head_tc_smem_bytes(int nt, int C) {
  return ((size_t)(2 * nt * 2 * 32) * 2 + (size_t)(nt * 2 * 2 * 32) * 2 + (size_t)kWarps * head_tc_warp_floats(C)) * sizeof(float);
}

// out[i] = sum over the CTAs' partial blocks, ascending: 32 outputs x 8 groups per CTA (the fixed tree of slab_reduce_kernel)
__global__ void __launch_bounds__(256) head_reduce_kernel(const float *__restrict__ part, float *__restrict__ out, int elems, int blocks) {
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (i < elems)
    for (int z = g; z < blocks; z += 8) s += __ldg(part + (size_t)z * elems + i);
  red[g][lane] = s;
  __syncthreads();
  if (g == 0 && i < elems) {
    float t = red[0][lane];
#pragma unroll
    for (int k = 1; k < 8; k++) t += red[k][lane];
    out[i] = t;
  }
}

// CTAs per SM: the kernel is bound by shared memory -> register bandwidth (every broadcast LDS.128 of W delivers 512 bytes
// for four FMAs per lane), not by latency: a third CTA per SM (85 registers, a few spills) measured 83 us against 73 with two
int head_ctas_per_sm() {
  static const int v = [] {
    const char *e = getenv("GCNB_HEAD_CTAS");  // tuning probe
    return e && atoi(e) == 3 ? 3 : 2;
  }();
  return v;
}
int head_blocks(int64_t n) {
  const int sm = std::max(1, device_info().sm_count);
  const int64_t want = ((n + 31) / 32 + kWarps - 1) / kWarps;
  return (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(want, (int64_t)sm * head_ctas_per_sm()), kMaxBlocks));
}

size_t head_smem_bytes(int in, int C) {
  const int CW = (C + 3) & ~3;
  return ((size_t)2 * in * CW + (size_t)kWarps * head_warp_floats(in, C)) * sizeof(float);
}

template <int IN, int CTAS>
int launch_head(const float *Y, const float *W, const int32_t *truth, int64_t n, int C, uint32_t num_samples, int training,
                float *logits, float *grad, float *dY, float *dw_part, float *result, float *part_loss, uint32_t *part_cnt,
                unsigned int *ticket, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    GCNB_CHECK(cudaFuncSetAttribute(head_kernel<IN, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)head_smem_bytes(IN, kMaxC)));
    attr_set = true;
  }
  const int aligned16 = (((uintptr_t)logits | (uintptr_t)grad) % 16) == 0;
  const uint32_t magic = (1u << 17) / (uint32_t)C + 1u;  // i / C == (i * magic) >> 17 for i < 2048, C <= 64
  head_kernel<IN, CTAS><<<head_blocks(n), kT, head_smem_bytes(IN, C), st>>>(Y, W, truth, n, C, num_samples, training, aligned16, magic,
                                                                      logits, grad, dY, dw_part, result, part_loss, part_cnt, ticket);
  GCNB_LAUNCH_CHECK();
  return 0;
}

template <int NT>
int launch_head_tc(const float *Y, const float *W, const int32_t *truth, int64_t n, int C, uint32_t num_samples, int training,
                   float *logits, float *grad, float *dY, float *dw_part, float *result, float *part_loss, uint32_t *part_cnt,
                   unsigned int *ticket, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    GCNB_CHECK(cudaFuncSetAttribute(head_tc_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)head_tc_smem_bytes(NT, 8 * NT)));
    attr_set = true;
  }
  const int aligned16 = (((uintptr_t)logits | (uintptr_t)grad) % 16) == 0;
  const uint32_t magic = (1u << 17) / (uint32_t)C + 1u;
  head_tc_kernel<NT><<<head_blocks(n), kT, head_tc_smem_bytes(NT, C), st>>>(Y, W, truth, n, C, num_samples, training, aligned16, magic,
                                                                            logits, grad, dY, dw_part, result, part_loss, part_cnt, ticket);
  GCNB_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" {

int gcnb_head_supported(int in_dim, int num_classes) {
  return (in_dim == 8 || in_dim == 16 || in_dim == 32) && num_classes >= 1 && num_classes <= kMaxC;
}

int64_t gcnb_head_workspace(int64_t n, int in_dim, int num_classes) {
  (void)n;
  return 16 + (int64_t)kMaxBlocks * 12 + (int64_t)kMaxBlocks * in_dim * num_classes * (int64_t)sizeof(float);
}

int gcnb_head_f32(const float *d_y, const float *d_w, const int32_t *d_truth, int64_t n, int in_dim, int num_classes,
                  uint32_t num_samples, int training, float *d_logits, float *d_grad, float *d_dy, float *d_result, void *d_ws,
                  int64_t ws_bytes, gcnb_stream_t s) {
  if (!d_y || !d_w || !d_truth || !d_logits || !d_result || !d_ws || n < 0 || (training && !d_dy)) return GCNB_E_BADARG;
  if (!gcnb_head_supported(in_dim, num_classes)) return GCNB_E_UNSUPPORTED;
  if (ws_bytes < gcnb_head_workspace(n, in_dim, num_classes)) return GCNB_E_BADARG;
  if ((((uintptr_t)d_y | (uintptr_t)d_dy) % 16) != 0) return GCNB_E_BADARG;  // rows are read / written as float4
  unsigned int *ticket = (unsigned int *)d_ws;
  float *part_loss = (float *)d_ws + 4;
  uint32_t *part_cnt = (uint32_t *)d_ws + 4 + kMaxBlocks;
  float *dw_part = (float *)d_ws + 4 + 3 * kMaxBlocks;
  cudaStream_t st = as_stream(s);
#define HEAD(IN_)                                                                                                              \
  return head_ctas_per_sm() == 3                                                                                               \
             ? launch_head<IN_, 3>(d_y, d_w, d_truth, n, num_classes, num_samples, training, d_logits, d_grad, d_dy, dw_part,  \
                                   d_result, part_loss, part_cnt, ticket, st)                                                  \
             : launch_head<IN_, 2>(d_y, d_w, d_truth, n, num_classes, num_samples, training, d_logits, d_grad, d_dy, dw_part,  \
                                   d_result, part_loss, part_cnt, ticket, st)
  if (in_dim == 8) HEAD(8);
  if (in_dim == 16) HEAD(16);
  HEAD(32);
#undef HEAD
}

int gcnb_head_tc_supported(int in_dim, int num_classes) { return in_dim == 16 && num_classes >= 1 && num_classes <= 48; }

// the tensor-core variant (head_tc_kernel): same arguments, workspace and partial-sum layout as gcnb_head_f32
int gcnb_head_tc_f32(const float *d_y, const float *d_w, const int32_t *d_truth, int64_t n, int in_dim, int num_classes,
                     uint32_t num_samples, int training, float *d_logits, float *d_grad, float *d_dy, float *d_result, void *d_ws,
                     int64_t ws_bytes, gcnb_stream_t s) {
  if (!d_y || !d_w || !d_truth || !d_logits || !d_result || !d_ws || n < 0 || (training && !d_dy)) return GCNB_E_BADARG;
  if (!gcnb_head_tc_supported(in_dim, num_classes)) return GCNB_E_UNSUPPORTED;
  if (ws_bytes < gcnb_head_workspace(n, in_dim, num_classes)) return GCNB_E_BADARG;
  if ((((uintptr_t)d_y | (uintptr_t)d_dy) % 16) != 0) return GCNB_E_BADARG;
  unsigned int *ticket = (unsigned int *)d_ws;
  float *part_loss = (float *)d_ws + 4;
  uint32_t *part_cnt = (uint32_t *)d_ws + 4 + kMaxBlocks;
  float *dw_part = (float *)d_ws + 4 + 3 * kMaxBlocks;
  cudaStream_t st = as_stream(s);
#define HEAD_TC(NT_) \
  return launch_head_tc<NT_>(d_y, d_w, d_truth, n, num_classes, num_samples, training, d_logits, d_grad, d_dy, dw_part, d_result, \
                             part_loss, part_cnt, ticket, st)
  if (num_classes <= 16) HEAD_TC(2);
  if (num_classes <= 32) HEAD_TC(4);
  HEAD_TC(6);
#undef HEAD_TC
}

int gcnb_head_reduce_dw_f32(const void *d_ws, float *d_dw, int64_t n, int in_dim, int num_classes, gcnb_stream_t s) {
  if (!d_ws || !d_dw || n < 0 || !gcnb_head_supported(in_dim, num_classes)) return GCNB_E_BADARG;
  const float *dw_part = (const float *)d_ws + 4 + 3 * kMaxBlocks;
  const int elems = in_dim * num_classes;
  head_reduce_kernel<<<(elems + 31) / 32, 256, 0, as_stream(s)>>>(dw_part, d_dw, elems, head_blocks(n));
  GCNB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
