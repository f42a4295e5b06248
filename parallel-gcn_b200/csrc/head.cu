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

int gcnb_head_reduce_dw_f32(const void *d_ws, float *d_dw, int64_t n, int in_dim, int num_classes, gcnb_stream_t s) {
  if (!d_ws || !d_dw || n < 0 || !gcnb_head_supported(in_dim, num_classes)) return GCNB_E_BADARG;
  const float *dw_part = (const float *)d_ws + 4 + 3 * kMaxBlocks;
  const int elems = in_dim * num_classes;
  head_reduce_kernel<<<(elems + 31) / 32, 256, 0, as_stream(s)>>>(dw_part, d_dw, elems, head_blocks(n));
  GCNB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
