/*
 * gcn_oracle.c -- CPU restatement of the parallel-GCN hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the B200 engine.  It is never linked into, imported by
 * or executed from the product path (parallel-gcn_b200/); only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Every function restates (in plain C, scalar fp32, no FMA contraction: build with
 * -ffp-contract=off) one piece of the reference, cited as file:line relative to the reference
 * repository root.  "ref-CPU" = hpdga-spring23/ (sequential C++), "ref-GPU" = src/ + include/
 * (CUDA).  Where the two differ the function takes a flag or comes in two flavours.
 *
 * Pinning: tests/test_oracle_pin.py checks these functions bit-for-bit against the reference's
 * own CPU implementation compiled in place (oracle/_ref/libref_cpu.so, built by oracle/Makefile
 * from /root/reference/hpdga-spring23/src) and against the committed fixtures in tests/golden/
 * that were generated from it (tests/golden/make_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * RNG 1: ref-CPU xorshift128+ (hpdga-spring23/src/rand.cpp:6-28, include/rand.h:6,16)
 * ---------------------------------------------------------------------------------------- */
static uint64_t g_xs[2];

/* hpdga-spring23/src/rand.cpp:6-14: seed from two non-zero libc rand() draws, no srand(). */
ORC_API void orc_xorshift_seed_from_libc(void) {
  int a = 0, b = 0;
  while (a == 0 || b == 0) {
    a = rand();
    b = rand();
  }
  g_xs[0] = (uint64_t)a;
  g_xs[1] = (uint64_t)b;
}
ORC_API void orc_xorshift_set(uint64_t s0, uint64_t s1) { g_xs[0] = s0; g_xs[1] = s1; }
ORC_API void orc_xorshift_get(uint64_t *out) { out[0] = g_xs[0]; out[1] = g_xs[1]; }
ORC_API void orc_libc_srand(unsigned seed) { srand(seed); }

/* hpdga-spring23/src/rand.cpp:17-28: returns the low 31 bits of (t + s). */
static inline uint32_t xs_next(void) {
  uint64_t t = g_xs[0];
  const uint64_t s = g_xs[1];
  g_xs[0] = s;
  t ^= t << 23;
  t ^= t >> 17;
  t ^= s ^ (s >> 26);
  g_xs[1] = t;
  return (uint32_t)((t + s) & 0x7fffffffu);
}
ORC_API uint32_t orc_xorshift_next(void) { return xs_next(); }

/* ------------------------------------------------------------------------------------------
 * RNG 2: ref-GPU Philox4x32-10 as driven by cuRAND (src/variable.cu:5-11 curand_init(seed, i, 0),
 * src/variable.cu:51 / src/module.cu:25 curand_uniform4).  Public algorithm (Salmon et al. 2011);
 * cuRAND's use of it: key = (seed_lo, seed_hi), counter = (draw t, 0, subsequence i, 0), and
 * uniform = x * 2^-32 + 2^-33 evaluated as one fused multiply-add on the device.
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
  uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
  uint32_t k0 = key_in[0], k1 = key_in[1];
  for (int r = 0; r < 10; r++) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* t-th curand_uniform4 of state i (see header comment).  fused=1 mirrors device FMA contraction. */
ORC_API void orc_curand_uniform4(uint32_t seed, uint32_t state_i, uint32_t draw_t, int fused, float out[4]) {
  const uint32_t ctr[4] = {draw_t, 0u, state_i, 0u};
  const uint32_t key[2] = {seed, 0u};
  uint32_t x[4];
  orc_philox4x32_10(ctr, key, x);
  const float a = 2.3283064e-10f, b = 2.3283064e-10f / 2.0f;
  for (int k = 0; k < 4; k++) {
    if (fused) out[k] = fmaf((float)x[k], a, b);
    else { volatile float m = (float)x[k] * a; out[k] = m + b; }
  }
}

/* ------------------------------------------------------------------------------------------
 * Weight init
 * ---------------------------------------------------------------------------------------- */
/* ref-CPU Glorot: hpdga-spring23/src/variable.cpp:15-19 (the 0.5 literal is double). */
ORC_API void orc_glorot_xorshift(int64_t size, int in_size, int out_size, float *w) {
  const float range = sqrtf(6.0f / (in_size + out_size));
  for (int64_t i = 0; i < size; i++)
    w[i] = (float)(((double)((float)xs_next() / 0x7fffffff) - 0.5) * range * 2);
}

/* ref-GPU Glorot: src/variable.cu:44-61 (kernel) + :63-83 (host: range=sqrtf(6.0f/(rows+cols)) widened
 * to double, scale = range*2; element = (u - 0.5) * scale in double, stored fp32).  draw_t = how many
 * RNG ops already consumed state i; one value per call site because every state below
 * ceil(size/4) has seen the same number of earlier ops only when sizes are nested -- the caller passes
 * the per-state draw index through draw_of_state (NULL => all zero). */
ORC_API void orc_glorot_philox(int64_t size, unsigned rows, unsigned cols, uint32_t seed,
                               const uint32_t *draw_of_state, int fused, float *w) {
  const double range = sqrtf(6.0f / (rows + cols));
  const double scale = range * 2;
  const int64_t groups = (size + 3) / 4;
  for (int64_t g = 0; g < groups; g++) {
    float u[4];
    orc_curand_uniform4(seed, (uint32_t)g, draw_of_state ? draw_of_state[g] : 0u, fused, u);
    for (int k = 0; k < 4; k++) {
      const int64_t j = g * 4 + k;
      if (j < size) w[j] = (float)((u[k] - 0.5) * scale);
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * Parser: .graph / .svmlight / .split  ->  CSR  (src/parser.cpp:15-48, 59-112, 114-132;
 * ref-CPU twin hpdga-spring23/src/parser.cpp:18-116).  Behaviour restated:
 *   - a line is consumed with getline(); the loop stops when getline hits EOF, so a last line
 *     without trailing newline is DROPPED (eof() is set while reading it);
 *   - graph row i = [i (implicit self), neighbours in file order], duplicates kept;
 *   - svmlight: "label k:v k:v ...", failed label extraction => label -1 and no features;
 *     input_dim = max k + 1, output_dim = max label + 1;
 *   - split: one integer per line (std::stoi), counts of 1/2/3 -> train/val/test dims.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int64_t num_nodes, graph_nnz, feat_rows, feat_nnz, input_dim, output_dim, n_split;
  int64_t train_dim, val_dim, test_dim;
  uint32_t *g_indptr, *g_indices, *f_indptr, *f_indices, *split;
  int32_t *label;
  float *f_value;
} orc_dataset;

typedef struct { char *p; size_t len, pos; } textbuf;

static int slurp(const char *path, textbuf *tb) {
  FILE *f = fopen(path, "rb");
  if (!f) return 0;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  tb->p = (char *)malloc((size_t)n + 1);
  tb->len = fread(tb->p, 1, (size_t)n, f);
  tb->p[tb->len] = 0;
  tb->pos = 0;
  fclose(f);
  return 1;
}
/* Emulates `getline(file, line); if (file.eof()) break;`: a line counts only if terminated by '\n'. */
static int next_line(textbuf *tb, char **beg, char **end) {
  if (tb->pos >= tb->len) return 0;
  char *nl = (char *)memchr(tb->p + tb->pos, '\n', tb->len - tb->pos);
  if (!nl) return 0;
  *beg = tb->p + tb->pos;
  *end = nl;
  tb->pos = (size_t)(nl - tb->p) + 1;
  return 1;
}
#define PUSH(arr, cnt, cap, val, T)                                    \
  do {                                                                 \
    if ((cnt) == (cap)) { (cap) = (cap) ? (cap)*2 : 1024; (arr) = (T *)realloc((arr), (size_t)(cap) * sizeof(T)); } \
    (arr)[(cnt)++] = (val);                                            \
  } while (0)

static int is_ws(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f' || c == '\n'; }

/* `ss >> int`: skip whitespace, parse a decimal integer; fail if none. */
static int scan_int(char **cur, char *end, long *out) {
  char *p = *cur;
  while (p < end && is_ws(*p)) p++;
  if (p >= end) return 0;
  char *q;
  char save = *end;
  *end = 0;
  long v = strtol(p, &q, 10);
  *end = save;
  if (q == p) return 0;
  *cur = q;
  *out = v;
  return 1;
}

ORC_API orc_dataset *orc_dataset_parse(const char *prefix) {
  char path[4096];
  textbuf g = {0}, s = {0}, v = {0};
  snprintf(path, sizeof path, "%s.graph", prefix);
  if (!slurp(path, &g)) return NULL;
  snprintf(path, sizeof path, "%s.split", prefix);
  if (!slurp(path, &s)) { free(g.p); return NULL; }
  snprintf(path, sizeof path, "%s.svmlight", prefix);
  if (!slurp(path, &v)) { free(g.p); free(s.p); return NULL; }

  orc_dataset *d = (orc_dataset *)calloc(1, sizeof *d);
  char *b, *e;
  /* graph: src/parser.cpp:15-48 */
  {
    int64_t np = 0, cp = 0, ni = 0, ci = 0;
    PUSH(d->g_indptr, np, cp, 0u, uint32_t);
    int64_t node = 0;
    while (next_line(&g, &b, &e)) {
      PUSH(d->g_indices, ni, ci, (uint32_t)node, uint32_t);
      PUSH(d->g_indptr, np, cp, d->g_indptr[np - 1] + 1, uint32_t);
      node++;
      char *cur = b;
      long nb;
      while (scan_int(&cur, e, &nb)) {
        PUSH(d->g_indices, ni, ci, (uint32_t)nb, uint32_t);
        d->g_indptr[np - 1] += 1;
      }
    }
    d->num_nodes = node;
    d->graph_nnz = ni;
  }
  /* svmlight: src/parser.cpp:59-112 */
  {
    int64_t np = 0, cp = 0, ni = 0, ci = 0, nv = 0, cv = 0, nl = 0, cl = 0;
    long max_idx = 0, max_label = 0;
    PUSH(d->f_indptr, np, cp, 0u, uint32_t);
    while (next_line(&v, &b, &e)) {
      PUSH(d->f_indptr, np, cp, d->f_indptr[np - 1], uint32_t);
      char *cur = b;
      long label = -1;
      int ok = scan_int(&cur, e, &label);
      PUSH(d->label, nl, cl, (int32_t)(ok ? label : -1), int32_t);
      if (!ok) continue;
      if (label > max_label) max_label = label;
      for (;;) {
        /* `ss >> kv` : next whitespace-delimited token, then `kv_ss >> k >> col >> v` */
        while (cur < e && is_ws(*cur)) cur++;
        if (cur >= e) break;
        char *tok_end = cur;
        while (tok_end < e && !is_ws(*tok_end)) tok_end++;
        char save = *tok_end;
        *tok_end = 0;
        char *q;
        long k = strtol(cur, &q, 10);
        float val = 0.f;
        if (q != cur && *q) { q++; /* the separator char */ val = strtof(q, NULL); }
        *tok_end = save;
        cur = tok_end;
        PUSH(d->f_value, nv, cv, val, float);
        PUSH(d->f_indices, ni, ci, (uint32_t)k, uint32_t);
        d->f_indptr[np - 1] += 1;
        if (k > max_idx) max_idx = k;
      }
    }
    d->feat_rows = np - 1;
    d->feat_nnz = ni;
    d->input_dim = max_idx + 1;
    d->output_dim = max_label + 1;
  }
  /* split: src/parser.cpp:114-132 */
  {
    int64_t ns = 0, cs = 0;
    while (next_line(&s, &b, &e)) {
      char *cur = b;
      long val = 0;
      scan_int(&cur, e, &val);
      if (val == 1) d->train_dim++;
      else if (val == 2) d->val_dim++;
      else if (val == 3) d->test_dim++;
      PUSH(d->split, ns, cs, (uint32_t)val, uint32_t);
    }
    d->n_split = ns;
  }
  free(g.p); free(s.p); free(v.p);
  return d;
}
ORC_API void orc_dataset_dims(const orc_dataset *d, int64_t out[10]) {
  out[0] = d->num_nodes; out[1] = d->graph_nnz; out[2] = d->feat_rows; out[3] = d->feat_nnz;
  out[4] = d->input_dim; out[5] = d->output_dim; out[6] = d->n_split;
  out[7] = d->train_dim; out[8] = d->val_dim; out[9] = d->test_dim;
}
/* which: 0 g_indptr 1 g_indices 2 f_indptr 3 f_indices 4 f_value 5 label 6 split */
ORC_API void orc_dataset_copy(const orc_dataset *d, int which, void *dst) {
  switch (which) {
    case 0: memcpy(dst, d->g_indptr, (size_t)(d->num_nodes + 1) * 4); break;
    case 1: memcpy(dst, d->g_indices, (size_t)d->graph_nnz * 4); break;
    case 2: memcpy(dst, d->f_indptr, (size_t)(d->feat_rows + 1) * 4); break;
    case 3: memcpy(dst, d->f_indices, (size_t)d->feat_nnz * 4); break;
    case 4: memcpy(dst, d->f_value, (size_t)d->feat_nnz * 4); break;
    case 5: memcpy(dst, d->label, (size_t)d->feat_rows * 4); break;
    case 6: memcpy(dst, d->split, (size_t)d->n_split * 4); break;
  }
}
ORC_API void orc_dataset_free(orc_dataset *d) {
  if (!d) return;
  free(d->g_indptr); free(d->g_indices); free(d->f_indptr); free(d->f_indices);
  free(d->split); free(d->label); free(d->f_value); free(d);
}

/* src/parser.cpp:164-181: graph_value[e] = 1. / sqrtf(unsigned(deg(src) * deg(dst))):
 * 32-bit unsigned product -> float -> sqrtf -> double divide -> stored fp32. */
ORC_API void orc_graph_values(int64_t n, const uint32_t *indptr, const uint32_t *indices, float *out) {
  for (int64_t src = 0; src < n; src++)
    for (uint32_t e = indptr[src]; e < indptr[src + 1]; e++) {
      const uint32_t dst = indices[e];
      const uint32_t prod = (indptr[src + 1] - indptr[src]) * (indptr[dst + 1] - indptr[dst]);
      out[e] = (float)(1. / sqrtf((float)prod));
    }
}

/* ------------------------------------------------------------------------------------------
 * Operators.  Summation orders are the reference's.
 * ---------------------------------------------------------------------------------------- */

/* GraphSum forward / backward (same loop on grads: symmetric-graph assumption).
 * values != NULL : ref-GPU graphsum_kernel, src/module.cu:172-186 (sum over ascending jj of a[jj]*b[col]).
 * values == NULL : ref-CPU, hpdga-spring23/src/module.cpp:82-96: coef recomputed per edge as
 *                  1.0 / sqrtf(int product) (double) narrowed to float, accumulation directly in out[]. */
ORC_API void orc_graphsum(int64_t n, int dim, const uint32_t *indptr, const uint32_t *indices,
                          const float *values, const float *in, float *out) {
  for (int64_t r = 0; r < n; r++) {
    float *o = out + r * dim;
    for (int j = 0; j < dim; j++) o[j] = 0.f;
    for (uint32_t e = indptr[r]; e < indptr[r + 1]; e++) {
      const uint32_t c = indices[e];
      float coef;
      if (values) coef = values[e];
      else {
        const int prod = (int)(indptr[r + 1] - indptr[r]) * (int)(indptr[c + 1] - indptr[c]);
        coef = (float)(1.0 / sqrtf((float)prod));
      }
      const float *x = in + (int64_t)c * dim;
      for (int j = 0; j < dim; j++) o[j] += coef * x[j];
    }
  }
}

/* SparseMatmul forward: C[m x p] = A_csr * B[n x p]; src/module.cu:108-122, ref-CPU module.cpp:46-59. */
ORC_API void orc_spmm(int64_t m, int p, const uint32_t *indptr, const uint32_t *indices,
                      const float *a_val, const float *b, float *c) {
  for (int64_t i = 0; i < m; i++) {
    float *o = c + i * p;
    for (int k = 0; k < p; k++) o[k] = 0.f;
    for (uint32_t jj = indptr[i]; jj < indptr[i + 1]; jj++) {
      const float *brow = b + (int64_t)indices[jj] * p;
      const float av = a_val[jj];
      for (int k = 0; k < p; k++) o[k] += av * brow[k];
    }
  }
}

/* SparseMatmul backward: B.grad[n x p] = A_csr^T * C.grad; ref-CPU order (ascending i, jj):
 * hpdga-spring23/src/module.cpp:61-72.  (ref-GPU src/module.cu:136-152 is unordered atomics.) */
ORC_API void orc_spmm_bwd(int64_t m, int64_t n, int p, const uint32_t *indptr, const uint32_t *indices,
                          const float *a_val, const float *c_grad, float *b_grad) {
  for (int64_t i = 0; i < n * p; i++) b_grad[i] = 0.f;
  for (int64_t i = 0; i < m; i++)
    for (uint32_t jj = indptr[i]; jj < indptr[i + 1]; jj++) {
      float *g = b_grad + (int64_t)indices[jj] * p;
      const float *cg = c_grad + i * p;
      for (int k = 0; k < p; k++) g[k] += cg[k] * a_val[jj];
    }
}

/* Matmul forward: C[m x p] = A[m x n] * B[n x p], ascending j; module.cpp:14-22, src/module.cu:274-317. */
ORC_API void orc_matmul(int64_t m, int n, int p, const float *a, const float *b, float *c) {
  for (int64_t i = 0; i < m; i++) {
    float *o = c + i * p;
    for (int k = 0; k < p; k++) o[k] = 0.f;
    for (int j = 0; j < n; j++) {
      const float av = a[i * n + j];
      for (int k = 0; k < p; k++) o[k] += av * b[(int64_t)j * p + k];
    }
  }
}

/* Matmul backward: A.grad = C.grad * B^T ; B.grad = A^T * C.grad (ascending i).
 * hpdga-spring23/src/module.cpp:24-38; ref-GPU src/module.cu:332-391. */
ORC_API void orc_matmul_bwd(int64_t m, int n, int p, const float *a, const float *b, const float *c_grad,
                            float *a_grad, float *b_grad) {
  for (int64_t i = 0; i < (int64_t)n * p; i++) b_grad[i] = 0.f;
  for (int64_t i = 0; i < m; i++)
    for (int j = 0; j < n; j++) {
      float tmp = 0.f;
      for (int k = 0; k < p; k++) {
        tmp += c_grad[i * p + k] * b[(int64_t)j * p + k];
        b_grad[(int64_t)j * p + k] += c_grad[i * p + k] * a[i * n + j];
      }
      a_grad[i * n + j] = tmp;
    }
}

/* ReLU: src/module.cu:222-256, module.cpp:164-188.  mask written only when training. */
ORC_API void orc_relu_fwd(int64_t n, float *x, uint8_t *mask, int training) {
  for (int64_t i = 0; i < n; i++) {
    const int keep = x[i] > 0;
    if (training) mask[i] = (uint8_t)keep;
    if (!keep) x[i] = 0.f;
  }
}
ORC_API void orc_relu_bwd(int64_t n, float *g, const uint8_t *mask) {
  for (int64_t i = 0; i < n; i++)
    if (!mask[i]) g[i] = 0.f;
}

/* Dropout masks.  ref-CPU: module.cpp:207-217, keep = (int)RAND() >= int(p * 0x7fffffff). */
ORC_API void orc_dropout_mask_xorshift(int64_t n, float p, uint8_t *mask) {
  const int threshold = (int)(p * 0x7fffffff);
  for (int64_t i = 0; i < n; i++) mask[i] = (uint8_t)((int)xs_next() >= threshold);
}
/* ref-GPU: src/module.cu:16-63, keep = u >= p, element 4g+k uses lane k of state g's draw. */
ORC_API void orc_dropout_mask_philox(int64_t n, float p, uint32_t seed, const uint32_t *draw_of_state,
                                     int fused, uint8_t *mask) {
  const int64_t groups = (n + 3) / 4;
  for (int64_t g = 0; g < groups; g++) {
    float u[4];
    orc_curand_uniform4(seed, (uint32_t)g, draw_of_state ? draw_of_state[g] : 0u, fused, u);
    for (int k = 0; k < 4; k++)
      if (g * 4 + k < n) mask[g * 4 + k] = (uint8_t)(u[k] >= p);
  }
}
/* x *= keep ? scale : 0 with scale given by the caller (ref-CPU: float 1/(1-p), module.cpp:212;
 * ref-GPU: double 1.0/(1.0-p) narrowed, src/module.cu:69). */
ORC_API void orc_dropout_apply(int64_t n, float *x, const uint8_t *mask, float scale) {
  for (int64_t i = 0; i < n; i++) x[i] *= mask[i] ? scale : 0.f;
}
ORC_API float orc_dropout_scale(float p, int gpu_flavour) {
  if (gpu_flavour) return (float)(1.0 / (1.0 - p));
  return 1 / (1 - p);
}

/* CrossEntropyLoss forward (+grad).  Mutates logits in place (-= row max) for labelled rows.
 * num_samples <= 0 : ref-CPU (module.cpp:119-153): count labelled rows; grad = prob, grad[t] -= 1.0
 *                    (double), then every grad /= count; returns total/count.
 * num_samples  > 0 : ref-GPU (src/module.cu:484-524): grad = prob/num_samples; grad[t] -= 1.0/num_samples
 *                    (double); rows with truth<0 keep grad 0 (zero_grad first, :528-529); returns the
 *                    UN-normalised sum (GCN::finalize divides, src/gcn.cu:447), summed in ascending node order
 *                    here (the reference's warp/atomic order is not defined). */
ORC_API float orc_cross_entropy(int64_t n, int C, float *logits, const int32_t *truth, float *grad,
                                int64_t num_samples, int training, int64_t *count_out) {
  float total = 0.f;
  int64_t count = 0;
  if (training) memset(grad, 0, (size_t)(n * C) * sizeof(float));
  for (int64_t i = 0; i < n; i++) {
    if (truth[i] < 0) continue;
    count++;
    float *lg = logits + i * C;
    float mx = num_samples > 0 ? lg[0] : -1e30f, se = 0.f;
    for (int j = 0; j < C; j++) mx = fmaxf(mx, lg[j]); /* ref-CPU uses fmax (double) on floats: same value */
    for (int j = 0; j < C; j++) { lg[j] -= mx; se += expf(lg[j]); }
    total += logf(se) - lg[truth[i]];
    if (training) {
      for (int j = 0; j < C; j++) {
        const float prob = expf(lg[j]) / se;
        grad[i * C + j] = num_samples > 0 ? prob / (float)num_samples : prob;
      }
      if (num_samples > 0) grad[i * C + truth[i]] = (float)(grad[i * C + truth[i]] - 1.0 / (double)num_samples);
      else grad[i * C + truth[i]] = (float)(grad[i * C + truth[i]] - 1.0);
    }
  }
  if (count_out) *count_out = count;
  if (num_samples > 0) return total;
  if (training)
    for (int64_t i = 0; i < n * C; i++) grad[i] /= count;
  return total / count;
}

/* Accuracy.  ref-CPU gcn.cpp:144-161 (any logit > truth logit => wrong); ref-GPU src/gcn.cu:264-278
 * (shifted truth logit < 0 => wrong).  Identical on CE-shifted logits; both return the wrong count. */
ORC_API int64_t orc_wrong_count(int64_t n, int C, const float *logits, const int32_t *truth, int64_t *total_out) {
  int64_t wrong = 0, total = 0;
  for (int64_t i = 0; i < n; i++) {
    if (truth[i] < 0) continue;
    total++;
    const float t = logits[i * C + truth[i]];
    for (int j = 0; j < C; j++)
      if (logits[i * C + j] > t) { wrong++; break; }
  }
  if (total_out) *total_out = total;
  return wrong;
}

/* sum of squares, ascending order: gcn.cpp:167-174 / src/gcn.cu:230-243. */
ORC_API float orc_sumsq(int64_t n, const float *w) {
  float s = 0.f;
  for (int64_t i = 0; i < n; i++) { const float x = w[i]; s += x * x; }
  return s;
}

/* set_truth: gcn.cpp:139-143 / src/gcn.cu:204-210. */
ORC_API void orc_set_truth(int64_t n, const uint32_t *split, const int32_t *label, uint32_t cur, int32_t *truth) {
  for (int64_t i = 0; i < n; i++) truth[i] = split[i] == cur ? label[i] : -1;
}

/* Adam: optim.cpp:23-35 / src/optim.cu:42-62.  (1.0 - beta) products and sums in double. */
ORC_API float orc_adam_step_size(float lr, float beta1, float beta2, int step) {
  return lr * sqrtf(1 - powf(beta2, step)) / (1 - powf(beta1, step));
}
ORC_API void orc_adam_step(int64_t n, float *w, const float *g, float *m, float *v, int decay,
                           float weight_decay, float beta1, float beta2, float eps, float step_size) {
  for (int64_t i = 0; i < n; i++) {
    float grad = g[i];
    if (decay) grad += weight_decay * w[i];
    m[i] = (float)(beta1 * m[i] + (1.0 - beta1) * grad);
    v[i] = (float)(beta2 * v[i] + (1.0 - beta2) * grad * grad);
    w[i] -= step_size * m[i] / (sqrtf(v[i]) + eps);
  }
}

/* FNV-1a 64 over raw bytes (SURVEY Appendix B known answers). */
ORC_API uint64_t orc_fnv1a64(const void *p, int64_t nbytes) {
  const uint8_t *b = (const uint8_t *)p;
  uint64_t h = 0xcbf29ce484222325ull;
  for (int64_t i = 0; i < nbytes; i++) { h ^= b[i]; h *= 0x100000001b3ull; }
  return h;
}
