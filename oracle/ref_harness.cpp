// ref_harness.cpp -- C-ABI shim around the UNMODIFIED reference CPU implementation
// (hpdga-spring23/src/*.cpp, compiled where it lies under /root/reference by oracle/Makefile into
// oracle/_ref/libref_cpu.so).  TEST INFRASTRUCTURE ONLY: used to pin oracle/gcn_oracle.c, to generate
// tests/golden/ and as bench.py's CPU baseline ("kind": "reference").  Nothing here is product code and
// no reference source is copied: this file only *includes the reference headers* and calls their classes.
#include <assert.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <unistd.h>
#include <utility>
#include <vector>

// The reference keeps train_epoch()/eval()/variables private; the harness needs them to drive single
// epochs and to read tensors.  Access specifiers do not change layout with GCC.  (`class` -> `struct` because
// GCN's leading members are private by default; every std header the reference pulls in is included above.)
#define private public
#define class struct
#include "include/gcn.h"
#include "include/module.h"
#include "include/optim.h"
#include "include/parser.h"
#include "include/rand.h"
#include "include/timer.h"
#undef class
#undef private

extern "C" {

struct RefDataset {
  GCNParams params;
  GCNData data;
};

void ref_srand(unsigned seed) { srand(seed); }
void ref_rand_state_get(uint64_t *out) { out[0] = rand_state[0]; out[1] = rand_state[1]; }
void ref_rand_state_set(uint64_t a, uint64_t b) { rand_state[0] = a; rand_state[1] = b; }
uint32_t ref_rand_next() { return RAND(); }

// Parser opens "data/<name>.*" relative to the CWD (hpdga-spring23/src/parser.cpp:6-9).
RefDataset *ref_dataset_parse(const char *root_dir, const char *name) {
  char cwd[4096];
  if (!getcwd(cwd, sizeof cwd)) return nullptr;
  if (chdir(root_dir) != 0) return nullptr;
  RefDataset *d = new RefDataset();
  d->params = GCNParams::get_default();
  bool ok;
  {
    Parser parser(&d->params, &d->data, name);
    ok = parser.parse();
  }
  if (chdir(cwd) != 0) ok = false;
  if (!ok) { delete d; return nullptr; }
  return d;
}

// In-memory dataset (synthetic graphs): public fields of GCNData, hpdga-spring23/include/gcn.h:17-23.
RefDataset *ref_dataset_from_arrays(int64_t n, int64_t gnnz, const int32_t *g_indptr, const int32_t *g_indices,
                                    int64_t fnnz, const int32_t *f_indptr, const int32_t *f_indices,
                                    const float *f_value, const int32_t *label, const int32_t *split,
                                    int input_dim, int output_dim) {
  RefDataset *d = new RefDataset();
  d->params = GCNParams::get_default();
  d->params.num_nodes = (int)n;
  d->params.input_dim = input_dim;
  d->params.output_dim = output_dim;
  d->data.graph.indptr.assign(g_indptr, g_indptr + n + 1);
  d->data.graph.indices.assign(g_indices, g_indices + gnnz);
  d->data.feature_index.indptr.assign(f_indptr, f_indptr + n + 1);
  d->data.feature_index.indices.assign(f_indices, f_indices + fnnz);
  d->data.feature_value.assign(f_value, f_value + fnnz);
  d->data.label.assign(label, label + n);
  d->data.split.assign(split, split + n);
  return d;
}

void ref_dataset_dims(const RefDataset *d, int64_t out[8]) {
  out[0] = d->params.num_nodes;
  out[1] = (int64_t)d->data.graph.indices.size();
  out[2] = (int64_t)d->data.feature_index.indptr.size() - 1;
  out[3] = (int64_t)d->data.feature_index.indices.size();
  out[4] = d->params.input_dim;
  out[5] = d->params.output_dim;
  out[6] = (int64_t)d->data.split.size();
  out[7] = (int64_t)d->data.label.size();
}
// which: 0 g_indptr 1 g_indices 2 f_indptr 3 f_indices 4 f_value 5 label 6 split
void ref_dataset_copy(const RefDataset *d, int which, void *dst) {
  auto cp = [&](const void *src, size_t n) { memcpy(dst, src, n); };
  switch (which) {
    case 0: cp(d->data.graph.indptr.data(), d->data.graph.indptr.size() * 4); break;
    case 1: cp(d->data.graph.indices.data(), d->data.graph.indices.size() * 4); break;
    case 2: cp(d->data.feature_index.indptr.data(), d->data.feature_index.indptr.size() * 4); break;
    case 3: cp(d->data.feature_index.indices.data(), d->data.feature_index.indices.size() * 4); break;
    case 4: cp(d->data.feature_value.data(), d->data.feature_value.size() * 4); break;
    case 5: cp(d->data.label.data(), d->data.label.size() * 4); break;
    case 6: cp(d->data.split.data(), d->data.split.size() * 4); break;
  }
}
void ref_dataset_free(RefDataset *d) { delete d; }

// ---- whole model ---------------------------------------------------------------------------------
GCN *ref_gcn_create(RefDataset *d, int hidden_dim, float dropout, float lr, float weight_decay, int epochs,
                    int early_stopping) {
  GCNParams p = d->params;
  p.hidden_dim = hidden_dim;
  p.dropout = dropout;
  p.learning_rate = lr;
  p.weight_decay = weight_decay;
  p.epochs = epochs;
  p.early_stopping = early_stopping;
  return new GCN(p, &d->data);  // ctor seeds xorshift from libc rand() and draws Glorot weights
}
void ref_gcn_free(GCN *g) { delete g; }
void ref_gcn_train_epoch(GCN *g, float out[2]) {
  auto r = g->train_epoch();
  out[0] = r.first;
  out[1] = r.second;
}
void ref_gcn_eval(GCN *g, int split, float out[2]) {
  auto r = g->eval(split);
  out[0] = r.first;
  out[1] = r.second;
}
int ref_gcn_num_variables(GCN *g) { return (int)g->variables.size(); }
int64_t ref_gcn_variable_size(GCN *g, int idx, int grad) {
  return (int64_t)(grad ? g->variables[idx].grad.size() : g->variables[idx].data.size());
}
void ref_gcn_variable_get(GCN *g, int idx, int grad, float *dst) {
  auto &v = grad ? g->variables[idx].grad : g->variables[idx].data;
  memcpy(dst, v.data(), v.size() * sizeof(float));
}
void ref_gcn_variable_set(GCN *g, int idx, const float *src) {
  auto &v = g->variables[idx].data;
  memcpy(v.data(), src, v.size() * sizeof(float));
}
// Full reference run() (built with -DEVAL: silent, fills gcn.time = TMR_TRAIN/(epochs+1) in ms,
// hpdga-spring23/src/gcn.cpp:270-273).  Returns wall seconds of run(); *ref_time_ms gets gcn.time.
double ref_gcn_run(GCN *g, float *ref_time_ms) {
  reset_timer();
  auto t0 = std::chrono::steady_clock::now();
  g->run();
  auto t1 = std::chrono::steady_clock::now();
  if (ref_time_ms) *ref_time_ms = g->time;
  return std::chrono::duration<double>(t1 - t0).count();
}
// per-module accumulated seconds (timer.h enum order), after ref_gcn_run / train_epoch calls
float ref_timer_total(int which) { return timer_total((timer_instance)which); }
void ref_timer_reset() { reset_timer(); }

// ---- single modules on caller arrays ---------------------------------------------------------------
static SparseIndex make_index(int64_t rows, const int32_t *indptr, const int32_t *indices) {
  SparseIndex s;
  s.indptr.assign(indptr, indptr + rows + 1);
  s.indices.assign(indices, indices + indptr[rows]);
  return s;
}
static void fill(std::vector<float> &v, const float *src) { if (src) memcpy(v.data(), src, v.size() * 4); }
static void dump(const std::vector<float> &v, float *dst) { if (dst) memcpy(dst, v.data(), v.size() * 4); }

void ref_graphsum(int64_t n, int dim, const int32_t *indptr, const int32_t *indices, const float *in,
                  float *out, const float *out_grad, float *in_grad) {
  SparseIndex g = make_index(n, indptr, indices);
  Variable vin((int)(n * dim)), vout((int)(n * dim));
  fill(vin.data, in);
  GraphSum m(&vin, &vout, &g, dim);
  m.forward(true);
  dump(vout.data, out);
  if (out_grad) {
    fill(vout.grad, out_grad);
    m.backward();
    dump(vin.grad, in_grad);
  }
}
void ref_sparse_matmul(int64_t m_, int n_, int p_, const int32_t *indptr, const int32_t *indices,
                       const float *a_val, const float *b, float *c, const float *c_grad, float *b_grad) {
  SparseIndex sp = make_index(m_, indptr, indices);
  Variable a(indptr[m_], false), bb(n_ * p_), cc((int)(m_ * p_));
  fill(a.data, a_val);
  fill(bb.data, b);
  SparseMatmul mod(&a, &bb, &cc, &sp, (int)m_, n_, p_);
  mod.forward(true);
  dump(cc.data, c);
  if (c_grad) {
    fill(cc.grad, c_grad);
    mod.backward();
    dump(bb.grad, b_grad);
  }
}
void ref_matmul(int64_t m_, int n_, int p_, const float *a, const float *b, float *c, const float *c_grad,
                float *a_grad, float *b_grad) {
  Variable va((int)(m_ * n_)), vb(n_ * p_), vc((int)(m_ * p_));
  fill(va.data, a);
  fill(vb.data, b);
  Matmul mod(&va, &vb, &vc, (int)m_, n_, p_);
  mod.forward(true);
  dump(vc.data, c);
  if (c_grad) {
    fill(vc.grad, c_grad);
    mod.backward();
    dump(va.grad, a_grad);
    dump(vb.grad, b_grad);
  }
}
float ref_cross_entropy(int64_t n, int C, float *logits_inout, const int32_t *truth, float *grad, int training) {
  Variable lg((int)(n * C));
  fill(lg.data, logits_inout);
  float loss = 0;
  std::vector<int> t(truth, truth + n);
  CrossEntropyLoss mod(&lg, t.data(), &loss, C);
  mod.forward(training != 0);
  dump(lg.data, logits_inout);
  if (training) dump(lg.grad, grad);
  return loss;
}
// Dropout forward with the reference's global xorshift stream; returns data and the mask via backward on ones.
void ref_dropout(int64_t n, float p, float *x_inout, float *ones_grad_out) {
  Variable v((int)n);
  fill(v.data, x_inout);
  Dropout mod(&v, p);
  mod.forward(true);
  dump(v.data, x_inout);
  if (ones_grad_out) {
    for (auto &g : v.grad) g = 1.f;
    mod.backward();
    dump(v.grad, ones_grad_out);
  }
}
void ref_relu(int64_t n, float *x_inout, float *grad_inout) {
  Variable v((int)n);
  fill(v.data, x_inout);
  ReLU mod(&v);
  mod.forward(true);
  dump(v.data, x_inout);
  if (grad_inout) {
    fill(v.grad, grad_inout);
    mod.backward();
    dump(v.grad, grad_inout);
  }
}
void ref_glorot(int64_t size, int in_size, int out_size, float *w) {
  Variable v((int)size, false);
  v.glorot(in_size, out_size);
  dump(v.data, w);
}
void ref_adam(int64_t n, int steps, float *w_inout, const float *grads /* steps x n */, int decay, float lr,
              float weight_decay) {
  Variable v((int)n);
  fill(v.data, w_inout);
  AdamParams ap = AdamParams::get_default();
  ap.lr = lr;
  ap.weight_decay = weight_decay;
  Adam opt({{&v, decay != 0}}, ap);
  for (int s = 0; s < steps; s++) {
    memcpy(v.grad.data(), grads + (int64_t)s * n, (size_t)n * 4);
    opt.step();
  }
  dump(v.data, w_inout);
}
}  // extern "C"
