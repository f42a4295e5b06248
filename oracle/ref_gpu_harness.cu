// ref_gpu_harness.cu -- stand-alone driver around the UNMODIFIED reference GPU implementation (src/*.cu of
// davide-gurrieri/parallel-GCN, compiled where it lies under /root/reference by oracle/Makefile with the flags of the
// reference's `performance-gpu` target, -arch=sm_100 instead of sm_75) -> oracle/_ref/ref_gpu_bench.
// TEST / MEASUREMENT INFRASTRUCTURE ONLY: it gives the same-box number "the reference's own CUDA code on this B200" for
// scripts/bench_ref_gpu.py.  A separate PROCESS on purpose: the reference's classes have the names of the product's host
// mirror (and `inline static` members are process-unique symbols), and a fault in the reference must not take a bench down.
// No reference source is copied: this file includes the reference headers and calls GCN::run() the way
// test/performance_gpu.cpp does (:24-76), with GCNData filled from raw arrays instead of Parser (its istringstream
// parser needs hours for a 115 M-entry text file).
//
// usage: ref_gpu_bench <dir> <epochs> <reps>      <dir> holds meta.txt and the raw little-endian arrays
//   meta.txt: num_nodes input_dim output_dim graph_nnz feat_nnz
//   g_indptr.u32 g_indices.u32 f_indptr.u32 f_indices.u32 f_value.f32 label.i32 split.u32
#include <unistd.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "include/gcn.cuh"
#include "include/optim.cuh"
#include "include/timer.h"

template <class T>
static bool read_raw(const std::string &path, std::vector<T> &v, size_t n) {
  v.resize(n);
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) return false;
  const size_t got = n ? fread(v.data(), sizeof(T), n, f) : 0;
  fclose(f);
  return got == n;
}

int main(int argc, char **argv) {
  if (argc < 4) {
    fprintf(stderr, "usage: %s <dir> <epochs> <reps>\n", argv[0]);
    return 2;
  }
  const std::string dir = argv[1];
  const natural epochs = (natural)atoi(argv[2]);
  const int reps = atoi(argv[3]);
  size_t n = 0, in_dim = 0, out_dim = 0, gnnz = 0, fnnz = 0;
  {
    std::ifstream meta(dir + "/meta.txt");
    if (!(meta >> n >> in_dim >> out_dim >> gnnz >> fnnz)) {
      fprintf(stderr, "bad meta.txt\n");
      return 2;
    }
  }
  GCNParams params;
  AdamParams adam_params;
  GCNData data;
  bool ok = read_raw(dir + "/g_indptr.u32", data.graph.indptr, n + 1) && read_raw(dir + "/g_indices.u32", data.graph.indices, gnnz) &&
            read_raw(dir + "/f_indptr.u32", data.feature_index.indptr, n + 1) &&
            read_raw(dir + "/f_indices.u32", data.feature_index.indices, fnnz) && read_raw(dir + "/f_value.f32", data.feature_value, fnnz) &&
            read_raw(dir + "/label.i32", data.label, n) && read_raw(dir + "/split.u32", data.split, n);
  if (!ok) {
    fprintf(stderr, "cannot read the arrays under %s\n", dir.c_str());
    return 2;
  }
  params.num_nodes = (natural)n;
  params.input_dim = (natural)in_dim;
  params.output_dim = (natural)out_dim;
  params.epochs = epochs;
  for (natural s : data.split) {  // Parser::parseSplit, src/parser.cpp:114-132
    if (s == 1) params.train_dim++;
    else if (s == 2) params.val_dim++;
    else if (s == 3) params.test_dim++;
  }
  data.graph_value.resize(gnnz);  // Parser::calculateGraphValues, src/parser.cpp:164-181
  for (size_t src = 0; src < n; src++)
    for (natural i = data.graph.indptr[src]; i < data.graph.indptr[src + 1]; i++) {
      const natural dst = data.graph.indices[i];
      data.graph_value[i] = 1. / sqrtf((data.graph.indptr[src + 1] - data.graph.indptr[src]) *
                                       (data.graph.indptr[dst + 1] - data.graph.indptr[dst]));
    }
  int dev = 0;
  cudaDeviceProp prop;
  cudaGetDevice(&dev);
  cudaGetDeviceProperties(&prop, dev);
  // launch shapes of test/performance_gpu.cpp:37-49 (pubmed's for small graphs, reddit's for large ones)
  if (n > 100000) {
    CudaParams::N_BLOCKS = 16 * prop.multiProcessorCount;
    CudaParams::N_THREADS = 512;
  } else {
    CudaParams::N_BLOCKS = 8 * prop.multiProcessorCount;
    CudaParams::N_THREADS = 256;
  }
  GCN gcn(&params, &adam_params, &data);
  double sum_avg = 0, sum_total = 0, best = 1e30;
  for (int r = 0; r < reps; r++) {
    reset_timer();
    gcn.run();
    sum_avg += gcn.avg_epoch_time;
    sum_total += gcn.total_time;
    if (gcn.avg_epoch_time < best) best = gcn.avg_epoch_time;
  }
  cudaDeviceSynchronize();
  const cudaError_t err = cudaGetLastError();
  printf("{\"impl\": \"reference_gpu\", \"device\": \"%s\", \"nodes\": %zu, \"graph_nnz\": %zu, \"feat_nnz\": %zu, \"epochs\": %u, "
         "\"reps\": %d, \"avg_epoch_ms\": %.6f, \"best_avg_epoch_ms\": %.6f, \"total_s\": %.6f, "
         "\"cuda_error\": \"%s\"}\n",
         prop.name, n, gnnz, fnnz, epochs, reps, sum_avg / reps, best, sum_total / reps, cudaGetErrorString(err));
  Variable::sizes.clear();
  // the reference's static device pointers are freed after the CUDA runtime has shut down (cudaFree -> "driver shutting
  // down", exit code 1 from its own CHECK macro); the measurement is complete, so leave without running static destructors
  fflush(stdout);
  _exit(err == cudaSuccess ? 0 : 1);
}
