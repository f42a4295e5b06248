// runtime.cpp -- L0 helpers of the host mirror: fatal error convention, GPU info, timers, SparseIndex.
// Reference: include/utils.cuh:25-111, src/timer.cpp, src/sparse.cu.
#include <cstdio>
#include "../include/sparse.cuh"
#include "../include/timer.h"
#include "../include/utils.cuh"
#include "../../../include/gcnb.h"

void gcnb_check_fatal(int code, const char *what, const char *file, int line) {
  if (code == 0) return;
  std::cerr << "libgcn_b200 error at: " << file << ":" << line << std::endl;
  std::cerr << gcnb_error_string(code) << " " << what << std::endl;
  std::exit(EXIT_FAILURE);
}

void print_gpu_info() {
  int dev;
  cudaDeviceProp devProp;
  CHECK_CUDA_ERROR(cudaGetDevice(&dev));
  CHECK_CUDA_ERROR(cudaGetDeviceProperties(&devProp, dev));
  std::cout << std::endl;
  std::cout << "GPU INFORMATIONS:" << std::endl;
  std::cout << "multiProcessorCount: " << devProp.multiProcessorCount << std::endl;
  std::cout << "maxBlocksPerMultiProcessor: " << devProp.maxBlocksPerMultiProcessor << std::endl;
  std::cout << "maxThreadsPerMultiProcessor: " << devProp.maxThreadsPerMultiProcessor << std::endl;
  std::cout << "maxThreadsPerBlock: " << devProp.maxThreadsPerBlock << std::endl;
  std::cout << "warpSize: " << devProp.warpSize << std::endl;
  std::cout << "sharedMemPerBlock [KB]: " << devProp.sharedMemPerBlock / 1024 << std::endl;
  std::cout << "sharedMemPerMultiprocessor [KB]: " << devProp.sharedMemPerMultiprocessor / 1024 << std::endl;
  std::cout << "totalGlobalMem [MB]: " << devProp.totalGlobalMem / 1048576 << std::endl;
  std::cout << std::endl;
}

void timer_start(timer_instance t) { tmr_t0[t] = std::chrono::high_resolution_clock::now(); }
float timer_stop(timer_instance t) {
  const float count =
      std::chrono::duration_cast<std::chrono::duration<float>>(std::chrono::high_resolution_clock::now() - tmr_t0[t])
          .count();
  tmr_sum[t] += count;
  return count;
}
float timer_total(timer_instance t) { return tmr_sum[t]; }
void reset_timer() {
  for (int i = 0; i < __NUM_TMR; i++) tmr_sum[i] = 0.0f;
}

void SparseIndex::print() {
  std::cout << "---sparse index info--" << std::endl;
  std::cout << "indptr: ";
  for (auto i : indptr) std::cout << i << " ";
  std::cout << std::endl;
  std::cout << "indices: ";
  for (auto i : indices) std::cout << i << " ";
  std::cout << std::endl;
}

DevSparseIndex::DevSparseIndex(const SparseIndex &sparse_index) {
  indices_size = static_cast<natural>(sparse_index.indices.size());
  indptr_size = static_cast<natural>(sparse_index.indptr.size());
  dev_indices = dev_shared_ptr<natural>(indices_size);
  dev_indptr = dev_shared_ptr<natural>(indptr_size);
  dev_indices.copy_to_device(sparse_index.indices.data());
  dev_indptr.copy_to_device(sparse_index.indptr.data());
}

DevSparseIndex::DevSparseIndex(const natural *indptr, size_t indptr_size_, const natural *indices, size_t indices_size_) {
  indices_size = static_cast<natural>(indices_size_);
  indptr_size = static_cast<natural>(indptr_size_);
  dev_indices = dev_shared_ptr<natural>(indices_size);
  dev_indptr = dev_shared_ptr<natural>(indptr_size);
  dev_indices.copy_to_device(indices);
  dev_indptr.copy_to_device(indptr);
}
