// sparse.cuh -- host CSR index and its device mirror, as include/sparse.cuh:11-29 / src/sparse.cu:20-30.
#ifndef SPARSE_CUH
#define SPARSE_CUH
#include <vector>
#include "../include/shared_ptr.cuh"
#include "../include/utils.cuh"

class SparseIndex {
 public:
  std::vector<natural> indices;
  std::vector<natural> indptr;
  void print();
};

class DevSparseIndex {
 public:
  dev_shared_ptr<natural> dev_indices;
  dev_shared_ptr<natural> dev_indptr;
  natural indices_size;
  natural indptr_size;
  DevSparseIndex(const SparseIndex &sparse_index);
  // extension: upload straight from caller-owned (ideally pinned) host arrays, no std::vector staging
  DevSparseIndex(const natural *indptr, size_t indptr_size_, const natural *indices, size_t indices_size_);
};
#endif
