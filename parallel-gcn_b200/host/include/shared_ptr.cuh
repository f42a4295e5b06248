// shared_ptr.cuh -- ref-counted device / pinned-host buffers, mirror of include/shared_ptr.cuh:8-330
// (get, isNull, useCount, copy_to_device[_async], copy_to_host[_async], set_zero, get_n_elements, print).
// Element counts are size_t here (the reference stores them in an int: tensors > 2^31 elements impossible there).
#ifndef SHARED_PTR_CUH
#define SHARED_PTR_CUH
#include <cuda_runtime.h>
#include <memory>
#include <vector>
#include "../include/smart_object.cuh"
#include "../include/utils.cuh"

template <typename T>
class dev_shared_ptr {
 public:
  dev_shared_ptr() = default;
  explicit dev_shared_ptr(size_t n_elements_) : n_elements(n_elements_) {
    T *p = nullptr;
    CHECK_CUDA_ERROR(cudaMalloc(&p, (n_elements_ ? n_elements_ : 1) * sizeof(T)));
    h_ = std::shared_ptr<T>(p, [](T *q) { if (q) cudaFree(q); });
  }
  T *get() const { return h_.get(); }
  bool isNull() const { return h_.get() == nullptr; }
  int useCount() const { return static_cast<int>(h_.use_count()); }
  void copy_to_device(const T *source) const {
    CHECK_CUDA_ERROR(cudaMemcpy(h_.get(), source, n_elements * sizeof(T), cudaMemcpyHostToDevice));
  }
  void copy_to_device_async(const T *source, smart_stream stream) const {
    CHECK_CUDA_ERROR(cudaMemcpyAsync(h_.get(), source, n_elements * sizeof(T), cudaMemcpyHostToDevice, stream.get()));
  }
  void copy_to_host(T *destination) const {
    CHECK_CUDA_ERROR(cudaMemcpy(destination, h_.get(), n_elements * sizeof(T), cudaMemcpyDeviceToHost));
  }
  void copy_to_host_async(T *destination, smart_stream stream) const {
    CHECK_CUDA_ERROR(cudaMemcpyAsync(destination, h_.get(), n_elements * sizeof(T), cudaMemcpyDeviceToHost, stream.get()));
  }
  void set_zero(smart_stream stream) const {
    CHECK_CUDA_ERROR(cudaMemsetAsync(h_.get(), 0, n_elements * sizeof(T), stream.get()));
  }
  size_t get_n_elements() const { return n_elements; }
  void print(unsigned col) {
    std::vector<T> host(n_elements);
    copy_to_host(host.data());
    unsigned count = 0;
    for (size_t i = 0; i < n_elements; i++) {
      std::cout << host[i] << " ";
      if (++count % col == 0) std::cout << std::endl;
    }
  }

 private:
  std::shared_ptr<T> h_;
  size_t n_elements = 0;
};

template <typename T>
class pinned_host_ptr {
 public:
  pinned_host_ptr() = default;
  explicit pinned_host_ptr(size_t n_elements_) : n_elements(n_elements_) {
    T *p = nullptr;
    CHECK_CUDA_ERROR(cudaMallocHost(&p, (n_elements_ ? n_elements_ : 1) * sizeof(T)));
    h_ = std::shared_ptr<T>(p, [](T *q) { if (q) cudaFreeHost(q); });
  }
  T *get() const { return h_.get(); }
  bool isNull() const { return h_.get() == nullptr; }
  int useCount() const { return static_cast<int>(h_.use_count()); }
  T &operator*() const { return *h_.get(); }
  void set_zero(smart_stream) const { for (size_t i = 0; i < n_elements; i++) h_.get()[i] = T(); }
  size_t get_n_elements() const { return n_elements; }

 private:
  std::shared_ptr<T> h_;
  size_t n_elements = 0;
};
#endif
