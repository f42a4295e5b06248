// smart_object.cuh -- ref-counted CUDA stream / event handles, mirror of include/smart_object.cuh:13-52 and
// src/smart_object.cu.  Copies share the handle; the last owner destroys it.  Single host thread, like the reference.
#ifndef SMART_OBJECT_CUH
#define SMART_OBJECT_CUH
#include <cuda_runtime.h>
#include <memory>
#include "../include/utils.cuh"

enum StreamPriority { Low = 0, High = -5 };

class smart_stream {
 public:
  smart_stream() : smart_stream(Low) {}
  explicit smart_stream(StreamPriority priority) {
    cudaStream_t s = nullptr;
    CHECK_CUDA_ERROR(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, static_cast<int>(priority)));
    h_ = std::shared_ptr<CUstream_st>(s, [](cudaStream_t p) { if (p) cudaStreamDestroy(p); });
  }
  cudaStream_t get() const { return h_.get(); }
  size_t getRefCount() const { return static_cast<size_t>(h_.use_count()); }

 private:
  std::shared_ptr<CUstream_st> h_;
};

class smart_event {
 public:
  smart_event() {
    cudaEvent_t e = nullptr;
    CHECK_CUDA_ERROR(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h_ = std::shared_ptr<CUevent_st>(e, [](cudaEvent_t p) { if (p) cudaEventDestroy(p); });
  }
  cudaEvent_t get() const { return h_.get(); }
  size_t getRefCount() const { return static_cast<size_t>(h_.use_count()); }

 private:
  std::shared_ptr<CUevent_st> h_;
};
#endif
