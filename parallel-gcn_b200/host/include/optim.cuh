// optim.cuh -- Adam, mirror of include/optim.cuh:16-50 / src/optim.cu.  step() is ONE multi-tensor launch
// (the reference launches one kernel per weight on two streams); events are still recorded so that
// SparseMatmul/Matmul::forward waits keep working.
#ifndef OPTIM_CUH
#define OPTIM_CUH
#include <cuda_runtime.h>
#include <memory>
#include <utility>
#include <vector>
#include "../include/shared_ptr.cuh"
#include "../include/smart_object.cuh"
#include "../include/timer.h"
#include "../include/utils.cuh"
#include "../include/variable.cuh"
using std::shared_ptr;

struct AdamParams {
  real learning_rate{0.01}, beta1{0.9}, beta2{0.999}, eps{1e-8}, weight_decay{5e-4};
};

struct AdamVariable {
 public:
  dev_shared_ptr<real> dev_data, dev_grad, dev_m, dev_v;
  natural size;
  bool decay;
  smart_stream forward_training_stream;
  AdamVariable(shared_ptr<Variable>, bool, smart_stream &forward_training_stream_);
};

class Adam {
  const AdamParams *params{nullptr};
  natural step_count{0};
  std::vector<AdamVariable> vars;
  smart_stream forward_training_stream;
  std::vector<smart_stream> backward_streams;
  std::vector<smart_event> start_matmul_forward;

 public:
  Adam() {}
  Adam(const std::vector<shared_ptr<Variable>> &weights, const std::vector<bool> &decays, AdamParams const *params_,
       const std::vector<smart_stream> &backward_streams_, std::vector<smart_event> &start_matmul_forward_,
       smart_stream &forward_training_stream_);
  void step();
  // extension used by the fused GCN driver: the whole step on one caller-chosen stream, no events
  void step_on(cudaStream_t stream);
  // the two halves of step_on, for CUDA-graph replay: advance the step counter and return this step's step size; launch
  // the update with a given step size
  real advance();
  void launch_on(cudaStream_t stream, real step_size);
  natural steps() const { return step_count; }
};
#endif
