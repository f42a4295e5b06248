// optim.cpp -- Adam (src/optim.cu:7-95): step_size = lr*sqrtf(1-b2^t)/(1-b1^t) on the host in fp32, then one
// multi-tensor kernel launch for all weights (reference: one launch per weight, on two streams).
#include "../include/optim.cuh"
#include <cmath>
#include "../../../include/gcnb.h"

AdamVariable::AdamVariable(shared_ptr<Variable> var, bool decay_, smart_stream &forward_training_stream_)
    : dev_data(var->dev_data), dev_grad(var->dev_grad), size(var->size), decay(decay_),
      forward_training_stream(forward_training_stream_) {
  dev_m = dev_shared_ptr<real>(size);
  dev_v = dev_shared_ptr<real>(size);
  dev_m.set_zero(forward_training_stream);
  dev_v.set_zero(forward_training_stream);
  CHECK_CUDA_ERROR(cudaStreamSynchronize(forward_training_stream.get()));
}

Adam::Adam(const std::vector<shared_ptr<Variable>> &weights, const std::vector<bool> &decays, AdamParams const *params_,
           const std::vector<smart_stream> &backward_streams_, std::vector<smart_event> &start_matmul_forward_,
           smart_stream &forward_training_stream_)
    : params(params_), step_count(0), forward_training_stream(forward_training_stream_),
      backward_streams(backward_streams_), start_matmul_forward(start_matmul_forward_) {
  if (weights.size() != decays.size()) {
    std::cout << "Error in Adam constructor: weights and decays must have the same size" << std::endl;
    exit(1);
  }
  if (weights.size() > GCNB_MAX_TENSORS) {
    std::cout << "Error in Adam constructor: at most " << GCNB_MAX_TENSORS << " weights" << std::endl;
    exit(1);
  }
  for (natural i = 0; i < weights.size(); i++) vars.emplace_back(weights[i], decays[i], forward_training_stream);
}

real Adam::advance() {
  step_count++;
  return params->learning_rate * sqrtf(1 - powf(params->beta2, step_count)) / (1 - powf(params->beta1, step_count));
}

void Adam::launch_on(cudaStream_t stream, real step_size) {
  gcnb_adam_tensors_t t{};
  t.n_tensors = static_cast<int>(vars.size());
  for (size_t i = 0; i < vars.size(); i++) {
    t.w[i] = vars[i].dev_data.get();
    t.g[i] = vars[i].dev_grad.get();
    t.m[i] = vars[i].dev_m.get();
    t.v[i] = vars[i].dev_v.get();
    t.size[i] = vars[i].size;
    t.decay[i] = vars[i].decay ? 1 : 0;
  }
  GCNB_CALL(gcnb_adam_step_f32(&t, params->weight_decay, params->beta1, params->beta2, params->eps, step_size, stream));
}

void Adam::step_on(cudaStream_t stream) { launch_on(stream, advance()); }

void Adam::step() {
  // The reference updates W0 on backward_streams[0] and the rest on backward_streams[1] (src/optim.cu:76-92).
  // One fused launch needs both producers: join stream 1 into stream 0, launch there, then publish the events
  // every forward matmul waits on.
  if (backward_streams.size() < 2 || start_matmul_forward.size() < vars.size()) {
    std::cout << "Error in Adam::step: needs 2 backward streams and one event per weight" << std::endl;
    exit(1);
  }
  smart_event joined;
  CHECK_CUDA_ERROR(cudaEventRecord(joined.get(), backward_streams[1].get()));
  CHECK_CUDA_ERROR(cudaStreamWaitEvent(backward_streams[0].get(), joined.get()));
  step_on(backward_streams[0].get());
  for (size_t i = 0; i < vars.size(); i++)
    CHECK_CUDA_ERROR(cudaEventRecord(start_matmul_forward[i].get(), backward_streams[0].get()));
  // later work queued on stream 1 must also see the update
  CHECK_CUDA_ERROR(cudaStreamWaitEvent(backward_streams[1].get(), start_matmul_forward[0].get()));
}
