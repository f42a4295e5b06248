// module.cuh -- the operator chain, mirror of include/module.cuh:21-145 (same class names, constructor signatures,
// const forward/backward, in-place semantics, masks kept between forward(true) and backward).  Each method is a
// thin call into the C ABI of libgcn_b200 (include/gcnb.h); stream/event choreography follows src/module.cu so
// that code composing these modules by hand (as the reference's GCN does) keeps its ordering guarantees.
#ifndef MODULE_CUH
#define MODULE_CUH
#include <cuda_runtime.h>
#include <memory>
using std::shared_ptr;
using std::unique_ptr;

#include "../include/reduction.cuh"
#include "../include/shared_ptr.cuh"
#include "../include/smart_object.cuh"
#include "../include/sparse.cuh"
#include "../include/timer.h"
#include "../include/utils.cuh"
#include "../include/variable.cuh"

class Module {
 public:
  virtual void forward(bool, const smart_stream &) const = 0;
  virtual void backward(const smart_stream &) const = 0;
  virtual void set_num_samples(natural){};
  virtual natural get_num_samples() const { return 0; };
  virtual ~Module(){};
};

// shared, lazily built load-balancing plan of one DevSparseIndex (keyed by its device pointers)
struct SpmmPlanCache;

class Dropout : public Module {
  shared_ptr<Variable> in;
  dev_shared_ptr<bool> dev_mask;
  real p;

 public:
  Dropout(shared_ptr<Variable> in_, real p_);
  void forward(bool, const smart_stream &) const;
  void backward(const smart_stream &) const;
};

class SparseMatmul : public Module {
  shared_ptr<Variable> a, b, c;
  DevSparseIndex *sp;
  natural m, n, p;
  smart_event start_matmul_forward;
  smart_event start_set_input;
  shared_ptr<SpmmPlanCache> plans;

 public:
  SparseMatmul(shared_ptr<Variable> a_, shared_ptr<Variable> b_, shared_ptr<Variable> c_, DevSparseIndex *sp_,
               natural m_, natural n_, natural p_, smart_event &start_matmul_forward_, smart_event &start_set_input_);
  ~SparseMatmul(){};
  void forward(bool, const smart_stream &) const;
  void backward(const smart_stream &) const;
};

class GraphSum : public Module {
  shared_ptr<Variable> in, out;
  DevSparseIndex *graph;
  dev_shared_ptr<real> dev_graph_value;
  natural dim;
  bool generate_event;
  smart_event start_matmul_backward;
  shared_ptr<SpmmPlanCache> plans;

 public:
  GraphSum(shared_ptr<Variable> in_, shared_ptr<Variable> out_, DevSparseIndex *graph_,
           dev_shared_ptr<real> dev_graph_value_, natural dim_, bool generate_event_,
           smart_event &start_matmul_backward_);
  ~GraphSum() {}
  void forward(bool, const smart_stream &) const;
  void backward(const smart_stream &) const;
};

class ReLU : public Module {
  shared_ptr<Variable> in;
  dev_shared_ptr<bool> dev_mask;

 public:
  ReLU(shared_ptr<Variable> in_);
  void forward(bool, const smart_stream &) const;
  void backward(const smart_stream &) const;
};

class Matmul : public Module {
  shared_ptr<Variable> a, b, c;
  natural m, n, p;
  smart_event event_forward;
  smart_event event_backward;
  smart_stream my_stream;
  dev_shared_ptr<real> workspace;  // split-K partials of the weight gradient (fixed-order reduce)

 public:
  Matmul(shared_ptr<Variable> a_, shared_ptr<Variable> b_, shared_ptr<Variable> c_, natural m_, natural n_,
         natural p_, smart_event &event_forward_, smart_event &event_backward_, const smart_stream &stream_);
  ~Matmul() {}
  void forward(bool, const smart_stream &) const;
  void backward(const smart_stream &) const;
};

class CrossEntropyLoss : public Module {
  shared_ptr<Variable> logits;
  dev_shared_ptr<integer> dev_truth;
  pinned_host_ptr<real> loss;
  natural num_classes;
  dev_shared_ptr<real> dev_loss_res;  // [loss sum, wrong count bits, labelled count bits, pad]
  dev_shared_ptr<natural> workspace;
  smart_event start_backward;

 public:
  natural num_samples;
  CrossEntropyLoss(shared_ptr<Variable> logits_, dev_shared_ptr<integer> dev_truth_, pinned_host_ptr<real> loss_,
                   natural num_classes_, smart_event &event);
  ~CrossEntropyLoss(){};
  void set_num_samples(natural num_samples_);
  natural get_num_samples() const;
  void forward(bool, const smart_stream &) const;
  void backward(const smart_stream &) const;
  // extension: the fused kernel also yields the wrong-prediction count (GCN::get_accuracy, src/gcn.cu:264-289)
  const real *device_result() const { return dev_loss_res.get(); }
};
#endif
