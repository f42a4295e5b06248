// variable.cuh -- fp32 device tensor with optional gradient, mirror of include/variable.cuh:11-29 /
// src/variable.cu.  RNG: the reference keeps one Philox state per 4 elements of the largest `rand` Variable
// (dev_rand_states) and every RNG op advances a prefix of them; here the same streams are reproduced statelessly:
// `rng_history` records which prefix lengths have been consumed how often (see gcnb_rng_t in include/gcnb.h).
#ifndef VARIABLE_CUH
#define VARIABLE_CUH
#include <fstream>
#include <map>
#include <string>
#include <vector>
#include "../../../include/gcnb.h"
#include "../include/shared_ptr.cuh"
#include "../include/smart_object.cuh"
#include "../include/utils.cuh"

// RNG streams of ONE model: which prefix lengths were consumed how often + the seed the streams were started with.  The
// reference keeps this state process-wide (static dev_rand_states, CudaParams::SEED), so two live models corrupt each
// other's streams; here every GCN owns a context and binds it to the calling thread for the duration of its calls
// (Variable::rng_bind), which also lets several models train concurrently from several host threads (gcnb_sweep_run).
// Code that uses Variable / Dropout directly, without a GCN, keeps the process-wide history below.
struct GCNRngContext {
  std::map<natural, natural> history;
  natural seed = 19990304;
};

class Variable {
 public:
  inline static std::vector<natural> sizes;                      // sizes of the Variables created with rand=true
  inline static dev_shared_ptr<randState> dev_rand_states;        // always null: kept for source compatibility
  inline static std::map<natural, natural> rng_history;           // groups (ceil(size/4)) -> times consumed
  inline static bool rng_initialized = false;
  dev_shared_ptr<real> dev_data;
  dev_shared_ptr<real> dev_grad;
  natural size, rows, cols;

  Variable(const natural size_, const bool requires_grad = true, const bool rand = false, const natural rows_ = 0,
           const natural cols_ = 0);
  Variable() = default;
  void print(const std::string &what, natural col) const;
  void save(const std::string &file_name, const std::string &what, natural col) const;
  void zero(smart_stream stream) const;
  void zero_grad(smart_stream stream) const;
  void glorot(cudaStream_t stream = nullptr) const;  // (the reference: glorot() on the default stream)
  void set_value(const real value, smart_stream stream) const;
  static void initialize_random();  // src/variable.cu:13-26: (re)starts every Philox stream at draw 0

  // RNG bookkeeping shared with Dropout
  static gcnb_rng_t rng_descriptor();            // history so far, seed = CudaParams::SEED
  static void rng_consume(size_t n_elements);    // an RNG op over n_elements just ran (64-bit: global counts of partitioned models)
  // binds a model's context to the calling thread (nullptr: back to the process-wide history); returns the previous one
  static GCNRngContext *rng_bind(GCNRngContext *ctx);
};
#endif
