// variable.cpp -- Variable (src/variable.cu:13-171 of the reference) on top of the C ABI.
#include "../include/variable.cuh"
#include <algorithm>

static thread_local GCNRngContext *t_rng_ctx = nullptr;

GCNRngContext *Variable::rng_bind(GCNRngContext *ctx) {
  GCNRngContext *prev = t_rng_ctx;
  t_rng_ctx = ctx;
  return prev;
}

Variable::Variable(const natural size_, const bool requires_grad, const bool rand, const natural rows_,
                   const natural cols_)
    : size(size_), rows(rows_), cols(cols_) {
  dev_data = dev_shared_ptr<real>(size);
  dev_grad = requires_grad ? dev_shared_ptr<real>(size) : dev_shared_ptr<real>();
  if (rand && !t_rng_ctx) sizes.push_back(size);  // (bookkeeping of the reference's state array; models with their own context skip it)
}

void Variable::initialize_random() {
  // reference: allocates ceil(max(sizes)/4) Philox states and curand_init()s them (src/variable.cu:13-26).
  // Stateless equivalent: every stream restarts at draw 0.
  if (t_rng_ctx) t_rng_ctx->history.clear();
  else rng_history.clear();
  rng_initialized = true;
}

gcnb_rng_t Variable::rng_descriptor() {
  gcnb_rng_t r{};
  r.seed = t_rng_ctx ? t_rng_ctx->seed : CudaParams::SEED;
  r.n_hist = 0;
  for (const auto &kv : (t_rng_ctx ? t_rng_ctx->history : rng_history)) {
    if (kv.second == 0) continue;
    if (r.n_hist == GCNB_MAX_RNG_HIST) {
      std::cerr << "Variable: more than " << GCNB_MAX_RNG_HIST << " distinct RNG consumer sizes" << std::endl;
      exit(EXIT_FAILURE);
    }
    r.hist_groups[r.n_hist] = kv.first;
    r.hist_count[r.n_hist] = kv.second;
    r.n_hist++;
  }
  return r;
}

void Variable::rng_consume(size_t n_elements) {
  // GLOBAL element counts of a row-partitioned model exceed 32 bits long before a rank's share does: count the Philox
  // groups in 64 bits; the descriptor's group index is 32-bit (up to 16 Gi elements per consumer), beyond that fail loudly
  const uint64_t groups = ((uint64_t)n_elements + 3) / 4;
  if (groups > 0xffffffffull) {
    std::cerr << "Variable::rng_consume: " << n_elements << " elements exceed the 2^34-element range of the Philox descriptor" << std::endl;
    exit(EXIT_FAILURE);
  }
  (t_rng_ctx ? t_rng_ctx->history : rng_history)[(natural)groups] += 1;
}

void Variable::glorot(cudaStream_t stream) const {
  if (!rng_initialized) {
    std::cerr << "Variable::glorot: Variable must be initialized with rand = true" << std::endl;
    exit(EXIT_FAILURE);
  }
  if (rows == 0 || cols == 0) {
    std::cerr << "Variable::glorot: rows and cols must be set" << std::endl;
    exit(EXIT_FAILURE);
  }
  const gcnb_rng_t rng = rng_descriptor();
  GCNB_CALL(gcnb_glorot_f32(dev_data.get(), size, rows, cols, &rng, stream));  // default stream unless told otherwise, as the reference
  rng_consume(size);
}

void Variable::zero(smart_stream stream) const { dev_data.set_zero(stream); }
void Variable::zero_grad(smart_stream stream) const { dev_grad.set_zero(stream); }

void Variable::set_value(const real value, smart_stream stream) const {
  std::vector<real> host(size, value);
  CHECK_CUDA_ERROR(cudaMemcpyAsync(dev_data.get(), host.data(), size * sizeof(real), cudaMemcpyHostToDevice, stream.get()));
  CHECK_CUDA_ERROR(cudaStreamSynchronize(stream.get()));
}

void Variable::print(const std::string &what, natural col) const {
  std::vector<real> host(size);
  if (what == "data") dev_data.copy_to_host(host.data());
  else if (what == "grad") dev_grad.copy_to_host(host.data());
  else {
    std::cerr << "Variable::print: what must be either 'data' or 'grad'" << std::endl;
    exit(EXIT_FAILURE);
  }
  int count = 0;
  for (natural i = 0; i < 20 * col && i < size; i++) {
    printf("%.4f ", host[i]);
    if (++count % col == 0) printf("\n");
  }
}

void Variable::save(const std::string &file_name, const std::string &what, natural col) const {
  std::vector<real> host(size);
  if (what == "data") dev_data.copy_to_host(host.data());
  else if (what == "grad") dev_grad.copy_to_host(host.data());
  else {
    std::cerr << "Variable::print: what must be either 'data' or 'grad'" << std::endl;
    exit(EXIT_FAILURE);
  }
  std::ofstream file(file_name);
  if (!file.is_open()) {
    std::cerr << "Unable to open file: " << file_name << std::endl;
    return;
  }
  int count = 0;
  for (const auto &element : host) {
    file << element << " ";
    if (++count % col == 0) file << "\n";
  }
  std::cout << "Vector saved to file: " << file_name << std::endl;
}
