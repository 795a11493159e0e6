// C++ host-side mirror of the reference's interface for the hot path, over the C ABI of include/jne.h.
// Names, argument meaning and error behaviour follow the reference (paths relative to its root):
//   jne::Model                              JohansenModel                         src/johansen_models.rs:6-51
//   jne::calculate_eigenvalues              calculate_eigenvalues                 src/johansen_statistics.rs:59-85
//   jne::calculate_eigenvalues_parallel     calculate_eigenvalues_parallel        src/data_storage/parallel_compute.rs:14-41
//   jne::run_model_simulation               run_model_simulation                  src/data_storage/parallel_compute.rs:150-232
//   jne::run_models_simulation              the CLI's model loop over one dim     src/main.rs:109-114
//   jne::get_filename                       EigenvalueSimulation::get_filename    src/data_storage/simulation.rs:98-122
// The reference panics on hot-path failures; here they surface as jne::Error (never abort, never a CPU fallback).
#pragma once
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/jne.h"
#include "../../include/jne_dat.h"

namespace jne {

struct Error : std::runtime_error {
  int status;
  Error(int s, const std::string& m) : std::runtime_error("jne status " + std::to_string(s) + ": " + m), status(s) {}
};

// src/johansen_models.rs:6-51 -- the number <-> variant table is part of the .dat contract
struct Model {
  uint8_t number;
  explicit Model(uint8_t n) : number(n) { if (n > 4) throw Error(JNE_ERR_INVALID_ARG, "model must be 0..4"); }
  uint8_t to_number() const { return number; }
  bool has_intercept() const { return number != 0; }
  bool has_trend() const { return number >= 3; }
  int num_eigs(uint32_t dim) const { return (number == 1 || number == 3) ? (int)dim + 1 : (int)dim; }
  static Model default_model() { return Model(2); }
};

class Engine {
 public:
  explicit Engine(const std::vector<int>& devices = {}) {
    const int rc = jne_init(devices.empty() ? nullptr : devices.data(), (int)devices.size(), &ctx_);
    if (rc != JNE_OK) throw Error(rc, jne_last_error(nullptr));
  }
  // borrow an existing context (not shut down by the returned object)
  static Engine* borrow(jne_ctx* ctx) {
    if (!ctx) throw Error(JNE_ERR_INVALID_ARG, "ctx is NULL");
    return new Engine(ctx, 0);
  }
  ~Engine() { if (owned_) jne_shutdown(ctx_); }
  Engine(const Engine&) = delete;
  Engine& operator=(const Engine&) = delete;
  jne_ctx* ctx() const { return ctx_; }
  void check(int64_t rc) const { if (rc < 0) throw Error((int)rc, jne_last_error(ctx_)); }

  std::vector<double> eigs_batch(Model m, uint32_t dim, uint32_t steps, const std::vector<uint32_t>& seeds) const {
    std::vector<double> out(seeds.size() * m.num_eigs(dim));
    check(jne_eigs_batch(ctx_, m.number, dim, steps, seeds.data(), seeds.size(), out.data()));
    return out;
  }

 private:
  Engine(jne_ctx* ctx, int) : ctx_(ctx), owned_(false) {}
  jne_ctx* ctx_ = nullptr;
  bool owned_ = true;
};

// src/johansen_statistics.rs:59-85: one run, eigenvalues descending
inline std::vector<double> calculate_eigenvalues(const Engine& gpu, uint32_t dim, uint32_t steps, uint32_t seed, Model model) {
  return gpu.eigs_batch(model, dim, steps, {seed});
}

// The mpsc::Sender<(u32, Vec<f64>)> of the reference, as a callback: (seed, eigenvalues, count)
using Sender = std::function<void(uint32_t, const double*, int)>;

// src/data_storage/parallel_compute.rs:14-41.  The reference walks 10 000-seed chunks through rayon; here every
// (large) chunk is one jne_submit, overlapped with delivering the previous chunk's rows to `sender`.
inline void calculate_eigenvalues_parallel(const Engine& gpu, uint32_t dim, uint32_t steps, const std::vector<uint32_t>& seeds,
                                           Model model, const Sender& sender, bool /*quiet*/, size_t chunk = 1u << 20) {
  const int p = model.num_eigs(dim);
  std::vector<double> buf[2];
  // declared after buf: if `sender` throws while a ticket is in flight, unwinding joins the worker (jne_wait) before
  // the buffer it writes into is freed, and the context is left without a pending ticket
  struct Inflight {
    jne_ctx* ctx; int64_t ticket;
    ~Inflight() { if (ticket > 0) jne_wait(ctx, ticket); }
  } inflight{gpu.ctx(), 0};
  size_t prev_a = 0, prev_n = 0;
  int which = 0;
  for (size_t a = 0; a < seeds.size() || prev_n; a += chunk) {
    const size_t n = a < seeds.size() ? std::min(chunk, seeds.size() - a) : 0;
    if (n) {
      buf[which].resize(n * p);
      const int64_t ticket = jne_submit(gpu.ctx(), model.number, dim, steps, seeds.data() + a, n, buf[which].data());
      gpu.check(ticket);
      inflight.ticket = ticket;
    }
    for (size_t i = 0; i < prev_n; ++i) sender(seeds[prev_a + i], buf[which ^ 1].data() + i * p, p);
    if (n) { const int64_t t = inflight.ticket; inflight.ticket = 0; gpu.check(jne_wait(gpu.ctx(), t)); }
    prev_a = a; prev_n = n; which ^= 1;
    if (!n) break;
  }
}

// src/data_storage/simulation.rs:98-122: data/eigenvalues_model{m}_dim{d}_steps{T}.dat
inline std::string get_filename(const std::string& dir, Model model, uint32_t dim, uint32_t steps) {
  return dir + "/eigenvalues_model" + std::to_string(model.number) + "_dim" + std::to_string(dim) + "_steps" +
         std::to_string(steps) + ".dat";
}

struct SimulationStats { uint64_t completed_before = 0, computed = 0, total_in_file = 0; };

// src/data_storage/parallel_compute.rs:150-232: resume scan -> remaining seeds -> compute -> append -> trailer.
// Parameter mismatch in an existing file: delete and restart (:159-175).  Implemented in jne_host.cpp.
SimulationStats run_model_simulation(const Engine& gpu, Model model, uint32_t dim, uint32_t steps, uint64_t num_runs,
                                     const std::string& filename, bool quiet);


// The CLI's loop `for &model in &models_vec { EigenvalueSimulation::new(model, dim, steps, num_runs).run_simulation() }`
// (src/main.rs:109-114) as ONE pass: the Brownian path of a seed does not depend on the model (src/rng_matrix.rs:11),
// so the fused kernel (jne_eigs_batch_multi) serves every selected model's file from one evaluation of the path.
// Each file is resumed on its own (scan, mismatch -> restart, already complete -> untouched), seeds are grouped by
// the set of models that still lack them, and every file ends up byte-identical to what run_model_simulation
// writes for that model.  filenames[m] is used when bit m of model_mask is set.  stats[m] per selected model.
void run_models_simulation(const Engine& gpu, uint32_t model_mask, uint32_t dim, uint32_t steps, uint64_t num_runs,
                           const std::string (&filenames)[5], bool quiet, SimulationStats (&stats)[5]);

}  // namespace jne
