// run_model_simulation: the reference's orchestration of one (model, dim, steps, num_runs) job
// (src/data_storage/parallel_compute.rs:150-232) over the GPU hot path and the batched .dat writer.
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "jne_host.hpp"

namespace jne {

SimulationStats run_model_simulation(const Engine& gpu, Model model, uint32_t dim, uint32_t steps, uint64_t num_runs,
                                     const std::string& filename, bool quiet) {
  SimulationStats st;
  if (dim > 255) throw Error(JNE_ERR_INVALID_ARG, "dim must fit the u8 header field");
  std::vector<uint8_t> bitmap((num_runs + 7) / 8);
  uint64_t completed = 0;
  int rc = jne_dat_completed_bitmap(filename.c_str(), model.number, (uint8_t)dim, steps, num_runs, bitmap.data(), &completed);
  if (rc != JNE_OK) {
    const std::string msg = jne_dat_last_error();
    if (msg.find("mismatch") == std::string::npos) throw Error(rc, msg);
    // incompatible parameters: remove and start fresh (parallel_compute.rs:159-175)
    if (!quiet) printf("WARNING: Existing file has incompatible parameters:\n  %s\n", msg.c_str());
    std::remove(filename.c_str());
    std::fill(bitmap.begin(), bitmap.end(), 0);
    completed = 0;
  }
  st.completed_before = completed;
  std::vector<uint32_t> remaining(jne_dat_remaining_seeds(bitmap.data(), num_runs, nullptr, 0));
  jne_dat_remaining_seeds(bitmap.data(), num_runs, remaining.data(), remaining.size());
  if (remaining.empty()) {                              // already complete (:182-198)
    st.total_in_file = completed;
    return st;
  }
  jne_dat_writer* w = nullptr;
  uint64_t existing = 0;
  rc = jne_dat_open(filename.c_str(), model.number, (uint8_t)dim, steps, &existing, &w);
  if (rc != JNE_OK) throw Error(rc, jne_dat_last_error());
  const int p = model.num_eigs(dim);
  // double-buffered: the GPU computes chunk i+1 while chunk i is encoded and written
  const size_t chunk = 1u << 20;
  std::vector<double> buf[2];
  size_t prev_a = 0, prev_n = 0;
  int which = 0;
  try {
    for (size_t a = 0;; a += chunk) {
      const size_t n = a < remaining.size() ? std::min(chunk, remaining.size() - a) : 0;
      int64_t ticket = 0;
      if (n) {
        buf[which].resize(n * p);
        ticket = jne_submit(gpu.ctx(), model.number, dim, steps, remaining.data() + a, n, buf[which].data());
        gpu.check(ticket);
      }
      if (prev_n) {
        rc = jne_dat_append_batch(w, remaining.data() + prev_a, buf[which ^ 1].data(), prev_n, (uint32_t)p);
        if (rc != JNE_OK) throw Error(rc, jne_dat_last_error());
        st.computed += prev_n;
        if (!quiet) printf("Simulation progress: %llu/%llu\n", (unsigned long long)(completed + st.computed), (unsigned long long)num_runs);
      }
      if (n) gpu.check(jne_wait(gpu.ctx(), ticket));
      prev_a = a; prev_n = n; which ^= 1;
      if (!n) break;
    }
  } catch (...) {
    jne_dat_abandon(w);     // leave a trailer-less, resumable file behind, like an interrupted reference run
    throw;
  }
  rc = jne_dat_finish(w);
  if (rc != JNE_OK) throw Error(rc, jne_dat_last_error());
  st.total_in_file = existing + st.computed;
  return st;
}

}  // namespace jne

extern "C" {

// C entry for run_model_simulation (used by the Python mirror and the tests).  stats: 3 x u64
// {completed_before, computed, total_in_file}.  Not in include/jne.h's hot-path section: orchestration helper.
int jne_run_model_simulation(jne_ctx* ctx_unused, uint8_t model, uint32_t dim, uint32_t steps, uint64_t num_runs,
                             const char* filename, int quiet, const int* device_ids, int n_devices, uint64_t* stats) {
  (void)ctx_unused;
  try {
    std::vector<int> devs(device_ids, device_ids + (device_ids ? n_devices : 0));
    jne::Engine gpu(devs);
    const jne::SimulationStats st = jne::run_model_simulation(gpu, jne::Model(model), dim, steps, num_runs, filename, quiet != 0);
    if (stats) { stats[0] = st.completed_before; stats[1] = st.computed; stats[2] = st.total_in_file; }
    return JNE_OK;
  } catch (const jne::Error& e) {
    fprintf(stderr, "jne_run_model_simulation: %s\n", e.what());
    return e.status;
  }
}

}  // extern "C"
