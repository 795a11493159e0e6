// run_model_simulation: the reference's orchestration of one (model, dim, steps, num_runs) job
// (src/data_storage/parallel_compute.rs:150-232) over the GPU hot path and the batched .dat writer.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>

#include "jne_host.hpp"

namespace jne {

namespace {
// An outstanding jne_submit ticket writes into a caller buffer from the context's worker thread.  If anything throws
// while it is in flight (a full disk in the writer, a failing sender), the buffer must outlive the worker and the
// context must not be left with a pending ticket: the guard is declared AFTER the buffers, so unwinding joins the
// worker (jne_wait) before they are freed.
struct TicketGuard {
  jne_ctx* ctx;
  int64_t ticket = 0;
  explicit TicketGuard(jne_ctx* c) : ctx(c) {}
  ~TicketGuard() { if (ticket > 0) jne_wait(ctx, ticket); }
  int wait() { const int64_t t = ticket; ticket = 0; return t > 0 ? jne_wait(ctx, t) : JNE_OK; }
  TicketGuard(const TicketGuard&) = delete;
  TicketGuard& operator=(const TicketGuard&) = delete;
};
}  // namespace

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
  if (completed >= num_runs) {                          // the reference counts ALL records of the file, also seeds beyond
    st.total_in_file = completed;                       // num_runs, and skips the job (:182-188)
    return st;
  }
  std::vector<uint32_t> remaining(jne_dat_remaining_seeds(bitmap.data(), num_runs, nullptr, 0));
  jne_dat_remaining_seeds(bitmap.data(), num_runs, remaining.data(), remaining.size());
  if (remaining.empty()) {                              // already complete (:190-198)
    st.total_in_file = completed;
    return st;
  }
  if (!quiet && completed)   // a resumed file continues with THIS library's stream (Philox), whoever wrote the head of it
    printf("Resuming %s: %llu of %llu runs present, %zu to compute\n", filename.c_str(), (unsigned long long)completed,
           (unsigned long long)num_runs, remaining.size());
  jne_dat_writer* w = nullptr;
  uint64_t existing = 0;
  rc = jne_dat_open(filename.c_str(), model.number, (uint8_t)dim, steps, &existing, &w);
  if (rc != JNE_OK) throw Error(rc, jne_dat_last_error());
  const int p = model.num_eigs(dim);
  // double-buffered: the GPU computes chunk i+1 while chunk i is encoded and written
  const size_t chunk = 1u << 20;
  std::vector<double> buf[2];
  TicketGuard inflight(gpu.ctx());     // after buf: joined before the buffers die on any exit path
  size_t prev_a = 0, prev_n = 0;
  int which = 0;
  try {
    for (size_t a = 0;; a += chunk) {
      const size_t n = a < remaining.size() ? std::min(chunk, remaining.size() - a) : 0;
      if (n) {
        buf[which].resize(n * p);
        const int64_t ticket = jne_submit(gpu.ctx(), model.number, dim, steps, remaining.data() + a, n, buf[which].data());
        gpu.check(ticket);
        inflight.ticket = ticket;
      }
      if (prev_n) {
        rc = jne_dat_append_batch(w, remaining.data() + prev_a, buf[which ^ 1].data(), prev_n, (uint32_t)p);
        if (rc != JNE_OK) throw Error(rc, jne_dat_last_error());
        st.computed += prev_n;
        if (!quiet) printf("Simulation progress: %llu/%llu\n", (unsigned long long)(completed + st.computed), (unsigned long long)num_runs);
      }
      if (n) gpu.check(inflight.wait());
      prev_a = a; prev_n = n; which ^= 1;
      if (!n) break;
    }
  } catch (...) {
    inflight.wait();        // the worker still writes into buf: join it before anything is torn down
    jne_dat_abandon(w);     // leave a trailer-less, resumable file behind, like an interrupted reference run
    throw;
  }
  rc = jne_dat_finish(w);
  if (rc != JNE_OK) throw Error(rc, jne_dat_last_error());
  st.total_in_file = existing + st.computed;
  return st;
}

void run_models_simulation(const Engine& gpu, uint32_t model_mask, uint32_t dim, uint32_t steps, uint64_t num_runs,
                           const std::string (&filenames)[5], bool quiet, SimulationStats (&stats)[5]) {
  if (model_mask == 0 || model_mask > 31u) throw Error(JNE_ERR_INVALID_ARG, "model_mask must select models 0..4");
  if (dim > 255) throw Error(JNE_ERR_INVALID_ARG, "dim must fit the u8 header field");
  // ---- resume scan per file; need[s-1] = models that still lack seed s ----
  std::vector<uint8_t> need(num_runs, 0);
  {
    std::vector<uint8_t> bitmap((num_runs + 7) / 8);
    for (int m = 0; m < 5; ++m) {
      stats[m] = SimulationStats{};
      if (!((model_mask >> m) & 1u)) continue;
      uint64_t completed = 0;
      int rc = jne_dat_completed_bitmap(filenames[m].c_str(), (uint8_t)m, (uint8_t)dim, steps, num_runs, bitmap.data(), &completed);
      if (rc != JNE_OK) {
        const std::string msg = jne_dat_last_error();
        if (msg.find("mismatch") == std::string::npos) throw Error(rc, msg);
        if (!quiet) printf("WARNING: Existing file has incompatible parameters:\n  %s\n", msg.c_str());
        std::remove(filenames[m].c_str());               // parallel_compute.rs:159-175
        std::fill(bitmap.begin(), bitmap.end(), 0);
        completed = 0;
      }
      stats[m].completed_before = completed;
      stats[m].total_in_file = completed;
      if (completed >= num_runs) continue;               // parallel_compute.rs:182-188: enough records, whatever their seeds
      for (uint64_t s = 0; s < num_runs; ++s)
        if (!((bitmap[s >> 3] >> (s & 7)) & 1u)) need[s] |= (uint8_t)(1u << m);
    }
  }
  // ---- writers for the files that lack something (a complete file is not touched, parallel_compute.rs:182-198) ----
  jne_dat_writer* w[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  uint64_t existing[5] = {0, 0, 0, 0, 0};
  uint32_t lacking = 0;
  for (uint64_t s = 0; s < num_runs; ++s) lacking |= need[s];
  auto abandon_all = [&]() { for (auto& x : w) if (x) { jne_dat_abandon(x); x = nullptr; } };
  for (int m = 0; m < 5; ++m) {
    if (!((lacking >> m) & 1u)) continue;
    const int rc = jne_dat_open(filenames[m].c_str(), (uint8_t)m, (uint8_t)dim, steps, &existing[m], &w[m]);
    if (rc != JNE_OK) { const std::string msg = jne_dat_last_error(); abandon_all(); throw Error(rc, msg); }
  }
  // ---- one fused pass per distinct set of lacking models, double-buffered against the writers ----
  // ~20 ms of GPU work per device at dim 12, T 10 000: short tail, writers well ahead
  const int n_dev = std::max(1, jne_ctx_device_count(gpu.ctx()));
  const size_t chunk_max = (size_t)(1u << 16) * (size_t)n_dev;
  const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
  std::vector<double> buf[2];
  TicketGuard inflight(gpu.ctx());     // after buf: joined before the buffers die on any exit path
  // seeds by the set of models that lack them, one pass (ascending within a group)
  std::vector<uint32_t> groups[32];
  {
    uint64_t counts[32] = {0};
    for (uint64_t s = 0; s < num_runs; ++s) ++counts[need[s]];
    for (uint32_t mask = 1; mask < 32; ++mask) groups[mask].reserve(counts[mask]);
    for (uint64_t s = 0; s < num_runs; ++s)
      if (need[s]) groups[need[s]].push_back((uint32_t)(s + 1));
    std::vector<uint8_t>().swap(need);
  }
  try {
    for (uint32_t mask = 1; mask < 32; ++mask) {
      const std::vector<uint32_t>& seeds = groups[mask];
      if (seeds.empty()) continue;
      // at least four chunks per group, so that the writers overlap the devices in small jobs as well
      const size_t chunk = std::min(chunk_max, std::max<size_t>(1u << 16, (seeds.size() + 3) / 4));
      const uint32_t width = (uint32_t)jne_multi_width(mask, dim);
      uint32_t off[5], pm[5];
      { uint32_t o = 0; for (int m = 0; m < 5; ++m) { pm[m] = (uint32_t)Model((uint8_t)m).num_eigs(dim); off[m] = o; if ((mask >> m) & 1u) o += pm[m]; } }
      size_t prev_a = 0, prev_n = 0;
      int which = 0;
      for (size_t a = 0;; a += chunk) {
        const size_t n = a < seeds.size() ? std::min(chunk, seeds.size() - a) : 0;
        if (n) {
          buf[which].resize(n * width);
          const int64_t ticket = jne_submit_multi(gpu.ctx(), mask, dim, steps, seeds.data() + a, n, buf[which].data());
          gpu.check(ticket);
          inflight.ticket = ticket;
        }
        if (prev_n) {                                  // one writer thread per file, as many files as models in the mask
          std::thread th[5];
          int rcs[5] = {0, 0, 0, 0, 0};
          std::string errs[5];
          const double* rows = buf[which ^ 1].data();
          try {
            for (int m = 0; m < 5; ++m) {
              if (!((mask >> m) & 1u)) continue;
              th[m] = std::thread([&, m]() {
                // encoders per file: what the host has beyond one thread per file, as far as the devices need it
                static const int enc_env = [] { const char* e = getenv("JNE_DAT_ENCODERS"); return e ? atoi(e) : 0; }();
                const int enc = enc_env > 0 ? enc_env : std::max(1, std::min({4, n_dev, hw / (2 * __builtin_popcount(mask))}));
                rcs[m] = jne_dat_append_batch_strided_mt(w[m], seeds.data() + prev_a, rows + off[m], prev_n, pm[m], width, enc);
                if (rcs[m] != JNE_OK) errs[m] = jne_dat_last_error();
              });
            }
          } catch (...) {                              // std::thread could not start: join the ones that did, then unwind
            for (auto& t : th) if (t.joinable()) t.join();
            throw;
          }
          for (auto& t : th) if (t.joinable()) t.join();
          for (int m = 0; m < 5; ++m) {
            if (rcs[m] != JNE_OK) throw Error(rcs[m], errs[m]);
            if ((mask >> m) & 1u) stats[m].computed += prev_n;
          }
          if (!quiet) printf("Simulation progress (models mask 0x%x): %llu/%llu seeds\n", mask,
                             (unsigned long long)(prev_a + prev_n), (unsigned long long)seeds.size());
        }
        if (n) gpu.check(inflight.wait());
        prev_a = a; prev_n = n; which ^= 1;
        if (!n) break;
      }
    }
  } catch (...) {
    inflight.wait();      // the worker still writes into buf: join it before anything is torn down
    abandon_all();        // trailer-less, resumable files, like an interrupted reference run
    throw;
  }
  for (int m = 0; m < 5; ++m) {
    if (!w[m]) continue;
    const int rc = jne_dat_finish(w[m]);
    w[m] = nullptr;
    if (rc != JNE_OK) { const std::string msg = jne_dat_last_error(); abandon_all(); throw Error(rc, msg); }
    stats[m].total_in_file = existing[m] + stats[m].computed;
  }
}

}  // namespace jne

extern "C" {

// C entry for run_models_simulation.  filenames: 5 entries (NULL for models outside model_mask); stats: 5 x 3 x u64.
int jne_run_models_simulation(jne_ctx* ctx, uint32_t model_mask, uint32_t dim, uint32_t steps, uint64_t num_runs,
                              const char* const* filenames, int quiet, const int* device_ids, int n_devices, uint64_t* stats) {
  try {
    std::vector<int> devs(device_ids, device_ids + (device_ids ? n_devices : 0));
    std::string names[5];
    for (int m = 0; m < 5; ++m) {
      if (!((model_mask >> m) & 1u)) continue;
      if (!filenames || !filenames[m]) throw jne::Error(JNE_ERR_INVALID_ARG, "filename missing for a selected model");
      names[m] = filenames[m];
    }
    std::unique_ptr<jne::Engine> own(ctx ? jne::Engine::borrow(ctx) : new jne::Engine(devs));
    const jne::Engine& gpu = *own;
    jne::SimulationStats st[5];
    jne::run_models_simulation(gpu, model_mask, dim, steps, num_runs, names, quiet != 0, st);
    if (stats)
      for (int m = 0; m < 5; ++m) { stats[3 * m] = st[m].completed_before; stats[3 * m + 1] = st[m].computed; stats[3 * m + 2] = st[m].total_in_file; }
    return JNE_OK;
  } catch (const jne::Error& e) {
    fprintf(stderr, "jne_run_models_simulation: %s\n", e.what());
    return e.status;
  } catch (const std::exception& e) {   // bad_alloc, system_error (thread creation) ...: nothing may cross the C ABI
    fprintf(stderr, "jne_run_models_simulation: %s\n", e.what());
    return JNE_ERR_INTERNAL;
  } catch (...) {
    fprintf(stderr, "jne_run_models_simulation: unknown exception\n");
    return JNE_ERR_INTERNAL;
  }
}

// C entry for run_model_simulation (used by the Python mirror and the tests).  stats: 3 x u64
// {completed_before, computed, total_in_file}.  Not in include/jne.h's hot-path section: orchestration helper.
// ctx != NULL: run on that context (device_ids ignored); ctx == NULL: a context over device_ids for this call.
int jne_run_model_simulation(jne_ctx* ctx, uint8_t model, uint32_t dim, uint32_t steps, uint64_t num_runs,
                             const char* filename, int quiet, const int* device_ids, int n_devices, uint64_t* stats) {
  try {
    std::vector<int> devs(device_ids, device_ids + (device_ids ? n_devices : 0));
    std::unique_ptr<jne::Engine> own(ctx ? jne::Engine::borrow(ctx) : new jne::Engine(devs));
    const jne::Engine& gpu = *own;
    const jne::SimulationStats st = jne::run_model_simulation(gpu, jne::Model(model), dim, steps, num_runs, filename, quiet != 0);
    if (stats) { stats[0] = st.completed_before; stats[1] = st.computed; stats[2] = st.total_in_file; }
    return JNE_OK;
  } catch (const jne::Error& e) {
    fprintf(stderr, "jne_run_model_simulation: %s\n", e.what());
    return e.status;
  } catch (const std::exception& e) {
    fprintf(stderr, "jne_run_model_simulation: %s\n", e.what());
    return JNE_ERR_INTERNAL;
  } catch (...) {
    fprintf(stderr, "jne_run_model_simulation: unknown exception\n");
    return JNE_ERR_INTERNAL;
  }
}

}  // extern "C"
