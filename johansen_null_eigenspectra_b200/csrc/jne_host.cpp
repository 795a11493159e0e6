// run_model_simulation: the reference's orchestration of one (model, dim, steps, num_runs) job
// (src/data_storage/parallel_compute.rs:150-232) over the GPU hot path and the batched .dat writer.
#include <algorithm>
#include <chrono>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>

#include "jne_host.hpp"

namespace jne {

namespace {

// JNE_DAT_TIMING=1: phase times of run_models_simulation on stderr (where does a job's wall time go)
struct PhaseTimer {
  bool on = [] { const char* e = getenv("JNE_DAT_TIMING"); return e && e[0] == '1'; }();
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void lap(const char* what) {
    if (!on) return;
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[jne_dat timing] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

// The row sink of one batch: every device's host thread hands its rows (pinned staging memory) to this function as
// they arrive, and the records of every selected model are encoded straight into that model's file.
struct BatchSink {
  jne_dat_batch* batch[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  uint32_t mask = 0, width = 0, off[5] = {0, 0, 0, 0, 0}, pm[5] = {0, 0, 0, 0, 0};
  int threads = 1;
  std::mutex mu;
  std::string err;
};

int batch_sink(void* user, uint64_t first, uint64_t count, const double* rows) {
  BatchSink* c = static_cast<BatchSink*>(user);
  int models[5], nm = 0;
  for (int m = 0; m < 5; ++m) if ((c->mask >> m) & 1u) models[nm++] = m;
  std::atomic<int> bad{0};
  auto fill = [&](int t, int T) {           // models t, t + T, ... of the mask
    for (int k = t; k < nm; k += T) {
      const int m = models[k];
      // a fill large enough is split further inside (small dims: tens of MB per file and sub-chunk)
      if (jne_dat_batch_fill(c->batch[m], first, count, rows + c->off[m], c->width, c->threads) != JNE_OK) {
        std::lock_guard<std::mutex> lk(c->mu);
        if (c->err.empty()) c->err = jne_dat_last_error();
        bad.store(1);
      }
    }
  };
  // One device thread cannot encode five files' records AND fault in their new pages as fast as its GPU produces rows
  // (~1 KB of memory traffic per seed, 3.2 M seeds/s at dim 12): the models of a sub-chunk are spread over a few
  // short-lived helpers.
  const int T = std::min(nm, (count * c->width * sizeof(double) >= ((size_t)1 << 20)) ? c->threads : 1);
  if (T <= 1) { fill(0, 1); return bad.load(); }
  std::thread th[4];
  int started = 0;
  try {
    for (int t = 1; t < T && t <= 4; ++t) { th[t - 1] = std::thread(fill, t, T); ++started; }
  } catch (...) { /* not started: filled below */ }
  fill(0, T);
  for (int t = started + 1; t < T; ++t) fill(t, T);
  for (int t = 0; t < started; ++t) th[t].join();
  return bad.load();
}

}  // namespace

SimulationStats run_model_simulation(const Engine& gpu, Model model, uint32_t dim, uint32_t steps, uint64_t num_runs,
                                     const std::string& filename, bool quiet) {
  // one model of the CLI's loop: the same job with a one-bit mask (resume scan, batches, trailer)
  std::string names[5];
  names[model.number] = filename;
  SimulationStats st[5];
  run_models_simulation(gpu, 1u << model.number, dim, steps, num_runs, names, quiet, st);
  return st[model.number];
}

void run_models_simulation(const Engine& gpu, uint32_t model_mask, uint32_t dim, uint32_t steps, uint64_t num_runs,
                           const std::string (&filenames)[5], bool quiet, SimulationStats (&stats)[5]) {
  if (model_mask == 0 || model_mask > 31u) throw Error(JNE_ERR_INVALID_ARG, "model_mask must select models 0..4");
  if (dim > 255) throw Error(JNE_ERR_INVALID_ARG, "dim must fit the u8 header field");
  PhaseTimer timer;
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
      if (!quiet && completed)   // a resumed file continues with THIS library's stream (JNE2), whoever wrote the head of it
        printf("Resuming %s: %llu of %llu runs present\n", filenames[m].c_str(), (unsigned long long)completed, (unsigned long long)num_runs);
      // need[s] |= bit for every seed whose bitmap bit is clear, eight seeds per bitmap byte
      const uint8_t bit = (uint8_t)(1u << m);
      const uint64_t full = num_runs & ~7ull;
      for (uint64_t s = 0; s < full; s += 8) {
        const uint8_t byte = bitmap[s >> 3];
        if (byte == 0xFF) continue;
        if (byte == 0) {
          uint64_t v;
          std::memcpy(&v, need.data() + s, 8);
          v |= 0x0101010101010101ull * bit;
          std::memcpy(need.data() + s, &v, 8);
        } else {
          for (int k = 0; k < 8; ++k) if (!((byte >> k) & 1u)) need[s + k] |= bit;
        }
      }
      for (uint64_t s = full; s < num_runs; ++s)
        if (!((bitmap[s >> 3] >> (s & 7)) & 1u)) need[s] |= bit;
    }
  }
  timer.lap("resume scan");
  // ---- writers for the files that lack something (a complete file is not touched, parallel_compute.rs:182-198) ----
  jne_dat_writer* w[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  uint64_t existing[5] = {0, 0, 0, 0, 0};
  // fresh jobs (and most resumes) have ONE set of lacking models for every seed: detect it in words
  bool uniform = num_runs > 0;
  {
    const uint64_t pat = 0x0101010101010101ull * (num_runs ? need[0] : 0);
    const uint64_t full = num_runs & ~7ull;
    for (uint64_t s = 0; s < full && uniform; s += 8) { uint64_t v; std::memcpy(&v, need.data() + s, 8); uniform = v == pat; }
    for (uint64_t s = full; s < num_runs && uniform; ++s) uniform = need[s] == need[0];
  }
  uint32_t lacking = 0;
  if (uniform) lacking = need[0];
  else for (uint64_t s = 0; s < num_runs; ++s) lacking |= need[s];
  auto abandon_all = [&]() { for (auto& x : w) if (x) { jne_dat_abandon(x); x = nullptr; } };
  for (int m = 0; m < 5; ++m) {
    if (!((lacking >> m) & 1u)) continue;
    const int rc = jne_dat_open(filenames[m].c_str(), (uint8_t)m, (uint8_t)dim, steps, &existing[m], &w[m]);
    if (rc != JNE_OK) { const std::string msg = jne_dat_last_error(); abandon_all(); throw Error(rc, msg); }
  }
  // seeds by the set of models that lack them, one pass (ascending within a group)
  std::vector<uint32_t> groups[32];
  if (uniform) {
    if (need[0]) {
      std::vector<uint32_t>& g = groups[need[0]];
      g.resize(num_runs);
      for (uint64_t s = 0; s < num_runs; ++s) g[s] = (uint32_t)(s + 1);
    }
    std::vector<uint8_t>().swap(need);
  } else {
    uint64_t counts[32] = {0};
    for (uint64_t s = 0; s < num_runs; ++s) ++counts[need[s]];
    for (uint32_t mask = 1; mask < 32; ++mask) groups[mask].reserve(counts[mask]);
    for (uint64_t s = 0; s < num_runs; ++s)
      if (need[s]) groups[need[s]].push_back((uint32_t)(s + 1));
    std::vector<uint8_t>().swap(need);
  }
  timer.lap("open + group seeds");
  // ---- one fused pass per distinct set of lacking models, in batches.  A batch reserves its bytes in every file (the
  //      record sizes follow from the seeds), then the GPUs stream their rows to the sink, where the device threads
  //      encode them into the files' pages as they arrive: no intermediate array and no separate writer phase (the
  //      double-buffered arrays + writer threads of round 1 ended at 5 GB/s on 8 GPUs although the box writes new
  //      pages at 20 GB/s, profiles/r2_io_bench_8gpu_box.txt).  A batch is also the unit an interrupted job loses. ----
  const int n_dev = std::max(1, jne_ctx_device_count(gpu.ctx()));
  const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
  static const int enc_env = [] { const char* e = getenv("JNE_DAT_ENCODERS"); return e ? atoi(e) : 0; }();
  const size_t batch_max = (size_t)n_dev << (dim <= 6 ? 23 : 21);   // 0.1-0.7 s of GPU work: the pipelines drain once per batch
  try {
    for (uint32_t mask = 1; mask < 32; ++mask) {
      const std::vector<uint32_t>& seeds = groups[mask];
      if (seeds.empty()) continue;
      BatchSink sink;
      sink.mask = mask;
      sink.width = (uint32_t)jne_multi_width(mask, dim);
      // a sub-chunk of a small dim is tens of MB per file: give the device thread's encoder a few helpers
      sink.threads = enc_env > 0 ? std::min(enc_env, 5) : std::max(1, std::min(4, hw / (2 * n_dev)));
      { uint32_t o = 0; for (int m = 0; m < 5; ++m) { sink.pm[m] = (uint32_t)Model((uint8_t)m).num_eigs(dim); sink.off[m] = o; if ((mask >> m) & 1u) o += sink.pm[m]; } }
      for (size_t a = 0; a < seeds.size(); a += batch_max) {
        const size_t nb = std::min(batch_max, seeds.size() - a);
        auto end_all = [&](int commit) {            // unmapping ~10^5 populated pages per file takes tens of ms: one thread per file
          int erc[5] = {0, 0, 0, 0, 0};
          std::string emsg[5];
          std::thread th[5];
          auto end_one = [&](int m) {
            erc[m] = jne_dat_batch_end(sink.batch[m], commit);
            sink.batch[m] = nullptr;
            if (erc[m] != JNE_OK) emsg[m] = jne_dat_last_error();
          };
          int last = -1;
          for (int m = 0; m < 5; ++m) if (sink.batch[m]) last = m;
          if (last < 0) return;
          try {
            for (int m = 0; m < last; ++m) if (sink.batch[m]) th[m] = std::thread(end_one, m);
          } catch (...) { /* ended below */ }
          end_one(last);
          for (int m = 0; m < 5; ++m) if (th[m].joinable()) th[m].join();
          for (int m = 0; m < 5; ++m) if (sink.batch[m]) end_one(m);     // a thread that never started
          for (int m = 0; m < 5; ++m) if (erc[m] != JNE_OK) throw Error(erc[m], emsg[m]);
        };
        {   // reserve the batch in every file (a pass over the seeds per file: one thread each)
          int brc[5] = {0, 0, 0, 0, 0};
          std::string bmsg[5];
          std::thread th[5];
          auto begin = [&](int m) {
            brc[m] = jne_dat_batch_begin(w[m], seeds.data() + a, nb, sink.pm[m], &sink.batch[m]);
            if (brc[m] != JNE_OK) bmsg[m] = jne_dat_last_error();
          };
          int last = -1;
          for (int m = 0; m < 5; ++m) if ((mask >> m) & 1u) last = m;
          try {
            for (int m = 0; m < 5; ++m) if (((mask >> m) & 1u) && m != last) th[m] = std::thread(begin, m);
          } catch (...) { for (auto& t : th) if (t.joinable()) t.join(); try { end_all(0); } catch (...) {} throw; }
          begin(last);
          for (int m = 0; m < 5; ++m) if (th[m].joinable()) th[m].join();
          for (int m = 0; m < 5; ++m) {
            if (((mask >> m) & 1u) && m != last && sink.batch[m] == nullptr && brc[m] == JNE_OK) begin(m);   // its thread never started
            if (brc[m] != JNE_OK) { try { end_all(0); } catch (...) {} throw Error(brc[m], bmsg[m]); }
          }
        }
        timer.lap("batch begin");
        const int rc = jne_eigs_batch_multi_stream(gpu.ctx(), mask, dim, steps, seeds.data() + a, nb, batch_sink, &sink);
        timer.lap("stream (GPU + encode)");
        if (rc != JNE_OK) {
          const std::string msg = !sink.err.empty() ? sink.err : std::string(jne_last_error(gpu.ctx()));
          try { end_all(0); } catch (...) {}
          throw Error(sink.err.empty() ? rc : JNE_ERR_IO, msg);
        }
        end_all(1);
        timer.lap("batch end");
        for (int m = 0; m < 5; ++m) if ((mask >> m) & 1u) stats[m].computed += nb;
        if (!quiet) printf("Simulation progress (models mask 0x%x): %llu/%llu seeds\n", mask,
                           (unsigned long long)(a + nb), (unsigned long long)seeds.size());
      }
    }
  } catch (...) {
    abandon_all();        // trailer-less, resumable files, like an interrupted reference run
    throw;
  }
  timer.lap("(loop exit)");
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
