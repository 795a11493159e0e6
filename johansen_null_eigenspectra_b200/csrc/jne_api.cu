// Host side of the C ABI declared in include/jne.h: context, device sharding, chunked
// double-buffered transfers, kernel dispatch.  No CPU fallback anywhere in this file: every
// numeric result is produced by the kernels of jne_kernels.cuh.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/jne.h"
#include "jne_kernels.cuh"
#include "jne_kernels_lane.cuh"

#define JNE_VERSION_STR "jne-b200 0.1.0 (sm_100a)"

namespace {

constexpr uint64_t kSlotSeeds = 1ull << 18;   // capacity of a staging slot, in runs (seed buffer) ...
constexpr int kSlots = 3;                     // staging slots per device, each with its own stream: while the host hands chunk i-3
                                              // to the caller (~6 GB/s), chunks i-2 and i-1 keep the device busy (two slots left
                                              // a bubble per chunk at the small dims: c2 80 % of the device-resident rate)
constexpr uint64_t kMaxWidth = 5 * 16;        // doubles per run: all five models at dim 15
constexpr uint64_t kSlotDoubles = (1ull << 15) * kMaxWidth;   // ... and in eigenvalues (21 MB): 2^15 runs at the widest row,
                                                               // 2^17 runs up to 20 doubles per run
constexpr uint64_t kChunkTarget = 1ull << 13; // runs per launch on the host-buffer path (D2H of chunk i overlaps the kernel of chunk i+1)

inline uint64_t slot_runs(uint32_t width) { return std::min<uint64_t>(kSlotSeeds, kSlotDoubles / std::max<uint32_t>(width, 1)); }

std::mutex g_init_err_mu;
std::string g_init_err;

struct Slot {
  uint32_t* d_seeds = nullptr;
  double* d_out = nullptr;
  uint32_t* h_seeds = nullptr;   // pinned
  double* h_out = nullptr;       // pinned
  cudaStream_t stream = nullptr;   // H2D -> kernel -> D2H of this slot; the two slots' streams overlap each other
  cudaEvent_t done = nullptr;      // D2H of this slot finished
  uint64_t n = 0, offset = 0;    // runs in flight and their position in the caller's arrays
  bool busy = false;
};

struct Device {
  int id = -1;
  cudaStream_t stream = nullptr;
  Slot slot[kSlots];
  unsigned int* d_err = nullptr;
  unsigned int* h_err = nullptr;  // pinned
  uint32_t* d_jtab = nullptr;   // 8 Jacobi step tables (ne = 2, 4, .., 16), kTabWords words each
  double* d_mom[kSlots] = {};              // lane family: per-run moments between the moments and the solve kernel,
  size_t mom_doubles[kSlots] = {};         // one buffer per concurrently used stream (capacity in doubles)
  int sm_count = 0;
  std::map<uint32_t, double*> aux_tabs;   // trend-weight tables of the AUX kernels, by steps (make_aux_table)
  double* d_scratch = nullptr;    // increments / pencil inputs, grown on demand
  size_t scratch_bytes = 0;
};

}  // namespace

struct jne_ctx {
  bool use_aux = true;        // trend moments through the MMA for dim <= 6 and 9..12 (env JNE_AUX=0: scalar FP64 sums)
  std::mutex aux_mu;
  bool use_lane = true;    // env JNE_LANE=0: tensor family for every dim (regression tooling)
  bool use_group = true;   // env JNE_GROUP=0: dims 9, 10 on the tensor family (regression tooling)
  bool lane_thread_solve = true;   // env JNE_LANE_SOLVE=warp: the lane family's moments through the warp-per-run epilogue
  std::vector<Device> devs;
  std::string err;
  std::mutex err_mu;
  std::atomic<uint64_t> launches{0};
  // async ticket
  std::thread worker;
  int64_t next_ticket = 1, pending_ticket = 0;
  int pending_status = JNE_OK;
};

namespace {

int fail(jne_ctx* ctx, int code, const std::string& msg) {
  if (ctx) { std::lock_guard<std::mutex> lk(ctx->err_mu); ctx->err = msg; }
  else { std::lock_guard<std::mutex> lk(g_init_err_mu); g_init_err = msg; }
  return code;
}

#define JNE_CUDA(ctx, call)                                                                         \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess) {                                                                        \
      char b_[512];                                                                                 \
      snprintf(b_, sizeof b_, "CUDA error %s (%s) at %s:%d: %s", cudaGetErrorName(e_),              \
               cudaGetErrorString(e_), __FILE__, __LINE__, #call);                                  \
      return fail(ctx, JNE_ERR_CUDA, b_);                                                           \
    }                                                                                               \
  } while (0)

// The library switches CUDA devices on the caller's thread (a context may span several); an FFI library must hand the
// thread back as it found it, or the host application's next allocation lands on the wrong GPU.
struct DeviceGuard {
  int prev = -1;
  DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); } }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

int validate(jne_ctx* ctx, uint8_t model, uint32_t dim, uint32_t steps) {
  if (model > 4) return fail(ctx, JNE_ERR_INVALID_ARG, "model must be 0..4");
  if (dim < 1 || dim > 255) return fail(ctx, JNE_ERR_INVALID_ARG, "dim must be 1..255");
  if (dim > JNE_MAX_DIM) return fail(ctx, JNE_ERR_UNSUPPORTED, "dim > JNE_MAX_DIM (15) is not supported");
  if (steps < 1 || steps > (1u << 30)) return fail(ctx, JNE_ERR_INVALID_ARG, "steps must be 1..2^30");
  if (model == 4 && steps < 2)
    return fail(ctx, JNE_ERR_INVALID_ARG, "model 4 needs steps >= 2 (singular Z Z', src/johansen_statistics.rs:191-192)");
  return JNE_OK;
}

// sum_{i=a}^{b-1} of (2i+1-T)  and of  3(2i+1-T)^2 - (T^2-1), exactly, in 128-bit integers.
void segment_weights(uint64_t a, uint64_t b, uint64_t T, double* w1, double* w2) {
  const __int128 A = a, B = b, TT = T, n = B - A;
  *w1 = (double)(n * (A + B - TT));
  auto sum1 = [](__int128 m) { return m * (m - 1) / 2; };                 // sum_{i<m} i
  auto sum2 = [](__int128 m) { return (m - 1) * m * (2 * m - 1) / 6; };   // sum_{i<m} i^2
  const __int128 S1 = sum1(B) - sum1(A), S2 = sum2(B) - sum2(A);
  const __int128 q = 1 - TT;
  const __int128 m2 = 4 * S2 + 4 * q * S1 + q * q * n;                    // sum (2i+1-T)^2
  *w2 = (double)(3 * m2 - (TT * TT - 1) * n);
}

// Trend weights fed to the MMA by the AUX kernels (jne_step): table[j][m][k] for local step j of segment k
// (global step i = a_k + j, segment [a_k, e_k)), with w1_i = 2i + 1 - T and w2_i = 3 w1_i^2 - (T^2 - 1):
//   m = 0: sum_{i < i' < e_k} w1_i'   (so that sum_i w1_i c_i = sum_s dB_s * weight, c = segment-local path before step i)
//   m = 1: sum_{i < i' < e_k} w2_i'
//   m = 2: w2_i        m = 3: w1_i
// Integer-valued, evaluated in 128-bit integers and rounded once; steps outside the segment carry 0.
constexpr uint32_t kAuxMaxSteps = 1u << 22;   // 128 MiB of table; longer runs use the scalar-sum kernels
// Length of the four time segments of the tensor family.  Long horizons: whole generator epochs (jne_rng.cuh), so that
// the lanes of a warp -- one segment each -- start new substreams in the same block of the time loop (a warp-uniform
// branch, 3 Philox calls per 16 blocks).  Short horizons: whole 8-step blocks; the segments then start inside an
// epoch, every lane skips to its place in its substreams once and re-keys in its own phase (the branch diverges: up
// to four passes through the key generation per 16 blocks, ~5 % of the loop) -- cheaper than leaving segments empty
// (T = 200: 89 % of the lane-steps useful instead of 39 %; T = 2 049: 98 % instead of 80 %).  The rule picks the
// better of the two estimates: T = 1 000, 5 000, 10 000 and 100 000 run epoch-aligned.
uint32_t seg_len_for(uint32_t steps) {
  const uint32_t by_block = 8u * ((steps + 31u) / 32u);
#ifdef JNE_EXP_SEGLEN8   // experiment only: segments of whole 8-step blocks at every horizon
  return by_block;
#endif
  const uint32_t by_epoch =
      (uint32_t)(JNE_EPOCH_STEPS * (((uint64_t)steps + 4u * JNE_EPOCH_STEPS - 1u) / (4u * JNE_EPOCH_STEPS)));
  return 0.95 * (double)by_epoch > (double)by_block ? by_block : by_epoch;
}
std::vector<double> make_aux_table(uint32_t steps) {
  const uint32_t seg_len = seg_len_for(steps);
  std::vector<double> tab((size_t)seg_len * 16, 0.0);
  const __int128 T = steps;
  for (int k = 0; k < 4; ++k) {
    const uint64_t a = std::min<uint64_t>((uint64_t)k * seg_len, steps);
    const uint64_t e = std::min<uint64_t>(a + seg_len, steps);
    __int128 u1 = 0, u2 = 0;
    for (uint64_t i = e; i-- > a;) {
      const __int128 w1 = 2 * (__int128)i + 1 - T, w2 = 3 * w1 * w1 - (T * T - 1);
      double* row = tab.data() + (size_t)(i - a) * 16;
      row[0 * 4 + k] = (double)u1;
      row[1 * 4 + k] = (double)u2;
      row[2 * 4 + k] = (double)w2;
      row[3 * 4 + k] = (double)w1;
      u1 += w1;
      u2 += w2;
    }
  }
  return tab;
}

uint32_t mask_width(uint32_t mask, uint32_t dim) {
  uint32_t w = 0;
  for (int m = 0; m < 5; ++m)
    if ((mask >> m) & 1u) w += (m == 1 || m == 3) ? dim + 1 : dim;
  return w;
}

// Jacobi step tables (jne_warp_jacobi).  Round-robin (circle method) pairing of ne players: in step `step`
// slot 0 pairs player ne-1 with `step`, slot l pairs (step + l) with (step - l) modulo ne-1; each pair is (min, max).
// Per step: npairs rotation words o(p,p) | o(q,q) << 8 | o(p,q) << 16, then one word per 2x2 block (P1 <= P2) of
// pair slots, o(p1,p2) | o(p1,q2) << 8 | o(q1,p2) << 16 | o(q1,q2) << 24, where o(i,j) is the offset of the
// element in the packed upper triangle of the ne x ne matrix.  One table per ne = 2, 4, .., 16.
constexpr int kTabWords = 15 * (8 + 36);
void make_jacobi_tables(uint32_t* tables /* 8 x kTabWords */) {
  std::memset(tables, 0, 8 * kTabWords * sizeof(uint32_t));
  for (int ne = 2; ne <= 16; ne += 2) {
    uint32_t* t = tables + (ne / 2 - 1) * kTabWords;
    const int np = ne / 2, m = ne - 1, nblk = np * (np + 1) / 2;
    auto tri = [ne](int i, int j) -> uint32_t {
      if (i > j) std::swap(i, j);
      return (uint32_t)(i * ne - i * (i - 1) / 2 + (j - i));
    };
    for (int step = 0; step < m; ++step) {
      int pp[8], qq[8];
      for (int l = 0; l < np; ++l) {
        const int a = (l == 0) ? m : (step + l) % m, b = (step + m - l) % m;
        pp[l] = std::min(a, b);
        qq[l] = std::max(a, b);
      }
      uint32_t* row = t + step * (np + nblk);
      for (int l = 0; l < np; ++l) row[l] = tri(pp[l], pp[l]) | tri(qq[l], qq[l]) << 8 | tri(pp[l], qq[l]) << 16;
      int blk = 0;
      for (int P1 = 0; P1 < np; ++P1)
        for (int P2 = P1; P2 < np; ++P2)
          row[np + blk++] = tri(pp[P1], pp[P2]) | tri(pp[P1], qq[P2]) << 8 | tri(qq[P1], pp[P2]) << 16 | tri(qq[P1], qq[P2]) << 24;
    }
  }
}

const uint32_t* jtab_for(const Device& dv, uint32_t d) { return dv.d_jtab + (((d + 1) / 2) - 1) * kTabWords; }

JneRunParams make_params_mask(uint32_t mask, uint32_t dim, uint32_t steps, bool from_increments) {
  JneRunParams p{};
  p.dim = dim; p.steps = steps;
  p.model_mask = mask;
  p.out_stride = mask_width(mask, dim);
  p.model = 0;
  for (int m = 0; m < 5; ++m) if ((mask >> m) & 1u) p.model = m;   // highest selected model
  p.p = p.out_stride;   // doubles per run (== eigenvalues per run for a single model)
  p.seg_len = seg_len_for(steps);
  p.T = (double)steps;
  p.factor = from_increments ? (double)steps : 1.0;
  uint64_t full = p.seg_len / 8;
  for (int k = 0; k < 4; ++k) {
    const uint64_t a = std::min<uint64_t>((uint64_t)k * p.seg_len, steps);
    const uint64_t b = std::min<uint64_t>(a + p.seg_len, steps);
    p.seg_n[k] = (double)(b - a);
    full = std::min<uint64_t>(full, (b - a) / 8);
    segment_weights(a, b, steps, &p.seg_w1[k], &p.seg_w2[k]);
  }
  p.full_blocks = (uint32_t)full;
  return p;
}

JneRunParams make_params(uint8_t model, uint32_t dim, uint32_t steps, bool from_increments) {
  return make_params_mask(1u << model, dim, steps, from_increments);
}

template <int DP, bool MULTI> constexpr size_t cta_smem() {
  return (size_t)JNE_WARPS_PER_CTA * JneEpi<DP, MULTI ? 5 : 1>::WARP_SMEM * sizeof(double);
}
constexpr size_t pencil_smem() { return (size_t)JNE_WARPS_PER_CTA * JneEpi<16, 1>::END * sizeof(double); }

template <int DP, int DET, bool RNG, bool MULTI, bool AUXT = false>
cudaError_t launch_one(const uint32_t* d_seeds, const double* d_dB, uint64_t n, const JneRunParams& prm,
                       double* d_out, unsigned int* d_err, double* d_dbg, cudaStream_t st) {
  // short horizons run the instance whose lanes re-key their substreams in their own phase (seg_len_for)
  const bool unaligned = RNG && (prm.seg_len % JNE_EPOCH_STEPS) != 0u;
  auto kern = unaligned ? jne_run_kernel<DP, DET, RNG, MULTI, AUXT, RNG> : jne_run_kernel<DP, DET, RNG, MULTI, AUXT, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cta_smem<DP, MULTI>());
  if (e != cudaSuccess) return e;
  const unsigned grid = (unsigned)((n + JNE_WARPS_PER_CTA - 1) / JNE_WARPS_PER_CTA);
  kern<<<grid, 32 * JNE_WARPS_PER_CTA, cta_smem<DP, MULTI>(), st>>>(d_seeds, d_dB, n, prm, d_out, d_err, d_dbg);
  return cudaGetLastError();
}

template <int DP, bool RNG>
cudaError_t launch_det(int det, const uint32_t* s, const double* b, uint64_t n, const JneRunParams& prm, double* o,
                       unsigned int* e, double* dbg, cudaStream_t st) {
  const bool multi = (prm.model_mask & (prm.model_mask - 1u)) != 0;   // more than one model: superset accumulation
  if constexpr (JneGeo<DP>::B == 4 || DP == 8) {
    if (prm.aux_tab != nullptr && (multi || det >= 1)) {   // trend moments through the MMA
      if (multi) return launch_one<DP, 2, RNG, true, true>(s, b, n, prm, o, e, dbg, st);
      if (det == 1) return launch_one<DP, 1, RNG, false, true>(s, b, n, prm, o, e, dbg, st);
      return launch_one<DP, 2, RNG, false, true>(s, b, n, prm, o, e, dbg, st);
    }
  }
  if (multi) return launch_one<DP, 2, RNG, true>(s, b, n, prm, o, e, dbg, st);
  switch (det) {
    case 0: return launch_one<DP, 0, RNG, false>(s, b, n, prm, o, e, dbg, st);
    case 1: return launch_one<DP, 1, RNG, false>(s, b, n, prm, o, e, dbg, st);
    default: return launch_one<DP, 2, RNG, false>(s, b, n, prm, o, e, dbg, st);
  }
}


// Runs in one full wave of the kernel launch_run would pick for prm (resident CTAs per SM x SMs x runs per CTA):
// the host-buffer path sizes its chunks in whole waves so that a chunk does not end on a mostly empty wave.
template <int DP, int DET, bool MULTI, bool AUXT = false>
uint64_t wave_one(const Device& dv) {
  auto kern = jne_run_kernel<DP, DET, true, MULTI, AUXT>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cta_smem<DP, MULTI>()) != cudaSuccess) return 0;
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 32 * JNE_WARPS_PER_CTA, cta_smem<DP, MULTI>()) != cudaSuccess) return 0;
  return (uint64_t)nb * dv.sm_count * JNE_WARPS_PER_CTA;
}
template <int DP>
uint64_t wave_det(const Device& dv, const JneRunParams& prm, bool aux) {
  const int det = (prm.model <= 1) ? 0 : (prm.model <= 3 ? 1 : 2);
  const bool multi = (prm.model_mask & (prm.model_mask - 1u)) != 0;
  if constexpr (JneGeo<DP>::B == 4 || DP == 8) {
    if (aux && det >= 1)
      return multi ? wave_one<DP, 2, true, true>(dv) : det == 1 ? wave_one<DP, 1, false, true>(dv) : wave_one<DP, 2, false, true>(dv);
  }
  if (multi) return wave_one<DP, 2, true>(dv);
  return det == 0 ? wave_one<DP, 0, false>(dv) : det == 1 ? wave_one<DP, 1, false>(dv) : wave_one<DP, 2, false>(dv);
}
uint64_t wave_runs(const jne_ctx* ctx, const Device& dv, const JneRunParams& prm);

// The AUX kernels serve dim <= 4 and 9..12 (the MMA tile layouts with four F rows and four dB rows in one group) and
// dim 5, 6 (two padding rows in the 8-row F group) when a selected model has a trend row and the weight table fits.
bool aux_wanted(const jne_ctx* ctx, const JneRunParams& prm) {
  return ctx->use_aux && prm.model >= 2 && prm.steps <= kAuxMaxSteps &&
         (prm.dim <= 6 || (prm.dim >= 9 && prm.dim <= 12));
}
// Device copy of make_aux_table(steps), built on first use; nullptr when it cannot be had (callers fall back).
const double* aux_table_for(jne_ctx* ctx, Device& dv, uint32_t steps) {
  std::lock_guard<std::mutex> lk(ctx->aux_mu);
  auto it = dv.aux_tabs.find(steps);
  if (it != dv.aux_tabs.end()) return it->second;
  if (dv.aux_tabs.size() >= 8) {   // bounded cache: drop everything once eight horizons are resident
    cudaDeviceSynchronize();
    for (auto& kv : dv.aux_tabs) cudaFree(kv.second);
    dv.aux_tabs.clear();
  }
  const std::vector<double> tab = make_aux_table(steps);
  double* d = nullptr;
  if (cudaMalloc(&d, tab.size() * sizeof(double)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (cudaMemcpy(d, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaGetLastError(); cudaFree(d); return nullptr;
  }
  // cudaMemcpy from pageable memory may return while the DMA is still in flight on the legacy default stream, and the
  // kernels run on NON-BLOCKING streams, which do not wait for it: without this the first launch for a new horizon
  // could read a half-written table (seen as non-finite runs in a freshly created multi-device context).
  if (cudaDeviceSynchronize() != cudaSuccess) { cudaGetLastError(); cudaFree(d); return nullptr; }
  dv.aux_tabs[steps] = d;
  return d;
}

// ---- lane family (jne_kernels_lane.cuh): dim <= 6, one thread per run + one warp per run for the solve ----
constexpr uint64_t kMomChunk = 1ull << 18;   // upper bound on the runs per moments / solve pair (195 MB of moments at dim 6)

// dims served by the group kernels (L lanes per run).  Measured (profiles/r2_variants_group.txt, fused pass, T 10 000):
// dim 9 with 3 lanes x 3 rows 3.83 M seeds/s against 3.28 M on the tensor family; dim 10 with 5 x 2 3.13 M against
// 3.27 M (two idle lanes per warp, the replicated path costs 10 of 50 FP64 operations) -- so only dim 9 is routed here.
#define JNE_GROUP_L9 3
#ifdef JNE_EXP_GROUP_78   // experiment: dims 7, 8 as 2 lanes x 4 rows, dim 10 as 5 x 2
bool group_dim(uint32_t dim) { return dim >= 7 && dim <= 10; }
#else
bool group_dim(uint32_t dim) { return dim == 9; }
#endif
bool lane_wanted(const jne_ctx* ctx, const JneRunParams& prm) {
  return ctx->use_lane && (prm.dim <= JNE_LANE_MAX_DIM || (ctx->use_group && group_dim(prm.dim)));
}
int lane_det(const JneRunParams& prm) {
  const bool multi = (prm.model_mask & (prm.model_mask - 1u)) != 0;
  return multi ? 2 : ((prm.model <= 1) ? 0 : (prm.model <= 3 ? 1 : 2));
}

template <int D, bool RNG>
cudaError_t launch_lane_moments(int det, const uint32_t* s, const double* b, uint64_t m, uint32_t steps, double* mom, cudaStream_t st) {
  constexpr int TH = JneLaneGeo<D>::THREADS;
  const unsigned grid = (unsigned)((m + TH - 1) / TH);
  if constexpr (RNG) {
    switch (det) {
      case 0: jne_lane_moments_kernel<D, 0, true><<<grid, TH, 0, st>>>(s, b, m, steps, mom); break;
      case 1: jne_lane_moments_kernel<D, 1, true><<<grid, TH, 0, st>>>(s, b, m, steps, mom); break;
      default: jne_lane_moments_kernel<D, 2, true><<<grid, TH, 0, st>>>(s, b, m, steps, mom); break;
    }
  } else {   // caller increments (parity gate 1): one instantiation, the superset of the trend sums
    jne_lane_moments_kernel<D, 2, false><<<grid, TH, 0, st>>>(s, b, m, steps, mom);
  }
  return cudaGetLastError();
}
template <int D, int L, bool RNG>
cudaError_t launch_group_moments(int det, const uint32_t* s, const double* b, uint64_t m, uint32_t steps, double* mom, cudaStream_t st) {
  using Q = JneGroupGeo<D, L>;
  const unsigned grid = (unsigned)((m + Q::RUNS_PER_CTA - 1) / Q::RUNS_PER_CTA);
  if constexpr (RNG) {
    switch (det) {
      case 0: jne_group_moments_kernel<D, L, 0, true><<<grid, Q::THREADS, 0, st>>>(s, b, m, steps, mom); break;
      case 1: jne_group_moments_kernel<D, L, 1, true><<<grid, Q::THREADS, 0, st>>>(s, b, m, steps, mom); break;
      default: jne_group_moments_kernel<D, L, 2, true><<<grid, Q::THREADS, 0, st>>>(s, b, m, steps, mom); break;
    }
  } else {
    jne_group_moments_kernel<D, L, 2, false><<<grid, Q::THREADS, 0, st>>>(s, b, m, steps, mom);
  }
  return cudaGetLastError();
}
template <int D, int L> uint64_t group_wave_one(const Device& dv, int det) {
  using Q = JneGroupGeo<D, L>;
  int nb = 0;
  cudaError_t rc = det == 0 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, jne_group_moments_kernel<D, L, 0, true>, Q::THREADS, 0)
                 : det == 1 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, jne_group_moments_kernel<D, L, 1, true>, Q::THREADS, 0)
                            : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, jne_group_moments_kernel<D, L, 2, true>, Q::THREADS, 0);
  if (rc != cudaSuccess) { cudaGetLastError(); return 0; }
  return (uint64_t)nb * dv.sm_count * Q::RUNS_PER_CTA;
}

template <bool RNG>
cudaError_t launch_lane_moments_dim(uint32_t dim, int det, const uint32_t* s, const double* b, uint64_t m, uint32_t steps,
                                    double* mom, cudaStream_t st) {
  switch (dim) {
    case 9: return launch_group_moments<9, JNE_GROUP_L9, RNG>(det, s, b, m, steps, mom, st);
#ifdef JNE_EXP_GROUP_78
    case 7: return launch_group_moments<7, 2, RNG>(det, s, b, m, steps, mom, st);
    case 8: return launch_group_moments<8, 2, RNG>(det, s, b, m, steps, mom, st);
    case 10: return launch_group_moments<10, 5, RNG>(det, s, b, m, steps, mom, st);
#endif
    case 1: return launch_lane_moments<1, RNG>(det, s, b, m, steps, mom, st);
    case 2: return launch_lane_moments<2, RNG>(det, s, b, m, steps, mom, st);
    case 3: return launch_lane_moments<3, RNG>(det, s, b, m, steps, mom, st);
    case 4: return launch_lane_moments<4, RNG>(det, s, b, m, steps, mom, st);
    case 5: return launch_lane_moments<5, RNG>(det, s, b, m, steps, mom, st);
    default: return launch_lane_moments<6, RNG>(det, s, b, m, steps, mom, st);
  }
}
template <int DP, bool MULTI>
cudaError_t launch_lane_solve(const double* mom, uint64_t m, const JneRunParams& prm, double* o, unsigned int* e, double* dbg,
                              cudaStream_t st) {
  auto kern = jne_lane_solve_kernel<DP, MULTI>;
  cudaError_t rc = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)jne_lane_solve_smem<DP, MULTI>());
  if (rc != cudaSuccess) return rc;
  const unsigned grid = (unsigned)((m + JNE_WARPS_PER_CTA - 1) / JNE_WARPS_PER_CTA);
  kern<<<grid, 32 * JNE_WARPS_PER_CTA, jne_lane_solve_smem<DP, MULTI>(), st>>>(mom, m, prm, o, e, dbg);
  return cudaGetLastError();
}

template <int D>
cudaError_t launch_lane_tsolve(const double* mom, uint64_t m, const JneRunParams& prm, double* o, unsigned int* e, cudaStream_t st) {
  jne_lane_tsolve_kernel<D><<<(unsigned)((m + 127) / 128), 128, 0, st>>>(mom, m, prm.model_mask, prm.T, prm.factor, prm.out_stride, o, e);
  return cudaGetLastError();
}
cudaError_t launch_lane_tsolve_dim(const double* mom, uint64_t m, const JneRunParams& prm, double* o, unsigned int* e, cudaStream_t st) {
  switch (prm.dim) {
    case 1: return launch_lane_tsolve<1>(mom, m, prm, o, e, st);
    case 2: return launch_lane_tsolve<2>(mom, m, prm, o, e, st);
    case 3: return launch_lane_tsolve<3>(mom, m, prm, o, e, st);
    case 4: return launch_lane_tsolve<4>(mom, m, prm, o, e, st);
    case 5: return launch_lane_tsolve<5>(mom, m, prm, o, e, st);
    default: return launch_lane_tsolve<6>(mom, m, prm, o, e, st);
  }
}

// Runs resident at once in the lane moments kernel (0 if the query fails).
template <int D> uint64_t lane_wave_one(const Device& dv, int det) {
  int nb = 0;
  constexpr int TH = JneLaneGeo<D>::THREADS;
  cudaError_t rc = det == 0 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, jne_lane_moments_kernel<D, 0, true>, TH, 0)
                 : det == 1 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, jne_lane_moments_kernel<D, 1, true>, TH, 0)
                            : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, jne_lane_moments_kernel<D, 2, true>, TH, 0);
  if (rc != cudaSuccess) { cudaGetLastError(); return 0; }
  return (uint64_t)nb * dv.sm_count * TH;
}
uint64_t lane_wave(const Device& dv, const JneRunParams& prm) {
  const int det = lane_det(prm);
  switch (prm.dim) {
    case 9: return group_wave_one<9, JNE_GROUP_L9>(dv, det);
#ifdef JNE_EXP_GROUP_78
    case 7: return group_wave_one<7, 2>(dv, det);
    case 8: return group_wave_one<8, 2>(dv, det);
    case 10: return group_wave_one<10, 5>(dv, det);
#endif
    case 1: return lane_wave_one<1>(dv, det);
    case 2: return lane_wave_one<2>(dv, det);
    case 3: return lane_wave_one<3>(dv, det);
    case 4: return lane_wave_one<4>(dv, det);
    case 5: return lane_wave_one<5>(dv, det);
    default: return lane_wave_one<6>(dv, det);
  }
}

// moments kernel + solve kernel per chunk of runs (whole waves of the moments kernel), both on stream st
template <bool RNG>
cudaError_t launch_lane(jne_ctx* ctx, Device& dv, const uint32_t* s, const double* b, uint64_t n, const JneRunParams& prm,
                        double* o, unsigned int* e, double* dbg, cudaStream_t st, int ms) {
  const size_t per_run = (size_t)jne_lane_mom_size((int)prm.dim);
  uint64_t chunk = kMomChunk;
  const uint64_t wave = lane_wave(dv, prm);
  if (wave > 0 && wave <= kMomChunk) chunk = (kMomChunk / wave) * wave;
  const size_t need = (size_t)std::min<uint64_t>(n, chunk) * per_run;
  if (dv.mom_doubles[ms] < need) {
    if (dv.d_mom[ms]) {
      cudaError_t f = cudaFree(dv.d_mom[ms]);   // (synchronises the device: earlier launches reading it are done)
      dv.d_mom[ms] = nullptr; dv.mom_doubles[ms] = 0;
      if (f != cudaSuccess) return f;
    }
    // a full chunk asks for the largest layout of the family at once (dim 9: 180 doubles per run, 377 MB): a sweep over
    // the dims would otherwise free and re-allocate the buffer at every dim, and each such pair synchronises the device
    const size_t cap = n >= chunk ? std::max(need, (size_t)kMomChunk * (size_t)jne_lane_mom_size(9)) : need;
    size_t got = cap;
    cudaError_t a = cudaMalloc(&dv.d_mom[ms], cap * sizeof(double));
    if (a != cudaSuccess && cap > need) { cudaGetLastError(); got = need; a = cudaMalloc(&dv.d_mom[ms], need * sizeof(double)); }
    if (a != cudaSuccess) return a;
    dv.mom_doubles[ms] = got;
  }
  double* d_mom = dv.d_mom[ms];
  const bool multi = (prm.model_mask & (prm.model_mask - 1u)) != 0;
  const int det = lane_det(prm);
  for (uint64_t off = 0; off < n; off += chunk) {
    const uint64_t m = std::min(chunk, n - off);
    const uint32_t* sp = s ? s + off : nullptr;
    const double* bp = b ? b + off * (uint64_t)prm.dim * prm.steps : nullptr;
    cudaError_t rc = launch_lane_moments_dim<RNG>(prm.dim, det, sp, bp, m, prm.steps, d_mom, st);
    if (rc != cudaSuccess) return rc;
    double* op = o + off * prm.out_stride;
    double* dp = dbg ? dbg + off * 512 : nullptr;
    // solve: one thread per run; the warp-per-run epilogue of the tensor family only when the caller wants the
    // assembled matrices back (jne_eigs_batch_debug) or asks for it (JNE_LANE_SOLVE=warp, regression tooling)
    if (prm.dim > 6) rc = multi ? launch_lane_solve<12, true>(d_mom, m, prm, op, e, dp, st) : launch_lane_solve<12, false>(d_mom, m, prm, op, e, dp, st);
    else if (dp == nullptr && ctx->lane_thread_solve) rc = launch_lane_tsolve_dim(d_mom, m, prm, op, e, st);
    else if (prm.dim <= 4) rc = multi ? launch_lane_solve<4, true>(d_mom, m, prm, op, e, dp, st) : launch_lane_solve<4, false>(d_mom, m, prm, op, e, dp, st);
    else rc = multi ? launch_lane_solve<8, true>(d_mom, m, prm, op, e, dp, st) : launch_lane_solve<8, false>(d_mom, m, prm, op, e, dp, st);
    if (rc != cudaSuccess) return rc;
    ctx->launches.fetch_add(1);   // the solve kernel; the caller counts the moments kernel
  }
  return cudaSuccess;
}


template <bool RNG>
cudaError_t launch_run(jne_ctx* ctx, Device& dv, const uint32_t* s, const double* b, uint64_t n, const JneRunParams& prm,
                       double* o, unsigned int* e, double* dbg, cudaStream_t st, int mom_slot = 0) {
  if (lane_wanted(ctx, prm)) return launch_lane<RNG>(ctx, dv, s, b, n, prm, o, e, dbg, st, mom_slot);
  const int det = (prm.model <= 1) ? 0 : (prm.model <= 3 ? 1 : 2);
  JneRunParams q = prm;
  q.aux_tab = aux_wanted(ctx, prm) ? aux_table_for(ctx, dv, prm.steps) : nullptr;
  if (prm.dim <= 4) return launch_det<4, RNG>(det, s, b, n, q, o, e, dbg, st);
  if (prm.dim <= 8) return launch_det<8, RNG>(det, s, b, n, q, o, e, dbg, st);
  if (prm.dim <= 12) return launch_det<12, RNG>(det, s, b, n, q, o, e, dbg, st);
  return launch_det<16, RNG>(det, s, b, n, q, o, e, dbg, st);
}

uint64_t wave_runs(const jne_ctx* ctx, const Device& dv, const JneRunParams& prm) {
  if (lane_wanted(ctx, prm)) return lane_wave(dv, prm);
  const bool aux = aux_wanted(ctx, prm);
  uint64_t w = prm.dim <= 4 ? wave_det<4>(dv, prm, aux) : prm.dim <= 8 ? wave_det<8>(dv, prm, aux)
             : prm.dim <= 12 ? wave_det<12>(dv, prm, aux) : wave_det<16>(dv, prm, aux);
  cudaGetLastError();
  return w;
}

int ensure_scratch(jne_ctx* ctx, Device& dv, size_t bytes) {
  if (dv.scratch_bytes >= bytes) return JNE_OK;
  if (dv.d_scratch) { cudaFree(dv.d_scratch); dv.d_scratch = nullptr; dv.scratch_bytes = 0; }
  JNE_CUDA(ctx, cudaMalloc(&dv.d_scratch, bytes));
  dv.scratch_bytes = bytes;
  return JNE_OK;
}

// Hand-over copy pinned staging -> caller's array.  One host thread moves ~4 GB/s here, and the small dims produce
// their rows faster than that (c2: 26 M seeds/s x 216 B); large chunks are split over a few short-lived threads.
// (Page-locking the caller's array for a direct DMA was measured and lost: cudaHostRegister + Unregister cost more than
// the copy they save -- dim 12: 99.7 % -> 81 % of the device-resident rate, profiles/r2_exp_e2e_small_dims.txt.)
void handover_copy(double* dst, const double* src, size_t n_doubles) {
  const size_t bytes = n_doubles * sizeof(double);
  if (bytes < ((size_t)4 << 20)) { std::memcpy(dst, src, bytes); return; }
  constexpr int kThreads = 4;
  const size_t per = (n_doubles + kThreads - 1) / kThreads;
  std::thread th[kThreads - 1];
  int started = 0;
  try {
    for (int t = 1; t < kThreads; ++t) {
      const size_t a = std::min(n_doubles, t * per), b = std::min(n_doubles, a + per);
      th[t - 1] = std::thread([=]() { std::memcpy(dst + a, src + a, (b - a) * sizeof(double)); });
      ++started;
    }
  } catch (...) {   // could not start a helper: copy the rest here
    const size_t a = std::min(n_doubles, (size_t)(started + 1) * per);
    std::memcpy(dst + a, src + a, (n_doubles - a) * sizeof(double));
  }
  std::memcpy(dst, src, std::min(n_doubles, per) * sizeof(double));
  for (int t = 0; t < started; ++t) th[t].join();
}

// Where a share's rows go: into the caller's array, or to the caller's sink (jne_eigs_batch_multi_stream).
struct RowSink {
  double* out = nullptr;            // array of this share (row 0 = the share's first seed)
  jne_rows_sink fn = nullptr;
  void* user = nullptr;
  uint64_t base = 0;                // index of the share's first seed in the caller's seed list
};

// Drain one slot: wait for its D2H, hand the rows over.
int drain(jne_ctx* ctx, Slot& s, uint32_t p, const RowSink& sink) {
  if (!s.busy) return JNE_OK;
  JNE_CUDA(ctx, cudaEventSynchronize(s.done));
  s.busy = false;
  if (sink.fn) {
    if (sink.fn(sink.user, sink.base + s.offset, s.n, s.h_out) != 0) return fail(ctx, JNE_ERR_IO, "the row sink reported a failure");
  } else {
    handover_copy(sink.out + s.offset * p, s.h_out, s.n * p);
  }
  return JNE_OK;
}

// One device's share of a host-buffer batch (called on its own host thread).
int run_share(jne_ctx* ctx, Device& dv, const JneRunParams& prm_in, const uint32_t* seeds, uint64_t n, const RowSink& out,
              std::string* err_out) {
  auto body = [&]() -> int {
    JNE_CUDA(ctx, cudaSetDevice(dv.id));
    JneRunParams prm = prm_in;
    prm.jtab = jtab_for(dv, prm.dim);
    *dv.h_err = 0;
    for (auto& s : dv.slot) s.busy = false;   // nothing of an earlier (possibly failed) call is ever drained into this one
    JNE_CUDA(ctx, cudaMemsetAsync(dv.d_err, 0, sizeof(unsigned int), dv.stream));
    JNE_CUDA(ctx, cudaStreamSynchronize(dv.stream));
    uint64_t done = 0;
    int which = 0;
    // Chunks are whole waves (a launch does not end on a mostly idle wave) and small (what stays exposed at the end
    // of a call is the last chunk's D2H and the hand-over of the last two chunks to the caller's array).
    const uint64_t cap = slot_runs(prm.out_stride);
    uint64_t chunk = std::min(kChunkTarget, cap);
    const uint64_t wave = wave_runs(ctx, dv, prm);
    if (wave > 0 && wave <= cap) chunk = std::max<uint64_t>(1, kChunkTarget / wave) * wave;
    else if (wave > cap) chunk = cap;          // a wave does not fit the slot: fill the slot rather than 8 k runs
    while (done < n) {
      Slot& s = dv.slot[which];
      int rc = drain(ctx, s, prm.p, out);
      if (rc) return rc;
      const uint64_t m = std::min<uint64_t>(chunk, n - done);
      std::memcpy(s.h_seeds, seeds + done, m * sizeof(uint32_t));
      // each slot has its own stream: the copies of one chunk and the tail wave of its kernel overlap the
      // kernel of the other slot's chunk
      JNE_CUDA(ctx, cudaMemcpyAsync(s.d_seeds, s.h_seeds, m * sizeof(uint32_t), cudaMemcpyHostToDevice, s.stream));
      JNE_CUDA(ctx, launch_run<true>(ctx, dv, s.d_seeds, nullptr, m, prm, s.d_out, dv.d_err, nullptr, s.stream, which));
      ctx->launches.fetch_add(1);
      JNE_CUDA(ctx, cudaMemcpyAsync(s.h_out, s.d_out, m * prm.p * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
      JNE_CUDA(ctx, cudaEventRecord(s.done, s.stream));
      s.n = m; s.offset = done; s.busy = true;
      done += m;
      which = (which + 1) % kSlots;
    }
    for (int i = 0; i < kSlots; ++i) {      // oldest first
      int rc = drain(ctx, dv.slot[(which + i) % kSlots], prm.p, out);
      if (rc) return rc;
    }
    JNE_CUDA(ctx, cudaMemcpyAsync(dv.h_err, dv.d_err, sizeof(unsigned int), cudaMemcpyDeviceToHost, dv.stream));
    JNE_CUDA(ctx, cudaStreamSynchronize(dv.stream));
    if (*dv.h_err) {
      char b[160];
      snprintf(b, sizeof b, "%u run(s) produced non-finite eigenvalues (the reference panics at src/johansen_statistics.rs:45)", *dv.h_err);
      return fail(ctx, JNE_ERR_NONFINITE, b);
    }
    return JNE_OK;
  };
  // ctx->err is shared between device threads: serialise through a local copy
  int rc = body();
  if (rc) {
    // error exit: let whatever is still in flight on the slots' streams finish (no DMA may touch the caller's array or
    // the staging buffers after this call returns), and forget it -- a later call must not copy stale rows
    for (auto& s : dv.slot) { if (s.stream) cudaStreamSynchronize(s.stream); s.busy = false; }
    cudaGetLastError();
    if (err_out) { std::lock_guard<std::mutex> lk(ctx->err_mu); *err_out = ctx->err; }
  }
  return rc;
}

int eigs_batch_sync_impl(jne_ctx* ctx, uint32_t mask, uint32_t dim, uint32_t steps, const uint32_t* seeds, uint64_t n,
                         double* out, jne_rows_sink sink_fn = nullptr, void* sink_user = nullptr) {
  const JneRunParams prm = make_params_mask(mask, dim, steps, false);
  const size_t nd = ctx->devs.size();
  if (n == 0) return JNE_OK;
  if (nd == 1) return run_share(ctx, ctx->devs[0], prm, seeds, n, RowSink{out, sink_fn, sink_user, 0}, nullptr);
  // contiguous slices, one host thread per device, no collective
  std::vector<std::thread> th;
  std::vector<int> rc(nd, JNE_OK);
  std::vector<std::string> errs(nd);
  const uint64_t per = (n + nd - 1) / nd;
  try {
    for (size_t i = 0; i < nd; ++i) {
      const uint64_t a = std::min<uint64_t>(i * per, n), b = std::min<uint64_t>(a + per, n);
      if (a == b) continue;
      th.emplace_back([&, i, a, b]() {
        rc[i] = run_share(ctx, ctx->devs[i], prm, seeds + a, b - a, RowSink{out ? out + a * prm.p : nullptr, sink_fn, sink_user, a}, &errs[i]);
      });
    }
  } catch (...) {            // a thread could not be started: the ones that run still use rc / errs / out
    for (auto& t : th) t.join();
    throw;
  }
  for (auto& t : th) t.join();
  for (size_t i = 0; i < nd; ++i)
    if (rc[i]) { std::lock_guard<std::mutex> lk(ctx->err_mu); ctx->err = errs[i]; return rc[i]; }
  return JNE_OK;
}

// Nothing may cross the C ABI: host-side failures (std::bad_alloc, std::system_error from std::thread ...) become
// JNE_ERR_INTERNAL with the message in jne_last_error.
template <class F> auto guarded(jne_ctx* ctx, F&& f) -> decltype(f()) {
  try {
    return f();
  } catch (const std::exception& e) {
    return (decltype(f()))fail(ctx, JNE_ERR_INTERNAL, std::string("host-side failure: ") + e.what());
  } catch (...) {
    return (decltype(f()))fail(ctx, JNE_ERR_INTERNAL, "host-side failure: unknown exception");
  }
}

int eigs_batch_sync(jne_ctx* ctx, uint32_t mask, uint32_t dim, uint32_t steps, const uint32_t* seeds, uint64_t n,
                    double* out) {
  return guarded(ctx, [&]() { return eigs_batch_sync_impl(ctx, mask, dim, steps, seeds, n, out); });
}

void join_worker(jne_ctx* ctx) {
  if (ctx->worker.joinable()) ctx->worker.join();
}

// ---- FP64 peak microbenchmark (roofline denominator) ----
template <int MODE>
__global__ void __launch_bounds__(256) fp64_peak_kernel(int iters, double seed, double* sink) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = seed * (i + threadIdx.x);
  const double a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) jne_dmma(acc[2 * i], acc[2 * i + 1], a, b);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 123.456) sink[0] = s;
}

// ---- streaming statistics ----
// SumAggregator / MaxAggregator of src/simulation_analyzers.rs:25-40: the sum runs over the record in stored
// (descending) order, the maximum folds from f64::MIN.
__global__ void jne_aggregate_kernel(const double* __restrict__ eigs, uint64_t n, uint32_t p, uint32_t stride,
                                     double* __restrict__ trace, double* __restrict__ maxeig) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* e = eigs + i * stride;
  double sum = 0.0, mx = -1.7976931348623157e308;
  for (uint32_t k = 0; k < p; ++k) { sum += e[k]; mx = fmax(mx, e[k]); }
  trace[i] = sum;
  maxeig[i] = mx;
}

// ---- exact order statistics without a sort (row f3) ----
// Doubles map to 64-bit keys whose unsigned order is the doubles' order.  The k-th smallest key of a sample that is
// spread over several arrays (one per device) is found digit by digit: eight passes of 8 bits, each a histogram of
// the next digit over the keys that match the prefixes found so far.  Several ranks are resolved together (one
// histogram row of 256 counters per distinct prefix), and several samples per pass (the ten statistics of a five-model
// job share the passes).  Nothing but these histograms leaves a device -- a few KB per pass; with 16-bit digits the
// 256 KB rows, copied and summed per device, sample and pass, were two thirds of the wall time of the 8-GPU default
// sweep at small dims -- and the result does not depend on how the sample is partitioned.
__device__ __forceinline__ uint64_t jne_order_key(double x) {
  const uint64_t b = (uint64_t)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
constexpr int kSelMaxGroups = 32;
struct JneSelPrefixes { uint64_t prefix[kSelMaxGroups]; int n; };
// hist[g][digit] += 1 for every key whose bits above the digit equal prefix[g] (pass 0: one group, every key)
__global__ void __launch_bounds__(256)
jne_select_hist_kernel(const double* __restrict__ vals, uint64_t n, int shift, JneSelPrefixes pf, unsigned int* __restrict__ hist) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < n; base += stride) {   // whole warps stay together
    const uint64_t i = base + threadIdx.x;
    int slot = -1;
    if (i < n) {
      const uint64_t key = jne_order_key(vals[i]);
      const uint64_t hi = shift >= 56 ? 0ull : key >> (shift + 8);
      int g = -1;
      for (int k = 0; k < pf.n; ++k) if (pf.prefix[k] == hi) g = k;
      if (g >= 0) slot = (g << 8) | (int)((key >> shift) & 0xffull);
    }
    // warp-aggregated: one atomic per distinct (group, digit) in the warp (the leading digits of a sample are few)
    const unsigned active = __ballot_sync(0xffffffffu, slot >= 0);
    if (slot >= 0) {
      const unsigned peers = __match_any_sync(active, slot);
      if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(hist + slot, (unsigned int)__popc(peers));
    }
  }
}

__global__ void jne_iota_kernel(uint32_t first, uint64_t n, uint32_t* __restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = first + (uint32_t)i;
}

}  // namespace

extern "C" {

const char* jne_version(void) { return JNE_VERSION_STR; }

int jne_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int jne_num_eigs(uint8_t model, uint32_t dim) {
  if (model > 4 || dim < 1 || dim > 255) return JNE_ERR_INVALID_ARG;
  return (model == 1 || model == 3) ? (int)dim + 1 : (int)dim;
}

double jne_flops_per_run(uint8_t model, uint32_t dim, uint32_t steps) {
  const double p = (model == 1 || model == 3) ? dim + 1.0 : dim;
  return 2.0 * steps * (p * (p + 1.0) / 2.0 + p * dim);
}

const char* jne_last_error(const jne_ctx* ctx) {
  if (ctx) return ctx->err.c_str();
  std::lock_guard<std::mutex> lk(g_init_err_mu);
  static thread_local std::string copy;
  copy = g_init_err;
  return copy.c_str();
}

uint64_t jne_launch_count(const jne_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }

int jne_jacobi_table(uint32_t ne, uint32_t* words, uint32_t capacity) {
  if (ne < 2 || ne > 16 || (ne & 1u)) return JNE_ERR_INVALID_ARG;
  const uint32_t np = ne / 2, n = (ne - 1) * (np + np * (np + 1) / 2);
  if (n <= capacity) {
    if (!words) return JNE_ERR_INVALID_ARG;
    std::vector<uint32_t> all(8 * kTabWords);
    make_jacobi_tables(all.data());
    std::memcpy(words, all.data() + (np - 1) * kTabWords, n * sizeof(uint32_t));
  }
  return (int)n;
}

int64_t jne_trend_weight_table(uint32_t steps, double* table, uint64_t capacity) {
  if (steps < 1 || steps > kAuxMaxSteps) return JNE_ERR_INVALID_ARG;
  const uint64_t n = (uint64_t)seg_len_for(steps) * 16u;
  if (n <= capacity) {
    if (!table) return JNE_ERR_INVALID_ARG;
    const std::vector<double> tab = make_aux_table(steps);
    std::memcpy(table, tab.data(), n * sizeof(double));
  }
  return (int64_t)n;
}

void jne_shutdown(jne_ctx* ctx) {
  DeviceGuard device_guard;
  if (!ctx) return;
  join_worker(ctx);
  for (auto& dv : ctx->devs) {
    if (cudaSetDevice(dv.id) != cudaSuccess) continue;
    for (auto& s : dv.slot) {
      if (s.d_seeds) cudaFree(s.d_seeds);
      if (s.d_out) cudaFree(s.d_out);
      if (s.h_seeds) cudaFreeHost(s.h_seeds);
      if (s.h_out) cudaFreeHost(s.h_out);
      if (s.done) cudaEventDestroy(s.done);
      if (s.stream) cudaStreamDestroy(s.stream);
    }
    if (dv.d_err) cudaFree(dv.d_err);
    if (dv.d_jtab) cudaFree(dv.d_jtab);
    for (auto& kv : dv.aux_tabs) cudaFree(kv.second);
    for (double* m : dv.d_mom) if (m) cudaFree(m);
    if (dv.h_err) cudaFreeHost(dv.h_err);
    if (dv.d_scratch) cudaFree(dv.d_scratch);
    if (dv.stream) cudaStreamDestroy(dv.stream);
  }
  delete ctx;
}

int jne_init(const int* device_ids, int n_devices, jne_ctx** out) {
  DeviceGuard device_guard;
  if (!out) return fail(nullptr, JNE_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  int visible = 0;
  cudaError_t e = cudaGetDeviceCount(&visible);
  if (e != cudaSuccess || visible == 0) {
    cudaGetLastError();
    return fail(nullptr, JNE_ERR_CUDA,
                std::string("no usable CUDA device (") + (e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e)) +
                    "); this library has no CPU fallback");
  }
  if (n_devices < 0) return fail(nullptr, JNE_ERR_INVALID_ARG, "n_devices < 0");
  if (n_devices == 0) n_devices = visible;
  jne_ctx* ctx = new (std::nothrow) jne_ctx();
  if (!ctx) return fail(nullptr, JNE_ERR_INTERNAL, "out of host memory");
  if (const char* ln = std::getenv("JNE_LANE")) ctx->use_lane = std::strcmp(ln, "0") != 0;
  if (const char* gr = std::getenv("JNE_GROUP")) ctx->use_group = std::strcmp(gr, "0") != 0;
  if (const char* ls = std::getenv("JNE_LANE_SOLVE")) ctx->lane_thread_solve = std::strcmp(ls, "warp") != 0;
  if (const char* ax = std::getenv("JNE_AUX")) ctx->use_aux = std::strcmp(ax, "0") != 0;
  try { ctx->devs.resize(n_devices); } catch (...) { delete ctx; return fail(nullptr, JNE_ERR_INTERNAL, "out of host memory"); }
  for (int i = 0; i < n_devices; ++i) {
    Device& dv = ctx->devs[i];
    dv.id = device_ids ? device_ids[i] : i;
    if (dv.id < 0 || dv.id >= visible) {
      fail(nullptr, JNE_ERR_INVALID_ARG, "device id out of range");
      jne_shutdown(ctx);
      return JNE_ERR_INVALID_ARG;
    }
    auto setup = [&]() -> int {
      JNE_CUDA(nullptr, cudaSetDevice(dv.id));
      cudaDeviceProp prop;
      JNE_CUDA(nullptr, cudaGetDeviceProperties(&prop, dv.id));
      if (prop.major != 10)
        return fail(nullptr, JNE_ERR_CUDA, std::string("device ") + prop.name + " is not sm_100 (B200); the kernels are built for sm_100a only");
      dv.sm_count = prop.multiProcessorCount;
      JNE_CUDA(nullptr, cudaStreamCreateWithFlags(&dv.stream, cudaStreamNonBlocking));
      for (auto& s : dv.slot) {
        JNE_CUDA(nullptr, cudaMalloc(&s.d_seeds, kSlotSeeds * sizeof(uint32_t)));
        JNE_CUDA(nullptr, cudaMalloc(&s.d_out, kSlotDoubles * sizeof(double)));
        JNE_CUDA(nullptr, cudaMallocHost(&s.h_seeds, kSlotSeeds * sizeof(uint32_t)));
        JNE_CUDA(nullptr, cudaMallocHost(&s.h_out, kSlotDoubles * sizeof(double)));
        JNE_CUDA(nullptr, cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        JNE_CUDA(nullptr, cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
      }
      std::vector<uint32_t> tables(8 * kTabWords);
      make_jacobi_tables(tables.data());
      JNE_CUDA(nullptr, cudaMalloc(&dv.d_jtab, tables.size() * sizeof(uint32_t)));
      JNE_CUDA(nullptr, cudaMemcpy(dv.d_jtab, tables.data(), tables.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
      JNE_CUDA(nullptr, cudaMalloc(&dv.d_err, sizeof(unsigned int)));
      JNE_CUDA(nullptr, cudaMemset(dv.d_err, 0, sizeof(unsigned int)));
      JNE_CUDA(nullptr, cudaMallocHost(&dv.h_err, sizeof(unsigned int)));
      *dv.h_err = 0;
      JNE_CUDA(nullptr, cudaDeviceSynchronize());   // the uploads above ride the legacy stream; the work streams are non-blocking
      return JNE_OK;
    };
    int rc = setup();
    if (rc) { jne_shutdown(ctx); return rc; }
  }
  *out = ctx;
  return JNE_OK;
}

int jne_eigs_batch(jne_ctx* ctx, uint8_t model, uint32_t dim, uint32_t steps, const uint32_t* seeds, uint64_t n,
                   double* out) {
  DeviceGuard device_guard;
  if (!ctx) return JNE_ERR_INVALID_ARG;
  join_worker(ctx);
  int rc = validate(ctx, model, dim, steps);
  if (rc) return rc;
  if (n && (!seeds || !out)) return fail(ctx, JNE_ERR_INVALID_ARG, "seeds/out is NULL");
  return eigs_batch_sync(ctx, 1u << model, dim, steps, seeds, n, out);
}

int jne_eigs_batch_multi(jne_ctx* ctx, uint32_t model_mask, uint32_t dim, uint32_t steps, const uint32_t* seeds,
                         uint64_t n, double* out) {
  DeviceGuard device_guard;
  if (!ctx) return JNE_ERR_INVALID_ARG;
  join_worker(ctx);
  if (model_mask == 0 || model_mask > 31u) return fail(ctx, JNE_ERR_INVALID_ARG, "model_mask must select models 0..4");
  for (int m = 0; m < 5; ++m)
    if ((model_mask >> m) & 1u) { int rc = validate(ctx, (uint8_t)m, dim, steps); if (rc) return rc; }
  if (n && (!seeds || !out)) return fail(ctx, JNE_ERR_INVALID_ARG, "seeds/out is NULL");
  return eigs_batch_sync(ctx, model_mask, dim, steps, seeds, n, out);
}

int jne_eigs_batch_multi_stream(jne_ctx* ctx, uint32_t model_mask, uint32_t dim, uint32_t steps, const uint32_t* seeds,
                                uint64_t n, jne_rows_sink sink, void* user) {
  DeviceGuard device_guard;
  if (!ctx) return JNE_ERR_INVALID_ARG;
  join_worker(ctx);
  if (model_mask == 0 || model_mask > 31u) return fail(ctx, JNE_ERR_INVALID_ARG, "model_mask must select models 0..4");
  for (int m = 0; m < 5; ++m)
    if ((model_mask >> m) & 1u) { int rc = validate(ctx, (uint8_t)m, dim, steps); if (rc) return rc; }
  if (!sink || (n && !seeds)) return fail(ctx, JNE_ERR_INVALID_ARG, "seeds/sink is NULL");
  return guarded(ctx, [&]() { return eigs_batch_sync_impl(ctx, model_mask, dim, steps, seeds, n, nullptr, sink, user); });
}

int jne_ctx_device_count(const jne_ctx* ctx) { return ctx ? (int)ctx->devs.size() : JNE_ERR_INVALID_ARG; }

int jne_multi_width(uint32_t model_mask, uint32_t dim) {
  if (model_mask == 0 || model_mask > 31u || dim < 1 || dim > 255) return JNE_ERR_INVALID_ARG;
  return (int)mask_width(model_mask, dim);
}

int jne_eigs_batch_multi_device(jne_ctx* ctx, uint32_t model_mask, uint32_t dim, uint32_t steps, const void* d_seeds,
                                uint64_t n, void* d_out, void* stream) {
  DeviceGuard device_guard;
  if (!ctx) return JNE_ERR_INVALID_ARG;
  if (model_mask == 0 || model_mask > 31u) return fail(ctx, JNE_ERR_INVALID_ARG, "model_mask must select models 0..4");
  for (int m = 0; m < 5; ++m)
    if ((model_mask >> m) & 1u) { int rc = validate(ctx, (uint8_t)m, dim, steps); if (rc) return rc; }
  if (n == 0) return JNE_OK;
  if (!d_seeds || !d_out) return fail(ctx, JNE_ERR_INVALID_ARG, "d_seeds/d_out is NULL");
  Device& dv = ctx->devs[0];
  JNE_CUDA(ctx, cudaSetDevice(dv.id));
  JneRunParams prm = make_params_mask(model_mask, dim, steps, false);
  prm.jtab = jtab_for(dv, dim);
  const uint64_t max_runs = (uint64_t)0x7fffffffu * JNE_WARPS_PER_CTA;
  for (uint64_t off = 0; off < n; off += max_runs) {
    const uint64_t m = std::min(max_runs, n - off);
    JNE_CUDA(ctx, launch_run<true>(ctx, dv, (const uint32_t*)d_seeds + off, nullptr, m, prm, (double*)d_out + off * prm.out_stride,
                                   dv.d_err, nullptr, (cudaStream_t)stream));
    ctx->launches.fetch_add(1);
  }
  return JNE_OK;
}

int64_t jne_submit(jne_ctx* ctx, uint8_t model, uint32_t dim, uint32_t steps, const uint32_t* seeds, uint64_t n,
                   double* out) {
  if (!ctx) return JNE_ERR_INVALID_ARG;
  if (ctx->pending_ticket) return fail(ctx, JNE_ERR_INVALID_ARG, "a ticket is already outstanding; call jne_wait first");
  int rc = validate(ctx, model, dim, steps);
  if (rc) return rc;
  if (n && (!seeds || !out)) return fail(ctx, JNE_ERR_INVALID_ARG, "seeds/out is NULL");
  join_worker(ctx);
  return guarded(ctx, [&]() -> int64_t {
    auto copy = std::make_shared<std::vector<uint32_t>>(seeds, seeds + n);
    ctx->worker = std::thread([ctx, model, dim, steps, copy, n, out]() {
      ctx->pending_status = eigs_batch_sync(ctx, 1u << model, dim, steps, copy->data(), n, out);
    });
    ctx->pending_ticket = ctx->next_ticket++;     // only once the worker exists
    return ctx->pending_ticket;
  });
}

int64_t jne_submit_multi(jne_ctx* ctx, uint32_t model_mask, uint32_t dim, uint32_t steps, const uint32_t* seeds, uint64_t n,
                         double* out) {
  if (!ctx) return JNE_ERR_INVALID_ARG;
  if (ctx->pending_ticket) return fail(ctx, JNE_ERR_INVALID_ARG, "a ticket is already outstanding; call jne_wait first");
  if (model_mask == 0 || model_mask > 31u) return fail(ctx, JNE_ERR_INVALID_ARG, "model_mask must select models 0..4");
  for (int m = 0; m < 5; ++m)
    if ((model_mask >> m) & 1u) { int rc = validate(ctx, (uint8_t)m, dim, steps); if (rc) return rc; }
  if (n && (!seeds || !out)) return fail(ctx, JNE_ERR_INVALID_ARG, "seeds/out is NULL");
  join_worker(ctx);
  return guarded(ctx, [&]() -> int64_t {
    auto copy = std::make_shared<std::vector<uint32_t>>(seeds, seeds + n);
    ctx->worker = std::thread([ctx, model_mask, dim, steps, copy, n, out]() {
      ctx->pending_status = eigs_batch_sync(ctx, model_mask, dim, steps, copy->data(), n, out);
    });
    ctx->pending_ticket = ctx->next_ticket++;     // only once the worker exists
    return ctx->pending_ticket;
  });
}

int jne_wait(jne_ctx* ctx, int64_t ticket) {
  DeviceGuard device_guard;
  if (!ctx) return JNE_ERR_INVALID_ARG;
  if (ticket <= 0 || ticket != ctx->pending_ticket) return fail(ctx, JNE_ERR_INVALID_ARG, "unknown ticket");
  join_worker(ctx);
  ctx->pending_ticket = 0;
  return ctx->pending_status;
}

int jne_eigs_batch_device(jne_ctx* ctx, uint8_t model, uint32_t dim, uint32_t steps, const void* d_seeds, uint64_t n,
                          void* d_out, void* stream) {
  DeviceGuard device_guard;
  if (!ctx) return JNE_ERR_INVALID_ARG;
  int rc = validate(ctx, model, dim, steps);
  if (rc) return rc;
  if (n == 0) return JNE_OK;
  if (!d_seeds || !d_out) return fail(ctx, JNE_ERR_INVALID_ARG, "d_seeds/d_out is NULL");
  Device& dv = ctx->devs[0];
  JNE_CUDA(ctx, cudaSetDevice(dv.id));
  JneRunParams prm = make_params(model, dim, steps, false);
  prm.jtab = jtab_for(dv, dim);
  // grid.x is 32-bit: split very large batches
  const uint64_t max_runs = (uint64_t)0x7fffffffu * JNE_WARPS_PER_CTA;
  for (uint64_t off = 0; off < n; off += max_runs) {
    const uint64_t m = std::min(max_runs, n - off);
    JNE_CUDA(ctx, launch_run<true>(ctx, dv, (const uint32_t*)d_seeds + off, nullptr, m, prm, (double*)d_out + off * prm.p,
                                   dv.d_err, nullptr, (cudaStream_t)stream));
    ctx->launches.fetch_add(1);
  }
  return JNE_OK;
}

int jne_check_async(jne_ctx* ctx) {
  DeviceGuard device_guard;
  if (!ctx) return JNE_ERR_INVALID_ARG;
  Device& dv = ctx->devs[0];
  JNE_CUDA(ctx, cudaSetDevice(dv.id));
  JNE_CUDA(ctx, cudaDeviceSynchronize());
  unsigned int h = 0;
  JNE_CUDA(ctx, cudaMemcpy(&h, dv.d_err, sizeof h, cudaMemcpyDeviceToHost));
  JNE_CUDA(ctx, cudaMemset(dv.d_err, 0, sizeof h));
  if (h) {
    char b[128];
    snprintf(b, sizeof b, "%u run(s) produced non-finite eigenvalues", h);
    return fail(ctx, JNE_ERR_NONFINITE, b);
  }
  return JNE_OK;
}

static int single_device_run(jne_ctx* ctx, const JneRunParams& prm_in, const uint32_t* seeds, const double* dB,
                             uint64_t n, double* out, double* mats) {
  Device& dv = ctx->devs[0];
  JNE_CUDA(ctx, cudaSetDevice(dv.id));
  JneRunParams prm = prm_in;
  prm.jtab = jtab_for(dv, prm.dim);
  const bool rng = dB == nullptr;
  const uint64_t per_run_in = rng ? 0 : (uint64_t)prm.dim * prm.steps;
  // bound the scratch: <= 1 GiB of increments, <= 2^18 runs per launch
  uint64_t chunk = 1ull << 15;
  if (!rng) chunk = std::max<uint64_t>(1, std::min<uint64_t>(chunk, (1ull << 27) / std::max<uint64_t>(1, per_run_in)));
  if (mats) chunk = std::min<uint64_t>(chunk, 1ull << 14);
  const size_t in_bytes = rng ? chunk * sizeof(uint32_t) : chunk * per_run_in * sizeof(double);
  const size_t out_bytes = chunk * prm.p * sizeof(double);
  const size_t dbg_bytes = mats ? chunk * 512 * sizeof(double) : 0;
  int rc = ensure_scratch(ctx, dv, in_bytes + out_bytes + dbg_bytes + 512);
  if (rc) return rc;
  char* base = (char*)dv.d_scratch;
  double* d_out = (double*)base;
  double* d_dbg = mats ? (double*)(base + out_bytes) : nullptr;
  void* d_in = base + out_bytes + dbg_bytes;
  JNE_CUDA(ctx, cudaMemsetAsync(dv.d_err, 0, sizeof(unsigned int), dv.stream));
  for (uint64_t off = 0; off < n; off += chunk) {
    const uint64_t m = std::min(chunk, n - off);
    if (rng) {
      JNE_CUDA(ctx, cudaMemcpyAsync(d_in, seeds + off, m * sizeof(uint32_t), cudaMemcpyHostToDevice, dv.stream));
      JNE_CUDA(ctx, launch_run<true>(ctx, dv, (const uint32_t*)d_in, nullptr, m, prm, d_out, dv.d_err, d_dbg, dv.stream));
    } else {
      JNE_CUDA(ctx, cudaMemcpyAsync(d_in, dB + off * per_run_in, m * per_run_in * sizeof(double), cudaMemcpyHostToDevice, dv.stream));
      JNE_CUDA(ctx, launch_run<false>(ctx, dv, nullptr, (const double*)d_in, m, prm, d_out, dv.d_err, d_dbg, dv.stream));
    }
    ctx->launches.fetch_add(1);
    JNE_CUDA(ctx, cudaMemcpyAsync(out + off * prm.p, d_out, m * prm.p * sizeof(double), cudaMemcpyDeviceToHost, dv.stream));
    if (mats) JNE_CUDA(ctx, cudaMemcpyAsync(mats + off * 512, d_dbg, m * 512 * sizeof(double), cudaMemcpyDeviceToHost, dv.stream));
    JNE_CUDA(ctx, cudaStreamSynchronize(dv.stream));
  }
  JNE_CUDA(ctx, cudaMemcpyAsync(dv.h_err, dv.d_err, sizeof(unsigned int), cudaMemcpyDeviceToHost, dv.stream));
  JNE_CUDA(ctx, cudaStreamSynchronize(dv.stream));
  if (*dv.h_err) {
    char b[128];
    snprintf(b, sizeof b, "%u run(s) produced non-finite eigenvalues", *dv.h_err);
    return fail(ctx, JNE_ERR_NONFINITE, b);
  }
  return JNE_OK;
}

int jne_eigs_from_increments(jne_ctx* ctx, uint8_t model, uint32_t dim, uint32_t steps, const double* dB, uint64_t n,
                             double* out) {
  DeviceGuard device_guard;
  if (!ctx) return JNE_ERR_INVALID_ARG;
  join_worker(ctx);
  int rc = validate(ctx, model, dim, steps);
  if (rc) return rc;
  if (n == 0) return JNE_OK;
  if (!dB || !out) return fail(ctx, JNE_ERR_INVALID_ARG, "dB/out is NULL");
  return single_device_run(ctx, make_params(model, dim, steps, true), nullptr, dB, n, out, nullptr);
}

int jne_eigs_batch_debug(jne_ctx* ctx, uint8_t model, uint32_t dim, uint32_t steps, const uint32_t* seeds, uint64_t n,
                         double* out, double* mats) {
  DeviceGuard device_guard;
  if (!ctx) return JNE_ERR_INVALID_ARG;
  join_worker(ctx);
  int rc = validate(ctx, model, dim, steps);
  if (rc) return rc;
  if (n == 0) return JNE_OK;
  if (!seeds || !out || !mats) return fail(ctx, JNE_ERR_INVALID_ARG, "seeds/out/mats is NULL");
  return single_device_run(ctx, make_params(model, dim, steps, false), seeds, nullptr, n, out, mats);
}

int jne_gen_normal_matrix(jne_ctx* ctx, uint32_t dim, uint32_t steps, uint32_t seed, double* out) {
  DeviceGuard device_guard;
  if (!ctx) return JNE_ERR_INVALID_ARG;
  join_worker(ctx);
  if (dim < 1 || steps < 1 || !out) return fail(ctx, JNE_ERR_INVALID_ARG, "dim/steps/out invalid");
  if ((uint64_t)dim * steps > (1ull << 31)) return fail(ctx, JNE_ERR_INVALID_ARG, "Matrix too large");  // src/rng_matrix.rs:12
  Device& dv = ctx->devs[0];
  JNE_CUDA(ctx, cudaSetDevice(dv.id));
  const size_t bytes = (size_t)dim * steps * sizeof(double);
  int rc = ensure_scratch(ctx, dv, bytes);
  if (rc) return rc;
  const uint64_t items = (uint64_t)((steps + 3) / 4) * dim;
  jne_normal_matrix_kernel<<<(unsigned)((items + 255) / 256), 256, 0, dv.stream>>>(seed, dim, steps, dv.d_scratch);
  JNE_CUDA(ctx, cudaGetLastError());
  ctx->launches.fetch_add(1);
  JNE_CUDA(ctx, cudaMemcpyAsync(out, dv.d_scratch, bytes, cudaMemcpyDeviceToHost, dv.stream));
  JNE_CUDA(ctx, cudaStreamSynchronize(dv.stream));
  return JNE_OK;
}

int jne_brownian_motion_matrix(jne_ctx* ctx, uint32_t dim, uint32_t steps, double delta_t, uint32_t seed, double* out) {
  DeviceGuard device_guard;
  if (!ctx) return JNE_ERR_INVALID_ARG;
  join_worker(ctx);
  if (dim < 1 || steps < 1 || !out) return fail(ctx, JNE_ERR_INVALID_ARG, "dim/steps/out invalid");
  if ((uint64_t)dim * ((uint64_t)steps + 1) > (1ull << 31)) return fail(ctx, JNE_ERR_INVALID_ARG, "Matrix too large");
  Device& dv = ctx->devs[0];
  JNE_CUDA(ctx, cudaSetDevice(dv.id));
  const size_t bytes = (size_t)dim * ((size_t)steps + 1) * sizeof(double);
  int rc = ensure_scratch(ctx, dv, bytes);
  if (rc) return rc;
  jne_brownian_kernel<<<(dim + 31) / 32, 32, 0, dv.stream>>>(seed, dim, steps, delta_t, dv.d_scratch);
  JNE_CUDA(ctx, cudaGetLastError());
  ctx->launches.fetch_add(1);
  JNE_CUDA(ctx, cudaMemcpyAsync(out, dv.d_scratch, bytes, cudaMemcpyDeviceToHost, dv.stream));
  JNE_CUDA(ctx, cudaStreamSynchronize(dv.stream));
  return JNE_OK;
}

int jne_pencil_eigs_batch(jne_ctx* ctx, uint32_t p, uint32_t d, const double* S1, const double* S2, uint64_t n,
                          double* out) {
  DeviceGuard device_guard;
  if (!ctx) return JNE_ERR_INVALID_ARG;
  join_worker(ctx);
  if (p < 1 || p > 16 || d < 1 || d > p) return fail(ctx, JNE_ERR_INVALID_ARG, "need 1 <= d <= p <= 16");
  if (n == 0) return JNE_OK;
  if (!S1 || !S2 || !out) return fail(ctx, JNE_ERR_INVALID_ARG, "S1/S2/out is NULL");
  Device& dv = ctx->devs[0];
  JNE_CUDA(ctx, cudaSetDevice(dv.id));
  JNE_CUDA(ctx, cudaFuncSetAttribute(jne_pencil_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pencil_smem()));
  const uint64_t chunk = 1ull << 16;
  const size_t b1 = chunk * p * d * sizeof(double), b2 = chunk * p * p * sizeof(double), bo = chunk * p * sizeof(double);
  int rc = ensure_scratch(ctx, dv, b1 + b2 + bo);
  if (rc) return rc;
  double* d1 = dv.d_scratch;
  double* d2 = d1 + chunk * p * d;
  double* dout = d2 + chunk * p * p;
  JNE_CUDA(ctx, cudaMemsetAsync(dv.d_err, 0, sizeof(unsigned int), dv.stream));
  for (uint64_t off = 0; off < n; off += chunk) {
    const uint64_t m = std::min(chunk, n - off);
    JNE_CUDA(ctx, cudaMemcpyAsync(d1, S1 + off * p * d, m * p * d * sizeof(double), cudaMemcpyHostToDevice, dv.stream));
    JNE_CUDA(ctx, cudaMemcpyAsync(d2, S2 + off * p * p, m * p * p * sizeof(double), cudaMemcpyHostToDevice, dv.stream));
    jne_pencil_kernel<<<(unsigned)((m + JNE_WARPS_PER_CTA - 1) / JNE_WARPS_PER_CTA), 32 * JNE_WARPS_PER_CTA, pencil_smem(), dv.stream>>>(
        d1, d2, m, (int)p, (int)d, 1.0, dout, dv.d_err, jtab_for(dv, d));
    JNE_CUDA(ctx, cudaGetLastError());
    ctx->launches.fetch_add(1);
    JNE_CUDA(ctx, cudaMemcpyAsync(out + off * p, dout, m * p * sizeof(double), cudaMemcpyDeviceToHost, dv.stream));
    JNE_CUDA(ctx, cudaStreamSynchronize(dv.stream));
  }
  JNE_CUDA(ctx, cudaMemcpyAsync(dv.h_err, dv.d_err, sizeof(unsigned int), cudaMemcpyDeviceToHost, dv.stream));
  JNE_CUDA(ctx, cudaStreamSynchronize(dv.stream));
  if (*dv.h_err) return fail(ctx, JNE_ERR_NONFINITE, "non-finite eigenvalues (S2 not positive definite?)");
  return JNE_OK;
}

}  // extern "C"

namespace {

// One statistic of one model: its values as they lie on the context's devices (n[i] doubles at vals[i] on device i).
struct SelSample { std::vector<const double*> vals; std::vector<uint64_t> n; };

// One selection job: x_(k) for every k in ranks (0-based, any order, duplicates allowed) of the union of the sample's
// arrays.  All jobs of a call share the eight passes: per pass every device histograms every job (one launch each, no
// synchronisation in between), then one small copy per device brings all rows back and the host advances every rank.
struct SelJob { SelSample smp; std::vector<uint64_t> ranks; std::vector<double> x; };

int select_ranks(jne_ctx* ctx, const std::vector<int>& dev_index, std::vector<SelJob>& jobs) {
  const size_t nd = dev_index.size(), nj = jobs.size();
  constexpr size_t kRow = 256;
  std::vector<std::vector<uint64_t>> prefix(nj), within(nj);
  std::vector<std::vector<int>> group(nj);
  for (size_t j = 0; j < nj; ++j) { prefix[j].assign(jobs[j].ranks.size(), 0); within[j] = jobs[j].ranks; group[j].resize(jobs[j].ranks.size()); }
  std::vector<unsigned int*> d_hist(nd, nullptr);
  const size_t cap_words = nj * (size_t)kSelMaxGroups * kRow;
  std::vector<unsigned int> h_hist, h_part(cap_words);
  auto body = [&]() -> int {
    for (size_t i = 0; i < nd; ++i) {
      JNE_CUDA(ctx, cudaSetDevice(ctx->devs[dev_index[i]].id));
      JNE_CUDA(ctx, cudaMalloc(&d_hist[i], cap_words * sizeof(unsigned int)));
    }
    std::vector<JneSelPrefixes> pf(nj);
    std::vector<size_t> row0(nj);
    for (int pass = 0; pass < 8; ++pass) {
      const int shift = 56 - 8 * pass;
      size_t words = 0;
      for (size_t j = 0; j < nj; ++j) {                    // distinct prefixes among the job's ranks
        pf[j] = JneSelPrefixes{};
        for (size_t r = 0; r < prefix[j].size(); ++r) {
          int g = -1;
          for (int k = 0; k < pf[j].n; ++k) if (pf[j].prefix[k] == prefix[j][r]) g = k;
          if (g < 0) {
            if (pf[j].n == kSelMaxGroups) return fail(ctx, JNE_ERR_INVALID_ARG, "too many distinct percentiles in one call (at most 16)");
            g = pf[j].n; pf[j].prefix[pf[j].n++] = prefix[j][r];
          }
          group[j][r] = g;
        }
        row0[j] = words;
        words += (size_t)pf[j].n * kRow;
      }
      h_hist.assign(words, 0u);
      for (size_t i = 0; i < nd; ++i) {
        Device& dv = ctx->devs[dev_index[i]];
        JNE_CUDA(ctx, cudaSetDevice(dv.id));
        JNE_CUDA(ctx, cudaMemsetAsync(d_hist[i], 0, words * sizeof(unsigned int), dv.stream));
        for (size_t j = 0; j < nj; ++j) {
          const uint64_t n = jobs[j].smp.n[i];
          if (n == 0) continue;
          const unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)dv.sm_count * 16);
          jne_select_hist_kernel<<<grid, 256, 0, dv.stream>>>(jobs[j].smp.vals[i], n, shift, pf[j], d_hist[i] + row0[j]);
          ctx->launches.fetch_add(1);
        }
        JNE_CUDA(ctx, cudaGetLastError());
      }
      for (size_t i = 0; i < nd; ++i) {                    // merge: the only data that leaves a device
        Device& dv = ctx->devs[dev_index[i]];
        JNE_CUDA(ctx, cudaSetDevice(dv.id));
        JNE_CUDA(ctx, cudaMemcpyAsync(h_part.data(), d_hist[i], words * sizeof(unsigned int), cudaMemcpyDeviceToHost, dv.stream));
        JNE_CUDA(ctx, cudaStreamSynchronize(dv.stream));
        for (size_t w = 0; w < words; ++w) h_hist[w] += h_part[w];
      }
      for (size_t j = 0; j < nj; ++j)
        for (size_t r = 0; r < prefix[j].size(); ++r) {
          const unsigned int* h = h_hist.data() + row0[j] + (size_t)group[j][r] * kRow;
          uint64_t cum = 0;
          size_t digit = 0;
          for (; digit < kRow; ++digit) {
            if (within[j][r] < cum + h[digit]) break;
            cum += h[digit];
          }
          if (digit == kRow) return fail(ctx, JNE_ERR_INVALID_ARG, "rank outside the sample");
          within[j][r] -= cum;
          prefix[j][r] = (prefix[j][r] << 8) | (uint64_t)digit;
        }
    }
    for (size_t j = 0; j < nj; ++j) {
      jobs[j].x.resize(prefix[j].size());
      for (size_t r = 0; r < prefix[j].size(); ++r) {
        const uint64_t key = prefix[j][r];
        const uint64_t bits = (key >> 63) ? (key & 0x7fffffffffffffffull) : ~key;
        std::memcpy(&jobs[j].x[r], &bits, 8);
      }
    }
    return JNE_OK;
  };
  const int rc = body();
  for (size_t i = 0; i < nd; ++i)
    if (d_hist[i]) { cudaSetDevice(ctx->devs[dev_index[i]].id); cudaFree(d_hist[i]); }
  return rc;
}

// get_percentile_value of src/simulation_analyzers.rs:4-18 for every q and every sample: rank = q (n - 1), linear
// interpolation between the two neighbouring order statistics; two products and one sum, each rounded (no FMA
// contraction), as the reference.  outs[j] receives the nq percentiles of samples[j].
int percentiles_of_samples(jne_ctx* ctx, const std::vector<int>& dev_index, const std::vector<SelSample>& samples,
                           const double* qs, uint32_t nq, const std::vector<double*>& outs) {
  std::vector<SelJob> jobs;
  std::vector<size_t> job_of(samples.size(), (size_t)-1);
  std::vector<uint64_t> total(samples.size(), 0);
  for (size_t j = 0; j < samples.size(); ++j) {
    uint64_t n = 0;
    for (uint64_t m : samples[j].n) n += m;
    total[j] = n;
    if (n == 0) { for (uint32_t k = 0; k < nq; ++k) outs[j][k] = std::nan(""); continue; }
    SelJob job;
    job.smp = samples[j];
    for (uint32_t k = 0; k < nq; ++k) {
      const double rank = qs[k] * (double)(n - 1);
      if (!(rank >= 0.0) || rank > (double)(n - 1)) return fail(ctx, JNE_ERR_INVALID_ARG, "percentile outside [0, 1]");
      job.ranks.push_back((uint64_t)std::floor(rank));
      job.ranks.push_back((uint64_t)std::ceil(rank));
    }
    job_of[j] = jobs.size();
    jobs.push_back(std::move(job));
  }
  if (jobs.empty()) return JNE_OK;
  const int rc = select_ranks(ctx, dev_index, jobs);
  if (rc) return rc;
  for (size_t j = 0; j < samples.size(); ++j) {
    if (job_of[j] == (size_t)-1) continue;
    const SelJob& job = jobs[job_of[j]];
    for (uint32_t k = 0; k < nq; ++k) {
      const double rank = qs[k] * (double)(total[j] - 1);
      const uint64_t lo = job.ranks[2 * k], hi = job.ranks[2 * k + 1];
      if (lo == hi) { outs[j][k] = job.x[2 * k]; continue; }
      const double w = rank - (double)lo;
      volatile double a = job.x[2 * k] * (1.0 - w), b = job.x[2 * k + 1] * w;   // volatile: two roundings, then the sum
      outs[j][k] = a + b;
    }
  }
  return JNE_OK;
}

// Simulates seeds first_seed .. first_seed + n - 1 for every model of the mask on ALL devices of the context (contiguous
// shares, one host thread per device) and keeps, per model, trace and max-eig of every run on the device that computed
// it; then the percentiles by exact distributed selection.
int simulate_percentiles_impl(jne_ctx* ctx, uint32_t model_mask, uint32_t dim, uint32_t steps, uint32_t first_seed, uint64_t n,
                              const double* qs, uint32_t nq, double* trace_out, double* maxeig_out) {
  const size_t nd = ctx->devs.size();
  const int nm = __builtin_popcount(model_mask);
  const JneRunParams prm0 = make_params_mask(model_mask, dim, steps, false);
  std::vector<double*> d_agg(nd, nullptr);                 // per device: [model slot][trace | max][share]
  std::vector<uint64_t> share(nd, 0), start(nd, 0);
  const uint64_t per = (n + nd - 1) / nd;
  for (size_t i = 0; i < nd; ++i) { start[i] = std::min<uint64_t>(i * per, n); share[i] = std::min<uint64_t>(start[i] + per, n) - start[i]; }
  std::vector<int> rcs(nd, JNE_OK);
  std::vector<std::string> errs(nd);
  auto device_part = [&](size_t i) -> int {
    Device& dv = ctx->devs[i];
    const uint64_t ni = share[i];
    if (ni == 0) return JNE_OK;
    JNE_CUDA(ctx, cudaSetDevice(dv.id));
    JneRunParams prm = prm0;
    prm.jtab = jtab_for(dv, dim);
    const uint64_t chunk = 1ull << 19;
    // one region of the device's grow-only scratch buffer: [statistics | eigenvalues of a chunk | seeds of a chunk].  A sweep
    // calls this once per dim; allocating and freeing ~1 GB per call cost the one-GPU default sweep a second
    const size_t agg_doubles = ((size_t)nm * 2 * ni + 31) & ~(size_t)31, eig_doubles = (size_t)chunk * prm.out_stride;   // 256-byte aligned parts
    const size_t eig_reserve = (size_t)chunk * mask_width(model_mask, JNE_MAX_DIM);   // sized for any dim: no regrowth along a sweep
    {
      const int rc = ensure_scratch(ctx, dv, (agg_doubles + std::max(eig_doubles, eig_reserve)) * sizeof(double) + chunk * sizeof(uint32_t));
      if (rc) return rc;
    }
    d_agg[i] = dv.d_scratch;
    double* d_eigs = dv.d_scratch + agg_doubles;
    uint32_t* d_seeds = reinterpret_cast<uint32_t*>(d_eigs + eig_doubles);
    auto body = [&]() -> int {
      JNE_CUDA(ctx, cudaMemsetAsync(dv.d_err, 0, sizeof(unsigned int), dv.stream));
      for (uint64_t off = 0; off < ni; off += chunk) {
        const uint64_t m = std::min(chunk, ni - off);
        jne_iota_kernel<<<(unsigned)((m + 255) / 256), 256, 0, dv.stream>>>(first_seed + (uint32_t)(start[i] + off), m, d_seeds);
        JNE_CUDA(ctx, launch_run<true>(ctx, dv, d_seeds, nullptr, m, prm, d_eigs, dv.d_err, nullptr, dv.stream));
        ctx->launches.fetch_add(2);
        uint32_t col = 0;
        int slot = 0;
        for (int mod = 0; mod < 5; ++mod) {              // one Brownian path per seed serves every selected model
          if (!((model_mask >> mod) & 1u)) continue;
          const uint32_t p = (mod == 1 || mod == 3) ? dim + 1 : dim;
          double* agg = d_agg[i] + (size_t)slot * 2 * ni;
          jne_aggregate_kernel<<<(unsigned)((m + 255) / 256), 256, 0, dv.stream>>>(d_eigs + col, m, p, prm.out_stride, agg + off, agg + ni + off);
          ctx->launches.fetch_add(1);
          col += p; ++slot;
        }
        JNE_CUDA(ctx, cudaGetLastError());
      }
      JNE_CUDA(ctx, cudaMemcpyAsync(dv.h_err, dv.d_err, sizeof(unsigned int), cudaMemcpyDeviceToHost, dv.stream));
      JNE_CUDA(ctx, cudaStreamSynchronize(dv.stream));
      if (*dv.h_err) return fail(ctx, JNE_ERR_NONFINITE, "non-finite eigenvalues");
      return JNE_OK;
    };
    return body();
  };
  if (nd == 1) {
    rcs[0] = device_part(0);
  } else {
    std::vector<std::thread> th;
    try {
      for (size_t i = 0; i < nd; ++i)
        th.emplace_back([&, i]() { rcs[i] = device_part(i); if (rcs[i]) { std::lock_guard<std::mutex> lk(ctx->err_mu); errs[i] = ctx->err; } });
    } catch (...) { for (auto& t : th) t.join(); throw; }
    for (auto& t : th) t.join();
  }
  int rc = JNE_OK;
  for (size_t i = 0; i < nd; ++i)
    if (rcs[i] && !rc) { rc = rcs[i]; if (nd > 1) { std::lock_guard<std::mutex> lk(ctx->err_mu); ctx->err = errs[i]; } }
  if (!rc) {
    std::vector<int> dev_index(nd);
    for (size_t i = 0; i < nd; ++i) dev_index[i] = (int)i;
    std::vector<SelSample> samples;                        // trace and max-eig of every selected model: one selection
    std::vector<double*> outs;
    for (int slot = 0; slot < nm; ++slot) {
      SelSample tr, mx;
      for (size_t i = 0; i < nd; ++i) {
        const double* agg = d_agg[i] ? d_agg[i] + (size_t)slot * 2 * share[i] : nullptr;
        tr.vals.push_back(agg); tr.n.push_back(share[i]);
        mx.vals.push_back(agg ? agg + share[i] : nullptr); mx.n.push_back(share[i]);
      }
      samples.push_back(std::move(tr)); outs.push_back(trace_out + (size_t)slot * nq);
      samples.push_back(std::move(mx)); outs.push_back(maxeig_out + (size_t)slot * nq);
    }
    rc = percentiles_of_samples(ctx, dev_index, samples, qs, nq, outs);
  }
  return rc;     // the statistics stay in the devices' scratch buffers (reused by the next call)
}

}  // namespace

extern "C" {

int jne_percentiles_device(jne_ctx* ctx, const void* d_eigs, uint64_t n, uint32_t p, uint32_t stride, const double* qs,
                           uint32_t nq, double* trace_out, double* maxeig_out, void* stream) {
  DeviceGuard device_guard;
  if (!ctx) return JNE_ERR_INVALID_ARG;
  if (!d_eigs || !qs || !trace_out || !maxeig_out || p < 1 || stride < p || nq < 1)
    return fail(ctx, JNE_ERR_INVALID_ARG, "jne_percentiles_device: bad arguments");
  join_worker(ctx);
  return guarded(ctx, [&]() -> int {
    Device& dv = ctx->devs[0];
    JNE_CUDA(ctx, cudaSetDevice(dv.id));
    double* d_buf = nullptr;
    JNE_CUDA(ctx, cudaMalloc(&d_buf, 2 * std::max<uint64_t>(n, 1) * sizeof(double)));
    cudaStream_t st = (cudaStream_t)stream;
    auto body = [&]() -> int {
      if (n) jne_aggregate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const double*)d_eigs, n, p, stride, d_buf, d_buf + n);
      JNE_CUDA(ctx, cudaGetLastError());
      ctx->launches.fetch_add(1);
      JNE_CUDA(ctx, cudaStreamSynchronize(st));       // the selection runs on the context's own stream
      return percentiles_of_samples(ctx, {0}, {SelSample{{d_buf}, {n}}, SelSample{{d_buf + n}, {n}}}, qs, nq, {trace_out, maxeig_out});
    };
    const int rc = body();
    cudaFree(d_buf);
    return rc;
  });
}

int jne_simulate_percentiles(jne_ctx* ctx, uint8_t model, uint32_t dim, uint32_t steps, uint32_t first_seed, uint64_t n,
                             const double* qs, uint32_t nq, double* trace_out, double* maxeig_out) {
  DeviceGuard device_guard;
  if (!ctx) return JNE_ERR_INVALID_ARG;
  join_worker(ctx);
  int rc = validate(ctx, model, dim, steps);
  if (rc) return rc;
  if (!qs || !trace_out || !maxeig_out || nq < 1) return fail(ctx, JNE_ERR_INVALID_ARG, "jne_simulate_percentiles: bad arguments");
  if ((uint64_t)first_seed + n > (1ull << 32)) return fail(ctx, JNE_ERR_INVALID_ARG, "seed range exceeds u32");
  return guarded(ctx, [&]() { return simulate_percentiles_impl(ctx, 1u << model, dim, steps, first_seed, n, qs, nq, trace_out, maxeig_out); });
}

int jne_simulate_percentiles_multi(jne_ctx* ctx, uint32_t model_mask, uint32_t dim, uint32_t steps, uint32_t first_seed,
                                   uint64_t n, const double* qs, uint32_t nq, double* trace_out, double* maxeig_out) {
  DeviceGuard device_guard;
  if (!ctx) return JNE_ERR_INVALID_ARG;
  join_worker(ctx);
  if (model_mask == 0 || model_mask > 31u) return fail(ctx, JNE_ERR_INVALID_ARG, "model_mask must select models 0..4");
  for (int m = 0; m < 5; ++m)
    if ((model_mask >> m) & 1u) { int rc = validate(ctx, (uint8_t)m, dim, steps); if (rc) return rc; }
  if (!qs || !trace_out || !maxeig_out || nq < 1) return fail(ctx, JNE_ERR_INVALID_ARG, "jne_simulate_percentiles_multi: bad arguments");
  if ((uint64_t)first_seed + n > (1ull << 32)) return fail(ctx, JNE_ERR_INVALID_ARG, "seed range exceeds u32");
  return guarded(ctx, [&]() { return simulate_percentiles_impl(ctx, model_mask, dim, steps, first_seed, n, qs, nq, trace_out, maxeig_out); });
}

int jne_fp64_peak_tflops(jne_ctx* ctx, int mode, double ms_target, double* tflops) {
  DeviceGuard device_guard;
  if (!ctx || !tflops) return JNE_ERR_INVALID_ARG;
  join_worker(ctx);
  Device& dv = ctx->devs[0];
  JNE_CUDA(ctx, cudaSetDevice(dv.id));
  cudaDeviceProp prop;
  JNE_CUDA(ctx, cudaGetDeviceProperties(&prop, dv.id));
  int rc = ensure_scratch(ctx, dv, 4096);
  if (rc) return rc;
  const int blocks = prop.multiProcessorCount * 2;
  cudaEvent_t e0, e1;
  JNE_CUDA(ctx, cudaEventCreate(&e0));
  JNE_CUDA(ctx, cudaEventCreate(&e1));
  auto launch = [&](int iters) {
    if (mode == 0) fp64_peak_kernel<0><<<blocks, 256, 0, dv.stream>>>(iters, 1.0000001, dv.d_scratch);
    else fp64_peak_kernel<1><<<blocks, 256, 0, dv.stream>>>(iters, 1.0000001, dv.d_scratch);
    ctx->launches.fetch_add(1);
  };
  // flops per thread per iter: mode 0: 64 FMA; mode 1: 32 DMMA x 256 FMA / 32 lanes = 256 FMA
  const double fma_per_thread_iter = mode == 0 ? 64.0 : 256.0;
  int iters = mode == 0 ? 4000 : 1000;
  launch(iters);  // warm-up + calibration
  JNE_CUDA(ctx, cudaStreamSynchronize(dv.stream));
  float ms = 0.f;
  for (int rep = 0; rep < 2; ++rep) {
    JNE_CUDA(ctx, cudaEventRecord(e0, dv.stream));
    launch(iters);
    JNE_CUDA(ctx, cudaEventRecord(e1, dv.stream));
    JNE_CUDA(ctx, cudaEventSynchronize(e1));
    JNE_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    if (rep == 0) iters = (int)std::max(1.0, std::min(2.0e6, iters * (ms_target / std::max(1e-3f, ms))));
  }
  JNE_CUDA(ctx, cudaGetLastError());
  *tflops = 2.0 * fma_per_thread_iter * iters * 256.0 * blocks / (ms * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return JNE_OK;
}

}  // extern "C"
