/*
 * jne.h -- C ABI of the B200-native Johansen null-eigenspectra hot path.
 *
 * Drop-in boundary.  The reference (Kuan-Lun/johansen-null-eigenspectra
 * v0.8.0, paths relative to its root) has no FFI today; these entry points are what a thin
 * `extern "C"` block in the Rust crate would bind so that
 *     src/data_storage/parallel_compute.rs:14-41  calculate_eigenvalues_parallel
 * calls the GPU instead of
 *     src/johansen_statistics.rs:59-85            calculate_eigenvalues
 * once per seed.  INTEGRATION.md shows that binding.
 *
 * Conventions: plain pointers and sizes, no structs by value; every function returns
 * JNE_OK (0) or a negative jne_status and never aborts; the message of the last failure on a
 * context is available through jne_last_error().  The caller owns every buffer it passes in;
 * the library keeps no caller pointer after a synchronous call returns.  One host thread per
 * context at a time.  There is NO CPU fallback: without a usable CUDA device jne_init fails.
 */
#ifndef JNE_H
#define JNE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct jne_ctx jne_ctx;

typedef enum jne_status {
  JNE_OK = 0,
  JNE_ERR_INVALID_ARG = -1,  /* model > 4, dim outside 1..JNE_MAX_DIM, steps < 1, null pointer ... */
  JNE_ERR_CUDA = -2,         /* any CUDA runtime failure (no device, OOM, launch failure, ECC ...) */
  JNE_ERR_NONFINITE = -3,    /* a run produced a NaN/Inf eigenvalue; the reference panics at
                                src/johansen_statistics.rs:45 (partial_cmp().unwrap()) */
  JNE_ERR_UNSUPPORTED = -4,  /* dim > JNE_MAX_DIM */
  JNE_ERR_IO = -5,           /* .dat writer / reader failures (jne_dat_* functions) */
  JNE_ERR_INTERNAL = -6      /* host-side failure inside the library (out of memory, a thread could not be started):
                                caught at the boundary, never thrown across it */
} jne_status;

#define JNE_MAX_DIM 15

/* ---- library / device ------------------------------------------------------------------ */

/* Version string "jne-b200 <semver> (sm_100a)". */
const char* jne_version(void);

/* Number of visible CUDA devices (0 when none / no driver). */
int jne_device_count(void);

/* Create a context over `n_devices` devices (device_ids == NULL: devices 0..n-1; n_devices == 0:
 * all visible devices).  Seed lists are sharded contiguously over the context's devices with no
 * collective.  Replaces nothing in the reference (rayon's global pool,
 * src/cli.rs:311-329, plays this role there). */
int jne_init(const int* device_ids, int n_devices, jne_ctx** out);
void jne_shutdown(jne_ctx* ctx);
const char* jne_last_error(const jne_ctx* ctx);   /* ctx may be NULL: last jne_init failure */

/* Eigenvalues per run: dim+1 for models 1 and 3, else dim; negative status on invalid input.
 * Mirrors src/data_storage/thread_manager.rs:40-44. */
int jne_num_eigs(uint8_t model, uint32_t dim);

/* ---- the hot path ---------------------------------------------------------------------- */

/* Batched calculate_eigenvalues (src/johansen_statistics.rs:59-85) for an arbitrary list of
 * seeds (src/data_storage/progress.rs:56-61 hands out arbitrary subsets of 1..=num_runs).
 * out: n x p doubles, row i = eigenvalues of seeds[i], descending (:45).  Host buffers; the
 * call copies seeds H2D, runs the fused kernels on every device of the context, and copies the
 * eigenvalues D2H before returning.  Same (model, dim, steps, seed) => bit-identical row,
 * whatever n, the batch split or the number of devices. */
int jne_eigs_batch(jne_ctx* ctx, uint8_t model, uint32_t dim, uint32_t steps,
                   const uint32_t* seeds, uint64_t n, double* out);

/* Fused multi-model batch.  The reference's Brownian path depends on
 * (dim, steps, seed) only (src/rng_matrix.rs:11) and its CLI evaluates every model on the same seeds
 * (src/main.rs:109), so one pass over the path yields the eigenvalues of every model selected in
 * model_mask (bit m = model m).  out: n rows of jne_multi_width(model_mask, dim) doubles; within a row the
 * selected models follow each other in ascending model order, each block descending.  Every block is
 * bit-identical to what jne_eigs_batch returns for that (model, dim, steps, seed). */
int jne_eigs_batch_multi(jne_ctx* ctx, uint32_t model_mask, uint32_t dim, uint32_t steps,
                         const uint32_t* seeds, uint64_t n, double* out);
int jne_eigs_batch_multi_device(jne_ctx* ctx, uint32_t model_mask, uint32_t dim, uint32_t steps,
                                const void* d_seeds, uint64_t n, void* d_out, void* stream);
int jne_multi_width(uint32_t model_mask, uint32_t dim);   /* sum of jne_num_eigs over the selected models */
int jne_ctx_device_count(const jne_ctx* ctx);             /* devices this context shards its batches over */

/* Streaming form of the fused batch -- the closest counterpart of the reference's interface, which SENDS every run's
 * record through a channel as it is finished (`sender: mpsc::Sender<(u32, Vec<f64>)>`,
 * src/data_storage/parallel_compute.rs:14-41): the rows of seeds[first .. first + count) -- `count` rows of
 * jne_multi_width(model_mask, dim) doubles, same content as jne_eigs_batch_multi -- are handed to `sink` as soon as
 * they are back on the host, straight from the library's pinned staging memory (valid during the call only), without
 * the copy into a caller array.  `sink` runs on the library's per-device host threads: it may be called concurrently
 * for disjoint ranges, in no particular order, and must not call back into the same context; a non-zero return value
 * aborts the batch (JNE_ERR_IO).  Returns after every row has been delivered.  run_models_simulation encodes the .dat
 * records inside the sink. */
typedef int (*jne_rows_sink)(void* user, uint64_t first, uint64_t count, const double* rows);
int jne_eigs_batch_multi_stream(jne_ctx* ctx, uint32_t model_mask, uint32_t dim, uint32_t steps,
                                const uint32_t* seeds, uint64_t n, jne_rows_sink sink, void* user);

/* Asynchronous pair: jne_submit enqueues the batch (seeds are copied before it returns; `out`
 * must stay valid until jne_wait) and returns a ticket > 0, or a negative status.  jne_wait
 * blocks until that batch's eigenvalues are in `out`.  At most one ticket may be outstanding
 * per context; lets the caller overlap GPU work with the .dat writer thread
 * (src/data_storage/thread_manager.rs:25-92). */
int64_t jne_submit(jne_ctx* ctx, uint8_t model, uint32_t dim, uint32_t steps,
                   const uint32_t* seeds, uint64_t n, double* out);
/* The same for the fused multi-model batch (rows of jne_multi_width(model_mask, dim) doubles). */
int64_t jne_submit_multi(jne_ctx* ctx, uint32_t model_mask, uint32_t dim, uint32_t steps,
                         const uint32_t* seeds, uint64_t n, double* out);
int jne_wait(jne_ctx* ctx, int64_t ticket);

/* Same computation with everything already resident on device 0 of the context: d_seeds and
 * d_out are DEVICE pointers, `stream` is a cudaStream_t (NULL = default stream).  Enqueues only;
 * the caller synchronises.  Non-finite runs are counted and reported by the next synchronous
 * call or by jne_check_async(). */
int jne_eigs_batch_device(jne_ctx* ctx, uint8_t model, uint32_t dim, uint32_t steps,
                          const void* d_seeds, uint64_t n, void* d_out, void* stream);
int jne_check_async(jne_ctx* ctx);   /* syncs device 0; JNE_ERR_NONFINITE if any run since the last check failed */

/* calculate_eigenvalues_from_matrices fed by caller-supplied increments (parity gate 1):
 * dB holds n runs, each dim x steps column-major (element (r, t) at t*dim + r, the layout of
 * src/rng_matrix.rs:36), already scaled (dB = sqrt(dt) z).  The Brownian path is rebuilt as
 * the reference does: B_0 = 0, naive cumulative sum (src/matrix_utils.rs:51-63), dB re-derived
 * by subtraction (src/johansen_statistics.rs:80-82), delta_t = 1/steps (:70). */
int jne_eigs_from_increments(jne_ctx* ctx, uint8_t model, uint32_t dim, uint32_t steps,
                             const double* dB, uint64_t n, double* out);

/* ---- pieces of the path exposed for parity tests ------------------------------------------ */

/* gen_normal_matrix(nrows = dim, ncols = steps, seed) (src/rng_matrix.rs:11-37): dim x steps
 * column-major standard normals of the device stream (stream JNE2: Philox4x32-10-keyed xoshiro128++ substreams + FP32 Box-Muller, jne_rng.cuh). */
int jne_gen_normal_matrix(jne_ctx* ctx, uint32_t dim, uint32_t steps, uint32_t seed, double* out);

/* brownian_motion_matrix(dim, steps, delta_t, AlongColumns, zeros, seed) (src/rng_matrix.rs:57-141):
 * dim x (steps+1) column-major, first column zero. */
int jne_brownian_motion_matrix(jne_ctx* ctx, uint32_t dim, uint32_t steps, double delta_t,
                               uint32_t seed, double* out);

/* The eigen-solve alone: for each of n problems, eigenvalues |alpha|/beta of the pencil
 * (S1' S1, S2), descending -- what GeneralizedEigen::new + raw_eigenvalues + sort compute at
 * src/johansen_statistics.rs:35-46.  S1: d x p column-major per problem (sum dB F'),
 * S2: p x p symmetric positive definite.  1 <= d <= p <= 16. */
int jne_pencil_eigs_batch(jne_ctx* ctx, uint32_t p, uint32_t d, const double* S1, const double* S2,
                          uint64_t n, double* out);

/* Debug: like jne_eigs_batch on device 0, additionally returning per run the assembled
 * S2 (16 x 16, row-major, zero padded) followed by R = S1' (16 x 16): 512 doubles per run. */
int jne_eigs_batch_debug(jne_ctx* ctx, uint8_t model, uint32_t dim, uint32_t steps,
                         const uint32_t* seeds, uint64_t n, double* out, double* mats);

/* ---- streaming statistics ------------------------------------------- */

/* Percentiles of the trace (sum of a record's eigenvalues) and of the max-eigenvalue over n records resident on
 * device 0 of the context (row i at d_eigs + i*stride, p values) -- calculate_trace_percentiles / calculate_maxeig_percentiles,
 * src/simulation_analyzers.rs:42-81, which today re-read and sort the whole .dat file on the host.  qs: n_q
 * fractions in [0,1]; value at rank q (n-1) with linear interpolation (:4-18).  qs / outputs are HOST arrays. */
int jne_percentiles_device(jne_ctx* ctx, const void* d_eigs, uint64_t n, uint32_t p, uint32_t stride,
                           const double* qs, uint32_t n_q, double* trace_out, double* maxeig_out, void* stream);

/* Simulate seeds first_seed .. first_seed+n-1 on EVERY device of the context (contiguous shares) and return only those
 * percentiles: eigenvalues never leave the GPUs (seeds are generated on the device, records are reduced to (trace, max)
 * as they are produced, the order statistics are found by an exact digit-by-digit selection whose per-device
 * histograms are the only data merged on the host -- no sort, no collective).  The result does not depend on the
 * number of devices.  At most 16 percentiles per call. */
int jne_simulate_percentiles(jne_ctx* ctx, uint8_t model, uint32_t dim, uint32_t steps, uint32_t first_seed,
                             uint64_t n, const double* qs, uint32_t n_q, double* trace_out, double* maxeig_out);

/* The same for every model selected in model_mask from ONE fused pass over the seeds (rows f2 x f3: the CLI prints
 * these statistics after each model of its loop, src/main.rs:117-126).  trace_out / maxeig_out: n_q doubles per
 * selected model, models in ascending order. */
int jne_simulate_percentiles_multi(jne_ctx* ctx, uint32_t model_mask, uint32_t dim, uint32_t steps, uint32_t first_seed,
                                   uint64_t n, const double* qs, uint32_t n_q, double* trace_out, double* maxeig_out);

/* ---- orchestration helper (C++ mirror in csrc/jne_host.hpp) -------------------------------------- */

/* run_model_simulation (src/data_storage/parallel_compute.rs:150-232) for one (model, dim, steps, num_runs) job:
 * resume scan of `filename` -> remaining seeds of 1..=num_runs -> GPU batches -> batched EIGENVALS_V6 append ->
 * trailer.  Runs on `ctx` when it is not NULL, else on a context of its own over device_ids (NULL / 0 = all GPUs).
 * stats (3 x u64, may be NULL): records present before, records computed now, records in the file after. */
int jne_run_model_simulation(jne_ctx* ctx, uint8_t model, uint32_t dim, uint32_t steps, uint64_t num_runs,
                             const char* filename, int quiet, const int* device_ids, int n_devices, uint64_t* stats);

/* The CLI's model loop over one dim (`for &model in &models_vec { ... run_simulation() }`, src/main.rs:109-114) as ONE
 * fused pass: every model selected in model_mask gets its own EIGENVALS_V6 file (filenames[m] for bit m; entries of
 * unselected models may be NULL), resumed independently (scan, parameter mismatch -> restart, complete -> untouched);
 * seeds are grouped by the set of models that still lack them and each group is one jne_eigs_batch_multi stream, so a
 * fresh five-model job evaluates every Brownian path once instead of five times.  Each file is byte-identical to the
 * one jne_run_model_simulation writes for that model.  stats (5 x 3 x u64, may be NULL): per model as above. */
int jne_run_models_simulation(jne_ctx* ctx, uint32_t model_mask, uint32_t dim, uint32_t steps, uint64_t num_runs,
                              const char* const* filenames, int quiet, const int* device_ids, int n_devices,
                              uint64_t* stats);

/* ---- measurement helpers ------------------------------------------------------------------ */

/* Register-resident DFMA (mode 0) or mma.sync.m8n8k4.f64 (mode 1) throughput on device 0 of the
 * context, in TFLOP/s, timed with CUDA events over `ms_target` milliseconds of work.  This is the
 * FP64 roofline denominator (MEASURED_PEAKS.json holds no FP64 figure). */
int jne_fp64_peak_tflops(jne_ctx* ctx, int mode, double ms_target, double* tflops);

/* Kernel launches issued by this context since creation (bench.py's gpu_launches). */
uint64_t jne_launch_count(const jne_ctx* ctx);

/* Algorithmic flops per run: 2 T [p(p+1)/2 + p d]. */
double jne_flops_per_run(uint8_t model, uint32_t dim, uint32_t steps);

/* The step table the device Jacobi walks for an ne x ne problem (ne even, 2..16): ne-1 round-robin steps, each
 * ne/2 rotation words  o(p,p) | o(q,q) << 8 | o(p,q) << 16  followed by one word per 2x2 block of pair slots
 * P1 <= P2,  o(p1,p2) | o(p1,q2) << 8 | o(q1,p2) << 16 | o(q1,q2) << 24,  o(i,j) = offset of element (min, max)
 * in the packed upper triangle.  Host-only (no device needed): lets the CPU tests check the schedule.  Returns the
 * number of words (written when it is <= capacity) or a negative error code. */
int jne_jacobi_table(uint32_t ne, uint32_t* words, uint32_t capacity);

/* The trend-weight table the AUX kernels (dim <= 6 and 9..12, a model with a trend row) feed to the tensor pipe for a
 * run of `steps` steps: seg_len local steps (128 ceil(steps / 512): whole generator epochs; short horizons 8 ceil(steps / 32), whichever wastes fewer lane-steps) x 4 weights x 4 time segments, doubles,
 * table[(j * 4 + m) * 4 + k] for local step j of segment k (global step i = k seg_len + j, segment end e_k), with
 * w1_i = 2i + 1 - T, w2_i = 3 w1_i^2 - (T^2 - 1) (the integer forms of the reference's trend regressors
 * (i+1)/T - 1/2 and the residual of ((i+1)/T)^2 on [1, (i+1)/T], src/johansen_statistics.rs:127-135,170-194):
 *   m = 0: sum_{i < i' < e_k} w1_i'    m = 1: sum_{i < i' < e_k} w2_i'    m = 2: w2_i    m = 3: w1_i
 * and 0 for steps at or beyond the end of the segment.  Host-only (no device needed): lets the CPU tests check the
 * table against exact integer arithmetic.  Returns the number of doubles (written when it is <= capacity) or a
 * negative error code; steps must be 1..2^22 (longer runs use the scalar-sum kernels). */
int64_t jne_trend_weight_table(uint32_t steps, double* table, uint64_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* JNE_H */
