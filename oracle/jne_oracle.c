/*
 * jne_oracle.c -- CPU restatement of the reference hot path.  TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (johansen_null_eigenspectra_b200/) never does.
 *
 * PARITY STATUS: "parity unpinned" -- the reference (Kuan-Lun/johansen-null-eigenspectra v0.8.0)
 * cannot be compiled here (no cargo/rustc; nalgebra is a git dependency that is not vendored;
 * no system LAPACK) and holds no golden vector for this path.  This port is validated against
 * oracle/johansen_oracle.py (same increments => same eigenvalues) in tests/test_oracle.py.
 *
 * What follows the reference, function by function (paths relative to /root/reference):
 *   xoshiro256++ / SplitMix64 / ziggurat : third-party, NOT under /root/reference --
 *       rand_xoshiro 0.7.0 (Xoshiro256PlusPlus::seed_from_u64 = SplitMix64 fill; next_u64),
 *       rand 0.9.1 (random::<u64>() = next_u64; random::<f64>() = (next_u64 >> 11) 2^-53; Open01),
 *       rand_distr 0.5.1 (StandardNormal = 256-layer ziggurat on next_u64), Cargo.lock:285-330.
 *       Restated from the published algorithms (Blackman & Vigna 2019; Marsaglia & Tsang 2000 /
 *       Doornik 2005 ZIGNOR with R = 3.654152885361009, V = 4.92867323399e-3).  The tables are
 *       recomputed, so the stream is not claimed bit-identical to rand_distr's.
 *   gen_normal_matrix              src/rng_matrix.rs:11-37   (chunk scheme incl. physical-core dependence)
 *   brownian_motion_matrix         src/rng_matrix.rs:57-141  (concat, scale, row-wise cumsum: 3 buffers)
 *   dmatrix_cumsum RowWise         src/matrix_utils.rs:51-63
 *   sum_of_outer_products          src/matrix_utils.rs:67-85 (sequential order; the reference's rayon
 *                                                              reduce order is nondeterministic)
 *   construct_f_matrix             src/johansen_statistics.rs:102-197
 *   calculate_eigenvalues[_from_matrices]  src/johansen_statistics.rs:24-85, eigen-solve through LAPACK
 *       dggev('V','V') -- the routine nalgebra-lapack 0.25.0 GeneralizedEigen::new calls -- taken from
 *       scipy's bundled OpenBLAS (symbol scipy_dggev_).
 *   calculate_eigenvalues_parallel src/data_storage/parallel_compute.rs:14-41 (threads over seeds)
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

extern void scipy_dggev_(const char* jobvl, const char* jobvr, const int* n, double* a, const int* lda, double* b,
                         const int* ldb, double* alphar, double* alphai, double* beta, double* vl, const int* ldvl,
                         double* vr, const int* ldvr, double* work, const int* lwork, int* info, size_t, size_t);
extern void scipy_dsygv_(const int* itype, const char* jobz, const char* uplo, const int* n, double* a, const int* lda,
                         double* b, const int* ldb, double* w, double* work, const int* lwork, int* info, size_t, size_t);
extern void scipy_openblas_set_num_threads(int);

/* ---------------- RNG: xoshiro256++ seeded by SplitMix64 ---------------- */
typedef struct { uint64_t s[4]; } xo_t;
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t splitmix64(uint64_t* st) {
  uint64_t z = (*st += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static void xo_seed(xo_t* r, uint64_t seed) { for (int i = 0; i < 4; ++i) r->s[i] = splitmix64(&seed); }
static inline uint64_t xo_next(xo_t* r) {
  uint64_t* s = r->s;
  const uint64_t res = rotl(s[0] + s[3], 23) + s[0];
  const uint64_t t = s[1] << 17;
  s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
  return res;
}
/* Known-answer hooks (tests/test_oracle.py): the published vectors of the two generators -- SplitMix64 from seed
 * 1234567 (Vigna's splitmix64.c) and xoshiro256++ from the state {1, 2, 3, 4} (the `reference` test of
 * rand_xoshiro's xoshiro256plusplus.rs) -- pin this restatement of the third-party arithmetic. */
void jne_oracle_splitmix64(uint64_t seed, size_t n, uint64_t* out) { for (size_t i = 0; i < n; ++i) out[i] = splitmix64(&seed); }
void jne_oracle_xoshiro_from_state(const uint64_t* state, size_t n, uint64_t* out) {
  xo_t r; memcpy(r.s, state, sizeof r.s);
  for (size_t i = 0; i < n; ++i) out[i] = xo_next(&r);
}
void jne_oracle_xoshiro_seed_from_u64(uint64_t seed, uint64_t* state_out) { xo_t r; xo_seed(&r, seed); memcpy(state_out, r.s, sizeof r.s); }

static inline double float_with_exponent(uint64_t bits52, int e) {
  uint64_t u = ((uint64_t)(1023 + e) << 52) | bits52;
  double d; memcpy(&d, &u, 8); return d;
}
static inline double open01(xo_t* r) { return float_with_exponent(xo_next(r) >> 12, 0) - (1.0 - 2.220446049250313e-16 / 2.0); }
static inline double unit_f64(xo_t* r) { return (double)(xo_next(r) >> 11) * (1.0 / 9007199254740992.0); }

/* ---------------- ziggurat StandardNormal ---------------- */
#define ZIG_R 3.654152885361008796
#define ZIG_V 4.92867323399e-3
static double zig_x[257], zig_f[257];
static pthread_once_t zig_once = PTHREAD_ONCE_INIT;
static void zig_init(void) {
  const double fr = exp(-0.5 * ZIG_R * ZIG_R);
  zig_x[0] = ZIG_V / fr; zig_x[1] = ZIG_R;
  for (int i = 1; i < 255; ++i) zig_x[i + 1] = sqrt(-2.0 * log(ZIG_V / zig_x[i] + exp(-0.5 * zig_x[i] * zig_x[i])));
  zig_x[256] = 0.0;
  for (int i = 0; i < 257; ++i) zig_f[i] = exp(-0.5 * zig_x[i] * zig_x[i]);
}
static double standard_normal(xo_t* r) {
  for (;;) {
    const uint64_t bits = xo_next(r);
    const int i = (int)(bits & 0xff);
    const double u = float_with_exponent(bits >> 12, 1) - 3.0;
    const double x = u * zig_x[i];
    if (fabs(x) < zig_x[i + 1]) return x;
    if (i == 0) {
      double xx = 1.0, yy = 0.0;
      while (-2.0 * yy < xx * xx) { xx = log(open01(r)) / ZIG_R; yy = log(open01(r)); }
      return u < 0.0 ? xx - ZIG_R : ZIG_R - xx;
    }
    if (zig_f[i + 1] + (zig_f[i] - zig_f[i + 1]) * unit_f64(r) < exp(-0.5 * x * x)) return x;
  }
}

/* gen_normal_matrix (src/rng_matrix.rs:11-37); serial over chunks (inside the outer per-seed
 * parallel loop every core is already busy, so the nested rayon pool degenerates to this). */
void jne_oracle_gen_normal_matrix(size_t nrows, size_t ncols, uint64_t seed, size_t n_physical_cpus, double* data) {
  pthread_once(&zig_once, zig_init);
  const size_t total = nrows * ncols;
  const size_t min_chunk = 10000;
  size_t chunk_count = total / min_chunk;
  if (chunk_count < n_physical_cpus) chunk_count = n_physical_cpus;
  if (chunk_count > total) chunk_count = total;
  const size_t chunk_size = (total + chunk_count - 1) / chunk_count;
  xo_t base; xo_seed(&base, seed);
  uint64_t* derived = (uint64_t*)malloc(chunk_count * sizeof(uint64_t));
  for (size_t c = 0; c < chunk_count; ++c) derived[c] = xo_next(&base);
  size_t c = 0;
  for (size_t off = 0; off < total; off += chunk_size, ++c) {
    xo_t r; xo_seed(&r, derived[c]);
    const size_t end = off + chunk_size < total ? off + chunk_size : total;
    for (size_t k = off; k < end; ++k) data[k] = standard_normal(&r);
  }
  free(derived);
}

/* ---------------- linear algebra helpers (column-major, like nalgebra DMatrix) ---------------- */
/* sum_of_outer_products(a, b): a is ar x T, b is br x T  ->  ar x br  (src/matrix_utils.rs:67-85) */
static void sum_of_outer_products(const double* a, int ar, const double* b, int br, size_t T, double* out) {
  memset(out, 0, sizeof(double) * ar * br);
  for (size_t t = 0; t < T; ++t) {
    const double* ca = a + t * ar; const double* cb = b + t * br;
    for (int j = 0; j < br; ++j) { const double bj = cb[j]; for (int i = 0; i < ar; ++i) out[j * ar + i] += ca[i] * bj; }
  }
}

/* construct_f_matrix (src/johansen_statistics.rs:102-197); bm_prev is d x T column-major; returns p x T */
static double* construct_f_matrix(const double* bp, int d, size_t T, int model, int* p_out) {
  const int p = (model == 1 || model == 3) ? d + 1 : d;
  *p_out = p;
  double* f = (double*)calloc((size_t)p * T, sizeof(double));
  const double tt = (double)T;
  if (model == 0) { memcpy(f, bp, sizeof(double) * d * T); return f; }
  if (model == 1) {
    for (size_t t = 0; t < T; ++t) { memcpy(f + t * p, bp + t * d, sizeof(double) * d); f[t * p + d] = 1.0; }
    return f;
  }
  if (model == 2 || model == 3) {
    const int nb = model == 2 ? d - 1 : d;
    for (int r = 0; r < nb; ++r) {
      double s = 0.0; for (size_t t = 0; t < T; ++t) s += bp[t * d + r];
      const double mean = s / tt;
      for (size_t t = 0; t < T; ++t) f[t * p + r] = bp[t * d + r] - mean;
    }
    for (size_t t = 0; t < T; ++t) f[t * p + nb] = (double)(t + 1) / tt - 0.5;
    return f;
  }
  /* model 4: X = [B_{1..d-1}; tau^2], Z = [1; tau], F = X - (X Z')(Z Z')^-1 Z */
  double* x = (double*)malloc(sizeof(double) * d * T);
  double z00 = 0, z01 = 0, z11 = 0;
  for (size_t t = 0; t < T; ++t) {
    const double y = (double)(t + 1) / tt;
    for (int r = 0; r < d - 1; ++r) x[t * d + r] = bp[t * d + r];
    x[t * d + d - 1] = y * y;
    z00 += 1.0; z01 += y; z11 += y * y;
  }
  const double det = z00 * z11 - z01 * z01;
  const double i00 = z11 / det, i01 = -z01 / det, i11 = z00 / det;
  double* xz = (double*)calloc((size_t)d * 2, sizeof(double));   /* X Z' : d x 2 */
  for (size_t t = 0; t < T; ++t) {
    const double y = (double)(t + 1) / tt;
    for (int r = 0; r < d; ++r) { xz[r] += x[t * d + r]; xz[d + r] += x[t * d + r] * y; }
  }
  for (size_t t = 0; t < T; ++t) {
    const double y = (double)(t + 1) / tt;
    for (int r = 0; r < d; ++r) {
      const double c0 = xz[r] * i00 + xz[d + r] * i01, c1 = xz[r] * i01 + xz[d + r] * i11;
      f[t * p + r] = x[t * d + r] - (c0 + c1 * y);
    }
  }
  free(x); free(xz);
  return f;
}

/* calculate_eigenvalues_from_matrices (src/johansen_statistics.rs:24-47). out: p doubles, descending.
 * returns 0, or LAPACK info / -1000 on NaN. */
static int cmp_desc(const void* a, const void* b) { const double x = *(const double*)a, y = *(const double*)b; return (x < y) - (x > y); }
static int eigs_from_matrices(const double* bm_prev, const double* dbm, int d, size_t T, double delta_t, int model, double* out) {
  int p;
  double* fm = construct_f_matrix(bm_prev, d, T, model, &p);
  double* s1 = (double*)malloc(sizeof(double) * d * p);        /* d x p */
  double* s2 = (double*)malloc(sizeof(double) * p * p);
  sum_of_outer_products(dbm, d, fm, p, T, s1);
  sum_of_outer_products(fm, p, fm, p, T, s2);
  for (int i = 0; i < p * p; ++i) s2[i] *= delta_t;
  double* a = (double*)calloc((size_t)p * p, sizeof(double));  /* S1' S1 */
  for (int j = 0; j < p; ++j) for (int i = 0; i < p; ++i) { double s = 0; for (int k = 0; k < d; ++k) s += s1[i * d + k] * s1[j * d + k]; a[j * p + i] = s; }
  double alphar[32], alphai[32], beta[32];
  double* vl = (double*)malloc(sizeof(double) * p * p); double* vr = (double*)malloc(sizeof(double) * p * p);
  int lwork = 64 * p + 64, info = 0;
  double* work = (double*)malloc(sizeof(double) * lwork);
  scipy_dggev_("V", "V", &p, a, &p, s2, &p, alphar, alphai, beta, vl, &p, vr, &p, work, &lwork, &info, 1, 1);
  int rc = info;
  for (int i = 0; i < p; ++i) { out[i] = hypot(alphar[i], alphai[i]) / beta[i]; if (isnan(out[i])) rc = -1000; }
  qsort(out, p, sizeof(double), cmp_desc);
  free(fm); free(s1); free(s2); free(a); free(vl); free(vr); free(work);
  return rc;
}

/* eigenvalues from caller-supplied increments dB (d x T column-major) -- oracle side of parity gate (1). */
int jne_oracle_eigs_from_increments(int model, int d, size_t T, const double* dB, double* out) {
  double* bm = (double*)malloc(sizeof(double) * d * (T + 1));
  for (int r = 0; r < d; ++r) bm[r] = 0.0;
  for (size_t t = 0; t < T; ++t) for (int r = 0; r < d; ++r) bm[(t + 1) * d + r] = bm[t * d + r] + dB[t * d + r];
  double* dbm = (double*)malloc(sizeof(double) * d * T);
  for (size_t k = 0; k < (size_t)d * T; ++k) dbm[k] = bm[k + d] - bm[k];
  const int rc = eigs_from_matrices(bm, dbm, d, T, 1.0 / (double)T, model, out);
  free(bm); free(dbm);
  return rc;
}

/* calculate_eigenvalues (src/johansen_statistics.rs:59-85) */
int jne_oracle_calculate_eigenvalues(int d, size_t T, uint32_t seed, int model, size_t n_physical_cpus, double* out) {
  const double delta_t = 1.0 / (double)T;
  /* make_z_matrix: [start | normals]  (src/rng_matrix.rs:95-111) */
  double* gen = (double*)malloc(sizeof(double) * d * T);
  jne_oracle_gen_normal_matrix(d, T, (uint64_t)seed, n_physical_cpus, gen);
  double* z = (double*)malloc(sizeof(double) * d * (T + 1));
  memset(z, 0, sizeof(double) * d);
  memcpy(z + d, gen, sizeof(double) * d * T);
  free(gen);
  /* scaled = z * sqrt(dt) (:138-139) */
  const double sq = sqrt(delta_t);
  double* scaled = (double*)malloc(sizeof(double) * d * (T + 1));
  for (size_t k = 0; k < (size_t)d * (T + 1); ++k) scaled[k] = z[k] * sq;
  free(z);
  /* dmatrix_cumsum RowWise (src/matrix_utils.rs:51-63) */
  double* bm = (double*)malloc(sizeof(double) * d * (T + 1));
  for (int r = 0; r < d; ++r) { double acc = 0.0; for (size_t t = 0; t <= T; ++t) { acc += scaled[t * d + r]; bm[t * d + r] = acc; } }
  free(scaled);
  /* dbm = current - previous; bm_previous.into_owned() (:80-84) */
  double* dbm = (double*)malloc(sizeof(double) * d * T);
  for (size_t k = 0; k < (size_t)d * T; ++k) dbm[k] = bm[k + d] - bm[k];
  double* prev = (double*)malloc(sizeof(double) * d * T);
  memcpy(prev, bm, sizeof(double) * d * T);
  const int rc = eigs_from_matrices(prev, dbm, d, T, delta_t, model, out);
  free(bm); free(dbm); free(prev);
  return rc;
}

/* calculate_eigenvalues_parallel (src/data_storage/parallel_compute.rs:14-41): threads over seeds. */
typedef struct { int d, model, nthreads, tid, rc; size_t T, n, ncpu; const uint32_t* seeds; double* out; } job_t;
static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  const int p = (j->model == 1 || j->model == 3) ? j->d + 1 : j->d;
  for (size_t i = j->tid; i < j->n; i += j->nthreads) {
    const int rc = jne_oracle_calculate_eigenvalues(j->d, j->T, j->seeds[i], j->model, j->ncpu, j->out + i * p);
    if (rc) j->rc = rc;
  }
  return NULL;
}
int jne_oracle_eigs_batch(int model, int d, size_t T, const uint32_t* seeds, size_t n, int nthreads, size_t n_physical_cpus, double* out) {
  scipy_openblas_set_num_threads(1);
  if (nthreads < 1) nthreads = 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  job_t* jobs = (job_t*)malloc(sizeof(job_t) * nthreads);
  for (int t = 0; t < nthreads; ++t) {
    jobs[t] = (job_t){d, model, nthreads, t, 0, T, n, n_physical_cpus, seeds, out};
    pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  int rc = 0;
  for (int t = 0; t < nthreads; ++t) { pthread_join(th[t], NULL); if (jobs[t].rc) rc = jobs[t].rc; }
  free(th); free(jobs);
  return rc;
}


/* ------------------------------------------------------------------------------------------------------------
 * "Optimised CPU" variant: NOT the reference's algorithm, a fair best-effort CPU one, so
 * that the GPU speed-up is not flattered by the reference's temporaries.  One pass over the path with raw moments
 * (no d x T buffers), demeaning / detrending as Schur complements (same algebra as the GPU epilogue), LAPACK dsygv
 * (symmetric-definite, eigenvalues only) instead of dggev('V','V').  Same xoshiro256++/ziggurat normals, one stream
 * per run.  Checked against the faithful port through jne_oracle_fast_from_increments (tests/test_oracle.py).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct { double bb[16][16], bz[16][16], sb[16], s1b[16], s2b[16], sz[16], s1z[16], s2z[16]; } mom_t;

static int fast_solve(const mom_t* m, int d, double T, int model, double factor, double* out) {
  const int p = (model == 1 || model == 3) ? d + 1 : d, nb = (model == 2 || model == 4) ? d - 1 : d;
  double S2[17 * 17], R[17][16], A[17 * 17], w[17], work[17 * 40];
  memset(S2, 0, sizeof S2); memset(R, 0, sizeof R);
  const double nu = T * (T * T - 1.0) / 3.0;
  for (int i = 0; i < nb; ++i) {
    for (int j = 0; j < nb; ++j) {
      double v = i <= j ? m->bb[i][j] : m->bb[j][i];
      if (model >= 2) v -= m->sb[i] * m->sb[j] / T;
      if (model == 4) v -= m->s1b[i] * m->s1b[j] / nu;
      S2[j * p + i] = v;
    }
    for (int j = 0; j < d; ++j) {
      double v = m->bz[i][j];
      if (model >= 2) v -= m->sb[i] * m->sz[j] / T;
      if (model == 4) v -= m->s1b[i] * m->s1z[j] / nu;
      R[i][j] = v;
    }
  }
  if (p > nb) {
    for (int j = 0; j < d; ++j) {
      double s2v, rv;
      if (model == 1) { s2v = m->sb[j]; rv = m->sz[j]; }
      else if (model == 4) { s2v = m->s2b[j] / (T * T); rv = m->s2z[j] / (T * T); }
      else { s2v = m->s1b[j] / T; rv = (m->s1z[j] + m->sz[j]) / T; }
      if (j < nb) { S2[nb * p + j] = s2v; S2[j * p + nb] = s2v; }
      R[nb][j] = rv;
    }
    S2[nb * p + nb] = model == 1 ? T : model == 4 ? 0.8 * T * (T * T - 1.0) * (T * T - 4.0) / (T * T * T * T) : (nu + T) / (T * T);
  }
  for (int i = 0; i < p; ++i) for (int j = 0; j < p; ++j) { double v = 0; for (int k = 0; k < d; ++k) v += R[i][k] * R[j][k]; A[j * p + i] = v; }
  const int itype = 1, lwork = 17 * 40; int info = 0, pp = p;
  scipy_dsygv_(&itype, "N", "U", &pp, A, &pp, S2, &pp, w, work, &lwork, &info, 1, 1);
  for (int i = 0; i < p; ++i) out[i] = fabs(w[p - 1 - i]) * factor;
  return info;
}

static void fast_accumulate(mom_t* m, int d, size_t T, const double* dB /* d x T col-major or NULL */, xo_t* rng) {
  memset(m, 0, sizeof *m);
  double B[16] = {0}, z[16];
  const double TT = (double)T, w2c = -(TT * TT - 1.0);
  for (size_t t = 0; t < T; ++t) {
    const double w1 = 2.0 * (double)t + 1.0 - TT, w2 = 3.0 * w1 * w1 + w2c;
    for (int i = 0; i < d; ++i) z[i] = dB ? dB[t * d + i] : standard_normal(rng);
    for (int i = 0; i < d; ++i) {
      const double b = B[i];
      for (int j = i; j < d; ++j) m->bb[i][j] += b * B[j];
      for (int j = 0; j < d; ++j) m->bz[i][j] += b * z[j];
      m->sb[i] += b; m->s1b[i] += w1 * b; m->s2b[i] += w2 * b;
      m->s1z[i] += w1 * z[i]; m->s2z[i] += w2 * z[i];
    }
    for (int i = 0; i < d; ++i) B[i] += z[i];
  }
  for (int i = 0; i < d; ++i) m->sz[i] = B[i];
}

int jne_oracle_fast_from_increments(int model, int d, size_t T, const double* dB, double* out) {
  if (d > 16) return -1;
  mom_t m;
  fast_accumulate(&m, d, T, dB, NULL);
  return fast_solve(&m, d, (double)T, model, (double)T, out);
}

typedef struct { int d, model, nthreads, tid, rc; size_t T, n; const uint32_t* seeds; double* out; } fjob_t;
static void* fast_worker(void* arg) {
  fjob_t* j = (fjob_t*)arg;
  const int p = (j->model == 1 || j->model == 3) ? j->d + 1 : j->d;
  for (size_t i = j->tid; i < j->n; i += j->nthreads) {
    xo_t rng; xo_seed(&rng, (uint64_t)j->seeds[i]);
    mom_t m;
    fast_accumulate(&m, j->d, j->T, NULL, &rng);
    const int rc = fast_solve(&m, j->d, (double)j->T, j->model, 1.0, j->out + i * p);
    if (rc) j->rc = rc;
  }
  return NULL;
}
int jne_oracle_fast_batch(int model, int d, size_t T, const uint32_t* seeds, size_t n, int nthreads, double* out) {
  if (d > 16) return -1;
  pthread_once(&zig_once, zig_init);
  scipy_openblas_set_num_threads(1);
  if (nthreads < 1) nthreads = 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  fjob_t* jobs = (fjob_t*)malloc(sizeof(fjob_t) * nthreads);
  for (int t = 0; t < nthreads; ++t) { jobs[t] = (fjob_t){d, model, nthreads, t, 0, T, n, seeds, out}; pthread_create(&th[t], NULL, fast_worker, &jobs[t]); }
  int rc = 0;
  for (int t = 0; t < nthreads; ++t) { pthread_join(th[t], NULL); if (jobs[t].rc) rc = jobs[t].rc; }
  free(th); free(jobs);
  return rc;
}

/* ------------------------------------------------------------------------------------------------------------
 * Gate-(2) sampler (tools/gate2_cpu_samples.py): for every seed ONE Brownian path from f64 ziggurat normals
 * (xoshiro256++ seeded by the seed, like jne_oracle_fast_batch), the raw moments once, then all five models;
 * per seed and model the two statistics the reference's analysers report (src/simulation_analyzers.rs:25-40):
 * trace = sum of the eigenvalues, max = the largest.  out: n x 10 doubles, [model][trace, max].
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct { int d, nthreads, tid, rc; size_t T, n; const uint32_t* seeds; double* out; } sjob_t;
static void* stats_worker(void* arg) {
  sjob_t* j = (sjob_t*)arg;
  double ev[17];
  for (size_t i = j->tid; i < j->n; i += j->nthreads) {
    xo_t rng; xo_seed(&rng, (uint64_t)j->seeds[i]);
    mom_t m;
    fast_accumulate(&m, j->d, j->T, NULL, &rng);
    for (int model = 0; model < 5; ++model) {
      const int p = (model == 1 || model == 3) ? j->d + 1 : j->d;
      const int rc = fast_solve(&m, j->d, (double)j->T, model, 1.0, ev);
      if (rc) j->rc = rc;
      double tr = 0.0, mx = -1.7976931348623157e308;
      for (int k = 0; k < p; ++k) { tr += ev[k]; if (ev[k] > mx) mx = ev[k]; }
      j->out[i * 10 + 2 * model] = tr;
      j->out[i * 10 + 2 * model + 1] = mx;
    }
  }
  return NULL;
}
int jne_oracle_fast_multi_stats(int d, size_t T, const uint32_t* seeds, size_t n, int nthreads, double* out) {
  if (d > 16) return -1;
  pthread_once(&zig_once, zig_init);
  scipy_openblas_set_num_threads(1);
  if (nthreads < 1) nthreads = 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  sjob_t* jobs = (sjob_t*)malloc(sizeof(sjob_t) * nthreads);
  for (int t = 0; t < nthreads; ++t) { jobs[t] = (sjob_t){d, nthreads, t, 0, T, n, seeds, out}; pthread_create(&th[t], NULL, stats_worker, &jobs[t]); }
  int rc = 0;
  for (int t = 0; t < nthreads; ++t) { pthread_join(th[t], NULL); if (jobs[t].rc) rc = jobs[t].rc; }
  free(th); free(jobs);
  return rc;
}
