// Lane family: one THREAD per run, register-resident FP64 moments (sm_100a).  Serves dim <= 6.
//
// Why a second family.  On the tensor path (jne_kernels.cuh) a warp owns one run and feeds V = [F ; dB] to
// mma.sync.m8n8k4.f64 tiles; the tile layouts are sized for 4, 8 and 12 rows, so dim 1 pays for dim 4 and
// dim 5 for dim 8 -- both in padded products and in normals generated for rows that do not exist
// (profiles/r1_bench_all_configs.jsonl: 10.5 M seeds/s for every dim 1..4, 7.2 M for dim 5 and 6).  The reference's
// cost is proportional to the products it needs (src/matrix_utils.rs:67-85: p x d per step), and below about
// seven rows all of a run's moments fit one thread's registers:
//     sum c c' (upper triangle)   D (D + 1) / 2        sum c dB'   D^2        sum c, sum w1 c, sum w2 c   3 D
// = 81 doubles at D = 6.  So here a lane owns a whole run: it generates exactly the D rows the run has (one generator
// call per row and four steps), keeps the path in registers, and every product is an FP64 FMA on two registers:
// no padded products, no operand exchange, no time segments (the sums run left to right over the T steps, as the
// reference's do), D known at compile time.  The moments go to global memory in a compact layout
// (JneLaneMom<D>) and jne_lane_solve_kernel -- one warp per run, the epilogue of the tensor path unchanged
// (jne_warp_models: assembly per model, elimination, Jacobi) -- turns them into eigenvalues.
//
//   K1  jne_sub_normals4          (jne_rng.cuh)  replaces gen_normal_matrix        src/rng_matrix.rs:11-37
//   K2  jne_lane_moments_kernel                  replaces brownian_motion_matrix   src/rng_matrix.rs:57-141,
//                                                dmatrix_cumsum RowWise            src/matrix_utils.rs:51-63,
//                                                construct_f_matrix                src/johansen_statistics.rs:102-197,
//                                                2 x sum_of_outer_products         src/matrix_utils.rs:67-85
//   K3  jne_lane_solve_kernel                    replaces GeneralizedEigen::new + |alpha|/beta + sort
//                                                                                  src/johansen_statistics.rs:35-46
// The random stream is the same function of (seed, row, step) as everywhere else.
#pragma once
#include <type_traits>
#include "jne_kernels.cuh"

#define JNE_LANE_MAX_DIM 6
// Launch geometry: 128 threads per CTA (one warp per SM sub-partition) and the resident CTAs per SM the register
// budget is sized for.  Registers are allocated per sub-partition (16 384 each): 6 / 5 / 4 / 3 / 2 resident warps
// leave 85 / 102 / 128 / 168 / 255 registers per thread.  A run carries 8 D registers of generator state (two
// substreams per row, jne_rng.cuh) next to its moments.
#ifndef JNE_LANE_MINB5
#define JNE_LANE_MINB5 2      // dim 5: 255 registers, 2 resident CTAs per SM; 3 = 168 registers (measured equal without
#endif                        // the software pipeline, whose second block of normals needs the room)
#ifndef JNE_LANE_REG_STATES
#define JNE_LANE_REG_STATES 5   // generator states in registers up to this dim, in shared memory above (dim 6)
#endif
#ifndef JNE_LANE_PIPELINE
#define JNE_LANE_PIPELINE 1   // 0: generate a block, then consume it (regression / ablation)
#endif
#ifndef JNE_LANE_MINB3
#define JNE_LANE_MINB3 4      // dim 3: 128 registers (5 CTAs = 102 registers spill the substream states: 27.3 -> 31.0 M seeds/s)
#endif
#ifndef JNE_LANE_MINB4
#define JNE_LANE_MINB4 3      // dim 4: 168 registers (17.9 -> 19.6 M seeds/s)
#endif
template <int D> struct JneLaneGeo {
  static constexpr int THREADS = 128;
  static constexpr int MINB = D <= 2 ? 6 : D == 3 ? JNE_LANE_MINB3 : D == 4 ? JNE_LANE_MINB4 : D == 5 ? JNE_LANE_MINB5 : 2;
};

// Per-run moments in global memory (doubles):
//   bb [D (D + 1) / 2]  sum c_i c_j, i <= j, row-major upper triangle        (c = path BEFORE the step = F rows)
//   bz [D][D]           sum c_i dB_j
//   tot[6][D]           0 sum c, 1 sum w1 c, 2 sum w2 c, 3 sum dB, 4 sum w1 dB, 5 sum w2 dB
// with w1_t = 2 t + 1 - T and w2_t = 3 w1_t^2 - (T^2 - 1): the integer forms of the reference's trend regressors
// (t + 1) / T - 1/2 and the residual of ((t + 1) / T)^2 on [1, (t + 1) / T] (src/johansen_statistics.rs:127-135,170-194).
template <int D> struct JneLaneMom {
  static constexpr int NBB = D * (D + 1) / 2;
  static constexpr int OFF_BZ = NBB, OFF_TOT = NBB + D * D, SZ = OFF_TOT + 6 * D;
};
__host__ __device__ constexpr int jne_lane_mom_size(int d) { return d * (d + 1) / 2 + d * d + 6 * d; }

// DET: 0 = models 0, 1 (sum c only), 1 = models 2, 3 (+ w1 moments), 2 = model 4 / several models (+ w2 moments).
template <int D, int DET> struct JneLaneState {
  double c[D], bb[JneLaneMom<D>::NBB], bz[D * D], s0[D], s1[D], s2[D];
  double w1, w2, w2c;
};

// One step: the products of the path before the step with itself and with the step's increments, the trend sums,
// then B_t = B_{t-1} + dB_t (src/matrix_utils.rs:51-63).
template <int D, int DET, bool SRC_RNG>
__device__ __forceinline__ void jne_lane_step(JneLaneState<D, DET>& S, const double (&zin)[D]) {
  double dz[D], cn[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    cn[i] = S.c[i] + zin[i];
    dz[i] = SRC_RNG ? zin[i] : cn[i] - S.c[i];   // caller increments: dB re-derived by subtraction (src/johansen_statistics.rs:80-82)
  }
  int idx = 0;
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = i; j < D; ++j) { S.bb[idx] = fma(S.c[i], S.c[j], S.bb[idx]); ++idx; }
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) S.bz[i * D + j] = fma(S.c[i], dz[j], S.bz[i * D + j]);
#pragma unroll
  for (int i = 0; i < D; ++i) {
    S.s0[i] += S.c[i];
    if (DET >= 1) S.s1[i] = fma(S.w1, S.c[i], S.s1[i]);
    if (DET >= 2) S.s2[i] = fma(S.w2, S.c[i], S.s2[i]);
    S.c[i] = cn[i];
  }
  if (DET >= 1) S.w1 += 2.0;
  if (DET >= 2) S.w2 = fma(3.0 * S.w1, S.w1, S.w2c);
}

template <int D, int DET, bool SRC_RNG>
__global__ void __launch_bounds__(JneLaneGeo<D>::THREADS, JneLaneGeo<D>::MINB)
jne_lane_moments_kernel(const uint32_t* __restrict__ seeds, const double* __restrict__ dB, uint64_t n, uint32_t T,
                        double* __restrict__ mom) {
  using M = JneLaneMom<D>;
  const uint64_t run_raw = (uint64_t)blockIdx.x * JneLaneGeo<D>::THREADS + threadIdx.x;
  const bool live = run_raw < n;
  const uint64_t run = live ? run_raw : n - 1;         // idle lanes shadow the last run (no divergence in the loop)
  // generator state: both substreams (halves) of every row of the current epoch, jne_rng.cuh
  const uint32_t seed = SRC_RNG ? seeds[run] : 0u;
  // D <= JNE_LANE_REG_STATES: in registers.  Above (dim 6: 81 moments, the path and 48 state words do not fit 255
  // registers; ptxas spilled 72 bytes of the hot loop), the states live in shared memory, one 16-byte word per thread
  // and state: a conflict-free LDS.128 / STS.128 pair per generator call.
  constexpr bool ST_SMEM = SRC_RNG && D > JNE_LANE_REG_STATES && sizeof(jne_sub) == 16;
  __shared__ uint4 st_sm[ST_SMEM ? 2 * D : 1][ST_SMEM ? JneLaneGeo<D>::THREADS : 1];
  jne_sub st[ST_SMEM ? 1 : D][2];
  auto seed_all = [&](uint32_t epoch) {
#pragma unroll
    for (int r = 0; r < D; ++r) {
      if constexpr (ST_SMEM) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          jne_sub x;
          jne_sub_seed(x, seed, (uint32_t)r, epoch, (uint32_t)h);
          st_sm[2 * r + h][threadIdx.x] = make_uint4(x.s0, x.s1, x.s2, x.s3);
        }
      } else {
        jne_sub_seed(st[r][0], seed, (uint32_t)r, epoch, 0u);
        jne_sub_seed(st[r][1], seed, (uint32_t)r, epoch, 1u);
      }
    }
  };
  auto draw4 = [&](int r, int h, jne_zt* z) {            // the next four-step block of row r, half h (constants after unrolling)
    if constexpr (ST_SMEM) {
      const uint4 v = st_sm[2 * r + h][threadIdx.x];
      jne_sub x;
      x.s0 = v.x; x.s1 = v.y; x.s2 = v.z; x.s3 = v.w;
      jne_sub_normals4(x, z);
      st_sm[2 * r + h][threadIdx.x] = make_uint4(x.s0, x.s1, x.s2, x.s3);
    } else {
      jne_sub_normals4(st[r][h], z);
    }
  };
  const double* dBrun = SRC_RNG ? nullptr : dB + run * (uint64_t)D * T;

  JneLaneState<D, DET> S;
#pragma unroll
  for (int i = 0; i < D; ++i) { S.c[i] = S.s0[i] = S.s1[i] = S.s2[i] = 0.0; }
#pragma unroll
  for (int i = 0; i < M::NBB; ++i) S.bb[i] = 0.0;
#pragma unroll
  for (int i = 0; i < D * D; ++i) S.bz[i] = 0.0;
  const double Td = (double)T;
  const double w1_first = 1.0 - Td;
  const double w2c = -(Td * Td - 1.0);
  S.w1 = w1_first;
  S.w2c = w2c;
  S.w2 = fma(3.0 * w1_first, w1_first, w2c);

  // Four steps per generator call and row.  The ragged tail (T mod 4 steps) is uniform over the grid: T is a launch
  // parameter.
  const uint32_t nfull = T >> 2, tail = T & 3u;
  if constexpr (SRC_RNG && JNE_LANE_PIPELINE) {
    // Software pipeline: the normals of block b + 1 are generated BETWEEN the steps of block b (rows spread over the
    // four steps), so that every stretch of the instruction stream carries both a generator / MUFU dependency chain
    // and a batch of independent FP64 FMAs.  With two or three resident warps per sub-partition the generator's latency
    // is otherwise exposed (ncu: 55 % issue-slot use, FP64 and XU pipes 45 / 42 % busy, `wait` the top stall).
    // Blocks alternate between the two substreams of a row (half = parity of the block), so the loop handles an even
    // and an odd block per iteration and every state index is a compile-time constant.
    jne_zt zc[D][4], zn[D][4];
    seed_all(0u);
#pragma unroll
    for (int r = 0; r < D; ++r) draw4(r, 0, zc[r]);
    // block tb (parity H) is consumed from zc while block tb + 1 (parity 1 - H) is generated into zn
    auto block = [&](auto Hc, uint32_t tb) {
      constexpr int H = decltype(Hc)::value;
      if (H == 1 && ((tb + 1u) & (JNE_EPOCH_BLOCKS - 1u)) == 0u) seed_all((tb + 1u) / JNE_EPOCH_BLOCKS);   // next block opens an epoch
#pragma unroll
      for (int s = 0; s < 4; ++s) {
#pragma unroll
        for (int r = 0; r < D; ++r)
          if (r * 4 / D == s) draw4(r, 1 - H, zn[r]);
        double zz[D];
#pragma unroll
        for (int r = 0; r < D; ++r) zz[r] = (double)zc[r][s];
        jne_lane_step<D, DET, SRC_RNG>(S, zz);
      }
#pragma unroll
      for (int r = 0; r < D; ++r)
#pragma unroll
        for (int s = 0; s < 4; ++s) zc[r][s] = zn[r][s];
    };
    uint32_t tb = 0;
    for (; tb + 1u < nfull; tb += 2u) {
      block(std::integral_constant<int, 0>{}, tb);
      block(std::integral_constant<int, 1>{}, tb + 1u);
    }
    if (tb < nfull) block(std::integral_constant<int, 0>{}, tb);
#pragma unroll 1
    for (uint32_t s = 0; s < tail; ++s) {            // zc holds block nfull
      double zz[D];
#pragma unroll
      for (int r = 0; r < D; ++r) zz[r] = (double)(s == 0u ? zc[r][0] : s == 1u ? zc[r][1] : zc[r][2]);
      jne_lane_step<D, DET, SRC_RNG>(S, zz);
    }
  } else {
    auto load_block = [&](uint32_t tb, uint32_t ns, double (&z)[4][D]) {
      if constexpr (SRC_RNG) {
#pragma unroll
        for (int r = 0; r < D; ++r) {
          jne_zt zf[4];
          jne_normals4(seed, (uint32_t)r, tb, zf);     // random access (ablation builds only: JNE_LANE_PIPELINE=0)
#pragma unroll
          for (int s = 0; s < 4; ++s) z[s][r] = (double)zf[s];
        }
      } else {
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
          for (int r = 0; r < D; ++r) z[s][r] = (uint32_t)s < ns ? dBrun[(uint64_t)(4u * tb + s) * D + r] : 0.0;
      }
    };
    for (uint32_t tb = 0; tb < nfull; ++tb) {
      double z[4][D];
      load_block(tb, 4u, z);
      jne_lane_step<D, DET, SRC_RNG>(S, z[0]);
      jne_lane_step<D, DET, SRC_RNG>(S, z[1]);
      jne_lane_step<D, DET, SRC_RNG>(S, z[2]);
      jne_lane_step<D, DET, SRC_RNG>(S, z[3]);
    }
    if (tail != 0u) {
      double z[4][D];
      load_block(nfull, tail, z);
#pragma unroll 1
      for (uint32_t s = 0; s < tail; ++s) {
        double zz[D];
#pragma unroll
        for (int r = 0; r < D; ++r) zz[r] = s == 0u ? z[0][r] : s == 1u ? z[1][r] : z[2][r];
        jne_lane_step<D, DET, SRC_RNG>(S, zz);
      }
    }
  }
  if (!live) return;

  double* m = mom + run * (uint64_t)M::SZ;
#pragma unroll
  for (int i = 0; i < M::NBB; ++i) m[i] = S.bb[i];
#pragma unroll
  for (int i = 0; i < D * D; ++i) m[M::OFF_BZ + i] = S.bz[i];
  // the increment moments by summation by parts (dB_t = c_{t+1} - c_t, c_0 = 0, w1_t - w1_{t-1} = 2,
  // w2_t - w2_{t-1} = 12 w1_t - 12):  sum w1 dB = w1_last c_T - 2 sum c,  sum w2 dB = w2_last c_T - 12 sum w1 c + 12 sum c
  const double w1_last = w1_first + 2.0 * (Td - 1.0);
  const double w2_last = fma(3.0 * w1_last, w1_last, w2c);
#pragma unroll
  for (int i = 0; i < D; ++i) {
    double* t = m + M::OFF_TOT + i;
    t[0 * D] = S.s0[i];
    t[1 * D] = S.s1[i];
    t[2 * D] = S.s2[i];
    t[3 * D] = S.c[i];
    t[4 * D] = fma(w1_last, S.c[i], -2.0 * S.s0[i]);
    t[5 * D] = fma(w2_last, S.c[i], 12.0 * (S.s0[i] - S.s1[i]));
  }
}

// One warp per run: the moments of jne_lane_moments_kernel in the stitched layout of jne_warp_stitch, then the
// epilogue of the tensor path (per-model assembly with demeaning / detrending as Schur complements, elimination,
// Jacobi, sort).  DP = 4 (dim <= 4) or 8 sizes the work matrices.
template <int DP, bool MULTI>
__global__ void __launch_bounds__(32 * JNE_WARPS_PER_CTA)
jne_lane_solve_kernel(const double* __restrict__ mom, uint64_t n, JneRunParams prm, double* __restrict__ out,
                      unsigned int* __restrict__ err_count, double* __restrict__ dbg) {
  using G = JneGeo<DP>;
  using E = JneEpi<DP, MULTI ? 5 : 1>;
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint64_t run = (uint64_t)blockIdx.x * JNE_WARPS_PER_CTA + warp;
  if (run >= n) return;
  double* wsm = smem + (size_t)warp * E::END;
  double* tot = wsm;
  double* MBB = tot + G::TOT_SZ;
  double* MBZ = MBB + G::STITCH_HALF;
  const int d = prm.dim;
  const int nbb = d * (d + 1) / 2;
  const double* M = mom + run * (uint64_t)jne_lane_mom_size(d);
  for (int e = lane; e < G::STITCH_HALF; e += 32) {
    const int i = e >> 4, j = e & 15;
    const bool in = i < d && j < d;
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    MBB[e] = in ? M[lo * d - ((lo * (lo - 1)) >> 1) + (hi - lo)] : 0.0;
    MBZ[e] = in ? M[nbb + i * d + j] : 0.0;
  }
  for (int e = lane; e < 96; e += 32) {
    const int k = e >> 4, i = e & 15;
    tot[e] = i < d ? M[nbb + d * d + k * d + i] : 0.0;
  }
  __syncwarp();
  const bool ok = jne_warp_models<DP, MULTI ? 5 : 1>(wsm, prm, out + run * prm.out_stride,
                                                     dbg != nullptr ? dbg + run * 512 : nullptr);
  if (!ok && lane == 0) atomicAdd(err_count, 1u);
}

template <int DP, bool MULTI> constexpr size_t jne_lane_solve_smem() {
  return (size_t)JNE_WARPS_PER_CTA * sizeof(double) * JneEpi<DP, MULTI ? 5 : 1>::END;
}

// ---------------------------------------------------------------------------------------------
// K3 for the lane family: one THREAD per run (dim <= 6), everything in registers.
//
// The warp-per-run epilogue of the tensor path costs ~7 000 warp instructions per run at dim 5 whatever the size of the
// problem (ncu: issue-bound, 17 % of the c2 step next to the lane moments kernel); a 6 x 6 problem does not need 32
// lanes.  Here a thread assembles S2 = sum F F' and R = S1' = sum F dB' of one model from the run's moments exactly as
// jne_warp_assemble does (demeaning / detrending as Schur complements, src/johansen_statistics.rs:102-197), reduces the
// pencil (S1'S1, S2) to the d x d Gram matrix G = R' S2^-1 R by a Cholesky factorisation and a forward substitution,
// and diagonalises G by cyclic Jacobi -- all loops unrolled over the compile-time D, every array in registers.
// Replaces GeneralizedEigen::new (LAPACK dggev) + |alpha| / beta + sort, src/johansen_statistics.rs:35-46.
// Every model is computed on its own, so a block of a fused launch is bit-identical to the single-model launch.
// Models whose F has D rows are embedded in the (D + 1)-row problem by a decoupled unit row (one code path).
// ---------------------------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ bool jne_thread_model(const double* __restrict__ M, int model, double T, double factor,
                                                 double* __restrict__ out) {
  constexpr int P = D + 1, NBB = D * (D + 1) / 2;
  const double* bb = M;
  const double* bz = M + NBB;
  const double* tot = M + NBB + D * D;
  double SB[D], S1B[D], S2B[D], Sz[D], S1z[D], S2z[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    SB[i] = tot[0 * D + i]; S1B[i] = tot[1 * D + i]; S2B[i] = tot[2 * D + i];
    Sz[i] = tot[3 * D + i]; S1z[i] = tot[4 * D + i]; S2z[i] = tot[5 * D + i];
  }
  const double invT = 1.0 / T;
  const double nu = T * (T * T - 1.0) / 3.0;               // sum w1^2
  const double inv_nu = 1.0 / nu;                          // inf at T = 1 (model 4 needs T >= 3)
  const bool demean = model >= 2, detrend = model == 4;
  const bool trim = model == 2 || model == 4;              // F keeps D - 1 Brownian rows; the deterministic row is row D - 1
  const bool extra = model == 1 || model == 3;             // F has D + 1 rows; the deterministic row is row D

  double S2[P][P], R[P][D];                                // S2: lower triangle (i >= j)
#pragma unroll
  for (int i = 0; i < D; ++i) {
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      double v = bb[j * D - ((j * (j - 1)) >> 1) + (i - j)];
      if (demean) v -= SB[i] * SB[j] * invT;               // src/johansen_statistics.rs:120-125,145-150,186-194
      if (detrend) v -= S1B[i] * S1B[j] * inv_nu;          // residual on [1, tau]   :186-194
      S2[i][j] = v;
    }
#pragma unroll
    for (int j = 0; j < D; ++j) {
      double v = bz[i * D + j];
      if (demean) v -= SB[i] * Sz[j] * invT;
      if (detrend) v -= S1B[i] * S1z[j] * inv_nu;
      R[i][j] = v;
    }
  }
  // deterministic row: constant (model 1, :108-113), trend (i+1)/T - 1/2 = (w1+1)/(2T) carried as (w1+1)/T (models 2, 3,
  // :127-135,152-160; NOT demeaned), residual of tau^2 on [1, tau] = w2 / (12 T^2) carried as w2 / T^2 (model 4, :170-194).
  // Row scalings of F do not change the pencil's eigenvalues.
  double s2v[D], rv[D], dg;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    s2v[j] = model == 1 ? SB[j] : model == 4 ? S2B[j] * invT * invT : S1B[j] * invT;
    rv[j] = model == 1 ? Sz[j] : model == 4 ? S2z[j] * invT * invT : (S1z[j] + Sz[j]) * invT;
  }
  dg = model == 1 ? T : model == 4 ? 0.8 * T * (T * T - 1.0) * (T * T - 4.0) * invT * invT * invT * invT
                                   : (nu + T) * invT * invT;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    if (trim) {
      if (j < D - 1) S2[D - 1][j] = s2v[j]; else S2[D - 1][D - 1] = dg;
      R[D - 1][j] = rv[j];
    }
    S2[D][j] = extra ? s2v[j] : 0.0;
    R[D][j] = extra ? rv[j] : 0.0;
  }
  S2[D][D] = extra ? dg : 1.0;

  // Cholesky S2 = L L' (left-looking; pivots through a wide-range rsqrt so that caller increments of any scale work,
  // NaN for a non-positive pivot -> flagged below), W = L^-1 R, G = W' W.
  double W[P][D];
#pragma unroll
  for (int j = 0; j < P; ++j) {
    double dj = S2[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) dj = fma(-S2[j][k], S2[j][k], dj);
    const double inv = jne_rsqrt_wide(dj);
#pragma unroll
    for (int i = j + 1; i < P; ++i) {
      double v = S2[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) v = fma(-S2[i][k], S2[j][k], v);
      S2[i][j] = v * inv;
    }
#pragma unroll
    for (int c = 0; c < D; ++c) {
      double v = R[j][c];
#pragma unroll
      for (int k = 0; k < j; ++k) v = fma(-S2[j][k], W[k][c], v);
      W[j][c] = v * inv;
    }
  }
  double G[D][D];                                          // upper triangle (a <= b)
  double tr = 0.0;
#pragma unroll
  for (int a = 0; a < D; ++a)
#pragma unroll
    for (int b = a; b < D; ++b) {
      double v = 0.0;
#pragma unroll
      for (int j = 0; j < P; ++j) v = fma(W[j][a], W[j][b], v);
      G[a][b] = v;
      if (a == b) tr += v;
    }
  const double inv_tr = 1.0 / tr;                          // unit trace: the rotation thresholds below are absolute
#pragma unroll
  for (int a = 0; a < D; ++a)
#pragma unroll
    for (int b = a; b < D; ++b) G[a][b] *= inv_tr;

  // Cyclic Jacobi, rotation (p, q) applied as the exact similarity J' G J for the (c, s) actually used: tan(theta) comes
  // from FP32 arithmetic (it only steers convergence), (c, s) is normalised in FP64.  Same skip rule as jne_warp_jacobi.
  if (D > 1) {
    const double tol = 8.8817841970012523e-16;
    for (int sweep = 0; sweep < 30; ++sweep) {
      bool rotated = false;
#pragma unroll
      for (int p = 0; p < D - 1; ++p)
#pragma unroll
        for (int q = p + 1; q < D; ++q) {
          const double app = G[p][p], aqq = G[q][q], apq = G[p][q];
          const double diff = aqq - app;
          if (fabs(apq) > tol && apq * apq > 1e-14 * fabs(diff) * fmin(app, aqq)) {
            float th, hy, tf;
            asm("div.approx.ftz.f32 %0, %1, %2;" : "=f"(th) : "f"((float)diff), "f"(2.0f * (float)apq));
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(hy) : "f"(fmaf(th, th, 1.0f)));
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(tf) : "f"(fabsf(th) + hy));
            const double t = (double)copysignf(tf, th);
            const double c = jne_rsqrt(fma(t, t, 1.0)), s = t * c;
            rotated = true;
            const double cc = c * c, ss = s * s, cs2 = 2.0 * c * s;
            G[p][p] = fma(cc, app, fma(-cs2, apq, ss * aqq));
            G[q][q] = fma(ss, app, fma(cs2, apq, cc * aqq));
            G[p][q] = fma(c * s, app - aqq, (cc - ss) * apq);
#pragma unroll
            for (int k = 0; k < D; ++k) {
              if (k == p || k == q) continue;
              double& gkp = k < p ? G[k][p] : G[p][k];
              double& gkq = k < q ? G[k][q] : G[q][k];
              const double x = gkp, y = gkq;
              gkp = fma(c, x, -s * y);
              gkq = fma(s, x, c * y);
            }
          }
        }
      if (!rotated) break;
    }
  }
  // lambda_i = factor * trace * |g_ii| (src/johansen_statistics.rs:40-44: |alpha| / beta), sorted descending (:45) by
  // rank counting; models 1 and 3 carry one extra eigenvalue that is exactly 0 here (rank-D pencil).
  double ev[P];
  const double scale = factor * tr;
  bool finite = true;
#pragma unroll
  for (int k = 0; k < D; ++k) { ev[k] = fabs(G[k][k]) * scale; finite &= isfinite(ev[k]); }
  ev[D] = 0.0 * scale;                                     // 0 * NaN keeps a failure visible
  const int np = extra ? P : D;
#pragma unroll
  for (int k = 0; k < P; ++k) {
    if (k < np) {
      int rank = 0;
#pragma unroll
      for (int j = 0; j < P; ++j)
        if (j < np) rank += (ev[j] > ev[k]) || (ev[j] == ev[k] && j < k);
      out[finite ? rank : k] = ev[k];                      // NaN compares false everywhere: keep the slots distinct
    }
  }
  return finite;
}

template <int D>
__global__ void __launch_bounds__(128)
jne_lane_tsolve_kernel(const double* __restrict__ mom, uint64_t n, uint32_t model_mask, double T, double factor,
                       uint32_t out_stride, double* __restrict__ out, unsigned int* __restrict__ err_count) {
  const uint64_t run = (uint64_t)blockIdx.x * 128 + threadIdx.x;
  if (run >= n) return;
  const double* M = mom + run * (uint64_t)JneLaneMom<D>::SZ;
  double* o = out + run * out_stride;
  bool ok = true;
#pragma unroll 1
  for (int model = 0; model < 5; ++model) {
    if (!((model_mask >> model) & 1u)) continue;
    ok &= jne_thread_model<D>(M, model, T, factor, o);
    o += (model == 1 || model == 3) ? D + 1 : D;
  }
  if (!ok) atomicAdd(err_count, 1u);
}

// ---------------------------------------------------------------------------------------------
// Group kernels: L lanes per run, R = D / L rows per lane (in use: dim 9 as 3 x 3).
//
// Above six rows a run's moments no longer fit one thread, and on the tensor path dims 9..11 pay for dim 12's five
// DMMA tiles (profiles/r2_bench_all_configs_v2.jsonl: 3.28 M seeds/s at dim 9 against 3.19 M at dim 12).  Here the
// D rows of a run are dealt to L lanes; a lane generates the normals of its own R rows and owns the products that have
// one of its rows as the left factor:
//     sum c_own dB'_all   R x D        sum c_own c_own' (upper)   R (R + 1) / 2
//     sum c_own c'_partner for the next floor((L - 1) / 2) lanes of the group (full R x R blocks)
//     L even: the half block a <= b with the lane opposite (both lanes compute its diagonal: identical bits)
// so every unordered pair of rows is accumulated exactly once (or twice with the same bits), with no padded products.
// Every lane keeps the WHOLE path c (D adds per step) in a frame rotated by its own first row -- own rows first,
// then the next lane's, ... -- which makes every register index a compile-time constant while the lane dependence
// sits in shared-memory addresses.  The only exchange is the step's increments: each lane writes the four steps of
// its R rows of a four-step block to shared memory once per block (double-buffered, one __syncwarp per block) and reads
// the D increments of a step back as broadcasts; the path itself never travels.  Moments leave in the layout of
// JneLaneMom<D>; jne_lane_solve_kernel (one warp per run) finishes.
// ---------------------------------------------------------------------------------------------
template <int D, int L> struct JneGroupGeo {
  static constexpr int R = (D + L - 1) / L;          // rows per lane
  static constexpr int LR = L * R;                   // rows incl. padding (== D for the shapes in use)
  static constexpr int G = 32 / L;                   // runs per warp (lanes >= G * L shadow the last group)
  static constexpr int WARPS = 4, THREADS = 128;
  static constexpr int RUNS_PER_CTA = WARPS * G;
  static constexpr int NFULL = (L - 1) / 2;          // partners l + 1 .. l + NFULL: full R x R blocks
  static constexpr bool HALF = (L % 2) == 0 && L > 1;   // partner l + L / 2: the half block a <= b
  static constexpr int NOWN = R * (R + 1) / 2;
  static constexpr int ZS = 2 * 4 * LR * G;          // doubles of shared memory per warp (two blocks of four steps)
};

template <int D, int L, int DET> struct JneGroupState {
  using Q = JneGroupGeo<D, L>;
  double c[Q::LR];                                   // the whole path, rotated frame: c[k] = row (l R + k) mod LR
  double bz[Q::R][Q::LR], bown[Q::NOWN], bfull[Q::NFULL > 0 ? Q::NFULL : 1][Q::R][Q::R], bhalf[Q::NOWN];
  double s0[Q::R], s1[Q::R], s2[Q::R];
  double w1, w2, w2c;
};

template <int D, int L, int DET, bool SRC_RNG>
__device__ __forceinline__ void jne_group_step(JneGroupState<D, L, DET>& S, const double (&zin)[JneGroupGeo<D, L>::LR]) {
  using Q = JneGroupGeo<D, L>;
  double dz[Q::LR], cn[Q::LR];
#pragma unroll
  for (int k = 0; k < Q::LR; ++k) {
    cn[k] = S.c[k] + zin[k];                         // B_t = B_{t-1} + dB_t (src/matrix_utils.rs:51-63), every row, every lane
    dz[k] = SRC_RNG ? zin[k] : cn[k] - S.c[k];       // caller increments: re-derived by subtraction (src/johansen_statistics.rs:80-82)
  }
#pragma unroll
  for (int a = 0; a < Q::R; ++a) {
#pragma unroll
    for (int k = 0; k < Q::LR; ++k) S.bz[a][k] = fma(S.c[a], dz[k], S.bz[a][k]);
  }
  {
    int idx = 0;
#pragma unroll
    for (int a = 0; a < Q::R; ++a)
#pragma unroll
      for (int b = a; b < Q::R; ++b) {
        S.bown[idx] = fma(S.c[a], S.c[b], S.bown[idx]);
        if (Q::HALF) S.bhalf[idx] = fma(S.c[a], S.c[(L / 2) * Q::R + b], S.bhalf[idx]);
        ++idx;
      }
  }
#pragma unroll
  for (int p = 0; p < Q::NFULL; ++p)
#pragma unroll
    for (int a = 0; a < Q::R; ++a)
#pragma unroll
      for (int b = 0; b < Q::R; ++b) S.bfull[p][a][b] = fma(S.c[a], S.c[(p + 1) * Q::R + b], S.bfull[p][a][b]);
#pragma unroll
  for (int a = 0; a < Q::R; ++a) {
    S.s0[a] += S.c[a];
    if (DET >= 1) S.s1[a] = fma(S.w1, S.c[a], S.s1[a]);
    if (DET >= 2) S.s2[a] = fma(S.w2, S.c[a], S.s2[a]);
  }
#pragma unroll
  for (int k = 0; k < Q::LR; ++k) S.c[k] = cn[k];
  if (DET >= 1) S.w1 += 2.0;
  if (DET >= 2) S.w2 = fma(3.0 * S.w1, S.w1, S.w2c);
}

template <int D, int L, int DET, bool SRC_RNG>
__global__ void __launch_bounds__(JneGroupGeo<D, L>::THREADS, 2)
jne_group_moments_kernel(const uint32_t* __restrict__ seeds, const double* __restrict__ dB, uint64_t n, uint32_t T,
                         double* __restrict__ mom) {
  using Q = JneGroupGeo<D, L>;
  using M = JneLaneMom<D>;
  constexpr int R = Q::R, LR = Q::LR, G = Q::G;
  __shared__ double zs_all[Q::WARPS][Q::ZS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool shadow = lane >= G * L;                   // spare lanes repeat the last group (identical values, no stores to global)
  const int g = shadow ? G - 1 : lane / L, l = lane % L;
  const uint64_t run_raw = ((uint64_t)blockIdx.x * Q::WARPS + warp) * G + g;
  const bool live = run_raw < n && !shadow;
  const uint64_t run = run_raw < n ? run_raw : n - 1;
  double* zs = zs_all[warp];
  // generator state: both substreams (halves) of this lane's R rows for the current epoch, jne_rng.cuh
  const uint32_t seed = SRC_RNG ? seeds[run] : 0u;
  // (registers: the same states in shared memory, as in the dim-6 lane kernel, cost this kernel 9 %)
  jne_sub st[R][2];
  auto seed_all = [&](uint32_t epoch) {
#pragma unroll
    for (int a = 0; a < R; ++a) {
      jne_sub_seed(st[a][0], seed, (uint32_t)(l * R + a), epoch, 0u);
      jne_sub_seed(st[a][1], seed, (uint32_t)(l * R + a), epoch, 1u);
    }
  };
  const double* dBrun = SRC_RNG ? nullptr : dB + run * (uint64_t)D * T;

  JneGroupState<D, L, DET> S;
#pragma unroll
  for (int k = 0; k < LR; ++k) S.c[k] = 0.0;
#pragma unroll
  for (int a = 0; a < R; ++a) {
    S.s0[a] = S.s1[a] = S.s2[a] = 0.0;
#pragma unroll
    for (int k = 0; k < LR; ++k) S.bz[a][k] = 0.0;
#pragma unroll
    for (int p = 0; p < (Q::NFULL > 0 ? Q::NFULL : 1); ++p)
#pragma unroll
      for (int b = 0; b < R; ++b) S.bfull[p][a][b] = 0.0;
  }
#pragma unroll
  for (int i = 0; i < Q::NOWN; ++i) { S.bown[i] = 0.0; S.bhalf[i] = 0.0; }
  const double Td = (double)T;
  const double w1_first = 1.0 - Td;
  const double w2c = -(Td * Td - 1.0);
  S.w1 = w1_first;
  S.w2c = w2c;
  S.w2 = fma(3.0 * w1_first, w1_first, w2c);

  // shared-memory addresses: element (buffer, step s, row, group) at ((buffer * 4 + s) * LR + row) * G + group
  int rd_off[LR];                                      // this lane's rotated frame: k -> row (l R + k) mod LR
#pragma unroll
  for (int k = 0; k < LR; ++k) rd_off[k] = ((l * R + k) % LR) * G + g;
  const int wr_off = (l * R) * G + g;

  // writes the four steps of row a of this lane of block tb (the next block of substream tb & 1 = BUF, or the caller's
  // increments) into buffer BUF
  auto publish_row = [&](auto Bc, uint32_t tb, int a) {
    constexpr int BUF = decltype(Bc)::value;
    const int row = l * R + a;
    double v[4];
    if constexpr (SRC_RNG) {
      jne_zt zf[4];
      jne_sub_normals4(st[a][BUF], zf, row < D ? 1.0f : 0.0f);
#pragma unroll
      for (int s = 0; s < 4; ++s) v[s] = (double)zf[s];
    } else {
#pragma unroll
      for (int s = 0; s < 4; ++s) v[s] = (row < D && 4u * tb + s < T) ? dBrun[(uint64_t)(4u * tb + s) * D + row] : 0.0;
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) zs[((BUF * 4 + s) * LR + a) * G + wr_off] = v[s];
  };
  auto read_step = [&](int buf, int s, double (&zz)[LR]) {
#pragma unroll
    for (int k = 0; k < LR; ++k) zz[k] = zs[(buf * 4 + s) * LR * G + rd_off[k]];
  };

  // Schedule of a block tb (its four steps lie complete in buffer tb & 1): block tb + 1 is generated and published
  // during steps 0..2, every step's increments are read one step ahead of their use (the shared-memory latency hides
  // behind the previous step's FMAs), and ONE warp barrier sits between steps 2 and 3: by then block tb + 1 is complete
  // (step 3 prefetches its first step) and every lane already holds step 3 of block tb in registers, so nobody reads
  // buffer tb & 1 again before block tb + 2 is written into it.  The buffer of a block is also the substream (half)
  // it comes from, so the loop takes an even and an odd block per iteration: every index is a compile-time constant.
  const uint32_t nfull = T >> 2, tail = T & 3u;
  if (SRC_RNG) seed_all(0u);
#pragma unroll
  for (int a = 0; a < R; ++a) publish_row(std::integral_constant<int, 0>{}, 0u, a);
  __syncwarp();
  double z0[LR], z1[LR];
  read_step(0, 0, z0);
  auto block = [&](auto Hc, uint32_t tb) {
    constexpr int buf = decltype(Hc)::value;
    using NB = std::integral_constant<int, 1 - buf>;
    if (SRC_RNG && buf == 1 && ((tb + 1u) & (JNE_EPOCH_BLOCKS - 1u)) == 0u) seed_all((tb + 1u) / JNE_EPOCH_BLOCKS);
#pragma unroll
    for (int a = 0; a < R; ++a) if (a * 3 / R == 0) publish_row(NB{}, tb + 1u, a);
    read_step(buf, 1, z1);
    jne_group_step<D, L, DET, SRC_RNG>(S, z0);
#pragma unroll
    for (int a = 0; a < R; ++a) if (a * 3 / R == 1) publish_row(NB{}, tb + 1u, a);
    read_step(buf, 2, z0);
    jne_group_step<D, L, DET, SRC_RNG>(S, z1);
#pragma unroll
    for (int a = 0; a < R; ++a) if (a * 3 / R == 2) publish_row(NB{}, tb + 1u, a);
    read_step(buf, 3, z1);
    jne_group_step<D, L, DET, SRC_RNG>(S, z0);
    __syncwarp();
    read_step(buf ^ 1, 0, z0);
    jne_group_step<D, L, DET, SRC_RNG>(S, z1);
  };
  {
    uint32_t tb = 0;
    for (; tb + 1u < nfull; tb += 2u) {
      block(std::integral_constant<int, 0>{}, tb);
      block(std::integral_constant<int, 1>{}, tb + 1u);
    }
    if (tb < nfull) block(std::integral_constant<int, 0>{}, tb);
  }
  {
    const int buf = (int)(nfull & 1u);
#pragma unroll 1
    for (uint32_t s = 0; s < tail; ++s) {
      read_step(buf, (int)s, z0);
      jne_group_step<D, L, DET, SRC_RNG>(S, z0);
    }
  }
  if (!live) return;

  // ---- moments to global memory (JneLaneMom<D>); rotated column k is row (l R + k) mod LR ----
  double* m = mom + run * (uint64_t)M::SZ;
  auto tri = [](int i, int j) { const int lo = i < j ? i : j, hi = i < j ? j : i; return lo * D - ((lo * (lo - 1)) >> 1) + (hi - lo); };
  const double w1_last = w1_first + 2.0 * (Td - 1.0);
  const double w2_last = fma(3.0 * w1_last, w1_last, w2c);
#pragma unroll
  for (int a = 0; a < R; ++a) {
    const int i = l * R + a;
    if (i >= D) continue;
#pragma unroll
    for (int k = 0; k < LR; ++k) {
      const int j = (l * R + k) % LR;
      if (j < D) m[M::OFF_BZ + i * D + j] = S.bz[a][k];
    }
#pragma unroll
    for (int p = 0; p < Q::NFULL; ++p)
#pragma unroll
      for (int b = 0; b < R; ++b) {
        const int j = (l * R + (p + 1) * R + b) % LR;
        if (j < D) m[tri(i, j)] = S.bfull[p][a][b];
      }
    double* t = m + M::OFF_TOT + i;
    const double cT = S.c[a];                          // own rows are the first R of the rotated frame
    t[0 * D] = S.s0[a];
    t[1 * D] = S.s1[a];
    t[2 * D] = S.s2[a];
    t[3 * D] = cT;
    t[4 * D] = fma(w1_last, cT, -2.0 * S.s0[a]);       // summation by parts, see jne_lane_moments_kernel
    t[5 * D] = fma(w2_last, cT, 12.0 * (S.s0[a] - S.s1[a]));
  }
  {
    int idx = 0;
#pragma unroll
    for (int a = 0; a < R; ++a)
#pragma unroll
      for (int b = a; b < R; ++b) {
        const int i = l * R + a, j = l * R + b, jh = (l * R + (L / 2) * R + b) % LR;
        if (i < D && j < D) m[tri(i, j)] = S.bown[idx];
        if (Q::HALF && i < D && jh < D) m[tri(i, jh)] = S.bhalf[idx];
        ++idx;
      }
  }
}
