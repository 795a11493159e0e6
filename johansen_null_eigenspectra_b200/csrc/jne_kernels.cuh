// Device side of the Johansen null-eigenspectra hot path (sm_100a).
//
//   K1  jne_sub_normals4        (jne_rng.cuh)  replaces gen_normal_matrix        src/rng_matrix.rs:11-37
//   K2  jne_run_kernel main loop               replaces brownian_motion_matrix   src/rng_matrix.rs:57-141,
//                                              dmatrix_cumsum RowWise            src/matrix_utils.rs:51-63,
//                                              construct_f_matrix                src/johansen_statistics.rs:102-197,
//                                              2 x sum_of_outer_products         src/matrix_utils.rs:67-85
//   K3  jne_warp_gram / _jacobi / _emit        replaces GeneralizedEigen::new + |alpha|/beta + sort
//                                                                                src/johansen_statistics.rs:35-46
//
// One warp owns one run.  Lane (g = lane>>2, k = lane&3) owns Brownian rows g, g+8 of time
// SEGMENT k (the T steps are cut into 4 contiguous segments, one per MMA k-slot), carries the
// segment-local cumulative sum c in registers and feeds V = [F ; dB] straight into
// mma.sync.m8n8k4.f64 fragments: for that shape the A fragment (row g, k) and the B fragment
// (k, col g) are both "the lane's own value", so sum_t V V' needs no operand staging at all.
// The deterministic regressors (1, tau, tau^2) never enter the MMA: their cross moments are
// five FP64 FMAs per row and step, and demeaning / detrending is applied once per run as a
// Schur complement in the epilogue (SURVEY.md section 7 hard part 4).  Segment-local sums are stitched
// with the rank-one corrections of SURVEY.md section 5 ("long-context") in the epilogue.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "jne_rng.cuh"

#define JNE_MAX_DIM 15           // p = dim+1 <= 16: one half-warp covers a row / column of the work matrices
#ifndef JNE_WARPS_PER_CTA
#define JNE_WARPS_PER_CTA 4
#endif

struct JneRunParams {
  uint32_t dim;        // d
  uint32_t steps;      // T
  uint32_t seg_len;    // steps per segment, multiple of 8
  uint32_t model;      // 0..4
  uint32_t p;          // eigenvalues per run
  uint32_t full_blocks;  // leading 8-step blocks that are inside the segment for every lane
  uint32_t model_mask;   // bit m set: solve model m (multi-model launches; single-model launches set one bit)
  uint32_t out_stride;   // doubles per run in `out` (sum of p over the selected models)
  const uint32_t* jtab;  // device pointer: Jacobi step table for ne = even(dim) (jne_api.cu, make_jacobi_tables)
  const double* aux_tab;   // trend weights for the MMA (jne_step, AUX): [seg_len][4 weights][4 segments], see make_aux_table
  double T;            // (double)steps
  double factor;       // s^2 * T: 1 for the RNG path (s^2 = dt), T for caller-supplied increments
  double seg_n[4];     // steps in segment k
  double seg_w1[4];    // sum over segment k of w1_i = 2i + 1 - T
  double seg_w2[4];    // sum over segment k of w2_i = 3 w1_i^2 - (T^2 - 1)
};

template <int DP> struct JneGeo {
  static constexpr int A = DP / 8, B = DP % 8;
  static constexpr int NRT = (DP + 7) / 8;       // row tiles == Brownian rows per lane
  static constexpr int NCT = (2 * DP + 7) / 8;   // column tiles of V = [F ; dB]
  static constexpr int NT = NRT * NCT - NRT * (NRT - 1) / 2;  // tiles (a, b >= a)
  static constexpr int VV_LD = 8 * NCT + 1;
  // per-warp shared memory, in doubles
  static constexpr int VV_SZ = 8 * NRT * VV_LD;
  static constexpr int VEC_SZ = 6 * 4 * 16;
  static constexpr int TOT_SZ = 7 * 16;
  static constexpr int RAW_SZ = TOT_SZ + VV_SZ + VEC_SZ;
  static constexpr int STITCH_HALF = DP * 16;                    // M_BB (and M_Bz): DP rows x 16 columns
  static constexpr int STITCH_SZ = 2 * STITCH_HALF;
  // pencil matrices S2 (p x p) and R = S1' (p x d), p <= min(DP + 1, 16): odd leading dimension (bank spread)
  static constexpr int LDW = DP + 1;
  static constexpr int WROWS = DP < 16 ? DP + 1 : 16;
  static constexpr int MAT_SZ = WROWS * LDW + (WROWS * LDW & 1);
  // Jacobi: ne = even(d) <= DP players, NPAIR pair slots, NBLK 2x2 blocks (P1 <= P2), packed upper triangle
  static constexpr int NPAIR = DP / 2;
  static constexpr int NBLK = NPAIR * (NPAIR + 1) / 2;
  static constexpr int G_SZ = DP * (DP + 1) / 2;
};

// Per-warp shared-memory layout of the epilogue for up to NM models per run (1 or 5), in doubles:
//   [0, TOT)  totals | stitched M_BB, M_Bz | S2, R | G[NM] packed | cs[NM][NPAIR] (c, s) | ev[16] | fac[8]
// The raw moment dump of the time loop (VV, vec) aliases everything behind the totals.  NM == 1: S2 / R alias the
// stitched moments (jne_warp_assemble passes the entries through registers).
template <int DP, int NM> struct JneEpi {
  using G = JneGeo<DP>;
  static constexpr int OFF_S2 = NM > 1 ? G::TOT_SZ + G::STITCH_SZ : G::TOT_SZ;
  static constexpr int OFF_R = OFF_S2 + G::MAT_SZ;
  static constexpr int OFF_G = NM > 1 ? OFF_R + G::MAT_SZ
                                      : G::TOT_SZ + (G::STITCH_SZ > 2 * G::MAT_SZ ? G::STITCH_SZ : 2 * G::MAT_SZ);
  static constexpr int OFF_CS = OFF_G + NM * G::G_SZ;
  static constexpr int OFF_EV = OFF_CS + 2 * NM * G::NPAIR;
  static constexpr int OFF_FAC = OFF_EV + 16;
  static constexpr int END = OFF_FAC + 8;
  static constexpr int WARP_SMEM = G::RAW_SZ > END ? G::RAW_SZ : END;
  static constexpr int NPASS_B = (NM * G::NBLK + 31) / 32;    // block passes per Jacobi step
  static constexpr int NPASS_R = (NM * G::NPAIR + 31) / 32;   // rotation passes per Jacobi step
  static_assert(OFF_CS % 2 == 0 && WARP_SMEM % 2 == 0 && END % 2 == 0, "double2 alignment of cs");
};

__device__ __forceinline__ void jne_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------------------------
// K3: eigenvalues of the pencil (S1'S1, S2), one warp per run, up to five models of the run at once.
//   jne_warp_gram    per model: S2 = L L' (Cholesky), W = L^-1 R, G = W'W in ONE right-looking elimination
//                    (the d x d Gram matrix has the pencil's non-zero spectrum), unit trace, packed triangle
//   jne_warp_jacobi  cyclic two-sided Jacobi on all the run's G matrices together: the ne/2 disjoint pairs of
//                    a round-robin step rotate concurrently and G <- J'GJ is applied per 2x2 BLOCK (pair slot
//                    P1 x pair slot P2), one lane per block; rotation and block roles of every model share the
//                    warp's 32 lanes, element addresses come from a host-built step table
//   jne_warp_emit    lambda_i = factor * |g_ii| sorted descending (src/johansen_statistics.rs:40-45)
// Replaces GeneralizedEigen::new (LAPACK dggev) + |alpha|/beta + sort, src/johansen_statistics.rs:35-46.
// ---------------------------------------------------------------------------------------------
// 1/x and 1/sqrt(x) to ~1 ulp from an FP32 seed + two Newton steps; x must be inside the float range
// (true for every call site below).  Avoids the ~35-instruction IEEE division / sqrt sequences.
__device__ __forceinline__ double jne_rcp(double x) {
  float r0;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"((float)x));
  double r = (double)r0;
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
}
__device__ __forceinline__ double jne_rsqrt(double x) {
  float y0;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"((float)x));
  double y = (double)y0;
  const double hx = 0.5 * x;
  y = fma(y, fma(-hx * y, y, 0.5), y);
  y = fma(y, fma(-hx * y, y, 0.5), y);
  return y;
}

// 1/sqrt(x) for any positive finite double (the Cholesky pivots scale with the caller's increments squared,
// so they may lie outside the float range): x = m 4^k, m in [1, 4).  Non-positive or NaN -> NaN.
__device__ __forceinline__ double jne_rsqrt_wide(double x) {
  if (!(x > 0.0) || !(x < 1.7e308)) return __longlong_as_double(0x7ff8000000000000ll);
  double post = 1.0;
  if (x < 0x1p-900) { x *= 0x1p200; post = 0x1p100; }         // denormals / tiny values: rsqrt(x 2^200) 2^100
  const int hi = __double2hiint(x);
  const int k2 = (((hi >> 20) & 0x7ff) - 1023) & ~1;          // even part of the exponent
  const double m = __hiloint2double(hi - (k2 << 20), __double2loint(x));
  const double y = jne_rsqrt(m);
  return __hiloint2double(__double2hiint(y) - ((k2 >> 1) << 20), __double2loint(y)) * post;
}

// offset of element (i <= j) in the packed upper triangle of an ne x ne symmetric matrix (row i holds ne - i entries)
__device__ __forceinline__ int jne_tri(int i, int j, int ne) { return i * ne - ((i * (i - 1)) >> 1) + (j - i); }

// G = R' S2^-1 R for one model.
//   S2 : p x p symmetric positive definite, lower triangle used, ld LDW   (destroyed)
//   R  : p x d = S1' (row j = sum_t F_j dB_t'), ld LDW                     (destroyed)
// Right-looking elimination, one pivot j per step: lanes 0..15 scale column j of L (row x = lane), lanes
// 16..31 scale row j of W = L^-1 R (column x = lane - 16); both travel by shuffle, are never stored, and
// feed (a) the trailing update of S2, (b) the forward substitution of the remaining rows of R and (c) the
// Gram accumulation G += w_j w_j' held in registers (lane (h, x) owns G[2q + h][x], x <= 2q + h).
// Models 1 and 3 (p = d + 1) have rank d: their extra eigenvalue is exactly 0 here (dggev returns O(1e-14)
// noise for it).  G is zero-padded to ne x ne (ne even), scaled to unit trace (makes the solve invariant to
// the scale of the caller's increments and keeps the FP32-seeded reciprocals of the Jacobi in range) and
// written as a packed upper triangle; the trace is returned on every lane.
// unit-trace packed store of the lane-distributed accumulators; returns the trace on every lane
template <int NQ>
__device__ __forceinline__ double jne_gram_store(const double (&acc)[NQ], int ne, double* __restrict__ Gm) {
  const int lane = threadIdx.x & 31, x = lane & 15, h = lane >> 4;
  double tr = 0.0;
#pragma unroll
  for (int q = 0; q < NQ; ++q)
    if (2 * q + h == x) tr += acc[q];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tr += __shfl_xor_sync(0xffffffffu, tr, o);
  const double inv_tr = 1.0 / tr;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int i = 2 * q + h;
    if (i < ne && x <= i) Gm[jne_tri(x, i, ne)] = acc[q] * inv_tr;
  }
  return tr;
}

// snap_at > 0: after snap_at pivots the Gram matrix of the leading snap_at rows of F is also stored (Gsnap, its trace
// to *tr_snap): the elimination of a model whose F extends another model's F by trailing rows serves both.
template <int DP, bool SNAP = false>
__device__ __forceinline__ double jne_warp_gram(double* __restrict__ S2, double* __restrict__ R, int p, int d, int ne,
                                                double* __restrict__ Gm, int snap_at = -1,
                                                double* __restrict__ Gsnap = nullptr, double* tr_snap = nullptr) {
  constexpr int LDW = JneGeo<DP>::LDW, NQ = (DP + 1) / 2;
  const int lane = threadIdx.x & 31, x = lane & 15, h = lane >> 4;
  const bool row_s = x < p, col_r = x < d;
  double acc[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) acc[q] = 0.0;
  for (int j = 0; j < p; ++j) {
    const double inv = jne_rsqrt_wide(S2[j * LDW + j]);       // NaN for a non-positive pivot -> flagged at emit
    double v = 0.0;
    if (h == 0) { if (x > j && row_s) v = S2[x * LDW + j] * inv; }     // L[x][j]
    else if (col_r) v = R[j * LDW + x] * inv;                            // W[j][x]
    const double lxj = __shfl_sync(0xffffffffu, v, x);
    const double wjx = __shfl_sync(0xffffffffu, v, 16 + x);
    for (int k0 = j + 1; k0 < p; k0 += 2) {                   // two trailing columns / rows per pass
      const int kk = k0 + h;
      const double lkj = __shfl_sync(0xffffffffu, v, kk & 15);
      if (kk < p) {
        if (x >= kk && row_s) S2[x * LDW + kk] = fma(-lxj, lkj, S2[x * LDW + kk]);
        if (col_r) R[kk * LDW + x] = fma(-lkj, wjx, R[kk * LDW + x]);
      }
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q)
      if (2 * q < d) acc[q] = fma(__shfl_sync(0xffffffffu, v, 16 + ((2 * q + h) & 15)), wjx, acc[q]);
    if (SNAP && j + 1 == snap_at) {
      const double trs = jne_gram_store<NQ>(acc, ne, Gsnap);
      if (lane == 0) *tr_snap = trs;
    }
    __syncwarp();
  }
  return jne_gram_store<NQ>(acc, ne, Gm);
}

// Cyclic Jacobi on nm packed matrices Gs[m * gsz ..] at once.  Step table (per step: npairs rotation words, then
// nblk block words; bytes are packed-triangle offsets):
//   rotation word of pair slot l = (p, q):      o(p,p) | o(q,q) << 8 | o(p,q) << 16
//   block word of (P1 <= P2) = (p1,q1),(p2,q2): o(p1,p2) | o(p1,q2) << 8 | o(q1,p2) << 16 | o(q1,q2) << 24
// tan(theta) comes from FP32 arithmetic (it only steers convergence); (c, s) is normalised in FP64 so that every
// J is orthogonal to rounding.  Leaving a_pq in place moves the two eigenvalues by about a_pq^2 / |a_qq - a_pp|
// (second order) and never by more than |a_pq|: a rotation is skipped when that is below 1e-14 of the smaller
// one, or when |a_pq| <= 2^-50 outright (trace = 1).  A model whose sweep rotates nothing stays untouched while
// the others finish, so every matrix sees exactly the rotations it would see alone.
template <int NPB, int NPR>
__device__ __forceinline__ void jne_warp_jacobi(double* __restrict__ Gs, double2* __restrict__ cs,
                                                const uint32_t* __restrict__ tab, int nm, int ne, int gsz) {
  const int lane = threadIdx.x & 31;
  const int npairs = ne >> 1, nblk = npairs * (npairs + 1) / 2, stride = npairs + nblk;
  constexpr uint32_t NONE = 0xffffffffu;
  // roles: rotation role r = m * npairs + slot; block role b = m * nblk + blk, blk -> (P1 <= P2)
  uint32_t rrole[NPR], brole[NPB];
#pragma unroll
  for (int q = 0; q < NPR; ++q) {
    const int r = lane + 32 * q;
    rrole[q] = NONE;
    if (r < nm * npairs) {
      const int m = r / npairs;
      rrole[q] = (uint32_t)(m * gsz) | ((uint32_t)r << 10) | ((uint32_t)(r - m * npairs) << 16);
    }
  }
#pragma unroll
  for (int q = 0; q < NPB; ++q) {
    const int b = lane + 32 * q;
    brole[q] = NONE;
    if (b < nm * nblk) {
      const int m = b / nblk, blk = b - m * nblk;
      int r = 0, k = blk;
      while (k >= npairs - r) { k -= npairs - r; ++r; }     // block (P1 = r, P2 = r + k)
      brole[q] = (uint32_t)(m * gsz) | ((uint32_t)(m * npairs + r) << 10) | ((uint32_t)(m * npairs + r + k) << 16) |
                 ((uint32_t)blk << 22) | (k == 0 ? (1u << 28) : 0u);
    }
  }
  const double tol = 8.8817841970012523e-16;
  for (int sweep = 0; sweep < 30; ++sweep) {
    int rotated = 0;
    const uint32_t* ts = tab;
    for (int step = 0; step < ne - 1; ++step, ts += stride) {
#pragma unroll
      for (int q = 0; q < NPR; ++q) {
        if (rrole[q] != NONE) {
          const uint32_t w = __ldg(ts + ((rrole[q] >> 16) & 63u));
          const double* gm = Gs + (rrole[q] & 1023u);
          const double app = gm[w & 255u], aqq = gm[(w >> 8) & 255u], apq = gm[(w >> 16) & 255u];
          double c = 1.0, s = 0.0;
          const double diff = aqq - app;
          if (fabs(apq) > tol && apq * apq > 1e-14 * fabs(diff) * fmin(app, aqq)) {
            float th, hy, tf;
            asm("div.approx.ftz.f32 %0, %1, %2;" : "=f"(th) : "f"((float)diff), "f"(2.0f * (float)apq));
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(hy) : "f"(fmaf(th, th, 1.0f)));
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(tf) : "f"(fabsf(th) + hy));
            const double t = (double)copysignf(tf, th);
            c = jne_rsqrt(fma(t, t, 1.0));
            s = t * c;
            rotated = 1;
          }
          cs[(rrole[q] >> 10) & 63u] = make_double2(c, s);
        }
      }
      __syncwarp();
      // block pass: B <- J1' B J2 for the 2x2 block (rows of slot P1) x (cols of slot P2)
#pragma unroll
      for (int q = 0; q < NPB; ++q) {
        if (brole[q] != NONE) {
          const double2 r1 = cs[(brole[q] >> 10) & 63u], r2 = cs[(brole[q] >> 16) & 63u];
          const double c1 = r1.x, s1 = r1.y, c2 = r2.x, s2 = r2.y;
          if (s1 != 0.0 || s2 != 0.0) {
            const uint32_t w = __ldg(ts + npairs + ((brole[q] >> 22) & 63u));
            double* gm = Gs + (brole[q] & 1023u);
            double* e00 = gm + (w & 255u);
            double* e01 = gm + ((w >> 8) & 255u);
            double* e10 = gm + ((w >> 16) & 255u);
            double* e11 = gm + (w >> 24);
            const double x00 = *e00, x01 = *e01, x10 = *e10, x11 = *e11;
            const double r00 = fma(c1, x00, -s1 * x10), r01 = fma(c1, x01, -s1 * x11);   // J1' B
            const double r10 = fma(s1, x00, c1 * x10), r11 = fma(s1, x01, c1 * x11);
            const double y00 = fma(c2, r00, -s2 * r01), y11 = fma(s2, r10, c2 * r11);    // (.) J2
            double y01 = fma(s2, r00, c2 * r01), y10 = fma(c2, r10, -s2 * r11);
            if (brole[q] & (1u << 28)) { y01 = 0.5 * (y01 + y10); y10 = y01; }   // diagonal block: the (nearly) annihilated element (e01 == e10)
            *e00 = y00; *e01 = y01; *e10 = y10; *e11 = y11;
          }
        }
      }
      __syncwarp();
    }
    if (!__any_sync(0xffffffffu, rotated)) break;
  }
}

// Eigenvalues of one model = factor * |diag| (d of them, the other p - d are 0), sorted descending by rank
// counting.  Returns false when a value is not finite (the reference panics at src/johansen_statistics.rs:45).
__device__ __forceinline__ bool jne_warp_emit(const double* __restrict__ Gm, double factor, int p, int d, int ne,
                                              double* __restrict__ ev, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  double v = 0.0;
  if (lane < p) {
    v = (lane < d) ? fabs(Gm[jne_tri(lane, lane, ne)]) * factor : 0.0 * factor;   // 0 * NaN keeps a failure visible
    ev[lane] = v;
  }
  __syncwarp();
  bool finite = true;
  if (lane < p) {
    int rank = 0;
    for (int j = 0; j < p; ++j) {
      const double u = ev[j];
      rank += (u > v) || (u == v && j < lane);
    }
    finite = isfinite(v);
    if (!finite) rank = lane;   // NaN compares false everywhere: keep slots distinct
    out[rank] = v;
  }
  __syncwarp();
  return __all_sync(0xffffffffu, finite);
}

// ---------------------------------------------------------------------------------------------
// Epilogue part 1, once per run: stitch the four segments (model-independent).
//   VV  [8*NRT][VV_LD] : sum over segments and steps of V V', V = [c (DP rows) ; z (DP rows)]
//   vec [6][4][16]     : per segment k and row r:  0 e = c_end, 1 s0 = sum c, 2 s1 = sum w1 c,
//                        3 s2 = sum w2 c, 4 u1 = sum w1 z, 5 u2 = sum w2 z
// out:
//   tot [6][16]        : whole-run totals  0 S_B = sum B, 1 S_1B = sum w1 B, 2 S_2B = sum w2 B,
//                        3 S_z = sum dB, 4 S_1z = sum w1 dB, 5 S_2z = sum w2 dB
//   MBB, MBZ [16][16]  : sum B B' and sum B dB' with B = b0 + c:  sum (b0+c)(b0+c)' =
//                        n b0 b0' + b0 (sum c)' + (sum c) b0' + sum c c'   
// MBB / MBZ ALIAS the raw area: every entry is computed into registers, then, after a warp barrier, stored.
// ---------------------------------------------------------------------------------------------
template <int DP, bool AUX = false>
__device__ __forceinline__ void jne_warp_stitch(const double* VV, const double* vec, double* tot,
                                                double* MBB, double* MBZ, const JneRunParams& prm) {
  using G = JneGeo<DP>;
  const int lane = threadIdx.x & 31;
  const int d = prm.dim;
  if (lane < 16) {
    const int r = lane;
    double b0 = 0.0, sB = 0.0, s1B = 0.0, s2B = 0.0, sz = 0.0, s1z = 0.0, s2z = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double e = vec[(0 * 4 + k) * 16 + r];
      sB += vec[(1 * 4 + k) * 16 + r] + prm.seg_n[k] * b0;
      if (AUX) {
        s1B += prm.seg_w1[k] * b0;
        s2B += prm.seg_w2[k] * b0;
      } else {
        s1B += vec[(2 * 4 + k) * 16 + r] + prm.seg_w1[k] * b0;
        s2B += vec[(3 * 4 + k) * 16 + r] + prm.seg_w2[k] * b0;
        s1z += vec[(4 * 4 + k) * 16 + r];
        s2z += vec[(5 * 4 + k) * 16 + r];
      }
      sz += e;
      b0 += e;
    }
    if (AUX && r < DP) {
      if (G::B == 4) {   // rows 8A+4+m of VV: sum over steps and segments of weight_m(step) * dB_r (jne_step)
        const double* wr = VV + (8 * G::A + 4) * G::VV_LD + DP + r;
        s1B += wr[0 * G::VV_LD];   // weight 0: sum_{i > s, same segment} w1_i  ->  sum w1_i c_i
        s2B += wr[1 * G::VV_LD];   // weight 1: the same for w2
        s2z = wr[2 * G::VV_LD];    // weight 2: w2_s
        s1z = wr[3 * G::VV_LD];    // weight 3: w1_s
      } else {           // DP = 8, dim <= 6: rows 6, 7 of V carry w1, w2 themselves: products with c and with dB
        s1B += VV[6 * G::VV_LD + r];
        s2B += VV[7 * G::VV_LD + r];
        s1z = VV[6 * G::VV_LD + DP + r];
        s2z = VV[7 * G::VV_LD + DP + r];
      }
    }
    tot[0 * 16 + r] = sB;  tot[1 * 16 + r] = s1B; tot[2 * 16 + r] = s2B;
    tot[3 * 16 + r] = sz;  tot[4 * 16 + r] = s1z; tot[5 * 16 + r] = s2z;
  }
  constexpr int NQ = (DP + 1) / 2;
  double r_bb[NQ], r_bz[NQ];
  const int j = lane & 15;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int i = 2 * q + (lane >> 4);
    double mbb = 0.0, mbz = 0.0;
    if (i < d && j < d) {
      double bi = 0.0, bj = 0.0;
      mbb = (i <= j) ? VV[i * G::VV_LD + j] : VV[j * G::VV_LD + i];
      mbz = VV[i * G::VV_LD + DP + j];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const double ei = vec[(0 * 4 + k) * 16 + i], ej = vec[(0 * 4 + k) * 16 + j];
        const double s0i = vec[(1 * 4 + k) * 16 + i], s0j = vec[(1 * 4 + k) * 16 + j];
        mbb += prm.seg_n[k] * bi * bj + bi * s0j + s0i * bj;
        mbz += bi * ej;
        bi += ei; bj += ej;
      }
    }
    r_bb[q] = mbb;
    r_bz[q] = mbz;
  }
  __syncwarp();   // all reads of the raw area are done: the stitched moments may now overwrite it
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int i = 2 * q + (lane >> 4);
    if (i < DP) { MBB[i * 16 + j] = r_bb[q]; MBZ[i * 16 + j] = r_bz[q]; }
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// Epilogue part 2, per model: S2 = sum F F' (p x p) and R = S1' = sum F dB' (p x d) from the stitched
// moments, with demeaning / detrending as Schur complements.  F per model follows
// src/johansen_statistics.rs:102-197.  Row scalings of F leave the pencil's
// eigenvalues unchanged, so the trend row is carried as (w1+1)/T (= 2 (tau - 1/2)) and the detrended
// tau^2 row as w2/T^2 (= 12 x its residual on [1, tau]).  Row ORDER does not change them either (G = R'S2^-1 R is
// invariant under any invertible map of F's rows): model 3 is laid out as [B_0 .. B_{d-2}, trend, B_{d-1}] so that
// its first d rows are exactly model 2's F and one elimination serves both (jne_warp_models).  ALIASED: S2 / R may
// overlap MBB / MBZ (single-model launches); entries pass through registers and a warp barrier.
// ---------------------------------------------------------------------------------------------
template <int DP>
__device__ __forceinline__ void jne_warp_assemble(const double* MBB, const double* MBZ, const double* tot,
                                                  double* S2, double* R, const JneRunParams& prm, int model, int p) {
  constexpr int LDW = JneGeo<DP>::LDW;
  const int lane = threadIdx.x & 31;
  const int d = prm.dim;
  const double T = prm.T;
  const int nb = (model == 2 || model == 4) ? d - 1 : d;   // Brownian rows kept in F
  const int rdet = (model == 3) ? d - 1 : nb;              // position of the deterministic row
  const double invT = 1.0 / T;
  const double nu = T * (T * T - 1.0) / 3.0;               // sum w1^2
  const double inv_nu = 1.0 / nu;                          // inf at T = 1 (model 4 needs T >= 3)
  constexpr int NQ = (DP + 1) / 2;
  double r_bb[NQ], r_bz[NQ];
  const int j = lane & 15;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int i = 2 * q + (lane >> 4);
    double mbb = 0.0, mbz = 0.0;
    if (i < nb) {
      mbb = MBB[i * 16 + j];
      mbz = MBZ[i * 16 + j];
      const double SBi = tot[0 * 16 + i], S1Bi = tot[1 * 16 + i];
      if (model >= 2) {   // demean (models 2,3,4)  src/johansen_statistics.rs:120-125,145-150,186-194
        mbb -= SBi * tot[0 * 16 + j] * invT;
        mbz -= SBi * tot[3 * 16 + j] * invT;
      }
      if (model == 4) {   // detrend: residual on [1, tau]   :186-194
        mbb -= S1Bi * tot[1 * 16 + j] * inv_nu;
        mbz -= S1Bi * tot[4 * 16 + j] * inv_nu;
      }
    }
    r_bb[q] = mbb;
    r_bz[q] = mbz;
  }
  // --- deterministic row (position rdet), lanes 0..15 ---
  double s2v = 0.0, rv = 0.0, dg = 0.0;
  if (p > nb && lane < 16) {
    if (model == 1) {                 // constant row  :108-113
      s2v = tot[0 * 16 + j];          // sum B_j
      rv = tot[3 * 16 + j];           // sum dB_j
      dg = T;
    } else if (model == 2 || model == 3) {   // trend row (i+1)/T - 1/2 = (w1+1)/(2T)  :127-135,152-160
      s2v = tot[1 * 16 + j] * invT;                   // sum (B_j - mean) (w1+1)/T = S_1B/T
      rv = (tot[4 * 16 + j] + tot[3 * 16 + j]) * invT;
      dg = (nu + T) * invT * invT;                    // sum (w1+1)^2 / T^2
    } else if (model == 4) {          // tau^2 residual on [1, tau] = w2 / (12 T^2)  :170-194
      s2v = tot[2 * 16 + j] * invT * invT;
      rv = tot[5 * 16 + j] * invT * invT;
      dg = 0.8 * T * (T * T - 1.0) * (T * T - 4.0) * invT * invT * invT * invT;   // sum w2^2 / T^4
    }
  }
  __syncwarp();   // all reads of the stitched moments are done: S2 / R may overwrite them (single-model)
  const int jp = (j >= rdet && j < nb) ? j + 1 : j;        // position of Brownian row j (model 3: B_{d-1} sits behind the trend)
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int i = 2 * q + (lane >> 4);
    if (i < nb) {
      const int ip = (i >= rdet) ? i + 1 : i;
      if (j < nb) S2[ip * LDW + jp] = r_bb[q];
      if (j < d) R[ip * LDW + j] = r_bz[q];
    }
  }
  if (p > nb && lane < 16) {
    if (j < nb) { S2[rdet * LDW + jp] = s2v; S2[jp * LDW + rdet] = s2v; }
    if (j == nb) S2[rdet * LDW + rdet] = dg;
    if (j < d) R[rdet * LDW + j] = rv;
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// Epilogue part 3: every selected model of the run.  wsm holds the totals and the stitched moments (layout JneEpi).
// dbg (optional): the last selected model's S2 (16 x 16) then R (16 x 16).
// ---------------------------------------------------------------------------------------------
template <int DP, int NM>
__device__ __forceinline__ bool jne_warp_models(double* __restrict__ wsm, const JneRunParams& prm, double* __restrict__ out,
                                                double* __restrict__ dbg) {
  using G = JneGeo<DP>;
  using E = JneEpi<DP, NM>;
  const int lane = threadIdx.x & 31;
  const double* tot = wsm;
  const double* MBB = wsm + G::TOT_SZ;
  const double* MBZ = MBB + G::STITCH_HALF;
  double* S2 = wsm + E::OFF_S2;
  double* R = wsm + E::OFF_R;
  double* Gs = wsm + E::OFF_G;
  double2* cs = reinterpret_cast<double2*>(wsm + E::OFF_CS);
  double* ev = wsm + E::OFF_EV;
  double* fac = wsm + E::OFF_FAC;    // per slot: trace of G (x prm.factor at emit)
  const int d = prm.dim, ne = (d + 1) & ~1, gsz = ne * (ne + 1) / 2;   // odd d: index d is an all-zero row / column
  const uint32_t mask = prm.model_mask;
  // Three eliminations serve the five models: F of model 1 is F of model 0 plus the constant row, F of model 3
  // (in the row order of jne_warp_assemble) is F of model 2 plus B_{d-1}; the shorter model's Gram matrix is the
  // snapshot after its own d pivots -- bit for bit what its own elimination produces.  Slots are in model order.
#pragma unroll 1
  for (int c = 0; c < 3; ++c) {
    const uint32_t bits = (mask >> (2 * c)) & (c < 2 ? 3u : 1u);
    if (bits == 0u) continue;
    const int model = (bits & 2u) ? 2 * c + 1 : 2 * c;
    const int p = (model == 1 || model == 3) ? d + 1 : d;
    const int slot_lo = __popc(mask & ((1u << (2 * c)) - 1u));
    jne_warp_assemble<DP>(MBB, MBZ, tot, S2, R, prm, model, p);
    if (NM == 1 && dbg != nullptr) {   // reference row order: model 3 carries its trend row last (:152-160)
      const int rdet = (model == 3) ? d - 1 : p;
      for (int e = lane; e < 256; e += 32) {
        const int i = e >> 4, j = e & 15;
        const int ip = (model != 3 || i < rdet) ? i : (i == d ? rdet : i + 1);
        const int jp = (model != 3 || j < rdet) ? j : (j == d ? rdet : j + 1);
        dbg[e] = (i < p && j < p) ? S2[ip * G::LDW + jp] : 0.0;
        dbg[256 + e] = (i < p && j < d) ? R[ip * G::LDW + j] : 0.0;
      }
      __syncwarp();
    }
    const bool both = bits == 3u;
    const double tr = jne_warp_gram<DP, (NM > 1)>(S2, R, p, d, ne, Gs + (slot_lo + (both ? 1 : 0)) * gsz, both ? d : -1,
                                        Gs + slot_lo * gsz, fac + slot_lo);
    if (lane == 0) fac[slot_lo + (both ? 1 : 0)] = tr;
    if (NM == 1) break;
    __syncwarp();
  }
  __syncwarp();
  const int nm = __popc(mask);
  jne_warp_jacobi<E::NPASS_B, E::NPASS_R>(Gs, cs, prm.jtab, nm, ne, gsz);
  bool ok = true;
  int k = 0;
#pragma unroll 1
  for (int model = 0; model < 5; ++model) {
    if (!((mask >> model) & 1u)) continue;
    const int p = (model == 1 || model == 3) ? d + 1 : d;
    ok &= jne_warp_emit(Gs + k * gsz, prm.factor * fac[k], p, d, ne, ev, out);
    out += p;
    ++k;
    if (NM == 1) break;
  }
  return ok;
}

// ---------------------------------------------------------------------------------------------
// The time loop works on blocks of 8 consecutive steps of the lane's segment.  jne_gen8 produces the
// block's increments (generator + Box-Muller, or global loads), jne_consume8 feeds them to the MMAs.  Both
// are branch-free.
//
// RNG balance: a generator call yields 4 steps of one row, and a lane owns DP/8 rows (0.5, 1, 1.5 or 2).  With
// 1.5 rows (DP = 12) lanes g >= 4 would idle through the second row's calls, so the calls are spread over
// all lanes instead: per 8 steps every lane generates its row g twice (steps 0-3, 4-7) and ONE block of a
// row 8 + (g & 3) -- lanes g < 4 for steps 0-3, lanes g >= 4 for steps 4-7 -- and the upper lanes hand
// theirs to lane g - 4 with four FP32 shuffles: 3 calls per lane instead of 4.  DP = 4 likewise: 1 instead of 2.
// This is what the two substreams per row and epoch (jne_rng.cuh: the halves, by parity of the four-step block) are
// for: the two lanes that share a row each advance their own generator.  The value of element (row, step) is
// unchanged: it depends on (seed, row, step) only.
// ---------------------------------------------------------------------------------------------
// Generator state of a lane: both halves of the rows it owns whole (state 2 j + h), one half of the shared row slot
// (DP = 4, 12; the last state).  Registers: keeping the states in the warp's loop-idle shared memory instead (one
// LDS.128 / STS.128 pair per call) measured 1 % slower (profiles/r2_variants_generators.txt).
#ifdef JNE_EXP_NORESEED   // experiment only: substreams keyed once per run and segment (NOT a valid stream)
#define JNE_EPOCH_START(t, t_begin, unaligned) ((t) == (t_begin))
#else   // aligned segments: every lane of the warp at once; unaligned ones: every lane in its own phase
#define JNE_EPOCH_START(t, t_begin, unaligned) ((((unaligned) ? (t) : (t) - (t_begin)) & (JNE_EPOCH_STEPS - 1u)) == 0u)
#endif
template <int DP> struct JneGen {
  using G = JneGeo<DP>;
  static constexpr int NOWN = (G::B == 0) ? G::NRT : G::NRT - 1;
  static constexpr int NST = 2 * NOWN + (G::B != 0 ? 1 : 0);
  jne_sub st[NST];
};
// Start of an epoch (every JNE_EPOCH_STEPS steps of the lane's segment, warp-uniform): new substream keys.
template <int DP>
__device__ __forceinline__ void jne_gen_seed(JneGen<DP>& gs, uint32_t seed, uint32_t epoch, int g) {
  using G = JneGeo<DP>;
#pragma unroll
  for (int j = 0; j < JneGen<DP>::NOWN; ++j) {
    jne_sub_seed(gs.st[2 * j], seed, 8 * j + g, epoch, 0u);
    jne_sub_seed(gs.st[2 * j + 1], seed, 8 * j + g, epoch, 1u);
  }
  if (G::B != 0) jne_sub_seed(gs.st[JneGen<DP>::NST - 1], seed, 8 * (G::NRT - 1) + (g & 3), epoch, (uint32_t)(g >> 2));
}
template <int DP, bool SRC_RNG> struct JneZ { using type = jne_zt; };
template <int DP> struct JneZ<DP, false> { using type = double; };

template <int DP, bool SRC_RNG>
__device__ __forceinline__ void jne_gen8(uint32_t t, uint32_t t_end, uint32_t d, int g, JneGen<DP>& gs,
                                         const float (&rowscale)[JneGeo<DP>::NRT], float xscale,
                                         const double* __restrict__ dBrun,
                                         typename JneZ<DP, SRC_RNG>::type (&z)[JneGeo<DP>::NRT][8]) {
  using G = JneGeo<DP>;
  if constexpr (!SRC_RNG) {
#pragma unroll
    for (int j = 0; j < G::NRT; ++j) {
      const uint32_t row = 8 * j + g;
#pragma unroll
      for (int s = 0; s < 8; ++s) z[j][s] = (row < d && t + s < t_end) ? dBrun[(uint64_t)(t + s) * d + row] : 0.0;
    }
  } else if constexpr (G::B == 0) {           // DP = 8, 16: every lane owns whole rows
#pragma unroll
    for (int j = 0; j < G::NRT; ++j) {     // t is a multiple of 8: steps 0-3 are an even block (half 0), 4-7 an odd one
      jne_sub_normals4(gs.st[2 * j], &z[j][0], rowscale[j]);
      jne_sub_normals4(gs.st[2 * j + 1], &z[j][4], rowscale[j]);
    }
  } else {                                    // DP = 4, 12: the last row slot is shared by lanes g and g ^ 4
    constexpr int L = G::NRT - 1;             // the shared slot
#pragma unroll
    for (int j = 0; j < L; ++j) {
      jne_sub_normals4(gs.st[2 * j], &z[j][0], rowscale[j]);
      jne_sub_normals4(gs.st[2 * j + 1], &z[j][4], rowscale[j]);
    }
    jne_zt x[4];
    jne_sub_normals4(gs.st[2 * L], x, xscale);     // row 8 L + (g & 3), block (t >> 2) + (g >> 2)
    // Lanes g >= 4 own no row in this slot: what they accumulate there (a path nobody reads: their operand in
    // the mixed group is the received increment or the trend weight, and rows >= DP of the dump are ignored) is
    // left unmasked -- zeroing it cost two FSEL per step after the widening.
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const jne_zt other = __shfl_xor_sync(0xffffffffu, x[s], 16);    // lane g ^ 4, same segment
      z[L][s] = x[s];
      z[L][4 + s] = other;
    }
  }
}

// State of a run's time loop carried from block to block.
template <int DP> struct JneLoopState {
  using G = JneGeo<DP>;
  double c[G::NRT], s0[G::NRT], s1[G::NRT], s2[G::NRT];
  double acc[G::NT][2];
  double w1;
};
// Per-block trend sums (see jne_consume8).
template <int DP> struct JneBlockSums { double bA[JneGeo<DP>::NRT], bB[JneGeo<DP>::NRT], bQ[JneGeo<DP>::NRT]; };

// One step s of a block: path update, operand exchange, the tile MMAs, the block-local trend sums.
// MASKED blocks (only the ragged tail of the last segment, or tiny T) zero the contributions of
// steps at or beyond t_end.
template <int DP, int DET, bool SRC_RNG, bool MASKED, int S, bool AUX = false>
__device__ __forceinline__ void jne_step(uint32_t t, uint32_t t_end, int g, int src_lane,
                                         const typename JneZ<DP, SRC_RNG>::type (&z)[JneGeo<DP>::NRT][8],
                                         JneLoopState<DP>& L, JneBlockSums<DP>& bs, double auxv = 0.0) {
  using G = JneGeo<DP>;
  constexpr int s = S;
  const bool active = !MASKED || (t + s) < t_end;
  double f[G::NRT], dz[G::NRT], cn[G::NRT];
#pragma unroll
  for (int j = 0; j < G::NRT; ++j) {
    f[j] = active ? L.c[j] : 0.0;
    dz[j] = active ? (double)z[j][s] : 0.0;
    cn[j] = L.c[j] + dz[j];                  // B_t = B_{t-1} + dB_t   (src/matrix_utils.rs:51-63)
    if (!SRC_RNG) dz[j] = cn[j] - L.c[j];    // dB re-derived by subtraction (src/johansen_statistics.rs:80-82)
  }
  // V tiles: index i = 8*jt + g;  i < DP -> F_i (own);  DP <= i < 2DP -> dB_{i-DP}, owned by lane
  // g' = (g - B) & 7 in its slot jt-A (receivers g >= B) or jt-A-1 (receivers g < B).  Every lane
  // publishes each of its increments once; the receiver picks.
  double recv[G::NRT];
#pragma unroll
  for (int m = 0; m < G::NRT; ++m) recv[m] = (G::B == 0) ? dz[m] : __shfl_sync(0xffffffffu, dz[m], src_lane);
  double V[G::NCT];
#pragma unroll
  for (int jt = 0; jt < G::NCT; ++jt) {
    const int i = 8 * jt + g;
    const int ja = jt - G::A, jb = jt - G::A - 1;
    const double ra = (ja >= 0 && ja < G::NRT) ? recv[(ja >= 0 && ja < G::NRT) ? ja : 0] : 0.0;
    const double rb = (jb >= 0 && jb < G::NRT) ? recv[(jb >= 0 && jb < G::NRT) ? jb : 0] : 0.0;
    const double r = (G::B == 0 || g >= G::B) ? ra : rb;
    const double own = (jt < G::NRT) ? f[jt < G::NRT ? jt : 0] : 0.0;
    V[jt] = (i < DP) ? own : ((i < 2 * DP) ? r : 0.0);
  }
  // AUX (DP = 4, 12: the group G::A holds four F rows and four dB rows): as the A operand of its own tiles that
  // group's dB half only produces dB x dB products nobody reads, so lanes g >= 4 put a trend weight of the step
  // there instead and the tensor pipe delivers sum_t weight_t dB_t' for all rows in the slots it wasted before
  // (rows 8 G::A + 4 + m of VV, m = weight index g - 4).  See make_aux_table (jne_api.cu) for the four weights.
  // AUX with DP = 8 (dim <= 6 only): rows 6 and 7 of F are padding, so lanes g = 6, 7 feed w1 and w2 of the step as
  // the A operand of both tiles: rows 6, 7 of VV then hold sum w c' and sum w dB' directly.
  constexpr int AUXG = (G::B == 4) ? G::A : 0;            // the operand group that carries the weights
  double Vx = V[AUXG < G::NCT ? AUXG : 0];
  if (AUX && DET >= 1) Vx = (g >= (G::B == 4 ? 4 : 6)) ? auxv : Vx;
  int ti = 0;
#pragma unroll
  for (int a = 0; a < G::NRT; ++a)
#pragma unroll
    for (int b = a; b < G::NCT; ++b) {
#ifdef JNE_EXP_NOMMA   // experiment only: no tensor work
      L.acc[ti][0] += V[a]; L.acc[ti][1] += V[b];
#else
      jne_dmma(L.acc[ti][0], L.acc[ti][1], (AUX && a == AUXG) ? Vx : V[a], V[b]);
#endif
      ++ti;
    }
  // deterministic cross moments of the path (those of the increments follow by summation by parts in
  // the epilogue: sum w z = w_last c_end - sum (w_t - w_{t-1}) c_t) and the running path.  The first term of
  // each block sum is assigned, not added to zero (x + 0 is not a no-op the compiler may drop).
#pragma unroll
  for (int j = 0; j < G::NRT; ++j) {
#ifdef JNE_EXP_NODET   // experiment only: no block-local trend sums
    if (s == 0) { bs.bA[j] = f[j]; bs.bB[j] = f[j]; bs.bQ[j] = f[j]; }
    L.c[j] = cn[j];
    continue;
#endif
    if (s == 0) bs.bA[j] = f[j]; else bs.bA[j] += f[j];   // every DET sums sum c the same way: records stay bit-identical across kernels
    if (!AUX) {
      if (DET >= 1 && s == 1) bs.bB[j] = f[j];
      if (DET >= 1 && s > 1) bs.bB[j] = fma((double)s, f[j], bs.bB[j]);
      if (DET >= 2 && s == 1) bs.bQ[j] = f[j];
      if (DET >= 2 && s > 1) bs.bQ[j] = fma((double)(s * s), f[j], bs.bQ[j]);
    }
    L.c[j] = cn[j];
  }
}

// Trend moments per 8-step block.  With w1_s = w1 + 2s and w2_s = 3 w1_s^2 + w2c = w2 + 12 w1 s + 12 s^2:
//   sum_s w1_s f_s = w1 A + 2 B,   sum_s w2_s f_s = w2 A + 12 w1 B + 12 Q,   A = sum f_s, B = sum s f_s, Q = sum s^2 f_s
// so the per-step work is one add and one or two FMAs with small exact multipliers, and the weights are
// touched once per block instead of three FP64 operations per lane and step.
template <int DP, int DET, bool AUX = false>
__device__ __forceinline__ void jne_block_end(JneLoopState<DP>& L, const JneBlockSums<DP>& bs, double w2c) {
  using G = JneGeo<DP>;
  if (DET == 0 || AUX) {
#pragma unroll
    for (int j = 0; j < G::NRT; ++j) L.s0[j] += bs.bA[j];
  } else {
    const double w1 = L.w1;
    const double w2 = fma(3.0 * w1, w1, w2c), w1x12 = 12.0 * w1;
#pragma unroll
    for (int j = 0; j < G::NRT; ++j) {
      L.s0[j] += bs.bA[j];
      L.s1[j] = fma(w1, bs.bA[j], L.s1[j]);
      L.s1[j] = fma(2.0, bs.bB[j], L.s1[j]);
      if (DET >= 2) {
        L.s2[j] = fma(w2, bs.bA[j], L.s2[j]);
        L.s2[j] = fma(w1x12, bs.bB[j], L.s2[j]);
        L.s2[j] = fma(12.0, bs.bQ[j], L.s2[j]);
      }
    }
    L.w1 += 16.0;
  }
}

// aux (AUX only): the lane's column of the weight table at the block's first step; a step is 16 doubles further on.
template <int DP, int DET, bool SRC_RNG, bool MASKED, bool AUX = false>
__device__ __forceinline__ void jne_consume8(uint32_t t, uint32_t t_end, int g, int src_lane,
                                             const typename JneZ<DP, SRC_RNG>::type (&z)[JneGeo<DP>::NRT][8],
                                             JneLoopState<DP>& L, double w2c, const double* __restrict__ aux = nullptr) {
  JneBlockSums<DP> bs;
  double av[8];
#pragma unroll
  for (int s = 0; s < 8; ++s) av[s] = (AUX && DET >= 1) ? __ldg(aux + 16 * s) : 0.0;
  jne_step<DP, DET, SRC_RNG, MASKED, 0, AUX>(t, t_end, g, src_lane, z, L, bs, av[0]);
  jne_step<DP, DET, SRC_RNG, MASKED, 1, AUX>(t, t_end, g, src_lane, z, L, bs, av[1]);
  jne_step<DP, DET, SRC_RNG, MASKED, 2, AUX>(t, t_end, g, src_lane, z, L, bs, av[2]);
  jne_step<DP, DET, SRC_RNG, MASKED, 3, AUX>(t, t_end, g, src_lane, z, L, bs, av[3]);
  jne_step<DP, DET, SRC_RNG, MASKED, 4, AUX>(t, t_end, g, src_lane, z, L, bs, av[4]);
  jne_step<DP, DET, SRC_RNG, MASKED, 5, AUX>(t, t_end, g, src_lane, z, L, bs, av[5]);
  jne_step<DP, DET, SRC_RNG, MASKED, 6, AUX>(t, t_end, g, src_lane, z, L, bs, av[6]);
  jne_step<DP, DET, SRC_RNG, MASKED, 7, AUX>(t, t_end, g, src_lane, z, L, bs, av[7]);
  jne_block_end<DP, DET, AUX>(L, bs, w2c);
}

// End of the time loop: the increment moments by summation by parts, then the lane-distributed raw moments go to
// the warp's shared memory (VV, vec: the raw view read by jne_warp_stitch).
template <int DP, bool AUX = false>
__device__ __forceinline__ void jne_warp_dump(JneLoopState<DP>& L, double* __restrict__ VV, double* __restrict__ vec,
                                              uint32_t t_begin, uint32_t t_end, double w1_first, double w2c, int g, int k) {
  using G = JneGeo<DP>;
  double (&c)[G::NRT] = L.c;
  double (&s0)[G::NRT] = L.s0;
  double (&s1)[G::NRT] = L.s1;
  double (&s2)[G::NRT] = L.s2;
  double (&acc)[G::NT][2] = L.acc;

  // sum w1 z and sum w2 z over the lane's segment by summation by parts (z_t = c_{t+1} - c_t, c_0 = 0,
  // w1_t - w1_{t-1} = 2, w2_t - w2_{t-1} = 12 w1_t - 12):
  //   u1 = w1_last c_end - 2 sum c,      u2 = w2_last c_end - 12 sum w1 c + 12 sum c
  double u1[G::NRT], u2[G::NRT];
  if (AUX) {   // the trend moments come out of the MMA (jne_step); vec rows 2..5 are not read
#pragma unroll
    for (int j = 0; j < G::NRT; ++j) { u1[j] = 0.0; u2[j] = 0.0; }
  } else {
    const double nseg = (double)(t_end - t_begin);
    const double w1_last = w1_first + 2.0 * (nseg - 1.0);
    const double w2_last = fma(3.0 * w1_last, w1_last, w2c);
#pragma unroll
    for (int j = 0; j < G::NRT; ++j) {
      const bool any = t_end > t_begin;
      u1[j] = any ? fma(w1_last, c[j], -2.0 * s0[j]) : 0.0;
      u2[j] = any ? fma(w2_last, c[j], 12.0 * (s0[j] - s1[j])) : 0.0;
    }
  }

  // ---- dump raw moments to the warp's shared memory ----
  {
    int ti = 0;
#pragma unroll
    for (int a = 0; a < G::NRT; ++a)
#pragma unroll
      for (int b = a; b < G::NCT; ++b) {
        VV[(8 * a + g) * G::VV_LD + 8 * b + 2 * k] = acc[ti][0];
        VV[(8 * a + g) * G::VV_LD + 8 * b + 2 * k + 1] = acc[ti][1];
        ++ti;
      }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int r = 8 * j + g;
      const bool has = j < G::NRT;
      vec[(0 * 4 + k) * 16 + r] = has ? c[has ? j : 0] : 0.0;
      vec[(1 * 4 + k) * 16 + r] = has ? s0[has ? j : 0] : 0.0;
      vec[(2 * 4 + k) * 16 + r] = has ? s1[has ? j : 0] : 0.0;
      vec[(3 * 4 + k) * 16 + r] = has ? s2[has ? j : 0] : 0.0;
      vec[(4 * 4 + k) * 16 + r] = has ? u1[has ? j : 0] : 0.0;
      vec[(5 * 4 + k) * 16 + r] = has ? u2[has ? j : 0] : 0.0;
    }
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// Fused per-run kernel.  SRC_RNG: increments come from the random stream keyed by seeds[run];
// otherwise from caller-supplied dB (n runs x (dim x steps) column-major per run, the layout of
// src/rng_matrix.rs:36) and the path is rebuilt exactly as src/johansen_statistics.rs:80-82 does.
// DET: 0 = models 0,1 (sum c only), 1 = models 2,3 (+ w1 moments), 2 = model 4 (+ w2 moments).
// ---------------------------------------------------------------------------------------------
#ifndef JNE_MULTI_MINB
#define JNE_MULTI_MINB 5
#endif
#ifdef JNE_MINB_OVERRIDE
#define JNE_MINB(MULTI, DET) JNE_MINB_OVERRIDE
#else
#define JNE_MINB(MULTI, DET) (((MULTI) ? JNE_MULTI_MINB : ((DET) == 0 ? 6 : 5)) * 4 / JNE_WARPS_PER_CTA)
#endif
// UNALIGNED (short horizons, seg_len_for in jne_api.cu): the segments are whole 8-step blocks, not whole generator epochs.
template <int DP, int DET, bool SRC_RNG, bool MULTI, bool AUXT = false, bool UNALIGNED = false>
__global__ void __launch_bounds__(32 * JNE_WARPS_PER_CTA, SRC_RNG ? JNE_MINB(MULTI, DET) : 1)
jne_run_kernel(const uint32_t* __restrict__ seeds, const double* __restrict__ dB, uint64_t n,
               JneRunParams prm, double* __restrict__ out, unsigned int* __restrict__ err_count,
               double* __restrict__ dbg /* optional: per run S2 (16x16) then R (16x16) */) {
  using G = JneGeo<DP>;
  using ZT = typename JneZ<DP, SRC_RNG>::type;
  // trend moments through the MMA (jne_step): DP = 4, 12 always; DP = 8 when the host guarantees dim <= 6
  constexpr bool AUX = AUXT && (G::B == 4 || DP == 8) && DET >= 1;
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint64_t run = (uint64_t)blockIdx.x * JNE_WARPS_PER_CTA + warp;
  if (run >= n) return;
  using E = JneEpi<DP, MULTI ? 5 : 1>;
  double* wsm = smem + (size_t)warp * E::WARP_SMEM;
  double* tot = wsm;                // whole-run totals (live through the epilogue)
  double* VV = tot + G::TOT_SZ;     // raw view
  double* vec = VV + G::VV_SZ;
  double* MBB = tot + G::TOT_SZ;    // stitched view (aliases the raw view, see jne_warp_stitch)
  double* MBZ = MBB + G::STITCH_HALF;

  const int g = lane >> 2, k = lane & 3;
  const uint32_t d = prm.dim, T = prm.steps;
  const uint32_t t_begin = min((uint32_t)k * prm.seg_len, T);
  const uint32_t t_end = min(T, t_begin + prm.seg_len);
  const uint32_t seed = SRC_RNG ? seeds[run] : 0u;
  JneGen<DP> gs;
  const double* dBrun = SRC_RNG ? nullptr : dB + run * (uint64_t)d * T;
  float rowscale[G::NRT];
#pragma unroll
  for (int j = 0; j < G::NRT; ++j) rowscale[j] = (8u * j + g < d) ? 1.0f : 0.0f;
  const float xscale = (8u * (G::NRT - 1) + (g & 3) < d) ? 1.0f : 0.0f;   // the shared row slot (DP = 4, 12)

  JneLoopState<DP> L;
#pragma unroll
  for (int j = 0; j < G::NRT; ++j) { L.c[j] = L.s0[j] = L.s1[j] = L.s2[j] = 0.0; }
#pragma unroll
  for (int i = 0; i < G::NT; ++i) { L.acc[i][0] = 0.0; L.acc[i][1] = 0.0; }

  const double w1_first = 2.0 * (double)t_begin + 1.0 - prm.T;
  L.w1 = w1_first;
  const double w2c = -(prm.T * prm.T - 1.0);
  const int src_lane = (((g - G::B) & 7) << 2) | k;

  // blocks in which every lane's eight steps are inside its segment need no masking
  ZT z[G::NRT][8];
  uint32_t t = t_begin;
  const uint32_t t_full = t_begin + 8u * prm.full_blocks, t_stop = t_begin + prm.seg_len;
  // weight table: [local step][weight m][segment k]; lane (g, k) reads weight m = g & 3 (used when g >= 4)
  // (DP = 8: lane g = 6 reads weight 3 = w1, lane g = 7 weight 2 = w2)
  const int aux_m = (G::B == 4) ? (g & 3) : (g == 7 ? 2 : 3);
  const double* aux = AUX ? prm.aux_tab + ((aux_m << 2) | k) : nullptr;
  // Long horizons: segments are whole epochs long (seg_len_for, jne_api.cu), every lane of the warp enters a new epoch
  // in the same block and the branch below is warp-uniform.  Short horizons (the UNALIGNED instance: a kernel of its
  // own, because the extra code cost the fused dim-12 loop 1.4 % through register allocation alone): segments of whole
  // 8-step blocks start inside an epoch; a lane keys the epoch it starts in, skips the blocks before its own (at most 15 per substream;
  // t_begin is a multiple of 8, so both halves are at the same block number) and from then on re-keys in its own phase.
  // The lanes of an empty segment (t_begin = T) compute substreams nobody reads.
  if (SRC_RNG && UNALIGNED && (t_begin & (JNE_EPOCH_STEPS - 1u)) != 0u) {
    jne_gen_seed<DP>(gs, seed, t_begin / JNE_EPOCH_STEPS, g);
    const uint32_t skip = ((t_begin >> 2) & (JNE_EPOCH_BLOCKS - 1u)) >> 1;
    for (uint32_t i = 0; i < skip; ++i) {
      jne_zt dump[4];
#pragma unroll
      for (int q = 0; q < JneGen<DP>::NST; ++q) jne_sub_normals4(gs.st[q], dump, 0.0f);
    }
  }
  for (; t < t_full; t += 8) {
    if (SRC_RNG && JNE_EPOCH_START(t, t_begin, UNALIGNED)) jne_gen_seed<DP>(gs, seed, t / JNE_EPOCH_STEPS, g);
    jne_gen8<DP, SRC_RNG>(t, t_end, d, g, gs, rowscale, xscale, dBrun, z);
    jne_consume8<DP, DET, SRC_RNG, false, AUX>(t, t_end, g, src_lane, z, L, w2c, aux);
    if (AUX) aux += 128;
  }
  for (; t < t_stop; t += 8) {
    if (SRC_RNG && JNE_EPOCH_START(t, t_begin, UNALIGNED)) jne_gen_seed<DP>(gs, seed, t / JNE_EPOCH_STEPS, g);
    jne_gen8<DP, SRC_RNG>(t, t_end, d, g, gs, rowscale, xscale, dBrun, z);
    jne_consume8<DP, DET, SRC_RNG, true, AUX>(t, t_end, g, src_lane, z, L, w2c, aux);
    if (AUX) aux += 128;
  }
  jne_warp_dump<DP, AUX>(L, VV, vec, t_begin, t_end, w1_first, w2c, g, k);
#ifdef JNE_EXP_NOEPI   // experiment only: time loop without the epilogue (NOT valid records)
  { double acc0 = 0.0;
    for (int e = lane; e < G::VV_SZ + G::VEC_SZ; e += 32) acc0 += VV[e];
    if (lane < (int)prm.out_stride) out[run * prm.out_stride + lane] = acc0;
    return; }
#endif
  // ---- stitch once, then per model: assemble + reduce to the Gram matrix; one Jacobi for all; emit ----
  // One Brownian path serves every selected model: the reference draws the path from (dim, steps, seed)
  // only (src/rng_matrix.rs:11) and its CLI loops the models over the same seeds (src/main.rs:109).
  jne_warp_stitch<DP, AUX>(VV, vec, tot, MBB, MBZ, prm);
  const bool ok = jne_warp_models<DP, MULTI ? 5 : 1>(wsm, prm, out + run * prm.out_stride,
                                                     dbg != nullptr ? dbg + run * 512 : nullptr);
  if (!ok && lane == 0) atomicAdd(err_count, 1u);
}

// ---------------------------------------------------------------------------------------------
// Standalone K3: batched pencil solve on caller-supplied matrices (parity with dggev, the routine
// behind GeneralizedEigen::new at src/johansen_statistics.rs:35-38).
//   S1 : n x (d x p) column-major per problem  (== p x d row-major S1')
//   S2 : n x (p x p) symmetric
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * JNE_WARPS_PER_CTA)
jne_pencil_kernel(const double* __restrict__ S1, const double* __restrict__ S2in, uint64_t n, int p, int d,
                  double factor, double* __restrict__ out, unsigned int* __restrict__ err_count,
                  const uint32_t* __restrict__ jtab) {
  using G = JneGeo<16>;
  using E = JneEpi<16, 1>;
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint64_t run = (uint64_t)blockIdx.x * JNE_WARPS_PER_CTA + warp;
  if (run >= n) return;
  double* wsm = smem + (size_t)warp * E::END;
  double* S2 = wsm + E::OFF_S2;
  double* R = wsm + E::OFF_R;
  double* Gm = wsm + E::OFF_G;
  for (int e = lane; e < p * p; e += 32) S2[(e / p) * G::LDW + (e % p)] = S2in[run * p * p + e];
  for (int e = lane; e < p * d; e += 32) R[(e / d) * G::LDW + (e % d)] = S1[run * p * d + e];
  __syncwarp();
  const int ne = (d + 1) & ~1;
  const double tr = jne_warp_gram<16>(S2, R, p, d, ne, Gm);
  __syncwarp();
  jne_warp_jacobi<E::NPASS_B, E::NPASS_R>(Gm, reinterpret_cast<double2*>(wsm + E::OFF_CS), jtab, 1, ne, ne * (ne + 1) / 2);
  const bool ok = jne_warp_emit(Gm, factor * tr, p, d, ne, wsm + E::OFF_EV, out + run * p);
  if (!ok && lane == 0) atomicAdd(err_count, 1u);
}

// ---------------------------------------------------------------------------------------------
// Exposed pieces of the stream (parity with the reference's RNG / Brownian tests).
//   jne_normal_matrix_kernel : gen_normal_matrix(nrows = d, ncols = T, seed)   src/rng_matrix.rs:11-37
//   jne_brownian_kernel      : brownian_motion_matrix(..., AlongColumns, 0)    src/rng_matrix.rs:57-141
// Output column-major d x T (resp. d x (T+1)), like DMatrix::from_vec.
// ---------------------------------------------------------------------------------------------
__global__ void jne_normal_matrix_kernel(uint32_t seed, uint32_t d, uint32_t T, double* __restrict__ out) {
  const uint32_t nb = (T + 3) / 4;
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (uint64_t)nb * d) return;
  const uint32_t row = idx % d, tb = idx / d;
  jne_zt z[4];
  jne_normals4(seed, row, tb, z);
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const uint32_t t = 4 * tb + s;
    if (t < T) out[(uint64_t)t * d + row] = (double)z[s];
  }
}

// One thread per row: sequential left-to-right cumsum of sqrt(dt) * z (src/matrix_utils.rs:51-63).
__global__ void jne_brownian_kernel(uint32_t seed, uint32_t d, uint32_t T, double delta_t, double* __restrict__ out) {
  const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= d) return;
  const double sq = sqrt(delta_t);
  double acc = 0.0;
  out[row] = acc;
  for (uint32_t tb = 0; tb < (T + 3) / 4; ++tb) {
    jne_zt z[4];
    jne_normals4(seed, row, tb, z);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const uint32_t t = 4 * tb + s;
      if (t < T) {   // scaled = z * sqrt(dt), then acc += scaled: two roundings, never fused (src/rng_matrix.rs:138-140)
        acc = __dadd_rn(acc, __dmul_rn((double)z[s], sq));
        out[(uint64_t)(t + 1) * d + row] = acc;
      }
    }
  }
}
