// Second kernel family for 9 <= dim <= 12: register-tiled FP64 FMA accumulation, 4 lanes per run.
//
// Why: on the tensor path (jne_kernels.cuh) dim 12 needs five DMMA.8x8x4 per step-round for 222 useful products out
// of 320 (12 is not a multiple of 8 and the symmetric tiles are computed whole), which caps that formulation at
// ~55 % of the FP64 roofline however well the pipe is fed.  DFMA has the same peak as DMMA on B200
// (profiles/r1_microbench_pipes.txt), so a formulation with NO padded products has a higher ceiling, provided both
// operands of every product are in the lane's registers.  Layout: lane l (= lane & 3) of a 4-lane group owns Brownian
// rows 3l..3l+2 of ONE run (8 runs per warp) and accumulates
//     F_own (3) x dB (all 12)                                  36 products   -> sum B dB'
//     F_own x F_own (upper), x F_next (all), x F_next2 (lower)  21 products   -> sum B B' (each unordered pair once or twice)
// per step; the 9 foreign increments and 6 foreign path values arrive by 15 shuffles of a double.  Time runs
// sequentially (no segments, no stitching); the path's cumulative sum, sum B, sum w1 B, sum w2 B are per-lane FP64
// adds on the own rows only.  The moments go to global memory (608 doubles per run, the "stitched" layout of
// jne_kernels.cuh) and jne_solve_kernel runs the per-model assembly + Cholesky/Jacobi solve at full occupancy.
// The random stream is the same function of (seed, row, step) as everywhere else.
#pragma once
#include "../jne_kernels.cuh"

#define JNE_MOM_DOUBLES (2 * 256 + 6 * 16)   // MBB[16][16], MBZ[16][16], tot[6][16]
#ifndef JNE_V2_WARPS
#define JNE_V2_WARPS 4        // warps per CTA of the moments kernel (8 runs per warp)
#endif
#ifndef JNE_V2_MINB
#define JNE_V2_MINB 2         // min resident CTAs per SM (register cap = 65536 / (32 * WARPS * MINB))
#endif

#ifndef JNE_V2_SMEM
#define JNE_V2_SMEM 1         // 1: operands travel through shared memory (4 stores + 10 loads); 0: 30 shuffles
#endif

template <int DET, bool SRC_RNG, bool MASKED>
__device__ __forceinline__ void jne_v2_step(const double (&zs)[3], bool active, int src1, int src2, double* xch, double (&c)[3],
                                            double (&s0)[3], double (&s1)[3], double (&s2)[3], double (&azd)[3][4][3],
                                            double (&aown)[6], double (&an1)[3][3], double (&an2)[6], double& w1,
                                            double w2c) {
  double f[3], dz[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    f[a] = (!MASKED || active) ? c[a] : 0.0;
    dz[a] = (!MASKED || active) ? zs[a] : 0.0;
    const double cn = c[a] + dz[a];            // B_t = B_{t-1} + dB_t   (src/matrix_utils.rs:51-63)
    if (!SRC_RNG) dz[a] = cn - c[a];           // dB re-derived by subtraction (src/johansen_statistics.rs:80-82)
    c[a] = cn;
  }
  // increments of the other three lanes of the group (rows 3(l^x) + b) and path values of the next two lanes
  double r[3][3], fn1[3], fn2[3];
#if JNE_V2_SMEM
  {
    // xch: the WARP's exchange area, lane-major so that every access is bank-conflict free:
    //   double2 A[32] = (dz0, dz1), double B[32] = dz2, double2 C[32] = (f0, f1), double D[32] = f2.
    // All traffic stays inside the warp, whose shared-memory operations execute in program order.
    const int lane = threadIdx.x & 31;
    double2* A = reinterpret_cast<double2*>(xch);
    double* B = xch + 64;
    double2* C = reinterpret_cast<double2*>(xch + 96);
    double* D = xch + 160;
    A[lane] = make_double2(dz[0], dz[1]);
    B[lane] = dz[2];
    C[lane] = make_double2(f[0], f[1]);
    D[lane] = f[2];
    __syncwarp();
#pragma unroll
    for (int x = 1; x <= 3; ++x) {
      const double2 a = A[lane ^ x];
      r[x - 1][0] = a.x; r[x - 1][1] = a.y; r[x - 1][2] = B[lane ^ x];
    }
    const double2 a1 = C[src1], a2 = C[src2];
    fn1[0] = a1.x; fn1[1] = a1.y; fn1[2] = D[src1];
    fn2[0] = a2.x; fn2[1] = a2.y; fn2[2] = D[src2];
    __syncwarp();   // the next step's stores must not overtake lagging lanes' loads of this step
  }
#else
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    r[0][b] = __shfl_xor_sync(0xffffffffu, dz[b], 1);
    r[1][b] = __shfl_xor_sync(0xffffffffu, dz[b], 2);
    r[2][b] = __shfl_xor_sync(0xffffffffu, dz[b], 3);
    fn1[b] = __shfl_sync(0xffffffffu, f[b], src1);
    fn2[b] = __shfl_sync(0xffffffffu, f[b], src2);
  }
#endif
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      azd[a][0][b] = fma(f[a], dz[b], azd[a][0][b]);
      azd[a][1][b] = fma(f[a], r[0][b], azd[a][1][b]);
      azd[a][2][b] = fma(f[a], r[1][b], azd[a][2][b]);
      azd[a][3][b] = fma(f[a], r[2][b], azd[a][3][b]);
      an1[a][b] = fma(f[a], fn1[b], an1[a][b]);
    }
  }
  aown[0] = fma(f[0], f[0], aown[0]); aown[1] = fma(f[0], f[1], aown[1]); aown[2] = fma(f[0], f[2], aown[2]);
  aown[3] = fma(f[1], f[1], aown[3]); aown[4] = fma(f[1], f[2], aown[4]); aown[5] = fma(f[2], f[2], aown[5]);
  an2[0] = fma(f[0], fn2[0], an2[0]);
  an2[1] = fma(f[1], fn2[0], an2[1]); an2[2] = fma(f[1], fn2[1], an2[2]);
  an2[3] = fma(f[2], fn2[0], an2[3]); an2[4] = fma(f[2], fn2[1], an2[4]); an2[5] = fma(f[2], fn2[2], an2[5]);
  double w2 = 0.0;
  if (DET >= 2) w2 = fma(3.0 * w1, w1, w2c);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    s0[a] += f[a];
    if (DET >= 1) s1[a] = fma(w1, f[a], s1[a]);
    if (DET >= 2) s2[a] = fma(w2, f[a], s2[a]);
  }
  if (DET >= 1) w1 += 2.0;
}

template <int DET, bool SRC_RNG>
__global__ void __launch_bounds__(32 * JNE_V2_WARPS, JNE_V2_MINB)
jne_moments12_kernel(const uint32_t* __restrict__ seeds, const double* __restrict__ dB, uint64_t n, JneRunParams prm,
                     double* __restrict__ mom) {
  __shared__ uint32_t key_stage[JNE_V2_WARPS][32][10];
  __shared__ __align__(16) double xch_all[JNE_V2_WARPS][192];   // per warp: A[32] double2, B[32], C[32] double2, D[32]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int l = lane & 3;
  const uint64_t run_raw = ((uint64_t)blockIdx.x * JNE_V2_WARPS + warp) * 8 + (lane >> 2);
  const bool live = run_raw < n;
  const uint64_t run = live ? run_raw : n - 1;          // idle groups shadow the last run (shuffles stay convergent)
  const uint32_t d = prm.dim, T = prm.steps;

  // per-lane round keys through shared memory (see jne_make_keys: stops ptxas re-adding them in the loop)
  jne_keys keys;
  {
    const uint32_t seed = SRC_RNG ? seeds[run] : 0u;
    volatile uint32_t* st = key_stage[warp][lane];
#pragma unroll
    for (int r = 0; r < 10; ++r) st[r] = seed + (uint32_t)r * 0x9E3779B9u;
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 10; ++r) keys.k[r] = st[r];
  }
  const double* dBrun = SRC_RNG ? nullptr : dB + run * (uint64_t)d * T;
  float scale[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) scale[a] = (3u * l + a < d) ? 1.0f : 0.0f;
  const int src1 = (lane & ~3) | ((l + 1) & 3), src2 = (lane & ~3) | ((l + 2) & 3);
  double* xch = &xch_all[warp][0];

  double c[3] = {0, 0, 0}, s0[3] = {0, 0, 0}, s1[3] = {0, 0, 0}, s2[3] = {0, 0, 0};
  double azd[3][4][3], aown[6], an1[3][3], an2[6];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) { azd[a][0][b] = azd[a][1][b] = azd[a][2][b] = azd[a][3][b] = 0.0; an1[a][b] = 0.0; }
#pragma unroll
  for (int i = 0; i < 6; ++i) { aown[i] = 0.0; an2[i] = 0.0; }
  const double w1_first = 1.0 - prm.T;
  double w1 = w1_first;
  const double w2c = -(prm.T * prm.T - 1.0);

  const uint32_t nfull = T >> 2;
  uint32_t tb = 0;
  for (; tb < nfull; ++tb) {
    double z[3][4];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if constexpr (SRC_RNG) {
        float zf[4];
        jne_normals4_keyed(keys, 3 * l + a, tb, zf, scale[a]);
#pragma unroll
        for (int s = 0; s < 4; ++s) z[a][s] = (double)zf[s];
      } else {
#pragma unroll
        for (int s = 0; s < 4; ++s) z[a][s] = (3u * l + a < d) ? dBrun[(uint64_t)(4 * tb + s) * d + 3 * l + a] : 0.0;
      }
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const double zs[3] = {z[0][s], z[1][s], z[2][s]};
      jne_v2_step<DET, SRC_RNG, false>(zs, true, src1, src2, xch, c, s0, s1, s2, azd, aown, an1, an2, w1, w2c);
    }
  }
  if ((T & 3u) != 0u) {   // ragged tail: steps at or beyond T contribute nothing
    double z[3][4];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if constexpr (SRC_RNG) {
        float zf[4];
        jne_normals4_keyed(keys, 3 * l + a, tb, zf, scale[a]);
#pragma unroll
        for (int s = 0; s < 4; ++s) z[a][s] = (double)zf[s];
      } else {
#pragma unroll
        for (int s = 0; s < 4; ++s)
          z[a][s] = (3u * l + a < d && 4 * tb + s < T) ? dBrun[(uint64_t)(4 * tb + s) * d + 3 * l + a] : 0.0;
      }
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const double zs[3] = {z[0][s], z[1][s], z[2][s]};
      jne_v2_step<DET, SRC_RNG, true>(zs, 4 * tb + s < T, src1, src2, xch, c, s0, s1, s2, azd, aown, an1, an2, w1, w2c);
    }
  }
  if (!live) return;

  // ---- moments to global memory, "stitched" layout: MBB[16][16], MBZ[16][16], tot[6][16] ----
  double* M = mom + run * (uint64_t)JNE_MOM_DOUBLES;
  double* MBB = M;
  double* MBZ = M + 256;
  double* tot = M + 512;
  const int l1 = (l + 1) & 3, l2 = (l + 2) & 3;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int i = 3 * l + a;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      MBZ[i * 16 + 3 * l + b] = azd[a][0][b];
      MBZ[i * 16 + 3 * (l ^ 1) + b] = azd[a][1][b];
      MBZ[i * 16 + 3 * (l ^ 2) + b] = azd[a][2][b];
      MBZ[i * 16 + 3 * (l ^ 3) + b] = azd[a][3][b];
      MBB[i * 16 + 3 * l1 + b] = an1[a][b];          // block (l, l+1): written by lane l ...
      MBB[(3 * l1 + b) * 16 + i] = an1[a][b];        // ... together with its mirror
    }
  }
  {
    const int i0 = 3 * l;
    const double o[3][3] = {{aown[0], aown[1], aown[2]}, {aown[1], aown[3], aown[4]}, {aown[2], aown[4], aown[5]}};
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) MBB[(i0 + a) * 16 + i0 + b] = o[a][b];
    // block (l, l+2): this lane holds entries (a, b <= a); lane l+2 holds the transposed complement.  Each writes
    // its entries and their mirrors; the diagonal a == b is written twice with identical bits.
    const int j0 = 3 * l2;
    const double lo[6] = {an2[0], an2[1], an2[2], an2[3], an2[4], an2[5]};
    int idx = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b <= a; ++b) {
        MBB[(i0 + a) * 16 + j0 + b] = lo[idx];
        MBB[(j0 + b) * 16 + i0 + a] = lo[idx];
        ++idx;
      }
  }
  // totals; the increment moments follow by summation by parts (see jne_run_kernel)
  {
    const double w1_last = w1_first + 2.0 * (prm.T - 1.0);
    const double w2_last = fma(3.0 * w1_last, w1_last, w2c);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int i = 3 * l + a;
      tot[0 * 16 + i] = s0[a];
      tot[1 * 16 + i] = s1[a];
      tot[2 * 16 + i] = s2[a];
      tot[3 * 16 + i] = c[a];
      tot[4 * 16 + i] = fma(w1_last, c[a], -2.0 * s0[a]);
      tot[5 * 16 + i] = fma(w2_last, c[a], 12.0 * (s0[a] - s1[a]));
    }
  }
}

// Per-run epilogue for the v2 moments: loads the stitched layout and runs the same per-model assembly + solve as the
// tensor path (jne_warp_assemble / jne_warp_pencil_solve), one warp per run, at full occupancy.
template <bool MULTI>
__global__ void __launch_bounds__(32 * JNE_WARPS_PER_CTA)
jne_solve_kernel(const double* __restrict__ mom, uint64_t n, JneRunParams prm, double* __restrict__ out,
                 unsigned int* __restrict__ err_count, double* __restrict__ dbg) {
  using G = JneGeo<12>;
  using E = JneEpi<12, MULTI ? 5 : 1>;
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint64_t run = (uint64_t)blockIdx.x * JNE_WARPS_PER_CTA + warp;
  if (run >= n) return;
  double* wsm = smem + (size_t)warp * E::END;
  double* tot = wsm;
  double* MBB = tot + G::TOT_SZ;
  double* MBZ = MBB + G::STITCH_HALF;
  const int d = prm.dim;
  const double* M = mom + run * (uint64_t)JNE_MOM_DOUBLES;
  for (int e = lane; e < G::STITCH_HALF; e += 32) {   // 12 rows x 16 columns of each 16 x 16 global array
    const int i = e >> 4, j = e & 15;
    MBB[e] = (i < d && j < d) ? M[e] : 0.0;
    MBZ[e] = (i < d && j < d) ? M[256 + e] : 0.0;
  }
  for (int e = lane; e < 96; e += 32) tot[e] = ((e & 15) < d) ? M[512 + e] : 0.0;
  __syncwarp();
  const bool ok = jne_warp_models<12, MULTI ? 5 : 1>(wsm, prm, out + run * prm.out_stride,
                                                     dbg != nullptr ? dbg + run * 512 : nullptr);
  if (!ok && lane == 0) atomicAdd(err_count, 1u);
}

template <bool MULTI> constexpr size_t jne_solve_smem() {
  return (size_t)JNE_WARPS_PER_CTA * sizeof(double) * JneEpi<12, MULTI ? 5 : 1>::END;
}
