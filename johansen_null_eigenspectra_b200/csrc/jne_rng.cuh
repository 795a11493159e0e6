// Counter-based Gaussian stream, in registers (sm_100a).
//
// Replaces gen_normal_matrix (reference src/rng_matrix.rs:11-37: per-chunk Xoshiro256++ +
// ziggurat StandardNormal, whose stream depends on the machine's physical core count,
// :16-20).  Here element (row r, step t) of the d x T normal matrix is a pure function of
// (seed, r, t):   Philox4x32-10(key = (seed, JNE_KEY1), ctr = (t >> 2, r, 0, 0))
// yields four 32-bit words; words (0,1) -> Box-Muller pair for steps 4b, 4b+1 and words
// (2,3) -> steps 4b+2, 4b+3.  The stream does not depend on model, dim, batch size, GPU
// count or launch geometry (SURVEY.md section 8b "semantics that must hold").
//
// Pipes (measured, profiles/r1_microbench_pipes.txt): IMAD.WIDE 32 lanes/clk/SM, MUFU 16,
// F2F.F64.F32 16; FP64 37.0 TFLOP/s.  The transform therefore stays in FP32 + MUFU and never
// touches the FP64 pipe until the final widening.
#pragma once
#include <cstdint>

#define JNE_KEY1 0x4A4E4531u  // "JNE1"
// Philox rounds.  10 is the Random123 / cuRAND default and what this library ships; 7 is the smallest count the
// Random123 paper reports as Crush-resistant -- available as a build-time ablation only (profiles/: the lever table).
#ifndef JNE_PHILOX_ROUNDS
#define JNE_PHILOX_ROUNDS 10
#endif
// Type of a generated normal before it is widened for the FP64 accumulation: float in the product (FP32 + MUFU
// transform), double in the validation build (-DJNE_RNG_F64, see jne_box_muller_f64).
#ifdef JNE_RNG_F64
typedef double jne_zt;
#else
typedef float jne_zt;
#endif

struct jne_u4 { uint32_t x, y, z, w; };

// 32 x 32 -> (hi, lo) in one IMAD.WIDE.U32
__device__ __forceinline__ void jne_mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
  asm("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%1, %0}, p;\n\t}" : "=r"(hi), "=r"(lo) : "r"(a), "r"(b));
}

__device__ __forceinline__ jne_u4 jne_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                    uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < JNE_PHILOX_ROUNDS; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0;
    const uint64_t p1 = (uint64_t)M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ (k0 + (uint32_t)r * W0);
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ (k1 + (uint32_t)r * W1);
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
  }
  return jne_u4{c0, c1, c2, c3};
}

// Same function with the round keys of word 0 precomputed (key0[r] = seed + r * W0); the round keys of
// word 1 are compile-time constants.  Lets the per-run key schedule live in registers across the time loop.
#ifdef JNE_EXP_XOSHIRO
struct jne_keys { mutable uint32_t k[10]; };   // experiment: words 2..5 are a sequential generator's state
#else
struct jne_keys { uint32_t k[10]; };
#endif
// `stage` is 10 words of the warp's shared memory: the round trip through memory stops ptxas from
// rematerialising "seed + r*W0" inside the time loop (it re-added all nine keys on every call).
__device__ __forceinline__ jne_keys jne_make_keys(uint32_t seed, volatile uint32_t* stage) {
  jne_keys ks;
  const int lane = threadIdx.x & 31;
  if (lane < 10) stage[lane] = seed + (uint32_t)lane * 0x9E3779B9u;
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 10; ++r) ks.k[r] = stage[r];
  __syncwarp();
  return ks;
}
#ifdef JNE_EXP_THREEFRY   // experiment only (= number of rounds): Threefry4x32 (Random123) in place of Philox4x32-10
__device__ __forceinline__ jne_u4 jne_threefry4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                   uint32_t k0, uint32_t k1, uint32_t k2, uint32_t k3) {
  constexpr int R[8][2] = {{10, 26}, {11, 21}, {13, 27}, {23, 5}, {6, 20}, {17, 11}, {25, 10}, {18, 20}};
  const uint32_t ks[5] = {k0, k1, k2, k3, 0x1BD11BDAu ^ k0 ^ k1 ^ k2 ^ k3};
  uint32_t X0 = c0 + ks[0], X1 = c1 + ks[1], X2 = c2 + ks[2], X3 = c3 + ks[3];
#pragma unroll
  for (int r = 0; r < JNE_EXP_THREEFRY; ++r) {
    if ((r & 1) == 0) {
      X0 += X1; X1 = __funnelshift_l(X1, X1, R[r & 7][0]) ^ X0;
      X2 += X3; X3 = __funnelshift_l(X3, X3, R[r & 7][1]) ^ X2;
    } else {
      X0 += X3; X3 = __funnelshift_l(X3, X3, R[r & 7][0]) ^ X0;
      X2 += X1; X1 = __funnelshift_l(X1, X1, R[r & 7][1]) ^ X2;
    }
    if ((r & 3) == 3) {
      const int s = (r + 1) >> 2;
      X0 += ks[s % 5]; X1 += ks[(s + 1) % 5]; X2 += ks[(s + 2) % 5]; X3 += ks[(s + 3) % 5] + (uint32_t)s;
    }
  }
  return jne_u4{X0, X1, X2, X3};
}
#endif
__device__ __forceinline__ jne_u4 jne_philox4x32_10_keyed(uint32_t c0, uint32_t c1, const jne_keys& ks, uint32_t c2 = 0u) {
#ifdef JNE_EXP_THREEFRY
  return jne_threefry4x32(c0, c1, c2, 0u, ks.k[0], JNE_KEY1, 0u, 0u);
#endif
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W1 = 0xBB67AE85u;
  uint32_t c3 = 0u;
#pragma unroll
  for (int r = 0; r < JNE_PHILOX_ROUNDS; ++r) {
    uint32_t h0, l0, h1, l1;
    jne_mulhilo(M0, c0, h0, l0);
    jne_mulhilo(M1, c2, h1, l1);
    const uint32_t n0 = h1 ^ c1 ^ ks.k[r];
    const uint32_t n2 = h0 ^ c3 ^ (JNE_KEY1 + (uint32_t)r * W1);
    c1 = l1;
    c3 = l0;
    c0 = n0;
    c2 = n2;
  }
  return jne_u4{c0, c1, c2, c3};
}

// Two N(0,1) variates from two 32-bit words.  u = (wa + 1/2) 2^-32 in (0, 1], radius
// r = sqrt(-2 ln u) <= 6.76; angle theta = 2 pi * int32(wb) * 2^-32 in [-pi, pi) so the MUFU
// sin/cos see their most accurate range.
__device__ __forceinline__ void jne_box_muller(uint32_t wa, uint32_t wb, float& z0, float& z1, float scale = 1.0f) {
  const float u = fmaf(__uint2float_rn(wa), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
  float l, r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(u));   // MUFU.LG2 (u >= 2^-33: never denormal)
  // -2 ln 2 as a float is -0x1.62e430p+0; the constant below sits three ulps further out (+2.58e-7 relative).  It
  // cancels the variance deficit of this FP32 / MUFU pipeline, measured on 2^27 normals against the FP64 transform of
  // the same Philox blocks: E[z^2]_fp32 - E[z^2]_fp64 = -2.5265e-7 +- 8e-11 (profiles/r2_rng_moments_before_calibration.txt; every
  // eigenvalue statistic carried the same -2.53e-7 relative shift, profiles/r2_gate2_ab_*).  The eigenvalues scale
  // with Var(z), so this is the one moment worth calibrating; residual +5e-9.
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(l * -0x1.62e436p+0f));  // MUFU.SQRT
  r *= scale;   // 1 (or 0 for a row the run does not have: keeps the generation branch-free)
  const float th = __int2float_rn((int)wb) * 1.4629180792671596e-9f;  // 2 pi 2^-32
  z0 = r * __cosf(th);
  z1 = r * __sinf(th);
}

// VALIDATION STREAM (-DJNE_RNG_F64, never in libjne.so): the same counters, keys and word assignment, but the radius
// uniform and the angle carry 64 bits -- the product's 32-bit word on top, the matching word of a second Philox block
// (counter word 2 = 1) below -- and the transform runs in FP64 (log, sqrt, sincospi).  u = (x + 1/2) 2^-64, so the
// radius reaches 9.5 instead of 6.76 and the tail is not quantised at 2^-32.  Element (row, step) of this stream differs
// from the product's by the MUFU / FP32 rounding (~1e-6) and the low-order refinement only, which makes the A/B of
// tools/validate_gate2.py a PAIRED comparison of the statistics, run by run.
__device__ __forceinline__ void jne_box_muller_f64(uint32_t wa, uint32_t wa_lo, uint32_t wb, uint32_t wb_lo, double& z0,
                                                   double& z1, double scale) {
  const uint64_t xa = ((uint64_t)wa << 32) | wa_lo;
  const long long xb = (long long)(((uint64_t)wb << 32) | wb_lo);
  const double u = fma((double)xa, 0x1p-64, 0x1p-65);           // (0, 1]; log(1) = 0 gives radius 0
  const double r = sqrt(-2.0 * log(u)) * scale;
  double sn, cs;
  sincospi((double)xb * 0x1p-63, &sn, &cs);                     // angle = 2 pi xb 2^-64 in [-pi, pi)
  z0 = r * cs;
  z1 = r * sn;
}

// The four normals of (row, time block tb = t >> 2) for one seed.
__device__ __forceinline__ void jne_normals4(uint32_t seed, uint32_t row, uint32_t tb, jne_zt z[4]) {
  const jne_u4 w = jne_philox4x32_10(tb, row, 0u, 0u, seed, JNE_KEY1);
#ifdef JNE_RNG_F64
  const jne_u4 v = jne_philox4x32_10(tb, row, 1u, 0u, seed, JNE_KEY1);
  jne_box_muller_f64(w.x, v.x, w.y, v.y, z[0], z[1], 1.0);
  jne_box_muller_f64(w.z, v.z, w.w, v.w, z[2], z[3], 1.0);
#else
  jne_box_muller(w.x, w.y, z[0], z[1]);
  jne_box_muller(w.z, w.w, z[2], z[3]);
#endif
}
__device__ __forceinline__ void jne_normals4_keyed(const jne_keys& ks, uint32_t row, uint32_t tb, jne_zt* z,
                                                   float scale = 1.0f) {
#ifdef JNE_RNG_F64
  const jne_u4 w = jne_philox4x32_10_keyed(tb, row, ks), v = jne_philox4x32_10_keyed(tb, row, ks, 1u);
  jne_box_muller_f64(w.x, v.x, w.y, v.y, z[0], z[1], (double)scale);
  jne_box_muller_f64(w.z, v.z, w.w, v.w, z[2], z[3], (double)scale);
#else
#ifdef JNE_EXP_NORNG   // experiment only: no Philox / Box-Muller (NOT a valid stream)
  z[0] = scale * 0.5f; z[1] = -scale * (float)(row + 1) * 0.25f; z[2] = scale * 0.125f * (tb & 3); z[3] = -scale;
  return;
#endif
#ifdef JNE_EXP_XOSHIRO   // experiment only: xoshiro128++ advanced per call, (row, tb) ignored (NOT a valid stream): the
  {                      // cost ceiling of a sequential generator seeded once per run, row and segment
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t &s0 = ks.k[2], &s1 = ks.k[3], &s2 = ks.k[4], &s3 = ks.k[5];
      const uint32_t a = s0 + s3;
      w[i] = __funnelshift_l(a, a, 7) + s0;
      const uint32_t t = s1 << 9;
      s2 ^= s0; s3 ^= s1; s1 ^= s2; s0 ^= s3; s2 ^= t; s3 = __funnelshift_l(s3, s3, 11);
    }
    jne_box_muller(w[0], w[1], z[0], z[1], scale);
    jne_box_muller(w[2], w[3], z[2], z[3], scale);
    return;
  }
#endif
#ifdef JNE_EXP_NOBM    // experiment only: Philox but no Box-Muller
  { const jne_u4 w = jne_philox4x32_10_keyed(tb, row, ks);
    z[0] = scale * __int_as_float((w.x >> 9) | 0x3f800000); z[1] = scale * __int_as_float((w.y >> 9) | 0x3f800000);
    z[2] = scale * __int_as_float((w.z >> 9) | 0x3f800000); z[3] = scale * __int_as_float((w.w >> 9) | 0x3f800000); return; }
#endif
  const jne_u4 w = jne_philox4x32_10_keyed(tb, row, ks);
  jne_box_muller(w.x, w.y, z[0], z[1], scale);
  jne_box_muller(w.z, w.w, z[2], z[3], scale);
#endif
}
