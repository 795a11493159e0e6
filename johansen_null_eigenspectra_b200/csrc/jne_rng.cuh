// Gaussian stream of a run, in registers (sm_100a).
//
// Replaces gen_normal_matrix (reference src/rng_matrix.rs:11-37: per-chunk Xoshiro256++ +
// ziggurat StandardNormal, whose stream depends on the machine's physical core count,
// :16-20).  Here element (row r, step t) of the d x T normal matrix is a pure function of
// (seed, r, t) -- stream "JNE2":
//   * time is cut into EPOCHS of 128 steps = 32 four-step blocks; block b = t >> 2 belongs to epoch e = b >> 5 and
//     to HALF h = b & 1 of it, where it is block number j = (b & 31) >> 1 of that half;
//   * every (row, epoch, half) owns a SUBSTREAM: the counter-based Philox4x32-10(key = (seed, "JNE2"),
//     ctr = (e, r, h, 0)) yields the 128-bit state of a xoshiro128++ generator (Blackman & Vigna), whose outputs
//     4j .. 4j+3 are the four words of block j: words (0,1) -> Box-Muller pair for steps 4b, 4b+1, words (2,3) ->
//     steps 4b+2, 4b+3.
// The stream does not depend on model, dim, batch size, GPU count or launch geometry (SURVEY.md section 8b
// "semantics that must hold"); element (r, t) costs one Philox call and at most 16 generator steps to reach.
//
// Why not a Philox call per block (the stream of round 1, "JNE1"): IMAD.WIDE.U32 executes on the FP64 datapath
// (profiles/r1_microbench_dmma_interference.txt), so the 20 wide multiplies per four normals competed with the DMMA /
// DFMA work the kernels are bound by: the generator cost 36 % of the dim-12 throughput.  ARX counter generators move
// the cost to the issue slots instead (Threefry4x32-20: -10 %, Threefry4x32-12: +1.6 %); a xoshiro128++ step is 8
// ALU instructions per word (profiles/r2_variants_generators.txt).  Philox stays where random access is needed -- the
// substream keys -- at 1/16 of its former rate.  Two substreams per row and epoch (the halves) let two lanes of the
// tensor kernels share a row without exchanging generator state.
//
// Pipes (measured, profiles/r1_microbench_pipes.txt): MUFU 16 lanes/clk/SM, F2F.F64.F32 16; FP64 37.0 TFLOP/s.  The
// transform stays in FP32 + MUFU and never touches the FP64 pipe until the final widening.
#pragma once
#include <cstdint>

#define JNE_KEY1 0x4A4E4532u  // "JNE2"
#define JNE_EPOCH_STEPS 128u  // steps per epoch; JNE_EPOCH_BLOCKS four-step blocks, half of them per substream
#define JNE_EPOCH_BLOCKS 32u
// Philox rounds of the substream keys.  10 is the Random123 / cuRAND default.
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

// One step of xoshiro128++ (Blackman & Vigna, "Scrambled linear pseudorandom number generators", 2021).
__device__ __forceinline__ uint32_t jne_xoshiro128pp(uint32_t& s0, uint32_t& s1, uint32_t& s2, uint32_t& s3) {
  const uint32_t a = s0 + s3;
  const uint32_t out = __funnelshift_l(a, a, 7) + s0;
  const uint32_t t = s1 << 9;
  s2 ^= s0;
  s3 ^= s1;
  s1 ^= s2;
  s0 ^= s3;
  s2 ^= t;
  s3 = __funnelshift_l(s3, s3, 11);
  return out;
}

// State of one substream (row, epoch, half).  The validation build carries a second generator for the low halves of
// its 64-bit uniforms (substream counter word 3 = 1).
struct jne_sub {
  uint32_t s0, s1, s2, s3;
#ifdef JNE_RNG_F64
  uint32_t l0, l1, l2, l3;
#endif
};

__device__ __forceinline__ void jne_sub_seed(jne_sub& st, uint32_t seed, uint32_t row, uint32_t epoch, uint32_t half) {
  jne_u4 w = jne_philox4x32_10(epoch, row, half, 0u, seed, JNE_KEY1);
  if ((w.x | w.y | w.z | w.w) == 0u) w.x = 1u;   // the generator's one forbidden state (probability 2^-128)
  st.s0 = w.x; st.s1 = w.y; st.s2 = w.z; st.s3 = w.w;
#ifdef JNE_RNG_F64
  jne_u4 v = jne_philox4x32_10(epoch, row, half, 1u, seed, JNE_KEY1);
  if ((v.x | v.y | v.z | v.w) == 0u) v.x = 1u;
  st.l0 = v.x; st.l1 = v.y; st.l2 = v.z; st.l3 = v.w;
#endif
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
  // the same uniform words: E[z^2]_fp32 - E[z^2]_fp64 = -2.5265e-7 +- 8e-11 (profiles/r2_rng_moments_before_calibration.txt; every
  // eigenvalue statistic carried the same -2.53e-7 relative shift, profiles/r2_gate2_ab_*).  The eigenvalues scale
  // with Var(z), so this is the one moment worth calibrating; residual +5e-9.
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(l * -0x1.62e436p+0f));  // MUFU.SQRT
  r *= scale;   // 1 (or 0 for a row the run does not have: keeps the generation branch-free)
  const float th = __int2float_rn((int)wb) * 1.4629180792671596e-9f;  // 2 pi 2^-32
  z0 = r * __cosf(th);
  z1 = r * __sinf(th);
}

// VALIDATION STREAM (-DJNE_RNG_F64, never in libjne.so): the same substreams and word assignment, but the radius
// uniform and the angle carry 64 bits -- the product's 32-bit word on top, the matching word of a second generator
// (substream key with counter word 3 = 1) below -- and the transform runs in FP64 (log, sqrt, sincospi).  u = (x + 1/2) 2^-64, so the
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

// The next four-step block of a substream: four words -> four normals (times `scale`: 1, or 0 for a row the run
// does not have, which keeps the generation branch-free).
__device__ __forceinline__ void jne_sub_normals4(jne_sub& st, jne_zt* z, float scale = 1.0f) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) w[i] = jne_xoshiro128pp(st.s0, st.s1, st.s2, st.s3);
#ifdef JNE_RNG_F64
  uint32_t v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = jne_xoshiro128pp(st.l0, st.l1, st.l2, st.l3);
  jne_box_muller_f64(w[0], v[0], w[1], v[1], z[0], z[1], (double)scale);
  jne_box_muller_f64(w[2], v[2], w[3], v[3], z[2], z[3], (double)scale);
#else
#ifdef JNE_EXP_NORNG   // experiment only: no generator, no transform (NOT a valid stream)
  z[0] = scale * 0.5f; z[1] = -scale * 0.25f; z[2] = scale * 0.125f; z[3] = -scale;
  return;
#endif
#ifdef JNE_EXP_NOBM    // experiment only: the generator without the normal transform (NOT a valid stream)
  z[0] = scale * __int_as_float((w[0] >> 9) | 0x3f800000); z[1] = scale * __int_as_float((w[1] >> 9) | 0x3f800000);
  z[2] = scale * __int_as_float((w[2] >> 9) | 0x3f800000); z[3] = scale * __int_as_float((w[3] >> 9) | 0x3f800000);
  return;
#endif
  jne_box_muller(w[0], w[1], z[0], z[1], scale);
  jne_box_muller(w[2], w[3], z[2], z[3], scale);
#endif
}

// Random access (test entries, never on the hot path): the four normals of (row, block tb = t >> 2).
__device__ __forceinline__ void jne_normals4(uint32_t seed, uint32_t row, uint32_t tb, jne_zt z[4]) {
  jne_sub st;
  jne_sub_seed(st, seed, row, tb / JNE_EPOCH_BLOCKS, tb & 1u);
  const uint32_t j = (tb % JNE_EPOCH_BLOCKS) >> 1;
  for (uint32_t i = 0; i < j; ++i) {
    jne_zt skip[4];
    jne_sub_normals4(st, skip);
  }
  jne_sub_normals4(st, z);
}

