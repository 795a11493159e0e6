// Warp-specialised form of the fused per-run kernel (RNG path), sm_100a.
//
// Why: Philox's 32 x 32 -> 64 multiplies (IMAD.WIDE) share a dispatch path with the FP64 / DMMA pipe
// (profiles/r1_microbench_dmma_interference.txt), and in jne_run_kernel the generator and tensor phases of a warp
// overlap little with those of its neighbours (profiles/r1_exp_ablation_v2.txt: without the generator 299 k,
// without the MMAs 237 k, together 389 k SMSP-cycles per seed).  This family tests whether putting the two kinds
// of work into different warps of one persistent CTA decouples them:
//   consumer warps (JNE_WS_CONS, one run at a time each)  wait for a block of normals in a shared-memory ring,
//       run the eight steps of jne_consume8 on it and, at the end of the run, the usual epilogue;
//   producer warps (JNE_WS_PROD, one per SM sub-partition) fill the rings of their consumers with exactly the
//       values jne_gen8 would have produced in the consumer's lanes (same lane <-> (row, segment) map).
// A record is therefore bit-identical to jne_run_kernel's (checked).  Hand-over is one mbarrier pair (full / empty)
// per ring slot; the ring aliases the epilogue workspace of its consumer, so the consumer hands the last
// JNE_WS_SLOTS slots back only after its epilogue.
// MEASURED (profiles/r1_exp_warp_specialised.txt): consumers alone 3.87 M seeds/s, producers alone 4.8 M (one per
// sub-partition) / 7.0 M (two), together 2.43 - 2.76 M against 2.98 M for jne_run_kernel: under a saturated DMMA
// stream every IMAD.WIDE of a producer waits for about one DMMA slot, whichever warp issues it.  Not the default
// (JNE_KERNEL=ws selects it).
#pragma once
#include "../jne_kernels.cuh"

#ifndef JNE_WS_CONS
#define JNE_WS_CONS 20        // consumer warps per CTA: 5 per SM sub-partition
#endif
#ifndef JNE_WS_PROD
#define JNE_WS_PROD 4         // producer warps per CTA: 1 per SM sub-partition
#endif
#define JNE_WS_SLOTS 4        // ring depth, in 8-step blocks
#ifndef JNE_WS_EXP
#define JNE_WS_EXP 0          // timing experiments only (1: consumers without producers, 2: producers without consumers)
#endif
#define JNE_WS_SERVE (JNE_WS_CONS / JNE_WS_PROD)

__device__ __forceinline__ uint32_t jne_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void jne_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(jne_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void jne_mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(jne_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool jne_mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(jne_smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

// non-blocking probe (the producer moves on to its next consumer instead of waiting)
__device__ __forceinline__ bool jne_mbar_poll(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(jne_smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

template <int DP, bool MULTI> struct JneWs {
  using G = JneGeo<DP>;
  using E = JneEpi<DP, MULTI ? 5 : 1>;
  static constexpr int SLOT_F4 = G::NRT * 2 * 32;                               // float4 per slot: NRT x 8 floats per lane
  static constexpr int RING_DOUBLES = JNE_WS_SLOTS * SLOT_F4 * 2;
  static constexpr int WARP_SMEM = E::WARP_SMEM > G::TOT_SZ + RING_DOUBLES ? E::WARP_SMEM : G::TOT_SZ + RING_DOUBLES;
  static constexpr size_t CTA_SMEM = (size_t)JNE_WS_CONS * WARP_SMEM * sizeof(double) +
                                     (size_t)JNE_WS_CONS * 2 * JNE_WS_SLOTS * sizeof(uint64_t);
};

template <int DP, int DET, bool MULTI>
__global__ void __launch_bounds__(32 * (JNE_WS_CONS + JNE_WS_PROD), 1)
jne_run_kernel_ws(const uint32_t* __restrict__ seeds, uint64_t n, JneRunParams prm, double* __restrict__ out,
                  unsigned int* __restrict__ err_count) {
  using G = JneGeo<DP>;
  using W = JneWs<DP, MULTI>;
  extern __shared__ double smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)JNE_WS_CONS * W::WARP_SMEM);   // [consumer][full | empty][slot]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < JNE_WS_CONS * 2 * JNE_WS_SLOTS) jne_mbar_init(bars + threadIdx.x, 1u);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  const uint64_t ncons = (uint64_t)gridDim.x * JNE_WS_CONS;        // consumer warps in the grid: run r belongs to r % ncons
  const uint32_t NB = prm.seg_len / 8u;                            // blocks per run and lane
  const int g = lane >> 2, k = lane & 3;
  const uint32_t d = prm.dim, T = prm.steps;
  const uint32_t t_begin = min((uint32_t)k * prm.seg_len, T);
  const uint32_t t_end = min(T, t_begin + prm.seg_len);

  if (warp >= JNE_WS_CONS) {
    // ------------------------------- producer -------------------------------
    const int pw = warp - JNE_WS_CONS;
    float rowscale[G::NRT];
#pragma unroll
    for (int j = 0; j < G::NRT; ++j) rowscale[j] = (8u * j + g < d) ? 1.0f : 0.0f;
    const float xscale = (8u * (G::NRT - 1) + (g & 3) < d) ? 1.0f : 0.0f;
    uint64_t run[JNE_WS_SERVE];
    uint32_t blk[JNE_WS_SERVE], B[JNE_WS_SERVE];
    int live = 0;
#pragma unroll
    for (int i = 0; i < JNE_WS_SERVE; ++i) {
      run[i] = (uint64_t)blockIdx.x * JNE_WS_CONS + (uint32_t)(pw + JNE_WS_PROD * i);
      blk[i] = 0u; B[i] = 0u;
      live += run[i] < n;
    }
    while (live > 0) {
#pragma unroll
      for (int i = 0; i < JNE_WS_SERVE; ++i) {
        if (run[i] >= n) continue;
        const int cw = pw + JNE_WS_PROD * i;
        const uint32_t slot = B[i] % JNE_WS_SLOTS, parity = ((B[i] / JNE_WS_SLOTS) & 1u) ^ 1u;
        uint64_t* full = bars + (size_t)cw * 2 * JNE_WS_SLOTS + slot;
        if (JNE_WS_EXP == 1) return;                                         // experiment: consumers only
        if (JNE_WS_EXP != 2 && !jne_mbar_poll(full + JNE_WS_SLOTS, parity)) continue;   // slot still in use: serve the next consumer
        // First block of a further run: the ring aliases the consumer's epilogue workspace, so nothing may be written
        // before the consumer has handed back the LAST block of its previous run (it does so after the epilogue).
        // Runs shorter than the ring (NB < JNE_WS_SLOTS) leave free slots that the test above would let through.
        if (JNE_WS_EXP != 2 && blk[i] == 0u && B[i] > 0u) {
          const uint32_t Bp = B[i] - 1u;
          if (!jne_mbar_poll(full + JNE_WS_SLOTS - slot + Bp % JNE_WS_SLOTS, (Bp / JNE_WS_SLOTS) & 1u)) continue;
        }
        jne_keys ks;
        const uint32_t seed = seeds[run[i]];
#pragma unroll
        for (int r = 0; r < 10; ++r) ks.k[r] = seed + (uint32_t)r * 0x9E3779B9u;
        float z[G::NRT][8];
        jne_gen8<DP, true>(t_begin + 8u * blk[i], t_end, d, g, ks, rowscale, xscale, nullptr, z);
        float4* ring = reinterpret_cast<float4*>(smem + (size_t)cw * W::WARP_SMEM + G::TOT_SZ) + (size_t)slot * W::SLOT_F4;
#pragma unroll
        for (int j = 0; j < G::NRT; ++j) {
          ring[(2 * j) * 32 + lane] = make_float4(z[j][0], z[j][1], z[j][2], z[j][3]);
          ring[(2 * j + 1) * 32 + lane] = make_float4(z[j][4], z[j][5], z[j][6], z[j][7]);
        }
        __syncwarp();
        if (lane == 0) jne_mbar_arrive(full);
        ++B[i];
        if (++blk[i] == NB) {
          blk[i] = 0u;
          run[i] += ncons;
          live -= run[i] >= n;
        }
      }
    }
    return;
  }

  // ------------------------------- consumer -------------------------------
  double* wsm = smem + (size_t)warp * W::WARP_SMEM;
  double* tot = wsm;
  double* VV = tot + G::TOT_SZ;
  double* vec = VV + G::VV_SZ;
  double* MBB = tot + G::TOT_SZ;
  double* MBZ = MBB + G::STITCH_HALF;
  const float4* ring = reinterpret_cast<const float4*>(wsm + G::TOT_SZ);
  uint64_t* full = bars + (size_t)warp * 2 * JNE_WS_SLOTS;
  uint64_t* empty = full + JNE_WS_SLOTS;
  const double w2c = -(prm.T * prm.T - 1.0);
  const int src_lane = (((g - G::B) & 7) << 2) | k;
  uint32_t B = 0u;                                                   // blocks consumed so far (slot and phase)
  for (uint64_t run = (uint64_t)blockIdx.x * JNE_WS_CONS + (uint32_t)warp; run < n; run += ncons) {
    JneLoopState<DP> L;
#pragma unroll
    for (int j = 0; j < G::NRT; ++j) { L.c[j] = L.s0[j] = L.s1[j] = L.s2[j] = 0.0; }
#pragma unroll
    for (int i = 0; i < G::NT; ++i) { L.acc[i][0] = 0.0; L.acc[i][1] = 0.0; }
    const double w1_first = 2.0 * (double)t_begin + 1.0 - prm.T;
    L.w1 = w1_first;
    for (uint32_t b = 0; b < NB; ++b, ++B) {
      const uint32_t slot = B % JNE_WS_SLOTS, parity = (B / JNE_WS_SLOTS) & 1u;
      if (JNE_WS_EXP == 2) return;                                          // experiment: producers only
      while (JNE_WS_EXP != 1 && !jne_mbar_test(full + slot, parity)) {}
      float z[G::NRT][8];
      const float4* rs = ring + (size_t)slot * W::SLOT_F4;
#pragma unroll
      for (int j = 0; j < G::NRT; ++j) {
        const float4 a = rs[(2 * j) * 32 + lane], c = rs[(2 * j + 1) * 32 + lane];
        z[j][0] = a.x; z[j][1] = a.y; z[j][2] = a.z; z[j][3] = a.w;
        z[j][4] = c.x; z[j][5] = c.y; z[j][6] = c.z; z[j][7] = c.w;
      }
      const uint32_t t = t_begin + 8u * b;
      if (b < prm.full_blocks) jne_consume8<DP, DET, true, false>(t, t_end, g, src_lane, z, L, w2c);
      else jne_consume8<DP, DET, true, true>(t, t_end, g, src_lane, z, L, w2c);
      // the ring aliases the epilogue workspace: the last JNE_WS_SLOTS blocks of a run are handed back after it
      if (b + JNE_WS_SLOTS < NB) {
        __syncwarp();
        if (lane == 0) jne_mbar_arrive(empty + slot);
      }
    }
    jne_warp_dump<DP>(L, VV, vec, t_begin, t_end, w1_first, w2c, g, k);
    jne_warp_stitch<DP>(VV, vec, tot, MBB, MBZ, prm);
    const bool ok = jne_warp_models<DP, MULTI ? 5 : 1>(wsm, prm, out + run * prm.out_stride, nullptr);
    if (!ok && lane == 0) atomicAdd(err_count, 1u);
    __syncwarp();
    if (lane == 0) {
      const uint32_t held = NB < JNE_WS_SLOTS ? NB : JNE_WS_SLOTS;
      for (uint32_t h = held; h > 0; --h) jne_mbar_arrive(empty + (B - h) % JNE_WS_SLOTS);
    }
  }
}
