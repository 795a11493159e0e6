// Pipe-rate microbenchmarks for the design decisions in DESIGN.md (sm_100a).
// Measures per-SM per-clock rates (clock64 inside the kernel) and wall TFLOP/s
// (CUDA events) of: DFMA, DMMA (mma.sync f64), DFMA+DMMA mixed, IMAD.WIDE,
// MUFU (lg2/sin/sqrt), I2F, F2F.F64.F32, SHFL, and DMMA+INT mixes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

struct Out { double sink; long long cycles; };

// mode 0: DFMA x16 chains; 1: DMMA884 x8; 2: mixed 8 DMMA + NF DFMA per iter
template <int MODE, int NF>
__global__ void __launch_bounds__(256) k_fp64(int iters, double seed, Out* out) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = seed * (i + threadIdx.x);
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = 0.0;
  double a = seed + threadIdx.x, b = seed * 0.5;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    } else if (MODE == 1) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(c[2 * i], c[2 * i + 1], a, b);
    } else if (MODE == 2) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          dmma884(c[2 * i], c[2 * i + 1], a, b);
#pragma unroll
          for (int j = 0; j < NF; ++j) acc[(i * NF + j) & 15] = fma(acc[(i * NF + j) & 15], a, b);
        }
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i] + c[i];
  if (s == 123.456) out[0].sink = s;
  if (threadIdx.x == 0) out[blockIdx.x].cycles = t1 - t0;
}

template <int SHAPE>
__global__ void __launch_bounds__(256) k_dmma_big(int iters, double seed, Out* out) {
  double c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = 0;
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + i + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = seed * 0.5 + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (SHAPE == 8) dmma1688(c[i], a, b); else dmma16816(c[i], a, b);
      }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += c[i][j];
  if (s == 123.456) out[0].sink = s;
  if (threadIdx.x == 0) out[blockIdx.x].cycles = t1 - t0;
}

// integer / sfu / conversion / shuffle pipes. 8 independent chains, 32 ops per iter.
// OP 0: IMAD.WIDE.U32 (mulhi+lo chained, philox style) 1: LOP3 2: MUFU.LG2 3: MUFU.SIN 4: MUFU.SQRT(approx)
// 5: I2F.U32 6: F2F.F64.F32 7: SHFL.BFLY 8: FFMA 9: IADD3 10: MUFU.RSQ 11: MUFU.EX2 12: F2F.F32.F64
template <int OP>
__global__ void __launch_bounds__(256) k_misc(int iters, unsigned seed, Out* out) {
  unsigned x[8]; float f[8]; double d[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = seed * (i + 1) + threadIdx.x; f[i] = 1.0f + 0.001f * (i + threadIdx.x); d[i] = f[i]; }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (OP == 0) { unsigned long long p = (unsigned long long)x[i] * 0xD2511F53u; x[i] = (unsigned)(p >> 32) ^ (unsigned)p; }
        else if (OP == 1) { x[i] = (x[i] ^ x[(i + 1) & 7]) ^ seed; }
        else if (OP == 2) { f[i] = __log2f(f[i]) + 3.0f; }
        else if (OP == 3) { f[i] = __sinf(f[i]); }
        else if (OP == 4) { asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(f[i])); }
        else if (OP == 5) { f[i] = (float)x[i]; x[i] = __float_as_uint(f[i]) + i; }
        else if (OP == 6) { d[i] = (double)f[i]; f[i] = __int_as_float(__double2hiint(d[i]) ^ __double2loint(d[i])); }
        else if (OP == 7) { x[i] = __shfl_xor_sync(0xffffffffu, x[i], 16); }
        else if (OP == 8) { f[i] = fmaf(f[i], 1.0001f, 0.5f); }
        else if (OP == 9) { x[i] = x[i] + x[(i + 1) & 7] + seed; }
        else if (OP == 10) { asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(f[i])); }
        else if (OP == 11) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i])); }
        else if (OP == 12) { f[i] = (float)d[i]; d[i] = __hiloint2double(__float_as_int(f[i]), __float_as_int(f[i])); }
      }
  }
  long long t1 = clock64();
  unsigned s = 0; float fs = 0; double ds = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { s += x[i]; fs += f[i]; ds += d[i]; }
  if (s == 12345u && fs == 1.5f && ds == 2.5) out[0].sink = s;
  if (threadIdx.x == 0) out[blockIdx.x].cycles = t1 - t0;
}

// DMMA (5 per iter) co-issued with NI philox-like INT ops (IMAD.WIDE + LOP3 pairs) per iter.
template <int NI, bool USE_DFMA>
__global__ void __launch_bounds__(256) k_mix_int(int iters, double seed, Out* out) {
  double c[10]; double acc[8];
#pragma unroll
  for (int i = 0; i < 10; ++i) c[i] = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = seed * i;
  double a = seed + threadIdx.x, b = seed * 0.5;
  unsigned x[4] = {threadIdx.x, 2u * threadIdx.x, 77u, 99u};
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (!USE_DFMA) {
#pragma unroll
        for (int i = 0; i < 5; ++i) dmma884(c[2 * i], c[2 * i + 1], a, b);
      } else {
#pragma unroll
        for (int i = 0; i < 40; ++i) acc[i & 7] = fma(acc[i & 7], a, b);
      }
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        unsigned long long p = (unsigned long long)x[j & 3] * 0xD2511F53u;
        x[j & 3] = (unsigned)(p >> 32) ^ x[(j + 1) & 3] ^ (unsigned)p;
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 10; ++i) s += c[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  s += x[0] + x[1] + x[2] + x[3];
  if (s == 123.456) out[0].sink = s;
  if (threadIdx.x == 0) out[blockIdx.x].cycles = t1 - t0;
}

template <typename F>
static void run(const char* name, F launch, int blocks, int iters, double ops_per_thread_iter, double flop_per_op,
                Out* d_out, int nsm) {
  launch(blocks, 10);  // warm
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  launch(blocks, iters);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<Out> h(blocks);
  CK(cudaMemcpy(h.data(), d_out, sizeof(Out) * blocks, cudaMemcpyDeviceToHost));
  std::vector<long long> cyc(blocks);
  for (int i = 0; i < blocks; ++i) cyc[i] = h[i].cycles;
  std::sort(cyc.begin(), cyc.end());
  double med = (double)cyc[blocks / 2];
  double blocks_per_sm = (double)blocks / nsm;
  double total_ops = (double)blocks * 256.0 * iters * ops_per_thread_iter;   // lane-ops
  double ops_clk_sm = blocks_per_sm * 256.0 * iters * ops_per_thread_iter / med;
  printf("%-34s blocks/SM=%4.1f ms=%8.3f lane-ops/clk/SM=%8.2f  warp-instr/clk/SM=%6.3f  T(fl)op/s=%8.3f  eff.MHz=%6.0f\n",
         name, blocks_per_sm, ms, ops_clk_sm, ops_clk_sm / 32.0, total_ops * flop_per_op / (ms * 1e-3) / 1e12,
         med / (ms * 1e-3) / 1e6);
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int nsm = p.multiProcessorCount;
  printf("device %s  SMs=%d  clock=%d kHz  cc=%d.%d\n", p.name, nsm, p.clockRate, p.major, p.minor);
  Out* d_out; CK(cudaMalloc(&d_out, sizeof(Out) * nsm * 16));
  for (int bps : {1, 2, 4, 8}) {
    int blocks = nsm * bps;
    int iters = 20000;
    printf("---- %d blocks/SM x 256 threads ----\n", bps);
    run("DFMA", [&](int b, int it) { k_fp64<0, 0><<<b, 256>>>(it, 1.0000001, d_out); }, blocks, iters, 64, 2, d_out, nsm);
    // per thread-iter: 32 dmma warp-instr => per lane 32 * (256 FMA / 32 lanes) = 256 FMA lane-ops
    run("DMMA m8n8k4 (FMA-equiv lane-ops)", [&](int b, int it) { k_fp64<1, 0><<<b, 256>>>(it, 1.0000001, d_out); }, blocks, iters / 4, 32 * 8, 2, d_out, nsm);
    run("DMMA m16n8k8", [&](int b, int it) { k_dmma_big<8><<<b, 256>>>(it, 1.0000001, d_out); }, blocks, iters / 8, 32 * 32, 2, d_out, nsm);
    run("DMMA m16n8k16", [&](int b, int it) { k_dmma_big<16><<<b, 256>>>(it, 1.0000001, d_out); }, blocks, iters / 16, 32 * 64, 2, d_out, nsm);
    run("MIX 1 DMMA : 2 DFMA", [&](int b, int it) { k_fp64<2, 2><<<b, 256>>>(it, 1.0000001, d_out); }, blocks, iters / 4, 32 * (8 + 2), 2, d_out, nsm);
    run("MIX 1 DMMA : 4 DFMA", [&](int b, int it) { k_fp64<2, 4><<<b, 256>>>(it, 1.0000001, d_out); }, blocks, iters / 4, 32 * (8 + 4), 2, d_out, nsm);
    run("MIX 1 DMMA : 8 DFMA", [&](int b, int it) { k_fp64<2, 8><<<b, 256>>>(it, 1.0000001, d_out); }, blocks, iters / 4, 32 * (8 + 8), 2, d_out, nsm);
  }
  {
    int blocks = nsm * 4, iters = 20000;
    printf("---- misc pipes, 4 blocks/SM x 256 threads (lane-ops/clk/SM) ----\n");
    run("IMAD.WIDE.U32 (+xor)", [&](int b, int it) { k_misc<0><<<b, 256>>>(it, 12345u, d_out); }, blocks, iters, 32, 1, d_out, nsm);
    run("LOP3", [&](int b, int it) { k_misc<1><<<b, 256>>>(it, 12345u, d_out); }, blocks, iters, 32, 1, d_out, nsm);
    run("IADD3", [&](int b, int it) { k_misc<9><<<b, 256>>>(it, 12345u, d_out); }, blocks, iters, 32, 1, d_out, nsm);
    run("FFMA", [&](int b, int it) { k_misc<8><<<b, 256>>>(it, 12345u, d_out); }, blocks, iters, 32, 1, d_out, nsm);
    run("MUFU.LG2 (+FADD)", [&](int b, int it) { k_misc<2><<<b, 256>>>(it, 12345u, d_out); }, blocks, iters, 32, 1, d_out, nsm);
    run("MUFU.SIN (__sinf incl. range mul)", [&](int b, int it) { k_misc<3><<<b, 256>>>(it, 12345u, d_out); }, blocks, iters, 32, 1, d_out, nsm);
    run("MUFU.SQRT approx", [&](int b, int it) { k_misc<4><<<b, 256>>>(it, 12345u, d_out); }, blocks, iters, 32, 1, d_out, nsm);
    run("MUFU.RSQ approx", [&](int b, int it) { k_misc<10><<<b, 256>>>(it, 12345u, d_out); }, blocks, iters, 32, 1, d_out, nsm);
    run("MUFU.EX2 approx", [&](int b, int it) { k_misc<11><<<b, 256>>>(it, 12345u, d_out); }, blocks, iters, 32, 1, d_out, nsm);
    run("I2F.U32 (+IADD)", [&](int b, int it) { k_misc<5><<<b, 256>>>(it, 12345u, d_out); }, blocks, iters, 32, 1, d_out, nsm);
    run("F2F.F64.F32 (+LOP)", [&](int b, int it) { k_misc<6><<<b, 256>>>(it, 12345u, d_out); }, blocks, iters, 32, 1, d_out, nsm);
    run("F2F.F32.F64", [&](int b, int it) { k_misc<12><<<b, 256>>>(it, 12345u, d_out); }, blocks, iters, 32, 1, d_out, nsm);
    run("SHFL.BFLY", [&](int b, int it) { k_misc<7><<<b, 256>>>(it, 12345u, d_out); }, blocks, iters, 32, 1, d_out, nsm);
    printf("---- 5 DMMA (or 40 DFMA) + NI philox-style INT ops per round; FMA-equiv lane-ops ----\n");
    run("5 DMMA + 0 INT", [&](int b, int it) { k_mix_int<0, false><<<b, 256>>>(it, 1.0000001, d_out); }, blocks, iters / 4, 4 * 5 * 8, 2, d_out, nsm);
    run("5 DMMA + 16 INT", [&](int b, int it) { k_mix_int<16, false><<<b, 256>>>(it, 1.0000001, d_out); }, blocks, iters / 4, 4 * 5 * 8, 2, d_out, nsm);
    run("5 DMMA + 32 INT", [&](int b, int it) { k_mix_int<32, false><<<b, 256>>>(it, 1.0000001, d_out); }, blocks, iters / 4, 4 * 5 * 8, 2, d_out, nsm);
    run("5 DMMA + 48 INT", [&](int b, int it) { k_mix_int<48, false><<<b, 256>>>(it, 1.0000001, d_out); }, blocks, iters / 4, 4 * 5 * 8, 2, d_out, nsm);
    run("40 DFMA + 0 INT", [&](int b, int it) { k_mix_int<0, true><<<b, 256>>>(it, 1.0000001, d_out); }, blocks, iters / 4, 4 * 40, 2, d_out, nsm);
    run("40 DFMA + 16 INT", [&](int b, int it) { k_mix_int<16, true><<<b, 256>>>(it, 1.0000001, d_out); }, blocks, iters / 4, 4 * 40, 2, d_out, nsm);
    run("40 DFMA + 32 INT", [&](int b, int it) { k_mix_int<32, true><<<b, 256>>>(it, 1.0000001, d_out); }, blocks, iters / 4, 4 * 40, 2, d_out, nsm);
  }
  return 0;
}
