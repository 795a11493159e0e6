// Which instruction classes steal time from the DMMA / FP64 "shared" pipe on sm_100a?
// Each warp does rounds of 5 DMMA + N ops of class OP (4 independent chains); reports SMSP cycles per round.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
struct Out { double sink; long long cycles; };
// OP: 0 IMAD.WIDE 1 IMAD.HI+IMAD.LO 2 IMAD.LO 3 LOP3 4 FFMA 5 MUFU.EX2 6 F2F.F64.F32 7 DADD 8 SHFL 9 I2FP 10 IMAD.HI 11 DFMA 12 F2F via int trick(ALU)
template <int OP, int N, int ND>
__global__ void __launch_bounds__(256) k(int iters, double seed, Out* out) {
  double c[10]; for (int i = 0; i < 10; ++i) c[i] = 0;
  double a = seed + threadIdx.x, b = seed * 0.5;
  unsigned x[4] = {threadIdx.x + 1u, 2u * threadIdx.x + 3u, 77u + threadIdx.x, 99u};
  float f[4] = {1.0f + threadIdx.x * 1e-3f, 1.1f, 1.2f, 1.3f};
  double d[4] = {1.0 + threadIdx.x * 1e-3, 1.1, 1.2, 1.3};
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ND; ++i) dmma884(c[2 * (i % 5)], c[2 * (i % 5) + 1], a, b);
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const int q = j & 3;
      if (OP == 0) { unsigned hi, lo; asm volatile("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%1, %0}, p;\n\t}" : "=r"(hi), "=r"(lo) : "r"(x[q]), "r"(0xD2511F53u)); x[q] = hi ^ lo; }
      else if (OP == 1) { unsigned hi = __umulhi(x[q], 0xD2511F53u); unsigned lo = x[q] * 0xD2511F53u; asm volatile("" : "+r"(hi), "+r"(lo)); x[q] = hi ^ lo; }
      else if (OP == 2) { x[q] = x[q] * 0xD2511F53u + 12345u; }
      else if (OP == 3) { x[q] = (x[q] ^ x[(q + 1) & 3]) & 0xfffffff7u | 5u; asm volatile("" : "+r"(x[q])); }
      else if (OP == 4) { f[q] = fmaf(f[q], 1.0001f, 0.5f); }
      else if (OP == 5) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[q])); }
      else if (OP == 6) { d[q] = (double)f[q]; f[q] = __int_as_float(__double2hiint(d[q])); }
      else if (OP == 7) { d[q] = d[q] + 1.000001; }
      else if (OP == 8) { x[q] = __shfl_xor_sync(0xffffffffu, x[q], 16); }
      else if (OP == 9) { f[q] = __uint2float_rn(x[q]); x[q] = __float_as_uint(f[q]); }
      else if (OP == 10) { x[q] = __umulhi(x[q], 0xD2511F53u) + 1u; }
      else if (OP == 11) { d[q] = fma(d[q], 1.000001, 0.5); }
      else if (OP == 12) { unsigned bb = __float_as_uint(f[q]); unsigned hi = (bb & 0x80000000u) | (((bb & 0x7fffffffu) >> 3) + 0x38000000u); unsigned lo = bb << 29; d[q] = __hiloint2double(hi, lo); f[q] = __int_as_float(__double2hiint(d[q]) ^ lo); }
    }
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < 10; ++i) s += c[i];
  for (int i = 0; i < 4; ++i) s += x[i] + f[i] + d[i];
  if (s == 123.456) out[0].sink = s;
  if (threadIdx.x == 0) out[blockIdx.x].cycles = t1 - t0;
}
template <int OP, int N, int ND>
double run(Out* d_out, int nsm, int iters, float* ms_out) {
  int blocks = nsm * 4;   // 32 warps / SM
  k<OP, N, ND><<<blocks, 256>>>(100, 1.0000001, d_out);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  k<OP, N, ND><<<blocks, 256>>>(iters, 1.0000001, d_out);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); *ms_out = ms;
  // SMSP cycles per warp-round at the nominal clock, from wall time: 8 warps per SMSP
  return ms * 1e-3 * 1.965e9 / ((double)iters * 8.0);
}
#define ROW(OP, NAME) { float m0, m1, m2, m3, m4; \
  double a0 = run<OP, 16, 0>(d_out, nsm, iters, &m0); \
  double a1 = run<OP, 0, 5>(d_out, nsm, iters, &m1); \
  double a2 = run<OP, 8, 5>(d_out, nsm, iters, &m2); \
  double a3 = run<OP, 16, 5>(d_out, nsm, iters, &m3); \
  double a4 = run<OP, 32, 5>(d_out, nsm, iters, &m4); \
  printf("%-26s alone(16 ops)=%7.1f (%.2f/op) | 5DMMA+0=%6.1f  +8=%6.1f  +16=%6.1f  +32=%6.1f | marginal/op: %5.2f %5.2f %5.2f\n", NAME, a0, a0 / 16, a1, a2, a3, a4, (a2 - a1) / 8, (a3 - a1) / 16, (a4 - a1) / 32); }
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int nsm = p.multiProcessorCount, iters = 20000;
  Out* d_out; CK(cudaMalloc(&d_out, sizeof(Out) * nsm * 8));
  printf("cycles per warp-round per SMSP at 1.965 GHz-equivalent wall time, 8 warps/SMSP\n");
  ROW(0, "IMAD.WIDE.U32(+xor)");
  ROW(1, "IMAD.HI + IMAD.LO(+xor)");
  ROW(2, "IMAD.LO");
  ROW(10, "IMAD.HI(+add)");
  ROW(3, "LOP3 x2");
  ROW(4, "FFMA");
  ROW(5, "MUFU.EX2");
  ROW(6, "F2F.F64.F32");
  ROW(12, "f32->f64 by ALU bit trick");
  ROW(7, "DADD");
  ROW(11, "DFMA");
  ROW(8, "SHFL.BFLY");
  ROW(9, "I2FP.F32.U32");
  return 0;
}
