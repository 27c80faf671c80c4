// f32x2.cu -- issue/pipe throughput of FFMA vs FFMA2 (packed fp32) vs MUFU on one B200, to state the bound of
// the packed epoch sweep (csrc/epoch_adam.cu).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_f32x2 tools/ubench/f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
  float2 x[8];
  unsigned y[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i); y[i] = threadIdx.x + i; }
  const float2 A = make_float2(a, a), B = make_float2(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { x[i].x = __fmaf_rn(x[i].x, a, b); x[i].y = __fmaf_rn(x[i].y, a, b); }       // 2 FFMA
      if (MODE == 1) { x[i] = __ffma2_rn(x[i], A, B); }                                            // 1 FFMA2
      if (MODE == 2) { asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x[i].x)); asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x[i].y)); }  // 2 MUFU
      if (MODE == 3) { x[i] = __ffma2_rn(x[i], A, B); asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x[i].x)); }  // mix: 1 FFMA2 + 1 MUFU
      if (MODE == 4) { x[i] = __fmul2_rn(x[i], A); x[i] = __fadd2_rn(x[i], B); }                   // FMUL2 + FADD2: ptxas contracts them to ONE FFMA2
      if (MODE == 5) { x[i] = __ffma2_rn(x[i], A, B); y[i] = (y[i] ^ (y[i] << 1)) + 0x9e3779b9u; }   // FFMA2 + 2 independent ALU ops (LOP3/IADD)
      if (MODE == 6) { x[i].x = fminf(fminf(x[i].x, a), x[i].y); x[i].y = fmaxf(x[i].y, b); }       // FMNMX3 + FMNMX
      if (MODE == 7) { x[i] = __fmul2_rn(x[i], A); }                                               // FMUL2 alone
      if (MODE == 8) { x[i] = __fadd2_rn(x[i], B); }                                               // FADD2 alone
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y + (MODE == 5 ? (float)y[i] : 0.f);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, double inst_per_iter_thread) {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const int grid = sms * 8, iters = 20000;
  float* out; cudaMalloc(&out, grid * 256 * sizeof(float));
  k<MODE><<<grid, 256>>>(out, 100, 0.999f, 1e-3f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<MODE><<<grid, 256>>>(out, iters, 0.999f, 1e-3f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double warp_inst = (double)grid * 8 * iters * inst_per_iter_thread;  // 8 warps per CTA
  printf("{\"kernel\": \"%s\", \"ms\": %.3f, \"warp_instr_per_clk_per_sm_at_max_clock\": %.3f, \"max_clock_mhz\": %d}\n", name, ms,
         warp_inst / (ms * 1e-3) / sms / (khz * 1e3), khz / 1000);
  cudaFree(out);
}
int main() {
  run<0>("FFMA x16 per iter (scalar)", 16);
  run<1>("FFMA2 x8 per iter (packed)", 8);
  run<2>("MUFU.RSQ x16 per iter", 16);
  run<3>("FFMA2 x8 + MUFU x8 per iter", 16);
  run<4>("FMUL2+FADD2 x8 per iter (contracted by ptxas: 8 FFMA2)", 8);
  run<5>("FFMA2 x8 + ALU x16 per iter", 24);
  run<6>("FMNMX3 x8 + FMNMX x8 per iter", 16);
  run<7>("FMUL2 x8 per iter", 8);
  run<8>("FADD2 x8 per iter", 8);
  return 0;
}
