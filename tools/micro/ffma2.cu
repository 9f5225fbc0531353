// Microbenchmark (development): rates of the packed FP32 instructions of sm_100a (FFMA2 / FADD2 / FMUL2) alone and mixed.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2 tools/micro/ffma2.cu && /tmp/ffma2
// Result on B200 (profiles/r1_ffma2_microbench.txt): FFMA2 alone issues every 2nd cycle per scheduler (same FMA rate as scalar
// FFMA), but FADD2 + FFMA2 alternate at one packed instruction per cycle: a subtract-then-square distance loop doubles.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 v) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }
#define FMA2(p, a, b) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p) : "l"(a), "l"(b))
#define ADD2(p, a) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(a))
#define MUL2(p, a) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(a))

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s) {
  float a[16];
  u64 p[8], q[8];
  for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 0.001f + i;
  for (int i = 0; i < 8; i++) p[i] = pk(a[2 * i], a[2 * i + 1]), q[i] = pk(a[2 * i + 1], a[2 * i]);
  const u64 ss = pk(s * 1e-6f, s * 1e-6f);
  const float c = 0.999f;
  const u64 cc = pk(c, c);
  for (int t = 0; t < iters; t++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) { a[2 * i] = __fmaf_rn(a[2 * i], c, s); a[2 * i + 1] = __fmaf_rn(a[2 * i + 1], c, s); }   // 16 FFMA
      if (MODE == 1) { FMA2(p[i], cc, ss); }                                 // 8 FFMA2
      if (MODE == 2) { ADD2(q[i], ss); FMA2(p[i], cc, ss); }                 // 8 FADD2 + 8 FFMA2
      if (MODE == 3) { ADD2(p[i], ss); }                                     // 8 FADD2
      if (MODE == 4) { MUL2(p[i], cc); }                                     // 8 FMUL2
      if (MODE == 5) { MUL2(q[i], cc); FMA2(p[i], cc, ss); }                 // 8 FMUL2 + 8 FFMA2
      if (MODE == 6) { ADD2(q[i], ss); MUL2(p[i], cc); }                     // 8 FADD2 + 8 FMUL2
      if (MODE == 7) { a[i] = __fmaf_rn(a[i], c, s); FMA2(p[i], cc, ss); }   // 8 FFMA + 8 FFMA2
      if (MODE == 8) { a[i] = __fadd_rn(a[i], s); FMA2(p[i], cc, ss); }      // 8 FADD + 8 FFMA2
      if (MODE == 9) { ADD2(q[i], ss); ADD2(p[i], ss); }                     // 16 FADD2
      if (MODE == 10) { a[i] = fmaxf(a[i], s); FMA2(p[i], cc, ss); }         // 8 FMNMX + 8 FFMA2
      if (MODE == 11) { ADD2(q[i], ss); FMA2(p[i], cc, ss); a[i] = fmaxf(a[i], s); }  // FADD2 + FFMA2 + FMNMX
    }
  }
  float r = 0;
  for (int i = 0; i < 16; i++) r += a[i];
  for (int i = 0; i < 8; i++) r += lo(p[i]) + lo(q[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, double instr_per_iter) {
  float* out;
  const int blocks = 148 * 8, iters = 20000;
  cudaMalloc(&out, blocks * 256 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<blocks, 256>>>(out, iters, 1.0f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(out, iters, 1.0f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double winstr = (double)blocks * 8 * iters * instr_per_iter;   // warp instructions
  printf("%-28s %8.3f ms  %6.3f warp-instr / clk / scheduler (at %d MHz nominal)\n", name, ms, winstr / (148.0 * 4) / (ms * 1e-3 * clk * 1e3), clk / 1000);
  cudaFree(out);
}
int main() {
  run<0>("16 FFMA", 16);
  run<1>("8 FFMA2", 8);
  run<2>("8 FADD2 + 8 FFMA2", 16);
  run<3>("8 FADD2", 8);
  run<4>("8 FMUL2", 8);
  run<5>("8 FMUL2 + 8 FFMA2", 16);
  run<6>("8 FADD2 + 8 FMUL2", 16);
  run<7>("8 FFMA + 8 FFMA2", 16);
  run<8>("8 FADD + 8 FFMA2", 16);
  run<9>("16 FADD2", 16);
  run<10>("8 FMNMX + 8 FFMA2", 16);
  run<11>("8 FADD2 + 8 FFMA2 + 8 FMNMX", 24);
  return 0;
}
