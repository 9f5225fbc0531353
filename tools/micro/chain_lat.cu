// Microbenchmark (development): dependent-chain latencies of the warp-collective / special instructions the MDS replay warp's
// per-pick chain is made of, one warp alone on an SM.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/chain_lat chain_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
template <int MODE>
__global__ void k(float* out, long long* cyc, float seed) {
  const int lane = threadIdx.x & 31;
  __shared__ float4 sm[64];
  unsigned v = (unsigned)(lane * 2654435761u) ^ __float_as_uint(seed);
  float f = seed + lane * 1e-3f;
  sm[lane] = make_float4(f, f, f, f);
  __syncwarp();
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < ITERS; i++) {
    if (MODE == 0) { v = __reduce_min_sync(0xffffffffu, v) + lane + 1; }                       // CREDUX.MIN + mov + add
    if (MODE == 1) { unsigned b = __ballot_sync(0xffffffffu, (v & 1) == 0); v = v + b; }       // VOTE + add
    if (MODE == 2) { unsigned b = __ballot_sync(0xffffffffu, (v & 31) == lane); int ol = __ffs(b | 0x80000000u) - 1; v = v + ol + 1; }  // VOTE+FLO
    if (MODE == 3) { v = __shfl_sync(0xffffffffu, v, (v + 1) & 31) + 1; }                      // SHFL.IDX (lane from data)
    if (MODE == 4) { f = exp2f(f) * 0.25f; }                                                  // MUFU.EX2 (+FMUL)  [exp2f accurate = ex2.approx + scaling]
    if (MODE == 5) { f = __fmaf_rn(f, 0.999f, 0.001f); }                                      // FFMA
    if (MODE == 6) { f = expf(-f) + 0.5f; }                                                   // full expf
    if (MODE == 7) { f = sm[(__float_as_uint(f) >> 3) & 31].x + 0.001f; }                     // LDS dependent
    if (MODE == 8) { if (lane == 0) *(volatile float*)&sm[32].x = f; __syncwarp(); f = *(volatile float*)&sm[32].x + 1.0f; }  // STS -> LDS round trip
    if (MODE == 9) {  // the replay chain's collective part: redux -> vote -> flo -> 4 shfl
      unsigned mh = __reduce_min_sync(0xffffffffu, v);
      unsigned b = __ballot_sync(0xffffffffu, v == mh);
      int ol = __ffs(b) - 1;
      float a = __shfl_sync(0xffffffffu, f, ol);
      unsigned w = __shfl_sync(0xffffffffu, v, ol);
      f = a + 1.0f; v = (w * 1664525u + 1013904223u) ^ (lane * 40503u);
    }
    if (MODE == 10) { v = __reduce_min_sync(0xffffffffu, v) + lane + 1; v = __reduce_min_sync(0xffffffffu, v ^ 5u) + lane; }  // 2 dependent redux
    if (MODE == 11) { v = __shfl_xor_sync(0xffffffffu, v, 1) + 1; }                            // SHFL.BFLY
    if (MODE == 12) { unsigned m = __match_any_sync(0xffffffffu, v & 3); v = v + m; }          // MATCH
    if (MODE == 13) { v = min(v, __shfl_xor_sync(0xffffffffu, v, 16)); v = min(v, __shfl_xor_sync(0xffffffffu, v, 8)); v = min(v, __shfl_xor_sync(0xffffffffu, v, 4));
                      v = min(v, __shfl_xor_sync(0xffffffffu, v, 2)); v = min(v, __shfl_xor_sync(0xffffffffu, v, 1)); v += lane + 1; }  // 5-level butterfly min
  }
  long long t1 = clock64();
  if (lane == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = f + v;
}
template <int MODE>
void run(const char* name) {
  float* out; long long* cyc;
  cudaMalloc(&out, 4 * 32); cudaMalloc(&cyc, 8);
  k<MODE><<<1, 32>>>(out, cyc, 1.0f);
  k<MODE><<<1, 32>>>(out, cyc, 1.0f);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-44s %7.1f cycles / iteration\n", name, (double)h / ITERS);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("CREDUX.MIN -> vector add");
  run<10>("2 dependent CREDUX.MIN");
  run<1>("VOTE.ANY (ballot) -> add");
  run<2>("ballot -> ffs -> add");
  run<3>("SHFL.IDX -> add");
  run<11>("SHFL.BFLY -> add");
  run<13>("5-level butterfly min");
  run<12>("MATCH.ANY");
  run<4>("exp2f (MUFU.EX2 + scale)");
  run<5>("FFMA");
  run<6>("expf(-x) + add");
  run<7>("LDS (dependent address) + add");
  run<8>("STS -> syncwarp -> LDS + add");
  run<9>("redux -> ballot -> ffs -> 2 shfl -> add");
  return 0;
}
