// thinconv.cu -- 1x1 convolutions with a THIN side (<= 8 channels in or out), exact fp32, sm_100a.
//
// The generator's first and last layers: xyz / lattice inputs (EdgeConv conv1 on 3 channels, the folding decoders' Conv1d(2 -> 1026),
// PointNetRes conv1 on 4) and xyz outputs (decoder conv4 256 -> 3, PointNetRes conv7 128 -> 3) -- reference
// models/sparenet_generator.py:146-160, 984-991, 1044-1062, 593-646.  Their rows are shorter than a TMA box of the tensor-core GEMM and
// their arithmetic is nothing (< 0.5 % of the step's flops): each is ONE pass over its wide tensor, so the kernels below are plain
// streaming loops with 128-bit accesses (the library GEMMs they replace ran at 50-75 % of that).  The reference computes these layers
// with cuDNN (TF32 allowed); fp32 FMAs here.
//   expand   y[g,co,n] = sum_{ci<S} W[g?][co,ci] x[g?][ci,n]     (S <= 8 input channels; also the data gradient of a thin-OUTPUT layer)
//   reduce   y[g,s,n]  = sum_{c<L} W[g?][s,c]   x[g][c,n]        (S <= 8 output channels; also the data gradient of a thin-INPUT layer)
//   wgrad    G[g][l,s] = sum_n big[g][l,n] small[g?][s,n]         (per batch entry; a shared weight's sum over g is a second tiny pass)
// Strides: every operand has its own batch stride (0 = shared by the batch); weights are addressed as W[row * w_rs + col * w_cs], so the
// transposed use in a data gradient needs no copy.
#include "common.cuh"

namespace snb {

constexpr int THIN_MAXS = 8;

// ---- expand: grid (ceil(N/4/256), ceil(Co/32), G); thread = 4 consecutive positions, loops over its block's 32 output channels --------
__global__ void __launch_bounds__(256) thin_expand_kernel(const float* __restrict__ x, long long x_bs, const float* __restrict__ W, long long w_bs,
                                                           int w_rs, int w_cs, int S, int Co, int N, float* __restrict__ y) {
  __shared__ float ws[32][THIN_MAXS];
  const int g = blockIdx.z, co0 = blockIdx.y * 32;
  const float* __restrict__ wg = W + (size_t)g * w_bs;
  for (int e = threadIdx.x; e < 32 * THIN_MAXS; e += 256) {
    const int r = e / THIN_MAXS, s = e - r * THIN_MAXS;
    ws[r][s] = (co0 + r < Co && s < S) ? wg[(size_t)(co0 + r) * w_rs + (size_t)s * w_cs] : 0.f;
  }
  __syncthreads();
  const int n = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (n >= N) return;
  const float* __restrict__ xg = x + (size_t)g * x_bs;
  float4 xv[THIN_MAXS];
#pragma unroll
  for (int s = 0; s < THIN_MAXS; s++) xv[s] = s < S ? *reinterpret_cast<const float4*>(xg + (size_t)s * N + n) : make_float4(0.f, 0.f, 0.f, 0.f);
  float* __restrict__ yg = y + ((size_t)g * Co + co0) * N + n;
  const int nr = min(32, Co - co0);
  for (int r = 0; r < nr; r++) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int s = 0; s < THIN_MAXS; s++) {
      const float w = ws[r][s];
      a.x = __fmaf_rn(w, xv[s].x, a.x);
      a.y = __fmaf_rn(w, xv[s].y, a.y);
      a.z = __fmaf_rn(w, xv[s].z, a.z);
      a.w = __fmaf_rn(w, xv[s].w, a.w);
    }
    *reinterpret_cast<float4*>(yg + (size_t)r * N) = a;
  }
}

// ---- reduce: grid (ceil(N/4/256), G); thread = 4 consecutive positions, loops over all L wide channels ------------------------------
__global__ void __launch_bounds__(256) thin_reduce_kernel(const float* __restrict__ x, long long x_bs, const float* __restrict__ W, long long w_bs,
                                                           int w_rs, int w_cs, int S, int L, int N, float* __restrict__ y, long long y_bs) {
  extern __shared__ float wsm[];   // [L][THIN_MAXS]
  const int g = blockIdx.y;
  const float* __restrict__ wg = W + (size_t)g * w_bs;
  for (int e = threadIdx.x; e < L * THIN_MAXS; e += 256) {
    const int c = e / THIN_MAXS, s = e - c * THIN_MAXS;
    wsm[e] = s < S ? wg[(size_t)s * w_rs + (size_t)c * w_cs] : 0.f;
  }
  __syncthreads();
  const int n = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (n >= N) return;
  const float* __restrict__ xg = x + (size_t)g * x_bs + n;
  float4 acc[THIN_MAXS];
#pragma unroll
  for (int s = 0; s < THIN_MAXS; s++) acc[s] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
  for (int c = 0; c < L; c++) {
    const float4 v = *reinterpret_cast<const float4*>(xg + (size_t)c * N);
    const float4 w0 = *reinterpret_cast<const float4*>(&wsm[c * THIN_MAXS]), w1 = *reinterpret_cast<const float4*>(&wsm[c * THIN_MAXS + 4]);
    const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int s = 0; s < THIN_MAXS; s++) {
      acc[s].x = __fmaf_rn(w[s], v.x, acc[s].x);
      acc[s].y = __fmaf_rn(w[s], v.y, acc[s].y);
      acc[s].z = __fmaf_rn(w[s], v.z, acc[s].z);
      acc[s].w = __fmaf_rn(w[s], v.w, acc[s].w);
    }
  }
  float* __restrict__ yg = y + (size_t)g * y_bs + n;
#pragma unroll
  for (int s = 0; s < THIN_MAXS; s++)
    if (s < S) *reinterpret_cast<float4*>(yg + (size_t)s * N) = acc[s];
}

// ---- wgrad: grid (ceil(L/4), G); a block owns 4 wide rows of one batch entry, its 256 threads stride over the positions.  The thin
//      channel count is a template parameter (no padded FMAs, ~64 registers: three blocks per SM keep enough loads in flight -- the first
//      version, 8 rows x 8 padded channels at 200 registers and one block per SM, ran at a quarter of the streaming rate) ----------------
template <int ST>
__global__ void __launch_bounds__(256) thin_wgrad_kernel(const float* __restrict__ big, long long big_bs, const float* __restrict__ small,
                                                          long long small_bs, int S, int L, int N, float* __restrict__ out) {
  // out[g][l][s] (row-major [G, L, THIN_MAXS]; columns >= S are zero)
  __shared__ float red[8][4 * THIN_MAXS];
  const int g = blockIdx.y, l0 = blockIdx.x * 4;
  const float* __restrict__ bg = big + (size_t)g * big_bs;
  const float* __restrict__ sg = small + (size_t)g * small_bs;
  const float* __restrict__ br[4];
#pragma unroll
  for (int r = 0; r < 4; r++) br[r] = bg + (size_t)min(l0 + r, L - 1) * N;   // rows past L repeat the last one (never stored)
  float acc[4][ST];
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int s = 0; s < ST; s++) acc[r][s] = 0.f;
#pragma unroll 2
  for (int n = threadIdx.x * 4; n < N; n += 1024) {
    float4 sv[ST], bv[4];
#pragma unroll
    for (int r = 0; r < 4; r++) bv[r] = *reinterpret_cast<const float4*>(br[r] + n);
#pragma unroll
    for (int s = 0; s < ST; s++) sv[s] = s < S ? *reinterpret_cast<const float4*>(sg + (size_t)s * N + n) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int s = 0; s < ST; s++)
        acc[r][s] = __fmaf_rn(bv[r].x, sv[s].x, __fmaf_rn(bv[r].y, sv[s].y, __fmaf_rn(bv[r].z, sv[s].z, __fmaf_rn(bv[r].w, sv[s].w, acc[r][s]))));
  }
  // block reduction: warp shuffles, then the 8 warps through shared memory (fixed order)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int s = 0; s < ST; s++) {
      float v = acc[r][s];
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[warp][r * THIN_MAXS + s] = v;
    }
  __syncthreads();
  if (threadIdx.x < 4 * THIN_MAXS) {
    const int r = threadIdx.x / THIN_MAXS, s = threadIdx.x - r * THIN_MAXS;
    float v = 0.f;
    if (s < ST) {
#pragma unroll
      for (int w = 0; w < 8; w++) v += red[w][threadIdx.x];
    }
    if (l0 + r < L) out[((size_t)g * L + l0 + r) * THIN_MAXS + s] = v;
  }
}

}  // namespace snb

using namespace snb;

static int thin_check(int G, int S, int L, int N) {
  if (G < 0 || S <= 0 || L <= 0 || N < 0) return SNB_EINVAL;
  if (S > THIN_MAXS || (N & 3) != 0 || G > 65535) return SNB_ELIMIT;
  return SNB_OK;
}

// y [G,Co,N] = W x: x [.,S,N] with batch stride x_bs (0: shared), W element (co, s) at W[g * w_bs + co * w_rs + s * w_cs]
SNB_API int snb_thin_expand(const float* x, long long x_bs, const float* W, long long w_bs, int w_rs, int w_cs, int G, int S, int Co, int N, float* y,
                            void* stream) {
  int rc = thin_check(G, S, Co, N);
  if (rc) return rc;
  if (G == 0 || N == 0) return SNB_OK;
  if ((((uintptr_t)x | (uintptr_t)y) & 15) != 0 || (x_bs & 3) != 0) return SNB_EALIGN;
  thin_expand_kernel<<<dim3((unsigned)((N / 4 + 255) / 256), (unsigned)((Co + 31) / 32), (unsigned)G), 256, 0, (cudaStream_t)stream>>>(x, x_bs, W, w_bs,
                                                                                                                              w_rs, w_cs, S, Co, N, y);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

// y [.,S,N] (batch stride y_bs) = W x: x [G,L,N] with batch stride x_bs, W element (s, c) at W[g * w_bs + s * w_rs + c * w_cs]
SNB_API int snb_thin_reduce(const float* x, long long x_bs, const float* W, long long w_bs, int w_rs, int w_cs, int G, int S, int L, int N, float* y,
                            long long y_bs, void* stream) {
  int rc = thin_check(G, S, L, N);
  if (rc) return rc;
  if (G == 0 || N == 0) return SNB_OK;
  if ((((uintptr_t)x | (uintptr_t)y) & 15) != 0 || (x_bs & 3) != 0 || (y_bs & 3) != 0) return SNB_EALIGN;
  const size_t smem = (size_t)L * THIN_MAXS * sizeof(float);
  if (smem > 200 * 1024) return SNB_ELIMIT;
  if (smem > 48 * 1024) SNB_CUDA(cudaFuncSetAttribute(thin_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  thin_reduce_kernel<<<dim3((unsigned)((N / 4 + 255) / 256), (unsigned)G), 256, smem, (cudaStream_t)stream>>>(x, x_bs, W, w_bs, w_rs, w_cs, S, L, N, y,
                                                                                                          y_bs);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

// out [G, L, 8] (columns >= S zero): out[g][l][s] = sum_n big[g][l,n] small[g][s,n]
SNB_API int snb_thin_wgrad(const float* big, long long big_bs, const float* small_, long long small_bs, int G, int S, int L, int N, float* out,
                           void* stream) {
  int rc = thin_check(G, S, L, N);
  if (rc) return rc;
  if (G == 0) return SNB_OK;
  if ((((uintptr_t)big | (uintptr_t)small_) & 15) != 0 || (big_bs & 3) != 0 || (small_bs & 3) != 0) return SNB_EALIGN;
  const dim3 grid((unsigned)((L + 3) / 4), (unsigned)G);
  cudaStream_t st = (cudaStream_t)stream;
  if (S <= 2) thin_wgrad_kernel<2><<<grid, 256, 0, st>>>(big, big_bs, small_, small_bs, S, L, N, out);
  else if (S == 3) thin_wgrad_kernel<3><<<grid, 256, 0, st>>>(big, big_bs, small_, small_bs, S, L, N, out);
  else if (S == 4) thin_wgrad_kernel<4><<<grid, 256, 0, st>>>(big, big_bs, small_, small_bs, S, L, N, out);
  else thin_wgrad_kernel<8><<<grid, 256, 0, st>>>(big, big_bs, small_, small_bs, S, L, N, out);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
