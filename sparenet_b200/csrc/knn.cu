// knn.cu -- k nearest neighbours in feature space for the EdgeConv encoder, sm_100a.
//
// Replaces the un-vendored third-party knn_cuda.KNN(k, transpose_mode=True) call of
// models/sparenet_generator.py:852-877 (KNN_CUDA 0.2 wheel, setup_env.sh:5; brute force, fp32).  Contract
// (SURVEY.md 9.6): for every point the k points with the smallest squared feature distance, itself
// included; downstream only consumes the SET (max over k, BN/SE statistics are permutation invariant).
// Defined arithmetic: d(i,j) = sum_c (x[c,j]-x[c,i])^2 accumulated with FMAs in ascending c, fp32;
// ordering (d, j) lexicographic, i.e. ties go to the smaller index.
//
// The matrix is symmetric bit for bit, so only upper-triangular tiles are evaluated and mirrored.
// Two kernels: (1) a register-tiled 128x128 SIMT distance kernel on the channel-major [B,C,N] layout the
// encoder already uses (coalesced loads, 8x8 outputs per thread, direct-difference form -- the
// |a|^2+|b|^2-2ab GEMM form would lose the exact ordering to cancellation); (2) a warp-per-row top-k.
//
// The encoder's FIRST layer searches in xyz (C = 3): knn_small_kernel does distance and selection in one launch with the
// sample's points in shared memory and nothing in HBM but the indices (the two-kernel path wrote and re-read a [B,N,N] matrix:
// 0.52 ms per step for 134 M three-term distances).  A warp per query: every lane keeps its 64 distances in registers, the
// k-th smallest of the 32 lane minima is an upper bound of the k-th smallest distance (k distinct points are at or below
// it), the few entries at or below that bound go to a shared-memory list and k pops of the warp minimum by (d, j) order
// them -- no per-lane sorted lists (whose insertion path diverges on almost every element when a lane only sees 64).
// Same arithmetic (d = x_j - x_i, fma in ascending c from 0) and the same (d, j) order as the two-kernel path: identical
// indices (tests/test_gpu_ops.py::test_knn_small_identical_to_two_kernel_path).
#include <stdlib.h>
#include "common.cuh"

namespace snb {

constexpr int KNN_BT = 128;  // block tile (rows = cols)
constexpr int KNN_KT = 16;   // channels per smem stage
constexpr int KNN_MAXK = 32;

__global__ void __launch_bounds__(256, 2) knn_dist_kernel(const float* __restrict__ x, int C, int N, float* __restrict__ D) {
  __shared__ __align__(16) float sa[KNN_KT][KNN_BT];
  __shared__ __align__(16) float sb[KNN_KT][KNN_BT];
  const int b = blockIdx.z;
  if (blockIdx.x < blockIdx.y) return;  // d(i,j) == d(j,i) bit for bit ((-t)^2 == t^2): only upper-triangular tiles compute
  const int i0 = blockIdx.y * KNN_BT, j0 = blockIdx.x * KNN_BT;
  const float* __restrict__ xb = x + (size_t)b * C * N;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 8 x 8 outputs each
  float acc[8][8];
#pragma unroll
  for (int r = 0; r < 8; r++)
#pragma unroll
    for (int c = 0; c < 8; c++) acc[r][c] = 0.f;

  // software pipeline: the next stage's global loads are in flight (registers) while the current stage is computed from smem
  constexpr int PER = KNN_KT * KNN_BT / 256;  // elements of each operand a thread stages
  float ra[PER], rb[PER];
  auto fetch = [&](int c0) {
#pragma unroll
    for (int u = 0; u < PER; u++) {
      const int e = threadIdx.x + u * 256;
      const int cc = e / KNN_BT, p = e % KNN_BT;
      const int c = c0 + cc;
      const int ia = i0 + p, jb = j0 + p;
      ra[u] = (c < C && ia < N) ? xb[(size_t)c * N + ia] : 0.f;
      rb[u] = (c < C && jb < N) ? xb[(size_t)c * N + jb] : 0.f;
    }
  };
  fetch(0);
  for (int c0 = 0; c0 < C; c0 += KNN_KT) {
#pragma unroll
    for (int u = 0; u < PER; u++) {
      const int e = threadIdx.x + u * 256;
      sa[e / KNN_BT][e % KNN_BT] = ra[u];
      sb[e / KNN_BT][e % KNN_BT] = rb[u];
    }
    __syncthreads();
    if (c0 + KNN_KT < C) fetch(c0 + KNN_KT);
#pragma unroll
    for (int cc = 0; cc < KNN_KT; cc++) {
      const float4 a0 = *reinterpret_cast<const float4*>(&sa[cc][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&sa[cc][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&sb[cc][tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&sb[cc][tx * 8 + 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c < 8; c++) {
          const float d = __fsub_rn(bv[c], av[r]);
          acc[r][c] = __fmaf_rn(d, d, acc[r][c]);
        }
    }
    __syncthreads();
  }
  float* __restrict__ Db = D + (size_t)b * N * N;
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const int i = i0 + ty * 8 + r;
    if (i >= N) continue;
    const int j = j0 + tx * 8;
    if (j + 7 < N && ((N & 3) == 0)) {
      *reinterpret_cast<float4*>(&Db[(size_t)i * N + j]) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
      *reinterpret_cast<float4*>(&Db[(size_t)i * N + j + 4]) = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
    } else {
#pragma unroll
      for (int c = 0; c < 8; c++)
        if (j + c < N) Db[(size_t)i * N + j + c] = acc[r][c];
    }
  }
  if (blockIdx.x == blockIdx.y) return;
  // mirror tile: D[j0+.., i0+..] = acc^T; for a fixed column c the 8 rows r are contiguous in memory
#pragma unroll
  for (int c = 0; c < 8; c++) {
    const int j = j0 + tx * 8 + c;
    if (j >= N) continue;
    const int i = i0 + ty * 8;
    if (i + 7 < N && ((N & 3) == 0)) {
      *reinterpret_cast<float4*>(&Db[(size_t)j * N + i]) = make_float4(acc[0][c], acc[1][c], acc[2][c], acc[3][c]);
      *reinterpret_cast<float4*>(&Db[(size_t)j * N + i + 4]) = make_float4(acc[4][c], acc[5][c], acc[6][c], acc[7][c]);
    } else {
#pragma unroll
      for (int r = 0; r < 8; r++)
        if (i + r < N) Db[(size_t)j * N + i + r] = acc[r][c];
    }
  }
}

// warp per query row: per-lane sorted top-K over a strided scan, then K rounds of warp arg-min to merge.
template <int K>
__global__ void __launch_bounds__(256) knn_topk_kernel(const float* __restrict__ D, int N, size_t rows, int k, int* __restrict__ idx) {
  const size_t row = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* __restrict__ d = D + row * N;
  float bd[K];
  int bi[K];
#pragma unroll
  for (int t = 0; t < K; t++) {
    bd[t] = __int_as_float(0x7f800000);
    bi[t] = 0x7fffffff;
  }
  auto offer = [&](float v, int vi) {
    if (v < bd[K - 1]) {  // within a lane j ascends, so strict '<' keeps the smaller index on ties
      bool ins = false;
#pragma unroll
      for (int t = 0; t < K; t++) {
        if (ins || v < bd[t]) {  // insert, then shift every later entry down by one
          ins = true;
          const float tv = bd[t];
          const int ti = bi[t];
          bd[t] = v;
          bi[t] = vi;
          v = tv;
          vi = ti;
        }
      }
    }
  };
  if ((N & 3) == 0) {  // 128-bit loads, two in flight per lane; j still ascends within a lane
    const float4* __restrict__ d4 = reinterpret_cast<const float4*>(d);
    const int n4 = N >> 2;
    for (int q = lane; q < n4; q += 64) {
      const float4 v0 = d4[q];
      const bool two = q + 32 < n4;
      const float4 v1 = two ? d4[q + 32] : make_float4(0.f, 0.f, 0.f, 0.f);
      offer(v0.x, 4 * q);
      offer(v0.y, 4 * q + 1);
      offer(v0.z, 4 * q + 2);
      offer(v0.w, 4 * q + 3);
      if (two) {
        offer(v1.x, 4 * (q + 32));
        offer(v1.y, 4 * (q + 32) + 1);
        offer(v1.z, 4 * (q + 32) + 2);
        offer(v1.w, 4 * (q + 32) + 3);
      }
    }
  } else {
    for (int j = lane; j < N; j += 32) offer(d[j], j);
  }
  int* out = idx + row * k;
  for (int t = 0; t < k; t++) {
    float hv = bd[0];
    int hi = bi[0];
    float mv = hv;
    int mi = hi;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, mv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
      if (ov < mv || (ov == mv && oi < mi)) {
        mv = ov;
        mi = oi;
      }
    }
    if (lane == 0) out[t] = mi;
    if (hi == mi && hv == mv) {  // the owning lane pops its head
#pragma unroll
      for (int u = 0; u + 1 < K; u++) {
        bd[u] = bd[u + 1];
        bi[u] = bi[u + 1];
      }
      bd[K - 1] = __int_as_float(0x7f800000);
      bi[K - 1] = 0x7fffffff;
    }
  }
}

// ---- fused distance + selection for narrow features (C <= 4: the xyz layer) --------------------------------------------------------
constexpr int KS_MAXN = 2048;  // 64 distances per lane
constexpr int KS_CAP = 96;     // entries at or below the bound (3 per lane); more (many identical points) -> pops over the registers
constexpr int KS_QPW = 8;      // queries per warp (the CTA stages its sample's points once for 64 queries)

template <int C>
__global__ void __launch_bounds__(256) knn_small_kernel(const float* __restrict__ x, int N, int k, int* __restrict__ idx) {
  extern __shared__ __align__(16) float ks_x[];  // [C][N]
  __shared__ float cd[8][KS_CAP];
  __shared__ int cj[8][KS_CAP];
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* __restrict__ xb = x + (size_t)b * C * N;
  for (int e = threadIdx.x; e < C * N; e += 256) ks_x[e] = xb[e];
  __syncthreads();
  const float INF = __int_as_float(0x7f800000);
  const int q0 = (blockIdx.x * 8 + warp) * KS_QPW;
  for (int qq = 0; qq < KS_QPW; qq++) {
    const int i = q0 + qq;
    if (i >= N) break;  // warp-uniform
    float xi[C];
#pragma unroll
    for (int c = 0; c < C; c++) xi[c] = ks_x[c * N + i];
    float v[64];
    float lm = INF;
#pragma unroll
    for (int t = 0; t < 64; t++) {
      const int j = t * 32 + lane;
      float acc = INF;
      if (j < N) {
        acc = 0.f;
#pragma unroll
        for (int c = 0; c < C; c++) {
          const float d = __fsub_rn(ks_x[c * N + j], xi[c]);
          acc = __fmaf_rn(d, d, acc);
        }
      }
      v[t] = acc;
      lm = fminf(lm, acc);
    }
    // bound: the k-th smallest lane minimum (distances are >= +0: the bit patterns order like the values)
    unsigned key = __float_as_uint(lm), bound = 0x7f800000u;
    for (int t = 0; t < k; t++) {
      bound = __reduce_min_sync(0xffffffffu, key);
      const unsigned owners = __ballot_sync(0xffffffffu, key == bound);
      if (lane == __ffs(owners) - 1) key = 0x7f800000u;
    }
    const float tau = __uint_as_float(bound);
    // list positions from ballots (warp-uniform branch, taken for the ~10 of 64 steps that have a hit): a shared-memory atomic per
    // hit put its round trip on the warp's critical path at every reconvergence point
    int n = 0;
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int t = 0; t < 64; t++) {
      const bool hit = v[t] <= tau && t * 32 + lane < N;
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (m) {
        const int pos = n + __popc(m & lt);
        if (hit && pos < KS_CAP) {
          cd[warp][pos] = v[t];
          cj[warp][pos] = t * 32 + lane;
        }
        n += __popc(m);
      }
    }
    __syncwarp();
    int* out = idx + ((size_t)b * N + i) * k;
    if (n <= KS_CAP) {
      float ed[3];
      int ej[3];
#pragma unroll
      for (int u = 0; u < 3; u++) {
        const int e = lane + 32 * u;
        ed[u] = e < n ? cd[warp][e] : INF;
        ej[u] = e < n ? cj[warp][e] : 0x7fffffff;
      }
      for (int t = 0; t < k; t++) {
        float hv = ed[0];
        int hj = ej[0];
#pragma unroll
        for (int u = 1; u < 3; u++)
          if (ed[u] < hv || (ed[u] == hv && ej[u] < hj)) {
            hv = ed[u];
            hj = ej[u];
          }
        const unsigned md = __reduce_min_sync(0xffffffffu, __float_as_uint(hv));
        const int mj = (int)__reduce_min_sync(0xffffffffu, __float_as_uint(hv) == md ? (unsigned)hj : 0x7fffffffu);
        if (lane == 0) out[t] = mj;
#pragma unroll
        for (int u = 0; u < 3; u++)
          if (ej[u] == mj) {  // indices are unique: one slot of one lane
            ed[u] = INF;
            ej[u] = 0x7fffffff;
          }
      }
    } else {
      // many entries at the bound (duplicated points): k pops straight from the registers
      for (int t = 0; t < k; t++) {
        float hv = INF;
        int hj = 0x7fffffff;
#pragma unroll
        for (int u = 0; u < 64; u++)
          if (v[u] < hv) {  // ascending j within a lane: strict '<' keeps the smaller index
            hv = v[u];
            hj = u * 32 + lane;
          }
        const unsigned md = __reduce_min_sync(0xffffffffu, __float_as_uint(hv));
        const int mj = (int)__reduce_min_sync(0xffffffffu, __float_as_uint(hv) == md ? (unsigned)hj : 0x7fffffffu);
        if (lane == 0) out[t] = mj;
#pragma unroll
        for (int u = 0; u < 64; u++)
          if (u * 32 + lane == mj) v[u] = INF;
      }
    }
    __syncwarp();  // the lists are reused by the next query
  }
}

}  // namespace snb

using namespace snb;

SNB_API size_t snb_knn_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  return sizeof(float) * (size_t)B * N * N;
}

SNB_API int snb_knn(const float* x, int B, int C, int N, int k, int* idx, void* workspace, size_t workspace_bytes, void* stream) {
  if (B < 0 || C <= 0 || N < 0 || k <= 0) return SNB_EINVAL;
  if (k > KNN_MAXK || k > N || B > 65535) return SNB_ELIMIT;
  if (B == 0 || N == 0) return SNB_OK;
  if (!workspace || workspace_bytes < snb_knn_workspace_bytes(B, N)) return SNB_EWORKSPACE;
  if (((uintptr_t)workspace & 15) != 0) return SNB_EALIGN;
  cudaStream_t s = (cudaStream_t)stream;
  // narrow features (the xyz layer): fused distance + selection, nothing but the indices reaches HBM (SNB_KNN_SMALL=0: two-kernel path)
  if (C <= 4 && N <= KS_MAXN && (size_t)C * N * sizeof(float) <= 40 * 1024) {
    const char* sw = getenv("SNB_KNN_SMALL");
    if (!(sw && sw[0] == '0')) {
      const dim3 grid((unsigned)((N + 8 * KS_QPW - 1) / (8 * KS_QPW)), (unsigned)B);
      const size_t sm = (size_t)C * N * sizeof(float);
      switch (C) {
        case 1: knn_small_kernel<1><<<grid, 256, sm, s>>>(x, N, k, idx); break;
        case 2: knn_small_kernel<2><<<grid, 256, sm, s>>>(x, N, k, idx); break;
        case 3: knn_small_kernel<3><<<grid, 256, sm, s>>>(x, N, k, idx); break;
        default: knn_small_kernel<4><<<grid, 256, sm, s>>>(x, N, k, idx); break;
      }
      SNB_LAUNCH_CHECK();
      return SNB_OK;
    }
  }
  float* D = (float*)workspace;
  const int nt = (N + KNN_BT - 1) / KNN_BT;
  knn_dist_kernel<<<dim3(nt, nt, B), 256, 0, s>>>(x, C, N, D);
  SNB_LAUNCH_CHECK();
  const size_t rows = (size_t)B * N;
  const unsigned g = (unsigned)((rows * 32 + 255) / 256);
  if (k <= 8) knn_topk_kernel<8><<<g, 256, 0, s>>>(D, N, rows, k, idx);
  else if (k <= 16) knn_topk_kernel<16><<<g, 256, 0, s>>>(D, N, rows, k, idx);
  else knn_topk_kernel<32><<<g, 256, 0, s>>>(D, N, rows, k, idx);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
