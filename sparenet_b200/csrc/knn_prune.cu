// knn_prune.cu -- kNN for wide features (C >= 64) with a tensor-core Gram matrix as a PRUNING filter, sm_100a.
//
// Parity: tests/test_gpu_ops.py::test_knn_pruned_identical_to_brute_force (identical indices incl. the large-norm cancellation case
// and duplicated / all-zero points); the rule was also checked on the CPU on the encoder's real features
// (tests/perf/knn_prune_study.py: ~11 candidates per row survive for k = 8).  Timed on B200 inside the step (B=32, N=2048, k=8, CUPTI):
// per layer 0.43 ms (C=256) / 0.59 ms (C=512) for the selection + exact re-ranking below, plus 0.16 / 0.25 ms of Gram GEMM, 0.05 ms of
// norms and 0.03 / 0.05 ms for the point-major copy (the first version of this file: 0.68 / 0.82 ms for the kernel alone; the
// brute-force kernels: 1.45 ms per 256 channels); sparenet_b200.functional.knn_indices uses it for C >= 256 (SNB_KNN_PRUNE=1 / 0
// forces it on / off), the Gram matrix comes from snb_gemm_tf32.
//
// Same contract and the SAME BITS as snb_knn (knn.cu): for every point the k points with the smallest
//     d(i,j) = sum_c (x[c,j] - x[c,i])^2,  accumulated with FMAs in ascending c, fp32,
// ordered by (d, j).  SURVEY.md 8(d): the |a|^2 + |b|^2 - 2ab GEMM form may not DEFINE the result (cancellation), but it may
// PRUNE: with G~ = X^T X from the TF32 tensor-core GEMM (operands truncated to 10 mantissa bits, fp32 accumulation)
//     d~(i,j) = n_i + n_j - 2 G~(i,j),   |d~ - d| <= eps_i = 2^-7.5 |a_i| max_j |a_j| + 2^-13 (n_i + max_j n_j)
// (2 * 2^-10 relative per product from the truncation, Cauchy-Schwarz, the factor 2 of the formula; the second term covers the
// fp32 rounding of the norms and of the accumulations for C <= 1024).  Every j among the exact k nearest of i then satisfies
//     d~(i,j) <= kth_smallest_j d~(i,j) * (1 + 2^-10) + 2 eps_i,
// so the exact distances are evaluated only for those candidates (a handful per row) and the exact (d, j) order picks the k.
// Rows whose candidate list overflows fall back to evaluating every j exactly.
//
// Inputs: xT [B,N,C] point-major copy of the features (candidate rows are contiguous), gram [B,N,N] from the GEMM.
#include <stdlib.h>
#include "common.cuh"

namespace snb {

constexpr int KP_MAXK = 32;
constexpr int KP_CAND = 96;  // candidate capacity per row (3 per lane)

// point-major copy xT [B,N,C] of the channel-major features x [B,C,N]: 32 x 32 tiles through shared memory, 128-byte rows on both sides
// (the PyTorch permute-copy ran at ~1.7 TB/s: 0.3 ms per step for the three wide layers)
__global__ void __launch_bounds__(256) transpose_cn_kernel(const float* __restrict__ x, int C, int N, float* __restrict__ xT) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* __restrict__ xb = x + (size_t)b * C * N;
  float* __restrict__ tb = xT + (size_t)b * C * N;
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const int c = c0 + ty + 8 * r, n = n0 + tx;
    tile[ty + 8 * r][tx] = (c < C && n < N) ? xb[(size_t)c * N + n] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const int n = n0 + ty + 8 * r, c = c0 + tx;
    if (n < N && c < C) tb[(size_t)n * C + c] = tile[tx][ty + 8 * r];
  }
}

// squared norms of the points (fp32) and their per-sample maximum (bit pattern; norms are >= 0)
__global__ void __launch_bounds__(256) knn_norm_kernel(const float* __restrict__ xT, int C, int N, size_t rows, float* __restrict__ nrm,
                                                        unsigned* __restrict__ nmax_bits) {
  const size_t row = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* __restrict__ p = xT + row * C;
  float acc = 0.f;
  for (int c = lane; c < C; c += 32) acc = __fmaf_rn(p[c], p[c], acc);
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    nrm[row] = acc;
    atomicMax(&nmax_bits[row / N], __float_as_uint(acc));
  }
}

// exact distance of candidate j to query i, the arithmetic of knn_dist_kernel: d = x_j[c] - x_i[c]; acc = fma(d, d, acc), ascending c
__device__ __forceinline__ float knn_exact_dist(const float* __restrict__ xi, const float* __restrict__ xj, int C) {
  float acc = 0.f;
  int c = 0;
  if ((C & 3) == 0) {
    for (; c < C; c += 4) {
      const float4 a = *reinterpret_cast<const float4*>(xi + c);
      const float4 b = *reinterpret_cast<const float4*>(xj + c);
      float d = __fsub_rn(b.x, a.x);
      acc = __fmaf_rn(d, d, acc);
      d = __fsub_rn(b.y, a.y);
      acc = __fmaf_rn(d, d, acc);
      d = __fsub_rn(b.z, a.z);
      acc = __fmaf_rn(d, d, acc);
      d = __fsub_rn(b.w, a.w);
      acc = __fmaf_rn(d, d, acc);
    }
  } else {
    for (; c < C; c++) {
      const float d = __fsub_rn(xj[c], xi[c]);
      acc = __fmaf_rn(d, d, acc);
    }
  }
  return acc;
}

// warp per query row.  Selection without per-lane sorted lists (a lane only sees N/32 entries, so their insertion path ran -- for the
// whole warp -- on almost every element: 6 000 warp instructions per row, 0.68 ms per layer):
//   1. every lane computes its approximate distances (kept in registers when N <= 2048: one pass over the Gram row, 128-bit loads all
//      independent) and their minimum; the k-th smallest of the 32 lane minima, tau0, is an upper bound of the k-th smallest
//      approximate distance (k distinct entries are at or below it);
//   2. entries at or below the LOOSE threshold f(tau0), f(t) = t (1 + 2^-10) + 2 eps (monotone, so f(tau0) >= f(kth)), go to a
//      shared-memory list -- a superset of the candidates, ~a dozen entries;
//   3. k pops of the warp minimum over the list give the exact k-th smallest approximate distance (duplicates count separately, as
//      before), the list is filtered by f(kth): the SAME candidate set as the two-pass scan, hence the same indices.
template <int K, bool CACHE>
__global__ void __launch_bounds__(256, CACHE ? 2 : 4) knn_prune_kernel(const float* __restrict__ xT, const float* __restrict__ gram, const float* __restrict__ nrm,
                                                         const unsigned* __restrict__ nmax_bits, int C, int N, size_t rows, int k,
                                                         int* __restrict__ idx) {
  __shared__ int cand[8][KP_CAND];
  __shared__ float cval[8][KP_CAND];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t row = (size_t)blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const size_t b = row / N;
  const int i = (int)(row - b * N);
  const float* __restrict__ g = gram + row * N;
  const float* __restrict__ nb = nrm + b * N;
  const float ni = nb[i];
  const float INF = __int_as_float(0x7f800000);
  const bool cached = CACHE && (N & 3) == 0 && N <= 2048;  // 16 x float4 per lane
  const int n4 = N >> 2;
  float v[64];
  float lm = INF;
  if (cached) {
    const float4* __restrict__ g4 = reinterpret_cast<const float4*>(g);
    const float4* __restrict__ nb4 = reinterpret_cast<const float4*>(nb);
#pragma unroll
    for (int t = 0; t < 16; t++) {
      const int q = lane + 32 * t;
      float4 r = make_float4(INF, INF, INF, INF);
      if (q < n4) {
        const float4 gg = g4[q], nn = nb4[q];
        r.x = __fmaf_rn(-2.f, gg.x, ni + nn.x);
        r.y = __fmaf_rn(-2.f, gg.y, ni + nn.y);
        r.z = __fmaf_rn(-2.f, gg.z, ni + nn.z);
        r.w = __fmaf_rn(-2.f, gg.w, ni + nn.w);
      }
      v[4 * t] = r.x;
      v[4 * t + 1] = r.y;
      v[4 * t + 2] = r.z;
      v[4 * t + 3] = r.w;
      lm = fminf(lm, fminf(fminf(r.x, r.y), fminf(r.z, r.w)));
    }
  } else {
    for (int j = lane; j < N; j += 32) lm = fminf(lm, __fmaf_rn(-2.f, g[j], ni + nb[j]));
  }
  // ---- tau0: the k-th smallest lane minimum (approximate distances may be slightly negative: order-preserving keys) ----
  const unsigned KINF = float_key(INF);
  unsigned key = float_key(lm), kb = KINF;
  for (int t = 0; t < k; t++) {
    kb = __reduce_min_sync(0xffffffffu, key);
    const unsigned owners = __ballot_sync(0xffffffffu, key == kb);
    if (lane == __ffs(owners) - 1) key = KINF;
  }
  const float tau0 = key_float(kb);
  const float nmx = __uint_as_float(nmax_bits[b]);
  // 2^-7.5 |a_i| max|a_j| for the truncated products, 2^-13 (n_i + max n_j) for the fp32 rounding of the norms and the accumulations
  // (also what keeps the rule valid for an all-zero query point, whose first term vanishes)
  const float eps = __fmaf_rn(0.0055242717f * sqrtf(ni), sqrtf(nmx), 0.00012207031f * (ni + nmx));
  const float thr_loose = __fmaf_rn(fabsf(tau0), 0.0009765625f, tau0) + 2.f * eps;
  // ---- the list: entries at or below the loose threshold (positions from ballots: warp-uniform branch, no shared-memory atomics) ----
  int nlist = 0;
  const unsigned lt = (1u << lane) - 1u;
  if (cached) {
#pragma unroll
    for (int t = 0; t < 64; t++) {
      const int j = 4 * (lane + 32 * (t >> 2)) + (t & 3);
      const bool hit = v[t] <= thr_loose;  // padding entries are +inf: only reachable when the threshold is +inf, which overflows the list
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (m) {
        const int pos = nlist + __popc(m & lt);
        if (hit && pos < KP_CAND && j < N) {
          cand[warp][pos] = j;
          cval[warp][pos] = v[t];
        }
        nlist += __popc(m);
      }
    }
  } else {
    for (int j0 = 0; j0 < N; j0 += 32) {
      const int j = j0 + lane;
      const float vv = j < N ? __fmaf_rn(-2.f, g[j], ni + nb[j]) : INF;
      const bool hit = j < N && vv <= thr_loose;
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (m) {
        const int pos = nlist + __popc(m & lt);
        if (hit && pos < KP_CAND) {
          cand[warp][pos] = j;
          cval[warp][pos] = vv;
        }
        nlist += __popc(m);
      }
    }
  }
  __syncwarp();
  // ---- the exact k-th smallest approximate distance from the list, then the candidates proper ----
  float cd[3];
  int cj[3];
  bool isc[3];
  int ncand = KP_CAND + 1;  // list overflow (degenerate rows) -> every j evaluated exactly below
  if (nlist <= KP_CAND) {
    float ev[3];
    unsigned ek[3];
#pragma unroll
    for (int u = 0; u < 3; u++) {
      const int e = lane + 32 * u;
      ev[u] = e < nlist ? cval[warp][e] : INF;
      cj[u] = e < nlist ? cand[warp][e] : 0x7fffffff;
      ek[u] = float_key(ev[u]);
    }
    unsigned kk = KINF;
    for (int t = 0; t < k; t++) {
      const unsigned lb = min(ek[0], min(ek[1], ek[2]));
      kk = __reduce_min_sync(0xffffffffu, lb);
      const unsigned owners = __ballot_sync(0xffffffffu, lb == kk);
      if (lane == __ffs(owners) - 1) {  // ONE entry pops (equal approximate values count separately)
        if (ek[0] == kk) ek[0] = KINF;
        else if (ek[1] == kk) ek[1] = KINF;
        else ek[2] = KINF;
      }
    }
    const float kth = key_float(kk);
    const float thr = __fmaf_rn(fabsf(kth), 0.0009765625f, kth) + 2.f * eps;
    ncand = 0;
#pragma unroll
    for (int u = 0; u < 3; u++) {
      isc[u] = (lane + 32 * u) < nlist && ev[u] <= thr;
      ncand += __popc(__ballot_sync(0xffffffffu, isc[u]));
    }
  }
  // ---- exact distances of the candidates, then the k smallest by (d, j) ----
  const float* __restrict__ xb = xT + b * (size_t)N * C;
  const float* __restrict__ xi = xb + (size_t)i * C;
  int* out = idx + row * k;
  if (ncand <= KP_CAND) {
#pragma unroll
    for (int u = 0; u < 3; u++) {
      if (isc[u]) {
        cd[u] = knn_exact_dist(xi, xb + (size_t)cj[u] * C, C);
      } else {
        cd[u] = INF;
        cj[u] = 0x7fffffff;
      }
    }
    for (int t = 0; t < k; t++) {
      // this lane's best remaining candidate, then the warp's by (d, j): exact distances are >= +0, their bits order like the values
      float hv = cd[0];
      int hj = cj[0];
#pragma unroll
      for (int u = 1; u < 3; u++)
        if (cd[u] < hv || (cd[u] == hv && cj[u] < hj)) {
          hv = cd[u];
          hj = cj[u];
        }
      const unsigned md = __reduce_min_sync(0xffffffffu, __float_as_uint(hv));
      const int mj = (int)__reduce_min_sync(0xffffffffu, __float_as_uint(hv) == md ? (unsigned)hj : 0x7fffffffu);
      if (lane == 0) out[t] = mj;
#pragma unroll
      for (int u = 0; u < 3; u++)
        if (cj[u] == mj) {  // indices are unique: exactly one slot of one lane
          cd[u] = __int_as_float(0x7f800000);
          cj[u] = 0x7fffffff;
        }
    }
  } else {
    // overflow (degenerate rows, e.g. many identical points): every j evaluated exactly, per-lane sorted lists as in knn_topk_kernel
    float ed[K];
    int ei[K];
#pragma unroll
    for (int t = 0; t < K; t++) {
      ed[t] = __int_as_float(0x7f800000);
      ei[t] = 0x7fffffff;
    }
    for (int j = lane; j < N; j += 32) {
      float v = knn_exact_dist(xi, xb + (size_t)j * C, C);
      int vi = j;
      if (v < ed[K - 1]) {
        bool ins = false;
#pragma unroll
        for (int t = 0; t < K; t++) {
          if (ins || v < ed[t]) {
            ins = true;
            const float tv = ed[t];
            const int ti = ei[t];
            ed[t] = v;
            ei[t] = vi;
            v = tv;
            vi = ti;
          }
        }
      }
    }
    for (int t = 0; t < k; t++) {
      const float hv = ed[0];
      const int hi = ei[0];
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
      if (hi == mi && hv == mv) {
#pragma unroll
        for (int u = 0; u + 1 < K; u++) {
          ed[u] = ed[u + 1];
          ei[u] = ei[u + 1];
        }
        ed[K - 1] = __int_as_float(0x7f800000);
        ei[K - 1] = 0x7fffffff;
      }
    }
  }
}

}  // namespace snb

using namespace snb;

SNB_API int snb_transpose_cn(const float* x, int B, int C, int N, float* xT, void* stream) {
  if (B < 0 || C <= 0 || N < 0) return SNB_EINVAL;
  if (B > 65535 || (C + 31) / 32 > 65535) return SNB_ELIMIT;
  if (B == 0 || N == 0) return SNB_OK;
  transpose_cn_kernel<<<dim3((unsigned)((N + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)B), 256, 0, (cudaStream_t)stream>>>(x, C, N, xT);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

// workspace: norms [B,N] floats + per-sample maxima [B] (16-byte aligned pieces)
SNB_API size_t snb_knn_pruned_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  return (((size_t)B * N * sizeof(float) + 15) & ~(size_t)15) + (((size_t)B * sizeof(unsigned) + 15) & ~(size_t)15);
}

SNB_API int snb_knn_pruned(const float* xT, const float* gram, int B, int C, int N, int k, int* idx, void* workspace, size_t workspace_bytes,
                           void* stream) {
  if (B < 0 || C <= 0 || N < 0 || k <= 0) return SNB_EINVAL;
  if (k > KP_MAXK || k > N) return SNB_ELIMIT;
  if (B == 0 || N == 0) return SNB_OK;
  if (!workspace || workspace_bytes < snb_knn_pruned_workspace_bytes(B, N)) return SNB_EWORKSPACE;
  if (((uintptr_t)workspace & 15) != 0 || ((uintptr_t)xT & 15) != 0) return SNB_EALIGN;
  cudaStream_t s = (cudaStream_t)stream;
  float* nrm = (float*)workspace;
  unsigned* nmax = (unsigned*)((char*)workspace + (((size_t)B * N * sizeof(float) + 15) & ~(size_t)15));
  SNB_CUDA(cudaMemsetAsync(nmax, 0, sizeof(unsigned) * (size_t)B, s));
  const size_t rows = (size_t)B * N;
  knn_norm_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, s>>>(xT, C, N, rows, nrm, nmax);
  SNB_LAUNCH_CHECK();
  const unsigned grid = (unsigned)((rows + 7) / 8);
  // SNB_KNN_PRUNE_CACHE=0: the approximate distances are recomputed from the Gram row (L1 / L2) for the list pass instead of being
  // kept in 64 registers per lane -- 64 registers per thread and twice the resident warps (measurement switch)
  const char* sw = getenv("SNB_KNN_PRUNE_CACHE");
  const bool cache = !(sw && sw[0] == '0');
  if (cache) {
    if (k <= 8) knn_prune_kernel<8, true><<<grid, 256, 0, s>>>(xT, gram, nrm, nmax, C, N, rows, k, idx);
    else if (k <= 16) knn_prune_kernel<16, true><<<grid, 256, 0, s>>>(xT, gram, nrm, nmax, C, N, rows, k, idx);
    else knn_prune_kernel<32, true><<<grid, 256, 0, s>>>(xT, gram, nrm, nmax, C, N, rows, k, idx);
  } else {
    if (k <= 8) knn_prune_kernel<8, false><<<grid, 256, 0, s>>>(xT, gram, nrm, nmax, C, N, rows, k, idx);
    else if (k <= 16) knn_prune_kernel<16, false><<<grid, 256, 0, s>>>(xT, gram, nrm, nmax, C, N, rows, k, idx);
    else knn_prune_kernel<32, false><<<grid, 256, 0, s>>>(xT, gram, nrm, nmax, C, N, rows, k, idx);
  }
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
