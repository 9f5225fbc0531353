// chamfer.cu -- bidirectional nearest-neighbour search + its backward, sm_100a.
//
// Replaces chamfer_dist_kernel / chamfer_dist_grad_kernel of the reference
// (cuda/chamfer_dist/chamfer.cu:15-145,173-201 == cuda/chamfer_distance/chamfer_distance.cu:6-137,158-187).
// Contract: s(i,j) = fma(dz,dz,fma(dx,dx,dy*dy)) with d* = ref_j - query_i;
// dist[i] = min_j s, idx[i] = smallest j attaining it.  Results are bit-exact by construction:
// the minimum of a set of floats does not depend on traversal order and the tie rule is resolved
// explicitly (first strictly-smaller group, then first equal point inside the group).
//
// Design: one launch covers both directions and the whole batch (grid.z = direction, grid.y = sample).
// Reference points stream through shared memory as raw AoS tiles fetched by TMA 1-D bulk copies
// (cp.async.bulk, double-buffered on mbarriers); 4 points = 48 B are read as three broadcast LDS.128.
// Each thread keeps Q queries in registers; per group of 4 references it evaluates the 4 distances,
// reduces them with FMNMX and keeps only (best value, group index).  The index inside the group is
// recovered once in the epilogue, which removes 2 of the 3 compare/select slots per pair.
#include "common.cuh"

namespace snb {

constexpr int CH_THREADS = 128;
constexpr int CH_TILE = 1024;  // reference points per shared-memory tile (multiple of 4)

template <int Q>
__global__ void __launch_bounds__(CH_THREADS) chamfer_nn_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2, int N,
                                                                 int M, float* __restrict__ dist1, float* __restrict__ dist2,
                                                                 int* __restrict__ idx1, int* __restrict__ idx2, int use_tma) {
  const int dir = blockIdx.z, b = blockIdx.y;
  const int nq = dir ? M : N, nr = dir ? N : M;
  const int q0 = blockIdx.x * (CH_THREADS * Q);
  if (q0 >= nq) return;  // uniform per block
  const float* __restrict__ qp = (dir ? xyz2 : xyz1) + (size_t)b * nq * 3;
  const float* __restrict__ rp = (dir ? xyz1 : xyz2) + (size_t)b * nr * 3;
  float* __restrict__ dout = (dir ? dist2 : dist1) + (size_t)b * nq;
  int* __restrict__ iout = (dir ? idx2 : idx1) + (size_t)b * nq;

  __shared__ __align__(128) float tile[2][CH_TILE * 3];
  __shared__ __align__(8) uint64_t bar[2];

  float qx[Q], qy[Q], qz[Q], best[Q];
  int bgrp[Q];
#pragma unroll
  for (int t = 0; t < Q; t++) {
    int qi = q0 + threadIdx.x + t * CH_THREADS;
    qi = qi < nq ? qi : nq - 1;
    qx[t] = qp[qi * 3 + 0];
    qy[t] = qp[qi * 3 + 1];
    qz[t] = qp[qi * 3 + 2];
    best[t] = __int_as_float(0x7f800000);  // +inf
    bgrp[t] = 0;
  }

  const int ntiles = (nr + CH_TILE - 1) / CH_TILE;
  if (use_tma) {
    if (threadIdx.x == 0) {
      mbar_init(&bar[0], 1);
      mbar_init(&bar[1], 1);
      fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const int cnt = nr < CH_TILE ? nr : CH_TILE;
      mbar_expect_tx(&bar[0], cnt * 12);
      tma_load_1d(tile[0], rp, cnt * 12, &bar[0]);
    }
  }

  for (int t = 0; t < ntiles; t++) {
    const int buf = t & 1;
    const int base = t * CH_TILE;
    const int cnt = (nr - base) < CH_TILE ? (nr - base) : CH_TILE;
    int cnt4 = cnt & ~3;
    if (use_tma) {
      if (threadIdx.x == 0 && t + 1 < ntiles) {  // prefetch the next tile into the other buffer
        const int nb = base + CH_TILE;
        const int ncnt = (nr - nb) < CH_TILE ? (nr - nb) : CH_TILE;
        mbar_expect_tx(&bar[buf ^ 1], ncnt * 12);
        tma_load_1d(tile[buf ^ 1], rp + (size_t)nb * 3, ncnt * 12, &bar[buf ^ 1]);
      }
      mbar_wait(&bar[buf], (t >> 1) & 1);
    } else {
      for (int i = threadIdx.x; i < cnt * 3; i += CH_THREADS) tile[buf][i] = rp[(size_t)base * 3 + i];
      if (cnt4 != cnt) {  // pad the ragged tail of the last group with +inf coordinates
        for (int i = cnt * 3 + threadIdx.x; i < (cnt4 + 4) * 3; i += CH_THREADS) tile[buf][i] = __int_as_float(0x7f800000);
        cnt4 += 4;
      }
      __syncthreads();
    }

    const float4* __restrict__ tp = reinterpret_cast<const float4*>(tile[buf]);
#pragma unroll 2
    for (int k = 0; k < cnt4; k += 4) {
      const float4 a = tp[(k >> 2) * 3 + 0];
      const float4 bb = tp[(k >> 2) * 3 + 1];
      const float4 c = tp[(k >> 2) * 3 + 2];
#pragma unroll
      for (int u = 0; u < Q; u++) {
        const float d0 = sqdist3(__fsub_rn(a.x, qx[u]), __fsub_rn(a.y, qy[u]), __fsub_rn(a.z, qz[u]));
        const float d1 = sqdist3(__fsub_rn(a.w, qx[u]), __fsub_rn(bb.x, qy[u]), __fsub_rn(bb.y, qz[u]));
        const float d2 = sqdist3(__fsub_rn(bb.z, qx[u]), __fsub_rn(bb.w, qy[u]), __fsub_rn(c.x, qz[u]));
        const float d3 = sqdist3(__fsub_rn(c.y, qx[u]), __fsub_rn(c.z, qy[u]), __fsub_rn(c.w, qz[u]));
        const float m = fminf(fminf(d0, d1), fminf(d2, d3));
        if (m < best[u]) {
          best[u] = m;
          bgrp[u] = base + k;
        }
      }
    }
    __syncthreads();  // everyone is done with tile[buf] before it is refilled
  }

  // epilogue: first point of the winning group whose distance equals the minimum
#pragma unroll
  for (int u = 0; u < Q; u++) {
    const int qi = q0 + threadIdx.x + u * CH_THREADS;
    if (qi >= nq) continue;
    int bi = bgrp[u];
#pragma unroll
    for (int j = 3; j >= 0; j--) {
      const int r = bgrp[u] + j;
      if (r < nr) {
        const float d = sqdist3(__fsub_rn(rp[r * 3 + 0], qx[u]), __fsub_rn(rp[r * 3 + 1], qy[u]), __fsub_rn(rp[r * 3 + 2], qz[u]));
        if (d == best[u]) bi = r;
      }
    }
    dout[qi] = best[u];
    iout[qi] = bi;
  }
}

// ---- backward ---------------------------------------------------------------------------------------
// grad_a[j] (direct, plain store) = 2 g_j (a_j - c_idx[j]); grad_c[idx[j]] -= the same (RED.ADD).
// Pass 1 writes every row of both outputs, pass 2 scatters: no memset, 3 atomics per point instead of 6.
__global__ void __launch_bounds__(256) chamfer_grad_direct_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2, int N, int M,
                                                                   const int* __restrict__ idx1, const int* __restrict__ idx2,
                                                                   const float* __restrict__ g1, const float* __restrict__ g2,
                                                                   float* __restrict__ gx1, float* __restrict__ gx2) {
  const int dir = blockIdx.z, b = blockIdx.y;
  const int n = dir ? M : N, m = dir ? N : M;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const float* a = (dir ? xyz2 : xyz1) + (size_t)b * n * 3;
  const float* c = (dir ? xyz1 : xyz2) + (size_t)b * m * 3;
  const int j2 = ((dir ? idx2 : idx1) + (size_t)b * n)[j];
  const float g = ((dir ? g2 : g1) + (size_t)b * n)[j] * 2.f;
  float* ga = (dir ? gx2 : gx1) + (size_t)b * n * 3;
  ga[j * 3 + 0] = g * (a[j * 3 + 0] - c[j2 * 3 + 0]);
  ga[j * 3 + 1] = g * (a[j * 3 + 1] - c[j2 * 3 + 1]);
  ga[j * 3 + 2] = g * (a[j * 3 + 2] - c[j2 * 3 + 2]);
}

__global__ void __launch_bounds__(256) chamfer_grad_scatter_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2, int N, int M,
                                                                    const int* __restrict__ idx1, const int* __restrict__ idx2,
                                                                    const float* __restrict__ g1, const float* __restrict__ g2,
                                                                    float* __restrict__ gx1, float* __restrict__ gx2) {
  const int dir = blockIdx.z, b = blockIdx.y;
  const int n = dir ? M : N, m = dir ? N : M;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const float* a = (dir ? xyz2 : xyz1) + (size_t)b * n * 3;
  const float* c = (dir ? xyz1 : xyz2) + (size_t)b * m * 3;
  const int j2 = ((dir ? idx2 : idx1) + (size_t)b * n)[j];
  const float g = ((dir ? g2 : g1) + (size_t)b * n)[j] * 2.f;
  if (g == 0.f) return;  // adding -0/+0 leaves every sum unchanged (consistency loss passes a zero grad_dist2)
  float* gc = (dir ? gx1 : gx2) + (size_t)b * m * 3;
  atomicAdd(&gc[j2 * 3 + 0], -(g * (a[j * 3 + 0] - c[j2 * 3 + 0])));
  atomicAdd(&gc[j2 * 3 + 1], -(g * (a[j * 3 + 1] - c[j2 * 3 + 1])));
  atomicAdd(&gc[j2 * 3 + 2], -(g * (a[j * 3 + 2] - c[j2 * 3 + 2])));
}

// chamfer_bvh.cu: exact search with spatial pruning (same bits, O(N log N) work); 0 bytes = shape not served
size_t chamfer_bvh_workspace_bytes(int B, int N, int M);
int chamfer_bvh_launch(const float* xyz1, const float* xyz2, int B, int N, int M, float* dist1, float* dist2, int* idx1, int* idx2, void* workspace,
                       cudaStream_t s);

}  // namespace snb

using namespace snb;

SNB_API size_t snb_chamfer_workspace_bytes(int B, int N, int M) {
  if (B <= 0 || N <= 0 || M <= 0) return 0;
  return chamfer_bvh_workspace_bytes(B, N, M);
}

// include/sparenet_b200.h: snb_chamfer_fwd.  With a workspace of snb_chamfer_workspace_bytes() (> 0 for 256 <= N, M <= 16384) the
// pruned search runs; without one (or for other shapes) the brute-force TMA-tiled kernel does.  Both return identical bits.
SNB_API int snb_chamfer_fwd(const float* xyz1, const float* xyz2, int B, int N, int M, float* dist1, float* dist2, int* idx1, int* idx2,
                            void* workspace, size_t workspace_bytes, void* stream) {
  if (B < 0 || N < 0 || M < 0) return SNB_EINVAL;
  if (B == 0 || (N == 0 && M == 0)) return SNB_OK;
  if (N == 0 || M == 0) return SNB_EINVAL;  // a nearest neighbour in an empty set is undefined
  if (B > 65535) return SNB_ELIMIT;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t need = chamfer_bvh_workspace_bytes(B, N, M);
  if (need > 0 && workspace != nullptr && workspace_bytes >= need && (((uintptr_t)workspace & 15) == 0))
    return chamfer_bvh_launch(xyz1, xyz2, B, N, M, dist1, dist2, idx1, idx2, workspace, s);
  constexpr int Q = 4;
  const int qpb = CH_THREADS * Q;
  const int nmax = N > M ? N : M;
  // TMA bulk copies need 16-byte aligned sources and sizes: every per-sample base is 12*n*b bytes in
  const bool tma = (N % 4 == 0) && (M % 4 == 0) && (((uintptr_t)xyz1 & 15) == 0) && (((uintptr_t)xyz2 & 15) == 0);
  dim3 grid((nmax + qpb - 1) / qpb, B, 2);
  chamfer_nn_kernel<Q><<<grid, CH_THREADS, 0, s>>>(xyz1, xyz2, N, M, dist1, dist2, idx1, idx2, tma ? 1 : 0);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

// include/sparenet_b200.h: snb_chamfer_bwd
SNB_API int snb_chamfer_bwd(const float* xyz1, const float* xyz2, int B, int N, int M, const int* idx1, const int* idx2, const float* g1,
                            const float* g2, float* gx1, float* gx2, void* stream) {
  if (B < 0 || N < 0 || M < 0) return SNB_EINVAL;
  if (B == 0 || N == 0 || M == 0) return SNB_OK;
  if (B > 65535) return SNB_ELIMIT;
  cudaStream_t s = (cudaStream_t)stream;
  const int nmax = N > M ? N : M;
  dim3 grid((nmax + 255) / 256, B, 2);
  chamfer_grad_direct_kernel<<<grid, 256, 0, s>>>(xyz1, xyz2, N, M, idx1, idx2, g1, g2, gx1, gx2);
  SNB_LAUNCH_CHECK();
  chamfer_grad_scatter_kernel<<<grid, 256, 0, s>>>(xyz1, xyz2, N, M, idx1, idx2, g1, g2, gx1, gx2);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
