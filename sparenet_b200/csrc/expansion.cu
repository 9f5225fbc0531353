// expansion.cu -- expansion penalty (per-primitive minimum spanning tree), sm_100a.
//
// Replaces calc_penalty / calc_grad (cuda/expansion_penalty/expansion_penalty_cuda.cu:7-149,167-184).
// Contract: Prim from local vertex 0 on sqrtf(fma(dz,dz,fma(dx,dx,dy*dy))) edge costs,
// argmin ties -> larger index; mean edge by the pairwise (up-sweep shaped) sum / (p-1); leaf peeling in
// synchronous rounds, "larger index peels" between two facing leaves; edges longer than alpha*mean are
// written to the peeled endpoint.  mean_mst_length[b] = (sum over primitives, in index order) / (n/p).
//
// Design: ONE WARP per primitive, no block barriers at all.  Each lane owns p/32 vertices in registers
// (coordinates, tentative distance, tentative parent); a Prim round is a register relax + a 5-step
// shuffle arg-min carrying (distance, index).  The reference's 2 x B*n*512 global adjacency scratch
// (2.1 GB at B=32, n=16384) is replaced by a parent array + xor-of-neighbours in shared memory.
#include <math.h>
#include "common.cuh"

namespace snb {

constexpr int EX_WARPS = 2;  // warps (= primitives) per block

struct ExSmem {
  float xyz[512 * 3];
  float ecost[512];  // cost of the tree edge (v, parent[v])
  int parent[512];
  int cnt[512];  // remaining degree
  int xr[512];   // xor of the remaining neighbours' ids
};

// Everything after Prim, by ONE warp on the shared-memory tree: mean edge length, synchronous leaf peeling, outputs.
template <int VPL>
__device__ __forceinline__ void expansion_tail(ExSmem& s, int lane, int P, size_t pbase, int N, int prim, float alpha, float (&esum)[VPL],
                                               float* __restrict__ dist, int* __restrict__ idx, float* __restrict__ prim_mean) {
  // mean edge length: the reference's in-place up-sweep (:103-117) is the balanced pairwise tree in index
  // order.  Levels 1..5 pair neighbouring lanes (xor shuffles), the remaining levels pair register slots.
#pragma unroll
  for (int i = 0; i < VPL; i++) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      if (o < P) esum[i] = __fadd_rn(esum[i], __shfl_xor_sync(0xffffffffu, esum[i], o));
    }
  }
#pragma unroll
  for (int st = 1; st < VPL; st <<= 1) {
#pragma unroll
    for (int i = 0; i + st < VPL; i += 2 * st) esum[i] = __fadd_rn(esum[i], esum[i + st]);
  }
  // for P < 32 the xor tree above summed lanes >= P too; they hold 0 and x+0 == x exactly
  const float mean_dis = esum[0] / (float)(P - 1);
  if (lane == 0) prim_mean[prim] = mean_dis;
  const float thr = mean_dis * alpha;

  // synchronous leaf peeling (:123-146)
  float dv[VPL];
  int iv[VPL];
#pragma unroll
  for (int i = 0; i < VPL; i++) {
    dv[i] = 0.f;
    iv[i] = -1;
  }
  const int ybase = (int)(pbase % (size_t)N);  // index of local vertex 0 inside its sample
  for (;;) {
    int peel_u[VPL];
    bool any_leaf = false;
#pragma unroll
    for (int i = 0; i < VPL; i++) {
      const int v = lane + 32 * i;
      peel_u[i] = -1;
      if (v < P && s.cnt[v] == 1) {
        any_leaf = true;
        const int u = s.xr[v];
        const int cu = s.cnt[u];
        if (cu > 1 || (cu == 1 && v > u)) peel_u[i] = u;
      }
    }
    if (!__any_sync(0xffffffffu, any_leaf)) break;
    __syncwarp();  // all decisions are taken on the pre-round state
#pragma unroll
    for (int i = 0; i < VPL; i++) {
      const int u = peel_u[i];
      if (u >= 0) {
        const int v = lane + 32 * i;
        const float c = (s.parent[v] == u) ? s.ecost[v] : s.ecost[u];
        s.cnt[v] = 0;
        s.xr[v] = 0;
        atomicSub(&s.cnt[u], 1);
        atomicXor(&s.xr[u], v);
        if (c > thr) {
          dv[i] = c;
          iv[i] = ybase + u;
        }
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int i = 0; i < VPL; i++) {
    const int v = lane + 32 * i;
    if (v < P) {
      dist[pbase + v] = dv[i];
      idx[pbase + v] = iv[i];
    }
  }
}

template <int VPL>
__global__ void __launch_bounds__(EX_WARPS * 32) expansion_kernel(const float* __restrict__ xyz, int N, int P, int nprim_total, float alpha,
                                                                  float* __restrict__ dist, int* __restrict__ idx,
                                                                  float* __restrict__ prim_mean) {
  __shared__ ExSmem sm[EX_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int prim = blockIdx.x * EX_WARPS + warp;  // global primitive id = b*(N/P) + y
  if (prim >= nprim_total) return;
  ExSmem& s = sm[warp];
  const size_t pbase = (size_t)prim * P;  // == b*N + y*P
  const float* __restrict__ src = xyz + pbase * 3;
  for (int i = lane; i < P * 3; i += 32) s.xyz[i] = src[i];
  for (int i = lane; i < P; i += 32) {
    s.cnt[i] = 0;
    s.xr[i] = 0;
    s.parent[i] = -1;
    s.ecost[i] = 0.f;
  }
  __syncwarp();

  float x[VPL], y[VPL], z[VPL], cur[VPL], esum[VPL];
  int cidx[VPL];
  unsigned vis = 0;  // bit i: vertex lane+32*i already in the tree (or out of range)
#pragma unroll
  for (int i = 0; i < VPL; i++) {
    const int v = lane + 32 * i;
    const bool ok = v < P;
    x[i] = ok ? s.xyz[v * 3 + 0] : 0.f;
    y[i] = ok ? s.xyz[v * 3 + 1] : 0.f;
    z[i] = ok ? s.xyz[v * 3 + 2] : 0.f;
    cur[i] = 1e9f;
    cidx[i] = 0;
    esum[i] = 0.f;
    if (!ok || v == 0) vis |= 1u << i;
  }

  int last = 0;
  for (int r = 0; r < P - 1; r++) {
    const float xl = s.xyz[last * 3 + 0], yl = s.xyz[last * 3 + 1], zl = s.xyz[last * 3 + 2];
    float bd = 2e9f;
    int bi = -1;
#pragma unroll
    for (int i = 0; i < VPL; i++) {
      if (!((vis >> i) & 1u)) {
        const float d = sqrtf(sqdist3(__fsub_rn(x[i], xl), __fsub_rn(y[i], yl), __fsub_rn(z[i], zl)));
        if (d < cur[i]) {
          cur[i] = d;
          cidx[i] = last;
        }
        if (cur[i] <= bd) {  // ascending i == ascending index: '<=' lets the larger index win ties
          bd = cur[i];
          bi = lane + 32 * i;
        }
      }
    }
    // arg-min as two warp reductions instead of a 5-level (distance, index) shuffle butterfly:
    // distances are >= 0, so their bit patterns order like the values; ties go to the larger index
    const unsigned mbits = __reduce_min_sync(0xffffffffu, __float_as_uint(bd));
    last = __reduce_max_sync(0xffffffffu, __float_as_uint(bd) == mbits ? bi : -1);
    if ((last & 31) == lane) {  // owner lane records the tree edge
      const int slot = last >> 5;
      int u = 0;
      float c = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; i++)
        if (i == slot) {
          u = cidx[i];
          c = cur[i];
          esum[i] = c;
        }
      vis |= 1u << slot;
      s.parent[last] = u;
      s.ecost[last] = c;
      s.cnt[last] += 1;
      s.xr[last] ^= u;
      atomicAdd(&s.cnt[u], 1);  // u may live in another lane; same-warp smem RMW on distinct addresses
      atomicXor(&s.xr[u], last);
    }
    __syncwarp();
  }

  expansion_tail<VPL>(s, lane, P, pbase, N, prim, alpha, esum, dist, idx, prim_mean);
}

// P >= 128: FOUR warps per primitive for the Prim phase.  One warp per primitive leaves 7 warps on an SM (1024 primitives in a batch
// of 32) with ~530 dependent instructions per round each: 28.6 % of the issue slots.  Here every lane holds P/128 vertices, each
// warp reduces its arg-min as above, the four results meet in shared memory (double-buffered by round parity: ONE block barrier per
// round) and every thread takes the minimum -- same distances, same tie rule (larger index), same tree.  Warp 0 then runs the
// unchanged tail on the shared-memory tree.
template <int VPW, int WPP>   // vertices per lane, warps per primitive (VPW * WPP * 32 == P)
__global__ void __launch_bounds__(WPP * 32) expansion_kernel_mw(const float* __restrict__ xyz, int N, int P, float alpha, float* __restrict__ dist,
                                                           int* __restrict__ idx, float* __restrict__ prim_mean) {
  __shared__ ExSmem s;
  __shared__ unsigned long long slots[2][WPP];
  constexpr int T = WPP * 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int prim = blockIdx.x;  // global primitive id = b*(N/P) + y
  const size_t pbase = (size_t)prim * P;
  const float* __restrict__ src = xyz + pbase * 3;
  for (int i = tid; i < P * 3; i += T) s.xyz[i] = src[i];
  for (int i = tid; i < P; i += T) {
    s.cnt[i] = 0;
    s.xr[i] = 0;
    s.parent[i] = -1;
    s.ecost[i] = 0.f;
  }
  __syncthreads();
  float x[VPW], y[VPW], z[VPW], cur[VPW];
  int cidx[VPW];
  unsigned vis = 0;  // bit i: vertex tid+T*i already in the tree
#pragma unroll
  for (int i = 0; i < VPW; i++) {
    const int v = tid + T * i;
    x[i] = s.xyz[v * 3 + 0];
    y[i] = s.xyz[v * 3 + 1];
    z[i] = s.xyz[v * 3 + 2];
    cur[i] = 1e9f;
    cidx[i] = 0;
    if (v == 0) vis |= 1u << i;
  }
  int last = 0;
  for (int r = 0; r < P - 1; r++) {
    const float xl = s.xyz[last * 3 + 0], yl = s.xyz[last * 3 + 1], zl = s.xyz[last * 3 + 2];
    float bd = 2e9f;
    int bi = -1;
#pragma unroll
    for (int i = 0; i < VPW; i++) {
      if (!((vis >> i) & 1u)) {
        const float d = sqrtf(sqdist3(__fsub_rn(x[i], xl), __fsub_rn(y[i], yl), __fsub_rn(z[i], zl)));
        if (d < cur[i]) {
          cur[i] = d;
          cidx[i] = last;
        }
        if (cur[i] <= bd) {  // ascending i == ascending index: '<=' lets the larger index win ties
          bd = cur[i];
          bi = tid + T * i;
        }
      }
    }
    const unsigned mbits = __reduce_min_sync(0xffffffffu, __float_as_uint(bd));
    const int wl = __reduce_max_sync(0xffffffffu, __float_as_uint(bd) == mbits ? bi : -1);
    // (distance bits, 0x7fffffff - index): the minimum key is the smallest distance and, among equals, the largest index
    if (lane == 0) slots[r & 1][warp] = ((unsigned long long)mbits << 32) | (unsigned)(0x7fffffff - wl);
    __syncthreads();
    unsigned long long k = slots[r & 1][0];
#pragma unroll
    for (int q = 1; q < WPP; q++) {
      const unsigned long long o = slots[r & 1][q];
      k = o < k ? o : k;
    }
    last = 0x7fffffff - (int)(unsigned)k;
    if ((last % T) == tid) {  // owner thread records the tree edge
      const int slot = last / T;
      int u = 0;
      float c = 0.f;
#pragma unroll
      for (int i = 0; i < VPW; i++)
        if (i == slot) {
          u = cidx[i];
          c = cur[i];
        }
      vis |= 1u << slot;
      s.parent[last] = u;
      s.ecost[last] = c;
      atomicAdd(&s.cnt[last], 1);
      atomicXor(&s.xr[last], u);
      atomicAdd(&s.cnt[u], 1);
      atomicXor(&s.xr[u], last);
    }
  }
  __syncthreads();
  if (warp == 0) {
    constexpr int VPL = VPW * WPP;
    float esum[VPL];
#pragma unroll
    for (int i = 0; i < VPL; i++) esum[i] = s.ecost[lane + 32 * i];   // cost of the edge that brought vertex lane+32i in (vertex 0: 0)
    expansion_tail<VPL>(s, lane, P, pbase, N, prim, alpha, esum, dist, idx, prim_mean);
  }
}

// mean_mst_length[b] = (sum_y prim_mean[b, y], ascending y) / np   (deterministic; the reference atomicAdds)
__global__ void expansion_mean_kernel(const float* __restrict__ prim_mean, int B, int np, float* __restrict__ mml) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float acc = 0.f;
  for (int y = 0; y < np; y++) acc = __fadd_rn(acc, prim_mean[(size_t)b * np + y]);
  mml[b] = acc / (float)np;
}

// grad_xyz[j] = 2 g_j (x_j - x_idx[j]) where idx[j] != -1, else 0  (expansion_penalty_cuda.cu:167-184)
__global__ void __launch_bounds__(256) expansion_grad_kernel(const float* __restrict__ xyz, int N, size_t total, const float* __restrict__ g,
                                                              const int* __restrict__ idx, float* __restrict__ gx) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int j2 = idx[i];
  float a = 0.f, b = 0.f, c = 0.f;
  if (j2 != -1) {
    const size_t o = (i / N) * N + j2;
    const float gg = g[i] * 2.f;
    a = gg * (xyz[i * 3 + 0] - xyz[o * 3 + 0]);
    b = gg * (xyz[i * 3 + 1] - xyz[o * 3 + 1]);
    c = gg * (xyz[i * 3 + 2] - xyz[o * 3 + 2]);
  }
  gx[i * 3 + 0] = a;
  gx[i * 3 + 1] = b;
  gx[i * 3 + 2] = c;
}

}  // namespace snb

using namespace snb;

SNB_API size_t snb_expansion_workspace_bytes(int B, int N, int primitive_size) {
  if (B <= 0 || N <= 0 || primitive_size <= 0) return 0;
  return sizeof(float) * (size_t)B * (size_t)(N / primitive_size);
}

SNB_API int snb_expansion_fwd(const float* xyz, int B, int N, int P, float alpha, float* dist, int* assignment, float* mean_mst_length,
                              void* workspace, size_t workspace_bytes, void* stream) {
  if (B < 0 || N < 0) return SNB_EINVAL;
  if (P < 2 || P > 512 || (P & (P - 1)) || (N % P)) return SNB_ELIMIT;
  if (B == 0 || N == 0) return SNB_OK;
  if (workspace_bytes < snb_expansion_workspace_bytes(B, N, P) || !workspace) return SNB_EWORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  const int np = N / P;
  const long long nprim = (long long)B * np;
  if (nprim > 0x7fffffffLL) return SNB_ELIMIT;
  float* prim_mean = (float*)workspace;
  const int grid = (int)((nprim + EX_WARPS - 1) / EX_WARPS);
  const int vpl = P >= 32 ? P / 32 : 1;
#define EX_LAUNCH(V) expansion_kernel<V><<<grid, EX_WARPS * 32, 0, s>>>(xyz, N, P, (int)nprim, alpha, dist, assignment, prim_mean)
#define EX_LAUNCH_MW(V, W) expansion_kernel_mw<V, W><<<(int)nprim, W * 32, 0, s>>>(xyz, N, P, alpha, dist, assignment, prim_mean)
  switch (vpl) {
    case 16: EX_LAUNCH_MW(4, 4); break;   // P = 512, 256, 128: four warps per primitive in the Prim phase (measured at P = 512: 2 warps 0.63 ms, 4 warps 0.53, 8 warps 0.78)
    case 8: EX_LAUNCH_MW(2, 4); break;
    case 4: EX_LAUNCH_MW(1, 4); break;
    case 2: EX_LAUNCH(2); break;
    default: EX_LAUNCH(1); break;
  }
#undef EX_LAUNCH
#undef EX_LAUNCH_MW
  SNB_LAUNCH_CHECK();
  expansion_mean_kernel<<<(B + 127) / 128, 128, 0, s>>>(prim_mean, B, np, mean_mst_length);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

SNB_API int snb_expansion_bwd(const float* xyz, int B, int N, const float* grad_dist, const int* assignment, float* grad_xyz, void* stream) {
  if (B < 0 || N < 0) return SNB_EINVAL;
  if (B == 0 || N == 0) return SNB_OK;
  const size_t total = (size_t)B * N;
  expansion_grad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(xyz, N, total, grad_dist, assignment, grad_xyz);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
