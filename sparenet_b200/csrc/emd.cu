// emd.cu -- approximate Earth-Mover's distance by the auction algorithm, sm_100a.
//
// Replaces the 7-kernels-per-iteration host loop of the reference (cuda/emd/emd_cuda.cu:23-226,256-269:
// clear / calc_unass_cnt / calc_unass_cnt_sum / calc_unass_idx / Bid / GetMax / Assign, then CalcDist) and
// NmDistanceGradKernel (:284-300).  Contract, reproduced bit-exactly:
//   v_k = (float)((3.0 - (double)sqrtf(s_k)) - (double)price[k]),  s_k = fma(dz,dz,fma(dx,dx,dy*dy)),
//   best = max_k v_k, better = second largest (multiset), inc = (best - better) + eps,
//   exact-tie best_i = the reference's thread-partition order (slice of the 2048-tile, then k),
//   GetMax window +-1e-6 in double with the race resolved as "largest bidder index wins",
//   Assign with eviction, last iteration force-assigns every remaining bidder.
//
// Design: ONE persistent kernel; a thread-block cluster owns one sample and runs all `iters` rounds with
// cluster barriers between the phases (no host loop, no 351 launches).  Objects live in the workspace as
// packed float4 (x, y, z, price) so a tile is one TMA bulk copy and one LDS.128 per object.
//   Bid hot loop: max / second-max are order statistics, so only pairs that can still change them need
//   the exact (sqrt + fp64) evaluation.  A conservative fp32 filter  max(c - p, 0)^2 > s  with
//   c = 3 - better + slack  (9 fp32 ops, no sqrt, no fp64) discards the rest; flagged pairs -- O(log n)
//   per bidder plus near-ties -- take the exact path.  The filter only ever over-flags, so results are
//   identical to evaluating every pair exactly.
//   Work mapping per round: G = 1..32 threads per bidder (all cluster threads busy when few bidders are
//   left), Q = 1..8 bidders per thread when bidders outnumber threads (objects broadcast from smem).
#include <math.h>
#include "common.cuh"

namespace snb {

constexpr int EMD_THREADS = 512;
constexpr int EMD_TILE = 2048;       // objects per smem tile == the reference's Bid tile (emd_cuda.cu:97)
constexpr int EMD_MAX_CLUSTER = 8;
constexpr size_t EMD_SMEM = (size_t)2 * EMD_TILE * 16 + 64;

struct EmdWs {  // per-sample views into the caller's workspace
  float4* obj;          // [n] x,y,z,price
  int* assignment_inv;  // [n]
  int* bid;             // [n]
  float* bid_inc;       // [n]
  float* max_inc;       // [n]
  int* max_idx;         // [n]
  int* unass;           // [n]
  int* counter;         // [2]: ping-pong bidder counters
};

__host__ __device__ inline size_t emd_ws_per_sample(int n) { return (size_t)n * (16 + 6 * 4) + 64; }

__device__ __forceinline__ EmdWs emd_ws_view(void* ws, int b, int n) {
  char* p = (char*)ws + (size_t)b * emd_ws_per_sample(n);
  EmdWs w;
  w.obj = (float4*)p;                 p += (size_t)n * 16;
  w.assignment_inv = (int*)p;         p += (size_t)n * 4;
  w.bid = (int*)p;                    p += (size_t)n * 4;
  w.bid_inc = (float*)p;              p += (size_t)n * 4;
  w.max_inc = (float*)p;              p += (size_t)n * 4;
  w.max_idx = (int*)p;                p += (size_t)n * 4;
  w.unass = (int*)p;                  p += (size_t)n * 4;
  w.counter = (int*)p;
  return w;
}

__device__ __forceinline__ void atomic_max_float(float* addr, float val) {
  if (val >= 0.f) atomicMax((int*)addr, __float_as_int(val));
  else atomicMin((unsigned*)addr, __float_as_uint(val));
}

// position of object k in the reference's Bid traversal for this round (emd_cuda.cu:107-108,134-138,166-172):
// thread slice inside its 2048-tile first, then k.  Smaller key wins an exact tie.  k < 0 -> worst key.
__device__ __forceinline__ unsigned long long emd_tie_key(int k, int n, int tpu) {
  if (k < 0) return ~0ull;
  const int k2 = k & ~(EMD_TILE - 1);
  const int end_k = (n - k2) < EMD_TILE ? (n - k2) : EMD_TILE;
  const int delta = (end_k + tpu - 1) / tpu;
  const unsigned slice = (unsigned)((k - k2) / delta);
  return ((unsigned long long)slice << 32) | (unsigned)k;
}

struct BidState {
  float best, better, c;
  int best_i;
};

__device__ __forceinline__ float emd_thr(float better) {
  // c = 3 - better + slack; slack 1e-4 * max(1,|3-better|) dominates every rounding error of the fp32 filter
  const float a = 3.0f - better;
  return a + 1e-4f * fmaxf(1.0f, fabsf(a));
}

__device__ __forceinline__ void emd_exact_update(BidState& st, float s, float price, int k, int n, int tpu) {
  const float d = (float)((3.0 - (double)sqrtf(s)) - (double)price);
  if (d > st.best) {
    st.better = st.best;
    st.best = d;
    st.best_i = k;
    st.c = emd_thr(st.better);
  } else if (d == st.best) {
    st.better = d;
    if (emd_tie_key(k, n, tpu) < emd_tie_key(st.best_i, n, tpu)) st.best_i = k;
    st.c = emd_thr(st.better);
  } else if (d > st.better) {
    st.better = d;
    st.c = emd_thr(st.better);
  }
}

__device__ __forceinline__ void emd_merge(BidState& a, float obest, float obetter, int obi, int n, int tpu) {
  if (obest > a.best) {
    a.better = fmaxf(a.best, obetter);
    a.best = obest;
    a.best_i = obi;
  } else if (obest == a.best) {
    a.better = a.best;
    if (emd_tie_key(obi, n, tpu) < emd_tie_key(a.best_i, n, tpu)) a.best_i = obi;
  } else {
    a.better = fmaxf(a.better, obest);
  }
}

// One Bid pass: this thread serves bidders list[u0 + q*nslots] (q < Q) with sub-lane g of G.
template <int Q>
__device__ __forceinline__ void emd_bid_pass(const EmdWs& w, const float* __restrict__ xyz1, int n, int U, int tpu, float eps, int slot,
                                             int nslots, int u0, int g, int G, float4* tiles, uint64_t* bars, uint32_t (&phase)[2]) {
  BidState st[Q];
  float bx[Q], by[Q], bz[Q];
  int bj[Q];
#pragma unroll
  for (int q = 0; q < Q; q++) {
    const int u = u0 + slot + q * nslots;
    bj[q] = (u < U) ? w.unass[u] : -1;
    const int j = bj[q] >= 0 ? bj[q] : 0;
    bx[q] = xyz1[j * 3 + 0];
    by[q] = xyz1[j * 3 + 1];
    bz[q] = xyz1[j * 3 + 2];
    st[q].best = -1e9f;
    st[q].better = -1e9f;
    st[q].best_i = -1;
    st[q].c = emd_thr(-1e9f);
  }
  const int ntiles = (n + EMD_TILE - 1) / EMD_TILE;
  if (threadIdx.x == 0) {
    const int cnt = n < EMD_TILE ? n : EMD_TILE;
    fence_proxy_async_all();
    mbar_expect_tx(&bars[0], cnt * 16);
    tma_load_1d(tiles, w.obj, cnt * 16, &bars[0]);
  }
  for (int t = 0; t < ntiles; t++) {
    const int buf = t & 1;
    const int base = t * EMD_TILE;
    const int cnt = (n - base) < EMD_TILE ? (n - base) : EMD_TILE;
    if (threadIdx.x == 0 && t + 1 < ntiles) {
      const int nb = base + EMD_TILE;
      const int ncnt = (n - nb) < EMD_TILE ? (n - nb) : EMD_TILE;
      mbar_expect_tx(&bars[buf ^ 1], ncnt * 16);
      tma_load_1d(tiles + (size_t)(buf ^ 1) * EMD_TILE, w.obj + nb, ncnt * 16, &bars[buf ^ 1]);
    }
    mbar_wait(&bars[buf], phase[buf]);
    phase[buf] ^= 1u;
    const float4* __restrict__ tp = tiles + (size_t)buf * EMD_TILE;
    // The filter is evaluated for a GROUP of 8 (object, bidder) pairs first -- straight-line code, one flag word -- and the exact
    // path (rare: ~2 ln n times per bidder plus near-ties) sits behind ONE branch per group instead of one per pair.  That keeps
    // the hot loop small enough for the instruction cache (the per-pair inlined exact path made `no_instruction` the top stall)
    // and gives the few-bidder rounds (Q = 1) eight independent shared-memory loads in flight.  A pair filtered with a threshold
    // that is stale within its group is only ever over-flagged, and flagged pairs are evaluated in ascending k: results unchanged.
    constexpr int KU = Q >= 8 ? 1 : 8 / Q;  // objects per group
#pragma unroll 1
    for (int k0 = g; k0 < cnt; k0 += G * KU) {
      float4 o[KU];
      float sv[KU][Q];
      unsigned flags = 0u;
#pragma unroll
      for (int u = 0; u < KU; u++) {
        const int k = k0 + u * G;
        const bool valid = k < cnt;
        o[u] = tp[valid ? k : k0];
#pragma unroll
        for (int q = 0; q < Q; q++) {
          sv[u][q] = sqdist3(__fsub_rn(o[u].x, bx[q]), __fsub_rn(o[u].y, by[q]), __fsub_rn(o[u].z, bz[q]));
          const float qq = fmaxf(st[q].c - o[u].w, 0.f);
          flags |= (valid && __fmaf_rn(qq, qq, -sv[u][q]) > 0.f) ? (1u << (u * Q + q)) : 0u;
        }
      }
      if (flags) {
#pragma unroll
        for (int u = 0; u < KU; u++)
#pragma unroll
          for (int q = 0; q < Q; q++)
            if ((flags >> (u * Q + q)) & 1u) emd_exact_update(st[q], sv[u][q], o[u].w, base + k0 + u * G, n, tpu);
      }
    }
    __syncthreads();
  }
  // merge the G partial states of each bidder (xor butterfly inside the aligned lane group)
#pragma unroll
  for (int q = 0; q < Q; q++) {
    for (int off = 1; off < G; off <<= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, st[q].best, off);
      const float obb = __shfl_xor_sync(0xffffffffu, st[q].better, off);
      const int oi = __shfl_xor_sync(0xffffffffu, st[q].best_i, off);
      emd_merge(st[q], ob, obb, oi, n, tpu);
    }
    if (g == 0 && bj[q] >= 0) {
      const float inc = __fadd_rn(__fsub_rn(st[q].best, st[q].better), eps);
      w.bid[bj[q]] = st[q].best_i;
      w.bid_inc[bj[q]] = inc;
      atomic_max_float(&w.max_inc[st[q].best_i], inc);
    }
  }
}

__global__ void __launch_bounds__(EMD_THREADS, 1) emd_auction_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2, int n,
                                                                      float eps, int iters, float* __restrict__ dist,
                                                                      int* __restrict__ assignment, void* workspace) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* tiles = reinterpret_cast<float4*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)2 * EMD_TILE * 16);
  const uint32_t cs = cluster_nctarank(), rank = cluster_ctarank();
  const int b = blockIdx.x / cs;
  const int tid = threadIdx.x;
  const int T = cs * EMD_THREADS;          // threads serving this sample
  const int gtid = rank * EMD_THREADS + tid;
  xyz1 += (size_t)b * n * 3;
  xyz2 += (size_t)b * n * 3;
  dist += (size_t)b * n;
  assignment += (size_t)b * n;
  const EmdWs w = emd_ws_view(workspace, b, n);

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  uint32_t phase[2] = {0u, 0u};
  // initial state (emd_module.py:43-54)
  for (int k = gtid; k < n; k += T) {
    w.obj[k] = make_float4(xyz2[k * 3 + 0], xyz2[k * 3 + 1], xyz2[k * 3 + 2], 0.f);
    assignment[k] = -1;
    w.assignment_inv[k] = -1;
    w.max_inc[k] = 0.f;
    w.max_idx[k] = -1;
  }
  if (gtid == 0) {
    w.counter[0] = 0;
    w.counter[1] = 0;
  }
  fence_proxy_async_all();
  __syncthreads();
  cluster_sync_all();

  const int block_cnt = n / 1024;
  for (int it = 0; it < iters; it++) {
    int* cnt_cur = &w.counter[it & 1];
    // ---- compact the unassigned bidders (order is irrelevant to every result) ----
    for (int k0 = 0; k0 < n; k0 += T) {
      const int k = k0 + gtid;
      const bool un = (k < n) && (assignment[k] == -1);
      const unsigned m = __ballot_sync(0xffffffffu, un);
      if (m) {
        const int lane = tid & 31;
        int basepos = 0;
        if (lane == 0) basepos = atomicAdd(cnt_cur, __popc(m));
        basepos = __shfl_sync(0xffffffffu, basepos, 0);
        if (un) w.unass[basepos + __popc(m & ((1u << lane) - 1u))] = k;
      }
    }
    cluster_sync_all();
    const int U = *((volatile int*)cnt_cur);
    if (U == 0) break;  // uniform over the cluster: nothing can change any more
    if (gtid == 0) w.counter[(it + 1) & 1] = 0;
    const bool last = (it == iters - 1);
    const int unass_per_block = (U + block_cnt - 1) / block_cnt;
    const int tpu = 1024 / unass_per_block;  // the reference's thread_per_unass, needed only for tie keys

    // ---- Bid ----
    if (U >= T) {
      const int per = (U + T - 1) / T;
      const int Q = per >= 8 ? 8 : (per >= 4 ? 4 : (per >= 2 ? 2 : 1));
      for (int u0 = 0; u0 < U; u0 += T * Q) {
        if (Q == 8) emd_bid_pass<8>(w, xyz1, n, U, tpu, eps, gtid, T, u0, 0, 1, tiles, bars, phase);
        else if (Q == 4) emd_bid_pass<4>(w, xyz1, n, U, tpu, eps, gtid, T, u0, 0, 1, tiles, bars, phase);
        else if (Q == 2) emd_bid_pass<2>(w, xyz1, n, U, tpu, eps, gtid, T, u0, 0, 1, tiles, bars, phase);
        else emd_bid_pass<1>(w, xyz1, n, U, tpu, eps, gtid, T, u0, 0, 1, tiles, bars, phase);
      }
    } else {
      int G = 1;
      while (G < 32 && U * G * 2 <= T) G <<= 1;
      const int nslots = T / G;
      for (int u0 = 0; u0 < U; u0 += nslots)
        emd_bid_pass<1>(w, xyz1, n, U, tpu, eps, gtid / G, nslots, u0, gtid % G, G, tiles, bars, phase);
    }
    cluster_sync_all();

    // ---- GetMax (:181-194): largest qualifying bidder wins the object ----
    for (int u = gtid; u < U; u += T) {
      const int j = w.unass[u];
      const int o = w.bid[j];
      const double bi = (double)w.bid_inc[j], mi = (double)w.max_inc[o];
      if (bi - 1e-6 <= mi && mi <= bi + 1e-6) atomicMax(&w.max_idx[o], j);
    }
    cluster_sync_all();

    // ---- Assign (:196-215) ----
    for (int u = gtid; u < U; u += T) {
      const int j = w.unass[u];
      const int o = w.bid[j];
      if (last || w.max_idx[o] == j) {
        const int inv = w.assignment_inv[o];
        if (!last && inv != -1) assignment[inv] = -1;
        w.assignment_inv[o] = j;
        assignment[j] = o;
        if (!last) {  // after the last round prices are dead state; the forced many-to-one writes would race
          float* pr = &w.obj[o].w;
          *pr = __fadd_rn(*pr, w.bid_inc[j]);
          w.max_inc[o] = -1e9f;
          w.max_idx[o] = -1;
        }
      }
    }
    fence_proxy_async_all();  // price updates must be visible to the next round's TMA tile loads
    cluster_sync_all();
  }

  // ---- CalcDist (:217-226) ----
  for (int k = gtid; k < n; k += T) {
    int a = assignment[k];
    a = a < 0 ? 0 : a;  // only reachable with iters == 0 (the reference reads out of bounds there)
    dist[k] = sqdist3(__fsub_rn(xyz1[k * 3 + 0], xyz2[a * 3 + 0]), __fsub_rn(xyz1[k * 3 + 1], xyz2[a * 3 + 1]),
                      __fsub_rn(xyz1[k * 3 + 2], xyz2[a * 3 + 2]));
  }
}

// grad_xyz1[j] = 2 g_j (x1_j - x2_assignment[j])   (emd_cuda.cu:284-300)
__global__ void __launch_bounds__(256) emd_grad_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2, int n, size_t total,
                                                        const float* __restrict__ g, const int* __restrict__ ass, float* __restrict__ gx) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t o = (i / n) * n + ass[i];
  const float gg = g[i] * 2.f;
  gx[i * 3 + 0] = gg * (xyz1[i * 3 + 0] - xyz2[o * 3 + 0]);
  gx[i * 3 + 1] = gg * (xyz1[i * 3 + 1] - xyz2[o * 3 + 1]);
  gx[i * 3 + 2] = gg * (xyz1[i * 3 + 2] - xyz2[o * 3 + 2]);
}

}  // namespace snb

using namespace snb;

SNB_API size_t snb_emd_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  return (size_t)B * emd_ws_per_sample(N);
}

SNB_API int snb_emd_fwd(const float* xyz1, const float* xyz2, int B, int N, float eps, int iters, float* dist, int* assignment,
                        void* workspace, size_t workspace_bytes, void* stream) {
  if (B < 0 || N < 0 || iters < 0) return SNB_EINVAL;
  if (B > 512 || (N % 1024) != 0) return SNB_ELIMIT;  // emd_cuda.cu:236-249
  if (B == 0 || N == 0) return SNB_OK;
  if (!workspace || workspace_bytes < snb_emd_workspace_bytes(B, N)) return SNB_EWORKSPACE;
  if (((uintptr_t)workspace & 15) != 0) return SNB_EALIGN;
  cudaStream_t s = (cudaStream_t)stream;
  // per device/context and cheap: set before every launch (a process-wide flag would leave the other GPUs of one process without it)
  SNB_CUDA(cudaFuncSetAttribute(emd_auction_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EMD_SMEM));
  int cs = 1;
  while (cs * 2 <= EMD_MAX_CLUSTER && B * cs * 2 <= kNumSMs) cs *= 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * cs));
  cfg.blockDim = dim3(EMD_THREADS);
  cfg.dynamicSmemBytes = EMD_SMEM;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  SNB_CUDA(cudaLaunchKernelEx(&cfg, emd_auction_kernel, xyz1, xyz2, N, eps, iters, dist, assignment, workspace));
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

SNB_API int snb_emd_bwd(const float* xyz1, const float* xyz2, int B, int N, const float* grad_dist, const int* assignment, float* grad_xyz1,
                        void* stream) {
  if (B < 0 || N < 0) return SNB_EINVAL;
  if (B == 0 || N == 0) return SNB_OK;
  const size_t total = (size_t)B * N;
  emd_grad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(xyz1, xyz2, N, total, grad_dist, assignment, grad_xyz1);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
