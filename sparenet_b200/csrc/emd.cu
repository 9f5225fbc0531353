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
#include "bvh.cuh"

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

// ---- the pruned auction (n <= 16384) ----------------------------------------------------------------------------------------
// Same rounds, same arithmetic, same results as emd_auction_kernel; what changes is how many (bidder, object) pairs a Bid pass
// looks at.  Both clouds are put in Morton order by the Chamfer search's builder (bvh.cuh): objects in leaves of 32 with
// axis-aligned boxes, 16 leaves per super-box.  Every round each CTA recomputes the lowest price of every leaf / super-box
// (prices only change in Assign).  A bidder's value for any object of a box is at most 3 - dist(bidder, box) - min price(box),
// so the box is skipped when the pair filter above, evaluated with the box distance (shrunk by 1e-5: never above the computed
// distance of a point inside) and the box's minimum price, would reject -- then it would reject every object of the box:
//     qq_box = max(c - pmin, 0) >= qq_k,  lb <= s_k   =>   fma(qq_box, qq_box, -lb) <= 0  implies  fma(qq_k, qq_k, -s_k) <= 0.
// best / second best are order statistics and exact ties are decided by the explicit tie key, so the visiting order is free:
// a warp takes up to 32 bidders that are neighbours in Morton order, starts in the leaf with the best bound for its first bidder
// and then walks super-boxes and leaves in index order, entering one when ANY lane needs it; inside a leaf all lanes read the
// same object (one broadcast load) and run the unchanged filter + exact path.  Work per round drops from U*n pairs to about
// U * (n/512 + 16*few + 32*few).  All state is kept in (Morton rank of the bidder, Morton position of the object) space; only the
// GetMax race rule ("largest bidder index wins") and the tie keys use original indices.
constexpr int EMDT_THREADS = 512;
constexpr int EMDT_MAXNC = BVH_MAXN / BVH_LEAF;              // 512 leaves
constexpr int EMDT_MAXNS = EMDT_MAXNC / BVH_FAN;             // 32 super-boxes: one lane each in the home search

struct EmdTreeWs {
  const float4* bq;     // [n] bidders in Morton order: x, y, z, original index bits
  float4* obj;          // [n] objects in Morton order: x, y, z, price
  const float4* box;    // [2*nc] object leaf boxes
  const float4* sbox;   // [2*ns]
  int* oid;             // [n] original index of the object at a position
  int* ass;             // [n] bidder rank -> object position (-1: unassigned)
  int* ass_inv;         // [n] object position -> bidder rank
  int* bid;             // [n] by bidder rank
  float* bid_inc;       // [n] by bidder rank
  float* max_inc;       // [n] by object position
  int* max_idx;         // [n] by object position: ORIGINAL bidder index (the reference's race rule)
  int* unass;           // [n] ranks of the unassigned bidders
  float* lp;            // [n/32] lowest price of every leaf (refreshed each round)
  int* counter;         // [2]
};

__host__ __device__ inline size_t emdt_ws_per_sample(int n) {
  return 2 * bvh_cloud_floats4(n) * sizeof(float4) + (size_t)n * 8 * 4 + (size_t)(n / BVH_LEAF) * 4 + 64;
}

__device__ __forceinline__ EmdTreeWs emdt_ws_view(void* ws, int b, int n) {
  char* p = (char*)ws + (size_t)b * emdt_ws_per_sample(n);
  float4* tree = (float4*)p;
  const BvhView Bq = bvh_view(tree, n), Ob = bvh_view(tree + bvh_cloud_floats4(n), n);
  EmdTreeWs w;
  w.bq = Bq.pts;
  w.obj = Ob.pts;
  w.box = Ob.box;
  w.sbox = Ob.sbox;
  p += 2 * bvh_cloud_floats4(n) * sizeof(float4);
  w.oid = (int*)p;          p += (size_t)n * 4;
  w.ass = (int*)p;          p += (size_t)n * 4;
  w.ass_inv = (int*)p;      p += (size_t)n * 4;
  w.bid = (int*)p;          p += (size_t)n * 4;
  w.bid_inc = (float*)p;    p += (size_t)n * 4;
  w.max_inc = (float*)p;    p += (size_t)n * 4;
  w.max_idx = (int*)p;      p += (size_t)n * 4;
  w.unass = (int*)p;        p += (size_t)n * 4;
  w.lp = (float*)p;         p += (size_t)(n / BVH_LEAF) * 4;
  w.counter = (int*)p;
  return w;
}

__device__ __forceinline__ unsigned long long emdt_tie_key(int pos, const int* oid, int n, int tpu) {
  return pos < 0 ? ~0ull : emd_tie_key(oid[pos], n, tpu);
}

// st.c only ever tightens: with several lanes per bidder it may already hold the threshold of the GROUP's second best
__device__ __forceinline__ void emdt_exact_update(BidState& st, float s, float price, int pos, const int* oid, int n, int tpu) {
  const float d = (float)((3.0 - (double)sqrtf(s)) - (double)price);
  if (d > st.best) {
    st.better = st.best;
    st.best = d;
    st.best_i = pos;
    st.c = fminf(st.c, emd_thr(st.better));
  } else if (d == st.best) {
    st.better = d;
    if (emdt_tie_key(pos, oid, n, tpu) < emdt_tie_key(st.best_i, oid, n, tpu)) st.best_i = pos;
    st.c = fminf(st.c, emd_thr(st.better));
  } else if (d > st.better) {
    st.better = d;
    st.c = fminf(st.c, emd_thr(st.better));
  }
}

__device__ __forceinline__ void emdt_merge(BidState& a, float obest, float obetter, int obi, const int* oid, int n, int tpu) {
  if (obest > a.best) {
    a.better = fmaxf(a.best, obetter);
    a.best = obest;
    a.best_i = obi;
  } else if (obest == a.best) {
    a.better = a.best;
    if (emdt_tie_key(obi, oid, n, tpu) < emdt_tie_key(a.best_i, oid, n, tpu)) a.best_i = obi;
  } else {
    a.better = fmaxf(a.better, obest);
  }
}

// true when the pair filter could flag an object at squared distance >= lb with price >= pmin
__device__ __forceinline__ bool emdt_need(float c, float pmin, float lb) {
  const float qq = fmaxf(c - pmin, 0.f);
  return __fmaf_rn(qq, qq, -lb) > 0.f;
}

#ifdef SNB_EMD_STATS  // development counters (tools/emd_stats.py builds a private library with them)
__device__ unsigned long long g_emd_stats[16];
#define EMD_STAT(i, v) do { if ((threadIdx.x & 31) == 0) atomicAdd(&g_emd_stats[i], (unsigned long long)(v)); } while (0)
#define EMD_CLOCK() clock64()
#else
#define EMD_STAT(i, v) do { } while (0)
#define EMD_CLOCK() 0ll
#endif

// The lanes of a warp scan one leaf staged in the warp's shared-memory buffer.  G lanes serve one bidder (32/G bidders per warp):
// lane g of a group takes objects g, g+G, ... (every group reads the same addresses: broadcast LDS.128).  Pass 1 is the pair
// filter for the lane's objects, straight-line, collecting a flag word.  In a leaf the search actually opens, flagged pairs are
// COMMON (these are the objects near the top of the bidder's list), so the exact path must not sit behind a per-object
// warp-divergent branch: pass 2 is a converged loop in which every lane takes ITS next flagged object, re-checks it against its
// threshold (tightened by the previous exact evaluations) and evaluates it exactly -- trip count = the largest flag count of a
// lane.  With G > 1 a lane's own second best is a weak threshold, so the group's (best, second best) over all its lanes is merged
// by a butterfly after the leaf and every lane keeps the tighter threshold.
template <int G>
__device__ __forceinline__ void emdt_scan_leaf(const float4* buf, int pos0, int g, float bx, float by, float bz, BidState& st, const int* oid,
                                               int n, int tpu) {
  {
    unsigned fl = 0u;
#pragma unroll
    for (int i = 0; i < BVH_LEAF / G; i++) {
      const float4 o = buf[g + G * i];
      const float sv = sqdist3(__fsub_rn(o.x, bx), __fsub_rn(o.y, by), __fsub_rn(o.z, bz));
      const float qq = fmaxf(st.c - o.w, 0.f);
      fl |= (__fmaf_rn(qq, qq, -sv) > 0.f) ? (1u << i) : 0u;
    }
    while (__any_sync(0xffffffffu, fl != 0u)) {
      if (fl) {
        const int t = g + G * (__ffs(fl) - 1);
        fl &= fl - 1u;
        const float4 o = buf[t];
        const float sv = sqdist3(__fsub_rn(o.x, bx), __fsub_rn(o.y, by), __fsub_rn(o.z, bz));
        const float qq = fmaxf(st.c - o.w, 0.f);
        if (__fmaf_rn(qq, qq, -sv) > 0.f) emdt_exact_update(st, sv, o.w, pos0 + t, oid, n, tpu);
      }
    }
  }
  if (G > 1) {
    float b1 = st.best, b2 = st.better;
#pragma unroll
    for (int off = 1; off < G; off <<= 1) {
      const float o1 = __shfl_xor_sync(0xffffffffu, b1, off), o2 = __shfl_xor_sync(0xffffffffu, b2, off);
      b2 = fmaxf(fminf(b1, o1), fmaxf(b2, o2));   // second largest of the union of two disjoint multisets
      b1 = fmaxf(b1, o1);
    }
    st.c = fminf(st.c, emd_thr(b2));
  }
}

struct EmdtShared {
  float4 box[2 * EMDT_MAXNC];
  float4 sbox[2 * EMDT_MAXNS];
  float4 stage[EMDT_THREADS / 32][2][BVH_LEAF];   // per warp: the leaf being scanned + the next one
  float lp[EMDT_MAXNC];   // lowest price of a leaf
  float sp[EMDT_MAXNS];   // ... of a super-box
};

// One Bid pass of a warp: bidders unass[u0 .. u0 + 32/G), G lanes each.
template <int G>
__device__ __forceinline__ void emdt_bid_pass(const EmdTreeWs& w, EmdtShared& sh, int n, int nc, int ns, int U, int u0, int tpu, float eps) {
  constexpr int PW = 32 / G;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slot = lane / G, g = lane % G;
  const int u = u0 + slot;
  const bool act = u < U;
  const int r = w.unass[act ? u : u0];   // idle groups shadow the first bidder: they never widen the set of visited boxes
  const float4 bb = w.bq[r];
  BidState st;
  st.best = -1e9f;
  st.better = -1e9f;
  st.best_i = -1;
  st.c = emd_thr(-1e9f);
  const long long tp0 = EMD_CLOCK();
  // ---- every bidder's own most promising leaf (best bound 3 - dist - min price), the super-boxes / leaves dealt to the lanes of
  //      its group; the distinct ones are scanned first so that every bidder of the pass starts with a tight threshold ----
  int hl;
  {
    unsigned long long key = ~0ull;
    for (int s = g; s < ns; s += G) {
      const float v = sqrtf(box_lb(sh.sbox[2 * s], sh.sbox[2 * s + 1], bb.x, bb.y, bb.z)) + sh.sp[s];
      const unsigned long long k2 = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned)s;   // v >= 0: bit order == value order
      key = k2 < key ? k2 : key;
    }
#pragma unroll
    for (int off = 1; off < G; off <<= 1) {
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, key, off);
      key = o < key ? o : key;
    }
    const int bs = (int)(unsigned)key;
    key = ~0ull;
    const int c1 = (bs + 1) * BVH_FAN < nc ? (bs + 1) * BVH_FAN : nc;
    for (int c = bs * BVH_FAN + g; c < c1; c += G) {
      const float v = sqrtf(box_lb(sh.box[2 * c], sh.box[2 * c + 1], bb.x, bb.y, bb.z)) + sh.lp[c];
      const unsigned long long k2 = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned)c;
      key = k2 < key ? k2 : key;
    }
#pragma unroll
    for (int off = 1; off < G; off <<= 1) {
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, key, off);
      key = o < key ? o : key;
    }
    hl = (int)(unsigned)key;
  }
  unsigned hdone = 0u;   // lane s: leaves of super-box s already scanned as somebody's home
  int bufi = 0;
  {
    unsigned pend = 0xffffffffu;
    int c = __shfl_sync(0xffffffffu, hl, 0);
    pend &= ~__ballot_sync(0xffffffffu, hl == c);
    float4 pf = w.obj[c * BVH_LEAF + lane];
    while (c >= 0) {
      int cn = -1;
      if (pend) {
        cn = __shfl_sync(0xffffffffu, hl, __ffs(pend) - 1);
        pend &= ~__ballot_sync(0xffffffffu, hl == cn);
      }
      sh.stage[warp][bufi][lane] = pf;
      __syncwarp();
      if (cn >= 0) pf = w.obj[cn * BVH_LEAF + lane];
      EMD_STAT(9, 1);
      emdt_scan_leaf<G>(sh.stage[warp][bufi], c * BVH_LEAF, g, bb.x, bb.y, bb.z, st, w.oid, n, tpu);
      if (lane == c / BVH_FAN) hdone |= 1u << (c % BVH_FAN);
      bufi ^= 1;
      c = cn;
    }
  }
  // ---- which other leaves can still matter: lane s ends up with the 16-bit leaf mask of super-box s (a later, tighter
  //      threshold can only drop leaves, and every leaf is tested again right before it is scanned) ----
  unsigned mym = 0u;
#pragma unroll 1
  for (int s = 0; s < ns; s++) {
    const bool need_s = emdt_need(st.c, sh.sp[s], box_lb(sh.sbox[2 * s], sh.sbox[2 * s + 1], bb.x, bb.y, bb.z));
    if (!__any_sync(0xffffffffu, need_s)) continue;
    EMD_STAT(3, 1);
    unsigned m16 = 0u;
    const int c0 = s * BVH_FAN;
#pragma unroll 4
    for (int q = 0; q < BVH_FAN; q++) {
      const int c = c0 + q;
      const bool need_c = c < nc && emdt_need(st.c, sh.lp[c], box_lb(sh.box[2 * c], sh.box[2 * c + 1], bb.x, bb.y, bb.z));
      m16 |= __any_sync(0xffffffffu, need_c) ? (1u << q) : 0u;
    }
    if (lane == s) mym = m16 & ~hdone;
  }
  // ---- walk the marked leaves in index order; the next one is already in flight while this one is scanned ----
  {
    unsigned smask = __ballot_sync(0xffffffffu, mym != 0u);
    int cur_s = 0;
    unsigned cur_m = 0u;
    auto next_leaf = [&]() -> int {
      while (cur_m == 0u) {
        if (smask == 0u) return -1;
        cur_s = __ffs(smask) - 1;
        smask &= smask - 1u;
        cur_m = __shfl_sync(0xffffffffu, mym, cur_s);
      }
      const int q = __ffs(cur_m) - 1;
      cur_m &= cur_m - 1u;
      return cur_s * BVH_FAN + q;
    };
    int c = next_leaf();
    float4 pf = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c >= 0) pf = w.obj[c * BVH_LEAF + lane];
    while (c >= 0) {
      const int cn = next_leaf();
      sh.stage[warp][bufi][lane] = pf;
      __syncwarp();
      if (cn >= 0) pf = w.obj[cn * BVH_LEAF + lane];
      const bool need_c = emdt_need(st.c, sh.lp[c], box_lb(sh.box[2 * c], sh.box[2 * c + 1], bb.x, bb.y, bb.z));
      const unsigned nm = __ballot_sync(0xffffffffu, need_c);
      EMD_STAT(8, 1);
      if (nm) {
        EMD_STAT(4, 1);
        EMD_STAT(5, __popc(nm));
        emdt_scan_leaf<G>(sh.stage[warp][bufi], c * BVH_LEAF, g, bb.x, bb.y, bb.z, st, w.oid, n, tpu);
      }
      bufi ^= 1;
      c = cn;
    }
  }
  EMD_STAT(12, EMD_CLOCK() - tp0);
  __syncwarp();
  EMD_STAT(2, 1);
  // ---- merge the G partial states of each bidder (xor butterfly inside the aligned lane group) ----
#pragma unroll
  for (int off = 1; off < G; off <<= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, st.best, off);
    const float obb = __shfl_xor_sync(0xffffffffu, st.better, off);
    const int oi = __shfl_xor_sync(0xffffffffu, st.best_i, off);
    emdt_merge(st, ob, obb, oi, w.oid, n, tpu);
  }
  if (act && g == 0) {
    const float inc = __fadd_rn(__fsub_rn(st.best, st.better), eps);
    w.bid[r] = st.best_i;
    w.bid_inc[r] = inc;
    atomic_max_float(&w.max_inc[st.best_i], inc);
  }
  (void)PW;
}

__global__ void __launch_bounds__(EMDT_THREADS, 1) emd_auction_tree_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2, int n,
                                                                            float eps, int iters, float* __restrict__ dist,
                                                                            int* __restrict__ assignment, void* workspace) {
  __shared__ EmdtShared sh;
  const uint32_t cs = cluster_nctarank(), rank = cluster_ctarank();
  const int b = blockIdx.x / cs;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = cs * EMDT_THREADS;
  const int gtid = rank * EMDT_THREADS + tid;
  const int Wn = T / 32, gw = gtid >> 5;
  xyz1 += (size_t)b * n * 3;
  xyz2 += (size_t)b * n * 3;
  dist += (size_t)b * n;
  assignment += (size_t)b * n;
  const EmdTreeWs w = emdt_ws_view(workspace, b, n);
  const int nc = n / BVH_LEAF, ns = (nc + BVH_FAN - 1) / BVH_FAN;   // n % 1024 == 0: no padded leaves, every super-box is full
  const int R = ((n + Wn - 1) / Wn + 31) & ~31;                     // ranks per warp in the compaction (a multiple of 32)

  for (int k = gtid; k < n; k += T) {
    float4 p = w.obj[k];
    w.oid[k] = __float_as_int(p.w);
    p.w = 0.f;   // price (emd_module.py:43-54)
    w.obj[k] = p;
    w.ass[k] = -1;
    w.ass_inv[k] = -1;
    w.max_inc[k] = 0.f;
    w.max_idx[k] = -1;
  }
  for (int i = tid; i < 2 * nc; i += EMDT_THREADS) sh.box[i] = w.box[i];
  for (int i = tid; i < 2 * ns; i += EMDT_THREADS) sh.sbox[i] = w.sbox[i];
  if (gtid == 0) {
    w.counter[0] = 0;
    w.counter[1] = 0;
    w.counter[2] = 0;   // [2], [3]: the Bid pass tickets of even / odd rounds
    w.counter[3] = 0;
  }
  __syncthreads();
  cluster_sync_all();

  const int block_cnt = n / 1024;
  for (int it = 0; it < iters; it++) {
    const long long tr0 = EMD_CLOCK();
    int* cnt_cur = &w.counter[it & 1];
    // ---- lowest price per leaf, the leaves dealt to the warps of the cluster (prices are final since the barrier that ended
    //      the last round) ----
    for (int c = gw; c < nc; c += Wn) {
      float p = w.obj[c * BVH_LEAF + lane].w;
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) p = fminf(p, __shfl_xor_sync(0xffffffffu, p, o));
      if (lane == 0) w.lp[c] = p;
    }
    // ---- compact the unassigned bidders: a warp owns R consecutive Morton ranks and appends its survivors with ONE atomic, so
    //      chunks of the list stay spatial neighbours ----
    {
      const int k0 = gw * R, k1 = (k0 + R) < n ? (k0 + R) : n;
      int cnt = 0;
      for (int k = k0 + lane; k < k1; k += 32) cnt += (w.ass[k] == -1) ? 1 : 0;   // n % 32 == 0: whole warps
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      if (cnt) {
        int base = 0;
        if (lane == 0) base = atomicAdd(cnt_cur, cnt);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int k = k0 + lane; k < k1; k += 32) {
          const bool un = w.ass[k] == -1;
          const unsigned m = __ballot_sync(0xffffffffu, un);
          if (un) w.unass[base + __popc(m & ((1u << lane) - 1u))] = k;
          base += __popc(m);
        }
      }
    }
    const long long tc1 = EMD_CLOCK();
    cluster_sync_all();
    const int U = *((volatile int*)cnt_cur);
    if (U == 0) break;
    for (int c = tid; c < nc; c += EMDT_THREADS) sh.lp[c] = w.lp[c];
    __syncthreads();
    if (tid < ns) {
      float p = sh.lp[tid * BVH_FAN];
      for (int c = 1; c < BVH_FAN && tid * BVH_FAN + c < nc; c++) p = fminf(p, sh.lp[tid * BVH_FAN + c]);
      sh.sp[tid] = p;
    }
    __syncthreads();
    if (gtid == 0) {
      w.counter[(it + 1) & 1] = 0;
      w.counter[2 + ((it + 1) & 1)] = 0;
    }
    const bool last = (it == iters - 1);
    const int unass_per_block = (U + block_cnt - 1) / block_cnt;
    const int tpu = 1024 / unass_per_block;  // the reference's thread_per_unass, needed only for tie keys
    const long long tb0 = EMD_CLOCK();

    // ---- Bid.  Dense rounds (a quarter of the bidders or more still unassigned): 32 neighbouring bidders per warp pass, one lane
    //      each.  Sparse rounds: the survivors are far apart in Morton order, the union of the boxes 32 of them open is several
    //      times what each needs, and the round lasts as long as its longest pass: 8 bidders with 4 lanes each, or -- when there
    //      are fewer than 4 bidders per warp of the cluster -- 2 bidders with 16 lanes each.  Passes differ a lot in length: the
    //      warps of the cluster draw them from a ticket counter instead of a fixed deal. ----
    {
      int* ticket = &w.counter[2 + (it & 1)];
      const int mode = ((long long)U * 4 >= n) ? 0 : (U > 4 * Wn ? 1 : 2);
      const int pw = mode == 0 ? 32 : (mode == 1 ? 8 : 2);
      for (;;) {
        int p = 0;
        if (lane == 0) p = atomicAdd(ticket, 1);
        p = __shfl_sync(0xffffffffu, p, 0);
        if ((long long)p * pw >= U) break;
        if (mode == 0) emdt_bid_pass<1>(w, sh, n, nc, ns, U, p * pw, tpu, eps);
        else if (mode == 1) emdt_bid_pass<4>(w, sh, n, nc, ns, U, p * pw, tpu, eps);
        else emdt_bid_pass<16>(w, sh, n, nc, ns, U, p * pw, tpu, eps);
      }
    }
    const long long tb1 = EMD_CLOCK();
    cluster_sync_all();
    const long long tg0 = EMD_CLOCK();

    // ---- GetMax (:181-194): largest qualifying ORIGINAL bidder index wins the object ----
    for (int u = gtid; u < U; u += T) {
      const int r = w.unass[u];
      const int o = w.bid[r];
      const double bi = (double)w.bid_inc[r], mi = (double)w.max_inc[o];
      if (bi - 1e-6 <= mi && mi <= bi + 1e-6) atomicMax(&w.max_idx[o], __float_as_int(w.bq[r].w));
    }
    const long long tg1 = EMD_CLOCK();
    cluster_sync_all();
    const long long ta0 = EMD_CLOCK();

    // ---- Assign (:196-215) ----
    for (int u = gtid; u < U; u += T) {
      const int r = w.unass[u];
      const int o = w.bid[r];
      if (last || w.max_idx[o] == __float_as_int(w.bq[r].w)) {
        const int inv = w.ass_inv[o];
        if (!last && inv != -1) w.ass[inv] = -1;
        w.ass_inv[o] = r;
        w.ass[r] = o;
        if (!last) {  // after the last round prices are dead state; the forced many-to-one writes would race
          float* pr = &w.obj[o].w;
          *pr = __fadd_rn(*pr, w.bid_inc[r]);
          w.max_inc[o] = -1e9f;
          w.max_idx[o] = -1;
        }
      }
    }
    cluster_sync_all();
#ifdef SNB_EMD_STATS
    if (gtid == 0 && b == 0) {
      atomicAdd(&g_emd_stats[0], 1ull);
      atomicAdd(&g_emd_stats[1], (unsigned long long)U);
      atomicAdd(&g_emd_stats[6], (unsigned long long)(tb1 - tb0));
      atomicAdd(&g_emd_stats[7], (unsigned long long)(EMD_CLOCK() - tr0));
      atomicAdd(&g_emd_stats[10], (unsigned long long)(tc1 - tr0));
      atomicAdd(&g_emd_stats[11], (unsigned long long)(tg1 - tg0));
      atomicAdd(&g_emd_stats[13], (unsigned long long)(EMD_CLOCK() - ta0));
      atomicAdd(&g_emd_stats[14], (unsigned long long)(tg0 - tb1));
    }
#endif
  }

  // ---- back to original indices + CalcDist (:217-226) ----
  for (int r = gtid; r < n; r += T) {
    const int j = __float_as_int(w.bq[r].w);
    const int pos = w.ass[r];
    const int a = pos < 0 ? -1 : w.oid[pos];
    assignment[j] = a;
    const int aa = a < 0 ? 0 : a;  // only reachable with iters == 0 (the reference reads out of bounds there)
    dist[j] = sqdist3(__fsub_rn(xyz1[j * 3 + 0], xyz2[aa * 3 + 0]), __fsub_rn(xyz1[j * 3 + 1], xyz2[aa * 3 + 1]),
                      __fsub_rn(xyz1[j * 3 + 2], xyz2[aa * 3 + 2]));
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
  size_t per = emd_ws_per_sample(N);
  if (N <= BVH_MAXN && emdt_ws_per_sample(N) > per) per = emdt_ws_per_sample(N);
  return (size_t)B * per;
}

static int emd_fwd_impl(bool allow_tree, const float* xyz1, const float* xyz2, int B, int N, float eps, int iters, float* dist, int* assignment,
                        void* workspace, size_t workspace_bytes, void* stream) {
  if (B < 0 || N < 0 || iters < 0) return SNB_EINVAL;
  if (B > 512 || (N % 1024) != 0) return SNB_ELIMIT;  // emd_cuda.cu:236-249
  if (B == 0 || N == 0) return SNB_OK;
  if (!workspace || workspace_bytes < snb_emd_workspace_bytes(B, N)) return SNB_EWORKSPACE;
  if (((uintptr_t)workspace & 15) != 0) return SNB_EALIGN;
  cudaStream_t s = (cudaStream_t)stream;
  int cs = 1;
  while (cs * 2 <= EMD_MAX_CLUSTER && B * cs * 2 <= kNumSMs) cs *= 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * cs));
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (N <= BVH_MAXN && allow_tree) {
    // bidders = xyz1 (cloud 0 of the hierarchy), objects = xyz2; the auction kernel turns the objects' index lane into the price
    const int rc = bvh_build_launch(xyz1, xyz2, B, N, N, (float4*)workspace, emdt_ws_per_sample(N) / sizeof(float4), s);
    if (rc != SNB_OK) return rc;
    cfg.blockDim = dim3(EMDT_THREADS);
    cfg.dynamicSmemBytes = 0;
    SNB_CUDA(cudaLaunchKernelEx(&cfg, emd_auction_tree_kernel, xyz1, xyz2, N, eps, iters, dist, assignment, workspace));
    SNB_LAUNCH_CHECK();
    return SNB_OK;
  }
  // per device/context and cheap: set before every launch (a process-wide flag would leave the other GPUs of one process without it)
  SNB_CUDA(cudaFuncSetAttribute(emd_auction_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EMD_SMEM));
  cfg.blockDim = dim3(EMD_THREADS);
  cfg.dynamicSmemBytes = EMD_SMEM;
  SNB_CUDA(cudaLaunchKernelEx(&cfg, emd_auction_kernel, xyz1, xyz2, N, eps, iters, dist, assignment, workspace));
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

SNB_API int snb_emd_fwd(const float* xyz1, const float* xyz2, int B, int N, float eps, int iters, float* dist, int* assignment,
                        void* workspace, size_t workspace_bytes, void* stream) {
  return emd_fwd_impl(true, xyz1, xyz2, B, N, eps, iters, dist, assignment, workspace, workspace_bytes, stream);
}

SNB_API int snb_emd_fwd_scan(const float* xyz1, const float* xyz2, int B, int N, float eps, int iters, float* dist, int* assignment,
                             void* workspace, size_t workspace_bytes, void* stream) {
  return emd_fwd_impl(false, xyz1, xyz2, B, N, eps, iters, dist, assignment, workspace, workspace_bytes, stream);
}

#ifdef SNB_EMD_STATS
extern "C" __attribute__((visibility("default"))) int snb_emd_debug_stats(unsigned long long* out, int reset) {
  cudaMemcpyFromSymbol(out, g_emd_stats, sizeof(g_emd_stats));
  if (reset) {
    unsigned long long z[16] = {};
    cudaMemcpyToSymbol(g_emd_stats, z, sizeof(z));
  }
  return 0;
}
#endif

SNB_API int snb_emd_bwd(const float* xyz1, const float* xyz2, int B, int N, const float* grad_dist, const int* assignment, float* grad_xyz1,
                        void* stream) {
  if (B < 0 || N < 0) return SNB_EINVAL;
  if (B == 0 || N == 0) return SNB_OK;
  const size_t total = (size_t)B * N;
  emd_grad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(xyz1, xyz2, N, total, grad_dist, assignment, grad_xyz1);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
