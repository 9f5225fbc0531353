// mds.cu -- minimum-density sampling + gather_points (fwd/bwd), sm_100a.
//
// Replaces minimum_density_sampling_kernel / gather_points[_grad]_kernel (cuda/MDS/MDS_cuda.cu:29-211).
// Contract: t = (float)(5.0*mml*mml); idx[0] = 0; each of the m-1 dependent rounds adds
// w = expf(-d/t) (d = fma(dz,dz,fma(dx,dx,dy*dy)), d* = x_k - x_old; IEEE divide, accurate expf; doubled for
// k >= 8192) to every point's accumulated density and picks the arg-min.  Ties follow the reference's reduction
// exactly: inside a thread the first k of its stride (k = tid, tid+bs, ...), across threads the smem tournament
// (MDS_cuda.cu:81-87,139-198: slot t absorbs slot t+s for s = bs/2..1, lower slot wins ties), whose winner among
// equals is the thread with the smallest BIT-REVERSED tid.  So the tie key is (bitrev(k % bs), k),
// bs = min(1024, 2^floor(log2 n)) -- verified against the reference extension on the GPU.  Chosen points park at 1e9.
//   The reference accumulates as (float)((double)temp + (double)w): for one addition of two floats, double
//   rounding through fp64 is innocuous (53 >= 2*24+2), so a plain fp32 add is bit-identical.
//
// Design.  The m-1 picks are a dependent chain, but only locally: a pick changes the densities of its neighbourhood
// (for 90 % of the live points the added weight is below half an ulp of their density) and the NEXT pick is almost always
// somewhere else.  So the chain is cut into GENERATIONS that are exact, not speculative:
//   1. every warp publishes its MDS_M lowest (density, tie key) candidates -- with coordinates -- and its (MDS_M+1)-th lowest
//      pair as a bound; all candidates of the cluster form the POOL (<= 128 warps x 4), theta = min over warps of the bounds.
//      Every point outside the pool is >= theta, and densities only ever grow;
//   2. ONE warp per CTA replays the sequential algorithm on the pool alone, in registers, with the very same arithmetic:
//      update the pool with the last pick's weights, take the arg-min, accept it while (density, key) < theta -- an accepted
//      pick is the global arg-min of the sequential algorithm (nothing outside the pool can have dropped below theta).  The
//      first candidate is always accepted, typically ~18 are (measured on the bench clouds: 932 generations for 16383 picks);
//   3. all warps apply the accepted picks, in order, to their own points (the same sequence of fp32 additions as the
//      reference's rounds), park the picked points, and select again.
// One cluster exchange (st.async + mbarrier complete_tx, no barrier.cluster) per generation instead of one per pick; the
// per-pick chain shrinks to ~130 instructions of one warp.  A thread-block CLUSTER (up to 8 CTAs) owns one sample; every point
// (xyz + density) lives in registers for the whole kernel; the reference does 11 block barriers and a global read-modify-write
// of `temp` per pick.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "common.cuh"

namespace snb {

constexpr int MDS_MAX_WARPS = 16;  // warps per CTA: 4 (128 threads, the default), 8 or 16
constexpr int MDS_MAX_CLUSTER = 8;
constexpr int MDS_M = 4;           // candidates each warp contributes to a generation's pool
constexpr int MDS_MAXK = 256;      // picks per generation (capacity of the pick list)
constexpr unsigned long long MDS_NONE = 0xffffffffffffffffull;
constexpr unsigned MDS_PARKED = 0x4e6e6b28u;  // bits of 1e9f: parked / padding entries compare >= this

// st.async writes the payload into the peer CTA's shared memory AND completes the same number of tx-bytes on the
// peer's mbarrier, so data and "it arrived" are one instruction; nobody executes barrier.cluster inside the loop.
__device__ __forceinline__ void st_async_b64(uint32_t remote_addr, unsigned long long v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote_addr), "l"(v), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void st_async_v4f32(uint32_t remote_addr, float a, float b, float c, float d, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1,%2,%3,%4}, [%5];" ::"r"(remote_addr), "f"(a),
               "f"(b), "f"(c), "f"(d), "r"(remote_bar)
               : "memory");
}

constexpr int MDS_SLOTS = MDS_MAX_CLUSTER * MDS_MAX_WARPS;  // up to 128 warps per sample

// The update is issue bound, so it is written to the bone:
//   * -d/t with the loop-invariant divisor t becomes Markstein's 3-instruction correctly-rounded division
//     (q = RN(d*r), rem = fma(-q,t,d) exact, q' = fma(rem,r,q) with r = RN(1/t)); the IEEE result is identical to
//     div.rn whenever t's significand is not all ones and the quotient is a normal number -- outside the normal
//     range expf(q') is exactly 1 or 0 either way; an all-ones significand falls back to div.rn (FAST_DIV=false).
//   * the x2 weight of points k >= 8192 is a per-slot register factor folded into one FMA (2w is exact),
//   * tie keys are per-slot registers, candidates are compared as packed u64 (density bits, key),
//   * LIVE-POINT COMPACTION: a chosen point is parked at 1e9 for good and m/n of the points end up chosen (89 % in
//     SpareNet's refiner), so the register layout is re-packed through shared memory whenever the CTA's live points
//     fit a narrower unrolled loop (PT 18 -> 14 -> 10 -> 7 -> 4 -> 2): on average ~56 % of the points are still
//     updated per pick.  Every live point sees exactly the same sequence of fp32 additions as before, so the sampled
//     indices are unchanged; dropped points could never be chosen again (the reference adds w to their 1e9 for nothing).
template <bool FAST_DIV>
__device__ __forceinline__ float mds_add(float temp, float fac, float x, float y, float z, float x1, float y1, float z1, float t, float r) {
  const float d = sqdist3(__fsub_rn(x, x1), __fsub_rn(y, y1), __fsub_rn(z, z1));
  float q;
  if (FAST_DIV) {
    const float q0 = __fmul_rn(d, r);
    q = __fmaf_rn(__fmaf_rn(-q0, t, d), r, q0);  // == div.rn(d, t) (see above)
  } else {
    q = __fdiv_rn(d, t);
  }
  return __fmaf_rn(expf(-q), fac, temp);  // temp + w or temp + 2w (2w exact): one rounding, as the reference
}
__device__ __forceinline__ unsigned long long mds_pack(float v, unsigned key) {
  return ((unsigned long long)__float_as_uint(v) << 32) | key;  // densities are >= 0: u64 order == (density, tie key)
}

struct MdsStage {          // staging area in dynamic shared memory, capacity = THREADS * PT0 entries
  float* t;                // running density of the entry's point (2e9 = padding)
  unsigned* k;             // tie key << 21 | point index (~0 = padding); coordinates and the x2 factor follow from the index
  unsigned short* loc;     // point (k - kbeg) -> its entry in the current layout
  int* count;
};

struct MdsShared {         // static shared memory of one CTA
  unsigned long long pack[2][MDS_SLOTS * MDS_M];   // pool candidates (density bits, key), double buffered by generation parity
  float4 xyz[2][MDS_SLOTS * MDS_M];                // ... their coordinates
  unsigned long long theta[2][MDS_SLOTS];          // per-warp bounds: the warp's (MDS_M+1)-th lowest pair
  float4 picks[MDS_MAXK];                          // accepted picks of the last generation: x, y, z, bits(index)
  uint64_t bars[2];
  int npicks;
};

struct MdsCtx {
  const float* dataset;
  int* idxs;
  const float* sxyz;   // this CTA's points (shared memory copy when it fits)
  MdsShared* sh;
  MdsStage st;
  int m, kbeg, kend;
  uint32_t cs, rank;
  float t, r;
};

__host__ __device__ constexpr int mds_next_pt(int pt) { return pt > 12 ? pt - 4 : (pt > 6 ? pt - 3 : (pt > 2 ? pt - 2 : 0)); }

template <int MDS_THREADS, int PT, bool FAST_DIV>
struct MdsLevel {
  // runs generations while the CTA still holds more live points than the next narrower layout can take; returns the next j
  static __device__ __forceinline__ int run(int j, const MdsCtx& c, int& live, int& gen) {
    constexpr int MDS_WARPS = MDS_THREADS / 32;
    constexpr int NQ = (MDS_MAX_CLUSTER * MDS_WARPS * MDS_M + 31) / 32;  // pool entries per lane of the replaying warp
    constexpr int NEXT = mds_next_pt(PT);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    MdsShared& sh = *c.sh;
    const float t = c.t, r = c.r;
    float x[PT], y[PT], z[PT], temp[PT], fac[PT];
    unsigned key[PT];
#pragma unroll
    for (int i = 0; i < PT; i++) {  // entry e = tid + i*THREADS of the staged layout
      const int e = tid + i * MDS_THREADS;
      temp[i] = c.st.t[e];
      key[i] = c.st.k[e];
      const int k = (int)(key[i] & 0x1fffffu);
      const int kk = key[i] != 0xffffffffu ? k - c.kbeg : 0;
      x[i] = c.sxyz[kk * 3 + 0];
      y[i] = c.sxyz[kk * 3 + 1];
      z[i] = c.sxyz[kk * 3 + 2];
      fac[i] = k < 8192 ? 1.0f : 2.0f;  // MDS_cuda.cu:111-112 (k > 8191 counts double)
    }
    const uint32_t cs = c.cs;
    const uint32_t my_slot = c.rank * MDS_WARPS + warp;
    const int total = cs * MDS_WARPS;
    // peer addresses for parity 0; parity 1 is a constant offset further (a CTA's shared::cluster window is contiguous)
    const uint32_t dst = lane < (int)cs ? (uint32_t)lane : 0u;
    const uint32_t r_pack0 = mapa_shared(smem_u32(&sh.pack[0][my_slot * MDS_M]), dst);
    const uint32_t r_xyz0 = mapa_shared(smem_u32(&sh.xyz[0][my_slot * MDS_M]), dst);
    const uint32_t r_theta0 = mapa_shared(smem_u32(&sh.theta[0][my_slot]), dst);
    const uint32_t r_bar0 = mapa_shared(smem_u32(&sh.bars[0]), dst);
    const uint32_t gen_bytes = (uint32_t)total * (MDS_M * 24u + 8u);

    for (;;) {
      // ---- apply the picks of the previous generation, in order, to this thread's points ----------------------------
      const int np = sh.npicks;
#pragma unroll 1
      for (int p = 0; p < np; p++) {
        const float4 pk = sh.picks[p];  // broadcast read
        const int pidx = __float_as_int(pk.w);
        if (pidx >= c.kbeg && pidx < c.kend) {  // park it: every warp keeps the live count, only the owner touches its registers
          live--;
          const int e = c.st.loc[pidx - c.kbeg];
          if ((e % MDS_THREADS) == tid) {
            const int slot = e / MDS_THREADS;
#pragma unroll
            for (int i = 0; i < PT; i++)
              if (i == slot) temp[i] = 1e9f;   // 1e9f + w == 1e9f for every later w <= 2
          }
        }
#pragma unroll
        for (int i = 0; i < PT; i++) temp[i] = mds_add<FAST_DIV>(temp[i], fac[i], x[i], y[i], z[i], pk.x, pk.y, pk.z, t, r);
      }
      if (NEXT > 0 && live <= NEXT * MDS_THREADS) break;  // re-pack into the narrower layout (uniform over the CTA)

      // ---- this warp's MDS_M lowest (density, key) pairs and the next one as its bound -------------------------------
      unsigned long long mine = MDS_NONE;
#pragma unroll
      for (int i = 0; i < PT; i++) {
        const unsigned long long p = mds_pack(temp[i], key[i]);
        mine = p < mine ? p : mine;
      }
      unsigned long long sel[MDS_M + 1];
#pragma unroll
      for (int s = 0; s <= MDS_M; s++) {
        unsigned long long w = warp_min_u64(mine);
        if ((unsigned)(w >> 32) >= MDS_PARKED) w = MDS_NONE;  // parked / padding: this warp has run out of live points
        sel[s] = w;
        if (s < MDS_M && mine == w && w != MDS_NONE) {         // the owning lane (keys are unique) moves on to its next entry
          unsigned long long nx = MDS_NONE;
#pragma unroll
          for (int i = 0; i < PT; i++) {
            const unsigned long long p = mds_pack(temp[i], key[i]);
            nx = (p > w && p < nx) ? p : nx;
          }
          mine = nx;
        }
      }
      // ---- publish them to every CTA of the cluster ------------------------------------------------------------------
      const int par = gen & 1;
      if (tid == 0) mbar_expect_tx(&sh.bars[par], gen_bytes);  // arm this generation's phase (the single expected arrival)
      if (lane < (int)cs) {
        const uint32_t r_bar = r_bar0 + par * (uint32_t)sizeof(uint64_t);
        const uint32_t r_pack = r_pack0 + par * (uint32_t)sizeof(sh.pack[0]);
        const uint32_t r_xyz = r_xyz0 + par * (uint32_t)sizeof(sh.xyz[0]);
#pragma unroll
        for (int s = 0; s < MDS_M; s++) {
          const int kc = (int)((unsigned)sel[s] & 0x1fffffu) - c.kbeg;  // a candidate of this warp is one of this CTA's points
          const int kk = (sel[s] != MDS_NONE && kc >= 0 && kc < c.kend - c.kbeg) ? kc : 0;
          st_async_b64(r_pack + s * 8u, sel[s], r_bar);
          st_async_v4f32(r_xyz + s * 16u, c.sxyz[kk * 3 + 0], c.sxyz[kk * 3 + 1], c.sxyz[kk * 3 + 2], 0.f, r_bar);
        }
        st_async_b64(r_theta0 + par * (uint32_t)sizeof(sh.theta[0]), sel[MDS_M], r_bar);
      }
      mbar_wait_tx(&sh.bars[par], (uint32_t)(gen >> 1) & 1u);  // k-th use of bars[par] (generations par, par+2, ...) has parity k & 1
      gen++;

      // ---- warp 0 replays the sequential algorithm on the pool --------------------------------------------------------
      if (warp == 0) {
        float px[NQ], py[NQ], pz[NQ], pt[NQ], pf[NQ];
        unsigned pkey[NQ];
        const int nent = total * MDS_M;
#pragma unroll
        for (int q = 0; q < NQ; q++) {
          const int e = lane + 32 * q;
          const unsigned long long p = e < nent ? sh.pack[par][e] : MDS_NONE;
          const float4 cx = sh.xyz[par][e < nent ? e : 0];
          const bool ok = (unsigned)(p >> 32) < MDS_PARKED;
          pt[q] = ok ? __uint_as_float((unsigned)(p >> 32)) : 2e9f;
          pkey[q] = ok ? (unsigned)p : 0xffffffffu;
          px[q] = cx.x;
          py[q] = cx.y;
          pz[q] = cx.z;
          pf[q] = (pkey[q] & 0x1fffffu) < 8192u ? 1.0f : 2.0f;
        }
        unsigned long long th = MDS_NONE;
        for (int e = lane; e < total; e += 32) {
          const unsigned long long v = sh.theta[par][e];
          th = v < th ? v : th;
        }
        th = warp_min_u64(th);
        const int kmax = (c.m - j) < MDS_MAXK ? (c.m - j) : MDS_MAXK;
        int K = 0;
        float lx = 0.f, ly = 0.f, lz = 0.f;
        while (K < kmax) {
          unsigned long long cand = MDS_NONE;
#pragma unroll
          for (int q = 0; q < NQ; q++) {
            if (K > 0) pt[q] = mds_add<FAST_DIV>(pt[q], pf[q], px[q], py[q], pz[q], lx, ly, lz, t, r);
            const unsigned long long p = mds_pack(pt[q], pkey[q]);
            cand = p < cand ? p : cand;
          }
          const unsigned long long g = warp_min_u64(cand);
          if (!(g < th) || (unsigned)(g >> 32) >= MDS_PARKED) break;  // something outside the pool may be lower: next generation
          const int ol = __ffs(__ballot_sync(0xffffffffu, cand == g)) - 1;  // keys are unique: exactly one lane holds it
          float ox = 0.f, oy = 0.f, oz = 0.f;
          if (lane == ol) {
#pragma unroll
            for (int q = 0; q < NQ; q++)
              if (mds_pack(pt[q], pkey[q]) == g) {
                ox = px[q];
                oy = py[q];
                oz = pz[q];
                pt[q] = 1e9f;  // parked
              }
          }
          lx = __shfl_sync(0xffffffffu, ox, ol);
          ly = __shfl_sync(0xffffffffu, oy, ol);
          lz = __shfl_sync(0xffffffffu, oz, ol);
          const int old = (int)((unsigned)g & 0x1fffffu);
          if (lane == 0) {
            sh.picks[K] = make_float4(lx, ly, lz, __int_as_float(old));
            if (c.rank == 0) c.idxs[j + K] = old;
          }
          K++;
        }
        if (lane == 0) sh.npicks = K;
      }
      __syncthreads();
      const int K = sh.npicks;
      if (K == 0) {  // nothing left anywhere: the reference keeps returning index 0 (MDS_cuda.cu:121-133)
        if (c.rank == 0)
          for (int q = j + tid; q < c.m; q += MDS_THREADS) c.idxs[q] = 0;
        return c.m;
      }
      j += K;
      if (j >= c.m) return j;
    }
    // re-pack the live points for the narrower layout (all warps take this branch in the same generation)
    __syncthreads();
    if (tid == 0) {
      *c.st.count = 0;
      sh.npicks = 0;  // the pending picks have been applied
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < PT; i++) {
      if (temp[i] < 1e9f) {
        const int e = atomicAdd(c.st.count, 1);
        c.st.t[e] = temp[i];
        c.st.k[e] = key[i];
        c.st.loc[(int)(key[i] & 0x1fffffu) - c.kbeg] = (unsigned short)e;
      }
    }
    __syncthreads();
    for (int e = *c.st.count + tid; e < NEXT * MDS_THREADS; e += MDS_THREADS) {  // padding entries can never win
      c.st.t[e] = 2e9f;
      c.st.k[e] = 0xffffffffu;
    }
    __syncthreads();
    return j;
  }
};

template <int MDS_THREADS, int PT, bool FAST_DIV>
struct MdsChain {
  static __device__ __forceinline__ void run(int j, const MdsCtx& c, int& live, int& gen) {
    j = MdsLevel<MDS_THREADS, PT, FAST_DIV>::run(j, c, live, gen);
    if (j < c.m) MdsChain<MDS_THREADS, mds_next_pt(PT), FAST_DIV>::run(j, c, live, gen);
  }
};
template <int MDS_THREADS, bool FAST_DIV>
struct MdsChain<MDS_THREADS, 0, FAST_DIV> {
  static __device__ __forceinline__ void run(int, const MdsCtx&, int&, int&) {}
};

// dynamic shared memory: [stage t | stage k | count | loc | this CTA's points]; the points stay in global memory (L2) when
// they do not fit next to the rest (only for > 9216 points per CTA, far beyond SpareNet's 2048)
static inline size_t mds_smem_bytes(int per, int threads, int pt, bool stage_xyz) {
  const size_t cap = (size_t)threads * pt;
  return cap * 8 + 16 + (((size_t)per * 2 + 15) & ~(size_t)15) + (stage_xyz ? (size_t)per * 12 : 0);
}

template <int MDS_THREADS, int PT>
__global__ void __launch_bounds__(MDS_THREADS, 1) mds_cluster_kernel(const float* __restrict__ dataset, int n, int m,
                                                                      const float* __restrict__ mean_mst_length, int* __restrict__ idxs,
                                                                      int bs_mask, int bs_log2, int stage_xyz) {
  __shared__ __align__(16) MdsShared sh;
  extern __shared__ __align__(16) unsigned char dyn[];
  const uint32_t cs = cluster_nctarank();
  const uint32_t rank = cluster_ctarank();
  const int b = blockIdx.x / cs;
  const int tid = threadIdx.x;
  dataset += (size_t)b * n * 3;
  idxs += (size_t)b * m;
  const int chunk = (n + cs - 1) / cs;
  const int kbeg = rank * chunk;
  const int kend = (kbeg + chunk) < n ? (kbeg + chunk) : n;
  const int per = kend > kbeg ? kend - kbeg : 0;
  // carve the dynamic shared memory: this CTA's points (AoS), the staging SoA, the point -> entry map
  constexpr int CAP = MDS_THREADS * PT;
  MdsCtx c;
  c.st.t = reinterpret_cast<float*>(dyn);
  c.st.k = reinterpret_cast<unsigned*>(c.st.t + CAP);
  c.st.count = reinterpret_cast<int*>(c.st.k + CAP);
  c.st.loc = reinterpret_cast<unsigned short*>(c.st.count + 4);
  const float* sxyz = dataset + (size_t)kbeg * 3;
  if (stage_xyz) {
    float* sx = reinterpret_cast<float*>(dyn + (size_t)CAP * 8 + 16 + (((size_t)chunk * 2 + 15) & ~(size_t)15));
    for (int i = tid; i < per * 3; i += MDS_THREADS) sx[i] = sxyz[i];
    sxyz = sx;
  }
  // initial layout: every point of the CTA except the pre-chosen point 0 (MDS.cpp:119-121), then padding
  for (int e = tid; e < CAP; e += MDS_THREADS) {
    const int k = kbeg + e + ((kbeg == 0) ? 1 : 0);  // rank 0 skips k = 0
    const bool ok = k < kend;
    c.st.t[e] = ok ? 0.f : 2e9f;
    const unsigned rev = bs_log2 ? (__brev((unsigned)(k & bs_mask)) >> (32 - bs_log2)) : 0u;
    c.st.k[e] = ok ? ((rev << 21) | (unsigned)k) : 0xffffffffu;
    if (ok) c.st.loc[k - kbeg] = (unsigned short)e;
  }
  int live = per - ((kbeg == 0 && per > 0) ? 1 : 0);
  const float mml = mean_mst_length[b];
  const float t = (float)(5.0 * (double)mml * (double)mml);
  if (tid == 0) {
    mbar_init(&sh.bars[0], 1);
    mbar_init(&sh.bars[1], 1);
    fence_mbar_init();
    // generation 0 applies the pre-chosen point 0; index -1: it was never part of the layout, nothing to park
    sh.picks[0] = make_float4(dataset[0], dataset[1], dataset[2], __int_as_float(-1));
    sh.npicks = 1;
  }
  if (rank == 0 && tid == 0) idxs[0] = 0;
  __syncthreads();
  cluster_sync_all();  // peers must see initialised barriers before the first remote complete_tx
  c.dataset = dataset;
  c.idxs = idxs;
  c.sxyz = sxyz;
  c.sh = &sh;
  c.m = m;
  c.kbeg = kbeg;
  c.kend = kend;
  c.cs = cs;
  c.rank = rank;
  c.t = t;
  c.r = __frcp_rn(t);
  const unsigned tb = __float_as_uint(t);
  const bool fast = ((tb & 0x7fffffu) != 0x7fffffu) && ((tb >> 23) & 0xffu) > 1u && ((tb >> 23) & 0xffu) < 254u && !(tb >> 31);
  int gen = 0;
  if (m > 1) {
    if (fast) MdsChain<MDS_THREADS, PT, true>::run(1, c, live, gen);
    else MdsChain<MDS_THREADS, PT, false>::run(1, c, live, gen);
  }
  cluster_sync_all();  // no CTA may exit while a peer can still write into its shared memory
}

// ---- gather_points: out[b,c,j] = f[b,c,idx[b,j]]; backward scatters with atomics (the reference's
// non-atomic '+=' (MDS_cuda.cu:63-65) is only correct for unique indices; RED.ADD is correct always) -----
__global__ void __launch_bounds__(256) gather_fwd_kernel(const float* __restrict__ f, const int* __restrict__ idx, int C, int n, int m,
                                                          float* __restrict__ out) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  out[((size_t)b * C + c) * m + j] = f[((size_t)b * C + c) * n + idx[(size_t)b * m + j]];
}
__global__ void __launch_bounds__(256) gather_bwd_kernel(const float* __restrict__ g, const int* __restrict__ idx, int C, int n, int m,
                                                          float* __restrict__ gf) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  atomicAdd(&gf[((size_t)b * C + c) * n + idx[(size_t)b * m + j]], g[((size_t)b * C + c) * m + j]);
}

template <int MDS_THREADS, int PT>
static int mds_launch(const float* xyz, int B, int n, int m, const float* mml, int* idx, int cs, int bs_mask, int bs_log2, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * cs));
  cfg.blockDim = dim3(MDS_THREADS);
  const int per = (n + cs - 1) / cs;
  int stage_xyz = mds_smem_bytes(per, MDS_THREADS, PT, true) + sizeof(MdsShared) <= (size_t)220 * 1024 ? 1 : 0;
  const size_t smem = mds_smem_bytes(per, MDS_THREADS, PT, stage_xyz != 0);
  cudaError_t ea = cudaFuncSetAttribute(mds_cluster_kernel<MDS_THREADS, PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (ea != cudaSuccess) return (int)ea;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, mds_cluster_kernel<MDS_THREADS, PT>, xyz, n, m, mml, idx, bs_mask, bs_log2, stage_xyz);
}

}  // namespace snb

using namespace snb;

SNB_API size_t snb_mds_workspace_bytes(int B, int n, int m) {
  (void)B; (void)n; (void)m;
  return 0;  // densities live in registers; kept in the ABI for the out-of-register fallback
}

SNB_API int snb_mds_sample(const float* xyz, int B, int n, int m, const float* mean_mst_length, int* idx, void* workspace,
                           size_t workspace_bytes, void* stream) {
  (void)workspace; (void)workspace_bytes;
  if (B < 0 || n <= 0 || m < 0) return SNB_EINVAL;
  if (B == 0 || m == 0) return SNB_OK;
  if (n >= (1 << 21)) return SNB_ELIMIT;
  cudaStream_t s = (cudaStream_t)stream;
  int bs = 1;
  int lg = 0;
  while (bs * 2 <= n && bs < 1024) { bs *= 2; lg++; }  // opt_n_threads(n), MDS_cuda.cu:8-12
  // cluster size: as many SMs per sample as the batch leaves free (<= 8, power of two)
  int cs = 1;
  while (cs * 2 <= MDS_MAX_CLUSTER && B * cs * 2 <= kNumSMs) cs *= 2;
  int per = (n + cs - 1) / cs;  // points per CTA
  while (per > 512 * 24 && cs < MDS_MAX_CLUSTER) {  // too many points for the register file: widen the cluster
    cs *= 2;
    per = (n + cs - 1) / cs;
  }
  int rc;
  const int bm = bs - 1;
#define MDS_GO(T, P) rc = mds_launch<T, P>(xyz, B, n, m, mean_mst_length, idx, cs, bm, lg, s)
  // Experiment hook (development): SNB_MDS_LAYOUT="<cluster size>,<threads>" picks another co-residency layout, e.g. "8,128":
  // 8 thin CTAs per sample, two samples' CTAs sharing an SM so one sample's exchange latency hides behind the other's math.
  int force_threads = 0;
  // Half-filled machine (e.g. B = 32: 4 SMs per sample): 8 thin CTAs per sample, two samples' CTAs sharing an SM, so one
  // sample's exchange latency hides behind the other's arithmetic (12.3 vs 13.1 ms at B=32, n=18432, m=16384).
  if (cs == 4 && B * 8 <= 2 * kNumSMs && (n + 7) / 8 <= 128 * 18) {
    cs = 8;
    force_threads = 128;
    per = (n + cs - 1) / cs;
  }
  if (const char* e = getenv("SNB_MDS_LAYOUT")) {
    int a = 0, t = 0;
    if (sscanf(e, "%d,%d", &a, &t) == 2 && (a == 1 || a == 2 || a == 4 || a == 8) && (t == 128 || t == 256 || t == 512)) {
      cs = a;
      force_threads = t;
      per = (n + cs - 1) / cs;
    }
  }
  if (force_threads == 128) {
    if (per <= 128 * 9) MDS_GO(128, 9);
    else if (per <= 128 * 18) MDS_GO(128, 18);
    else if (per <= 128 * 36) MDS_GO(128, 36);
    else return SNB_ELIMIT;
  } else if (force_threads == 512) {
    if (per <= 512 * 5) MDS_GO(512, 5);
    else if (per <= 512 * 9) MDS_GO(512, 9);
    else if (per <= 512 * 12) MDS_GO(512, 12);
    else return SNB_ELIMIT;
  } else if (per <= 256 * 2) MDS_GO(256, 2);
  else if (per <= 256 * 4) MDS_GO(256, 4);
  else if (per <= 256 * 6) MDS_GO(256, 6);
  else if (per <= 256 * 9) MDS_GO(256, 9);
  else if (per <= 256 * 12) MDS_GO(256, 12);
  else if (per <= 256 * 18) MDS_GO(256, 18);
  else if (per <= 512 * 12) MDS_GO(512, 12);
  else if (per <= 512 * 18) MDS_GO(512, 18);
  else if (per <= 512 * 24) MDS_GO(512, 24);
  else return SNB_ELIMIT;  // n > 8*512*24 = 98304 points per sample
#undef MDS_GO
  if (rc != 0) return rc;
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

SNB_API int snb_gather_fwd(const float* features, const int* idx, int B, int C, int n, int m, float* out, void* stream) {
  if (B < 0 || C < 0 || n < 0 || m < 0) return SNB_EINVAL;
  if (B == 0 || C == 0 || m == 0) return SNB_OK;
  if (B > 65535 || C > 65535) return SNB_ELIMIT;
  dim3 grid((m + 255) / 256, C, B);
  gather_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(features, idx, C, n, m, out);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

SNB_API int snb_gather_bwd(const float* grad_out, const int* idx, int B, int C, int n, int m, float* grad_features, void* stream) {
  if (B < 0 || C < 0 || n < 0 || m < 0) return SNB_EINVAL;
  if (B == 0 || C == 0 || n == 0) return SNB_OK;
  if (B > 65535 || C > 65535) return SNB_ELIMIT;
  cudaStream_t s = (cudaStream_t)stream;
  SNB_CUDA(cudaMemsetAsync(grad_features, 0, sizeof(float) * (size_t)B * C * n, s));
  if (m == 0) return SNB_OK;
  dim3 grid((m + 255) / 256, C, B);
  gather_bwd_kernel<<<grid, 256, 0, s>>>(grad_out, idx, C, n, m, grad_features);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
