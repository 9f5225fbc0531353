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
//   1. every worker warp publishes its M lowest (density, tie key) candidates -- with coordinates -- and its (M+1)-th lowest
//      pair as a bound; all candidates of the cluster form the POOL (<= 256 slots), theta = min over warps of the bounds.
//      Every point outside the pool is >= theta, and densities only ever grow;
//   2. a dedicated REPLAY warp per CTA keeps the pool entries below theta and replays the sequential algorithm on them, in
//      registers, with the very same arithmetic: update with the last pick's weights, take the arg-min, accept it while
//      (density, key) < theta -- an accepted pick is the global arg-min of the sequential algorithm (nothing outside the pool
//      can have dropped below theta).  The first candidate is always accepted;
//   3. every accepted pick is STREAMED to the worker warps through shared memory (one self-validating 16-byte entry) the moment
//      it is known: they apply it to their own points (the same sequence of fp32 additions as the reference's rounds) while the
//      replay warp is already working on the next one, park the picked points, and select again when the generation is done.
// One cluster exchange (a bulk copy per warp and peer CTA + mbarrier complete_tx, no barrier.cluster) per generation instead of
// one per pick; the per-pick chain is ~130 instructions of one warp, overlapped with the workers' arithmetic.  A thread-block CLUSTER (up to 8
// CTAs) owns one sample; every point (xyz + density) lives in registers for the whole kernel; the reference does 11 block
// barriers and a global read-modify-write of `temp` per pick.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "common.cuh"

namespace snb {

constexpr int MDS_MAX_WARPS = 16;  // worker warps per CTA: 4, 8 or 16 (plus the replay warp)
constexpr int MDS_MAX_CLUSTER = 8;
constexpr int MDS_MAXM = 8;        // candidates a worker warp contributes to a generation's pool (fewer when > 32 warps)
constexpr int MDS_POOL = 256;      // pool capacity = total worker warps x M
constexpr int MDS_MAXK = MDS_POOL; // picks per generation (capacity of the pick list)
constexpr unsigned long long MDS_NONE = 0xffffffffffffffffull;
constexpr unsigned MDS_PARKED = 0x4e6e6b28u;  // bits of 1e9f: parked / padding entries compare >= this
constexpr unsigned MDS_END = 0x1fffffu;        // index field of the entry that closes a generation's pick list

// The exchange: a warp stages its message (bound + candidates) in its own shared memory and sends it to every CTA of the
// cluster with ONE bulk copy each (cp.async.bulk shared::cta -> shared::cluster), which also completes the message's bytes on
// the peer's mbarrier -- data and "it arrived" are one operation and nobody executes barrier.cluster inside the loop.  (One
// st.async per 8/16-byte field made ~400 complete_tx operations per generation on every mbarrier: ~15k cycles of dead time.)
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t remote_dst, uint32_t local_src, uint32_t bytes, uint32_t remote_bar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(remote_dst),
               "r"(local_src), "r"(bytes), "r"(remote_bar)
               : "memory");
}
// A pick travels from the replay warp to the workers as ONE 16-byte shared-memory store (x, y, z, generation tag << 21 | index):
// a 128-bit access of one thread is a single shared-memory transaction, so the entry validates itself and the per-pick path
// needs neither a counter nor a fence.  (A st.release per pick -- MEMBAR + the pending global store of the index -- cost
// ~800 cycles per pick on the replay warp's dependent chain.)
__device__ __forceinline__ void st_volatile_v4(float4* p, float a, float b, float c, unsigned d) {
  asm volatile("st.volatile.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(smem_u32(p)), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)),
               "r"(__float_as_uint(c)), "r"(d)
               : "memory");
}
__device__ __forceinline__ void ld_volatile_v4(const float4* p, float& a, float& b, float& c, unsigned& d) {
  unsigned ua, ub, uc;
  asm volatile("ld.volatile.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(ua), "=r"(ub), "=r"(uc), "=r"(d) : "r"(smem_u32(p)) : "memory");
  a = __uint_as_float(ua);
  b = __uint_as_float(ub);
  c = __uint_as_float(uc);
}
__device__ __forceinline__ void mbar_complete_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.complete_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ unsigned mds_tag(int gen) { return ((unsigned)(gen % 2047) + 1u) << 21; }
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }


// Development statistics (SNB_MDS_STATS builds only): cycle accounting of block 0's replay warp and worker warp 0.
#ifdef SNB_MDS_STATS
__device__ unsigned long long g_mds_stats[16];
#define MDS_STAT_ADD(i, v) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) atomicAdd(&g_mds_stats[i], (unsigned long long)(v)); } while (0)
#define MDS_CLOCK() clock64()
#else
#define MDS_STAT_ADD(i, v) do { } while (0)
#define MDS_CLOCK() 0ll
#pragma nv_diag_suppress 177
#endif

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

struct MdsStage {          // staging area in dynamic shared memory, capacity = WORKERS * PT0 entries
  float* t;                // running density of the entry's point (2e9 = padding)
  unsigned* k;             // tie key << 21 | point index (~0 = padding); coordinates and the x2 factor follow from the index
  unsigned short* loc;     // point k (local index k / cs) -> its entry in the current layout
  int* count;
};

// Message of one worker warp: [bound (the warp's (M+1)-th lowest pair) | 8 bytes unused | M x float4 coordinates | M x u64 (density
// bits, key)], padded to a multiple of 16 bytes.
__host__ __device__ constexpr int mds_msg_bytes(int msel) { return (16 + 24 * msel + 15) & ~15; }
constexpr int MDS_POOL_BYTES = 9216;  // >= total worker warps x message bytes for every layout (32 x 208, 64 x 112, 128 x 64)

struct MdsShared {         // static shared memory of one CTA
  unsigned char pool[2][MDS_POOL_BYTES];           // the messages of all worker warps of the cluster, double buffered by generation parity
  unsigned char stage[2][MDS_MAX_WARPS * 208];     // this CTA's outgoing messages (source of the bulk copies)
  unsigned long long cpack[MDS_POOL];              // the replay warp's compacted pool (entries below theta)
  float4 cxyz[MDS_POOL];
  float4 bcast[2];                                 // the replay warp's winner broadcast slot (alternating)
  float4 picks[MDS_MAXK + 1];                      // accepted picks of the current generation: x, y, z, tag << 21 | index
  uint64_t bars[2];
  long long dbg_end, dbg_pub;                      // SNB_MDS_STATS only
  unsigned long long dbg_lastpub;
};

struct MdsCtx {
  const float* dataset;
  int* idxs;
  const float* sxyz;   // this CTA's points (shared memory copy when it fits), local point i at sxyz[i * xs]
  int xs;              // 3 (staged copy) or 3 * cs (global memory)
  MdsShared* sh;
  MdsStage st;
  int m, msel, csh;    // csh = log2(cluster size): point k belongs to CTA k & (cs-1), local index k >> csh
  uint32_t cs, rank;
  float t, r;
};

__host__ __device__ constexpr int mds_next_pt(int pt) { return pt > 12 ? pt - 4 : (pt > 6 ? pt - 3 : (pt > 2 ? pt - 2 : 0)); }  // 24 20 16 12 9 6 4 2

// Thread layout of a CTA.  The replay warp's per-pick chain is the critical path of the whole kernel and it is latency bound,
// while the workers' updates are always ready to issue, so the warp schedulers (warp id mod 4) matter:
//   WORKERS = 224 (the layout SpareNet's refiner size gets): 8 warps, the replay warp is warp 0 and shares its scheduler with ONE
//                 worker warp (warp 4); measured best on the bench clouds (8.3 / 10.1 ms per call);
//   WORKERS = 192 (experiment, SNB_MDS_LAYOUT="4,192"): warp 4 only waits at the final barrier, the replay warp has its scheduler to
//                 itself -- 450 instead of 740 cycles per pick, but the workers lose a quarter of their issue slots (9.2 ms);
//   otherwise     WORKERS + 32 threads, the replay warp comes after the workers and shares its scheduler with two of them.
__host__ __device__ constexpr bool mds_iso(int workers) { return workers == 192; }
__host__ __device__ constexpr bool mds_first(int workers) { return workers == 192 || workers == 224; }  // replay warp = warp 0
__host__ __device__ constexpr int mds_threads(int workers) { return mds_first(workers) ? 256 : workers + 32; }

// ---- worker warps ------------------------------------------------------------------------------------------------------
template <int WORKERS, int PT, bool FAST_DIV>
struct MdsLevel {
  // runs generations while the CTA still holds more live points than the next narrower layout can take;
  // returns true when the kernel is finished
  static __device__ __forceinline__ bool run(const MdsCtx& c, int& live, int& gen) {
    constexpr int WARPS = WORKERS / 32;
    constexpr int NEXT = mds_next_pt(PT);
    const int lane = threadIdx.x & 31;
    const int warp = mds_iso(WORKERS) ? (int)(threadIdx.x >> 5) - 1 - (int)(threadIdx.x >= 160)
                                      : (mds_first(WORKERS) ? (int)(threadIdx.x >> 5) - 1 : (int)(threadIdx.x >> 5));  // worker warp
    const int tid = warp * 32 + lane;                                                                                      // worker thread
    MdsShared& sh = *c.sh;
    const float t = c.t, r = c.r;
    float x[PT], y[PT], z[PT], temp[PT], fac[PT];
    unsigned key[PT];
#pragma unroll
    for (int i = 0; i < PT; i++) {  // entry e = tid + i*WORKERS of the staged layout
      const int e = tid + i * WORKERS;
      temp[i] = c.st.t[e];
      key[i] = c.st.k[e];
      const int k = (int)(key[i] & 0x1fffffu);
      const int kk = key[i] != 0xffffffffu ? k >> c.csh : 0;
      x[i] = c.sxyz[kk * c.xs + 0];
      y[i] = c.sxyz[kk * c.xs + 1];
      z[i] = c.sxyz[kk * c.xs + 2];
      fac[i] = k < 8192 ? 1.0f : 2.0f;  // MDS_cuda.cu:111-112 (k > 8191 counts double)
    }
    const int cs = (int)c.cs, msel = c.msel;
    const uint32_t my_slot = c.rank * WARPS + warp;
    const int total = cs * WARPS;
    (void)total;

    for (;;) {
      // ---- this warp's msel lowest (density, key) pairs and the next one as its bound: lane s keeps the s-th ------------
      const long long w0 = MDS_CLOCK();
      unsigned long long mine = MDS_NONE, second = MDS_NONE;  // this lane's two lowest pairs, in one pass
#pragma unroll
      for (int i = 0; i < PT; i++) {
        const unsigned long long p = mds_pack(temp[i], key[i]);
        const unsigned long long hi2 = p < mine ? mine : p;
        mine = p < mine ? p : mine;
        second = hi2 < second ? hi2 : second;
      }
      unsigned long long mysel = MDS_NONE;
      int taken = 0;
#pragma unroll 1
      for (int s = 0; s <= msel; s++) {
        unsigned long long w = warp_min_u64(mine);
        if ((unsigned)(w >> 32) >= MDS_PARKED) break;  // parked / padding: this warp has run out of live points
        if (lane == s) mysel = w;
        if (s < msel && mine == w) {                   // the owning lane (keys are unique) moves on to its next entry
          if (taken++ == 0) {
            mine = second;
          } else {
            unsigned long long nx = MDS_NONE;
#pragma unroll
            for (int i = 0; i < PT; i++) {
              const unsigned long long p = mds_pack(temp[i], key[i]);
              nx = (p > w && p < nx) ? p : nx;
            }
            mine = nx;
          }
        }
      }
      // ---- publish them to every CTA of the cluster ------------------------------------------------------------------
      const int par = gen & 1;
      {
        const int mb = mds_msg_bytes(msel);
        unsigned char* msg = &sh.stage[par][warp * mb];       // source of the bulk copies to the peers
        unsigned char* own = &sh.pool[par][my_slot * mb];     // this CTA's copy is written directly (no copy onto itself)
        if (lane < msel) {
          const int kk = mysel != MDS_NONE ? (int)((unsigned)mysel & 0x1fffffu) >> c.csh : 0;  // one of this CTA's points
          const float4 cx = make_float4(c.sxyz[kk * c.xs + 0], c.sxyz[kk * c.xs + 1], c.sxyz[kk * c.xs + 2], 0.f);
          reinterpret_cast<float4*>(msg + 16)[lane] = cx;
          reinterpret_cast<unsigned long long*>(msg + 16 + 16 * msel)[lane] = mysel;
          reinterpret_cast<float4*>(own + 16)[lane] = cx;
          reinterpret_cast<unsigned long long*>(own + 16 + 16 * msel)[lane] = mysel;
        } else if (lane == msel) {
          *reinterpret_cast<unsigned long long*>(msg) = mysel;
          *reinterpret_cast<unsigned long long*>(own) = mysel;
#ifdef SNB_MDS_STATS
          unsigned long long gt;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
          *reinterpret_cast<unsigned long long*>(msg + 8) = gt;
          *reinterpret_cast<unsigned long long*>(own + 8) = gt;
#endif
        }
        fence_proxy_async();  // the generic-proxy writes above, before the async-proxy reads of the bulk copies
        __syncwarp();
        if (lane < cs && lane != (int)c.rank)
          bulk_copy_to_peer(mapa_shared(smem_u32(&sh.pool[par][my_slot * mb]), (uint32_t)lane), smem_u32(msg), (uint32_t)mb,
                            mapa_shared(smem_u32(&sh.bars[par]), (uint32_t)lane));
        if (lane == (int)c.rank) {  // the local message: its bytes are complete (ordered by the __syncwarp and this fence)
          __threadfence_block();
          mbar_complete_tx(&sh.bars[par], (uint32_t)mb);
        }
      }
      gen++;
      const long long w1 = MDS_CLOCK();
      long long wapply = 0;

      // ---- apply the picks of generation `gen` as the replay warp streams them ----------------------------------------
      int applied = 0;
      bool fin = false;
      const unsigned tag = mds_tag(gen);
#ifdef SNB_MDS_STATS
      if (tid == 0) sh.dbg_pub = w1;
      if (lane == 0) atomicMax(&sh.dbg_lastpub, (unsigned long long)w1);
      long long a0 = MDS_CLOCK(), wlag = 0;
#endif
#pragma unroll 1
      for (;;) {
        float4 pk;
        unsigned pw;
        ld_volatile_v4(&sh.picks[applied], pk.x, pk.y, pk.z, pw);
        if ((pw & 0xffe00000u) != tag) {  // not written yet
          __nanosleep(20);
#ifdef SNB_MDS_STATS
          a0 = MDS_CLOCK();
#endif
          continue;
        }
        const int pidx = (int)(pw & 0x1fffffu);
        if (pidx == (int)MDS_END) {
          fin = pk.x != 0.f;
#ifdef SNB_MDS_STATS
          wlag = MDS_CLOCK() - *(volatile long long*)&sh.dbg_end;
#endif
          break;
        }
        if ((pidx & ((1 << c.csh) - 1)) == (int)c.rank) {  // park it: every warp keeps the live count, only the owner touches its registers
          live--;
          const int e = c.st.loc[pidx >> c.csh];
          if ((e % WORKERS) == tid) {
            const int slot = e / WORKERS;
#pragma unroll
            for (int i = 0; i < PT; i++)
              if (i == slot) temp[i] = 1e9f;   // 1e9f + w == 1e9f for every later w <= 2
          }
        }
#ifndef SNB_MDS_NOAPPLY  // (timing experiment: the replay warp's chain with idle workers; results are wrong)
        {
#pragma unroll
          for (int i = 0; i < PT; i++) temp[i] = mds_add<FAST_DIV>(temp[i], fac[i], x[i], y[i], z[i], pk.x, pk.y, pk.z, t, r);
        }
#endif
        applied++;
#ifdef SNB_MDS_STATS
        const long long a1 = MDS_CLOCK();
        wapply += a1 - a0;
        a0 = a1;
#endif
      }
      if (warp == 0) {
        MDS_STAT_ADD(8, w1 - w0);                       // worker warp 0: select + publish
        MDS_STAT_ADD(9, wapply);                        // ... applying picks
        MDS_STAT_ADD(10, MDS_CLOCK() - w1 - wapply);    // ... waiting for picks
#ifdef SNB_MDS_STATS
        MDS_STAT_ADD(11, wlag);                         // ... end of replay -> worker has applied everything
#endif
      }
      if (fin) return true;
      if (NEXT > 0 && live <= NEXT * WORKERS) break;  // re-pack into the narrower layout (uniform over the CTA's workers)
    }
    // re-pack the live points for the narrower layout (all worker warps take this branch in the same generation)
    bar_sync_named(1, WORKERS);
    if (tid == 0) *c.st.count = 0;
    bar_sync_named(1, WORKERS);
#pragma unroll
    for (int i = 0; i < PT; i++) {
      if (temp[i] < 1e9f) {
        const int e = atomicAdd(c.st.count, 1);
        c.st.t[e] = temp[i];
        c.st.k[e] = key[i];
        c.st.loc[(int)(key[i] & 0x1fffffu) >> c.csh] = (unsigned short)e;
      }
    }
    bar_sync_named(1, WORKERS);
    for (int e = *c.st.count + tid; e < NEXT * WORKERS; e += WORKERS) {  // padding entries can never win
      c.st.t[e] = 2e9f;
      c.st.k[e] = 0xffffffffu;
    }
    bar_sync_named(1, WORKERS);
    return false;
  }
};

template <int WORKERS, int PT, bool FAST_DIV>
struct MdsChain {
  static __device__ __forceinline__ void run(const MdsCtx& c, int& live, int& gen) {
    if (!MdsLevel<WORKERS, PT, FAST_DIV>::run(c, live, gen)) MdsChain<WORKERS, mds_next_pt(PT), FAST_DIV>::run(c, live, gen);
  }
};
template <int WORKERS, bool FAST_DIV>
struct MdsChain<WORKERS, 0, FAST_DIV> {
  static __device__ __forceinline__ void run(const MdsCtx&, int&, int&) {}
};

// One generation of the replay on a pool of NT x 32 register-resident entries.  The per-pick dependent chain is the critical
// path of the whole kernel, so it is kept to: update (distance -> division -> expf -> add, the tiers interleaved) -> per-lane
// integer min of the density bits -> ONE warp reduction (CREDUX.MIN) -> the owner of the minimum stores (x, y, z, key) to a
// 16-byte broadcast slot -> every lane loads it -> next update.  No ballot / find-first-set / shuffle round trip, no divergent
// branch.  Exactness: the winner is unique iff exactly one entry carries the minimal density bits (one CREDUX.ADD, off the
// chain); on a tie of densities, or when the minimum reaches theta's density, the 64-bit (density, key) rule decides.
// Returns the number of accepted picks.
template <int NT, bool FAST_DIV>
__device__ __forceinline__ int mds_replay_gen(MdsShared& sh, int P, int kmax, unsigned long long th, unsigned tag, float t, float r) {
  const int lane = threadIdx.x & 31;
  float px[NT], py[NT], pz[NT], pt[NT], pf[NT];
  unsigned pkey[NT];
#pragma unroll
  for (int q = 0; q < NT; q++) {
    const int e = lane + 32 * q;
    const bool ok = e < P;
    const unsigned long long p = ok ? sh.cpack[e] : MDS_NONE;
    const float4 cx = sh.cxyz[ok ? e : 0];
    pt[q] = ok ? __uint_as_float((unsigned)(p >> 32)) : 2e9f;
    pkey[q] = ok ? (unsigned)p : 0xffffffffu;
    px[q] = cx.x;
    py[q] = cx.y;
    pz[q] = cx.z;
    pf[q] = (pkey[q] & 0x1fffffu) < 8192u ? 1.0f : 2.0f;
  }
  const unsigned th_hi = (unsigned)(th >> 32);
  int K = 0;
  float lx = 0.f, ly = 0.f, lz = 0.f;
  while (K < kmax) {
    if (K > 0) {  // the first pick of a generation is taken from the pool as it arrived
#pragma unroll
      for (int q = 0; q < NT; q++) pt[q] = mds_add<FAST_DIV>(pt[q], pf[q], px[q], py[q], pz[q], lx, ly, lz, t, r);
    }
    unsigned chi = __float_as_uint(pt[0]);
#pragma unroll
    for (int q = 1; q < NT; q++) chi = min(chi, __float_as_uint(pt[q]));
    const unsigned mh = __reduce_min_sync(0xffffffffu, chi);
    if (mh >= MDS_PARKED) break;  // the pool is used up
    float4* slot = &sh.bcast[K & 1];
    int nm = 0;
#pragma unroll
    for (int q = 0; q < NT; q++) {
      const bool mt = __float_as_uint(pt[q]) == mh;
      nm += mt ? 1 : 0;
      if (mt) st_volatile_v4(slot, px[q], py[q], pz[q], pkey[q]);
    }
    const int tot = __reduce_add_sync(0xffffffffu, nm);
    if (tot != 1 || !(mh < th_hi)) {  // rare: equal densities, or the minimum has reached theta's density
      unsigned clo = 0xffffffffu;
#pragma unroll
      for (int q = 0; q < NT; q++) clo = (__float_as_uint(pt[q]) == mh && pkey[q] < clo) ? pkey[q] : clo;
      const unsigned ml = __reduce_min_sync(0xffffffffu, clo);
      if (!((((unsigned long long)mh << 32) | ml) < th)) break;  // something outside the pool may be lower: next generation
      __syncwarp();
#pragma unroll
      for (int q = 0; q < NT; q++) {
        const bool mt = __float_as_uint(pt[q]) == mh && pkey[q] == ml;  // keys are unique: exactly one entry
        if (mt) st_volatile_v4(slot, px[q], py[q], pz[q], pkey[q]);
        pt[q] = mt ? 1e9f : pt[q];  // parked
      }
    } else {
#pragma unroll
      for (int q = 0; q < NT; q++) pt[q] = __float_as_uint(pt[q]) == mh ? 1e9f : pt[q];  // parked
    }
    __syncwarp();
    unsigned key;
    ld_volatile_v4(slot, lx, ly, lz, key);
    if (lane == 0) st_volatile_v4(&sh.picks[K], lx, ly, lz, tag | (key & 0x1fffffu));
    K++;
  }
  return K;
}

// ---- the replay warp: the sequential algorithm on the pool ---------------------------------------------------------------
template <bool FAST_DIV>
__device__ __noinline__ void mds_replay(const MdsCtx& c, int total_warps) {
  MdsShared& sh = *c.sh;
  const int lane = threadIdx.x & 31;
  const float t = c.t, r = c.r;
  const int S = total_warps * c.msel;
  const int msel = c.msel, mb = mds_msg_bytes(c.msel), lsel = 31 - __clz(c.msel);
  const uint32_t gen_bytes = (uint32_t)(total_warps * mb);
  int j = 1, kprev = -1;
  for (int gen = 0;; gen++) {  // pool `gen` produces the picks of generation gen+1
    const int par = gen & 1;
    const long long c0 = MDS_CLOCK();
    if (lane == 0) mbar_expect_tx(&sh.bars[par], gen_bytes);  // arm this generation's phase (the single expected arrival)
    __syncwarp();
    mbar_wait_tx(&sh.bars[par], (uint32_t)(gen >> 1) & 1u);   // k-th use of bars[par] has parity k & 1
    const long long c1 = MDS_CLOCK();
    MDS_STAT_ADD(6, gen > 0 ? c1 - sh.dbg_pub : 0);  // worker warp 0 published -> pool complete here
    MDS_STAT_ADD(7, gen > 0 ? c1 - (long long)sh.dbg_lastpub : 0);  // last worker warp of THIS CTA published -> pool complete here
#ifdef SNB_MDS_STATS
    if (blockIdx.x == 0 && gen > 0) {  // per rank: pool complete here - the rank's last publish (global timer, ns)
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      const int wpc = total_warps / (int)c.cs;
      for (int rk = 0; rk < (int)c.cs && rk < 4; rk++) {
        unsigned long long last = 0;
        for (int w = 0; w < wpc; w++) {
          const unsigned long long v = *reinterpret_cast<const unsigned long long*>(&sh.pool[par][(rk * wpc + w) * mb + 8]);
          last = v > last ? v : last;
        }
        MDS_STAT_ADD(12 + rk, now - last);
      }
    }
#endif
    for (int q = lane; q <= kprev; q += 32) sh.picks[q] = make_float4(0.f, 0.f, 0.f, 0.f);  // every worker is past them
    __syncwarp();
    unsigned long long th = MDS_NONE;
    for (int e = lane; e < total_warps; e += 32) {
      const unsigned long long v = *reinterpret_cast<const unsigned long long*>(&sh.pool[par][e * mb]);
      th = v < th ? v : th;
    }
    th = warp_min_u64(th);
    // keep only the entries below theta: nothing else can be accepted in this generation (densities only grow)
    int P = 0;
#pragma unroll 1
    for (int e0 = 0; e0 < S; e0 += 128) {  // four independent tiers per trip (msel is a power of two)
      unsigned long long p[4];
      const unsigned char* msg[4];
      int es[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int e = e0 + 32 * u + lane;
        msg[u] = &sh.pool[par][(e < S ? e >> lsel : 0) * mb];
        es[u] = e & (msel - 1);
        p[u] = e < S ? reinterpret_cast<const unsigned long long*>(msg[u] + 16 + 16 * msel)[es[u]] : MDS_NONE;
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const bool ok = p[u] < th && (unsigned)(p[u] >> 32) < MDS_PARKED;
        const unsigned mask = __ballot_sync(0xffffffffu, ok);
        if (ok) {
          const int pos = P + __popc(mask & ((1u << lane) - 1u));
          sh.cpack[pos] = p[u];
          sh.cxyz[pos] = reinterpret_cast<const float4*>(msg[u] + 16)[es[u]];
        }
        P += __popc(mask);
      }
    }
    __syncwarp();
    const unsigned tag = mds_tag(gen + 1);
    if (P == 0) {  // nothing left anywhere: the reference keeps returning index 0 (MDS_cuda.cu:121-133)
      if (c.rank == 0)
        for (int q = j + lane; q < c.m; q += 32) c.idxs[q] = 0;
      if (lane == 0) st_volatile_v4(&sh.picks[0], 1.f, 0.f, 0.f, tag | MDS_END);
      return;
    }
    const int kmax = (c.m - j) < P ? (c.m - j) : P;
    const long long c2 = MDS_CLOCK();
    int K;
    switch ((P + 31) >> 5) {  // the per-pick chain is unrolled over exactly the tiers (32 entries each) the pool fills
      case 1: K = mds_replay_gen<1, FAST_DIV>(sh, P, kmax, th, tag, t, r); break;
      case 2: K = mds_replay_gen<2, FAST_DIV>(sh, P, kmax, th, tag, t, r); break;
      case 3: K = mds_replay_gen<3, FAST_DIV>(sh, P, kmax, th, tag, t, r); break;
      case 4: K = mds_replay_gen<4, FAST_DIV>(sh, P, kmax, th, tag, t, r); break;
      case 5: K = mds_replay_gen<5, FAST_DIV>(sh, P, kmax, th, tag, t, r); break;
      case 6: K = mds_replay_gen<6, FAST_DIV>(sh, P, kmax, th, tag, t, r); break;
      case 7: K = mds_replay_gen<7, FAST_DIV>(sh, P, kmax, th, tag, t, r); break;
      default: K = mds_replay_gen<8, FAST_DIV>(sh, P, kmax, th, tag, t, r); break;
    }
    const bool fin = j + K >= c.m;
    if (lane == 0) {
#ifdef SNB_MDS_STATS
      *(volatile long long*)&sh.dbg_end = MDS_CLOCK();
#endif
      st_volatile_v4(&sh.picks[K], fin ? 1.f : 0.f, 0.f, 0.f, tag | MDS_END);
    }
    __syncwarp();
    if (c.rank == 0)
      for (int q = lane; q < K; q += 32) c.idxs[j + q] = (int)(__float_as_uint(sh.picks[q].w) & 0x1fffffu);
    j += K;
    MDS_STAT_ADD(0, 1);                    // generations
    MDS_STAT_ADD(1, K);                    // picks
    MDS_STAT_ADD(2, c1 - c0);              // replay warp: waiting for the pool
    MDS_STAT_ADD(3, c2 - c1);              // ... compaction + load
    MDS_STAT_ADD(4, MDS_CLOCK() - c2);     // ... replay loop
    MDS_STAT_ADD(5, P);                    // pool entries below theta
    if (fin) return;
    kprev = K;
  }
}

// dynamic shared memory: [stage t | stage k | count | loc | this CTA's points]; the points stay in global memory (L2) when
// they do not fit next to the rest (only for > 9216 points per CTA, far beyond SpareNet's 2048)
static inline size_t mds_smem_bytes(int per, int threads, int pt, bool stage_xyz) {
  const size_t cap = (size_t)threads * pt;
  return cap * 8 + 16 + (((size_t)per * 2 + 15) & ~(size_t)15) + (stage_xyz ? (size_t)per * 12 : 0);
}

template <int WORKERS, int PT, int OCC>
__global__ void __launch_bounds__(mds_threads(WORKERS), OCC) mds_cluster_kernel(const float* __restrict__ dataset, int n, int m,
                                                                        const float* __restrict__ mean_mst_length, int* __restrict__ idxs,
                                                                        int bs_mask, int bs_log2, int stage_xyz, int msel) {
  __shared__ __align__(16) MdsShared sh;
  extern __shared__ __align__(16) unsigned char dyn[];
  const uint32_t cs = cluster_nctarank();
  const uint32_t rank = cluster_ctarank();
  const int b = blockIdx.x / cs;
  const int tid = threadIdx.x;
  dataset += (size_t)b * n * 3;
  idxs += (size_t)b * m;
  // Points are dealt round-robin to the CTAs of the cluster (k -> CTA k mod cs): points k < 8192 weigh half (MDS_cuda.cu:111),
  // so they are picked -- and parked -- first; contiguous chunks left the CTAs holding the high indices at the widest register
  // layout long after the others had shrunk, and the slowest CTA sets the pace of every generation.
  const int csh = 31 - __clz((int)cs);
  const int chunk = (n + cs - 1) / cs;
  const int per = (n > (int)rank) ? (n - (int)rank + (int)cs - 1) / (int)cs : 0;
  // carve the dynamic shared memory: this CTA's points (AoS), the staging SoA, the point -> entry map
  constexpr int CAP = WORKERS * PT;
  constexpr int THREADS = mds_threads(WORKERS);
  MdsCtx c;
  c.st.t = reinterpret_cast<float*>(dyn);
  c.st.k = reinterpret_cast<unsigned*>(c.st.t + CAP);
  c.st.count = reinterpret_cast<int*>(c.st.k + CAP);
  c.st.loc = reinterpret_cast<unsigned short*>(c.st.count + 4);
  const float* sxyz = dataset + (size_t)rank * 3;
  int xs = 3 * (int)cs;
  if (stage_xyz) {
    float* sx = reinterpret_cast<float*>(dyn + (size_t)CAP * 8 + 16 + (((size_t)chunk * 2 + 15) & ~(size_t)15));
    for (int i = tid; i < per * 3; i += THREADS) sx[i] = sxyz[(size_t)(i / 3) * xs + (i % 3)];
    sxyz = sx;
    xs = 3;
  }
  const float mml = mean_mst_length[b];
  const float t = (float)(5.0 * (double)mml * (double)mml);
  const float rcp = __frcp_rn(t);
  const unsigned tb = __float_as_uint(t);
  const bool fast = ((tb & 0x7fffffu) != 0x7fffffu) && ((tb >> 23) & 0xffu) > 1u && ((tb >> 23) & 0xffu) < 254u && !(tb >> 31);
  if (stage_xyz) __syncthreads();
  // initial layout: every point of the CTA except the pre-chosen point 0 (MDS.cpp:119-121), then padding; the densities
  // start at the weights of round 1 (the pick is point 0: 0 + w is exact)
  const float x0 = dataset[0], y0 = dataset[1], z0 = dataset[2];
  for (int e = tid; e < CAP; e += THREADS) {
    const int li = e + ((rank == 0) ? 1 : 0);  // local point index; rank 0 skips k = 0
    const int k = (li << csh) + (int)rank;
    const bool ok = li < per;
    float w0 = 2e9f;
    if (ok) {
      const float* p = sxyz + (size_t)li * xs;
      const float fac = k < 8192 ? 1.0f : 2.0f;
      w0 = fast ? mds_add<true>(0.f, fac, p[0], p[1], p[2], x0, y0, z0, t, rcp) : mds_add<false>(0.f, fac, p[0], p[1], p[2], x0, y0, z0, t, rcp);
    }
    c.st.t[e] = w0;
    const unsigned rev = bs_log2 ? (__brev((unsigned)(k & bs_mask)) >> (32 - bs_log2)) : 0u;
    c.st.k[e] = ok ? ((rev << 21) | (unsigned)k) : 0xffffffffu;
    if (ok) c.st.loc[li] = (unsigned short)e;
  }
  int live = per - ((rank == 0 && per > 0) ? 1 : 0);
  if (tid == 0) {
    mbar_init(&sh.bars[0], 1);
    mbar_init(&sh.bars[1], 1);
    fence_mbar_init();
  }
  if (rank == 0 && tid == 0) idxs[0] = 0;
  for (int e = tid; e <= MDS_MAXK; e += THREADS) sh.picks[e] = make_float4(0.f, 0.f, 0.f, 0.f);  // tag 0 = never written
  __syncthreads();
  cluster_sync_all();  // peers must see initialised barriers before the first remote complete_tx
  c.dataset = dataset;
  c.idxs = idxs;
  c.sxyz = sxyz;
  c.sh = &sh;
  c.m = m;
  c.xs = xs;
  c.csh = csh;
  c.msel = msel;
  c.cs = cs;
  c.rank = rank;
  c.t = t;
  c.r = rcp;
  if (m > 1) {
    const bool replayer = mds_first(WORKERS) ? tid < 32 : tid >= WORKERS;
    const bool idle = mds_iso(WORKERS) && (tid >> 5) == 4;
    if (replayer) {
      if (fast) mds_replay<true>(c, (int)cs * (WORKERS / 32));
      else mds_replay<false>(c, (int)cs * (WORKERS / 32));
    } else if (!idle) {
      int gen = 0;
      if (fast) MdsChain<WORKERS, PT, true>::run(c, live, gen);
      else MdsChain<WORKERS, PT, false>::run(c, live, gen);
    }
  }
  __syncthreads();
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

template <int WORKERS, int PT, int OCC>
static int mds_launch(const float* xyz, int B, int n, int m, const float* mml, int* idx, int cs, int bs_mask, int bs_log2, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * cs));
  cfg.blockDim = dim3(mds_threads(WORKERS));
  const int per = (n + cs - 1) / cs;
  int stage_xyz = mds_smem_bytes(per, WORKERS, PT, true) + sizeof(MdsShared) <= (size_t)(OCC > 1 ? 100 : 220) * 1024 ? 1 : 0;
  size_t smem = mds_smem_bytes(per, WORKERS, PT, stage_xyz != 0);
  const int total_warps = cs * (WORKERS / 32);
  int msel = MDS_POOL / total_warps;
  if (msel > MDS_MAXM) msel = MDS_MAXM;
  if (const char* e = getenv("SNB_MDS_M")) {
    const int v = atoi(e);
    if (v >= 1 && v <= msel) msel = v;
  }
  while (msel & (msel - 1)) msel &= msel - 1;  // a power of two: 1, 2, 4 or 8
  cudaError_t ea = cudaFuncSetAttribute(mds_cluster_kernel<WORKERS, PT, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
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
  return (int)cudaLaunchKernelEx(&cfg, mds_cluster_kernel<WORKERS, PT, OCC>, xyz, n, m, mml, idx, bs_mask, bs_log2, stage_xyz, msel);
}

}  // namespace snb

using namespace snb;

#ifdef SNB_MDS_STATS
SNB_API int snb_mds_debug_stats(unsigned long long* out16, int reset) {
  SNB_CUDA(cudaMemcpyFromSymbol(out16, g_mds_stats, sizeof(unsigned long long) * 16));
  if (reset) {
    unsigned long long z[16] = {};
    SNB_CUDA(cudaMemcpyToSymbol(g_mds_stats, z, sizeof(z)));
  }
  return SNB_OK;
}
#endif

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
  while (per > 512 * 18 && cs < MDS_MAX_CLUSTER) {  // too many points for the register file: widen the cluster
    cs *= 2;
    per = (n + cs - 1) / cs;
  }
  int rc;
  const int bm = bs - 1;
#define MDS_GO(T, P, O) rc = mds_launch<T, P, O>(xyz, B, n, m, mean_mst_length, idx, cs, bm, lg, s)
  // Experiment hook (development): SNB_MDS_LAYOUT="<cluster size>,<worker threads>" picks another layout, e.g. "8,128":
  // 8 thin CTAs per sample, two samples' CTAs sharing an SM.
  int force_threads = 0;
  if (const char* e = getenv("SNB_MDS_LAYOUT")) {
    int a = 0, t = 0;
    if (sscanf(e, "%d,%d", &a, &t) == 2 && (a == 1 || a == 2 || a == 4 || a == 8) && (t == 128 || t == 192 || t == 224 || t == 256)) {
      cs = a;
      force_threads = t;
      per = (n + cs - 1) / cs;
    }
  }
  if (force_threads == 128) {  // two samples' CTAs per SM
    if (per <= 128 * 18) MDS_GO(128, 18, 2);
    else return SNB_ELIMIT;
  } else if (force_threads == 224) {
    if (per <= 224 * 9) MDS_GO(224, 9, 1);
    else if (per <= 224 * 21) MDS_GO(224, 21, 1);
    else return SNB_ELIMIT;
  } else if (force_threads == 192) {
    if (per <= 192 * 24) MDS_GO(192, 24, 1);
    else return SNB_ELIMIT;
  } else if (force_threads == 256 && per > 256 * 12 && per <= 256 * 18) MDS_GO(256, 18, 1);
  else if (per <= 256 * 2) MDS_GO(256, 2, 1);
  else if (per <= 256 * 4) MDS_GO(256, 4, 1);
  else if (per <= 256 * 6) MDS_GO(256, 6, 1);
  else if (per <= 256 * 9) MDS_GO(256, 9, 1);
  else if (per <= 256 * 12) MDS_GO(256, 12, 1);
  else if (per <= 224 * 21) MDS_GO(224, 21, 1);  // SpareNet's refiner (4608 points per CTA): replay warp beside ONE worker warp
  else if (per <= 512 * 12) MDS_GO(512, 12, 1);
  else if (per <= 512 * 18) MDS_GO(512, 18, 1);
  else return SNB_ELIMIT;  // n > 8*512*18 = 73728 points per sample
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
