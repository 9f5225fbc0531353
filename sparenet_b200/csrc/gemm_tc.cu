// gemm_tc.cu -- the dense 1x1-conv / AdaIN-folding stacks of the generator on the 5th-generation tensor cores (sm_100a).
//
// One persistent, warp-specialised kernel family:  D[g] (M x N) = A[g] (M x K) . T(B[g]) (K x N)   in TF32 with fp32 accumulation
//   * operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a 4-stage shared-memory ring,
//   * tcgen05.mma kind::tf32, 128 x block_n x 8 per instruction, issued by ONE thread, accumulators in TMEM (2 x 256 columns:
//     the epilogue of tile i overlaps the MMAs of tile i+1),
//   * optional PROLOGUE on the activation operand in shared memory: T(x) = leaky_relu(scale[row]*x + shift[row]) -- the folded
//     AdaIN.BN.SE.ReLU tail of the previous layer, so the activated tensor never exists in HBM,
//   * EPILOGUE straight from TMEM (tcgen05.ld): TMA store of the tile and/or per-row statistics of it (mean and centred second
//     moment per tile: no cancellation), row max/min with their positions, or a TMA reduce-add for split-K weight gradients.
// Three operand arrangements cover a 1x1 convolution on channel-major activations [G, C, Npos] (positions contiguous):
//   FWD   Y[g]  = W        . T(X[g])      A = W  K-major  [Cout, Cin],   B = X   MN-major [Cin rows, positions]
//   DGRAD gX[g] = W^T      . gY[g]        A = W  MN-major (Cin contiguous), B = gY  MN-major
//   WGRAD gW    = sum_b gY[b] . T(X[b])^T  A = gY K-major  [Cout, positions], B = X  K-major [Cin rows, positions]
// Reference layers served: models/sparenet_generator.py:146-186,188-242 (EdgeConv convs), :593-646 (PointNetRes),
// :984-991,1044-1062 (GridDecoder).  The reference runs them through cuDNN with TF32 allowed (torch default for convolutions).
#include <cuda.h>
#include <stdio.h>

#include "common.cuh"
#include "sparenet_b200.h"

namespace snb {
namespace {

constexpr int BM = 128;                 // tile rows (UMMA M, cta_group::1)
constexpr int BK = 32;                  // k elements per stage = one 128-byte swizzle row of fp32
constexpr int UMMA_K = 8;               // kind::tf32
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 4;    // 16 KB
constexpr int B_BYTES_MAX = 256 * BK * 4;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES_MAX;            // 48 KB
constexpr int STG_BYTES = 32 * 128;                            // one epilogue staging box: 32 rows x 32 fp32
constexpr int SMEM_STAGING = STAGES * STAGE_BYTES;             // 196608
constexpr int SMEM_BARS = SMEM_STAGING + 8 * STG_BYTES;        // 229376: one staging box per epilogue warp
constexpr int SMEM_TOTAL = SMEM_BARS + 256 + 1024;             // + barriers + alignment slack
constexpr int TMEM_COLS = 512;

enum : int { MODE_FWD = 0, MODE_DGRAD = 1, MODE_WGRAD = 2 };

struct KParams {
  int mode;
  int M, N, K;          // GEMM dims of one output batch; WGRAD: K = positions per inner batch
  int G;                // output batches
  int BI;               // inner (reduction) batches per output batch (WGRAD), else 1
  int a_batched;        // FWD/DGRAD: A has a batch dim (decoder: one weight per primitive)
  int block_n;          // 64, 128, 192 or 256; the two epilogue warpgroups take block_n/2 columns each (= one statistics tile)
  int mt, nt;           // tiles along M, N
  int split;            // split-K factor (WGRAD)
  int kb_total;         // k-blocks of one output batch = BI * ceil(K/BK)  (FWD/DGRAD: ceil(K/BK))
  int kb_per_batch;     // ceil(K/BK)
  int total_tiles;
  int b_mod;            // > 0: the activation operand holds only b_mod positions and is tiled along the position axis (decoder layer 1:
                        //      one lattice response per primitive, shared by all samples; only scale/shift differ per sample)
  // prologue
  const float* scale;
  const float* shift;
  float slope;
  int xf_rows;          // rows of the parameter table per activation batch (= input channels)
  int xf_S;             // segments per row: scale index = (batch*xf_rows + row)*xf_S + pos0/seg
  int seg;
  // epilogue
  int store;            // 0 none, 1 store, 2 reduce-add
  float* pmean;         // [G, M, 2*nt] per-half-tile row mean          (nullptr: off)
  float* pm2;           // [G, M, 2*nt] per-half-tile centred 2nd moment
  float* pmax;          // [G, M, 2*nt]                                   (nullptr: off)
  float* pmin;
  int* pimax;           // position inside the full row (n index)
  int* pimin;
};

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src), "r"(c0),
               "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_map(const CUtensorMap* m) { asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory"); }

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
      "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
        "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// mbarrier wait with a watchdog: a protocol bug traps after ~4 s instead of hanging the GPU (the fast path is one try_wait)
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __noinline__ void mbar_wait_slow(uint64_t* bar, uint32_t parity) {
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (!mbar_try(bar, parity)) {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 4000000000ull) {
      printf("gemm_tf32_kernel: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x,
             smem_u32(bar) & 0xfffu, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity) {
  if (!mbar_try(bar, parity)) mbar_wait_slow(bar, parity);
}

// shared-memory matrix descriptor (UMMA, sm_100): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) |
// base_offset=0 [49,52) | layout_type [61,64): 2 = SWIZZLE_128B (16-byte chunks ^ row%8; the K-major operands),
// 1 = SWIZZLE_128B_BASE32B (32-byte chunks ^ row%4) -- the ONLY layout the tensor core accepts for MN-major TF32 operands.
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) |
         (1ull << 46) | ((uint64_t)layout_type << 61);
}
// K-major stage [rows][128 B]: 8-row groups 1024 B apart, k-step j = 32 bytes along the swizzled row.
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t base, int j) { return make_desc(base + j * 32, 16, 1024, 2); }
// MN-major stage [chunk of 32 mn][32 k-rows][128 B]: chunks (LBO) 4096 B apart, 4-row swizzle atoms (SBO) 512 B apart,
// k-step j = 8 rows = 1024 bytes.
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t base, int j) { return make_desc(base + j * 1024, 32 * 128, 512, 1); }

struct TileCoord {
  int m0, n0, g, sp, kb0, kb1;
};
__device__ __forceinline__ TileCoord decode_tile(const KParams& p, int t) {
  TileCoord c;
  const int mi = t % p.mt;
  t /= p.mt;
  const int ni = t % p.nt;
  t /= p.nt;
  c.sp = t % p.split;
  c.g = t / p.split;
  c.m0 = mi * BM;
  c.n0 = ni * p.block_n;
  const int per = (p.kb_total + p.split - 1) / p.split;
  c.kb0 = c.sp * per;
  c.kb1 = min(p.kb_total, c.kb0 + per);
  return c;
}

// ---------------------------------------------------------------------------------------------- the kernel
constexpr int EPI_WARPS = 8, XF_WARPS = 8;
constexpr int THREADS_PLAIN = (2 + EPI_WARPS) * 32;             // 320
constexpr int THREADS_XFORM = (2 + EPI_WARPS + XF_WARPS) * 32;  // 576

__device__ __forceinline__ float4 xf4(float4 x, float a, float b, float slope) {
  // leaky_relu(a*x + b) = max(t, slope*t) for 0 <= slope <= 1 (the host checks the range)
  x.x = fmaf(x.x, a, b); x.y = fmaf(x.y, a, b); x.z = fmaf(x.z, a, b); x.w = fmaf(x.w, a, b);
  x.x = fmaxf(x.x, x.x * slope); x.y = fmaxf(x.y, x.y * slope); x.z = fmaxf(x.z, x.z * slope); x.w = fmaxf(x.w, x.w * slope);
  return x;
}

template <bool XFORM>
__global__ void __launch_bounds__(XFORM ? THREADS_XFORM : THREADS_PLAIN, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_d,
                 const KParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_BARS);
  uint64_t* full = bars;                  // [STAGES] TMA -> consumers
  uint64_t* ready = bars + STAGES;        // [STAGES] transform warps -> MMA (XFORM only)
  uint64_t* empty = bars + 2 * STAGES;    // [STAGES] MMA commit -> TMA
  uint64_t* tfull = bars + 3 * STAGES;    // [2] MMA commit -> epilogue
  uint64_t* tempty = tfull + 2;           // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_map(&map_a);
    prefetch_map(&map_b);
    prefetch_map(&map_d);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&ready[s], XF_WARPS / 2);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const bool a_mn = (p.mode == MODE_DGRAD);
  const bool b_mn = (p.mode != MODE_WGRAD);
  const uint32_t b_bytes = (uint32_t)p.block_n * 128u;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TileCoord c = decode_tile(p, t);
        for (int kb = c.kb0; kb < c.kb1; ++kb) {
          mbar_wait_wd(&empty[s], ph ^ 1u);
          mbar_expect_tx(&full[s], (uint32_t)A_BYTES + b_bytes);
          const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_BYTES, bar = smem_u32(&full[s]);
          if (p.mode == MODE_WGRAD) {
            const int bi = kb / p.kb_per_batch, kpos = (kb - bi * p.kb_per_batch) * BK, batch = c.g * p.BI + bi;
            tma_load_4d(sa, &map_a, bar, kpos, c.m0, batch, 0);
            tma_load_4d(sb, &map_b, bar, p.b_mod > 0 ? kpos % p.b_mod : kpos, c.n0, batch, 0);
          } else {
            const int ga = p.a_batched ? c.g : 0;
            if (a_mn) tma_load_4d(sa, &map_a, bar, 0, kb * BK, c.m0 >> 5, ga);
            else tma_load_4d(sa, &map_a, bar, kb * BK, c.m0, ga, 0);
            tma_load_4d(sb, &map_b, bar, 0, kb * BK, (p.b_mod > 0 ? c.n0 % p.b_mod : c.n0) >> 5, c.g);
          }
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (one thread)
    if (lane == 0) {
      // instruction descriptor: D=f32 [4,6)=1, A=tf32 [7,10)=2, B=tf32 [10,13)=2, a_major bit15, b_major bit16, N>>3 [17,23), M>>4 [24,29)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
                             ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int s = 0, as = 0;
      uint32_t ph = 0, aph = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TileCoord c = decode_tile(p, t);
        mbar_wait_wd(&tempty[as], aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * 256);
        uint32_t acc = 0;
        for (int kb = c.kb0; kb < c.kb1; ++kb) {
          mbar_wait_wd(&full[s], ph);
          if (XFORM) mbar_wait_wd(&ready[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
          for (int j = 0; j < BK / UMMA_K; ++j) {
            const uint64_t da = a_mn ? desc_mnmajor(sa, j) : desc_kmajor(sa, j);
            const uint64_t db = b_mn ? desc_mnmajor(sb, j) : desc_kmajor(sb, j);
            tc_mma_tf32(d_tmem, da, db, idesc, acc);
            acc = 1;
          }
          tc_commit(smem_u32(&empty[s]));      // frees the smem slot once these MMAs have read it
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        tc_commit(smem_u32(&tfull[as]));       // accumulator complete
        if (++as == 2) { as = 0; aph ^= 1u; }
      }
    }
  } else if (warp < 2 + EPI_WARPS) {
    // ================================================================ epilogue: TMEM -> registers -> (stats) -> smem -> TMA store
    // eight warps: lane quarter q = warp % 4 (the TMEM lanes a warp may touch), column half = (warp - 2) / 4
    const int q = warp & 3, half = (warp - 2) >> 2;
    const uint32_t stg = smem_base + SMEM_STAGING + (uint32_t)(warp - 2) * STG_BYTES;
    const int nch = p.block_n >> 6;                          // 32-column chunks per half
    const int sw = p.block_n >> 1;                           // columns per half = width of one statistics tile
    int as = 0;
    uint32_t aph = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const TileCoord c = decode_tile(p, t);
      const int row = c.m0 + q * 32 + lane;
      const int col0 = half * sw;                            // first column of this half inside the tile
      mbar_wait_wd(&tfull[as], aph);
      tc_fence_after();
      float mean = 0.f, m2 = 0.f, vmax = -INFINITY, vmin = INFINITY;
      int imax = 0, imin = 0;
      const bool has_k = c.kb1 > c.kb0;                      // an empty split contributes nothing
      for (int ch = 0; ch < nch; ++ch) {
        uint32_t v[32];
        tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 256 + col0 + ch * 32), v);
        tc_wait_ld();
        if (p.pmean != nullptr) {
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            s0 += __uint_as_float(v[i]);
            s1 += __uint_as_float(v[i + 1]);
            s2 += __uint_as_float(v[i + 2]);
            s3 += __uint_as_float(v[i + 3]);
          }
          const float mc = ((s0 + s1) + (s2 + s3)) * (1.f / 32.f);
          float q0 = 0.f, q1 = 0.f;
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float d0 = __uint_as_float(v[i]) - mc, d1 = __uint_as_float(v[i + 1]) - mc;
            q0 = fmaf(d0, d0, q0);
            q1 = fmaf(d1, d1, q1);
          }
          // Chan merge of (32*ch, mean, m2) with (32, mc, q0+q1)
          const float na = 32.f * ch, nb = 32.f, nab = na + nb, dl = mc - mean;
          mean += dl * (nb / nab);
          m2 += (q0 + q1) + dl * dl * (na * nb / nab);
        }
        if (p.pmax != nullptr) {
          // chunk extrema with 3-input min/max, then the (rare) position scan only when the running extremum moves
          float cx = __uint_as_float(v[0]), cn = cx;
#pragma unroll
          for (int i = 1; i < 31; i += 2) {
            cx = fmaxf(fmaxf(cx, __uint_as_float(v[i])), __uint_as_float(v[i + 1]));
            cn = fminf(fminf(cn, __uint_as_float(v[i])), __uint_as_float(v[i + 1]));
          }
          cx = fmaxf(cx, __uint_as_float(v[31]));
          cn = fminf(cn, __uint_as_float(v[31]));
          if (cx > vmax) {
            vmax = cx;
            int f = 31;
#pragma unroll
            for (int i = 30; i >= 0; --i) f = (__uint_as_float(v[i]) == cx) ? i : f;
            imax = ch * 32 + f;
          }
          if (cn < vmin) {
            vmin = cn;
            int f = 31;
#pragma unroll
            for (int i = 30; i >= 0; --i) f = (__uint_as_float(v[i]) == cn) ? i : f;
            imin = ch * 32 + f;
          }
        }
        if (p.store != 0 && has_k) {
          if (lane == 0) bulk_wait_read<0>();                // the previous store has finished reading this warp's staging box
          __syncwarp();
          const uint32_t dst = stg + (uint32_t)lane * 128u;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t a = dst + (uint32_t)((j ^ (lane & 7)) << 4);   // 128-byte swizzle: 16-byte chunk index ^ (row % 8)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v[4 * j]), "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                         : "memory");
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            if (p.store == 1) tma_store_3d(&map_d, stg, c.n0 + col0 + ch * 32, c.m0 + q * 32, c.g);
            else tma_reduce_add_3d(&map_d, stg, c.n0 + col0 + ch * 32, c.m0 + q * 32, c.g);
            bulk_commit();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      if (row < p.M) {
        const size_t o = ((size_t)c.g * p.M + row) * (2 * p.nt) + 2 * (c.n0 / p.block_n) + half;
        if (p.pmean != nullptr) { p.pmean[o] = mean; p.pm2[o] = m2; }
        if (p.pmax != nullptr) { p.pmax[o] = vmax; p.pmin[o] = vmin; p.pimax[o] = c.n0 + col0 + imax; p.pimin[o] = c.n0 + col0 + imin; }
      }
      if (++as == 2) { as = 0; aph ^= 1u; }
    }
    if (lane == 0) bulk_wait_all();
  } else if (XFORM) {
    // ================================================================ prologue: activation operand transformed in shared memory
    // Two groups of four warps take alternate k-blocks, so the load -> fma/max -> store -> proxy fence -> arrive chain of one stage
    // overlaps the next stage's.  Thread tw of a group owns the 128-byte rows tw and tw + 128 of the stage (8 + 8 vectors, visited
    // in a lane-rotated order so that a quarter-warp touches 8 different 16-byte columns: no bank conflicts).
    const int xw = threadIdx.x - (2 + EPI_WARPS) * 32;       // 0..255
    const int grp = xw >> 7, tw = xw & 127;
    const int nrow = (tw < p.block_n ? 1 : 0) + (tw + 128 < p.block_n ? 1 : 0);   // rows of this thread that exist in the stage
    int s = 0;
    uint32_t ph = 0, cnt = 0;
    float a0 = 0.f, b0 = 0.f, a1 = 0.f, b1 = 0.f;
    long long key = -1;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const TileCoord c = decode_tile(p, t);
      for (int kb = c.kb0; kb < c.kb1; ++kb, ++cnt) {
        if ((cnt & 1u) == (uint32_t)grp) {
          bool v0, v1;
          if (b_mn) {
            // MN-major stage [chunk][32 k-rows][32 positions]: rows tw and tw + 128 are both k-row tw % 32 -> ONE channel per thread
            const int chn = kb * BK + (tw & 31);
            v0 = v1 = chn < p.xf_rows;                       // k-rows past Cin stay zero (TMA fill): A's columns there are zero too
            const size_t o = ((size_t)c.g * p.xf_rows + (v0 ? chn : 0)) * p.xf_S + c.n0 / p.seg;
            a0 = a1 = __ldg(p.scale + o);
            b0 = b1 = __ldg(p.shift + o);
          } else {
            // K-major stage [channel rows][32 positions]: row r is channel n0 + r, constant over the k-blocks of a segment
            const int bi = kb / p.kb_per_batch, batch = c.g * p.BI + bi, segi = ((kb - bi * p.kb_per_batch) * BK) / p.seg;
            const long long k2 = ((long long)batch * p.xf_S + segi) * p.nt + (c.n0 / p.block_n);
            const int c0 = c.n0 + tw, c1 = c0 + 128;
            v0 = c0 < p.xf_rows;                             // rows past Cin keep the TMA zero fill
            v1 = c1 < p.xf_rows;
            if (k2 != key) {
              key = k2;
              const size_t o0 = ((size_t)batch * p.xf_rows + (v0 ? c0 : 0)) * p.xf_S + segi;
              const size_t o1 = ((size_t)batch * p.xf_rows + (v1 ? c1 : 0)) * p.xf_S + segi;
              a0 = __ldg(p.scale + o0); b0 = __ldg(p.shift + o0);
              a1 = __ldg(p.scale + o1); b1 = __ldg(p.shift + o1);
            }
          }
          mbar_wait_wd(&full[s], ph);
          uint8_t* base = smem + s * STAGE_BYTES + A_BYTES + tw * 128;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h < nrow && (h == 0 ? v0 : v1)) {
              uint8_t* rp = base + h * (128 * 128);
              const float a = h == 0 ? a0 : a1, b = h == 0 ? b0 : b1;
              float4 x[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) x[j] = *reinterpret_cast<const float4*>(rp + (((j + tw) & 7) << 4));
#pragma unroll
              for (int j = 0; j < 8; ++j) x[j] = xf4(x[j], a, b, p.slope);
#pragma unroll
              for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(rp + (((j + tw) & 7) << 4)) = x[j];
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ready[s]);
        }
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// rank-4 fp32 map with 128-byte swizzle.  dims/strides in elements (stride[0] is implicitly 1).
int encode4(CUtensorMap* m, const void* base, const uint64_t dims[4], const uint64_t strides_elem[4], const uint32_t box[4], int rank = 4,
            CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = get_encode();
  if (fn == nullptr) return SNB_EINVAL;
  cuuint64_t gd[4], gs[3];
  cuuint32_t bx[4], es[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; ++i) { gd[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 1; i < 4; ++i) {
    gs[i - 1] = strides_elem[i] * 4ull;
    if (gs[i - 1] == 0) gs[i - 1] = 16;
    if (gs[i - 1] % 16 != 0) return SNB_EALIGN;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0) return SNB_EALIGN;
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? SNB_OK : SNB_EINVAL;
}

// K-major operand [batch, rows, kdim] (kdim contiguous): dims (kdim, rows, batch, 1), box (32, box_rows, 1, 1)
int encode_kmajor(CUtensorMap* m, const float* base, int kdim, int rows, int batch, long long ld, long long batch_stride, int box_rows) {
  const uint64_t dims[4] = {(uint64_t)kdim, (uint64_t)rows, (uint64_t)batch, 1};
  const uint64_t bs = batch_stride > 0 ? (uint64_t)batch_stride : (uint64_t)rows * ld;
  const uint64_t st[4] = {1, (uint64_t)ld, bs, bs * (uint64_t)batch};
  const uint32_t box[4] = {32, (uint32_t)box_rows, 1, 1};
  return encode4(m, base, dims, st, box);
}
// MN-major operand [batch, krows, mn] (mn contiguous) seen as (32, krows, mn/32, batch): box (32, 32, chunks, 1) lands in shared memory
// as [chunk][32 k-rows][128 bytes] with 32-byte chunks swizzled by row % 4 (SWIZZLE_128B_ATOM_32B) = the canonical MN-major
// SWIZZLE_128B_BASE32B layout with LBO = 4096 (next 32 mn), SBO = 512 (next 4 k-rows) -- the only one TF32 accepts MN-major
int encode_mnmajor(CUtensorMap* m, const float* base, int mn, int krows, int batch, long long ld, long long batch_stride, int chunks) {
  if (mn % 32 != 0) return SNB_EINVAL;
  const uint64_t dims[4] = {32, (uint64_t)krows, (uint64_t)(mn / 32), (uint64_t)batch};
  const uint64_t bs = batch_stride > 0 ? (uint64_t)batch_stride : (uint64_t)krows * ld;
  const uint64_t st[4] = {1, (uint64_t)ld, 32, bs};
  const uint32_t box[4] = {32, 32, (uint32_t)chunks, 1};
  return encode4(m, base, dims, st, box, 4, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
}
int encode_out(CUtensorMap* m, const float* base, int cols, int rows, int batch, long long ld, long long batch_stride) {
  const uint64_t dims[4] = {(uint64_t)cols, (uint64_t)rows, (uint64_t)batch, 1};
  const uint64_t bs = batch_stride > 0 ? (uint64_t)batch_stride : (uint64_t)rows * ld;
  const uint64_t st[4] = {1, (uint64_t)ld, bs, bs * (uint64_t)batch};
  const uint32_t box[4] = {32, 32, 1, 1};
  return encode4(m, base, dims, st, box, 3);
}

}  // namespace
}  // namespace snb

// The C ABI takes a plain descriptor struct (include/sparenet_b200.h: snb_gemm_desc).
static int pick_block_n(int N, int block_n) {
  if (block_n > 0) return block_n;
  if (N <= 256) return ((N + 63) / 64) * 64;
  int best = 256, pad = ((N + 255) / 256) * 256;     // least padding among 256 / 192 / 128 columns per tile, larger tile on ties
  for (int bn = 192; bn >= 128; bn -= 64) {
    const int q = ((N + bn - 1) / bn) * bn;
    if (q < pad) { pad = q; best = bn; }
  }
  return best;
}

SNB_API int snb_gemm_tf32_block_n(int N, int block_n) { return pick_block_n(N, block_n); }

// number of statistics tiles along N: every column tile is reduced in two halves of block_n/2 columns
SNB_API int snb_gemm_tf32_tiles(int N, int block_n) {
  block_n = pick_block_n(N, block_n);
  return 2 * ((N + block_n - 1) / block_n);
}

SNB_API int snb_gemm_tf32(const snb_gemm_desc* d, void* stream) {
  using namespace snb;
  if (d == nullptr || d->A == nullptr || d->B == nullptr) return SNB_EINVAL;
  if (d->M <= 0 || d->N <= 0 || d->K <= 0 || d->G <= 0) return SNB_EINVAL;
  if (d->mode < MODE_FWD || d->mode > MODE_WGRAD) return SNB_EINVAL;
  if (d->store != 0 && d->D == nullptr) return SNB_EINVAL;
  const int BI = d->mode == MODE_WGRAD ? (d->BI > 0 ? d->BI : 1) : 1;
  const int block_n = pick_block_n(d->N, d->block_n);
  if (block_n > 256 || block_n % 64 != 0) return SNB_EINVAL;
  const bool stats = d->pmean != nullptr || d->pmax != nullptr;
  if (stats && (d->N % block_n) != 0) return SNB_EINVAL;     // statistics need full tiles
  if ((d->pmean == nullptr) != (d->pm2 == nullptr)) return SNB_EINVAL;
  if (d->pmax != nullptr && (d->pmin == nullptr || d->pimax == nullptr || d->pimin == nullptr)) return SNB_EINVAL;
  const bool xform = d->scale != nullptr;
  if (xform && (d->shift == nullptr || d->seg <= 0 || d->mode == MODE_DGRAD || !(d->slope >= 0.f && d->slope <= 1.f))) return SNB_EINVAL;

  KParams p{};
  p.mode = d->mode;
  p.M = d->M; p.N = d->N; p.K = d->K; p.G = d->G; p.BI = BI;
  p.a_batched = d->a_batch_stride != 0;
  p.block_n = block_n;
  p.mt = (d->M + BM - 1) / BM;
  p.nt = (d->N + block_n - 1) / block_n;
  p.kb_per_batch = (d->K + BK - 1) / BK;
  p.kb_total = p.kb_per_batch * BI;
  int split = d->split > 1 ? d->split : 1;
  if (d->mode != MODE_WGRAD) split = 1;
  if (d->mode == MODE_WGRAD && d->split == 0) {
    // auto: the persistent grid runs ceil(items / 148) waves of tiles.  Pick the smallest split-K whose last wave is (nearly) full
    // -- a weight gradient has few output tiles and a very long K (the positions), e.g. 16 tiles x 2048 k-blocks: split 10 gave
    // 160 items = 2 waves at 54 %, split 9 gives 144 items = one wave at 97 %.  At least 8 k-blocks per item; every split adds one
    // reduce-add of the (small) output.
    const int tiles = p.mt * p.nt * d->G;
    const int smax = min(kNumSMs, p.kb_total / 8 > 0 ? p.kb_total / 8 : 1);
    float best = 0.f;
    for (int s_ = 1; s_ <= smax; ++s_) {
      const long long items = (long long)tiles * s_;
      const float eff = (float)items / (float)(((items + kNumSMs - 1) / kNumSMs) * kNumSMs);
      if (eff > best + 1e-6f) {
        best = eff;
        split = s_;
      }
      if (eff >= 0.93f) break;
    }
  }
  if (split > 1 && d->store != 2) return SNB_EINVAL;
  p.split = split;
  p.total_tiles = p.mt * p.nt * d->G * split;
  p.scale = d->scale; p.shift = d->shift; p.slope = d->slope; p.seg = d->seg > 0 ? d->seg : 1;
  p.store = d->store;
  p.pmean = d->pmean; p.pm2 = d->pm2; p.pmax = d->pmax; p.pmin = d->pmin; p.pimax = d->pimax; p.pimin = d->pimin;

  CUtensorMap ma, mb, md;
  int rc;
  if (d->mode == MODE_FWD) {
    // A = W [Ga, M, K] K-major; B = X [G, K rows, N] MN-major
    p.xf_rows = d->K;
    p.xf_S = xform ? (d->N + p.seg - 1) / p.seg : 1;
    if (xform && (p.seg % block_n) != 0) return SNB_EINVAL;
    if ((rc = encode_kmajor(&ma, d->A, d->K, d->M, p.a_batched ? d->G : 1, d->lda, d->a_batch_stride, BM)) != SNB_OK) return rc;
    const int nb = d->b_pos_mod > 0 ? d->b_pos_mod : d->N;
    if (d->b_pos_mod > 0 && ((d->N % nb) != 0 || (nb % block_n) != 0)) return SNB_EINVAL;
    p.b_mod = d->b_pos_mod > 0 ? nb : 0;
    if ((rc = encode_mnmajor(&mb, d->B, nb, d->K, d->G, d->ldb, d->b_batch_stride, block_n / 32)) != SNB_OK) return rc;
  } else if (d->mode == MODE_DGRAD) {
    if (d->b_pos_mod > 0) return SNB_EINVAL;
    // A = W^T: W stored [Ga, K rows (Cout), M (Cin) contiguous] MN-major; B = gY [G, K rows, N] MN-major
    if ((rc = encode_mnmajor(&ma, d->A, d->M, d->K, p.a_batched ? d->G : 1, d->lda, d->a_batch_stride, BM / 32)) != SNB_OK) return rc;
    if ((rc = encode_mnmajor(&mb, d->B, d->N, d->K, d->G, d->ldb, d->b_batch_stride, block_n / 32)) != SNB_OK) return rc;
  } else {
    // A = gY [G*BI, M rows, K positions] K-major; B = X [G*BI, N rows, K positions] K-major
    p.xf_rows = d->N;
    p.xf_S = xform ? (d->K + p.seg - 1) / p.seg : 1;
    if (xform && (p.seg % BK) != 0) return SNB_EINVAL;
    if ((rc = encode_kmajor(&ma, d->A, d->K, d->M, d->G * BI, d->lda, d->a_batch_stride, BM)) != SNB_OK) return rc;
    const int kb_ = d->b_pos_mod > 0 ? d->b_pos_mod : d->K;
    if (d->b_pos_mod > 0 && (BI != 1 || (d->K % kb_) != 0 || (kb_ % BK) != 0)) return SNB_EINVAL;
    p.b_mod = d->b_pos_mod > 0 ? kb_ : 0;
    if ((rc = encode_kmajor(&mb, d->B, kb_, d->N, d->G * BI, d->ldb, d->b_batch_stride, block_n)) != SNB_OK) return rc;
  }
  if (d->store != 0) {
    if ((rc = encode_out(&md, d->D, d->N, d->M, d->G, d->ldd, d->d_batch_stride)) != SNB_OK) return rc;
  } else {
    md = ma;   // never dereferenced
  }

  int dev = 0, sms = kNumSMs;
  SNB_CUDA(cudaGetDevice(&dev));
  SNB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = p.total_tiles < sms ? p.total_tiles : sms;
  cudaStream_t st = (cudaStream_t)stream;
  if (xform) {
    SNB_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    gemm_tf32_kernel<true><<<grid, THREADS_XFORM, SMEM_TOTAL, st>>>(ma, mb, md, p);
  } else {
    SNB_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    gemm_tf32_kernel<false><<<grid, THREADS_PLAIN, SMEM_TOTAL, st>>>(ma, mb, md, p);
  }
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
