// chamfer_bvh.cu -- exact nearest-neighbour search with spatial pruning (the fast path of snb_chamfer_fwd), sm_100a.
//
// Same contract as the brute-force kernel (chamfer.cu; reference cuda/chamfer_dist/chamfer.cu:15-145): dist = min_j s(i,j) with
// s = fma(dz,dz,fma(dx,dx,dy*dy)), d* = ref_j - query_i, idx = smallest j attaining it -- and the same BITS, because the minimum
// of a set of floats does not depend on the order they are visited in and every candidate is evaluated with the identical
// expression; a whole cluster is skipped only when a conservative lower bound proves none of its points can beat OR TIE the
// running minimum.  What changes is the work: O(N log N) instead of O(N*M) distance evaluations.
//
//   build  (one CTA per (cloud, sample)): bounding box -> 30-bit Morton codes -> bitonic sort in shared memory (<= 16384 points,
//          128 KB) -> sorted points as float4 (x,y,z,original index), axis-aligned boxes of 32 consecutive points ("clusters")
//          and of 16 consecutive clusters ("super-clusters").
//   query  (one thread per query, queries taken in THEIR cloud's Morton order so a warp's lanes walk the same boxes): visit the
//          super-cluster / cluster with the smallest box distance first, then every other box whose lower bound (shrunk by 1e-5
//          relative, far more than the fp32 rounding of the bound) does not exceed the running minimum.
// Ties: candidates compare as (distance, original index), so the lowest index wins regardless of the visiting order.
#include <math.h>
#include "bvh.cuh"

namespace snb {

__global__ void __launch_bounds__(BVH_BUILD_THREADS) chamfer_bvh_build_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2, int N,
                                                                              int M, float4* __restrict__ ws, size_t per_sample_f4) {
  extern __shared__ __align__(16) unsigned long long keys[];
  __shared__ float red[6][32];
  const int cloud = blockIdx.x, b = blockIdx.y;
  const int n = cloud ? M : N;
  const float* __restrict__ p = (cloud ? xyz2 : xyz1) + (size_t)b * n * 3;
  float4* base = ws + (size_t)b * per_sample_f4 + (cloud ? bvh_cloud_floats4(N) : 0);
  const BvhView v = bvh_view(base, n);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // ---- bounding box ----
  float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
  for (int i = tid; i < n; i += BVH_BUILD_THREADS)
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const float t = p[i * 3 + c];
      lo[c] = fminf(lo[c], t);
      hi[c] = fmaxf(hi[c], t);
    }
#pragma unroll
  for (int c = 0; c < 3; c++) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
      hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
    }
    if (lane == 0) {
      red[c][warp] = lo[c];
      red[3 + c][warp] = hi[c];
    }
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < 3; c++) {
    float a = red[c][lane], bb = red[3 + c][lane];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o));
      bb = fmaxf(bb, __shfl_xor_sync(0xffffffffu, bb, o));
    }
    lo[c] = a;
    hi[c] = bb;
  }
  // ---- Morton keys (code << 32 | index); padding sorts last ----
  int npad = 1;
  while (npad < n) npad <<= 1;
  float sc[3];
#pragma unroll
  for (int c = 0; c < 3; c++) sc[c] = hi[c] > lo[c] ? 1023.0f / (hi[c] - lo[c]) : 0.f;
  for (int i = tid; i < npad; i += BVH_BUILD_THREADS) {
    unsigned long long k = ~0ull;
    if (i < n) {
      const unsigned cx = (unsigned)fminf(fmaxf((p[i * 3 + 0] - lo[0]) * sc[0], 0.f), 1023.f);
      const unsigned cy = (unsigned)fminf(fmaxf((p[i * 3 + 1] - lo[1]) * sc[1], 0.f), 1023.f);
      const unsigned cz = (unsigned)fminf(fmaxf((p[i * 3 + 2] - lo[2]) * sc[2], 0.f), 1023.f);
      const unsigned code = (spread10(cx) << 2) | (spread10(cy) << 1) | spread10(cz);
      k = ((unsigned long long)code << 32) | (unsigned)i;
    }
    keys[i] = k;
  }
  __syncthreads();
  // ---- bitonic sort in shared memory ----
  for (int k = 2; k <= npad; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < npad; i += BVH_BUILD_THREADS) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = keys[i], c2 = keys[ixj];
          if ((a > c2) == ((i & k) == 0)) {
            keys[i] = c2;
            keys[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  // ---- sorted points + cluster boxes (a warp per cluster) ----
  const float qnan = __int_as_float(0x7fc00000);
  for (int c = warp; c < v.nc; c += BVH_BUILD_THREADS / 32) {
    const int t = c * BVH_LEAF + lane;
    float x = qnan, y = qnan, z = qnan;
    int id = -1;
    if (t < n) {
      id = (int)(unsigned)keys[t];
      x = p[id * 3 + 0];
      y = p[id * 3 + 1];
      z = p[id * 3 + 2];
    }
    v.pts[t] = make_float4(x, y, z, __int_as_float(id));
    float l0 = t < n ? x : 3.4e38f, l1 = t < n ? y : 3.4e38f, l2 = t < n ? z : 3.4e38f;
    float h0 = t < n ? x : -3.4e38f, h1 = t < n ? y : -3.4e38f, h2 = t < n ? z : -3.4e38f;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      l0 = fminf(l0, __shfl_xor_sync(0xffffffffu, l0, o));
      l1 = fminf(l1, __shfl_xor_sync(0xffffffffu, l1, o));
      l2 = fminf(l2, __shfl_xor_sync(0xffffffffu, l2, o));
      h0 = fmaxf(h0, __shfl_xor_sync(0xffffffffu, h0, o));
      h1 = fmaxf(h1, __shfl_xor_sync(0xffffffffu, h1, o));
      h2 = fmaxf(h2, __shfl_xor_sync(0xffffffffu, h2, o));
    }
    if (lane == 0) {
      v.box[2 * c] = make_float4(l0, l1, l2, 0.f);
      v.box[2 * c + 1] = make_float4(h0, h1, h2, 0.f);
    }
  }
  __syncthreads();
  __threadfence_block();
  // ---- super-cluster boxes ----
  for (int s = tid; s < v.ns; s += BVH_BUILD_THREADS) {
    float4 l = make_float4(3.4e38f, 3.4e38f, 3.4e38f, 0.f), h = make_float4(-3.4e38f, -3.4e38f, -3.4e38f, 0.f);
    for (int c = s * BVH_FAN; c < (s + 1) * BVH_FAN && c < v.nc; c++) {
      const float4 a = v.box[2 * c], bb = v.box[2 * c + 1];
      l.x = fminf(l.x, a.x); l.y = fminf(l.y, a.y); l.z = fminf(l.z, a.z);
      h.x = fmaxf(h.x, bb.x); h.y = fmaxf(h.y, bb.y); h.z = fmaxf(h.z, bb.z);
    }
    v.sbox[2 * s] = l;
    v.sbox[2 * s + 1] = h;
  }
}

struct Best {
  float d;
  int i;
};

__device__ __forceinline__ void visit_cluster(const float4* __restrict__ pts, int c, float qx, float qy, float qz, Best& best) {
  const float4* __restrict__ p = pts + (size_t)c * BVH_LEAF;
#pragma unroll 8
  for (int t = 0; t < BVH_LEAF; t++) {
    const float4 r = p[t];
    const float d = sqdist3(__fsub_rn(r.x, qx), __fsub_rn(r.y, qy), __fsub_rn(r.z, qz));  // NaN padding never compares true
    const int id = __float_as_int(r.w);
    if (d < best.d || (d == best.d && id < best.i)) {
      best.d = d;
      best.i = id;
    }
  }
}

__device__ __forceinline__ void visit_super(const BvhView& R, int s, float qx, float qy, float qz, Best& best) {
  const int c0 = s * BVH_FAN, c1 = (c0 + BVH_FAN) < R.nc ? (c0 + BVH_FAN) : R.nc;
  // nearest cluster first: a tight bound early prunes the rest
  float lbmin = 3.4e38f;
  int cmin = c0;
  for (int c = c0; c < c1; c++) {
    const float lb = box_lb(R.box[2 * c], R.box[2 * c + 1], qx, qy, qz);
    if (lb < lbmin) {
      lbmin = lb;
      cmin = c;
    }
  }
  if (lbmin <= best.d) visit_cluster(R.pts, cmin, qx, qy, qz, best);
  for (int c = c0; c < c1; c++) {
    if (c == cmin) continue;
    if (box_lb(R.box[2 * c], R.box[2 * c + 1], qx, qy, qz) <= best.d) visit_cluster(R.pts, c, qx, qy, qz, best);
  }
}

__global__ void __launch_bounds__(128) chamfer_bvh_query_kernel(int N, int M, float4* __restrict__ ws, size_t per_sample_f4, float* __restrict__ dist1,
                                                                 float* __restrict__ dist2, int* __restrict__ idx1, int* __restrict__ idx2) {
  const int dir = blockIdx.z, b = blockIdx.y;
  const int nq = dir ? M : N, nr = dir ? N : M;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nq) return;
  float4* base = ws + (size_t)b * per_sample_f4;
  const BvhView Qv = bvh_view(base + (dir ? bvh_cloud_floats4(N) : 0), nq);
  const BvhView R = bvh_view(base + (dir ? 0 : bvh_cloud_floats4(N)), nr);
  const float4 q = Qv.pts[t];  // queries in Morton order of their own cloud
  const int qid = __float_as_int(q.w);
  Best best;
  best.d = __int_as_float(0x7f800000);
  best.i = 0x7fffffff;
  float lbmin = 3.4e38f;
  int smin = 0;
  for (int s = 0; s < R.ns; s++) {
    const float lb = box_lb(R.sbox[2 * s], R.sbox[2 * s + 1], q.x, q.y, q.z);
    if (lb < lbmin) {
      lbmin = lb;
      smin = s;
    }
  }
  visit_super(R, smin, q.x, q.y, q.z, best);
  for (int s = 0; s < R.ns; s++) {
    if (s == smin) continue;
    if (box_lb(R.sbox[2 * s], R.sbox[2 * s + 1], q.x, q.y, q.z) <= best.d) visit_super(R, s, q.x, q.y, q.z, best);
  }
  float* __restrict__ dout = (dir ? dist2 : dist1) + (size_t)b * nq;
  int* __restrict__ iout = (dir ? idx2 : idx1) + (size_t)b * nq;
  dout[qid] = best.d;
  iout[qid] = best.i;
}

// host side, called from snb_chamfer_fwd (chamfer.cu)
size_t chamfer_bvh_workspace_bytes(int B, int N, int M) {
  if (N > BVH_MAXN || M > BVH_MAXN || N < 256 || M < 256) return 0;
  return (size_t)B * (bvh_cloud_floats4(N) + bvh_cloud_floats4(M)) * sizeof(float4);
}

int bvh_build_launch(const float* xyz1, const float* xyz2, int B, int N, int M, float4* ws, size_t per_sample_f4, cudaStream_t s) {
  const int nmax = N > M ? N : M;
  int npad = 1;
  while (npad < nmax) npad <<= 1;
  const size_t smem = (size_t)npad * sizeof(unsigned long long);
  // per device/context and cheap: set before every launch (a process-wide flag would leave the other GPUs of one process without it)
  SNB_CUDA(cudaFuncSetAttribute(chamfer_bvh_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(BVH_MAXN * sizeof(unsigned long long))));
  chamfer_bvh_build_kernel<<<dim3(2, B), BVH_BUILD_THREADS, smem, s>>>(xyz1, xyz2, N, M, ws, per_sample_f4);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

int chamfer_bvh_launch(const float* xyz1, const float* xyz2, int B, int N, int M, float* dist1, float* dist2, int* idx1, int* idx2, void* workspace,
                       cudaStream_t s) {
  const size_t per = bvh_cloud_floats4(N) + bvh_cloud_floats4(M);
  const int nmax = N > M ? N : M;
  const int rc = bvh_build_launch(xyz1, xyz2, B, N, M, (float4*)workspace, per, s);
  if (rc != SNB_OK) return rc;
  chamfer_bvh_query_kernel<<<dim3((nmax + 127) / 128, B, 2), 128, 0, s>>>(N, M, (float4*)workspace, per, dist1, dist2, idx1, idx2);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

}  // namespace snb
