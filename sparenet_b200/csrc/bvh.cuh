// bvh.cuh -- the Morton-ordered two-level box hierarchy shared by the Chamfer search (chamfer_bvh.cu) and the EMD auction (emd.cu).
#pragma once
#include "common.cuh"

namespace snb {

constexpr int BVH_MAXN = 16384;
constexpr int BVH_BUILD_THREADS = 1024;
constexpr int BVH_LEAF = 32;    // points per cluster
constexpr int BVH_FAN = 16;     // clusters per super-cluster

struct BvhView {
  float4* pts;    // [npad32] sorted points, w = original index bits; padding rows hold NaN coordinates
  float4* box;    // [2*nc]   cluster boxes (lo, hi)
  float4* sbox;   // [2*ns]   super-cluster boxes
  int n, nc, ns;
};

__host__ __device__ inline int bvh_nc(int n) { return (n + BVH_LEAF - 1) / BVH_LEAF; }
__host__ __device__ inline int bvh_ns(int n) { return (bvh_nc(n) + BVH_FAN - 1) / BVH_FAN; }
__host__ __device__ inline size_t bvh_cloud_floats4(int n) { return (size_t)bvh_nc(n) * BVH_LEAF + 2 * (size_t)bvh_nc(n) + 2 * (size_t)bvh_ns(n); }

__device__ __forceinline__ BvhView bvh_view(float4* base, int n) {
  BvhView v;
  v.n = n;
  v.nc = bvh_nc(n);
  v.ns = bvh_ns(n);
  v.pts = base;
  v.box = base + (size_t)v.nc * BVH_LEAF;
  v.sbox = v.box + 2 * (size_t)v.nc;
  return v;
}

__device__ __forceinline__ unsigned spread10(unsigned v) {  // 10 bits -> every third bit
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

__device__ __forceinline__ float box_lb(const float4 lo, const float4 hi, float qx, float qy, float qz) {
  const float dx = fmaxf(fmaxf(lo.x - qx, qx - hi.x), 0.f);
  const float dy = fmaxf(fmaxf(lo.y - qy, qy - hi.y), 0.f);
  const float dz = fmaxf(fmaxf(lo.z - qz, qz - hi.z), 0.f);
  return (dx * dx + dy * dy + dz * dz) * 0.99999f;  // shrunk: never above the computed distance of any point inside the box
}

// Builds the hierarchies of xyz1 [B,N,3] and xyz2 [B,M,3] into ws (per sample: bvh_cloud_floats4(N) + bvh_cloud_floats4(M) float4,
// cloud 1 first).  Defined in chamfer_bvh.cu.
int bvh_build_launch(const float* xyz1, const float* xyz2, int B, int N, int M, float4* ws, size_t per_sample_f4, cudaStream_t s);

}  // namespace snb
