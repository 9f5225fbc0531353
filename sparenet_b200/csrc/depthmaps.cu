// depthmaps.cu -- the fused multi-view depth-map renderer: view transform + depth normalisation + max-splat in one C-ABI call.
//
// Replaces ComputeDepthMaps.forward (utils/p2i_utils.py:211-252) + p2i(reduce="max") (cuda/p2i_op/__init__.py:99-131,
// p2i_max.h:7-143) for one view and one radius:
//   q = M [p;1], pos = q.xyz / q.w                       (transform :153-165; M = projection . look_at, row-major 4x4)
//   (row, col) = ((-pos.y, pos.x) + 1) / 2 * (H-1, W-1)  (:225 and p2i_op/__init__.py:116-121)
//   feat = 1 - (pos.z - zmin) / (zmax - zmin)            zmin / zmax over ALL points of the call (:226)
//   out = max(0, max_points feat * w(r)),  w = 0.5 + 0.5 cos(pi r / R) on the integer footprint r <= R (p2i_max.h:39-63)
// The reference does this with ~12 PyTorch ops per view (a [B*N,4,4] matrix expanded on the HOST and uploaded, a batched
// matmul, two global reductions, concatenations) around a splat that spins on a per-pixel lock.  Here: one transform kernel
// that also reduces (zmin, argmin) / (zmax, argmax) with 64-bit atomics, the lock-free packed-atomicMax splat of p2i.cu with the
// feature formed on the fly, one unpack kernel.  The backward is the exact chain rule of the same graph, including the
// gradient that reaches the two extreme points through zmin / zmax (first index on ties, like torch.min/max of a flat tensor).
//
// Workspace (caller-owned, must survive from _fwd to _bwd): header | per-point (row, col, z, w) | per-point gradient accumulators |
// packed cells.
#include <math.h>

#include "common.cuh"

namespace snb {
namespace {

constexpr double DM_PI = 3.14159265358979323846;

struct DmView {
  float m[16];
};

struct DmHeader {
  unsigned long long zmin_key;   // (float_key(z) << 32) | index          -> atomicMin: smallest z, then smallest index
  unsigned long long zmax_key;   // (float_key(z) << 32) | ~index         -> atomicMax: largest z, then smallest index
  float gzmin, gzmax;            // backward: gradient w.r.t. zmin / zmax
  float pad[2];
};

__device__ __forceinline__ float dm_weight(float r, float radius) { return (float)(cos((double)r * DM_PI / (double)radius) * 0.5 + 0.5); }
__device__ __forceinline__ int dm_clamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__global__ void __launch_bounds__(256) dm_init_kernel(DmHeader* __restrict__ hdr, unsigned long long* __restrict__ cell, size_t total) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    hdr->zmin_key = ~0ull;
    hdr->zmax_key = 0ull;
    hdr->gzmin = hdr->gzmax = 0.f;
  }
  if (i < total) cell[i] = ((unsigned long long)float_key(0.f) << 32) | 0xffffffffull;   // background 0 wins ties against points
}

__global__ void __launch_bounds__(256) dm_transform_kernel(const float* __restrict__ data, int n, DmView V, int H, int W, float4* __restrict__ pt,
                                                            DmHeader* __restrict__ hdr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long kmin = ~0ull, kmax = 0ull;
  if (i < n) {
    const float x = data[3 * (size_t)i], y = data[3 * (size_t)i + 1], z = data[3 * (size_t)i + 2];
    float q[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) q[r] = __fmaf_rn(V.m[4 * r + 2], z, __fmaf_rn(V.m[4 * r + 1], y, __fmaf_rn(V.m[4 * r], x, V.m[4 * r + 3])));
    const float px = q[0] / q[3], py = q[1] / q[3], pz = q[2] / q[3];
    const float row = __fmul_rn(__fmul_rn(__fadd_rn(-py, 1.f), 0.5f), (float)(H - 1));
    const float col = __fmul_rn(__fmul_rn(__fadd_rn(px, 1.f), 0.5f), (float)(W - 1));
    pt[i] = make_float4(row, col, pz, q[3]);
    const unsigned long long zk = (unsigned long long)float_key(pz) << 32;
    kmin = zk | (unsigned)i;
    kmax = zk | (unsigned)(~(unsigned)i);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const unsigned long long a = __shfl_xor_sync(0xffffffffu, kmin, o), b = __shfl_xor_sync(0xffffffffu, kmax, o);
    kmin = a < kmin ? a : kmin;
    kmax = b > kmax ? b : kmax;
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&hdr->zmin_key, kmin);
    atomicMax(&hdr->zmax_key, kmax);
  }
}

__device__ __forceinline__ float dm_feature(float z, float zmin, float zmax) {
  return __fsub_rn(1.f, __fdiv_rn(__fsub_rn(z, zmin), __fsub_rn(zmax, zmin)));
}

// one warp per point, lanes sweep the footprint; ONE packed 64-bit atomicMax per surviving hit (see p2i.cu for the filter's bound)
__global__ void __launch_bounds__(256) dm_splat_kernel(const float4* __restrict__ pt, const DmHeader* __restrict__ hdr, int n, int N, int H, int W,
                                                        float radius, unsigned long long* __restrict__ cell) {
  const int p = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (p >= n) return;
  const float zmin = key_float((unsigned)(hdr->zmin_key >> 32)), zmax = key_float((unsigned)(hdr->zmax_key >> 32));
  const float4 v = pt[p];
  const float py = v.x, px = v.y;
  const float f = dm_feature(v.z, zmin, zmax);
  const int b = p / N;
  const int x0 = dm_clamp((int)floorf(px - radius), 0, W - 1), x1 = dm_clamp((int)ceilf(px + radius), 0, W - 1);
  const int y0 = dm_clamp((int)floorf(py - radius), 0, H - 1), y1 = dm_clamp((int)ceilf(py + radius), 0, H - 1);
  const int bh = y1 - y0 + 1, npix = (x1 - x0 + 1) * bh;
  const float inv2r = 1.5707963267948966f / radius;
  unsigned long long* __restrict__ img = cell + (size_t)b * H * W;
  for (int t = lane; t < npix; t += 32) {
    const int xo = t / bh;
    const int x = x0 + xo, y = y0 + (t - xo * bh);
    const float dx = (float)x - px, dy = (float)y - py;
    const float r = sqrtf(__fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
    if (!(r <= radius)) continue;
    const float ch = fabsf(__cosf(r * inv2r)) + 2e-6f;
    const float bound = f >= 0.f ? f * (ch * ch * 1.0001f) : 0.f;
    unsigned long long* cp = img + (size_t)y * W + x;
    const float cur = key_float((unsigned)(*((volatile unsigned long long*)cp) >> 32));
    if (bound < cur) continue;                              // cannot beat (or tie) the cell: skip the fp64 cosine
    const float val = __fmul_rn(f, dm_weight(r, radius));
    atomicMax(cp, ((unsigned long long)float_key(val) << 32) | (unsigned long long)(0xfffffffeu - (unsigned)p));
  }
}

__global__ void __launch_bounds__(256) dm_finish_kernel(const unsigned long long* __restrict__ cell, size_t total, float* __restrict__ out,
                                                         int* __restrict__ ids) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const unsigned long long c = cell[i];
  out[i] = key_float((unsigned)(c >> 32));
  const unsigned pr = (unsigned)c;
  ids[i] = pr == 0xffffffffu ? -1 : (int)(0xfffffffeu - pr);
}

// ---- backward ------------------------------------------------------------------------------------------------------------------
// per pixel: route the gradient to the winning point (p2i_max.h:68-143): gacc[p] += (k*dy, k*dx, g*w)
__global__ void __launch_bounds__(256) dm_bwd_pixel_kernel(const float* __restrict__ gout, const int* __restrict__ ids, const float4* __restrict__ pt,
                                                            const DmHeader* __restrict__ hdr, int H, int W, size_t total, float radius,
                                                            float4* __restrict__ gacc) {
  const size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= total) return;
  const int p = ids[o];
  if (p < 0) return;
  const float g = gout[o];
  const int x = (int)(o % W), y = (int)((o / W) % H);
  const float zmin = key_float((unsigned)(hdr->zmin_key >> 32)), zmax = key_float((unsigned)(hdr->zmax_key >> 32));
  const float4 v = pt[p];
  const float dx = (float)x - v.y, dy = (float)y - v.x;
  const float r = sqrtf(__fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
  const float w = dm_weight(r, radius);
  const float f = dm_feature(v.z, zmin, zmax);
  const float rr = r > 1e-10f ? r : 1e-10f;
  const float k = (float)((double)(g * f) * sin((double)r * DM_PI / (double)radius) * 0.5 * DM_PI / (double)radius / (double)rr);
  float* a = reinterpret_cast<float*>(&gacc[p]);
  atomicAdd(a + 0, k * dy);
  atomicAdd(a + 1, k * dx);
  atomicAdd(a + 2, g * w);
}

// per point: chain (row, col, feat) gradients back to the cloud; accumulates the gradients of zmin / zmax
__global__ void __launch_bounds__(256) dm_bwd_point_kernel(const float* __restrict__ data, int n, DmView V, int H, int W, const float4* __restrict__ pt,
                                                            const float4* __restrict__ gacc, DmHeader* __restrict__ hdr,
                                                            float* __restrict__ gdata) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float smin = 0.f, smax = 0.f;
  if (i < n) {
    const float zmin = key_float((unsigned)(hdr->zmin_key >> 32)), zmax = key_float((unsigned)(hdr->zmax_key >> 32));
    const float inv = 1.f / (zmax - zmin);
    const float4 v = pt[i];
    const float4 g = gacc[i];
    // feat = 1 - (z - zmin) * inv:  d/dz = -inv,  d/dzmin = (zmax - z) * inv^2,  d/dzmax = (z - zmin) * inv^2
    smin = g.z * (zmax - v.z) * inv * inv;
    smax = g.z * (v.z - zmin) * inv * inv;
    const float dpx = g.y * 0.5f * (float)(W - 1);         // col = (pos.x + 1)/2 * (W-1)
    const float dpy = -g.x * 0.5f * (float)(H - 1);        // row = (-pos.y + 1)/2 * (H-1)
    const float dpz = -g.z * inv;
    const float x = data[3 * (size_t)i], y = data[3 * (size_t)i + 1], z = data[3 * (size_t)i + 2];
    float q[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) q[r] = __fmaf_rn(V.m[4 * r + 2], z, __fmaf_rn(V.m[4 * r + 1], y, __fmaf_rn(V.m[4 * r], x, V.m[4 * r + 3])));
    const float iw = 1.f / q[3];
    const float dq0 = dpx * iw, dq1 = dpy * iw, dq2 = dpz * iw;
    const float dq3 = -(dq0 * q[0] + dq1 * q[1] + dq2 * q[2]) * iw;
#pragma unroll
    for (int c = 0; c < 3; ++c) gdata[3 * (size_t)i + c] = V.m[c] * dq0 + V.m[4 + c] * dq1 + V.m[8 + c] * dq2 + V.m[12 + c] * dq3;
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    smin += __shfl_xor_sync(0xffffffffu, smin, o);
    smax += __shfl_xor_sync(0xffffffffu, smax, o);
  }
  if ((threadIdx.x & 31) == 0 && (smin != 0.f || smax != 0.f)) {
    atomicAdd(&hdr->gzmin, smin);
    atomicAdd(&hdr->gzmax, smax);
  }
}

// the gradients of zmin / zmax reach the cloud through the two extreme points (pos.z of the argmin / argmax)
__global__ void dm_bwd_extreme_kernel(const float* __restrict__ data, DmView V, const DmHeader* __restrict__ hdr, float* __restrict__ gdata) {
  const int which = threadIdx.x;   // 0: argmin, 1: argmax
  if (which > 1) return;
  const unsigned i = which == 0 ? (unsigned)hdr->zmin_key : ~(unsigned)hdr->zmax_key;
  const float dpz = which == 0 ? hdr->gzmin : hdr->gzmax;
  const float x = data[3 * (size_t)i], y = data[3 * (size_t)i + 1], z = data[3 * (size_t)i + 2];
  float q[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) q[r] = __fmaf_rn(V.m[4 * r + 2], z, __fmaf_rn(V.m[4 * r + 1], y, __fmaf_rn(V.m[4 * r], x, V.m[4 * r + 3])));
  const float iw = 1.f / q[3];
  const float dq2 = dpz * iw, dq3 = -(dq2 * q[2]) * iw;
  for (int c = 0; c < 3; ++c) atomicAdd(&gdata[3 * (size_t)i + c], V.m[8 + c] * dq2 + V.m[12 + c] * dq3);   // argmin may equal argmax (n == 1)
}

struct DmLayout {
  size_t pt, gacc, cell, total;
};
DmLayout dm_layout(int B, int N, int H, int W) {
  DmLayout l;
  const size_t n = (size_t)B * N;
  l.pt = 64;
  l.gacc = l.pt + n * sizeof(float4);
  l.cell = l.gacc + n * sizeof(float4);
  l.total = l.cell + (size_t)B * H * W * sizeof(unsigned long long);
  return l;
}

int dm_check(int B, int N, int H, int W, double radius) {
  if (B <= 0 || N <= 0 || H <= 0 || W <= 0 || !(radius > 0.0)) return SNB_EINVAL;
  if ((long long)B * N > 0x7ffffff0LL || (long long)B * H * W > 0x7ffffff0LL) return SNB_ELIMIT;
  return SNB_OK;
}

}  // namespace
}  // namespace snb

SNB_API size_t snb_depthmaps_workspace_bytes(int B, int N, int H, int W) {
  if (B <= 0 || N <= 0 || H <= 0 || W <= 0) return 0;
  return snb::dm_layout(B, N, H, W).total;
}

SNB_API int snb_depthmaps_fwd(const float* data, int B, int N, const float* view_matrix, int H, int W, double radius, float* out, int* ids,
                              void* workspace, size_t workspace_bytes, void* stream) {
  using namespace snb;
  int rc = dm_check(B, N, H, W, radius);
  if (rc) return rc;
  if (!data || !view_matrix || !out || !ids) return SNB_EINVAL;
  const DmLayout l = dm_layout(B, N, H, W);
  if (!workspace || workspace_bytes < l.total) return SNB_EWORKSPACE;
  if (((uintptr_t)workspace & 15) != 0) return SNB_EALIGN;
  cudaStream_t s = (cudaStream_t)stream;
  DmView V;
  for (int i = 0; i < 16; ++i) V.m[i] = view_matrix[i];
  char* ws = (char*)workspace;
  DmHeader* hdr = (DmHeader*)ws;
  float4* pt = (float4*)(ws + l.pt);
  unsigned long long* cell = (unsigned long long*)(ws + l.cell);
  const int n = B * N;
  const size_t total = (size_t)B * H * W;
  dm_init_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(hdr, cell, total);
  dm_transform_kernel<<<(n + 255) / 256, 256, 0, s>>>(data, n, V, H, W, pt, hdr);
  dm_splat_kernel<<<(unsigned)(((size_t)n * 32 + 255) / 256), 256, 0, s>>>(pt, hdr, n, N, H, W, (float)radius, cell);
  dm_finish_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(cell, total, out, ids);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

SNB_API int snb_depthmaps_bwd(const float* grad_out, const int* ids, const float* data, int B, int N, const float* view_matrix, int H, int W,
                              double radius, void* workspace, size_t workspace_bytes, float* grad_data, void* stream) {
  using namespace snb;
  int rc = dm_check(B, N, H, W, radius);
  if (rc) return rc;
  if (!grad_out || !ids || !data || !view_matrix || !grad_data) return SNB_EINVAL;
  const DmLayout l = dm_layout(B, N, H, W);
  if (!workspace || workspace_bytes < l.total) return SNB_EWORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  DmView V;
  for (int i = 0; i < 16; ++i) V.m[i] = view_matrix[i];
  char* ws = (char*)workspace;
  DmHeader* hdr = (DmHeader*)ws;
  float4* pt = (float4*)(ws + l.pt);
  float4* gacc = (float4*)(ws + l.gacc);
  const int n = B * N;
  const size_t total = (size_t)B * H * W;
  SNB_CUDA(cudaMemsetAsync(gacc, 0, (size_t)n * sizeof(float4), s));
  SNB_CUDA(cudaMemsetAsync(&hdr->gzmin, 0, 2 * sizeof(float), s));
  dm_bwd_pixel_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(grad_out, ids, pt, hdr, H, W, total, (float)radius, gacc);
  dm_bwd_point_kernel<<<(n + 255) / 256, 256, 0, s>>>(data, n, V, H, W, pt, gacc, hdr, grad_data);
  dm_bwd_extreme_kernel<<<1, 32, 0, s>>>(data, V, hdr, grad_data);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
