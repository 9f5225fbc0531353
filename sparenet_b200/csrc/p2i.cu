// p2i.cu -- differentiable point -> image splat ("max" and "sum" reduce), float32 and float64, sm_100a.
//
// Replaces p2i_{max,sum}_{forward,backward}_kernel (cuda/p2i_op/p2i_max.h:7-143, p2i_sum.h:7-131,
// footprint utility.h:82-100, launcher common.h:95-128).  Contract:
//   footprint: integer pixels in [clamp(floor(p-R)), clamp(ceil(p+R))]^2 with r = sqrt(fma(dy,dy,dx*dx)) <= R,
//   weight  w = (T)(cos((double)r*pi/(double)R)*0.5+0.5)  (fp64 math even for T=float, p2i_max.h:48),
//   max:  out = max(background, max f*w) with strict '<' (p2i_max.h:56), ids = winner (lowest point id on ties),
//   sum:  out = background + sum f*w.
//
// Design: the reference serialises every pixel hit behind a global CAS spin-lock.  Here a warp owns a point
// and its lanes sweep the footprint; "max" is ONE 64-bit atomicMax on a packed (ordered value bits, ~id)
// word per hit -- value and winner id update together, no lock, deterministic winner.  Before paying for
// the fp64 cosine a hit is filtered with an fp32 upper bound of f*w against the (possibly stale, which is
// safe: the cell only grows) current cell value.  "sum" is a plain atomicAdd.  float64 takes a two-pass
// (value, then id) route because value + id no longer fit one 64-bit word.
#include <math.h>
#include "common.cuh"

namespace snb {

constexpr double P2I_PI = 3.14159265358979323846;  // M_PI

template <typename T> struct P2I;
template <> struct P2I<float> {
  static __device__ __forceinline__ float sqrt_(float v) { return sqrtf(v); }
  static __device__ __forceinline__ float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
  static __device__ __forceinline__ float mul_(float a, float b) { return __fmul_rn(a, b); }
};
template <> struct P2I<double> {
  static __device__ __forceinline__ double sqrt_(double v) { return sqrt(v); }
  static __device__ __forceinline__ double fma_(double a, double b, double c) { return __fma_rn(a, b, c); }
  static __device__ __forceinline__ double mul_(double a, double b) { return __dmul_rn(a, b); }
};

template <typename T>
__device__ __forceinline__ T p2i_weight(T r, T radius) {
  return (T)(cos((double)r * P2I_PI / (double)radius) * 0.5 + 0.5);
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

template <typename T>
struct Footprint {
  int x0, y0, bw, bh;
  __device__ __forceinline__ Footprint(T py, T px, T radius, int H, int W) {
    x0 = clampi((int)floor(px - radius), 0, W - 1);
    const int x1 = clampi((int)ceil(px + radius), 0, W - 1);
    y0 = clampi((int)floor(py - radius), 0, H - 1);
    const int y1 = clampi((int)ceil(py + radius), 0, H - 1);
    bw = x1 - x0 + 1;
    bh = y1 - y0 + 1;
  }
};

__device__ __forceinline__ unsigned long long double_key(double d) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(d);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_double(unsigned long long k) {
  return __longlong_as_double((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
}

// ------------------------------------------------------------------------------------------ max, float32
// cell = (float_key(value) << 32) | prio ; prio = 0xFFFFFFFF for the background, 0xFFFFFFFE - id for points:
// larger value wins; equal value: background beats points (strict '<'), lower point id beats higher.
__global__ void __launch_bounds__(256) p2i_max_init_f32(const float* __restrict__ bg, size_t total, unsigned long long* __restrict__ cell) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) cell[i] = ((unsigned long long)float_key(bg[i]) << 32) | 0xffffffffull;
}

__global__ void __launch_bounds__(256) p2i_max_splat_f32(const float* __restrict__ points, const float* __restrict__ feat,
                                                          const int* __restrict__ binds, int npoints, int B, int C, int H, int W, float radius,
                                                          unsigned long long* __restrict__ cell) {
  const int p = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (p >= npoints) return;
  const int b = binds[p];
  if (b < 0 || b >= B) return;
  const float py = points[p * 2 + 0], px = points[p * 2 + 1];
  const Footprint<float> fp(py, px, radius, H, W);
  const int npix = fp.bw * fp.bh;
  const float inv2r = 1.5707963267948966f / radius;
  for (int t = lane; t < npix; t += 32) {
    const int xo = t / fp.bh;
    const int x = fp.x0 + xo, y = fp.y0 + (t - xo * fp.bh);
    const float dx = (float)x - px, dy = (float)y - py;
    const float r = sqrtf(__fmaf_rn(dy, dy, __fmul_rn(dx, dx)));  // x is the outer loop of utility.h:90-99: dx*dx is hoisted, dy fused
    if (!(r <= radius)) continue;
    // fp32 upper bound of w = cos^2(pi r / 2R): __cosf is within 2^-21.4 absolute on [0, pi/2] and the argument
    // within ~2e-7, so |cos| + 2e-6 bounds the true half-angle cosine; squared and padded by 1e-4 relative.
    const float ch = fabsf(__cosf(r * inv2r)) + 2e-6f;
    const float w_hi = ch * ch * 1.0001f;
    float w = -1.f;  // exact weight, computed lazily
    for (int c = 0; c < C; c++) {
      const float f = feat[(size_t)p * C + c];
      unsigned long long* cp = &cell[(((size_t)b * C + c) * H + y) * W + x];
      const float cur = key_float((unsigned)(*((volatile unsigned long long*)cp) >> 32));
      const float bound = f >= 0.f ? f * w_hi : 0.f;  // f<0: f*w <= 0
      if (bound < cur) continue;                        // cannot beat (or tie) the cell: skip the fp64 cosine
      if (w < 0.f) w = p2i_weight<float>(r, radius);
      const float v = __fmul_rn(f, w);
      const unsigned long long cand = ((unsigned long long)float_key(v) << 32) | (unsigned long long)(0xfffffffeu - (unsigned)p);
      atomicMax(cp, cand);
    }
  }
}

__global__ void __launch_bounds__(256) p2i_max_finish_f32(const unsigned long long* __restrict__ cell, size_t total, float* __restrict__ out,
                                                           int* __restrict__ ids) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const unsigned long long c = cell[i];
  out[i] = key_float((unsigned)(c >> 32));
  const unsigned pr = (unsigned)c;
  ids[i] = pr == 0xffffffffu ? -1 : (int)(0xfffffffeu - pr);
}

// ------------------------------------------------------------------------------------------ max, float64
__global__ void __launch_bounds__(256) p2i_max_init_f64(const double* __restrict__ bg, size_t total, unsigned long long* __restrict__ cell,
                                                         int* __restrict__ ids) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) {
    cell[i] = double_key(bg[i]);
    ids[i] = 0x7fffffff;
  }
}
// pass 0: cell = max key; pass 1: ids = min id among hits whose value equals the max and exceeds the background
__global__ void __launch_bounds__(256) p2i_max_splat_f64(const double* __restrict__ points, const double* __restrict__ feat,
                                                          const int* __restrict__ binds, const double* __restrict__ bg, int npoints, int B,
                                                          int C, int H, int W, double radius, unsigned long long* __restrict__ cell,
                                                          int* __restrict__ ids, int pass) {
  const int p = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (p >= npoints) return;
  const int b = binds[p];
  if (b < 0 || b >= B) return;
  const double py = points[p * 2 + 0], px = points[p * 2 + 1];
  const Footprint<double> fp(py, px, radius, H, W);
  const int npix = fp.bw * fp.bh;
  for (int t = lane; t < npix; t += 32) {
    const int xo = t / fp.bh;
    const int x = fp.x0 + xo, y = fp.y0 + (t - xo * fp.bh);
    const double dx = (double)x - px, dy = (double)y - py;
    const double r = sqrt(__fma_rn(dy, dy, __dmul_rn(dx, dx)));
    if (!(r <= radius)) continue;
    const double w = p2i_weight<double>(r, radius);
    for (int c = 0; c < C; c++) {
      const size_t o = (((size_t)b * C + c) * H + y) * W + x;
      const double v = __dmul_rn(feat[(size_t)p * C + c], w);
      if (pass == 0) atomicMax(&cell[o], double_key(v));
      else if (double_key(v) == cell[o] && bg[o] < v) atomicMin(&ids[o], p);
    }
  }
}
__global__ void __launch_bounds__(256) p2i_max_finish_f64(const unsigned long long* __restrict__ cell, size_t total, double* __restrict__ out,
                                                           int* __restrict__ ids) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  out[i] = key_double(cell[i]);
  if (ids[i] == 0x7fffffff) ids[i] = -1;
}

// ------------------------------------------------------------------------------------------ max backward
template <typename T>
__global__ void __launch_bounds__(256) p2i_max_bwd_kernel(const T* __restrict__ gout, const int* __restrict__ ids, const T* __restrict__ points,
                                                           const T* __restrict__ feat, int C, int H, int W, size_t total, T radius,
                                                           T* __restrict__ gpoints, T* __restrict__ gfeat, T* __restrict__ gbg) {
  const size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = o < total;                             // no early exit: the warp-wide scan below needs every lane
  const int x = (int)(o % W), y = (int)((o / W) % H), c = (int)((o / ((size_t)W * H)) % C);
  const T g = valid ? gout[o] : (T)0;
  const int p = valid ? ids[o] : -1;
  const bool hit = p >= 0;
  if (valid) gbg[o] = hit ? (T)0 : g;
  if (!__any_sync(0xffffffffu, hit)) return;                // a warp of background pixels (most of a sparse view) has nothing to scan
  // The pixels a point owns are runs along x, i.e. runs of lanes: the three gradient terms are summed over each run with a
  // segmented warp scan and the LAST lane of a run issues the atomics (the reference issues three per pixel; a footprint of
  // radius 10 piles ~300 of them onto the same three addresses).
  T t0 = (T)0, t1 = (T)0, t2 = (T)0;
  if (hit) {
    const T py = points[p * 2 + 0], px = points[p * 2 + 1];
    const T dx = (T)x - px, dy = (T)y - py;
    const T r = P2I<T>::sqrt_(P2I<T>::fma_(dx, dx, P2I<T>::mul_(dy, dy)));
    const T w = p2i_weight<T>(r, radius);
    const T f = feat[(size_t)p * C + c];
    t0 = g * w;
    const T wg = g * f;
    const T rr = r > (T)1e-10 ? r : (T)1e-10;
    const T k = (T)((double)wg * sin((double)r * P2I_PI / (double)radius) * 0.5 * P2I_PI / (double)radius / (double)rr);
    t1 = k * dy;
    t2 = k * dx;
  }
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int key = hit ? p : -1 - lane;                      // misses never join a run
  const int prev = __shfl_up_sync(full, key, 1);
  const bool head = lane == 0 || prev != key;
  const unsigned heads = __ballot_sync(full, head);
  const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));   // first lane of this lane's run
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const T a0 = __shfl_up_sync(full, t0, off), a1 = __shfl_up_sync(full, t1, off), a2 = __shfl_up_sync(full, t2, off);
    if (lane - off >= start) {
      t0 += a0;
      t1 += a1;
      t2 += a2;
    }
  }
  const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
  if (hit && tail) {
    atomicAdd(&gfeat[(size_t)p * C + c], t0);
    atomicAdd(&gpoints[p * 2 + 0], t1);
    atomicAdd(&gpoints[p * 2 + 1], t2);
  }
}

// ------------------------------------------------------------------------------------------ sum
template <typename T>
__global__ void __launch_bounds__(256) p2i_sum_fwd_kernel(const T* __restrict__ points, const T* __restrict__ feat, const int* __restrict__ binds,
                                                           int npoints, int B, int C, int H, int W, T radius, T* __restrict__ out) {
  const int p = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (p >= npoints) return;
  const int b = binds[p];
  if (b < 0 || b >= B) return;
  const T py = points[p * 2 + 0], px = points[p * 2 + 1];
  const Footprint<T> fp(py, px, radius, H, W);
  const int npix = fp.bw * fp.bh;
  for (int t = lane; t < npix; t += 32) {
    const int xo = t / fp.bh;
    const int x = fp.x0 + xo, y = fp.y0 + (t - xo * fp.bh);
    const T dx = (T)x - px, dy = (T)y - py;
    const T r = P2I<T>::sqrt_(P2I<T>::fma_(dy, dy, P2I<T>::mul_(dx, dx)));
    if (!(r <= radius)) continue;
    const T w = p2i_weight<T>(r, radius);
    for (int c = 0; c < C; c++) atomicAdd(&out[(((size_t)b * C + c) * H + y) * W + x], w * feat[(size_t)p * C + c]);
  }
}

// one warp per (point, channel): every gradient of a point comes from its own footprint -> warp reduce, plain
// stores for features; position gradients of different channels meet in atomics only when C > 1.
template <typename T>
__global__ void __launch_bounds__(256) p2i_sum_bwd_kernel(const T* __restrict__ gout, const T* __restrict__ points, const T* __restrict__ feat,
                                                           const int* __restrict__ binds, int npoints, int B, int C, int H, int W, T radius,
                                                           T* __restrict__ gpoints, T* __restrict__ gfeat) {
  const size_t wid = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= (size_t)npoints * C) return;
  const int p = (int)(wid / C), c = (int)(wid % C);
  const int b = binds[p];
  T af = 0, ay = 0, ax = 0;
  if (b >= 0 && b < B) {
    const T py = points[p * 2 + 0], px = points[p * 2 + 1];
    const T f = feat[(size_t)p * C + c];
    const Footprint<T> fp(py, px, radius, H, W);
    const int npix = fp.bw * fp.bh;
    for (int t = lane; t < npix; t += 32) {
      const int xo = t / fp.bh;
      const int x = fp.x0 + xo, y = fp.y0 + (t - xo * fp.bh);
      const T dx = (T)x - px, dy = (T)y - py;
      const T r = P2I<T>::sqrt_(P2I<T>::fma_(dy, dy, P2I<T>::mul_(dx, dx)));
      if (!(r <= radius)) continue;
      const T w = p2i_weight<T>(r, radius);
      const T g = gout[(((size_t)b * C + c) * H + y) * W + x];
      af += g * w;
      const T rr = r > (T)1e-10 ? r : (T)1e-10;
      const double kk = (double)(g * f) * sin((double)r * P2I_PI / (double)radius) * 0.5 * P2I_PI / (double)radius;
      ay += (T)(kk * (double)dy / (double)rr);
      ax += (T)(kk * (double)dx / (double)rr);
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    af += __shfl_xor_sync(0xffffffffu, af, o);
    ay += __shfl_xor_sync(0xffffffffu, ay, o);
    ax += __shfl_xor_sync(0xffffffffu, ax, o);
  }
  if (lane == 0) {
    gfeat[(size_t)p * C + c] = af;
    if (C == 1) {
      gpoints[p * 2 + 0] = ay;
      gpoints[p * 2 + 1] = ax;
    } else {
      atomicAdd(&gpoints[p * 2 + 0], ay);
      atomicAdd(&gpoints[p * 2 + 1], ax);
    }
  }
}

}  // namespace snb

using namespace snb;

static inline unsigned blocks_for(size_t n, int per) { return (unsigned)((n + per - 1) / per); }

SNB_API size_t snb_p2i_workspace_bytes(int B, int C, int H, int W, int is_double) {
  (void)is_double;
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
  return sizeof(unsigned long long) * (size_t)B * C * H * W;
}

static int p2i_check(int npoints, int B, int C, int H, int W, int kernel_kind, double radius) {
  if (npoints < 0 || B < 0 || C < 0 || H < 0 || W < 0) return SNB_EINVAL;
  if (kernel_kind != 0) return SNB_EINVAL;  // only the cosine kernel exists (p2i_op/__init__.py:96)
  if (!(radius > 0.0)) return SNB_EINVAL;
  return SNB_OK;
}

SNB_API int snb_p2i_max_fwd(const void* points, const void* features, const int* batch_inds, const void* background, int npoints, int B, int C,
                            int H, int W, int kernel_kind, double radius, int is_double, void* out, int* out_ids, void* workspace,
                            size_t workspace_bytes, void* stream) {
  int rc = p2i_check(npoints, B, C, H, W, kernel_kind, radius);
  if (rc) return rc;
  const size_t total = (size_t)B * C * H * W;
  if (total == 0) return SNB_OK;
  if (!workspace || workspace_bytes < snb_p2i_workspace_bytes(B, C, H, W, is_double)) return SNB_EWORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  unsigned long long* cell = (unsigned long long*)workspace;
  const unsigned gpix = blocks_for(total, 256), gpts = blocks_for((size_t)npoints * 32, 256);
  if (!is_double) {
    p2i_max_init_f32<<<gpix, 256, 0, s>>>((const float*)background, total, cell);
    if (npoints)
      p2i_max_splat_f32<<<gpts, 256, 0, s>>>((const float*)points, (const float*)features, batch_inds, npoints, B, C, H, W, (float)radius, cell);
    p2i_max_finish_f32<<<gpix, 256, 0, s>>>(cell, total, (float*)out, out_ids);
  } else {
    p2i_max_init_f64<<<gpix, 256, 0, s>>>((const double*)background, total, cell, out_ids);
    for (int pass = 0; pass < 2 && npoints; pass++)
      p2i_max_splat_f64<<<gpts, 256, 0, s>>>((const double*)points, (const double*)features, batch_inds, (const double*)background, npoints, B,
                                              C, H, W, radius, cell, out_ids, pass);
    p2i_max_finish_f64<<<gpix, 256, 0, s>>>(cell, total, (double*)out, out_ids);
  }
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

SNB_API int snb_p2i_max_bwd(const void* grad_out, const int* out_ids, const void* points, const void* features, int npoints, int B, int C, int H,
                            int W, int kernel_kind, double radius, int is_double, void* grad_points, void* grad_features,
                            void* grad_background, void* stream) {
  int rc = p2i_check(npoints, B, C, H, W, kernel_kind, radius);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t esz = is_double ? 8 : 4;
  if (npoints) {
    SNB_CUDA(cudaMemsetAsync(grad_points, 0, esz * (size_t)npoints * 2, s));
    if (C) SNB_CUDA(cudaMemsetAsync(grad_features, 0, esz * (size_t)npoints * C, s));
  }
  const size_t total = (size_t)B * C * H * W;
  if (total == 0) return SNB_OK;
  if (!is_double)
    p2i_max_bwd_kernel<float><<<blocks_for(total, 256), 256, 0, s>>>((const float*)grad_out, out_ids, (const float*)points,
                                                                     (const float*)features, C, H, W, total, (float)radius,
                                                                     (float*)grad_points, (float*)grad_features, (float*)grad_background);
  else
    p2i_max_bwd_kernel<double><<<blocks_for(total, 256), 256, 0, s>>>((const double*)grad_out, out_ids, (const double*)points,
                                                                      (const double*)features, C, H, W, total, radius, (double*)grad_points,
                                                                      (double*)grad_features, (double*)grad_background);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

SNB_API int snb_p2i_sum_fwd(const void* points, const void* features, const int* batch_inds, const void* background, int npoints, int B, int C,
                            int H, int W, int kernel_kind, double radius, int is_double, void* out, void* stream) {
  int rc = p2i_check(npoints, B, C, H, W, kernel_kind, radius);
  if (rc) return rc;
  const size_t total = (size_t)B * C * H * W;
  if (total == 0) return SNB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t esz = is_double ? 8 : 4;
  if (out != background) SNB_CUDA(cudaMemcpyAsync(out, background, esz * total, cudaMemcpyDeviceToDevice, s));
  if (npoints == 0) return SNB_OK;
  const unsigned gpts = blocks_for((size_t)npoints * 32, 256);
  if (!is_double)
    p2i_sum_fwd_kernel<float><<<gpts, 256, 0, s>>>((const float*)points, (const float*)features, batch_inds, npoints, B, C, H, W, (float)radius,
                                                   (float*)out);
  else
    p2i_sum_fwd_kernel<double><<<gpts, 256, 0, s>>>((const double*)points, (const double*)features, batch_inds, npoints, B, C, H, W, radius,
                                                    (double*)out);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

SNB_API int snb_p2i_sum_bwd(const void* grad_out, const void* points, const void* features, const int* batch_inds, int npoints, int B, int C, int H,
                            int W, int kernel_kind, double radius, int is_double, void* grad_points, void* grad_features, void* stream) {
  int rc = p2i_check(npoints, B, C, H, W, kernel_kind, radius);
  if (rc) return rc;
  if (npoints == 0 || C == 0) return SNB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t esz = is_double ? 8 : 4;
  if (C > 1) SNB_CUDA(cudaMemsetAsync(grad_points, 0, esz * (size_t)npoints * 2, s));
  const unsigned g = blocks_for((size_t)npoints * C * 32, 256);
  if (!is_double)
    p2i_sum_bwd_kernel<float><<<g, 256, 0, s>>>((const float*)grad_out, (const float*)points, (const float*)features, batch_inds, npoints, B, C, H,
                                                W, (float)radius, (float*)grad_points, (float*)grad_features);
  else
    p2i_sum_bwd_kernel<double><<<g, 256, 0, s>>>((const double*)grad_out, (const double*)points, (const double*)features, batch_inds, npoints, B, C,
                                                 H, W, radius, (double*)grad_points, (double*)grad_features);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
