// rowops.cu -- row-wise statistics and fused per-row affine + (leaky-)ReLU, forward and backward, sm_100a.
//
// These are the memory-bound tails of every dense layer of the generator once the normalisation algebra is folded
// (SURVEY.md 9.6): AdaIN o BatchNorm o SE o ReLU after a decoder conv (models/sparenet_generator.py:1053-1061), and
// BatchNorm o SE o ReLU in PointNetRes / conv5 of the encoder (:618-646, :234-236), all collapse to
//        y[r, :] = act( h[r, :] * scale[r] + shift[r] )          with one (scale, shift) per ROW r = (primitive, channel, sample)
// plus the row statistics (mean, biased variance) that the closed-form scale/shift are computed from.
// HBM-bound by construction: each kernel touches every element once with 128-bit accesses; one warp owns a row so the
// reductions are shuffle-only.  `in_div` lets several output rows share one input row (decoder layer 1: the folded
// lattice activations do not depend on the sample).
#include "common.cuh"

namespace snb {

constexpr int ROW_THREADS = 256;  // 8 warps = 8 rows per CTA

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// mean / biased variance per row; sums are shifted by the row's first element so E[x^2]-E[x]^2 does not cancel
__global__ void __launch_bounds__(ROW_THREADS) row_stats_kernel(const float* __restrict__ h, long long R, int L, float* __restrict__ mean,
                                                                 float* __restrict__ var) {
  const long long r = (long long)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const float* __restrict__ p = h + r * L;
  const float x0 = p[0];
  float s = 0.f, q = 0.f;
  if ((L & 3) == 0) {
    const float4* __restrict__ p4 = reinterpret_cast<const float4*>(p);
    for (int i = lane; i < (L >> 2); i += 32) {
      const float4 v = p4[i];
      const float a = v.x - x0, b = v.y - x0, c = v.z - x0, d = v.w - x0;
      s += (a + b) + (c + d);
      q = __fmaf_rn(a, a, __fmaf_rn(b, b, __fmaf_rn(c, c, __fmaf_rn(d, d, q))));
    }
  } else {
    for (int i = lane; i < L; i += 32) {
      const float a = p[i] - x0;
      s += a;
      q = __fmaf_rn(a, a, q);
    }
  }
  s = warp_sum(s);
  q = warp_sum(q);
  if (lane == 0) {
    const float m = s / (float)L;
    mean[r] = x0 + m;
    var[r] = fmaxf(q / (float)L - m * m, 0.f);
  }
}

// gh[r,l] = gmean[r]/L + gvar[r] * 2 (h[r,l] - mean[r]) / L      (adjoint of row_stats)
__global__ void __launch_bounds__(ROW_THREADS) row_stats_bwd_kernel(const float* __restrict__ h, const float* __restrict__ mean,
                                                                     const float* __restrict__ gmean, const float* __restrict__ gvar, long long R,
                                                                     int L, float* __restrict__ gh) {
  const long long r = (long long)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const float invL = 1.f / (float)L;
  const float a = gmean[r] * invL, b = 2.f * gvar[r] * invL, m = mean[r];
  const float* __restrict__ p = h + r * L;
  float* __restrict__ g = gh + r * L;
  if ((L & 3) == 0) {
    const float4* __restrict__ p4 = reinterpret_cast<const float4*>(p);
    float4* __restrict__ g4 = reinterpret_cast<float4*>(g);
    for (int i = lane; i < (L >> 2); i += 32) {
      const float4 v = p4[i];
      g4[i] = make_float4(__fmaf_rn(b, v.x - m, a), __fmaf_rn(b, v.y - m, a), __fmaf_rn(b, v.z - m, a), __fmaf_rn(b, v.w - m, a));
    }
  } else {
    for (int i = lane; i < L; i += 32) g[i] = __fmaf_rn(b, p[i] - m, a);
  }
}

__device__ __forceinline__ float act(float z, float slope) { return z > 0.f ? z : z * slope; }

// y[r,:] = act(h[r / in_div, :] * scale[r] + shift[r])
__global__ void __launch_bounds__(ROW_THREADS) row_affine_act_fwd_kernel(const float* __restrict__ h, const float* __restrict__ scale,
                                                                          const float* __restrict__ shift, long long R, int L, int in_div,
                                                                          float slope, float* __restrict__ y) {
  const long long r = (long long)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const float sc = scale[r], sh = shift[r];
  const float* __restrict__ p = h + (r / in_div) * L;
  float* __restrict__ o = y + r * L;
  if ((L & 3) == 0) {
    const float4* __restrict__ p4 = reinterpret_cast<const float4*>(p);
    float4* __restrict__ o4 = reinterpret_cast<float4*>(o);
    for (int i = lane; i < (L >> 2); i += 32) {
      const float4 v = p4[i];
      o4[i] = make_float4(act(__fmaf_rn(v.x, sc, sh), slope), act(__fmaf_rn(v.y, sc, sh), slope), act(__fmaf_rn(v.z, sc, sh), slope),
                          act(__fmaf_rn(v.w, sc, sh), slope));
    }
  } else {
    for (int i = lane; i < L; i += 32) o[i] = act(__fmaf_rn(p[i], sc, sh), slope);
  }
}

// Shared-input rows short enough to sit in registers (L = 128 * NV <= 1024): ONE pass over gy -- the input row and the running gh
// stay in registers while the in_div output rows stream by, each contributing its (gscale, gshift) through two warp sums.  (The
// general path below reads gy twice; on the decoders' first layer, gy = 2.2 GB, that was 1.33 ms at 3.3 TB/s effective.)
template <int NV>
__device__ __forceinline__ void row_affine_act_bwd_shared_regs(const float* __restrict__ gy, const float* __restrict__ p, const float* __restrict__ scale,
                                                               const float* __restrict__ shift, long long r0, int L, int in_div, float slope,
                                                               float* __restrict__ g, float* __restrict__ gscale, float* __restrict__ gshift,
                                                               int lane) {
  const float4* __restrict__ p4 = reinterpret_cast<const float4*>(p);
  float4 hv[NV], acc[NV];
#pragma unroll
  for (int u = 0; u < NV; u++) {
    hv[u] = p4[lane + 32 * u];
    acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll 2
  for (int s = 0; s < in_div; s++) {
    const long long r = r0 + s;
    const float sc = scale[r], sh = shift[r];
    const float4* __restrict__ q4 = reinterpret_cast<const float4*>(gy + r * L);
    float4 w[NV];
#pragma unroll
    for (int u = 0; u < NV; u++) w[u] = q4[lane + 32 * u];
    float as = 0.f, ab = 0.f;
#pragma unroll
    for (int u = 0; u < NV; u++) {
      const float d0 = __fmaf_rn(hv[u].x, sc, sh) > 0.f ? w[u].x : w[u].x * slope;
      const float d1 = __fmaf_rn(hv[u].y, sc, sh) > 0.f ? w[u].y : w[u].y * slope;
      const float d2 = __fmaf_rn(hv[u].z, sc, sh) > 0.f ? w[u].z : w[u].z * slope;
      const float d3 = __fmaf_rn(hv[u].w, sc, sh) > 0.f ? w[u].w : w[u].w * slope;
      acc[u].x = __fmaf_rn(d0, sc, acc[u].x);
      acc[u].y = __fmaf_rn(d1, sc, acc[u].y);
      acc[u].z = __fmaf_rn(d2, sc, acc[u].z);
      acc[u].w = __fmaf_rn(d3, sc, acc[u].w);
      as = __fmaf_rn(d0, hv[u].x, __fmaf_rn(d1, hv[u].y, __fmaf_rn(d2, hv[u].z, __fmaf_rn(d3, hv[u].w, as))));
      ab += (d0 + d1) + (d2 + d3);
    }
    as = warp_sum(as);
    ab = warp_sum(ab);
    if (lane == 0) {
      gscale[r] = as;
      gshift[r] = ab;
    }
  }
  float4* __restrict__ g4 = reinterpret_cast<float4*>(g);
#pragma unroll
  for (int u = 0; u < NV; u++) g4[lane + 32 * u] = acc[u];
}

// One warp per INPUT row: loops over the in_div output rows that share it.
//   d = gy * act'(z),  gh[rin,:] = sum_rows d * scale[r],  gscale[r] = sum_l d * h,  gshift[r] = sum_l d
__global__ void __launch_bounds__(ROW_THREADS) row_affine_act_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ h,
                                                                          const float* __restrict__ scale, const float* __restrict__ shift,
                                                                          long long Rin, int L, int in_div, float slope, float* __restrict__ gh,
                                                                          float* __restrict__ gscale, float* __restrict__ gshift) {
  const long long rin = (long long)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (rin >= Rin) return;
  const float* __restrict__ p = h + rin * L;
  float* __restrict__ g = gh + rin * L;
  if (in_div == 1) {
    const float sc = scale[rin], sh = shift[rin];
    const float* __restrict__ q = gy + rin * L;
    float as = 0.f, ab = 0.f;
    if ((L & 3) == 0) {
      const float4* __restrict__ p4 = reinterpret_cast<const float4*>(p);
      const float4* __restrict__ q4 = reinterpret_cast<const float4*>(q);
      float4* __restrict__ g4 = reinterpret_cast<float4*>(g);
      for (int i = lane; i < (L >> 2); i += 32) {
        const float4 v = p4[i], w = q4[i];
        const float d0 = __fmaf_rn(v.x, sc, sh) > 0.f ? w.x : w.x * slope;
        const float d1 = __fmaf_rn(v.y, sc, sh) > 0.f ? w.y : w.y * slope;
        const float d2 = __fmaf_rn(v.z, sc, sh) > 0.f ? w.z : w.z * slope;
        const float d3 = __fmaf_rn(v.w, sc, sh) > 0.f ? w.w : w.w * slope;
        g4[i] = make_float4(d0 * sc, d1 * sc, d2 * sc, d3 * sc);
        as = __fmaf_rn(d0, v.x, __fmaf_rn(d1, v.y, __fmaf_rn(d2, v.z, __fmaf_rn(d3, v.w, as))));
        ab += (d0 + d1) + (d2 + d3);
      }
    } else {
      for (int i = lane; i < L; i += 32) {
        const float v = p[i], w = q[i];
        const float d = __fmaf_rn(v, sc, sh) > 0.f ? w : w * slope;
        g[i] = d * sc;
        as = __fmaf_rn(d, v, as);
        ab += d;
      }
    }
    as = warp_sum(as);
    ab = warp_sum(ab);
    if (lane == 0) {
      gscale[rin] = as;
      gshift[rin] = ab;
    }
    return;
  }
  if ((L & 127) == 0 && L <= 1024) {
    const int nv = L >> 7;
    const long long r0 = rin * in_div;
    if (nv == 4) { row_affine_act_bwd_shared_regs<4>(gy, p, scale, shift, r0, L, in_div, slope, g, gscale, gshift, lane); return; }
    if (nv == 1) { row_affine_act_bwd_shared_regs<1>(gy, p, scale, shift, r0, L, in_div, slope, g, gscale, gshift, lane); return; }
    if (nv == 2) { row_affine_act_bwd_shared_regs<2>(gy, p, scale, shift, r0, L, in_div, slope, g, gscale, gshift, lane); return; }
    if (nv == 8) { row_affine_act_bwd_shared_regs<8>(gy, p, scale, shift, r0, L, in_div, slope, g, gscale, gshift, lane); return; }
  }
  // shared-input variant, any L: element-major so gh accumulates in registers across the in_div rows (two passes over gy)
  for (int i0 = lane; i0 < L; i0 += 32 * 4) {
    float hv[4], acc[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = i0 + 32 * u;
      hv[u] = i < L ? p[i] : 0.f;
      acc[u] = 0.f;
    }
    for (int s = 0; s < in_div; s++) {
      const long long r = rin * in_div + s;
      const float sc = scale[r], sh = shift[r];
      const float* __restrict__ q = gy + r * L;
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int i = i0 + 32 * u;
        if (i < L) {
          const float w = q[i];
          const float d = __fmaf_rn(hv[u], sc, sh) > 0.f ? w : w * slope;
          acc[u] = __fmaf_rn(d, sc, acc[u]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = i0 + 32 * u;
      if (i < L) g[i] = acc[u];
    }
  }
  for (int s = 0; s < in_div; s++) {
    const long long r = rin * in_div + s;
    const float sc = scale[r], sh = shift[r];
    const float* __restrict__ q = gy + r * L;
    float as = 0.f, ab = 0.f;
    for (int i = lane; i < L; i += 32) {
      const float v = p[i], w = q[i];
      const float d = __fmaf_rn(v, sc, sh) > 0.f ? w : w * slope;
      as = __fmaf_rn(d, v, as);
      ab += d;
    }
    as = warp_sum(as);
    ab = warp_sum(ab);
    if (lane == 0) {
      gscale[r] = as;
      gshift[r] = ab;
    }
  }
}

// max / min / argmax / argmin per row (PointNetRes global feature: max over points commutes with the monotone BN)
__global__ void __launch_bounds__(ROW_THREADS) row_minmax_kernel(const float* __restrict__ h, long long R, int L, float* __restrict__ vmax,
                                                                  float* __restrict__ vmin, int* __restrict__ imax, int* __restrict__ imin) {
  const long long r = (long long)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const float* __restrict__ p = h + r * L;
  float mx = -3.4e38f, mn = 3.4e38f;
  int ax = 0, an = 0;
  for (int i = lane; i < L; i += 32) {
    const float v = p[i];
    if (v > mx) { mx = v; ax = i; }
    if (v < mn) { mn = v; an = i; }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const float omx = __shfl_xor_sync(0xffffffffu, mx, o), omn = __shfl_xor_sync(0xffffffffu, mn, o);
    const int oax = __shfl_xor_sync(0xffffffffu, ax, o), oan = __shfl_xor_sync(0xffffffffu, an, o);
    if (omx > mx || (omx == mx && oax < ax)) { mx = omx; ax = oax; }
    if (omn < mn || (omn == mn && oan < an)) { mn = omn; an = oan; }
  }
  if (lane == 0) {
    vmax[r] = mx; vmin[r] = mn; imax[r] = ax; imin[r] = an;
  }
}

// row_stats + row_minmax in ONE pass over h (the refiner's 1024-channel global feature needs all six per row)
__global__ void __launch_bounds__(ROW_THREADS) row_stats_minmax_kernel(const float* __restrict__ h, long long R, int L, float* __restrict__ mean,
                                                                        float* __restrict__ var, float* __restrict__ vmax,
                                                                        float* __restrict__ vmin, int* __restrict__ imax, int* __restrict__ imin) {
  const long long r = (long long)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const float* __restrict__ p = h + r * L;
  const float x0 = p[0];
  float s = 0.f, q = 0.f, mx = -3.4e38f, mn = 3.4e38f;
  int ax = 0, an = 0;
  if ((L & 3) == 0) {
    const float4* __restrict__ p4 = reinterpret_cast<const float4*>(p);
    for (int i = lane; i < (L >> 2); i += 32) {
      const float4 v = p4[i];
      const float a = v.x - x0, b = v.y - x0, c = v.z - x0, d = v.w - x0;
      s += (a + b) + (c + d);
      q = __fmaf_rn(a, a, __fmaf_rn(b, b, __fmaf_rn(c, c, __fmaf_rn(d, d, q))));
      const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int u = 0; u < 4; u++) {  // strict comparisons keep the first position within the lane's increasing sequence
        if (e[u] > mx) { mx = e[u]; ax = 4 * i + u; }
        if (e[u] < mn) { mn = e[u]; an = 4 * i + u; }
      }
    }
  } else {
    for (int i = lane; i < L; i += 32) {
      const float v = p[i], a = v - x0;
      s += a;
      q = __fmaf_rn(a, a, q);
      if (v > mx) { mx = v; ax = i; }
      if (v < mn) { mn = v; an = i; }
    }
  }
  s = warp_sum(s);
  q = warp_sum(q);
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const float omx = __shfl_xor_sync(0xffffffffu, mx, o), omn = __shfl_xor_sync(0xffffffffu, mn, o);
    const int oax = __shfl_xor_sync(0xffffffffu, ax, o), oan = __shfl_xor_sync(0xffffffffu, an, o);
    if (omx > mx || (omx == mx && oax < ax)) { mx = omx; ax = oax; }
    if (omn < mn || (omn == mn && oan < an)) { mn = omn; an = oan; }
  }
  if (lane == 0) {
    const float m = s / (float)L;
    mean[r] = x0 + m;
    var[r] = fmaxf(q / (float)L - m * m, 0.f);
    vmax[r] = mx; vmin[r] = mn; imax[r] = ax; imin[r] = an;
  }
}

// Two-phase backward of  y = act(scale*h + shift)  when (scale, shift) are themselves functions of the row statistics of h
// (BatchNorm o SE o ReLU): phase A reduces  gscale[r] = sum_l d*h, gshift[r] = sum_l d  (d = gy * act'(z)) without writing
// anything per element; the host differentiates the small closed-form (mean, var) -> (scale, shift); phase B writes
//     gh = d*scale + gmean/L + 2 gvar (h - mean)/L
// in one pass.  5 tensor passes instead of the 8 of {affine_act_bwd, row_stats_bwd, add}.
__global__ void __launch_bounds__(ROW_THREADS) row_act_bwd_reduce_kernel(const float* __restrict__ gy, const float* __restrict__ h,
                                                                          const float* __restrict__ scale, const float* __restrict__ shift,
                                                                          long long R, int L, float slope, float* __restrict__ gscale,
                                                                          float* __restrict__ gshift, const float* __restrict__ gy_row) {
  const long long r = (long long)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const float sc = scale[r], sh = shift[r];
  const float gb = gy_row ? gy_row[r] : 0.f;   // a term of the upstream gradient that is constant along the row
  const float* __restrict__ p = h + r * L;
  const float* __restrict__ q = gy + r * L;
  float as = 0.f, ab = 0.f;
  if ((L & 3) == 0) {
    const float4* __restrict__ p4 = reinterpret_cast<const float4*>(p);
    const float4* __restrict__ q4 = reinterpret_cast<const float4*>(q);
    for (int i = lane; i < (L >> 2); i += 32) {
      const float4 v = p4[i];
      float4 w = q4[i];
      w.x += gb; w.y += gb; w.z += gb; w.w += gb;
      const float d0 = __fmaf_rn(v.x, sc, sh) > 0.f ? w.x : w.x * slope;
      const float d1 = __fmaf_rn(v.y, sc, sh) > 0.f ? w.y : w.y * slope;
      const float d2 = __fmaf_rn(v.z, sc, sh) > 0.f ? w.z : w.z * slope;
      const float d3 = __fmaf_rn(v.w, sc, sh) > 0.f ? w.w : w.w * slope;
      as = __fmaf_rn(d0, v.x, __fmaf_rn(d1, v.y, __fmaf_rn(d2, v.z, __fmaf_rn(d3, v.w, as))));
      ab += (d0 + d1) + (d2 + d3);
    }
  } else {
    for (int i = lane; i < L; i += 32) {
      const float v = p[i], w = q[i] + gb;
      const float d = __fmaf_rn(v, sc, sh) > 0.f ? w : w * slope;
      as = __fmaf_rn(d, v, as);
      ab += d;
    }
  }
  as = warp_sum(as);
  ab = warp_sum(ab);
  if (lane == 0) {
    gscale[r] = as;
    gshift[r] = ab;
  }
}

__global__ void __launch_bounds__(ROW_THREADS) row_norm_act_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ h,
                                                                        const float* __restrict__ scale, const float* __restrict__ shift,
                                                                        const float* __restrict__ mean, const float* __restrict__ gmean,
                                                                        const float* __restrict__ gvar, long long R, int L, float slope,
                                                                        float* __restrict__ gh, const float* __restrict__ gy_row) {
  const long long r = (long long)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const float sc = scale[r], sh = shift[r];
  const float gb = gy_row ? gy_row[r] : 0.f;
  const float invL = 1.f / (float)L;
  const float a = gmean[r] * invL, b = 2.f * gvar[r] * invL, m = mean[r];
  const float s1 = sc, s0 = sc * slope;
  const float* __restrict__ p = h + r * L;
  const float* __restrict__ q = gy + r * L;
  float* __restrict__ g = gh + r * L;
  if ((L & 3) == 0) {
    const float4* __restrict__ p4 = reinterpret_cast<const float4*>(p);
    const float4* __restrict__ q4 = reinterpret_cast<const float4*>(q);
    float4* __restrict__ g4 = reinterpret_cast<float4*>(g);
    for (int i = lane; i < (L >> 2); i += 32) {
      const float4 v = p4[i];
      float4 w = q4[i];
      w.x += gb; w.y += gb; w.z += gb; w.w += gb;
      float4 o;
      o.x = __fmaf_rn(w.x, __fmaf_rn(v.x, sc, sh) > 0.f ? s1 : s0, __fmaf_rn(b, v.x - m, a));
      o.y = __fmaf_rn(w.y, __fmaf_rn(v.y, sc, sh) > 0.f ? s1 : s0, __fmaf_rn(b, v.y - m, a));
      o.z = __fmaf_rn(w.z, __fmaf_rn(v.z, sc, sh) > 0.f ? s1 : s0, __fmaf_rn(b, v.z - m, a));
      o.w = __fmaf_rn(w.w, __fmaf_rn(v.w, sc, sh) > 0.f ? s1 : s0, __fmaf_rn(b, v.w - m, a));
      g4[i] = o;
    }
  } else {
    for (int i = lane; i < L; i += 32) {
      const float v = p[i];
      g[i] = __fmaf_rn(q[i] + gb, __fmaf_rn(v, sc, sh) > 0.f ? s1 : s0, __fmaf_rn(b, v - m, a));
    }
  }
}

// ---- pooled tail: max and mean over the row of act(scale*h + shift), without materialising the activated row -------------------
// (the encoder's output, models/sparenet_generator.py:234-242: LeakyReLU(BN(conv5)) -> [max over points | mean over points])
__global__ void __launch_bounds__(ROW_THREADS) row_act_pool_fwd_kernel(const float* __restrict__ h, const float* __restrict__ scale,
                                                                        const float* __restrict__ shift, long long R, int L, float slope,
                                                                        float* __restrict__ vmax, int* __restrict__ imax, float* __restrict__ vmean) {
  const long long r = (long long)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const float sc = scale[r], sh = shift[r];
  const float* __restrict__ p = h + r * L;
  float best = -INFINITY, sum = 0.f;
  int bi = 0x7fffffff;
  if ((L & 3) == 0) {
    const float4* __restrict__ p4 = reinterpret_cast<const float4*>(p);
    for (int i = lane; i < (L >> 2); i += 32) {
      const float4 v = p4[i];
      const float z0 = act(__fmaf_rn(v.x, sc, sh), slope), z1 = act(__fmaf_rn(v.y, sc, sh), slope);
      const float z2 = act(__fmaf_rn(v.z, sc, sh), slope), z3 = act(__fmaf_rn(v.w, sc, sh), slope);
      sum += (z0 + z1) + (z2 + z3);
      if (z0 > best) { best = z0; bi = 4 * i; }
      if (z1 > best) { best = z1; bi = 4 * i + 1; }
      if (z2 > best) { best = z2; bi = 4 * i + 2; }
      if (z3 > best) { best = z3; bi = 4 * i + 3; }
    }
  } else {
    for (int i = lane; i < L; i += 32) {
      const float z = act(__fmaf_rn(p[i], sc, sh), slope);
      sum += z;
      if (z > best) { best = z; bi = i; }
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }   // first position on ties
  }
  sum = warp_sum(sum);
  if (lane == 0) {
    vmax[r] = best;
    imax[r] = bi;
    vmean[r] = sum / (float)L;
  }
}

// gy[r,l] = gmean[r]/L + (l == imax[r]) gmax[r];  d = gy * act'(scale*h + shift);  gscale = sum_l d*h, gshift = sum_l d
__global__ void __launch_bounds__(ROW_THREADS) row_act_pool_bwd_reduce_kernel(const float* __restrict__ h, const float* __restrict__ scale,
                                                                               const float* __restrict__ shift, const float* __restrict__ gmax,
                                                                               const float* __restrict__ gmean, const int* __restrict__ imax,
                                                                               long long R, int L, float slope, float* __restrict__ gscale,
                                                                               float* __restrict__ gshift) {
  const long long r = (long long)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const float sc = scale[r], sh = shift[r], gm = gmean[r] / (float)L, gx = gmax[r];
  const int ix = imax[r];
  const float* __restrict__ p = h + r * L;
  float as = 0.f, ab = 0.f;
  if ((L & 3) == 0) {
    const float4* __restrict__ p4 = reinterpret_cast<const float4*>(p);
    for (int i = lane; i < (L >> 2); i += 32) {
      const float4 v = p4[i];
      const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float g = (4 * i + j) == ix ? gm + gx : gm;
        const float d = __fmaf_rn(e[j], sc, sh) > 0.f ? g : g * slope;
        as = __fmaf_rn(d, e[j], as);
        ab += d;
      }
    }
  } else {
    for (int i = lane; i < L; i += 32) {
      const float v = p[i];
      const float g = i == ix ? gm + gx : gm;
      const float d = __fmaf_rn(v, sc, sh) > 0.f ? g : g * slope;
      as = __fmaf_rn(d, v, as);
      ab += d;
    }
  }
  as = warp_sum(as);
  ab = warp_sum(ab);
  if (lane == 0) {
    gscale[r] = as;
    gshift[r] = ab;
  }
}

// gh = d*scale + gmean_stat/L + 2 gvar_stat (h - mean)/L   (the statistics' gradients come back from the closed-form tail)
__global__ void __launch_bounds__(ROW_THREADS) row_act_pool_bwd_kernel(const float* __restrict__ h, const float* __restrict__ scale,
                                                                        const float* __restrict__ shift, const float* __restrict__ mean,
                                                                        const float* __restrict__ gmax, const float* __restrict__ gmean,
                                                                        const int* __restrict__ imax, const float* __restrict__ gstat_mean,
                                                                        const float* __restrict__ gstat_var, long long R, int L, float slope,
                                                                        float* __restrict__ gh) {
  const long long r = (long long)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const float invL = 1.f / (float)L;
  const float sc = scale[r], sh = shift[r], gm = gmean[r] * invL, gx = gmax[r], m = mean[r];
  const float a = gstat_mean[r] * invL, b = 2.f * gstat_var[r] * invL;
  const int ix = imax[r];
  const float* __restrict__ p = h + r * L;
  float* __restrict__ o = gh + r * L;
  if ((L & 3) == 0) {
    const float4* __restrict__ p4 = reinterpret_cast<const float4*>(p);
    float4* __restrict__ o4 = reinterpret_cast<float4*>(o);
    for (int i = lane; i < (L >> 2); i += 32) {
      const float4 v = p4[i];
      const float e[4] = {v.x, v.y, v.z, v.w};
      float w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float g = (4 * i + j) == ix ? gm + gx : gm;
        const float d = __fmaf_rn(e[j], sc, sh) > 0.f ? g : g * slope;
        w[j] = __fmaf_rn(d, sc, __fmaf_rn(b, e[j] - m, a));
      }
      o4[i] = make_float4(w[0], w[1], w[2], w[3]);
    }
  } else {
    for (int i = lane; i < L; i += 32) {
      const float v = p[i];
      const float g = i == ix ? gm + gx : gm;
      const float d = __fmaf_rn(v, sc, sh) > 0.f ? g : g * slope;
      o[i] = __fmaf_rn(d, sc, __fmaf_rn(b, v - m, a));
    }
  }
}

}  // namespace snb

using namespace snb;

static inline unsigned row_grid(long long R) { return (unsigned)((R + ROW_THREADS / 32 - 1) / (ROW_THREADS / 32)); }

// ---- adjoint of the row extrema of h = W x (PointNetRes conv3 -> bn3 -> max over the points) ------------------------------------
// h [B,Co,N] is never stored; its row max / min sit in column imax / imin [b,co] of x [B,Ci,N].  A gradient g on such an extremum is
// g W[co,:] on that column of gx and g x[b,:,col] on row co of gW.  One launch for both extrema (entries with g == 0 -- the extremum
// the sign of the BatchNorm weight did not pick -- are skipped) instead of 2 x (mul, scatter_add_, gather, einsum + dtype copies) on
// [B,Ci,Co] tensors: ~270 us of PyTorch launches per call.  A block per (8 output channels, sample), threads over the input channels.
namespace snb {
__global__ void __launch_bounds__(128) conv_extrema_bwd_kernel(const float* __restrict__ x, const float* __restrict__ W, const int* __restrict__ imax,
                                                                const int* __restrict__ imin, const float* __restrict__ gmax,
                                                                const float* __restrict__ gmin, int Ci, int Co, int N, float* gx, float* gW) {
  const int b = blockIdx.y;
  const int co0 = blockIdx.x * 8;
  const float* __restrict__ xb = x + (size_t)b * Ci * N;
  float* gxb = gx ? gx + (size_t)b * Ci * N : nullptr;
  for (int r = 0; r < 8; r++) {
    const int co = co0 + r;
    if (co >= Co) break;
    const size_t e = (size_t)b * Co + co;
    for (int side = 0; side < 2; side++) {
      const float* __restrict__ gp = side ? gmin : gmax;
      if (!gp) continue;
      const float g = gp[e];
      if (g == 0.f) continue;
      const int col = side ? imin[e] : imax[e];
      for (int ci = threadIdx.x; ci < Ci; ci += blockDim.x) {
        if (gxb) atomicAdd(&gxb[(size_t)ci * N + col], g * W[(size_t)co * Ci + ci]);
        if (gW) atomicAdd(&gW[(size_t)co * Ci + ci], g * xb[(size_t)ci * N + col]);
      }
    }
  }
}
}  // namespace snb

// gx [B,Ci,N] and gW [Co,Ci] are ACCUMULATED into (atomics); either may be NULL, as may gmax / gmin.
SNB_API int snb_conv_extrema_bwd(const float* x, const float* W, const int* imax, const int* imin, const float* gmax, const float* gmin, int B, int Ci,
                                 int Co, int N, float* gx, float* gW, void* stream) {
  if (B < 0 || Ci <= 0 || Co <= 0 || N <= 0) return SNB_EINVAL;
  if (B > 65535) return SNB_ELIMIT;
  if (B == 0 || (!gx && !gW) || (!gmax && !gmin)) return SNB_OK;
  snb::conv_extrema_bwd_kernel<<<dim3((unsigned)((Co + 7) / 8), (unsigned)B), 128, 0, (cudaStream_t)stream>>>(x, W, imax, imin, gmax, gmin, Ci, Co, N, gx,
                                                                                                          gW);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

// ---- merge of the tensor-core GEMM's per-tile statistics (csrc/gemm_tc.cu epilogue) ------------------------------------------------
// The epilogue leaves, per output row and tile of `w` positions, the tile mean, the centred second moment and (optionally) the tile's
// extrema with their positions.  One launch folds them per segment of `tps` tiles -- Chan's pairwise update written for equal tile
// sizes: mean = avg of tile means, M2 = sum M2_t + w sum (mean_t - mean)^2 (no cancellation) -- and the extrema over ALL tiles of a
// row (first position attaining them).  This was 7 + 6 microsecond-sized PyTorch launches after every GEMM that returns statistics.
// A group of `lp` lanes (power of two <= 32) per (row, segment).
namespace snb {
__global__ void __launch_bounds__(256) gemm_stats_merge_kernel(const float* __restrict__ pm, const float* __restrict__ p2, long long pairs, int tps,
                                                                int lp, float w, float inv_seg, float* __restrict__ mean, float* __restrict__ var) {
  const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long pair = gt / lp;
  const int sub = (int)(gt % lp);
  const bool on = pair < pairs;
  const float* __restrict__ m_ = pm + (on ? pair : 0) * tps;
  const float* __restrict__ q_ = p2 + (on ? pair : 0) * tps;
  float s = 0.f, q = 0.f;
  if (on)
    for (int t = sub; t < tps; t += lp) {
      s += m_[t];
      q += q_[t];
    }
  for (int o = lp >> 1; o >= 1; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  const float mu = s / (float)tps;
  float d2 = 0.f;
  if (on)
    for (int t = sub; t < tps; t += lp) {
      const float d = m_[t] - mu;
      d2 = __fmaf_rn(d, d, d2);
    }
  for (int o = lp >> 1; o >= 1; o >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, o);
  if (on && sub == 0) {
    mean[pair] = mu;
    var[pair] = (q + w * d2) * inv_seg;
  }
}

__global__ void __launch_bounds__(256) gemm_minmax_merge_kernel(const float* __restrict__ px, const float* __restrict__ pn, const int* __restrict__ ix,
                                                                 const int* __restrict__ in_, long long rows, int T, int lp, float* __restrict__ vmax,
                                                                 float* __restrict__ vmin, int* __restrict__ imax, int* __restrict__ imin) {
  const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long row = gt / lp;
  const int sub = (int)(gt % lp);
  const bool on = row < rows;
  const size_t base = (size_t)(on ? row : 0) * T;
  float bx = -__int_as_float(0x7f800000), bn = __int_as_float(0x7f800000);
  int tx = 0x7fffffff, tn = 0x7fffffff;
  if (on)
    for (int t = sub; t < T; t += lp) {  // ascending tiles within a lane: strict comparisons keep the first
      const float a = px[base + t], b = pn[base + t];
      if (a > bx || tx == 0x7fffffff) { bx = a; tx = t; }
      if (b < bn || tn == 0x7fffffff) { bn = b; tn = t; }
    }
  for (int o = lp >> 1; o >= 1; o >>= 1) {
    const float ox = __shfl_xor_sync(0xffffffffu, bx, o), on_ = __shfl_xor_sync(0xffffffffu, bn, o);
    const int otx = __shfl_xor_sync(0xffffffffu, tx, o), otn = __shfl_xor_sync(0xffffffffu, tn, o);
    if (otx != 0x7fffffff && (tx == 0x7fffffff || ox > bx || (ox == bx && otx < tx))) { bx = ox; tx = otx; }
    if (otn != 0x7fffffff && (tn == 0x7fffffff || on_ < bn || (on_ == bn && otn < tn))) { bn = on_; tn = otn; }
  }
  if (on && sub == 0) {
    vmax[row] = bx;
    vmin[row] = bn;
    imax[row] = ix[base + tx];
    imin[row] = in_[base + tn];
  }
}
}  // namespace snb

// pm, p2 [pairs, tps] -> mean, var [pairs] (pairs = rows x segments; w = positions per tile, seg = tps * w positions per segment)
SNB_API int snb_gemm_stats_merge(const float* pmean, const float* pm2, long long pairs, int tps, int w, float* mean, float* var, void* stream) {
  if (pairs < 0 || tps <= 0 || w <= 0) return SNB_EINVAL;
  if (pairs == 0) return SNB_OK;
  int lp = 1;
  while (lp < tps && lp < 32) lp <<= 1;
  const long long threads = pairs * lp;
  snb::gemm_stats_merge_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pmean, pm2, pairs, tps, lp, (float)w,
                                                                                                 1.0f / ((float)tps * (float)w), mean, var);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

// tile extrema [rows, T] with their positions -> the rows' extrema (first position attaining them)
SNB_API int snb_gemm_minmax_merge(const float* pmax, const float* pmin, const int* pimax, const int* pimin, long long rows, int T, float* vmax,
                                  float* vmin, int* imax, int* imin, void* stream) {
  if (rows < 0 || T <= 0) return SNB_EINVAL;
  if (rows == 0) return SNB_OK;
  int lp = 1;
  while (lp < T && lp < 32) lp <<= 1;
  const long long threads = rows * lp;
  snb::gemm_minmax_merge_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pmax, pmin, pimax, pimin, rows, T, lp, vmax, vmin,
                                                                                                  imax, imin);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

SNB_API int snb_row_stats(const float* h, long long R, int L, float* mean, float* var, void* stream) {
  if (R < 0 || L <= 0) return SNB_EINVAL;
  if (R == 0) return SNB_OK;
  if (R > 0x3fffffffLL) return SNB_ELIMIT;
  row_stats_kernel<<<row_grid(R), ROW_THREADS, 0, (cudaStream_t)stream>>>(h, R, L, mean, var);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

SNB_API int snb_row_stats_bwd(const float* h, const float* mean, const float* gmean, const float* gvar, long long R, int L, float* gh,
                              void* stream) {
  if (R < 0 || L <= 0) return SNB_EINVAL;
  if (R == 0) return SNB_OK;
  if (R > 0x3fffffffLL) return SNB_ELIMIT;
  row_stats_bwd_kernel<<<row_grid(R), ROW_THREADS, 0, (cudaStream_t)stream>>>(h, mean, gmean, gvar, R, L, gh);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

SNB_API int snb_row_affine_act_fwd(const float* h, const float* scale, const float* shift, long long R, int L, int in_div, float slope,
                                   float* y, void* stream) {
  if (R < 0 || L <= 0 || in_div <= 0 || (R % in_div) != 0) return SNB_EINVAL;
  if (R == 0) return SNB_OK;
  if (R > 0x3fffffffLL) return SNB_ELIMIT;
  row_affine_act_fwd_kernel<<<row_grid(R), ROW_THREADS, 0, (cudaStream_t)stream>>>(h, scale, shift, R, L, in_div, slope, y);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

SNB_API int snb_row_affine_act_bwd(const float* gy, const float* h, const float* scale, const float* shift, long long R, int L, int in_div,
                                   float slope, float* gh, float* gscale, float* gshift, void* stream) {
  if (R < 0 || L <= 0 || in_div <= 0 || (R % in_div) != 0) return SNB_EINVAL;
  if (R == 0) return SNB_OK;
  if (R > 0x3fffffffLL) return SNB_ELIMIT;
  row_affine_act_bwd_kernel<<<row_grid(R / in_div), ROW_THREADS, 0, (cudaStream_t)stream>>>(gy, h, scale, shift, R / in_div, L, in_div, slope, gh,
                                                                                          gscale, gshift);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

SNB_API int snb_row_minmax(const float* h, long long R, int L, float* vmax, float* vmin, int* imax, int* imin, void* stream) {
  if (R < 0 || L <= 0) return SNB_EINVAL;
  if (R == 0) return SNB_OK;
  if (R > 0x3fffffffLL) return SNB_ELIMIT;
  row_minmax_kernel<<<row_grid(R), ROW_THREADS, 0, (cudaStream_t)stream>>>(h, R, L, vmax, vmin, imax, imin);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

SNB_API int snb_row_stats_minmax(const float* h, long long R, int L, float* mean, float* var, float* vmax, float* vmin, int* imax, int* imin,
                                 void* stream) {
  if (R < 0 || L <= 0) return SNB_EINVAL;
  if (R == 0) return SNB_OK;
  if (R > 0x3fffffffLL) return SNB_ELIMIT;
  row_stats_minmax_kernel<<<row_grid(R), ROW_THREADS, 0, (cudaStream_t)stream>>>(h, R, L, mean, var, vmax, vmin, imax, imin);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

SNB_API int snb_row_act_bwd_reduce(const float* gy, const float* h, const float* scale, const float* shift, long long R, int L, float slope,
                                   float* gscale, float* gshift, const float* gy_row, void* stream) {
  if (R < 0 || L <= 0) return SNB_EINVAL;
  if (R == 0) return SNB_OK;
  if (R > 0x3fffffffLL) return SNB_ELIMIT;
  row_act_bwd_reduce_kernel<<<row_grid(R), ROW_THREADS, 0, (cudaStream_t)stream>>>(gy, h, scale, shift, R, L, slope, gscale, gshift, gy_row);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

SNB_API int snb_row_norm_act_bwd(const float* gy, const float* h, const float* scale, const float* shift, const float* mean,
                                 const float* gmean, const float* gvar, long long R, int L, float slope, float* gh, const float* gy_row,
                                 void* stream) {
  if (R < 0 || L <= 0) return SNB_EINVAL;
  if (R == 0) return SNB_OK;
  if (R > 0x3fffffffLL) return SNB_ELIMIT;
  row_norm_act_bwd_kernel<<<row_grid(R), ROW_THREADS, 0, (cudaStream_t)stream>>>(gy, h, scale, shift, mean, gmean, gvar, R, L, slope, gh, gy_row);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

// pooled tail (max / mean over the row of act(scale*h + shift)); see the kernels above
SNB_API int snb_row_act_pool_fwd(const float* h, const float* scale, const float* shift, long long R, int L, float slope, float* vmax, int* imax,
                                 float* vmean, void* stream) {
  if (R < 0 || L <= 0) return SNB_EINVAL;
  if (R == 0) return SNB_OK;
  if (R > 0x3fffffffLL) return SNB_ELIMIT;
  row_act_pool_fwd_kernel<<<row_grid(R), ROW_THREADS, 0, (cudaStream_t)stream>>>(h, scale, shift, R, L, slope, vmax, imax, vmean);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

SNB_API int snb_row_act_pool_bwd_reduce(const float* h, const float* scale, const float* shift, const float* gmax, const float* gmean,
                                        const int* imax, long long R, int L, float slope, float* gscale, float* gshift, void* stream) {
  if (R < 0 || L <= 0) return SNB_EINVAL;
  if (R == 0) return SNB_OK;
  if (R > 0x3fffffffLL) return SNB_ELIMIT;
  row_act_pool_bwd_reduce_kernel<<<row_grid(R), ROW_THREADS, 0, (cudaStream_t)stream>>>(h, scale, shift, gmax, gmean, imax, R, L, slope, gscale,
                                                                                        gshift);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

SNB_API int snb_row_act_pool_bwd(const float* h, const float* scale, const float* shift, const float* mean, const float* gmax, const float* gmean,
                                 const int* imax, const float* gstat_mean, const float* gstat_var, long long R, int L, float slope, float* gh,
                                 void* stream) {
  if (R < 0 || L <= 0) return SNB_EINVAL;
  if (R == 0) return SNB_OK;
  if (R > 0x3fffffffLL) return SNB_ELIMIT;
  row_act_pool_bwd_kernel<<<row_grid(R), ROW_THREADS, 0, (cudaStream_t)stream>>>(h, scale, shift, mean, gmax, gmean, imax, gstat_mean, gstat_var, R,
                                                                                 L, slope, gh);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
