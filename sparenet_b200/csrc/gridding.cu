// gridding.cu -- GRNet's Gridding (point cloud -> 3-D occupancy grid) and GriddingReverse (grid -> points), sm_100a.
//
// Replaces gridding_kernel / gridding_grad_kernel (cuda/gridding/gridding.cu:29-177,213-312) and
// gridding_reverse_kernel / gridding_reverse_grad_kernel (cuda/gridding/gridding_reverse.cu:30-103,124-214).
// Contract: per point the 8 cell corners (lower = floor, upper = ceil, +1 if equal), per-axis weight
// 1 - |p - corner| saved as [B,n,8,3], corner indexes [B,n,8], grid += wx*wy*wz; the backward is -/+ g * (product of the
// other two weights) summed over the corners in order.  Reverse: every vertex with all offsets >= 1 emits the
// grid-weighted centroid of the cell below it (skipped when the weights sum < 1e-6).
// The reference runs ONE block per sample (<<<B, 512>>>); here every point / vertex is a thread of a full grid.
// Unlike the reference, corners that fall outside the grid are skipped (index -1) instead of written out of bounds.
//
// The same two kernels serve GRNet's gridding LOSS (cuda/gridding_loss/gridding_distance.cu:29-177,213-338): there every vertex
// keeps EIGHT accumulators, one per corner role (index = vertex * 8 + corner), so `slots` = 8 instead of 1.
// cubic_feature_sampling (cuda/cubic_feature_sampling/cubic_feature_sampling.cu:29-204) is at the end of this file.
#include <math.h>
#include "common.cuh"

namespace snb {

__global__ void __launch_bounds__(256) gridding_fwd_kernel(const float* __restrict__ pts, size_t total, int n, float minx, float miny, float minz,
                                                            int lx, int ly, int lz, int slots, float* __restrict__ grid,
                                                            float* __restrict__ weights, int* __restrict__ indexes) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t b = i / n;
  const float p[3] = {pts[i * 3 + 0], pts[i * 3 + 1], pts[i * 3 + 2]};
  const float mn[3] = {minx, miny, minz};
  const int len[3] = {lx, ly, lz};
  int lo[3], hi[3];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    lo[c] = (int)floorf(p[c]);
    hi[c] = (int)ceilf(p[c]);
    if (lo[c] == hi[c]) hi[c] += 1;
  }
  float* __restrict__ g = grid + b * (size_t)lx * ly * lz * slots;
#pragma unroll
  for (int t = 0; t < 8; t++) {
    const int u[3] = {(t >> 2) & 1, (t >> 1) & 1, t & 1};
    int off[3];
    float w[3];
    bool inside = true;
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const int corner = u[c] ? hi[c] : lo[c];
      w[c] = 1.f - fabsf(p[c] - (float)corner);
      off[c] = (int)((float)corner - mn[c]);
      inside = inside && off[c] >= 0 && off[c] < len[c];
      weights[i * 24 + t * 3 + c] = w[c];
    }
    const int ix = inside ? ((off[0] * ly + off[1]) * lz + off[2]) * slots + (slots == 8 ? t : 0) : -1;
    indexes[i * 8 + t] = ix;
    if (inside) atomicAdd(&g[ix], __fmul_rn(__fmul_rn(w[0], w[1]), w[2]));
  }
}

__global__ void __launch_bounds__(256) gridding_bwd_kernel(const float* __restrict__ weights, const int* __restrict__ indexes,
                                                            const float* __restrict__ ggrid, size_t total, int n, size_t nv,
                                                            float* __restrict__ gpts) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float* __restrict__ gg = ggrid + (i / n) * nv;
  float g0 = 0.f, g1 = 0.f, g2 = 0.f;
#pragma unroll
  for (int t = 0; t < 8; t++) {
    const int ix = indexes[i * 8 + t];
    if (ix < 0) continue;
    const float gv = gg[ix];
    const float wx = weights[i * 24 + t * 3 + 0], wy = weights[i * 24 + t * 3 + 1], wz = weights[i * 24 + t * 3 + 2];
    g0 += __fmul_rn(__fmul_rn(((t >> 2) & 1) ? gv : -gv, wy), wz);
    g1 += __fmul_rn(__fmul_rn(((t >> 1) & 1) ? gv : -gv, wx), wz);
    g2 += __fmul_rn(__fmul_rn((t & 1) ? gv : -gv, wx), wy);
  }
  gpts[i * 3 + 0] = g0;
  gpts[i * 3 + 1] = g1;
  gpts[i * 3 + 2] = g2;
}

__device__ __forceinline__ bool rev_cell(const float* __restrict__ g, int S, size_t j, int& x, int& y, int& z, size_t (&ix)[8], float (&w)[8],
                                         float& sum) {
  x = (int)(j / ((size_t)S * S));
  y = (int)(j % ((size_t)S * S) / S);
  z = (int)(j % S);
  if (x == 0 || y == 0 || z == 0) return false;
  sum = 0.f;
#pragma unroll
  for (int t = 0; t < 8; t++) {
    ix[t] = ((size_t)(x - 1 + ((t >> 2) & 1)) * S + (y - 1 + ((t >> 1) & 1))) * S + (z - 1 + (t & 1));
    w[t] = g[ix[t]];
    sum += w[t];
  }
  return !(sum < 1e-6f);
}

__global__ void __launch_bounds__(256) gridding_rev_fwd_kernel(const float* __restrict__ grid, size_t total, int S, float* __restrict__ pts) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t nv = (size_t)S * S * S;
  const size_t j = i % nv;
  int x, y, z;
  size_t ix[8];
  float w[8], sum;
  float px = 0.f, py = 0.f, pz = 0.f;
  if (rev_cell(grid + (i / nv) * nv, S, j, x, y, z, ix, w, sum)) {
    const float cx = (float)(x - S / 2), cy = (float)(y - S / 2), cz = (float)(z - S / 2);
#pragma unroll
    for (int t = 0; t < 8; t++) {
      const float wn = w[t] / sum;
      px += wn * (((t >> 2) & 1) ? cx : cx - 1.f);
      py += wn * (((t >> 1) & 1) ? cy : cy - 1.f);
      pz += wn * ((t & 1) ? cz : cz - 1.f);
    }
  }
  pts[i * 3 + 0] = px;
  pts[i * 3 + 1] = py;
  pts[i * 3 + 2] = pz;
}

__global__ void __launch_bounds__(256) gridding_rev_bwd_kernel(const float* __restrict__ pts, const float* __restrict__ grid,
                                                                const float* __restrict__ gpts, size_t total, int S, float* __restrict__ ggrid) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t nv = (size_t)S * S * S;
  const size_t j = i % nv, boff = (i / nv) * nv;
  int x, y, z;
  size_t ix[8];
  float w[8], sum;
  if (!rev_cell(grid + boff, S, j, x, y, z, ix, w, sum)) return;
  const float cx = (float)(x - S / 2), cy = (float)(y - S / 2), cz = (float)(z - S / 2);
  const float p0 = pts[i * 3 + 0], p1 = pts[i * 3 + 1], p2 = pts[i * 3 + 2];
  const float g0 = gpts[i * 3 + 0], g1 = gpts[i * 3 + 1], g2 = gpts[i * 3 + 2];
#pragma unroll
  for (int t = 0; t < 8; t++) {
    const float vx = ((t >> 2) & 1) ? cx : cx - 1.f, vy = ((t >> 1) & 1) ? cy : cy - 1.f, vz = (t & 1) ? cz : cz - 1.f;
    atomicAdd(&ggrid[boff + ix[t]], g0 * (vx - p0) / sum + g1 * (vy - p1) / sum + g2 * (vz - p2) / sum);
  }
}

// ---- cubic feature sampling (GRNet) ---------------------------------------------------------------------------------------------
// per point the (2 ns)^3 grid vertices around it (lower - (ns-1) .. upper + (ns-1) per axis, x outermost), -1 outside the grid
// (cubic_feature_sampling.cu:49-87); point_features[b, i, v, :] = cubic_features[b, :, vertex] (zeros where outside, :89-103).
// One thread per (point, vertex, channel): the channel axis of the output is contiguous, the gather reads the L2-resident volume.
__global__ void __launch_bounds__(256) cubic_sampling_index_kernel(const float* __restrict__ pts, size_t total, int S, int ns,
                                                                    int* __restrict__ indexes) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // point
  if (i >= total) return;
  int lo[3], hi[3];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const float p = pts[i * 3 + c];
    lo[c] = (int)floorf(p);
    hi[c] = (int)ceilf(p);
    if (lo[c] == hi[c]) hi[c] += 1;
  }
  const int e = ns - 1, side = 2 * ns;
  int v = 0;
  for (int j = lo[0] - e; j <= hi[0] + e; ++j)
    for (int k = lo[1] - e; k <= hi[1] + e; ++k)
      for (int m = lo[2] - e; m <= hi[2] + e; ++m) {
        const bool out = j < 0 || j >= S || k < 0 || k >= S || m < 0 || m >= S;
        indexes[i * (size_t)(side * side * side) + v++] = out ? -1 : (j * S + k) * S + m;
      }
}

__global__ void __launch_bounds__(256) cubic_sampling_gather_kernel(const float* __restrict__ feat, const int* __restrict__ indexes, size_t total,
                                                                     size_t per_batch, int C, size_t cub, float* __restrict__ out) {
  const size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // ((b*n + i)*V + v)*C + k
  if (o >= total) return;
  const int k = (int)(o % C);
  const size_t pv = o / C;                                           // (b*n + i)*V + v
  const int ix = indexes[pv];
  const size_t b = pv / per_batch;
  out[o] = ix < 0 ? 0.f : feat[(b * C + k) * cub + ix];
}

__global__ void __launch_bounds__(256) cubic_sampling_bwd_kernel(const float* __restrict__ gout, const int* __restrict__ indexes, size_t total,
                                                                  size_t per_batch, int C, size_t cub, float* __restrict__ gfeat) {
  const size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= total) return;
  const int k = (int)(o % C);
  const size_t pv = o / C;
  const int ix = indexes[pv];
  if (ix < 0) return;
  atomicAdd(&gfeat[((pv / per_batch) * C + k) * cub + ix], gout[o]);
}

}  // namespace snb

using namespace snb;

// cubic_feature_sampling.forward (cubic_feature_sampling_cuda.cpp, .cu:105-137): ptcloud [B,n,3] in grid units, cubic_features
// [B,C,S,S,S] -> point_features [B,n,(2 ns)^3,C], grid_pt_indexes [B,n,(2 ns)^3]
SNB_API int snb_cubic_sampling_fwd(const float* ptcloud, const float* cubic_features, int B, int n, int C, int scale, int neighborhood_size,
                                   float* point_features, int* grid_pt_indexes, void* stream) {
  if (B < 0 || n < 0 || C < 0 || scale <= 0 || neighborhood_size <= 0) return SNB_EINVAL;
  if (scale > 1024 || neighborhood_size > 4) return SNB_ELIMIT;
  const size_t V = (size_t)8 * neighborhood_size * neighborhood_size * neighborhood_size, pts = (size_t)B * n;
  if (pts == 0) return SNB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  cubic_sampling_index_kernel<<<(unsigned)((pts + 255) / 256), 256, 0, s>>>(ptcloud, pts, scale, neighborhood_size, grid_pt_indexes);
  const size_t total = pts * V * C;
  if (total)
    cubic_sampling_gather_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(cubic_features, grid_pt_indexes, total, (size_t)n * V, C,
                                                                                (size_t)scale * scale * scale, point_features);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

// cubic_feature_sampling.backward (.cu:139-204): grad_cubic_features [B,C,S,S,S] (fully written); the cloud receives no gradient
// (floor / ceil have zero derivative, :165-170): the wrapper returns zeros for it like the reference
SNB_API int snb_cubic_sampling_bwd(const float* grad_point_features, const int* grid_pt_indexes, int B, int n, int C, int scale,
                                   int neighborhood_size, float* grad_cubic_features, void* stream) {
  if (B < 0 || n < 0 || C < 0 || scale <= 0 || neighborhood_size <= 0) return SNB_EINVAL;
  if (scale > 1024 || neighborhood_size > 4) return SNB_ELIMIT;
  const size_t V = (size_t)8 * neighborhood_size * neighborhood_size * neighborhood_size, cub = (size_t)scale * scale * scale;
  cudaStream_t s = (cudaStream_t)stream;
  if ((size_t)B * C * cub) SNB_CUDA(cudaMemsetAsync(grad_cubic_features, 0, sizeof(float) * (size_t)B * C * cub, s));
  const size_t total = (size_t)B * n * V * C;
  if (total)
    cubic_sampling_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(grad_point_features, grid_pt_indexes, total, (size_t)n * V, C, cub,
                                                                             grad_cubic_features);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

SNB_API int snb_gridding_fwd(const float* ptcloud, int B, int n, float min_x, float max_x, float min_y, float max_y, float min_z, float max_z,
                             float* grid, float* grid_pt_weights, int* grid_pt_indexes, void* stream) {
  if (B < 0 || n < 0) return SNB_EINVAL;
  const int lx = (int)(max_x - min_x + 1), ly = (int)(max_y - min_y + 1), lz = (int)(max_z - min_z + 1);
  if (lx <= 0 || ly <= 0 || lz <= 0) return SNB_EINVAL;
  if ((long long)lx * ly * lz > 0x7fffffffLL) return SNB_ELIMIT;
  if (B == 0) return SNB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  SNB_CUDA(cudaMemsetAsync(grid, 0, sizeof(float) * (size_t)B * lx * ly * lz, s));
  const size_t total = (size_t)B * n;
  if (total == 0) return SNB_OK;
  gridding_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(ptcloud, total, n, min_x, min_y, min_z, lx, ly, lz, 1, grid, grid_pt_weights,
                                                                     grid_pt_indexes);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

// gridding_distance.forward (cuda/gridding_loss/gridding_distance_cuda.cpp, gridding_distance.cu:179-211): grid [B, V, 8]
SNB_API int snb_gridding_dist_fwd(const float* ptcloud, int B, int n, float min_x, float max_x, float min_y, float max_y, float min_z, float max_z,
                                  float* grid, float* grid_pt_weights, int* grid_pt_indexes, void* stream) {
  if (B < 0 || n < 0) return SNB_EINVAL;
  const int lx = (int)(max_x - min_x + 1), ly = (int)(max_y - min_y + 1), lz = (int)(max_z - min_z + 1);
  if (lx <= 0 || ly <= 0 || lz <= 0) return SNB_EINVAL;
  if ((long long)lx * ly * lz * 8 > 0x7fffffffLL) return SNB_ELIMIT;
  if (B == 0) return SNB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  SNB_CUDA(cudaMemsetAsync(grid, 0, sizeof(float) * (size_t)B * lx * ly * lz * 8, s));
  const size_t total = (size_t)B * n;
  if (total == 0) return SNB_OK;
  gridding_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(ptcloud, total, n, min_x, min_y, min_z, lx, ly, lz, 8, grid, grid_pt_weights,
                                                                     grid_pt_indexes);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

// gridding_distance.backward (gridding_distance.cu:213-338): identical routing, the gradient grid is [B, V, 8]
SNB_API int snb_gridding_dist_bwd(const float* grid_pt_weights, const int* grid_pt_indexes, const float* grad_grid, int B, int n,
                                  long long n_grid_vertices, float* grad_ptcloud, void* stream) {
  if (B < 0 || n < 0 || n_grid_vertices < 0) return SNB_EINVAL;
  const size_t total = (size_t)B * n;
  if (total == 0) return SNB_OK;
  gridding_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(grid_pt_weights, grid_pt_indexes, grad_grid, total, n,
                                                                                        (size_t)n_grid_vertices * 8, grad_ptcloud);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

SNB_API int snb_gridding_bwd(const float* grid_pt_weights, const int* grid_pt_indexes, const float* grad_grid, int B, int n,
                             long long n_grid_vertices, float* grad_ptcloud, void* stream) {
  if (B < 0 || n < 0 || n_grid_vertices < 0) return SNB_EINVAL;
  const size_t total = (size_t)B * n;
  if (total == 0) return SNB_OK;
  gridding_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(grid_pt_weights, grid_pt_indexes, grad_grid, total, n,
                                                                                        (size_t)n_grid_vertices, grad_ptcloud);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

SNB_API int snb_gridding_rev_fwd(const float* grid, int B, int scale, float* ptcloud, void* stream) {
  if (B < 0 || scale <= 0) return SNB_EINVAL;
  if (scale > 1024) return SNB_ELIMIT;
  const size_t total = (size_t)B * scale * scale * scale;
  if (total == 0) return SNB_OK;
  gridding_rev_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(grid, total, scale, ptcloud);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

SNB_API int snb_gridding_rev_bwd(const float* ptcloud, const float* grid, const float* grad_ptcloud, int B, int scale, float* grad_grid,
                                 void* stream) {
  if (B < 0 || scale <= 0) return SNB_EINVAL;
  if (scale > 1024) return SNB_ELIMIT;
  const size_t total = (size_t)B * scale * scale * scale;
  if (total == 0) return SNB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  SNB_CUDA(cudaMemsetAsync(grad_grid, 0, sizeof(float) * total, s));
  gridding_rev_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(ptcloud, grid, grad_ptcloud, total, scale, grad_grid);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
