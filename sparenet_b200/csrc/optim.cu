// optim.cu -- Adam over one flat parameter arena in ONE launch, sm_100a.
//
// The generator has ~800 parameter tensors (32 folding decoders); torch.optim.Adam(fused=True) walks them in 40 multi-tensor
// launches of 60-320 blocks each and reaches ~2.1 TB/s on the 330 MB of parameters (1.07 ms per step).  With parameters, gradients
// and both moments living in flat 16-byte-aligned arenas of identical layout (sparenet_b200/dist.py: GradArena; optim.py:
// FlatAdam) the step is one grid-stride pass of 128-bit accesses: 28 B per parameter at the HBM roofline.
// Update rule = torch.optim.Adam's (torch/optim/adam.py, _fused_adam; reference runners/sparenet_runner.py builds torch.optim.Adam):
//   g += wd * p;  m = m + (1 - b1) (g - m);  v = b2 v + (1 - b2) g g;  p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps).
#include <math.h>
#include "common.cuh"

namespace snb {

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, float step_size, float b1c, float b2, float b2c, float eps,
                                      float rbc2s, float wd) {
  g = __fmaf_rn(wd, p, g);
  m = __fmaf_rn(b1c, g - m, m);
  v = __fmaf_rn(b2, v, b2c * g * g);
  const float denom = __fmaf_rn(sqrtf(v), rbc2s, eps);
  p -= step_size * (m / denom);
}

__global__ void __launch_bounds__(256) adam_flat_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                                                         float4* __restrict__ v, size_t n4, float step_size, float b1c, float b2, float b2c,
                                                         float eps, float rbc2s, float wd) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 pp = p[i], mm = m[i], vv = v[i];
    const float4 gg = g[i];
    adam1(pp.x, gg.x, mm.x, vv.x, step_size, b1c, b2, b2c, eps, rbc2s, wd);
    adam1(pp.y, gg.y, mm.y, vv.y, step_size, b1c, b2, b2c, eps, rbc2s, wd);
    adam1(pp.z, gg.z, mm.z, vv.z, step_size, b1c, b2, b2c, eps, rbc2s, wd);
    adam1(pp.w, gg.w, mm.w, vv.w, step_size, b1c, b2, b2c, eps, rbc2s, wd);
    p[i] = pp;
    m[i] = mm;
    v[i] = vv;
  }
}

// ---- gather of the step's gradient tensors into the flat arena: ONE launch over a pointer table ---------------------------------------
// (sparenet_b200.dist.GradArena.pack: autograd hands back ~800 fresh gradient tensors per step; torch._foreach_copy_ moved their
// 330 MB in 11 multi-tensor launches at ~2.3 TB/s.)  The table travels as a KERNEL PARAMETER (24 B per tensor; CUDA 12.1+ accepts
// 32 KB of parameters): no device-side table, no host-to-device copy, and under CUDA-graph capture the launch is an ordinary kernel
// node that carries its table with it.  A block copies 4096 floats.
struct PackEntry {
  const float* src;
  float* dst;
  unsigned n;
  unsigned first_block;
};
constexpr int PACK_CHUNK = 4096;
constexpr int PACK_MAX = 1024;   // entries per launch (24.6 KB of parameters)
struct PackTable {
  PackEntry e[PACK_MAX];
};

__global__ void __launch_bounds__(256) multi_copy_kernel(const __grid_constant__ PackTable tab, int ntab) {
  // binary search: the last entry whose first block is <= this block
  int lo = 0, hi = ntab - 1;
  const unsigned blk = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab.e[mid].first_block <= blk) lo = mid;
    else hi = mid - 1;
  }
  const float* __restrict__ s0 = tab.e[lo].src;
  float* __restrict__ d0 = tab.e[lo].dst;
  const long long off = (long long)(blk - tab.e[lo].first_block) * PACK_CHUNK;
  const long long left = (long long)tab.e[lo].n - off;
  if (left <= 0) return;
  const int n = left < PACK_CHUNK ? (int)left : PACK_CHUNK;
  const float* __restrict__ s = s0 + off;
  float* __restrict__ d = d0 + off;
  if (((((uintptr_t)s) | ((uintptr_t)d)) & 15) == 0) {
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += 256) reinterpret_cast<float4*>(d)[i] = reinterpret_cast<const float4*>(s)[i];
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += 256) d[i] = s[i];
  } else {
    for (int i = threadIdx.x; i < n; i += 256) d[i] = s[i];
  }
}

}  // namespace snb

using namespace snb;

// srcs / dsts / ns: HOST arrays of ntab device pointers and element counts (contiguous float32 tensors, each below 2^32 elements).
SNB_API int snb_multi_copy(const void* const* srcs, void* const* dsts, const long long* ns, int ntab, void* stream) {
  if (ntab < 0) return SNB_EINVAL;
  if (ntab == 0) return SNB_OK;
  if (!srcs || !dsts || !ns) return SNB_EINVAL;
  PackTable tab;  // host staging of one launch's parameters (24.6 KB on the stack; the launch copies it)
  int i = 0;
  while (i < ntab) {
    int m = 0;
    unsigned long long blk = 0;
    for (; i < ntab && m < PACK_MAX; i++) {
      if (ns[i] < 0 || ns[i] > 0xffffffffLL) return SNB_ELIMIT;
      if (ns[i] == 0) continue;
      const unsigned long long nb = ((unsigned long long)ns[i] + PACK_CHUNK - 1) / PACK_CHUNK;
      if (blk + nb > 0x7fffffffULL) break;
      tab.e[m].src = (const float*)srcs[i];
      tab.e[m].dst = (float*)dsts[i];
      tab.e[m].n = (unsigned)ns[i];
      tab.e[m].first_block = (unsigned)blk;
      blk += nb;
      m++;
    }
    if (m == 0) {
      if (i < ntab && ns[i] != 0) return SNB_ELIMIT;   // a single tensor above the grid limit
      continue;
    }
    multi_copy_kernel<<<(unsigned)blk, 256, 0, (cudaStream_t)stream>>>(tab, m);
    SNB_LAUNCH_CHECK();
  }
  return SNB_OK;
}
SNB_API int snb_adam_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr, float beta1, float beta2,
                          float eps, float weight_decay, int step, void* stream) {
  if (step < 1 || (n & 3) != 0) return SNB_EINVAL;
  if (n == 0) return SNB_OK;
  if ((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) != 0) return SNB_EALIGN;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1), rbc2s = (float)(1.0 / sqrt(bc2));
  const size_t n4 = n / 4;
  const size_t want = (n4 + 255) / 256;
  const unsigned grid = (unsigned)(want < (size_t)kNumSMs * 8 ? want : (size_t)kNumSMs * 8);
  adam_flat_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((float4*)param, (const float4*)grad, (float4*)exp_avg, (float4*)exp_avg_sq, n4,
                                                          step_size, 1.0f - beta1, beta2, 1.0f - beta2, eps, rbc2s, weight_decay);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
