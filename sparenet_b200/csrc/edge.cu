// edge.cu -- fused EdgeConv neighbourhood reduction (forward + backward), sm_100a.
//
// Replaces, together with snb_knn, the get_graph_feature -> Conv2d -> BatchNorm2d -> SE -> LeakyReLU -> max_k chain of the
// reference's EdgeConvResFeat (models/sparenet_generator.py:188-242,880-906) WITHOUT ever forming its [B,2C,N,k] /
// [B,C',N,k] tensors.  With W = [W_a | W_b]:  conv([x_j - x_i ; x_i]) = a_j + c_i,  a = W_a x,  c = (W_b - W_a) x  (per-point
// GEMMs, done by the caller).  Everything the rest of the block needs from u[b,ch,i,m] = a[b,ch,idx[b,i,m]] + c[b,ch,i] is
//     umax/umin[b,ch,i] = max_m / min_m u      (max_k commutes with the monotone BN.SE.LeakyReLU tail, sign of gamma picks one)
//     S1[b,ch] = sum_{i,m} u,  S2[b,ch] = sum_{i,m} u^2     (BatchNorm2d batch statistics and the SE squeeze)
// which this kernel produces in one pass over a and c.  The backward kernel is the exact adjoint.
//
// Layout: a, c, umax, umin [B,C,N] channel-major (what the encoder already holds), idx [B,N,k] int32.
// One CTA per (sample, channel) row: the row of `a` (N floats) is staged in shared memory, so the k gathers per point are
// shared-memory reads; idx is read coalesced and re-used by all C channel-CTAs of a sample out of L2.
#include "common.cuh"

namespace snb {

constexpr int EDGE_THREADS = 256;
constexpr int EDGE_MAXK = 32;

__device__ __forceinline__ double block_sum_double(double v, double* red) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  double t = (threadIdx.x < EDGE_THREADS / 32) ? red[threadIdx.x] : 0.0;
  if (w == 0) {
#pragma unroll
    for (int o = 4; o >= 1; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;  // valid in thread 0
}

// CHB channel rows per CTA: the k neighbour indices of a point are loaded ONCE and used for CHB rows staged in shared memory (idx is
// N*k ints per sample = 8x a channel row: with one row per CTA the index reads out of L2 were 70 % of the kernel's traffic).
template <int CHB>
__global__ void __launch_bounds__(EDGE_THREADS) edge_reduce_fwd_kernel(const float* __restrict__ a, const float* __restrict__ c,
                                                                        const int* __restrict__ idx, int C, int N, int k, size_t in_bstride,
                                                                        float* __restrict__ umax, float* __restrict__ umin,
                                                                        unsigned char* __restrict__ smax, unsigned char* __restrict__ smin,
                                                                        double* __restrict__ S1, double* __restrict__ S2,
                                                                        const unsigned char* __restrict__ sel) {
  extern __shared__ float arow[];          // [CHB][N]
  __shared__ double red[EDGE_THREADS / 32];
  const int ch0 = blockIdx.x * CHB, b = blockIdx.y;
  const size_t row0 = ((size_t)b * C + ch0) * N;               // outputs: [B,C,N] contiguous
  const size_t in0 = (size_t)b * in_bstride + (size_t)ch0 * N;  // a, c: batch stride in_bstride (= C*N, or 2*C*N when a and c are the two
                                                                // channel halves of ONE [B,2C,N] GEMM output)
  for (int i = threadIdx.x; i < CHB * N; i += EDGE_THREADS) arow[i] = a[in0 + i];
  __syncthreads();
  const int* __restrict__ ib = idx + (size_t)b * N * k;
  double s1[CHB], s2[CHB];
#pragma unroll
  for (int q = 0; q < CHB; q++) s1[q] = s2[q] = 0.0;
  for (int i = threadIdx.x; i < N; i += EDGE_THREADS) {
    float mx[CHB], mn[CHB], sa[CHB], sq[CHB];
    int ax[CHB], an[CHB];
#pragma unroll
    for (int q = 0; q < CHB; q++) {
      mx[q] = -3.4e38f;
      mn[q] = 3.4e38f;
      sa[q] = sq[q] = 0.f;
      ax[q] = an[q] = 0;
    }
    for (int m = 0; m < k; m++) {
      const int j = ib[(size_t)i * k + m];
#pragma unroll
      for (int q = 0; q < CHB; q++) {
        const float v = arow[q * N + j];
        sa[q] += v;
        sq[q] = __fmaf_rn(v, v, sq[q]);
        if (v > mx[q]) { mx[q] = v; ax[q] = m; }   // first maximum / minimum wins, like torch.max over dim=-1 on CUDA is free to
        if (v < mn[q]) { mn[q] = v; an[q] = m; }
      }
    }
#pragma unroll
    for (int q = 0; q < CHB; q++) {
      const size_t row = row0 + (size_t)q * N;
      const float ci = c[in0 + (size_t)q * N + i];
      if (sel) {  // only the extremum the sign of the channel's BatchNorm weight asks for (written to umax / smax)
        const bool up = sel[ch0 + q] != 0;
        umax[row + i] = (up ? mx[q] : mn[q]) + ci;
        smax[row + i] = (unsigned char)(up ? ax[q] : an[q]);
      } else {
        umax[row + i] = mx[q] + ci;
        umin[row + i] = mn[q] + ci;
        smax[row + i] = (unsigned char)ax[q];
        smin[row + i] = (unsigned char)an[q];
      }
      // sum_m (a_j + c)   and   sum_m (a_j + c)^2 = sum a_j^2 + 2 c sum a_j + k c^2
      s1[q] += (double)sa[q] + (double)k * (double)ci;
      s2[q] += (double)sq[q] + 2.0 * (double)ci * (double)sa[q] + (double)k * (double)ci * (double)ci;
    }
  }
#pragma unroll
  for (int q = 0; q < CHB; q++) {
    const double t1 = block_sum_double(s1[q], red);
    const double t2 = block_sum_double(s2[q], red);
    if (threadIdx.x == 0) {
      S1[(size_t)b * C + ch0 + q] = t1;
      S2[(size_t)b * C + ch0 + q] = t2;
    }
  }
}

// adjoint:  gc_i = gmax_i + gmin_i + k gS1 + 2 gS2 (sum_m a_jm + k c_i)
//           ga_j = sum_{(i,m): idx[i,m] = j} [ gS1 + 2 gS2 (a_j + c_i) + [m = smax_i] gmax_i + [m = smin_i] gmin_i ]
template <int CHB>
__global__ void __launch_bounds__(EDGE_THREADS) edge_reduce_bwd_kernel(const float* __restrict__ a, const float* __restrict__ c,
                                                                        const int* __restrict__ idx, const unsigned char* __restrict__ smax,
                                                                        const unsigned char* __restrict__ smin, const float* __restrict__ gmax,
                                                                        const float* __restrict__ gmin, const double* __restrict__ gS1,
                                                                        const double* __restrict__ gS2, int C, int N, int k, size_t in_bstride,
                                                                        float* __restrict__ ga, float* __restrict__ gc) {
  extern __shared__ float sm[];
  float* arow = sm;                // [CHB][N]
  float* grow = sm + CHB * N;      // [CHB][N]
  const int ch0 = blockIdx.x * CHB, b = blockIdx.y;
  const size_t row0 = ((size_t)b * C + ch0) * N;               // gmax, gmin, slots: [B,C,N] contiguous
  const size_t in0 = (size_t)b * in_bstride + (size_t)ch0 * N;  // a, c, ga, gc: batch stride in_bstride
  for (int i = threadIdx.x; i < CHB * N; i += EDGE_THREADS) {
    arow[i] = a[in0 + i];
    grow[i] = 0.f;
  }
  __syncthreads();
  float g1[CHB], g2[CHB];
#pragma unroll
  for (int q = 0; q < CHB; q++) {
    g1[q] = (float)gS1[(size_t)b * C + ch0 + q];
    g2[q] = 2.f * (float)gS2[(size_t)b * C + ch0 + q];
  }
  const int* __restrict__ ib = idx + (size_t)b * N * k;
  for (int i = threadIdx.x; i < N; i += EDGE_THREADS) {
    float ci[CHB], gx[CHB], gn[CHB], sa[CHB];
    int ax[CHB], an[CHB];
#pragma unroll
    for (int q = 0; q < CHB; q++) {
      const size_t r = row0 + (size_t)q * N + i;
      ci[q] = c[in0 + (size_t)q * N + i];
      gx[q] = gmax[r];
      gn[q] = gmin ? gmin[r] : 0.f;                 // gmin == nullptr: the selected-extremum form
      ax[q] = smax[r];
      an[q] = gmin ? (int)smin[r] : -1;
      sa[q] = 0.f;
    }
    for (int m = 0; m < k; m++) {
      const int j = ib[(size_t)i * k + m];
#pragma unroll
      for (int q = 0; q < CHB; q++) {
        const float v = arow[q * N + j];
        sa[q] += v;
        float g = __fmaf_rn(g2[q], v + ci[q], g1[q]);
        if (m == ax[q]) g += gx[q];
        if (m == an[q]) g += gn[q];
        atomicAdd(&grow[q * N + j], g);
      }
    }
#pragma unroll
    for (int q = 0; q < CHB; q++)
      gc[in0 + (size_t)q * N + i] = gx[q] + gn[q] + (float)k * g1[q] + g2[q] * (sa[q] + (float)k * ci[q]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < CHB * N; i += EDGE_THREADS) ga[in0 + i] = grow[i];
}

}  // namespace snb

using namespace snb;

// channel rows per CTA: 4 when the channel count allows it and the rows fit in shared memory, else 1
static int edge_chb(int C, int N, int rows_per_channel) { return (C % 4 == 0 && (size_t)N * 4 * rows_per_channel * sizeof(float) <= 160 * 1024) ? 4 : 1; }

template <int CHB>
static int edge_launch_fwd(const float* a, const float* c, const int* idx, int B, int C, int N, int k, float* umax, float* umin, unsigned char* smax,
                           unsigned char* smin, double* S1, double* S2, const unsigned char* sel, cudaStream_t s, size_t in_bstride = 0) {
  if (in_bstride == 0) in_bstride = (size_t)C * N;
  const size_t smem = (size_t)N * CHB * sizeof(float);
  if (smem > 48 * 1024) SNB_CUDA(cudaFuncSetAttribute(edge_reduce_fwd_kernel<CHB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  edge_reduce_fwd_kernel<CHB><<<dim3(C / CHB, B), EDGE_THREADS, smem, s>>>(a, c, idx, C, N, k, in_bstride, umax, umin, smax, smin, S1, S2, sel);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

template <int CHB>
static int edge_launch_bwd(const float* a, const float* c, const int* idx, const unsigned char* smax, const unsigned char* smin, const float* gmax,
                           const float* gmin, const double* gS1, const double* gS2, int B, int C, int N, int k, float* ga, float* gc,
                           cudaStream_t s, size_t in_bstride = 0) {
  if (in_bstride == 0) in_bstride = (size_t)C * N;
  const size_t smem = (size_t)N * 2 * CHB * sizeof(float);
  if (smem > 48 * 1024) SNB_CUDA(cudaFuncSetAttribute(edge_reduce_bwd_kernel<CHB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  edge_reduce_bwd_kernel<CHB><<<dim3(C / CHB, B), EDGE_THREADS, smem, s>>>(a, c, idx, smax, smin, gmax, gmin, gS1, gS2, C, N, k, in_bstride, ga, gc);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

static int edge_check(int B, int C, int N, int k) {
  if (B < 0 || C < 0 || N < 0 || k <= 0) return SNB_EINVAL;
  if (k > EDGE_MAXK || k > 255 || B > 65535 || (size_t)N * 8 > 200 * 1024) return SNB_ELIMIT;
  return SNB_OK;
}

SNB_API int snb_edge_reduce_fwd(const float* a, const float* c, const int* idx, int B, int C, int N, int k, float* umax, float* umin,
                                unsigned char* slot_max, unsigned char* slot_min, double* S1, double* S2, void* stream) {
  int rc = edge_check(B, C, N, k);
  if (rc) return rc;
  if (B == 0 || C == 0 || N == 0) return SNB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  return edge_chb(C, N, 1) == 4 ? edge_launch_fwd<4>(a, c, idx, B, C, N, k, umax, umin, slot_max, slot_min, S1, S2, nullptr, s)
                                : edge_launch_fwd<1>(a, c, idx, B, C, N, k, umax, umin, slot_max, slot_min, S1, S2, nullptr, s);
}

SNB_API int snb_edge_reduce_sel_fwd(const float* a, const float* c, const int* idx, const unsigned char* sel_max, int B, int C, int N, int k,
                                    float* ustar, unsigned char* slot, double* S1, double* S2, void* stream) {
  int rc = edge_check(B, C, N, k);
  if (rc) return rc;
  if (!sel_max) return SNB_EINVAL;
  if (B == 0 || C == 0 || N == 0) return SNB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  return edge_chb(C, N, 1) == 4 ? edge_launch_fwd<4>(a, c, idx, B, C, N, k, ustar, nullptr, slot, nullptr, S1, S2, sel_max, s)
                                : edge_launch_fwd<1>(a, c, idx, B, C, N, k, ustar, nullptr, slot, nullptr, S1, S2, sel_max, s);
}

SNB_API int snb_edge_reduce_sel_bwd(const float* a, const float* c, const int* idx, const unsigned char* slot, const float* g_ustar,
                                    const double* gS1, const double* gS2, int B, int C, int N, int k, float* ga, float* gc, void* stream) {
  int rc = edge_check(B, C, N, k);
  if (rc) return rc;
  if (B == 0 || C == 0 || N == 0) return SNB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  return edge_chb(C, N, 2) == 4 ? edge_launch_bwd<4>(a, c, idx, slot, nullptr, g_ustar, nullptr, gS1, gS2, B, C, N, k, ga, gc, s)
                                : edge_launch_bwd<1>(a, c, idx, slot, nullptr, g_ustar, nullptr, gS1, gS2, B, C, N, k, ga, gc, s);
}

SNB_API int snb_edge_reduce_bwd(const float* a, const float* c, const int* idx, const unsigned char* slot_max, const unsigned char* slot_min,
                                const float* g_umax, const float* g_umin, const double* gS1, const double* gS2, int B, int C, int N, int k,
                                float* ga, float* gc, void* stream) {
  int rc = edge_check(B, C, N, k);
  if (rc) return rc;
  if (B == 0 || C == 0 || N == 0) return SNB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  return edge_chb(C, N, 2) == 4 ? edge_launch_bwd<4>(a, c, idx, slot_max, slot_min, g_umax, g_umin, gS1, gS2, B, C, N, k, ga, gc, s)
                                : edge_launch_bwd<1>(a, c, idx, slot_max, slot_min, g_umax, g_umin, gS1, gS2, B, C, N, k, ga, gc, s);
}

// The same two calls for a and c living in ONE [B, 2C, N] tensor (channels [0,C) = a, [C,2C) = c: the output of a single GEMM with the
// stacked weight [W_a ; W_b - W_a]); the backward writes ga and gc into the two halves of ONE [B, 2C, N] gradient, so the data and
// weight gradients of the pair are single GEMMs too and no elementwise add of two data gradients is left.
SNB_API int snb_edge_reduce_sel_fwd_stacked(const float* ac, const int* idx, const unsigned char* sel_max, int B, int C, int N, int k, float* ustar,
                                            unsigned char* slot, double* S1, double* S2, void* stream) {
  int rc = edge_check(B, C, N, k);
  if (rc) return rc;
  if (!sel_max) return SNB_EINVAL;
  if (B == 0 || C == 0 || N == 0) return SNB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const float* a = ac;
  const float* c = ac + (size_t)C * N;
  const size_t bs = (size_t)2 * C * N;
  return edge_chb(C, N, 1) == 4 ? edge_launch_fwd<4>(a, c, idx, B, C, N, k, ustar, nullptr, slot, nullptr, S1, S2, sel_max, s, bs)
                                : edge_launch_fwd<1>(a, c, idx, B, C, N, k, ustar, nullptr, slot, nullptr, S1, S2, sel_max, s, bs);
}

SNB_API int snb_edge_reduce_sel_bwd_stacked(const float* ac, const int* idx, const unsigned char* slot, const float* g_ustar, const double* gS1,
                                            const double* gS2, int B, int C, int N, int k, float* gac, void* stream) {
  int rc = edge_check(B, C, N, k);
  if (rc) return rc;
  if (B == 0 || C == 0 || N == 0) return SNB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t bs = (size_t)2 * C * N;
  return edge_chb(C, N, 2) == 4
             ? edge_launch_bwd<4>(ac, ac + (size_t)C * N, idx, slot, nullptr, g_ustar, nullptr, gS1, gS2, B, C, N, k, gac, gac + (size_t)C * N, s, bs)
             : edge_launch_bwd<1>(ac, ac + (size_t)C * N, idx, slot, nullptr, g_ustar, nullptr, gS1, gS2, B, C, N, k, gac, gac + (size_t)C * N, s, bs);
}
