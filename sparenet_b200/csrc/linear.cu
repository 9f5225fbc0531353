// linear.cu -- fp32 nn.Linear for SMALL batches (B <= 32 rows): forward, data gradient, weight gradient, sm_100a.
//
// The encoder -> decoder bridge of the generator is three fully connected layers on a [B, 4096] activation (reference
// models/sparenet_generator.py:85-120 SpareNetEncode.linear, :289-330 SpareNetDecode.mlp): 64 MB of fp32 weights each, 32 rows of
// activations.  The reference runs them as fp32 cuBLAS GEMMs (torch.backends.cuda.matmul.allow_tf32 is off by default), and so do we:
// exact fp32 FMAs, no tensor cores.  With 32 rows the product is a weight STREAM -- every weight is used 32 times -- and the library's
// 128x64 SIMT tiles reach a tenth of the streaming rate (~100 us forward, ~65 us data gradient per layer).  Here:
//   forward   y[b,o] = sum_k x[b,k] W[o,k] + bias[o]:   a block owns 128 outputs x one K slice; a thread = (4 outputs, 8 of the batch rows),
//             the weight rows streamed with 128-bit loads (the 4 threads of an output group share the addresses), the activation slice in
//             shared memory
//   dgrad     gx[b,k] = sum_o gy[b,o] W[o,k]:            a block owns 128 inputs x one O slice; a thread = (4 consecutive k, 8 batch rows),
//             weight rows read as coalesced 128-bit loads, the gradient slice in shared memory
//   wgrad     gW[o,k] = sum_b gy[b,o] x[b,k], gbias[o] = sum_b gy[b,o]:  128 x 128 output tile per block, 8 x 8 per thread
// Split slices write partial sums to a workspace and a second pass adds them in a FIXED order (deterministic, unlike atomics).
#include "common.cuh"

namespace snb {

constexpr int LIN_MAXB = 32;
constexpr int LIN_FWD_OT = 128;   // outputs per block (forward): 32 groups of 4
constexpr int LIN_FWD_KC = 256;   // activation chunk staged in shared memory (forward): 32 x 256 floats = 32 KB
constexpr int LIN_DG_KT = 128;    // inputs per block (dgrad)
constexpr int LIN_DG_OC = 256;    // gradient chunk staged in shared memory (dgrad)

// ---- forward: grid (ceil(O/128), S); block 128 threads = 32 output groups x 4 batch groups; a thread owns 4 outputs x 8 batch rows, so
//      the 8 shared-memory vectors of a k step feed 128 FMAs (with one output per thread the kernel was bound by the shared-memory pipe:
//      8 LDS.128 per 32 FMAs); partial[s][b][o] ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W, int B, int K, int O, int kslice,
                                                          float* __restrict__ part) {
  __shared__ __align__(16) float xs[LIN_MAXB][LIN_FWD_KC + 4];   // +16 B per row: the 4 rows a warp reads at once hit different banks
  const int og = threadIdx.x >> 2, bq = threadIdx.x & 3;
  const int o0 = blockIdx.x * LIN_FWD_OT + og * 4;
  const int k0 = blockIdx.y * kslice, k1 = min(K, k0 + kslice);
  const float* __restrict__ wr[4];
#pragma unroll
  for (int j = 0; j < 4; j++) wr[j] = W + (size_t)(o0 + j < O ? o0 + j : 0) * K;
  float acc[4][8];
#pragma unroll
  for (int j = 0; j < 4; j++)
#pragma unroll
    for (int i = 0; i < 8; i++) acc[j][i] = 0.f;
  for (int kc = k0; kc < k1; kc += LIN_FWD_KC) {
    const int kn = min(LIN_FWD_KC, k1 - kc);   // multiple of 4
    __syncthreads();
    for (int e = threadIdx.x; e < LIN_MAXB * (LIN_FWD_KC / 4); e += 128) {
      const int b = e / (LIN_FWD_KC / 4), q = e - b * (LIN_FWD_KC / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b < B && 4 * q < kn) v = *reinterpret_cast<const float4*>(x + (size_t)b * K + kc + 4 * q);
      *reinterpret_cast<float4*>(&xs[b][4 * q]) = v;
    }
    __syncthreads();
    if (o0 < O) {
#pragma unroll 2
      for (int k = 0; k < kn; k += 4) {
        float4 w[4];
#pragma unroll
        for (int j = 0; j < 4; j++) w[j] = *reinterpret_cast<const float4*>(wr[j] + kc + k);
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const float4 xv = *reinterpret_cast<const float4*>(&xs[i * 4 + bq][k]);   // batch row i*4 + bq: adjacent rows within a warp
#pragma unroll
          for (int j = 0; j < 4; j++) {
            acc[j][i] = __fmaf_rn(w[j].x, xv.x, acc[j][i]);
            acc[j][i] = __fmaf_rn(w[j].y, xv.y, acc[j][i]);
            acc[j][i] = __fmaf_rn(w[j].z, xv.z, acc[j][i]);
            acc[j][i] = __fmaf_rn(w[j].w, xv.w, acc[j][i]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; j++) {
    if (o0 + j >= O) continue;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int b = i * 4 + bq;
      if (b < B) part[((size_t)blockIdx.y * B + b) * O + o0 + j] = acc[j][i];
    }
  }
}

// out[b,j] = bias[j] + sum_s part[s][b][j]  (fixed order)
__global__ void __launch_bounds__(256) linear_reduce_kernel(const float* __restrict__ part, const float* __restrict__ bias, int B, int J, int S,
                                                             float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * J) return;
  float a = bias ? bias[i % J] : 0.f;
  for (int s = 0; s < S; s++) a += part[(size_t)s * B * J + i];
  out[i] = a;
}

// ---- dgrad: grid (ceil(K/128), S); block 128 threads: thread = (4 consecutive k, 8 batch rows); partial[s][b][k] ------------------------
__global__ void __launch_bounds__(128) linear_dgrad_kernel(const float* __restrict__ gy, const float* __restrict__ W, int B, int K, int O, int oslice,
                                                            float* __restrict__ part) {
  __shared__ __align__(16) float gs[LIN_DG_OC][LIN_MAXB];   // [o][b]: a thread reads its 8 consecutive b of one o as two 128-bit loads
  const int kq = threadIdx.x >> 2, bq = threadIdx.x & 3;
  const int k = blockIdx.x * LIN_DG_KT + 4 * kq;
  const int o0 = blockIdx.y * oslice, o1 = min(O, o0 + oslice);
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
  for (int oc = o0; oc < o1; oc += LIN_DG_OC) {
    const int on = min(LIN_DG_OC, o1 - oc);
    __syncthreads();
    for (int e = threadIdx.x; e < LIN_MAXB * LIN_DG_OC; e += 128) {
      const int b = e / LIN_DG_OC, oo = e - b * LIN_DG_OC;   // consecutive threads = consecutive o: coalesced rows of gy
      gs[oo][b] = (b < B && oo < on) ? gy[(size_t)b * O + oc + oo] : 0.f;
    }
    __syncthreads();
    if (k < K) {
      // software pipeline over groups of 8 weight rows: the next group's eight 128-bit loads are in flight while the current group's
      // 256 FMAs issue (the loop is a stream over W; without the explicit prefetch every group waited out its own load latency)
      float4 wn[8];
#pragma unroll
      for (int u = 0; u < 8; u++) wn[u] = u < on ? *reinterpret_cast<const float4*>(W + (size_t)(oc + u) * K + k) : make_float4(0.f, 0.f, 0.f, 0.f);
      for (int og = 0; og < on; og += 8) {
        float4 wc[8];
#pragma unroll
        for (int u = 0; u < 8; u++) wc[u] = wn[u];
        if (og + 8 < on) {
#pragma unroll
          for (int u = 0; u < 8; u++)
            wn[u] = og + 8 + u < on ? *reinterpret_cast<const float4*>(W + (size_t)(oc + og + 8 + u) * K + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
          // rows past the slice hold zeros in shared memory and in wc: they add nothing
          const float4 ga = *reinterpret_cast<const float4*>(&gs[og + u][bq * 8]), gb = *reinterpret_cast<const float4*>(&gs[og + u][bq * 8 + 4]);
          const float g[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
          for (int i = 0; i < 8; i++) {
            acc[i][0] = __fmaf_rn(g[i], wc[u].x, acc[i][0]);
            acc[i][1] = __fmaf_rn(g[i], wc[u].y, acc[i][1]);
            acc[i][2] = __fmaf_rn(g[i], wc[u].z, acc[i][2]);
            acc[i][3] = __fmaf_rn(g[i], wc[u].w, acc[i][3]);
          }
        }
      }
    }
  }
  if (k < K) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int b = bq * 8 + i;
      if (b < B) *reinterpret_cast<float4*>(part + ((size_t)blockIdx.y * B + b) * K + k) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
  }
}

// ---- wgrad: 128 x 128 tile of gW per block (256 threads, 8 x 8 each: four 128-bit shared-memory loads per 64 FMAs); gbias by the blocks of
//      the first k tile --------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) linear_wgrad_kernel(const float* __restrict__ gy, const float* __restrict__ x, int B, int K, int O,
                                                            float* __restrict__ gW, float* __restrict__ gbias) {
  __shared__ __align__(16) float gt[LIN_MAXB][128];
  __shared__ __align__(16) float xt[LIN_MAXB][128];
  const int o0 = blockIdx.y * 128, k0 = blockIdx.x * 128;
  for (int e = threadIdx.x; e < LIN_MAXB * 128; e += 256) {
    const int b = e >> 7, j = e & 127;
    gt[b][j] = (b < B && o0 + j < O) ? gy[(size_t)b * O + o0 + j] : 0.f;
    xt[b][j] = (b < B && k0 + j < K) ? x[(size_t)b * K + k0 + j] : 0.f;
  }
  __syncthreads();
  // thread (ty, tx): outputs o0 + 4 ty + {0..3} and o0 + 64 + 4 ty + {0..3}, inputs k0 + 4 tx + {0..3} and k0 + 64 + 4 tx + {0..3}
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = 0.f;
#pragma unroll 4
  for (int b = 0; b < LIN_MAXB; b++) {
    const float4 g0 = *reinterpret_cast<const float4*>(&gt[b][4 * ty]), g1 = *reinterpret_cast<const float4*>(&gt[b][64 + 4 * ty]);
    const float4 x0 = *reinterpret_cast<const float4*>(&xt[b][4 * tx]), x1 = *reinterpret_cast<const float4*>(&xt[b][64 + 4 * tx]);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, xx[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j < 8; j++) acc[i][j] = __fmaf_rn(gg[i], xx[j], acc[i][j]);
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int o = o0 + (i < 4 ? 4 * ty + i : 64 + 4 * ty + (i - 4));
    if (o >= O) continue;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int kk = k0 + 64 * h + 4 * tx;
      if (kk + 3 < K) {
        *reinterpret_cast<float4*>(gW + (size_t)o * K + kk) = make_float4(acc[i][4 * h], acc[i][4 * h + 1], acc[i][4 * h + 2], acc[i][4 * h + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; j++)
          if (kk + j < K) gW[(size_t)o * K + kk + j] = acc[i][4 * h + j];
      }
    }
  }
  if (gbias != nullptr && blockIdx.x == 0 && threadIdx.x < 128 && o0 + threadIdx.x < O) {
    float s = 0.f;
    for (int b = 0; b < B; b++) s += gt[b][threadIdx.x];
    gbias[o0 + threadIdx.x] = s;
  }
}

static int lin_splits(int blocks, int len, int chunk) {
  // enough blocks for ~4 per SM, slices a multiple of `chunk`
  int s = (4 * kNumSMs + blocks - 1) / blocks;
  const int maxs = (len + chunk - 1) / chunk;
  if (s > maxs) s = maxs;
  if (s < 1) s = 1;
  return s;
}

}  // namespace snb

using namespace snb;

static int lin_fwd_splits(int K, int O, int* kslice) {
  const int blocks = (O + LIN_FWD_OT - 1) / LIN_FWD_OT;
  int s = lin_splits(blocks, K, LIN_FWD_KC);
  int sl = ((K + s - 1) / s + LIN_FWD_KC - 1) / LIN_FWD_KC * LIN_FWD_KC;
  s = (K + sl - 1) / sl;
  *kslice = sl;
  return s;
}
static int lin_dg_splits(int K, int O, int* oslice) {
  const int blocks = (K + LIN_DG_KT - 1) / LIN_DG_KT;
  int s = lin_splits(blocks, O, LIN_DG_OC);
  int sl = ((O + s - 1) / s + LIN_DG_OC - 1) / LIN_DG_OC * LIN_DG_OC;
  s = (O + sl - 1) / sl;
  *oslice = sl;
  return s;
}

// floats of workspace either direction may need
SNB_API size_t snb_linear_workspace_floats(int B, int K, int O) {
  if (B <= 0 || K <= 0 || O <= 0) return 0;
  int sl;
  const size_t f = (size_t)lin_fwd_splits(K, O, &sl) * B * O, d = (size_t)lin_dg_splits(K, O, &sl) * B * K;
  return f > d ? f : d;
}

SNB_API int snb_linear_fwd(const float* x, const float* W, const float* bias, int B, int K, int O, float* y, float* workspace, void* stream) {
  if (B < 0 || K <= 0 || O <= 0) return SNB_EINVAL;
  if (B > LIN_MAXB || (K & 3) != 0) return SNB_ELIMIT;
  if (B == 0) return SNB_OK;
  if (!workspace) return SNB_EWORKSPACE;
  if ((((uintptr_t)x | (uintptr_t)W) & 15) != 0) return SNB_EALIGN;
  cudaStream_t s = (cudaStream_t)stream;
  int kslice;
  const int S = lin_fwd_splits(K, O, &kslice);
  linear_fwd_kernel<<<dim3((unsigned)((O + LIN_FWD_OT - 1) / LIN_FWD_OT), (unsigned)S), 128, 0, s>>>(x, W, B, K, O, kslice, workspace);
  SNB_LAUNCH_CHECK();
  linear_reduce_kernel<<<(unsigned)(((size_t)B * O + 255) / 256), 256, 0, s>>>(workspace, bias, B, O, S, y);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

SNB_API int snb_linear_dgrad(const float* gy, const float* W, int B, int K, int O, float* gx, float* workspace, void* stream) {
  if (B < 0 || K <= 0 || O <= 0) return SNB_EINVAL;
  if (B > LIN_MAXB || (K & 3) != 0) return SNB_ELIMIT;
  if (B == 0) return SNB_OK;
  if (!workspace) return SNB_EWORKSPACE;
  if ((((uintptr_t)W | (uintptr_t)workspace) & 15) != 0) return SNB_EALIGN;
  cudaStream_t s = (cudaStream_t)stream;
  int oslice;
  const int S = lin_dg_splits(K, O, &oslice);
  linear_dgrad_kernel<<<dim3((unsigned)((K + LIN_DG_KT - 1) / LIN_DG_KT), (unsigned)S), 128, 0, s>>>(gy, W, B, K, O, oslice, workspace);
  SNB_LAUNCH_CHECK();
  linear_reduce_kernel<<<(unsigned)(((size_t)B * K + 255) / 256), 256, 0, s>>>(workspace, nullptr, B, K, S, gx);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

SNB_API int snb_linear_wgrad(const float* gy, const float* x, int B, int K, int O, float* gW, float* gbias, void* stream) {
  if (B < 0 || K <= 0 || O <= 0) return SNB_EINVAL;
  if (B > LIN_MAXB || (K & 3) != 0) return SNB_ELIMIT;
  if (((uintptr_t)gW & 15) != 0) return SNB_EALIGN;
  linear_wgrad_kernel<<<dim3((unsigned)((K + 127) / 128), (unsigned)((O + 127) / 128)), 256, 0, (cudaStream_t)stream>>>(gy, x, B, K, O, gW, gbias);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
