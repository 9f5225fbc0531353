// tails.cu -- the closed-form BatchNorm . SE tail of a dense layer as ONE launch per direction, sm_100a.
//
// The generator never applies BatchNorm / SE / ReLU to an activation tensor: each dense layer's tail collapses to one
// (scale, shift) pair per (sample, channel) computed from the ROW statistics of the layer's pre-activation h [B,C,L]
// (reference models/sparenet_generator.py:593-646: conv -> BatchNorm1d -> SELayer1D -> ReLU of PointNetRes):
//     m      = row_mean(h) + rb                      rb = the conv bias ([C]) or a per-sample bias ([B,C]): never added to h
//     mean_c = avg_b m,  var_c = avg_b row_var(h) + avg_b (m - mean_c)^2        (within-row + between-row: no cancellation)
//     sc_c   = gamma_c * rsqrt(var_c + eps),  sh_c = beta_c - sc_c * mean_c     (BatchNorm, batch statistics in train mode)
//     z      = m * sc_c + sh_c                        (the SE squeeze: mean over the points of BN(h + rb))
//     gate   = sigmoid(W2 relu(W1 z))                 (SE excitation, W1 [H,C], W2 [C,H], no biases)
//     S      = gate * sc_c,   T = gate * sh_c + rb * S                        => tail(h) = relu(S*h + T)
// As PyTorch glue this is ~37 launches forward and ~48 backward (each ~2 us of kernel + ~2.4 us of launch gap inside the step's CUDA
// graph: ~360 us per tail, ten tails per step in the refiner).  The data is tiny (B*C <= 16 K values) but the work is a chain of
// dependent phases, i.e. LATENCY: a first version on ONE thread block took 60-100 us (the two small matrix products alone are
// ~50 K warp instructions on a single SM).  So one CLUSTER of 8 thread blocks runs each direction: every phase is spread over the
// 8192 threads / 256 warps of the cluster with all of its loads independent, phases are separated by cluster barriers, and what
// crosses a barrier goes through the caller's save / scratch buffers in global memory (L2; the barrier's acquire invalidates L1).
#include <math.h>
#include "common.cuh"

namespace snb {

constexpr int TAIL_THREADS = 1024;
constexpr int TAIL_CLUSTER = 8;

__device__ __forceinline__ float tail_rb(const float* __restrict__ rb, int rb_batched, int b, int c, int C) {
  return rb ? (rb_batched ? rb[b * C + c] : rb[c]) : 0.f;
}
__device__ __forceinline__ float tail_warp_sum(float v) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

// save layout (floats): mean[C] inv[C] sc[C] sh[C] z[B*C] gate[B*C] a1[B*H] w2T[H*C]
// w2T = W2 transposed, written by the forward's first phase: the gate phase (a thread per (sample, channel), consecutive lanes =
// consecutive channels) and the backward's W2^T ga2 phase read W2 along the channels, which in its own [C,H] layout is a stride of H
// floats between lanes -- 32 cache lines per warp load, 131 us per direction for the encoder's C = 1024, H = 64 tail.
__global__ void __launch_bounds__(TAIL_THREADS, 1) bn_se_tail_fwd_kernel(const float* __restrict__ m_bc, const float* __restrict__ v_bc,
                                                                          const float* __restrict__ rb, int rb_batched, const float* __restrict__ g,
                                                                          const float* __restrict__ beta, const float* __restrict__ w1,
                                                                          const float* __restrict__ w2, int B, int C, int H, float eps, int training,
                                                                          float momentum, float unbias, float* run_mean, float* run_var,
                                                                          long long* num_batches, float* __restrict__ S, float* __restrict__ T,
                                                                          float* save) {
  const int nt = (int)(cluster_nctarank() * blockDim.x), ct = (int)(cluster_ctarank() * blockDim.x + threadIdx.x);
  const int lane = ct & 31, cw = ct >> 5, nw = nt >> 5;
  const int BC = B * C;
  float* mean = save;
  float* inv = mean + C;
  float* sc = inv + C;
  float* sh = sc + C;
  float* z = sh + C;
  float* gate = z + (size_t)BC;
  float* a1 = gate + (size_t)BC;
  float* w2T = a1 + (size_t)B * H;
  const float rB = 1.0f / (float)B;
  for (int i = ct; i < C * H; i += nt) {
    const int c = i / H, h = i - c * H;
    w2T[(size_t)h * C + c] = w2[i];
  }
  // ---- BatchNorm statistics and the per-channel scale / shift: a warp per channel, lanes over the samples ----
  for (int c = cw; c < C; c += nw) {
    float mu, var;
    if (training) {
      float sm = 0.f, sv = 0.f;
      for (int b = lane; b < B; b += 32) {
        sm += m_bc[b * C + c] + tail_rb(rb, rb_batched, b, c, C);
        sv += v_bc[b * C + c];
      }
      mu = tail_warp_sum(sm) * rB;
      float sd = 0.f;
      for (int b = lane; b < B; b += 32) {
        const float d = (m_bc[b * C + c] + tail_rb(rb, rb_batched, b, c, C)) - mu;
        sd += d * d;
      }
      var = tail_warp_sum(sv) * rB + tail_warp_sum(sd) * rB;
      if (lane == 0 && run_mean) {  // nn.BatchNorm bookkeeping: running statistics with the unbiased variance
        run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * mu;
        run_var[c] = (1.f - momentum) * run_var[c] + momentum * (var * unbias);
      }
    } else {
      mu = run_mean[c];
      var = run_var[c];
    }
    if (lane == 0) {
      const float iv = rsqrtf(var + eps);
      const float s_c = g[c] * iv;
      mean[c] = mu;
      inv[c] = iv;
      sc[c] = s_c;
      sh[c] = beta[c] - s_c * mu;
    }
  }
  if (ct == 0 && training && num_batches) *num_batches += 1;
  cluster_sync_all();
  // ---- a1 = relu(W1 z), z = m*sc + sh formed on the fly (and stored by the warps of hidden unit 0): a warp per (sample, hidden
  //      unit), lanes over the channels ----
  for (int o = cw; o < B * H; o += nw) {
    const int b = o / H, h = o - b * H;
    float acc = 0.f;
#pragma unroll 4
    for (int c = lane; c < C; c += 32) {
      const float zz = (m_bc[b * C + c] + tail_rb(rb, rb_batched, b, c, C)) * sc[c] + sh[c];
      if (h == 0) z[b * C + c] = zz;
      acc = __fmaf_rn(w1[h * C + c], zz, acc);
    }
    acc = tail_warp_sum(acc);
    if (lane == 0) a1[o] = fmaxf(acc, 0.f);
  }
  cluster_sync_all();
  // ---- gate = sigmoid(W2 a1); the folded scale / shift ----
  for (int i = ct; i < BC; i += nt) {
    const int b = i / C, c = i - b * C;
    float acc = 0.f;
#pragma unroll 8
    for (int h = 0; h < H; h++) acc = __fmaf_rn(w2T[(size_t)h * C + c], a1[b * H + h], acc);
    const float gt = 1.0f / (1.0f + expf(-acc));
    gate[i] = gt;
    const float s_ = gt * sc[c];
    S[i] = s_;
    T[i] = gt * sh[c] + tail_rb(rb, rb_batched, b, c, C) * s_;
  }
}

// scratch layout (floats): ga2[B*C] gz[B*C] p1[B*C] p2[B*C] gp1[B*H]
__global__ void __launch_bounds__(TAIL_THREADS, 1) bn_se_tail_bwd_kernel(const float* __restrict__ gS, const float* __restrict__ gT,
                                                                          const float* __restrict__ m_bc, const float* __restrict__ rb, int rb_batched,
                                                                          const float* __restrict__ g, const float* __restrict__ w1,
                                                                          const float* __restrict__ w2, int B, int C, int H, int training,
                                                                          const float* save, float* scratch, float* __restrict__ gm,
                                                                          float* __restrict__ gv, float* __restrict__ grb, float* __restrict__ gg,
                                                                          float* __restrict__ gbeta, float* __restrict__ gw1, float* __restrict__ gw2) {
  const int nt = (int)(cluster_nctarank() * blockDim.x), ct = (int)(cluster_ctarank() * blockDim.x + threadIdx.x);
  const int lane = ct & 31, cw = ct >> 5, nw = nt >> 5;
  const int BC = B * C;
  const float* mean = save;
  const float* inv = mean + C;
  const float* sc = inv + C;
  const float* sh = sc + C;
  const float* z = sh + C;
  const float* gate = z + (size_t)BC;
  const float* a1 = gate + (size_t)BC;
  const float* w2T = a1 + (size_t)B * H;
  float* ga2 = scratch;
  float* gz = ga2 + (size_t)BC;
  float* p1 = gz + (size_t)BC;
  float* p2 = p1 + (size_t)BC;
  float* gp1 = p2 + (size_t)BC;
  const float rB = 1.0f / (float)B;
  // ---- through S, T to the gate's pre-activation; the direct terms of the per-channel scale / shift gradients ----
  for (int i = ct; i < BC; i += nt) {
    const int b = i / C, c = i - b * C;
    const float gt_ = gT[i], ga = gate[i];
    const float gst = gS[i] + gt_ * tail_rb(rb, rb_batched, b, c, C);
    ga2[i] = (gst * sc[c] + gt_ * sh[c]) * ga * (1.0f - ga);
    p1[i] = gst * ga;
    p2[i] = gt_ * ga;
  }
  cluster_sync_all();
  // ---- gp1 = (W2^T ga2) masked by the ReLU: a warp per (sample, hidden unit), lanes over the channels ----
  for (int o = cw; o < B * H; o += nw) {
    const int b = o / H, h = o - b * H;
    float acc = 0.f;
#pragma unroll 4
    for (int c = lane; c < C; c += 32) acc = __fmaf_rn(w2T[(size_t)h * C + c], ga2[b * C + c], acc);
    acc = tail_warp_sum(acc);
    if (lane == 0) gp1[o] = a1[o] > 0.f ? acc : 0.f;
  }
  cluster_sync_all();
  // ---- the two weight gradients (sums over the samples) and gz = W1^T gp1 with the remaining per-element terms ----
  for (int i = ct; i < C * H; i += nt) {
    const int c = i / H, h = i - c * H;     // gw2 [C,H]
    float acc = 0.f;
#pragma unroll 8
    for (int b = 0; b < B; b++) acc = __fmaf_rn(ga2[b * C + c], a1[b * H + h], acc);
    gw2[i] = acc;
    const int h1 = i / C, c1 = i - h1 * C;  // gw1 [H,C]
    float acc1 = 0.f;
#pragma unroll 8
    for (int b = 0; b < B; b++) acc1 = __fmaf_rn(gp1[b * H + h1], z[b * C + c1], acc1);
    gw1[i] = acc1;
  }
  for (int i = ct; i < BC; i += nt) {
    const int b = i / C, c = i - b * C;
    float acc = 0.f;
#pragma unroll 8
    for (int h = 0; h < H; h++) acc = __fmaf_rn(w1[h * C + c], gp1[b * H + h], acc);
    gz[i] = acc;
    p1[i] += acc * (m_bc[i] + tail_rb(rb, rb_batched, b, c, C));
    p2[i] += acc;
  }
  cluster_sync_all();
  // ---- per channel (a warp each, lanes over the samples): scale / shift gradients -> gamma, beta, batch mean and variance, then
  //      straight back to the row statistics and the bias ----
  for (int c = cw; c < C; c += nw) {
    float gsc = 0.f, gsh = 0.f;
    for (int b = lane; b < B; b += 32) {
      gsc += p1[b * C + c];
      gsh += p2[b * C + c];
    }
    gsc = tail_warp_sum(gsc);
    gsh = tail_warp_sum(gsh);
    const float mu = mean[c], iv = inv[c], s_c = sc[c];
    const float gst_c = gsc - gsh * mu;
    const float gvar = training ? -0.5f * gst_c * g[c] * iv * iv * iv : 0.f;
    const float gmean = training ? -gsh * s_c : 0.f;
    if (lane == 0) {
      gbeta[c] = gsh;
      gg[c] = gst_c * iv;
    }
    float grb_c = 0.f;
    for (int b = lane; b < B; b += 32) {
      const int i = b * C + c;
      const float r_ = tail_rb(rb, rb_batched, b, c, C);
      const float gm_ = gz[i] * s_c + (gmean + gvar * 2.0f * ((m_bc[i] + r_) - mu)) * rB;
      gm[i] = gm_;
      gv[i] = gvar * rB;
      const float gr = gm_ + gT[i] * gate[i] * s_c;   // m = row mean + rb, and T's direct term rb * S
      if (grb && rb_batched) grb[i] = gr;
      grb_c += gr;
    }
    grb_c = tail_warp_sum(grb_c);
    if (grb && !rb_batched && lane == 0) grb[c] = grb_c;
  }
}

static int tail_launch_cfg(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* at, cudaStream_t s) {
  cfg = {};
  cfg.gridDim = dim3(TAIL_CLUSTER);
  cfg.blockDim = dim3(TAIL_THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = TAIL_CLUSTER;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return 0;
}

}  // namespace snb

using namespace snb;

SNB_API size_t snb_bn_se_tail_save_floats(int B, int C, int H) {
  if (B <= 0 || C <= 0 || H <= 0) return 0;
  return (size_t)4 * C + (size_t)2 * B * C + (size_t)B * H + (size_t)C * H;
}

SNB_API size_t snb_bn_se_tail_scratch_floats(int B, int C, int H) {
  if (B <= 0 || C <= 0 || H <= 0) return 0;
  return (size_t)4 * B * C + (size_t)B * H;
}

SNB_API int snb_bn_se_tail_fwd(const float* row_mean, const float* row_var, const float* row_bias, int bias_per_sample, const float* gamma,
                               const float* beta, const float* w1, const float* w2, int B, int C, int H, float eps, int training, float momentum,
                               float unbias, float* running_mean, float* running_var, long long* num_batches_tracked, float* scale, float* shift,
                               float* save, void* stream) {
  if (B < 0 || C < 0 || H < 0) return SNB_EINVAL;
  if (B == 0 || C == 0) return SNB_OK;
  if (H == 0 || (!training && (!running_mean || !running_var))) return SNB_EINVAL;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute at[1];
  tail_launch_cfg(cfg, at, (cudaStream_t)stream);
  SNB_CUDA(cudaLaunchKernelEx(&cfg, bn_se_tail_fwd_kernel, row_mean, row_var, row_bias, bias_per_sample, gamma, beta, w1, w2, B, C, H, eps, training,
                              momentum, unbias, running_mean, running_var, num_batches_tracked, scale, shift, save));
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

SNB_API int snb_bn_se_tail_bwd(const float* grad_scale, const float* grad_shift, const float* row_mean, const float* row_bias, int bias_per_sample,
                               const float* gamma, const float* w1, const float* w2, int B, int C, int H, int training, const float* save,
                               float* scratch, float* grad_row_mean, float* grad_row_var, float* grad_row_bias, float* grad_gamma, float* grad_beta,
                               float* grad_w1, float* grad_w2, void* stream) {
  if (B < 0 || C < 0 || H < 0) return SNB_EINVAL;
  if (B == 0 || C == 0) return SNB_OK;
  if (H == 0) return SNB_EINVAL;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute at[1];
  tail_launch_cfg(cfg, at, (cudaStream_t)stream);
  SNB_CUDA(cudaLaunchKernelEx(&cfg, bn_se_tail_bwd_kernel, grad_scale, grad_shift, row_mean, row_bias, bias_per_sample, gamma, w1, w2, B, C, H,
                              training, save, (float*)scratch, grad_row_mean, grad_row_var, grad_row_bias, grad_gamma, grad_beta, grad_w1, grad_w2));
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

// ---- the decoders' tail: instance norm -> AdaIN -> BatchNorm -> SE -> ReLU folded to one scale/shift --------------------------------
// (reference models/sparenet_generator.py:984-1062: StyleBasedAdaIn / AdaptiveInstanceNorm1d / GridDecoder; all 32 primitives at once)
// Per primitive p, channel c, sample b, with the row statistics (m, s2) of the pre-activation h[p,c,b,:] and the style (w, t)[b,c]:
//     r = rsqrt(s2 + eps), v = s2 / (s2 + eps)                        instance norm: x_hat = (h - m) r has mean 0, variance v
//     mu_c = avg_b t,  q_c = avg_b (w^2 v + t^2) - mu_c^2,  inv = rsqrt(q + eps)          BatchNorm of AdaIN(x_hat) = w x_hat + t
//     z = gamma inv (t - mu) + beta                                   SE squeeze;  gate = sigmoid(W2 relu(W1 z))
//     A = gate gamma inv w,  D = gate z          =>   tail(h) = relu(sc h + sh),  sc = A r,  sh = D - sc m
// One cluster of 4 thread blocks per primitive (128 blocks for the 32 primitives), phases separated by cluster barriers, what
// crosses a barrier in global scratch.  Channel phases: a warp per channel, lanes = samples (B <= 32).
namespace snb {

constexpr int ADT_CLUSTER = 4;

// save per primitive (floats): inv[C] mu[C] z[C*B] gate[C*B] a1[H*B]
__host__ __device__ inline size_t adt_save_floats(int C, int B, int H) { return (size_t)2 * C + (size_t)2 * C * B + (size_t)H * B; }
// scratch per primitive (floats): ga2[C*B] gz1[C*B] gwt1[C*B] gp1[H*B] ggam_part[C] ginv_part[C]
__host__ __device__ inline size_t adt_scratch_floats(int C, int B, int H) { return (size_t)3 * C * B + (size_t)H * B + (size_t)2 * C; }

__global__ void __launch_bounds__(TAIL_THREADS, 1) adain_tail_fwd_kernel(const float* __restrict__ mean, const float* __restrict__ var,
                                                                          const float* __restrict__ wsty, const float* __restrict__ bsty,
                                                                          const float* __restrict__ gam, const float* __restrict__ bet,
                                                                          const float* __restrict__ w1, const float* __restrict__ w2, int C, int Cp,
                                                                          int B, int H, float eps, float* __restrict__ sc, float* __restrict__ sh,
                                                                          float* save_all, float* __restrict__ bn_mu, float* __restrict__ bn_var) {
  const int p = blockIdx.x / ADT_CLUSTER;
  const int nt = ADT_CLUSTER * blockDim.x, ct = (int)(cluster_ctarank() * blockDim.x + threadIdx.x);
  const int lane = ct & 31, cw = ct >> 5, nw = nt >> 5;
  float* save = save_all + (size_t)p * adt_save_floats(C, B, H);
  float* inv = save;
  float* mu = inv + C;
  float* z = mu + C;
  float* gate = z + (size_t)C * B;
  float* a1 = gate + (size_t)C * B;
  const float* __restrict__ mp = mean + (size_t)p * Cp * B;
  const float* __restrict__ vp = var + (size_t)p * Cp * B;
  const float* __restrict__ g_ = gam + (size_t)p * C;
  const float* __restrict__ b_ = bet + (size_t)p * C;
  const float* __restrict__ w1p = w1 + (size_t)p * H * C;
  const float* __restrict__ w2p = w2 + (size_t)p * C * H;
  const float rB = 1.0f / (float)B;
  const bool on = lane < B;
  // ---- BatchNorm statistics of AdaIN(x_hat) and the SE squeeze ----
  for (int c = cw; c < C; c += nw) {
    const float t = on ? bsty[lane * C + c] : 0.f, w = on ? wsty[lane * C + c] : 0.f;
    const float s2 = on ? vp[c * B + lane] : 0.f;
    const float v = s2 / (s2 + eps);
    const float m_ = tail_warp_sum(t) * rB;
    const float q = tail_warp_sum(on ? w * w * v + t * t : 0.f) * rB - m_ * m_;
    const float iv = rsqrtf(q + eps);
    if (on) z[c * B + lane] = g_[c] * (t - m_) * iv + b_[c];
    if (lane == 0) {
      inv[c] = iv;
      mu[c] = m_;
      bn_mu[(size_t)p * C + c] = m_;
      bn_var[(size_t)p * C + c] = q;
    }
  }
  cluster_sync_all();
  // ---- a1 = relu(W1 z): a warp per hidden unit, lanes = samples ----
  for (int h = cw; h < H; h += nw) {
    float acc = 0.f;
    // C dependent steps of one warp: two accumulators and 16 independent loads in flight (the phase is latency, not work)
    float acc2 = 0.f;
    int c = 0;
#pragma unroll 8
    for (; c + 1 < C; c += 2) {
      acc = __fmaf_rn(w1p[h * C + c], on ? z[c * B + lane] : 0.f, acc);
      acc2 = __fmaf_rn(w1p[h * C + c + 1], on ? z[(c + 1) * B + lane] : 0.f, acc2);
    }
    if (c < C) acc = __fmaf_rn(w1p[h * C + c], on ? z[c * B + lane] : 0.f, acc);
    acc += acc2;
    if (on) a1[h * B + lane] = fmaxf(acc, 0.f);
  }
  cluster_sync_all();
  // ---- gate and the folded scale / shift (padded channels: zeros) ----
  for (int i = ct; i < Cp * B; i += nt) {
    const int c = i / B, b = i - c * B;
    float s_ = 0.f, h_ = 0.f;
    if (c < C) {
      float acc = 0.f;
#pragma unroll 8
      for (int h = 0; h < H; h++) acc = __fmaf_rn(w2p[c * H + h], a1[h * B + b], acc);
      const float gt = 1.0f / (1.0f + expf(-acc));
      gate[c * B + b] = gt;
      const float s2 = vp[i];
      const float A = gt * g_[c] * inv[c] * wsty[b * C + c];
      s_ = A * rsqrtf(s2 + eps);
      h_ = gt * z[c * B + b] - s_ * mp[i];
    }
    sc[(size_t)p * Cp * B + i] = s_;
    sh[(size_t)p * Cp * B + i] = h_;
  }
}

__global__ void __launch_bounds__(TAIL_THREADS, 1) adain_tail_bwd_kernel(const float* __restrict__ gsc, const float* __restrict__ gsh,
                                                                          const float* __restrict__ mean, const float* __restrict__ var,
                                                                          const float* __restrict__ wsty, const float* __restrict__ bsty,
                                                                          const float* __restrict__ gam, const float* __restrict__ w1,
                                                                          const float* __restrict__ w2, int C, int Cp, int B, int H, float eps,
                                                                          const float* save_all, float* scratch_all, float* __restrict__ gmean,
                                                                          float* __restrict__ gvar, float* __restrict__ gws, float* __restrict__ gbs,
                                                                          float* __restrict__ ggam, float* __restrict__ gbet, float* __restrict__ gw1,
                                                                          float* __restrict__ gw2) {
  const int p = blockIdx.x / ADT_CLUSTER;
  const int nt = ADT_CLUSTER * blockDim.x, ct = (int)(cluster_ctarank() * blockDim.x + threadIdx.x);
  const int lane = ct & 31, cw = ct >> 5, nw = nt >> 5;
  const float* save = save_all + (size_t)p * adt_save_floats(C, B, H);
  const float* inv = save;
  const float* mu = inv + C;
  const float* z = mu + C;
  const float* gate = z + (size_t)C * B;
  const float* a1 = gate + (size_t)C * B;
  float* scratch = scratch_all + (size_t)p * adt_scratch_floats(C, B, H);
  float* ga2 = scratch;
  float* gz1 = ga2 + (size_t)C * B;
  float* gwt1 = gz1 + (size_t)C * B;
  float* gp1 = gwt1 + (size_t)C * B;
  float* ggam_part = gp1 + (size_t)H * B;
  float* ginv_part = ggam_part + C;
  const size_t pofs = (size_t)p * Cp * B;
  const float* __restrict__ g_ = gam + (size_t)p * C;
  const float* __restrict__ w1p = w1 + (size_t)p * H * C;
  const float* __restrict__ w2p = w2 + (size_t)p * C * H;
  const float rB = 1.0f / (float)B;
  const bool on = lane < B;
  // ---- through sc, sh to A, D, the instance statistics and the gate's pre-activation (a warp per channel, lanes = samples) ----
  for (int c = cw; c < Cp; c += nw) {
    const size_t i = pofs + (size_t)c * B + lane;
    if (c >= C) {  // padded channel: sc = sh = 0 whatever the statistics
      if (on) {
        gmean[i] = 0.f;
        gvar[i] = 0.f;
      }
      continue;
    }
    float t_gam = 0.f, t_inv = 0.f;
    if (on) {
      const float s2 = var[i], m_ = mean[i], r = rsqrtf(s2 + eps);
      const float w = wsty[lane * C + c], gt = gate[c * B + lane], zz = z[c * B + lane];
      const float A = gt * g_[c] * inv[c] * w;
      const float gD = gsh[i], gsct = gsc[i] - gD * m_;
      gmean[i] = -gD * (A * r);
      const float gA = gsct * r;
      gvar[i] = gsct * A * (-0.5f * r * r * r);                       // through r; the term through v is added in the last phase
      ga2[c * B + lane] = (gA * g_[c] * inv[c] * w + gD * zz) * gt * (1.0f - gt);
      gz1[c * B + lane] = gD * gt;
      gwt1[c * B + lane] = gA * gt * g_[c] * inv[c];
      t_gam = gA * gt * inv[c] * w;
      t_inv = gA * gt * g_[c] * w;
    }
    t_gam = tail_warp_sum(t_gam);
    t_inv = tail_warp_sum(t_inv);
    if (lane == 0) {
      ggam_part[c] = t_gam;
      ginv_part[c] = t_inv;
    }
  }
  cluster_sync_all();
  // ---- gp1 = (W2^T ga2) masked by the ReLU: a warp per hidden unit, lanes = samples ----
  for (int h = cw; h < H; h += nw) {
    float acc = 0.f;
    float acc2 = 0.f;
    int c = 0;
#pragma unroll 8
    for (; c + 1 < C; c += 2) {
      acc = __fmaf_rn(w2p[c * H + h], on ? ga2[c * B + lane] : 0.f, acc);
      acc2 = __fmaf_rn(w2p[(c + 1) * H + h], on ? ga2[(c + 1) * B + lane] : 0.f, acc2);
    }
    if (c < C) acc = __fmaf_rn(w2p[c * H + h], on ? ga2[c * B + lane] : 0.f, acc);
    acc += acc2;
    if (on) gp1[h * B + lane] = a1[h * B + lane] > 0.f ? acc : 0.f;
  }
  cluster_sync_all();
  // ---- per channel: the SE weight gradients, gz, and back through the squeeze and the BatchNorm statistics to the style ----
  // H <= 64: gp1 / a1 transposed into shared memory ([sample][hidden unit], padded), so the two weight gradients of a channel are
  // computed with lanes = hidden units and a loop over the samples -- no warp reduction per (channel, hidden unit) pair (two
  // dependent 5-step shuffle trees each: 0.42 ms for the decoders' first layer, C = 1026, H = 64).
  __shared__ float s_gp[32][65];
  __shared__ float s_a1[32][65];
  const bool staged = H <= 64;
  if (staged) {
    for (int e = threadIdx.x; e < H * B; e += blockDim.x) {
      const int h = e / B, b = e - h * B;
      s_gp[b][h] = gp1[e];
      s_a1[b][h] = a1[e];
    }
    __syncthreads();
  }
  for (int c = cw; c < C; c += nw) {
    const size_t i = pofs + (size_t)c * B + lane;
    const float zz = on ? z[c * B + lane] : 0.f, g2 = on ? ga2[c * B + lane] : 0.f;
    float gz = on ? gz1[c * B + lane] : 0.f;
    if (staged) {
      const int bl = on ? lane : 0;
      float gza = 0.f;
#pragma unroll 8
      for (int h = 0; h < H; h++) gza = __fmaf_rn(w1p[h * C + c], s_gp[bl][h], gza);
      if (on) gz += gza;
      for (int h0 = 0; h0 < H; h0 += 32) {
        const int h = h0 + lane;
        const int hl = h < H ? h : 0;
        float acc1 = 0.f, acc2 = 0.f;
        for (int b = 0; b < B; b++) {
          const float zb = __shfl_sync(0xffffffffu, zz, b), gb = __shfl_sync(0xffffffffu, g2, b);
          acc1 = __fmaf_rn(s_gp[b][hl], zb, acc1);
          acc2 = __fmaf_rn(gb, s_a1[b][hl], acc2);
        }
        if (h < H) {
          gw1[((size_t)p * H + h) * C + c] = acc1;
          gw2[((size_t)p * C + c) * H + h] = acc2;
        }
      }
    } else {
      for (int h = 0; h < H; h++) {
        const float gp = on ? gp1[h * B + lane] : 0.f, a_ = on ? a1[h * B + lane] : 0.f;
        gz = __fmaf_rn(w1p[h * C + c], gp, gz);
        const float s1 = tail_warp_sum(gp * zz), s2_ = tail_warp_sum(g2 * a_);
        if (lane == 0) {
          gw1[((size_t)p * H + h) * C + c] = s1;
          gw2[((size_t)p * C + c) * H + h] = s2_;
        }
      }
    }
    const float t = on ? bsty[lane * C + c] : 0.f, w = on ? wsty[lane * C + c] : 0.f;
    const float s2 = on ? var[i] : 1.f;
    const float v = s2 / (s2 + eps);
    const float iv = inv[c], m_ = mu[c], gm_ = g_[c];
    const float dgam = tail_warp_sum(gz * iv * (t - m_)) + ggam_part[c];
    const float dinv = tail_warp_sum(gz * gm_ * (t - m_)) + ginv_part[c];
    const float dbet = tail_warp_sum(gz);
    const float gq = dinv * (-0.5f * iv * iv * iv);
    const float dmu = -dbet * gm_ * iv - 2.0f * m_ * gq;
    if (lane == 0) {
      ggam[(size_t)p * C + c] = dgam;
      gbet[(size_t)p * C + c] = dbet;
    }
    if (on) {
      gbs[((size_t)p * B + lane) * C + c] = gz * gm_ * iv + (gq * 2.0f * t + dmu) * rB;
      gws[((size_t)p * B + lane) * C + c] = gwt1[c * B + lane] + gq * 2.0f * w * v * rB;
      const float d = s2 + eps;
      gvar[i] += gq * w * w * rB * (eps / (d * d));                   // v = s2 / (s2 + eps)
    }
  }
}

}  // namespace snb

SNB_API size_t snb_adain_tail_save_floats(int P, int C, int B, int H) {
  if (P <= 0 || C <= 0 || B <= 0 || H <= 0) return 0;
  return (size_t)P * adt_save_floats(C, B, H);
}

SNB_API size_t snb_adain_tail_scratch_floats(int P, int C, int B, int H) {
  if (P <= 0 || C <= 0 || B <= 0 || H <= 0) return 0;
  return (size_t)P * adt_scratch_floats(C, B, H);
}

static int adt_launch_cfg(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* at, int P, cudaStream_t s) {
  cfg = {};
  cfg.gridDim = dim3((unsigned)(P * ADT_CLUSTER));
  cfg.blockDim = dim3(TAIL_THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = ADT_CLUSTER;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return 0;
}

SNB_API int snb_adain_tail_fwd(const float* row_mean, const float* row_var, const float* style_scale, const float* style_shift, const float* gamma,
                               const float* beta, const float* w1, const float* w2, int P, int C, int Cpad, int B, int H, float eps, float* scale,
                               float* shift, float* save, float* bn_mean, float* bn_var, void* stream) {
  if (P < 0 || C < 0 || B < 0 || H < 0 || Cpad < C) return SNB_EINVAL;
  if (B > 32) return SNB_ELIMIT;
  if (P == 0 || C == 0 || B == 0) return SNB_OK;
  if (H == 0) return SNB_EINVAL;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute at[1];
  adt_launch_cfg(cfg, at, P, (cudaStream_t)stream);
  SNB_CUDA(cudaLaunchKernelEx(&cfg, adain_tail_fwd_kernel, row_mean, row_var, style_scale, style_shift, gamma, beta, w1, w2, C, Cpad, B, H, eps, scale,
                              shift, save, bn_mean, bn_var));
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

SNB_API int snb_adain_tail_bwd(const float* grad_scale, const float* grad_shift, const float* row_mean, const float* row_var,
                               const float* style_scale, const float* style_shift, const float* gamma, const float* w1, const float* w2, int P,
                               int C, int Cpad, int B, int H, float eps, const float* save, float* scratch, float* grad_row_mean,
                               float* grad_row_var, float* grad_style_scale, float* grad_style_shift, float* grad_gamma, float* grad_beta,
                               float* grad_w1, float* grad_w2, void* stream) {
  if (P < 0 || C < 0 || B < 0 || H < 0 || Cpad < C) return SNB_EINVAL;
  if (B > 32) return SNB_ELIMIT;
  if (P == 0 || C == 0 || B == 0) return SNB_OK;
  if (H == 0) return SNB_EINVAL;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute at[1];
  adt_launch_cfg(cfg, at, P, (cudaStream_t)stream);
  SNB_CUDA(cudaLaunchKernelEx(&cfg, adain_tail_bwd_kernel, grad_scale, grad_shift, row_mean, row_var, style_scale, style_shift, gamma, w1, w2, C, Cpad,
                              B, H, eps, save, scratch, grad_row_mean, grad_row_var, grad_style_scale, grad_style_shift, grad_gamma, grad_beta,
                              grad_w1, grad_w2));
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

// ---- the refiner's global feature: conv3 -> bn3 -> max over the points, from row statistics and extrema only -------------------------
// (reference models/sparenet_generator.py:626-629: x = bn3(conv3(x)); x, _ = torch.max(x, 2).)  With h = W3 x kept out of HBM the caller
// holds, per (sample, channel): the row mean m and biased row variance v of h (without the conv bias), and the row max / min.  BatchNorm
// is monotone per channel, so max_n BN(h + bias) = BN(h* + bias) with h* = max h where gamma > 0, min h otherwise:
//     mu = avg_b (m + bias),  q = avg_b v + avg_b (m + bias - mu)^2                      (batch statistics: within + between rows)
//     glob = (h* + bias - mu) gamma rsqrt(q + eps) + beta
// A thread per channel loops over the samples (coalesced across the block); forward and backward are one launch each instead of ~25
// and ~45 PyTorch launches on [B, 1024] tensors.  The conv bias cancels in train mode (its gradient is exactly 0).
namespace snb {

__global__ void __launch_bounds__(128) bn_max_tail_fwd_kernel(const float* __restrict__ m_bc, const float* __restrict__ v_bc,
                                                               const float* __restrict__ hmax, const float* __restrict__ hmin,
                                                               const float* __restrict__ bias, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, int B, int C, float eps, int training, float momentum,
                                                               float unbias, float* running_mean, float* running_var,
                                                               long long* num_batches_tracked, float* __restrict__ glob, float* __restrict__ save) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && training && num_batches_tracked) *num_batches_tracked += 1;
  if (c >= C) return;
  const float bi = bias ? bias[c] : 0.f;
  float mu, q;
  if (training) {
    float s = 0.f, sv = 0.f;
    for (int b = 0; b < B; b++) {
      s += m_bc[(size_t)b * C + c] + bi;
      sv += v_bc[(size_t)b * C + c];
    }
    mu = s / (float)B;
    float d2 = 0.f;
    for (int b = 0; b < B; b++) {
      const float d = m_bc[(size_t)b * C + c] + bi - mu;
      d2 = __fmaf_rn(d, d, d2);
    }
    q = (sv + d2) / (float)B;
    if (running_mean) {
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mu;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * q * unbias;
    }
  } else {
    mu = running_mean[c];
    q = running_var[c];
  }
  const float inv = rsqrtf(q + eps);
  const float g = gamma[c], sc = g * inv, be = beta[c];
  const float* __restrict__ hs = g > 0.f ? hmax : hmin;
  for (int b = 0; b < B; b++) glob[(size_t)b * C + c] = (hs[(size_t)b * C + c] + bi - mu) * sc + be;
  save[c] = mu;
  save[C + c] = inv;
}

__global__ void __launch_bounds__(128) bn_max_tail_bwd_kernel(const float* __restrict__ gglob, const float* __restrict__ m_bc,
                                                               const float* __restrict__ hmax, const float* __restrict__ hmin,
                                                               const float* __restrict__ bias, const float* __restrict__ gamma,
                                                               const float* __restrict__ save, int B, int C, int training, float* __restrict__ gm_bc,
                                                               float* __restrict__ gv_bc, float* __restrict__ ghmax, float* __restrict__ ghmin,
                                                               float* __restrict__ ggamma, float* __restrict__ gbeta, float* __restrict__ gbias) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float bi = bias ? bias[c] : 0.f;
  const float mu = save[c], inv = save[C + c], g = gamma[c], sc = g * inv;
  const bool up = g > 0.f;
  const float* __restrict__ hs = up ? hmax : hmin;
  float sg = 0.f, sgx = 0.f;
  for (int b = 0; b < B; b++) {
    const float gg = gglob[(size_t)b * C + c];
    sg += gg;
    sgx = __fmaf_rn(gg, hs[(size_t)b * C + c] + bi - mu, sgx);
  }
  gbeta[c] = sg;
  ggamma[c] = sgx * inv;
  // d glob / d mu = -sc;  d glob / d inv = (h* + bias - mu) gamma  ->  d / d q = . (-1/2) inv^3
  const float gmu = -sg * sc;
  const float gq = sgx * g * (-0.5f * inv * inv * inv);
  const float rB = 1.f / (float)B;
  for (int b = 0; b < B; b++) {
    const size_t i = (size_t)b * C + c;
    const float gh = gglob[i] * sc;
    ghmax[i] = up ? gh : 0.f;
    ghmin[i] = up ? 0.f : gh;
    if (training) {
      gm_bc[i] = (gmu + 2.f * gq * (m_bc[i] + bi - mu)) * rB;
      gv_bc[i] = gq * rB;
    } else {
      gm_bc[i] = 0.f;
      gv_bc[i] = 0.f;
    }
  }
  if (gbias) gbias[c] = training ? 0.f : sg * sc;   // train mode: the bias shifts h* and mu alike
}

}  // namespace snb

SNB_API int snb_bn_max_tail_fwd(const float* row_mean, const float* row_var, const float* row_max, const float* row_min, const float* conv_bias,
                                const float* gamma, const float* beta, int B, int C, float eps, int training, float momentum, float unbias,
                                float* running_mean, float* running_var, long long* num_batches_tracked, float* glob, float* save, void* stream) {
  if (B < 0 || C < 0) return SNB_EINVAL;
  if (B == 0 || C == 0) return SNB_OK;
  if (!training && (!running_mean || !running_var)) return SNB_EINVAL;
  snb::bn_max_tail_fwd_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(row_mean, row_var, row_max, row_min, conv_bias, gamma, beta, B, C, eps,
                                                                              training, momentum, unbias, running_mean, running_var,
                                                                              num_batches_tracked, glob, save);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

SNB_API int snb_bn_max_tail_bwd(const float* grad_glob, const float* row_mean, const float* row_max, const float* row_min, const float* conv_bias,
                                const float* gamma, const float* save, int B, int C, int training, float* grad_row_mean, float* grad_row_var,
                                float* grad_row_max, float* grad_row_min, float* grad_gamma, float* grad_beta, float* grad_conv_bias, void* stream) {
  if (B < 0 || C < 0) return SNB_EINVAL;
  if (B == 0 || C == 0) return SNB_OK;
  snb::bn_max_tail_bwd_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(grad_glob, row_mean, row_max, row_min, conv_bias, gamma, save, B, C,
                                                                              training, grad_row_mean, grad_row_var, grad_row_max, grad_row_min,
                                                                              grad_gamma, grad_beta, grad_conv_bias);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
