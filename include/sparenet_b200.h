/*
 * sparenet_b200.h -- C ABI of libsparenet_b200.so: the SpareNet per-batch point-cloud hot path on B200 (sm_100a).
 *
 * Drop-in boundary: each entry point replaces one pybind function of the reference's CUDA
 * extensions; the reference-side Python wrappers (torch.autograd.Function / nn.Module, mirrored under
 * sparenet_b200/dropin/) bind these through ctypes.
 *
 * Conventions
 *   - plain device pointers + sizes, no torch types.  fp32 contiguous AoS clouds [B,N,3]; int32 indices.
 *   - the caller owns every buffer, including scratch ("workspace"); the library never allocates, frees or
 *     keeps state between calls.  Query scratch sizes with snb_<op>_workspace_bytes().
 *   - stream-explicit: `stream` is a cudaStream_t (0 = legacy default stream); the device is the current one.
 *     No host synchronisation, no host reads of device memory: every call is CUDA-graph capturable.
 *   - return value: 0 = ok, < 0 = invalid argument (SNB_E*), > 0 = cudaError_t of the failed launch.
 *     snb_strerror() turns any of them into text.  (The reference printf()s and continues, chamfer.cu:166-169,
 *     emd_cuda.cu:276-280, or exit(-1)s, MDS_cuda.cu:15-24.)
 */
#ifndef SPARENET_B200_H
#define SPARENET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNB_OK 0
#define SNB_EINVAL (-1)     /* null/negative/inconsistent argument */
#define SNB_ELIMIT (-2)     /* shape outside the op's documented limits */
#define SNB_EWORKSPACE (-3) /* workspace too small */
#define SNB_EALIGN (-4)     /* pointer alignment */

int snb_version(void);
const char* snb_strerror(int code);

/* ---- Chamfer --------------------------------------------------------------------------------------
 * replaces chamfer.forward / chamfer.backward (cuda/chamfer_dist/chamfer_cuda.cpp:12-42, chamfer.cu:147-229)
 * and cd.forward_cuda / cd.backward_cuda (cuda/chamfer_distance/chamfer_distance.cpp:26-55,
 * chamfer_distance.cu:139-155,189-209).
 * dist1[b,i] = min_j |xyz1[b,i]-xyz2[b,j]|^2, idx1 = smallest argmin; dist2/idx2 the other direction.
 * Outputs are fully written (no pre-zeroing needed).  Gradient buffers are fully written too. */
size_t snb_chamfer_workspace_bytes(int B, int N, int M);  /* 0 when the shape is served by the brute-force kernel only */
int snb_chamfer_fwd(const float* xyz1, const float* xyz2, int B, int N, int M,
                    float* dist1, float* dist2, int* idx1, int* idx2,
                    void* workspace, size_t workspace_bytes, void* stream);  /* workspace may be NULL: brute force */
int snb_chamfer_bwd(const float* xyz1, const float* xyz2, int B, int N, int M,
                    const int* idx1, const int* idx2, const float* grad_dist1, const float* grad_dist2,
                    float* grad_xyz1, float* grad_xyz2, void* stream);

/* ---- EMD (auction) ---------------------------------------------------------------------------------
 * replaces emd.forward / emd.backward (cuda/emd/emd.cpp:6-28, emd_cuda.cu:228-316).  n == m, n % 1024 == 0,
 * B <= 512 (emd_cuda.cu:236-249 -> SNB_ELIMIT).  The 12 scratch tensors the reference allocates in Python
 * (emd_module.py:43-54) become one caller-owned workspace.  dist = squared distance to the assigned point.
 * snb_emd_fwd: N <= 16384 runs the auction over a Morton-ordered box hierarchy of both clouds (a Bid pass only opens boxes whose
 * bound 3 - dist - min price can still change a bidder's best / second best); larger clouds run the exhaustive Bid.
 * snb_emd_fwd_scan: always the exhaustive Bid.  Both return the same assignment and distances, bit for bit. */
size_t snb_emd_workspace_bytes(int B, int N);
int snb_emd_fwd(const float* xyz1, const float* xyz2, int B, int N, float eps, int iters,
                float* dist, int* assignment, void* workspace, size_t workspace_bytes, void* stream);
int snb_emd_fwd_scan(const float* xyz1, const float* xyz2, int B, int N, float eps, int iters,
                     float* dist, int* assignment, void* workspace, size_t workspace_bytes, void* stream);
int snb_emd_bwd(const float* xyz1, const float* xyz2, int B, int N, const float* grad_dist,
                const int* assignment, float* grad_xyz1, void* stream);

/* ---- Expansion penalty (per-primitive MST) ------------------------------------------------------------
 * replaces expansion_penalty.forward / .backward (cuda/expansion_penalty/expansion_penalty.cpp:4-22,
 * expansion_penalty_cuda.cu:151-198).  primitive_size: power of two, 2..512, dividing N (the reference's tree
 * reductions are only complete for powers of two).  The 2 x B*N*512 global scratch
 * (expansion_penalty_module.py:33-34) shrinks to B*N/p floats.  mean_mst_length is already divided by N/p. */
size_t snb_expansion_workspace_bytes(int B, int N, int primitive_size);
int snb_expansion_fwd(const float* xyz, int B, int N, int primitive_size, float alpha,
                      float* dist, int* assignment, float* mean_mst_length,
                      void* workspace, size_t workspace_bytes, void* stream);
int snb_expansion_bwd(const float* xyz, int B, int N, const float* grad_dist, const int* assignment,
                      float* grad_xyz, void* stream);

/* ---- Minimum-density sampling + gather ---------------------------------------------------------------
 * replaces MDS.minimum_density_sampling / gather_points / gather_points_grad (cuda/MDS/MDS.cpp:54-135,
 * MDS_cuda.cu:29-271).  xyz [B,n,3], idx [B,m] int32; features [B,C,n]. */
size_t snb_mds_workspace_bytes(int B, int n, int m);
int snb_mds_sample(const float* xyz, int B, int n, int m, const float* mean_mst_length, int* idx,
                   void* workspace, size_t workspace_bytes, void* stream);
int snb_gather_fwd(const float* features, const int* idx, int B, int C, int n, int m, float* out, void* stream);
int snb_gather_bwd(const float* grad_out, const int* idx, int B, int C, int n, int m, float* grad_features, void* stream);

/* ---- p2i (point -> image splat) ----------------------------------------------------------------------
 * replaces ext.p2i_{max,sum}_{forward,backward}_gpu (cuda/p2i_op/ext.cpp:8-15, p2i_max.h:145-232,
 * p2i_sum.h:133-214).  points [npoints,2] (row, col) in PIXEL space, features [npoints,C], batch_inds [npoints],
 * background/out [B,C,H,W]; kernel_kind 0 = cosine.  is_double selects float64 buffers (gradcheck path).
 * max: out = max(background, max f*w); ids = lowest point id attaining it, -1 where the background stands.
 * workspace: snb_p2i_workspace_bytes (zero-size allowed for sum). */
size_t snb_p2i_workspace_bytes(int B, int C, int H, int W, int is_double);
int snb_p2i_max_fwd(const void* points, const void* features, const int* batch_inds, const void* background,
                    int npoints, int B, int C, int H, int W, int kernel_kind, double radius, int is_double,
                    void* out, int* out_ids, void* workspace, size_t workspace_bytes, void* stream);
int snb_p2i_max_bwd(const void* grad_out, const int* out_ids, const void* points, const void* features,
                    int npoints, int B, int C, int H, int W, int kernel_kind, double radius, int is_double,
                    void* grad_points, void* grad_features, void* grad_background, void* stream);
int snb_p2i_sum_fwd(const void* points, const void* features, const int* batch_inds, const void* background,
                    int npoints, int B, int C, int H, int W, int kernel_kind, double radius, int is_double,
                    void* out, void* stream);
int snb_p2i_sum_bwd(const void* grad_out, const void* points, const void* features, const int* batch_inds,
                    int npoints, int B, int C, int H, int W, int kernel_kind, double radius, int is_double,
                    void* grad_points, void* grad_features, void* stream);

/* ---- fused depth-map renderer (csrc/depthmaps.cu) -----------------------------------------------------
 * replaces ComputeDepthMaps.forward (utils/p2i_utils.py:211-252) + p2i(reduce="max") (cuda/p2i_op/__init__.py:99-131) for ONE
 * view and ONE radius: q = M [p;1], pos = q.xyz/q.w, (row, col) = ((-pos.y, pos.x) + 1)/2 * (H-1, W-1),
 * feature = 1 - (pos.z - zmin)/(zmax - zmin) with zmin/zmax over all B*N points of the call, out = max(0, max feature * w(r)).
 * data [B,N,3]; view_matrix: 16 floats on the HOST (row-major projection . look_at, utils/p2i_utils.py:200-209);
 * out [B,1,H,W], ids [B,1,H,W] (winner point index in [0, B*N) or -1).  The workspace must be kept by the caller from _fwd to
 * _bwd (it holds the per-point pixel coordinates and the depth range).  _bwd writes grad_data [B,N,3] completely, including the
 * gradient that reaches the two extreme points through zmin / zmax. */
size_t snb_depthmaps_workspace_bytes(int B, int N, int H, int W);
int snb_depthmaps_fwd(const float* data, int B, int N, const float* view_matrix, int H, int W, double radius,
                      float* out, int* ids, void* workspace, size_t workspace_bytes, void* stream);
int snb_depthmaps_bwd(const float* grad_out, const int* ids, const float* data, int B, int N, const float* view_matrix,
                      int H, int W, double radius, void* workspace, size_t workspace_bytes, float* grad_data, void* stream);

/* ---- kNN (replaces the un-vendored knn_cuda.KNN used by models/sparenet_generator.py:852-877) -------
 * x [B,C,N] channel-major features; idx [B,N,k] int32 = the k nearest points (self included) by exact fp32
 * squared distance, ascending; ties by smaller index.  k <= 32. */
size_t snb_knn_workspace_bytes(int B, int N);
int snb_knn(const float* x, int B, int C, int N, int k, int* idx, void* workspace, size_t workspace_bytes, void* stream);
/* The same result (bit-identical indices) for wide features, with a caller-supplied Gram matrix gram [B,N,N] = X^T X from a
 * TF32 library GEMM used as a PRUNING filter and exact fp32 re-evaluation of the surviving candidates (csrc/knn_prune.cu).
 * xT [B,N,C] is the point-major copy of the features.  Parity-tested on B200. */
size_t snb_knn_pruned_workspace_bytes(int B, int N);
/* xT [B,N,C] = point-major copy of x [B,C,N] (what snb_knn_pruned and the Gram GEMM read; replaces the tensor permute + copy the
 * reference's knn() does before calling knn_cuda, models/sparenet_generator.py:866-868). */
int snb_transpose_cn(const float* x, int B, int C, int N, float* xT, void* stream);
int snb_knn_pruned(const float* xT, const float* gram, int B, int C, int N, int k, int* idx, void* workspace,
                   size_t workspace_bytes, void* stream);

/* ---- EdgeConv neighbourhood reduction (models/sparenet_generator.py:188-242,880-906) -----------------------
 * With conv([x_j - x_i ; x_i]) = a_j + c_i (a = W_a x, c = (W_b - W_a) x, per-point GEMMs done by the caller) this
 * produces everything the block needs from u[b,ch,i,m] = a[b,ch,idx[b,i,m]] + c[b,ch,i] without forming it:
 * umax/umin [B,C,N] = max_m / min_m u, the winning neighbour slots (uint8), S1/S2 [B,C] = sum u, sum u^2 (fp64).
 * a, c, umax, umin are channel-major [B,C,N]; idx [B,N,k] int32, k <= 32.  _bwd is the exact adjoint. */
int snb_edge_reduce_fwd(const float* a, const float* c, const int* idx, int B, int C, int N, int k,
                        float* umax, float* umin, unsigned char* slot_max, unsigned char* slot_min,
                        double* S1, double* S2, void* stream);
int snb_edge_reduce_bwd(const float* a, const float* c, const int* idx, const unsigned char* slot_max,
                        const unsigned char* slot_min, const float* g_umax, const float* g_umin,
                        const double* gS1, const double* gS2, int B, int C, int N, int k,
                        float* ga, float* gc, void* stream);
/* The same with the extremum chosen per channel (sel_max[ch] != 0: max, else min -- the sign of the channel's BatchNorm
 * weight, models/sparenet_generator.py:213-216): ustar [B,C,N] and ONE slot tensor instead of two of each. */
int snb_edge_reduce_sel_fwd(const float* a, const float* c, const int* idx, const unsigned char* sel_max, int B, int C,
                            int N, int k, float* ustar, unsigned char* slot, double* S1, double* S2, void* stream);
int snb_edge_reduce_sel_bwd(const float* a, const float* c, const int* idx, const unsigned char* slot,
                            const float* g_ustar, const double* gS1, const double* gS2, int B, int C, int N, int k,
                            float* ga, float* gc, void* stream);
/* The selected-extremum pair for a and c stored as the two channel halves of ONE [B,2C,N] tensor (the output of a single GEMM with the
 * stacked weight [W_a ; W_b - W_a]); the backward fills the two halves of ONE [B,2C,N] gradient. */
int snb_edge_reduce_sel_fwd_stacked(const float* ac, const int* idx, const unsigned char* sel_max, int B, int C, int N, int k,
                                    float* ustar, unsigned char* slot, double* S1, double* S2, void* stream);
int snb_edge_reduce_sel_bwd_stacked(const float* ac, const int* idx, const unsigned char* slot, const float* g_ustar, const double* gS1,
                                    const double* gS2, int B, int C, int N, int k, float* gac, void* stream);

/* ---- row-wise tails of the folded normalisation stacks (models/sparenet_generator.py:618-646,1053-1061) ----
 * h [R,L] contiguous rows.  row_stats: mean / biased variance per row.  row_affine_act:
 * y[r,:] = leaky_relu(h[r / in_div,:] * scale[r] + shift[r], slope) (slope 0 = ReLU); _bwd returns gh [R/in_div, L],
 * gscale [R], gshift [R].  row_minmax: per-row max/min and their first positions. */
int snb_row_stats(const float* h, long long R, int L, float* mean, float* var, void* stream);
int snb_row_stats_bwd(const float* h, const float* mean, const float* gmean, const float* gvar, long long R, int L,
                      float* gh, void* stream);
int snb_row_affine_act_fwd(const float* h, const float* scale, const float* shift, long long R, int L, int in_div,
                           float slope, float* y, void* stream);
int snb_row_affine_act_bwd(const float* gy, const float* h, const float* scale, const float* shift, long long R, int L,
                           int in_div, float slope, float* gh, float* gscale, float* gshift, void* stream);
int snb_row_minmax(const float* h, long long R, int L, float* vmax, float* vmin, int* imax, int* imin, void* stream);
/* row_stats + row_minmax in one pass (PointNetRes conv3 -> bn3 -> max over points, models/sparenet_generator.py:626-629). */
int snb_row_stats_minmax(const float* h, long long R, int L, float* mean, float* var, float* vmax, float* vmin,
                         int* imax, int* imin, void* stream);
/* Two-phase backward of y = leaky_relu(scale*h + shift) when (scale, shift) depend on h's row statistics (BN o SE o ReLU):
 * _reduce returns gscale[r] = sum_l d*h, gshift[r] = sum_l d with d = gy * act'(scale*h + shift); after the caller has
 * pulled (gscale, gshift) back to (gmean, gvar), row_norm_act_bwd writes gh = d*scale + gmean/L + 2 gvar (h - mean)/L.
 * gy_row (may be NULL): a per-row constant added to gy on the fly (the W^T b 1^T term of a row-statistics gradient). */
int snb_row_act_bwd_reduce(const float* gy, const float* h, const float* scale, const float* shift, long long R, int L,
                           float slope, float* gscale, float* gshift, const float* gy_row, void* stream);
int snb_row_norm_act_bwd(const float* gy, const float* h, const float* scale, const float* shift, const float* mean,
                         const float* gmean, const float* gvar, long long R, int L, float slope, float* gh,
                         const float* gy_row, void* stream);

/* Pooled tail: vmax / vmean [R] = max / mean over the row of leaky_relu(scale*h + shift) (first position imax on ties) without
 * storing the activated row -- the encoder's [max | mean over points] output (models/sparenet_generator.py:234-242).  Backward
 * in the same two phases as above: _bwd_reduce returns gscale / gshift for gy = gmean/L + [l == imax] gmax, _bwd writes
 * gh = d*scale + gstat_mean/L + 2 gstat_var (h - mean)/L. */
int snb_row_act_pool_fwd(const float* h, const float* scale, const float* shift, long long R, int L, float slope,
                         float* vmax, int* imax, float* vmean, void* stream);
int snb_row_act_pool_bwd_reduce(const float* h, const float* scale, const float* shift, const float* gmax, const float* gmean,
                                const int* imax, long long R, int L, float slope, float* gscale, float* gshift, void* stream);
int snb_row_act_pool_bwd(const float* h, const float* scale, const float* shift, const float* mean, const float* gmax,
                         const float* gmean, const int* imax, const float* gstat_mean, const float* gstat_var, long long R, int L,
                         float slope, float* gh, void* stream);

/* ---- closed-form BatchNorm . SE tail of a dense layer (csrc/tails.cu; models/sparenet_generator.py:593-646, 767-790) -------
 * (scale, shift) [B,C] of tail(h) = relu(scale*h + shift) from the ROW statistics of the pre-activation h [B,C,L]:
 * m = row_mean + row_bias ([C], or [B,C] with bias_per_sample), BatchNorm1d over (B, L) from m and row_var (train: batch
 * statistics, running_* advanced in place with `momentum` and the unbiased factor `unbias` = B*L/(B*L-1), the int64 counter
 * incremented; eval: running_*), SELayer1D gate = sigmoid(w2 relu(w1 z)) on the squeeze z = BN(m), w1 [H,C], w2 [C,H];
 * scale = gate*sc, shift = gate*sh + row_bias*scale.  One thread block per call (the data is B*C values).  `save` (fwd -> bwd)
 * holds snb_bn_se_tail_save_floats floats, `scratch` snb_bn_se_tail_scratch_floats.  bwd: gradients of every input from
 * (grad_scale, grad_shift); grad_row_bias is [C] or [B,C] like row_bias and may be NULL. */
size_t snb_bn_se_tail_save_floats(int B, int C, int H);
size_t snb_bn_se_tail_scratch_floats(int B, int C, int H);
int snb_bn_se_tail_fwd(const float* row_mean, const float* row_var, const float* row_bias, int bias_per_sample,
                       const float* gamma, const float* beta, const float* w1, const float* w2, int B, int C, int H,
                       float eps, int training, float momentum, float unbias, float* running_mean, float* running_var,
                       long long* num_batches_tracked, float* scale, float* shift, float* save, void* stream);
int snb_bn_se_tail_bwd(const float* grad_scale, const float* grad_shift, const float* row_mean, const float* row_bias,
                       int bias_per_sample, const float* gamma, const float* w1, const float* w2, int B, int C, int H,
                       int training, const float* save, float* scratch, float* grad_row_mean, float* grad_row_var,
                       float* grad_row_bias, float* grad_gamma, float* grad_beta, float* grad_w1, float* grad_w2, void* stream);

/* ---- the decoders' tail: instance norm -> AdaIN -> BatchNorm1d -> SELayer1D -> ReLU folded to one scale/shift (csrc/tails.cu;
 * models/sparenet_generator.py:909-1062) for all P primitives at once.  row_mean / row_var [P,Cpad,B]: statistics of the rows
 * h[p,c,b,:] of the layer's pre-activation (the GEMM epilogue's), style_scale / style_shift [B,C]: the AdaIN weight / bias of the
 * sample (shared by the primitives), gamma / beta [P,C], w1 [P,H,C], w2 [P,C,H]; scale / shift [P,Cpad,B] (channels >= C: zeros).
 * Train-mode BatchNorm (batch statistics; they are returned in bn_mean / bn_var [P,C] for the caller's running statistics).  B <= 32.
 * bwd: gradients of every input; grad_style_* are per-primitive partials [P,B,C] (the caller sums over P). */
/* The refiner's global feature, conv3 -> bn3 -> max over the points (models/sparenet_generator.py:626-629), from the row statistics and
 * extrema of h = W3 x alone: glob [B,C] = BN(h* + bias) with h* = row max where gamma > 0, row min otherwise; train mode uses (and
 * advances the running statistics with) the batch statistics rebuilt from the rows.  save: 2*C floats for the backward. */
int snb_bn_max_tail_fwd(const float* row_mean, const float* row_var, const float* row_max, const float* row_min, const float* conv_bias,
                        const float* gamma, const float* beta, int B, int C, float eps, int training, float momentum, float unbias,
                        float* running_mean, float* running_var, long long* num_batches_tracked, float* glob, float* save, void* stream);
int snb_bn_max_tail_bwd(const float* grad_glob, const float* row_mean, const float* row_max, const float* row_min, const float* conv_bias,
                        const float* gamma, const float* save, int B, int C, int training, float* grad_row_mean, float* grad_row_var,
                        float* grad_row_max, float* grad_row_min, float* grad_gamma, float* grad_beta, float* grad_conv_bias, void* stream);
size_t snb_adain_tail_save_floats(int P, int C, int B, int H);
size_t snb_adain_tail_scratch_floats(int P, int C, int B, int H);
int snb_adain_tail_fwd(const float* row_mean, const float* row_var, const float* style_scale, const float* style_shift,
                       const float* gamma, const float* beta, const float* w1, const float* w2, int P, int C, int Cpad, int B,
                       int H, float eps, float* scale, float* shift, float* save, float* bn_mean, float* bn_var, void* stream);
int snb_adain_tail_bwd(const float* grad_scale, const float* grad_shift, const float* row_mean, const float* row_var,
                       const float* style_scale, const float* style_shift, const float* gamma, const float* w1, const float* w2,
                       int P, int C, int Cpad, int B, int H, float eps, const float* save, float* scratch, float* grad_row_mean,
                       float* grad_row_var, float* grad_style_scale, float* grad_style_shift, float* grad_gamma, float* grad_beta,
                       float* grad_w1, float* grad_w2, void* stream);

/* ---- Adam over a flat parameter arena (csrc/optim.cu) -------------------------------------------------------------------------
 * One launch for all parameters: param / grad / exp_avg / exp_avg_sq are flat fp32 arrays of n elements (n % 4 == 0, 16-byte
 * aligned) with the SAME layout (sparenet_b200.dist.GradArena / sparenet_b200.optim.FlatAdam).  torch.optim.Adam's update rule
 * (no amsgrad, no maximize; weight_decay is the L2 form), `step` = 1, 2, ... is the number of this update. */
int snb_adam_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int step, void* stream);
/* ---- 1x1 convolutions with a thin side (<= 8 channels in or out; N % 4 == 0), exact fp32 streaming kernels (csrc/thinconv.cu): the
 * xyz / lattice input layers and xyz output layers of the generator (models/sparenet_generator.py:146-160, 593-646, 984-991, 1044-1062).
 * Every operand has its own batch stride in floats (0 = shared by the batch); W is addressed as W[g*w_bs + row*w_rs + col*w_cs], so a data
 * gradient uses the transposed weight without a copy.
 *   expand: y [G,Co,N] = W x, x [.,S,N], W rows = output channels, cols = the S thin input channels
 *   reduce: y [.,S,N]  = W x, x [G,L,N], W rows = the S thin output channels, cols = input channels
 *   wgrad:  out [G,L,8] (columns >= S zero), out[g][l][s] = sum_n big[g][l,n] small[g][s,n] */
int snb_thin_expand(const float* x, long long x_bs, const float* W, long long w_bs, int w_rs, int w_cs, int G, int S, int Co, int N, float* y,
                    void* stream);
int snb_thin_reduce(const float* x, long long x_bs, const float* W, long long w_bs, int w_rs, int w_cs, int G, int S, int L, int N, float* y,
                    long long y_bs, void* stream);
int snb_thin_wgrad(const float* big, long long big_bs, const float* small_, long long small_bs, int G, int S, int L, int N, float* out,
                   void* stream);

/* ---- fp32 nn.Linear for small batches (B <= 32 rows, K % 4 == 0): the encoder -> decoder bridge (models/sparenet_generator.py:85-120
 * SpareNetEncode.linear, :289-330 SpareNetDecode.mlp).  Exact fp32 FMAs like the reference's cuBLAS calls (no tensor cores); x [B,K],
 * W [O,K] row-major, y [B,O].  workspace: snb_linear_workspace_floats(B,K,O) floats (split-slice partial sums, added in a fixed order).
 * wgrad: gW [O,K] = gy^T x, gbias [O] (may be NULL). */
size_t snb_linear_workspace_floats(int B, int K, int O);
int snb_linear_fwd(const float* x, const float* W, const float* bias, int B, int K, int O, float* y, float* workspace, void* stream);
int snb_linear_dgrad(const float* gy, const float* W, int B, int K, int O, float* gx, float* workspace, void* stream);
int snb_linear_wgrad(const float* gy, const float* x, int B, int K, int O, float* gW, float* gbias, void* stream);

/* Gather of many contiguous float32 tensors into their slots of a flat arena in ONE launch per 1024 tensors
 * (sparenet_b200.dist.GradArena.pack): srcs / dsts / ns are HOST arrays of ntab device pointers and element counts; the pointer
 * table travels as a kernel parameter, so the call is graph-capturable without any host-to-device copy. */
int snb_multi_copy(const void* const* srcs, void* const* dsts, const long long* ns, int ntab, void* stream);

/* ---- TF32 tensor-core GEMM of the 1x1-conv / AdaIN-folding stacks (csrc/gemm_tc.cu: tcgen05.mma + TMEM + TMA) ----------
 * replaces the cuDNN/cuBLAS calls behind nn.Conv1d / nn.Conv2d(kernel_size=1) in models/sparenet_generator.py:146-186,
 * 188-242 (EdgeConv), :593-646 (PointNetRes), :984-991,1044-1062 (GridDecoder) on channel-major activations
 * [G, C, Npos] (positions contiguous), TF32 operands with fp32 accumulation like the reference's cuDNN default.
 *   mode 0 FWD    D[g] (M x N) = A (M x K) . T(B[g]) (K x N)      A = weight [M=Cout, K=Cin] (lda), B = X[g] [K rows, N]
 *   mode 1 DGRAD  D[g] (M x N) = A^T . B[g]                        A = weight stored [K=Cout rows, M=Cin] (lda), B = gY[g]
 *   mode 2 WGRAD  D[g] (M x N) = sum_{bi<BI} A[g*BI+bi] . T(B[g*BI+bi])^T   A = gY [M=Cout rows, K positions],
 *                                                                    B = X [N=Cin rows, K positions]
 * T(x) = leaky_relu(scale*x + shift, slope) per (batch, input channel, segment) when scale != NULL (FWD, WGRAD): the
 * folded AdaIN.BN.SE.ReLU tail of the previous layer applied while the operand sits in shared memory;
 * scale/shift are [batches, Cin, Npos/seg], seg a multiple of block_n (FWD) / 32 (WGRAD).
 * Epilogue (from TMEM): store (1) / TMA reduce-add into a caller-initialised D (2) / nothing (0), and optionally
 * per-tile row statistics pmean/pm2 [G, M, N/block_n] (mean and CENTRED second moment of each tile row) and
 * pmax/pmin/pimax/pimin [G, M, N/block_n] (extrema and their column in the full row).  Limits: leading dimensions and
 * batch strides multiples of 4 elements, pointers 16-byte aligned, N % 32 == 0 (FWD/DGRAD), M % 32 == 0 (DGRAD),
 * block_n in {32,64,...,256} (0 = auto), statistics need N % block_n == 0. */
typedef struct snb_gemm_desc {
  int mode;
  int G, BI;
  int M, N, K;
  const float* A; long long lda; long long a_batch_stride;   /* a_batch_stride == 0: one A for all batches */
  const float* B; long long ldb; long long b_batch_stride;
  float* D; long long ldd; long long d_batch_stride;
  int block_n;
  int store;
  int split;                                                  /* WGRAD split-K (0 = auto, needs store == 2 when > 1) */
  const float* scale; const float* shift; float slope; int seg;
  float* pmean; float* pm2;
  float* pmax; float* pmin; int* pimax; int* pimin;
  int b_pos_mod;                                              /* > 0 (FWD, WGRAD): B holds only b_pos_mod positions per batch entry
                                                                 and is tiled along the position axis (the scale/shift still vary) */
} snb_gemm_desc;
int snb_gemm_tf32(const snb_gemm_desc* desc, void* stream);
int snb_gemm_tf32_block_n(int N, int block_n);                /* the column-tile width the library uses (block_n = 0: its choice) */
int snb_gemm_tf32_tiles(int N, int block_n);                  /* statistics tiles along N: 2 per column tile (width block_n/2) */
/* Adjoint of the row extrema of h = W x without h (PointNetRes conv3 -> max over the points, models/sparenet_generator.py:626-629):
 * for every (sample, output channel) with a non-zero gradient on its row max / min, adds g W[co,:] to column imax / imin of gx [B,Ci,N]
 * and g x[b,:,col] to row co of gW [Co,Ci] (atomic accumulation; gx, gW, gmax, gmin may each be NULL). */
int snb_conv_extrema_bwd(const float* x, const float* W, const int* imax, const int* imin, const float* gmax, const float* gmin, int B, int Ci,
                         int Co, int N, float* gx, float* gW, void* stream);
/* Merge of the epilogue's per-tile statistics (csrc/rowops.cu), one launch each: tile means / centred second moments
 * [pairs, tiles_per_segment] of tiles of w positions -> mean and biased variance [pairs] of every segment (pairs = rows x segments);
 * tile extrema with their positions [rows, T] -> the rows' extrema (first position attaining them). */
int snb_gemm_stats_merge(const float* pmean, const float* pm2, long long pairs, int tiles_per_segment, int w, float* mean, float* var,
                         void* stream);
int snb_gemm_minmax_merge(const float* pmax, const float* pmin, const int* pimax, const int* pimin, long long rows, int T, float* vmax,
                          float* vmin, int* imax, int* imin, void* stream);

/* ---- Gridding / GriddingReverse (GRNet) --------------------------------------------------------------------
 * replaces gridding.forward / backward / rev_forward / rev_backward (cuda/gridding/gridding_cuda.cpp:43-99,
 * gridding.cu:179-335, gridding_reverse.cu:105-236).  ptcloud [B,n,3] already multiplied by scale/2; grid
 * [B, len_x*len_y*len_z] with len = max - min + 1; grid_pt_weights [B,n,8,3], grid_pt_indexes [B,n,8] (-1 where a
 * corner lies outside the grid: the reference writes out of bounds there).  Reverse: grid [B,S^3] -> ptcloud [B,S^3,3]. */
int snb_gridding_fwd(const float* ptcloud, int B, int n, float min_x, float max_x, float min_y, float max_y,
                     float min_z, float max_z, float* grid, float* grid_pt_weights, int* grid_pt_indexes, void* stream);
int snb_gridding_bwd(const float* grid_pt_weights, const int* grid_pt_indexes, const float* grad_grid, int B, int n,
                     long long n_grid_vertices, float* grad_ptcloud, void* stream);
int snb_gridding_rev_fwd(const float* grid, int B, int scale, float* ptcloud, void* stream);
int snb_gridding_rev_bwd(const float* ptcloud, const float* grid, const float* grad_ptcloud, int B, int scale,
                         float* grad_grid, void* stream);


/* ---- GRNet's gridding loss and cubic feature sampling -----------------------------------
 * snb_gridding_dist_*: replaces gridding_distance.forward / backward (cuda/gridding_loss/gridding_distance_cuda.cpp,
 * gridding_distance.cu:179-338): like snb_gridding_* but every vertex keeps EIGHT accumulators, one per corner role:
 * grid [B, V, 8], grid_pt_indexes = vertex * 8 + corner (-1 outside the grid).
 * snb_cubic_sampling_*: replaces cubic_feature_sampling.forward / backward (cubic_feature_sampling.cu:105-204): ptcloud
 * [B,n,3] in grid units, cubic_features [B,C,S,S,S] -> point_features [B,n,(2 ns)^3,C] (zeros outside the grid) and
 * grid_pt_indexes [B,n,(2 ns)^3]; the backward fully writes grad_cubic_features (the cloud receives no gradient). */
int snb_gridding_dist_fwd(const float* ptcloud, int B, int n, float min_x, float max_x, float min_y, float max_y,
                          float min_z, float max_z, float* grid, float* grid_pt_weights, int* grid_pt_indexes, void* stream);
int snb_gridding_dist_bwd(const float* grid_pt_weights, const int* grid_pt_indexes, const float* grad_grid, int B, int n,
                          long long n_grid_vertices, float* grad_ptcloud, void* stream);
int snb_cubic_sampling_fwd(const float* ptcloud, const float* cubic_features, int B, int n, int C, int scale,
                           int neighborhood_size, float* point_features, int* grid_pt_indexes, void* stream);
int snb_cubic_sampling_bwd(const float* grad_point_features, const int* grid_pt_indexes, int B, int n, int C, int scale,
                           int neighborhood_size, float* grad_cubic_features, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPARENET_B200_H */
