// gemm_test.cu -- standalone check + timing of snb_gemm_tf32 (csrc/gemm_tc.cu) against a float64 host reference, one case per
// process so that a trapped kernel cannot take the later cases with it.   build: tools/build_gemm_test.sh
//   gemm_test <case>      correctness cases 0..N-1 (small shapes, full comparison)
//   gemm_test perf <i>    layer-sized shapes: sampled comparison, CUDA-event timing, cuBLAS TF32 strided-batched GEMM beside it
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "sparenet_b200.h"

#define CK(x)                                                                     \
  do {                                                                            \
    cudaError_t e = (x);                                                          \
    if (e != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                    \
    }                                                                             \
  } while (0)

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static float frand() {  // [-1, 1)
  rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull;
  return (float)((rng_state >> 40) & 0xFFFFFF) / 8388608.0f - 1.0f;
}
static float tf32_trunc(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}
static float tf32_rn(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u += 0x00000FFFu + ((u >> 13) & 1u);
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}

struct Case {
  const char* name;
  int mode, G, BI, M, N, K;
  int lda_pad, a_batched, block_n, store, split, xform, seg, stats, minmax;
  int bmod;   // > 0: the activation operand holds bmod positions, tiled along the position axis
};

static const Case CASES[] = {
    //                         mode G BI   M    N    K  lda+ aB  bn st sp xf  seg stats mm
    {"fwd 1tile 1kb", 0, 1, 1, 128, 256, 32, 0, 0, 0, 1, 0, 0, 0, 0, 0},
    {"fwd 1tile 4kb", 0, 1, 1, 128, 256, 128, 0, 0, 0, 1, 0, 0, 0, 0, 0},
    {"fwd multi-tile ring wrap", 0, 2, 1, 256, 512, 160, 0, 0, 0, 1, 0, 0, 0, 0, 0},
    {"fwd tails M=200 K=72 lda+8 bn128", 0, 2, 1, 200, 512, 72, 8, 0, 128, 1, 0, 0, 0, 0, 0},
    {"fwd batched A M=136 K=40", 0, 3, 1, 136, 256, 40, 0, 1, 0, 1, 0, 0, 0, 0, 0},
    {"fwd stats+minmax", 0, 2, 1, 256, 1024, 64, 0, 0, 0, 1, 0, 0, 0, 1, 1},
    {"fwd xform seg=N stats", 0, 2, 1, 128, 512, 96, 0, 0, 0, 1, 0, 1, 512, 1, 0},
    {"fwd xform seg=256 batched A", 0, 2, 1, 128, 512, 104, 0, 1, 0, 1, 0, 1, 256, 1, 0},
    {"fwd nostore stats+minmax", 0, 2, 1, 256, 512, 64, 0, 0, 0, 0, 0, 0, 0, 1, 1},
    {"fwd bn64 N=64", 0, 2, 1, 128, 64, 64, 0, 0, 0, 1, 0, 0, 0, 0, 0},
    {"dgrad basic", 1, 2, 1, 128, 256, 64, 0, 0, 0, 1, 0, 0, 0, 0, 0},
    {"dgrad tails M=160 K=200", 1, 2, 1, 160, 512, 200, 0, 0, 0, 1, 0, 0, 0, 0, 0},
    {"dgrad batched A", 1, 3, 1, 256, 256, 72, 0, 1, 0, 1, 0, 0, 0, 0, 0},
    {"wgrad basic", 2, 1, 1, 128, 256, 64, 0, 0, 0, 1, 1, 0, 0, 0, 0},
    {"wgrad BI=3 M=200 N=96", 2, 1, 3, 200, 96, 512, 0, 0, 0, 1, 1, 0, 0, 0, 0},
    {"wgrad split4 reduce BI=4", 2, 1, 4, 128, 64, 1024, 0, 0, 0, 2, 4, 0, 0, 0, 0},
    {"wgrad xform G=2 seg=256", 2, 2, 1, 128, 256, 512, 0, 0, 0, 1, 1, 1, 256, 0, 0},
    {"wgrad xform BI=2 seg=K N=160", 2, 1, 2, 128, 160, 256, 0, 0, 0, 1, 1, 1, 256, 0, 0},
    {"wgrad auto split", 2, 1, 8, 256, 128, 2048, 0, 0, 0, 2, 0, 0, 0, 0, 0},
    {"fwd tiled B (bmod 512) xform seg=512 stats", 0, 2, 1, 136, 2048, 72, 0, 1, 0, 1, 0, 1, 512, 1, 0, 512},
    {"wgrad tiled B (bmod 256) xform seg=256", 2, 2, 1, 136, 96, 1024, 0, 0, 0, 2, 1, 1, 256, 0, 0, 256},
    {"wgrad N=1056 auto block_n", 2, 1, 1, 200, 1056, 256, 0, 0, 0, 2, 1, 0, 0, 0, 0, 0},
};
static const int NCASES = sizeof(CASES) / sizeof(CASES[0]);

// layer-sized shapes of the generator step (B=32)
static const Case PERF[] = {
    {"decoder conv2 fwd  P32 544x16384x1056 xform+stats", 0, 32, 1, 544, 16384, 1056, 0, 1, 0, 1, 0, 1, 512, 1, 0},
    {"decoder conv3 fwd  P32 256x16384x544 xform+stats", 0, 32, 1, 256, 16384, 544, 0, 1, 0, 1, 0, 1, 512, 1, 0},
    {"encoder conv5 fwd  B32 2048x2048x2048 plain", 0, 32, 1, 2048, 2048, 2048, 0, 0, 0, 1, 0, 0, 0, 1, 0},
    {"refiner conv3 fwd  B32 1024x16384x128 nostore stats+minmax xform", 0, 32, 1, 1024, 16384, 128, 0, 0, 0, 0, 0, 1, 16384, 1, 1},
    {"refiner conv5 fwd  B32 256x16384x512 xform+stats", 0, 32, 1, 256, 16384, 512, 0, 0, 0, 1, 0, 1, 16384, 1, 0},
    {"refiner conv2 fwd  B32 128x16384x64 xform+stats", 0, 32, 1, 128, 16384, 64, 0, 0, 0, 1, 0, 1, 16384, 1, 0},
    {"decoder conv2 dgrad P32 1056x16384x544", 1, 32, 1, 1056, 16384, 544, 0, 1, 0, 1, 0, 0, 0, 0, 0},
    {"encoder conv5 dgrad B32 2048x2048x2048", 1, 32, 1, 2048, 2048, 2048, 0, 0, 0, 1, 0, 0, 0, 0, 0},
    {"decoder conv2 wgrad P32 544x1056x16384 xform", 2, 32, 1, 544, 1056, 16384, 0, 0, 0, 1, 1, 1, 512, 0, 0},
    {"encoder conv5 wgrad BI32 2048x2048x2048", 2, 1, 32, 2048, 2048, 2048, 0, 0, 0, 1, 1, 0, 0, 0, 0},
    {"refiner conv5 wgrad BI32 256x512x16384 xform autosplit", 2, 1, 32, 256, 512, 16384, 0, 0, 0, 2, 0, 1, 16384, 0, 0},
};
static const int NPERF = sizeof(PERF) / sizeof(PERF[0]);

struct Problem {
  Case c;
  size_t a_elems, b_elems, d_elems, p_elems;
  long long lda, ldb, ldd, a_bs, b_bs, d_bs;
  int act_batches, cin, S, nt, bn;
  std::vector<float> A, B, sc, sh, D0;
};

static void setup(Problem& P, const Case& c) {
  P.c = c;
  const int BI = c.mode == 2 ? c.BI : 1;
  P.act_batches = c.G * BI;
  if (c.mode == 0) {  // A [Ga, M, K]; B [G, K, N]
    P.lda = c.K + c.lda_pad;
    P.a_bs = (long long)c.M * P.lda;
    P.a_elems = (size_t)(c.a_batched ? c.G : 1) * P.a_bs;
    P.ldb = c.bmod > 0 ? c.bmod : c.N;
    P.b_bs = (long long)c.K * P.ldb;
    P.b_elems = (size_t)c.G * P.b_bs;
    P.cin = c.K;
  } else if (c.mode == 1) {  // A = W [Ga, K rows, M]; B [G, K, N]
    P.lda = c.M + c.lda_pad;
    P.a_bs = (long long)c.K * P.lda;
    P.a_elems = (size_t)(c.a_batched ? c.G : 1) * P.a_bs;
    P.ldb = c.N;
    P.b_bs = (long long)c.K * c.N;
    P.b_elems = (size_t)c.G * P.b_bs;
    P.cin = c.K;
  } else {  // A = gY [G*BI, M, K]; B = X [G*BI, N, K]
    P.lda = c.K;
    P.a_bs = (long long)c.M * c.K;
    P.a_elems = (size_t)P.act_batches * P.a_bs;
    P.ldb = c.bmod > 0 ? c.bmod : c.K;
    P.b_bs = (long long)c.N * P.ldb;
    P.b_elems = (size_t)P.act_batches * P.b_bs;
    P.cin = c.N;
  }
  P.ldd = c.N;
  P.d_bs = (long long)c.M * c.N;
  P.d_elems = (size_t)c.G * P.d_bs;
  P.bn = snb_gemm_tf32_block_n(c.N, c.block_n) / 2;   // width of one statistics tile = block_n / 2
  P.nt = snb_gemm_tf32_tiles(c.N, c.block_n);
  P.p_elems = (size_t)c.G * c.M * P.nt;
  const int npos = c.mode == 2 ? c.K : c.N;
  P.S = c.xform ? (npos + c.seg - 1) / c.seg : 1;
  P.A.resize(P.a_elems);
  P.B.resize(P.b_elems);
  for (auto& x : P.A) x = frand();
  for (auto& x : P.B) x = frand();
  if (c.xform) {
    P.sc.resize((size_t)P.act_batches * P.cin * P.S);
    P.sh.resize(P.sc.size());
    for (auto& x : P.sc) x = 0.5f + 0.5f * frand();
    for (auto& x : P.sh) x = 0.3f * frand();
  }
  if (c.store == 2) {
    P.D0.resize(P.d_elems);
    for (auto& x : P.D0) x = frand();
  }
}

static inline float xf(const Problem& P, float x, int batch, int ch, int pos) {
  if (!P.c.xform) return x;
  const size_t i = ((size_t)batch * P.cin + ch) * P.S + pos / P.c.seg;
  const float t = fmaf(x, P.sc[i], P.sh[i]);
  return t > 0.f ? t : t * 0.2f;
}

// reference value of D[g][m][n] in float64; conv = 0 exact inputs, 1 tf32-truncated, 2 tf32 round-to-nearest
static double ref_elem(const Problem& P, int g, int m, int n, int conv) {
  const Case& c = P.c;
  auto cv = [&](float x) { return conv == 0 ? x : (conv == 1 ? tf32_trunc(x) : tf32_rn(x)); };
  double acc = 0;
  if (c.mode == 0) {
    const float* A = P.A.data() + (c.a_batched ? (size_t)g * P.a_bs : 0) + (size_t)m * P.lda;
    const float* B = P.B.data() + (size_t)g * P.b_bs;
    const int nb = c.bmod > 0 ? n % c.bmod : n;
    for (int k = 0; k < c.K; ++k) acc += (double)cv(A[k]) * (double)cv(xf(P, B[(size_t)k * P.ldb + nb], g, k, n));
  } else if (c.mode == 1) {
    const float* A = P.A.data() + (c.a_batched ? (size_t)g * P.a_bs : 0);
    const float* B = P.B.data() + (size_t)g * P.b_bs;
    for (int k = 0; k < c.K; ++k) acc += (double)cv(A[(size_t)k * P.lda + m]) * (double)cv(B[(size_t)k * P.ldb + n]);
  } else {
    for (int bi = 0; bi < c.BI; ++bi) {
      const int b = g * c.BI + bi;
      const float* A = P.A.data() + (size_t)b * P.a_bs + (size_t)m * P.lda;
      const float* B = P.B.data() + (size_t)b * P.b_bs + (size_t)n * P.ldb;
      for (int k = 0; k < c.K; ++k) acc += (double)cv(A[k]) * (double)cv(xf(P, B[c.bmod > 0 ? k % c.bmod : k], b, n, k));
    }
  }
  if (c.store == 2) acc += P.D0[((size_t)g * c.M + m) * c.N + n];
  return acc;
}

struct Dev {
  float *A, *B, *D, *sc, *sh, *pmean, *pm2, *pmax, *pmin;
  int *pimax, *pimin;
};

static void upload(const Problem& P, Dev& d) {
  memset(&d, 0, sizeof(d));
  CK(cudaMalloc(&d.A, P.a_elems * 4));
  CK(cudaMalloc(&d.B, P.b_elems * 4));
  CK(cudaMalloc(&d.D, P.d_elems * 4));
  CK(cudaMemcpy(d.A, P.A.data(), P.a_elems * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d.B, P.B.data(), P.b_elems * 4, cudaMemcpyHostToDevice));
  if (P.c.store == 2) CK(cudaMemcpy(d.D, P.D0.data(), P.d_elems * 4, cudaMemcpyHostToDevice));
  else CK(cudaMemset(d.D, 0xFF, P.d_elems * 4));  // NaN canary
  if (P.c.xform) {
    CK(cudaMalloc(&d.sc, P.sc.size() * 4));
    CK(cudaMalloc(&d.sh, P.sh.size() * 4));
    CK(cudaMemcpy(d.sc, P.sc.data(), P.sc.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.sh, P.sh.data(), P.sh.size() * 4, cudaMemcpyHostToDevice));
  }
  if (P.c.stats) {
    CK(cudaMalloc(&d.pmean, P.p_elems * 4));
    CK(cudaMalloc(&d.pm2, P.p_elems * 4));
  }
  if (P.c.minmax) {
    CK(cudaMalloc(&d.pmax, P.p_elems * 4));
    CK(cudaMalloc(&d.pmin, P.p_elems * 4));
    CK(cudaMalloc(&d.pimax, P.p_elems * 4));
    CK(cudaMalloc(&d.pimin, P.p_elems * 4));
  }
}

static snb_gemm_desc make_desc(const Problem& P, const Dev& d) {
  snb_gemm_desc g;
  memset(&g, 0, sizeof(g));
  const Case& c = P.c;
  g.mode = c.mode;
  g.G = c.G;
  g.BI = c.BI;
  g.M = c.M;
  g.N = c.N;
  g.K = c.K;
  g.A = d.A;
  g.lda = P.lda;
  g.a_batch_stride = (c.mode == 2 || c.a_batched) ? P.a_bs : 0;
  g.B = d.B;
  g.ldb = P.ldb;
  g.b_batch_stride = P.b_bs;
  g.D = d.D;
  g.ldd = P.ldd;
  g.d_batch_stride = P.d_bs;
  g.block_n = c.block_n;
  g.store = c.store;
  g.split = c.split;
  g.scale = d.sc;
  g.shift = d.sh;
  g.slope = 0.2f;
  g.seg = c.seg;
  g.pmean = d.pmean;
  g.pm2 = d.pm2;
  g.pmax = d.pmax;
  g.pmin = d.pmin;
  g.pimax = d.pimax;
  g.pimin = d.pimin;
  g.b_pos_mod = c.bmod;
  return g;
}

static int run_case(const Case& c, bool perf) {
  printf("=== %s: mode %d G %d BI %d M %d N %d K %d bn %d store %d split %d xform %d seg %d stats %d minmax %d\n", c.name, c.mode, c.G, c.BI, c.M,
         c.N, c.K, c.block_n, c.store, c.split, c.xform, c.seg, c.stats, c.minmax);
  fflush(stdout);
  Problem P;
  setup(P, c);
  Dev d;
  upload(P, d);
  snb_gemm_desc g = make_desc(P, d);
  int rc = snb_gemm_tf32(&g, 0);
  if (rc != 0) {
    printf("FAIL: snb_gemm_tf32 returned %d\n", rc);
    return 1;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("FAIL: kernel error %s\n", cudaGetErrorString(e));
    return 1;
  }
  int bad = 0;
  std::vector<float> D(P.d_elems);
  double scale_ref = 0;
  if (c.store) {
    CK(cudaMemcpy(D.data(), d.D, P.d_elems * 4, cudaMemcpyDeviceToHost));
    // full compare for small problems, 20000 samples otherwise
    const size_t total = P.d_elems;
    const bool full = !perf && total <= (1u << 21);
    const double cost = (double)c.K * (c.mode == 2 ? c.BI : 1) * 3.0;
    size_t nsamp = full ? total : 20000;
    if (!full && nsamp * cost > 6e8) nsamp = (size_t)(6e8 / cost) + 64;
    double e0 = 0, e1 = 0, e2 = 0;
    int shown = 0;
    long long errq[4] = {0, 0, 0, 0}, errc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (size_t s = 0; s < nsamp; ++s) {
      size_t i = full ? s : (size_t)((double)((rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull) >> 11) / 9007199254740992.0 * total);
      const int gi = (int)(i / P.d_bs), m = (int)((i % P.d_bs) / c.N), n = (int)(i % c.N);
      const double r0 = ref_elem(P, gi, m, n, 0), r1 = ref_elem(P, gi, m, n, 1), r2 = ref_elem(P, gi, m, n, 2);
      const double got = D[i];
      scale_ref = fmax(scale_ref, fabs(r0));
      const double d0 = fabs(got - r0), d1 = fabs(got - r1), d2 = fabs(got - r2);
      e0 = fmax(e0, isnan(got) ? 1e30 : d0);
      e1 = fmax(e1, isnan(got) ? 1e30 : d1);
      e2 = fmax(e2, isnan(got) ? 1e30 : d2);
      const double tol = 2e-3 * sqrt((double)c.K * (c.mode == 2 ? c.BI : 1)) + 1e-3;
      if (!(d0 <= tol)) {
        ++bad;
        errq[(m % 128) / 32]++;
        errc[(n % 256) / 32]++;
        if (shown < 12) {
          printf("  mismatch g %d m %d n %d: got %.6f want %.6f\n", gi, m, n, got, r0);
          ++shown;
        }
      }
    }
    printf("  D: max|err| vs exact %.3e, vs tf32-trunc %.3e, vs tf32-rn %.3e (max|ref| %.3f); mismatches %d / %zu\n", e0, e1, e2, scale_ref, bad, nsamp);
    if (bad)
      printf("  mismatch histogram by row quarter %lld %lld %lld %lld; by 32-col chunk %lld %lld %lld %lld %lld %lld %lld %lld\n", errq[0], errq[1], errq[2],
             errq[3], errc[0], errc[1], errc[2], errc[3], errc[4], errc[5], errc[6], errc[7]);
  }
  if (c.stats || c.minmax) {
    std::vector<float> pm(P.p_elems), p2(P.p_elems), px(P.p_elems), pn(P.p_elems);
    std::vector<int> ix(P.p_elems), in_(P.p_elems);
    if (c.stats) {
      CK(cudaMemcpy(pm.data(), d.pmean, P.p_elems * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(p2.data(), d.pm2, P.p_elems * 4, cudaMemcpyDeviceToHost));
    }
    if (c.minmax) {
      CK(cudaMemcpy(px.data(), d.pmax, P.p_elems * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(pn.data(), d.pmin, P.p_elems * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(ix.data(), d.pimax, P.p_elems * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(in_.data(), d.pimin, P.p_elems * 4, cudaMemcpyDeviceToHost));
    }
    // reference statistics from the device D when stored (isolates the epilogue), else from the float64 reference
    double em = 0, e2 = 0, ex = 0;
    int sbad = 0, shown = 0;
    const size_t rows = (size_t)c.G * c.M;
    const size_t step = perf ? (rows / 64 > 0 ? rows / 64 : 1) : 1;
    for (size_t r = 0; r < rows; r += step) {
      const int gi = (int)(r / c.M), m = (int)(r % c.M);
      for (int t = 0; t < P.nt; ++t) {
        double s = 0, mx = -1e30, mn = 1e30;
        int imx = -1, imn = -1;
        std::vector<double> vals(P.bn);
        for (int j = 0; j < P.bn; ++j) {
          const int n = t * P.bn + j;
          const double v = c.store ? (double)D[((size_t)gi * c.M + m) * c.N + n] : ref_elem(P, gi, m, n, 1);
          vals[j] = v;
          s += v;
          if (v > mx) { mx = v; imx = n; }
          if (v < mn) { mn = v; imn = n; }
        }
        const double mean = s / P.bn;
        double m2 = 0;
        for (int j = 0; j < P.bn; ++j) m2 += (vals[j] - mean) * (vals[j] - mean);
        const size_t o = r * P.nt + t;
        if (c.stats) {
          const double dm = fabs(pm[o] - mean), d2 = fabs(p2[o] - m2) / fmax(1.0, m2);
          em = fmax(em, dm);
          e2 = fmax(e2, d2);
          const double tolm = c.store ? 1e-4 : 2e-2, tol2 = c.store ? 1e-4 : 2e-2;
          if (!(dm <= tolm) || !(d2 <= tol2)) {
            ++sbad;
            if (shown++ < 6) printf("  stats mismatch g %d m %d tile %d: mean %.6f want %.6f, m2 %.6f want %.6f\n", gi, m, t, pm[o], mean, p2[o], m2);
          }
        }
        if (c.minmax) {
          const double dx = fmax(fabs(px[o] - mx), fabs(pn[o] - mn));
          ex = fmax(ex, dx);
          const bool idx_ok = c.store ? (ix[o] == imx && in_[o] == imn) : true;
          if (!(dx <= (c.store ? 0.0 : 2e-2)) || !idx_ok) {
            ++sbad;
            if (shown++ < 6)
              printf("  minmax mismatch g %d m %d tile %d: max %.6f@%d want %.6f@%d, min %.6f@%d want %.6f@%d\n", gi, m, t, px[o], ix[o], mx, imx, pn[o], in_[o], mn,
                     imn);
          }
        }
      }
    }
    printf("  stats: max|mean err| %.3e, max rel m2 err %.3e, max|extremum err| %.3e; mismatches %d\n", em, e2, ex, sbad);
    bad += sbad;
  }
  if (perf) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) snb_gemm_tf32(&g, 0);
    CK(cudaDeviceSynchronize());
    const int iters = 10;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) snb_gemm_tf32(&g, 0);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= iters;
    const double flops = 2.0 * c.G * (double)c.M * c.N * c.K * (c.mode == 2 ? c.BI : 1);
    double bytes = (double)P.b_elems * 4 + (double)P.a_elems * 4 + (c.store ? (double)P.d_elems * 4 : 0);
    printf("  ours: %.3f ms  %.1f TFLOP/s  %.0f GB/s (algorithmic bytes %.1f MB)\n", ms, flops / ms * 1e-9, bytes / ms * 1e-6, bytes * 1e-6);
    // cuBLAS TF32 beside it (plain GEMM of the same shape: no prologue, no statistics)
    cublasHandle_t h;
    if (cublasCreate(&h) == CUBLAS_STATUS_SUCCESS) {
      cublasSetMathMode(h, CUBLAS_TF32_TENSOR_OP_MATH);
      const float one = 1.f, zero = 0.f;
      // row-major D[M,N] = A[M,K] B[K,N]  <=>  column-major D^T[N,M] = B^T[N,K] A^T[K,M]
      auto call = [&]() {
        if (c.mode == 0)
          return cublasGemmStridedBatchedEx(h, CUBLAS_OP_N, CUBLAS_OP_N, c.N, c.M, c.K, &one, d.B, CUDA_R_32F, (int)P.ldb, P.b_bs, d.A, CUDA_R_32F, (int)P.lda,
                                            c.a_batched ? P.a_bs : 0, &zero, d.D, CUDA_R_32F, (int)P.ldd, P.d_bs, c.G, CUBLAS_COMPUTE_32F_FAST_TF32,
                                            CUBLAS_GEMM_DEFAULT);
        if (c.mode == 1)  // D[M,N] = W^T B: A^T in column-major terms is W [K rows, M] read as M x K col-major -> op T
          return cublasGemmStridedBatchedEx(h, CUBLAS_OP_N, CUBLAS_OP_T, c.N, c.M, c.K, &one, d.B, CUDA_R_32F, (int)P.ldb, P.b_bs, d.A, CUDA_R_32F, (int)P.lda,
                                            c.a_batched ? P.a_bs : 0, &zero, d.D, CUDA_R_32F, (int)P.ldd, P.d_bs, c.G, CUBLAS_COMPUTE_32F_FAST_TF32,
                                            CUBLAS_GEMM_DEFAULT);
        // WGRAD: D[M,N] = A[M,Kpos] B[N,Kpos]^T with (BI*Kpos) as one long K when BI == 1 or the batch is contiguous
        return cublasGemmStridedBatchedEx(h, CUBLAS_OP_T, CUBLAS_OP_N, c.N, c.M, c.K, &one, d.B, CUDA_R_32F, (int)P.ldb, P.b_bs * c.BI, d.A, CUDA_R_32F,
                                          (int)P.lda, P.a_bs * c.BI, &zero, d.D, CUDA_R_32F, (int)P.ldd, P.d_bs, c.G, CUBLAS_COMPUTE_32F_FAST_TF32,
                                          CUBLAS_GEMM_DEFAULT);
      };
      cublasStatus_t st = call();
      if (st == CUBLAS_STATUS_SUCCESS) {
        for (int i = 0; i < 2; ++i) call();
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int i = 0; i < iters; ++i) call();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms2 = 0;
        CK(cudaEventElapsedTime(&ms2, e0, e1));
        ms2 /= iters;
        const double f2 = c.mode == 2 ? flops / c.BI : flops;   // the cuBLAS call above covers one inner batch for WGRAD
        printf("  cuBLAS tf32 (plain%s): %.3f ms  %.1f TFLOP/s\n", c.mode == 2 && c.BI > 1 ? ", ONE inner batch" : "", ms2, f2 / ms2 * 1e-9);
      } else {
        printf("  cuBLAS call failed (%d)\n", (int)st);
      }
      cublasDestroy(h);
    }
  }
  printf("%s: %s\n", bad ? "FAIL" : "PASS", c.name);
  return bad ? 1 : 0;
}

int main(int argc, char** argv) {
  if (argc >= 3 && strcmp(argv[1], "perf") == 0) {
    const int i = atoi(argv[2]);
    if (i < 0 || i >= NPERF) return 3;
    return run_case(PERF[i], true);
  }
  if (argc >= 2 && strcmp(argv[1], "count") == 0) {
    printf("%d %d\n", NCASES, NPERF);
    return 0;
  }
  const int i = argc >= 2 ? atoi(argv[1]) : 0;
  if (i < 0 || i >= NCASES) return 3;
  return run_case(CASES[i], false);
}
