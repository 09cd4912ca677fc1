// Neural constitutive law, fused: SVD -> invariants -> 13->H->H->9 bias-free GELU
// MLP -> symmetrise -> polar recombination, forward and backward, one kernel each.
// Replaces ~35 torch launches + 1 Warp launch per call of
//   InvariantFullMetaElasticity.forward   modules/nclaw/material/meta.py:196-221
//   InvariantFullMetaPlasticity.forward   modules/nclaw/material/meta.py:468-489
// and the Warp batch_svd / SVDFunction of modules/nclaw/warp/svd.py:9-96.
//
// Structure (one persistent CTA per SM):
//   * NPIPE independent tile pipelines per CTA.  A pipeline = 128 threads = the 128 particles
//     of a tile = the 128 TMEM lanes; one thread owns one particle row for the SVD /
//     invariants, the GELU epilogues between the GEMMs, and the polar recombination.  A
//     pipeline synchronises on its own named barrier, its first warp issues the tcgen05.mma
//     of the pipeline's next GEMM (forward and data-gradient GEMMs: nothing else of that
//     pipeline can run until they finish anyway) and everybody waits on the commit mbarrier.
//     While one pipeline waits for the tensor core the others run their CUDA-core phases.
//   * Backward only: ONE extra warp issues every weight-gradient GEMM of the CTA.  The
//     accumulators are shared by the pipelines and stay in TMEM for the whole kernel, so a
//     single issuing thread keeps them well ordered; the pipelines hand it work through
//     mbarriers (wfull[p]: "both operand tiles are in shared memory", wdone[p]:
//     tcgen05.commit, after which the pipeline may overwrite the tiles) and never wait for
//     the weight gradients themselves.
//   * The three Linear layers -- and, in the backward, their data-gradient and
//     weight-gradient GEMMs -- are tcgen05.mma kind::f16 on bf16 operands in shared memory
//     with fp32 accumulators in TMEM.  Every fp32 value is split into NS bf16 terms
//     (x = x0 + x1 [+ x2]) and a GEMM is the products with i + j < NS, all accumulated into
//     the same TMEM columns.  Forward: NS = 3 (24 mantissa bits, ~2e-7 relative to fp64);
//     backward: NS = 2 (16+ mantissa bits, ~4e-6: gradients are compared at 5e-4).
//   * Weights are split and laid out ONCE per weight update by stage_weights_kernel into a
//     60 KB global image; each CTA pulls its copy with cp.async.bulk.
//
// Shared-memory operand layout (SWIZZLE_NONE core matrices of 8 rows x 16 bytes, see
// umma.cuh): an [R x K] tile is stored per split as [K/8][R][8] bf16.  The same bytes are a
// K-major operand (forward GEMMs: rows = particles or output features) and, with LBO/SBO
// swapped and the major bit set, an MN-major operand (the transposed matrix): the
// data-gradient GEMMs read the forward weight image transposed and the weight-gradient
// GEMMs read the activation tiles transposed (contraction over the tile's 128 particles)
// -- no transposed copy is ever written.
//
// LoRA adaptors are merged into the weights by the host (W + scaling * B A,
// loralib.py:216-224) and the weight gradients un-merged there.
//
// Gradient of the SVD part: everything downstream of the SVD depends on U, V only
// through R = U V^T, so the backward uses the polar-decomposition form
//   dF += U [ (M - M^T)_ij / (sigma_i + sigma_j) ] V^T,  M = U^T (dL/dR) V
// which is the exact sum of the dU and dV terms of the textbook SVD adjoint but has
// no 1/(sigma_i^2 - sigma_j^2) pole at F ~ I (SURVEY 7 "SVD-gradient degeneracy").
#include "common.cuh"
#include "mpm_common.cuh"
#include "svd3.cuh"
#include "umma.cuh"
#include <stdlib.h>

// Developer instrumentation (tools/phase_timing.py builds a copy with -DNEUMA_PHASE_TIMING):
// thread 0 of CTA 0 accumulates clock64() deltas per phase of its pipeline.
#ifdef NEUMA_PHASE_TIMING
__device__ long long g_phase_cycles[64];
#define PHASE_T0() long long pt__ = clock64()
#define PHASE(i)                                             \
  do {                                                       \
    if (blockIdx.x == 0 && threadIdx.x == 0) {               \
      const long long t__ = clock64();                       \
      g_phase_cycles[i] += t__ - pt__;                       \
      pt__ = t__;                                            \
    }                                                        \
  } while (0)
#else
#define PHASE_T0()
#define PHASE(i)
#endif

namespace {

constexpr int H = 64;      // hidden width (layer_widths: [64, 64] in every shipped config)
constexpr int NF = 13;     // 3 + 9 + 1 invariants
constexpr int TP = 128;    // particles per tile = TMEM lanes = rows of every GEMM
constexpr int KF = 16;     // invariants (13) and MLP head (9) padded to one bf16 K step
#ifndef NEUMA_NS_FWD
#define NEUMA_NS_FWD 3
#endif
#ifndef NEUMA_NPIPE_FWD
#define NEUMA_NPIPE_FWD 3
#endif
constexpr int NS_FWD = NEUMA_NS_FWD, NS_BWD = 2;          // bf16 splits per fp32 value
constexpr int NPIPE_FWD = NEUMA_NPIPE_FWD, NPIPE_BWD = 2;  // tile pipelines per CTA

struct Kinematics {
  Mat3 F, U, V, R;
  float sig[3];
  float feat[NF];
};

__device__ __forceinline__ void kinematics(const Mat3& F, bool normalize, Kinematics& k) {
  k.F = F;
  svd3_fast(k.F, k.U, k.sig, k.V);
  k.R = mat3_mul_nt(k.U, k.V);
  const Mat3 FtF = mat3_mul_tn(k.F, k.F);
  const float sub = normalize ? 1.0f : 0.0f;
#pragma unroll
  for (int i = 0; i < 3; ++i) k.feat[i] = k.sig[i] - sub;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) k.feat[3 + 3 * i + j] = FtF(i, j) - ((i == j) ? sub : 0.0f);
  k.feat[12] = mat3_det(k.F) - sub;
}
// The same with the right singular vectors V of F already known (saved by the forward kernel): only
// the sort-free tail of the SVD (U and sigma from F V by Gram-Schmidt) is left -- the Jacobi sweeps were
// 16 % of a backward tile's latency.  V is the SORTED V the forward produced, so the tail's column sort is
// a no-op and U, sigma come out as in the forward.
__device__ __forceinline__ void kinematics_from_v(const Mat3& F, const float* __restrict__ Vrow, bool normalize,
                                                  Kinematics& k) {
  k.F = F;
  float V[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) V[i][j] = Vrow ? Vrow[3 * i + j] : ((i == j) ? 1.0f : 0.0f);
  svd3_finish(k.F, V, k.U, k.sig, k.V);
  k.R = mat3_mul_nt(k.U, k.V);
  const Mat3 FtF = mat3_mul_tn(k.F, k.F);
  const float sub = normalize ? 1.0f : 0.0f;
#pragma unroll
  for (int i = 0; i < 3; ++i) k.feat[i] = k.sig[i] - sub;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) k.feat[3 + 3 * i + j] = FtF(i, j) - ((i == j) ? sub : 0.0f);
  k.feat[12] = mat3_det(k.F) - sub;
}
__device__ __forceinline__ Mat3 load_F_or_identity(const float* __restrict__ F, int p, int n) {
  Mat3 m;
  if (p < n) {
    m = mat3_load(F + (size_t)p * 9);
  } else {  // padding rows of the last tile run the same code on the identity
#pragma unroll
    for (int i = 0; i < 9; ++i) m.m[i] = (i % 4 == 0) ? 1.0f : 0.0f;
  }
  return m;
}

// pull a particle's 36-byte row towards L1 ahead of its use (next tile's F, this tile's dL/dout)
__device__ __forceinline__ void prefetch_row(const float* __restrict__ base, int p, int n) {
  if (p < n) {
    const char* a = reinterpret_cast<const char*>(base + (size_t)p * 9);
    asm volatile("prefetch.global.L1 [%0];" ::"l"(a));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(a + 32));
  }
}

// ---- shared-memory / staged-image map (bytes)
constexpr uint32_t SMALL_SPLIT = TP * KF * 2;  // [128 x 16] activation tile, one split
constexpr uint32_t BIG_SPLIT = TP * H * 2;     // [128 x 64] activation tile, one split
constexpr uint32_t W0_SPLIT = H * KF * 2, W1_SPLIT = H * H * 2, W2_SPLIT = KF * H * 2;
template <int NS>
struct WImg {  // [W0: NS splits][W1: NS splits][W2: NS splits], each split [K/8][R][8] bf16
  static constexpr uint32_t OFF_W0 = 0, OFF_W1 = NS * W0_SPLIT, OFF_W2 = OFF_W1 + NS * W1_SPLIT,
                            BYTES = OFF_W2 + NS * W2_SPLIT;
};
constexpr uint32_t STAGED_OFF_FWD = 0, STAGED_OFF_BWD = WImg<NS_FWD>::BYTES,
                   STAGED_BYTES = STAGED_OFF_BWD + WImg<NS_BWD>::BYTES;  // 61440
// per-pipeline activation tiles
template <int NS>
struct PipeFwd {
  static constexpr uint32_t OFF_FEAT = 0, OFF_BIGA = NS * SMALL_SPLIT, BYTES = OFF_BIGA + NS * BIG_SPLIT;
};
template <int NS>
struct PipeBwd {
  static constexpr uint32_t OFF_FEAT = 0, OFF_GO = NS * SMALL_SPLIT, OFF_BIGA = 2 * NS * SMALL_SPLIT,
                            OFF_BIGB = OFF_BIGA + NS * BIG_SPLIT, BYTES = OFF_BIGB + NS * BIG_SPLIT;
};
constexpr uint32_t FWD_SMEM = WImg<NS_FWD>::BYTES + NPIPE_FWD * PipeFwd<NS_FWD>::BYTES;  // 221184
constexpr uint32_t BWD_SMEM = WImg<NS_BWD>::BYTES + NPIPE_BWD * PipeBwd<NS_BWD>::BYTES;  // 188416
// RT = threads per particle row (1 or 2).  With RT = 2 a pipeline has 8 warps: warps 0-3 ("owners")
// hold the particle rows -- SVD, invariants, output head -- and the columns [0, 32) of the GELU /
// gradient phases, warps 4-7 ("helpers") the columns [32, 64) of the same rows (a warp reaches the
// TMEM lanes 32 (warp % 4) ..., so warp w and w + 4 share a lane quarter).  Twice the warps per SM
// for the phases that dominate (GELU + split + store: 60 % of a tile), same shared memory.
template <int RT> constexpr int fwd_threads() { return NPIPE_FWD * TP * RT + 32; }
template <int RT> constexpr int bwd_threads() { return NPIPE_BWD * TP * RT + 32; }
// backward, RT = 2: F, U, V, sigma of the row parked in shared memory between the phases
// ([element][row]: conflict-free), so that 544 threads fit the register file
constexpr int STASH_FLOATS = 30;
constexpr uint32_t STASH_BYTES = STASH_FLOATS * TP * 4;
template <int RT> constexpr uint32_t bwd_smem() { return BWD_SMEM + (RT == 2 ? NPIPE_BWD * STASH_BYTES : 0u); }

// ---- TMEM columns (fp32, lane = particle row unless noted)
constexpr uint32_t FWD_PIPE_COLS = 64, FWD_TMEM_COLS = 256;
// backward, per pipeline: h1 -> gelu'(h1) in place | h2 -> gelu'(h2) | current GEMM result
constexpr uint32_t BWD_PIPE_COLS = 192, COL_H1 = 0, COL_H2 = 64, COL_S = 128;
// shared weight-gradient accumulators (M = 64 layout: row o -> lane (o%16) + 32 (o/16))
constexpr uint32_t COL_GW1 = NPIPE_BWD * BWD_PIPE_COLS, COL_GW0 = COL_GW1 + 64, COL_GW2T = COL_GW0 + 16;
constexpr uint32_t BWD_TMEM_COLS = 512;
// per-CTA weight-gradient partial sums: [dW1 64x64 | dW0 64x13 | dW2 9x64]
constexpr int GW_OFF0 = H * H, GW_OFF2 = GW_OFF0 + H * NF, GW_FLOATS = GW_OFF2 + 9 * H;  // 5504
static_assert(COL_GW2T + 16 <= BWD_TMEM_COLS, "TMEM budget");
static_assert(NPIPE_FWD * FWD_PIPE_COLS <= FWD_TMEM_COLS, "TMEM budget");

// ---- fp32 -> NS bf16 terms, two values at a time (cvt.rn.bf16x2: first source -> upper half)
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
template <int NS>
__device__ __forceinline__ void split_pair(float a, float b, uint32_t (&s)[NS]) {
#pragma unroll
  for (int t = 0; t < NS; ++t) {
    s[t] = pack_bf16x2(a, b);
    if (t + 1 < NS) {
      a -= __uint_as_float(s[t] << 16);
      b -= __uint_as_float(s[t] & 0xFFFF0000u);
    }
  }
}
// eight consecutive K values of particle row r -> chunk c of the NS split images of a tile
template <int NS>
__device__ __forceinline__ void store_chunk(uint8_t* img, uint32_t split_bytes, int c, int r, const float* x) {
  uint32_t s[4][NS];
#pragma unroll
  for (int q = 0; q < 4; ++q) split_pair<NS>(x[2 * q], x[2 * q + 1], s[q]);
  uint8_t* p = img + ((size_t)c * TP + r) * 16;
#pragma unroll
  for (int t = 0; t < NS; ++t)
    *reinterpret_cast<uint4*>(p + t * split_bytes) = make_uint4(s[0][t], s[1][t], s[2][t], s[3][t]);
}

// weights -> the staged image (both split counts), run once per weight update.
// One thread per element of the zero-padded [R x K] matrices.
__global__ void __launch_bounds__(256)
stage_weights_kernel(const float* __restrict__ W0, const float* __restrict__ W1,
                     const float* __restrict__ W2, uint8_t* __restrict__ staged) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  constexpr int N0 = H * KF, N1 = H * H, N2 = KF * H;
  if (e >= N0 + N1 + N2) return;
  int r, k, R;
  float x;
  uint32_t off3, off2, split;
  if (e < N0) {
    r = e / KF; k = e % KF; R = H;
    x = (k < NF) ? W0[r * NF + k] : 0.f;
    off3 = WImg<NS_FWD>::OFF_W0; off2 = WImg<NS_BWD>::OFF_W0; split = W0_SPLIT;
  } else if (e < N0 + N1) {
    const int f = e - N0;
    r = f / H; k = f % H; R = H;
    x = W1[r * H + k];
    off3 = WImg<NS_FWD>::OFF_W1; off2 = WImg<NS_BWD>::OFF_W1; split = W1_SPLIT;
  } else {
    const int f = e - N0 - N1;
    r = f / H; k = f % H; R = KF;
    x = (r < 9) ? W2[r * H + k] : 0.f;
    off3 = WImg<NS_FWD>::OFF_W2; off2 = WImg<NS_BWD>::OFF_W2; split = W2_SPLIT;
  }
  const size_t o = (((size_t)(k >> 3) * R + r) * 8 + (k & 7)) * 2;
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    const uint16_t h = (uint16_t)(pack_bf16x2(x, 0.f) & 0xFFFFu);
    if (t < NS_FWD) *reinterpret_cast<uint16_t*>(staged + STAGED_OFF_FWD + off3 + t * split + o) = h;
    if (t < NS_BWD) *reinterpret_cast<uint16_t*>(staged + STAGED_OFF_BWD + off2 + t * split + o) = h;
    x -= __uint_as_float((uint32_t)h << 16);
  }
}

// ---- operand descriptors
struct Desc {
  uint64_t d;       // descriptor of split 0, K step 0
  uint32_t split;   // (bytes >> 4) between splits
  uint32_t kstep;   // (bytes >> 4) per K = 16 step
};
// [K/8][R][8] read K-major (rows = M/N index): LBO = R*16 (next 8 K), SBO = 128 (next 8 rows)
__device__ __forceinline__ Desc desc_kmajor(uint32_t addr, uint32_t split_bytes, int R) {
  return Desc{umma::smem_desc(addr, (uint32_t)R * 16u, 128u), split_bytes >> 4, ((uint32_t)R * 32u) >> 4};
}
// the same bytes read transposed (MN-major; rows = K index): SBO = R*16, LBO = 128, 16 rows per K step
__device__ __forceinline__ Desc desc_mnmajor(uint32_t addr, uint32_t split_bytes, int R) {
  return Desc{umma::smem_desc(addr, 128u, (uint32_t)R * 16u), split_bytes >> 4, 256u >> 4};
}

// D[tmem] (+)= sum_{i + j < NS} A_i B_j^T over NK K-steps.  Called by ALL lanes of the MMA warp.
template <int NS, int NK>
__device__ __forceinline__ void gemm_split(uint32_t tmem_d, const Desc a, const Desc b, uint32_t idesc,
                                           bool accumulate) {
  // K loop deliberately rolled: the MMA warp's code must stay small (instruction cache is shared
  // with two or three compute pipelines running different phases)
#pragma unroll 1
  for (int k = 0; k < NK; ++k) {
    const uint64_t ak = a.d + (uint64_t)(k * a.kstep), bk = b.d + (uint64_t)(k * b.kstep);
#pragma unroll
    for (int i = 0; i < NS; ++i)
#pragma unroll
      for (int j = 0; j < NS - i; ++j)
        umma::mma_bf16_ss(tmem_d, ak + (uint64_t)(i * a.split), bk + (uint64_t)(j * b.split), idesc,
                          (accumulate || k || i || j) ? 1u : 0u);
  }
}

// Phi(h) and phi(h) of the standard normal for the exact-erf GELU (nn.GELU() default,
// modules/nclaw/material/utils.py:16-17): gelu = h Phi(h), gelu' = Phi(h) + h phi(h).
// erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, on a par with fp32 erff); it shares
// exp(-h^2/2) with phi.  Branch-free: 2 MUFU (rcp, ex2) + ~10 FMA-pipe instructions.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void normal_cdf_pdf(float h, float& cdf, float& pdf) {
  const float ax = fabsf(h) * 0.70710678118654752f;
  const float t = rcp_approx(fmaf(0.3275911f, ax, 1.0f));
  const float e = ex2_approx((h * h) * -0.72134752044448170f);  // exp(-h^2 / 2)
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(t, p, 1.421413741f);
  p = fmaf(t, p, -0.284496736f);
  p = fmaf(t, p, 0.254829592f);
  const float q = 0.5f * (p * t) * e;               // 0.5 (1 - erf|x|)
  cdf = (h < 0.0f) ? q : 1.0f - q;
  pdf = 0.39894228040143268f * e;
}

// ---- pipeline <-> MMA warp handshakes
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t mbar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(done)
      : "r"(mbar), "r"(parity)
      : "memory");
  return done != 0;
}
struct PipeSync {
  uint32_t done, wfull, wdone;  // shared-memory addresses of the pipeline's mbarriers
  uint32_t bar_id;              // named barrier of the pipeline (1 + p)
  uint32_t ph_done, ph_wdone;
  uint32_t nthreads;            // threads of the pipeline (128 RT)
  // my smem operand writes / TMEM accesses of this phase are finished; returns once every
  // thread of the pipeline got here
  __device__ __forceinline__ void publish() const {
    umma::fence_async_smem();
    umma::fence_before_sync();
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
  }
  // non-blocking hand-over to a dedicated issuing warp polling wfull
  __device__ __forceinline__ void publish_async() const {
    umma::fence_async_smem();
    umma::fence_before_sync();
    mbar_arrive(wfull);
  }
  __device__ __forceinline__ void wait_done() {
    umma::mbar_wait(done, ph_done);
    ph_done ^= 1u;
    umma::fence_after_sync();
  }
  __device__ __forceinline__ void wait_wdone() {
    umma::mbar_wait(wdone, ph_wdone);
    ph_wdone ^= 1u;
    umma::fence_after_sync();
  }
};

// GELU phase of one particle row: read the 64 accumulator columns, write gelu() as the next
// GEMM's A operand and, in the backward, put gelu'() back into the same TMEM columns.
// RT threads share a row: this one owns the columns [half * 64 / RT, (half + 1) * 64 / RT).
template <int NS, bool KEEP_GRAD, int RT>
__device__ __forceinline__ void gelu_phase(uint32_t t_acc, uint8_t* img, int r, int half) {
#pragma unroll 1
  for (int q0 = 0; q0 < 4 / RT; ++q0) {
    const int q = half * (4 / RT) + q0;
    float h[16], a[16];
    umma::tmem_ld16(t_acc + 16 * q, h);
    umma::tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float cdf, pdf;
      normal_cdf_pdf(h[i], cdf, pdf);
      a[i] = h[i] * cdf;
      if (KEEP_GRAD) h[i] = fmaf(h[i], pdf, cdf);
    }
    store_chunk<NS>(img, BIG_SPLIT, 2 * q, r, a);
    store_chunk<NS>(img, BIG_SPLIT, 2 * q + 1, r, a + 8);
    if (KEEP_GRAD) umma::tmem_st16(t_acc + 16 * q, h);
  }
  if (KEEP_GRAD) umma::tmem_wait_st();
}
// dL/dh = dL/da * gelu'(h): both in TMEM; the result becomes an A operand tile at `img`
// once the weight-gradient GEMM that still reads the tile's previous content is done.
template <int NS, int RT>
__device__ __forceinline__ void grad_phase(uint32_t t_acc, uint32_t t_dgelu, uint8_t* img, int r, int half,
                                           PipeSync& ps, bool wait_w) {
  if (wait_w) ps.wait_wdone();
#pragma unroll 1
  for (int q0 = 0; q0 < 4 / RT; ++q0) {
    const int q = half * (4 / RT) + q0;
    float g[16], d[16];
    umma::tmem_ld16(t_acc + 16 * q, g);
    umma::tmem_ld16(t_dgelu + 16 * q, d);
    umma::tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) g[i] *= d[i];
    store_chunk<NS>(img, BIG_SPLIT, 2 * q, r, g);
    store_chunk<NS>(img, BIG_SPLIT, 2 * q + 1, r, g + 8);
  }
}

__device__ __forceinline__ Mat3 head_to_sym(const float (&o)[16]) {
  Mat3 X;  // x = 0.5 (x^T + x)
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) X(i, j) = 0.5f * (o[3 * i + j] + o[3 * j + i]);
  return X;
}

// ---- block prologue: barriers, TMEM, weight image
struct Bars {
  uint64_t w;  // weight image landed (cp.async.bulk complete_tx)
  uint64_t done[4], wfull[4], wdone[4];
};
template <int NPIPE>
__device__ __forceinline__ uint32_t block_setup(Bars* bars, uint32_t* tmem_slot, uint8_t* sm_w,
                                                const uint8_t* __restrict__ staged, uint32_t w_bytes,
                                                uint32_t tmem_cols, uint32_t pipe_threads) {
  const int warp = threadIdx.x >> 5;
  // no early griddepcontrol.launch_dependents here: the kernel that follows a persistent constitutive kernel
  // would sit resident next to it for its whole duration (measured: +3 % on the backward graph, profiles/r01s);
  // the small kernels DO trigger early, so that THIS kernel's CTAs run the prologue below on the SMs the
  // previous kernel has already left
  if (threadIdx.x == 0) {
    umma::mbar_init(umma::smem_u32(&bars->w), 1);
#pragma unroll
    for (int p = 0; p < NPIPE; ++p) {
      umma::mbar_init(umma::smem_u32(&bars->done[p]), 1);
      umma::mbar_init(umma::smem_u32(&bars->wfull[p]), pipe_threads);
      umma::mbar_init(umma::smem_u32(&bars->wdone[p]), 1);
    }
    umma::fence_mbar_init();
  }
  if (warp == 0) umma::tmem_alloc(umma::smem_u32(tmem_slot), tmem_cols);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  // everything above overlapped the previous kernel's tail; from here on we read its outputs
  pdl_wait();
  if (threadIdx.x == 0) {
    const uint32_t bar = umma::smem_u32(&bars->w);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(w_bytes) : "memory");
    constexpr uint32_t CH = 12288;
    for (uint32_t o = 0; o < w_bytes; o += CH) {
      const uint32_t sz = (w_bytes - o < CH) ? (w_bytes - o) : CH;
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
              umma::smem_u32(sm_w + o)),
          "l"(staged + o), "r"(sz), "r"(bar)
          : "memory");
    }
  }
  return *tmem_slot;
}

// ---- MPM work fused into the epilogues (the particle row is already in this thread's hands) -----
// forward, elasticity: P2G of the particle right after its stress is known -- the fire-and-forget
// vector reductions of one pipeline's tile overlap the SVD / GELU work of the other pipelines
// (a stand-alone P2G kernel is bound by the LSU / L2 atomic rate with idle ALUs, this kernel by
// the ALUs with an idle LSU); grid == NULL: not fused.
struct FusedP2G {
  neuma_mpm_const c;
  const float *x, *v, *C, *vol, *rho;
  const int32_t *enabled, *box;
  float4* grid;
};
// backward, plasticity: the adjoint scatter of G2P (mpm_rollout.cu::bwd_scatter_adj_kernel) right
// after dL/dF_trial is known; adj == NULL: not fused.
struct FusedScatter {
  neuma_mpm_const c;
  const float *x, *F, *clip, *v2, *gx2, *gv2, *gC2;
  const int32_t *enabled, *box;
  float4* adj;
};

// forward, plasticity: G2P of the particle (mpm.py:432-498) right before its return mapping -- the
// thread gathers its 27 nodes, writes x', v', C', F_trial (the backward needs them) and goes on
// with F_trial; the gather's memory latency hides under the other pipelines' arithmetic.
// grid == NULL: not fused.
struct FusedG2P {
  neuma_mpm_const c;
  const float *x, *F, *clip;
  const int32_t *enabled, *box;
  const float4* grid;
  float *x2, *v2, *C2, *F2;
  int32_t* box_next;
};
// backward, elasticity: the adjoint gather of the MPM substep (mpm_rollout.cu::bwd_gather_v_kernel)
// right before the elasticity backward that consumes its dL/dstress; grid == NULL: not fused.
struct FusedGather {
  neuma_mpm_const c;
  const float *x, *v, *C, *F, *stress, *vol, *rho, *clip, *v2, *C2, *gx2, *gv2, *gC2, *gF2;
  const int32_t *enabled, *box;
  const float4 *grid, *adj;
  float *gx, *gv, *gC, *gstress;
};

// node box of the next substep + Z-slab drift counter, from inside a tile pipeline (warp reductions
// and global atomics; the stand-alone G2P kernel reduces through shared memory first)
__device__ __forceinline__ void box_accumulate_warp(const neuma_mpm_const& c, bool active, const float px[3],
                                                    const int32_t* __restrict__ box_cur, int32_t* __restrict__ box) {
  int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
  if (active) {
    const int mn[3] = {0, 0, c.z_off}, mx[3] = {c.num_grids - 1, c.num_grids - 1, c.z_off + c.nz - 1};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int b = base_node(px[a], c.inv_dx);
      lo[a] = max(b, mn[a]);
      hi[a] = min(b + 2, mx[a]);
    }
    const int bz = base_node(px[2], c.inv_dx);
    const bool below = bz < c.z_off;
    const bool above = bz + 2 > c.z_off + c.nz - 1 && c.z_off + c.nz < c.num_grids;
    if (below || above) atomicAdd(&box[7], 1);
    // neuma_mpm_grid_op_next cleared the next slot over THIS box dilated by one node
    if (box_cur) {
      bool out = false;
#pragma unroll
      for (int a = 0; a < 3; ++a) out |= lo[a] < box_cur[a] - 1 || hi[a] > box_cur[3 + a] + 1;
      if (out) box[6] = 1;
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int l = __reduce_min_sync(0xffffffffu, lo[a]);
    const int h = __reduce_max_sync(0xffffffffu, hi[a]);
    if ((threadIdx.x & 31) == 0 && l <= h) {
      atomicMin(&box[a], l);
      atomicMax(&box[3 + a], h);
    }
  }
}

// ------------------------------------------------------------------ forward
template <int RT>
__global__ void __launch_bounds__(fwd_threads<RT>(), 1)
constitutive_fwd_kernel(int kind, int n, const float* __restrict__ F, const uint8_t* __restrict__ staged,
                        float alpha, float* __restrict__ out, const FusedP2G fp, const FusedG2P fg,
                        float* __restrict__ vsave) {
  constexpr int NS = NS_FWD, NPIPE = NPIPE_FWD, PW = 4 * RT;  // PW = warps per pipeline
  using W = WImg<NS>;
  using P = PipeFwd<NS>;
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ Bars bars;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t tmem = block_setup<NPIPE>(&bars, &tmem_slot, sm, staged + STAGED_OFF_FWD, W::BYTES, FWD_TMEM_COLS,
                                           TP * RT);
  const uint32_t smem = umma::smem_u32(sm);
  const int tiles = (n + TP - 1) / TP;
  const int chains = gridDim.x * NPIPE;
  constexpr uint32_t ID64 = umma::idesc_bf16(TP, H, 0, 0), ID16 = umma::idesc_bf16(TP, KF, 0, 0);

  if (warp == NPIPE * PW) {
    // ---------------- MMA warp (forward kernel: one dedicated issuing warp serves the pipelines'
    // wfull[p] requests in arrival order; measured ~8% faster here than per-pipeline issue)
    const Desc w0 = desc_kmajor(smem + W::OFF_W0, W0_SPLIT, H), w1 = desc_kmajor(smem + W::OFF_W1, W1_SPLIT, H),
               w2 = desc_kmajor(smem + W::OFF_W2, W2_SPLIT, KF);
    int njobs[NPIPE], job[NPIPE];
    uint32_t ph[NPIPE];
    int remaining = 0;
#pragma unroll
    for (int p = 0; p < NPIPE; ++p) {
      const int c = p * gridDim.x + blockIdx.x;
      njobs[p] = (c < tiles) ? 3 * ((tiles - 1 - c) / chains + 1) : 0;
      job[p] = 0;
      ph[p] = 0;
      remaining += njobs[p];
    }
    umma::mbar_wait(umma::smem_u32(&bars.w), 0);
    while (remaining > 0) {
#pragma unroll
      for (int p = 0; p < NPIPE; ++p) {
        if (job[p] >= njobs[p]) continue;
        uint32_t ok = mbar_test(umma::smem_u32(&bars.wfull[p]), ph[p]) ? 1u : 0u;
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (!ok) continue;
        ph[p] ^= 1u;
        umma::fence_after_sync();
        const uint32_t pbase = smem + W::BYTES + p * P::BYTES;
        const Desc feat = desc_kmajor(pbase + P::OFF_FEAT, SMALL_SPLIT, TP);
        const Desc act = desc_kmajor(pbase + P::OFF_BIGA, BIG_SPLIT, TP);
        const uint32_t acc = tmem + p * FWD_PIPE_COLS;
        const int j = job[p] % 3;
        if (j == 0) gemm_split<NS, 1>(acc, feat, w0, ID64, false);
        else if (j == 1) gemm_split<NS, 4>(acc, act, w1, ID64, false);
        else gemm_split<NS, 4>(acc, act, w2, ID16, false);
        umma::commit(umma::smem_u32(&bars.done[p]));
        ++job[p];
        --remaining;
      }
    }
  } else {
    // ---------------- compute threads: pipeline p, particle row r, column half `half`
    const int p = warp / PW, w8 = warp % PW, quarter = w8 & 3, half = w8 >> 2;
    const int r = quarter * 32 + (tid & 31);
    const bool owner = half == 0;
    uint8_t* const pm = sm + W::BYTES + p * P::BYTES;
    const uint32_t tacc = tmem + ((uint32_t)(quarter * 32) << 16) + p * FWD_PIPE_COLS;
    PipeSync ps{umma::smem_u32(&bars.done[p]), umma::smem_u32(&bars.wfull[p]), 0u, (uint32_t)(1 + p), 0u, 0u,
                (uint32_t)(TP * RT)};
    for (int tile = p * gridDim.x + blockIdx.x; tile < tiles; tile += chains) {
      const int pi = tile * TP + r;
      Mat3 kF, kR;
      if (owner) {
        Mat3 Fin;
        bool have_F = false;
        if (fg.grid) {   // fused G2P: F (= F_trial of this substep) is produced here, not read
          const bool act = pi < n && fg.enabled[pi] != 0;
          float nx3[3] = {0.f, 0.f, 0.f};
          if (act) {
            g2p_particle<false>(fg.c, grid_map(fg.c, fg.box), pi, fg.x, fg.F, fg.clip, fg.grid, fg.x2, fg.v2, fg.C2,
                                fg.F2, nx3, Fin.m);
            have_F = true;
          }
          if (fg.box_next) box_accumulate_warp(fg.c, act, nx3, fg.box, fg.box_next);
        }
        if (!have_F) Fin = load_F_or_identity(F, pi, n);
        Kinematics k;
        kinematics(Fin, kind != 2, k);
        if (vsave && pi < n) mat3_store(vsave + (size_t)pi * 9, k.V);
        kF = k.F;
        kR = k.R;
        float f[KF];
#pragma unroll
        for (int i = 0; i < KF; ++i) f[i] = (i < NF) ? k.feat[i] : 0.f;
        store_chunk<NS>(pm + P::OFF_FEAT, SMALL_SPLIT, 0, r, f);
        store_chunk<NS>(pm + P::OFF_FEAT, SMALL_SPLIT, 1, r, f + 8);
      }
      ps.publish_async();                                            // -> h1 = feat W0^T
      if (owner && !fg.grid) prefetch_row(F, pi + chains * TP, n);
      ps.wait_done();
      gelu_phase<NS, false, RT>(tacc, pm + P::OFF_BIGA, r, half);
      ps.publish_async();                                            // -> h2 = a1 W1^T
      ps.wait_done();
      gelu_phase<NS, false, RT>(tacc, pm + P::OFF_BIGA, r, half);
      ps.publish_async();                                            // -> o = a2 W2^T
      ps.wait_done();
      if (owner) {
        float o[16];
        umma::tmem_ld16(tacc, o);
        umma::tmem_wait_ld();
        if (pi < n) {
          const Mat3 X = head_to_sym(o);
          const Mat3 Pm = mat3_mul(kR, X);
          Mat3 O;
          if (kind == 1) {
#pragma unroll
            for (int i = 0; i < 9; ++i) O.m[i] = alpha * Pm.m[i] + kF.m[i];  // delta_Fp + F
          } else {
            O = mat3_mul_nt(Pm, kF);  // cauchy = P F^T
          }
          mat3_store(out + (size_t)pi * 9, O);
          if (fp.grid && fp.enabled[pi] != 0 && particle_selected(fp.c, fp.x[3 * pi + 2])) {
            check_interior_class(fp.c, fp.x[3 * pi + 2], fp.box);
            p2g_particle(fp.c, grid_map(fp.c, fp.box), pi, fp.x, fp.v, fp.C, O.m, fp.vol[pi], fp.rho[pi], fp.grid);
          }
        }
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, FWD_TMEM_COLS);
}

// ------------------------------------------------------------------ backward
template <bool WGRAD, int RT>
__global__ void __launch_bounds__(bwd_threads<RT>(), 1)
constitutive_bwd_kernel(int kind, int n, const float* __restrict__ F, const uint8_t* __restrict__ staged,
                        float alpha, const float* __restrict__ gout, float* __restrict__ gF,
                        float* __restrict__ gW0, float* __restrict__ gW1, float* __restrict__ gW2,
                        float* __restrict__ gWpart, int accumulate_gF, const FusedScatter fs, const FusedGather fa,
                        const float* __restrict__ vin) {
  constexpr int NS = NS_BWD, NPIPE = NPIPE_BWD, PW = 4 * RT;  // PW = warps per pipeline
  using W = WImg<NS>;
  using P = PipeBwd<NS>;
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ Bars bars;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
#ifdef NEUMA_PHASE_TIMING
  long long kt__ = clock64();
  if (threadIdx.x == 0) {
    unsigned long long ns; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    atomicMin((unsigned long long*)&g_phase_cycles[24], ns);
    if (blockIdx.x == 0) g_phase_cycles[26] = clock64();
  }
#define KPHASE(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { const long long t__ = clock64(); g_phase_cycles[i] += t__ - kt__; kt__ = t__; } } while (0)
#else
#define KPHASE(i)
#endif
  const uint32_t tmem = block_setup<NPIPE>(&bars, &tmem_slot, sm, staged + STAGED_OFF_BWD, W::BYTES, BWD_TMEM_COLS,
                                           TP * RT);
  KPHASE(16);
  const uint32_t smem = umma::smem_u32(sm);
  const int tiles = (n + TP - 1) / TP;
  const int chains = gridDim.x * NPIPE;
  constexpr uint32_t ID64 = umma::idesc_bf16(TP, H, 0, 0), ID16 = umma::idesc_bf16(TP, KF, 0, 0);
  constexpr uint32_t ID64_BT = umma::idesc_bf16(TP, H, 0, 1), ID16_BT = umma::idesc_bf16(TP, KF, 0, 1);
  constexpr uint32_t IDW64 = umma::idesc_bf16(64, H, 1, 1), IDW16 = umma::idesc_bf16(64, KF, 1, 1);

  if (warp == NPIPE * PW) {
    // ---------------- weight-gradient warp: serves wfull[p] in arrival order
    if (WGRAD) {
      int njobs[NPIPE], job[NPIPE];
      uint32_t ph[NPIPE];
      int remaining = 0;
#pragma unroll
      for (int p = 0; p < NPIPE; ++p) {
        const int c = p * gridDim.x + blockIdx.x;
        njobs[p] = (c < tiles) ? 3 * ((tiles - 1 - c) / chains + 1) : 0;
        job[p] = 0;
        ph[p] = 0;
        remaining += njobs[p];
      }
      bool acc2 = false, acc1 = false, acc0 = false;  // accumulators initialised?
      while (remaining > 0) {
#pragma unroll
        for (int p = 0; p < NPIPE; ++p) {
          if (job[p] >= njobs[p]) continue;
          uint32_t ok = mbar_test(umma::smem_u32(&bars.wfull[p]), ph[p]) ? 1u : 0u;
          ok = __shfl_sync(0xffffffffu, ok, 0);
          if (!ok) continue;
          ph[p] ^= 1u;
          umma::fence_after_sync();
          const uint32_t pbase = smem + W::BYTES + p * P::BYTES;
          const Desc featT = desc_mnmajor(pbase + P::OFF_FEAT, SMALL_SPLIT, TP), goT = desc_mnmajor(pbase + P::OFF_GO, SMALL_SPLIT, TP),
                     bigaT = desc_mnmajor(pbase + P::OFF_BIGA, BIG_SPLIT, TP), bigbT = desc_mnmajor(pbase + P::OFF_BIGB, BIG_SPLIT, TP);
          const int j = job[p] % 3;
          if (j == 0) {         // dW2^T += a2^T go
            gemm_split<NS, 8>(tmem + COL_GW2T, bigbT, goT, IDW16, acc2);
            acc2 = true;
          } else if (j == 1) {  // dW1 += g2^T a1
            gemm_split<NS, 8>(tmem + COL_GW1, bigbT, bigaT, IDW64, acc1);
            acc1 = true;
          } else {              // dW0 += g1^T feat
            gemm_split<NS, 8>(tmem + COL_GW0, bigaT, featT, IDW16, acc0);
            acc0 = true;
          }
          umma::commit(umma::smem_u32(&bars.wdone[p]));
          ++job[p];
          --remaining;
        }
      }
    }
  } else {
    // ---------------- pipeline p, particle row r, column half `half`; one owner warp issues its GEMMs
    const int p = warp / PW, w8 = warp % PW, quarter = w8 & 3, half = w8 >> 2;
    const int r = quarter * 32 + (tid & 31);
    const bool owner = half == 0;
    const bool issuer = owner && quarter == (p & 3);  // issuers of different pipelines sit on different SM sub-partitions
    uint8_t* const pm = sm + W::BYTES + p * P::BYTES;
    const uint32_t pbase = smem + W::BYTES + p * P::BYTES;
    const uint32_t tp = tmem + p * BWD_PIPE_COLS;
    const uint32_t tl = tp + ((uint32_t)(quarter * 32) << 16);
    // RT = 2: F, U, V, sigma of row r parked in shared memory, [element][row]
    float* const stash = reinterpret_cast<float*>(sm + W::BYTES + NPIPE * P::BYTES + p * STASH_BYTES) + r;
    const uint32_t dbar = umma::smem_u32(&bars.done[p]), wfull = umma::smem_u32(&bars.wfull[p]);
    PipeSync ps{dbar, wfull, umma::smem_u32(&bars.wdone[p]), (uint32_t)(1 + p), 0u, 0u, (uint32_t)(TP * RT)};
    if (issuer) umma::mbar_wait(umma::smem_u32(&bars.w), 0);  // weight image has landed
    // operand descriptors are rebuilt where they are used (issuer warp only): keeping ten of them
    // live costs every thread 40 registers
    auto wdesc = [&](int which, bool transposed) {
      const uint32_t off = which == 0 ? W::OFF_W0 : (which == 1 ? W::OFF_W1 : W::OFF_W2);
      const uint32_t split = which == 0 ? W0_SPLIT : (which == 1 ? W1_SPLIT : W2_SPLIT);
      const int R = which == 2 ? KF : H;
      return transposed ? desc_mnmajor(smem + off, split, R) : desc_kmajor(smem + off, split, R);
    };
    bool first = true;
    PHASE_T0();
    for (int tile = p * gridDim.x + blockIdx.x; tile < tiles; tile += chains, first = false) {
      // LATE trigger of the dependent kernel (see common.cuh): when this CTA's first pipeline starts its last
      // tile -- the small kernel that follows becomes resident during the tail, not during the whole kernel.
      // Backward only: measured -0.04 ms per frame here, +0.025 ms in the forward kernel (profiles/r03y_pdl_ab.txt)
      if (tid == 0 && tile + chains >= tiles) pdl_launch_dependents();
      const int pi = tile * TP + r;
      const bool live = pi < n;
      if (owner && fa.grid && live) {
        // fused adjoint gather: writes dL/dx, dL/dv, dL/dC, the MPM part of dL/dF (into gF, which this
        // kernel then accumulates onto) and dL/dstress (= this kernel's gout, read back below)
        if (fa.enabled[pi] == 0)
          zero_particle_grads(pi, fa.gx, fa.gv, fa.gC, gF, fa.gstress);
        else
          bwd_gather_particle<false, true>(fa.c, grid_map(fa.c, fa.box), pi, fa.x, fa.v, fa.C, fa.F, fa.stress, fa.vol,
                                           fa.rho, fa.clip, fa.v2, fa.C2, fa.gx2, fa.gv2, fa.gC2, fa.gF2, fa.grid,
                                           fa.adj, fa.gx, fa.gv, fa.gC, gF, fa.gstress);
      }
      Kinematics k;
      if (owner) {
        // vin: warp-uniform choice (svd3_fast votes across the warp); padding rows take V = I
        if (vin) kinematics_from_v(load_F_or_identity(F, pi, n), live ? vin + (size_t)pi * 9 : nullptr, kind != 2, k);
        else kinematics(load_F_or_identity(F, pi, n), kind != 2, k);
      }
      PHASE(0);
      // FEAT / BIGA were last read by the previous tile's dW0 GEMM
      if (WGRAD && !first) ps.wait_wdone();
      PHASE(15);
      if (owner) {
        float f[KF];
#pragma unroll
        for (int i = 0; i < KF; ++i) f[i] = (i < NF && live) ? k.feat[i] : 0.f;
        store_chunk<NS>(pm + P::OFF_FEAT, SMALL_SPLIT, 0, r, f);
        store_chunk<NS>(pm + P::OFF_FEAT, SMALL_SPLIT, 1, r, f + 8);
        if (RT == 2) {
#pragma unroll
          for (int i = 0; i < 9; ++i) {
            stash[i * TP] = k.F.m[i];
            stash[(9 + i) * TP] = k.U.m[i];
            stash[(18 + i) * TP] = k.V.m[i];
          }
#pragma unroll
          for (int i = 0; i < 3; ++i) stash[(27 + i) * TP] = k.sig[i];
        }
      }
      ps.publish();
      if (issuer) {                                              // h1 = feat W0^T
        umma::fence_after_sync();
        gemm_split<NS, 1>(tp + COL_H1, desc_kmajor(pbase + P::OFF_FEAT, SMALL_SPLIT, TP), wdesc(0, false), ID64, false);
        umma::commit(dbar);
      }
      if (owner) {
        if (!fa.grid) prefetch_row(gout, pi, n);
        prefetch_row(F, pi + chains * TP, n);
        if (vin) prefetch_row(vin, pi + chains * TP, n);
      }
      PHASE(1);
      ps.wait_done();
      PHASE(2);
      gelu_phase<NS, true, RT>(tl + COL_H1, pm + P::OFF_BIGA, r, half);    // a1; gelu'(h1) stays in TMEM
      ps.publish();
      if (issuer) {                                              // h2 = a1 W1^T
        umma::fence_after_sync();
        gemm_split<NS, 4>(tp + COL_H2, desc_kmajor(pbase + P::OFF_BIGA, BIG_SPLIT, TP), wdesc(1, false), ID64, false);
        umma::commit(dbar);
      }
      PHASE(3);
      ps.wait_done();
      PHASE(4);
      gelu_phase<NS, true, RT>(tl + COL_H2, pm + P::OFF_BIGB, r, half);    // a2; gelu'(h2)
      ps.publish();
      if (issuer) {                                              // o = a2 W2^T
        umma::fence_after_sync();
        gemm_split<NS, 4>(tp + COL_S, desc_kmajor(pbase + P::OFF_BIGB, BIG_SPLIT, TP), wdesc(2, false), ID16, false);
        umma::commit(dbar);
      }
      PHASE(5);
      ps.wait_done();
      PHASE(6);
      // ---- output head: dL/dout -> dL/d(MLP output), dL/dR, direct dL/dF
      Mat3 gR, gFd;
      if (owner) {
        float o[16], g9[KF];
        umma::tmem_ld16(tl + COL_S, o);
        umma::tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < KF; ++i) g9[i] = 0.f;
        if (live) {
          Mat3 kF, kR;
          if (RT == 2) {
            Mat3 kU, kV;
#pragma unroll
            for (int i = 0; i < 9; ++i) {
              kF.m[i] = stash[i * TP];
              kU.m[i] = stash[(9 + i) * TP];
              kV.m[i] = stash[(18 + i) * TP];
            }
            kR = mat3_mul_nt(kU, kV);
          } else {
            kF = k.F;
            kR = k.R;
          }
          const Mat3 X = head_to_sym(o);
          const Mat3 Pm = mat3_mul(kR, X);
          Mat3 G;
          if (fa.grid) {  // written by this thread's fused gather above: read it back through L2
#pragma unroll
            for (int i = 0; i < 9; ++i) G.m[i] = __ldcg(fa.gstress + (size_t)pi * 9 + i);
          } else {
            G = mat3_load(gout + (size_t)pi * 9);
          }
          Mat3 gP;  // out = P F^T (elastic)  |  out = F + alpha P (plastic)
          if (kind == 1) {
#pragma unroll
            for (int i = 0; i < 9; ++i) { gP.m[i] = alpha * G.m[i]; gFd.m[i] = G.m[i]; }
          } else {
            gP = mat3_mul(G, kF);
            gFd = mat3_mul_tn(G, Pm);
          }
          gR = mat3_mul(gP, X);  // X symmetric
          const Mat3 gXs = mat3_mul_tn(kR, gP);
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) g9[3 * i + j] = 0.5f * (gXs(i, j) + gXs(j, i));
        }
        store_chunk<NS>(pm + P::OFF_GO, SMALL_SPLIT, 0, r, g9);
        store_chunk<NS>(pm + P::OFF_GO, SMALL_SPLIT, 1, r, g9 + 8);
      }
      ps.publish();
      if (WGRAD) mbar_arrive(wfull);                             // -> dW2^T += a2^T go
      if (issuer) {                                              // dL/da2 = go W2
        umma::fence_after_sync();
        gemm_split<NS, 1>(tp + COL_S, desc_kmajor(pbase + P::OFF_GO, SMALL_SPLIT, TP), wdesc(2, true), ID64_BT, false);
        umma::commit(dbar);
      }
      PHASE(7);
      ps.wait_done();
      PHASE(8);
      grad_phase<NS, RT>(tl + COL_S, tl + COL_H2, pm + P::OFF_BIGB, r, half, ps, WGRAD);   // g2 over a2 (after dW2)
      ps.publish();
      if (WGRAD) mbar_arrive(wfull);                             // -> dW1 += g2^T a1
      if (issuer) {                                              // dL/da1 = g2 W1
        umma::fence_after_sync();
        gemm_split<NS, 4>(tp + COL_S, desc_kmajor(pbase + P::OFF_BIGB, BIG_SPLIT, TP), wdesc(1, true), ID64_BT, false);
        umma::commit(dbar);
      }
      PHASE(9);
      ps.wait_done();
      PHASE(10);
      grad_phase<NS, RT>(tl + COL_S, tl + COL_H1, pm + P::OFF_BIGA, r, half, ps, WGRAD);   // g1 over a1 (after dW1)
      ps.publish();
      if (WGRAD) mbar_arrive(wfull);                             // -> dW0 += g1^T feat
      if (issuer) {                                              // dL/dfeat = g1 W0
        umma::fence_after_sync();
        gemm_split<NS, 4>(tp + COL_S, desc_kmajor(pbase + P::OFF_BIGA, BIG_SPLIT, TP), wdesc(0, true), ID16_BT, false);
        umma::commit(dbar);
      }
      PHASE(11);
      ps.wait_done();
      PHASE(12);
      // ---- invariants -> F
      if (owner) {
        float gfeat[16];
        umma::tmem_ld16(tl + COL_S, gfeat);
        umma::tmem_wait_ld();
        if (live) {
          Mat3 kF, kU, kV;
          float sig[3];
          if (RT == 2) {
#pragma unroll
            for (int i = 0; i < 9; ++i) {
              kF.m[i] = stash[i * TP];
              kU.m[i] = stash[(9 + i) * TP];
              kV.m[i] = stash[(18 + i) * TP];
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) sig[i] = stash[(27 + i) * TP];
          } else {
            kF = k.F; kU = k.U; kV = k.V;
#pragma unroll
            for (int i = 0; i < 3; ++i) sig[i] = k.sig[i];
          }
          Mat3 gFt = gFd;
          {  // FtF = F^T F : dF += F (G + G^T)
            Mat3 Gs;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
              for (int j = 0; j < 3; ++j) Gs(i, j) = gfeat[3 + 3 * i + j] + gfeat[3 + 3 * j + i];
            const Mat3 t = mat3_mul(kF, Gs);
#pragma unroll
            for (int i = 0; i < 9; ++i) gFt.m[i] += t.m[i];
          }
          {  // det F : dF += gdet * cof(F)
            const Mat3& A = kF;
            const float gd = gfeat[12];
            gFt(0, 0) += gd * (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1));
            gFt(0, 1) += gd * (A(1, 2) * A(2, 0) - A(1, 0) * A(2, 2));
            gFt(0, 2) += gd * (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0));
            gFt(1, 0) += gd * (A(0, 2) * A(2, 1) - A(0, 1) * A(2, 2));
            gFt(1, 1) += gd * (A(0, 0) * A(2, 2) - A(0, 2) * A(2, 0));
            gFt(1, 2) += gd * (A(0, 1) * A(2, 0) - A(0, 0) * A(2, 1));
            gFt(2, 0) += gd * (A(0, 1) * A(1, 2) - A(0, 2) * A(1, 1));
            gFt(2, 1) += gd * (A(0, 2) * A(1, 0) - A(0, 0) * A(1, 2));
            gFt(2, 2) += gd * (A(0, 0) * A(1, 1) - A(0, 1) * A(1, 0));
          }
          {  // sigma and R = U V^T : dF += U (diag(gsig) + K) V^T
            const Mat3 M = mat3_mul(mat3_mul_tn(kU, gR), kV);
            Mat3 K;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
              for (int j = 0; j < 3; ++j)
                K(i, j) = (i == j) ? gfeat[i]
                                   : (M(i, j) - M(j, i)) / signed_floor(sig[i] + sig[j], 1e-6f);
            const Mat3 t = mat3_mul_nt(mat3_mul(kU, K), kV);
#pragma unroll
            for (int i = 0; i < 9; ++i) gFt.m[i] += t.m[i];
          }
          if (accumulate_gF) {  // dL/dF += ... (the caller's gF already holds the MPM substep's part)
            const Mat3 prev = mat3_load(gF + (size_t)pi * 9);
#pragma unroll
            for (int i = 0; i < 9; ++i) gFt.m[i] += prev.m[i];
          }
          mat3_store(gF + (size_t)pi * 9, gFt);
          if (fs.adj && fs.enabled[pi] != 0 && particle_selected(fs.c, fs.x[3 * pi + 2])) {
            check_interior_class(fs.c, fs.x[3 * pi + 2], fs.box);
            const G2PAdj g = g2p_adjoint_inputs(fs.c, pi, fs.x, fs.F, fs.clip, fs.v2, fs.gx2, fs.gv2, fs.gC2, gFt.m);
            g2p_adjoint_scatter(fs.c, grid_map(fs.c, fs.box), pi, fs.x, g, fs.adj);
          }
        }
      }
      PHASE(13);
    }
    if (WGRAD && !first) ps.wait_wdone();  // the last tile's dW0 GEMM
    PHASE(14);
  }
  KPHASE(17);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  KPHASE(18);
  if (WGRAD && warp < 4) {
    // every weight-gradient MMA of the CTA has completed (each pipeline waited for its last one).
    // M = 64 accumulators: lane 32*q + i (i < 16) holds row 16*q + i.
    const uint32_t tl0 = tmem + ((uint32_t)(warp * 32) << 16);
    const int i16 = tid & 31, row = warp * 16 + i16;
    float a[16];
    if (gWpart) {
      // this CTA's private slot of the partial-sum buffer: plain read-modify-write, no atomics
      // (148 CTAs x 5504 same-address atomics cost ~25% of the kernel); reduced once per frame
      // by reduce_partials_kernel.
      float* part = gWpart + (size_t)blockIdx.x * GW_FLOATS;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        umma::tmem_ld16(tl0 + COL_GW1 + 16 * c, a);
        umma::tmem_wait_ld();
        if (i16 < 16) {
          float4* q = reinterpret_cast<float4*>(part + row * H + 16 * c);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4 t = q[j];
            t.x += a[4 * j]; t.y += a[4 * j + 1]; t.z += a[4 * j + 2]; t.w += a[4 * j + 3];
            q[j] = t;
          }
        }
      }
      umma::tmem_ld16(tl0 + COL_GW0, a);
      umma::tmem_wait_ld();
      if (i16 < 16)
#pragma unroll
        for (int j = 0; j < NF; ++j) part[GW_OFF0 + row * NF + j] += a[j];
      umma::tmem_ld16(tl0 + COL_GW2T, a);
      umma::tmem_wait_ld();
      if (i16 < 16)
#pragma unroll
        for (int j = 0; j < 9; ++j) part[GW_OFF2 + j * H + row] += a[j];
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        umma::tmem_ld16(tl0 + COL_GW1 + 16 * c, a);
        umma::tmem_wait_ld();
        if (i16 < 16)
#pragma unroll
          for (int j = 0; j < 16; ++j) atomicAdd(gW1 + row * H + 16 * c + j, a[j]);
      }
      umma::tmem_ld16(tl0 + COL_GW0, a);
      umma::tmem_wait_ld();
      if (i16 < 16)
#pragma unroll
        for (int j = 0; j < NF; ++j) atomicAdd(gW0 + row * NF + j, a[j]);
      umma::tmem_ld16(tl0 + COL_GW2T, a);
      umma::tmem_wait_ld();
      if (i16 < 16)
#pragma unroll
        for (int j = 0; j < 9; ++j) atomicAdd(gW2 + j * H + row, a[j]);
    }
  }
  KPHASE(19);
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, BWD_TMEM_COLS);
  KPHASE(20);
#ifdef NEUMA_PHASE_TIMING
  if (threadIdx.x == 0) {
    unsigned long long ns; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    atomicMax((unsigned long long*)&g_phase_cycles[25], ns);
    if (blockIdx.x == 0) { g_phase_cycles[27] = clock64(); g_phase_cycles[28] = (long long)ns; }
  }
#endif
}

// ------------------------------------------------------------------ stand-alone SVD
__global__ void __launch_bounds__(256)
svd3_fwd_kernel(int n, const float* __restrict__ A, float* __restrict__ U,
                float* __restrict__ sigma, float* __restrict__ Vh) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  Mat3 a = mat3_load(A + (size_t)p * 9), u, v;
  float s[3];
  svd3(a, u, s, v);
  // batch_svd's "if det < 0 flip column 2" (svd.py:76-92) is the identity here:
  // svd3 returns det U = det V = +1 by construction.
  mat3_store(U + (size_t)p * 9, u);
  sigma[3 * p] = s[0]; sigma[3 * p + 1] = s[1]; sigma[3 * p + 2] = s[2];
  Mat3 vt;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) vt(i, j) = v(j, i);
  mat3_store(Vh + (size_t)p * 9, vt);
}

// Textbook SVD adjoint, A = U S V^T:
//   gA = U [ (Fm o (U^T gU - gU^T U)) S + S (Fm o (V^T gV - gV^T V)) + diag(gs) ] V^T,
//   Fm_ij = 1 / (s_j^2 - s_i^2).  The denominator is clamped away from zero the way
//   warp-lang's adj_svd3 does, min(s_j^2 - s_i^2, -1e-6) for i < j [UPSTREAM-UNVERIFIED].
__global__ void __launch_bounds__(256)
svd3_bwd_kernel(int n, const float* __restrict__ U, const float* __restrict__ sigma,
                const float* __restrict__ Vh, const float* __restrict__ gU,
                const float* __restrict__ gsigma, const float* __restrict__ gVh,
                float* __restrict__ gA) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const Mat3 u = mat3_load(U + (size_t)p * 9);
  const Mat3 vh = mat3_load(Vh + (size_t)p * 9);
  Mat3 v;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) v(i, j) = vh(j, i);
  const float s[3] = {sigma[3 * p], sigma[3 * p + 1], sigma[3 * p + 2]};
  Mat3 Mu, Mv;
#pragma unroll
  for (int i = 0; i < 9; ++i) { Mu.m[i] = 0.f; Mv.m[i] = 0.f; }
  if (gU) Mu = mat3_mul_tn(u, mat3_load(gU + (size_t)p * 9));
  if (gVh) {
    const Mat3 gvh = mat3_load(gVh + (size_t)p * 9);
    Mat3 gv;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) gv(i, j) = gvh(j, i);
    Mv = mat3_mul_tn(v, gv);
  }
  Mat3 inner;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (i == j) {
        inner(i, j) = gsigma ? gsigma[3 * p + i] : 0.f;
      } else {
        const int lo = i < j ? i : j, hi = i < j ? j : i;
        const float d = fminf(s[hi] * s[hi] - s[lo] * s[lo], -1e-6f);  // (lo,hi) entry
        const float f = (i < j) ? 1.0f / d : -1.0f / d;
        inner(i, j) = f * ((Mu(i, j) - Mu(j, i)) * s[j] + s[i] * (Mv(i, j) - Mv(j, i)));
      }
    }
  mat3_store(gA + (size_t)p * 9, mat3_mul(mat3_mul(u, inner), vh));
}

// gW += sum over the CTA slots of the partial buffer; the slots are zeroed for the next frame.
// Block = 32 consecutive elements x 32 slot lanes: every thread sums <= 5 slots (independent loads
// in flight), shared-memory tree at the end (the round-1 shape, 8 slot lanes, took 52 us for 3.3 MB).
__global__ void __launch_bounds__(1024)
reduce_partials_kernel(int slots, float* __restrict__ part, float* __restrict__ gW0,
                       float* __restrict__ gW1, float* __restrict__ gW2) {
  __shared__ float red[32][33];
  const int lane = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + lane;
  float acc = 0.f;
  if (e < GW_FLOATS) {
#pragma unroll 5
    for (int sidx = sl; sidx < slots; sidx += 32) {
      float* q = part + (size_t)sidx * GW_FLOATS + e;
      acc += *q;
      *q = 0.f;
    }
  }
  red[sl][lane] = acc;
  __syncthreads();
  if (sl == 0 && e < GW_FLOATS) {
#pragma unroll
    for (int i = 1; i < 32; ++i) acc += red[i][lane];
    if (e < GW_OFF0) gW1[e] += acc;
    else if (e < GW_OFF2) gW0[e - GW_OFF0] += acc;
    else gW2[e - GW_OFF2] += acc;
  }
}

// threads per particle row of the fused kernels: NEUMA_ROW_THREADS=1|2.  Default 1: measured on the
// B200 (profiles/r02i_phase_timing.txt) the GELU phases take the same 3.9 k cycles per tile with one
// or two threads per row -- they are bound by the XU pipe (rcp + ex2 + cvt.bf16x2 per value, quarter
// rate) and the FMA pipe together, not by per-warp latency -- and the backward is 8 % slower with
// two (register spills at 96 regs/thread).  Kept as a measured A/B.
int row_threads() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("NEUMA_ROW_THREADS");
    v = (e && e[0] == '2') ? 2 : 1;
  }
  return v;
}

int set_attrs(neuma_ctx* ctx) {
  if (ctx->attrs_set) return 0;
  int cur = -1;
  cudaGetDevice(&cur);
  if (cur != ctx->device) NEUMA_CUDA(ctx, cudaSetDevice(ctx->device));
  const auto attr = cudaFuncAttributeMaxDynamicSharedMemorySize;
  NEUMA_CUDA(ctx, cudaFuncSetAttribute(constitutive_fwd_kernel<1>, attr, (int)FWD_SMEM));
  NEUMA_CUDA(ctx, cudaFuncSetAttribute(constitutive_fwd_kernel<2>, attr, (int)FWD_SMEM));
  NEUMA_CUDA(ctx, cudaFuncSetAttribute(constitutive_bwd_kernel<true, 1>, attr, (int)bwd_smem<1>()));
  NEUMA_CUDA(ctx, cudaFuncSetAttribute(constitutive_bwd_kernel<false, 1>, attr, (int)bwd_smem<1>()));
  NEUMA_CUDA(ctx, cudaFuncSetAttribute(constitutive_bwd_kernel<true, 2>, attr, (int)bwd_smem<2>()));
  NEUMA_CUDA(ctx, cudaFuncSetAttribute(constitutive_bwd_kernel<false, 2>, attr, (int)bwd_smem<2>()));
  ctx->attrs_set = 1;
  if (cur >= 0 && cur != ctx->device) cudaSetDevice(cur);
  return 0;
}

// one place that picks the instantiation
int launch_fwd(neuma_ctx* ctx, int kind, int n, const float* F, const void* staged, float alpha, float* out,
               const FusedP2G& fp, const FusedG2P& fg, cudaStream_t st) {
  const int tiles = (n + TP - 1) / TP;
  const int grid = tiles < ctx->num_sms ? tiles : ctx->num_sms;
  float* vsave = ctx->svd_v_out;       // neuma_constitutive_svd_cache: consumed by this launch (either pointer)
  ctx->svd_v_out = nullptr;
  ctx->svd_v_in = nullptr;
  if (row_threads() == 2)
    NEUMA_CUDA(ctx, neuma_launch(constitutive_fwd_kernel<2>, dim3(grid), dim3(fwd_threads<2>()), FWD_SMEM, st, kind, n,
                                 F, (const uint8_t*)staged, alpha, out, fp, fg, vsave));
  else
    NEUMA_CUDA(ctx, neuma_launch(constitutive_fwd_kernel<1>, dim3(grid), dim3(fwd_threads<1>()), FWD_SMEM, st, kind, n,
                                 F, (const uint8_t*)staged, alpha, out, fp, fg, vsave));
  NEUMA_LAUNCH_CHECK(ctx, "constitutive_fwd_kernel");
  return 0;
}

int launch_bwd(neuma_ctx* ctx, bool wgrad, int kind, int n, const float* F, const void* staged, float alpha,
               const float* gout, float* gF, float* gW0, float* gW1, float* gW2, float* gWpart, int accumulate_gF,
               const FusedScatter& fs, const FusedGather& fa, cudaStream_t st) {
  const int tiles = (n + TP - 1) / TP;
  const int grid = tiles < ctx->num_sms ? tiles : ctx->num_sms;
  const uint8_t* sg = (const uint8_t*)staged;
  const float* vin = ctx->svd_v_in;    // neuma_constitutive_svd_cache: consumed by this launch (either pointer)
  ctx->svd_v_in = nullptr;
  ctx->svd_v_out = nullptr;
#define NEUMA_BWD(WG, RTV)                                                                                       \
  NEUMA_CUDA(ctx, neuma_launch(constitutive_bwd_kernel<WG, RTV>, dim3(grid), dim3(bwd_threads<RTV>()),            \
                               bwd_smem<RTV>(), st, kind, n, F, sg, alpha, gout, gF, gW0, gW1, gW2, gWpart,       \
                               accumulate_gF, fs, fa, vin))
  if (row_threads() == 2) {
    if (wgrad) NEUMA_BWD(true, 2); else NEUMA_BWD(false, 2);
  } else {
    if (wgrad) NEUMA_BWD(true, 1); else NEUMA_BWD(false, 1);
  }
#undef NEUMA_BWD
  NEUMA_LAUNCH_CHECK(ctx, "constitutive_bwd_kernel");
  return 0;
}

int check_args(neuma_ctx* ctx, int kind, int hidden) {
  if (!ctx) return 2;
  if (hidden != H) return neuma_fail_msg(ctx, "neuma_constitutive: only hidden width 64 is built");
  if (kind < 0 || kind > 2) return neuma_fail_msg(ctx, "neuma_constitutive: kind must be 0, 1 or 2");
  return 0;
}

}  // namespace

extern "C" {

int neuma_svd3_fwd(neuma_ctx* ctx, int n, const float* A, float* U, float* sigma, float* Vh,
                   void* stream) {
  if (!ctx) return 2;
  if (n <= 0) return 0;
  svd3_fwd_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, A, U, sigma, Vh);
  NEUMA_LAUNCH_CHECK(ctx, "svd3_fwd_kernel");
  return 0;
}

int neuma_svd3_bwd(neuma_ctx* ctx, int n, const float* U, const float* sigma, const float* Vh,
                   const float* gU, const float* gsigma, const float* gVh, float* gA,
                   void* stream) {
  if (!ctx) return 2;
  if (n <= 0) return 0;
  svd3_bwd_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, U, sigma, Vh, gU, gsigma,
                                                                      gVh, gA);
  NEUMA_LAUNCH_CHECK(ctx, "svd3_bwd_kernel");
  return 0;
}

int neuma_constitutive_staged_bytes(void) { return (int)STAGED_BYTES; }

int neuma_constitutive_svd_cache(neuma_ctx* ctx, float* v_out, const float* v_in) {
  if (!ctx) return 2;
  ctx->svd_v_out = v_out;
  ctx->svd_v_in = v_in;
  return 0;
}

#ifdef NEUMA_PHASE_TIMING
int neuma_debug_phase_cycles(long long* out_host, int reset) {
  cudaDeviceSynchronize();
  if (out_host) cudaMemcpyFromSymbol(out_host, g_phase_cycles, sizeof(g_phase_cycles));
  if (reset) {
    long long z[64] = {0};
    z[24] = 0x7fffffffffffffffLL;
    cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z));
  }
  return 0;
}
#endif

int neuma_constitutive_stage(neuma_ctx* ctx, int hidden, const float* W0, const float* W1,
                             const float* W2, void* staged, void* stream) {
  if (int r = check_args(ctx, 0, hidden)) return r;
  if (!W0 || !W1 || !W2 || !staged) return neuma_fail_msg(ctx, "neuma_constitutive_stage: NULL argument");
  constexpr int total = H * KF + H * H + KF * H;
  stage_weights_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(W0, W1, W2, (uint8_t*)staged);
  NEUMA_LAUNCH_CHECK(ctx, "stage_weights_kernel");
  return 0;
}

int neuma_constitutive_fwd_staged(neuma_ctx* ctx, int kind, int n, int hidden, const float* F,
                                  const void* staged, float alpha, float* out, void* stream) {
  if (int r = check_args(ctx, kind, hidden)) return r;
  if (n <= 0) return 0;
  if (int r = set_attrs(ctx)) return r;
  FusedP2G none;
  memset(&none, 0, sizeof(none));
  FusedG2P nog;
  memset(&nog, 0, sizeof(nog));
  return launch_fwd(ctx, kind, n, F, staged, alpha, out, none, nog, (cudaStream_t)stream);
}

int neuma_constitutive_fwd_p2g(neuma_ctx* ctx, int kind, int n, int hidden, const float* F, const void* staged,
                               float* stress, const neuma_mpm_const* c, const float* x, const float* v,
                               const float* C, const float* vol, const float* rho, const int32_t* enabled,
                               const int32_t* box, float* grid, void* stream) {
  if (int r = check_args(ctx, kind, hidden)) return r;
  if (kind == 1) return neuma_fail_msg(ctx, "neuma_constitutive_fwd_p2g: elasticity kinds only (0, 2)");
  if (!c || !x || !v || !C || !vol || !rho || !enabled || !grid || !stress)
    return neuma_fail_msg(ctx, "neuma_constitutive_fwd_p2g: NULL argument");
  if ((c->compact || c->sel_mode == 3) && !box)
    return neuma_fail_msg(ctx, "neuma_constitutive_fwd_p2g: a compact grid / sel_mode 3 needs the box");
  if (n <= 0) return 0;
  if (int r = set_attrs(ctx)) return r;
  FusedP2G fp{*c, x, v, C, vol, rho, enabled, box, (float4*)grid};
  FusedG2P nog;
  memset(&nog, 0, sizeof(nog));
  return launch_fwd(ctx, kind, n, F, staged, 0.0f, stress, fp, nog, (cudaStream_t)stream);
}

int neuma_constitutive_fwd_g2p(neuma_ctx* ctx, int n, int hidden, const void* staged, float alpha, float* F_out,
                               const neuma_mpm_const* c, const float* x, const float* F, const float* clip_bound,
                               const int32_t* enabled, const int32_t* box, const float* grid, float* x2, float* v2,
                               float* C2, float* Ftrial, int32_t* box_next, void* stream) {
  if (int r = check_args(ctx, 1, hidden)) return r;
  if (!c || !x || !F || !clip_bound || !enabled || !grid || !x2 || !v2 || !C2 || !Ftrial || !F_out)
    return neuma_fail_msg(ctx, "neuma_constitutive_fwd_g2p: NULL argument");
  if (c->sel_mode != 0) return neuma_fail_msg(ctx, "neuma_constitutive_fwd_g2p: sel_mode must be 0");
  if (c->compact && !box) return neuma_fail_msg(ctx, "neuma_constitutive_fwd_g2p: a compact grid needs the box");
  if (n <= 0) return 0;
  if (int r = set_attrs(ctx)) return r;
  FusedP2G nop;
  memset(&nop, 0, sizeof(nop));
  FusedG2P fg{*c, x, F, clip_bound, enabled, box, (const float4*)grid, x2, v2, C2, Ftrial, box_next};
  return launch_fwd(ctx, 1, n, Ftrial, staged, alpha, F_out, nop, fg, (cudaStream_t)stream);
}

int neuma_constitutive_bwd_staged(neuma_ctx* ctx, int kind, int n, int hidden, const float* F,
                                  const void* staged, float alpha, const float* gout, float* gF,
                                  float* gW0, float* gW1, float* gW2, void* stream) {
  if (int r = check_args(ctx, kind, hidden)) return r;
  if (n <= 0) return 0;
  if (int r = set_attrs(ctx)) return r;
  const bool wgrad = gW0 && gW1 && gW2;
  if (!wgrad && (gW0 || gW1 || gW2))
    return neuma_fail_msg(ctx, "neuma_constitutive_bwd: pass all of gW0,gW1,gW2 or none");
  FusedScatter none;
  memset(&none, 0, sizeof(none));
  FusedGather noa;
  memset(&noa, 0, sizeof(noa));
  return launch_bwd(ctx, wgrad, kind, n, F, staged, alpha, gout, gF, gW0, gW1, gW2, nullptr, 0, none, noa,
                    (cudaStream_t)stream);
}

int neuma_constitutive_partials_floats(neuma_ctx* ctx) { return ctx ? ctx->num_sms * GW_FLOATS : 0; }

int neuma_constitutive_bwd_partials(neuma_ctx* ctx, int kind, int n, int hidden, const float* F,
                                    const void* staged, float alpha, const float* gout, float* gF,
                                    int accumulate_gF, float* gWpart, void* stream) {
  if (int r = check_args(ctx, kind, hidden)) return r;
  if (!gWpart) return neuma_fail_msg(ctx, "neuma_constitutive_bwd_partials: gWpart is NULL");
  if (n <= 0) return 0;
  if (int r = set_attrs(ctx)) return r;
  FusedScatter none;
  memset(&none, 0, sizeof(none));
  FusedGather noa;
  memset(&noa, 0, sizeof(noa));
  return launch_bwd(ctx, true, kind, n, F, staged, alpha, gout, gF, nullptr, nullptr, nullptr, gWpart,
                    accumulate_gF, none, noa, (cudaStream_t)stream);
}

int neuma_constitutive_bwd_scatter(neuma_ctx* ctx, int n, int hidden, const float* Ftrial, const void* staged,
                                   float alpha, const float* gout, float* gFtrial, float* gWpart,
                                   const neuma_mpm_const* c, const float* x, const float* F,
                                   const float* clip_bound, const int32_t* enabled, const float* v2,
                                   const float* gx2, const float* gv2, const float* gC2, const int32_t* box,
                                   float* adj, void* stream) {
  if (int r = check_args(ctx, 1, hidden)) return r;
  if (!gWpart || !c || !x || !F || !clip_bound || !enabled || !v2 || !gx2 || !gv2 || !gC2 || !adj || !gFtrial)
    return neuma_fail_msg(ctx, "neuma_constitutive_bwd_scatter: NULL argument");
  if ((c->compact || c->sel_mode == 3) && !box)
    return neuma_fail_msg(ctx, "neuma_constitutive_bwd_scatter: a compact grid / sel_mode 3 needs the box");
  if (n <= 0) return 0;
  if (int r = set_attrs(ctx)) return r;
  FusedScatter fs{*c, x, F, clip_bound, v2, gx2, gv2, gC2, enabled, box, (float4*)adj};
  FusedGather noa;
  memset(&noa, 0, sizeof(noa));
  return launch_bwd(ctx, true, 1, n, Ftrial, staged, alpha, gout, gFtrial, nullptr, nullptr, nullptr, gWpart, 0,
                    fs, noa, (cudaStream_t)stream);
}

int neuma_constitutive_bwd_gather(neuma_ctx* ctx, int kind, int n, int hidden, const void* staged, float* gWpart,
                                  const neuma_mpm_const* c, const float* x, const float* v, const float* C,
                                  const float* F, const float* stress, const float* vol, const float* rho,
                                  const float* clip_bound, const int32_t* enabled, const float* v2, const float* C2,
                                  const float* gx2, const float* gv2, const float* gC2, const float* gF2,
                                  const int32_t* box, const float* grid, const float* adj, float* gx, float* gv,
                                  float* gC, float* gF, float* gstress, void* stream) {
  if (int r = check_args(ctx, kind, hidden)) return r;
  if (kind == 1) return neuma_fail_msg(ctx, "neuma_constitutive_bwd_gather: elasticity kinds only (0, 2)");
  if (!gWpart || !c || !x || !v || !C || !F || !stress || !vol || !rho || !clip_bound || !enabled || !v2 || !C2 ||
      !gx2 || !gv2 || !gC2 || !gF2 || !grid || !adj || !gx || !gv || !gC || !gF || !gstress)
    return neuma_fail_msg(ctx, "neuma_constitutive_bwd_gather: NULL argument");
  if (c->sel_mode != 0) return neuma_fail_msg(ctx, "neuma_constitutive_bwd_gather: sel_mode must be 0");
  if (c->compact && !box) return neuma_fail_msg(ctx, "neuma_constitutive_bwd_gather: a compact grid needs the box");
  if (n <= 0) return 0;
  if (int r = set_attrs(ctx)) return r;
  FusedScatter nos;
  memset(&nos, 0, sizeof(nos));
  FusedGather fa{*c, x, v, C, F, stress, vol, rho, clip_bound, v2, C2, gx2, gv2, gC2, gF2, enabled, box,
                 (const float4*)grid, (const float4*)adj, gx, gv, gC, gstress};
  return launch_bwd(ctx, true, kind, n, F, staged, 0.0f, gstress, gF, nullptr, nullptr, nullptr, gWpart, 1, nos, fa,
                    (cudaStream_t)stream);
}

int neuma_constitutive_reduce_partials(neuma_ctx* ctx, int hidden, float* gWpart, float* gW0,
                                       float* gW1, float* gW2, void* stream) {
  if (int r = check_args(ctx, 0, hidden)) return r;
  if (!gWpart || !gW0 || !gW1 || !gW2)
    return neuma_fail_msg(ctx, "neuma_constitutive_reduce_partials: NULL argument");
  reduce_partials_kernel<<<(GW_FLOATS + 31) / 32, 1024, 0, (cudaStream_t)stream>>>(ctx->num_sms, gWpart,
                                                                                 gW0, gW1, gW2);
  NEUMA_LAUNCH_CHECK(ctx, "reduce_partials_kernel");
  return 0;
}

int neuma_constitutive_fwd(neuma_ctx* ctx, int kind, int n, int hidden, const float* F,
                           const float* W0, const float* W1, const float* W2, float alpha,
                           float* out, void* stream) {
  if (int r = check_args(ctx, kind, hidden)) return r;
  if (n <= 0) return 0;
  if (int r = neuma_ensure_wstage(ctx, STAGED_BYTES)) return r;
  if (int r = neuma_constitutive_stage(ctx, hidden, W0, W1, W2, ctx->wstage, stream)) return r;
  return neuma_constitutive_fwd_staged(ctx, kind, n, hidden, F, ctx->wstage, alpha, out, stream);
}

int neuma_constitutive_bwd(neuma_ctx* ctx, int kind, int n, int hidden, const float* F,
                           const float* W0, const float* W1, const float* W2, float alpha,
                           const float* gout, float* gF, float* gW0, float* gW1, float* gW2,
                           void* stream) {
  if (int r = check_args(ctx, kind, hidden)) return r;
  if (n <= 0) return 0;
  if (int r = neuma_ensure_wstage(ctx, STAGED_BYTES)) return r;
  if (int r = neuma_constitutive_stage(ctx, hidden, W0, W1, W2, ctx->wstage, stream)) return r;
  return neuma_constitutive_bwd_staged(ctx, kind, n, hidden, F, ctx->wstage, alpha, gout, gF, gW0,
                                       gW1, gW2, stream);
}

}  // extern "C"
