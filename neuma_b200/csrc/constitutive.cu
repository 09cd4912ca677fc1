// Neural constitutive law, fused: SVD -> invariants -> 13->H->H->9 bias-free GELU
// MLP -> symmetrise -> polar recombination, forward and backward, one kernel each.
// Replaces ~35 torch launches + 1 Warp launch per call of
//   InvariantFullMetaElasticity.forward   modules/nclaw/material/meta.py:196-221
//   InvariantFullMetaPlasticity.forward   modules/nclaw/material/meta.py:468-489
// and the Warp batch_svd / SVDFunction of modules/nclaw/warp/svd.py:9-96.
//
// The three Linear layers -- and, in the backward, their data-gradient and weight-gradient
// GEMMs -- run on the 5th-generation tensor cores: tcgen05.mma kind::f16 on bf16 operands
// staged in shared memory, fp32 accumulators in TMEM.  Every fp32 value is split into
// three bf16 terms (x = x0 + x1 + x2, 24 mantissa bits) and each GEMM is the six products
// with i + j <= 2, accumulated smallest first ("bf16x3"): ~2e-7 relative error against
// fp64, i.e. fp32-class results from the tensor pipe.  One thread owns one particle row
// (= one TMEM lane) for the SVD / invariants / polar recombination, two threads share a
// row for the GELU epilogues between the GEMMs.
//
// Shared-memory operand layout (SWIZZLE_NONE core matrices of 8 rows x 16 bytes, see
// umma.cuh): an [R x K] fp32 tile is stored per split as [K/8][R][8] bf16.  The same bytes
// are a K-major operand (rows = particles or output features) for the forward /
// data-gradient GEMMs and, with LBO/SBO swapped and the major bit set, an MN-major
// operand X^T for the weight-gradient GEMMs that contract over the 128 particles of the
// tile -- no transposed copy is ever written.  (kind::tf32 would need one: its MN-major
// form exists only with the 128B_BASE32B swizzle, measured with tools/umma_probe.cu.)
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
#include "svd3.cuh"
#include "umma.cuh"

namespace {

constexpr int H = 64;      // hidden width (layer_widths: [64, 64] in every shipped config)
constexpr int NF = 13;     // 3 + 9 + 1 invariants
constexpr int TP = 128;    // particles per tile = TMEM lanes = rows of every GEMM
constexpr int NT = 256;    // threads per block: two per particle row in the GELU phases
constexpr int KF = 16;     // invariants (13) and MLP head (9) padded to one bf16 K step

struct Kinematics {
  Mat3 F, U, V, R;
  float sig[3];
  float feat[NF];
};

__device__ __forceinline__ void kinematics(const float* __restrict__ Fp, bool normalize,
                                           Kinematics& k) {
  k.F = mat3_load(Fp);
  svd3(k.F, k.U, k.sig, k.V);
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


// ---- shared-memory map (bytes)
// Activation tiles (A operands; wgrad operands): per split [K/8][128][8] bf16, splits adjacent.
// Weights, N-STACKED so that one MMA multiplies an A split with several B splits at once:
//   Wk (forward, K-major B):      [K/8][3R][8], row = split*R + n          (N = 3R | 2R | R)
//   Wt (data gradient, MN-major): [split][cols/8][R][8]; read transposed, the three splits
//                                 are consecutive 8-wide N chunks           (N = 3c | 2c | c)
constexpr uint32_t SMALL_SPLIT = TP * KF * 2;  // [128 x 16] activation tile, one split
constexpr uint32_t BIG_SPLIT = TP * H * 2;     // [128 x 64] activation tile, one split
constexpr uint32_t W0_BYTES = 3 * H * KF * 2, W1_BYTES = 3 * H * H * 2, W2_BYTES = 3 * KF * H * 2;
constexpr uint32_t OFF_W0K = 0;
constexpr uint32_t OFF_W1K = OFF_W0K + W0_BYTES;
constexpr uint32_t OFF_W2K = OFF_W1K + W1_BYTES;
constexpr uint32_t OFF_FEAT = OFF_W2K + W2_BYTES;          // invariants
constexpr uint32_t OFF_BIGA = OFF_FEAT + 3 * SMALL_SPLIT;  // gelu(h1), later dL/dh1
constexpr uint32_t FWD_SMEM = OFF_BIGA + 3 * BIG_SPLIT;    // 98304
constexpr uint32_t OFF_GO = FWD_SMEM;                      // dL/d(MLP output)
constexpr uint32_t OFF_BIGB = OFF_GO + 3 * SMALL_SPLIT;    // gelu(h2), later dL/dh2
constexpr uint32_t OFF_W0T = OFF_BIGB + 3 * BIG_SPLIT;
constexpr uint32_t OFF_W1T = OFF_W0T + W0_BYTES;
constexpr uint32_t OFF_W2T = OFF_W1T + W1_BYTES;
constexpr uint32_t BWD_SMEM = OFF_W2T + W2_BYTES;          // 196608

// ---- TMEM columns (fp32, lane = particle row unless noted).  A stacked GEMM leaves three
// partial sums (blocks of N columns) that the reading thread adds: blk0 + blk1 + blk2.
constexpr uint32_t COL_D1 = 0;      // gelu'(h1)                                          (bwd)
constexpr uint32_t COL_D2 = 64;     // gelu'(h2)                                          (bwd)
constexpr uint32_t COL_S = 128;     // 3 x 64 (or 3 x 16) partial sums of the current GEMM (bwd)
constexpr uint32_t COL_GW1 = 320;   // dL/dW1 [o][k]: 2 x 64; M=64 layout: row o -> lane (o%16)+32(o/16)
constexpr uint32_t COL_GW0 = 448;   // dL/dW0 [o][j]: 2 x 16  (same layout)
constexpr uint32_t COL_GW2T = 480;  // dL/dW2^T [k][j]: 2 x 16 (same layout)
constexpr uint32_t FWD_COL_S = 0;
constexpr uint32_t FWD_TMEM_COLS = 256, BWD_TMEM_COLS = 512;

// x -> three bf16 terms, two values at a time (cvt.rn.bf16x2: first source -> upper half)
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& s0, uint32_t& s1, uint32_t& s2) {
  s0 = pack_bf16x2(a, b);
  a -= __uint_as_float(s0 << 16);
  b -= __uint_as_float(s0 & 0xFFFF0000u);
  s1 = pack_bf16x2(a, b);
  a -= __uint_as_float(s1 << 16);
  b -= __uint_as_float(s1 & 0xFFFF0000u);
  s2 = pack_bf16x2(a, b);
}
// eight consecutive K values of particle row r -> chunk c of the three split images of a tile
__device__ __forceinline__ void store_chunk(uint8_t* img, uint32_t split_bytes, int c, int r, const float* x) {
  uint4 s0, s1, s2;
  split_pair(x[0], x[1], s0.x, s1.x, s2.x);
  split_pair(x[2], x[3], s0.y, s1.y, s2.y);
  split_pair(x[4], x[5], s0.z, s1.z, s2.z);
  split_pair(x[6], x[7], s0.w, s1.w, s2.w);
  uint8_t* p = img + ((size_t)c * TP + r) * 16;
  *reinterpret_cast<uint4*>(p) = s0;
  *reinterpret_cast<uint4*>(p + split_bytes) = s1;
  *reinterpret_cast<uint4*>(p + 2 * split_bytes) = s2;
}

// weights -> stacked split images.  src is row-major [rows x cols], zero padded to [R x K].
template <bool TRANSPOSED_VIEW>
__device__ __forceinline__ void stage_weight(uint8_t* img, const float* __restrict__ src, int rows, int cols,
                                             int R, int K) {
  uint16_t* dst = reinterpret_cast<uint16_t*>(img);
  for (int e = threadIdx.x; e < R * K; e += blockDim.x) {
    const int r = e / K, k = e % K;
    float x = (r < rows && k < cols) ? src[r * cols + k] : 0.f;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const uint32_t h = pack_bf16x2(x, 0.f) & 0xFFFFu;
      const size_t o = TRANSPOSED_VIEW ? ((size_t)t * (K >> 3) + (k >> 3)) * R + r   // [split][k/8][R]
                                       : ((size_t)(k >> 3) * 3 + t) * R + r;         // [k/8][3R]
      dst[o * 8 + (k & 7)] = (uint16_t)h;
      x -= __uint_as_float(h << 16);
    }
  }
}

// ---- MMA issue.  Called by ALL lanes of one warp (elect.sync inside the helpers).
struct Desc {
  uint64_t d;       // descriptor of split 0, K step 0
  uint32_t split;   // (bytes >> 4) between splits (A operands)
  uint32_t kstep;   // (bytes >> 4) per K = 16 step
};
// activation tile / plain operand, K-major (rows = particles): LBO = R*16, SBO = 128
__device__ __forceinline__ Desc desc_kmajor(uint32_t addr, uint32_t split_bytes, int R) {
  return Desc{umma::smem_desc(addr, (uint32_t)R * 16u, 128u), split_bytes >> 4, ((uint32_t)R * 32u) >> 4};
}
// the same bytes read transposed (MN-major): SBO = R*16, LBO = 128, 16 rows per K step
__device__ __forceinline__ Desc desc_mnmajor(uint32_t addr, uint32_t split_bytes, int R) {
  return Desc{umma::smem_desc(addr, 128u, (uint32_t)R * 16u), split_bytes >> 4, 256u >> 4};
}
// forward weights, N-stacked K-major: [K/8][3R][8]
__device__ __forceinline__ Desc desc_wk(uint32_t addr, int R) {
  return Desc{umma::smem_desc(addr, (uint32_t)R * 48u, 128u), 0u, ((uint32_t)R * 96u) >> 4};
}

// D[128 x 3N | 2N | N] (+)= A_i B_stack^T for the three A splits: after the call
// blk0 = (A0+A1+A2) B0, blk1 = (A0+A1) B1, blk2 = A0 B2  (the six bf16x3 products).
template <int NK>
__device__ __forceinline__ void gemm_stacked(uint32_t tmem_d, const Desc a, const Desc b, uint32_t idesc_n1) {
  // idesc_n1: descriptor with the N field of ONE block; N is a multiple of 16, so 2N and 3N
  // are the same field times 2 and 3
  const uint32_t nfield = idesc_n1 & (0x3Fu << 17), rest = idesc_n1 & ~(0x3Fu << 17);
#pragma unroll
  for (int k = 0; k < NK; ++k) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
      umma::mma_bf16_ss(tmem_d, a.d + (uint64_t)(i * a.split + k * a.kstep), b.d + (uint64_t)(k * b.kstep),
                        rest | (nfield * (3 - i)), (k | i) ? 1u : 0u);
  }
}
// weight gradient, contraction over the 128 particles of the tile (8 K steps), M = 64:
// blk0 (+)= (X0 + X1)^T Y0, blk1 (+)= X0^T Y1  (three leading bf16 products: the dropped
// ones are 2^-16 relative with random sign, far below the fp32 atomics' own ordering noise)
__device__ __forceinline__ void gemm_wgrad(uint32_t tmem_d, const Desc x, const Desc y, uint32_t idesc_n1,
                                           bool accumulate) {
  const uint32_t nfield = idesc_n1 & (0x3Fu << 17), rest = idesc_n1 & ~(0x3Fu << 17);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    umma::mma_bf16_ss(tmem_d, x.d + (uint64_t)(k * x.kstep), y.d + (uint64_t)(k * y.kstep), rest | (nfield * 2),
                      (accumulate || k) ? 1u : 0u);
    umma::mma_bf16_ss(tmem_d, x.d + (uint64_t)(x.split + k * x.kstep), y.d + (uint64_t)(k * y.kstep),
                      rest | nfield, 1u);
  }
}

// sum of the three stacked partial results for 16 consecutive columns of this thread's row
template <int BLOCK>
__device__ __forceinline__ void load_sum3(uint32_t taddr, float (&v)[16]) {
  float b1[16], b2[16];
  umma::tmem_ld16(taddr, v);
  umma::tmem_ld16(taddr + BLOCK, b1);
  umma::tmem_ld16(taddr + 2 * BLOCK, b2);
  umma::tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] += b1[i] + b2[i];
}

// Phi(h) and phi(h) of the standard normal for the exact-erf GELU (nn.GELU() default,
// modules/nclaw/material/utils.py:16-17): gelu = h Phi(h), gelu' = Phi(h) + h phi(h).
// erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, on a par with fp32 erff and measured
// 2.5x closer to the fp64 GELU than torch's own fp32 kernel); it shares exp(-h^2/2) with
// phi, so the backward's gelu' costs three extra instructions.  Branch-free: 2 MUFU + ~12 FMA.
__device__ __forceinline__ void normal_cdf_pdf(float h, float& cdf, float& pdf) {
  const float ax = fabsf(h) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
  const float e = __expf(-ax * ax);
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(t, p, 1.421413741f);
  p = fmaf(t, p, -0.284496736f);
  p = fmaf(t, p, 0.254829592f);
  const float q = 0.5f * (p * t) * e;               // 0.5 (1 - erf|x|)
  cdf = (h < 0.0f) ? q : 1.0f - q;
  pdf = 0.39894228040143268f * e;
}

// GELU phase: the thread reads its 32 accumulator columns, writes gelu() as the next GEMM's
// A operand (chunks 4*half .. 4*half+3 of the tile at `img`) and, in the backward, stores
// gelu'() in TMEM for the data-gradient pass.
template <bool KEEP_GRAD>
__device__ __forceinline__ void gelu_phase(uint32_t t_acc, uint32_t t_dgelu, uint8_t* img, int r, int half) {
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    float h[16], a[16];
    load_sum3<H>(t_acc + 32 * half + 16 * q, h);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float cdf, pdf;
      normal_cdf_pdf(h[i], cdf, pdf);
      a[i] = h[i] * cdf;
      if (KEEP_GRAD) h[i] = cdf + h[i] * pdf;
    }
    store_chunk(img, BIG_SPLIT, 4 * half + 2 * q, r, a);
    store_chunk(img, BIG_SPLIT, 4 * half + 2 * q + 1, r, a + 8);
    if (KEEP_GRAD) umma::tmem_st16(t_dgelu + 32 * half + 16 * q, h);
  }
  if (KEEP_GRAD) umma::tmem_wait_st();
}
// dL/dh = dL/da * gelu'(h): both in TMEM; result becomes an A operand tile at `img`
__device__ __forceinline__ void grad_phase(uint32_t t_acc, uint32_t t_dgelu, uint8_t* img, int r, int half,
                                           uint64_t* wait_bar, uint32_t& wait_phase) {
  float g[2][16];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    float d[16];
    load_sum3<H>(t_acc + 32 * half + 16 * q, g[q]);
    umma::tmem_ld16(t_dgelu + 32 * half + 16 * q, d);
    umma::tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) g[q][i] *= d[i];
  }
  if (wait_bar) {  // the weight-gradient GEMM still reading the tile we are about to overwrite
    umma::mbar_wait(umma::smem_u32(wait_bar), wait_phase);
    wait_phase ^= 1u;
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    store_chunk(img, BIG_SPLIT, 4 * half + 2 * q, r, g[q]);
    store_chunk(img, BIG_SPLIT, 4 * half + 2 * q + 1, r, g[q] + 8);
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

struct Setup {
  uint32_t tmem, smem;
};
// weights -> smem images, TMEM allocation, mbarrier init; returns after a block-wide sync
template <bool BWD>
__device__ __forceinline__ Setup block_setup(uint8_t* sm, uint64_t* mbar, uint32_t* tmem_slot,
                                             const float* W0, const float* W1, const float* W2,
                                             uint32_t tmem_cols) {
  stage_weight<false>(sm + OFF_W0K, W0, H, NF, H, KF);
  stage_weight<false>(sm + OFF_W1K, W1, H, H, H, H);
  stage_weight<false>(sm + OFF_W2K, W2, 9, H, KF, H);
  if (BWD) {
    stage_weight<true>(sm + OFF_W0T, W0, H, NF, H, KF);
    stage_weight<true>(sm + OFF_W1T, W1, H, H, H, H);
    stage_weight<true>(sm + OFF_W2T, W2, 9, H, KF, H);
  }
  if (threadIdx.x < 32) umma::tmem_alloc(umma::smem_u32(tmem_slot), tmem_cols);
  if (threadIdx.x == 32) {
    umma::mbar_init(umma::smem_u32(mbar), 1);
    if (BWD) umma::mbar_init(umma::smem_u32(mbar + 1), 1);
    umma::fence_mbar_init();
  }
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  return Setup{*tmem_slot, umma::smem_u32(sm)};
}

// publish this thread's smem operand writes / TMEM reads, then hand over to the MMA warps
__device__ __forceinline__ void publish_and_sync() {
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
}
__device__ __forceinline__ void wait_mma(uint64_t* mbar, uint32_t& phase) {
  umma::mbar_wait(umma::smem_u32(mbar), phase);
  phase ^= 1u;
  umma::fence_after_sync();
}

// ------------------------------------------------------------------ forward
__global__ void __launch_bounds__(NT, 2)
constitutive_fwd_kernel(int kind, int n, const float* __restrict__ F, const float* __restrict__ W0,
                        const float* __restrict__ W1, const float* __restrict__ W2, float alpha,
                        float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t mbar[1];
  __shared__ uint32_t tmem_slot;
  const Setup su = block_setup<false>(sm, mbar, &tmem_slot, W0, W1, W2, FWD_TMEM_COLS);
  const int tid = threadIdx.x, r = tid & (TP - 1), half = tid >> 7, warp = tid >> 5;
  const uint32_t tlane = su.tmem + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t phase = 0;
  const Desc w0 = desc_wk(su.smem + OFF_W0K, H), w1 = desc_wk(su.smem + OFF_W1K, H), w2 = desc_wk(su.smem + OFF_W2K, KF);
  const Desc feat = desc_kmajor(su.smem + OFF_FEAT, SMALL_SPLIT, TP), act = desc_kmajor(su.smem + OFF_BIGA, BIG_SPLIT, TP);
  constexpr uint32_t ID64 = umma::idesc_bf16(TP, H, 0, 0), ID16 = umma::idesc_bf16(TP, KF, 0, 0);
  const int tiles = (n + TP - 1) / TP;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int p = tile * TP + r;
    Kinematics k;
    if (half == 0) {
      float f[KF];
#pragma unroll
      for (int i = 0; i < KF; ++i) f[i] = 0.f;
      if (p < n) {
        kinematics(F + (size_t)p * 9, kind != 2, k);
#pragma unroll
        for (int i = 0; i < NF; ++i) f[i] = k.feat[i];
      }
      store_chunk(sm + OFF_FEAT, SMALL_SPLIT, 0, r, f);
      store_chunk(sm + OFF_FEAT, SMALL_SPLIT, 1, r, f + 8);
    }
    publish_and_sync();
    if (warp == 0) {
      umma::fence_after_sync();
      gemm_stacked<1>(su.tmem + FWD_COL_S, feat, w0, ID64);
      umma::commit(umma::smem_u32(mbar));
    }
    wait_mma(mbar, phase);
    gelu_phase<false>(tlane + FWD_COL_S, 0, sm + OFF_BIGA, r, half);
    publish_and_sync();
    if (warp == 0) {
      umma::fence_after_sync();
      gemm_stacked<4>(su.tmem + FWD_COL_S, act, w1, ID64);
      umma::commit(umma::smem_u32(mbar));
    }
    wait_mma(mbar, phase);
    gelu_phase<false>(tlane + FWD_COL_S, 0, sm + OFF_BIGA, r, half);
    publish_and_sync();
    if (warp == 0) {
      umma::fence_after_sync();
      gemm_stacked<4>(su.tmem + FWD_COL_S, act, w2, ID16);
      umma::commit(umma::smem_u32(mbar));
    }
    wait_mma(mbar, phase);
    if (half == 0) {
      float o[16];
      load_sum3<KF>(tlane + FWD_COL_S, o);
      if (p < n) {
        const Mat3 X = head_to_sym(o);
        const Mat3 P = mat3_mul(k.R, X);
        Mat3 O;
        if (kind == 1) {
#pragma unroll
          for (int i = 0; i < 9; ++i) O.m[i] = alpha * P.m[i] + k.F.m[i];  // delta_Fp + F
        } else {
          O = mat3_mul_nt(P, k.F);  // cauchy = P F^T
        }
        mat3_store(out + (size_t)p * 9, O);
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(su.tmem, FWD_TMEM_COLS);
}

// ------------------------------------------------------------------ backward
template <bool WGRAD>
__global__ void __launch_bounds__(NT, 1)
constitutive_bwd_kernel(int kind, int n, const float* __restrict__ F, const float* __restrict__ W0,
                        const float* __restrict__ W1, const float* __restrict__ W2, float alpha,
                        const float* __restrict__ gout, float* __restrict__ gF,
                        float* __restrict__ gW0, float* __restrict__ gW1,
                        float* __restrict__ gW2) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t mbar[2];  // [0]: the forward / data-gradient chain (warp 0); [1]: weight gradients (warp 1)
  __shared__ uint32_t tmem_slot;
  const Setup su = block_setup<true>(sm, mbar, &tmem_slot, W0, W1, W2, BWD_TMEM_COLS);
  const int tid = threadIdx.x, r = tid & (TP - 1), half = tid >> 7, warp = tid >> 5;
  const uint32_t tlane = su.tmem + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t phase = 0, wphase = 0;
  uint64_t* const wbar = WGRAD ? mbar + 1 : nullptr;
  const Desc w0 = desc_wk(su.smem + OFF_W0K, H), w1 = desc_wk(su.smem + OFF_W1K, H), w2 = desc_wk(su.smem + OFF_W2K, KF);
  // transposed weights: [split][cols/8][R][8] read MN-major; the splits are adjacent N chunks
  const Desc w0t = desc_mnmajor(su.smem + OFF_W0T, 0, H), w1t = desc_mnmajor(su.smem + OFF_W1T, 0, H),
             w2t = desc_mnmajor(su.smem + OFF_W2T, 0, KF);
  const Desc feat = desc_kmajor(su.smem + OFF_FEAT, SMALL_SPLIT, TP), go = desc_kmajor(su.smem + OFF_GO, SMALL_SPLIT, TP),
             biga = desc_kmajor(su.smem + OFF_BIGA, BIG_SPLIT, TP), bigb = desc_kmajor(su.smem + OFF_BIGB, BIG_SPLIT, TP);
  const Desc featT = desc_mnmajor(su.smem + OFF_FEAT, SMALL_SPLIT, TP), goT = desc_mnmajor(su.smem + OFF_GO, SMALL_SPLIT, TP),
             bigaT = desc_mnmajor(su.smem + OFF_BIGA, BIG_SPLIT, TP), bigbT = desc_mnmajor(su.smem + OFF_BIGB, BIG_SPLIT, TP);
  constexpr uint32_t ID64 = umma::idesc_bf16(TP, H, 0, 0), ID16 = umma::idesc_bf16(TP, KF, 0, 0);
  constexpr uint32_t ID64_BT = umma::idesc_bf16(TP, H, 0, 1), ID16_BT = umma::idesc_bf16(TP, KF, 0, 1);
  constexpr uint32_t IDW64 = umma::idesc_bf16(64, H, 1, 1), IDW16 = umma::idesc_bf16(64, KF, 1, 1);
  const int tiles = (n + TP - 1) / TP;
  bool first = true;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, first = false) {
    const int p = tile * TP + r;
    const bool live = (half == 0) && (p < n);
    Kinematics k;
    Mat3 gR, gFd;
    // ---- invariants -> h1.  (FEAT / BIGA were last read by the previous tile's dW0 GEMM.)
    if (WGRAD && !first) {
      umma::mbar_wait(umma::smem_u32(wbar), wphase);
      wphase ^= 1u;
    }
    if (half == 0) {
      float f[KF];
#pragma unroll
      for (int i = 0; i < KF; ++i) f[i] = 0.f;
      if (live) {
        kinematics(F + (size_t)p * 9, kind != 2, k);
#pragma unroll
        for (int i = 0; i < NF; ++i) f[i] = k.feat[i];
      }
      store_chunk(sm + OFF_FEAT, SMALL_SPLIT, 0, r, f);
      store_chunk(sm + OFF_FEAT, SMALL_SPLIT, 1, r, f + 8);
    }
    publish_and_sync();
    if (warp == 0) {
      umma::fence_after_sync();
      gemm_stacked<1>(su.tmem + COL_S, feat, w0, ID64);
      umma::commit(umma::smem_u32(mbar));
    }
    wait_mma(mbar, phase);
    // ---- a1 = gelu(h1) -> h2
    gelu_phase<true>(tlane + COL_S, tlane + COL_D1, sm + OFF_BIGA, r, half);
    publish_and_sync();
    if (warp == 0) {
      umma::fence_after_sync();
      gemm_stacked<4>(su.tmem + COL_S, biga, w1, ID64);
      umma::commit(umma::smem_u32(mbar));
    }
    wait_mma(mbar, phase);
    // ---- a2 = gelu(h2) -> o
    gelu_phase<true>(tlane + COL_S, tlane + COL_D2, sm + OFF_BIGB, r, half);
    publish_and_sync();
    if (warp == 0) {
      umma::fence_after_sync();
      gemm_stacked<4>(su.tmem + COL_S, bigb, w2, ID16);
      umma::commit(umma::smem_u32(mbar));
    }
    wait_mma(mbar, phase);
    // ---- output head: dL/dout -> dL/d(MLP output), dL/dR, direct dL/dF
    if (half == 0) {
      float o[16], g9[KF];
      load_sum3<KF>(tlane + COL_S, o);
#pragma unroll
      for (int i = 0; i < KF; ++i) g9[i] = 0.f;
      if (live) {
        const Mat3 X = head_to_sym(o);
        const Mat3 P = mat3_mul(k.R, X);
        const Mat3 G = mat3_load(gout + (size_t)p * 9);
        Mat3 gP;  // out = P F^T (elastic)  |  out = F + alpha P (plastic)
        if (kind == 1) {
#pragma unroll
          for (int i = 0; i < 9; ++i) { gP.m[i] = alpha * G.m[i]; gFd.m[i] = G.m[i]; }
        } else {
          gP = mat3_mul(G, k.F);
          gFd = mat3_mul_tn(G, P);
        }
        gR = mat3_mul(gP, X);  // X symmetric
        const Mat3 gXs = mat3_mul_tn(k.R, gP);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) g9[3 * i + j] = 0.5f * (gXs(i, j) + gXs(j, i));
      }
      store_chunk(sm + OFF_GO, SMALL_SPLIT, 0, r, g9);
      store_chunk(sm + OFF_GO, SMALL_SPLIT, 1, r, g9 + 8);
    }
    publish_and_sync();
    if (warp == 0) {
      umma::fence_after_sync();
      gemm_stacked<1>(su.tmem + COL_S, go, w2t, ID64_BT);  // dL/da2 = go W2
      umma::commit(umma::smem_u32(mbar));
    } else if (WGRAD && warp == 1) {
      umma::fence_after_sync();
      gemm_wgrad(su.tmem + COL_GW2T, bigbT, goT, IDW16, !first);  // dW2^T += a2^T go
      umma::commit(umma::smem_u32(wbar));
    }
    wait_mma(mbar, phase);
    // ---- dL/dh2 (overwrites a2 once dW2 has read it) -> dL/da1, dW1
    grad_phase(tlane + COL_S, tlane + COL_D2, sm + OFF_BIGB, r, half, wbar, wphase);
    publish_and_sync();
    if (warp == 0) {
      umma::fence_after_sync();
      gemm_stacked<4>(su.tmem + COL_S, bigb, w1t, ID64_BT);  // dL/da1 = g2 W1
      umma::commit(umma::smem_u32(mbar));
    } else if (WGRAD && warp == 1) {
      umma::fence_after_sync();
      gemm_wgrad(su.tmem + COL_GW1, bigbT, bigaT, IDW64, !first);  // dW1 += g2^T a1
      umma::commit(umma::smem_u32(wbar));
    }
    wait_mma(mbar, phase);
    // ---- dL/dh1 (overwrites a1 once dW1 has read it) -> dL/dfeat, dW0
    grad_phase(tlane + COL_S, tlane + COL_D1, sm + OFF_BIGA, r, half, wbar, wphase);
    publish_and_sync();
    if (warp == 0) {
      umma::fence_after_sync();
      gemm_stacked<4>(su.tmem + COL_S, biga, w0t, ID16_BT);  // dL/dfeat = g1 W0
      umma::commit(umma::smem_u32(mbar));
    } else if (WGRAD && warp == 1) {
      umma::fence_after_sync();
      gemm_wgrad(su.tmem + COL_GW0, bigaT, featT, IDW16, !first);  // dW0 += g1^T feat
      umma::commit(umma::smem_u32(wbar));
    }
    wait_mma(mbar, phase);
    // ---- invariants -> F
    if (half == 0) {
      float gfeat[16];
      load_sum3<KF>(tlane + COL_S, gfeat);
      if (live) {
        Mat3 gFt = gFd;
        {  // FtF = F^T F : dF += F (G + G^T)
          Mat3 Gs;
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) Gs(i, j) = gfeat[3 + 3 * i + j] + gfeat[3 + 3 * j + i];
          const Mat3 t = mat3_mul(k.F, Gs);
#pragma unroll
          for (int i = 0; i < 9; ++i) gFt.m[i] += t.m[i];
        }
        {  // det F : dF += gdet * cof(F)
          const Mat3& A = k.F;
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
          const Mat3 M = mat3_mul(mat3_mul_tn(k.U, gR), k.V);
          Mat3 K;
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
              K(i, j) = (i == j) ? gfeat[i]
                                 : (M(i, j) - M(j, i)) / signed_floor(k.sig[i] + k.sig[j], 1e-6f);
          const Mat3 t = mat3_mul_nt(mat3_mul(k.U, K), k.V);
#pragma unroll
          for (int i = 0; i < 9; ++i) gFt.m[i] += t.m[i];
        }
        mat3_store(gF + (size_t)p * 9, gFt);
      }
    }
  }
  if (WGRAD && !first) {
    umma::mbar_wait(umma::smem_u32(wbar), wphase);  // the last tile's dW0 GEMM
    umma::fence_after_sync();
    if (half == 0) {
      // M=64 accumulators: lane 32*q + i (i < 16) holds row 16*q + i; value = blk0 + blk1
      const int i16 = tid & 31, row = (warp & 3) * 16 + i16;
      float a[16], b[16];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        umma::tmem_ld16(tlane + COL_GW1 + 16 * c, a);
        umma::tmem_ld16(tlane + COL_GW1 + H + 16 * c, b);
        umma::tmem_wait_ld();
        if (i16 < 16)
#pragma unroll
          for (int j = 0; j < 16; ++j) atomicAdd(gW1 + row * H + 16 * c + j, a[j] + b[j]);
      }
      umma::tmem_ld16(tlane + COL_GW0, a);
      umma::tmem_ld16(tlane + COL_GW0 + KF, b);
      umma::tmem_wait_ld();
      if (i16 < 16)
#pragma unroll
        for (int j = 0; j < NF; ++j) atomicAdd(gW0 + row * NF + j, a[j] + b[j]);
      umma::tmem_ld16(tlane + COL_GW2T, a);
      umma::tmem_ld16(tlane + COL_GW2T + KF, b);
      umma::tmem_wait_ld();
      if (i16 < 16)
#pragma unroll
        for (int j = 0; j < 9; ++j) atomicAdd(gW2 + j * H + row, a[j] + b[j]);
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(su.tmem, BWD_TMEM_COLS);
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

bool g_attr_set = false;

int set_attrs(neuma_ctx* ctx) {
  if (g_attr_set) return 0;
  NEUMA_CUDA(ctx, cudaFuncSetAttribute(constitutive_fwd_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM));
  NEUMA_CUDA(ctx, cudaFuncSetAttribute(constitutive_bwd_kernel<true>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM));
  NEUMA_CUDA(ctx, cudaFuncSetAttribute(constitutive_bwd_kernel<false>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM));
  g_attr_set = true;
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

int neuma_constitutive_fwd(neuma_ctx* ctx, int kind, int n, int hidden, const float* F,
                           const float* W0, const float* W1, const float* W2, float alpha,
                           float* out, void* stream) {
  if (!ctx) return 2;
  if (hidden != H) return neuma_fail_msg(ctx, "neuma_constitutive: only hidden width 64 is built");
  if (kind < 0 || kind > 2) return neuma_fail_msg(ctx, "neuma_constitutive: kind must be 0, 1 or 2");
  if (n <= 0) return 0;
  if (int r = set_attrs(ctx)) return r;
  const int tiles = (n + TP - 1) / TP;
  const int grid = tiles < ctx->num_sms * 2 ? tiles : ctx->num_sms * 2;
  constitutive_fwd_kernel<<<grid, NT, FWD_SMEM, (cudaStream_t)stream>>>(kind, n, F, W0, W1, W2,
                                                                        alpha, out);
  NEUMA_LAUNCH_CHECK(ctx, "constitutive_fwd_kernel");
  return 0;
}

int neuma_constitutive_bwd(neuma_ctx* ctx, int kind, int n, int hidden, const float* F,
                           const float* W0, const float* W1, const float* W2, float alpha,
                           const float* gout, float* gF, float* gW0, float* gW1, float* gW2,
                           void* stream) {
  if (!ctx) return 2;
  if (hidden != H) return neuma_fail_msg(ctx, "neuma_constitutive: only hidden width 64 is built");
  if (kind < 0 || kind > 2) return neuma_fail_msg(ctx, "neuma_constitutive: kind must be 0, 1 or 2");
  if (n <= 0) return 0;
  if (int r = set_attrs(ctx)) return r;
  const bool wgrad = gW0 && gW1 && gW2;
  if (!wgrad && (gW0 || gW1 || gW2))
    return neuma_fail_msg(ctx, "neuma_constitutive_bwd: pass all of gW0,gW1,gW2 or none");
  const int tiles = (n + TP - 1) / TP;
  const int grid = tiles < ctx->num_sms ? tiles : ctx->num_sms;
  cudaStream_t st = (cudaStream_t)stream;
  if (wgrad)
    constitutive_bwd_kernel<true><<<grid, NT, BWD_SMEM, st>>>(kind, n, F, W0, W1, W2, alpha, gout,
                                                              gF, gW0, gW1, gW2);
  else
    constitutive_bwd_kernel<false><<<grid, NT, BWD_SMEM, st>>>(kind, n, F, W0, W1, W2, alpha, gout,
                                                               gF, gW0, gW1, gW2);
  NEUMA_LAUNCH_CHECK(ctx, "constitutive_bwd_kernel");
  return 0;
}

}  // extern "C"
