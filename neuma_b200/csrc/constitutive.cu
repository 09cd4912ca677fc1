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


// ---- shared-memory map (bytes).  Each operand: 3 bf16 splits, each [K/8][R][8].
constexpr uint32_t W0_SPLIT = H * KF * 2;    // W0 [H x 13->16]
constexpr uint32_t W1_SPLIT = H * H * 2;     // W1 [H x H]
constexpr uint32_t W2_SPLIT = KF * H * 2;    // W2 [9->16 x H]
constexpr uint32_t SMALL_SPLIT = TP * KF * 2;  // [128 x 16] activation tile
constexpr uint32_t BIG_SPLIT = TP * H * 2;     // [128 x 64] activation tile
constexpr uint32_t OFF_W0 = 0;
constexpr uint32_t OFF_W1 = OFF_W0 + 3 * W0_SPLIT;
constexpr uint32_t OFF_W2 = OFF_W1 + 3 * W1_SPLIT;
constexpr uint32_t OFF_FEAT = OFF_W2 + 3 * W2_SPLIT;     // invariants
constexpr uint32_t OFF_BIGA = OFF_FEAT + 3 * SMALL_SPLIT;  // gelu(h1), later dL/dh1
constexpr uint32_t FWD_SMEM = OFF_BIGA + 3 * BIG_SPLIT;    // 98304
constexpr uint32_t OFF_GO = FWD_SMEM;                      // dL/d(MLP output)
constexpr uint32_t OFF_BIGB = OFF_GO + 3 * SMALL_SPLIT;    // gelu(h2), later dL/dh2
constexpr uint32_t BWD_SMEM = OFF_BIGB + 3 * BIG_SPLIT;    // 159744

// ---- TMEM columns (fp32 accumulators, lane = particle row unless noted)
constexpr uint32_t COL_H1 = 0;     // h1, then gelu'(h1)
constexpr uint32_t COL_H2 = 64;    // h2, then gelu'(h2), then dL/da1                (bwd)
constexpr uint32_t COL_O = 128;    // MLP output (16), then dL/dfeat                  (bwd)
constexpr uint32_t COL_GA2 = 144;  // dL/da2                                          (bwd)
constexpr uint32_t COL_GW1 = 208;  // dL/dW1  [o][k]   M=64 layout: row o -> lane (o%16)+32(o/16)
constexpr uint32_t COL_GW0 = 272;  // dL/dW0  [o][j]   (same layout)
constexpr uint32_t COL_GW2T = 288; // dL/dW2^T [k][j]  (same layout)
constexpr uint32_t FWD_COL_O = 64;
constexpr uint32_t FWD_TMEM_COLS = 128, BWD_TMEM_COLS = 512;

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
// eight consecutive K values of row r -> chunk c of the three split images of an [R x K] tile
__device__ __forceinline__ void store_chunk(uint8_t* img, uint32_t split_bytes, int R, int c, int r,
                                            const float* x) {
  uint4 s0, s1, s2;
  split_pair(x[0], x[1], s0.x, s1.x, s2.x);
  split_pair(x[2], x[3], s0.y, s1.y, s2.y);
  split_pair(x[4], x[5], s0.z, s1.z, s2.z);
  split_pair(x[6], x[7], s0.w, s1.w, s2.w);
  uint8_t* p = img + ((size_t)c * R + r) * 16;
  *reinterpret_cast<uint4*>(p) = s0;
  *reinterpret_cast<uint4*>(p + split_bytes) = s1;
  *reinterpret_cast<uint4*>(p + 2 * split_bytes) = s2;
}

// weights -> split images.  src is row-major [rows x cols]; the image is [R x K] with zero padding.
__device__ __forceinline__ void stage_weight(uint8_t* img, uint32_t split_bytes, const float* __restrict__ src,
                                             int rows, int cols, int R, int K) {
  for (int e = threadIdx.x; e < R * K; e += blockDim.x) {
    const int r = e / K, k = e % K;
    float x = (r < rows && k < cols) ? src[r * cols + k] : 0.f;
    uint16_t* dst = reinterpret_cast<uint16_t*>(img) + ((size_t)(k >> 3) * R + r) * 8 + (k & 7);
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const uint32_t h = pack_bf16x2(x, 0.f) & 0xFFFFu;
      dst[(size_t)t * (split_bytes / 2)] = (uint16_t)h;
      x -= __uint_as_float(h << 16);
    }
  }
}

// One bf16x3 GEMM into TMEM: D (+)= A B^T, six split products, smallest first.  Called by ALL
// lanes of warp 0 (elect.sync inside).  Strides in bytes; *_kstep advances one K=16 step.
struct Operand {
  uint32_t addr, split, lbo, sbo, kstep;
};
__device__ __forceinline__ Operand op_kmajor(uint32_t addr, uint32_t split, int R) {
  return Operand{addr, split, (uint32_t)R * 16u, 128u, (uint32_t)R * 32u};
}
__device__ __forceinline__ Operand op_mnmajor(uint32_t addr, uint32_t split, int R) {
  return Operand{addr, split, 128u, (uint32_t)R * 16u, 256u};
}
__device__ __forceinline__ void gemm_bf16x3(uint32_t tmem_d, const Operand a, const Operand b, int nk,
                                            uint32_t idesc, bool accumulate) {
  uint32_t acc = accumulate ? 1u : 0u;
#pragma unroll
  for (int t = 0; t < 6; ++t) {
    constexpr int AI[6] = {0, 2, 1, 0, 1, 0}, BI[6] = {2, 0, 1, 1, 0, 0};
    const uint64_t ad0 = umma::smem_desc(a.addr + AI[t] * a.split, a.lbo, a.sbo);
    const uint64_t bd0 = umma::smem_desc(b.addr + BI[t] * b.split, b.lbo, b.sbo);
    for (int k = 0; k < nk; ++k) {
      umma::mma_bf16_ss(tmem_d, ad0 + (uint64_t)((k * a.kstep) >> 4), bd0 + (uint64_t)((k * b.kstep) >> 4),
                        idesc, acc);
      acc = 1u;
    }
  }
}

// GELU phase: the thread reads its 32 accumulator columns, writes gelu() as the next GEMM's
// A operand (chunks 4*half .. 4*half+3 of the tile at `img`) and, in the backward, replaces
// the accumulator by gelu'() for the data-gradient pass.
template <bool KEEP_GRAD>
__device__ __forceinline__ void gelu_phase(uint32_t taddr, uint8_t* img, int r, int half) {
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    float h[16];
    umma::tmem_ld16(taddr + 16 * q, h);
    umma::tmem_wait_ld();
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float cdf = 0.5f * (1.0f + erff(h[i] * 0.70710678118654752f));
      a[i] = h[i] * cdf;
      if (KEEP_GRAD) h[i] = cdf + h[i] * (0.39894228040143268f * __expf(-0.5f * h[i] * h[i]));
    }
    store_chunk(img, BIG_SPLIT, TP, 4 * half + 2 * q, r, a);
    store_chunk(img, BIG_SPLIT, TP, 4 * half + 2 * q + 1, r, a + 8);
    if (KEEP_GRAD) umma::tmem_st16(taddr + 16 * q, h);
  }
  if (KEEP_GRAD) umma::tmem_wait_st();
}
// dL/dh = dL/da * gelu'(h): both in TMEM; result becomes an A operand tile at `img`
__device__ __forceinline__ void grad_phase(uint32_t t_ga, uint32_t t_dgelu, uint8_t* img, int r, int half) {
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    float g[16], d[16];
    umma::tmem_ld16(t_ga + 16 * q, g);
    umma::tmem_ld16(t_dgelu + 16 * q, d);
    umma::tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) g[i] *= d[i];
    store_chunk(img, BIG_SPLIT, TP, 4 * half + 2 * q, r, g);
    store_chunk(img, BIG_SPLIT, TP, 4 * half + 2 * q + 1, r, g + 8);
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
__device__ __forceinline__ Setup block_setup(uint8_t* sm, uint64_t* mbar, uint32_t* tmem_slot,
                                             const float* W0, const float* W1, const float* W2,
                                             uint32_t tmem_cols) {
  stage_weight(sm + OFF_W0, W0_SPLIT, W0, H, NF, H, KF);
  stage_weight(sm + OFF_W1, W1_SPLIT, W1, H, H, H, H);
  stage_weight(sm + OFF_W2, W2_SPLIT, W2, 9, H, KF, H);
  if (threadIdx.x < 32) umma::tmem_alloc(umma::smem_u32(tmem_slot), tmem_cols);
  if (threadIdx.x == 32) {
    umma::mbar_init(umma::smem_u32(mbar), 1);
    umma::fence_mbar_init();
  }
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  return Setup{*tmem_slot, umma::smem_u32(sm)};
}

// publish this thread's smem operand writes / TMEM reads, then hand over to the MMA warp
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
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_slot;
  const Setup su = block_setup(sm, &mbar, &tmem_slot, W0, W1, W2, FWD_TMEM_COLS);
  const int tid = threadIdx.x, r = tid & (TP - 1), half = tid >> 7, warp = tid >> 5;
  const uint32_t tlane = su.tmem + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t phase = 0;
  const Operand w0 = op_kmajor(su.smem + OFF_W0, W0_SPLIT, H), w1 = op_kmajor(su.smem + OFF_W1, W1_SPLIT, H),
                w2 = op_kmajor(su.smem + OFF_W2, W2_SPLIT, KF);
  const Operand feat = op_kmajor(su.smem + OFF_FEAT, SMALL_SPLIT, TP),
                act = op_kmajor(su.smem + OFF_BIGA, BIG_SPLIT, TP);
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
      store_chunk(sm + OFF_FEAT, SMALL_SPLIT, TP, 0, r, f);
      store_chunk(sm + OFF_FEAT, SMALL_SPLIT, TP, 1, r, f + 8);
    }
    publish_and_sync();
    if (warp == 0) {
      umma::fence_after_sync();
      gemm_bf16x3(su.tmem + COL_H1, feat, w0, 1, ID64, false);
      umma::commit(umma::smem_u32(&mbar));
    }
    wait_mma(&mbar, phase);
    gelu_phase<false>(tlane + COL_H1 + 32 * half, sm + OFF_BIGA, r, half);
    publish_and_sync();
    if (warp == 0) {
      umma::fence_after_sync();
      gemm_bf16x3(su.tmem + COL_H1, act, w1, 4, ID64, false);
      umma::commit(umma::smem_u32(&mbar));
    }
    wait_mma(&mbar, phase);
    gelu_phase<false>(tlane + COL_H1 + 32 * half, sm + OFF_BIGA, r, half);
    publish_and_sync();
    if (warp == 0) {
      umma::fence_after_sync();
      gemm_bf16x3(su.tmem + FWD_COL_O, act, w2, 4, ID16, false);
      umma::commit(umma::smem_u32(&mbar));
    }
    wait_mma(&mbar, phase);
    if (half == 0) {
      float o[16];
      umma::tmem_ld16(tlane + FWD_COL_O, o);
      umma::tmem_wait_ld();
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
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_slot;
  const Setup su = block_setup(sm, &mbar, &tmem_slot, W0, W1, W2, BWD_TMEM_COLS);
  const int tid = threadIdx.x, r = tid & (TP - 1), half = tid >> 7, warp = tid >> 5;
  const uint32_t tlane = su.tmem + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t phase = 0;
  // forward views (K-major) and the transposed views of the same bytes (MN-major)
  const Operand w0 = op_kmajor(su.smem + OFF_W0, W0_SPLIT, H), w1 = op_kmajor(su.smem + OFF_W1, W1_SPLIT, H),
                w2 = op_kmajor(su.smem + OFF_W2, W2_SPLIT, KF);
  const Operand w0t = op_mnmajor(su.smem + OFF_W0, W0_SPLIT, H), w1t = op_mnmajor(su.smem + OFF_W1, W1_SPLIT, H),
                w2t = op_mnmajor(su.smem + OFF_W2, W2_SPLIT, KF);
  const Operand feat = op_kmajor(su.smem + OFF_FEAT, SMALL_SPLIT, TP), go = op_kmajor(su.smem + OFF_GO, SMALL_SPLIT, TP),
                biga = op_kmajor(su.smem + OFF_BIGA, BIG_SPLIT, TP), bigb = op_kmajor(su.smem + OFF_BIGB, BIG_SPLIT, TP);
  const Operand featT = op_mnmajor(su.smem + OFF_FEAT, SMALL_SPLIT, TP), goT = op_mnmajor(su.smem + OFF_GO, SMALL_SPLIT, TP),
                bigaT = op_mnmajor(su.smem + OFF_BIGA, BIG_SPLIT, TP), bigbT = op_mnmajor(su.smem + OFF_BIGB, BIG_SPLIT, TP);
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
    // ---- invariants -> h1
    if (half == 0) {
      float f[KF];
#pragma unroll
      for (int i = 0; i < KF; ++i) f[i] = 0.f;
      if (live) {
        kinematics(F + (size_t)p * 9, kind != 2, k);
#pragma unroll
        for (int i = 0; i < NF; ++i) f[i] = k.feat[i];
      }
      store_chunk(sm + OFF_FEAT, SMALL_SPLIT, TP, 0, r, f);
      store_chunk(sm + OFF_FEAT, SMALL_SPLIT, TP, 1, r, f + 8);
    }
    publish_and_sync();
    if (warp == 0) {
      umma::fence_after_sync();
      gemm_bf16x3(su.tmem + COL_H1, feat, w0, 1, ID64, false);
      umma::commit(umma::smem_u32(&mbar));
    }
    wait_mma(&mbar, phase);
    // ---- a1 = gelu(h1) -> h2
    gelu_phase<true>(tlane + COL_H1 + 32 * half, sm + OFF_BIGA, r, half);
    publish_and_sync();
    if (warp == 0) {
      umma::fence_after_sync();
      gemm_bf16x3(su.tmem + COL_H2, biga, w1, 4, ID64, false);
      umma::commit(umma::smem_u32(&mbar));
    }
    wait_mma(&mbar, phase);
    // ---- a2 = gelu(h2) -> o
    gelu_phase<true>(tlane + COL_H2 + 32 * half, sm + OFF_BIGB, r, half);
    publish_and_sync();
    if (warp == 0) {
      umma::fence_after_sync();
      gemm_bf16x3(su.tmem + COL_O, bigb, w2, 4, ID16, false);
      umma::commit(umma::smem_u32(&mbar));
    }
    wait_mma(&mbar, phase);
    // ---- output head: dL/dout -> dL/d(MLP output), dL/dR, direct dL/dF
    if (half == 0) {
      float o[16], g9[KF];
      umma::tmem_ld16(tlane + COL_O, o);
      umma::tmem_wait_ld();
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
      store_chunk(sm + OFF_GO, SMALL_SPLIT, TP, 0, r, g9);
      store_chunk(sm + OFF_GO, SMALL_SPLIT, TP, 1, r, g9 + 8);
    }
    publish_and_sync();
    if (warp == 0) {
      umma::fence_after_sync();
      gemm_bf16x3(su.tmem + COL_GA2, go, w2t, 1, ID64_BT, false);           // dL/da2 = go W2
      if (WGRAD) gemm_bf16x3(su.tmem + COL_GW2T, bigbT, goT, 8, IDW16, !first);  // dW2^T += a2^T go
      umma::commit(umma::smem_u32(&mbar));
    }
    wait_mma(&mbar, phase);
    // ---- dL/dh2 -> dL/da1 (over the h2 columns), dW1
    grad_phase(tlane + COL_GA2 + 32 * half, tlane + COL_H2 + 32 * half, sm + OFF_BIGB, r, half);
    publish_and_sync();
    if (warp == 0) {
      umma::fence_after_sync();
      gemm_bf16x3(su.tmem + COL_H2, bigb, w1t, 4, ID64_BT, false);          // dL/da1 = g2 W1
      if (WGRAD) gemm_bf16x3(su.tmem + COL_GW1, bigbT, bigaT, 8, IDW64, !first);  // dW1 += g2^T a1
      umma::commit(umma::smem_u32(&mbar));
    }
    wait_mma(&mbar, phase);
    // ---- dL/dh1 -> dL/dfeat (over the o columns), dW0
    grad_phase(tlane + COL_H2 + 32 * half, tlane + COL_H1 + 32 * half, sm + OFF_BIGA, r, half);
    publish_and_sync();
    if (warp == 0) {
      umma::fence_after_sync();
      gemm_bf16x3(su.tmem + COL_O, biga, w0t, 4, ID16_BT, false);           // dL/dfeat = g1 W0
      if (WGRAD) gemm_bf16x3(su.tmem + COL_GW0, bigaT, featT, 8, IDW16, !first);  // dW0 += g1^T feat
      umma::commit(umma::smem_u32(&mbar));
    }
    wait_mma(&mbar, phase);
    // ---- invariants -> F
    if (half == 0) {
      float gfeat[16];
      umma::tmem_ld16(tlane + COL_O, gfeat);
      umma::tmem_wait_ld();
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
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  if (WGRAD && !first && half == 0) {
    // M=64 accumulators: lane 32*q + i (i < 16) holds row 16*q + i
    const int i16 = tid & 31, row = (warp & 3) * 16 + i16;
    float a[16];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      umma::tmem_ld16(tlane + COL_GW1 + 16 * c, a);
      umma::tmem_wait_ld();
      if (i16 < 16)
#pragma unroll
        for (int j = 0; j < 16; ++j) atomicAdd(gW1 + row * H + 16 * c + j, a[j]);
    }
    umma::tmem_ld16(tlane + COL_GW0, a);
    umma::tmem_wait_ld();
    if (i16 < 16)
#pragma unroll
      for (int j = 0; j < NF; ++j) atomicAdd(gW0 + row * NF + j, a[j]);
    umma::tmem_ld16(tlane + COL_GW2T, a);
    umma::tmem_wait_ld();
    if (i16 < 16)
#pragma unroll
      for (int j = 0; j < 9; ++j) atomicAdd(gW2 + j * H + row, a[j]);
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
