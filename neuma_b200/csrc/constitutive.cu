// Neural constitutive law, fused: SVD -> invariants -> 13->H->H->9 bias-free GELU
// MLP -> symmetrise -> polar recombination, forward and backward, one kernel each.
// Replaces ~35 torch launches + 1 Warp launch per call of
//   InvariantFullMetaElasticity.forward   modules/nclaw/material/meta.py:196-221
//   InvariantFullMetaPlasticity.forward   modules/nclaw/material/meta.py:468-489
// and the Warp batch_svd / SVDFunction of modules/nclaw/warp/svd.py:9-96.
//
// This file is the fp32 CUDA-core implementation (bit-for-bit fp32 arithmetic,
// weights broadcast from shared memory); LoRA adaptors are merged into the
// weights by the host (W + scaling * B A, loralib.py:216-224) and the weight
// gradients un-merged there.
//
// Gradient of the SVD part: everything downstream of the SVD depends on U, V only
// through R = U V^T, so the backward uses the polar-decomposition form
//   dF += U [ (M - M^T)_ij / (sigma_i + sigma_j) ] V^T,  M = U^T (dL/dR) V
// which is the exact sum of the dU and dV terms of the textbook SVD adjoint but has
// no 1/(sigma_i^2 - sigma_j^2) pole at F ~ I (SURVEY 7 "SVD-gradient degeneracy").
#include "common.cuh"
#include "svd3.cuh"

namespace {

constexpr int H = 64;      // hidden width (layer_widths: [64, 64] in every shipped config)
constexpr int NF = 13;     // 3 + 9 + 1 invariants
constexpr int TP = 128;    // particles per tile = threads per block
constexpr int LD = TP + 1; // padded row stride of the [feature][particle] smem tiles
constexpr int NW = NF * H + H * H + 9 * H;  // 5504 weights per net

__device__ __forceinline__ float gelu(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
}
__device__ __forceinline__ float gelu_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

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

// stage the weights: W0 [H,13] -> W0t [13][H]; W1 [H,H] -> W1t [k][o]; W2 [9,H] as is.
__device__ __forceinline__ void stage_weights(const float* __restrict__ W0,
                                              const float* __restrict__ W1,
                                              const float* __restrict__ W2, float* W0t,
                                              float* W1t, float* W2s) {
  for (int i = threadIdx.x; i < H * NF; i += blockDim.x) W0t[(i % NF) * H + (i / NF)] = W0[i];
  for (int i = threadIdx.x; i < H * H; i += blockDim.x) W1t[(i % H) * H + (i / H)] = W1[i];
  for (int i = threadIdx.x; i < 9 * H; i += blockDim.x) W2s[i] = W2[i];
}

// acc[o] = sum_k Wt[k][o] * in[k],  in[] in registers (k loop unrolled)
template <int K>
__device__ __forceinline__ void dense_regs(const float* __restrict__ Wt, const float (&in)[K],
                                           float (&acc)[H]) {
#pragma unroll
  for (int o = 0; o < H; ++o) acc[o] = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const float a = in[k];
    const float4* w4 = reinterpret_cast<const float4*>(Wt + k * H);
#pragma unroll
    for (int o = 0; o < H / 4; ++o) {
      const float4 w = w4[o];
      acc[4 * o + 0] += w.x * a; acc[4 * o + 1] += w.y * a;
      acc[4 * o + 2] += w.z * a; acc[4 * o + 3] += w.w * a;
    }
  }
}
// acc[o] = sum_k Wt[k][o] * col[k*LD]   (input read from this thread's smem column)
__device__ __forceinline__ void dense_smem(const float* __restrict__ Wt,
                                           const float* __restrict__ col, float (&acc)[H]) {
#pragma unroll
  for (int o = 0; o < H; ++o) acc[o] = 0.f;
#pragma unroll 2
  for (int k = 0; k < H; ++k) {
    const float a = col[k * LD];
    const float4* w4 = reinterpret_cast<const float4*>(Wt + k * H);
#pragma unroll
    for (int o = 0; o < H / 4; ++o) {
      const float4 w = w4[o];
      acc[4 * o + 0] += w.x * a; acc[4 * o + 1] += w.y * a;
      acc[4 * o + 2] += w.z * a; acc[4 * o + 3] += w.w * a;
    }
  }
}

__device__ __forceinline__ Mat3 mlp_head(const float* __restrict__ W2s, const float (&a2)[H]) {
  float o9[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    float s = 0.f;
    const float4* w4 = reinterpret_cast<const float4*>(W2s + j * H);
#pragma unroll
    for (int o = 0; o < H / 4; ++o) {
      const float4 w = w4[o];
      s += w.x * a2[4 * o] + w.y * a2[4 * o + 1] + w.z * a2[4 * o + 2] + w.w * a2[4 * o + 3];
    }
    o9[j] = s;
  }
  Mat3 X;  // x = 0.5 (x^T + x)
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) X(i, j) = 0.5f * (o9[3 * i + j] + o9[3 * j + i]);
  return X;
}

// ------------------------------------------------------------------ forward
__global__ void __launch_bounds__(TP)
constitutive_fwd_kernel(int kind, int n, const float* __restrict__ F, const float* __restrict__ W0,
                        const float* __restrict__ W1, const float* __restrict__ W2, float alpha,
                        float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  float* W0t = sm;
  float* W1t = W0t + NF * H;
  float* W2s = W1t + H * H;
  float* A1 = W2s + 9 * H;  // [H][LD]
  stage_weights(W0, W1, W2, W0t, W1t, W2s);
  __syncthreads();
  const int tid = threadIdx.x;
  const int tiles = (n + TP - 1) / TP;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int p = tile * TP + tid;
    if (p >= n) continue;  // no block-level sync inside the tile loop
    Kinematics k;
    kinematics(F + (size_t)p * 9, kind != 2, k);
    float acc[H];
    dense_regs<NF>(W0t, k.feat, acc);
    float* col = A1 + tid;
#pragma unroll
    for (int o = 0; o < H; ++o) col[o * LD] = gelu(acc[o]);
    dense_smem(W1t, col, acc);
#pragma unroll
    for (int o = 0; o < H; ++o) acc[o] = gelu(acc[o]);
    const Mat3 X = mlp_head(W2s, acc);
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

// ------------------------------------------------------------------ backward
template <bool WGRAD>
__global__ void __launch_bounds__(TP, 1)
constitutive_bwd_kernel(int kind, int n, const float* __restrict__ F, const float* __restrict__ W0,
                        const float* __restrict__ W1, const float* __restrict__ W2, float alpha,
                        const float* __restrict__ gout, float* __restrict__ gF,
                        float* __restrict__ gW0, float* __restrict__ gW1,
                        float* __restrict__ gW2) {
  extern __shared__ __align__(16) float sm[];
  float* W0t = sm;
  float* W1t = W0t + NF * H;
  float* W2s = W1t + H * H;
  float* GW = W2s + 9 * H;        // [NW] block-local weight-gradient accumulator
  float* A1 = GW + NW;            // [H][LD]  gelu(h1)
  float* H1 = A1 + H * LD;        // [H][LD]  h1, then dL/dh1
  float* H2 = H1 + H * LD;        // [H][LD]  h2, then dL/dh2
  float* A2 = H2 + H * LD;        // [H][LD]  gelu(h2)
  float* FT = A2 + H * LD;        // [NF][LD] invariants
  float* GO = FT + NF * LD;       // [9][LD]  dL/d(MLP output)
  stage_weights(W0, W1, W2, W0t, W1t, W2s);
  if (WGRAD)
    for (int i = threadIdx.x; i < NW; i += blockDim.x) GW[i] = 0.f;
  __syncthreads();
  const int tid = threadIdx.x;
  const int tiles = (n + TP - 1) / TP;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int p = tile * TP + tid;
    if (p < n) {
      Kinematics k;
      kinematics(F + (size_t)p * 9, kind != 2, k);
#pragma unroll
      for (int j = 0; j < NF; ++j) FT[j * LD + tid] = k.feat[j];
      float acc[H];
      dense_regs<NF>(W0t, k.feat, acc);
#pragma unroll
      for (int o = 0; o < H; ++o) {
        H1[o * LD + tid] = acc[o];
        A1[o * LD + tid] = gelu(acc[o]);
      }
      dense_smem(W1t, A1 + tid, acc);
#pragma unroll
      for (int o = 0; o < H; ++o) {
        H2[o * LD + tid] = acc[o];
        acc[o] = gelu(acc[o]);
        A2[o * LD + tid] = acc[o];
      }
      const Mat3 X = mlp_head(W2s, acc);
      const Mat3 P = mat3_mul(k.R, X);
      const Mat3 G = mat3_load(gout + (size_t)p * 9);
      // out = P F^T (elastic)  |  out = F + alpha P (plastic)
      Mat3 gP, gFd;
      if (kind == 1) {
#pragma unroll
        for (int i = 0; i < 9; ++i) { gP.m[i] = alpha * G.m[i]; gFd.m[i] = G.m[i]; }
      } else {
        gP = mat3_mul(G, k.F);
        gFd = mat3_mul_tn(G, P);
      }
      const Mat3 gR = mat3_mul(gP, X);       // X symmetric
      const Mat3 gXs = mat3_mul_tn(k.R, gP);
      float go[9];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) go[3 * i + j] = 0.5f * (gXs(i, j) + gXs(j, i));
#pragma unroll
      for (int j = 0; j < 9; ++j) GO[j * LD + tid] = go[j];
      // dL/da2 = W2^T go ; dL/dh2 = . * gelu'(h2)
      float g2[H];
#pragma unroll
      for (int o = 0; o < H; ++o) g2[o] = 0.f;
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const float4* w4 = reinterpret_cast<const float4*>(W2s + j * H);
#pragma unroll
        for (int o = 0; o < H / 4; ++o) {
          const float4 w = w4[o];
          g2[4 * o] += w.x * go[j]; g2[4 * o + 1] += w.y * go[j];
          g2[4 * o + 2] += w.z * go[j]; g2[4 * o + 3] += w.w * go[j];
        }
      }
#pragma unroll
      for (int o = 0; o < H; ++o) {
        g2[o] *= gelu_grad(H2[o * LD + tid]);
        H2[o * LD + tid] = g2[o];
      }
      // dL/da1[k] = sum_o W1[o][k] g2[o] ; dL/dh1 ; dL/dfeat = W0^T dL/dh1
      float gfeat[NF];
#pragma unroll
      for (int j = 0; j < NF; ++j) gfeat[j] = 0.f;
#pragma unroll 2
      for (int kk = 0; kk < H; ++kk) {
        const float4* w4 = reinterpret_cast<const float4*>(W1t + kk * H);
        float s = 0.f;
#pragma unroll
        for (int o = 0; o < H / 4; ++o) {
          const float4 w = w4[o];
          s += w.x * g2[4 * o] + w.y * g2[4 * o + 1] + w.z * g2[4 * o + 2] + w.w * g2[4 * o + 3];
        }
        const float g1 = s * gelu_grad(H1[kk * LD + tid]);
        H1[kk * LD + tid] = g1;
#pragma unroll
        for (int j = 0; j < NF; ++j) gfeat[j] += W0t[j * H + kk] * g1;
      }
      // invariants -> F
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
    } else if (WGRAD) {
      // tail of the last tile: contribute exact zeros to the weight-gradient GEMMs
      for (int o = 0; o < H; ++o) {
        A1[o * LD + tid] = 0.f; H1[o * LD + tid] = 0.f;
        H2[o * LD + tid] = 0.f; A2[o * LD + tid] = 0.f;
      }
      for (int j = 0; j < NF; ++j) FT[j * LD + tid] = 0.f;
      for (int j = 0; j < 9; ++j) GO[j * LD + tid] = 0.f;
    }
    if (WGRAD) {
      __syncthreads();
      {  // gW1[o][k] += sum_p dh2[o][p] a1[k][p] ; thread owns an 8 x 4 patch
        const int o0 = (tid / 16) * 8, k0 = (tid % 16) * 4;
        float a[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) a[i][j] = 0.f;
#pragma unroll 4
        for (int q = 0; q < TP; ++q) {
          float l[8], r[4];
#pragma unroll
          for (int i = 0; i < 8; ++i) l[i] = H2[(o0 + i) * LD + q];
#pragma unroll
          for (int j = 0; j < 4; ++j) r[j] = A1[(k0 + j) * LD + q];
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) a[i][j] += l[i] * r[j];
        }
        float* g = GW + NF * H;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) g[(o0 + i) * H + k0 + j] += a[i][j];
      }
      // gW2[j][k] += sum_p go[j][p] a2[k][p]
      for (int e = tid; e < 9 * H; e += TP) {
        const int j = e / H, kk = e % H;
        float s = 0.f;
#pragma unroll 8
        for (int q = 0; q < TP; ++q) s += GO[j * LD + q] * A2[kk * LD + q];
        GW[NF * H + H * H + e] += s;
      }
      // gW0[o][j] += sum_p dh1[o][p] feat[j][p]
      for (int e = tid; e < H * NF; e += TP) {
        const int o = e / NF, j = e % NF;
        float s = 0.f;
#pragma unroll 8
        for (int q = 0; q < TP; ++q) s += H1[o * LD + q] * FT[j * LD + q];
        GW[e] += s;
      }
      __syncthreads();
    }
  }
  if (WGRAD) {
    __syncthreads();
    for (int i = tid; i < NF * H; i += TP) atomicAdd(gW0 + i, GW[i]);
    for (int i = tid; i < H * H; i += TP) atomicAdd(gW1 + i, GW[NF * H + i]);
    for (int i = tid; i < 9 * H; i += TP) atomicAdd(gW2 + i, GW[NF * H + H * H + i]);
  }
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

constexpr size_t FWD_SMEM = (size_t)(NW + H * LD) * sizeof(float);
constexpr size_t BWD_SMEM = (size_t)(2 * NW + 4 * H * LD + NF * LD + 9 * LD) * sizeof(float);
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
  const int grid = tiles < ctx->num_sms * 4 ? tiles : ctx->num_sms * 4;
  constitutive_fwd_kernel<<<grid, TP, FWD_SMEM, (cudaStream_t)stream>>>(kind, n, F, W0, W1, W2,
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
    constitutive_bwd_kernel<true><<<grid, TP, BWD_SMEM, st>>>(kind, n, F, W0, W1, W2, alpha, gout,
                                                              gF, gW0, gW1, gW2);
  else
    constitutive_bwd_kernel<false><<<grid, TP, BWD_SMEM, st>>>(kind, n, F, W0, W1, W2, alpha, gout,
                                                               gF, gW0, gW1, gW2);
  NEUMA_LAUNCH_CHECK(ctx, "constitutive_bwd_kernel");
  return 0;
}

}  // extern "C"
