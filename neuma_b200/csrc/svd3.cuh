// Per-particle 3x3 SVD for sm_100a, replacing wp.svd3 + the sign fix of
// modules/nclaw/warp/svd.py:59-96.
//
// Algorithm (own implementation of the published "Jacobi on A^T A, sort, QR"
// scheme of McAdams et al., which is what warp-lang's svd3 implements
// [UPSTREAM-UNVERIFIED]):
//   1. cyclic Jacobi sweeps with exact Givens rotations diagonalise S = A^T A,
//      accumulating the rotation V (det V = +1);
//   2. B = A V; columns sorted by descending norm (a swap negates one column so V
//      stays a rotation);
//   3. Gram-Schmidt on B gives U (u2 = u0 x u1, so det U = +1) and the singular
//      values on the diagonal: sigma0 >= sigma1 >= |sigma2|, with sigma2 carrying the
//      sign of det A.
// This is the convention batch_svd produces after its det<0 fix (both U and V
// rotations, third singular value signed), so that fix is the identity here.
#pragma once
#include "common.cuh"

template <int P, int Q>
__device__ __forceinline__ void jacobi_rotate(float S[3][3], float V[3][3]) {
  const float apq = S[P][Q];
  if (fabsf(apq) < 1e-37f) return;
  const float theta = (S[Q][Q] - S[P][P]) / (2.0f * apq);
  const float t = copysignf(1.0f, theta) / (fabsf(theta) + sqrtf(theta * theta + 1.0f));
  const float c = rsqrtf(t * t + 1.0f);
  const float s = t * c;
  constexpr int R = 3 - P - Q;
  S[P][P] -= t * apq;
  S[Q][Q] += t * apq;
  S[P][Q] = S[Q][P] = 0.0f;
  const float arp = S[R][P], arq = S[R][Q];
  S[R][P] = S[P][R] = c * arp - s * arq;
  S[R][Q] = S[Q][R] = s * arp + c * arq;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float vp = V[k][P], vq = V[k][Q];
    V[k][P] = c * vp - s * vq;
    V[k][Q] = s * vp + c * vq;
  }
}

template <int I, int J>
__device__ __forceinline__ void sort_swap(float B[3][3], float V[3][3], float n[3]) {
  if (n[I] < n[J]) {
    const float tn = n[I]; n[I] = n[J]; n[J] = tn;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float tb = B[k][I]; B[k][I] = B[k][J]; B[k][J] = -tb;
      float tv = V[k][I]; V[k][I] = V[k][J]; V[k][J] = -tv;
    }
  }
}

// sort + Gram-Schmidt tail shared by svd3 / svd3_fast: given V (rotation diagonalising
// A^T A) produce U, sigma and the column-sorted V.
__device__ __forceinline__ void svd3_finish(const Mat3& A, float V[3][3], Mat3& U, float sig[3], Mat3& Vout) {
  float B[3][3], n[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) B[i][j] = A(i, 0) * V[0][j] + A(i, 1) * V[1][j] + A(i, 2) * V[2][j];
#pragma unroll
  for (int j = 0; j < 3; ++j) n[j] = B[0][j] * B[0][j] + B[1][j] * B[1][j] + B[2][j] * B[2][j];
  sort_swap<0, 1>(B, V, n);
  sort_swap<0, 2>(B, V, n);
  sort_swap<1, 2>(B, V, n);
  // Gram-Schmidt
  float u0[3], u1[3], u2[3];
  const float s0 = sqrtf(n[0]);
  if (s0 > 1e-30f) {
    const float r = 1.0f / s0;
    u0[0] = B[0][0] * r; u0[1] = B[1][0] * r; u0[2] = B[2][0] * r;
  } else {
    u0[0] = 1.f; u0[1] = 0.f; u0[2] = 0.f;
  }
  const float d01 = u0[0] * B[0][1] + u0[1] * B[1][1] + u0[2] * B[2][1];
  float b1[3] = {B[0][1] - d01 * u0[0], B[1][1] - d01 * u0[1], B[2][1] - d01 * u0[2]};
  const float s1 = sqrtf(b1[0] * b1[0] + b1[1] * b1[1] + b1[2] * b1[2]);
  if (s1 > 1e-30f && s1 > 1e-7f * s0) {
    const float r = 1.0f / s1;
    u1[0] = b1[0] * r; u1[1] = b1[1] * r; u1[2] = b1[2] * r;
  } else {
    // rank <= 1: any unit vector orthogonal to u0
    const float ax = fabsf(u0[0]), ay = fabsf(u0[1]), az = fabsf(u0[2]);
    float e[3] = {0.f, 0.f, 0.f};
    if (ax <= ay && ax <= az) e[0] = 1.f; else if (ay <= az) e[1] = 1.f; else e[2] = 1.f;
    const float d = e[0] * u0[0] + e[1] * u0[1] + e[2] * u0[2];
    float t[3] = {e[0] - d * u0[0], e[1] - d * u0[1], e[2] - d * u0[2]};
    const float r = rsqrtf(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
    u1[0] = t[0] * r; u1[1] = t[1] * r; u1[2] = t[2] * r;
  }
  u2[0] = u0[1] * u1[2] - u0[2] * u1[1];
  u2[1] = u0[2] * u1[0] - u0[0] * u1[2];
  u2[2] = u0[0] * u1[1] - u0[1] * u1[0];
  sig[0] = s0;
  sig[1] = u1[0] * B[0][1] + u1[1] * B[1][1] + u1[2] * B[2][1];
  sig[2] = u2[0] * B[0][2] + u2[1] * B[1][2] + u2[2] * B[2][2];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    U(k, 0) = u0[k]; U(k, 1) = u1[k]; U(k, 2) = u2[k];
#pragma unroll
    for (int j = 0; j < 3; ++j) Vout(k, j) = V[k][j];
  }
}

// A = U diag(sig) V^T, U and V rotations.
__device__ __forceinline__ void svd3(const Mat3& A, Mat3& U, float sig[3], Mat3& Vout) {
  float S[3][3], V[3][3] = {{1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, {0.f, 0.f, 1.f}};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) S[i][j] = A(0, i) * A(0, j) + A(1, i) * A(1, j) + A(2, i) * A(2, j);
#pragma unroll 1
  for (int sweep = 0; sweep < 6; ++sweep) {
    jacobi_rotate<0, 1>(S, V);
    jacobi_rotate<0, 2>(S, V);
    jacobi_rotate<1, 2>(S, V);
  }
  svd3_finish(A, V, U, sig, Vout);
}

// ---- the fused constitutive kernels' variant: same algorithm, but
//   * the Givens parameters use the MUFU approximations (rcp / sqrt / rsqrt .approx): the
//     rotation ANGLE only steers convergence; c^2 + s^2 = 1 holds to ~2 ulp per rotation;
//   * sweeps stop as soon as every lane of the warp has converged (quadratic convergence:
//     3 sweeps for F within ~0.1 of a rotation, 6 at most).
// Must be called by all 32 lanes of a warp.
__device__ __forceinline__ float mufu_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float mufu_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float mufu_rsqrt(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
template <int P, int Q>
__device__ __forceinline__ void jacobi_rotate_fast(float S[3][3], float V[3][3]) {
  const float apq = S[P][Q];
  const bool skip = fabsf(apq) < 1e-37f;
  const float theta = (S[Q][Q] - S[P][P]) * mufu_rcp(skip ? 2.0f : 2.0f * apq);
  float t = copysignf(mufu_rcp(fabsf(theta) + mufu_sqrt(fmaf(theta, theta, 1.0f))), theta);
  t = skip ? 0.0f : t;
  const float c = mufu_rsqrt(fmaf(t, t, 1.0f));
  const float s = t * c;
  constexpr int R = 3 - P - Q;
  S[P][P] -= t * apq;
  S[Q][Q] += t * apq;
  S[P][Q] = S[Q][P] = 0.0f;
  const float arp = S[R][P], arq = S[R][Q];
  S[R][P] = S[P][R] = c * arp - s * arq;
  S[R][Q] = S[Q][R] = s * arp + c * arq;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float vp = V[k][P], vq = V[k][Q];
    V[k][P] = c * vp - s * vq;
    V[k][Q] = s * vp + c * vq;
  }
}
__device__ __forceinline__ void svd3_fast(const Mat3& A, Mat3& U, float sig[3], Mat3& Vout) {
  float S[3][3], V[3][3] = {{1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, {0.f, 0.f, 1.f}};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) S[i][j] = A(0, i) * A(0, j) + A(1, i) * A(1, j) + A(2, i) * A(2, j);
#pragma unroll 1
  for (int sweep = 0; sweep < 6; ++sweep) {
    jacobi_rotate_fast<0, 1>(S, V);
    jacobi_rotate_fast<0, 2>(S, V);
    jacobi_rotate_fast<1, 2>(S, V);
    const float off = fabsf(S[0][1]) + fabsf(S[0][2]) + fabsf(S[1][2]);
    const float tr = S[0][0] + S[1][1] + S[2][2];
    if (__all_sync(0xffffffffu, off <= 1e-7f * tr)) break;
  }
  svd3_finish(A, V, U, sig, Vout);
}

__device__ __forceinline__ float signed_floor(float d, float eps) {
  // keep the sign, bound the magnitude away from 0
  return (fabsf(d) < eps) ? copysignf(eps, d) : d;
}
