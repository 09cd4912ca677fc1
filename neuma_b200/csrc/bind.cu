// Particle -> Gaussian-kernel binding and covariance deformation.
//   neuma_bind_spmm  : compute_bindings_xyz / compute_bindings_F,
//                      modules/tune/utils.py:424-472 (torch.sparse.mm of a COO
//                      [K x Np] matrix with <= 10 uniform-weight entries per row)
//   neuma_deform_cov : deform_cov_by_F, modules/d3gs/utils/simulation_utils.py:25-48
#include "common.cuh"

namespace {

// forward: one thread per (row, column) -- rows have <= ~10 entries, ncol is 3 or 9.
__global__ void __launch_bounds__(256)
spmm_csr_kernel(int K, int ncol, const int* __restrict__ rowptr, const int* __restrict__ col,
                const float* __restrict__ w, const float* __restrict__ in,
                float* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)K * ncol) return;
  const int r = (int)(t / ncol), cidx = (int)(t % ncol);
  float acc = 0.f;
  for (int e = rowptr[r]; e < rowptr[r + 1]; ++e) acc += w[e] * in[(size_t)col[e] * ncol + cidx];
  out[t] = acc;
}

// transpose (the backward of the above): scatter rows of `in` [K,ncol] into
// out [Np,ncol] (zero-filled by the caller side of this entry point).
__global__ void __launch_bounds__(256)
spmm_csr_t_kernel(int K, int ncol, const int* __restrict__ rowptr, const int* __restrict__ col,
                  const float* __restrict__ w, const float* __restrict__ in,
                  float* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)K * ncol) return;
  const int r = (int)(t / ncol), cidx = (int)(t % ncol);
  const float g = in[t];
  for (int e = rowptr[r]; e < rowptr[r + 1]; ++e)
    atomicAdd(out + (size_t)col[e] * ncol + cidx, w[e] * g);
}

__global__ void __launch_bounds__(256)
deform_cov_kernel(int K, const float* __restrict__ cov6, const float* __restrict__ F,
                  float* __restrict__ out6) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= K) return;
  const float* c = cov6 + (size_t)t * 6;
  Mat3 S;
  S(0, 0) = c[0]; S(0, 1) = c[1]; S(0, 2) = c[2];
  S(1, 0) = c[1]; S(1, 1) = c[3]; S(1, 2) = c[4];
  S(2, 0) = c[2]; S(2, 1) = c[4]; S(2, 2) = c[5];
  const Mat3 Fm = mat3_load(F + (size_t)t * 9);
  const Mat3 R = mat3_mul_nt(mat3_mul(Fm, S), Fm);  // F S F^T
  float* o = out6 + (size_t)t * 6;
  o[0] = R(0, 0); o[1] = R(0, 1); o[2] = R(0, 2);
  o[3] = R(1, 1); o[4] = R(1, 2); o[5] = R(2, 2);
}

// ---------------------------------------------------------------- particles -> Gaussians, fused
// One kernel for the whole sim -> render hand-over of a frame (tune/utils.py:424-472 +
// nclaw/utils.py:110-118 + simulation_utils.py:25-48):
//   de_x      = (x - center) / size                         denormalize_points_helper_func
//   means3D_k = k_prev_k + sum_e w_e (de_x[c_e] - de_x_prev[c_e])     compute_bindings_xyz
//   F_k       = sum_e w_e F[c_e]                                       compute_bindings_F
//   cov'_k    = F_k Sigma_k F_k^T                                      deform_cov_by_F
// `slot` (may be NULL) maps a caller-order particle index to its row in x / F (the engine keeps
// the particles in cell order); de_x_prev is in caller order.
struct Vec3 {
  float v[3];
};
__global__ void __launch_bounds__(128)
particles_to_gaussians_kernel(int K, const int* __restrict__ rowptr, const int* __restrict__ col,
                              const float* __restrict__ w, const int* __restrict__ slot,
                              const float* __restrict__ x, const float* __restrict__ F, Vec3 center,
                              Vec3 size, const float* __restrict__ de_x_prev,
                              const float* __restrict__ k_prev, const float* __restrict__ cov6,
                              float* __restrict__ means3D, float* __restrict__ cov_out,
                              float* __restrict__ Fk_out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float dx[3] = {0.f, 0.f, 0.f};
  Mat3 Fk;
#pragma unroll
  for (int i = 0; i < 9; ++i) Fk.m[i] = 0.f;
  for (int e = rowptr[k]; e < rowptr[k + 1]; ++e) {
    const int c = col[e];
    const int r = slot ? slot[c] : c;
    const float we = w[e];
#pragma unroll
    for (int a = 0; a < 3; ++a)
      dx[a] += we * ((x[3 * (size_t)r + a] - center.v[a]) / size.v[a] - de_x_prev[3 * (size_t)c + a]);
#pragma unroll
    for (int i = 0; i < 9; ++i) Fk.m[i] += we * F[9 * (size_t)r + i];
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) means3D[3 * (size_t)k + a] = k_prev[3 * (size_t)k + a] + dx[a];
  const float* c6 = cov6 + (size_t)k * 6;
  Mat3 S;
  S(0, 0) = c6[0]; S(0, 1) = c6[1]; S(0, 2) = c6[2];
  S(1, 0) = c6[1]; S(1, 1) = c6[3]; S(1, 2) = c6[4];
  S(2, 0) = c6[2]; S(2, 1) = c6[4]; S(2, 2) = c6[5];
  const Mat3 R = mat3_mul_nt(mat3_mul(Fk, S), Fk);
  float* o = cov_out + (size_t)k * 6;
  o[0] = R(0, 0); o[1] = R(0, 1); o[2] = R(0, 2);
  o[3] = R(1, 1); o[4] = R(1, 2); o[5] = R(2, 2);
  if (Fk_out) mat3_store(Fk_out + (size_t)k * 9, Fk);
}

// the transpose: dL/dmeans3D [K,3] -> dL/dx accumulated into gx [Np,3] (rows = slot[c]; the caller
// zero-fills).  cov' carries no gradient.
__global__ void __launch_bounds__(128)
gaussians_to_particles_kernel(int K, const int* __restrict__ rowptr, const int* __restrict__ col,
                              const float* __restrict__ w, const int* __restrict__ slot,
                              const float* __restrict__ g_means3D, Vec3 size, float* __restrict__ gx) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float g[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) g[a] = g_means3D[3 * (size_t)k + a] / size.v[a];
  for (int e = rowptr[k]; e < rowptr[k + 1]; ++e) {
    const int c = col[e];
    const int r = slot ? slot[c] : c;
    const float we = w[e];
#pragma unroll
    for (int a = 0; a < 3; ++a) atomicAdd(gx + 3 * (size_t)r + a, we * g[a]);
  }
}

// ---------------------------------------------------------------- binding construction
// gaussian_binding_with_clip_v1, modules/d3gs/utils/binding_utils.py:199-285: for every kernel k
// the particles with Mahalanobis distance p = d^T Sigma^-1 d <= chi2 (test_point_in_gaussians_with_
// inv_cov_warp, :88-120), at most `max_p` of them, those with the smallest p.  The reference runs
// a Python loop of K launches over all Np particles and fills a dense K x Np matrix; here one
// warp owns one kernel, streams the particles once and keeps per-lane sorted candidate lists
// that are merged with warp shuffles at the end.
constexpr int BIND_MAX = 16;

__global__ void __launch_bounds__(256)
bind_build_kernel(int K, int Np, const float* __restrict__ means, const float* __restrict__ cov6,
                  const float* __restrict__ x, float threshold, int max_p, int* __restrict__ counts,
                  int* __restrict__ cols, float* __restrict__ pvals) {
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (k >= K) return;
  // Sigma^-1 (compute_inv_cov, binding_utils.py:22-45): adjugate / determinant of the symmetric 3x3
  const float* c = cov6 + (size_t)k * 6;
  const float a00 = c[0], a01 = c[1], a02 = c[2], a11 = c[3], a12 = c[4], a22 = c[5];
  const float c00 = a11 * a22 - a12 * a12, c01 = a02 * a12 - a01 * a22, c02 = a01 * a12 - a02 * a11;
  const float c11 = a00 * a22 - a02 * a02, c12 = a01 * a02 - a00 * a12, c22 = a00 * a11 - a01 * a01;
  const float idet = 1.0f / (a00 * c00 + a01 * c01 + a02 * c02);
  const float i00 = c00 * idet, i01 = c01 * idet, i02 = c02 * idet, i11 = c11 * idet, i12 = c12 * idet,
              i22 = c22 * idet;
  const float m0 = means[3 * k], m1 = means[3 * k + 1], m2 = means[3 * k + 2];
  float bp[BIND_MAX];
  int bi[BIND_MAX];
#pragma unroll
  for (int i = 0; i < BIND_MAX; ++i) { bp[i] = 3.0e38f; bi[i] = -1; }
  int inside = 0;
  for (int q = lane; q < Np; q += 32) {
    const float d0 = x[3 * q] - m0, d1 = x[3 * q + 1] - m1, d2 = x[3 * q + 2] - m2;
    const float p1 = d0 * i00 + d1 * i01 + d2 * i02;
    const float p2 = d0 * i01 + d1 * i11 + d2 * i12;
    const float p3 = d0 * i02 + d1 * i12 + d2 * i22;
    const float pv = p1 * d0 + p2 * d1 + p3 * d2;
    if (pv <= threshold) {
      ++inside;
      if (pv < bp[BIND_MAX - 1]) {  // sorted insert (rare path: a kernel holds a few dozen particles)
        float cp = pv;
        int ci = q;
#pragma unroll
        for (int i = 0; i < BIND_MAX; ++i) {
          if (cp < bp[i]) {
            const float tp = bp[i]; bp[i] = cp; cp = tp;
            const int ti = bi[i]; bi[i] = ci; ci = ti;
          }
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) inside += __shfl_xor_sync(0xffffffffu, inside, o);
  if (lane == 0) counts[k] = inside;
  // merge: max_p rounds of "smallest head over the 32 lanes"
  for (int r = 0; r < max_p; ++r) {
    float hp = bp[0];
    int hl = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float op = __shfl_xor_sync(0xffffffffu, hp, o);
      const int ol = __shfl_xor_sync(0xffffffffu, hl, o);
      if (op < hp || (op == hp && ol < hl)) { hp = op; hl = ol; }
    }
    const int widx = __shfl_sync(0xffffffffu, bi[0], hl);
    if (lane == 0) {
      cols[(size_t)k * max_p + r] = (hp < 3.0e38f) ? widx : -1;
      if (pvals) pvals[(size_t)k * max_p + r] = hp;
    }
    if (lane == hl) {  // pop
#pragma unroll
      for (int i = 0; i + 1 < BIND_MAX; ++i) { bp[i] = bp[i + 1]; bi[i] = bi[i + 1]; }
      bp[BIND_MAX - 1] = 3.0e38f;
      bi[BIND_MAX - 1] = -1;
    }
  }
}

}  // namespace

extern "C" {

int neuma_particles_to_gaussians(neuma_ctx* ctx, int K, const int32_t* rowptr, const int32_t* col,
                                 const float* w, const int32_t* slot, const float* x, const float* F,
                                 const float* center_host, const float* size_host,
                                 const float* de_x_prev, const float* k_prev, const float* cov6,
                                 float* means3D, float* cov_out, float* Fk_out, void* stream) {
  if (!ctx) return 2;
  if (K <= 0) return 0;
  Vec3 c, s;
  for (int a = 0; a < 3; ++a) { c.v[a] = center_host ? center_host[a] : 0.f; s.v[a] = size_host ? size_host[a] : 1.f; }
  particles_to_gaussians_kernel<<<(K + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      K, rowptr, col, w, slot, x, F, c, s, de_x_prev, k_prev, cov6, means3D, cov_out, Fk_out);
  NEUMA_LAUNCH_CHECK(ctx, "particles_to_gaussians_kernel");
  return 0;
}

int neuma_gaussians_to_particles(neuma_ctx* ctx, int K, const int32_t* rowptr, const int32_t* col,
                                 const float* w, const int32_t* slot, const float* g_means3D,
                                 const float* size_host, float* gx, void* stream) {
  if (!ctx) return 2;
  if (K <= 0) return 0;
  Vec3 s;
  for (int a = 0; a < 3; ++a) s.v[a] = size_host ? size_host[a] : 1.f;
  gaussians_to_particles_kernel<<<(K + 127) / 128, 128, 0, (cudaStream_t)stream>>>(K, rowptr, col, w, slot,
                                                                                  g_means3D, s, gx);
  NEUMA_LAUNCH_CHECK(ctx, "gaussians_to_particles_kernel");
  return 0;
}

int neuma_bind_build(neuma_ctx* ctx, int K, int Np, const float* means, const float* cov6,
                     const float* x, float threshold, int max_particles, int32_t* counts,
                     int32_t* cols, float* pvals, void* stream) {
  if (!ctx) return 2;
  if (max_particles < 1 || max_particles > BIND_MAX)
    return neuma_fail_msg(ctx, "neuma_bind_build: max_particles must be in [1, 16]");
  if (K <= 0) return 0;
  const long long threads = (long long)K * 32;
  bind_build_kernel<<<(int)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      K, Np, means, cov6, x, threshold, max_particles, counts, cols, pvals);
  NEUMA_LAUNCH_CHECK(ctx, "bind_build_kernel");
  return 0;
}

int neuma_bind_spmm(neuma_ctx* ctx, int K, int Np, const int32_t* rowptr, const int32_t* col,
                    const float* w, int ncol, const float* in, float* out, int transpose,
                    void* stream) {
  if (!ctx) return 2;
  if (K < 0 || Np < 0 || ncol <= 0) return neuma_fail_msg(ctx, "neuma_bind_spmm: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  if (transpose) {
    NEUMA_CUDA(ctx, cudaMemsetAsync(out, 0, (size_t)Np * ncol * sizeof(float), st));
  }
  if (K == 0) return 0;
  const long long total = (long long)K * ncol;
  const int blocks = (int)((total + 255) / 256);
  if (!transpose)
    spmm_csr_kernel<<<blocks, 256, 0, st>>>(K, ncol, rowptr, col, w, in, out);
  else
    spmm_csr_t_kernel<<<blocks, 256, 0, st>>>(K, ncol, rowptr, col, w, in, out);
  NEUMA_LAUNCH_CHECK(ctx, "spmm_csr_kernel");
  return 0;
}

int neuma_deform_cov(neuma_ctx* ctx, int K, const float* cov6, const float* F, float* out6,
                     void* stream) {
  if (!ctx) return 2;
  if (K <= 0) return 0;
  deform_cov_kernel<<<(K + 255) / 256, 256, 0, (cudaStream_t)stream>>>(K, cov6, F, out6);
  NEUMA_LAUNCH_CHECK(ctx, "deform_cov_kernel");
  return 0;
}

}  // extern "C"
