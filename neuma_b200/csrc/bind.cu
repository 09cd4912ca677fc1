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

}  // namespace

extern "C" {

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
