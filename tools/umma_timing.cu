// Timing probe for small tcgen05.mma (kind::f16, bf16) shapes on sm_100a: cycles per MMA for
// chains that accumulate into the same TMEM columns vs. round-robin over several accumulators.
// This is what sizes the constitutive kernels' GEMM decomposition (csrc/constitutive.cu).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/umma_timing tools/umma_timing.cu && /tmp/umma_timing
#include <cuda_runtime.h>
#include <stdio.h>
#include "../neuma_b200/csrc/umma.cuh"

struct Cfg {
  uint32_t idesc;
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo, a_kstep, b_kstep;
  int L;        // MMAs in the chain
  int ndist;    // distinct accumulators, round-robin
  int dstride;  // TMEM columns between accumulators
};

__global__ void __launch_bounds__(128) timing(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tbase_s;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x < 32) umma::tmem_alloc(umma::smem_u32(&tbase_s), 512);
  if (threadIdx.x == 0) { umma::mbar_init(umma::smem_u32(&mbar), 1); umma::fence_mbar_init(); }
  umma::fence_async_smem(); umma::fence_before_sync(); __syncthreads(); umma::fence_after_sync();
  const uint32_t tb = tbase_s, s0 = umma::smem_u32(sm);
  if (threadIdx.x < 32) {
    uint32_t phase = 0;
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      for (int i = 0; i < c.L; ++i) {
        const uint64_t ad = umma::smem_desc(s0 + (i & 3) * c.a_kstep, c.a_lbo, c.a_sbo);
        const uint64_t bd = umma::smem_desc(s0 + 65536 + (i & 3) * c.b_kstep, c.b_lbo, c.b_sbo);
        umma::mma_bf16_ss(tb + (i % c.ndist) * c.dstride, ad, bd, c.idesc, i >= c.ndist ? 1u : 0u);
      }
      const long long t1 = clock64();
      umma::commit(umma::smem_u32(&mbar));
      umma::mbar_wait(umma::smem_u32(&mbar), phase); phase ^= 1u;
      const long long t2 = clock64();
      if (threadIdx.x == 0) { out[2 * rep] = t1 - t0; out[2 * rep + 1] = t2 - t0; }
    }
  }
  umma::fence_before_sync(); __syncthreads();
  if (threadIdx.x < 32) umma::tmem_dealloc(tb, 512);
}

static void run(const char* name, Cfg c) {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(timing, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  timing<<<1, 128, 160 * 1024>>>(c, d);
  long long h[6]; cudaError_t e = cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("%-44s ERROR %s\n", name, cudaGetErrorString(e)); exit(1); }
  printf("%-44s L=%3d nd=%d  issue %6lld cyc (%5.1f/mma)  done %6lld cyc (%5.1f/mma)\n", name, c.L, c.ndist, h[4],
         (double)h[4] / c.L, h[5], (double)h[5] / c.L);
  cudaFree(d);
}

int main() {
  const int Ls[3] = {1, 12, 48};
  for (int N : {16, 64, 128, 192}) {
    for (int L : Ls) for (int nd : {1, 2, 4}) {
      if (nd * N > 512 || (L == 1 && nd > 1)) continue;
      Cfg c{}; c.idesc = umma::idesc_bf16(128, N, 0, 0);
      c.a_lbo = 128 * 16; c.a_sbo = 128; c.b_lbo = N * 16; c.b_sbo = 128; c.a_kstep = 128 * 32; c.b_kstep = N * 32;
      c.L = L; c.ndist = nd; c.dstride = N;
      char nm[64]; snprintf(nm, 64, "M128 N%d K-major/K-major", N); run(nm, c);
    }
  }
  for (int N : {16, 64, 128}) for (int L : Ls) for (int nd : {1, 2}) {   // dgrad: B read MN-major
    if (L == 1 && nd > 1) continue;
    Cfg c{}; c.idesc = umma::idesc_bf16(128, N, 0, 1);
    c.a_lbo = 128 * 16; c.a_sbo = 128; c.a_kstep = 128 * 32; c.b_lbo = 128; c.b_sbo = 64 * 16; c.b_kstep = 256;
    c.L = L; c.ndist = nd; c.dstride = N;
    char nm[64]; snprintf(nm, 64, "M128 N%d A K-major, B MN-major", N); run(nm, c);
  }
  for (int M : {64, 128}) for (int N : {16, 64, 128}) for (int L : Ls) for (int nd : {1, 2}) {   // wgrad: both MN-major
    if (L == 1 && nd > 1) continue;
    Cfg c{}; c.idesc = umma::idesc_bf16(M, N, 1, 1);
    c.a_lbo = 128; c.a_sbo = 128 * 16; c.a_kstep = 256; c.b_lbo = 128; c.b_sbo = 128 * 16; c.b_kstep = 256;
    c.L = L; c.ndist = nd; c.dstride = N;
    char nm[64]; snprintf(nm, 64, "M%d N%d MN-major/MN-major (wgrad)", M, N); run(nm, c);
  }
  return 0;
}
