// Stand-alone probe of the tcgen05 conventions the constitutive kernels rely on:
// shared-memory descriptor fields (no swizzle, K-major and MN-major), the instruction
// descriptor, 3xTF32 accuracy, the M=128 and M=64 accumulator layouts in TMEM, and
// A-from-TMEM.  Host builds raw smem images; the kernel copies them in, issues the
// MMAs it is told to, and dumps TMEM [128 lanes x ncols].
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe tools/umma_probe.cu && ./umma_probe
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../neuma_b200/csrc/umma.cuh"

struct Job {
  uint32_t idesc;
  uint32_t a_off[3], b_off[3];  // byte offsets of the split images in smem
  int bf16;                     // 0: 3xTF32 (2 splits, 3 terms), 1: bf16x3 (3 splits, 6 terms)
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
  uint32_t a_kstep, b_kstep;  // bytes to advance the start address per K=8 step
  int nk;
  int ncols;       // accumulator columns to dump
  int a_in_tmem;   // 1: A is first stored to TMEM columns [64, 64+K) hi and [128,128+K) lo
  int K;           // (a_in_tmem) columns of A
  int smem_bytes;
  int per_term;     // 1: term t accumulates into columns [t*ncols, (t+1)*ncols)
};

__global__ void __launch_bounds__(128) probe(Job j, const uint8_t* __restrict__ img,
                                             const float* __restrict__ a_rows_hi,
                                             const float* __restrict__ a_rows_lo,
                                             float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tbase_s;
  for (int i = threadIdx.x; i < j.smem_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(sm)[i] = reinterpret_cast<const uint4*>(img)[i];
  if (threadIdx.x < 32) umma::tmem_alloc(umma::smem_u32(&tbase_s), 512);
  if (threadIdx.x == 0) {
    umma::mbar_init(umma::smem_u32(&mbar), 1);
    umma::fence_mbar_init();
  }
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tb = tbase_s;
  const int warp = threadIdx.x >> 5;
  const uint32_t lane_base = tb + ((uint32_t)(warp * 32) << 16);
  if (j.a_in_tmem) {
    for (int c = 0; c < j.K; c += 16) {
      float h[16], l[16];
      for (int i = 0; i < 16; ++i) {
        h[i] = a_rows_hi[threadIdx.x * j.K + c + i];
        l[i] = a_rows_lo[threadIdx.x * j.K + c + i];
      }
      umma::tmem_st16(lane_base + 64 + c, h);
      umma::tmem_st16(lane_base + 128 + c, l);
    }
    umma::tmem_wait_st();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
  }
  if (threadIdx.x < 32) {
    const uint32_t s0 = umma::smem_u32(sm);
    uint32_t acc = 0;
    // tf32: hi*hi, lo*hi, hi*lo ; bf16: all split pairs with i + j <= 2
    const int ai[6] = {0, 1, 0, 1, 2, 0}, bi[6] = {0, 0, 1, 1, 0, 2};
    const int nterm = j.bf16 ? 6 : 3;
    for (int t = 0; t < nterm; ++t) {
      if (j.per_term) acc = 0;
      const uint32_t td = tb + (j.per_term ? t * j.ncols : 0);
      for (int k = 0; k < j.nk; ++k) {
        const uint64_t bd = umma::smem_desc(s0 + j.b_off[bi[t]] + k * j.b_kstep, j.b_lbo, j.b_sbo);
        if (j.a_in_tmem) {
          umma::mma_tf32_ts(td, tb + (ai[t] ? 128 : 64) + k * 8, bd, j.idesc, acc);
        } else {
          const uint64_t ad = umma::smem_desc(s0 + j.a_off[ai[t]] + k * j.a_kstep, j.a_lbo, j.a_sbo);
          if (j.bf16) umma::mma_bf16_ss(td, ad, bd, j.idesc, acc);
          else umma::mma_tf32_ss(td, ad, bd, j.idesc, acc);
        }
        acc = 1;
      }
    }
    umma::commit(umma::smem_u32(&mbar));
  }
  umma::mbar_wait(umma::smem_u32(&mbar), 0);
  umma::fence_after_sync();
  const int dumpc = j.per_term ? j.ncols * (j.bf16 ? 6 : 3) : j.ncols;
  for (int c = 0; c < dumpc; c += 16) {
    float v[16];
    umma::tmem_ld16(lane_base + c, v);
    umma::tmem_wait_ld();
    for (int i = 0; i < 16; ++i) out[threadIdx.x * dumpc + c + i] = v[i];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) umma::tmem_dealloc(tb, 512);
}

static float tf32_hi_host(float a) {
  uint32_t u;
  memcpy(&u, &a, 4);
  u += 0x1000u;  // round to nearest (ties away), like cvt.rna
  u &= 0xFFFFE000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}

// image of X[R][K] (row-major source) as [k/4][r][4] floats; split into hi/lo
static void image(const std::vector<float>& X, int R, int K, std::vector<float>& hi, std::vector<float>& lo) {
  hi.assign((size_t)R * K, 0.f);
  lo.assign((size_t)R * K, 0.f);
  for (int r = 0; r < R; ++r)
    for (int k = 0; k < K; ++k) {
      const size_t o = ((size_t)(k / 4) * R + r) * 4 + (k % 4);
      const float h = tf32_hi_host(X[(size_t)r * K + k]);
      hi[o] = h;
      lo[o] = X[(size_t)r * K + k] - h;
    }
}

static uint16_t bf16_rn(float a) {
  uint32_t u; memcpy(&u, &a, 4);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static float bf16_f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float r; memcpy(&r, &u, 4); return r; }
// bf16 image of X[R][K] as [k/8][r][8] bf16, three splits; returned as float vectors holding raw bytes
static void image16(const std::vector<float>& X, int R, int K, std::vector<float> (&sp)[3]) {
  std::vector<uint16_t> s[3];
  for (auto& v : s) v.assign((size_t)R * K, 0);
  for (int r = 0; r < R; ++r)
    for (int k = 0; k < K; ++k) {
      const size_t o = ((size_t)(k / 8) * R + r) * 8 + (k % 8);
      float x = X[(size_t)r * K + k];
      for (int t = 0; t < 3; ++t) { s[t][o] = bf16_rn(x); x -= bf16_f(s[t][o]); }
    }
  for (int t = 0; t < 3; ++t) { sp[t].assign((size_t)R * K / 2, 0.f); memcpy(sp[t].data(), s[t].data(), (size_t)R * K * 2); }
}

static std::vector<float> rnd(size_t n, unsigned seed) {
  std::vector<float> v(n);
  srand(seed);
  for (auto& x : v) x = (float)rand() / RAND_MAX * 2.f - 1.f;
  return v;
}

static int run(const char* name, Job j, const std::vector<float>& Ahi, const std::vector<float>& Alo,
               const std::vector<float>& Bhi, const std::vector<float>& Blo,
               const std::vector<float>& A3, const std::vector<float>& B3,
               const std::vector<float>& arows_hi, const std::vector<float>& arows_lo,
               const std::vector<double>& ref /* [rows][ncols] */, int rows, bool m64) {
  std::vector<uint8_t> img;
  auto put = [&](const std::vector<float>& v) {
    uint32_t off = (uint32_t)img.size();
    img.insert(img.end(), (const uint8_t*)v.data(), (const uint8_t*)(v.data() + v.size()));
    while (img.size() % 1024) img.push_back(0);
    return off;
  };
  j.a_off[0] = put(Ahi); j.a_off[1] = put(Alo); j.a_off[2] = put(A3);
  j.b_off[0] = put(Bhi); j.b_off[1] = put(Blo); j.b_off[2] = put(B3);
  j.smem_bytes = (int)img.size();
  uint8_t* dimg; float *dout, *dah, *dal;
  cudaMalloc(&dimg, img.size());
  cudaMemcpy(dimg, img.data(), img.size(), cudaMemcpyHostToDevice);
  const int dumpc = j.per_term ? j.ncols * (j.bf16 ? 6 : 3) : j.ncols;
  cudaMalloc(&dout, 128 * dumpc * 4);
  cudaMemset(dout, 0, 128 * dumpc * 4);
  cudaMalloc(&dah, arows_hi.size() * 4 + 4);
  cudaMalloc(&dal, arows_lo.size() * 4 + 4);
  cudaMemcpy(dah, arows_hi.data(), arows_hi.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dal, arows_lo.data(), arows_lo.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe<<<1, 128, j.smem_bytes, 0>>>(j, dimg, dah, dal, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-28s CUDA ERROR %s\n", name, cudaGetErrorString(e));
    return 1;
  }
  std::vector<float> raw(128 * dumpc);
  cudaMemcpy(raw.data(), dout, raw.size() * 4, cudaMemcpyDeviceToHost);
  std::vector<float> out(128 * j.ncols, 0.f);
  for (int l = 0; l < 128; ++l)
    for (int c = 0; c < dumpc; ++c) out[l * j.ncols + c % j.ncols] += raw[l * dumpc + c];
  if (getenv("PROBE_DUMP")) {
    char fn[256]; static int idx = 0;
    snprintf(fn, 256, "%s/probe_%02d.bin", getenv("PROBE_DUMP"), idx++);
    FILE* f = fopen(fn, "wb");
    if (f) { int hdr[4] = {j.ncols, rows, (int)img.size(), j.bf16}; fwrite(hdr, 4, 4, f); fwrite(out.data(), 4, out.size(), f); { int d = dumpc; fwrite(&d, 4, 1, f); fwrite(raw.data(), 4, raw.size(), f); }
             fwrite(ref.data(), 8, ref.size(), f); fwrite(img.data(), 1, img.size(), f); fclose(f); printf("  [dumped %s: %s]\n", fn, name); }
  }
  double maxerr = 0, maxref = 0;
  for (int r = 0; r < rows; ++r) {
    const int lane = m64 ? (r % 16) + 32 * (r / 16) : r;
    for (int c = 0; c < j.ncols; ++c) {
      maxerr = fmax(maxerr, fabs(out[lane * j.ncols + c] - ref[(size_t)r * j.ncols + c]));
      maxref = fmax(maxref, fabs(ref[(size_t)r * j.ncols + c]));
    }
  }
  printf("%-28s max|err| = %.3e  (max|ref| %.3f)  rel %.3e  %s\n", name, maxerr, maxref, maxerr / maxref,
         maxerr / maxref < 1e-5 ? "OK" : "MISMATCH");
  if (maxerr / maxref >= 1e-5) {
    int nz = 0; for (float v : out) nz += (v != 0.f);
    printf("   nonzero outputs: %d of %zu\n", nz, out.size());
    printf("   out[lane0][0..3] = %g %g %g %g   ref = %g %g %g %g\n", out[0], out[1], out[2], out[3], ref[0],
           ref[1], ref[2], ref[3]);
  }
  cudaFree(dimg); cudaFree(dout); cudaFree(dah); cudaFree(dal);
  return maxerr / maxref < 1e-5 ? 0 : 1;
}

int main() {
  int bad = 0;
  const int PT = getenv("PROBE_PER_TERM") ? 1 : 0;
  std::vector<float> Ahi, Alo, Bhi, Blo, none;
  {  // 0: D[128][64] = A[128][64] B[64][64]^T, both K-major
    const int M = 128, N = 64, K = 64;
    auto A = rnd(M * K, 1), B = rnd(N * K, 2);
    image(A, M, K, Ahi, Alo); image(B, N, K, Bhi, Blo);
    std::vector<double> ref(M * N);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k]; ref[m * N + n] = s; }
    Job j{}; j.idesc = umma::idesc_tf32(M, N, 0, 0);
    j.a_lbo = M * 16; j.a_sbo = 128; j.b_lbo = N * 16; j.b_sbo = 128;
    j.a_kstep = 2 * M * 16; j.b_kstep = 2 * N * 16; j.nk = K / 8; j.ncols = N;
    bad += run("M128 N64 K64 K-major/K-major", j, Ahi, Alo, Bhi, Blo, none, none, none, none, ref, M, false);
    // 5: same product, A from TMEM
    std::vector<float> ah(M * K), al(M * K);
    for (int i = 0; i < M * K; ++i) { ah[i] = tf32_hi_host(A[i]); al[i] = A[i] - ah[i]; }
    j.a_in_tmem = 1; j.K = K;
    bad += run("M128 N64 K64 A in TMEM", j, Ahi, Alo, Bhi, Blo, none, none, ah, al, ref, M, false);
  }
  {  // 1: N = 16
    const int M = 128, N = 16, K = 64;
    auto A = rnd(M * K, 3), B = rnd(N * K, 4);
    image(A, M, K, Ahi, Alo); image(B, N, K, Bhi, Blo);
    std::vector<double> ref(M * N);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k]; ref[m * N + n] = s; }
    Job j{}; j.idesc = umma::idesc_tf32(M, N, 0, 0);
    j.a_lbo = M * 16; j.a_sbo = 128; j.b_lbo = N * 16; j.b_sbo = 128;
    j.a_kstep = 2 * M * 16; j.b_kstep = 2 * N * 16; j.nk = K / 8; j.ncols = N;
    bad += run("M128 N16 K64 K-major/K-major", j, Ahi, Alo, Bhi, Blo, none, none, none, none, ref, M, false);
  }
  {  // 2: dgrad: D[128][64] = A[128][16] W[16][64]; W stored as the forward's K-major operand
     //    W[j][hid] (rows j, K = hid) and read here MN-major (N = hid, K = j)
    const int M = 128, N = 64, K = 16;
    auto A = rnd(M * K, 5), W = rnd(K * N, 6);
    image(A, M, K, Ahi, Alo); image(W, K, N, Bhi, Blo);  // image rows = j (16), cols = hid (64)
    std::vector<double> ref(M * N);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * W[k * N + n]; ref[m * N + n] = s; }
    Job j{}; j.idesc = umma::idesc_tf32(M, N, 0, 1);
    j.a_lbo = M * 16; j.a_sbo = 128; j.a_kstep = 2 * M * 16;
    // image [hid/4][j][4]: MN(hid) chunk stride = 16 rows * 16 B; K(j) 8-row group stride = 128
    j.b_sbo = K * 16; j.b_lbo = 128; j.b_kstep = 128;
    j.nk = K / 8; j.ncols = N;
    bad += run("M128 N64 K16 B MN-major", j, Ahi, Alo, Bhi, Blo, none, none, none, none, ref, M, false);
  }
  {  // 3: wgrad: D[64][64] = X^T Y, X,Y [128 p][64 f], both MN-major, M = 64
    const int P = 128, F1 = 64, F2 = 64;
    auto X = rnd(P * F1, 7), Y = rnd(P * F2, 8);
    image(X, P, F1, Ahi, Alo); image(Y, P, F2, Bhi, Blo);
    std::vector<double> ref(F1 * F2);
    for (int a = 0; a < F1; ++a) for (int b = 0; b < F2; ++b) { double s = 0; for (int p = 0; p < P; ++p) s += (double)X[p * F1 + a] * Y[p * F2 + b]; ref[a * F2 + b] = s; }
    Job j{}; j.idesc = umma::idesc_tf32(64, F2, 1, 1);
    j.a_sbo = P * 16; j.a_lbo = 128; j.a_kstep = 128;
    j.b_sbo = P * 16; j.b_lbo = 128; j.b_kstep = 128;
    j.nk = P / 8; j.ncols = F2;
    bad += run("M64 N64 K128 MN/MN (wgrad)", j, Ahi, Alo, Bhi, Blo, none, none, none, none, ref, F1, true);
  }
  {  // 4: wgrad N = 16
    const int P = 128, F1 = 64, F2 = 16;
    auto X = rnd(P * F1, 9), Y = rnd(P * F2, 10);
    image(X, P, F1, Ahi, Alo); image(Y, P, F2, Bhi, Blo);
    std::vector<double> ref(F1 * F2);
    for (int a = 0; a < F1; ++a) for (int b = 0; b < F2; ++b) { double s = 0; for (int p = 0; p < P; ++p) s += (double)X[p * F1 + a] * Y[p * F2 + b]; ref[a * F2 + b] = s; }
    Job j{}; j.idesc = umma::idesc_tf32(64, F2, 1, 1);
    j.a_sbo = P * 16; j.a_lbo = 128; j.a_kstep = 128;
    j.b_sbo = P * 16; j.b_lbo = 128; j.b_kstep = 128;
    j.nk = P / 8; j.ncols = F2;
    bad += run("M64 N16 K128 MN/MN (wgrad)", j, Ahi, Alo, Bhi, Blo, none, none, none, none, ref, F1, true);
  }
  {  // 6: dgrad with N = 16: D[128][16] = A[128][64] W0[64][16]; W0 image rows = o(64), cols = j(16)
    const int M = 128, N = 16, K = 64;
    auto A = rnd(M * K, 11), W = rnd(K * N, 12);
    image(A, M, K, Ahi, Alo); image(W, K, N, Bhi, Blo);
    std::vector<double> ref(M * N);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * W[k * N + n]; ref[m * N + n] = s; }
    Job j{}; j.idesc = umma::idesc_tf32(M, N, 0, 1);
    j.a_lbo = M * 16; j.a_sbo = 128; j.a_kstep = 2 * M * 16;
    j.b_sbo = K * 16; j.b_lbo = 128; j.b_kstep = 128;
    j.nk = K / 8; j.ncols = N;
    bad += run("M128 N16 K64 B MN-major", j, Ahi, Alo, Bhi, Blo, none, none, none, none, ref, M, false);
  }


  std::vector<float> As[3], Bs[3];
  {  // bf16x3, K-major/K-major, M128 N64 K64
    const int M = 128, N = 64, K = 64;
    auto A = rnd(M * K, 21), B = rnd(N * K, 22);
    image16(A, M, K, As); image16(B, N, K, Bs);
    std::vector<double> ref(M * N);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k]; ref[m * N + n] = s; }
    Job j{}; j.per_term = PT; j.bf16 = 1; j.idesc = umma::idesc_bf16(M, N, 0, 0);
    j.a_lbo = M * 16; j.a_sbo = 128; j.b_lbo = N * 16; j.b_sbo = 128;
    j.a_kstep = 2 * M * 16; j.b_kstep = 2 * N * 16; j.nk = K / 16; j.ncols = N;
    bad += run("bf16x3 M128 N64 K64 K/K", j, As[0], As[1], Bs[0], Bs[1], As[2], Bs[2], none, none, ref, M, false);
  }
  {  // bf16x3 dgrad: D[128][64] = A[128][16] W[16][64], W image rows = j (16), cols = hid (64), read MN-major
    const int M = 128, N = 64, K = 16;
    auto A = rnd(M * K, 23), W = rnd(K * N, 24);
    image16(A, M, K, As); image16(W, K, N, Bs);
    std::vector<double> ref(M * N);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * W[k * N + n]; ref[m * N + n] = s; }
    Job j{}; j.per_term = PT; j.bf16 = 1; j.idesc = umma::idesc_bf16(M, N, 0, 1);
    j.a_lbo = M * 16; j.a_sbo = 128; j.a_kstep = 2 * M * 16;
    j.b_sbo = K * 16; j.b_lbo = 128; j.b_kstep = 256;   // image [hid/8][j][8]: MN chunk stride 16 rows*16B; 8-row K group 128 B
    j.nk = K / 16; j.ncols = N;
    bad += run("bf16x3 M128 N64 K16 B MN", j, As[0], As[1], Bs[0], Bs[1], As[2], Bs[2], none, none, ref, M, false);
  }
  {  // bf16x3 dgrad N=16: D[128][16] = A[128][64] W0[64][16]
    const int M = 128, N = 16, K = 64;
    auto A = rnd(M * K, 25), W = rnd(K * N, 26);
    image16(A, M, K, As); image16(W, K, N, Bs);
    std::vector<double> ref(M * N);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * W[k * N + n]; ref[m * N + n] = s; }
    Job j{}; j.per_term = PT; j.bf16 = 1; j.idesc = umma::idesc_bf16(M, N, 0, 1);
    j.a_lbo = M * 16; j.a_sbo = 128; j.a_kstep = 2 * M * 16;
    j.b_sbo = K * 16; j.b_lbo = 128; j.b_kstep = 256;
    j.nk = K / 16; j.ncols = N;
    bad += run("bf16x3 M128 N16 K64 B MN", j, As[0], As[1], Bs[0], Bs[1], As[2], Bs[2], none, none, ref, M, false);
  }
  for (int F2 : {64, 16}) {  // bf16x3 wgrad: D[64][F2] = X^T Y, X [128 p][64 f], Y [128 p][F2]; both MN-major, M = 64
    const int P = 128, F1 = 64;
    auto X = rnd(P * F1, 27), Y = rnd(P * F2, 28);
    image16(X, P, F1, As); image16(Y, P, F2, Bs);
    std::vector<double> ref(F1 * F2);
    for (int a = 0; a < F1; ++a) for (int b = 0; b < F2; ++b) { double s = 0; for (int p = 0; p < P; ++p) s += (double)X[p * F1 + a] * Y[p * F2 + b]; ref[a * F2 + b] = s; }
    Job j{}; j.per_term = PT; j.bf16 = 1; j.idesc = umma::idesc_bf16(64, F2, 1, 1);
    j.a_sbo = P * 16; j.a_lbo = 128; j.a_kstep = 256;
    j.b_sbo = P * 16; j.b_lbo = 128; j.b_kstep = 256;
    j.nk = P / 16; j.ncols = F2;
    bad += run(F2 == 64 ? "bf16x3 M64 N64 K128 MN/MN" : "bf16x3 M64 N16 K128 MN/MN", j, As[0], As[1], Bs[0], Bs[1], As[2], Bs[2], none, none, ref, F1, true);
  }
  printf(bad ? "PROBE FAILED (%d)\n" : "PROBE OK\n", bad);
  return bad;
}
