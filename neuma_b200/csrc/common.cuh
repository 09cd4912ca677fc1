// Shared device/host helpers for libneuma_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/neuma_b200.h"

#define NEUMA_NUM_SMS_B200 148

struct RasterState;  // raster.cu

struct neuma_ctx {
  int device;
  int num_sms;
  char err[512];
  long long launches;
  // MPM grid workspace: one float4 per node (mv.x, mv.y, mv.z, m), plane-major
  // index (iz - z_off) * G * G + iy * G + ix  (z slowest so that a z-plane -- the
  // slab halo unit -- is contiguous; the reference's [x][y][z] order is only
  // produced by neuma_mpm_grid_export).
  float4* grid_mvm;   // forward grid
  float4* grid_v;     // dense grid_op output (v.x, v.y, v.z, 0)
  float4* grid_adj;   // backward: (gv.x, gv.y, gv.z, 0)
  size_t grid_nodes;  // capacity, in nodes, of each of the above
  // scratch for partial reductions (constitutive weight grads)
  float* scratch;
  size_t scratch_bytes;
  // split/laid-out constitutive weights for the un-staged entry points (constitutive.cu)
  void* wstage;
  size_t wstage_bytes;
  RasterState* raster;
  int attrs_set;  // cudaFuncSetAttribute (opt-in shared memory) is per DEVICE: tracked per ctx
  int rollout_attrs_set;
  // neuma_constitutive_svd_cache: consumed by the NEXT constitutive launch on this ctx
  float* svd_v_out;        // forward kernels: the right singular vectors of every input row go here
  const float* svd_v_in;   // backward kernels: read them instead of running the Jacobi iteration
};

static inline int neuma_fail(neuma_ctx* ctx, const char* what, cudaError_t e) {
  if (ctx) snprintf(ctx->err, sizeof(ctx->err), "%s: %s", what, cudaGetErrorString(e));
  return 1;
}
static inline int neuma_fail_msg(neuma_ctx* ctx, const char* what) {
  if (ctx) snprintf(ctx->err, sizeof(ctx->err), "%s", what);
  return 2;
}

#define NEUMA_CUDA(ctx, call)                                   \
  do {                                                          \
    cudaError_t e__ = (call);                                   \
    if (e__ != cudaSuccess) return neuma_fail(ctx, #call, e__); \
  } while (0)

#define NEUMA_LAUNCH_CHECK(ctx, name)                            \
  do {                                                           \
    cudaError_t e__ = cudaGetLastError();                        \
    if (e__ != cudaSuccess) return neuma_fail(ctx, name, e__);   \
    (ctx)->launches++;                                           \
  } while (0)

// ---- programmatic dependent launch (griddepcontrol, sm_90+) -----------------------------
// The rollout is a chain of ~260 short dependent kernels per frame.  Kernels launched with
// neuma_launch() carry the programmatic-stream-serialization attribute: the next kernel of
// the stream is scheduled as soon as every CTA of this one has started (launch_dependents at
// the top), runs its prologue (barrier init, TMEM allocation, index math) while this one
// drains, and blocks in pdl_wait() until this grid has completed and its writes are visible.
// Semantics are those of ordinary stream order.  EVERY kernel launched through neuma_launch() must call
// pdl_wait() (pdl_prologue()) before it touches anything a predecessor wrote.
// Round 1 had every kernel trigger at its top: +1.7 % on the forward graph, -3 % on the backward one (the
// kernel after a persistent constitutive kernel then sits resident beside it for its whole duration) -> off.
// Now the SMALL kernels trigger early and the constitutive kernels do not: a constitutive kernel that
// follows a small one (plasticity after G2P, elasticity backward after the node pass) runs its prologue
// (barriers, TMEM allocation) on the SMs the small kernel has already left.  Forward graph -5.7 %,
// backward -0.4 %, frame +2.1 % (profiles/r03y_pdl_ab.txt): ON by default, NEUMA_PDL=0 disables.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
  pdl_launch_dependents();
  pdl_wait();
}
#endif
bool neuma_pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t neuma_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = neuma_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

int neuma_ensure_grid(neuma_ctx* ctx, size_t nodes);
int neuma_ensure_scratch(neuma_ctx* ctx, size_t bytes);
int neuma_ensure_wstage(neuma_ctx* ctx, size_t bytes);

// ---- small 3x3 helpers (row-major float[9]) ------------------------------
struct Mat3 {
  float m[9];
  __host__ __device__ float& operator()(int r, int c) { return m[r * 3 + c]; }
  __host__ __device__ const float& operator()(int r, int c) const { return m[r * 3 + c]; }
};

__device__ __forceinline__ Mat3 mat3_load(const float* p) {
  Mat3 a;
#pragma unroll
  for (int i = 0; i < 9; ++i) a.m[i] = p[i];
  return a;
}
__device__ __forceinline__ void mat3_store(float* p, const Mat3& a) {
#pragma unroll
  for (int i = 0; i < 9; ++i) p[i] = a.m[i];
}
__device__ __forceinline__ Mat3 mat3_mul(const Mat3& a, const Mat3& b) {
  Mat3 c;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      c(i, j) = a(i, 0) * b(0, j) + a(i, 1) * b(1, j) + a(i, 2) * b(2, j);
  return c;
}
// a * b^T
__device__ __forceinline__ Mat3 mat3_mul_nt(const Mat3& a, const Mat3& b) {
  Mat3 c;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      c(i, j) = a(i, 0) * b(j, 0) + a(i, 1) * b(j, 1) + a(i, 2) * b(j, 2);
  return c;
}
// a^T * b
__device__ __forceinline__ Mat3 mat3_mul_tn(const Mat3& a, const Mat3& b) {
  Mat3 c;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      c(i, j) = a(0, i) * b(0, j) + a(1, i) * b(1, j) + a(2, i) * b(2, j);
  return c;
}
__device__ __forceinline__ float mat3_det(const Mat3& a) {
  return a(0, 0) * (a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1)) -
         a(0, 1) * (a(1, 0) * a(2, 2) - a(1, 2) * a(2, 0)) +
         a(0, 2) * (a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0));
}

// red.global.add.v4.f32: one 16-byte vector reduction per grid node visit
// (sm_90+), instead of the reference's 4 scalar atomics (mpm.py:370-371).
__device__ __forceinline__ void red_add_f4(float4* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

__device__ __forceinline__ float finite_or_zero(float a) {
  // torch.nan_to_num_(g, 0, 0, 0), modules/nclaw/sim/interface.py:65-74
  return (isnan(a) || isinf(a)) ? 0.f : a;
}
