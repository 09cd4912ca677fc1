// MLS-MPM substep for sm_100a: P2G scatter, grid momentum solve, G2P gather, and
// the hand-derived adjoints.  Replaces the Warp kernels of
// modules/nclaw/sim/mpm.py:321-498 and the tape-generated adjoints driven at
// mpm.py:299-319.
//
// Data layout in HBM
//   particles : the reference's AoS fp32 arrays (x,v [Np,3]; C,F,stress [Np,3,3]).
//   grid      : ONE float4 per node, (mv.x, mv.y, mv.z, m), plane-major (z slowest).
//               A node visit in P2G is a single 16-byte red.global.add.v4.f32
//               instead of the reference's four scalar atomics, and a node visit in
//               G2P is one LDG.128.
//   grid_op   : evaluated lazily inside G2P per gathered node (same arithmetic as
//               mpm.py:402-429, done on the 27 nodes a particle reads) so the dense
//               G^3 sweep disappears from the fused substep; the dense kernel is kept
//               for the stand-alone neuma_mpm_grid_op entry point.
#include "mpm_common.cuh"

namespace {

// ------------------------------------------------------------------ clear
__global__ void grid_clear_kernel(float4* __restrict__ g, size_t n) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    g[i] = z;
}

__device__ __forceinline__ GridMap dense_map(const neuma_mpm_const& c) {
  return GridMap{0, 0, c.z_off, c.num_grids, c.num_grids, c.nz};
}

// ------------------------------------------------------------------ P2G
__global__ void __launch_bounds__(256)
p2g_kernel(neuma_mpm_const c, int n, const float* __restrict__ x, const float* __restrict__ v,
           const float* __restrict__ C, const float* __restrict__ stress,
           const float* __restrict__ vol, const float* __restrict__ rho,
           const int* __restrict__ enabled, float4* __restrict__ grid) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  if (enabled[p] == 0) return;
  p2g_particle(c, dense_map(c), p, x, v, C, stress + 9 * (size_t)p, vol[p], rho[p], grid);
}

// ------------------------------------------------------------------ grid_op (dense)
__global__ void __launch_bounds__(256)
grid_op_kernel(neuma_mpm_const c, const float4* __restrict__ grid, float4* __restrict__ gv) {
  const size_t total = (size_t)c.nz * c.num_grids * c.num_grids;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int ix = (int)(i % c.num_grids);
    const int iy = (int)((i / c.num_grids) % c.num_grids);
    const int iz = (int)(i / ((size_t)c.num_grids * c.num_grids)) + c.z_off;
    float v[3], pass[3], inv_m, vp[3];
    node_velocity(c, grid[i], ix, iy, iz, v, pass, inv_m, vp);
    gv[i] = make_float4(v[0], v[1], v[2], 0.f);
  }
}

// ------------------------------------------------------------------ G2P
template <bool LAZY>
__global__ void __launch_bounds__(256)
g2p_kernel(neuma_mpm_const c, int n, const float* x, const float* F,   // x2.. may alias x.. (in-place step)
           const float* __restrict__ clip, const int* __restrict__ enabled,
           const float4* __restrict__ grid, float* x2, float* v2, float* C2, float* F2) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  if (enabled[p] == 0) return;
  float nx3[3];
  g2p_particle<LAZY>(c, dense_map(c), p, x, F, clip, grid, x2, v2, C2, F2, nx3);
}

// ------------------------------------------------------------------ backward
// Pass 1 of the backward: recompute P2G into grid_mvm (mpm.py:312-314) and, in
// the same sweep over the particles, scatter adj(G2P)'s dL/dv_node into grid_adj.
__global__ void __launch_bounds__(256)
bwd_scatter_kernel(neuma_mpm_const c, int n, const float* __restrict__ x,
                   const float* __restrict__ v, const float* __restrict__ C,
                   const float* __restrict__ F, const float* __restrict__ stress,
                   const float* __restrict__ vol, const float* __restrict__ rho,
                   const float* __restrict__ clip, const int* __restrict__ enabled,
                   const float* __restrict__ v2, const float* __restrict__ gx2,
                   const float* __restrict__ gv2, const float* __restrict__ gC2,
                   const float* __restrict__ gF2, float4* __restrict__ grid,
                   float4* __restrict__ gadj) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  if (enabled[p] == 0) return;
  const GridMap gm = dense_map(c);
  p2g_particle(c, gm, p, x, v, C, stress + 9 * (size_t)p, vol[p], rho[p], grid);
  const G2PAdj g = g2p_adjoint_inputs(c, p, x, F, clip, v2, gx2, gv2, gC2, gF2 + 9 * (size_t)p);
  g2p_adjoint_scatter(c, gm, p, x, g, gadj);
}

// Pass 2: per particle, gather (mv, m) and dL/dv of its 27 nodes, push dL/dv
// through grid_op's adjoint on the fly, and accumulate adj(G2P) (the parts that
// need v_node) and adj(P2G) into the particle's input gradients.
__global__ void __launch_bounds__(128)
bwd_gather_kernel(neuma_mpm_const c, int n, const float* __restrict__ x,
                  const float* __restrict__ v, const float* __restrict__ C,
                  const float* __restrict__ F, const float* __restrict__ stress,
                  const float* __restrict__ vol, const float* __restrict__ rho,
                  const float* __restrict__ clip, const int* __restrict__ enabled,
                  const float* __restrict__ v2, const float* __restrict__ gx2,
                  const float* __restrict__ gv2, const float* __restrict__ gC2,
                  const float* __restrict__ gF2, const float4* __restrict__ grid,
                  const float4* __restrict__ gadj, float* __restrict__ gx,
                  float* __restrict__ gv, float* __restrict__ gC, float* __restrict__ gF,
                  float* __restrict__ gstress) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  if (enabled[p] == 0) {
    zero_particle_grads(p, gx, gv, gC, gF, gstress);
    return;
  }
  bwd_gather_particle<true, false>(c, dense_map(c), p, x, v, C, F, stress, vol, rho, clip, v2, nullptr, gx2,
                                   gv2, gC2, gF2, grid, gadj, gx, gv, gC, gF, gstress);
}

// ------------------------------------------------------------------ export / halo
__global__ void grid_export_kernel(neuma_mpm_const c, const float4* __restrict__ grid,
                                   const float4* __restrict__ gvel, float* __restrict__ mv,
                                   float* __restrict__ m, float* __restrict__ v) {
  const int G = c.num_grids;
  const size_t total = (size_t)c.nz * G * G;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int ix = (int)(i % G);
    const int iy = (int)((i / G) % G);
    const int lz = (int)(i / ((size_t)G * G));
    const size_t o = ((size_t)ix * G + iy) * c.nz + lz;  // reference order [x][y][z]
    const float4 q = grid[i];
    if (mv) { mv[3 * o] = q.x; mv[3 * o + 1] = q.y; mv[3 * o + 2] = q.z; }
    if (m) m[o] = q.w;
    if (v) {
      const float4 u = gvel[i];
      v[3 * o] = u.x; v[3 * o + 1] = u.y; v[3 * o + 2] = u.z;
    }
  }
}

__global__ void halo_add_kernel(float4* __restrict__ dst, const float4* __restrict__ src, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    float4 a = dst[i];
    const float4 b = src[i];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    dst[i] = a;
  }
}

inline int blocks_for(size_t n, int threads, int cap) {
  size_t b = (n + threads - 1) / threads;
  if (b > (size_t)cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

int check_const(neuma_ctx* ctx, const neuma_mpm_const* c) {
  if (!ctx || !c) return 2;
  if (c->num_grids <= 0 || c->nz <= 0 || c->z_off < 0 || c->z_off + c->nz > c->num_grids + 4)
    return neuma_fail_msg(ctx, "neuma_mpm: bad grid extents");
  if (c->bc != 0 && c->bc != 1) return neuma_fail_msg(ctx, "neuma_mpm: bc must be 0 or 1");
  if (c->compact || c->sel_mode)
    return neuma_fail_msg(ctx, "neuma_mpm: compact / sel_mode only apply to the rollout entry points");
  return neuma_ensure_grid(ctx, (size_t)c->nz * c->num_grids * c->num_grids);
}

}  // namespace

extern "C" {

int neuma_mpm_grid_clear(neuma_ctx* ctx, const neuma_mpm_const* c, void* stream) {
  if (int r = check_const(ctx, c)) return r;
  const size_t nodes = (size_t)c->nz * c->num_grids * c->num_grids;
  cudaStream_t st = (cudaStream_t)stream;
  grid_clear_kernel<<<blocks_for(nodes, 256, ctx->num_sms * 8), 256, 0, st>>>(ctx->grid_mvm, nodes);
  NEUMA_LAUNCH_CHECK(ctx, "grid_clear_kernel");
  return 0;
}

int neuma_mpm_p2g(neuma_ctx* ctx, const neuma_mpm_const* c, int n, const float* x, const float* v,
                  const float* C, const float* stress, const float* vol, const float* rho,
                  const int32_t* enabled, void* stream) {
  if (int r = check_const(ctx, c)) return r;
  if (n <= 0) return 0;
  p2g_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*c, n, x, v, C, stress, vol, rho,
                                                                 enabled, ctx->grid_mvm);
  NEUMA_LAUNCH_CHECK(ctx, "p2g_kernel");
  return 0;
}

int neuma_mpm_grid_op(neuma_ctx* ctx, const neuma_mpm_const* c, void* stream) {
  if (int r = check_const(ctx, c)) return r;
  const size_t nodes = (size_t)c->nz * c->num_grids * c->num_grids;
  grid_op_kernel<<<blocks_for(nodes, 256, ctx->num_sms * 8), 256, 0, (cudaStream_t)stream>>>(
      *c, ctx->grid_mvm, ctx->grid_v);
  NEUMA_LAUNCH_CHECK(ctx, "grid_op_kernel");
  return 0;
}

int neuma_mpm_g2p(neuma_ctx* ctx, const neuma_mpm_const* c, int n, const float* x, const float* F,
                  const float* clip_bound, const int32_t* enabled, float* x2, float* v2, float* C2,
                  float* F2, int lazy_grid_op, void* stream) {
  if (int r = check_const(ctx, c)) return r;
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (lazy_grid_op)
    g2p_kernel<true><<<(n + 255) / 256, 256, 0, st>>>(*c, n, x, F, clip_bound, enabled,
                                                      ctx->grid_mvm, x2, v2, C2, F2);
  else
    g2p_kernel<false><<<(n + 255) / 256, 256, 0, st>>>(*c, n, x, F, clip_bound, enabled,
                                                       ctx->grid_v, x2, v2, C2, F2);
  NEUMA_LAUNCH_CHECK(ctx, "g2p_kernel");
  return 0;
}

int neuma_mpm_grid_export(neuma_ctx* ctx, const neuma_mpm_const* c, float* mv, float* m, float* v,
                          void* stream) {
  if (int r = check_const(ctx, c)) return r;
  const size_t nodes = (size_t)c->nz * c->num_grids * c->num_grids;
  grid_export_kernel<<<blocks_for(nodes, 256, ctx->num_sms * 8), 256, 0, (cudaStream_t)stream>>>(
      *c, ctx->grid_mvm, ctx->grid_v, mv, m, v);
  NEUMA_LAUNCH_CHECK(ctx, "grid_export_kernel");
  return 0;
}

int neuma_mpm_halo_pack(neuma_ctx* ctx, const neuma_mpm_const* c, int z, int nplanes, float* buf,
                        void* stream) {
  if (int r = check_const(ctx, c)) return r;
  if (z < c->z_off || z + nplanes > c->z_off + c->nz)
    return neuma_fail_msg(ctx, "neuma_mpm_halo_pack: planes outside the slab");
  const size_t plane = (size_t)c->num_grids * c->num_grids;
  NEUMA_CUDA(ctx, cudaMemcpyAsync(buf, ctx->grid_mvm + (size_t)(z - c->z_off) * plane,
                                  plane * nplanes * sizeof(float4), cudaMemcpyDeviceToDevice,
                                  (cudaStream_t)stream));
  return 0;
}

int neuma_mpm_halo_unpack_add(neuma_ctx* ctx, const neuma_mpm_const* c, int z, int nplanes,
                              const float* buf, void* stream) {
  if (int r = check_const(ctx, c)) return r;
  if (z < c->z_off || z + nplanes > c->z_off + c->nz)
    return neuma_fail_msg(ctx, "neuma_mpm_halo_unpack_add: planes outside the slab");
  const size_t plane = (size_t)c->num_grids * c->num_grids;
  const size_t n = plane * nplanes;
  halo_add_kernel<<<blocks_for(n, 256, ctx->num_sms * 8), 256, 0, (cudaStream_t)stream>>>(
      ctx->grid_mvm + (size_t)(z - c->z_off) * plane, (const float4*)buf, n);
  NEUMA_LAUNCH_CHECK(ctx, "halo_add_kernel");
  return 0;
}

int neuma_mpm_substep_fwd(neuma_ctx* ctx, const neuma_mpm_const* c, int n, const float* x,
                          const float* v, const float* C, const float* F, const float* stress,
                          const float* vol, const float* rho, const float* clip_bound,
                          const int32_t* enabled, float* x2, float* v2, float* C2, float* F2,
                          void* stream) {
  if (int r = neuma_mpm_grid_clear(ctx, c, stream)) return r;
  if (int r = neuma_mpm_p2g(ctx, c, n, x, v, C, stress, vol, rho, enabled, stream)) return r;
  return neuma_mpm_g2p(ctx, c, n, x, F, clip_bound, enabled, x2, v2, C2, F2, 1, stream);
}

int neuma_mpm_substep_bwd(neuma_ctx* ctx, const neuma_mpm_const* c, int n, const float* x,
                          const float* v, const float* C, const float* F, const float* stress,
                          const float* vol, const float* rho, const float* clip_bound,
                          const int32_t* enabled, const float* v2, const float* gx2,
                          const float* gv2, const float* gC2, const float* gF2, float* gx,
                          float* gv, float* gC, float* gF, float* gstress, void* stream) {
  if (int r = check_const(ctx, c)) return r;
  if (n <= 0) return 0;
  if (!v2 || !gx2 || !gv2 || !gC2 || !gF2)
    return neuma_fail_msg(ctx, "neuma_mpm_substep_bwd: upstream gradients must be non-NULL");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t nodes = (size_t)c->nz * c->num_grids * c->num_grids;
  const int cb = blocks_for(nodes, 256, ctx->num_sms * 8);
  grid_clear_kernel<<<cb, 256, 0, st>>>(ctx->grid_mvm, nodes);
  NEUMA_LAUNCH_CHECK(ctx, "grid_clear_kernel");
  grid_clear_kernel<<<cb, 256, 0, st>>>(ctx->grid_adj, nodes);
  NEUMA_LAUNCH_CHECK(ctx, "grid_clear_kernel");
  bwd_scatter_kernel<<<(n + 255) / 256, 256, 0, st>>>(*c, n, x, v, C, F, stress, vol, rho,
                                                      clip_bound, enabled, v2, gx2, gv2, gC2, gF2,
                                                      ctx->grid_mvm, ctx->grid_adj);
  NEUMA_LAUNCH_CHECK(ctx, "bwd_scatter_kernel");
  bwd_gather_kernel<<<(n + 127) / 128, 128, 0, st>>>(
      *c, n, x, v, C, F, stress, vol, rho, clip_bound, enabled, v2, gx2, gv2, gC2, gF2,
      ctx->grid_mvm, ctx->grid_adj, gx, gv, gC, gF, gstress);
  NEUMA_LAUNCH_CHECK(ctx, "bwd_gather_kernel");
  return 0;
}

}  // extern "C"
