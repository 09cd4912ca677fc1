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

// ------------------------------------------------------------------ P2G
__global__ void __launch_bounds__(256)
p2g_kernel(neuma_mpm_const c, int n, const float* __restrict__ x, const float* __restrict__ v,
           const float* __restrict__ C, const float* __restrict__ stress,
           const float* __restrict__ vol, const float* __restrict__ rho,
           const int* __restrict__ enabled, float4* __restrict__ grid) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  if (enabled[p] == 0) return;
  p2g_particle(c, p, x, v, C, stress, vol[p], rho[p], grid);
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
g2p_kernel(neuma_mpm_const c, int n, const float* __restrict__ x, const float* __restrict__ F,
           const float* __restrict__ clip, const int* __restrict__ enabled,
           const float4* __restrict__ grid, float* __restrict__ x2, float* __restrict__ v2,
           float* __restrict__ C2, float* __restrict__ F2) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  if (enabled[p] == 0) return;
  const float px[3] = {x[3 * p], x[3 * p + 1], x[3 * p + 2]};
  const Stencil s = make_stencil(px[0], px[1], px[2], c.inv_dx);
  float nv[3] = {0.f, 0.f, 0.f};
  float nC[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const float k4 = 4.0f * c.inv_dx * c.inv_dx;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int kk = 0; kk < 3; ++kk) {
        const int ix = s.b[0] + i, iy = s.b[1] + j, iz = s.b[2] + kk;
        if (!node_in_range(c, ix, iy, iz)) continue;
        const float4 q = grid[node_index(c, ix, iy, iz)];
        float gv[3];
        if (LAZY) {
          float pass[3], inv_m, vp[3];
          node_velocity(c, q, ix, iy, iz, gv, pass, inv_m, vp);
        } else {
          gv[0] = q.x; gv[1] = q.y; gv[2] = q.z;
        }
        const float d[3] = {((float)i - s.f[0]) * c.dx, ((float)j - s.f[1]) * c.dx,
                            ((float)kk - s.f[2]) * c.dx};
        const float w = s.w[0][i] * s.w[1][j] * s.w[2][kk];
        const float w4 = k4 * w;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          nv[a] += w * gv[a];
#pragma unroll
          for (int b = 0; b < 3; ++b) nC[3 * a + b] += w4 * (gv[a] * d[b]);
        }
      }
  // F' = (I + dt C') F   (mpm.py:487)
  float Fo[9], Fn[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) Fo[i] = F[9 * p + i];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k)
        acc += ((a == k ? 1.0f : 0.0f) + c.dt * nC[3 * a + k]) * Fo[3 * k + b];
      Fn[3 * a + b] = acc;
    }
  const float bnd = clip[p] * c.dx;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float nx = px[a] + c.dt * nv[a];
    nx = fminf(fmaxf(nx, 0.0f + bnd), 1.0f - bnd);
    x2[3 * p + a] = nx;
    v2[3 * p + a] = nv[a];
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    C2[9 * p + i] = nC[i];
    F2[9 * p + i] = Fn[i];
  }
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
  p2g_particle(c, p, x, v, C, stress, vol[p], rho[p], grid);
  const G2PAdj g = g2p_adjoint_inputs(c, p, x, F, clip, v2, gx2, gv2, gC2, gF2);
  const Stencil s = make_stencil(x[3 * p], x[3 * p + 1], x[3 * p + 2], c.inv_dx);
  const float k4 = 4.0f * c.inv_dx * c.inv_dx;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int kk = 0; kk < 3; ++kk) {
        const int ix = s.b[0] + i, iy = s.b[1] + j, iz = s.b[2] + kk;
        if (!node_in_range(c, ix, iy, iz)) continue;
        const float d[3] = {((float)i - s.f[0]) * c.dx, ((float)j - s.f[1]) * c.dx,
                            ((float)kk - s.f[2]) * c.dx};
        const float w = s.w[0][i] * s.w[1][j] * s.w[2][kk];
        float o[3];
#pragma unroll
        for (int a = 0; a < 3; ++a)
          o[a] = w * (g.gvt[a] + k4 * (g.gCt[3 * a] * d[0] + g.gCt[3 * a + 1] * d[1] +
                                       g.gCt[3 * a + 2] * d[2]));
        red_add_f4(gadj + node_index(c, ix, iy, iz), o[0], o[1], o[2], 0.0f);
      }
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
    // the reference's kernels return early: no adjoint contribution
    if (gx) for (int a = 0; a < 3; ++a) gx[3 * p + a] = 0.f;
    if (gv) for (int a = 0; a < 3; ++a) gv[3 * p + a] = 0.f;
    if (gC) for (int a = 0; a < 9; ++a) gC[9 * p + a] = 0.f;
    if (gF) for (int a = 0; a < 9; ++a) gF[9 * p + a] = 0.f;
    if (gstress) for (int a = 0; a < 9; ++a) gstress[9 * p + a] = 0.f;
    return;
  }
  const float p_vol = vol[p];
  const float p_mass = p_vol * rho[p];
  const Stencil s = make_stencil(x[3 * p], x[3 * p + 1], x[3 * p + 2], c.inv_dx);
  const float ks = -c.dt * p_vol * 4.0f * c.inv_dx * c.inv_dx;
  const float k4 = 4.0f * c.inv_dx * c.inv_dx;
  float A[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) A[i] = ks * stress[9 * p + i] + p_mass * C[9 * p + i];
  const float pv[3] = {p_mass * v[3 * p], p_mass * v[3 * p + 1], p_mass * v[3 * p + 2]};
  const G2PAdj g = g2p_adjoint_inputs(c, p, x, F, clip, v2, gx2, gv2, gC2, gF2);

  float acc_gmv[3] = {0.f, 0.f, 0.f};  // sum_n W g_mv
  float gA[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float gd[3] = {0.f, 0.f, 0.f};       // sum_n dL/d dpos_n
  float gw[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
  float nC[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // recomputed C'
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int kk = 0; kk < 3; ++kk) {
        const int ix = s.b[0] + i, iy = s.b[1] + j, iz = s.b[2] + kk;
        if (!node_in_range(c, ix, iy, iz)) continue;
        const size_t ni = node_index(c, ix, iy, iz);
        const float4 q = grid[ni];
        const float4 ga = gadj[ni];
        float vn[3], pass[3], inv_m, vp[3];
        node_velocity(c, q, ix, iy, iz, vn, pass, inv_m, vp);
        // adj(grid_op): v = mv/(m+eps) + g dt
        const float gvn[3] = {ga.x * pass[0], ga.y * pass[1], ga.z * pass[2]};
        const float gmv[3] = {gvn[0] * inv_m, gvn[1] * inv_m, gvn[2] * inv_m};
        const float gm = -(vp[0] * gvn[0] + vp[1] * gvn[1] + vp[2] * gvn[2]) * inv_m;
        const float d[3] = {((float)i - s.f[0]) * c.dx, ((float)j - s.f[1]) * c.dx,
                            ((float)kk - s.f[2]) * c.dx};
        const float w = s.w[0][i] * s.w[1][j] * s.w[2][kk];
        // --- adj(G2P) terms that need v_node
        float gCt_d[3], gCtT_v[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          gCt_d[a] = g.gCt[3 * a] * d[0] + g.gCt[3 * a + 1] * d[1] + g.gCt[3 * a + 2] * d[2];
          gCtT_v[a] = g.gCt[a] * vn[0] + g.gCt[3 + a] * vn[1] + g.gCt[6 + a] * vn[2];
        }
        float gW = vn[0] * g.gvt[0] + vn[1] * g.gvt[1] + vn[2] * g.gvt[2] +
                   k4 * (vn[0] * gCt_d[0] + vn[1] * gCt_d[1] + vn[2] * gCt_d[2]);
        // --- adj(P2G)
        float Ad[3], ATg[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          Ad[a] = A[3 * a] * d[0] + A[3 * a + 1] * d[1] + A[3 * a + 2] * d[2];
          ATg[a] = A[a] * gmv[0] + A[3 + a] * gmv[1] + A[6 + a] * gmv[2];
        }
        gW += gmv[0] * (pv[0] + Ad[0]) + gmv[1] * (pv[1] + Ad[1]) + gmv[2] * (pv[2] + Ad[2]) +
              gm * p_mass;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          acc_gmv[a] += w * gmv[a];
          gd[a] += w * (k4 * gCtT_v[a] + ATg[a]);
#pragma unroll
          for (int b = 0; b < 3; ++b) {
            gA[3 * a + b] += w * (gmv[a] * d[b]);
            nC[3 * a + b] += (k4 * w) * (vn[a] * d[b]);
          }
        }
        gw[0][i] += gW * (s.w[1][j] * s.w[2][kk]);
        gw[1][j] += gW * (s.w[0][i] * s.w[2][kk]);
        gw[2][kk] += gW * (s.w[0][i] * s.w[1][j]);
      }
  // weights -> fx: dw0/df = -(1.5 - f), dw1/df = -2 (f - 1), dw2/df = f - 0.5
  // dpos = (offset - fx) dx -> dL/dfx -= dx * gd ;  fx = x * inv_dx - base
  if (gx) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float f = s.f[a];
      float gf = gw[a][0] * (-(1.5f - f)) + gw[a][1] * (-2.0f * (f - 1.0f)) + gw[a][2] * (f - 0.5f);
      gf -= c.dx * gd[a];
      gx[3 * p + a] = finite_or_zero(g.gxn[a] + c.inv_dx * gf);
    }
  }
  if (gv) {
#pragma unroll
    for (int a = 0; a < 3; ++a) gv[3 * p + a] = finite_or_zero(p_mass * acc_gmv[a]);
  }
  if (gC) {
#pragma unroll
    for (int a = 0; a < 9; ++a) gC[9 * p + a] = finite_or_zero(p_mass * gA[a]);
  }
  if (gstress) {
#pragma unroll
    for (int a = 0; a < 9; ++a) gstress[9 * p + a] = finite_or_zero(ks * gA[a]);
  }
  if (gF) {
    // gF = (I + dt C')^T gF'
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k)
          acc += ((a == k ? 1.0f : 0.0f) + c.dt * nC[3 * k + a]) * gF2[9 * p + 3 * k + b];
        gF[9 * p + 3 * a + b] = finite_or_zero(acc);
      }
  }
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
