// Device helpers shared by the MPM kernels (mpm.cu: stand-alone / per-substep entry
// points; mpm_rollout.cu: the rollout path with caller-owned per-substep grids).
// Arithmetic follows modules/nclaw/sim/mpm.py:321-498.
#pragma once
#include "common.cuh"

namespace {

struct Stencil {
  int b[3];       // base node (trunc toward zero, mpm.py:337-339)
  float f[3];     // fx = x*inv_dx - base
  float w[3][3];  // w[axis][i], quadratic B-spline (mpm.py:345-354)
};

__device__ __forceinline__ Stencil make_stencil(float x0, float x1, float x2, float inv_dx) {
  Stencil s;
  const float p[3] = {x0 * inv_dx, x1 * inv_dx, x2 * inv_dx};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    s.b[a] = (int)(p[a] - 0.5f);  // int() truncates, it is not floor
    s.f[a] = p[a] - (float)s.b[a];
    const float wa = 1.5f - s.f[a];
    const float wb = s.f[a] - 1.0f;
    const float wc = s.f[a] - 0.5f;
    s.w[a][0] = wa * wa * 0.5f;
    s.w[a][1] = 0.75f - wb * wb;
    s.w[a][2] = wc * wc * 0.5f;
  }
  return s;
}

__device__ __forceinline__ bool node_in_range(const neuma_mpm_const& c, int ix, int iy, int iz) {
  // The reference does not bounds-check (Warp release build): a particle at the
  // upper clamp writes nodes G, G+1 ("latent OOB").  Here such node
  // visits are dropped, in P2G, G2P and both adjoints alike.
  return (unsigned)ix < (unsigned)c.num_grids && (unsigned)iy < (unsigned)c.num_grids &&
         (unsigned)(iz - c.z_off) < (unsigned)c.nz;
}
__device__ __forceinline__ size_t node_index(const neuma_mpm_const& c, int ix, int iy, int iz) {
  return ((size_t)(iz - c.z_off) * c.num_grids + iy) * c.num_grids + ix;
}

// grid_op on one node (mpm.py:373-429).  Returns v; `pass` gets 1/0 per
// component: whether dL/dv flows back to (mv, m) (0 where the wall BC zeroed it
// or the node is empty).  `vpre_nog` = mv / (m + eps) (needed by the adjoint).
__device__ __forceinline__ void node_velocity(const neuma_mpm_const& c, const float4 mvm, int ix,
                                              int iy, int iz, float v[3], float pass[3],
                                              float& inv_m, float vpre_nog[3]) {
  const float gdt[3] = {c.gravity[0] * c.dt, c.gravity[1] * c.dt, c.gravity[2] * c.dt};
  const float mv[3] = {mvm.x, mvm.y, mvm.z};
  const bool has_mass = mvm.w > 0.0f;
  inv_m = has_mass ? 1.0f / (mvm.w + c.eps) : 0.0f;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    vpre_nog[a] = has_mass ? mv[a] * inv_m : 0.0f;
    v[a] = has_mass ? (mv[a] / (mvm.w + c.eps) + gdt[a]) : gdt[a];
    pass[a] = has_mass ? 1.0f : 0.0f;
  }
  const int idx[3] = {ix, iy, iz};
  const int hi = c.num_grids - c.bound;
  if (c.bc == 0) {
    // noslip: the six tests run in sequence on the current v; once v is zeroed
    // every later test is false, so "any test fires on the pre-BC v" is equivalent.
    bool zero = false;
#pragma unroll
    for (int a = 0; a < 3; ++a) zero |= (idx[a] < c.bound && v[a] < 0.0f);
#pragma unroll
    for (int a = 0; a < 3; ++a) zero |= (idx[a] >= hi && v[a] > 0.0f);
    if (zero) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        v[a] = 0.0f;
        pass[a] = 0.0f;
      }
    }
  } else {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (idx[a] < c.bound && v[a] < 0.0f) {
        v[a] = 0.0f;
        pass[a] = 0.0f;
      }
      if (idx[a] >= hi && v[a] > 0.0f) {
        v[a] = 0.0f;
        pass[a] = 0.0f;
      }
    }
  }
}


// one particle's 27 node visits of p2g (mpm.py:321-371) into a packed (mv, m) grid
__device__ __forceinline__ void p2g_particle(const neuma_mpm_const& c, int p,
                                             const float* __restrict__ x,
                                             const float* __restrict__ v,
                                             const float* __restrict__ C,
                                             const float* __restrict__ stress, float p_vol,
                                             float p_rho, float4* __restrict__ grid) {
  const float p_mass = p_vol * p_rho;
  const Stencil s = make_stencil(x[3 * p], x[3 * p + 1], x[3 * p + 2], c.inv_dx);
  const float k = -c.dt * p_vol * 4.0f * c.inv_dx * c.inv_dx;
  float A[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) A[i] = k * stress[9 * p + i] + p_mass * C[9 * p + i];
  const float pv[3] = {p_mass * v[3 * p], p_mass * v[3 * p + 1], p_mass * v[3 * p + 2]};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int kk = 0; kk < 3; ++kk) {
        const int ix = s.b[0] + i, iy = s.b[1] + j, iz = s.b[2] + kk;
        if (!node_in_range(c, ix, iy, iz)) continue;
        const float d0 = ((float)i - s.f[0]) * c.dx;
        const float d1 = ((float)j - s.f[1]) * c.dx;
        const float d2 = ((float)kk - s.f[2]) * c.dx;
        const float w = s.w[0][i] * s.w[1][j] * s.w[2][kk];
        const float m0 = w * (pv[0] + (A[0] * d0 + A[1] * d1 + A[2] * d2));
        const float m1 = w * (pv[1] + (A[3] * d0 + A[4] * d1 + A[5] * d2));
        const float m2 = w * (pv[2] + (A[6] * d0 + A[7] * d1 + A[8] * d2));
        red_add_f4(grid + node_index(c, ix, iy, iz), m0, m1, m2, w * p_mass);
      }
}


// Upstream gradients of one particle's G2P outputs folded into the two
// quantities the grid sees: gvt = dL/dv' (incl. the x' = clamp(x + dt v') path)
// and gCt = dL/dC' (incl. the F' = (I + dt C') F path).
struct G2PAdj {
  float gxn[3];
  float gvt[3];
  float gCt[9];
};
__device__ __forceinline__ G2PAdj g2p_adjoint_inputs(const neuma_mpm_const& c, int p,
                                                     const float* __restrict__ x,
                                                     const float* __restrict__ F,
                                                     const float* __restrict__ clip,
                                                     const float* __restrict__ v2,
                                                     const float* __restrict__ gx2,
                                                     const float* __restrict__ gv2,
                                                     const float* __restrict__ gC2,
                                                     const float* __restrict__ gF2) {
  G2PAdj r;
  const float bnd = clip[p] * c.dx;
  const float lo = 0.0f + bnd, hi = 1.0f - bnd;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float nx = x[3 * p + a] + c.dt * v2[3 * p + a];
    // adjoint of wp.clamp: passes the gradient iff lo <= nx <= hi
    r.gxn[a] = (nx >= lo && nx <= hi) ? gx2[3 * p + a] : 0.0f;
    r.gvt[a] = gv2[3 * p + a] + c.dt * r.gxn[a];
  }
  // gCt = gC' + dt * gF' F^T
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) acc += gF2[9 * p + 3 * a + k] * F[9 * p + 3 * b + k];
      r.gCt[3 * a + b] = gC2[9 * p + 3 * a + b] + c.dt * acc;
    }
  return r;
}


}  // namespace
