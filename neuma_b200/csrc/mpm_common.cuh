// Device code shared by the MPM kernels (mpm.cu: stand-alone / per-substep entry points on
// the ctx grid; mpm_rollout.cu: the rollout path with caller-owned per-substep grids).
// Arithmetic follows modules/nclaw/sim/mpm.py:321-498.  ONE copy of every per-particle body
// (P2G visit loop, G2P gather, adjoint scatter, adjoint gather); the kernels are thin wrappers.
#pragma once
#include "common.cuh"
#include <limits.h>

namespace {

struct Stencil {
  int b[3];       // base node (trunc toward zero, mpm.py:337-339)
  float f[3];     // fx = x*inv_dx - base
  float w[3][3];  // w[axis][i], quadratic B-spline (mpm.py:345-354)
};

// base node of a coordinate: int(x * inv_dx - 0.5), truncation toward zero (mpm.py:337-339).
// Explicitly un-fused (product rounded, then the subtraction), like the reference's fp32
// expression evaluated op by op, and so that EVERY kernel (stencil, box, cell key, slab
// selection) derives the identical integer from the same x.
__device__ __forceinline__ int base_node(float x, float inv_dx) {
  return (int)__fsub_rn(__fmul_rn(x, inv_dx), 0.5f);
}

__device__ __forceinline__ Stencil make_stencil(float x0, float x1, float x2, float inv_dx) {
  Stencil s;
  const float p[3] = {__fmul_rn(x0, inv_dx), __fmul_rn(x1, inv_dx), __fmul_rn(x2, inv_dx)};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    s.b[a] = (int)__fsub_rn(p[a], 0.5f);  // int() truncates, it is not floor
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

// ---- where a node lives in a grid buffer --------------------------------------------------
// dense   (c.compact == 0): the buffer holds planes [z_off, z_off + nz) of the G x G x G grid,
//          plane-major (z slowest) -- the per-substep entry points and the Z-slab path, whose
//          halo planes must sit at fixed addresses;
// compact (c.compact != 0): the buffer holds ONLY the bounding box of the occupied nodes
//          (`box`, on the device, x-range widened to 128-byte rows), z slowest inside the box:
//          S saved grids of a rollout cost S x |box| x 16 B instead of S x G^3 x 16 B
//          (SURVEY 8f.1).  A box larger than c.cap_nodes is an overflow: the map is empty, every
//          kernel skips its work and box[6] is raised by box_clear (checked on the host).
struct GridMap {
  int x0, y0, z0, nx, ny, nz;
};
struct Box {
  int lo[3], hi[3];
};
__device__ __forceinline__ Box load_box(const int32_t* __restrict__ b) {
  Box r;
#pragma unroll
  for (int a = 0; a < 3; ++a) { r.lo[a] = b[a]; r.hi[a] = b[3 + a]; }
  return r;
}
__device__ __forceinline__ bool box_empty(const Box& b) {
  return b.lo[0] > b.hi[0] || b.lo[1] > b.hi[1] || b.lo[2] > b.hi[2];
}
// the box as an iteration window: x-range widened to 8 nodes = 128-byte rows
__device__ __forceinline__ GridMap box_window(const neuma_mpm_const& c, const Box& b) {
  GridMap m;
  if (box_empty(b)) { m.x0 = m.y0 = m.z0 = 0; m.nx = m.ny = m.nz = 0; return m; }
  m.x0 = b.lo[0] & ~7;
  const int x1 = c.compact ? (b.hi[0] | 7) : min(c.num_grids - 1, b.hi[0] | 7);
  m.nx = x1 - m.x0 + 1;
  m.y0 = b.lo[1]; m.ny = b.hi[1] - b.lo[1] + 1;
  m.z0 = b.lo[2]; m.nz = b.hi[2] - b.lo[2] + 1;
  return m;
}
__device__ __forceinline__ long long map_nodes(const GridMap& m) {
  return (long long)m.nx * m.ny * m.nz;
}
__device__ __forceinline__ bool map_overflows(const neuma_mpm_const& c, const GridMap& m) {
  return c.compact && map_nodes(m) > (long long)c.cap_nodes;
}
__device__ __forceinline__ GridMap grid_map(const neuma_mpm_const& c, const int32_t* __restrict__ box) {
  if (!c.compact) return GridMap{0, 0, c.z_off, c.num_grids, c.num_grids, c.nz};
  GridMap m = box_window(c, load_box(box));
  if (map_overflows(c, m)) m.nx = m.ny = m.nz = 0;
  return m;
}
__device__ __forceinline__ bool node_in_range(const GridMap& m, int ix, int iy, int iz) {
  // The reference does not bounds-check (Warp release build): a particle at the upper clamp
  // writes nodes G, G+1 ("latent OOB").  Here such node visits are dropped, in P2G,
  // G2P and both adjoints alike (a compact box is already clipped to the domain).
  return (unsigned)(ix - m.x0) < (unsigned)m.nx && (unsigned)(iy - m.y0) < (unsigned)m.ny &&
         (unsigned)(iz - m.z0) < (unsigned)m.nz;
}
__device__ __forceinline__ size_t node_index(const GridMap& m, int ix, int iy, int iz) {
  return ((size_t)(iz - m.z0) * m.ny + (iy - m.y0)) * m.nx + (ix - m.x0);
}
// A valid index for ANY node: out-of-range coordinates are clamped to the nearest node of the map
// (index 0 for an empty map).  The gathers load through it unconditionally and zero what they loaded
// when !node_in_range -- a load behind a branch cannot be hoisted, and 27 of them in a row cost 27
// serialized L2 round trips per warp (ncu: profiles/r03_gather_latency.txt).
__device__ __forceinline__ size_t node_index_clamped(const GridMap& m, int ix, int iy, int iz) {
  const int cx = min(max(ix - m.x0, 0), max(m.nx - 1, 0));
  const int cy = min(max(iy - m.y0, 0), max(m.ny - 1, 0));
  const int cz = min(max(iz - m.z0, 0), max(m.nz - 1, 0));
  return ((size_t)cz * m.ny + cy) * m.nx + cx;
}
// A/B switch of the gathers' node loads (tools/gather_variants.sh; the default is what ships):
//   0 = a load per visited node behind its range check, 1 = unconditional clamped loads,
//   2 = 1 + the whole stencil prefetched to L1 first
#ifndef NEUMA_GATHER_VARIANT
#define NEUMA_GATHER_VARIANT 1
#endif

// The 27 node indices of a particle's stencil as 9 row bases + 3 x offsets (12 registers), each
// clamped to the map, and the per-axis in-range bits.
struct StencilIdx {
  int row[9];    // [3 * kk + j]: index of (x0, b1 + j, b2 + kk), y / z clamped
  int cx[3];     // clamped x offset of b0 + i
  unsigned ok;   // bit i: x in range, bit 3 + j: y, bit 6 + kk: z
  __device__ __forceinline__ bool in_range(int i, int j, int kk) const {
    return ((ok >> i) & (ok >> (3 + j)) & (ok >> (6 + kk)) & 1u) != 0;
  }
  __device__ __forceinline__ int at(int i, int j, int kk) const { return row[3 * kk + j] + cx[i]; }
};
__device__ __forceinline__ StencilIdx make_stencil_idx(const GridMap& m, const int b[3]) {
  StencilIdx si;
  si.ok = 0;
  int cy[3], cz[3];
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    const int rx = b[0] + t - m.x0, ry = b[1] + t - m.y0, rz = b[2] + t - m.z0;
    si.ok |= ((unsigned)rx < (unsigned)m.nx ? 1u : 0u) << t;
    si.ok |= ((unsigned)ry < (unsigned)m.ny ? 1u : 0u) << (3 + t);
    si.ok |= ((unsigned)rz < (unsigned)m.nz ? 1u : 0u) << (6 + t);
    si.cx[t] = min(max(rx, 0), max(m.nx - 1, 0));
    cy[t] = min(max(ry, 0), max(m.ny - 1, 0));
    cz[t] = min(max(rz, 0), max(m.nz - 1, 0));
  }
#pragma unroll
  for (int kk = 0; kk < 3; ++kk)
#pragma unroll
    for (int j = 0; j < 3; ++j) si.row[3 * kk + j] = (cz[kk] * m.ny + cy[j]) * m.nx;
  return si;
}
// pull the 27 stencil nodes of a particle towards L1 in ONE round trip (a row's 3 nodes are 48
// contiguous bytes: its two ends cover the one or two 128-byte lines it lies in).  The gathers are short
// of registers and of warps (100 k particles are 21 warps per SM): the compiler issues their node loads
// a few at a time and, without this, the warp pays a memory latency per group.
__device__ __forceinline__ void prefetch_stencil(const StencilIdx& si, const float4* __restrict__ g0,
                                                 const float4* __restrict__ g1) {
#pragma unroll
  for (int r = 0; r < 9; ++r) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(g0 + si.row[r] + si.cx[0]));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(g0 + si.row[r] + si.cx[2]));
    if (g1) {
      asm volatile("prefetch.global.L1 [%0];" ::"l"(g1 + si.row[r] + si.cx[0]));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(g1 + si.row[r] + si.cx[2]));
    }
  }
}
__device__ __forceinline__ float4 select4(bool ok, const float4& q) {
  return make_float4(ok ? q.x : 0.f, ok ? q.y : 0.f, ok ? q.z : 0.f, ok ? q.w : 0.f);
}

// ---- Z-slab overlap selection (neuma_mpm_const::sel_*): planes below sel_lo_end and from
// sel_up_start on are shared with a neighbour rank.
__device__ __forceinline__ bool plane_selected(const neuma_mpm_const& c, int iz) {
  if (c.sel_mode == 0) return true;
  const bool shared = iz < c.sel_lo_end || iz >= c.sel_up_start;
  return (c.sel_mode == 1) == shared;
}
__device__ __forceinline__ bool particle_selected(const neuma_mpm_const& c, float z) {
  if (c.sel_mode == 0 || c.sel_mode == 3) return true;
  const int bz = base_node(z, c.inv_dx);   // the stencil touches planes bz .. bz + 2
  const bool shared = bz < c.sel_lo_end || bz + 2 >= c.sel_up_start;
  return (c.sel_mode == 1) == shared;
}
// sel_mode 3: the caller launched this kernel over particles it CLASSIFIED as interior (its scatter
// runs next to the halo exchange of the boundary class).  One that touches a shared plane after all
// would be missed by the neighbour: count it in box[7] so that the host raises (SlabRollout.check_drift).
__device__ __forceinline__ void check_interior_class(const neuma_mpm_const& c, float z, const int32_t* box) {
  if (c.sel_mode != 3) return;
  const int bz = base_node(z, c.inv_dx);
  if (bz < c.sel_lo_end || bz + 2 >= c.sel_up_start) atomicAdd(const_cast<int32_t*>(box) + 7, 1);
}

// grid_op on one node (mpm.py:373-429).  Returns v; `pass` gets 1/0 per component: whether
// dL/dv flows back to (mv, m) (0 where the wall BC zeroed it or the node is empty).
// `vpre_nog` = mv / (m + eps) (needed by the adjoint).
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

// the rollout grids store a BC-zeroed velocity component as -0.0f (numerically 0 for G2P, and
// exactly the "no gradient passes" mask the adjoint needs)
constexpr uint32_t NEG_ZERO = 0x80000000u;

// ---- P2G: one particle's 27 node visits (mpm.py:321-371) into a packed (mv, m) grid --------
// `stress9`: the particle's 9 stress values (a row of the stress array, or registers of the kernel
// that just computed them -- constitutive.cu fuses this into the elasticity kernel's epilogue)
__device__ __forceinline__ void p2g_particle(const neuma_mpm_const& c, const GridMap& gm, int p,
                                             const float* __restrict__ x,
                                             const float* __restrict__ v,
                                             const float* __restrict__ C,
                                             const float* stress9, float p_vol,
                                             float p_rho, float4* __restrict__ grid) {
  const float p_mass = p_vol * p_rho;
  const Stencil s = make_stencil(x[3 * p], x[3 * p + 1], x[3 * p + 2], c.inv_dx);
  const float k = -c.dt * p_vol * 4.0f * c.inv_dx * c.inv_dx;
  float A[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) A[i] = k * stress9[i] + p_mass * C[9 * p + i];
  const float pv[3] = {p_mass * v[3 * p], p_mass * v[3 * p + 1], p_mass * v[3 * p + 2]};
  // mv_n += w (m v + affine dpos_n), dpos_n = (o_n - f) dx with integer offsets o_n: the bracket is
  // Z_n = Z0 + As o_n (As = dx affine, Z0 = m v - As f), built up along the unrolled loops by adding
  // matrix columns -- 3 adds per visit instead of a 3x3 matrix-vector product
  float As[9], Z0[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float af = 0.f;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      As[3 * a + b] = c.dx * A[3 * a + b];
      af += As[3 * a + b] * s.f[b];
    }
    Z0[a] = pv[a] - af;
  }
#pragma unroll
  for (int kk = 0; kk < 3; ++kk) {
    float Zz[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) Zz[a] = Z0[a] + (float)kk * As[3 * a + 2];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float Zy[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) Zy[a] = Zz[a] + (float)j * As[3 * a + 1];
      const float wjk = s.w[1][j] * s.w[2][kk];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int ix = s.b[0] + i, iy = s.b[1] + j, iz = s.b[2] + kk;
        if (!node_in_range(gm, ix, iy, iz)) continue;
        const float w = s.w[0][i] * wjk;
        red_add_f4(grid + node_index(gm, ix, iy, iz), w * (Zy[0] + (float)i * As[0]), w * (Zy[1] + (float)i * As[3]),
                   w * (Zy[2] + (float)i * As[6]), w * p_mass);
      }
    }
  }
}

// ---- G2P: one particle (mpm.py:432-498).  LAZY: the grid holds (mv, m) and grid_op is
// evaluated per gathered node; otherwise it holds v.  Writes x', v', C', F'; returns x' in nx3 and,
// when Fret is given, F' in registers (the fused plasticity kernel goes on with it).
// x2 / F2 MAY ALIAS x / F (MPMForwardSim steps in place, interface.py:133): no __restrict__ on them, and the
// row of F is read into registers before anything is written.
template <bool LAZY>
__device__ __forceinline__ void g2p_particle(const neuma_mpm_const& c, const GridMap& gm, int p,
                                             const float* x, const float* F,
                                             const float* __restrict__ clip,
                                             const float4* __restrict__ grid, float* x2,
                                             float* v2, float* C2,
                                             float* F2, float nx3[3], float* Fret = nullptr) {
  const float px[3] = {x[3 * p], x[3 * p + 1], x[3 * p + 2]};
  const Stencil s = make_stencil(px[0], px[1], px[2], c.inv_dx);
  float nv[3] = {0.f, 0.f, 0.f};
  float nC[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const float k4 = 4.0f * c.inv_dx * c.inv_dx;
  const StencilIdx si = make_stencil_idx(gm, s.b);
  if (NEUMA_GATHER_VARIANT >= 2) prefetch_stencil(si, grid, nullptr);
#pragma unroll
  for (int kk = 0; kk < 3; ++kk) {
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int ix = s.b[0] + i, iy = s.b[1] + j, iz = s.b[2] + kk;
        const bool ok = si.in_range(i, j, kk);           // dropped visits contribute exact zeros
#if NEUMA_GATHER_VARIANT == 0
        if (!ok) continue;
#endif
        const float4 q = select4(ok, grid[si.at(i, j, kk)]);
        float gv[3];
        if (LAZY) {
          float pass[3], inv_m, vp[3];
          node_velocity(c, q, ix, iy, iz, gv, pass, inv_m, vp);
        } else {
          gv[0] = q.x; gv[1] = q.y; gv[2] = q.z;
        }
        const float d[3] = {((float)i - s.f[0]) * c.dx, ((float)j - s.f[1]) * c.dx,
                            ((float)kk - s.f[2]) * c.dx};
        const float w = ok ? s.w[0][i] * s.w[1][j] * s.w[2][kk] : 0.0f;
        const float w4 = k4 * w;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          nv[a] += w * gv[a];
#pragma unroll
          for (int b = 0; b < 3; ++b) nC[3 * a + b] += w4 * (gv[a] * d[b]);
        }
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
    nx3[a] = nx;
    x2[3 * p + a] = nx;
    v2[3 * p + a] = nv[a];
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    C2[9 * p + i] = nC[i];
    F2[9 * p + i] = Fn[i];
    if (Fret) Fret[i] = Fn[i];
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
                                                     const float* gF2row /* dL/dF' of THIS particle */) {
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
      for (int k = 0; k < 3; ++k) acc += gF2row[3 * a + k] * F[9 * p + 3 * b + k];
      r.gCt[3 * a + b] = gC2[9 * p + 3 * a + b] + c.dt * acc;
    }
  return r;
}

// ---- backward pass 1: scatter adj(G2P)'s dL/dv_node of one particle into gadj (xyz; w unused)
__device__ __forceinline__ void g2p_adjoint_scatter(const neuma_mpm_const& c, const GridMap& gm, int p,
                                                    const float* __restrict__ x, const G2PAdj& g,
                                                    float4* __restrict__ gadj) {
  const Stencil s = make_stencil(x[3 * p], x[3 * p + 1], x[3 * p + 2], c.inv_dx);
  const float k4 = 4.0f * c.inv_dx * c.inv_dx;
  // dL/dv_n += w (gvt + k4 gCt dpos_n) = w U_n,  U_n = U0 + Gs o_n  (integer offsets, see p2g_particle)
  float Gs[9], U0[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float cf = 0.f;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      Gs[3 * a + b] = (k4 * c.dx) * g.gCt[3 * a + b];
      cf += Gs[3 * a + b] * s.f[b];
    }
    U0[a] = g.gvt[a] - cf;
  }
#pragma unroll
  for (int kk = 0; kk < 3; ++kk) {
    float Uz[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) Uz[a] = U0[a] + (float)kk * Gs[3 * a + 2];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float Uy[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) Uy[a] = Uz[a] + (float)j * Gs[3 * a + 1];
      const float wjk = s.w[1][j] * s.w[2][kk];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int ix = s.b[0] + i, iy = s.b[1] + j, iz = s.b[2] + kk;
        if (!node_in_range(gm, ix, iy, iz)) continue;
        const float w = s.w[0][i] * wjk;
        red_add_f4(gadj + node_index(gm, ix, iy, iz), w * (Uy[0] + (float)i * Gs[0]), w * (Uy[1] + (float)i * Gs[3]),
                   w * (Uy[2] + (float)i * Gs[6]), 0.0f);
      }
    }
  }
}

// ---- backward pass 3, one particle: gather its 27 nodes and accumulate adj(G2P) (the terms
// that need v_node) and adj(P2G).
//   LAZY  : `grid` holds (mv, m), `gadj` holds dL/dv_node; grid_op and its adjoint are evaluated
//           per visit (per-substep entry points: no node pass);
//   !LAZY : `grid` holds (v, m), `gadj` holds (dL/dmv, dL/dm) left by the node pass (rollout).
// HAVE_C2: C2 is the forward's C'; otherwise it is recomputed from the gathered velocities.
template <bool LAZY, bool HAVE_C2>
__device__ __forceinline__ void bwd_gather_particle(
    const neuma_mpm_const& c, const GridMap& gm, int p, const float* __restrict__ x,
    const float* __restrict__ v, const float* __restrict__ C, const float* __restrict__ F,
    const float* __restrict__ stress, const float* __restrict__ vol, const float* __restrict__ rho,
    const float* __restrict__ clip, const float* __restrict__ v2, const float* __restrict__ C2,
    const float* __restrict__ gx2, const float* __restrict__ gv2, const float* __restrict__ gC2,
    const float* __restrict__ gF2, const float4* __restrict__ grid, const float4* __restrict__ gadj,
    float* __restrict__ gx, float* __restrict__ gv, float* __restrict__ gC, float* __restrict__ gF,
    float* __restrict__ gstress) {
  const Stencil s = make_stencil(x[3 * p], x[3 * p + 1], x[3 * p + 2], c.inv_dx);
  const StencilIdx si = make_stencil_idx(gm, s.b);
  if (NEUMA_GATHER_VARIANT >= 2) prefetch_stencil(si, grid, gadj);
  const float p_vol = vol[p];
  const float p_mass = p_vol * rho[p];
  const float ks = -c.dt * p_vol * 4.0f * c.inv_dx * c.inv_dx;
  const float k4 = 4.0f * c.inv_dx * c.inv_dx;
  float A[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) A[i] = ks * stress[9 * p + i] + p_mass * C[9 * p + i];
  const float pv[3] = {p_mass * v[3 * p], p_mass * v[3 * p + 1], p_mass * v[3 * p + 2]};
  const G2PAdj g = g2p_adjoint_inputs(c, p, x, F, clip, v2, gx2, gv2, gC2, gF2 + 9 * (size_t)p);
  constexpr bool need_C = !HAVE_C2;

  // dpos_n = (o_n - f) dx with INTEGER stencil offsets o_n in {0,1,2}^3: everything that is linear in
  // dpos splits into a per-particle part and a part that is a small-integer combination of matrix
  // columns, so the per-(particle, node) work is ~35 flops instead of ~95:
  //   sum_n w gmv (x) dpos      = dx (M1 - acc_gmv (x) f),     M1 = sum_n w gmv (x) o
  //   sum_n w (k4 gCt^T v_n + A^T gmv_n) = k4 gCt^T V0 + A^T acc_gmv,   V0 = sum_n w v_n
  //   gW_n = v_n . U_n + gmv_n . Z_n + gm_n m_p,   U_n = U0 + Gs o_n,  Z_n = Z0 + As o_n
  float Gs[9], As[9], U0[3], Z0[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float cf = 0.f, af = 0.f;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      Gs[3 * a + b] = (k4 * c.dx) * g.gCt[3 * a + b];
      As[3 * a + b] = c.dx * A[3 * a + b];
      cf += Gs[3 * a + b] * s.f[b];
      af += As[3 * a + b] * s.f[b];
    }
    U0[a] = g.gvt[a] - cf;
    Z0[a] = pv[a] - af;
  }
  float acc_gmv[3] = {0.f, 0.f, 0.f};                                 // sum_n w gmv
  float V0[3] = {0.f, 0.f, 0.f};                                      // sum_n w v_n
  float M1[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};        // sum_n w gmv (x) o
  float Mv[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};        // sum_n w v_n (x) o   (only if C' is recomputed)
  float gw[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
#pragma unroll
  for (int kk = 0; kk < 3; ++kk) {
    float Uz[3], Zz[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      Uz[a] = U0[a] + (float)kk * Gs[3 * a + 2];
      Zz[a] = Z0[a] + (float)kk * As[3 * a + 2];
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float Uy[3], Zy[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        Uy[a] = Uz[a] + (float)j * Gs[3 * a + 1];
        Zy[a] = Zz[a] + (float)j * As[3 * a + 1];
      }
      const float wjk = s.w[1][j] * s.w[2][kk];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int ix = s.b[0] + i, iy = s.b[1] + j, iz = s.b[2] + kk;
        const bool ok = si.in_range(i, j, kk);           // dropped visits contribute exact zeros
#if NEUMA_GATHER_VARIANT == 0
        if (!ok) continue;
#endif
        const float4 q = select4(ok, grid[si.at(i, j, kk)]);
        const float4 ga = select4(ok, gadj[si.at(i, j, kk)]);
        float vn[3], gmv[3], gmn;
        if (LAZY) {
          float pass[3], inv_m, vp[3];
          node_velocity(c, q, ix, iy, iz, vn, pass, inv_m, vp);
          // adj(grid_op): v = mv/(m+eps) + g dt
          const float gvn[3] = {ga.x * pass[0], ga.y * pass[1], ga.z * pass[2]};
          gmv[0] = gvn[0] * inv_m; gmv[1] = gvn[1] * inv_m; gmv[2] = gvn[2] * inv_m;
          gmn = -(vp[0] * gvn[0] + vp[1] * gvn[1] + vp[2] * gvn[2]) * inv_m;
        } else {
          vn[0] = q.x; vn[1] = q.y; vn[2] = q.z;
          gmv[0] = ga.x; gmv[1] = ga.y; gmv[2] = ga.z;
          gmn = ga.w;
        }
        const float w = ok ? s.w[0][i] * wjk : 0.0f;
        const float off[3] = {(float)i, (float)j, (float)kk};   // compile-time 0 / 1 / 2 after unrolling
        float gW = gmn * p_mass;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const float U = Uy[a] + off[0] * Gs[3 * a], Z = Zy[a] + off[0] * As[3 * a];
          gW += vn[a] * U + gmv[a] * Z;
          const float t = w * gmv[a];
          acc_gmv[a] += t;
          V0[a] += w * vn[a];
#pragma unroll
          for (int b = 0; b < 3; ++b) {
            M1[3 * a + b] += off[b] * t;
            if (need_C) Mv[3 * a + b] += off[b] * (w * vn[a]);
          }
        }
        if (LAZY) gW = ok ? gW : 0.0f;   // (!LAZY: q = ga = 0 already makes it an exact zero)
        gw[0][i] += gW * wjk;
        gw[1][j] += gW * (s.w[0][i] * s.w[2][kk]);
        gw[2][kk] += gW * (s.w[0][i] * s.w[1][j]);
      }
    }
  }
  float gA[9], gd[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    gd[a] = k4 * (g.gCt[a] * V0[0] + g.gCt[3 + a] * V0[1] + g.gCt[6 + a] * V0[2]) +
            (A[a] * acc_gmv[0] + A[3 + a] * acc_gmv[1] + A[6 + a] * acc_gmv[2]);
#pragma unroll
    for (int b = 0; b < 3; ++b) gA[3 * a + b] = c.dx * (M1[3 * a + b] - s.f[b] * acc_gmv[a]);
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
    float nC[9];
    if (!need_C) {
#pragma unroll
      for (int a = 0; a < 9; ++a) nC[a] = C2[9 * p + a];
    } else {
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) nC[3 * a + b] = (k4 * c.dx) * (Mv[3 * a + b] - s.f[b] * V0[a]);
    }
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

// ---- coalesced row I/O.  The particle arrays are arrays of 3- or 9-float rows; a thread reading "its"
// row makes every warp instruction touch 12 / 32 sectors (ncu on g2p_from_kernel: 18 sectors per load
// request, 27 per store request -- the kernels were bound by L1 sector lookups, not by latency or
// bandwidth).  A block copies the rows of its B consecutive particles between global and shared memory
// with 128 contiguous bytes per warp instruction (4 sectors) and the per-particle code runs on the
// shared copy (stride-K words with K odd: no bank conflicts).
template <int K, int B>
__device__ __forceinline__ void rows_load(float* __restrict__ sm, const float* __restrict__ g, int p0, int cnt) {
  const float* __restrict__ src = g + (size_t)p0 * K;
  const int nf = cnt * K;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const int i = j * B + (int)threadIdx.x;
    if (i < nf) sm[i] = src[i];
  }
}
template <int K, int B>
__device__ __forceinline__ void rows_store(float* __restrict__ g, const float* __restrict__ sm, int p0, int cnt) {
  float* __restrict__ dst = g + (size_t)p0 * K;
  const int nf = cnt * K;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const int i = j * B + (int)threadIdx.x;
    if (i < nf) dst[i] = sm[i];
  }
}
// one thread's row (the fallback when only some particles of the block write)
template <int K>
__device__ __forceinline__ void row_store(float* __restrict__ g, const float* __restrict__ sm, int p, int t) {
#pragma unroll
  for (int j = 0; j < K; ++j) g[(size_t)p * K + j] = sm[t * K + j];
}

// disabled particles: the reference's kernels return early, no adjoint contribution
__device__ __forceinline__ void zero_particle_grads(int p, float* gx, float* gv, float* gC, float* gF,
                                                    float* gstress) {
  if (gx) for (int a = 0; a < 3; ++a) gx[3 * p + a] = 0.f;
  if (gv) for (int a = 0; a < 3; ++a) gv[3 * p + a] = 0.f;
  if (gC) for (int a = 0; a < 9; ++a) gC[9 * p + a] = 0.f;
  if (gF) for (int a = 0; a < 9; ++a) gF[9 * p + a] = 0.f;
  if (gstress) for (int a = 0; a < 9; ++a) gstress[9 * p + a] = 0.f;
}

}  // namespace
