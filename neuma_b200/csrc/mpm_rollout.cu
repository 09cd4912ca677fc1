// Rollout path of the MLS-MPM substep for sm_100a: the same arithmetic as mpm.cu
// (modules/nclaw/sim/mpm.py:279-498) laid out for a differentiable S-substep rollout
// (experiments/finetune.py:360-414) on one B200:
//
//   * every substep writes its grid into its OWN caller-owned slot (S x G^3 x 16 B: 0.7 GB
//     for S = 20 at 128^3 -- nothing next to 180 GB of HBM), so the backward reads the
//     forward's grid instead of re-running P2G (the reference recomputes it, mpm.py:312-315);
//   * grid_op runs once per node over the particles' bounding box, IN PLACE:
//     (mv, m) -> (v, m).  A component the wall BC zeroed is stored as -0.0f, which is
//     numerically 0 for G2P and tells the adjoint exactly which components pass a
//     gradient (a naturally computed -0.0f cannot occur: the grid starts at +0 and
//     +0 + -0 = +0);
//   * clears and node passes touch only the bounding box of the occupied nodes, kept
//     on the device (int32[8]: lo.xyz, hi.xyz inclusive) and updated by G2P for the
//     next substep -- exact, no host round trip, CUDA-graph friendly;
//   * the backward is  scatter adj(G2P) -> one node pass (adj(grid_op)) -> gather,
//     and the per-(particle, node) work holds no divisions.
//
// Particles are expected in cell (Morton) order (neuma_mpm_cell_order): consecutive
// lanes then touch neighbouring nodes, so the 27 node visits of a warp coalesce in L1/L2.
#include "mpm_common.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <stdlib.h>

namespace {

__global__ void box_init_kernel(int32_t* box) {
  if (threadIdx.x < 3) box[threadIdx.x] = INT_MAX;
  else if (threadIdx.x < 6) box[threadIdx.x] = INT_MIN;
  else if (threadIdx.x < 8) box[threadIdx.x] = 0;
}

// folds one particle's node range [base, base+2] (clamped to the slab) into the box
__device__ __forceinline__ void box_accumulate(const neuma_mpm_const& c, bool active, float x0,
                                               float x1, float x2, int32_t* __restrict__ box,
                                               int* sbox /* shared int[6] */) {
  int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
  if (active) {
    const float p[3] = {x0, x1, x2};
    const int mn[3] = {0, 0, c.z_off}, mx[3] = {c.num_grids - 1, c.num_grids - 1, c.z_off + c.nz - 1};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int b = base_node(p[a], c.inv_dx);
      lo[a] = max(b, mn[a]);
      hi[a] = min(b + 2, mx[a]);
    }
    // Z-slab ownership: a particle whose stencil needs a plane this rank does not
    // hold would silently lose node visits -- count it in box[7]; the host raises
    // (SlabRollout.check_drift).  The top of the DOMAIN is the reference's latent OOB, not drift.
    const int bz = base_node(p[2], c.inv_dx);
    const bool below = bz < c.z_off;
    const bool above = bz + 2 > c.z_off + c.nz - 1 && c.z_off + c.nz < c.num_grids;
    if (below || above) atomicAdd(&box[7], 1);
  }
  if (threadIdx.x < 6) sbox[threadIdx.x] = threadIdx.x < 3 ? INT_MAX : INT_MIN;
  __syncthreads();
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int l = __reduce_min_sync(0xffffffffu, lo[a]);
    const int h = __reduce_max_sync(0xffffffffu, hi[a]);
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&sbox[a], l);
      atomicMax(&sbox[3 + a], h);
    }
  }
  __syncthreads();
  if (threadIdx.x < 3) atomicMin(&box[threadIdx.x], sbox[threadIdx.x]);
  else if (threadIdx.x < 6) atomicMax(&box[threadIdx.x], sbox[threadIdx.x]);
}

__global__ void __launch_bounds__(256)
box_reduce_kernel(neuma_mpm_const c, int n, const float* __restrict__ x,
                  const int* __restrict__ enabled, int32_t* __restrict__ box) {
  __shared__ int sbox[6];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = p < n && enabled[p] != 0;
  float px[3] = {0.f, 0.f, 0.f};
  if (active) { px[0] = x[3 * p]; px[1] = x[3 * p + 1]; px[2] = x[3 * p + 2]; }
  box_accumulate(c, active, px[0], px[1], px[2], box, sbox);
}

// iteration space of the node passes: the box window (x-range widened to 128-byte rows)
struct BoxIter {
  GridMap w;       // window iterated over
  GridMap m;       // where the nodes live (== w in compact mode)
  long long total;
};
__device__ __forceinline__ BoxIter box_iter(const neuma_mpm_const& c, const int32_t* __restrict__ box) {
  BoxIter it;
  it.w = box_window(c, load_box(box));
  if (c.compact) {
    if (map_overflows(c, it.w)) it.w.nx = it.w.ny = it.w.nz = 0;
    it.m = it.w;
  } else {
    it.m = GridMap{0, 0, c.z_off, c.num_grids, c.num_grids, c.nz};
  }
  it.total = map_nodes(it.w);
  return it;
}
__device__ __forceinline__ size_t box_node(const BoxIter& it, long long i, int& ix, int& iy, int& iz) {
  const int row = (int)(i / it.w.nx);
  ix = it.w.x0 + (int)(i - (long long)row * it.w.nx);
  const int pl = row / it.w.ny;
  iy = it.w.y0 + (row - pl * it.w.ny);
  iz = it.w.z0 + pl;
  return node_index(it.m, ix, iy, iz);
}

// zero the box region of up to two grids; optionally reset another box to "empty".
// Compact mode: raises box[6] when the box does not fit c.cap_nodes.
__global__ void __launch_bounds__(256)
box_clear_kernel(neuma_mpm_const c, int32_t* __restrict__ box, float4* __restrict__ g0,
                 float4* __restrict__ g1, int32_t* __restrict__ box_reset) {
  pdl_prologue();
  if (box_reset && blockIdx.x == 0 && threadIdx.x < 8)
    box_reset[threadIdx.x] = threadIdx.x < 3 ? INT_MAX : (threadIdx.x < 6 ? INT_MIN : 0);
  if (c.compact && blockIdx.x == 0 && threadIdx.x == 0 &&
      map_overflows(c, box_window(c, load_box(box))))
    box[6] = 1;
  const BoxIter it = box_iter(c, box);
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < it.total;
       i += (long long)gridDim.x * blockDim.x) {
    int ix, iy, iz;
    const size_t ni = box_node(it, i, ix, iy, iz);
    if (g0) g0[ni] = z;
    if (g1) g1[ni] = z;
  }
}

constexpr int SCATTER_BLOCK = 256;   // p2g_to_kernel / bwd_scatter_adj_kernel: rows staged in shared memory
__global__ void __launch_bounds__(SCATTER_BLOCK)
p2g_to_kernel(neuma_mpm_const c, int n, const float* __restrict__ x, const float* __restrict__ v,
              const float* __restrict__ C, const float* __restrict__ stress,
              const float* __restrict__ vol, const float* __restrict__ rho,
              const int* __restrict__ enabled, const int32_t* __restrict__ box,
              float4* __restrict__ grid) {
  pdl_prologue();
  constexpr int B = SCATTER_BLOCK;
  __shared__ float sm[B * 24];                         // x 3, v 3, C 9, stress 9
  float* const sx = sm;
  float* const sv = sx + 3 * B;
  float* const sC = sv + 3 * B;
  float* const ss = sC + 9 * B;
  const int t = threadIdx.x, p0 = blockIdx.x * B, p = p0 + t, cnt = min(B, n - p0);
  const GridMap gm = grid_map(c, box);
  rows_load<3, B>(sx, x, p0, cnt);
  rows_load<3, B>(sv, v, p0, cnt);
  rows_load<9, B>(sC, C, p0, cnt);
  rows_load<9, B>(ss, stress, p0, cnt);
  const int en = p < n ? enabled[p] : 0;
  __syncthreads();
  if (p >= n || en == 0 || !particle_selected(c, sx[3 * t + 2])) return;
  check_interior_class(c, sx[3 * t + 2], box);
  p2g_particle(c, gm, t, sx, sv, sC, ss + 9 * t, vol[p], rho[p], grid);
}

// ---- the north_star's P2G shape, kept as a MEASURED ALTERNATIVE (NEUMA_P2G=tiled; tools/p2g_ab.py,
// profiles/r02_p2g_ab.txt): a CTA takes 256 consecutive particles of the cell-ordered list, stages a
// 16 x 16 x 8-node tile of the grid in shared memory anchored at the chunk's lowest base node,
// accumulates (mv, m) there with shared-memory atomics and flushes the touched nodes with one
// red.global.add.v4.f32 each (visits that fall outside the tile go to global memory directly).
// It trades 27 x 16-byte global reductions per particle (2.6 LSU cycles per lane-visit measured) for
// 108 scalar shared-memory atomics per particle (2 cycles per lane EACH, B300_MICROARCH "ATOMS
// spread-addr") plus the flush -- and loses; see DESIGN.md 5.2.
constexpr int PT_X = 16, PT_Y = 16, PT_Z = 8, PT_NODES = PT_X * PT_Y * PT_Z;
__global__ void __launch_bounds__(256)
p2g_tiled_kernel(neuma_mpm_const c, int n, const float* __restrict__ x, const float* __restrict__ v,
                 const float* __restrict__ C, const float* __restrict__ stress,
                 const float* __restrict__ vol, const float* __restrict__ rho,
                 const int* __restrict__ enabled, const int32_t* __restrict__ box,
                 float4* __restrict__ grid) {
  pdl_prologue();
  __shared__ float4 tile[PT_NODES];
  __shared__ int s_lo[3];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = p < n && enabled[p] != 0 && particle_selected(c, x[3 * p + 2]);
  if (threadIdx.x < 3) s_lo[threadIdx.x] = INT_MAX;
  for (int i = threadIdx.x; i < PT_NODES; i += blockDim.x) tile[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  Stencil s;
  if (active) s = make_stencil(x[3 * p], x[3 * p + 1], x[3 * p + 2], c.inv_dx);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int m = __reduce_min_sync(0xffffffffu, active ? s.b[a] : INT_MAX);
    if ((threadIdx.x & 31) == 0) atomicMin(&s_lo[a], m);
  }
  __syncthreads();
  const int lo[3] = {s_lo[0], s_lo[1], s_lo[2]};
  const GridMap gm = grid_map(c, box);
  if (active) {
    const float p_vol = vol[p], p_mass = p_vol * rho[p];
    const float k = -c.dt * p_vol * 4.0f * c.inv_dx * c.inv_dx;
    float A[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) A[i] = k * stress[9 * p + i] + p_mass * C[9 * p + i];
    const float pv[3] = {p_mass * v[3 * p], p_mass * v[3 * p + 1], p_mass * v[3 * p + 2]};
#pragma unroll
    for (int kk = 0; kk < 3; ++kk)
#pragma unroll
      for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int ix = s.b[0] + i, iy = s.b[1] + j, iz = s.b[2] + kk;
          if (!node_in_range(gm, ix, iy, iz)) continue;
          const float d0 = ((float)i - s.f[0]) * c.dx, d1 = ((float)j - s.f[1]) * c.dx,
                      d2 = ((float)kk - s.f[2]) * c.dx;
          const float w = s.w[0][i] * s.w[1][j] * s.w[2][kk];
          const float m0 = w * (pv[0] + (A[0] * d0 + A[1] * d1 + A[2] * d2));
          const float m1 = w * (pv[1] + (A[3] * d0 + A[4] * d1 + A[5] * d2));
          const float m2 = w * (pv[2] + (A[6] * d0 + A[7] * d1 + A[8] * d2));
          const int lx = ix - lo[0], ly = iy - lo[1], lz = iz - lo[2];
          if ((unsigned)lx < PT_X && (unsigned)ly < PT_Y && (unsigned)lz < PT_Z) {
            float* q = reinterpret_cast<float*>(&tile[(lz * PT_Y + ly) * PT_X + lx]);
            atomicAdd(q, m0); atomicAdd(q + 1, m1); atomicAdd(q + 2, m2); atomicAdd(q + 3, w * p_mass);
          } else {
            red_add_f4(grid + node_index(gm, ix, iy, iz), m0, m1, m2, w * p_mass);
          }
        }
  }
  __syncthreads();
  if (lo[0] == INT_MAX) return;
  for (int i = threadIdx.x; i < PT_NODES; i += blockDim.x) {
    const float4 q = tile[i];
    if (q.w == 0.0f) continue;                        // w p_mass > 0 wherever anything was added
    const int lx = i % PT_X, ly = (i / PT_X) % PT_Y, lz = i / (PT_X * PT_Y);
    red_add_f4(grid + node_index(gm, lo[0] + lx, lo[1] + ly, lo[2] + lz), q.x, q.y, q.z, q.w);
  }
}

// grid_op over the box, in place: (mv, m) -> (v, m); BC-zeroed components become -0.0f.
// Optionally zeroes the same nodes of `adj` (the backward's scatter target).
// next_grid != NULL: also prepares the NEXT substep's grid slot -- zeroes the window of this box
// dilated by one node (the next box lies inside it unless a particle crosses more than a cell in one
// substep, which the fused G2P flags in box_next[6]) -- and resets box_reset to "empty".  With that the
// forward needs no separate clear kernel between G2P and the next P2G.
__global__ void __launch_bounds__(256)
grid_op_box_kernel(neuma_mpm_const c, int32_t* __restrict__ box, float4* __restrict__ grid,
                   float4* __restrict__ adj, float4* __restrict__ next_grid, int32_t* __restrict__ box_reset) {
  pdl_prologue();
  if (box_reset && blockIdx.x == 0 && threadIdx.x < 8)
    box_reset[threadIdx.x] = threadIdx.x < 3 ? INT_MAX : (threadIdx.x < 6 ? INT_MIN : 0);
  if (c.compact && blockIdx.x == 0 && threadIdx.x == 0 && map_overflows(c, box_window(c, load_box(box))))
    box[6] = 1;
  const BoxIter it = box_iter(c, box);
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  if (next_grid) {
    Box d = load_box(box);
    if (!box_empty(d)) {
      const int mn[3] = {0, 0, c.z_off}, mx[3] = {c.num_grids - 1, c.num_grids - 1, c.z_off + c.nz - 1};
#pragma unroll
      for (int a = 0; a < 3; ++a) { d.lo[a] = max(d.lo[a] - 1, mn[a]); d.hi[a] = min(d.hi[a] + 1, mx[a]); }
    }
    const GridMap wd = box_window(c, d);
    const GridMap md = c.compact ? wd : GridMap{0, 0, c.z_off, c.num_grids, c.num_grids, c.nz};
    long long cnt = map_nodes(wd);
    if (c.compact && cnt > (long long)c.cap_nodes) cnt = c.cap_nodes;   // the whole slot
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cnt;
         i += (long long)gridDim.x * blockDim.x) {
      if (c.compact) {
        next_grid[i] = z;
      } else {
        const int row = (int)(i / wd.nx);
        const int ix = wd.x0 + (int)(i - (long long)row * wd.nx);
        const int pl = row / wd.ny;
        next_grid[node_index(md, ix, wd.y0 + (row - pl * wd.ny), wd.z0 + pl)] = z;
      }
    }
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < it.total;
       i += (long long)gridDim.x * blockDim.x) {
    int ix, iy, iz;
    const size_t ni = box_node(it, i, ix, iy, iz);
    if (!plane_selected(c, iz)) continue;
    const float4 q = grid[ni];
    float v[3], pass[3], inv_m, vp[3];
    node_velocity(c, q, ix, iy, iz, v, pass, inv_m, vp);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      // canonical +0: a computed -0.0f (e.g. gravity passed as -0.0 on an empty node) must not
      // be mistaken for the sentinel below
      v[a] = v[a] + 0.0f;
      if (q.w > 0.0f && pass[a] == 0.0f) v[a] = __uint_as_float(NEG_ZERO);
    }
    grid[ni] = make_float4(v[0], v[1], v[2], q.w);
    if (adj) adj[ni] = z;
  }
}

// g2p from a (v, m) grid (mpm.py:432-498); optionally folds the new positions into the
// next substep's box.
// 100 k particles are 21 warps per SM: (256, 3) keeps 24 warps resident so the kernel is ONE wave.
// Row I/O goes through shared memory (rows_load / rows_store in mpm_common.cuh).
constexpr int G2P_BLOCK = 256;
template <bool WITH_BOX>
__global__ void __launch_bounds__(G2P_BLOCK, 3)
g2p_from_kernel(neuma_mpm_const c, int n, const float* __restrict__ x, const float* __restrict__ F,
                const float* __restrict__ clip, const int* __restrict__ enabled,
                const int32_t* __restrict__ box, const float4* __restrict__ grid,
                float* __restrict__ x2, float* __restrict__ v2, float* __restrict__ C2,
                float* __restrict__ F2, int32_t* __restrict__ box_next) {
  pdl_prologue();
  __shared__ int sbox[6];
  constexpr int B = G2P_BLOCK;
  __shared__ float sm[B * 36];                         // in: x 3, F 9; out: x' 3, v' 3, C' 9, F' 9
  float* const sx = sm;
  float* const sF = sx + 3 * B;
  float* const ox = sF + 9 * B;
  float* const ov = ox + 3 * B;
  float* const oC = ov + 3 * B;
  float* const oF = oC + 9 * B;
  const int t = threadIdx.x, p0 = blockIdx.x * B, p = p0 + t, cnt = min(B, n - p0);
  const GridMap gm = grid_map(c, box);
  rows_load<3, B>(sx, x, p0, cnt);
  rows_load<9, B>(sF, F, p0, cnt);
  const int en = p < n ? enabled[p] : 0;
  __syncthreads();
  const bool active = p < n && en != 0 && particle_selected(c, sx[3 * t + 2]);
  float nx3[3] = {0.f, 0.f, 0.f};
  if (active) g2p_particle<false>(c, gm, t, sx, sF, clip + p0, grid, ox, ov, oC, oF, nx3);
  // rows of particles that do not take part keep what they hold: the block stores cooperatively only
  // when every particle wrote its row
  if (__syncthreads_and(active || p >= n)) {
    rows_store<3, B>(x2, ox, p0, cnt);
    rows_store<3, B>(v2, ov, p0, cnt);
    rows_store<9, B>(C2, oC, p0, cnt);
    rows_store<9, B>(F2, oF, p0, cnt);
  } else if (active) {
    row_store<3>(x2, ox, p, t);
    row_store<3>(v2, ov, p, t);
    row_store<9>(C2, oC, p, t);
    row_store<9>(F2, oF, p, t);
  }
  if (WITH_BOX) box_accumulate(c, active, nx3[0], nx3[1], nx3[2], box_next, sbox);
}

// backward pass 1: scatter adj(G2P)'s dL/dv_node into adj (xyz; w unused)
__global__ void __launch_bounds__(SCATTER_BLOCK)
bwd_scatter_adj_kernel(neuma_mpm_const c, int n, const float* __restrict__ x,
                       const float* __restrict__ F, const float* __restrict__ clip,
                       const int* __restrict__ enabled, const float* __restrict__ v2,
                       const float* __restrict__ gx2, const float* __restrict__ gv2,
                       const float* __restrict__ gC2, const float* __restrict__ gF2,
                       const int32_t* __restrict__ box, float4* __restrict__ gadj) {
  pdl_prologue();
  constexpr int B = SCATTER_BLOCK;
  __shared__ float sm[B * 39];                         // x 3, F 9, v' 3, gx' 3, gv' 3, gC' 9, gF' 9
  float* q = sm;
  auto take = [&](int k) { float* r = q; q += k * B; return r; };
  float* const sx = take(3); float* const sF = take(9); float* const sv2 = take(3); float* const sgx2 = take(3);
  float* const sgv2 = take(3); float* const sgC2 = take(9); float* const sgF2 = take(9);
  const int t = threadIdx.x, p0 = blockIdx.x * B, p = p0 + t, cnt = min(B, n - p0);
  const GridMap gm = grid_map(c, box);
  rows_load<3, B>(sx, x, p0, cnt);     rows_load<9, B>(sF, F, p0, cnt);       rows_load<3, B>(sv2, v2, p0, cnt);
  rows_load<3, B>(sgx2, gx2, p0, cnt); rows_load<3, B>(sgv2, gv2, p0, cnt);   rows_load<9, B>(sgC2, gC2, p0, cnt);
  rows_load<9, B>(sgF2, gF2, p0, cnt);
  const int en = p < n ? enabled[p] : 0;
  __syncthreads();
  if (p >= n || en == 0 || !particle_selected(c, sx[3 * t + 2])) return;
  check_interior_class(c, sx[3 * t + 2], box);
  const G2PAdj g = g2p_adjoint_inputs(c, t, sx, sF, clip + p0, sv2, sgx2, sgv2, sgC2, sgF2 + 9 * t);
  g2p_adjoint_scatter(c, gm, t, sx, g, gadj);
}

// backward pass 2, per node of the box: adj(grid_op).  (v, m) and dL/dv -> (dL/dmv, dL/dm),
// in place in adj.   v = mv / (m + eps) + g dt  =>  dL/dmv = pass * gv / (m + eps),
// dL/dm = -(mv / (m + eps)) . dL/dmv
__global__ void __launch_bounds__(256)
bwd_node_kernel(neuma_mpm_const c, const int32_t* __restrict__ box,
                const float4* __restrict__ grid, float4* __restrict__ adj) {
  pdl_prologue();
  const BoxIter it = box_iter(c, box);
  const float gdt[3] = {c.gravity[0] * c.dt, c.gravity[1] * c.dt, c.gravity[2] * c.dt};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < it.total;
       i += (long long)gridDim.x * blockDim.x) {
    int ix, iy, iz;
    const size_t ni = box_node(it, i, ix, iy, iz);
    if (!plane_selected(c, iz)) continue;
    const float4 q = grid[ni];
    const float4 ga = adj[ni];
    const bool has = q.w > 0.0f;
    const float inv_m = has ? 1.0f / (q.w + c.eps) : 0.0f;
    const float v[3] = {q.x, q.y, q.z}, g[3] = {ga.x, ga.y, ga.z};
    float gmv[3], gm = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const bool pass = has && __float_as_uint(v[a]) != NEG_ZERO;
      gmv[a] = pass ? g[a] * inv_m : 0.0f;
      gm -= pass ? (v[a] - gdt[a]) * gmv[a] : 0.0f;
    }
    adj[ni] = make_float4(gmv[0], gmv[1], gmv[2], gm);
  }
}

// backward pass 3, per particle: gather v_node and (dL/dmv, dL/dm) of its 27 nodes and
// accumulate adj(G2P) (the terms that need v_node) and adj(P2G).
constexpr int GATHER_BLOCK = 128;
constexpr int GATHER_SMEM_FLOATS = 69 + 33;           // staged inputs + outputs per particle
constexpr int GATHER_SMEM = GATHER_SMEM_FLOATS * GATHER_BLOCK * 4;
template <bool HAVE_C2>
__global__ void __launch_bounds__(GATHER_BLOCK, 3)
bwd_gather_v_kernel(neuma_mpm_const c, int n, const float* __restrict__ x,
                    const float* __restrict__ v, const float* __restrict__ C,
                    const float* __restrict__ F, const float* __restrict__ stress,
                    const float* __restrict__ vol, const float* __restrict__ rho,
                    const float* __restrict__ clip, const int* __restrict__ enabled,
                    const float* __restrict__ v2, const float* __restrict__ C2,
                    const float* __restrict__ gx2, const float* __restrict__ gv2,
                    const float* __restrict__ gC2, const float* __restrict__ gF2,
                    const int32_t* __restrict__ box, const float4* __restrict__ grid,
                    const float4* __restrict__ gadj, float* __restrict__ gx, float* __restrict__ gv,
                    float* __restrict__ gC, float* __restrict__ gF, float* __restrict__ gstress) {
  pdl_prologue();
  constexpr int B = GATHER_BLOCK;
  extern __shared__ float gsm[];                       // GATHER_SMEM_FLOATS per particle, see below
  float* q = gsm;
  auto take = [&](int k) { float* r = q; q += k * B; return r; };
  float* const sx = take(3); float* const sv = take(3); float* const sC = take(9); float* const sF = take(9);
  float* const ss = take(9); float* const sv2 = take(3); float* const sC2 = take(9); float* const sgx2 = take(3);
  float* const sgv2 = take(3); float* const sgC2 = take(9); float* const sgF2 = take(9);
  float* const ogx = take(3); float* const ogv = take(3); float* const ogC = take(9); float* const ogF = take(9);
  float* const ogs = take(9);
  const int t = threadIdx.x, p0 = blockIdx.x * B, p = p0 + t, cnt = min(B, n - p0);
  const GridMap gm = grid_map(c, box);
  rows_load<3, B>(sx, x, p0, cnt);   rows_load<3, B>(sv, v, p0, cnt);     rows_load<9, B>(sC, C, p0, cnt);
  rows_load<9, B>(sF, F, p0, cnt);   rows_load<9, B>(ss, stress, p0, cnt); rows_load<3, B>(sv2, v2, p0, cnt);
  if (HAVE_C2) rows_load<9, B>(sC2, C2, p0, cnt);
  rows_load<3, B>(sgx2, gx2, p0, cnt); rows_load<3, B>(sgv2, gv2, p0, cnt); rows_load<9, B>(sgC2, gC2, p0, cnt);
  rows_load<9, B>(sgF2, gF2, p0, cnt);
  const int en = p < n ? enabled[p] : 0;
  __syncthreads();
  // rows of the other selection class belong to the other launch
  const bool mine = p < n && particle_selected(c, sx[3 * t + 2]);
  if (mine) {
    // any of the five outputs may be NULL (not wanted by the caller)
    if (en == 0)
      zero_particle_grads(t, gx ? ogx : nullptr, gv ? ogv : nullptr, gC ? ogC : nullptr, gF ? ogF : nullptr,
                          gstress ? ogs : nullptr);
    else
      bwd_gather_particle<false, HAVE_C2>(c, gm, t, sx, sv, sC, sF, ss, vol + p0, rho + p0, clip + p0, sv2,
                                          HAVE_C2 ? sC2 : nullptr, sgx2, sgv2, sgC2, sgF2, grid, gadj,
                                          gx ? ogx : nullptr, gv ? ogv : nullptr, gC ? ogC : nullptr,
                                          gF ? ogF : nullptr, gstress ? ogs : nullptr);
  }
  if (__syncthreads_and(mine || p >= n)) {
    if (gx) rows_store<3, B>(gx, ogx, p0, cnt);
    if (gv) rows_store<3, B>(gv, ogv, p0, cnt);
    if (gC) rows_store<9, B>(gC, ogC, p0, cnt);
    if (gF) rows_store<9, B>(gF, ogF, p0, cnt);
    if (gstress) rows_store<9, B>(gstress, ogs, p0, cnt);
  } else if (mine) {
    if (gx) row_store<3>(gx, ogx, p, t);
    if (gv) row_store<3>(gv, ogv, p, t);
    if (gC) row_store<9>(gC, ogC, p, t);
    if (gF) row_store<9>(gF, ogF, p, t);
    if (gstress) row_store<9>(gstress, ogs, p, t);
  }
}

// ------------------------------------------------------------------ cell order
__device__ __forceinline__ uint32_t spread3(uint32_t v) {  // 10 bits -> every third bit
  v &= 0x3FFu;
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
__global__ void __launch_bounds__(256)
cell_key_kernel(neuma_mpm_const c, int n, const float* __restrict__ x, uint32_t* __restrict__ keys,
                int32_t* __restrict__ ids) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  uint32_t b[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int t = base_node(x[3 * p + a], c.inv_dx);
    b[a] = (uint32_t)min(max(t, 0), 1023);
  }
  // z is the slowest grid axis: give it the most significant interleaved bit
  keys[p] = spread3(b[0]) | (spread3(b[1]) << 1) | (spread3(b[2]) << 2);
  ids[p] = p;
}

__global__ void __launch_bounds__(256)
gather_rows_kernel(int n, int words, const uint32_t* __restrict__ src, const int32_t* __restrict__ idx,
                   uint32_t* __restrict__ dst) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * words) return;
  const int r = (int)(t / words), w = (int)(t - (long long)r * words);
  dst[t] = src[(size_t)idx[r] * words + w];
}
__global__ void __launch_bounds__(256)
invert_perm_kernel(int n, const int32_t* __restrict__ perm, int32_t* __restrict__ inv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) inv[perm[i]] = i;
}

// dst += src over n float4 nodes: the receiving side of the slab-halo exchange
__global__ void __launch_bounds__(256)
planes_add_kernel(float4* __restrict__ dst, const float4* __restrict__ src, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float4 a = dst[i];
    const float4 b = src[i];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    dst[i] = a;
  }
}

// ---- halo exchange over NVLink peer memory, two kernels per neighbour --------------
// Flags live in symmetric memory and carry monotonically increasing exchange numbers, so nothing
// is ever reset and the kernels -- pointers only, no per-launch scalars -- replay from a CUDA graph.
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// push my shared planes into the neighbour's mailbox, then publish exchange number `seq` in ITS
// data flag.  Waits first until the neighbour has acknowledged exchange seq-1 (mailbox free).
__global__ void __launch_bounds__(256)
halo_push_kernel(const float4* __restrict__ src, float4* __restrict__ peer_mailbox, size_t n,
                 int* __restrict__ peer_data_flag, const int* __restrict__ my_ack_flag,
                 int* __restrict__ my_seq, int* __restrict__ done) {
  __shared__ int s_seq;
  if (threadIdx.x == 0) {
    const int seq = *reinterpret_cast<volatile int*>(my_seq) + 1;
    while (ld_acquire_sys(my_ack_flag) < seq - 1) __nanosleep(64);
    s_seq = seq;
  }
  __syncthreads();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    peer_mailbox[i] = src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    // the block that finishes last has seen every other block's fence: it publishes
    if (atomicAdd(done, 1) == (int)gridDim.x - 1) {
      *done = 0;
      *my_seq = s_seq;
      __threadfence_system();
      st_release_sys(peer_data_flag, s_seq);
    }
  }
}
// wait for the neighbour's push number seq, add the mailbox into my planes, acknowledge
__global__ void __launch_bounds__(256)
halo_add_ack_kernel(float4* __restrict__ dst, const float4* mailbox, size_t n,
                    const int* __restrict__ my_data_flag, int* __restrict__ peer_ack_flag,
                    int* __restrict__ my_seq, int* __restrict__ done) {
  __shared__ int s_seq;
  if (threadIdx.x == 0) {
    const int seq = *reinterpret_cast<volatile int*>(my_seq) + 1;
    while (ld_acquire_sys(my_data_flag) < seq) __nanosleep(64);
    s_seq = seq;
  }
  __syncthreads();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float4 a = dst[i];
    const float4 b = __ldcv(mailbox + i);  // written by the peer GPU: never from a stale L1 line
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    dst[i] = a;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(done, 1) == (int)gridDim.x - 1) {
      *done = 0;
      *my_seq = s_seq;
      __threadfence_system();
      st_release_sys(peer_ack_flag, s_seq);   // the mailbox may be overwritten
    }
  }
}

// ---- ONE kernel per summing halo exchange, both neighbours ----------------------
// Per side: push my shared planes -- only the window (x, y and z range) my particles occupy
// (`box`: what box_clear zeroed and P2G wrote) -- into the neighbour's mailbox over NVLink,
// publish the exchange number; wait for the neighbour's push, add the part of ITS window that
// overlaps mine into my planes, acknowledge.  Mailboxes are double-buffered
// (slot = exchange number & 1), so the push waits only for the acknowledgement of exchange
// seq - 2.  Work is split by (plane, y) ROW over the warps, identically in both phases: a warp
// only ever adds into rows it pushed itself, so no block has to wait for another block of this
// grid (no grid barrier; the grid may be partially resident next to a concurrent P2G).
// Counters are monotonic: cnt[0] = exchanges done, cnt[1] / cnt[2] = blocks that finished the
// push / add phase since the beginning of time.
__global__ void __launch_bounds__(256)
halo_exchange_kernel(neuma_halo_side lo, neuma_halo_side hi, neuma_mpm_const c,
                     const int32_t* __restrict__ box, int32_t* __restrict__ cnt) {
  __shared__ int s_seq;
  __shared__ int s_hdr[2][6];
  const neuma_halo_side side[2] = {lo, hi};
  const int G = c.num_grids;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * (blockDim.x >> 5) + warp, nw = gridDim.x * (blockDim.x >> 5);
  if (threadIdx.x == 0) {
    const int seq = *reinterpret_cast<volatile int*>(cnt) + 1;
#pragma unroll
    for (int s = 0; s < 2; ++s)
      if (side[s].planes)
        while (ld_acquire_sys(side[s].my_ack_flag) < seq - 2) __nanosleep(32);
    s_seq = seq;
  }
  __syncthreads();
  const int seq = s_seq, slot = seq & 1;
  // my window: dense-grid rows [y0, y0 + ny), nodes [x0, x0 + nx) (x0, nx multiples of 8), planes
  // [z0, z0 + nz) -- the region box_clear zeroed and P2G wrote; everything else in my grid slot is
  // stale and must neither be sent nor read
  GridMap w = box_window(c, load_box(box));
  if (c.compact) w.nx = w.ny = w.nz = 0;   // the slab path keeps x, y dense (fixed halo addresses)
  const size_t plane = (size_t)G * G;
  // per side: my planes inside the shared region [zs, zs + npl), as offsets into it
  int mp0[2], mpn[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int lo_pl = max(w.z0, side[s].z0), hi_pl = min(w.z0 + w.nz, side[s].z0 + side[s].npl);
    mp0[s] = lo_pl - side[s].z0;
    mpn[s] = (side[s].planes && hi_pl > lo_pl && w.nx > 0 && w.ny > 0) ? hi_pl - lo_pl : 0;
  }
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    if (!side[s].planes) continue;
    const float4* __restrict__ src = reinterpret_cast<const float4*>(side[s].planes);
    float4* __restrict__ dst = reinterpret_cast<float4*>(side[s].peer_mail) + (size_t)slot * side[s].npl * plane;
    for (int row = gw; row < mpn[s] * w.ny; row += nw) {
      const int pl = mp0[s] + row / w.ny, y = w.y0 + row % w.ny;
      const size_t o = (size_t)pl * plane + (size_t)y * G + w.x0;
      for (int xx = lane; xx < w.nx; xx += 32) dst[o + xx] = src[o + xx];
    }
    if (blockIdx.x == 0 && threadIdx.x < 6) {
      const int h[6] = {w.x0, w.nx, w.y0, w.ny, mp0[s], mpn[s]};
      side[s].peer_hdr[slot * 8 + threadIdx.x] = h[threadIdx.x];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    // the block that finishes the push phase last has seen every other block's fence: it publishes
    if (atomicAdd(&cnt[1], 1) == seq * (int)gridDim.x - 1) {
      __threadfence_system();
#pragma unroll
      for (int s = 0; s < 2; ++s)
        if (side[s].planes) st_release_sys(side[s].peer_data_flag, seq);
    }
#pragma unroll
    for (int s = 0; s < 2; ++s)
      if (side[s].planes) {
        while (ld_acquire_sys(side[s].my_data_flag) < seq) __nanosleep(32);
        for (int k = 0; k < 6; ++k)
          s_hdr[s][k] = *reinterpret_cast<const volatile int*>(side[s].my_hdr + slot * 8 + k);
      }
  }
  __syncthreads();
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    if (!side[s].planes) continue;
    const int x0 = s_hdr[s][0], nx = s_hdr[s][1], y0 = s_hdr[s][2], ny = s_hdr[s][3];
    const int sp0 = s_hdr[s][4], spn = s_hdr[s][5];          // the sender's planes inside the shared region
    if (nx <= 0 || ny <= 0 || spn <= 0) continue;
    float4* __restrict__ dst = reinterpret_cast<float4*>(side[s].planes);
    const float4* src = reinterpret_cast<const float4*>(side[s].my_mail) + (size_t)slot * side[s].npl * plane;
    // (a) rows of MY window, dealt exactly as in the push phase: the warp that pushed a row is
    //     the one that adds into it
    for (int row = gw; row < mpn[s] * w.ny; row += nw) {
      const int pl = mp0[s] + row / w.ny, y = w.y0 + row % w.ny;
      if (y < y0 || y >= y0 + ny || pl < sp0 || pl >= sp0 + spn) continue;
      // nodes outside MY window are never read by my kernels (they lie outside my node box)
      const int xa = max(x0, w.x0), xb = min(x0 + nx, w.x0 + w.nx);
      const size_t o = (size_t)pl * plane + (size_t)y * G + xa;
      for (int xx = lane; xx < xb - xa; xx += 32) {
        float4 a = dst[o + xx];
        const float4 b = __ldcv(src + o + xx);  // written by the peer GPU: never from a stale L1 line
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        dst[o + xx] = a;
      }
    }
    // (b) rows of the SENDER's window outside mine: nobody pushed them and my kernels never read them
    //     (they lie outside my node box) -- nothing to do.
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&cnt[2], 1) == seq * (int)gridDim.x - 1) {
      cnt[0] = seq;
      __threadfence_system();
#pragma unroll
      for (int s = 0; s < 2; ++s)
        if (side[s].planes) st_release_sys(side[s].peer_ack_flag, seq);   // slot may be overwritten at seq + 2
    }
  }
}

constexpr int NEUMA_HALO_BLOCKS = 64;

inline int node_pass_blocks(neuma_ctx* ctx) { return ctx->num_sms * 8; }

// opt-in dynamic shared memory of the gather kernel (per device, tracked in the ctx)
int set_rollout_attrs(neuma_ctx* ctx) {
  if (ctx->rollout_attrs_set) return 0;
  int cur = -1;
  cudaGetDevice(&cur);
  if (cur != ctx->device) NEUMA_CUDA(ctx, cudaSetDevice(ctx->device));
  const auto attr = cudaFuncAttributeMaxDynamicSharedMemorySize;
  NEUMA_CUDA(ctx, cudaFuncSetAttribute(bwd_gather_v_kernel<true>, attr, GATHER_SMEM));
  NEUMA_CUDA(ctx, cudaFuncSetAttribute(bwd_gather_v_kernel<false>, attr, GATHER_SMEM));
  ctx->rollout_attrs_set = 1;
  if (cur >= 0 && cur != ctx->device) cudaSetDevice(cur);
  return 0;
}

int check_slab(neuma_ctx* ctx, const neuma_mpm_const* c) {
  if (!ctx || !c) return 2;
  if (c->num_grids <= 0 || c->nz <= 0 || c->z_off < 0 || c->z_off + c->nz > c->num_grids + 4)
    return neuma_fail_msg(ctx, "neuma_mpm: bad grid extents");
  if (c->bc != 0 && c->bc != 1) return neuma_fail_msg(ctx, "neuma_mpm: bc must be 0 or 1");
  if (c->compact && c->cap_nodes <= 0) return neuma_fail_msg(ctx, "neuma_mpm: compact grid without cap_nodes");
  return 0;
}

}  // namespace

extern "C" {

int neuma_mpm_box_from_x(neuma_ctx* ctx, const neuma_mpm_const* c, int n, const float* x,
                         const int32_t* enabled, int32_t* box, void* stream) {
  if (int r = check_slab(ctx, c)) return r;
  cudaStream_t st = (cudaStream_t)stream;
  box_init_kernel<<<1, 32, 0, st>>>(box);
  NEUMA_LAUNCH_CHECK(ctx, "box_init_kernel");
  if (n <= 0) return 0;
  box_reduce_kernel<<<(n + 255) / 256, 256, 0, st>>>(*c, n, x, enabled, box);
  NEUMA_LAUNCH_CHECK(ctx, "box_reduce_kernel");
  return 0;
}

int neuma_mpm_box_clear(neuma_ctx* ctx, const neuma_mpm_const* c, int32_t* box, float* grid0,
                        float* grid1, int32_t* box_reset, void* stream) {
  if (int r = check_slab(ctx, c)) return r;
  NEUMA_CUDA(ctx, neuma_launch(box_clear_kernel, dim3(node_pass_blocks(ctx)), dim3(256), 0, (cudaStream_t)stream,
                               *c, box, (float4*)grid0, (float4*)grid1, box_reset));
  NEUMA_LAUNCH_CHECK(ctx, "box_clear_kernel");
  return 0;
}

int neuma_mpm_p2g_to(neuma_ctx* ctx, const neuma_mpm_const* c, int n, const float* x, const float* v,
                     const float* C, const float* stress, const float* vol, const float* rho,
                     const int32_t* enabled, const int32_t* box, float* grid, void* stream) {
  if (int r = check_slab(ctx, c)) return r;
  if (n <= 0) return 0;
  if ((c->compact || c->sel_mode == 3) && !box)
    return neuma_fail_msg(ctx, "neuma_mpm_p2g_to: a compact grid / sel_mode 3 needs the box");
  static int tiled = -1;
  if (tiled < 0) {
    const char* e = getenv("NEUMA_P2G");
    tiled = (e && !strcmp(e, "tiled")) ? 1 : 0;
  }
  if (tiled) {   // the measured alternative (see p2g_tiled_kernel)
    NEUMA_CUDA(ctx, neuma_launch(p2g_tiled_kernel, dim3((n + 255) / 256), dim3(256), 0, (cudaStream_t)stream, *c, n,
                                 x, v, C, stress, vol, rho, enabled, box, (float4*)grid));
    NEUMA_LAUNCH_CHECK(ctx, "p2g_tiled_kernel");
    return 0;
  }
  const int pb = SCATTER_BLOCK;
  NEUMA_CUDA(ctx, neuma_launch(p2g_to_kernel, dim3((n + pb - 1) / pb), dim3(pb), 0, (cudaStream_t)stream, *c, n, x,
                               v, C, stress, vol, rho, enabled, box, (float4*)grid));
  NEUMA_LAUNCH_CHECK(ctx, "p2g_to_kernel");
  return 0;
}

int neuma_mpm_grid_op_box(neuma_ctx* ctx, const neuma_mpm_const* c, int32_t* box, float* grid,
                          float* adj_to_clear, void* stream) {
  return neuma_mpm_grid_op_next(ctx, c, box, grid, adj_to_clear, nullptr, nullptr, stream);
}

int neuma_mpm_grid_op_next(neuma_ctx* ctx, const neuma_mpm_const* c, int32_t* box, float* grid, float* adj_to_clear,
                           float* next_grid, int32_t* box_reset, void* stream) {
  if (int r = check_slab(ctx, c)) return r;
  if (!box || !grid) return neuma_fail_msg(ctx, "neuma_mpm_grid_op: box and grid are required");
  if (next_grid && c->sel_mode != 0) return neuma_fail_msg(ctx, "neuma_mpm_grid_op_next: sel_mode must be 0");
  NEUMA_CUDA(ctx, neuma_launch(grid_op_box_kernel, dim3(node_pass_blocks(ctx)), dim3(256), 0, (cudaStream_t)stream,
                               *c, box, (float4*)grid, (float4*)adj_to_clear, (float4*)next_grid, box_reset));
  NEUMA_LAUNCH_CHECK(ctx, "grid_op_box_kernel");
  return 0;
}

int neuma_mpm_g2p_from(neuma_ctx* ctx, const neuma_mpm_const* c, int n, const float* x,
                       const float* F, const float* clip_bound, const int32_t* enabled,
                       const int32_t* box, const float* grid_v, float* x2, float* v2, float* C2,
                       float* F2, int32_t* box_next, void* stream) {
  if (int r = check_slab(ctx, c)) return r;
  if (n <= 0) return 0;
  if (c->compact && !box) return neuma_fail_msg(ctx, "neuma_mpm_g2p_from: a compact grid needs its box");
  cudaStream_t st = (cudaStream_t)stream;
  const int pb = G2P_BLOCK;
  NEUMA_CUDA(ctx, neuma_launch(box_next ? g2p_from_kernel<true> : g2p_from_kernel<false>, dim3((n + pb - 1) / pb),
                               dim3(pb), 0, st, *c, n, x, F, clip_bound, enabled, box, (const float4*)grid_v, x2, v2,
                               C2, F2, box_next));
  NEUMA_LAUNCH_CHECK(ctx, "g2p_from_kernel");
  return 0;
}

int neuma_mpm_rollout_fwd(neuma_ctx* ctx, const neuma_mpm_const* c, int n, const float* x,
                          const float* v, const float* C, const float* F, const float* stress,
                          const float* vol, const float* rho, const float* clip_bound,
                          const int32_t* enabled, float* x2, float* v2, float* C2, float* F2,
                          float* grid, float* adj, int32_t* box, int32_t* box_next,
                          void* stream) {
  if (!grid || !box) return neuma_fail_msg(ctx, "neuma_mpm_rollout_fwd: grid and box are required");
  if (int r = neuma_mpm_box_clear(ctx, c, box, grid, nullptr, box_next, stream)) return r;
  if (int r = neuma_mpm_p2g_to(ctx, c, n, x, v, C, stress, vol, rho, enabled, box, grid, stream)) return r;
  if (int r = neuma_mpm_grid_op_box(ctx, c, box, grid, adj, stream)) return r;
  return neuma_mpm_g2p_from(ctx, c, n, x, F, clip_bound, enabled, box, grid, x2, v2, C2, F2, box_next,
                            stream);
}

int neuma_mpm_bwd_scatter(neuma_ctx* ctx, const neuma_mpm_const* c, int n, const float* x,
                          const float* F, const float* clip_bound, const int32_t* enabled,
                          const float* v2, const float* gx2, const float* gv2, const float* gC2,
                          const float* gF2, const int32_t* box, float* adj, void* stream) {
  if (int r = check_slab(ctx, c)) return r;
  if (n <= 0) return 0;
  if ((c->compact || c->sel_mode == 3) && !box)
    return neuma_fail_msg(ctx, "neuma_mpm_bwd_scatter: a compact grid / sel_mode 3 needs the box");
  if (!v2 || !gx2 || !gv2 || !gC2 || !gF2 || !adj)
    return neuma_fail_msg(ctx, "neuma_mpm_bwd_scatter: NULL argument");
  const int pb = SCATTER_BLOCK;
  NEUMA_CUDA(ctx, neuma_launch(bwd_scatter_adj_kernel, dim3((n + pb - 1) / pb), dim3(pb), 0, (cudaStream_t)stream,
                               *c, n, x, F, clip_bound, enabled, v2, gx2, gv2, gC2, gF2, box, (float4*)adj));
  NEUMA_LAUNCH_CHECK(ctx, "bwd_scatter_adj_kernel");
  return 0;
}

int neuma_mpm_bwd_node(neuma_ctx* ctx, const neuma_mpm_const* c, const float* grid_v, float* adj, const int32_t* box,
                       void* stream) {
  if (int r = check_slab(ctx, c)) return r;
  if (!grid_v || !adj || !box) return neuma_fail_msg(ctx, "neuma_mpm_bwd_node: NULL argument");
  NEUMA_CUDA(ctx, neuma_launch(bwd_node_kernel, dim3(node_pass_blocks(ctx)), dim3(256), 0, (cudaStream_t)stream, *c,
                               box, (const float4*)grid_v, (float4*)adj));
  NEUMA_LAUNCH_CHECK(ctx, "bwd_node_kernel");
  return 0;
}

int neuma_mpm_bwd_node_gather(neuma_ctx* ctx, const neuma_mpm_const* c, int n, const float* x,
                              const float* v, const float* C, const float* F, const float* stress,
                              const float* vol, const float* rho, const float* clip_bound,
                              const int32_t* enabled, const float* v2, const float* C2,
                              const float* gx2, const float* gv2, const float* gC2, const float* gF2,
                              const float* grid_v, float* adj, const int32_t* box, float* gx,
                              float* gv, float* gC, float* gF, float* gstress, void* stream) {
  if (int r = check_slab(ctx, c)) return r;
  if (n <= 0) return 0;
  if (!v2 || !gx2 || !gv2 || !gC2 || !gF2 || !grid_v || !adj || !box)
    return neuma_fail_msg(ctx, "neuma_mpm_bwd_node_gather: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  NEUMA_CUDA(ctx, neuma_launch(bwd_node_kernel, dim3(node_pass_blocks(ctx)), dim3(256), 0, st, *c, box,
                               (const float4*)grid_v, (float4*)adj));
  NEUMA_LAUNCH_CHECK(ctx, "bwd_node_kernel");
  const int pb = GATHER_BLOCK;
  if (int r = set_rollout_attrs(ctx)) return r;
  NEUMA_CUDA(ctx, neuma_launch(C2 ? bwd_gather_v_kernel<true> : bwd_gather_v_kernel<false>, dim3((n + pb - 1) / pb),
                               dim3(pb), GATHER_SMEM, st, *c, n, x, v, C, F, stress, vol, rho, clip_bound, enabled, v2, C2, gx2,
                               gv2, gC2, gF2, box, (const float4*)grid_v, (const float4*)adj, gx, gv, gC, gF, gstress));
  NEUMA_LAUNCH_CHECK(ctx, "bwd_gather_v_kernel");
  return 0;
}

int neuma_mpm_rollout_bwd(neuma_ctx* ctx, const neuma_mpm_const* c, int n, const float* x,
                          const float* v, const float* C, const float* F, const float* stress,
                          const float* vol, const float* rho, const float* clip_bound,
                          const int32_t* enabled, const float* v2, const float* C2,
                          const float* gx2, const float* gv2, const float* gC2, const float* gF2,
                          const float* grid_v, float* adj, const int32_t* box, float* gx, float* gv,
                          float* gC, float* gF, float* gstress, void* stream) {
  if (int r = neuma_mpm_bwd_scatter(ctx, c, n, x, F, clip_bound, enabled, v2, gx2, gv2, gC2, gF2, box, adj,
                                    stream))
    return r;
  return neuma_mpm_bwd_node_gather(ctx, c, n, x, v, C, F, stress, vol, rho, clip_bound, enabled, v2, C2,
                                   gx2, gv2, gC2, gF2, grid_v, adj, box, gx, gv, gC, gF, gstress,
                                   stream);
}

int neuma_mpm_planes_add(neuma_ctx* ctx, float* dst, const float* src, size_t nodes, void* stream) {
  if (!ctx) return 2;
  if (nodes == 0) return 0;
  size_t b = (nodes + 255) / 256;
  if (b > (size_t)ctx->num_sms * 8) b = (size_t)ctx->num_sms * 8;
  planes_add_kernel<<<(int)b, 256, 0, (cudaStream_t)stream>>>((float4*)dst, (const float4*)src, nodes);
  NEUMA_LAUNCH_CHECK(ctx, "planes_add_kernel");
  return 0;
}

int neuma_halo_push(neuma_ctx* ctx, const float* src, float* peer_mailbox, size_t nodes,
                    int32_t* peer_data_flag, const int32_t* my_ack_flag, int32_t* my_seq, int32_t* done,
                    void* stream) {
  if (!ctx) return 2;
  if (nodes == 0) return 0;
  size_t b = (nodes + 255) / 256;
  if (b > (size_t)ctx->num_sms * 2) b = (size_t)ctx->num_sms * 2;   // all blocks co-resident: they spin
  halo_push_kernel<<<(int)b, 256, 0, (cudaStream_t)stream>>>((const float4*)src, (float4*)peer_mailbox, nodes,
                                                              peer_data_flag, my_ack_flag, my_seq, done);
  NEUMA_LAUNCH_CHECK(ctx, "halo_push_kernel");
  return 0;
}

int neuma_halo_add_ack(neuma_ctx* ctx, float* dst, const float* mailbox, size_t nodes,
                       const int32_t* my_data_flag, int32_t* peer_ack_flag, int32_t* my_seq, int32_t* done,
                       void* stream) {
  if (!ctx) return 2;
  if (nodes == 0) return 0;
  size_t b = (nodes + 255) / 256;
  if (b > (size_t)ctx->num_sms * 2) b = (size_t)ctx->num_sms * 2;
  halo_add_ack_kernel<<<(int)b, 256, 0, (cudaStream_t)stream>>>((float4*)dst, (const float4*)mailbox, nodes,
                                                                 my_data_flag, peer_ack_flag, my_seq, done);
  NEUMA_LAUNCH_CHECK(ctx, "halo_add_ack_kernel");
  return 0;
}

int neuma_halo_exchange(neuma_ctx* ctx, const neuma_mpm_const* c, const neuma_halo_side* lo,
                        const neuma_halo_side* hi, const int32_t* box, int32_t* cnt, void* stream) {
  if (int r = check_slab(ctx, c)) return r;
  if (c->compact) return neuma_fail_msg(ctx, "neuma_halo_exchange: the slab grids are dense in x, y");
  if (!box || !cnt) return neuma_fail_msg(ctx, "neuma_halo_exchange: NULL argument");
  neuma_halo_side none;
  memset(&none, 0, sizeof(none));
  const neuma_halo_side a = lo ? *lo : none, b = hi ? *hi : none;
  if (!a.planes && !b.planes) return 0;
  // a FIXED grid: the monotonic counters count blocks; all blocks co-resident with room to spare
  halo_exchange_kernel<<<NEUMA_HALO_BLOCKS, 256, 0, (cudaStream_t)stream>>>(a, b, *c, box, cnt);
  NEUMA_LAUNCH_CHECK(ctx, "halo_exchange_kernel");
  return 0;
}

int neuma_mpm_cell_order(neuma_ctx* ctx, const neuma_mpm_const* c, int n, const float* x,
                         int32_t* perm, void* stream) {
  if (int r = check_slab(ctx, c)) return r;
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int bits = 1;
  while ((1 << bits) < c->num_grids && bits < 10) ++bits;
  size_t temp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, n, 0, 3 * bits, st);
  const size_t karr = ((size_t)n * 4 + 255) & ~(size_t)255;
  if (int r = neuma_ensure_scratch(ctx, 3 * karr + temp)) return r;
  uint8_t* base = (uint8_t*)ctx->scratch;
  uint32_t* keys = (uint32_t*)base;
  uint32_t* keys2 = (uint32_t*)(base + karr);
  int32_t* ids = (int32_t*)(base + 2 * karr);
  cell_key_kernel<<<(n + 255) / 256, 256, 0, st>>>(*c, n, x, keys, ids);
  NEUMA_LAUNCH_CHECK(ctx, "cell_key_kernel");
  NEUMA_CUDA(ctx, cub::DeviceRadixSort::SortPairs(base + 3 * karr, temp, keys, keys2, ids, perm, n, 0,
                                                  3 * bits, st));
  ctx->launches += 3;
  return 0;
}

int neuma_gather_rows(neuma_ctx* ctx, int n, int words_per_row, const void* src, const int32_t* idx,
                      void* dst, void* stream) {
  if (!ctx) return 2;
  if (n <= 0 || words_per_row <= 0) return 0;
  const long long total = (long long)n * words_per_row;
  gather_rows_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      n, words_per_row, (const uint32_t*)src, idx, (uint32_t*)dst);
  NEUMA_LAUNCH_CHECK(ctx, "gather_rows_kernel");
  return 0;
}

int neuma_invert_perm(neuma_ctx* ctx, int n, const int32_t* perm, int32_t* inv, void* stream) {
  if (!ctx) return 2;
  if (n <= 0) return 0;
  invert_perm_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, perm, inv);
  NEUMA_LAUNCH_CHECK(ctx, "invert_perm_kernel");
  return 0;
}

}  // extern "C"
