/*
 * neuma_b200.h -- C ABI of libneuma_b200.so
 *
 * B200-native (sm_100a) replacement for the one hot path of XJay18/NeuMA: the
 * differentiable MLS-MPM substep, the neural constitutive law, the
 * particle->Gaussian binding, and the Particle-GS tile rasterizer.
 *
 * The reference has NO native interface for this path: its kernels are Warp-DSL
 * Python JIT-compiled at run time (modules/nclaw/sim/mpm.py:321-498,
 * modules/nclaw/warp/svd.py:59-96, modules/d3gs/utils/simulation_utils.py:25-48)
 * and an external CUDA extension (diff_gaussian_rasterization, call site
 * modules/tune/utils.py:385-419).  Each entry point below names the reference
 * symbol it replaces; INTEGRATION.md shows the ctypes stub a maintainer adds on
 * the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - all arrays are fp32 / int32, C-contiguous, in the reference's layouts:
 *       x,v [Np,3]; C,F,stress [Np,3,3] row-major; vol,rho,clip_bound [Np];
 *       enabled int32[Np]; packed covariances [K,6] = xx,xy,xz,yy,yz,zz;
 *   - the caller owns every array; the library owns only the ctx workspace
 *     (grid, sort buffers), grown lazily and never returned;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no
 *     entry point synchronises the device unless documented;
 *   - every entry returns 0 on success, non-zero on failure; the message is
 *     available from neuma_last_error(ctx).  No exceptions cross the ABI;
 *   - a ctx is bound to one device and is not thread-safe (the reference is
 *     single-threaded, experiments/finetune.py:507).
 */
#ifndef NEUMA_B200_H
#define NEUMA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct neuma_ctx neuma_ctx;

/* MPMConstant, modules/nclaw/sim/mpm.py:158-167 (+ bc from :545-550).
 * z_off / nz describe the z-slab of the grid this ctx holds: grid planes
 * [z_off, z_off + nz).  Single GPU: z_off = 0, nz = num_grids.
 * The fields after nz only concern the rollout entry points (zero = the plain dense case):
 *   compact / cap_nodes : the caller's per-substep grid buffers hold only the bounding box of
 *       the occupied nodes (cap_nodes float4 nodes each) instead of nz*G*G nodes;
 *   sel_mode / sel_lo_end / sel_up_start : Z-slab overlap.  Planes below sel_lo_end
 *       and from sel_up_start on are SHARED with a neighbour rank.  sel_mode 1 restricts a
 *       particle kernel to the particles whose stencil touches a shared plane and a node pass
 *       to the shared planes; sel_mode 2 to the complement; 0 = everything; sel_mode 3 = everything,
 *       but the launch covers particles the caller classified as interior: one that touches a
 *       shared plane is counted in box[7] (scatter kernels only). */
typedef struct {
  int32_t num_grids;
  float dt;
  int32_t bound;
  float gravity[3];
  float dx;
  float inv_dx;
  float eps;
  int32_t bc;    /* 0 = noslip (mpm.py:402-429), 1 = freeslip (mpm.py:373-400) */
  int32_t z_off; /* first grid plane held by this ctx */
  int32_t nz;    /* number of grid planes held by this ctx */
  int32_t compact;
  int32_t cap_nodes;
  int32_t sel_mode;
  int32_t sel_lo_end;
  int32_t sel_up_start;
} neuma_mpm_const;

/* One side (lower or upper z-neighbour) of the Z-slab halo exchange.  All pointers
 * are DEVICE pointers; peer_* point into the NEIGHBOUR's memory (NVLink peer mapping).  A
 * mailbox holds 2 slots x npl x G x G float4 nodes, a header 2 slots x 8 int32 (the sender's
 * window x0, nx, y0, ny, first plane, planes); flags are single int32, zero-initialised once.
 * planes == NULL: no neighbour on this side. */
typedef struct {
  float* planes;               /* my shared planes inside the grid being exchanged */
  float* peer_mail;            /* neighbour's mailbox for what I send */
  int32_t* peer_hdr;
  const float* my_mail;        /* my mailbox for what that neighbour sends */
  const int32_t* my_hdr;
  int32_t* peer_data_flag;
  const int32_t* my_data_flag;
  int32_t* peer_ack_flag;
  const int32_t* my_ack_flag;
  int32_t npl;                 /* number of shared planes */
  int32_t z0;                  /* global index of the first shared plane */
} neuma_halo_side;

/* GaussianRasterizationSettings, modules/d3gs/gaussian_renderer/__init__.py:103-116.
 * viewmatrix / projmatrix are the reference's transposed (row-vector) 4x4s
 * (modules/d3gs/scene/cameras.py:54-56), flattened row-major, HOST arrays. */
typedef struct {
  int32_t image_height;
  int32_t image_width;
  float tanfovx;
  float tanfovy;
  float bg[3];
  float scale_modifier;
  float viewmatrix[16];
  float projmatrix[16];
  int32_t sh_degree;
  float campos[3];
  int32_t prefiltered;
  int32_t debug;
} neuma_raster_settings;

/* ---- context ---------------------------------------------------------- */
int neuma_ctx_create(int device, neuma_ctx** out);
int neuma_ctx_destroy(neuma_ctx* ctx);
const char* neuma_last_error(neuma_ctx* ctx);
int neuma_version(void);
/* number of kernels this ctx has launched since creation (bench.py gpu_launches) */
long long neuma_launch_count(neuma_ctx* ctx);

/* ---- MPM substep: modules/nclaw/sim/mpm.py ---------------------------- */
/* grid.clear(), mpm.py:147-150 / :290 */
int neuma_mpm_grid_clear(neuma_ctx* ctx, const neuma_mpm_const* c, void* stream);
/* p2g, mpm.py:321-371 (accumulates into the ctx grid) */
int neuma_mpm_p2g(neuma_ctx* ctx, const neuma_mpm_const* c, int n,
                  const float* x, const float* v, const float* C, const float* stress,
                  const float* vol, const float* rho, const int32_t* enabled, void* stream);
/* grid_op_noslip / grid_op_freeslip, mpm.py:373-429 (dense sweep, writes grid.v) */
int neuma_mpm_grid_op(neuma_ctx* ctx, const neuma_mpm_const* c, void* stream);
/* g2p, mpm.py:432-498.  In-place safe (x2 may alias x, F2 may alias F), as
 * MPMForwardSim requires (modules/nclaw/sim/interface.py:133).
 * lazy_grid_op != 0: evaluate the grid_op arithmetic per gathered node from
 * (mv, m) instead of reading grid.v (no dense sweep needed). */
int neuma_mpm_g2p(neuma_ctx* ctx, const neuma_mpm_const* c, int n,
                  const float* x, const float* F, const float* clip_bound, const int32_t* enabled,
                  float* x2, float* v2, float* C2, float* F2, int lazy_grid_op, void* stream);
/* copies the ctx grid out in the reference's layout (mpm.py:131-142): mv, v
 * [G,G,nz,3] and m [G,G,nz], index [ix][iy][iz - z_off]; any pointer may be NULL. */
int neuma_mpm_grid_export(neuma_ctx* ctx, const neuma_mpm_const* c,
                          float* mv, float* m, float* v, void* stream);
/* adds (dir=+1) / overwrites (dir=0) `nplanes` z-planes of the packed (mv,m)
 * grid starting at plane z (global index) from/to a contiguous buffer of
 * nplanes*G*G*4 floats: the slab-halo pack/unpack. */
int neuma_mpm_halo_pack(neuma_ctx* ctx, const neuma_mpm_const* c, int z, int nplanes,
                        float* buf, void* stream);
int neuma_mpm_halo_unpack_add(neuma_ctx* ctx, const neuma_mpm_const* c, int z, int nplanes,
                              const float* buf, void* stream);

/* MPMModel.forward, mpm.py:279-297: clear + p2g + grid_op + g2p, fused
 * (clear, p2g, g2p with lazy grid_op). */
int neuma_mpm_substep_fwd(neuma_ctx* ctx, const neuma_mpm_const* c, int n,
                          const float* x, const float* v, const float* C, const float* F,
                          const float* stress,
                          const float* vol, const float* rho, const float* clip_bound,
                          const int32_t* enabled,
                          float* x2, float* v2, float* C2, float* F2, void* stream);
/* MPMModel.backward, mpm.py:299-319 + MPMSimFunction.backward,
 * modules/nclaw/sim/interface.py:39-76: recompute p2g + grid_op, then
 * adj(g2p), adj(grid_op), adj(p2g); NaN/Inf in the results are replaced by 0
 * (interface.py:65-74).  Any output gradient pointer may be NULL
 * (the reference returns None for inputs without requires_grad, mpm.py:198-204).
 * v2 is the v' the forward produced (the reference keeps it in state_next): it
 * decides which components of x' = clamp(x + dt v') were clamped. */
int neuma_mpm_substep_bwd(neuma_ctx* ctx, const neuma_mpm_const* c, int n,
                          const float* x, const float* v, const float* C, const float* F,
                          const float* stress,
                          const float* vol, const float* rho, const float* clip_bound,
                          const int32_t* enabled,
                          const float* v2 /* the forward's v output */,
                          const float* gx2, const float* gv2, const float* gC2, const float* gF2,
                          float* gx, float* gv, float* gC, float* gF, float* gstress,
                          void* stream);

/* ---- MPM rollout path (MPMCacheDiffSim, modules/nclaw/sim/interface.py:108-123) --------
 * The same substep with CALLER-OWNED packed grids, one per substep of a differentiable
 * rollout, so the backward reads the forward's grid instead of recomputing P2G
 * (mpm.py:312-315).  `grid` / `adj`: nz*G*G float4 nodes, plane-major (z slowest) -- or, with
 * c->compact, c->cap_nodes nodes holding just the box (x-range widened to multiples of 8), z
 * slowest inside the box; a box that does not fit raises box[6] and the substep is skipped.
 * box[7] counts particles whose stencil left the planes [z_off, z_off + nz) (Z-slab drift).
 * `box`: int32[8] on the device = inclusive node bounding box (lo.xyz, hi.xyz, 2 pad) of
 * the occupied nodes; clears and node passes touch only the box. */
/* box of the nodes the particles at x touch (init + reduce) */
int neuma_mpm_box_from_x(neuma_ctx* ctx, const neuma_mpm_const* c, int n, const float* x,
                         const int32_t* enabled, int32_t* box, void* stream);
/* grid.clear() (mpm.py:147-150) restricted to the box, on up to two grids (either may be
 * NULL); box_reset (may be NULL) is set to the empty box */
int neuma_mpm_box_clear(neuma_ctx* ctx, const neuma_mpm_const* c, int32_t* box, float* grid0,
                        float* grid1, int32_t* box_reset, void* stream);
/* p2g (mpm.py:321-371) accumulating (mv, m) into `grid` */
int neuma_mpm_p2g_to(neuma_ctx* ctx, const neuma_mpm_const* c, int n, const float* x, const float* v,
                     const float* C, const float* stress, const float* vol, const float* rho,
                     const int32_t* enabled, const int32_t* box /* NULL unless c->compact */,
                     float* grid, void* stream);
/* grid_op_* (mpm.py:373-429) over the box, in place: (mv, m) -> (v, m); components zeroed
 * by the wall BC are stored as -0.0f.  adj_to_clear (may be NULL) is zeroed over the box. */
int neuma_mpm_grid_op_box(neuma_ctx* ctx, const neuma_mpm_const* c, int32_t* box, float* grid,
                          float* adj_to_clear, void* stream);
/* the same, and prepares the NEXT substep's slot so that no clear kernel sits between this substep's
 * G2P and the next P2G: next_grid (may be NULL) is zeroed over this box dilated by one node (the
 * next box lies inside unless a particle crossed more than one cell in the substep -- the fused
 * neuma_constitutive_fwd_g2p raises box_next[6] then), box_reset (may be NULL) is set to "empty". */
int neuma_mpm_grid_op_next(neuma_ctx* ctx, const neuma_mpm_const* c, int32_t* box, float* grid, float* adj_to_clear,
                           float* next_grid, int32_t* box_reset, void* stream);
/* g2p (mpm.py:432-498) from a (v, m) grid; box_next (may be NULL; must hold the empty box
 * or a partial reduction) receives the box of the nodes the NEW positions touch */
int neuma_mpm_g2p_from(neuma_ctx* ctx, const neuma_mpm_const* c, int n, const float* x,
                       const float* F, const float* clip_bound, const int32_t* enabled,
                       const int32_t* box /* NULL unless c->compact */,
                       const float* grid_v, float* x2, float* v2, float* C2, float* F2,
                       int32_t* box_next, void* stream);
/* MPMModel.forward (mpm.py:279-297) = box_clear(grid; resets box_next) + p2g_to +
 * grid_op_box (also zeroes adj over the box if non-NULL) + g2p_from */
int neuma_mpm_rollout_fwd(neuma_ctx* ctx, const neuma_mpm_const* c, int n, const float* x,
                          const float* v, const float* C, const float* F, const float* stress,
                          const float* vol, const float* rho, const float* clip_bound,
                          const int32_t* enabled, float* x2, float* v2, float* C2, float* F2,
                          float* grid, float* adj, int32_t* box, int32_t* box_next,
                          void* stream);
/* MPMModel.backward (mpm.py:299-319) without the P2G recompute: grid_v is the (v, m) grid
 * neuma_mpm_rollout_fwd left, adj a zeroed (over the box) scratch grid (overwritten), C2 the
 * forward's C output (may be NULL: recomputed).  Gradient outputs as neuma_mpm_substep_bwd. */
int neuma_mpm_rollout_bwd(neuma_ctx* ctx, const neuma_mpm_const* c, int n, const float* x,
                          const float* v, const float* C, const float* F, const float* stress,
                          const float* vol, const float* rho, const float* clip_bound,
                          const int32_t* enabled, const float* v2, const float* C2,
                          const float* gx2, const float* gv2, const float* gC2, const float* gF2,
                          const float* grid_v, float* adj, const int32_t* box, float* gx, float* gv,
                          float* gC, float* gF, float* gstress, void* stream);
/* the two halves of neuma_mpm_rollout_bwd, for the Z-slab path: scatter adj(G2P) into adj;
 * [sum adj over the slab overlap planes with the neighbours]; node pass + gather */
int neuma_mpm_bwd_scatter(neuma_ctx* ctx, const neuma_mpm_const* c, int n, const float* x,
                          const float* F, const float* clip_bound, const int32_t* enabled,
                          const float* v2, const float* gx2, const float* gv2, const float* gC2,
                          const float* gF2, const int32_t* box /* NULL unless c->compact */,
                          float* adj, void* stream);
int neuma_mpm_bwd_node_gather(neuma_ctx* ctx, const neuma_mpm_const* c, int n, const float* x,
                              const float* v, const float* C, const float* F, const float* stress,
                              const float* vol, const float* rho, const float* clip_bound,
                              const int32_t* enabled, const float* v2, const float* C2,
                              const float* gx2, const float* gv2, const float* gC2, const float* gF2,
                              const float* grid_v, float* adj, const int32_t* box, float* gx,
                              float* gv, float* gC, float* gF, float* gstress, void* stream);
/* the node pass alone (adjoint of grid_op over the box, in place on adj); the gather that follows it is
 * either neuma_mpm_bwd_node_gather's second kernel or neuma_constitutive_bwd_gather */
int neuma_mpm_bwd_node(neuma_ctx* ctx, const neuma_mpm_const* c, const float* grid_v, float* adj, const int32_t* box,
                       void* stream);
/* dst += src over `nodes` packed float4 nodes: the receiving side of the Z-slab halo exchange
 * (a slab's overlap planes are one contiguous block in the plane-major layout) */
int neuma_mpm_planes_add(neuma_ctx* ctx, float* dst, const float* src, size_t nodes, void* stream);
/* The summing halo exchange over NVLink PEER memory, two kernels per neighbour (no NCCL, CUDA-graph
 * capturable): neuma_halo_push copies `nodes` float4 nodes into the neighbour's mailbox (a peer
 * pointer) once the neighbour has acknowledged the previous exchange (my_ack_flag, written by the
 * neighbour), then stores the exchange number into the neighbour's data flag;
 * neuma_halo_add_ack waits for its own data flag, adds its mailbox into `dst` and acknowledges.
 * Flags and my_seq / done counters are int32, zero-initialised once; flags carry monotonically
 * increasing exchange numbers. */
int neuma_halo_push(neuma_ctx* ctx, const float* src, float* peer_mailbox, size_t nodes,
                    int32_t* peer_data_flag, const int32_t* my_ack_flag, int32_t* my_seq, int32_t* done,
                    void* stream);
int neuma_halo_add_ack(neuma_ctx* ctx, float* dst, const float* mailbox, size_t nodes,
                       const int32_t* my_data_flag, int32_t* peer_ack_flag, int32_t* my_seq, int32_t* done,
                       void* stream);
/* The same exchange as ONE kernel for both neighbours (what SlabRollout uses): pushes only the
 * x-y window `box` says my particles occupy, waits for the neighbours' pushes, adds THEIR
 * windows, acknowledges; double-buffered mailboxes.  cnt: int32[4] local counters, zeroed once.
 * Either side may be NULL. */
int neuma_halo_exchange(neuma_ctx* ctx, const neuma_mpm_const* c, const neuma_halo_side* lo,
                        const neuma_halo_side* hi, const int32_t* box, int32_t* cnt, void* stream);
/* permutation that puts the particles in cell (Morton) order of their stencil base:
 * perm[i] = index of the i-th particle in that order.  Not capturable (may grow the ctx
 * scratch). */
int neuma_mpm_cell_order(neuma_ctx* ctx, const neuma_mpm_const* c, int n, const float* x,
                         int32_t* perm, void* stream);
/* dst[i, :] = src[idx[i], :] for rows of `words_per_row` 32-bit words; inv[perm[i]] = i */
int neuma_gather_rows(neuma_ctx* ctx, int n, int words_per_row, const void* src, const int32_t* idx,
                      void* dst, void* stream);
int neuma_invert_perm(neuma_ctx* ctx, int n, const int32_t* perm, int32_t* inv, void* stream);

/* ---- SVD: modules/nclaw/warp/svd.py:59-96 ------------------------------ */
/* batch_svd incl. the det<0 sign fix; outputs U [n,3,3], sigma [n,3], Vh [n,3,3] */
int neuma_svd3_fwd(neuma_ctx* ctx, int n, const float* A, float* U, float* sigma, float* Vh,
                   void* stream);
/* adjoint of the above given dL/dU, dL/dsigma, dL/dVh (any may be NULL = 0) */
int neuma_svd3_bwd(neuma_ctx* ctx, int n, const float* U, const float* sigma, const float* Vh,
                   const float* gU, const float* gsigma, const float* gVh, float* gA,
                   void* stream);

/* ---- neural constitutive law: modules/nclaw/material/meta.py ----------- */
/* kind 0: InvariantFullMetaElasticity.forward (meta.py:196-221), out = Kirchhoff
 *         stress  R sym(MLP(feat)) F^T
 * kind 1: InvariantFullMetaPlasticity.forward (meta.py:468-489), out = F + alpha R sym(MLP(feat))
 * W0 [H,13], W1 [H,H], W2 [9,H]: bias-free Linear weights (H = hidden = 64),
 * with any LoRA adaptor already merged (W + scaling * B A, loralib.py:216-224);
 * the host wrapper un-merges the weight gradients. */
int neuma_constitutive_fwd(neuma_ctx* ctx, int kind, int n, int hidden,
                           const float* F, const float* W0, const float* W1, const float* W2,
                           float alpha, float* out, void* stream);
/* gW* are [same shapes], ACCUMULATED into (caller zero-fills); may be NULL to skip
 * parameter gradients.  gF is overwritten. */
int neuma_constitutive_bwd(neuma_ctx* ctx, int kind, int n, int hidden,
                           const float* F, const float* W0, const float* W1, const float* W2,
                           float alpha, const float* gout,
                           float* gF, float* gW0, float* gW1, float* gW2, void* stream);

/* The kernels read the weights from a "staged" image (every fp32 weight split into bf16
 * terms, laid out as tensor-core operands): neuma_constitutive_staged_bytes() bytes of
 * caller-owned device memory filled by neuma_constitutive_stage.  The two entry points
 * above stage into a ctx-owned buffer on every call; a rollout that evaluates the same
 * weights many times (finetune.py:360-365, 2 x S calls per frame) stages once and calls
 * the *_staged variants. */
int neuma_constitutive_staged_bytes(void);
int neuma_constitutive_stage(neuma_ctx* ctx, int hidden, const float* W0, const float* W1,
                             const float* W2, void* staged, void* stream);
int neuma_constitutive_fwd_staged(neuma_ctx* ctx, int kind, int n, int hidden, const float* F,
                                  const void* staged, float alpha, float* out, void* stream);
int neuma_constitutive_bwd_staged(neuma_ctx* ctx, int kind, int n, int hidden, const float* F,
                                  const void* staged, float alpha, const float* gout, float* gF,
                                  float* gW0, float* gW1, float* gW2, void* stream);

/* Weight gradients without atomics: every CTA adds into its own slot of `gWpart`
 * (neuma_constitutive_partials_floats(ctx) floats, zero-initialised by the caller once);
 * neuma_constitutive_reduce_partials adds the slot sum into gW0/gW1/gW2 and zeroes the
 * slots.  A rollout calls the first 2 x S times per frame and the second once. */
int neuma_constitutive_partials_floats(neuma_ctx* ctx);
/* accumulate_gF != 0: gF += dL/dF instead of overwriting it */
int neuma_constitutive_bwd_partials(neuma_ctx* ctx, int kind, int n, int hidden, const float* F,
                                    const void* staged, float alpha, const float* gout, float* gF,
                                    int accumulate_gF, float* gWpart, void* stream);
int neuma_constitutive_reduce_partials(neuma_ctx* ctx, int hidden, float* gWpart, float* gW0,
                                       float* gW1, float* gW2, void* stream);

/* MPM work fused into the constitutive kernels' epilogues (what FusedFrame / SlabRollout run):
 *   neuma_constitutive_fwd_p2g     elasticity (kind 0 or 2) + p2g (mpm.py:321-371) of every enabled
 *                                  particle into `grid` -- finetune.py:362-363 as ONE kernel; the
 *                                  stress is still written (the backward reads it);
 *   neuma_constitutive_bwd_scatter plasticity backward + the adjoint scatter of g2p into `adj`
 *                                  (= neuma_constitutive_bwd_partials(kind 1) followed by
 *                                  neuma_mpm_bwd_scatter with gF2 = the gFtrial just computed).
 *   neuma_constitutive_fwd_g2p     g2p (mpm.py:432-498) + plasticity of every particle: the thread gathers
 *                                  x', v', C', F_trial (all written, as neuma_mpm_g2p_from does, box_next
 *                                  included) and maps F_trial to F_out = F[s+1] -- finetune.py:363-364;
 *   neuma_constitutive_bwd_gather  the adjoint gather of the substep (neuma_mpm_bwd_gather_v) + the
 *                                  elasticity backward that consumes its dL/dstress: writes gx, gv, gC,
 *                                  gstress and the complete gF (MPM part + elasticity part).
 * c / box / grid / adj as in the rollout entry points (c->sel_mode applies to the MPM part; the g2p /
 * gather fusions need sel_mode 0). */
int neuma_constitutive_fwd_p2g(neuma_ctx* ctx, int kind, int n, int hidden, const float* F, const void* staged,
                               float* stress, const neuma_mpm_const* c, const float* x, const float* v,
                               const float* C, const float* vol, const float* rho, const int32_t* enabled,
                               const int32_t* box, float* grid, void* stream);
int neuma_constitutive_bwd_scatter(neuma_ctx* ctx, int n, int hidden, const float* Ftrial, const void* staged,
                                   float alpha, const float* gout, float* gFtrial, float* gWpart,
                                   const neuma_mpm_const* c, const float* x, const float* F,
                                   const float* clip_bound, const int32_t* enabled, const float* v2,
                                   const float* gx2, const float* gv2, const float* gC2, const int32_t* box,
                                   float* adj, void* stream);
/* Rollouts evaluate the backward of a net on exactly the rows its forward saw (finetune.py:360-414): the
 * forward can hand the right singular vectors V of every input row (9 floats per particle) to the backward,
 * which then skips the Jacobi iteration of the 3x3 SVD (warp/svd.py:59-96 recomputes U, sigma, V^T from the
 * saved tensors as well).  Applies to the NEXT constitutive launch on this ctx only, whichever entry point:
 * a forward kernel writes v_out[n][9] (if not NULL), a backward kernel reads v_in[n][9] (if not NULL). */
int neuma_constitutive_svd_cache(neuma_ctx* ctx, float* v_out, const float* v_in);
int neuma_constitutive_fwd_g2p(neuma_ctx* ctx, int n, int hidden, const void* staged, float alpha, float* F_out,
                               const neuma_mpm_const* c, const float* x, const float* F, const float* clip_bound,
                               const int32_t* enabled, const int32_t* box, const float* grid, float* x2, float* v2,
                               float* C2, float* Ftrial, int32_t* box_next, void* stream);
int neuma_constitutive_bwd_gather(neuma_ctx* ctx, int kind, int n, int hidden, const void* staged, float* gWpart,
                                  const neuma_mpm_const* c, const float* x, const float* v, const float* C,
                                  const float* F, const float* stress, const float* vol, const float* rho,
                                  const float* clip_bound, const int32_t* enabled, const float* v2, const float* C2,
                                  const float* gx2, const float* gv2, const float* gC2, const float* gF2,
                                  const int32_t* box, const float* grid, const float* adj, float* gx, float* gv,
                                  float* gC, float* gF, float* gstress, void* stream);

/* ---- particle -> Gaussian binding: modules/tune/utils.py:424-472 ------- */
/* out[K,ncol] = B in[Np,ncol] (transpose=0) or out[Np,ncol] = B^T in[K,ncol]
 * (transpose=1, the backward), B in CSR (rowptr[K+1], col[nnz], w[nnz]). */
int neuma_bind_spmm(neuma_ctx* ctx, int K, int Np, const int32_t* rowptr, const int32_t* col,
                    const float* w, int ncol, const float* in, float* out, int transpose,
                    void* stream);
/* The whole sim -> render hand-over of a frame in one launch:
 * denormalize_points_helper_func (nclaw/utils.py:110-118) + compute_bindings_xyz + compute_bindings_F
 * (tune/utils.py:424-472) + deform_cov_by_F (simulation_utils.py:25-48).  `slot` (may be NULL)
 * maps a caller-order particle index to its row in x / F; center_host / size_host: float[3] on
 * the HOST (NULL = 0 / 1); Fk_out may be NULL.  The second entry is the transpose for
 * dL/dmeans3D -> dL/dx (accumulates into gx, which the caller zero-fills). */
int neuma_particles_to_gaussians(neuma_ctx* ctx, int K, const int32_t* rowptr, const int32_t* col,
                                 const float* w, const int32_t* slot, const float* x, const float* F,
                                 const float* center_host, const float* size_host,
                                 const float* de_x_prev, const float* k_prev, const float* cov6,
                                 float* means3D, float* cov_out, float* Fk_out, void* stream);
int neuma_gaussians_to_particles(neuma_ctx* ctx, int K, const int32_t* rowptr, const int32_t* col,
                                 const float* w, const int32_t* slot, const float* g_means3D,
                                 const float* size_host, float* gx, void* stream);
/* gaussian_binding_with_clip_v1, modules/d3gs/utils/binding_utils.py:199-285: for
 * every kernel k the particles with d^T Sigma^-1 d <= threshold (chi2.ppf(confidence, 3)), the
 * `max_particles` (<= 16) nearest of them in that metric.  counts[K] = number inside (uncapped),
 * cols[K, max_particles] = selected particle ids in ascending distance, -1 padded; pvals (may be
 * NULL) the distances.  The weights are uniform 1/n (binding_utils.py:265-274). */
int neuma_bind_build(neuma_ctx* ctx, int K, int Np, const float* means, const float* cov6,
                     const float* x, float threshold, int max_particles, int32_t* counts,
                     int32_t* cols, float* pvals, void* stream);
/* deform_cov_by_F, modules/d3gs/utils/simulation_utils.py:25-48 */
int neuma_deform_cov(neuma_ctx* ctx, int K, const float* cov6, const float* F, float* out6,
                     void* stream);

/* ---- Gaussian rasterizer: diff_gaussian_rasterization (external), call site
 *      modules/tune/utils.py:385-419 ------------------------------------- */
/* shs [K,M,3] (M >= (deg+1)^2) or colors_precomp [K,3] (exactly one non-NULL);
 * opacities [K]; cov6 [K,6]; image [3,H,W]; radii int32[K].
 * Keeps per-call state in the ctx for the matching backward. */
int neuma_raster_fwd(neuma_ctx* ctx, const neuma_raster_settings* s, int K, int M,
                     const float* means3D, const float* shs, const float* colors_precomp,
                     const float* opacities, const float* cov6,
                     float* image, int32_t* radii, void* stream);
/* dL/dimage [3,H,W] -> dL/dmeans3D [K,3], dL/dmeans2D [K,3] (z = 0); optional
 * (may be NULL) dL/dcov6 [K,6], dL/dopacity [K], dL/dcolors or dL/dshs. */
int neuma_raster_bwd(neuma_ctx* ctx, const neuma_raster_settings* s, int K, int M,
                     const float* means3D, const float* shs, const float* colors_precomp,
                     const float* opacities, const float* cov6, const int32_t* radii,
                     const float* dL_dimage,
                     const void* saved_state /* from neuma_raster_state_save, or NULL = the
                                                ctx's last forward */,
                     int saved_num_rendered /* instance count of that forward */,
                     float* dL_dmeans3D, float* dL_dmeans2D, float* dL_dcov6,
                     float* dL_dopacity, float* dL_dcolor_or_sh, void* stream);
/* A rollout runs several forwards before their backwards (finetune.py:360-414), so the
 * per-call forward state the backward needs (projected means, conics, colours, the sorted
 * instance list, tile ranges, final transmittance, contributor counts -- upstream's
 * geomBuffer / binningBuffer / imgBuffer) can be copied out into a caller-owned device
 * buffer of neuma_raster_state_bytes() bytes and handed back to neuma_raster_bwd. */
int neuma_raster_state_bytes(neuma_ctx* ctx, size_t* out_host);
int neuma_raster_state_save(neuma_ctx* ctx, void* dst, void* stream);
/* introspection for the bit-exact tile/sort parity tests: number of
 * (tile, gaussian) instances of the last forward, then the sorted 64-bit keys
 * (tile << 32 | depth bits), Gaussian ids, and per-tile [start,end) ranges.
 * Synchronises `stream`. */
int neuma_raster_num_rendered(neuma_ctx* ctx, void* stream, int* out_host);
int neuma_raster_export_binning(neuma_ctx* ctx, uint64_t* keys, int32_t* gaussian_ids,
                                int32_t* tile_ranges /* [tiles,2] */, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NEUMA_B200_H */
