// Tile-based Gaussian-splat rasterizer, forward and backward, for sm_100a.
// Replaces the external `diff_gaussian_rasterization` extension
// (graphdeco-inria/gaussian-splatting@b17ded92, pinned at README.md:56-60; source
// NOT under /root/reference) at its call site modules/tune/utils.py:385-419, with
// the settings of modules/d3gs/gaussian_renderer/__init__.py:92-119.
//
// The arithmetic restates the published 3DGS algorithm [UPSTREAM-UNVERIFIED]:
// cull z_view <= 0.2; EWA cov2D = J W S W^T J^T with x/z, y/z clamped to
// 1.3 tan(fov/2) and +0.3 on the diagonal; conic = inverse; radius =
// ceil(3 sqrt(lambda_max)); 16x16-pixel tiles; instance key = tile << 32 | float
// bits of depth, stable LSD radix sort; alpha = min(0.99, o exp(-1/2 d^T Q d)),
// skip alpha < 1/255, stop before T < 1e-4; pixel = sum c alpha T + T_final bg;
// SH degree <= 3 with +0.5 and clamp at 0.
//
// B200 notes: the per-pixel blend keeps the classic one-CTA-per-tile shape
// (the work is latency/occupancy bound at NeuMA's ~200 active tiles); the
// backward reduces the per-Gaussian gradient contributions of a warp's 32 pixels
// with shuffles before one atomic per warp (instead of 32), and the per-Gaussian
// records are read as 128-bit vectors.
#include "common.cuh"
#include <cub/cub.cuh>

namespace {
constexpr int BX = 16, BY = 16, BLOCK = BX * BY;
}

struct RasterState {
  // per Gaussian
  int capK = 0;
  float* depths = nullptr;
  float2* xy = nullptr;
  float4* conic_o = nullptr;
  float* rgb = nullptr;
  unsigned char* clamped = nullptr;
  uint32_t* tiles_touched = nullptr;
  uint32_t* offsets = nullptr;
  float4* dconic = nullptr;  // backward: dL/d(A,B,C) of the conic, .w = dL/dopacity
  float* dcolor = nullptr;   // backward: dL/drgb [K,3]
  // per instance
  size_t capD = 0;
  uint64_t *keys_a = nullptr, *keys_b = nullptr;
  uint32_t *vals_a = nullptr, *vals_b = nullptr;
  // per tile / pixel
  int capT = 0;
  uint2* ranges = nullptr;
  int capPix = 0;
  float* final_T = nullptr;
  uint32_t* n_contrib = nullptr;
  // cub temp
  void* temp = nullptr;
  size_t temp_bytes = 0;
  // last forward
  int num_rendered = 0, K = 0, tiles_x = 0, tiles_y = 0, W = 0, Himg = 0;
  int* num_rendered_pinned = nullptr;
};

void neuma_raster_state_free(RasterState* s) {
  if (!s) return;
  cudaFree(s->depths); cudaFree(s->xy); cudaFree(s->conic_o); cudaFree(s->rgb);
  cudaFree(s->clamped); cudaFree(s->tiles_touched); cudaFree(s->offsets);
  cudaFree(s->dconic); cudaFree(s->dcolor);
  cudaFree(s->keys_a); cudaFree(s->keys_b); cudaFree(s->vals_a); cudaFree(s->vals_b);
  cudaFree(s->ranges); cudaFree(s->final_T); cudaFree(s->n_contrib); cudaFree(s->temp);
  if (s->num_rendered_pinned) cudaFreeHost(s->num_rendered_pinned);
  delete s;
}

namespace {

struct Cam {
  float view[16], proj[16];
  float cam[3], bg[3];
  float tanx, tany, fx, fy;
  int W, H, tiles_x, tiles_y, deg;
};

Cam make_cam(const neuma_raster_settings* s) {
  Cam c;
  memcpy(c.view, s->viewmatrix, sizeof(c.view));
  memcpy(c.proj, s->projmatrix, sizeof(c.proj));
  memcpy(c.cam, s->campos, sizeof(c.cam));
  memcpy(c.bg, s->bg, sizeof(c.bg));
  c.tanx = s->tanfovx; c.tany = s->tanfovy;
  c.W = s->image_width; c.H = s->image_height;
  c.fx = c.W / (2.0f * c.tanx);
  c.fy = c.H / (2.0f * c.tany);
  c.tiles_x = (c.W + BX - 1) / BX;
  c.tiles_y = (c.H + BY - 1) / BY;
  c.deg = s->sh_degree;
  return c;
}

__device__ __forceinline__ void xform43(const float* m, const float p[3], float o[3]) {
  o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
  o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
  o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
}
__device__ __forceinline__ void xform44(const float* m, const float p[3], float o[4]) {
  o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
  o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
  o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
  o[3] = m[3] * p[0] + m[7] * p[1] + m[11] * p[2] + m[15];
}
__device__ __forceinline__ float ndc2pix(float v, int S) { return ((v + 1.0f) * S - 1.0f) * 0.5f; }

// real SH basis (degree <= 3) and its gradient wrt the unit direction
#define SH_C0 0.28209479177387814f
#define SH_C1 0.4886025119029199f
__device__ __forceinline__ void sh_basis(int deg, float x, float y, float z, float b[16],
                                         float gx[16], float gy[16], float gz[16]) {
  const float C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                       -1.0925484305920792f, 0.5462742152960396f};
  const float C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                       0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                       -0.5900435899266435f};
#pragma unroll
  for (int i = 0; i < 16; ++i) { b[i] = 0.f; gx[i] = 0.f; gy[i] = 0.f; gz[i] = 0.f; }
  b[0] = SH_C0;
  if (deg > 0) {
    b[1] = -SH_C1 * y; gy[1] = -SH_C1;
    b[2] = SH_C1 * z;  gz[2] = SH_C1;
    b[3] = -SH_C1 * x; gx[3] = -SH_C1;
    if (deg > 1) {
      const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
      b[4] = C2[0] * xy; gx[4] = C2[0] * y; gy[4] = C2[0] * x;
      b[5] = C2[1] * yz; gy[5] = C2[1] * z; gz[5] = C2[1] * y;
      b[6] = C2[2] * (2.0f * zz - xx - yy);
      gx[6] = C2[2] * -2.0f * x; gy[6] = C2[2] * -2.0f * y; gz[6] = C2[2] * 4.0f * z;
      b[7] = C2[3] * xz; gx[7] = C2[3] * z; gz[7] = C2[3] * x;
      b[8] = C2[4] * (xx - yy); gx[8] = C2[4] * 2.0f * x; gy[8] = C2[4] * -2.0f * y;
      if (deg > 2) {
        b[9] = C3[0] * y * (3.0f * xx - yy);
        gx[9] = C3[0] * 6.0f * xy; gy[9] = C3[0] * (3.0f * xx - 3.0f * yy);
        b[10] = C3[1] * xy * z;
        gx[10] = C3[1] * yz; gy[10] = C3[1] * xz; gz[10] = C3[1] * xy;
        b[11] = C3[2] * y * (4.0f * zz - xx - yy);
        gx[11] = C3[2] * -2.0f * xy; gy[11] = C3[2] * (4.0f * zz - xx - 3.0f * yy);
        gz[11] = C3[2] * 8.0f * yz;
        b[12] = C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
        gx[12] = C3[3] * -6.0f * xz; gy[12] = C3[3] * -6.0f * yz;
        gz[12] = C3[3] * (6.0f * zz - 3.0f * xx - 3.0f * yy);
        b[13] = C3[4] * x * (4.0f * zz - xx - yy);
        gx[13] = C3[4] * (4.0f * zz - 3.0f * xx - yy); gy[13] = C3[4] * -2.0f * xy;
        gz[13] = C3[4] * 8.0f * xz;
        b[14] = C3[5] * z * (xx - yy);
        gx[14] = C3[5] * 2.0f * xz; gy[14] = C3[5] * -2.0f * yz; gz[14] = C3[5] * (xx - yy);
        b[15] = C3[6] * x * (xx - 3.0f * yy);
        gx[15] = C3[6] * (3.0f * xx - 3.0f * yy); gy[15] = C3[6] * -6.0f * xy;
      }
    }
  }
}

// EWA projection of a packed-6 covariance; returns cov2D (a, b, c) incl. the 0.3
// low-pass, plus the intermediates the backward needs.
struct Cov2D {
  float a, b, c;
  float T[2][3];   // T = J W  (rows 0,1; row 2 of J is zero)
  float t[3];      // clamped view-space mean
  float xmul, ymul;
};
__device__ __forceinline__ Cov2D cov2d(const Cam& cam, const float p[3], const float* cv) {
  Cov2D r;
  float t[3];
  xform43(cam.view, p, t);
  const float limx = 1.3f * cam.tanx, limy = 1.3f * cam.tany;
  const float txtz = t[0] / t[2], tytz = t[1] / t[2];
  r.xmul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
  r.ymul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
  t[0] = fminf(limx, fmaxf(-limx, txtz)) * t[2];
  t[1] = fminf(limy, fmaxf(-limy, tytz)) * t[2];
  r.t[0] = t[0]; r.t[1] = t[1]; r.t[2] = t[2];
  const float J00 = cam.fx / t[2], J02 = -(cam.fx * t[0]) / (t[2] * t[2]);
  const float J11 = cam.fy / t[2], J12 = -(cam.fy * t[1]) / (t[2] * t[2]);
  // Wr[i][j]: view = Wr p + trans ; Wr[i][j] = view[4*j + i]
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    r.T[0][j] = J00 * cam.view[4 * j + 0] + J02 * cam.view[4 * j + 2];
    r.T[1][j] = J11 * cam.view[4 * j + 1] + J12 * cam.view[4 * j + 2];
  }
  const float S[3][3] = {{cv[0], cv[1], cv[2]}, {cv[1], cv[3], cv[4]}, {cv[2], cv[4], cv[5]}};
  float TS[2][3];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      TS[i][j] = r.T[i][0] * S[0][j] + r.T[i][1] * S[1][j] + r.T[i][2] * S[2][j];
  r.a = TS[0][0] * r.T[0][0] + TS[0][1] * r.T[0][1] + TS[0][2] * r.T[0][2] + 0.3f;
  r.b = TS[0][0] * r.T[1][0] + TS[0][1] * r.T[1][1] + TS[0][2] * r.T[1][2];
  r.c = TS[1][0] * r.T[1][0] + TS[1][1] * r.T[1][1] + TS[1][2] * r.T[1][2] + 0.3f;
  return r;
}

__device__ __forceinline__ void tile_rect(const Cam& cam, float px, float py, int radius,
                                          int& x0, int& y0, int& x1, int& y1) {
  x0 = min(cam.tiles_x, max(0, (int)((px - radius) / BX)));
  y0 = min(cam.tiles_y, max(0, (int)((py - radius) / BY)));
  x1 = min(cam.tiles_x, max(0, (int)((px + radius + BX - 1) / BX)));
  y1 = min(cam.tiles_y, max(0, (int)((py + radius + BY - 1) / BY)));
}

// ---------------------------------------------------------------- preprocess
__global__ void __launch_bounds__(256)
preprocess_kernel(Cam cam, int K, int M, const float* __restrict__ means,
                  const float* __restrict__ shs, const float* __restrict__ colors_precomp,
                  const float* __restrict__ opac, const float* __restrict__ cov6,
                  int* __restrict__ radii, float* __restrict__ depths, float2* __restrict__ xy,
                  float4* __restrict__ conic_o, float* __restrict__ rgb,
                  unsigned char* __restrict__ clamped, uint32_t* __restrict__ tiles_touched) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K) return;
  radii[i] = 0;
  tiles_touched[i] = 0;
  const float p[3] = {means[3 * i], means[3 * i + 1], means[3 * i + 2]};
  float pv[3];
  xform43(cam.view, p, pv);
  if (pv[2] <= 0.2f) return;  // near-plane cull
  float ph[4];
  xform44(cam.proj, p, ph);
  const float pw = 1.0f / (ph[3] + 0.0000001f);
  const float pp[2] = {ph[0] * pw, ph[1] * pw};
  const Cov2D cv = cov2d(cam, p, cov6 + (size_t)i * 6);
  const float det = cv.a * cv.c - cv.b * cv.b;
  if (det == 0.0f) return;
  const float di = 1.0f / det;
  const float mid = 0.5f * (cv.a + cv.c);
  const float l1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
  const float l2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
  const int radius = (int)ceilf(3.0f * sqrtf(fmaxf(l1, l2)));
  const float px = ndc2pix(pp[0], cam.W), py = ndc2pix(pp[1], cam.H);
  int x0, y0, x1, y1;
  tile_rect(cam, px, py, radius, x0, y0, x1, y1);
  if ((x1 - x0) * (y1 - y0) == 0) return;
  if (colors_precomp) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      rgb[3 * i + c] = colors_precomp[3 * i + c];
      clamped[3 * i + c] = 0;
    }
  } else {
    float d[3] = {p[0] - cam.cam[0], p[1] - cam.cam[1], p[2] - cam.cam[2]};
    const float il = 1.0f / sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    d[0] *= il; d[1] *= il; d[2] *= il;
    float b[16], gx[16], gy[16], gz[16];
    sh_basis(cam.deg, d[0], d[1], d[2], b, gx, gy, gz);
    const int nb = (cam.deg + 1) * (cam.deg + 1);
    const float* sh = shs + (size_t)i * M * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float r = 0.f;
      for (int k = 0; k < nb; ++k) r += b[k] * sh[3 * k + c];
      r += 0.5f;
      clamped[3 * i + c] = r < 0.f;
      rgb[3 * i + c] = fmaxf(r, 0.f);
    }
  }
  depths[i] = pv[2];
  radii[i] = radius;
  xy[i] = make_float2(px, py);
  conic_o[i] = make_float4(cv.c * di, -cv.b * di, cv.a * di, opac[i]);
  tiles_touched[i] = (uint32_t)((x1 - x0) * (y1 - y0));
}

__global__ void __launch_bounds__(256)
duplicate_kernel(Cam cam, int K, const float2* __restrict__ xy, const float* __restrict__ depths,
                 const uint32_t* __restrict__ offsets, const int* __restrict__ radii,
                 uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K) return;
  if (radii[i] <= 0) return;
  uint32_t off = (i == 0) ? 0 : offsets[i - 1];
  int x0, y0, x1, y1;
  tile_rect(cam, xy[i].x, xy[i].y, radii[i], x0, y0, x1, y1);
  const uint32_t dbits = __float_as_uint(depths[i]);
  for (int y = y0; y < y1; ++y)
    for (int x = x0; x < x1; ++x) {
      const uint64_t key = ((uint64_t)(y * cam.tiles_x + x) << 32) | dbits;
      keys[off] = key;
      vals[off] = (uint32_t)i;
      ++off;
    }
}

__global__ void __launch_bounds__(256)
tile_ranges_kernel(int L, const uint64_t* __restrict__ keys, uint2* __restrict__ ranges) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  const uint32_t t = (uint32_t)(keys[i] >> 32);
  if (i == 0) {
    ranges[t].x = 0;
  } else {
    const uint32_t tp = (uint32_t)(keys[i - 1] >> 32);
    if (t != tp) {
      ranges[tp].y = i;
      ranges[t].x = i;
    }
  }
  if (i == L - 1) ranges[t].y = L;
}

// ---------------------------------------------------------------- render fwd
__global__ void __launch_bounds__(BLOCK)
render_fwd_kernel(Cam cam, const uint2* __restrict__ ranges, const uint32_t* __restrict__ plist,
                  const float2* __restrict__ xy, const float* __restrict__ rgb,
                  const float4* __restrict__ conic_o, float* __restrict__ final_T,
                  uint32_t* __restrict__ n_contrib, float* __restrict__ out) {
  const int tx = blockIdx.x, ty = blockIdx.y;
  const int pxi = tx * BX + threadIdx.x, pyi = ty * BY + threadIdx.y;
  const int tid = threadIdx.y * BX + threadIdx.x;
  const bool inside = pxi < cam.W && pyi < cam.H;
  const float pfx = (float)pxi, pfy = (float)pyi;
  const uint2 range = ranges[ty * cam.tiles_x + tx];
  const int rounds = ((int)(range.y - range.x) + BLOCK - 1) / BLOCK;
  int todo = (int)(range.y - range.x);
  __shared__ float2 s_xy[BLOCK];
  __shared__ float4 s_co[BLOCK];
  __shared__ float4 s_rgb[BLOCK];
  bool done = !inside;
  float T = 1.0f;
  uint32_t contributor = 0, last = 0;
  float C[3] = {0.f, 0.f, 0.f};
  for (int r = 0; r < rounds; ++r, todo -= BLOCK) {
    if (__syncthreads_count(done) == BLOCK) break;
    const int prog = r * BLOCK + tid;
    if ((int)range.x + prog < (int)range.y) {
      const uint32_t id = plist[range.x + prog];
      s_xy[tid] = xy[id];
      s_co[tid] = conic_o[id];
      s_rgb[tid] = make_float4(rgb[3 * id], rgb[3 * id + 1], rgb[3 * id + 2], 0.f);
    }
    __syncthreads();
    for (int j = 0; !done && j < min(BLOCK, todo); ++j) {
      ++contributor;
      const float2 q = s_xy[j];
      const float dx = q.x - pfx, dy = q.y - pfy;
      const float4 co = s_co[j];
      const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
      if (power > 0.0f) continue;
      const float alpha = fminf(0.99f, co.w * __expf(power));
      if (alpha < 1.0f / 255.0f) continue;
      const float test_T = T * (1.0f - alpha);
      if (test_T < 0.0001f) {
        done = true;
        continue;
      }
      const float4 c = s_rgb[j];
      const float wgt = alpha * T;
      C[0] += c.x * wgt; C[1] += c.y * wgt; C[2] += c.z * wgt;
      T = test_T;
      last = contributor;
    }
  }
  if (inside) {
    const int pix = pyi * cam.W + pxi;
    final_T[pix] = T;
    n_contrib[pix] = last;
    const size_t HW = (size_t)cam.H * cam.W;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) out[ch * HW + pix] = C[ch] + T * cam.bg[ch];
  }
}

// ---------------------------------------------------------------- render bwd
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(BLOCK)
render_bwd_kernel(Cam cam, const uint2* __restrict__ ranges, const uint32_t* __restrict__ plist,
                  const float2* __restrict__ xy, const float* __restrict__ rgb,
                  const float4* __restrict__ conic_o, const float* __restrict__ final_T,
                  const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpix,
                  float* __restrict__ dmean2d /* [K,3] */, float4* __restrict__ dconic,
                  float* __restrict__ dcolor) {
  const int tx = blockIdx.x, ty = blockIdx.y;
  const int pxi = tx * BX + threadIdx.x, pyi = ty * BY + threadIdx.y;
  const int tid = threadIdx.y * BX + threadIdx.x;
  const int lane = tid & 31;
  const bool inside = pxi < cam.W && pyi < cam.H;
  const float pfx = (float)pxi, pfy = (float)pyi;
  const uint2 range = ranges[ty * cam.tiles_x + tx];
  const int total = (int)(range.y - range.x);
  const int rounds = (total + BLOCK - 1) / BLOCK;
  int todo = total;
  __shared__ uint32_t s_id[BLOCK];
  __shared__ float2 s_xy[BLOCK];
  __shared__ float4 s_co[BLOCK];
  __shared__ float4 s_rgb[BLOCK];
  const int pix = pyi * cam.W + pxi;
  const size_t HW = (size_t)cam.H * cam.W;
  const float T_final = inside ? final_T[pix] : 0.f;
  float T = T_final;
  uint32_t contributor = (uint32_t)total;
  const uint32_t last_contrib = inside ? n_contrib[pix] : 0;
  float dp[3] = {0.f, 0.f, 0.f};
  if (inside) {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) dp[ch] = dL_dpix[ch * HW + pix];
  }
  const float bg_dot = cam.bg[0] * dp[0] + cam.bg[1] * dp[1] + cam.bg[2] * dp[2];
  float accum[3] = {0.f, 0.f, 0.f}, last_color[3] = {0.f, 0.f, 0.f};
  float last_alpha = 0.f;
  const float ddelx = 0.5f * cam.W, ddely = 0.5f * cam.H;
  for (int r = 0; r < rounds; ++r, todo -= BLOCK) {
    __syncthreads();
    const int prog = r * BLOCK + tid;
    if (prog < total) {
      const uint32_t id = plist[range.y - prog - 1];  // back to front
      s_id[tid] = id;
      s_xy[tid] = xy[id];
      s_co[tid] = conic_o[id];
      s_rgb[tid] = make_float4(rgb[3 * id], rgb[3 * id + 1], rgb[3 * id + 2], 0.f);
    }
    __syncthreads();
    const int nj = min(BLOCK, todo);
    for (int j = 0; j < nj; ++j) {
      --contributor;
      // per-pixel contribution (zero if this Gaussian did not blend into the pixel)
      float g_mx = 0.f, g_my = 0.f, g_ca = 0.f, g_cb = 0.f, g_cc = 0.f, g_op = 0.f;
      float g_col[3] = {0.f, 0.f, 0.f};
      bool active = inside && contributor < last_contrib;
      if (active) {
        const float2 q = s_xy[j];
        const float dx = q.x - pfx, dy = q.y - pfy;
        const float4 co = s_co[j];
        const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
        if (power > 0.0f) active = false;
        if (active) {
          const float G = __expf(power);
          const float alpha = fminf(0.99f, co.w * G);
          if (alpha < 1.0f / 255.0f) active = false;
          if (active) {
            T = T / (1.0f - alpha);
            const float dch = alpha * T;
            const float4 c4 = s_rgb[j];
            const float col[3] = {c4.x, c4.y, c4.z};
            float dL_dalpha = 0.f;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
              accum[ch] = last_alpha * last_color[ch] + (1.0f - last_alpha) * accum[ch];
              last_color[ch] = col[ch];
              dL_dalpha += (col[ch] - accum[ch]) * dp[ch];
              g_col[ch] = dch * dp[ch];
            }
            dL_dalpha *= T;
            last_alpha = alpha;
            dL_dalpha += (-T_final / (1.0f - alpha)) * bg_dot;
            const float dL_dG = co.w * dL_dalpha;
            const float gdx = G * dx, gdy = G * dy;
            const float dG_dx = -gdx * co.x - gdy * co.y;
            const float dG_dy = -gdy * co.z - gdx * co.y;
            g_mx = dL_dG * dG_dx * ddelx;
            g_my = dL_dG * dG_dy * ddely;
            g_ca = -0.5f * gdx * dx * dL_dG;
            g_cb = -gdx * dy * dL_dG;  // true d/dB (B multiplies dx*dy once)
            g_cc = -0.5f * gdy * dy * dL_dG;
            g_op = G * dL_dalpha;
          }
        }
      }
      // one atomic per warp per Gaussian instead of one per pixel
      if (__any_sync(0xffffffffu, active)) {
        g_mx = warp_sum(g_mx); g_my = warp_sum(g_my);
        g_ca = warp_sum(g_ca); g_cb = warp_sum(g_cb); g_cc = warp_sum(g_cc);
        g_op = warp_sum(g_op);
        g_col[0] = warp_sum(g_col[0]); g_col[1] = warp_sum(g_col[1]); g_col[2] = warp_sum(g_col[2]);
        if (lane == 0) {
          const uint32_t id = s_id[j];
          atomicAdd(dmean2d + 3 * id, g_mx);
          atomicAdd(dmean2d + 3 * id + 1, g_my);
          red_add_f4(dconic + id, g_ca, g_cb, g_cc, g_op);
          atomicAdd(dcolor + 3 * id, g_col[0]);
          atomicAdd(dcolor + 3 * id + 1, g_col[1]);
          atomicAdd(dcolor + 3 * id + 2, g_col[2]);
        }
      }
    }
  }
}

// ---------------------------------------------------------------- preprocess bwd
__global__ void __launch_bounds__(256)
preprocess_bwd_kernel(Cam cam, int K, int M, const float* __restrict__ means,
                      const float* __restrict__ shs, const float* __restrict__ colors_precomp,
                      const float* __restrict__ cov6, const int* __restrict__ radii,
                      const unsigned char* __restrict__ clamped,
                      const float* __restrict__ dmean2d, const float4* __restrict__ dconic,
                      const float* __restrict__ dcolor, float* __restrict__ dmeans3d,
                      float* __restrict__ dcov6, float* __restrict__ dopac,
                      float* __restrict__ dcol_or_sh) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K) return;
  float gm[3] = {0.f, 0.f, 0.f};
  const bool vis = radii[i] > 0;
  const float4 dc = vis ? dconic[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  if (dopac) dopac[i] = dc.w;
  float gcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (vis) {
    const float p[3] = {means[3 * i], means[3 * i + 1], means[3 * i + 2]};
    const float* cvp = cov6 + (size_t)i * 6;
    // ---- conic -> cov2D -> (cov3D, mean)
    const Cov2D cv = cov2d(cam, p, cvp);
    const float a = cv.a, b = cv.b, c = cv.c;
    const float det = a * c - b * b;
    // Q = inverse; dSigma = -Q Gq Q with Gq = [[gA, gB/2],[gB/2, gC]]; the published
    // backward regularises 1/det^2 as 1/(det^2 + 1e-7) -- same here.
    const float reg = (det * det) / (det * det + 0.0000001f);
    const float di = 1.0f / det;
    const float A = c * di, B = -b * di, Cq = a * di;
    const float g00 = dc.x, g01 = 0.5f * dc.y, g11 = dc.z;
    // M = Q Gq
    const float m00 = A * g00 + B * g01, m01 = A * g01 + B * g11;
    const float m10 = B * g00 + Cq * g01, m11 = B * g01 + Cq * g11;
    // S = -(M Q)
    const float s00 = -(m00 * A + m01 * B) * reg;
    const float s01 = -(m00 * B + m01 * Cq) * reg;
    const float s10 = -(m10 * A + m11 * B) * reg;
    const float s11 = -(m10 * B + m11 * Cq) * reg;
    const float ga = s00, gb = s01 + s10, gc = s11;  // dL/d(a,b,c)
    // cov2D = T S3 T^T (+0.3 I):  a = T0 S3 T0^T, b = T0 S3 T1^T, c = T1 S3 T1^T
    const float S3[3][3] = {{cvp[0], cvp[1], cvp[2]}, {cvp[1], cvp[3], cvp[4]}, {cvp[2], cvp[4], cvp[5]}};
    // dL/dS3[j][k] (full matrix) = ga T0j T0k + gb/2 (T0j T1k + T1j T0k) + gc T1j T1k
    float gS[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int k = 0; k < 3; ++k)
        gS[j][k] = ga * cv.T[0][j] * cv.T[0][k] +
                   0.5f * gb * (cv.T[0][j] * cv.T[1][k] + cv.T[1][j] * cv.T[0][k]) +
                   gc * cv.T[1][j] * cv.T[1][k];
    gcov[0] = gS[0][0]; gcov[3] = gS[1][1]; gcov[5] = gS[2][2];
    gcov[1] = gS[0][1] + gS[1][0]; gcov[2] = gS[0][2] + gS[2][0]; gcov[4] = gS[1][2] + gS[2][1];
    // dL/dT0 = 2 ga S3 T0 + gb S3 T1 ; dL/dT1 = 2 gc S3 T1 + gb S3 T0
    float ST0[3], ST1[3], gT[2][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      ST0[j] = S3[j][0] * cv.T[0][0] + S3[j][1] * cv.T[0][1] + S3[j][2] * cv.T[0][2];
      ST1[j] = S3[j][0] * cv.T[1][0] + S3[j][1] * cv.T[1][1] + S3[j][2] * cv.T[1][2];
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      gT[0][j] = 2.0f * ga * ST0[j] + gb * ST1[j];
      gT[1][j] = 2.0f * gc * ST1[j] + gb * ST0[j];
    }
    // T0j = J00 Wr[0][j] + J02 Wr[2][j] ; T1j = J11 Wr[1][j] + J12 Wr[2][j]
    float gJ00 = 0.f, gJ02 = 0.f, gJ11 = 0.f, gJ12 = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      gJ00 += gT[0][j] * cam.view[4 * j + 0];
      gJ02 += gT[0][j] * cam.view[4 * j + 2];
      gJ11 += gT[1][j] * cam.view[4 * j + 1];
      gJ12 += gT[1][j] * cam.view[4 * j + 2];
    }
    const float tz = 1.0f / cv.t[2], tz2 = tz * tz, tz3 = tz2 * tz;
    float gt[3];
    gt[0] = cv.xmul * (-cam.fx * tz2 * gJ02);
    gt[1] = cv.ymul * (-cam.fy * tz2 * gJ12);
    gt[2] = -cam.fx * tz2 * gJ00 - cam.fy * tz2 * gJ11 + 2.0f * cam.fx * cv.t[0] * tz3 * gJ02 +
            2.0f * cam.fy * cv.t[1] * tz3 * gJ12;
    // t = Wr p + trans  ->  dL/dp = Wr^T gt
#pragma unroll
    for (int j = 0; j < 3; ++j)
      gm[j] += cam.view[4 * j + 0] * gt[0] + cam.view[4 * j + 1] * gt[1] + cam.view[4 * j + 2] * gt[2];
    // ---- mean2D -> mean (perspective projection)
    float ph[4];
    xform44(cam.proj, p, ph);
    const float mw = 1.0f / (ph[3] + 0.0000001f);
    const float mul1 = ph[0] * mw * mw, mul2 = ph[1] * mw * mw;
    const float gx2 = dmean2d[3 * i], gy2 = dmean2d[3 * i + 1];
#pragma unroll
    for (int j = 0; j < 3; ++j)
      gm[j] += (cam.proj[4 * j + 0] * mw - cam.proj[4 * j + 3] * mul1) * gx2 +
               (cam.proj[4 * j + 1] * mw - cam.proj[4 * j + 3] * mul2) * gy2;
    // ---- colour -> SH / view direction
    if (!colors_precomp) {
      float d[3] = {p[0] - cam.cam[0], p[1] - cam.cam[1], p[2] - cam.cam[2]};
      const float len2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
      const float il = 1.0f / sqrtf(len2);
      const float u[3] = {d[0] * il, d[1] * il, d[2] * il};
      float bb[16], gx[16], gy[16], gz[16];
      sh_basis(cam.deg, u[0], u[1], u[2], bb, gx, gy, gz);
      const int nb = (cam.deg + 1) * (cam.deg + 1);
      const float* sh = shs + (size_t)i * M * 3;
      float gu[3] = {0.f, 0.f, 0.f};
      for (int ch = 0; ch < 3; ++ch) {
        const float g = clamped[3 * i + ch] ? 0.f : dcolor[3 * i + ch];
        for (int k = 0; k < nb; ++k) {
          const float s = sh[3 * k + ch];
          gu[0] += g * gx[k] * s; gu[1] += g * gy[k] * s; gu[2] += g * gz[k] * s;
          if (dcol_or_sh) dcol_or_sh[((size_t)i * M + k) * 3 + ch] = g * bb[k];
        }
        if (dcol_or_sh)
          for (int k = nb; k < M; ++k) dcol_or_sh[((size_t)i * M + k) * 3 + ch] = 0.f;
      }
      // u = d / |d| : dL/dd = (gu - u (u . gu)) / |d|
      const float ud = u[0] * gu[0] + u[1] * gu[1] + u[2] * gu[2];
#pragma unroll
      for (int j = 0; j < 3; ++j) gm[j] += (gu[j] - u[j] * ud) * il;
    } else if (dcol_or_sh) {
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) dcol_or_sh[3 * i + ch] = dcolor[3 * i + ch];
    }
  } else if (dcol_or_sh) {
    const int cnt = colors_precomp ? 3 : M * 3;
    for (int k = 0; k < cnt; ++k) dcol_or_sh[(size_t)i * cnt + k] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) dmeans3d[3 * i + j] = gm[j];
  if (dcov6) {
#pragma unroll
    for (int j = 0; j < 6; ++j) dcov6[(size_t)i * 6 + j] = gcov[j];
  }
}

template <typename T>
int grow(neuma_ctx* ctx, T*& ptr, size_t count) {
  if (ptr) cudaFree(ptr);
  ptr = nullptr;
  NEUMA_CUDA(ctx, cudaMalloc(&ptr, count * sizeof(T)));
  return 0;
}

int ensure_state(neuma_ctx* ctx, int K, int tiles, int pix) {
  if (!ctx->raster) {
    ctx->raster = new RasterState();
    NEUMA_CUDA(ctx, cudaMallocHost(&ctx->raster->num_rendered_pinned, sizeof(int)));
  }
  RasterState* s = ctx->raster;
  if (K > s->capK) {
    const int cap = K + K / 4 + 1024;
    if (grow(ctx, s->depths, cap) || grow(ctx, s->xy, cap) || grow(ctx, s->conic_o, cap) ||
        grow(ctx, s->rgb, (size_t)cap * 3) || grow(ctx, s->clamped, (size_t)cap * 3) ||
        grow(ctx, s->tiles_touched, cap) || grow(ctx, s->offsets, cap) ||
        grow(ctx, s->dconic, cap) || grow(ctx, s->dcolor, (size_t)cap * 3))
      return 1;
    s->capK = cap;
  }
  if (tiles > s->capT) {
    if (grow(ctx, s->ranges, tiles)) return 1;
    s->capT = tiles;
  }
  if (pix > s->capPix) {
    if (grow(ctx, s->final_T, pix) || grow(ctx, s->n_contrib, pix)) return 1;
    s->capPix = pix;
  }
  return 0;
}

int ensure_temp(neuma_ctx* ctx, size_t bytes) {
  RasterState* s = ctx->raster;
  if (bytes <= s->temp_bytes) return 0;
  if (s->temp) cudaFree(s->temp);
  s->temp = nullptr;
  s->temp_bytes = 0;
  NEUMA_CUDA(ctx, cudaMalloc(&s->temp, bytes + bytes / 4));
  s->temp_bytes = bytes + bytes / 4;
  return 0;
}

int check_settings(neuma_ctx* ctx, const neuma_raster_settings* s, int K, int M) {
  if (!ctx || !s) return 2;
  if (s->image_height <= 0 || s->image_width <= 0) return neuma_fail_msg(ctx, "raster: bad image size");
  if (s->sh_degree < 0 || s->sh_degree > 3) return neuma_fail_msg(ctx, "raster: sh_degree must be 0..3");
  if (K < 0 || M < 0) return neuma_fail_msg(ctx, "raster: bad K / M");
  return 0;
}

}  // namespace

extern "C" {

int neuma_raster_fwd(neuma_ctx* ctx, const neuma_raster_settings* set, int K, int M,
                     const float* means3D, const float* shs, const float* colors_precomp,
                     const float* opacities, const float* cov6, float* image, int32_t* radii,
                     void* stream) {
  if (int r = check_settings(ctx, set, K, M)) return r;
  if ((shs == nullptr) == (colors_precomp == nullptr))
    return neuma_fail_msg(ctx, "raster: provide exactly one of shs / colors_precomp");
  if (shs && M < (set->sh_degree + 1) * (set->sh_degree + 1))
    return neuma_fail_msg(ctx, "raster: shs has fewer coefficients than sh_degree needs");
  const Cam cam = make_cam(set);
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles = cam.tiles_x * cam.tiles_y, pix = cam.W * cam.H;
  if (int r = ensure_state(ctx, K > 0 ? K : 1, tiles, pix)) return r;
  RasterState* s = ctx->raster;
  s->K = K; s->tiles_x = cam.tiles_x; s->tiles_y = cam.tiles_y; s->W = cam.W; s->Himg = cam.H;
  s->num_rendered = 0;
  NEUMA_CUDA(ctx, cudaMemsetAsync(s->ranges, 0, (size_t)tiles * sizeof(uint2), st));
  if (K > 0) {
    preprocess_kernel<<<(K + 255) / 256, 256, 0, st>>>(cam, K, M, means3D, shs, colors_precomp,
                                                       opacities, cov6, radii, s->depths, s->xy,
                                                       s->conic_o, s->rgb, s->clamped,
                                                       s->tiles_touched);
    NEUMA_LAUNCH_CHECK(ctx, "preprocess_kernel");
    size_t tb = 0;
    cub::DeviceScan::InclusiveSum(nullptr, tb, s->tiles_touched, s->offsets, K, st);
    if (int r = ensure_temp(ctx, tb)) return r;
    NEUMA_CUDA(ctx, cub::DeviceScan::InclusiveSum(s->temp, s->temp_bytes, s->tiles_touched,
                                                  s->offsets, K, st));
    ctx->launches++;
    // the instance count sizes the sort: one 4-byte D2H + stream sync, as upstream does
    NEUMA_CUDA(ctx, cudaMemcpyAsync(s->num_rendered_pinned, s->offsets + (K - 1), sizeof(int),
                                    cudaMemcpyDeviceToHost, st));
    NEUMA_CUDA(ctx, cudaStreamSynchronize(st));
    s->num_rendered = *s->num_rendered_pinned;
  }
  const int L = s->num_rendered;
  if (L > 0) {
    if ((size_t)L > s->capD) {
      const size_t cap = (size_t)L + L / 2 + 4096;
      if (grow(ctx, s->keys_a, cap) || grow(ctx, s->keys_b, cap) || grow(ctx, s->vals_a, cap) ||
          grow(ctx, s->vals_b, cap))
        return 1;
      s->capD = cap;
    }
    duplicate_kernel<<<(K + 255) / 256, 256, 0, st>>>(cam, K, s->xy, s->depths, s->offsets, radii,
                                                      s->keys_a, s->vals_a);
    NEUMA_LAUNCH_CHECK(ctx, "duplicate_kernel");
    int tile_bits = 0;
    while ((1 << tile_bits) < tiles) ++tile_bits;
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, s->keys_a, s->keys_b, s->vals_a, s->vals_b, L, 0,
                                    32 + tile_bits, st);
    if (int r = ensure_temp(ctx, tb)) return r;
    NEUMA_CUDA(ctx, cub::DeviceRadixSort::SortPairs(s->temp, s->temp_bytes, s->keys_a, s->keys_b,
                                                    s->vals_a, s->vals_b, L, 0, 32 + tile_bits, st));
    ctx->launches += 4;
    tile_ranges_kernel<<<(L + 255) / 256, 256, 0, st>>>(L, s->keys_b, s->ranges);
    NEUMA_LAUNCH_CHECK(ctx, "tile_ranges_kernel");
  }
  render_fwd_kernel<<<dim3(cam.tiles_x, cam.tiles_y), dim3(BX, BY), 0, st>>>(
      cam, s->ranges, s->vals_b, s->xy, s->rgb, s->conic_o, s->final_T, s->n_contrib, image);
  NEUMA_LAUNCH_CHECK(ctx, "render_fwd_kernel");
  return 0;
}

int neuma_raster_bwd(neuma_ctx* ctx, const neuma_raster_settings* set, int K, int M,
                     const float* means3D, const float* shs, const float* colors_precomp,
                     const float* opacities, const float* cov6, const int32_t* radii,
                     const float* dL_dimage, float* dL_dmeans3D, float* dL_dmeans2D,
                     float* dL_dcov6, float* dL_dopacity, float* dL_dcolor_or_sh, void* stream) {
  if (int r = check_settings(ctx, set, K, M)) return r;
  RasterState* s = ctx->raster;
  if (!s || s->K != K || s->W != set->image_width || s->Himg != set->image_height)
    return neuma_fail_msg(ctx, "raster_bwd: no matching forward state in this ctx");
  if (!dL_dmeans3D || !dL_dmeans2D) return neuma_fail_msg(ctx, "raster_bwd: dL_dmeans3D/2D required");
  (void)opacities;
  if (K == 0) return 0;
  const Cam cam = make_cam(set);
  cudaStream_t st = (cudaStream_t)stream;
  NEUMA_CUDA(ctx, cudaMemsetAsync(dL_dmeans2D, 0, (size_t)K * 3 * sizeof(float), st));
  NEUMA_CUDA(ctx, cudaMemsetAsync(s->dconic, 0, (size_t)K * sizeof(float4), st));
  NEUMA_CUDA(ctx, cudaMemsetAsync(s->dcolor, 0, (size_t)K * 3 * sizeof(float), st));
  if (s->num_rendered > 0) {
    render_bwd_kernel<<<dim3(cam.tiles_x, cam.tiles_y), dim3(BX, BY), 0, st>>>(
        cam, s->ranges, s->vals_b, s->xy, s->rgb, s->conic_o, s->final_T, s->n_contrib, dL_dimage,
        dL_dmeans2D, s->dconic, s->dcolor);
    NEUMA_LAUNCH_CHECK(ctx, "render_bwd_kernel");
  }
  preprocess_bwd_kernel<<<(K + 255) / 256, 256, 0, st>>>(
      cam, K, M, means3D, shs, colors_precomp, cov6, radii, s->clamped, dL_dmeans2D, s->dconic,
      s->dcolor, dL_dmeans3D, dL_dcov6, dL_dopacity, dL_dcolor_or_sh);
  NEUMA_LAUNCH_CHECK(ctx, "preprocess_bwd_kernel");
  return 0;
}

int neuma_raster_num_rendered(neuma_ctx* ctx, void* stream, int* out_host) {
  if (!ctx || !out_host) return 2;
  (void)stream;
  *out_host = ctx->raster ? ctx->raster->num_rendered : 0;
  return 0;
}

int neuma_raster_export_binning(neuma_ctx* ctx, uint64_t* keys, int32_t* gaussian_ids,
                                int32_t* tile_ranges, void* stream) {
  if (!ctx || !ctx->raster) return neuma_fail_msg(ctx, "raster_export_binning: no forward state");
  RasterState* s = ctx->raster;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t L = (size_t)s->num_rendered;
  if (keys && L)
    NEUMA_CUDA(ctx, cudaMemcpyAsync(keys, s->keys_b, L * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
  if (gaussian_ids && L)
    NEUMA_CUDA(ctx, cudaMemcpyAsync(gaussian_ids, s->vals_b, L * sizeof(uint32_t),
                                    cudaMemcpyDeviceToDevice, st));
  if (tile_ranges)
    NEUMA_CUDA(ctx, cudaMemcpyAsync(tile_ranges, s->ranges,
                                    (size_t)s->tiles_x * s->tiles_y * sizeof(uint2),
                                    cudaMemcpyDeviceToDevice, st));
  NEUMA_CUDA(ctx, cudaStreamSynchronize(st));
  return 0;
}

}  // extern "C"
