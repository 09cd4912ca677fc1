// Shared by raster_pre.cu (compiled with -fmad=false so that depth bits, radii and tile
// rectangles are bit-reproducible against the fp32 numpy oracle) and raster.cu.
#pragma once
#include "common.cuh"

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
  float4* dconic = nullptr;  // backward: dL/d(A,B,C) of the conic, .w = dL/dopacity
  float* dcolor = nullptr;   // backward: dL/drgb [K,3]
  // per (Gaussian, tile) instance.  Binning is a two-level bucket sort, no global sort: count the
  // instances of every tile, scan the counts, scatter every instance into its tile's bucket, then
  // ONE CTA per tile sorts its bucket by (depth bits << 32 | Gaussian id) in shared memory -- the
  // same total order as upstream's stable radix sort of (tile << 32 | depth) keys.
  size_t capD = 0;
  uint64_t* keys = nullptr;   // [capD] depth bits << 32 | Gaussian id, bucketed by tile, sorted in place
  uint32_t* plist = nullptr;  // [capD] Gaussian ids in blend order
  // per tile / pixel
  int capT = 0;
  uint32_t* tile_count = nullptr;   // [tiles]
  uint32_t* tile_start = nullptr;   // [tiles + 1] exclusive scan of the counts
  uint32_t* tile_cursor = nullptr;  // [tiles]
  uint2* ranges = nullptr;          // [tiles] (start, end) of the tile's instances; (0, 0) if none
  int capPix = 0;
  float* final_T = nullptr;
  uint32_t* n_contrib = nullptr;
  // instance count of the last forward: on the device ([0] = count, [1] = 1 if it exceeded capD)
  // and, copied asynchronously after every forward, in pinned host memory
  int32_t* d_num = nullptr;
  int32_t* h_num = nullptr;
  cudaEvent_t num_ready = nullptr;  // recorded after the async copy into h_num
  bool num_pending = false;
  // last forward
  int num_rendered = -1;            // -1: not read back yet (neuma_raster_num_rendered synchronises)
  int K = 0, tiles_x = 0, tiles_y = 0, W = 0, Himg = 0;
};


struct Cam {
  float view[16], proj[16];
  float cam[3], bg[3];
  float tanx, tany, fx, fy;
  int W, H, tiles_x, tiles_y, deg;
};

namespace {


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
  // explicit fp32 evaluation order (mirrored by oracle/raster_oracle.py::bin_np)
  const float rf = (float)radius;
  x0 = min(cam.tiles_x, max(0, (int)((px - rf) / (float)BX)));
  y0 = min(cam.tiles_y, max(0, (int)((py - rf) / (float)BY)));
  x1 = min(cam.tiles_x, max(0, (int)(((px + rf) + (float)(BX - 1)) / (float)BX)));
  y1 = min(cam.tiles_y, max(0, (int)(((py + rf) + (float)(BY - 1)) / (float)BY)));
}


}  // namespace

// raster_pre.cu
void neuma_launch_preprocess(const Cam& cam, int K, int M, const float* means, const float* shs,
                             const float* colors_precomp, const float* opac, const float* cov6,
                             int* radii, RasterState* s, cudaStream_t st);
void neuma_launch_bucket_scatter(const Cam& cam, int K, const int* radii, RasterState* s, cudaStream_t st);
