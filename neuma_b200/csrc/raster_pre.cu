// Per-Gaussian preprocess + instance bucketing of the rasterizer forward.
// Compiled with -fmad=false: every fp32 operation below rounds separately, in the
// order written, exactly like oracle/raster_oracle.py::bin_np (numpy fp32), so the
// depth bits, radii, tile rectangles -- and hence the per-tile blend order, equal to
// upstream's sorted (tile << 32 | depth) instance keys -- are bit-exact against the oracle.
#include "raster_common.cuh"

namespace {

// ---------------------------------------------------------------- preprocess
// Per-tile instance counts are first accumulated in a per-CTA shared-memory histogram (`tiles`
// bins) and flushed with one global atomic per touched (CTA, tile): ~80 k global atomics per frame
// at the bench workload instead of one per instance (514 k, serialised on ~400 hot counters).
// smem_tiles == 0: the histogram does not fit shared memory, global atomics per instance.
__device__ __forceinline__ bool preprocess_one(const Cam& cam, int i, int M, const float* __restrict__ means,
                                               const float* __restrict__ shs,
                                               const float* __restrict__ colors_precomp,
                                               const float* __restrict__ opac, const float* __restrict__ cov6,
                                               int* __restrict__ radii, float* __restrict__ depths,
                                               float2* __restrict__ xy, float4* __restrict__ conic_o,
                                               float* __restrict__ rgb, unsigned char* __restrict__ clamped,
                                               int& x0, int& y0, int& x1, int& y1) {
  radii[i] = 0;
  const float p[3] = {means[3 * i], means[3 * i + 1], means[3 * i + 2]};
  float pv[3];
  xform43(cam.view, p, pv);
  if (pv[2] <= 0.2f) return false;  // near-plane cull
  float ph[4];
  xform44(cam.proj, p, ph);
  const float pw = 1.0f / (ph[3] + 0.0000001f);
  const float pp[2] = {ph[0] * pw, ph[1] * pw};
  const Cov2D cv = cov2d(cam, p, cov6 + (size_t)i * 6);
  const float det = cv.a * cv.c - cv.b * cv.b;
  if (det == 0.0f) return false;
  const float di = 1.0f / det;
  const float mid = 0.5f * (cv.a + cv.c);
  const float l1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
  const float l2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
  const int radius = (int)ceilf(3.0f * sqrtf(fmaxf(l1, l2)));
  const float px = ndc2pix(pp[0], cam.W), py = ndc2pix(pp[1], cam.H);
  tile_rect(cam, px, py, radius, x0, y0, x1, y1);
  if ((x1 - x0) * (y1 - y0) == 0) return false;
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
  return true;
}

__global__ void __launch_bounds__(256)
preprocess_kernel(Cam cam, int K, int M, const float* __restrict__ means,
                  const float* __restrict__ shs, const float* __restrict__ colors_precomp,
                  const float* __restrict__ opac, const float* __restrict__ cov6,
                  int* __restrict__ radii, float* __restrict__ depths, float2* __restrict__ xy,
                  float4* __restrict__ conic_o, float* __restrict__ rgb,
                  unsigned char* __restrict__ clamped, uint32_t* __restrict__ tile_count, int smem_tiles) {
  extern __shared__ uint32_t s_cnt[];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  for (int t = threadIdx.x; t < smem_tiles; t += blockDim.x) s_cnt[t] = 0;
  if (smem_tiles) __syncthreads();
  int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
  const bool live = i < K && preprocess_one(cam, i, M, means, shs, colors_precomp, opac, cov6, radii, depths, xy,
                                            conic_o, rgb, clamped, x0, y0, x1, y1);
  if (live) {
    // first level of the bucket sort: how many instances every tile gets
    uint32_t* cnt = smem_tiles ? s_cnt : tile_count;
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x) atomicAdd(&cnt[y * cam.tiles_x + x], 1u);
  }
  if (smem_tiles) {
    __syncthreads();
    for (int t = threadIdx.x; t < smem_tiles; t += blockDim.x) {
      const uint32_t c = s_cnt[t];
      if (c) atomicAdd(&tile_count[t], c);
    }
  }
}

// second level: every instance goes into its tile's bucket (any order inside the bucket; the
// per-tile sort fixes it).  A CTA counts its instances per tile in shared memory, reserves one
// contiguous range per touched tile with ONE global atomic, then places its instances inside the
// ranges with shared-memory atomics.  Instances beyond the capacity are dropped and flagged.
__global__ void __launch_bounds__(256)
bucket_scatter_kernel(Cam cam, int K, const float2* __restrict__ xy, const float* __restrict__ depths,
                      const int* __restrict__ radii, const uint32_t* __restrict__ tile_start,
                      uint32_t* __restrict__ tile_cursor, uint64_t* __restrict__ keys, uint32_t cap,
                      int32_t* __restrict__ d_num, int smem_tiles) {
  extern __shared__ uint32_t s_mem[];
  uint32_t* s_cnt = s_mem;
  uint32_t* s_base = s_mem + smem_tiles;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < K && radii[i] > 0;
  int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
  uint64_t key = 0;
  if (live) {
    tile_rect(cam, xy[i].x, xy[i].y, radii[i], x0, y0, x1, y1);
    key = ((uint64_t)__float_as_uint(depths[i]) << 32) | (uint32_t)i;
  }
  if (!smem_tiles) {  // histogram does not fit shared memory: one global atomic per instance
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x) {
        const int t = y * cam.tiles_x + x;
        const uint32_t pos = tile_start[t] + atomicAdd(&tile_cursor[t], 1u);
        if (pos < cap) keys[pos] = key;
        else d_num[1] = 1;
      }
    return;
  }
  for (int t = threadIdx.x; t < smem_tiles; t += blockDim.x) s_cnt[t] = 0;
  __syncthreads();
  for (int y = y0; y < y1; ++y)
    for (int x = x0; x < x1; ++x) atomicAdd(&s_cnt[y * cam.tiles_x + x], 1u);
  __syncthreads();
  for (int t = threadIdx.x; t < smem_tiles; t += blockDim.x) {
    const uint32_t c = s_cnt[t];
    if (c) {
      s_base[t] = tile_start[t] + atomicAdd(&tile_cursor[t], c);
      s_cnt[t] = 0;
    }
  }
  __syncthreads();
  for (int y = y0; y < y1; ++y)
    for (int x = x0; x < x1; ++x) {
      const int t = y * cam.tiles_x + x;
      const uint32_t pos = s_base[t] + atomicAdd(&s_cnt[t], 1u);
      if (pos < cap) keys[pos] = key;
      else d_num[1] = 1;
    }
}

}  // namespace

// shared-memory histograms of `tiles` bins: up to 96 KB (2 arrays in the scatter), else global atomics
static int smem_tiles_for(const Cam& cam, int arrays) {
  const int tiles = cam.tiles_x * cam.tiles_y;
  return (size_t)tiles * 4 * arrays <= 96 * 1024 ? tiles : 0;
}

void neuma_launch_preprocess(const Cam& cam, int K, int M, const float* means, const float* shs,
                             const float* colors_precomp, const float* opac, const float* cov6,
                             int* radii, RasterState* s, cudaStream_t st) {
  const int sm = smem_tiles_for(cam, 1);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(preprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(bucket_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr = true;
  }
  preprocess_kernel<<<(K + 255) / 256, 256, (size_t)sm * 4, st>>>(cam, K, M, means, shs, colors_precomp, opac,
                                                                  cov6, radii, s->depths, s->xy, s->conic_o,
                                                                  s->rgb, s->clamped, s->tile_count, sm);
}

void neuma_launch_bucket_scatter(const Cam& cam, int K, const int* radii, RasterState* s, cudaStream_t st) {
  const int sm = smem_tiles_for(cam, 2);
  bucket_scatter_kernel<<<(K + 255) / 256, 256, (size_t)sm * 8, st>>>(cam, K, s->xy, s->depths, radii,
                                                                      s->tile_start, s->tile_cursor, s->keys,
                                                                      (uint32_t)s->capD, s->d_num, sm);
}
