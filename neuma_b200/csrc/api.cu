// Context management for libneuma_b200 (C ABI, include/neuma_b200.h).
#include "common.cuh"
#include <stdlib.h>

void neuma_raster_state_free(RasterState* s);  // raster.cu

int neuma_ensure_grid(neuma_ctx* ctx, size_t nodes) {
  if (nodes <= ctx->grid_nodes) return 0;
  if (ctx->grid_mvm) cudaFree(ctx->grid_mvm);
  if (ctx->grid_v) cudaFree(ctx->grid_v);
  if (ctx->grid_adj) cudaFree(ctx->grid_adj);
  ctx->grid_mvm = ctx->grid_v = ctx->grid_adj = nullptr;
  ctx->grid_nodes = 0;
  NEUMA_CUDA(ctx, cudaMalloc(&ctx->grid_mvm, nodes * sizeof(float4)));
  NEUMA_CUDA(ctx, cudaMalloc(&ctx->grid_v, nodes * sizeof(float4)));
  NEUMA_CUDA(ctx, cudaMalloc(&ctx->grid_adj, nodes * sizeof(float4)));
  ctx->grid_nodes = nodes;
  return 0;
}

int neuma_ensure_scratch(neuma_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->scratch_bytes) return 0;
  if (ctx->scratch) cudaFree(ctx->scratch);
  ctx->scratch = nullptr;
  ctx->scratch_bytes = 0;
  NEUMA_CUDA(ctx, cudaMalloc(&ctx->scratch, bytes));
  ctx->scratch_bytes = bytes;
  return 0;
}

bool neuma_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("NEUMA_PDL");
    v = (e && e[0] == '0') ? 0 : 1;  // on by default since the constitutive kernels stopped triggering early
  }                                  // (profiles/r03y_pdl_ab.txt: +2.1 % frames/s); NEUMA_PDL=0 turns it off
  return v != 0;
}

int neuma_ensure_wstage(neuma_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->wstage_bytes) return 0;
  if (ctx->wstage) cudaFree(ctx->wstage);
  ctx->wstage = nullptr;
  ctx->wstage_bytes = 0;
  NEUMA_CUDA(ctx, cudaMalloc(&ctx->wstage, bytes));
  ctx->wstage_bytes = bytes;
  return 0;
}

extern "C" {

int neuma_version(void) { return 100; }

int neuma_ctx_create(int device, neuma_ctx** out) {
  if (!out) return 2;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return 3;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 3;
  if (prop.major != 10) return 4;  // sm_100a only: there is no fallback path
  // the caller's current device is left as it was (entry points run on the device of the
  // pointers / stream they are given; the Python front end makes ctx->device current per call)
  neuma_ctx* ctx = (neuma_ctx*)calloc(1, sizeof(neuma_ctx));
  if (!ctx) return 5;
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  *out = ctx;
  return 0;
}

int neuma_ctx_destroy(neuma_ctx* ctx) {
  if (!ctx) return 0;
  cudaSetDevice(ctx->device);
  if (ctx->grid_mvm) cudaFree(ctx->grid_mvm);
  if (ctx->grid_v) cudaFree(ctx->grid_v);
  if (ctx->grid_adj) cudaFree(ctx->grid_adj);
  if (ctx->scratch) cudaFree(ctx->scratch);
  if (ctx->wstage) cudaFree(ctx->wstage);
  if (ctx->raster) neuma_raster_state_free(ctx->raster);
  free(ctx);
  return 0;
}

const char* neuma_last_error(neuma_ctx* ctx) { return ctx ? ctx->err : "null ctx"; }

long long neuma_launch_count(neuma_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
