// Context, error reporting and device vectors of the C ABI (include/geosgpu.h).
#include "gg_common.cuh"

namespace gg {
static thread_local std::string g_last_error;
void set_last_error(const std::string& m) { g_last_error = m; }

__global__ void k_fill(int64_t n, double v, double* __restrict__ x) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = v;
}
}  // namespace gg

using namespace gg;

extern "C" {

int gg_init(int device_id, gg_ctx** out) {
  GG_API_BEGIN
  GG_REQUIRE(out, GG_ERR_INVALID, "null output");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    throw Error(GG_ERR_CUDA, std::string("no usable CUDA device (this library has no CPU fallback): ") +
                                 (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
  GG_REQUIRE(device_id >= 0 && device_id < ndev, GG_ERR_INVALID, "device_id out of range");
  GG_CUDA(cudaSetDevice(device_id));
  cudaDeviceProp prop;
  GG_CUDA(cudaGetDeviceProperties(&prop, device_id));
  GG_REQUIRE(prop.major == 10, GG_ERR_UNSUPPORTED,
             std::string("libgeosgpu is built for sm_100a (B200) only; found ") + prop.name + " (sm_" +
                 std::to_string(prop.major) + std::to_string(prop.minor) + ")");
  std::unique_ptr<gg_ctx> ctx(new gg_ctx);
  ctx->device = device_id;
  ctx->sm_count = prop.multiProcessorCount;
  GG_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  *out = ctx.release();
  GG_API_END
}

void gg_finalize(gg_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) {
    cudaStreamSynchronize(ctx->stream);
  }
  ctx->tabs.clear();
  ctx->ebuf.release();
  ctx->vbuf.release();
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* gg_last_error(gg_ctx* ctx) {
  (void)ctx;
  return g_last_error.c_str();
}

void* gg_stream(gg_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int gg_synchronize(gg_ctx* ctx) {
  GG_API_BEGIN
  GG_REQUIRE(ctx, GG_ERR_INVALID, "null context");
  GG_CUDA(cudaSetDevice(ctx->device));
  GG_CUDA(cudaStreamSynchronize(ctx->stream));
  GG_API_END
}

int64_t gg_launch_count(gg_ctx* ctx) { return ctx ? ctx->launches : 0; }

int gg_vec_create(gg_ctx* ctx, int64_t n, gg_vec** out) {
  GG_API_BEGIN
  GG_REQUIRE(ctx && out && n >= 0, GG_ERR_INVALID, "bad argument");
  GG_CUDA(cudaSetDevice(ctx->device));
  std::unique_ptr<gg_vec> v(new gg_vec);
  v->ctx = ctx;
  v->d.alloc(n);
  if (n) GG_CUDA(cudaMemsetAsync(v->d.p, 0, n * sizeof(double), ctx->stream));
  GG_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = v.release();
  GG_API_END
}

int gg_vec_set(gg_vec* v, const double* host) {
  GG_API_BEGIN
  GG_REQUIRE(v && (host || v->d.n == 0), GG_ERR_INVALID, "null argument");
  GG_CUDA(cudaSetDevice(v->ctx->device));
  if (v->d.n) GG_CUDA(cudaMemcpyAsync(v->d.p, host, v->d.n * sizeof(double), cudaMemcpyHostToDevice, v->ctx->stream));
  GG_CUDA(cudaStreamSynchronize(v->ctx->stream));
  GG_API_END
}

int gg_vec_get(gg_vec* v, double* host) {
  GG_API_BEGIN
  GG_REQUIRE(v && (host || v->d.n == 0), GG_ERR_INVALID, "null argument");
  GG_CUDA(cudaSetDevice(v->ctx->device));
  v->d.download(host, v->ctx->stream);
  GG_API_END
}

int gg_vec_fill(gg_vec* v, double value) {
  GG_API_BEGIN
  GG_REQUIRE(v, GG_ERR_INVALID, "null vector");
  GG_CUDA(cudaSetDevice(v->ctx->device));
  if (v->d.n) {
    k_fill<<<(unsigned)((v->d.n + 255) / 256), 256, 0, v->ctx->stream>>>((int64_t)v->d.n, value, v->d.p);
    v->ctx->launches++;
    GG_CUDA(cudaGetLastError());
  }
  GG_CUDA(cudaStreamSynchronize(v->ctx->stream));
  GG_API_END
}

int64_t gg_vec_size(gg_vec* v) { return v ? (int64_t)v->d.n : -1; }
double* gg_vec_device_ptr(gg_vec* v) { return v ? v->d.p : nullptr; }
void gg_vec_destroy(gg_vec* v) { delete v; }

}  // extern "C"
