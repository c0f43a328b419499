// Context, error reporting and device vectors of the C ABI (include/geosgpu.h).
#include <mutex>
#include "gg_common.cuh"

namespace gg {
static thread_local std::string g_last_error;
void set_last_error(const std::string& m) { g_last_error = m; }

// FP64 FMA throughput probe: 8 independent dependent-chains per thread, 2 flops per FMA
__global__ void __launch_bounds__(256) k_fp64_peak(int iters, double seed, double* __restrict__ out) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 12345.678) out[0] = s;  // never true: keeps the chains alive
}

__global__ void k_fill(int64_t n, double v, double* __restrict__ x) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = v;
}
}  // namespace gg

using namespace gg;

extern "C" {

// live contexts per device: the __constant__ tables the contexts key (const_order, swe_key, ebe_key, ...) are per device and
// per process, so a second context on the same device would run with the first one's tables
static std::mutex g_live_mutex;
static int g_live_ctx[64] = {0};

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
  {
    std::lock_guard<std::mutex> lock(g_live_mutex);
    GG_REQUIRE(device_id < 64 && g_live_ctx[device_id] == 0, GG_ERR_INVALID,
               "a context already exists on this device in this process (one gg_ctx per process and GPU: the resident "
               "__constant__ tables are keyed per context); gg_finalize it first");
    g_live_ctx[device_id] = 1;
  }
  struct Unclaim { int dev; bool armed = true; ~Unclaim() { if (armed) { std::lock_guard<std::mutex> l(g_live_mutex); g_live_ctx[dev] = 0; } } } unclaim{device_id};
  std::unique_ptr<gg_ctx> ctx(new gg_ctx);
  ctx->device = device_id;
  ctx->sm_count = prop.multiProcessorCount;
  GG_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  {
    // the fill stream outranks the main stream: its CTAs must get onto the SMs while the element kernel still has CTAs queued
    int lo = 0, hi = 0;
    GG_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    GG_CUDA(cudaStreamCreateWithPriority(&ctx->stream2, cudaStreamNonBlocking, hi));
  }
  for (auto& e : ctx->ev_batch) GG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  ctx->geometry_classes = getenv("GG_LB_CLASSES") && getenv("GG_LB_CLASSES")[0] == '1';
  GG_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
  unclaim.armed = false;
  *out = ctx.release();
  GG_API_END
}

void gg_finalize(gg_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  { std::lock_guard<std::mutex> lock(g_live_mutex); if (ctx->device >= 0 && ctx->device < 64) g_live_ctx[ctx->device] = 0; }
  if (ctx->stream) {
    cudaStreamSynchronize(ctx->stream);
  }
  ctx->tabs.clear();
  ctx->ebuf.release();
  ctx->vbuf.release();
  if (ctx->stream2) { cudaStreamSynchronize(ctx->stream2); cudaStreamDestroy(ctx->stream2); }
  for (auto& e : ctx->ev_batch) if (e) cudaEventDestroy(e);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
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

int gg_set_async(gg_ctx* ctx, int enabled) {
  GG_API_BEGIN
  GG_REQUIRE(ctx, GG_ERR_INVALID, "null context");
  ctx->async = enabled != 0;
  GG_API_END
}

int gg_set_geometry_classes(gg_ctx* ctx, int enabled) {
  GG_API_BEGIN
  GG_REQUIRE(ctx, GG_ERR_INVALID, "null context");
  ctx->geometry_classes = enabled != 0;
  GG_API_END
}

int gg_profile_enable(gg_ctx* ctx, int enabled) {
  GG_API_BEGIN
  GG_REQUIRE(ctx, GG_ERR_INVALID, "null context");
  GG_CUDA(cudaSetDevice(ctx->device));
  GG_CUDA(cudaStreamSynchronize(ctx->stream));
  for (auto& e : ctx->prof) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  ctx->prof.clear();
  ctx->profile = enabled != 0;
  GG_API_END
}

int gg_profile_query(gg_ctx* ctx, const char* kernel_name, int64_t* count, double* total_ms) {
  GG_API_BEGIN
  GG_REQUIRE(ctx && kernel_name && count && total_ms, GG_ERR_INVALID, "null argument");
  GG_CUDA(cudaSetDevice(ctx->device));
  GG_CUDA(cudaStreamSynchronize(ctx->stream));
  *count = 0; *total_ms = 0.0;
  for (auto& e : ctx->prof)
    if (e.name == kernel_name) {
      float ms = 0.f;
      GG_CUDA(cudaEventElapsedTime(&ms, e.a, e.b));
      *total_ms += ms; ++*count;
    }
  GG_API_END
}

int gg_measure_fp64_peak(gg_ctx* ctx, double* tflops) {
  GG_API_BEGIN
  GG_REQUIRE(ctx && tflops, GG_ERR_INVALID, "null argument");
  GG_CUDA(cudaSetDevice(ctx->device));
  DevBuf<double> out(1);
  const int iters = 4096, threads = 256, blocks = ctx->sm_count * 8;
  cudaEvent_t a, b;
  GG_CUDA(cudaEventCreate(&a)); GG_CUDA(cudaEventCreate(&b));
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    GG_CUDA(cudaEventRecord(a, ctx->stream));
    k_fp64_peak<<<blocks, threads, 0, ctx->stream>>>(iters, 1.0 + rep, out.p);
    GG_CUDA(cudaEventRecord(b, ctx->stream));
    GG_CUDA(cudaEventSynchronize(b));
    float ms = 0.f;
    GG_CUDA(cudaEventElapsedTime(&ms, a, b));
    double fl = 2.0 * 64.0 * iters * (double)threads * blocks;
    if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
    ctx->launches++;
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  *tflops = best;
  GG_API_END
}

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
