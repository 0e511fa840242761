// pitt_common.cuh — context, cloud handle and launch helpers shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pitt_b200.h"

#define PITT_SM_COUNT_DEFAULT 148

struct pitt_ctx {
  int device = 0;
  uint64_t seed = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int sm_count = PITT_SM_COUNT_DEFAULT;
  std::string err;
  double last_ms = 0.0;
  int64_t launches = 0;
  int timing_depth = 0;
  // pinned host scratch (grow-only)
  void* h_pin = nullptr;
  size_t h_pin_bytes = 0;
  // device scratch (grow-only arena, bump allocated per API call, reset at call entry)
  char* d_arena = nullptr;
  size_t d_arena_bytes = 0;
  size_t d_arena_off = 0;
  size_t d_call_total = 0;
  std::vector<void*> d_overflow;  // blocks allocated when the arena was too small (freed at reset)
  // released cloud buffers kept for reuse (frame streams stage a same-sized cloud every step)
  struct PoolBuf { void* p; size_t bytes; };
  std::vector<PoolBuf> cloud_pool;
  // helper contexts + host threads that run the independent primitive fits of a frame concurrently
  // (services.cu); created on first use, n_workers = 0 keeps everything on this ctx's stream
  struct pitt_workers* workers = nullptr;
  int n_workers = 4;
  cudaEvent_t ev_fan = nullptr;
  // pitt_sac_segment_host: the cloud travels host -> device in chunks on a second stream while the first chunks are scored
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_chunk[16] = {};
  cudaEvent_t ev_copy_gate = nullptr;
  // PITT_BLOCKING_SYNC=1: host waits sleep on an event (cudaEventBlockingSync) instead of spinning; for frame streams
  // with more host threads than cores (16 contexts x 8 ranks on one box)
  cudaEvent_t ev_block = nullptr;
  cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;  // bracket of the last plane_tc_kernel launch (pitt_debug_plane_tc_kernel_ms)
  bool time_tc_kernel = false;                   // pitt_debug_plane_tc_time_kernel(ctx, 1)
  int* d_ready = nullptr;   // 16 arrival flags (device) and the pinned word they are raised from
  int* h_one = nullptr;
  // services.cu: the primitive fits of a frame run as asynchronous chains on helper streams; pinned staging for their sample
  // tables and result blocks
  cudaStream_t fit_streams[4] = {};
  cudaEvent_t ev_fit_join[4] = {};
  cudaEvent_t ev_fit_fork = nullptr;
  void* h_stage = nullptr;
  size_t h_stage_bytes = 0;
  const int* skip_flag = nullptr;  // sac.cu: device word read by the estimate / score kernels launched while it is set (non-zero: return)
  void* mg_tables = nullptr;  // knn.cu: the two dense cell tables of the multi-level grid (allocated on first use)
  int* knn_scr = nullptr;     // knn.cu: scratch of the last large-cloud k-NN (diagnostics, arena memory)
  // services.cu, pitt_segment_frame: the whole-cloud normals (which findSupports never reads) run on this stream beside the supports loop
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t ev_aux_fork = nullptr, ev_aux_join = nullptr;
  // services.cu, batched frame streams: the next frame's host -> device copy runs on this stream while the current frame is segmented
  cudaStream_t h2d_stream = nullptr;
  cudaEvent_t ev_h2d[2] = {};
  void* h_pin2 = nullptr;  // pinned block for the gathered sample points (h_pin holds the sample indices at that time)
  size_t h_pin2_bytes = 0;
};

struct pitt_cloud {
  int n = 0;
  float4* d_xyz = nullptr;
  float4* d_nrm = nullptr;
  bool has_normals = false;
  std::vector<float> h_xyz;  // lazy host mirror (n*4)
  bool h_valid = false;
  // streaming stage (pitt_sac_segment_host only): chunk k is on the device once
  // ctx->ev_chunk[k] has fired; h_src = the caller's buffer (point_step 16), valid for the duration of the fused call
  mutable int stream_chunks = 0;
  int stream_off[17] = {};  // chunk k = points [stream_off[k], stream_off[k + 1])
  const float* h_src = nullptr;
  const int* d_ready = nullptr;  // per-chunk arrival flags for the single-launch tensor path (equal chunks of stream_off[1] points)
};

namespace pitt {

// every host wait on the context's stream goes through here
inline cudaError_t stream_sync(pitt_ctx* ctx) {
  if (!ctx->ev_block) return cudaStreamSynchronize(ctx->stream);
  cudaError_t e = cudaEventRecord(ctx->ev_block, ctx->stream);
  if (e != cudaSuccess) return e;
  return cudaEventSynchronize(ctx->ev_block);
}

inline int fail(pitt_ctx* ctx, int code, const char* what, cudaError_t e = cudaSuccess) {
  if (ctx) {
    ctx->err = what;
    if (e != cudaSuccess) {
      ctx->err += ": ";
      ctx->err += cudaGetErrorString(e);
    }
  }
  return code;
}

#define PITT_CUDA(ctx, call)                                                  \
  do {                                                                        \
    cudaError_t e__ = (call);                                                 \
    if (e__ != cudaSuccess) return pitt::fail((ctx), PITT_ERR_CUDA, #call, e__); \
  } while (0)

#define PITT_TRY(expr)              \
  do {                              \
    int s__ = (expr);               \
    if (s__ != PITT_OK) return s__; \
  } while (0)

// -------- device arena: one cudaMalloc that grows; bump pointer reset at the start of an API call.
inline void arena_reset(pitt_ctx* ctx) {
  // blocks that did not fit last time are released and the arena grown to last call's total
  if (!ctx->d_overflow.empty() || ctx->d_call_total > ctx->d_arena_bytes) {
    pitt::stream_sync(ctx);
    for (void* p : ctx->d_overflow) cudaFree(p);
    ctx->d_overflow.clear();
    if (ctx->d_call_total > ctx->d_arena_bytes) {
      if (ctx->d_arena) cudaFree(ctx->d_arena);
      ctx->d_arena = nullptr;
      ctx->d_arena_bytes = 0;
      size_t want = ctx->d_call_total + ctx->d_call_total / 4 + (1 << 20);
      if (cudaMalloc((void**)&ctx->d_arena, want) == cudaSuccess) ctx->d_arena_bytes = want;
      else cudaGetLastError();
    }
  }
  ctx->d_arena_off = 0;
  ctx->d_call_total = 0;
}
template <typename T>
inline int arena_alloc(pitt_ctx* ctx, size_t count, T** out) {
  size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
  if (bytes == 0) bytes = 256;
  ctx->d_call_total += bytes;
  if (ctx->d_arena_off + bytes <= ctx->d_arena_bytes) {
    *out = (T*)(ctx->d_arena + ctx->d_arena_off);
    ctx->d_arena_off += bytes;
    return PITT_OK;
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return fail(ctx, PITT_ERR_CUDA, "cudaMalloc(overflow)", e);
  ctx->d_overflow.push_back(p);
  *out = (T*)p;
  return PITT_OK;
}
inline int pinned_reserve(pitt_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->h_pin_bytes) return PITT_OK;
  pitt::stream_sync(ctx);
  if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
  ctx->h_pin = nullptr;
  ctx->h_pin_bytes = 0;
  size_t want = bytes + bytes / 4 + 4096;
  cudaError_t e = cudaMallocHost(&ctx->h_pin, want);
  if (e != cudaSuccess) return fail(ctx, PITT_ERR_CUDA, "cudaMallocHost", e);
  ctx->h_pin_bytes = want;
  return PITT_OK;
}

inline int pinned2_reserve(pitt_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->h_pin2_bytes) return PITT_OK;
  pitt::stream_sync(ctx);
  if (ctx->h_pin2) cudaFreeHost(ctx->h_pin2);
  ctx->h_pin2 = nullptr;
  ctx->h_pin2_bytes = 0;
  size_t want = bytes + bytes / 4 + 4096;
  cudaError_t e = cudaMallocHost(&ctx->h_pin2, want);
  if (e != cudaSuccess) return fail(ctx, PITT_ERR_CUDA, "cudaMallocHost", e);
  ctx->h_pin2_bytes = want;
  return PITT_OK;
}

// RAII bracket for the per-call device timing (outermost API call only).
struct CallTimer {
  pitt_ctx* ctx;
  bool outer;
  explicit CallTimer(pitt_ctx* c) : ctx(c) {
    outer = (ctx->timing_depth++ == 0);
    if (outer) {
      arena_reset(ctx);
      cudaEventRecord(ctx->ev0, ctx->stream);
    }
  }
  // call after the stream has been synchronised by the API function
  void finish() {
    if (outer) {
      cudaEventRecord(ctx->ev1, ctx->stream);
      cudaEventSynchronize(ctx->ev1);
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
      ctx->last_ms = ms;
    }
  }
  ~CallTimer() { ctx->timing_depth--; }
};

inline int ensure_host_mirror(pitt_ctx* ctx, const pitt_cloud* cc) {
  pitt_cloud* c = const_cast<pitt_cloud*>(cc);
  if (c->h_valid) return PITT_OK;
  c->h_xyz.resize((size_t)c->n * 4);
  if (c->n > 0) {
    PITT_CUDA(ctx, cudaMemcpyAsync(c->h_xyz.data(), c->d_xyz, (size_t)c->n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    PITT_CUDA(ctx, pitt::stream_sync(ctx));
  }
  c->h_valid = true;
  return PITT_OK;
}

// device buffers of clouds come from a small per-context pool: cudaMalloc/cudaFree synchronise the
// device and cost ~100 us each, which would dominate a 1 ms frame step
inline int pool_alloc(pitt_ctx* ctx, size_t bytes, void** out) {
  size_t best = (size_t)-1;
  int bi = -1;
  for (size_t i = 0; i < ctx->cloud_pool.size(); ++i)
    if (ctx->cloud_pool[i].bytes >= bytes && ctx->cloud_pool[i].bytes < best) { best = ctx->cloud_pool[i].bytes; bi = (int)i; }
  if (bi >= 0 && best <= bytes * 2 + 4096) {
    *out = ctx->cloud_pool[bi].p;
    ctx->cloud_pool.erase(ctx->cloud_pool.begin() + bi);
    return PITT_OK;
  }
  cudaError_t e = cudaMalloc(out, bytes ? bytes : 16);
  if (e != cudaSuccess) return fail(ctx, PITT_ERR_CUDA, "cudaMalloc(cloud)", e);
  return PITT_OK;
}
inline size_t pool_bytes_of(size_t n_points) { return n_points * 16; }
inline void pool_free(pitt_ctx* ctx, void* p, size_t bytes) {
  if (!p) return;
  if (ctx && ctx->cloud_pool.size() < 8) ctx->cloud_pool.push_back({p, bytes});
  else cudaFree(p);
}

// services.cu: stops the worker threads and destroys the helper contexts
void workers_destroy(pitt_ctx* ctx);

// PITT_TRACE=1: wall-clock phase log on stderr (development aid; one getenv per process)
struct TraceScope {
  const char* name;
  pitt_ctx* ctx;
  double t0;
  static bool enabled() {
    static int e = -1;
    if (e < 0) { const char* v = getenv("PITT_TRACE"); e = (v && v[0] == '1') ? 1 : 0; }
    return e == 1;
  }
  static double now() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
  }
  TraceScope(pitt_ctx* c, const char* n) : name(n), ctx(c), t0(0) {
    if (enabled()) t0 = now();
  }
  ~TraceScope() {
    if (enabled()) {
      pitt::stream_sync(ctx);
      fprintf(stderr, "[pitt trace] %-28s %8.3f ms\n", name, now() - t0);
    }
  }
};

// PITT_TRACE=2: host time stamps without synchronisation (where the host thread is when)
inline void trace_mark(const char* what) {
  static int e = -1;
  static double t_prev = 0.0;
  if (e < 0) { const char* v = getenv("PITT_TRACE"); e = (v && v[0] == '2') ? 1 : 0; }
  if (e != 1) return;
  const double t = TraceScope::now();
  fprintf(stderr, "[pitt mark] %-40s +%8.3f ms\n", what, t_prev == 0.0 ? 0.0 : t - t_prev);
  t_prev = t;
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace pitt
