// pitt_api.cu — the C ABI declared in include/pitt_b200.h: lifecycle, staging (K0), defaults and
// the seg.segment()-shaped entry points. Service-shaped entry points live in services.cu.
#include <cfloat>
#include <cmath>

#include "pitt_common.cuh"
#include "sac.cuh"
#include "../../include/pitt_b200_debug.h"

using namespace pitt;

namespace pitt {
// K0: AoS records with arbitrary stride -> float4 {x,y,z,1}
__global__ void pack_xyz_kernel(const unsigned char* __restrict__ src, int stride, int n, float4* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = reinterpret_cast<const float*>(src + (size_t)i * stride);
  dst[i] = make_float4(p[0], p[1], p[2], 1.0f);
}
__global__ void pack_normals_kernel(const unsigned char* __restrict__ src, int stride, int curv_off, int n,
                                    float4* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = reinterpret_cast<const float*>(src + (size_t)i * stride);
  dst[i] = make_float4(p[0], p[1], p[2], p[curv_off]);
}
}  // namespace pitt

extern "C" {

const char* pitt_version(void) { return "pitt_b200 0.1 (sm_100a)"; }

int pitt_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

static pitt_ctx* create_impl(int device, uint64_t seed, void* stream, bool own) {
  int n = pitt_device_count();
  if (n <= 0 || device < 0 || device >= n) return nullptr;  // no CPU fallback
  if (cudaSetDevice(device) != cudaSuccess) return nullptr;
  pitt_ctx* ctx = new pitt_ctx();
  ctx->device = device;
  ctx->seed = seed;
  ctx->own_stream = own;
  if (own) {
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return nullptr; }
  } else {
    ctx->stream = (cudaStream_t)stream;
  }
  cudaEventCreate(&ctx->ev0);
  cudaEventCreate(&ctx->ev1);
  {
    const char* v = getenv("PITT_BLOCKING_SYNC");
    if (v && v[0] == '1') cudaEventCreateWithFlags(&ctx->ev_block, cudaEventBlockingSync | cudaEventDisableTiming);
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
  return ctx;
}

pitt_ctx* pitt_create(int device, uint64_t seed) { return create_impl(device, seed, nullptr, true); }
pitt_ctx* pitt_create_on_stream(int device, uint64_t seed, void* cuda_stream) {
  return create_impl(device, seed, cuda_stream, false);
}

void pitt_destroy(pitt_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  workers_destroy(ctx);
  if (ctx->ev_fan) cudaEventDestroy(ctx->ev_fan);
  pitt::stream_sync(ctx);
  for (void* p : ctx->d_overflow) cudaFree(p);
  for (auto& b : ctx->cloud_pool) cudaFree(b.p);
  if (ctx->d_arena) cudaFree(ctx->d_arena);
  if (ctx->mg_tables) cudaFree(ctx->mg_tables);
  for (int i = 0; i < 4; ++i) {
    if (ctx->fit_streams[i]) { cudaStreamSynchronize(ctx->fit_streams[i]); cudaStreamDestroy(ctx->fit_streams[i]); }
    if (ctx->ev_fit_join[i]) cudaEventDestroy(ctx->ev_fit_join[i]);
  }
  if (ctx->ev_fit_fork) cudaEventDestroy(ctx->ev_fit_fork);
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
  if (ctx->h_pin2) cudaFreeHost(ctx->h_pin2);
  if (ctx->h_one) cudaFreeHost(ctx->h_one);
  if (ctx->d_ready) cudaFree(ctx->d_ready);
  for (cudaEvent_t e : ctx->ev_chunk) if (e) cudaEventDestroy(e);
  if (ctx->ev_copy_gate) cudaEventDestroy(ctx->ev_copy_gate);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->aux_stream) { cudaStreamSynchronize(ctx->aux_stream); cudaStreamDestroy(ctx->aux_stream); }
  if (ctx->ev_aux_fork) cudaEventDestroy(ctx->ev_aux_fork);
  if (ctx->ev_aux_join) cudaEventDestroy(ctx->ev_aux_join);
  if (ctx->h2d_stream) { cudaStreamSynchronize(ctx->h2d_stream); cudaStreamDestroy(ctx->h2d_stream); }
  for (cudaEvent_t e : ctx->ev_h2d) if (e) cudaEventDestroy(e);
  if (ctx->ev_block) cudaEventDestroy(ctx->ev_block);
  if (ctx->ev_k0) { cudaEventDestroy(ctx->ev_k0); cudaEventDestroy(ctx->ev_k1); }
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* pitt_last_error(const pitt_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context (no CUDA device?)"; }
double pitt_last_device_ms(const pitt_ctx* ctx) { return ctx ? ctx->last_ms : 0.0; }
int64_t pitt_kernel_launches(const pitt_ctx* ctx) { return ctx ? ctx->launches : 0; }
int pitt_set_workers(pitt_ctx* ctx, int n_workers) {
  if (!ctx) return PITT_ERR_CUDA;
  if (n_workers < 0 || n_workers > 64) return fail(ctx, PITT_ERR_INVALID, "pitt_set_workers: 0..64");
  if (n_workers != ctx->n_workers) {
    workers_destroy(ctx);
    ctx->n_workers = n_workers;
  }
  return PITT_OK;
}
int pitt_set_blocking_sync(pitt_ctx* ctx, int enable) {
  if (!ctx) return PITT_ERR_CUDA;
  cudaSetDevice(ctx->device);
  if (enable && !ctx->ev_block) {
    PITT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_block, cudaEventBlockingSync | cudaEventDisableTiming));
  } else if (!enable && ctx->ev_block) {
    cudaEventDestroy(ctx->ev_block);
    ctx->ev_block = nullptr;
  }
  return PITT_OK;
}
int pitt_synchronize(pitt_ctx* ctx) {
  if (!ctx) return PITT_ERR_CUDA;
  PITT_CUDA(ctx, pitt::stream_sync(ctx));
  return PITT_OK;
}

// ------------------------------------------------------------------ defaults
void pitt_default_sac_params(int model, pitt_sac_params* p) {
  memset(p, 0, sizeof(*p));
  p->model = model;
  p->max_iterations = 1000;
  p->probability = 0.99;
  p->radius_min = -DBL_MAX;
  p->radius_max = DBL_MAX;
  p->min_angle = -DBL_MAX;
  p->max_angle = DBL_MAX;
  p->optimize = 1;
  p->sampler = PITT_SAMPLER_PCL_MT19937;
  p->stop = PITT_STOP_PCL_ADAPTIVE;
  const double deg = M_PI / 180.0;
  (void)deg;
  switch (model) {
    case PITT_MODEL_PLANE:  // plane_segmentation_srv.cpp:19-24
      p->normal_distance_weight = 0.001; p->distance_threshold = 0.007; p->eps_angle = 0.0;
      p->min_angle = 0.0 / 180.0 * M_PI; p->max_angle = 10.0 / 180.0 * M_PI;
      break;
    case PITT_MODEL_SPHERE:  // sphere_segmentation_srv.cpp:19-26
      p->normal_distance_weight = 0.001; p->distance_threshold = 0.007; p->eps_angle = 0.0;
      p->radius_min = 0.005; p->radius_max = 0.500;
      p->min_angle = 100.0 / 180.0 * M_PI; p->max_angle = 180.0 / 180.0 * M_PI;
      break;
    case PITT_MODEL_CYLINDER:  // cylinder_segmentation_srv.cpp:23-30
      p->normal_distance_weight = 0.001; p->distance_threshold = 0.008; p->eps_angle = 0.0001;
      p->radius_min = 0.005; p->radius_max = 0.500;
      p->min_angle = 50.0 / 180.0 * M_PI; p->max_angle = 180.0 / 180.0 * M_PI;
      break;
    default:  // cone_segmentation_srv.cpp:24-31
      p->normal_distance_weight = 0.0006; p->distance_threshold = 0.0055; p->eps_angle = 0.4;
      p->radius_min = 0.001; p->radius_max = 0.500;
      p->min_angle = 10.0 / 180.0 * M_PI; p->max_angle = 170.0 / 180.0 * M_PI;
      break;
  }
}

void pitt_default_support_sac_params(pitt_sac_params* p) {  // supports_segmentation_srv.cpp:35-37,89-111
  pitt_default_sac_params(PITT_MODEL_PLANE, p);
  p->distance_threshold = (double)0.02f;
  p->normal_distance_weight = (double)0.9f;
  p->max_iterations = 10;
  p->min_angle = -DBL_MAX;
  p->max_angle = DBL_MAX;
}

void pitt_default_support_params(pitt_support_params* p) {
  memset(p, 0, sizeof(*p));
  p->min_iterative_cloud_percentual_size = -1.0f;
  p->min_iterative_plane_percentual_size = -1.0f;
  p->variance_threshold_for_horizontal = -1.0f;
  p->ransac_distance_point_in_shape_threshold = -1.0f;
  p->ransac_model_normal_distance_weigth = -1.0f;
  p->ransac_max_iteration_threshold = -1;
  p->horizontal_axis_len = 1;  // {-1}: srvm::DEFAULT_SERVICE_VEC_PARAMETER_REQUEST
  p->horizontal_axis[0] = -1.0f;
  p->support_edge_remove_offset_len = 1;
  p->support_edge_remove_offset[0] = -1.0f;
  p->normals_k = 50;
  p->compute_discarded_normals = 0;
}

void pitt_default_cluster_params(pitt_cluster_params* p) {  // cluster_segmentation_srv.cpp:32-35
  p->tolerance = 0.03;
  p->min_rate = 0.01;
  p->max_rate = 0.99;
  p->min_input_size = 30;
  p->reserved = 0;
}

void pitt_default_frame_params(pitt_frame_params* p) {
  memset(p, 0, sizeof(*p));
  pitt_default_support_params(&p->support);
  pitt_default_cluster_params(&p->cluster);
  pitt_default_sac_params(PITT_MODEL_PLANE, &p->plane);
  pitt_default_sac_params(PITT_MODEL_SPHERE, &p->sphere);
  pitt_default_sac_params(PITT_MODEL_CYLINDER, &p->cylinder);
  pitt_default_sac_params(PITT_MODEL_CONE, &p->cone);
  p->normals_k = 50;
  p->min_points = 30;
  p->viewpoint[0] = p->viewpoint[1] = p->viewpoint[2] = 0.0f;
  p->cone_over_cylinder_priority = 0.9f;
}

// ------------------------------------------------------------------ staging
int pitt_stage_cloud(pitt_ctx* ctx, const void* xyz, int stride_bytes, int n, pitt_cloud** out) {
  if (!ctx) return PITT_ERR_CUDA;
  if (!out || n < 0 || (n > 0 && !xyz) || stride_bytes < 12 || (stride_bytes & 3)) return fail(ctx, PITT_ERR_INVALID, "pitt_stage_cloud arguments");
  cudaSetDevice(ctx->device);
  CallTimer timer(ctx);
  pitt_cloud* c = new pitt_cloud();
  c->n = n;
  if (n > 0) {
    cudaError_t e = cudaSuccess;
    if (pool_alloc(ctx, (size_t)n * sizeof(float4), (void**)&c->d_xyz) != PITT_OK) { delete c; return PITT_ERR_CUDA; }
    if (stride_bytes == 16) {
      // pcl::PointXYZ / PointCloud2 point_step 16: already the HBM layout, one DMA
      e = cudaMemcpyAsync(c->d_xyz, xyz, (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream);
      if (e != cudaSuccess) { cudaFree(c->d_xyz); delete c; return fail(ctx, PITT_ERR_CUDA, "cudaMemcpyAsync(H2D cloud)", e); }
    } else {
      unsigned char* d_raw = nullptr;
      int s = arena_alloc(ctx, (size_t)n * stride_bytes, &d_raw);
      if (s != PITT_OK) { cudaFree(c->d_xyz); delete c; return s; }
      e = cudaMemcpyAsync(d_raw, xyz, (size_t)n * stride_bytes, cudaMemcpyHostToDevice, ctx->stream);
      if (e != cudaSuccess) { cudaFree(c->d_xyz); delete c; return fail(ctx, PITT_ERR_CUDA, "cudaMemcpyAsync(H2D cloud)", e); }
      pack_xyz_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(d_raw, stride_bytes, n, c->d_xyz);
      ctx->launches++;
    }
    e = pitt::stream_sync(ctx);  // the caller's buffer is only borrowed for the call
    if (e != cudaSuccess) { cudaFree(c->d_xyz); delete c; return fail(ctx, PITT_ERR_CUDA, "stage sync", e); }
  }
  timer.finish();
  *out = c;
  return PITT_OK;
}

int pitt_stage_cloud_device(pitt_ctx* ctx, const void* d_xyz4, int n, pitt_cloud** out) {
  if (!ctx) return PITT_ERR_CUDA;
  if (!out || n < 0 || (n > 0 && !d_xyz4)) return fail(ctx, PITT_ERR_INVALID, "pitt_stage_cloud_device arguments");
  cudaSetDevice(ctx->device);
  pitt_cloud* c = new pitt_cloud();
  c->n = n;
  if (n > 0) {
    cudaError_t e = cudaSuccess;
    if (pool_alloc(ctx, (size_t)n * sizeof(float4), (void**)&c->d_xyz) != PITT_OK) { delete c; return PITT_ERR_CUDA; }
    e = cudaMemcpyAsync(c->d_xyz, d_xyz4, (size_t)n * 16, cudaMemcpyDeviceToDevice, ctx->stream);
    if (e != cudaSuccess) { cudaFree(c->d_xyz); delete c; return fail(ctx, PITT_ERR_CUDA, "cudaMemcpyAsync(D2D cloud)", e); }
  }
  *out = c;
  return PITT_OK;
}

int pitt_set_normals(pitt_ctx* ctx, pitt_cloud* c, const void* normals, int stride_bytes) {
  if (!ctx) return PITT_ERR_CUDA;
  if (!c || (c->n > 0 && !normals) || (stride_bytes != 16 && stride_bytes != 32)) return fail(ctx, PITT_ERR_INVALID, "pitt_set_normals arguments");
  cudaSetDevice(ctx->device);
  CallTimer timer(ctx);
  if (c->n > 0) {
    if (!c->d_nrm) PITT_TRY(pool_alloc(ctx, (size_t)c->n * sizeof(float4), (void**)&c->d_nrm));
    if (stride_bytes == 16) {
      PITT_CUDA(ctx, cudaMemcpyAsync(c->d_nrm, normals, (size_t)c->n * 16, cudaMemcpyHostToDevice, ctx->stream));
    } else {
      unsigned char* d_raw = nullptr;
      PITT_TRY(arena_alloc(ctx, (size_t)c->n * stride_bytes, &d_raw));
      PITT_CUDA(ctx, cudaMemcpyAsync(d_raw, normals, (size_t)c->n * stride_bytes, cudaMemcpyHostToDevice, ctx->stream));
      pack_normals_kernel<<<cdiv(c->n, 256), 256, 0, ctx->stream>>>(d_raw, stride_bytes, 4, c->n, c->d_nrm);
      ctx->launches++;
    }
    PITT_CUDA(ctx, pitt::stream_sync(ctx));
  }
  c->has_normals = true;
  timer.finish();
  return PITT_OK;
}

int pitt_cloud_size(const pitt_cloud* c) { return c ? c->n : 0; }
int pitt_cloud_has_normals(const pitt_cloud* c) { return (c && c->has_normals) ? 1 : 0; }
const void* pitt_cloud_device_points(const pitt_cloud* c) { return c ? c->d_xyz : nullptr; }
const void* pitt_cloud_device_normals(const pitt_cloud* c) { return (c && c->has_normals) ? c->d_nrm : nullptr; }

void pitt_release_cloud(pitt_ctx* ctx, pitt_cloud* c) {
  if (!c) return;
  if (ctx) {
    cudaSetDevice(ctx->device);
    pitt::stream_sync(ctx);
  }
  pool_free(ctx, c->d_xyz, (size_t)c->n * sizeof(float4));
  pool_free(ctx, c->d_nrm, (size_t)c->n * sizeof(float4));
  delete c;
}

int pitt_get_normals(pitt_ctx* ctx, const pitt_cloud* c, float* out4) {
  if (!ctx) return PITT_ERR_CUDA;
  if (!c || !out4) return fail(ctx, PITT_ERR_INVALID, "pitt_get_normals arguments");
  if (!c->has_normals) return fail(ctx, PITT_ERR_STATE, "cloud has no normals");
  cudaSetDevice(ctx->device);
  if (c->n > 0) {
    PITT_CUDA(ctx, cudaMemcpyAsync(out4, c->d_nrm, (size_t)c->n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    PITT_CUDA(ctx, pitt::stream_sync(ctx));
  }
  return PITT_OK;
}

// ------------------------------------------------------------------ seg.segment()
int pitt_sac_segment(pitt_ctx* ctx, const pitt_cloud* c, const pitt_sac_params* p, int32_t* inliers, int cap,
                     int* n_inliers, float coeffs[8], int* n_coeffs, pitt_sac_info* info) {
  if (!ctx) return PITT_ERR_CUDA;
  if (!c || !p || !n_inliers || !coeffs || !n_coeffs) return fail(ctx, PITT_ERR_INVALID, "pitt_sac_segment arguments");
  cudaSetDevice(ctx->device);
  CallTimer timer(ctx);
  SacDeviceResult r;
  *n_inliers = 0;
  *n_coeffs = 0;
  PITT_TRY(sac_segment_impl(ctx, c, *p, &r));
  *n_inliers = r.n_inliers;
  *n_coeffs = r.n_coeffs;
  for (int i = 0; i < 8; ++i) coeffs[i] = r.coeffs[i];
  int status = PITT_OK;
  if (r.n_inliers > 0 && inliers) {  // inliers == NULL: the caller only wants the count and the coefficients
    if (cap < r.n_inliers) {
      status = fail(ctx, PITT_ERR_CAPACITY, "inlier buffer too small");
    } else {
      PITT_CUDA(ctx, cudaMemcpyAsync(inliers, r.d_inliers, (size_t)r.n_inliers * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      PITT_CUDA(ctx, pitt::stream_sync(ctx));
    }
  }
  timer.finish();
  if (info) {
    *info = r.info;
    info->device_ms = ctx->last_ms;
  }
  return status;
}

static int g_stream_chunks = -1;  // -1: read PITT_STREAM_CHUNKS on first use; 0: by cloud size; k: k equal chunks
void pitt_debug_stream_chunks(int k) { g_stream_chunks = k < 0 ? 0 : (k > 8 ? 8 : k); }

/* seg.segment() on a cloud that is still in host memory (the PointCloud2 payload of the service request): stage + segment +
 * release in one call. Plane models on point_step-16 clouds of at least 2^18 points take the fused path (sample points
 * fetched ahead, no intermediate synchronisation; clouds of 16 M points and more travel in chunks on a second stream and
 * are scored chunk by chunk under the copy); everything else is the three calls one after the other.
 * Results are identical to pitt_stage_cloud + pitt_sac_segment (tests/test_gpu_plane.py::test_segment_host_*). */
int pitt_sac_segment_host(pitt_ctx* ctx, const void* xyz, int stride_bytes, int n, const pitt_sac_params* p, int32_t* inliers, int cap,
                          int* n_inliers, float coeffs[8], int* n_coeffs, pitt_sac_info* info) {
  if (!ctx) return PITT_ERR_CUDA;
  if (!p || !n_inliers || !coeffs || !n_coeffs || n < 0 || (n > 0 && !xyz)) return fail(ctx, PITT_ERR_INVALID, "pitt_sac_segment_host arguments");
  cudaSetDevice(ctx->device);
  const bool streaming = (p->model == PITT_MODEL_PLANE) && stride_bytes == 16 && n >= (1 << 18) && p->sampler != PITT_SAMPLER_PHILOX;
  if (!streaming) {
    pitt_cloud* c = nullptr;
    PITT_TRY(pitt_stage_cloud(ctx, xyz, stride_bytes, n, &c));
    const int st = pitt_sac_segment(ctx, c, p, inliers, cap, n_inliers, coeffs, n_coeffs, info);
    pitt_release_cloud(ctx, c);
    return st;
  }
  // Chunk plan. Every extra chunk costs one more scoring launch (set-up kernels, pipeline fill, an imbalanced last round:
  // ~0.1 ms measured), a chunk's copy hides only while the previous chunk is being scored: chunks of at least 8 M points
  // (128 MB, 2.3 ms of PCIe time). A 1 M-point cloud (C2) therefore travels in one piece - the fused call still saves the
  // stage/segment/release round trips (1.27 vs 1.40 ms) -, a 50 M-point cloud (C5) in 6.
  // PITT_STREAM_CHUNKS=k / pitt_debug_stream_chunks(k) (1..8): k equal chunks whatever the size (tuning hook, tests).
  if (g_stream_chunks < 0) {
    const char* v = getenv("PITT_STREAM_CHUNKS");
    g_stream_chunks = v ? std::max(1, std::min(8, atoi(v))) : 0;
  }
  const int K_env = g_stream_chunks;
  // Large scoring jobs take the tensor path: ONE launch whose CTAs poll per-chunk arrival flags (16 chunks; a chunk costs
  // nothing but a 4-byte flag copy there), so only the first chunk's copy is exposed.
  const int H_first = p->stop == PITT_STOP_ALL_H ? p->max_iterations : p->max_iterations + 1;
  const bool single_launch = K_env == 0 && plane_job_takes_tensor_path(n, H_first) &&
                             (p->sampler != PITT_SAMPLER_REPLAY || p->replay_count >= H_first);
  const int K_equal = single_launch ? 16 : (K_env > 0 ? K_env : std::max(1, std::min(8, n / (8 << 20))));
  if (!ctx->copy_stream) {
    PITT_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 16; ++k) PITT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_chunk[k], cudaEventDisableTiming));
    PITT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_copy_gate, cudaEventDisableTiming));
    PITT_CUDA(ctx, cudaMalloc((void**)&ctx->d_ready, 16 * sizeof(int)));
    PITT_CUDA(ctx, cudaMallocHost((void**)&ctx->h_one, sizeof(int)));
    *ctx->h_one = 1;
  }
  trace_mark("segment_host: enter");
  CallTimer timer(ctx);
  pitt_cloud* c = new pitt_cloud();
  c->n = n;
  if (pool_alloc(ctx, (size_t)n * sizeof(float4), (void**)&c->d_xyz) != PITT_OK) { delete c; return PITT_ERR_CUDA; }
  trace_mark("segment_host: pool_alloc done");
  c->h_src = static_cast<const float*>(xyz);
  {
    auto round512 = [](long long v) { return (int)(((v + 511) / 512) * 512); };  // whole 512-point chunks of the scoring kernels
    int k = 0;
    c->stream_off[0] = 0;
    const int step = round512(cdiv(n, K_equal));
    while (c->stream_off[k] < n) { c->stream_off[k + 1] = std::min(n, c->stream_off[k] + step); ++k; }
    c->stream_chunks = k;
  }
  // the buffer may have been used by earlier work of this context's stream
  cudaEventRecord(ctx->ev_copy_gate, ctx->stream);
  cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy_gate, 0);
  if (single_launch) {
    cudaMemsetAsync(ctx->d_ready, 0, 16 * sizeof(int), ctx->copy_stream);
    c->d_ready = ctx->d_ready;
  }
  for (int k = 0; k < c->stream_chunks; ++k) {
    const size_t off = (size_t)c->stream_off[k];
    const size_t cnt = (size_t)c->stream_off[k + 1] - off;
    cudaError_t e = cudaMemcpyAsync(c->d_xyz + off, (const char*)xyz + off * 16, cnt * 16, cudaMemcpyHostToDevice, ctx->copy_stream);
    if (e == cudaSuccess && single_launch)  // the flag travels behind its chunk on the same stream
      e = cudaMemcpyAsync(ctx->d_ready + k, ctx->h_one, sizeof(int), cudaMemcpyHostToDevice, ctx->copy_stream);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_chunk[k], ctx->copy_stream);
    if (e != cudaSuccess) {
      cudaStreamSynchronize(ctx->copy_stream);
      pitt_release_cloud(ctx, c);
      return fail(ctx, PITT_ERR_CUDA, "chunked H2D of the cloud", e);
    }
  }
  trace_mark("segment_host: chunk copies issued");
  SacDeviceResult r;
  *n_inliers = 0;
  *n_coeffs = 0;
  int status = sac_segment_impl(ctx, c, *p, &r);
  trace_mark("segment_host: sac_segment_impl returned");
  if (status == PITT_OK) {
    *n_inliers = r.n_inliers;
    *n_coeffs = r.n_coeffs;
    for (int i = 0; i < 8; ++i) coeffs[i] = r.coeffs[i];
    if (r.n_inliers > 0 && inliers) {
      if (cap < r.n_inliers) {
        status = fail(ctx, PITT_ERR_CAPACITY, "inlier buffer too small");
      } else {
        cudaError_t e = cudaMemcpyAsync(inliers, r.d_inliers, (size_t)r.n_inliers * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) status = fail(ctx, PITT_ERR_CUDA, "D2H of the inlier list", e);
      }
    }
  }
  // the caller's buffer is only borrowed for the call: every copy has left it (early returns of the segmentation included)
  cudaStreamSynchronize(ctx->copy_stream);
  pitt::stream_sync(ctx);
  trace_mark("segment_host: inliers on the host");
  c->stream_chunks = 0;
  timer.finish();
  if (info && status == PITT_OK) {
    *info = r.info;
    info->device_ms = ctx->last_ms;
  }
  pitt_release_cloud(ctx, c);
  return status;
}

static int score_common(pitt_ctx* ctx, const pitt_cloud* c, const pitt_sac_params* p, const int* d_samples, int H,
                        int* d_counts, float* d_coeffs8, uint8_t* d_flags) {
  if ((p->model == PITT_MODEL_CYLINDER || p->model == PITT_MODEL_CONE) && !c->has_normals)
    return fail(ctx, PITT_ERR_STATE, "cylinder/cone scoring needs normals on the cloud");
  const Limits L = limits_for(*p);
  const ScoreParams sp = score_params_for(*p, L);
  HypRec* d_recs = nullptr;
  PITT_TRY(arena_alloc(ctx, (size_t)H, &d_recs));
  PITT_TRY(sac_estimate(ctx, c, p->model, d_samples, H, L, sp, d_recs, d_coeffs8, d_flags));
  PITT_TRY(sac_score(ctx, c, p->model, d_recs, H, sp, d_counts));
  return PITT_OK;
}

int pitt_sac_score(pitt_ctx* ctx, const pitt_cloud* c, const pitt_sac_params* p, const int32_t* samples, int H,
                   int32_t* counts, float* coeffs8, uint8_t* valid) {
  if (!ctx) return PITT_ERR_CUDA;
  if (!c || !p || !samples || H < 0 || p->model < 0 || p->model > 3) return fail(ctx, PITT_ERR_INVALID, "pitt_sac_score arguments");
  if (H == 0) return PITT_OK;
  cudaSetDevice(ctx->device);
  CallTimer timer(ctx);
  const int S = p->model == PITT_MODEL_PLANE ? 3 : p->model == PITT_MODEL_SPHERE ? 4 : p->model == PITT_MODEL_CYLINDER ? 2 : 3;
  for (size_t i = 0; i < (size_t)H * S; ++i)
    if (samples[i] < 0 || samples[i] >= c->n) return fail(ctx, PITT_ERR_INVALID, "sample index out of range");
  int* d_samples = nullptr;
  int* d_counts = nullptr;
  float* d_coeffs8 = nullptr;
  uint8_t* d_flags = nullptr;
  PITT_TRY(arena_alloc(ctx, (size_t)H * S, &d_samples));
  PITT_TRY(arena_alloc(ctx, (size_t)H, &d_counts));
  PITT_TRY(arena_alloc(ctx, (size_t)H * 8, &d_coeffs8));
  PITT_TRY(arena_alloc(ctx, (size_t)H, &d_flags));
  PITT_CUDA(ctx, cudaMemcpyAsync(d_samples, samples, (size_t)H * S * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  PITT_TRY(score_common(ctx, c, p, d_samples, H, d_counts, d_coeffs8, d_flags));
  if (counts) PITT_CUDA(ctx, cudaMemcpyAsync(counts, d_counts, (size_t)H * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (coeffs8) PITT_CUDA(ctx, cudaMemcpyAsync(coeffs8, d_coeffs8, (size_t)H * 8 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  std::vector<uint8_t> flags;
  if (valid) {
    flags.resize(H);
    PITT_CUDA(ctx, cudaMemcpyAsync(flags.data(), d_flags, (size_t)H, cudaMemcpyDeviceToHost, ctx->stream));
  }
  PITT_CUDA(ctx, pitt::stream_sync(ctx));
  if (valid)
    for (int h = 0; h < H; ++h) valid[h] = flags[h] & 1;
  timer.finish();
  return PITT_OK;
}

int pitt_sac_score_device(pitt_ctx* ctx, const pitt_cloud* c, const pitt_sac_params* p, const void* d_samples, int H,
                          void* d_counts) {
  if (!ctx) return PITT_ERR_CUDA;
  if (!c || !p || !d_samples || !d_counts || H <= 0 || p->model < 0 || p->model > 3) return fail(ctx, PITT_ERR_INVALID, "pitt_sac_score_device arguments");
  cudaSetDevice(ctx->device);
  // no CallTimer: nothing here synchronises; callers bracket with their own events
  arena_reset(ctx);
  uint8_t* d_flags = nullptr;
  PITT_TRY(arena_alloc(ctx, (size_t)H, &d_flags));
  return score_common(ctx, c, p, (const int*)d_samples, H, (int*)d_counts, nullptr, d_flags);
}

int pitt_sac_finish_device(pitt_ctx* ctx, const pitt_cloud* c, const pitt_sac_params* p, const void* d_samples_all, int H_all,
                           const void* d_best, int32_t* inliers, int cap, int* n_inliers, float* coeffs, int* n_coeffs,
                           pitt_sac_info* info) {
  if (!ctx) return PITT_ERR_CUDA;
  if (!c || !p || !d_samples_all || !d_best || !n_inliers || !coeffs || !n_coeffs || H_all <= 0 || p->model < 0 || p->model > 3)
    return fail(ctx, PITT_ERR_INVALID, "pitt_sac_finish_device arguments");
  if ((p->model == PITT_MODEL_CYLINDER || p->model == PITT_MODEL_CONE) && !c->has_normals)
    return fail(ctx, PITT_ERR_STATE, "cylinder/cone segmentation needs normals on the cloud");
  cudaSetDevice(ctx->device);
  CallTimer timer(ctx);
  SacDeviceResult r;
  PITT_TRY(sac_finish_from_winner(ctx, c, *p, (const int*)d_samples_all, H_all, (const int*)d_best, &r));
  *n_inliers = r.n_inliers;
  *n_coeffs = r.n_coeffs;
  for (int i = 0; i < r.n_coeffs; ++i) coeffs[i] = r.coeffs[i];
  int status = PITT_OK;
  if (inliers && r.n_inliers > 0) {
    if (cap < r.n_inliers) status = fail(ctx, PITT_ERR_CAPACITY, "inlier buffer too small");
    else {
      PITT_CUDA(ctx, cudaMemcpyAsync(inliers, r.d_inliers, (size_t)r.n_inliers * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      PITT_CUDA(ctx, pitt::stream_sync(ctx));
    }
  }
  timer.finish();
  if (info) {
    *info = r.info;
    info->device_ms = ctx->last_ms;
  }
  return status;
}

int pitt_sac_select(pitt_ctx* ctx, const pitt_cloud* c, const pitt_sac_params* p, const float* coeffs,
                    int32_t* inliers, int cap, int* n_inliers) {
  if (!ctx) return PITT_ERR_CUDA;
  if (!c || !p || !coeffs || !n_inliers) return fail(ctx, PITT_ERR_INVALID, "pitt_sac_select arguments");
  if ((p->model == PITT_MODEL_CYLINDER || p->model == PITT_MODEL_CONE) && !c->has_normals)
    return fail(ctx, PITT_ERR_STATE, "cylinder/cone selection needs normals on the cloud");
  cudaSetDevice(ctx->device);
  CallTimer timer(ctx);
  const Limits L = limits_for(*p);
  const ScoreParams sp = score_params_for(*p, L);
  float* d_co = nullptr;
  int* d_out = nullptr;
  int* d_total = nullptr;
  PITT_TRY(arena_alloc(ctx, 8, &d_co));
  PITT_TRY(arena_alloc(ctx, (size_t)c->n + 1, &d_out));
  PITT_TRY(arena_alloc(ctx, 1, &d_total));
  float h_co[8] = {0};
  const int NC = (p->model == PITT_MODEL_PLANE || p->model == PITT_MODEL_SPHERE) ? 4 : 7;
  for (int i = 0; i < NC; ++i) h_co[i] = coeffs[i];
  PITT_CUDA(ctx, cudaMemcpyAsync(d_co, h_co, sizeof(h_co), cudaMemcpyHostToDevice, ctx->stream));
  PITT_TRY(sac_select(ctx, c, p->model, d_co, L, sp, d_out, d_total));
  int total = 0;
  PITT_CUDA(ctx, cudaMemcpyAsync(&total, d_total, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PITT_CUDA(ctx, pitt::stream_sync(ctx));
  *n_inliers = total;
  int status = PITT_OK;
  if (total > 0) {
    if (!inliers || cap < total) status = fail(ctx, PITT_ERR_CAPACITY, "inlier buffer too small");
    else {
      PITT_CUDA(ctx, cudaMemcpyAsync(inliers, d_out, (size_t)total * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      PITT_CUDA(ctx, pitt::stream_sync(ctx));
    }
  }
  timer.finish();
  return status;
}

int pitt_sac_refine(pitt_ctx* ctx, const pitt_cloud* c, const pitt_sac_params* p, const float* coeffs,
                    const int32_t* inliers, int n_inliers, float* refined, pitt_sac_info* info) {
  if (!ctx) return PITT_ERR_CUDA;
  if (!c || !p || !coeffs || !refined || n_inliers < 0 || (n_inliers > 0 && !inliers)) return fail(ctx, PITT_ERR_INVALID, "pitt_sac_refine arguments");
  cudaSetDevice(ctx->device);
  CallTimer timer(ctx);
  const Limits L = limits_for(*p);
  const ScoreParams sp = score_params_for(*p, L);
  const int NC = (p->model == PITT_MODEL_PLANE || p->model == PITT_MODEL_SPHERE) ? 4 : 7;
  float* d_co = nullptr;
  float* d_ref = nullptr;
  int* d_idx = nullptr;
  int* d_ints = nullptr;
  PITT_TRY(arena_alloc(ctx, 8, &d_co));
  PITT_TRY(arena_alloc(ctx, 8, &d_ref));
  PITT_TRY(arena_alloc(ctx, (size_t)n_inliers + 1, &d_idx));
  PITT_TRY(arena_alloc(ctx, 4, &d_ints));
  float h_co[8] = {0};
  for (int i = 0; i < NC; ++i) h_co[i] = coeffs[i];
  int h_ints[4] = {n_inliers, 0, 0, 0};
  PITT_CUDA(ctx, cudaMemcpyAsync(d_co, h_co, sizeof(h_co), cudaMemcpyHostToDevice, ctx->stream));
  PITT_CUDA(ctx, cudaMemcpyAsync(d_ints, h_ints, sizeof(h_ints), cudaMemcpyHostToDevice, ctx->stream));
  if (n_inliers > 0) PITT_CUDA(ctx, cudaMemcpyAsync(d_idx, inliers, (size_t)n_inliers * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  if (p->model == PITT_MODEL_PLANE) {
    PITT_TRY(plane_refine(ctx, c, d_co, d_idx, d_ints, L, sp, d_ref, d_ints + 1));
  } else {
    PITT_TRY(lm_refine(ctx, c, p->model, d_co, d_idx, d_ints, n_inliers, d_ref, d_ints + 2));
  }
  float h_ref[8];
  PITT_CUDA(ctx, cudaMemcpyAsync(h_ref, d_ref, sizeof(h_ref), cudaMemcpyDeviceToHost, ctx->stream));
  PITT_CUDA(ctx, cudaMemcpyAsync(h_ints, d_ints, sizeof(h_ints), cudaMemcpyDeviceToHost, ctx->stream));
  PITT_CUDA(ctx, pitt::stream_sync(ctx));
  for (int i = 0; i < NC; ++i) refined[i] = h_ref[i];
  timer.finish();
  if (info) {
    memset(info, 0, sizeof(*info));
    info->lm_info = h_ints[2];
    info->lm_nfev = h_ints[3];
    info->device_ms = ctx->last_ms;
  }
  return PITT_OK;
}

int pitt_pcl_sample_stream(pitt_ctx* ctx, const pitt_cloud* c, int model, int count, int32_t* out) {
  if (!ctx) return PITT_ERR_CUDA;
  if (!c || !out || count < 0 || model < 0 || model > 3) return fail(ctx, PITT_ERR_INVALID, "pitt_pcl_sample_stream arguments");
  cudaSetDevice(ctx->device);
  if (model == PITT_MODEL_PLANE) PITT_TRY(ensure_host_mirror(ctx, c));
  PclSampleStream s(c->n, model, model == PITT_MODEL_PLANE ? c->h_xyz.data() : nullptr);
  const int S = model == PITT_MODEL_PLANE ? 3 : model == PITT_MODEL_SPHERE ? 4 : model == PITT_MODEL_CYLINDER ? 2 : 3;
  for (int h = 0; h < count; ++h)
    if (!s.next(out + (size_t)h * S)) return fail(ctx, PITT_ERR_INVALID, "cloud smaller than the sample size");
  return PITT_OK;
}

int pitt_fp32_peak(pitt_ctx* ctx, int kind, double* tflops) {
  if (!ctx) return PITT_ERR_CUDA;
  if (!tflops || kind < 0 || kind > 2) return fail(ctx, PITT_ERR_INVALID, "pitt_fp32_peak arguments");
  cudaSetDevice(ctx->device);
  CallTimer timer(ctx);
  int s = fp32_peak(ctx, kind, tflops);
  timer.finish();
  return s;
}

/* arg-max of gathered per-hypothesis counts (multi-GPU hypothesis split, config 5): d_counts are H
 * int32 on the device, d_best receives {index, count}; ties keep the earliest hypothesis. */
int pitt_argmax_counts_device(pitt_ctx* ctx, const void* d_counts, int H, void* d_best) {
  if (!ctx) return PITT_ERR_CUDA;
  if (!d_counts || !d_best || H <= 0) return fail(ctx, PITT_ERR_INVALID, "pitt_argmax_counts_device arguments");
  cudaSetDevice(ctx->device);
  arena_reset(ctx);
  uint8_t* d_flags = nullptr;
  float* d_dummy = nullptr;
  PITT_TRY(arena_alloc(ctx, (size_t)H, &d_flags));
  PITT_TRY(arena_alloc(ctx, (size_t)H * 8 + 8, &d_dummy));
  PITT_CUDA(ctx, cudaMemsetAsync(d_flags, 1, (size_t)H, ctx->stream));
  return sac_winner(ctx, (const int*)d_counts, d_flags, H, d_dummy, (int*)d_best, d_dummy + (size_t)H * 8);
}


namespace pitt {
// hypothesis split: hypotheses whose model could not be estimated (PCL skips them) and the padding of the last rank's slice
// must never win the arg-max: their count becomes -1
__global__ void split_mask_counts_kernel(int* __restrict__ counts, const uint8_t* __restrict__ flags, int h_have, int h_loc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= h_loc) return;
  if (i >= h_have || !(flags[i] & 1)) counts[i] = -1;
}
}  // namespace pitt

/* SURVEY 8e / BASELINE configs[4]: seg.segment() of ONE cloud with the hypothesis set split over `world` ranks (one process
 * per GPU, every rank holds the cloud). Rank r estimates and scores hypotheses [r*Hl, (r+1)*Hl) of the sample stream,
 * Hl = ceil(H / world); `allgather` (the caller's NCCL all-gather over NVLink) collects the Hl int32
 * counts of every rank in rank order; every rank then takes the earliest arg-max (PCL keeps the first best model: strict '>'
 * in ransac.hpp), re-estimates the winner from its sample, refines it and selects the final inliers, so all ranks return
 * the same result as pitt_sac_segment on one GPU with stop = ALL_H. */
int pitt_sac_segment_split(pitt_ctx* ctx, const pitt_cloud* c, const pitt_sac_params* p, int rank, int world,
                           pitt_allgather_fn allgather, void* user, int32_t* inliers, int cap, int* n_inliers, float coeffs[8],
                           int* n_coeffs, pitt_sac_info* info) {
  if (!ctx) return PITT_ERR_CUDA;
  if (!c || !p || !n_inliers || !coeffs || !n_coeffs || world < 1 || rank < 0 || rank >= world || (world > 1 && !allgather) ||
      p->model < 0 || p->model > 3)
    return fail(ctx, PITT_ERR_INVALID, "pitt_sac_segment_split arguments");
  if (p->stop != PITT_STOP_ALL_H) return fail(ctx, PITT_ERR_INVALID, "pitt_sac_segment_split: only stop = PITT_STOP_ALL_H can be split");
  if ((p->model == PITT_MODEL_CYLINDER || p->model == PITT_MODEL_CONE) && !c->has_normals)
    return fail(ctx, PITT_ERR_STATE, "cylinder/cone segmentation needs normals on the cloud");
  cudaSetDevice(ctx->device);
  CallTimer timer(ctx);
  *n_inliers = 0;
  *n_coeffs = 0;
  if (info) {
    memset(info, 0, sizeof(*info));
    info->best_hypothesis = -1;
  }
  const int S = p->model == PITT_MODEL_PLANE ? 3 : p->model == PITT_MODEL_SPHERE ? 4 : p->model == PITT_MODEL_CYLINDER ? 2 : 3;
  int H_all = p->max_iterations;
  if (p->sampler == PITT_SAMPLER_REPLAY) {
    if (!p->replay_samples) return fail(ctx, PITT_ERR_INVALID, "replay_samples is null");
    H_all = std::min(H_all, p->replay_count);
  }
  if (H_all <= 0 || c->n < S) {
    timer.finish();
    return PITT_OK;
  }
  const int H_loc = (H_all + world - 1) / world;
  const int h0 = std::min(H_all, rank * H_loc);
  const int H_have = std::max(0, std::min(H_all, h0 + H_loc) - h0);
  int* d_samples_all = nullptr;
  int* d_counts = nullptr;
  int* d_counts_all = nullptr;
  int* d_best = nullptr;
  uint8_t* d_flags = nullptr;
  float* d_dummy = nullptr;
  uint8_t* d_ones = nullptr;
  PITT_TRY(arena_alloc(ctx, (size_t)H_all * S, &d_samples_all));
  PITT_TRY(arena_alloc(ctx, (size_t)H_loc, &d_counts));
  PITT_TRY(arena_alloc(ctx, (size_t)H_loc * world, &d_counts_all));
  PITT_TRY(arena_alloc(ctx, 2, &d_best));
  PITT_TRY(arena_alloc(ctx, (size_t)H_loc, &d_flags));
  PITT_TRY(arena_alloc(ctx, (size_t)H_loc * world, &d_ones));
  PITT_TRY(arena_alloc(ctx, 16, &d_dummy));
  // the whole sample stream on every rank (the winner's sample is needed everywhere)
  if (p->sampler == PITT_SAMPLER_PHILOX) {
    PITT_TRY(sac_philox_samples(ctx, d_samples_all, H_all, S, c->n, 1u));
  } else {
    PITT_TRY(pinned_reserve(ctx, (size_t)H_all * S * sizeof(int)));
    int* h = (int*)ctx->h_pin;
    if (p->sampler == PITT_SAMPLER_REPLAY) {
      for (size_t i = 0; i < (size_t)H_all * S; ++i) {
        const int v = p->replay_samples[i];
        if (v < 0 || v >= c->n) return fail(ctx, PITT_ERR_INVALID, "sample index out of range");
        h[i] = v;
      }
    } else {
      if (p->model == PITT_MODEL_PLANE) PITT_TRY(ensure_host_mirror(ctx, c));
      PclSampleStream s(c->n, p->model, p->model == PITT_MODEL_PLANE ? c->h_xyz.data() : nullptr);
      for (int i = 0; i < H_all; ++i)
        if (!s.next(h + (size_t)i * S)) return fail(ctx, PITT_ERR_INVALID, "cloud smaller than the sample size");
    }
    PITT_CUDA(ctx, cudaMemcpyAsync(d_samples_all, h, (size_t)H_all * S * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  }
  PITT_CUDA(ctx, cudaMemsetAsync(d_flags, 0, (size_t)H_loc, ctx->stream));
  PITT_CUDA(ctx, cudaMemsetAsync(d_counts, 0, (size_t)H_loc * sizeof(int), ctx->stream));
  if (H_have > 0) PITT_TRY(score_common(ctx, c, p, d_samples_all + (size_t)h0 * S, H_have, d_counts, nullptr, d_flags));
  split_mask_counts_kernel<<<cdiv(H_loc, 256), 256, 0, ctx->stream>>>(d_counts, d_flags, H_have, H_loc);
  ctx->launches++;
  const int* d_gathered = d_counts;
  if (world > 1) {
    const int st = allgather(user, d_counts, d_counts_all, H_loc, (void*)ctx->stream);
    if (st != 0) return fail(ctx, PITT_ERR_CUDA, "pitt_sac_segment_split: the caller's all-gather failed");
    d_gathered = d_counts_all;
  }
  // earliest arg-max over the gathered counts: global stream position = rank * H_loc + i (padding and failed models are -1)
  PITT_CUDA(ctx, cudaMemsetAsync(d_ones, 1, (size_t)H_loc * world, ctx->stream));
  PITT_TRY(sac_winner(ctx, d_gathered, d_ones, H_loc * world, nullptr, d_best, d_dummy));
  SacDeviceResult r;
  PITT_TRY(sac_finish_from_winner(ctx, c, *p, d_samples_all, H_all, d_best, &r));
  int status = PITT_OK;
  if (r.info.best_hypothesis >= 0 && r.info.best_count >= 0) {
    *n_inliers = r.n_inliers;
    *n_coeffs = r.n_coeffs;
    for (int i = 0; i < r.n_coeffs; ++i) coeffs[i] = r.coeffs[i];
    if (inliers && r.n_inliers > 0) {
      if (cap < r.n_inliers) status = fail(ctx, PITT_ERR_CAPACITY, "inlier buffer too small");
      else {
        PITT_CUDA(ctx, cudaMemcpyAsync(inliers, r.d_inliers, (size_t)r.n_inliers * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        PITT_CUDA(ctx, pitt::stream_sync(ctx));
      }
    }
  }
  timer.finish();
  if (info) {
    *info = r.info;
    info->iterations = H_all;
    info->hypotheses = H_have;
    info->device_ms = ctx->last_ms;
  }
  return status;
}

/* test hook (not in the public header): route plane scoring through the generic kernel */
void pitt_debug_force_generic_plane(int on) { g_force_generic_plane = on; }
/* test hooks: plane scoring mode (0 = automatic: tensor path on large jobs, 1 = exact packed kernel only, 2 = FFMA filter + exact
 * re-evaluation always, 3 = tensor-core path + exact re-evaluation always) and
 * the filter statistics of the last scoring call made while collection was enabled:
 * out[0] = (hypothesis, point tile) pairs scored, out[1] = pairs re-evaluated exactly */
void pitt_debug_plane_mode(int mode) { g_plane_mode = mode; }
void pitt_debug_score_mode(int mode) { g_score_mode = mode; }
void pitt_debug_select_no_fuse(int v) { g_select_no_fuse = v; }
void pitt_debug_frame_mode(int mode) {
  g_frame_legacy = (mode == 1);
  g_frame_no_batch = (mode == 2);
}
void pitt_debug_lm_cluster_min(int rows) { g_lm_cluster_min = rows; }
void pitt_debug_plane_filter_stats(int enable, uint64_t* out2) {
  g_plane_filter_collect_stats = enable;
  if (out2) { out2[0] = g_plane_filter_stats[0]; out2[1] = g_plane_filter_stats[1]; }
}

/* test hooks of the tensor-core plane path (plane_tc.cu): statistics of the last call made while collection was enabled
 * (out[0] = (hypothesis, 128-point segment) pairs scored, out[1] = pairs re-evaluated exactly); raw accumulator dump of
 * hypothesis block 0 x point tile 0 (128 x 256 floats, then sigma and C); accumulation error bound in units of u m */
void pitt_debug_plane_tc_stats(int enable, uint64_t* out2) {
  g_plane_tc_collect_stats = enable;
  if (out2) { out2[0] = g_plane_tc_stats[0]; out2[1] = g_plane_tc_stats[1]; }
}
/* per-CTA (SM id << 48 | cycles) of the last call made while statistics were enabled, 160 entries */
void pitt_debug_plane_tc_cta_cycles(uint64_t* out160) {
  for (int i = 0; i < 176; ++i) out160[i] = g_plane_tc_stats[2 + i];
}
int pitt_debug_plane_tc_dump(int enable, float* out /*128*256 + 2, nullable*/) {
  g_plane_tc_dump = enable;
  if (out && !g_plane_tc_dump_host.empty()) {
    memcpy(out, g_plane_tc_dump_host.data(), g_plane_tc_dump_host.size() * sizeof(float));
    return (int)g_plane_tc_dump_host.size();
  }
  return 0;
}
void pitt_debug_plane_tc_acc_ulps(float ulps) { g_plane_tc_acc_ulps = ulps; }
void pitt_debug_plane_tc_variant(int v) { g_plane_tc_variant = v; }
/* enable: every tensor-path scoring call records two CUDA events around the plane_tc_kernel launch alone (on the
 * context's stream); pitt_debug_plane_tc_kernel_ms waits for the last pair and returns the kernel's duration (< 0: none) */
void pitt_debug_plane_tc_time_kernel(pitt_ctx* ctx, int enable) {
  if (ctx) ctx->time_tc_kernel = enable != 0;
}
double pitt_debug_plane_tc_kernel_ms(pitt_ctx* ctx) {
  if (!ctx || !ctx->ev_k0) return -1.0;
  float ms = -1.0f;
  if (cudaEventSynchronize(ctx->ev_k1) != cudaSuccess) return -1.0;
  if (cudaEventElapsedTime(&ms, ctx->ev_k0, ctx->ev_k1) != cudaSuccess) return -1.0;
  return (double)ms;
}
void pitt_debug_plane_tc_nwq(int v) { g_plane_tc_nwq = v; }
/* the device's Philox-4x32-10: one raw block for a given counter (4 words) and key (2 words) */
int pitt_debug_philox(pitt_ctx* ctx, const uint32_t* ctr4, const uint32_t* key2, uint32_t* out4) {
  if (!ctx || !ctr4 || !key2 || !out4) return PITT_ERR_INVALID;
  cudaSetDevice(ctx->device);
  arena_reset(ctx);
  uint32_t ck[6] = {ctr4[0], ctr4[1], ctr4[2], ctr4[3], key2[0], key2[1]};
  return sac_philox_raw(ctx, ck, out4);
}
/* the minimal sample sets the PHILOX sampler draws for hypotheses 0..H-1 of batch `stream_id` on a cloud of n points (S indices
 * each, written to the host array out) */
int pitt_debug_philox_samples(pitt_ctx* ctx, int H, int S, int n, uint32_t stream_id, int32_t* out) {
  if (!ctx || !out || H <= 0 || S < 1 || S > 4 || n < S) return PITT_ERR_INVALID;
  cudaSetDevice(ctx->device);
  arena_reset(ctx);
  int* d = nullptr;
  PITT_TRY(arena_alloc(ctx, (size_t)H * S, &d));
  PITT_TRY(sac_philox_samples(ctx, d, H, S, n, stream_id));
  PITT_CUDA(ctx, cudaMemcpyAsync(out, d, (size_t)H * S * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PITT_CUDA(ctx, pitt::stream_sync(ctx));
  return PITT_OK;
}

}  // extern "C"
