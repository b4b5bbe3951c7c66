// chs_api.cu — version, error string, configuration validation, workspace query and the NCCL
// communicator of libchs (include/chs.h).
#include <dlfcn.h>

#include <atomic>
#include <string.h>

#include "chs_common.cuh"

static thread_local char g_err[512] = "";

void chs_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<uint64_t> g_launches{0};
void chs_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" uint64_t chs_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int chs_version(void) { return CHS_VERSION; }
extern "C" uint64_t chs_sizeof(int32_t which) {
  return which == 0 ? sizeof(chs_config) : which == 1 ? sizeof(chs_workspace_sizes) : which == 2 ? sizeof(chs_tensors) : 0;
}
extern "C" const char* chs_last_error(void) { return g_err; }

int chs_make_dims(const chs_config* cfg, ChsDims* d) {
  CHS_REQUIRE(cfg && d, "null chs_config");
  CHS_REQUIRE(cfg->tile_size == CHS_TILE, "tile_size must be %d (got %d)", CHS_TILE, cfg->tile_size);
  CHS_REQUIRE(cfg->n_gauss >= 0 && cfg->n_frames >= 0 && cfg->n_virtual >= 1, "bad n_gauss / n_frames / n_virtual");
  CHS_REQUIRE(cfg->width >= 1 && cfg->height >= 1, "bad image size %d x %d", cfg->width, cfg->height);
  CHS_REQUIRE(cfg->crf_kind == CHS_CRF_IDENTITY || cfg->crf_kind == CHS_CRF_MLP || cfg->crf_kind == CHS_CRF_LUT, "unknown crf_kind %d",
              cfg->crf_kind);
  CHS_REQUIRE(cfg->crf_kind != CHS_CRF_MLP || (cfg->crf_hidden >= 1 && cfg->crf_hidden <= 128), "crf_hidden must be in [1,128]");
  CHS_REQUIRE(cfg->crf_kind != CHS_CRF_LUT || (cfg->crf_hidden >= 2 && cfg->crf_hidden <= 1024), "LUT CRF needs 2..1024 knots");
  CHS_REQUIRE(cfg->sort_mode == CHS_SORT_KEY64 || cfg->sort_mode == CHS_SORT_DEPTH_PRESORT, "unknown sort_mode %d", cfg->sort_mode);
  d->N = cfg->n_gauss; d->B = cfg->n_frames; d->n = cfg->n_virtual;
  d->C = d->B * d->n;
  d->W = cfg->width; d->H = cfg->height;
  d->tile_w = (d->W + CHS_TILE - 1) / CHS_TILE;
  d->tile_h = (d->H + CHS_TILE - 1) / CHS_TILE;
  d->tiles = d->tile_w * d->tile_h;
  d->tile_bits = chs_bit_length((uint64_t)d->tiles);
  d->cam_bits = chs_bit_length((uint64_t)d->C);
  d->CN = (int64_t)d->C * d->N;
  d->P = (int64_t)d->W * d->H;
  d->Cb = cfg->pose_fused ? d->B : d->C;
  d->CbN = (int64_t)d->Cb * d->N;
  CHS_REQUIRE(!cfg->pose_fused || (cfg->sort_mode == CHS_SORT_DEPTH_PRESORT && cfg->tune_bin == 0 && d->tile_w <= 256),
              "pose_fused needs the default binning route (CHS_SORT_DEPTH_PRESORT, tune_bin = 0, width <= 4096)");
  // the projection kernels keep the whole camera table in shared memory (112 B per camera in the backward)
  CHS_REQUIRE(d->C <= 1024, "too many cameras in one call (%d > 1024 = frames x virtual poses); split the frame batch", d->C);
  CHS_REQUIRE(d->CN < ((int64_t)1 << 31), "C*N = %lld exceeds int32 ids; split the frame batch", (long long)d->CN);
  CHS_REQUIRE((int64_t)d->C * d->tiles < ((int64_t)1 << 31), "C*tiles exceeds int32");
  CHS_REQUIRE(32 + d->tile_bits + d->cam_bits <= 64, "key does not fit 64 bits");
  return CHS_OK;
}

int chs_bin_count_bytes(const ChsDims& d, int sort_mode, uint64_t* bytes);
int chs_bin_sort_bytes(const ChsDims& d, const chs_config* cfg, int64_t M, uint64_t* bytes);

extern "C" int chs_workspace_query(const chs_config* cfg, int64_t n_isect, int32_t n_knots, chs_workspace_sizes* out) {
  ChsDims d;
  int st = chs_make_dims(cfg, &d);
  if (st) return st;
  CHS_REQUIRE(out, "chs_workspace_query: null output");
  memset(out, 0, sizeof(*out));
  if (d.CN > 0) {
    st = chs_bin_count_bytes(d, cfg->sort_mode, &out->bin_count_bytes);
    if (st) return st;
  }
  out->bin_count_bytes += 256;
  if (n_isect > 0) {
    st = chs_bin_sort_bytes(d, cfg, n_isect, &out->bin_sort_bytes);
    if (st) return st;
  }
  out->bin_sort_bytes += 256;
  uint64_t crf = (uint64_t)3 * chs_crf_stride(cfg->crf_kind, cfg->crf_hidden);
  uint64_t a = crf + d.B, b = (uint64_t)d.C * 12, c = (uint64_t)(n_knots > 0 ? n_knots : 0) * 7 + 2 * (uint64_t)d.B;
  uint64_t m = a > b ? a : b;
  m = m > c ? m : c;
  out->reduce_bytes = chs_align_up(m * sizeof(double), 256) + 256;
  return CHS_OK;
}

// ---------------------------------------------------------------------------------------------
// K10: NCCL all-reduce of the flat gradient buffer.  libnccl is resolved at run time (the copy
// already mapped into the process — e.g. the one PyTorch bundles — or the system one), so libchs
// has no link-time NCCL dependency and cannot clash with the host application's NCCL.
// ---------------------------------------------------------------------------------------------
namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.handle) return CHS_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    chs_set_error("NCCL: cannot dlopen libnccl.so.2: %s", dlerror());
    return CHS_ERR_NCCL;
  }
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
    chs_set_error("NCCL: missing symbols in libnccl.so.2");
    return CHS_ERR_NCCL;
  }
  g_nccl.handle = h;
  return CHS_OK;
}

int nccl_fail(const char* what, ncclResult_t r) {
  chs_set_error("NCCL: %s failed: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  return CHS_ERR_NCCL;
}
}  // namespace

struct chs_comm {
  ncclComm_t comm;
  int rank, world;
};

extern "C" int chs_comm_unique_id(void* unique_id_host_128) {
  CHS_REQUIRE(unique_id_host_128, "chs_comm_unique_id: null output");
  int st = load_nccl();
  if (st) return st;
  ncclUniqueId id;
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r) return nccl_fail("ncclGetUniqueId", r);
  memcpy(unique_id_host_128, &id, 128);
  return CHS_OK;
}

extern "C" int chs_comm_init(const void* unique_id_host_128, int32_t rank, int32_t world, chs_comm** out) {
  CHS_REQUIRE(unique_id_host_128 && out, "chs_comm_init: null pointer");
  CHS_REQUIRE(world >= 1 && rank >= 0 && rank < world, "chs_comm_init: bad rank %d / world %d", rank, world);
  int st = load_nccl();
  if (st) return st;
  ncclUniqueId id;
  memcpy(&id, unique_id_host_128, 128);
  ncclComm_t c;
  ncclResult_t r = g_nccl.CommInitRank(&c, world, id, rank);
  if (r) return nccl_fail("ncclCommInitRank", r);
  chs_comm* h = new chs_comm{c, rank, world};
  *out = h;
  return CHS_OK;
}

extern "C" int chs_allreduce_grads(chs_comm* comm, float* buf, uint64_t count, void* stream) {
  CHS_REQUIRE(comm && buf, "chs_allreduce_grads: null pointer");
  if (count == 0) return CHS_OK;
  // ncclFloat32 = 7, ncclSum = 0
  ncclResult_t r = g_nccl.AllReduce(buf, buf, (size_t)count, 7, 0, comm->comm, (cudaStream_t)stream);
  if (r) return nccl_fail("ncclAllReduce", r);
  return CHS_OK;
}

extern "C" int chs_comm_destroy(chs_comm* comm) {
  if (!comm) return CHS_OK;
  if (g_nccl.CommDestroy) g_nccl.CommDestroy(comm->comm);
  delete comm;
  return CHS_OK;
}

// ---------------------------------------------------------------------------------------------
// K10, NVLS variant: one-shot all-reduce over the NVSwitch multicast mapping (no NCCL involved).
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) nvls_allreduce_kernel(float* mc, uint64_t begin4, uint64_t end4, uint64_t tail_begin,
                                                             uint64_t tail_end) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  // four switch reductions in flight per thread before the first broadcast store: the loop is latency-bound (a multimem.ld_reduce
  // is a round trip through the NVSwitch), so memory-level parallelism per thread is what sets the time
  constexpr int kU = 4;
  for (uint64_t i0 = begin4 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < end4; i0 += stride * kU) {
    float4 v[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const uint64_t i = i0 + u * stride;
      if (i < end4)
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(mc + i * 4) : "memory");
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const uint64_t i = i0 + u * stride;
      if (i < end4)
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc + i * 4), "f"(v[u].x), "f"(v[u].y), "f"(v[u].z), "f"(v[u].w) : "memory");
    }
  }
  // scalar tail (count % 4 floats), owned by the last rank
  for (uint64_t i = tail_begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tail_end; i += stride) {
    float v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(v) : "l"(mc + i) : "memory");
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc + i), "f"(v) : "memory");
  }
}
}  // namespace

extern "C" int chs_nvls_allreduce(float* mc_ptr, uint64_t count, int32_t rank, int32_t world, void* stream) {
  CHS_REQUIRE(mc_ptr, "chs_nvls_allreduce: null multicast pointer (no NVLS multicast support?)");
  CHS_REQUIRE(world >= 1 && rank >= 0 && rank < world, "chs_nvls_allreduce: bad rank %d / world %d", rank, world);
  CHS_REQUIRE(((uintptr_t)mc_ptr) % 16 == 0, "chs_nvls_allreduce: pointer must be 16-byte aligned");
  if (count == 0) return CHS_OK;
  const uint64_t n4 = count / 4;
  const uint64_t per = (n4 + world - 1) / world;
  const uint64_t b4 = (uint64_t)rank * per < n4 ? (uint64_t)rank * per : n4;
  const uint64_t e4 = b4 + per < n4 ? b4 + per : n4;
  const uint64_t tb = rank == world - 1 ? n4 * 4 : count, te = count;
  uint64_t work = (e4 - b4) > (te - tb) ? (e4 - b4) : (te - tb);
  int blocks = (int)((work + 4 * 256 - 1) / (4 * 256));  // four float4 per thread and trip
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  nvls_allreduce_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(mc_ptr, b4, e4, tb, te);
  CHS_LAUNCH_CHECK();
  return CHS_OK;
}

// Broadcast of this rank's slice of a symmetric buffer to every rank through the multicast address (parameter replication:
// each rank uploads 1/G of the parameters from its host and the switch fans the slice out).
namespace {
__global__ void __launch_bounds__(256) nvls_broadcast_kernel(float* mc, const float* __restrict__ local, uint64_t begin, uint64_t end) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t b4 = (begin + 3) / 4, e4 = end / 4;  // whole float4 inside [begin, end)
  for (uint64_t i = b4 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(local)[i];
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc + i * 4), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  }
  // scalar head and tail
  const uint64_t head_end = b4 * 4 < end ? b4 * 4 : end, tail_begin = e4 * 4 > begin ? e4 * 4 : begin;
  for (uint64_t i = begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < head_end; i += stride)
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc + i), "f"(local[i]) : "memory");
  if (e4 >= b4)
    for (uint64_t i = tail_begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += stride)
      asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc + i), "f"(local[i]) : "memory");
}
}  // namespace

extern "C" int chs_nvls_broadcast(float* mc_ptr, const float* local_ptr, uint64_t begin, uint64_t count, void* stream) {
  CHS_REQUIRE(mc_ptr && local_ptr, "chs_nvls_broadcast: null pointer (no NVLS multicast support?)");
  CHS_REQUIRE(((uintptr_t)mc_ptr) % 16 == 0 && ((uintptr_t)local_ptr) % 16 == 0, "chs_nvls_broadcast: pointers must be 16-byte aligned");
  if (count == 0) return CHS_OK;
  int blocks = (int)((count / 4 + 255) / 256);
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  nvls_broadcast_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(mc_ptr, local_ptr, begin, begin + count);
  CHS_LAUNCH_CHECK();
  return CHS_OK;
}

// ---------------------------------------------------------------------------------------------
// one-shot entry points
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void add_inplace_kernel(float* dst, const float* src, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}
}  // namespace

extern "C" int chs_rasterize_fwd(const chs_config* cfg, const chs_tensors* t, int64_t* n_isect_out, void* stream) {
  ChsDims d;
  int st = chs_make_dims(cfg, &d);
  if (st) return st;
  CHS_REQUIRE(t && n_isect_out, "chs_rasterize_fwd: null argument");
  CHS_REQUIRE(!cfg->rgbo_per_camera, "chs_rasterize_fwd: per-camera colour records (SH) need the staged entry points (chs_sh_fwd + chs_blend_fwd)");
  CHS_REQUIRE(t->viewmats && t->workspace, "chs_rasterize_fwd: viewmats / workspace required");
  if (t->spline_kind >= 0) {
    st = chs_spline_fwd(t->spline_kind, t->knots, t->n_knots, t->knot_t0, t->knot_dt, t->frame_times, t->exposure, d.B, d.n,
                        t->viewmats, stream);
    if (st) return st;
  }
  st = chs_project_fwd(cfg, t->means, t->quats, t->scales, t->opacities, t->colors, t->viewmats, t->Ks, t->geom, t->conic_c, t->depths,
                       t->radii, t->tiles_touched, t->rgbo, stream);
  if (st) return st;
  int64_t M = 0;
  st = chs_bin_count(cfg, t->tiles_touched, t->depths, t->isect_offsets, t->order, t->n_isect_dev, &M, t->workspace, t->workspace_bytes,
                     stream);
  *n_isect_out = M;
  if (st) return st;
  if (M > t->isect_capacity) {
    chs_set_error("chs_rasterize_fwd: %lld intersections exceed isect_capacity %lld", (long long)M, (long long)t->isect_capacity);
    return CHS_ERR_WORKSPACE_TOO_SMALL;
  }
  st = chs_bin_sort(cfg, M, t->geom, t->radii, t->depths, t->isect_offsets, t->order, nullptr, t->vals_sorted, t->tile_offsets,
                    t->workspace, t->workspace_bytes, stream);
  if (st) return st;
  return chs_blend_fwd(cfg, t->geom, t->conic_c, t->rgbo, t->vals_sorted, t->tile_offsets, t->exposure, t->crf_params, t->ldr, t->alpha,
                       t->hdr_mean, t->final_T, t->last_id, stream);
}

extern "C" int chs_rasterize_bwd(const chs_config* cfg, const chs_tensors* t, int64_t n_isect, void* stream) {
  ChsDims d;
  int st = chs_make_dims(cfg, &d);
  if (st) return st;
  CHS_REQUIRE(t && t->workspace && t->v_ldr, "chs_rasterize_bwd: null argument");
  CHS_REQUIRE(!cfg->rgbo_per_camera, "chs_rasterize_bwd: per-camera colour records (SH) need the staged entry points");
  (void)n_isect;
  st = chs_crf_bwd(cfg, t->hdr_mean, t->exposure, t->crf_params, t->v_ldr, t->v_hdr, t->v_crf_params, t->v_exposure, t->workspace,
                   t->workspace_bytes, stream);
  if (st) return st;
  st = chs_blend_bwd(cfg, t->geom, t->conic_c, t->rgbo, t->vals_sorted, t->tile_offsets, t->final_T, t->last_id, t->v_hdr, t->v_alpha,
                     t->v_geom, t->v_cogr, t->v_blue, stream);
  if (st) return st;
  st = chs_project_bwd(cfg, t->means, t->quats, t->scales, t->viewmats, t->Ks, t->radii, t->v_geom, t->v_cogr, t->v_blue, t->grads_flat,
                       t->v_viewmats, t->workspace, t->workspace_bytes, stream);
  if (st) return st;
  if (t->spline_kind >= 0) {
    CHS_REQUIRE(t->v_knots && t->v_frame_times && t->v_exposure_window, "chs_rasterize_bwd: spline gradient outputs required");
    st = chs_spline_bwd(t->spline_kind, t->knots, t->n_knots, t->knot_t0, t->knot_dt, t->frame_times, t->exposure, d.B, d.n, t->v_viewmats,
                        t->v_knots, t->v_frame_times, t->v_exposure_window, t->workspace, t->workspace_bytes, stream);
    if (st) return st;
    if (d.B > 0) {
      add_inplace_kernel<<<(d.B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(t->v_exposure, t->v_exposure_window, d.B);
      CHS_LAUNCH_CHECK();
    }
  }
  return CHS_OK;
}
