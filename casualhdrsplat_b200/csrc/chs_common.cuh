// chs_common.cuh — shared host-side plumbing of libchs (error convention, shape bookkeeping)
// and small device helpers (vector loads, warp reductions, vector atomics).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/chs.h"
#include "chs_math.cuh"

// ---------------------------------------------------------------------------------------------
// error convention: every entry point returns a chs_status and leaves a thread-local message
// ---------------------------------------------------------------------------------------------
void chs_set_error(const char* fmt, ...);

#define CHS_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      chs_set_error(__VA_ARGS__);              \
      return CHS_ERR_INVALID_ARG;              \
    }                                          \
  } while (0)

#define CHS_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      chs_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return CHS_ERR_CUDA;                                                                      \
    }                                                                                           \
  } while (0)

// every hand-written kernel launch is followed by CHS_LAUNCH_CHECK(), which also counts it
// (chs_launch_count() — the bench reports it as gpu_launches; CUB and memsets are not counted)
void chs_count_launch();
#define CHS_LAUNCH_CHECK()          \
  do {                              \
    chs_count_launch();             \
    CHS_CUDA(cudaGetLastError());   \
  } while (0)

struct ChsDims {
  int N, B, n, C, W, H;
  int tile_w, tile_h, tiles;
  int tile_bits, cam_bits;
  int64_t CN;  // C * N
  int64_t P;   // W * H
  int Cb;       // cameras of the BINNING stage: C, or B with cfg->pose_fused (one tile list per frame)
  int64_t CbN;  // Cb * N
};

static inline int chs_bit_length(uint64_t v) {
  int b = 0;
  while (v) {
    ++b;
    v >>= 1;
  }
  return b;
}

// Validates the configuration and derives the shape bookkeeping. Returns a chs_status.
int chs_make_dims(const chs_config* cfg, ChsDims* d);

static inline uint64_t chs_align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

// Carves 256-byte aligned pieces out of a caller-provided scratch buffer.
struct ChsArena {
  char* base;
  uint64_t size, used;
  bool ok;
  ChsArena(void* p, uint64_t bytes) : base((char*)p), size(bytes), used(0), ok(true) {}
  template <class T> T* take(uint64_t count) {
    uint64_t bytes = chs_align_up(count * sizeof(T), 256);
    if (used + bytes > size) {
      ok = false;
      return nullptr;
    }
    T* r = (T*)(base + used);
    used += bytes;
    return r;
  }
};

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#define CHS_FULL_MASK 0xffffffffu

// streaming 128-bit load that does not pollute L1 (Gaussian attributes are read once per kernel)
__device__ __forceinline__ float4 chs_ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// 16-byte vector reduction (sm_90+): one L2 atomic transaction for four consecutive floats
__device__ __forceinline__ void chs_red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ float chs_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(CHS_FULL_MASK, v, o);
  return v;
}

__device__ __forceinline__ double chs_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(CHS_FULL_MASK, v, o);
  return v;
}

// Stage `rows` consecutive rows of a row-major [*, 3] fp32 array into shared memory with 128-bit
// coalesced loads: the block's slab starts at row `row0` (a multiple of 4 rows => 16-byte aligned
// when the base pointer is) and holds `rows` valid rows (<= blockDim.x).
__device__ __forceinline__ void chs_stage_rows3(const float* __restrict__ base, int64_t row0, int rows, float* smem) {
  const float* src = base + row0 * 3;
  int total = rows * 3;
  int vec = total >> 2;
  const float4* src4 = reinterpret_cast<const float4*>(src);
  for (int i = threadIdx.x; i < vec; i += blockDim.x) reinterpret_cast<float4*>(smem)[i] = chs_ldg_stream(src4 + i);
  for (int i = (vec << 2) + threadIdx.x; i < total; i += blockDim.x) smem[i] = src[i];
}
#endif
