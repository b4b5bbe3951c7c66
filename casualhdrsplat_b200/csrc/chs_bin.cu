// chs_bin.cu — K2 intersection count, K3 key generation, K4 radix sort / tile multisplit, K5 tile offsets
// (SURVEY.md section 2.4, Appendix A.4).  Integer work, bit-exact against the oracle.
//
// Two sort strategies with identical outputs:
//   CHS_SORT_KEY64          the literal algorithm: one 64-bit key cam|tile|depth per intersection, stable LSD radix sort of
//                           the low 32+tile_bits+cam_bits bits (7 passes of 12-byte pairs at 1080p).
//   CHS_SORT_DEPTH_PRESORT  depth is a property of the (camera, Gaussian) pair, not of the intersection: sort the C*N pairs by
//                           depth inside every camera once (C*N << M), emit intersections in that order, then a *stable*
//                           multisplit on the tile id alone leaves every tile list depth-ordered with ties in Gaussian order
//                           — the same permutation as the 64-bit sort.
// Every pass is hand-written (chs_sort.cuh: count -> column scan -> rank / local reorder / run-wise scatter, MATCH.ANY
// ranking, no look-back spinning).  The presort route never forms a cam|tile key at all: emission order is camera-major, so
// cameras are SEGMENTS of the item array, and the 16-bit tile id is split most-significant digit first:
//   pass 1  inside each camera, by tile >> 8   (a band of 256 consecutive tile ids, ~2 tile rows at 1080p): a tile of 4096
//           consecutive items scatters into the ~32 bands of ONE camera -> ~128-item (512 B) runs, well coalesced;
//   pass 2  inside each (camera, band) bucket, by tile & 255: the destination is the bucket's own ~1 MB region, which stays in L2
//           while its source tiles are processed, so the 16-item runs merge into full lines before they reach HBM.  The pass
//           reads a 1-byte digit + the value and writes the value only.
// The column totals of pass 2 ARE the per-(camera, tile) counts, so tile_offsets is an exclusive scan of them (K5 needs no
// pass over the sorted keys) and every list's final start is known before the last scatter.
// cub::DeviceRadixSort / DeviceScan remain behind chs_config.tune_bin = 1 as the measured baseline (profiles/), never the default.
//
// CHS_SORT_DEPTH_PRESORT has a third, sort-free implementation (opt-in: chs_config.tune_bin = 2): COUNTING PLACEMENT.  Once
// the (camera, Gaussian) pairs are in depth order, a tile list is simply
// "the pairs that touch the tile, in the order they appear" — a stable multisplit into C*tiles
// buckets, for which no keys ever need to exist.  The depth-ordered pairs of a camera are cut into
// chunks; one warp per (camera, chunk, band of tile rows) walks its chunk IN ORDER and counts, in a
// private shared-memory row, how many of its Gaussians touch each tile of the band (count_kernel);
// a column scan over the chunks turns the [C, chunks, tiles] counts into start positions and the
// per-(camera, tile) totals into tile_offsets; the same walk then places every Gaussian id at its
// final position with one shared-memory atomic per intersection (place_kernel).
// Measured on B200 (c3, M = 94.1 M, chunk 4096; profiles/r1g_bin_variants.md): rects 0.08 ms, count
// 0.55 ms, column scan 0.06 ms, place 2.7 ms — slower than either sort route.  The
// count walk is issue-bound (~40 instructions per visited rectangle at 37 % lane use); the place walk
// additionally thrashes L2 with long-lived partially written sectors.  Kept as a bit-exact, tested alternative.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "chs_common.cuh"
#include "chs_sort.cuh"

namespace {

constexpr int kThreads = 256;

struct GatherTouched {
  const int32_t* touched;
  const int32_t* order;
  __host__ __device__ uint32_t operator()(int64_t i) const { return (uint32_t)touched[order ? order[i] : i]; }
};
struct GatherTouched64 {
  const int32_t* touched;
  __host__ __device__ int64_t operator()(int64_t i) const { return (int64_t)touched[i]; }
};

__global__ void depth_keys_kernel(const int32_t* __restrict__ touched, const float* __restrict__ depths, int64_t CN, int N,
                                  uint64_t* __restrict__ keys, int32_t* __restrict__ vals) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= CN) return;
  uint64_t c = (uint64_t)(i / N);
  uint32_t d = touched[i] > 0 ? __float_as_uint(depths[i]) : 0xFFFFFFFFu;
  keys[i] = (c << 32) | d;
  vals[i] = (int32_t)i;
}

// Key emission (K3), warp-cooperative: a warp takes 32 consecutive (camera, Gaussian) entries of the
// emission order; their intersections occupy one contiguous output range (the exclusive scan is in
// the same order), which the 32 lanes write fully coalesced.  The owner of each output slot is found
// by a 5-step binary search over the lanes' start offsets (shuffles), its tile rectangle fetched by
// shuffle.  MODE 0: 64-bit cam|tile|depth keys; MODE 1: linear (cam * tiles + tile) keys; MODE 2: the tile id alone (the camera
// is implied by the item's position: emission order is camera-major).  Slots at or past `cap` (the capacity of the
// intersection buffers when the host sized them without knowing M) are not written.
template <int MODE, class LinT>
__global__ void __launch_bounds__(kThreads) emit_kernel(int64_t CN, int N, int tile_w, int tile_h, int tiles, int tile_bits, int tight, uint32_t cap,
                                                        const float4* __restrict__ geom, const int32_t* __restrict__ radii,
                                                        const float* __restrict__ depths, const uint32_t* __restrict__ offsets,
                                                        const int32_t* __restrict__ order, uint64_t* __restrict__ keys64,
                                                        LinT* __restrict__ keys32, int32_t* __restrict__ vals) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int32_t id = 0;
  int w = 1, cnt = 0;
  uint32_t off = 0, dbits = 0, xy0 = 0, magic = 0;
  if (i < CN) {
    id = order ? order[i] : (int32_t)i;
    off = offsets[i];
    const int radius = radii[id];
    if (radius != 0) {
      const float4 gm = geom[id];
      const ChsTileRect r = chs_tile_bounds_of(gm.x, gm.y, radius, tight, tile_w, tile_h);
      xy0 = (uint32_t)r.x0 | ((uint32_t)r.y0 << 16);
      w = max(r.x1 - r.x0, 1);
      cnt = (r.x1 - r.x0) * (r.y1 - r.y0);
      // ceil(2^32 / w): local / w == umulhi(local, magic) while local * w < 2^32 (local < tiles <= 2^16 here, w <= tile_w);
      // one division per pair instead of one per intersection
      magic = w > 1 ? 0xFFFFFFFFu / (uint32_t)w + 1u : 0u;
      if (MODE == 0) dbits = __float_as_uint(depths[id]);
    }
  }
  // lanes past CN (or culled) own empty ranges that start where the previous lane's range ends
  const uint32_t end_prev = __shfl_up_sync(CHS_FULL_MASK, off + (uint32_t)cnt, 1);
  if (i >= CN && lane > 0) off = end_prev;
  // a tail lane's `off` must be >= every earlier end; propagate with a max-scan over invalid lanes
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o2 = __shfl_up_sync(CHS_FULL_MASK, off + (uint32_t)cnt, d);
    if (i >= CN && lane >= d) off = max(off, o2);
  }
  const uint32_t w_begin = __shfl_sync(CHS_FULL_MASK, off, 0);
  const uint32_t w_end = min(__shfl_sync(CHS_FULL_MASK, off + (uint32_t)cnt, 31), cap);
  const bool use_magic = (uint64_t)tiles * (uint64_t)tile_w < (1ull << 32);
  for (uint32_t o = w_begin + lane; __any_sync(CHS_FULL_MASK, o < w_end); o += 32) {
    int lo_l = 0, hi_l = 31;
#pragma unroll
    for (int it = 0; it < 5; ++it) {
      const int mid = (lo_l + hi_l + 1) >> 1;
      const uint32_t v = __shfl_sync(CHS_FULL_MASK, off, mid);
      if (v <= o) lo_l = mid; else hi_l = mid - 1;
    }
    const uint32_t o_off = __shfl_sync(CHS_FULL_MASK, off, lo_l);
    const uint32_t o_xy = __shfl_sync(CHS_FULL_MASK, xy0, lo_l);
    const int o_w = __shfl_sync(CHS_FULL_MASK, w, lo_l);
    const uint32_t o_mg = __shfl_sync(CHS_FULL_MASK, magic, lo_l);
    const int32_t o_id = __shfl_sync(CHS_FULL_MASK, id, lo_l);
    const uint32_t o_d = MODE == 0 ? __shfl_sync(CHS_FULL_MASK, dbits, lo_l) : 0u;
    if (o < w_end) {
      const uint32_t local = o - o_off;
      const uint32_t ly = o_w > 1 ? (use_magic ? __umulhi(local, o_mg) : local / (uint32_t)o_w) : local;
      const uint32_t lx = local - ly * (uint32_t)o_w;
      const uint32_t tile = ((o_xy >> 16) + ly) * (uint32_t)tile_w + (o_xy & 0xffffu) + lx;
      if (MODE == 0) {
        const uint32_t c = (uint32_t)(o_id / N);
        keys64[o] = ((uint64_t)c << (32 + tile_bits)) | ((uint64_t)tile << 32) | (uint64_t)o_d;
      } else if (MODE == 1) {
        const uint32_t c = (uint32_t)(o_id / N);
        keys32[o] = (LinT)(c * (uint32_t)tiles + tile);
      } else {
        keys32[o] = (LinT)tile;
      }
      vals[o] = o_id;
    }
  }
}

__device__ __forceinline__ uint32_t lin_of_key64(uint64_t key, int tile_bits, int tiles) {
  uint32_t b = (uint32_t)(key >> 32);
  return (b >> tile_bits) * (uint32_t)tiles + (b & ((1u << tile_bits) - 1u));
}

// K5: tile_offsets[lin] = first sorted index whose (cam, tile) >= lin; tile_offsets[C*tiles] = M.
// Four sorted entries per thread.
template <int MODE, class LinT>
__global__ void __launch_bounds__(kThreads) tile_offsets_kernel(int64_t M, const int64_t* __restrict__ m_dev, int n_lin, int tile_bits, int tiles,
                                                                const uint64_t* __restrict__ keys64, const LinT* __restrict__ keys32,
                                                                uint32_t* __restrict__ tile_offsets) {
  if (m_dev) {  // M is then the capacity of the buffers; the live count only exists on the device
    const int64_t live = *m_dev;
    M = live < 0 ? 0 : (live < M ? live : M);
    if (M == 0) {
      for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b <= n_lin; b += (int64_t)gridDim.x * blockDim.x) tile_offsets[b] = 0u;
      return;
    }
  }
  const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i0 >= M) return;
  uint32_t lin[4];
  const int nv = (int)min((int64_t)4, M - i0);
  if (MODE == 0) {
    for (int k = 0; k < nv; ++k) lin[k] = lin_of_key64(keys64[i0 + k], tile_bits, tiles);
  } else if (nv == 4 && sizeof(LinT) == 4) {
    const uint4 v = *reinterpret_cast<const uint4*>(keys32 + i0);
    lin[0] = v.x; lin[1] = v.y; lin[2] = v.z; lin[3] = v.w;
  } else if (nv == 4) {
    const uint2 v = *reinterpret_cast<const uint2*>(keys32 + i0);  // four 16-bit keys
    lin[0] = v.x & 0xffffu; lin[1] = v.x >> 16; lin[2] = v.y & 0xffffu; lin[3] = v.y >> 16;
  } else {
    for (int k = 0; k < nv; ++k) lin[k] = keys32[i0 + k];
  }
  uint32_t prev;
  if (i0 == 0) {
    for (uint32_t b = 0; b <= lin[0]; ++b) tile_offsets[b] = 0;
    prev = lin[0];
  } else {
    prev = MODE == 0 ? lin_of_key64(keys64[i0 - 1], tile_bits, tiles) : keys32[i0 - 1];
  }
  for (int k = 0; k < nv; ++k) {
    for (uint32_t b = prev + 1; b <= lin[k]; ++b) tile_offsets[b] = (uint32_t)(i0 + k);
    prev = lin[k];
  }
  if (i0 + nv == M)
    for (uint32_t b = prev + 1; b <= (uint32_t)n_lin; ++b) tile_offsets[b] = (uint32_t)M;
}

template <class LinT>
__global__ void rebuild_keys_kernel(int64_t M, int tiles, int tile_bits, const LinT* __restrict__ lin_sorted,
                                    const int32_t* __restrict__ vals_sorted, const float* __restrict__ depths,
                                    uint64_t* __restrict__ keys_sorted) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  uint32_t lin = lin_sorted[i];
  uint64_t c = lin / (uint32_t)tiles, t = lin % (uint32_t)tiles;
  keys_sorted[i] = (c << (32 + tile_bits)) | (t << 32) | (uint64_t)__float_as_uint(depths[vals_sorted[i]]);
}

__global__ void store_total_kernel(const int64_t* total, int64_t* n_isect_dev) { *n_isect_dev = *total; }


// ---- counting placement -----------------------------------------------------------------------
constexpr int kPlaceWarps = 8;        // warps per block, one private cursor row each
constexpr int kBandTilesMax = 2048;   // tiles per band: 8 KB of u32 cursors per warp, 64 KB per block

struct PlacePlan {
  bool ok;
  int chunk, n_chunks, band_rows, band_tiles, n_bands;
  uint64_t matrix_elems;  // C * n_chunks * tiles
};

PlacePlan place_plan(const ChsDims& d, int chunk_knob) {
  PlacePlan p;
  memset(&p, 0, sizeof(p));
  if (d.tile_w > kBandTilesMax || d.tile_w >= 65536 || d.tile_h >= 65536 || d.N <= 0 || d.C <= 0) return p;
  p.band_rows = kBandTilesMax / d.tile_w;
  if (p.band_rows > d.tile_h) p.band_rows = d.tile_h;
  p.band_tiles = p.band_rows * d.tile_w;
  p.n_bands = (d.tile_h + p.band_rows - 1) / p.band_rows;
  int chunk = chunk_knob > 0 ? chunk_knob : 4096;  // chs_config.tune_bin_chunk
  const uint64_t budget = (uint64_t)256 << 20;  // bytes of count matrix
  for (;;) {
    p.chunk = chunk;
    p.n_chunks = (d.N + chunk - 1) / chunk;
    p.matrix_elems = (uint64_t)d.C * p.n_chunks * d.tiles;
    if (p.matrix_elems * 4 <= budget || chunk >= (1 << 20)) break;
    chunk *= 2;
  }
  p.ok = p.matrix_elems * 4 <= 2 * budget;
  return p;
}

// tile rectangle of every (camera, Gaussian) pair, in depth order; empty for culled pairs
__global__ void rects_kernel(int64_t CN, int tile_w, int tile_h, int tight, const float4* __restrict__ geom, const int32_t* __restrict__ radii,
                             const int32_t* __restrict__ order, ushort4* __restrict__ rects) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= CN) return;
  const int32_t id = order[i];
  const int radius = radii[id];
  ushort4 r = make_ushort4(0, 0, 0, 0);
  if (radius != 0) {
    const float4 gm = geom[id];
    const ChsTileRect t = chs_tile_bounds_of(gm.x, gm.y, radius, tight, tile_w, tile_h);
    r = make_ushort4((unsigned short)t.x0, (unsigned short)t.y0, (unsigned short)t.x1, (unsigned short)t.y1);
  }
  rects[i] = r;
}

struct PlaceArgs {
  int N, tile_w, tile_h, tiles;
  int chunk, n_chunks, band_rows, band_tiles, n_bands;
  int64_t n_warps;
  const ushort4* rects;         // [C, N] depth order
  const int32_t* order;         // [C, N] flat ids c * N + g in depth order
  uint32_t* matrix;             // [C, n_chunks, tiles] counts, then exclusive prefixes over chunks
  const uint32_t* tile_offsets; // [C * tiles + 1]
  int32_t* vals;                // [M]
};

// One warp = one (camera, chunk, band).  The warp walks the chunk's Gaussians strictly in order — 32 rectangles are
// loaded at once, then visited one by one (ballot + find-first-set), the lanes spreading over the tiles of the
// current rectangle — so the order in which a tile's cursor is bumped IS the depth order.  kPlace = false counts
// (cursor rows start at 0 and are written to the matrix); kPlace = true starts every cursor at the tile's final
// start position and writes the Gaussian id there.
template <bool kPlace>
__global__ void __launch_bounds__(kPlaceWarps * 32) place_kernel(PlaceArgs a) {
  extern __shared__ uint32_t s_rows[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* row = s_rows + warp * a.band_tiles;
  const int64_t wid = (int64_t)blockIdx.x * kPlaceWarps + warp;
  if (wid >= a.n_warps) return;  // whole warps leave together; the kernel has no block-wide barrier
  const int band = (int)(wid % a.n_bands);
  const int64_t rest = wid / a.n_bands;
  const int chunk = (int)(rest % a.n_chunks);
  const int c = (int)(rest / a.n_chunks);
  const int by0 = band * a.band_rows, by1 = min(by0 + a.band_rows, a.tile_h);
  const int n_row = (by1 - by0) * a.tile_w;
  uint32_t* mrow = a.matrix + ((int64_t)c * a.n_chunks + chunk) * a.tiles + (int64_t)by0 * a.tile_w;
  if (kPlace) {
    const uint32_t* toff = a.tile_offsets + (int64_t)c * a.tiles + (int64_t)by0 * a.tile_w;
    for (int i = lane; i < n_row; i += 32) row[i] = toff[i] + mrow[i];
  } else {
    for (int i = lane; i < n_row; i += 32) row[i] = 0u;
  }
  __syncwarp();
  const int p0 = chunk * a.chunk, p1 = min(p0 + a.chunk, a.N);
  const int64_t seg = (int64_t)c * a.N;
  for (int base = p0; base < p1; base += 32) {
    const int pos = base + lane;
    uint32_t packx = 0, packy = 0, magic = 0;
    int32_t id = 0;
    bool hit = false;
    if (pos < p1) {
      const ushort4 rc = a.rects[seg + pos];
      const int y0 = max((int)rc.y, by0), y1 = min((int)rc.w, by1);
      const int w = (int)rc.z - (int)rc.x;
      hit = w > 0 && y1 > y0;
      if (hit) {
        packx = (uint32_t)rc.x | ((uint32_t)w << 16);
        packy = (uint32_t)(y0 - by0) | ((uint32_t)(y1 - y0) << 16);
        magic = w > 1 ? 0xFFFFFFFFu / (uint32_t)w + 1u : 0u;  // ceil(2^32 / w): t / w == umulhi(t, magic) for t * w < 2^32
        if (kPlace) id = a.order[seg + pos];
      }
    }
    unsigned mask = __ballot_sync(CHS_FULL_MASK, hit);
    while (mask) {
      const int src = __ffs(mask) - 1;
      mask &= mask - 1;
      const uint32_t px = __shfl_sync(CHS_FULL_MASK, packx, src);
      const uint32_t py = __shfl_sync(CHS_FULL_MASK, packy, src);
      const uint32_t mg = __shfl_sync(CHS_FULL_MASK, magic, src);
      const int32_t gid = kPlace ? __shfl_sync(CHS_FULL_MASK, id, src) : 0;
      const int x0 = (int)(px & 0xffffu), w = (int)(px >> 16), ry0 = (int)(py & 0xffffu), h = (int)(py >> 16);
      const int n = w * h;
#pragma unroll 1
      for (int t = lane; t < n; t += 32) {
        const int ly = w > 1 ? (int)__umulhi((uint32_t)t, mg) : t;
        const int lx = t - ly * w;
        const int idx = (ry0 + ly) * a.tile_w + x0 + lx;
        const uint32_t o = atomicAdd(&row[idx], 1u);  // the lanes of one rectangle hit distinct tiles
        if (kPlace) a.vals[o] = gid;
      }
      __syncwarp();  // the next rectangle's bumps are ordered after this one's
    }
  }
  if (!kPlace) {
    __syncwarp();
    for (int i = lane; i < n_row; i += 32) mrow[i] = row[i];
  }
}

// counts [C, n_chunks, tiles] -> exclusive prefix over the chunk axis (in place) + per-(camera, tile) totals
__global__ void __launch_bounds__(kThreads) column_scan_kernel(int64_t n_lin, int tiles, int n_chunks, uint32_t* __restrict__ matrix,
                                                               uint32_t* __restrict__ totals) {
  const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (col == 0) totals[n_lin] = 0u;
  if (col >= n_lin) return;
  const int64_t c = col / tiles, tile = col % tiles;
  uint32_t* p = matrix + c * (int64_t)n_chunks * tiles + tile;
  uint32_t run = 0;
  for (int k0 = 0; k0 < n_chunks; k0 += 16) {
    uint32_t v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = (k0 + j < n_chunks) ? p[(int64_t)(k0 + j) * tiles] : 0u;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (k0 + j < n_chunks) {
        p[(int64_t)(k0 + j) * tiles] = run;
        run += v[j];
      }
  }
  totals[col] = run;
}

// keys of the sorted list from (tile_offsets, vals): one block per (camera, tile) bucket
// fused_n > 0 (pose_fused): a list entry frame * N + g is keyed by its depth at the frame's middle pose
__global__ void __launch_bounds__(128) rebuild_keys_from_offsets_kernel(int tiles, int tile_bits, const uint32_t* __restrict__ tile_offsets,
                                                                       const int32_t* __restrict__ vals_sorted,
                                                                       const float* __restrict__ depths, uint64_t* __restrict__ keys_sorted,
                                                                       int fused_n = 0, int N = 0) {
  const uint32_t lin = blockIdx.x;
  const uint64_t c = lin / (uint32_t)tiles, t = lin % (uint32_t)tiles;
  const uint64_t hi = (c << (32 + tile_bits)) | (t << 32);
  const uint32_t e = tile_offsets[lin + 1];
  for (uint32_t i = tile_offsets[lin] + threadIdx.x; i < e; i += blockDim.x) {
    int64_t id = vals_sorted[i];
    if (fused_n > 0) id = ((int64_t)c * fused_n + fused_n / 2) * N + (id - (int64_t)c * N);
    keys_sorted[i] = hi | (uint64_t)__float_as_uint(depths[id]);
  }
}

size_t place_scan_temp(int64_t n) {
  size_t b = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, b, (const uint32_t*)nullptr, (uint32_t*)nullptr, n);
  return b;
}

uint64_t place_bytes(const ChsDims& d, const PlacePlan& p) {
  const int64_t n_lin = (int64_t)d.C * d.tiles;
  return chs_align_up(place_scan_temp(n_lin + 1), 256) + chs_align_up((uint64_t)d.CN * sizeof(ushort4), 256) +
         chs_align_up(p.matrix_elems * 4, 256) + chs_align_up((uint64_t)(n_lin + 1) * 4, 256);
}

inline int grid_for(int64_t n) { return (int)((n + kThreads - 1) / kThreads); }

// CUB temp-storage sizes (need a CUDA context).
struct CountTemp {
  size_t scan, reduce, sort;
};
int count_temp_sizes(const ChsDims& d, int sort_mode, CountTemp* t) {
  t->scan = t->reduce = t->sort = 0;
  GatherTouched gt{nullptr, nullptr};
  cub::TransformInputIterator<uint32_t, GatherTouched, cub::CountingInputIterator<int64_t>> it(cub::CountingInputIterator<int64_t>(0), gt);
  CHS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, t->scan, it, (uint32_t*)nullptr, d.CN));
  GatherTouched64 g64{nullptr};
  cub::TransformInputIterator<int64_t, GatherTouched64, cub::CountingInputIterator<int64_t>> it64(cub::CountingInputIterator<int64_t>(0), g64);
  CHS_CUDA(cub::DeviceReduce::Sum(nullptr, t->reduce, it64, (int64_t*)nullptr, d.CN));
  if (sort_mode == CHS_SORT_DEPTH_PRESORT)
    CHS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t->sort, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const int32_t*)nullptr,
                                             (int32_t*)nullptr, d.CN, 0, 32 + d.cam_bits));
  return CHS_OK;
}

int sort_temp_size(const ChsDims& d, int sort_mode, int64_t M, size_t* bytes) {
  *bytes = 0;
  if (M <= 0) return CHS_OK;
  if (sort_mode == CHS_SORT_KEY64)
    CHS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, *bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const int32_t*)nullptr,
                                             (int32_t*)nullptr, M, 0, 32 + d.tile_bits + d.cam_bits));
  else {
    size_t b16 = 0;
    CHS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, *bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int32_t*)nullptr,
                                             (int32_t*)nullptr, M, 0, chs_bit_length((uint64_t)d.C * d.tiles)));
    CHS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, b16, (const uint16_t*)nullptr, (uint16_t*)nullptr, (const int32_t*)nullptr,
                                             (int32_t*)nullptr, M, 0, 16));
    if (b16 > *bytes) *bytes = b16;
  }
  return CHS_OK;
}

}  // namespace

// ---- baseline route (chs_config.tune_bin = 1): cub::DeviceRadixSort / DeviceScan; tune_bin = 2: counting placement ----
static int cub_bin_count_bytes(const ChsDims& d, int sort_mode, uint64_t* bytes) {
  CountTemp t;
  int st = count_temp_sizes(d, sort_mode, &t);
  if (st) return st;
  uint64_t b = chs_align_up(t.scan, 256) + chs_align_up(t.reduce, 256) + 256;
  if (sort_mode == CHS_SORT_DEPTH_PRESORT)
    b += chs_align_up(t.sort, 256) + 2 * chs_align_up((uint64_t)d.CN * 8, 256) + chs_align_up((uint64_t)d.CN * 4, 256);
  *bytes = b;
  return CHS_OK;
}

static int legacy_bin_sort_bytes(const ChsDims& d, const chs_config* cfg, int64_t M, uint64_t* bytes) {
  const int sort_mode = cfg->sort_mode;
  size_t t;
  int st = sort_temp_size(d, sort_mode, M, &t);
  if (st) return st;
  uint64_t m = (uint64_t)(M > 0 ? M : 0);
  uint64_t b = chs_align_up(t, 256) + chs_align_up(m * 4, 256);  // vals_in
  if (sort_mode == CHS_SORT_KEY64)
    b += 2 * chs_align_up(m * 8, 256);  // keys_in + keys_out (when the caller does not want the keys)
  else
    b += 2 * chs_align_up(m * 4, 256);  // lin_in + lin_out
  if (sort_mode == CHS_SORT_DEPTH_PRESORT) {
    const PlacePlan p = place_plan(d, cfg->tune_bin_chunk);
    if (p.ok) {
      const uint64_t pb = place_bytes(d, p);
      if (pb > b) b = pb;
    }
  }
  *bytes = b;
  return CHS_OK;
}

static int cub_bin_count(const chs_config* cfg, const int32_t* tiles_touched, const float* depths, uint32_t* isect_offsets,
                         int32_t* order, int64_t* n_isect_dev, int64_t* n_isect_host, void* workspace,
                         uint64_t workspace_bytes, void* stream) {
  ChsDims d;
  int st = chs_make_dims(cfg, &d);
  if (st) return st;
  CHS_REQUIRE(tiles_touched && depths && isect_offsets && n_isect_dev && workspace, "chs_bin_count: null pointer");
  CHS_REQUIRE(cfg->sort_mode == CHS_SORT_KEY64 || order, "chs_bin_count: order buffer required for CHS_SORT_DEPTH_PRESORT");
  cudaStream_t s = (cudaStream_t)stream;
  if (d.CN == 0) {
    CHS_CUDA(cudaMemsetAsync(n_isect_dev, 0, sizeof(int64_t), s));
    if (n_isect_host) *n_isect_host = 0;
    return CHS_OK;
  }
  CountTemp t;
  st = count_temp_sizes(d, cfg->sort_mode, &t);
  if (st) return st;
  ChsArena ar(workspace, workspace_bytes);
  char* scan_tmp = ar.take<char>(t.scan);
  char* red_tmp = ar.take<char>(t.reduce);
  int64_t* total = ar.take<int64_t>(1);
  const int32_t* ord = nullptr;
  if (cfg->sort_mode == CHS_SORT_DEPTH_PRESORT) {
    char* sort_tmp = ar.take<char>(t.sort);
    uint64_t* k_in = ar.take<uint64_t>(d.CN);
    uint64_t* k_out = ar.take<uint64_t>(d.CN);
    int32_t* v_in = ar.take<int32_t>(d.CN);
    if (!ar.ok) {
      chs_set_error("chs_bin_count: workspace too small (%llu bytes)", (unsigned long long)workspace_bytes);
      return CHS_ERR_WORKSPACE_TOO_SMALL;
    }
    depth_keys_kernel<<<grid_for(d.CN), kThreads, 0, s>>>(tiles_touched, depths, d.CN, d.N, k_in, v_in);
    CHS_LAUNCH_CHECK();
    size_t tb = t.sort;
    CHS_CUDA(cub::DeviceRadixSort::SortPairs(sort_tmp, tb, (const uint64_t*)k_in, k_out, (const int32_t*)v_in, order, d.CN, 0,
                                             32 + d.cam_bits, s));
    ord = order;
  }
  if (!ar.ok) {
    chs_set_error("chs_bin_count: workspace too small (%llu bytes)", (unsigned long long)workspace_bytes);
    return CHS_ERR_WORKSPACE_TOO_SMALL;
  }
  GatherTouched gt{tiles_touched, ord};
  cub::TransformInputIterator<uint32_t, GatherTouched, cub::CountingInputIterator<int64_t>> it(cub::CountingInputIterator<int64_t>(0), gt);
  size_t sb = t.scan;
  CHS_CUDA(cub::DeviceScan::ExclusiveSum(scan_tmp, sb, it, isect_offsets, d.CN, s));
  GatherTouched64 g64{tiles_touched};
  cub::TransformInputIterator<int64_t, GatherTouched64, cub::CountingInputIterator<int64_t>> it64(cub::CountingInputIterator<int64_t>(0), g64);
  size_t rb = t.reduce;
  CHS_CUDA(cub::DeviceReduce::Sum(red_tmp, rb, it64, total, d.CN, s));
  store_total_kernel<<<1, 1, 0, s>>>(total, n_isect_dev);
  CHS_LAUNCH_CHECK();
  if (n_isect_host) {
    CHS_CUDA(cudaMemcpyAsync(n_isect_host, n_isect_dev, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    CHS_CUDA(cudaStreamSynchronize(s));
    if (*n_isect_host >= ((int64_t)1 << 32) - 1) {
      chs_set_error("chs_bin_count: %lld intersections exceed the 2^32 limit of one launch; split the frame batch",
                    (long long)*n_isect_host);
      return CHS_ERR_UNSUPPORTED;
    }
  }
  return CHS_OK;
}

extern "C" int chs_bin_emit_keys(const chs_config* cfg, int64_t n_isect, const float* geom, const int32_t* radii, const float* depths,
                                 const uint32_t* isect_offsets, const int32_t* order, uint64_t* keys, int32_t* vals, void* stream) {
  ChsDims d;
  int st = chs_make_dims(cfg, &d);
  if (st) return st;
  CHS_REQUIRE(geom && radii && depths && isect_offsets, "chs_bin_emit_keys: null input");
  if (n_isect == 0 || d.CN == 0) return CHS_OK;
  CHS_REQUIRE(keys && vals, "chs_bin_emit_keys: null output");
  emit_kernel<0, uint32_t><<<grid_for(d.CN), kThreads, 0, (cudaStream_t)stream>>>(d.CN, d.N, d.tile_w, d.tile_h, d.tiles, d.tile_bits, cfg->tight_bounds != 0, 0xFFFFFFFFu,
                                                                         (const float4*)geom, radii, depths, isect_offsets, order, keys,
                                                                         nullptr, vals);
  CHS_LAUNCH_CHECK();
  return CHS_OK;
}

static int legacy_bin_sort(const chs_config* cfg, int64_t n_isect, const float* geom, const int32_t* radii, const float* depths,
                           const uint32_t* isect_offsets, const int32_t* order, uint64_t* keys_sorted, int32_t* vals_sorted,
                           uint32_t* tile_offsets, void* workspace, uint64_t workspace_bytes, void* stream) {
  ChsDims d;
  int st = chs_make_dims(cfg, &d);
  if (st) return st;
  CHS_REQUIRE(geom && radii && depths && isect_offsets && tile_offsets, "chs_bin_sort: null pointer");
  CHS_REQUIRE(n_isect >= 0 && n_isect < ((int64_t)1 << 32) - 1, "chs_bin_sort: n_isect out of range");
  CHS_REQUIRE(cfg->sort_mode == CHS_SORT_KEY64 || order, "chs_bin_sort: order required for CHS_SORT_DEPTH_PRESORT");
  cudaStream_t s = (cudaStream_t)stream;
  const int n_lin = d.C * d.tiles;
  const int64_t M = n_isect;
  if (M == 0) {
    CHS_CUDA(cudaMemsetAsync(tile_offsets, 0, ((size_t)n_lin + 1) * sizeof(uint32_t), s));
    return CHS_OK;
  }
  CHS_REQUIRE(vals_sorted && workspace, "chs_bin_sort: null output/workspace");
  if (cfg->sort_mode == CHS_SORT_DEPTH_PRESORT && cfg->tune_bin == 2) {  // counting placement instead of emit + sort
    const PlacePlan p = place_plan(d, cfg->tune_bin_chunk);
    if (p.ok) {
      ChsArena pa(workspace, workspace_bytes);
      size_t scan_b = place_scan_temp((int64_t)n_lin + 1);
      char* scan_tmp = pa.take<char>(scan_b);
      ushort4* rects = pa.take<ushort4>(d.CN);
      uint32_t* matrix = pa.take<uint32_t>(p.matrix_elems);
      uint32_t* totals = pa.take<uint32_t>((uint64_t)n_lin + 1);
      if (!pa.ok) {
        chs_set_error("chs_bin_sort: workspace too small (%llu bytes)", (unsigned long long)workspace_bytes);
        return CHS_ERR_WORKSPACE_TOO_SMALL;
      }
      const size_t smem = (size_t)kPlaceWarps * p.band_tiles * sizeof(uint32_t);
      // up to 64 KB of dynamic shared memory: opt in (per device, so not cached in a static)
      CHS_CUDA(cudaFuncSetAttribute(place_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPlaceWarps * kBandTilesMax * 4));
      CHS_CUDA(cudaFuncSetAttribute(place_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPlaceWarps * kBandTilesMax * 4));
      rects_kernel<<<grid_for(d.CN), kThreads, 0, s>>>(d.CN, d.tile_w, d.tile_h, cfg->tight_bounds != 0, (const float4*)geom, radii, order, rects);
      CHS_LAUNCH_CHECK();
      PlaceArgs a;
      a.N = d.N; a.tile_w = d.tile_w; a.tile_h = d.tile_h; a.tiles = d.tiles;
      a.chunk = p.chunk; a.n_chunks = p.n_chunks; a.band_rows = p.band_rows; a.band_tiles = p.band_tiles; a.n_bands = p.n_bands;
      a.n_warps = (int64_t)d.C * p.n_chunks * p.n_bands;
      a.rects = rects; a.order = order; a.matrix = matrix; a.tile_offsets = tile_offsets; a.vals = vals_sorted;
      const unsigned blocks = (unsigned)((a.n_warps + kPlaceWarps - 1) / kPlaceWarps);
      place_kernel<false><<<blocks, kPlaceWarps * 32, smem, s>>>(a);
      CHS_LAUNCH_CHECK();
      column_scan_kernel<<<grid_for(n_lin), kThreads, 0, s>>>(n_lin, d.tiles, p.n_chunks, matrix, totals);
      CHS_LAUNCH_CHECK();
      CHS_CUDA(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_b, (const uint32_t*)totals, tile_offsets, n_lin + 1, s));
      place_kernel<true><<<blocks, kPlaceWarps * 32, smem, s>>>(a);
      CHS_LAUNCH_CHECK();
      if (keys_sorted) {
        rebuild_keys_from_offsets_kernel<<<n_lin, 128, 0, s>>>(d.tiles, d.tile_bits, tile_offsets, vals_sorted, depths, keys_sorted);
        CHS_LAUNCH_CHECK();
      }
      return CHS_OK;
    }
  }
  size_t tb;
  st = sort_temp_size(d, cfg->sort_mode, M, &tb);
  if (st) return st;
  ChsArena ar(workspace, workspace_bytes);
  char* tmp = ar.take<char>(tb);
  int32_t* v_in = ar.take<int32_t>(M);
  if (cfg->sort_mode == CHS_SORT_KEY64) {
    uint64_t* k_in = ar.take<uint64_t>(M);
    uint64_t* k_out = keys_sorted ? keys_sorted : ar.take<uint64_t>(M);
    if (!ar.ok) {
      chs_set_error("chs_bin_sort: workspace too small (%llu bytes)", (unsigned long long)workspace_bytes);
      return CHS_ERR_WORKSPACE_TOO_SMALL;
    }
    emit_kernel<0, uint32_t><<<grid_for(d.CN), kThreads, 0, s>>>(d.CN, d.N, d.tile_w, d.tile_h, d.tiles, d.tile_bits, cfg->tight_bounds != 0, 0xFFFFFFFFu, (const float4*)geom, radii,
                                                       depths, isect_offsets, nullptr, k_in, nullptr, v_in);
    CHS_LAUNCH_CHECK();
    CHS_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, (const uint64_t*)k_in, k_out, (const int32_t*)v_in, vals_sorted, M, 0,
                                             32 + d.tile_bits + d.cam_bits, s));
    tile_offsets_kernel<0, uint32_t><<<grid_for((M + 3) / 4), kThreads, 0, s>>>(M, nullptr, n_lin, d.tile_bits, d.tiles, k_out, nullptr, tile_offsets);
    CHS_LAUNCH_CHECK();
  } else {
    // linear (cam * tiles + tile) keys: 16 bits suffice for one 1080p frame of 8 poses (65280 buckets), which cuts
    // the bytes every radix pass moves from 8 to 6 per intersection
    const bool k16 = n_lin <= 65536;
    void* l_in = k16 ? (void*)ar.take<uint16_t>(M) : (void*)ar.take<uint32_t>(M);
    void* l_out = k16 ? (void*)ar.take<uint16_t>(M) : (void*)ar.take<uint32_t>(M);
    if (!ar.ok) {
      chs_set_error("chs_bin_sort: workspace too small (%llu bytes)", (unsigned long long)workspace_bytes);
      return CHS_ERR_WORKSPACE_TOO_SMALL;
    }
    const int bits = n_lin > 1 ? chs_bit_length((uint64_t)n_lin - 1) : 1;
    if (k16) {
      emit_kernel<1, uint16_t><<<grid_for(d.CN), kThreads, 0, s>>>(d.CN, d.N, d.tile_w, d.tile_h, d.tiles, d.tile_bits, cfg->tight_bounds != 0, 0xFFFFFFFFu, (const float4*)geom,
                                                                   radii, depths, isect_offsets, order, nullptr, (uint16_t*)l_in, v_in);
      CHS_LAUNCH_CHECK();
      size_t tb16 = tb;
      CHS_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb16, (const uint16_t*)l_in, (uint16_t*)l_out, (const int32_t*)v_in, vals_sorted, M, 0,
                                               bits, s));
      tile_offsets_kernel<1, uint16_t><<<grid_for((M + 3) / 4), kThreads, 0, s>>>(M, nullptr, n_lin, d.tile_bits, d.tiles, nullptr,
                                                                                   (const uint16_t*)l_out, tile_offsets);
      CHS_LAUNCH_CHECK();
      if (keys_sorted) {
        rebuild_keys_kernel<uint16_t><<<grid_for(M), kThreads, 0, s>>>(M, d.tiles, d.tile_bits, (const uint16_t*)l_out, vals_sorted, depths,
                                                                       keys_sorted);
        CHS_LAUNCH_CHECK();
      }
    } else {
      emit_kernel<1, uint32_t><<<grid_for(d.CN), kThreads, 0, s>>>(d.CN, d.N, d.tile_w, d.tile_h, d.tiles, d.tile_bits, cfg->tight_bounds != 0, 0xFFFFFFFFu, (const float4*)geom,
                                                                   radii, depths, isect_offsets, order, nullptr, (uint32_t*)l_in, v_in);
      CHS_LAUNCH_CHECK();
      CHS_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, (const uint32_t*)l_in, (uint32_t*)l_out, (const int32_t*)v_in, vals_sorted, M, 0,
                                               bits, s));
      tile_offsets_kernel<1, uint32_t><<<grid_for((M + 3) / 4), kThreads, 0, s>>>(M, nullptr, n_lin, d.tile_bits, d.tiles, nullptr,
                                                                                   (const uint32_t*)l_out, tile_offsets);
      CHS_LAUNCH_CHECK();
      if (keys_sorted) {
        rebuild_keys_kernel<uint32_t><<<grid_for(M), kThreads, 0, s>>>(M, d.tiles, d.tile_bits, (const uint32_t*)l_out, vals_sorted, depths,
                                                                       keys_sorted);
        CHS_LAUNCH_CHECK();
      }
    }
  }
  return CHS_OK;
}

// =================================================================================================
// hand-written route (default): segmented depth presort, gathered scan, tile multisplit — all from chs_sort.cuh
// =================================================================================================
namespace {

namespace cs = chs_sort;

struct TouchedIn {
  const int32_t* touched;
  const int32_t* order;
  __device__ __forceinline__ uint32_t operator()(uint64_t i) const { return (uint32_t)touched[order ? order[i] : (int64_t)i]; }
};
struct ArrayIn {
  const uint32_t* p;
  __device__ __forceinline__ uint32_t operator()(uint64_t i) const { return p[i]; }
};
struct StoreU32 {
  uint32_t* out;
  __device__ __forceinline__ void operator()(uint64_t i, uint32_t v) const { out[i] = v; }
};

// live item count: the device counter clamped to the capacity of the buffers (or the capacity itself when the host knows M)
__device__ __forceinline__ uint32_t live_items(const int64_t* n_dev, uint32_t cap) {
  if (!n_dev) return cap;
  const int64_t m = *n_dev;
  return m < 0 ? 0u : (m < (int64_t)cap ? (uint32_t)m : cap);
}

// segment lengths of pass 1: camera c owns the items [offsets[c * N], offsets[(c + 1) * N]) of the camera-major emission order
struct CamLen {
  const uint32_t* offsets;
  int N, C;
  const int64_t* n_dev;
  uint32_t cap;
  __device__ __forceinline__ uint32_t begin(int c) const {
    const uint32_t live = live_items(n_dev, cap);
    return c < C ? min(offsets[(int64_t)c * N], live) : live;
  }
  __device__ __forceinline__ uint32_t operator()(int c) const { return begin(c + 1) - begin(c); }
};
// segment lengths of pass 2: bucket (camera c, band b) holds what pass 1 counted for digit b of segment c
struct BucketLen {
  const uint32_t* totals1;
  int nb;
  __device__ __forceinline__ uint32_t operator()(int i) const { return totals1[(int64_t)(i / nb) * cs::kDigits + (i % nb)]; }
};

// single block: seg_begin = exclusive scan of the lengths, tile_first = exclusive scan of ceil(length / tile); entry n_seg = totals
template <class LenFn>
__global__ void __launch_bounds__(1024) segments_kernel(LenFn len_of, int n_seg, uint32_t tile_items, uint32_t* __restrict__ seg_begin,
                                                        uint32_t* __restrict__ tile_first) {
  __shared__ uint64_t s_warp[33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t carry = 0;  // tiles << 40 | items  (items < 2^32, tiles < 2^21)
  for (int b0 = 0; b0 < n_seg; b0 += 1024) {
    const int i = b0 + threadIdx.x;
    const uint32_t len = i < n_seg ? len_of(i) : 0u;
    const uint64_t v = ((uint64_t)((len + tile_items - 1) / tile_items) << 40) | (uint64_t)len;
    uint64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint64_t u = __shfl_up_sync(CHS_FULL_MASK, incl, o);
      if (lane >= o) incl += u;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint64_t run = 0;
      for (int w = 0; w < 32; ++w) {
        const uint64_t c = s_warp[w];
        s_warp[w] = run;
        run += c;
      }
      s_warp[32] = run;
    }
    __syncthreads();
    if (i < n_seg) {
      const uint64_t e = carry + s_warp[warp] + incl - v;
      seg_begin[i] = (uint32_t)(e & 0xFFFFFFFFFFull);
      tile_first[i] = (uint32_t)(e >> 40);
    }
    carry += s_warp[32];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    seg_begin[n_seg] = (uint32_t)(carry & 0xFFFFFFFFFFull);
    tile_first[n_seg] = (uint32_t)(carry >> 40);
  }
}

// K5 without a pass over sorted keys: list starts of every (camera, tile) from the scanned column totals of pass 2
__global__ void __launch_bounds__(kThreads) tile_offsets_from_base_kernel(int64_t n_lin, int tiles, int nb, const uint32_t* __restrict__ base2,
                                                                           const uint32_t* __restrict__ seg_begin2, int n_seg2,
                                                                           uint32_t* __restrict__ tile_offsets) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_lin) {
    const int64_t c = i / tiles;
    const int t = (int)(i - c * tiles);
    tile_offsets[i] = base2[(c * nb + (t >> 8)) * cs::kDigits + (t & 255)];
  } else if (i == n_lin) {
    tile_offsets[n_lin] = seg_begin2[n_seg2];
  }
}

// K5 of the LSD routes when M only exists on the device: a one-thread shim that forwards it to tile_offsets_kernel's argument
// is not possible, so that kernel reads the live count itself (m_dev) and uses `M` as the capacity.

template <class Keys>
int launch_count(const cs::TileMap& map, Keys keys, int shift, uint32_t* counts, uint32_t t_cap, cudaStream_t s) {
  cs::radix_count_kernel<Keys><<<t_cap, cs::kThreads, 0, s>>>(map, keys, shift, counts, t_cap);
  CHS_LAUNCH_CHECK();
  return CHS_OK;
}
int launch_colscan(const cs::TileMap& map, uint32_t* counts, uint32_t t_cap, uint32_t* totals, cudaStream_t s) {
  const int64_t cols = (int64_t)map.n_seg * cs::kDigits;
  cs::radix_colscan_kernel<<<(unsigned)((cols + cs::kWarps - 1) / cs::kWarps), cs::kThreads, 0, s>>>(map, counts, t_cap, totals);
  CHS_LAUNCH_CHECK();
  return CHS_OK;
}
// nbits <= 5 selects the instantiation with five ballots per round (digits < 32: the tile bands of one camera)
template <class Keys, class KeyOutT>
int launch_scatter(const cs::ScatterArgs<Keys, KeyOutT>& a, int nbits, cudaStream_t s) {
  typedef typename Keys::key_type KeyT;
  const size_t smem = cs::scatter_smem_bytes<KeyT>();
  if (nbits <= 5) {
    if (smem > 48 * 1024)
      CHS_CUDA(cudaFuncSetAttribute(cs::radix_scatter_kernel<5, Keys, KeyOutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cs::radix_scatter_kernel<5, Keys, KeyOutT><<<a.t_cap, cs::kThreads, smem, s>>>(a);
  } else {
    if (smem > 48 * 1024)
      CHS_CUDA(cudaFuncSetAttribute(cs::radix_scatter_kernel<8, Keys, KeyOutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cs::radix_scatter_kernel<8, Keys, KeyOutT><<<a.t_cap, cs::kThreads, smem, s>>>(a);
  }
  CHS_LAUNCH_CHECK();
  return CHS_OK;
}

// one stable pass by the 8-bit digit at `shift`: count -> column scan -> scatter
template <class Keys, class KeyOutT>
int radix_pass(const cs::TileMap& map, Keys keys, const int32_t* vals_in, KeyOutT* keys_out, int32_t* vals_out, int shift, int nbits,
               uint32_t* counts, uint32_t t_cap, uint32_t* totals, const uint32_t* digit_base, cudaStream_t s) {
  int st = launch_count(map, keys, shift, counts, t_cap, s);
  if (st) return st;
  st = launch_colscan(map, counts, t_cap, totals, s);
  if (st) return st;
  cs::ScatterArgs<Keys, KeyOutT> a;
  a.map = map; a.keys = keys; a.vals_in = vals_in; a.keys_out = keys_out; a.vals_out = vals_out; a.shift = shift;
  a.prefix = counts; a.t_cap = t_cap; a.totals = totals; a.digit_base = digit_base;
  return launch_scatter(a, nbits, s);
}

template <class In, class Out>
int exclusive_scan(In in, uint64_t n, uint64_t* sums, Out out, int64_t* total_out, cudaStream_t s) {
  const uint32_t blocks = (uint32_t)((n + cs::kScanTile - 1) / cs::kScanTile);
  if (blocks == 0) {
    if (total_out) CHS_CUDA(cudaMemsetAsync(total_out, 0, sizeof(int64_t), s));
    return CHS_OK;
  }
  cs::scan_sums_kernel<In><<<blocks, cs::kThreads, 0, s>>>(in, n, sums);
  CHS_LAUNCH_CHECK();
  cs::scan_of_sums_kernel<<<1, 1024, 0, s>>>(sums, blocks, total_out, nullptr);
  CHS_LAUNCH_CHECK();
  cs::scan_final_kernel<In, Out><<<blocks, cs::kThreads, 0, s>>>(in, n, sums, out);
  CHS_LAUNCH_CHECK();
  return CHS_OK;
}


// =================================================================================================
// Banded placement (default route of CHS_SORT_DEPTH_PRESORT): the tile lists without emitting or sorting intersections.
//
// With the pairs of a camera in depth order, the list of tile t is "the pairs whose rectangle covers t, in that order".  The
// rank of a pair in a list is a COUNT of earlier pairs, and counting over a fixed chunk is a population count over a bit
// matrix: bit p of row t says "pair p of the chunk covers t".  Two levels keep the bit matrices in shared memory:
//   P1  band split.  A band is BR consecutive tile rows with BR * tile_w <= 256 tiles (two rows at 1080p).  For a chunk of 1024
//       depth-ordered pairs the kernel sets bm[band][pair] for every band a pair's rectangle reaches (1.9 per pair at c3), takes
//       per-band word prefixes, and writes one 8-byte record (pair id, sub-rectangle inside the band) per (pair, band) to the
//       band's bucket at  bucket start + chunk prefix + rank.  ~15 M records instead of 62.6 M intersections.
//   P2  tile placement.  A chunk of 1024 records of ONE bucket sets bm[tile][record] for the <= 256 tiles of the band, takes
//       per-tile word prefixes, and writes the pair id of every (record, tile) straight to its final position
//       tile_offsets[tile] + chunk prefix + rank.  The destination of a bucket is its own ~1 MB region, which stays in L2 while
//       the bucket's chunks are processed.
// Chunk prefixes come from a count kernel + the column scan of chs_sort.cuh at each level; the column totals of P2 are the
// per-(camera, tile) counts, so tile_offsets is their exclusive scan.  Nothing is keyed, ranked by ballots or moved twice:
// per intersection the work is one shared-memory atomicOr, two shared loads, a popc and one 4-byte store.
// =================================================================================================
constexpr int kBandChunk = 1024;  // pairs per P1 chunk = records per P2 chunk = 32 words of a bit-matrix row
constexpr int kBandWords = kBandChunk / 32;
constexpr int kBandRow = kBandWords + 1;  // padded row: the per-row prefix walk is bank-conflict free

struct BandGeom {
  int ok, BR, band_tiles, n_bands;
};
BandGeom band_geom(const ChsDims& d) {
  BandGeom g;
  g.ok = d.tile_w >= 1 && d.tile_w <= 256;
  g.BR = 1; g.band_tiles = d.tile_w; g.n_bands = d.tile_h;
  if (!g.ok) return g;
  g.BR = 256 / d.tile_w;
  if (g.BR > d.tile_h) g.BR = d.tile_h;
  if (g.BR < 1) g.BR = 1;
  g.band_tiles = g.BR * d.tile_w;
  g.n_bands = (d.tile_h + g.BR - 1) / g.BR;
  g.ok = g.n_bands <= 256;
  return g;
}

bool uses_band_route(const chs_config* cfg, const ChsDims& d) {
  return cfg->sort_mode == CHS_SORT_DEPTH_PRESORT && cfg->tune_bin == 0 && band_geom(d).ok;
}

struct BandArgs {
  int N, C, tile_w, tile_h, tight, BR, n_bands, chunks1;  // C = binning cameras (frames when fused); chunks1 = chunks per camera in P1
  int fused, n_virtual;      // pose_fused: pair (frame f, g) covers the union of the rectangles of cameras f * n_virtual + k
  uint32_t t_cap1, t_cap2, rec_cap, val_cap;
  const float4* geom;
  const int32_t* radii;
  const int32_t* order;
  ushort4* rects;            // [C * N] depth order (written by the P1 count, read by the P1 scatter)
  uint32_t* counts1;         // [256][t_cap1]
  const uint32_t* seg_begin2;  // [C * n_bands + 1] bucket starts in the record array
  uint2* recs;               // [rec_cap] (pair id, box)
  uint4* desc;               // [t_cap2] P2 chunks: (bucket, begin, end, 0); bucket = 0xffffffff past the end
  uint32_t* counts2;         // [256][t_cap2]
  const uint32_t* base2;     // [C * n_bands][256] final list starts
  int32_t* vals;             // [val_cap]
};

// box of a record: x0 | (x1 - 1) << 8 | ry0 << 16 | (ry1 - 1) << 24, rows relative to the band
__device__ __forceinline__ uint32_t make_box(int x0, int x1, int ry0, int ry1) {
  return (uint32_t)x0 | ((uint32_t)(x1 - 1) << 8) | ((uint32_t)ry0 << 16) | ((uint32_t)(ry1 - 1) << 24);
}

// P1 count: bands reached by the chunk's pairs -> counts1[band][chunk]; also leaves every pair's rectangle in depth order
__global__ void __launch_bounds__(256) band_count_kernel(BandArgs a) {
  __shared__ uint32_t hist[256];
  const int t = blockIdx.x;  // = c * chunks1 + q
  const int c = t / a.chunks1, q = t - c * a.chunks1;
  hist[threadIdx.x] = 0u;
  __syncthreads();
  const int64_t seg = (int64_t)c * a.N;
#pragma unroll
  for (int k = 0; k < kBandChunk / 256; ++k) {
    const int i = q * kBandChunk + k * 256 + threadIdx.x;
    if (i >= a.N) continue;
    const int32_t id = a.order[seg + i];
    ChsTileRect r;
    r.x0 = r.y0 = r.x1 = r.y1 = 0;
    if (!a.fused) {
      const int radius = a.radii[id];
      if (radius != 0) {
        const float4 gm = a.geom[id];
        r = chs_tile_bounds_of(gm.x, gm.y, radius, a.tight, a.tile_w, a.tile_h);
      }
    } else {  // union over the frame's poses that see the Gaussian
      const int64_t first = ((int64_t)c * a.n_virtual) * a.N + (id - (int32_t)seg);
      for (int k = 0; k < a.n_virtual; ++k) {
        const int64_t cid = first + (int64_t)k * a.N;
        const int radius = a.radii[cid];
        if (radius == 0) continue;
        const float4 gm = a.geom[cid];
        const ChsTileRect q2 = chs_tile_bounds_of(gm.x, gm.y, radius, a.tight, a.tile_w, a.tile_h);
        if (q2.x1 <= q2.x0 || q2.y1 <= q2.y0) continue;
        if (r.x1 > r.x0) {
          r.x0 = min(r.x0, q2.x0); r.y0 = min(r.y0, q2.y0); r.x1 = max(r.x1, q2.x1); r.y1 = max(r.y1, q2.y1);
        } else {
          r = q2;
        }
      }
    }
    ushort4 rc = make_ushort4(0, 0, 0, 0);
    if (r.x1 > r.x0 && r.y1 > r.y0) {
      rc = make_ushort4((unsigned short)r.x0, (unsigned short)r.y0, (unsigned short)r.x1, (unsigned short)r.y1);
      const int b1 = (r.y1 - 1) / a.BR;
      for (int b = r.y0 / a.BR; b <= b1; ++b) atomicAdd(&hist[b], 1u);
    }
    a.rects[seg + i] = rc;
  }
  __syncthreads();
  a.counts1[(size_t)threadIdx.x * a.t_cap1 + t] = hist[threadIdx.x];
}

// P1 scatter: one record per (pair, band) at bucket start + chunk prefix + rank
__global__ void __launch_bounds__(256) band_scatter_kernel(BandArgs a) {
  extern __shared__ uint32_t smem_band[];
  uint32_t* bm = smem_band;                        // [n_bands][kBandRow]
  uint32_t* pfx = bm + a.n_bands * kBandRow;       // [n_bands][kBandRow]
  uint32_t* base = pfx + a.n_bands * kBandRow;     // [n_bands]
  const int t = blockIdx.x;
  const int c = t / a.chunks1, q = t - c * a.chunks1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < a.n_bands * kBandRow; i += 256) bm[i] = 0u;
  for (int b = threadIdx.x; b < a.n_bands; b += 256) base[b] = a.seg_begin2[c * a.n_bands + b] + a.counts1[(size_t)b * a.t_cap1 + t];
  __syncthreads();
  const int64_t seg = (int64_t)c * a.N;
  ushort4 rc[kBandChunk / 256];
#pragma unroll
  for (int k = 0; k < kBandChunk / 256; ++k) {
    const int i = q * kBandChunk + k * 256 + threadIdx.x;  // pair p = k * 256 + tid of the chunk: word p >> 5 = k * 8 + warp, bit = lane
    rc[k] = i < a.N ? a.rects[seg + i] : make_ushort4(0, 0, 0, 0);
    if (rc[k].z > rc[k].x) {
      const int b1 = ((int)rc[k].w - 1) / a.BR;
      for (int b = (int)rc[k].y / a.BR; b <= b1; ++b) atomicOr(&bm[b * kBandRow + k * 8 + warp], 1u << lane);
    }
  }
  __syncthreads();
  for (int b = warp; b < a.n_bands; b += 8) {  // exclusive prefix of the row's word populations
    const uint32_t cnt = __popc(bm[b * kBandRow + lane]);
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(CHS_FULL_MASK, incl, o);
      if (lane >= o) incl += u;
    }
    pfx[b * kBandRow + lane] = incl - cnt;
  }
  __syncthreads();
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int k = 0; k < kBandChunk / 256; ++k) {
    if (rc[k].z <= rc[k].x) continue;
    const int i = q * kBandChunk + k * 256 + threadIdx.x;
    const int32_t id = a.order[seg + i];
    const int w = k * 8 + warp;
    const int y0 = rc[k].y, y1 = rc[k].w;
    const int b1 = (y1 - 1) / a.BR;
    for (int b = y0 / a.BR; b <= b1; ++b) {
      const uint32_t rank = pfx[b * kBandRow + w] + __popc(bm[b * kBandRow + w] & lt);
      const uint32_t dst = base[b] + rank;
      const int r0 = b * a.BR;
      if (dst < a.rec_cap) a.recs[dst] = make_uint2((uint32_t)id, make_box(rc[k].x, rc[k].z, max(y0, r0) - r0, min(y1, r0 + a.BR) - r0));
    }
  }
}

// P2 chunk descriptors: chunk t -> (bucket, first record, end record); one thread per chunk does the table search once
__global__ void __launch_bounds__(256) chunk_desc_kernel(int n_seg2, const uint32_t* __restrict__ seg_begin2, const uint32_t* __restrict__ tile_first2,
                                                         uint32_t t_cap2, uint32_t rec_cap, uint4* __restrict__ desc) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= t_cap2) return;
  uint4 r = make_uint4(0xffffffffu, 0u, 0u, 0u);
  if (t < tile_first2[n_seg2]) {
    int lo = 0, hi = n_seg2 - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (tile_first2[mid] <= t) lo = mid; else hi = mid - 1;
    }
    const uint64_t b = (uint64_t)seg_begin2[lo] + (uint64_t)(t - tile_first2[lo]) * kBandChunk;
    const uint64_t e = min((uint64_t)seg_begin2[lo + 1], b + kBandChunk);
    r = make_uint4((uint32_t)lo, (uint32_t)min(b, (uint64_t)rec_cap), (uint32_t)min(e, (uint64_t)rec_cap), 0u);
  }
  desc[t] = r;
}

// P2 count: tiles covered by the chunk's records -> counts2[tile in band][chunk]
__global__ void __launch_bounds__(256) tile_count_kernel(BandArgs a) {
  __shared__ uint32_t hist[256];
  const uint32_t t = blockIdx.x;
  const uint4 ds = a.desc[t];
  hist[threadIdx.x] = 0u;
  __syncthreads();
  if (ds.x != 0xffffffffu) {
#pragma unroll
    for (int k = 0; k < kBandChunk / 256; ++k) {
      const uint32_t i = ds.y + k * 256 + threadIdx.x;
      if (i >= ds.z) continue;
      const uint32_t box = a.recs[i].y;
      const int x0 = box & 0xff, x1 = (box >> 8) & 0xff, ry0 = (box >> 16) & 0xff, ry1 = box >> 24;
      for (int ry = ry0; ry <= ry1; ++ry)
        for (int x = x0; x <= x1; ++x) atomicAdd(&hist[ry * a.tile_w + x], 1u);
    }
  }
  __syncthreads();
  a.counts2[(size_t)threadIdx.x * a.t_cap2 + t] = hist[threadIdx.x];
}

// P2 scatter: every (record, tile) of the chunk straight to its final position in the tile's list.  Record-parallel threads set
// bm[tile][record]; then the roles swap: thread = tile of the band walks its row of the bit matrix in record (= depth) order and
// appends the pair id of every set bit at the tile's cursor.  The second pass touches exactly the chunk's intersections, is
// balanced across the tiles of a band, and needs neither rank arithmetic nor a second expansion of the rectangles.
__global__ void __launch_bounds__(256) tile_scatter_kernel(BandArgs a) {
  extern __shared__ uint32_t smem_tile[];
  uint32_t* bm = smem_tile;                    // [256][kBandRow]
  uint32_t* ids = bm + 256 * kBandRow;         // [kBandChunk] pair id of every record of the chunk
  const uint32_t t = blockIdx.x;
  const uint4 ds = a.desc[t];
  if (ds.x == 0xffffffffu) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t cursor = a.base2[(size_t)ds.x * 256 + threadIdx.x] + a.counts2[(size_t)threadIdx.x * a.t_cap2 + t];
  for (int i = threadIdx.x; i < 256 * kBandRow; i += 256) bm[i] = 0u;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kBandChunk / 256; ++k) {
    const uint32_t i = ds.y + k * 256 + threadIdx.x;  // record p = k * 256 + tid of the chunk: word k * 8 + warp, bit = lane
    if (i < ds.z) {
      const uint2 rec = a.recs[i];
      ids[k * 256 + threadIdx.x] = rec.x;
      const uint32_t box = rec.y;
      const int x0 = box & 0xff, x1 = (box >> 8) & 0xff, ry0 = (box >> 16) & 0xff, ry1 = box >> 24;
      for (int ry = ry0; ry <= ry1; ++ry)
        for (int x = x0; x <= x1; ++x) atomicOr(&bm[(ry * a.tile_w + x) * kBandRow + k * 8 + warp], 1u << lane);
    }
  }
  __syncthreads();
  const uint32_t* row = bm + threadIdx.x * kBandRow;  // padded rows: the 32 lanes of a warp read 32 different banks
#pragma unroll 4
  for (int w = 0; w < kBandWords; ++w) {
    uint32_t m = row[w];
    while (m) {
      const int b = __ffs(m) - 1;
      m &= m - 1;
      if (cursor < a.val_cap) a.vals[cursor] = (int32_t)ids[w * 32 + b];
      ++cursor;
    }
  }
}

// P2 scatter, warp-per-tile output (chs_config.tune_bin_chunk = 1 on the banded route).  Same bit matrix; the walk differs: a warp
// takes one tile at a time, lane w owns word w of the tile's row, a warp scan of the word populations gives every lane its
// offset in the tile's run, and the lanes append their set bits side by side.  All addresses of one store instruction then fall
// into the tile's run of this chunk (a few sectors) instead of 32 different lists (32 sectors).
__global__ void __launch_bounds__(256) tile_scatter_warp_kernel(BandArgs a) {
  extern __shared__ uint32_t smem_tile[];
  uint32_t* bm = smem_tile;                    // [256][kBandRow]
  uint32_t* ids = bm + 256 * kBandRow;         // [kBandChunk] pair id of every record of the chunk
  uint32_t* cur0 = ids + kBandChunk;           // [256] start of every tile's run
  const uint32_t t = blockIdx.x;
  const uint4 ds = a.desc[t];
  if (ds.x == 0xffffffffu) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  cur0[threadIdx.x] = a.base2[(size_t)ds.x * 256 + threadIdx.x] + a.counts2[(size_t)threadIdx.x * a.t_cap2 + t];
  for (int i = threadIdx.x; i < 256 * kBandRow; i += 256) bm[i] = 0u;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kBandChunk / 256; ++k) {
    const uint32_t i = ds.y + k * 256 + threadIdx.x;  // record p = k * 256 + tid of the chunk: word k * 8 + warp, bit = lane
    if (i < ds.z) {
      const uint2 rec = a.recs[i];
      ids[k * 256 + threadIdx.x] = rec.x;
      const uint32_t box = rec.y;
      const int x0 = box & 0xff, x1 = (box >> 8) & 0xff, ry0 = (box >> 16) & 0xff, ry1 = box >> 24;
      for (int ry = ry0; ry <= ry1; ++ry)
        for (int x = x0; x <= x1; ++x) atomicOr(&bm[(ry * a.tile_w + x) * kBandRow + k * 8 + warp], 1u << lane);
    }
  }
  __syncthreads();
  static_assert(kBandWords == 32, "one lane per word of a bit-matrix row");
  const int n_tiles = a.BR * a.tile_w;
  for (int tl = warp; tl < n_tiles; tl += 8) {
    uint32_t m = bm[tl * kBandRow + lane];
    if (!__any_sync(CHS_FULL_MASK, m != 0u)) continue;
    const uint32_t cnt = __popc(m);
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(CHS_FULL_MASK, incl, o);
      if (lane >= o) incl += u;
    }
    uint32_t cursor = cur0[tl] + incl - cnt;
    const uint32_t* my_ids = ids + lane * 32;
    while (m) {
      const int b = __ffs(m) - 1;
      m &= m - 1;
      if (cursor < a.val_cap) a.vals[cursor] = (int32_t)my_ids[b];
      ++cursor;
    }
  }
}

// K5 of the banded route: list starts of every (camera, tile) from the scanned column totals of P2 (clamped to the capacity of
// the value buffer, so that a step whose M outgrew it walks truncated lists instead of reading past the end)
__global__ void __launch_bounds__(kThreads) band_tile_offsets_kernel(int64_t n_lin, int tiles, int tile_w, int BR, int n_bands,
                                                                      const uint32_t* __restrict__ base2, const uint64_t* __restrict__ total,
                                                                      uint32_t val_cap, uint32_t* __restrict__ tile_offsets) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_lin) {
    const int64_t c = i / tiles;
    const int t = (int)(i - c * tiles);
    const int b = (t / tile_w) / BR;
    tile_offsets[i] = min(base2[(c * n_bands + b) * 256 + (t - b * BR * tile_w)], val_cap);
  } else if (i == n_lin) {
    const uint64_t m = *total;
    tile_offsets[n_lin] = m < (uint64_t)val_cap ? (uint32_t)m : val_cap;
  }
}

struct BandPlan {
  BandGeom g;
  int chunks1;
  int64_t n_seg2;
  uint32_t t_cap1, t_cap2;
  uint64_t rec_cap, sum_blocks;
};
BandPlan band_plan(const ChsDims& d, uint64_t cap) {
  BandPlan p;
  memset(&p, 0, sizeof(p));
  p.g = band_geom(d);
  p.chunks1 = (d.N + kBandChunk - 1) / kBandChunk;
  p.t_cap1 = (uint32_t)p.chunks1 * (uint32_t)d.Cb;
  p.n_seg2 = (int64_t)d.Cb * p.g.n_bands;
  // a pair whose rectangle spans r tile rows reaches at most r / BR + 2 bands and has at least r intersections, so the number of
  // (pair, band) records is bounded by M / BR + 2 C N (and by M): the bound sizes the record array and the P2 grid
  p.rec_cap = cap / (uint64_t)p.g.BR + 2 * (uint64_t)d.CbN;
  if (p.rec_cap > cap) p.rec_cap = cap;
  p.t_cap2 = (uint32_t)((p.rec_cap + kBandChunk - 1) / kBandChunk) + (uint32_t)p.n_seg2;
  p.sum_blocks = ((uint64_t)p.n_seg2 * 256 + cs::kScanTile - 1) / cs::kScanTile;
  return p;
}
uint64_t band_bytes(const ChsDims& d, uint64_t cap) {
  const BandPlan p = band_plan(d, cap);
  uint64_t b = chs_align_up((uint64_t)d.CbN * sizeof(ushort4), 256) + chs_align_up((uint64_t)256 * p.t_cap1 * 4, 256);
  b += chs_align_up((uint64_t)d.Cb * 256 * 4, 256) + 2 * chs_align_up((uint64_t)(p.n_seg2 + 1) * 4, 256);
  b += chs_align_up(p.rec_cap * sizeof(uint2), 256) + chs_align_up((uint64_t)p.t_cap2 * sizeof(uint4), 256);
  b += chs_align_up((uint64_t)256 * p.t_cap2 * 4, 256) + 2 * chs_align_up((uint64_t)p.n_seg2 * 256 * 4, 256);
  b += chs_align_up((p.sum_blocks + 1) * 8, 256);
  return b + 256;
}

inline uint32_t tiles_of(uint64_t n) { return (uint32_t)((n + cs::kTile - 1) / cs::kTile); }

// ---- workspace layouts (one definition for the size query and for the launch) ----
struct CountPlan {
  uint32_t tps, t_cap;
  uint64_t sum_blocks;
};
CountPlan count_plan(const ChsDims& d) {
  CountPlan p;
  p.tps = tiles_of((uint64_t)d.N);
  p.t_cap = p.tps * (uint32_t)d.Cb;
  p.sum_blocks = ((uint64_t)d.CbN + cs::kScanTile - 1) / cs::kScanTile;
  return p;
}
uint64_t hw_count_bytes(const ChsDims& d, int sort_mode) {
  const CountPlan p = count_plan(d);
  uint64_t b = chs_align_up(p.sum_blocks * 8, 256);
  if (sort_mode == CHS_SORT_DEPTH_PRESORT)
    b += chs_align_up((uint64_t)cs::kDigits * p.t_cap * 4, 256) + chs_align_up((uint64_t)d.Cb * cs::kDigits * 4, 256) +
         4 * chs_align_up((uint64_t)d.CbN * 4, 256);
  return b + 256;
}

struct SortPlan {
  bool multisplit;   // presort with tile ids that fit 16 bits: the two-pass segmented route
  int nb;            // bands of 256 tile ids per camera
  int64_t n_seg2;    // C * nb
  int passes;        // LSD routes: number of 8-bit passes
  int key_bytes;     // LSD routes: 8 (cam|tile|depth) or 4 (linear keys)
  uint32_t t_cap;
  uint64_t sum_blocks;
};
SortPlan sort_plan(const ChsDims& d, int sort_mode, uint64_t cap) {
  SortPlan p;
  memset(&p, 0, sizeof(p));
  p.multisplit = sort_mode == CHS_SORT_DEPTH_PRESORT && d.tiles <= 65536;
  if (p.multisplit) {
    p.nb = (d.tiles + 255) / 256;
    p.n_seg2 = (int64_t)d.C * p.nb;
    p.t_cap = tiles_of(cap) + (uint32_t)p.n_seg2;
    p.sum_blocks = ((uint64_t)p.n_seg2 * cs::kDigits + cs::kScanTile - 1) / cs::kScanTile;
  } else {
    p.key_bytes = sort_mode == CHS_SORT_KEY64 ? 8 : 4;
    const int bits = sort_mode == CHS_SORT_KEY64 ? 32 + d.tile_bits + d.cam_bits
                                                 : (d.C * (int64_t)d.tiles > 1 ? chs_bit_length((uint64_t)d.C * d.tiles - 1) : 1);
    p.passes = (bits + 7) / 8;
    p.t_cap = tiles_of(cap);
  }
  return p;
}
uint64_t hw_sort_bytes(const ChsDims& d, int sort_mode, uint64_t cap) {
  const SortPlan p = sort_plan(d, sort_mode, cap);
  uint64_t b = chs_align_up((uint64_t)cs::kDigits * p.t_cap * 4, 256);
  if (p.multisplit) {
    b += chs_align_up(cap * 2, 256) + 2 * chs_align_up(cap * 4, 256) + chs_align_up(cap, 256);       // tile keys, vals x2, digits
    b += chs_align_up((uint64_t)d.C * cs::kDigits * 4, 256) + 2 * chs_align_up((uint64_t)(d.C + 1) * 4, 256);  // totals1, tables1
    b += 2 * chs_align_up((uint64_t)p.n_seg2 * cs::kDigits * 4, 256) + 2 * chs_align_up((uint64_t)(p.n_seg2 + 1) * 4, 256);
    b += chs_align_up(p.sum_blocks * 8, 256);
  } else {
    b += 2 * chs_align_up(cap * p.key_bytes, 256) + chs_align_up(cap * 4, 256) + chs_align_up((uint64_t)cs::kDigits * 4, 256);
  }
  return b + 256;
}

int hw_bin_count(const chs_config* cfg, const ChsDims& d, const int32_t* tiles_touched, const float* depths, uint32_t* isect_offsets,
                 int32_t* order, int64_t* n_isect_dev, void* workspace, uint64_t workspace_bytes, cudaStream_t s) {
  const CountPlan p = count_plan(d);
  ChsArena ar(workspace, workspace_bytes);
  uint64_t* sums = ar.take<uint64_t>(p.sum_blocks);
  const int32_t* ord = nullptr;
  if (cfg->sort_mode == CHS_SORT_DEPTH_PRESORT) {
    uint32_t* counts = ar.take<uint32_t>((uint64_t)cs::kDigits * p.t_cap);
    uint32_t* totals = ar.take<uint32_t>((uint64_t)d.Cb * cs::kDigits);
    uint32_t* ka = ar.take<uint32_t>(d.CbN);
    uint32_t* kb = ar.take<uint32_t>(d.CbN);
    int32_t* va = ar.take<int32_t>(d.CbN);
    int32_t* vb = ar.take<int32_t>(d.CbN);
    if (!ar.ok) {
      chs_set_error("chs_bin_count: workspace too small (%llu bytes)", (unsigned long long)workspace_bytes);
      return CHS_ERR_WORKSPACE_TOO_SMALL;
    }
    // cameras (frames with pose_fused) are segments of N pairs: four stable LSD passes over the depth bits inside every segment
    cs::TileMap map;
    memset(&map, 0, sizeof(map));
    map.n_seg = d.Cb; map.seg_len = (uint32_t)d.N; map.tiles_per_seg = p.tps;
    int st;
    if (cfg->pose_fused)
      st = radix_pass<cs::FusedDepthKeys, uint32_t>(map, cs::FusedDepthKeys{depths, tiles_touched, d.N, d.n}, nullptr, ka, va, 0, 8, counts, p.t_cap,
                                                    totals, nullptr, s);
    else
      st = radix_pass<cs::DepthKeys, uint32_t>(map, cs::DepthKeys{depths, tiles_touched}, nullptr, ka, va, 0, 8, counts, p.t_cap, totals, nullptr, s);
    if (st) return st;
    st = radix_pass<cs::ArrayKeys<uint32_t>, uint32_t>(map, cs::ArrayKeys<uint32_t>{ka}, va, kb, vb, 8, 8, counts, p.t_cap, totals, nullptr, s);
    if (st) return st;
    st = radix_pass<cs::ArrayKeys<uint32_t>, uint32_t>(map, cs::ArrayKeys<uint32_t>{kb}, vb, ka, va, 16, 8, counts, p.t_cap, totals, nullptr, s);
    if (st) return st;
    st = radix_pass<cs::ArrayKeys<uint32_t>, uint32_t>(map, cs::ArrayKeys<uint32_t>{ka}, va, (uint32_t*)nullptr, order, 24, 8, counts, p.t_cap,
                                                       totals, nullptr, s);
    if (st) return st;
    ord = order;
  }
  if (!ar.ok) {
    chs_set_error("chs_bin_count: workspace too small (%llu bytes)", (unsigned long long)workspace_bytes);
    return CHS_ERR_WORKSPACE_TOO_SMALL;
  }
  if (uses_band_route(cfg, d)) {  // the banded placement never reads isect_offsets: M is a plain (coalesced) sum
    const uint32_t blocks = (uint32_t)p.sum_blocks;
    cs::scan_sums_kernel<TouchedIn><<<blocks, cs::kThreads, 0, s>>>(TouchedIn{tiles_touched, nullptr}, (uint64_t)d.CbN, sums);
    CHS_LAUNCH_CHECK();
    cs::scan_of_sums_kernel<<<1, 1024, 0, s>>>(sums, blocks, n_isect_dev, nullptr);
    CHS_LAUNCH_CHECK();
    return CHS_OK;
  }
  return exclusive_scan(TouchedIn{tiles_touched, ord}, (uint64_t)d.CN, sums, StoreU32{isect_offsets}, n_isect_dev, s);
}

// LSD sort of (key, value) pairs over the low 8 * passes bits, one segment whose live length may only exist on the device.
// The pairs start in (k0, v0); buffers alternate so that the last pass lands in (k_final, v_final).
template <class KeyT>
int lsd_sort(int passes, uint64_t cap, const int64_t* n_dev, KeyT* k_a, int32_t* v_a, KeyT* k_final, int32_t* v_final, uint32_t* counts,
             uint32_t t_cap, uint32_t* totals, cudaStream_t s) {
  cs::TileMap map;
  memset(&map, 0, sizeof(map));
  map.n_seg = 1; map.seg_len = (uint32_t)cap; map.tiles_per_seg = t_cap; map.n_items_dev = n_dev;
  // the caller emitted into (k_a, v_a) if passes is odd, into (k_final, v_final) if even
  KeyT* kin = passes % 2 ? k_a : k_final;
  int32_t* vin = passes % 2 ? v_a : v_final;
  KeyT* kout = passes % 2 ? k_final : k_a;
  int32_t* vout = passes % 2 ? v_final : v_a;
  for (int p = 0; p < passes; ++p) {
    int st = radix_pass<cs::ArrayKeys<KeyT>, KeyT>(map, cs::ArrayKeys<KeyT>{kin}, vin, kout, vout, 8 * p, 8, counts, t_cap, totals, nullptr, s);
    if (st) return st;
    KeyT* tk = kin; kin = kout; kout = tk;
    int32_t* tv = vin; vin = vout; vout = tv;
  }
  return CHS_OK;
}

int band_bin_sort(const chs_config* cfg, const ChsDims& d, uint64_t cap, const float4* geom, const int32_t* radii, const float* depths,
                  const int32_t* order, uint64_t* keys_sorted, int32_t* vals_sorted, uint32_t* tile_offsets, void* workspace,
                  uint64_t workspace_bytes, cudaStream_t s) {
  const BandPlan p = band_plan(d, cap);
  const int64_t n_lin = (int64_t)d.Cb * d.tiles;
  ChsArena ar(workspace, workspace_bytes);
  ushort4* rects = ar.take<ushort4>(d.CbN);
  uint32_t* counts1 = ar.take<uint32_t>((uint64_t)256 * p.t_cap1);
  uint32_t* totals1 = ar.take<uint32_t>((uint64_t)d.Cb * 256);
  uint32_t* seg_begin2 = ar.take<uint32_t>((uint64_t)p.n_seg2 + 1);
  uint32_t* tile_first2 = ar.take<uint32_t>((uint64_t)p.n_seg2 + 1);
  uint2* recs = ar.take<uint2>(p.rec_cap);
  uint4* desc = ar.take<uint4>(p.t_cap2);
  uint32_t* counts2 = ar.take<uint32_t>((uint64_t)256 * p.t_cap2);
  uint32_t* totals2 = ar.take<uint32_t>((uint64_t)p.n_seg2 * 256);
  uint32_t* base2 = ar.take<uint32_t>((uint64_t)p.n_seg2 * 256);
  uint64_t* sums = ar.take<uint64_t>(p.sum_blocks + 1);
  if (!ar.ok) {
    chs_set_error("chs_bin_sort: workspace too small (%llu bytes)", (unsigned long long)workspace_bytes);
    return CHS_ERR_WORKSPACE_TOO_SMALL;
  }
  int64_t* total = reinterpret_cast<int64_t*>(sums + p.sum_blocks);
  BandArgs a;
  memset(&a, 0, sizeof(a));
  a.N = d.N; a.C = d.Cb; a.tile_w = d.tile_w; a.tile_h = d.tile_h; a.tight = cfg->tight_bounds != 0; a.BR = p.g.BR; a.n_bands = p.g.n_bands;
  a.fused = cfg->pose_fused != 0; a.n_virtual = d.n;
  a.chunks1 = p.chunks1; a.t_cap1 = p.t_cap1; a.t_cap2 = p.t_cap2; a.rec_cap = (uint32_t)p.rec_cap; a.val_cap = (uint32_t)cap;
  a.geom = geom; a.radii = radii; a.order = order; a.rects = rects; a.counts1 = counts1; a.seg_begin2 = seg_begin2; a.recs = recs;
  a.desc = desc; a.counts2 = counts2; a.base2 = base2; a.vals = vals_sorted;
  // P1: band split
  band_count_kernel<<<p.t_cap1, 256, 0, s>>>(a);
  CHS_LAUNCH_CHECK();
  cs::TileMap m1;
  memset(&m1, 0, sizeof(m1));
  m1.n_seg = d.Cb; m1.tiles_per_seg = (uint32_t)p.chunks1;
  int st = launch_colscan(m1, counts1, p.t_cap1, totals1, s);
  if (st) return st;
  segments_kernel<BucketLen><<<1, 1024, 0, s>>>(BucketLen{totals1, p.g.n_bands}, (int)p.n_seg2, kBandChunk, seg_begin2, tile_first2);
  CHS_LAUNCH_CHECK();
  const size_t smem1 = ((size_t)2 * p.g.n_bands * kBandRow + p.g.n_bands) * sizeof(uint32_t);
  if (smem1 > 48 * 1024) CHS_CUDA(cudaFuncSetAttribute(band_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
  band_scatter_kernel<<<p.t_cap1, 256, smem1, s>>>(a);
  CHS_LAUNCH_CHECK();
  // P2: tile placement inside every (camera, band) bucket
  chunk_desc_kernel<<<(p.t_cap2 + 255) / 256, 256, 0, s>>>((int)p.n_seg2, seg_begin2, tile_first2, p.t_cap2, (uint32_t)p.rec_cap, desc);
  CHS_LAUNCH_CHECK();
  tile_count_kernel<<<p.t_cap2, 256, 0, s>>>(a);
  CHS_LAUNCH_CHECK();
  cs::TileMap m2;
  memset(&m2, 0, sizeof(m2));
  m2.table = 1; m2.n_seg = (int)p.n_seg2; m2.seg_begin = seg_begin2; m2.tile_first = tile_first2;
  st = launch_colscan(m2, counts2, p.t_cap2, totals2, s);
  if (st) return st;
  st = exclusive_scan(ArrayIn{totals2}, (uint64_t)p.n_seg2 * 256, sums, StoreU32{base2}, total, s);
  if (st) return st;
  band_tile_offsets_kernel<<<grid_for(n_lin + 1), kThreads, 0, s>>>(n_lin, d.tiles, d.tile_w, p.g.BR, p.g.n_bands, base2,
                                                                     reinterpret_cast<const uint64_t*>(total), (uint32_t)cap, tile_offsets);
  CHS_LAUNCH_CHECK();
  const size_t smem2 = (size_t)256 * kBandRow * 4 + (size_t)kBandChunk * 4;
  if (cfg->tune_bin_chunk == 1) {  // development knob: warp-per-tile output
    CHS_CUDA(cudaFuncSetAttribute(tile_scatter_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem2 + 1024)));
    tile_scatter_warp_kernel<<<p.t_cap2, 256, smem2 + 1024, s>>>(a);
  } else {
    CHS_CUDA(cudaFuncSetAttribute(tile_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    tile_scatter_kernel<<<p.t_cap2, 256, smem2, s>>>(a);
  }
  CHS_LAUNCH_CHECK();
  if (keys_sorted) {
    rebuild_keys_from_offsets_kernel<<<(unsigned)n_lin, 128, 0, s>>>(d.tiles, d.tile_bits, tile_offsets, vals_sorted, depths, keys_sorted,
                                                                     cfg->pose_fused ? d.n : 0, d.N);
    CHS_LAUNCH_CHECK();
  }
  return CHS_OK;
}

int hw_bin_sort(const chs_config* cfg, const ChsDims& d, uint64_t cap, const int64_t* n_dev, const float* geom, const int32_t* radii,
                const float* depths, const uint32_t* isect_offsets, const int32_t* order, uint64_t* keys_sorted, int32_t* vals_sorted,
                uint32_t* tile_offsets, void* workspace, uint64_t workspace_bytes, cudaStream_t s) {
  if (uses_band_route(cfg, d))  // default: banded placement (no keys, no emitted intersections; isect_offsets is not read)
    return band_bin_sort(cfg, d, cap, (const float4*)geom, radii, depths, order, keys_sorted, vals_sorted, tile_offsets, workspace,
                         workspace_bytes, s);
  const SortPlan p = sort_plan(d, cfg->sort_mode, cap);
  const int64_t n_lin = (int64_t)d.C * d.tiles;
  const int tight = cfg->tight_bounds != 0;
  ChsArena ar(workspace, workspace_bytes);
  uint32_t* counts = ar.take<uint32_t>((uint64_t)cs::kDigits * p.t_cap);
#define CHS_SORT_WS_CHECK()                                                                                        \
  if (!ar.ok) {                                                                                                    \
    chs_set_error("chs_bin_sort: workspace too small (%llu bytes)", (unsigned long long)workspace_bytes);          \
    return CHS_ERR_WORKSPACE_TOO_SMALL;                                                                            \
  }
  if (p.multisplit) {
    uint16_t* k16 = ar.take<uint16_t>(cap);
    int32_t* v_in = ar.take<int32_t>(cap);
    int32_t* v_mid = ar.take<int32_t>(cap);
    uint8_t* d8 = ar.take<uint8_t>(cap);
    uint32_t* totals1 = ar.take<uint32_t>((uint64_t)d.C * cs::kDigits);
    uint32_t* seg_begin1 = ar.take<uint32_t>((uint64_t)d.C + 1);
    uint32_t* tile_first1 = ar.take<uint32_t>((uint64_t)d.C + 1);
    uint32_t* totals2 = ar.take<uint32_t>((uint64_t)p.n_seg2 * cs::kDigits);
    uint32_t* base2 = ar.take<uint32_t>((uint64_t)p.n_seg2 * cs::kDigits);
    uint32_t* seg_begin2 = ar.take<uint32_t>((uint64_t)p.n_seg2 + 1);
    uint32_t* tile_first2 = ar.take<uint32_t>((uint64_t)p.n_seg2 + 1);
    uint64_t* sums = ar.take<uint64_t>(p.sum_blocks);
    CHS_SORT_WS_CHECK();
    emit_kernel<2, uint16_t><<<grid_for(d.CN), kThreads, 0, s>>>(d.CN, d.N, d.tile_w, d.tile_h, d.tiles, d.tile_bits, tight, (uint32_t)cap,
                                                                 (const float4*)geom, radii, depths, isect_offsets, order, nullptr, k16, v_in);
    CHS_LAUNCH_CHECK();
    // pass 1: inside each camera, by the band tile >> 8
    segments_kernel<CamLen><<<1, 1024, 0, s>>>(CamLen{isect_offsets, d.N, d.C, n_dev, (uint32_t)cap}, d.C, cs::kTile, seg_begin1, tile_first1);
    CHS_LAUNCH_CHECK();
    cs::TileMap m1;
    memset(&m1, 0, sizeof(m1));
    m1.table = 1; m1.n_seg = d.C; m1.seg_begin = seg_begin1; m1.tile_first = tile_first1;
    int st = radix_pass<cs::ArrayKeys<uint16_t>, uint8_t>(m1, cs::ArrayKeys<uint16_t>{k16}, v_in, d8, v_mid, 8, p.nb > 1 ? chs_bit_length((uint64_t)p.nb - 1) : 1, counts, p.t_cap,
                                                          totals1, nullptr, s);
    if (st) return st;
    // pass 2: inside each (camera, band) bucket, by tile & 255, straight to every list's final position
    segments_kernel<BucketLen><<<1, 1024, 0, s>>>(BucketLen{totals1, p.nb}, (int)p.n_seg2, cs::kTile, seg_begin2, tile_first2);
    CHS_LAUNCH_CHECK();
    cs::TileMap m2;
    memset(&m2, 0, sizeof(m2));
    m2.table = 1; m2.n_seg = (int)p.n_seg2; m2.seg_begin = seg_begin2; m2.tile_first = tile_first2;
    st = launch_count(m2, cs::ArrayKeys<uint8_t>{d8}, 0, counts, p.t_cap, s);
    if (st) return st;
    st = launch_colscan(m2, counts, p.t_cap, totals2, s);
    if (st) return st;
    st = exclusive_scan(ArrayIn{totals2}, (uint64_t)p.n_seg2 * cs::kDigits, sums, StoreU32{base2}, nullptr, s);
    if (st) return st;
    tile_offsets_from_base_kernel<<<grid_for(n_lin + 1), kThreads, 0, s>>>(n_lin, d.tiles, p.nb, base2, seg_begin2, (int)p.n_seg2, tile_offsets);
    CHS_LAUNCH_CHECK();
    cs::ScatterArgs<cs::ArrayKeys<uint8_t>, uint8_t> a;
    a.map = m2; a.keys = cs::ArrayKeys<uint8_t>{d8}; a.vals_in = v_mid; a.keys_out = nullptr; a.vals_out = vals_sorted; a.shift = 0;
    a.prefix = counts; a.t_cap = p.t_cap; a.totals = totals2; a.digit_base = base2;
    st = launch_scatter(a, d.tiles > 1 ? chs_bit_length((uint64_t)d.tiles - 1) : 1, s);
    if (st) return st;
    if (keys_sorted) {
      rebuild_keys_from_offsets_kernel<<<(unsigned)n_lin, 128, 0, s>>>(d.tiles, d.tile_bits, tile_offsets, vals_sorted, depths, keys_sorted);
      CHS_LAUNCH_CHECK();
    }
    return CHS_OK;
  }
  // LSD routes: the literal 64-bit key sort, or linear cam * tiles + tile keys when tile ids do not fit 16 bits
  uint32_t* totals = ar.take<uint32_t>(cs::kDigits);
  int32_t* v_a = ar.take<int32_t>(cap);
  if (cfg->sort_mode == CHS_SORT_KEY64) {
    uint64_t* k_a = ar.take<uint64_t>(cap);
    uint64_t* k_final = keys_sorted ? keys_sorted : ar.take<uint64_t>(cap);
    CHS_SORT_WS_CHECK();
    uint64_t* k0 = p.passes % 2 ? k_a : k_final;
    int32_t* v0 = p.passes % 2 ? v_a : vals_sorted;
    emit_kernel<0, uint32_t><<<grid_for(d.CN), kThreads, 0, s>>>(d.CN, d.N, d.tile_w, d.tile_h, d.tiles, d.tile_bits, tight, (uint32_t)cap,
                                                                 (const float4*)geom, radii, depths, isect_offsets, nullptr, k0, nullptr, v0);
    CHS_LAUNCH_CHECK();
    int st = lsd_sort<uint64_t>(p.passes, cap, n_dev, k_a, v_a, k_final, vals_sorted, counts, p.t_cap, totals, s);
    if (st) return st;
    tile_offsets_kernel<0, uint32_t><<<grid_for(((int64_t)cap + 3) / 4), kThreads, 0, s>>>((int64_t)cap, n_dev, (int)n_lin, d.tile_bits, d.tiles, k_final,
                                                                                          nullptr, tile_offsets);
    CHS_LAUNCH_CHECK();
  } else {
    uint32_t* k_a = ar.take<uint32_t>(cap);
    uint32_t* k_final = ar.take<uint32_t>(cap);
    CHS_SORT_WS_CHECK();
    uint32_t* k0 = p.passes % 2 ? k_a : k_final;
    int32_t* v0 = p.passes % 2 ? v_a : vals_sorted;
    emit_kernel<1, uint32_t><<<grid_for(d.CN), kThreads, 0, s>>>(d.CN, d.N, d.tile_w, d.tile_h, d.tiles, d.tile_bits, tight, (uint32_t)cap,
                                                                 (const float4*)geom, radii, depths, isect_offsets, order, nullptr, k0, v0);
    CHS_LAUNCH_CHECK();
    int st = lsd_sort<uint32_t>(p.passes, cap, n_dev, k_a, v_a, k_final, vals_sorted, counts, p.t_cap, totals, s);
    if (st) return st;
    tile_offsets_kernel<1, uint32_t><<<grid_for(((int64_t)cap + 3) / 4), kThreads, 0, s>>>((int64_t)cap, n_dev, (int)n_lin, d.tile_bits, d.tiles, nullptr,
                                                                                          k_final, tile_offsets);
    CHS_LAUNCH_CHECK();
    if (keys_sorted) {
      rebuild_keys_from_offsets_kernel<<<(unsigned)n_lin, 128, 0, s>>>(d.tiles, d.tile_bits, tile_offsets, vals_sorted, depths, keys_sorted);
      CHS_LAUNCH_CHECK();
    }
  }
#undef CHS_SORT_WS_CHECK
  return CHS_OK;
}

}  // namespace

// exported to chs_api.cu: every route fits the same workspace, so a knob never changes the sizes a caller queried
int chs_bin_count_bytes(const ChsDims& d, int sort_mode, uint64_t* bytes) {
  uint64_t b = 0;
  int st = cub_bin_count_bytes(d, sort_mode, &b);
  if (st) return st;
  const uint64_t h = hw_count_bytes(d, sort_mode);
  *bytes = b > h ? b : h;
  return CHS_OK;
}

int chs_bin_sort_bytes(const ChsDims& d, const chs_config* cfg, int64_t M, uint64_t* bytes) {
  uint64_t b = 0;
  int st = legacy_bin_sort_bytes(d, cfg, M, &b);
  if (st) return st;
  uint64_t h = hw_sort_bytes(d, cfg->sort_mode, (uint64_t)(M > 0 ? M : 0));
  if (cfg->sort_mode == CHS_SORT_DEPTH_PRESORT && band_geom(d).ok) {
    const uint64_t bb = band_bytes(d, (uint64_t)(M > 0 ? M : 0));
    if (bb > h) h = bb;
  }
  *bytes = b > h ? b : h;
  return CHS_OK;
}

extern "C" int chs_bin_count(const chs_config* cfg, const int32_t* tiles_touched, const float* depths, uint32_t* isect_offsets,
                             int32_t* order, int64_t* n_isect_dev, int64_t* n_isect_host, void* workspace,
                             uint64_t workspace_bytes, void* stream) {
  if (cfg && cfg->tune_bin == 1)
    return cub_bin_count(cfg, tiles_touched, depths, isect_offsets, order, n_isect_dev, n_isect_host, workspace, workspace_bytes, stream);
  ChsDims d;
  int st = chs_make_dims(cfg, &d);
  if (st) return st;
  CHS_REQUIRE(tiles_touched && depths && isect_offsets && n_isect_dev && workspace, "chs_bin_count: null pointer");
  CHS_REQUIRE(cfg->sort_mode == CHS_SORT_KEY64 || order, "chs_bin_count: order buffer required for CHS_SORT_DEPTH_PRESORT");
  cudaStream_t s = (cudaStream_t)stream;
  if (d.CbN == 0) {
    CHS_CUDA(cudaMemsetAsync(n_isect_dev, 0, sizeof(int64_t), s));
    if (n_isect_host) *n_isect_host = 0;
    return CHS_OK;
  }
  st = hw_bin_count(cfg, d, tiles_touched, depths, isect_offsets, order, n_isect_dev, workspace, workspace_bytes, s);
  if (st) return st;
  if (n_isect_host) {
    CHS_CUDA(cudaMemcpyAsync(n_isect_host, n_isect_dev, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    CHS_CUDA(cudaStreamSynchronize(s));
    if (*n_isect_host >= ((int64_t)1 << 32) - 1) {
      chs_set_error("chs_bin_count: %lld intersections exceed the 2^32 limit of one launch; split the frame batch",
                    (long long)*n_isect_host);
      return CHS_ERR_UNSUPPORTED;
    }
  }
  return CHS_OK;
}

static int bin_sort_common(const chs_config* cfg, int64_t cap, const int64_t* n_dev, const float* geom, const int32_t* radii,
                           const float* depths, const uint32_t* isect_offsets, const int32_t* order, uint64_t* keys_sorted,
                           int32_t* vals_sorted, uint32_t* tile_offsets, void* workspace, uint64_t workspace_bytes, void* stream) {
  ChsDims d;
  int st = chs_make_dims(cfg, &d);
  if (st) return st;
  CHS_REQUIRE(geom && radii && depths && isect_offsets && tile_offsets, "chs_bin_sort: null pointer");
  CHS_REQUIRE(cap >= 0 && cap < ((int64_t)1 << 32) - 1, "chs_bin_sort: n_isect out of range");
  CHS_REQUIRE(cfg->sort_mode == CHS_SORT_KEY64 || order, "chs_bin_sort: order required for CHS_SORT_DEPTH_PRESORT");
  cudaStream_t s = (cudaStream_t)stream;
  if (cap == 0 || d.CN == 0) {
    CHS_CUDA(cudaMemsetAsync(tile_offsets, 0, ((size_t)d.Cb * d.tiles + 1) * sizeof(uint32_t), s));
    return CHS_OK;
  }
  CHS_REQUIRE(vals_sorted && workspace, "chs_bin_sort: null output/workspace");
  return hw_bin_sort(cfg, d, (uint64_t)cap, n_dev, geom, radii, depths, isect_offsets, order, keys_sorted, vals_sorted, tile_offsets, workspace,
                     workspace_bytes, s);
}

extern "C" int chs_bin_sort(const chs_config* cfg, int64_t n_isect, const float* geom, const int32_t* radii, const float* depths,
                            const uint32_t* isect_offsets, const int32_t* order, uint64_t* keys_sorted, int32_t* vals_sorted,
                            uint32_t* tile_offsets, void* workspace, uint64_t workspace_bytes, void* stream) {
  if (cfg && (cfg->tune_bin == 1 || cfg->tune_bin == 2))
    return legacy_bin_sort(cfg, n_isect, geom, radii, depths, isect_offsets, order, keys_sorted, vals_sorted, tile_offsets, workspace,
                           workspace_bytes, stream);
  return bin_sort_common(cfg, n_isect, nullptr, geom, radii, depths, isect_offsets, order, keys_sorted, vals_sorted, tile_offsets, workspace,
                         workspace_bytes, stream);
}

extern "C" int chs_bin_sort_dev(const chs_config* cfg, int64_t isect_capacity, const int64_t* n_isect_dev, const float* geom,
                                const int32_t* radii, const float* depths, const uint32_t* isect_offsets, const int32_t* order,
                                uint64_t* keys_sorted, int32_t* vals_sorted, uint32_t* tile_offsets, void* workspace,
                                uint64_t workspace_bytes, void* stream) {
  CHS_REQUIRE(n_isect_dev, "chs_bin_sort_dev: null n_isect_dev");
  CHS_REQUIRE(!cfg || cfg->tune_bin == 0 || cfg->tune_bin == 3, "chs_bin_sort_dev: only the hand-written binning routes run without the host knowing M");
  return bin_sort_common(cfg, isect_capacity, n_isect_dev, geom, radii, depths, isect_offsets, order, keys_sorted, vals_sorted, tile_offsets,
                         workspace, workspace_bytes, stream);
}

// ---- verification helper: the radix machinery on caller-supplied keys (tests/test_gpu_sort.py) ----
extern "C" int chs_radix_sort_pairs(const uint32_t* keys_in, int32_t n_seg, uint32_t seg_len, int32_t bits, uint32_t* keys_out,
                                    int32_t* vals_out, void* workspace, uint64_t workspace_bytes, void* stream) {
  CHS_REQUIRE(n_seg >= 0 && bits >= 1 && bits <= 32, "chs_radix_sort_pairs: bad arguments");
  const uint64_t n = (uint64_t)n_seg * seg_len;
  CHS_REQUIRE(n < ((uint64_t)1 << 32) - 1, "chs_radix_sort_pairs: too many items");
  if (n == 0) return CHS_OK;
  CHS_REQUIRE(keys_in && keys_out && vals_out && workspace, "chs_radix_sort_pairs: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const uint32_t tps = tiles_of(seg_len), t_cap = tps * (uint32_t)n_seg;
  ChsArena ar(workspace, workspace_bytes);
  uint32_t* counts = ar.take<uint32_t>((uint64_t)cs::kDigits * t_cap);
  uint32_t* totals = ar.take<uint32_t>((uint64_t)n_seg * cs::kDigits);
  uint32_t* kt = ar.take<uint32_t>(n);
  int32_t* vt = ar.take<int32_t>(n);
  if (!ar.ok) {
    chs_set_error("chs_radix_sort_pairs: workspace too small (need ~%llu bytes)",
                  (unsigned long long)(1024 + (uint64_t)cs::kDigits * t_cap * 4 + (uint64_t)n_seg * 1024 + n * 8 + 2048));
    return CHS_ERR_WORKSPACE_TOO_SMALL;
  }
  cs::TileMap map;
  memset(&map, 0, sizeof(map));
  map.n_seg = n_seg; map.seg_len = seg_len; map.tiles_per_seg = tps;
  const int passes = (bits + 7) / 8;
  // buffers alternate so that the last pass lands in the caller's outputs
  const uint32_t* kin = keys_in;
  const int32_t* vin = nullptr;
  for (int p = 0; p < passes; ++p) {
    const bool to_out = (passes - 1 - p) % 2 == 0;
    uint32_t* ko = to_out ? keys_out : kt;
    int32_t* vo = to_out ? vals_out : vt;
    int st = radix_pass<cs::ArrayKeys<uint32_t>, uint32_t>(map, cs::ArrayKeys<uint32_t>{kin}, vin, ko, vo, 8 * p, 8, counts, t_cap, totals, nullptr, s);
    if (st) return st;
    kin = ko;
    vin = vo;
  }
  return CHS_OK;
}
