// chs_sort.cuh — hand-written stable radix passes and scans for the binning stage (K2-K5).
//
// Everything the binning needs from a sort library, specialised for what the keys here are:
//   * radix pass over 8-bit digits of (key, int32 value) pairs, stable, over SEGMENTS that never mix (cameras in the depth
//     presort; cameras and then (camera, tile-band) buckets in the tile multisplit), as three deterministic kernels
//         count   : per-tile digit histogram            -> counts[digit][tile]
//         colscan : exclusive scan of every (segment, digit) column over the segment's tiles, + column totals
//         scatter : rank inside the tile, local reorder in shared memory, coalesced run-wise store
//     No decoupled look-back, no spinning: a bug can produce wrong output but never a hang.  The price is one extra read of
//     the keys per pass (2-4 B per item).
//   * ranking inside a tile is warp-synchronous: the 32 items a warp looks at in one round are consecutive in memory, lanes
//     with the same digit find each other with one ballot per digit bit, their order is the lane order, and the warp's private
//     digit counter advances by the group size.  Rounds are in memory order, warps are combined in warp order: stable by
//     construction, and the cost does not depend on the digit distribution (a tile whose items all share one digit costs the
//     same as a uniform one; MATCH.ANY, tried first, resolves one distinct value per iteration and was 3-4x slower on
//     uniformly distributed digits).
//   * a reduce-then-scan exclusive scan (block sums -> scan of the sums -> final) with a functor input, used for the
//     intersection offsets (gathered through the depth order) and for turning per-(camera, tile) counts into tile_offsets.
#pragma once
#include "chs_common.cuh"

namespace chs_sort {

constexpr int kThreads = 256;                  // also the number of digits: thread d owns digit d in the per-digit steps
constexpr int kItems = 16;                     // items per thread
constexpr int kTile = kThreads * kItems;       // 4096 items per tile
constexpr int kWarps = kThreads / 32;
constexpr int kDigits = 256;

// Which items a tile covers.  Uniform: n_seg segments of seg_len items, each cut into tiles_per_seg tiles (n_seg == 1 may take
// its live length from a device counter, which is how the tile multisplit runs without the host knowing M).  Table: segment s
// covers items [seg_begin[s], seg_begin[s+1]) and tiles [tile_first[s], tile_first[s+1]).
struct TileMap {
  int table;
  int n_seg;
  uint32_t seg_len, tiles_per_seg;
  const int64_t* n_items_dev;
  const uint32_t* seg_begin;
  const uint32_t* tile_first;
};

struct TileRange {
  int seg;
  uint32_t begin, end, seg_begin;
};

// resolved by thread 0, broadcast through shared memory; returns false (for the whole block) if the tile is past the end
__device__ __forceinline__ bool resolve_tile(const TileMap& m, uint32_t t, TileRange* s_range) {
  if (threadIdx.x == 0) {
    TileRange r;
    r.seg = -1; r.begin = r.end = r.seg_begin = 0;
    if (!m.table) {
      const uint32_t seg = t / m.tiles_per_seg, lt = t - seg * m.tiles_per_seg;
      uint32_t len = m.seg_len;
      if (m.n_items_dev) {
        const int64_t live = *m.n_items_dev;
        len = live < (int64_t)len ? (uint32_t)(live < 0 ? 0 : live) : len;
      }
      const uint64_t b = (uint64_t)lt * kTile;
      if ((int)seg < m.n_seg && b < len) {
        r.seg = (int)seg;
        r.seg_begin = seg * m.seg_len;
        r.begin = r.seg_begin + (uint32_t)b;
        r.end = r.seg_begin + (uint32_t)min((uint64_t)len, b + kTile);
      }
    } else if (t < m.tile_first[m.n_seg]) {
      int lo = 0, hi = m.n_seg - 1;  // last segment whose first tile is <= t (empty segments have equal firsts and are skipped)
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (m.tile_first[mid] <= t) lo = mid; else hi = mid - 1;
      }
      r.seg = lo;
      r.seg_begin = m.seg_begin[lo];
      const uint64_t b = (uint64_t)r.seg_begin + (uint64_t)(t - m.tile_first[lo]) * kTile;
      r.begin = (uint32_t)b;
      r.end = (uint32_t)min((uint64_t)m.seg_begin[lo + 1], b + kTile);
    }
    *s_range = r;
  }
  __syncthreads();
  return s_range->seg >= 0;
}

// ---- key sources ----
template <class T> struct ArrayKeys {
  const T* p;
  typedef T key_type;
  __device__ __forceinline__ T operator()(uint32_t i) const { return p[i]; }
};
// depth of a (camera, Gaussian) pair as a sortable key: positive floats order like their bit patterns; culled pairs go last
struct DepthKeys {
  const float* depths;
  const int32_t* touched;
  typedef uint32_t key_type;
  __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return touched[i] > 0 ? __float_as_uint(depths[i]) : 0xFFFFFFFFu; }
};

// pose_fused (one tile list per frame): pair i = frame * N + g is keyed by its depth at the frame's middle pose; `touched` is
// the per-frame union count [B, N]
struct FusedDepthKeys {
  const float* depths;      // [C, N]
  const int32_t* touched;   // [B, N]
  int N, n_virtual;
  typedef uint32_t key_type;
  __device__ __forceinline__ uint32_t operator()(uint32_t i) const {
    if (touched[i] <= 0) return 0xFFFFFFFFu;
    const uint32_t f = i / (uint32_t)N, g = i - f * (uint32_t)N;
    return __float_as_uint(depths[((size_t)f * n_virtual + n_virtual / 2) * N + g]);
  }
};

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// Lanes holding the same digit, found with one ballot per digit bit (cost independent of how many distinct digits the warp
// holds: MATCH.ANY resolves one group per iteration and measured 3-4x slower on uniformly distributed digits, r2b).  Four
// instructions per bit: test, ballot, conditional complement, and.
template <int kBits>
__device__ __forceinline__ unsigned match_digit(uint32_t d, bool valid) {
  unsigned peers = __ballot_sync(CHS_FULL_MASK, valid);
#pragma unroll
  for (int b = 0; b < kBits; ++b) {
    asm("{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b32 m;\n\t"
        "setp.ne.b32 p, %1, 0;\n\t"
        "vote.sync.ballot.b32 m, p, 0xffffffff;\n\t"
        "@!p not.b32 m, m;\n\t"
        "and.b32 %0, %0, m;\n\t"
        "}"
        : "+r"(peers)
        : "r"(d & (1u << b)));
  }
  return peers;
}

// Rank the tile's items by digit.  Item (warp w, round k, lane l) is element begin + w * 32 * kItems + k * 32 + l.  On return
// rank[k] is the item's position among the items OF ITS WARP with the same digit, and warp_hist[w][d] the warp's digit counts.
// kBits: number of significant digit bits (digits < 2^kBits), which bounds the ballots per round.
template <int kBits, class Keys, class KeyT>
__device__ __forceinline__ void rank_tile(const Keys& keys, const TileRange& r, int shift, KeyT key[kItems], uint16_t rank[kItems],
                                          uint32_t (*warp_hist)[kDigits]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kWarps * kDigits; i += kThreads) (&warp_hist[0][0])[i] = 0u;
  const uint32_t base = r.begin + (uint32_t)warp * (32 * kItems) + lane;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const uint32_t idx = base + k * 32;
    key[k] = idx < r.end ? keys(idx) : (KeyT)0;
  }
  __syncthreads();
  const unsigned lt = lanemask_lt();
  uint32_t* wh = warp_hist[warp];
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const bool valid = base + k * 32 < r.end;
    const uint32_t d = (uint32_t)((key[k] >> shift) & (KeyT)0xFF);
    const unsigned peers = match_digit<kBits>(d, valid);
    const uint32_t before = __popc(peers & lt);
    uint32_t start = 0;
    if (before == 0 && valid) {  // the first lane of the group advances the warp's counter for the whole group
      start = wh[d];
      wh[d] = start + __popc(peers);
    }
    __syncwarp();
    rank[k] = (uint16_t)(__shfl_sync(CHS_FULL_MASK, start, __ffs(peers | (1u << lane)) - 1) + before);
  }
  __syncthreads();
}

// ---- count: counts[d * t_cap + t] = number of items of tile t with digit d (order does not matter: shared-memory atomics) ----
template <class Keys>
__global__ void __launch_bounds__(kThreads) radix_count_kernel(TileMap map, Keys keys, int shift, uint32_t* __restrict__ counts, uint32_t t_cap) {
  typedef typename Keys::key_type KeyT;
  __shared__ uint32_t hist[kDigits];
  __shared__ TileRange s_range;
  const uint32_t t = blockIdx.x;
  const int d = threadIdx.x;
  hist[d] = 0u;
  if (!resolve_tile(map, t, &s_range)) {  // (contains the barrier that publishes the zeroed histogram)
    counts[(size_t)d * t_cap + t] = 0u;
    return;
  }
  const TileRange r = s_range;
  KeyT key[kItems];
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const uint32_t idx = r.begin + k * kThreads + threadIdx.x;
    key[k] = idx < r.end ? keys(idx) : (KeyT)0;
  }
#pragma unroll
  for (int k = 0; k < kItems; ++k)
    if (r.begin + k * kThreads + threadIdx.x < r.end) atomicAdd(&hist[(uint32_t)((key[k] >> shift) & (KeyT)0xFF)], 1u);
  __syncthreads();
  counts[(size_t)d * t_cap + t] = hist[d];
}

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t u = __shfl_up_sync(CHS_FULL_MASK, v, o);
    if (lane >= o) v += u;
  }
  return v;
}

// block-wide exclusive scan of one value per thread (kThreads threads); also returns the block total
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp /* [kWarps + 1] */, uint32_t* total) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t incl = warp_inclusive_scan(v, lane);
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int w = 0; w < kWarps; ++w) {
      const uint32_t c = s_warp[w];
      s_warp[w] = run;
      run += c;
    }
    s_warp[kWarps] = run;
  }
  __syncthreads();
  const uint32_t r = s_warp[warp] + incl - v;
  if (total) *total = s_warp[kWarps];
  __syncthreads();  // s_warp may be reused by the caller
  return r;
}

// ---- colscan: one warp per (segment, digit) column: counts -> exclusive prefix over the segment's tiles; totals[seg][d] ----
__global__ void __launch_bounds__(kThreads) radix_colscan_kernel(TileMap map, uint32_t* __restrict__ counts, uint32_t t_cap,
                                                                 uint32_t* __restrict__ totals) {
  const int lane = threadIdx.x & 31;
  const int64_t col = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int seg = (int)(col / kDigits), d = (int)(col % kDigits);
  if (seg >= map.n_seg) return;
  uint32_t t0, t1;
  if (map.table) {
    t0 = map.tile_first[seg];
    t1 = map.tile_first[seg + 1];
  } else {
    t0 = (uint32_t)seg * map.tiles_per_seg;
    t1 = t0 + map.tiles_per_seg;
  }
  uint32_t* c = counts + (size_t)d * t_cap;
  uint32_t run = 0;
  for (uint32_t tb = t0; tb < t1; tb += 32) {
    const uint32_t t = tb + lane;
    const uint32_t v = t < t1 ? c[t] : 0u;
    const uint32_t incl = warp_inclusive_scan(v, lane);
    if (t < t1) c[t] = run + incl - v;
    run += __shfl_sync(CHS_FULL_MASK, incl, 31);
  }
  if (lane == 0) totals[(size_t)seg * kDigits + d] = run;
}

// ---- scatter ----
// Output position of an item with digit d in segment s:  base(s, d) + prefix[d][t] + (position among the tile's items with d)
//   base(s, d) = digit_base[s * 256 + d] when given (the tile multisplit knows every list's final start),
//                else out_seg_begin(s) + sum_{d' < d} totals[s][d']  (an LSD pass inside the segment).
// keys_out / vals_out may be null (the last pass of a sort drops what nobody reads); vals_in null = the item's own index.
template <class Keys, class KeyOutT>
struct ScatterArgs {
  TileMap map;
  Keys keys;
  const int32_t* vals_in;
  KeyOutT* keys_out;
  int32_t* vals_out;
  int shift;
  const uint32_t* prefix;      // counts after colscan
  uint32_t t_cap;
  const uint32_t* totals;      // [n_seg][256]
  const uint32_t* digit_base;  // [n_seg][256] or null
};

template <int kBits, class Keys, class KeyOutT>
__global__ void __launch_bounds__(kThreads, 3) radix_scatter_kernel(ScatterArgs<Keys, KeyOutT> a) {
  typedef typename Keys::key_type KeyT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t(*warp_hist)[kDigits] = reinterpret_cast<uint32_t(*)[kDigits]>(smem_raw);
  int32_t* s_val = reinterpret_cast<int32_t*>(smem_raw + sizeof(uint32_t) * kWarps * kDigits);
  KeyT* s_key = reinterpret_cast<KeyT*>(smem_raw + sizeof(uint32_t) * kWarps * kDigits + sizeof(int32_t) * kTile);
  __shared__ uint32_t s_loc[kDigits];  // first local slot of digit d
  __shared__ uint32_t s_dst[kDigits];  // global position of local slot 0 of digit d's run, minus s_loc[d]
  __shared__ uint32_t s_scan[kWarps + 1];
  __shared__ TileRange s_range;
  const uint32_t t = blockIdx.x;
  if (!resolve_tile(a.map, t, &s_range)) return;
  const TileRange r = s_range;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, d = threadIdx.x;

  KeyT key[kItems];
  uint16_t rank[kItems];
  rank_tile<kBits, Keys, KeyT>(a.keys, r, a.shift, key, rank, warp_hist);

  // per digit: exclusive prefix over the warps, tile count
  uint32_t cnt = 0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) {
    const uint32_t c = warp_hist[w][d];
    warp_hist[w][d] = cnt;
    cnt += c;
  }
  const uint32_t loc = block_exclusive_scan(cnt, s_scan, nullptr);
  uint32_t gbase;
  if (a.digit_base) {
    gbase = a.digit_base[(size_t)r.seg * kDigits + d];
  } else {
    gbase = r.seg_begin + block_exclusive_scan(a.totals[(size_t)r.seg * kDigits + d], s_scan, nullptr);
  }
  gbase += a.prefix[(size_t)d * a.t_cap + t];
  s_loc[d] = loc;
  s_dst[d] = gbase - loc;
  __syncthreads();

  // local reorder
  const uint32_t base = r.begin + (uint32_t)warp * (32 * kItems) + lane;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const uint32_t idx = base + k * 32;
    if (idx < r.end) {
      const uint32_t dd = (uint32_t)((key[k] >> a.shift) & (KeyT)0xFF);
      const uint32_t pos = s_loc[dd] + warp_hist[warp][dd] + rank[k];
      s_key[pos] = key[k];
      s_val[pos] = a.vals_in ? a.vals_in[idx] : (int32_t)idx;
    }
  }
  __syncthreads();
  // run-wise store: consecutive threads write consecutive addresses inside every digit run
  const uint32_t n = r.end - r.begin;
  for (uint32_t i = threadIdx.x; i < n; i += kThreads) {
    const KeyT kk = s_key[i];
    const uint32_t dd = (uint32_t)((kk >> a.shift) & (KeyT)0xFF);
    const uint32_t dst = s_dst[dd] + i;
    if (a.keys_out) a.keys_out[dst] = (KeyOutT)kk;
    if (a.vals_out) a.vals_out[dst] = s_val[i];
  }
}

template <class KeyT> constexpr size_t scatter_smem_bytes() {
  return sizeof(uint32_t) * kWarps * kDigits + (sizeof(int32_t) + sizeof(KeyT)) * kTile;
}

// ---- exclusive scan, reduce-then-scan, functor input (uint32 values, uint64 running sums) ----
constexpr int kScanTile = kThreads * kItems;

template <class In>
__global__ void __launch_bounds__(kThreads) scan_sums_kernel(In in, uint64_t n, uint64_t* __restrict__ sums) {
  __shared__ uint64_t s_part[kWarps];
  const uint64_t b0 = (uint64_t)blockIdx.x * kScanTile;
  uint64_t acc = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const uint64_t i = b0 + (uint64_t)k * kThreads + threadIdx.x;
    if (i < n) acc += in(i);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(CHS_FULL_MASK, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t s = 0;
    for (int w = 0; w < kWarps; ++w) s += s_part[w];
    sums[blockIdx.x] = s;
  }
}

// single block: sums -> exclusive prefixes in place; total -> *total_out (int64) and, if given, *total_u32 (saturating)
__global__ void __launch_bounds__(1024) scan_of_sums_kernel(uint64_t* __restrict__ sums, uint32_t n_blocks, int64_t* total_out,
                                                            uint32_t* total_u32) {
  __shared__ uint64_t s_warp[33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t carry = 0;
  for (uint32_t b0 = 0; b0 < n_blocks; b0 += 1024) {
    const uint32_t i = b0 + threadIdx.x;
    const uint64_t v = i < n_blocks ? sums[i] : 0;
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
    if (i < n_blocks) sums[i] = carry + s_warp[warp] + incl - v;
    carry += s_warp[32];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (total_out) *total_out = (int64_t)carry;
    if (total_u32) *total_u32 = carry > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)carry;
  }
}

// out(i, exclusive prefix) for every i.  Loads and stores are striped (coalesced, and the gathers of a functor input are all
// in flight at once); the scan itself runs blocked (thread owns kItems consecutive items) out of shared memory, whose rows are
// padded by one word per 16 so that both access patterns are conflict-free.
__device__ __forceinline__ int scan_slot(int j) { return j + (j >> 4); }

template <class In, class Out>
__global__ void __launch_bounds__(kThreads) scan_final_kernel(In in, uint64_t n, const uint64_t* __restrict__ sums, Out out) {
  __shared__ uint32_t s_scan[kWarps + 1];
  __shared__ uint32_t s_v[kScanTile + kScanTile / 16];
  const uint64_t b0 = (uint64_t)blockIdx.x * kScanTile;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const int j = k * kThreads + threadIdx.x;
    s_v[scan_slot(j)] = b0 + j < n ? in(b0 + j) : 0u;
  }
  __syncthreads();
  uint32_t v[kItems];
  uint32_t mine = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    v[k] = s_v[scan_slot(threadIdx.x * kItems + k)];
    mine += v[k];
  }
  uint32_t run = block_exclusive_scan(mine, s_scan, nullptr) + (uint32_t)sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    s_v[scan_slot(threadIdx.x * kItems + k)] = run;
    run += v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const int j = k * kThreads + threadIdx.x;
    if (b0 + j < n) out(b0 + j, s_v[scan_slot(j)]);
  }
}

}  // namespace chs_sort
