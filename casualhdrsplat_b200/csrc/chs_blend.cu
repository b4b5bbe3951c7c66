// chs_blend.cu — K6 blend_fwd (+ fused formation epilogue) and K8 blend_bwd
// (SURVEY.md section 2.4, Appendix A.5 / A.6 / A.7).
//
// Layout of one CTA: 128 threads = one 16x16 pixel tile; warp w owns the 8x8 pixel block at
// ((w & 1) * 8, (w >> 1) * 8); lane l owns the TWO pixels (l & 7, l >> 3) and (l & 7, (l >> 3) + 4)
// of that block.  Every per-pixel quantity is a register pair (pixel A, pixel B) and every
// multiply/add on it is ONE packed fp32x2 instruction (Blackwell FFMA2 / FMUL2 / FADD2);
// per-Gaussian scalars enter as broadcast operands (the .F32 operand mode), so no splat moves are
// needed.  Compared with one pixel per thread (rounds r1a-c, see git history) this halves the
// per-pixel cost of the arithmetic and amortises the per-(warp, Gaussian) fixed cost — bit scan,
// shared-memory loads, the 14-shuffle butterfly and the RED — over 64 pixels instead of 32.
//
// Per batch of up to 256 tile-list entries the CTA gathers the per-(camera, Gaussian) records
// (three 128-bit loads each) into shared memory, with the conic in completed-square form pre-scaled
// for exp2 (chs_make_splat).  Each warp then *culls the batch against its own 8x8 block* with an
// exact ellipse-vs-rectangle bound (chs_block_max_power): 32 lanes test 32 Gaussians, a ballot
// yields the survivors, and only those are evaluated for the warp's 64 pixels.  Skipped pairs are
// exactly pairs with alpha < 1/255, so results are unchanged, but most of the 256 pair evaluations
// per intersection of a naive tile kernel vanish.  Early termination is warp-granular (ballot of the
// per-pixel "done" state) and CTA-granular (__syncthreads_and) per batch.
//
// Forward fuses the whole formation epilogue: the CTA loops over the n virtual poses of its frame,
// accumulates sum_k H_k in registers, and writes B = F(dt/n * sum_k H_k) once (decision D0 order).
//
// Backward walks each pixel's list back to front from last_id.  Kernel generations in this file (the defaults are the
// last of each; the earlier ones stay selectable through chs_config.tune_* as measured baselines):
//   forward   blend_fwd_kernel (round 1) -> blend_fwd2_kernel (round 2: survivor list, T -= w; with kGroup, the default: whole
//             batch culled first, survivors in groups of four with a speculative transmittance chain and one stop vote)
//   backward  blend_bwd_kernel ("direct": nine partials per Gaussian reduced across the warp with a transposing butterfly)
//             -> blend_bwd2_kernel ("tabled": the per-pixel sequential part runs pixel-parallel and leaves two scalars per
//             (pixel, Gaussian) in a per-warp shared-memory table; every 8 Gaussians the lanes switch to one Gaussian each and
//             sum their table rows privately) -> blend_bwd3_kernel (division-free colour state, running table pointer)
//             -> blend_bwd5_kernel (default: phase A in two stages over groups of four, for instruction-level parallelism)
//             and blend_bwd4_kernel (opt-in: phase B as tensor-core products over fp16 hi + lo tables).
// Negative results kept out of the code (r1g, c3): prefetching the next survivor's staged record inside the
// forward pair loop (to hide the bit-scan -> address -> LDS chain) made K6 2.55 -> 2.95 ms at 64 registers and
// 3.00 ms at 72; software-pipelining phase A of the tabled backward one Gaussian ahead cost +0.5 ms (the staged groups of
// blend_bwd5_kernel are what finally overlapped those chains).
#include <cuda_fp16.h>

#include <type_traits>

#include "chs_common.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int kBatch = 256;
#define CHS_LOG2_ALPHA_MIN (-7.994353436858858f) /* log2(1/255) */

// ---- packed fp32x2 helpers ----
struct P2 {
  unsigned long long v;
};
__device__ __forceinline__ P2 p2(float a, float b) {
  P2 r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r.v) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ P2 p2s(float a) { return p2(a, a); }
__device__ __forceinline__ float p2lo(P2 a) {
  float x, y;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v));
  (void)y;
  return x;
}
__device__ __forceinline__ float p2hi(P2 a) {
  float x, y;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v));
  (void)x;
  return y;
}
__device__ __forceinline__ P2 operator*(P2 a, P2 b) {
  P2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ P2 operator+(P2 a, P2 b) {
  P2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ P2 operator-(P2 a, P2 b) {
  P2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ P2 fma2(P2 a, P2 b, P2 c) {
  P2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}
__device__ __forceinline__ P2 neg2(P2 a) { return p2(-p2lo(a), -p2hi(a)); }
__device__ __forceinline__ float p2sum(P2 a) { return p2lo(a) + p2hi(a); }

// ---- staged batch of tile-list entries ----
template <int kB>
struct SplatSmemT {
  float4 a[kB];  // mx, my, qa, r
  float4 b[kB];  // kc, lo, val (int bits; c * N + g), rbc
  float4 c[kB];  // r, g, b, 1/opacity
};
using SplatSmem = SplatSmemT<kBatch>;

template <class Smem>
__device__ __forceinline__ void stage_splat(Smem& sm, int slot, int32_t val, int cam_base, const float4* __restrict__ geom,
                                            const float* __restrict__ conic_c, const float4* __restrict__ rgbo) {
  const float4 gm = __ldg(geom + val);
  const float cc = __ldg(conic_c + val);
  const float4 col = __ldg(rgbo + (val - cam_base));
  ChsSplat<float> s;
  chs_make_splat(gm.x, gm.y, gm.z, gm.w, cc, col.w, col.x, col.y, col.z, s);
  sm.a[slot] = make_float4(s.mx, s.my, s.qa, s.r);
  sm.b[slot] = make_float4(s.kc, s.lo, __int_as_float(val), s.rbc);
  sm.c[slot] = make_float4(s.cr, s.cg, s.cb, s.inv_opac);
}

// can staged splat `slot` reach alpha >= 1/255 anywhere in the warp's rectangle of pixel centres?
template <class Smem>
__device__ __forceinline__ bool splat_hits_block(const Smem& sm, int slot, float bx0, float bx1, float by0, float by1) {
  const float4 a = sm.a[slot];
  const float4 b = sm.b[slot];
  ChsSplat<float> s;
  s.mx = a.x; s.my = a.y; s.qa = a.z; s.r = a.w;
  s.kc = b.x; s.lo = b.y; s.rbc = b.w;
  return chs_block_max_power(s, bx0, bx1, by0, by1) >= CHS_LOG2_ALPHA_MIN - 1e-3f;
}

// log2(alpha) (before the 0.999 clamp) of a staged splat at this thread's two pixels (same column,
// rows y and y + 4).  Identical operation order in forward and backward, so both take the same skip
// decisions.  power = qa u^2 + kc dy^2 + lo with u = dx + r dy.
__device__ __forceinline__ P2 pair_power2(const float4 sa, float kc, float lo, float px, P2 py2, float& dx, P2& dy2, P2& u2) {
  dx = sa.x - px;
  dy2 = p2s(sa.y) - py2;
  u2 = fma2(p2s(sa.w), dy2, p2s(dx));
  const P2 t2 = fma2(p2s(kc) * dy2, dy2, p2s(lo));
  return fma2(p2s(sa.z) * u2, u2, t2);
}

struct BlendFwdArgs {
  int N, n_virtual, W, H, tile_w, tiles;
  int crf_kind, crf_hidden, crf_before_average, rgbo_per_camera;
  int fused;  // pose_fused: tile lists are per frame (entries frame * N + g); every pose of the frame walks the same list
  float bg[3];
  const float4* geom;
  const float* conic_c;
  const float4* rgbo;
  const int32_t* vals;
  const uint32_t* tile_offsets;
  const float* exposure;
  const float* crf_params;
  float *ldr, *alpha, *hdr_mean, *final_T;
  int32_t* last_id;
};

template <int kMinBlocks, bool kPerPoseCrf>
__global__ void __launch_bounds__(kThreads, kMinBlocks) blend_fwd_kernel(BlendFwdArgs a) {
  __shared__ SplatSmem sm;
  extern __shared__ float s_crf[];  // the CRF parameters [3, stride] when the CRF is learned

  const int tile = blockIdx.x, frame = blockIdx.y;
  const int tx = tile % a.tile_w, ty = tile / a.tile_w;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bx = tx * CHS_TILE + (warp & 1) * 8, by = ty * CHS_TILE + (warp >> 1) * 8;
  const int ix = bx + (lane & 7), iyA = by + (lane >> 3), iyB = iyA + 4;
  const bool insideA = ix < a.W && iyA < a.H, insideB = ix < a.W && iyB < a.H;
  const float px = ix + 0.5f;
  const P2 py2 = p2(iyA + 0.5f, iyB + 0.5f);
  const float bx0 = bx + 0.5f, bx1 = bx + 7.5f, by0 = by + 0.5f, by1 = by + 7.5f;
  const int64_t P = (int64_t)a.W * a.H;
  const int64_t pixA = (int64_t)iyA * a.W + ix, pixB = (int64_t)iyB * a.W + ix;
  const float kInf = __int_as_float(0x7f800000);

  const int crf_stride = chs_crf_stride(a.crf_kind, a.crf_hidden);  // 0 for the identity CRF
  for (int i = tid; i < 3 * crf_stride; i += kThreads) s_crf[i] = a.crf_params[i];

  // crf_before_average (figure order): sum_* accumulate F(dt * H_k) instead of H_k, and the per-pose
  // HDR images are written to hdr_mean, which is then [C,H,W,3] (the backward needs every H_k)
  constexpr bool per_pose_crf = kPerPoseCrf;
  const float dt = a.exposure[frame];
  if (per_pose_crf) __syncthreads();  // s_crf is read inside the pose loop
  P2 sum_r2 = p2s(0.f), sum_g2 = p2s(0.f), sum_b2 = p2s(0.f), sum_al2 = p2s(0.f);
  for (int k = 0; k < a.n_virtual; ++k) {
    const int c = frame * a.n_virtual + k;
    const int cam_base = a.rgbo_per_camera ? 0 : c * a.N;  // index of the camera's first record in rgbo
    const uint32_t start = a.tile_offsets[(int64_t)c * a.tiles + tile];
    const uint32_t end = a.tile_offsets[(int64_t)c * a.tiles + tile + 1];
    P2 T2 = p2s(1.f), acc_r2 = p2s(0.f), acc_g2 = p2s(0.f), acc_b2 = p2s(0.f);
    int lastA = 0, lastB = 0;
    // "done" is folded into the pixel's alpha threshold: a finished pixel has threshold +inf
    float thrA = insideA ? CHS_LOG2_ALPHA_MIN : kInf, thrB = insideB ? CHS_LOG2_ALPHA_MIN : kInf;
    bool warp_done = __all_sync(CHS_FULL_MASK, thrA == kInf && thrB == kInf);
    for (uint32_t base = start; base < end; base += kBatch) {
      // barrier + CTA-wide early exit; also protects the staged batch of the previous iteration
      if (__syncthreads_and(thrA == kInf && thrB == kInf)) break;
      const int cnt = min((uint32_t)kBatch, end - base);
      for (int i = tid; i < cnt; i += kThreads) stage_splat(sm, i, a.vals[base + i], cam_base, a.geom, a.conic_c, a.rgbo);
      __syncthreads();
      if (warp_done) continue;
      const int idx0 = (int)(base - start) + 1;
      for (int sub = 0; sub < cnt; sub += 32) {
        const int j = sub + lane;
        const bool hit = (j < cnt) && splat_hits_block(sm, j, bx0, bx1, by0, by1);
        unsigned mask = __ballot_sync(CHS_FULL_MASK, hit);
        while (mask) {
          const int jj = sub + __ffs(mask) - 1;
          mask &= mask - 1;
          const float4 sa = sm.a[jj];
          const float2 sb = *reinterpret_cast<const float2*>(&sm.b[jj]);  // kc, log2(opacity)
          float dx;
          P2 dy2, u2;
          const P2 pw2 = pair_power2(sa, sb.x, sb.y, px, py2, dx, dy2, u2);
          const float pA = p2lo(pw2), pB = p2hi(pw2);
          const bool actA = pA >= thrA, actB = pB >= thrB;
          if (actA || actB) {
            const float alA = actA ? fminf(CHS_ALPHA_MAX, chs_exp2_fast(pA)) : 0.f;
            const float alB = actB ? fminf(CHS_ALPHA_MAX, chs_exp2_fast(pB)) : 0.f;
            const P2 al2 = p2(alA, alB);
            const P2 Tn2 = T2 * (p2s(1.f) - al2);
            // stop (this Gaussian is not accumulated) when the transmittance would drop to <= 1e-4
            const bool accA = actA && p2lo(Tn2) > CHS_T_STOP, accB = actB && p2hi(Tn2) > CHS_T_STOP;
            thrA = (actA && !accA) ? kInf : thrA;
            thrB = (actB && !accB) ? kInf : thrB;
            const P2 w2 = p2(accA ? alA : 0.f, accB ? alB : 0.f) * T2;
            const float4 col = sm.c[jj];
            acc_r2 = fma2(p2s(col.x), w2, acc_r2);
            acc_g2 = fma2(p2s(col.y), w2, acc_g2);
            acc_b2 = fma2(p2s(col.z), w2, acc_b2);
            T2 = p2(accA ? p2lo(Tn2) : p2lo(T2), accB ? p2hi(Tn2) : p2hi(T2));
            lastA = accA ? idx0 + jj : lastA;
            lastB = accB ? idx0 + jj : lastB;
          }
        }
        if (__all_sync(CHS_FULL_MASK, thrA == kInf && thrB == kInf)) {
          warp_done = true;
          break;
        }
      }
    }
    __syncthreads();  // the next pose restages shared memory
    if (insideA) {
      a.final_T[(int64_t)c * P + pixA] = p2lo(T2);
      a.last_id[(int64_t)c * P + pixA] = lastA;
    }
    if (insideB) {
      a.final_T[(int64_t)c * P + pixB] = p2hi(T2);
      a.last_id[(int64_t)c * P + pixB] = lastB;
    }
    const P2 h_r2 = fma2(T2, p2s(a.bg[0]), acc_r2), h_g2 = fma2(T2, p2s(a.bg[1]), acc_g2), h_b2 = fma2(T2, p2s(a.bg[2]), acc_b2);
    sum_al2 = sum_al2 + (p2s(1.f) - T2);
    if (!per_pose_crf) {
      sum_r2 = sum_r2 + h_r2;
      sum_g2 = sum_g2 + h_g2;
      sum_b2 = sum_b2 + h_b2;
    } else {
      float y[2][3];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float hr = h == 0 ? p2lo(h_r2) : p2hi(h_r2), hg = h == 0 ? p2lo(h_g2) : p2hi(h_g2), hb = h == 0 ? p2lo(h_b2) : p2hi(h_b2);
        y[h][0] = dt * hr; y[h][1] = dt * hg; y[h][2] = dt * hb;
        if (a.crf_kind != CHS_CRF_IDENTITY) {
          y[h][0] = chs_crf_fwd(a.crf_kind, y[h][0], s_crf, a.crf_hidden);
          y[h][1] = chs_crf_fwd(a.crf_kind, y[h][1], s_crf + crf_stride, a.crf_hidden);
          y[h][2] = chs_crf_fwd(a.crf_kind, y[h][2], s_crf + 2 * crf_stride, a.crf_hidden);
        }
        if (h == 0 ? insideA : insideB) {
          const int64_t o = ((int64_t)c * P + (h == 0 ? pixA : pixB)) * 3;
          a.hdr_mean[o] = hr; a.hdr_mean[o + 1] = hg; a.hdr_mean[o + 2] = hb;
        }
      }
      sum_r2 = sum_r2 + p2(y[0][0], y[1][0]);
      sum_g2 = sum_g2 + p2(y[0][1], y[1][1]);
      sum_b2 = sum_b2 + p2(y[0][2], y[1][2]);
    }
  }
  // formation epilogue (A.7, decision D0): mean over poses, x exposure, CRF
  const float inv_n = 1.f / (float)a.n_virtual;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (!(h == 0 ? insideA : insideB)) continue;
    const float hr = (h == 0 ? p2lo(sum_r2) : p2hi(sum_r2)) * inv_n;
    const float hg = (h == 0 ? p2lo(sum_g2) : p2hi(sum_g2)) * inv_n;
    const float hb = (h == 0 ? p2lo(sum_b2) : p2hi(sum_b2)) * inv_n;
    const float al = (h == 0 ? p2lo(sum_al2) : p2hi(sum_al2)) * inv_n;
    const int64_t pix = h == 0 ? pixA : pixB;
    if (per_pose_crf) {  // the sums already hold F(dt * H_k)
      const int64_t o = ((int64_t)frame * P + pix) * 3;
      a.ldr[o] = hr; a.ldr[o + 1] = hg; a.ldr[o + 2] = hb;
      a.alpha[(int64_t)frame * P + pix] = al;
      continue;
    }
    float o0 = dt * hr, o1 = dt * hg, o2 = dt * hb;
    if (a.crf_kind != CHS_CRF_IDENTITY) {
      o0 = chs_crf_fwd(a.crf_kind, o0, s_crf, a.crf_hidden);
      o1 = chs_crf_fwd(a.crf_kind, o1, s_crf + crf_stride, a.crf_hidden);
      o2 = chs_crf_fwd(a.crf_kind, o2, s_crf + 2 * crf_stride, a.crf_hidden);
    }
    const int64_t o = ((int64_t)frame * P + pix) * 3;
    a.ldr[o] = o0; a.ldr[o + 1] = o1; a.ldr[o + 2] = o2;
    a.hdr_mean[o] = hr; a.hdr_mean[o + 1] = hg; a.hdr_mean[o + 2] = hb;
    a.alpha[(int64_t)frame * P + pix] = al;
  }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
struct BlendBwdArgs {
  int N, n_virtual, W, H, tile_w, tiles, v_hdr_per_camera, rgbo_per_camera, fused;
  float bg[3];
  const float4* geom;
  const float* conic_c;
  const float4* rgbo;
  const int32_t* vals;
  const uint32_t* tile_offsets;
  const float* final_T;
  const int32_t* last_id;
  const float* v_hdr;    // [B,H,W,3] gradient w.r.t. each pose's HDR image
  const float* v_alpha;  // [B,H,W] or null (gradient w.r.t. the pose-averaged alpha)
  float4* v_geom;        // [C,N] (v_mx, v_my, v_A, v_B)
  float4* v_cogr;        // [C,N] (v_C, v_opacity, v_r, v_g)
  float* v_blue;         // [C,N]  v_b
};

// Transposing butterfly: on entry every lane holds 8 partials v[0..7]; on exit every lane l holds
// the warp total of v[l & 7].  9 shuffles instead of 40.
__device__ __forceinline__ float warp_transpose_reduce8(float v[8], int lane) {
  {
    const bool odd = lane & 1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = odd ? v[2 * i] : v[2 * i + 1];
      const float keep = odd ? v[2 * i + 1] : v[2 * i];
      v[i] = keep + __shfl_xor_sync(CHS_FULL_MASK, send, 1);  // v[i] = value 2i + bit0
    }
  }
  {
    const bool odd = lane & 2;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = odd ? v[2 * i] : v[2 * i + 1];
      const float keep = odd ? v[2 * i + 1] : v[2 * i];
      v[i] = keep + __shfl_xor_sync(CHS_FULL_MASK, send, 2);  // v[i] = value 4i + 2 bit1 + bit0
    }
  }
  {
    const bool odd = lane & 4;
    const float send = odd ? v[0] : v[1];
    const float keep = odd ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(CHS_FULL_MASK, send, 4);  // value lane & 7
  }
  v[0] += __shfl_xor_sync(CHS_FULL_MASK, v[0], 8);
  v[0] += __shfl_xor_sync(CHS_FULL_MASK, v[0], 16);
  return v[0];
}

// NP = fp32x2 pixel pairs per thread: 1 -> warp owns 8x8 pixels (4 warps per tile), 2 -> warp owns 8x16 pixels
// (2 warps per tile); lane l owns column (l & 7) and rows (l >> 3) + 4 j, j = 0 .. 2 NP - 1.
template <int NP, int kMinBlocks>
__global__ void __launch_bounds__(kThreads / NP, kMinBlocks) blend_bwd_kernel(BlendBwdArgs a) {
  constexpr int kT = kThreads / NP;
  __shared__ SplatSmem sm;
  __shared__ int s_max_last;

  const int tile = blockIdx.x, c = blockIdx.y;
  const int frame = c / a.n_virtual;
  const int cam_base = a.rgbo_per_camera ? 0 : c * a.N;  // index of the camera's first record in rgbo
  const int tx = tile % a.tile_w, ty = tile / a.tile_w;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bx = tx * CHS_TILE + (warp & 1) * 8, by = ty * CHS_TILE + (NP == 1 ? (warp >> 1) * 8 : 0);
  const int ix = bx + (lane & 7), iy0 = by + (lane >> 3);
  const float px = ix + 0.5f;
  const float bx0 = bx + 0.5f, bx1 = bx + 7.5f, by0 = by + 0.5f, by1 = by + (NP == 1 ? 7.5f : 15.5f);
  const int64_t P = (int64_t)a.W * a.H;
  const uint32_t start = a.tile_offsets[(int64_t)c * a.tiles + tile];
  const uint32_t end = a.tile_offsets[(int64_t)c * a.tiles + tile + 1];
  if (end <= start) return;

  const float inv_nv = 1.f / (float)a.n_virtual;
  const int vimg = a.v_hdr_per_camera ? c : frame;  // figure-order CRF: every pose has its own HDR gradient
  P2 py2[NP], Tr2[NP], vh_r2[NP], vh_g2[NP], vh_b2[NP], vat2[NP], buf_r2[NP], buf_g2[NP], buf_b2[NP];
  int last[2 * NP];
  int warp_last = 0;
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    float Tf[2] = {1.f, 1.f}, v[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}}, vat[2] = {0.f, 0.f};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int iy = iy0 + 4 * (2 * p + h);
      last[2 * p + h] = 0;
      if (ix < a.W && iy < a.H) {
        const int64_t pix = (int64_t)iy * a.W + ix;
        Tf[h] = a.final_T[(int64_t)c * P + pix];
        last[2 * p + h] = a.last_id[(int64_t)c * P + pix];
        const int64_t o = ((int64_t)vimg * P + pix) * 3;
        v[h][0] = a.v_hdr[o]; v[h][1] = a.v_hdr[o + 1]; v[h][2] = a.v_hdr[o + 2];
        const float v_al = a.v_alpha ? a.v_alpha[(int64_t)frame * P + pix] * inv_nv : 0.f;
        vat[h] = Tf[h] * (v_al - (a.bg[0] * v[h][0] + a.bg[1] * v[h][1] + a.bg[2] * v[h][2]));
      }
      warp_last = max(warp_last, last[2 * p + h]);
    }
    py2[p] = p2(iy0 + 8 * p + 0.5f, iy0 + 8 * p + 4.5f);
    Tr2[p] = p2(Tf[0], Tf[1]);
    vh_r2[p] = p2(v[0][0], v[1][0]); vh_g2[p] = p2(v[0][1], v[1][1]); vh_b2[p] = p2(v[0][2], v[1][2]);
    vat2[p] = p2(vat[0], vat[1]);
    buf_r2[p] = buf_g2[p] = buf_b2[p] = p2s(0.f);
  }
  if (tid == 0) s_max_last = 0;
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) warp_last = max(warp_last, __shfl_xor_sync(CHS_FULL_MASK, warp_last, o));
  if (lane == 0 && warp_last > 0) atomicMax(&s_max_last, warp_last);
  __syncthreads();
  const int n_walk = s_max_last;  // entries [0, n_walk) of the tile list can matter

  // lanes 0..7 own the eight totals of the butterfly, lane 8 the ninth (blue): one RED instruction
  // with nine active lanes adds a Gaussian's whole gradient record
  char* lane_base = nullptr;
  uint32_t lane_stride = 0;
  if (lane < 4) {
    lane_base = reinterpret_cast<char*>(a.v_geom) + lane * 4;
    lane_stride = 16;
  } else if (lane < 8) {
    lane_base = reinterpret_cast<char*>(a.v_cogr) + (lane - 4) * 4;
    lane_stride = 16;
  } else if (lane == 8) {
    lane_base = reinterpret_cast<char*>(a.v_blue);
    lane_stride = 4;
  }

  const float k2 = -2.0f / CHS_LOG2E;
  for (int hi = n_walk; hi > 0; hi -= kBatch) {
    const int lo = max(0, hi - kBatch);
    const int cnt = hi - lo;
    __syncthreads();
    for (int i = tid; i < cnt; i += kT) stage_splat(sm, i, a.vals[start + lo + i], cam_base, a.geom, a.conic_c, a.rgbo);
    __syncthreads();
    if (warp_last <= lo) continue;  // none of this warp's pixels reaches into this batch
    for (int sub_hi = cnt; sub_hi > 0; sub_hi -= 32) {
      const int sub_lo = max(0, sub_hi - 32);
      if (warp_last <= lo + sub_lo) continue;
      const int j = sub_lo + lane;
      const bool hit = (j < sub_hi) && (lo + j < warp_last) && splat_hits_block(sm, j, bx0, bx1, by0, by1);
      unsigned mask = __ballot_sync(CHS_FULL_MASK, hit);
      while (mask) {
        const int bit = 31 - __clz(mask);
        mask &= ~(1u << bit);
        const int jj = sub_lo + bit;
        const int rel = lo + jj + 1;  // 1-based index in the tile list
        const float4 sa = sm.a[jj];   // mx, my, qa, r
        const float4 sb = sm.b[jj];   // kc, log2(opacity), val, rbc
        float dx;
        P2 dy2[NP], u2[NP];
        float pw[2 * NP];
        bool valid[2 * NP];
        bool any_valid = false;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const P2 pw2 = pair_power2(sa, sb.x, sb.y, px, py2[p], dx, dy2[p], u2[p]);
          pw[2 * p] = p2lo(pw2);
          pw[2 * p + 1] = p2hi(pw2);
          valid[2 * p] = (rel <= last[2 * p]) && pw[2 * p] >= CHS_LOG2_ALPHA_MIN;
          valid[2 * p + 1] = (rel <= last[2 * p + 1]) && pw[2 * p + 1] >= CHS_LOG2_ALPHA_MIN;
          any_valid |= valid[2 * p] || valid[2 * p + 1];
        }
        if (!__any_sync(CHS_FULL_MASK, any_valid)) continue;
        const float4 col = sm.c[jj];  // r, g, b, 1/opacity
        // packed, branch-free chs_pair_bwd (csrc/chs_math.cuh) for this thread's pixels; a pixel that does not
        // contribute runs with alpha = 0, which leaves its T / buf untouched and yields zero partials
        P2 G[9];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const float auA = valid[2 * p] ? chs_exp2_fast(pw[2 * p]) : 0.f;
          const float auB = valid[2 * p + 1] ? chs_exp2_fast(pw[2 * p + 1]) : 0.f;
          const P2 al2 = p2(fminf(CHS_ALPHA_MAX, auA), fminf(CHS_ALPHA_MAX, auB));
          const P2 om2 = p2s(1.f) - al2;
          const P2 ra2 = p2(chs_rcp_fast(p2lo(om2)), chs_rcp_fast(p2hi(om2)));
          Tr2[p] = Tr2[p] * ra2;  // transmittance before this Gaussian
          const P2 f2 = al2 * Tr2[p];
          const P2 g6 = f2 * vh_r2[p], g7 = f2 * vh_g2[p], g8 = f2 * vh_b2[p];
          const P2 nra2 = neg2(ra2);
          const P2 c0 = fma2(p2s(col.x), Tr2[p], buf_r2[p] * nra2);
          const P2 c1 = fma2(p2s(col.y), Tr2[p], buf_g2[p] * nra2);
          const P2 c2 = fma2(p2s(col.z), Tr2[p], buf_b2[p] * nra2);
          const P2 v_al2 = fma2(c0, vh_r2[p], fma2(c1, vh_g2[p], fma2(c2, vh_b2[p], vat2[p] * ra2)));
          buf_r2[p] = fma2(p2s(col.x), f2, buf_r2[p]);
          buf_g2[p] = fma2(p2s(col.y), f2, buf_g2[p]);
          buf_b2[p] = fma2(p2s(col.z), f2, buf_b2[p]);
          // v_sigma = -alpha_unclamped * v_alpha, and no gradient through the 0.999 clamp
          const P2 vs2 = p2(auA <= CHS_ALPHA_MAX ? -auA : 0.f, auB <= CHS_ALPHA_MAX ? -auB : 0.f) * v_al2;
          const P2 vk2 = vs2 * p2s(k2);
          const P2 g0 = (vk2 * p2s(sa.z)) * u2[p];                        // v_sigma A u
          const P2 g1 = fma2(p2s(sa.w), g0, (vk2 * p2s(sb.x)) * dy2[p]);  // v_sigma (B dx + C dy)
          const P2 hs2 = vs2 * p2s(0.5f);
          const P2 hx2 = hs2 * p2s(dx);
          const P2 g2 = hx2 * p2s(dx);
          const P2 g3 = (hx2 + hx2) * dy2[p];
          const P2 g4 = (hs2 * dy2[p]) * dy2[p];
          const P2 g5 = neg2(vs2) * p2s(col.w);
          if (p == 0) {
            G[0] = g0; G[1] = g1; G[2] = g2; G[3] = g3; G[4] = g4; G[5] = g5; G[6] = g6; G[7] = g7; G[8] = g8;
          } else {
            G[0] = G[0] + g0; G[1] = G[1] + g1; G[2] = G[2] + g2; G[3] = G[3] + g3; G[4] = G[4] + g4;
            G[5] = G[5] + g5; G[6] = G[6] + g6; G[7] = G[7] + g7; G[8] = G[8] + g8;
          }
        }
        float g[9] = {p2sum(G[0]), p2sum(G[1]), p2sum(G[2]), p2sum(G[3]), p2sum(G[4]), p2sum(G[5]), p2sum(G[6]), p2sum(G[7]), p2sum(G[8])};
        const uint32_t val = (uint32_t)__float_as_int(sb.z);
        const float blue = chs_warp_sum(g[8]);
        const float r8 = warp_transpose_reduce8(g, lane);
        const float add = lane == 8 ? blue : r8;
        if (lane < 9 && add != 0.f) atomicAdd(reinterpret_cast<float*>(lane_base + (uint64_t)val * lane_stride), add);
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------
// backward, tabled form (default).
//
// The per-(warp, Gaussian) cost of blend_bwd_kernel above is dominated not by the pair arithmetic
// but by what surrounds it: 9 partials x 2 pixels per lane have to be folded across the warp
// (~60 instructions of p2sum / select / shuffle / add) before one RED.  Here the roles are swapped
// for that part of the work:
//   phase A (lane = pixel pair, as before) walks the list back to front and computes, per
//           contributing Gaussian, only what is inherently sequential per pixel — alpha, the
//           transmittance before it, the colour behind it — and from those the two scalars every
//           gradient of that (pixel, Gaussian) pair is a multiple of:  vs = dL/dsigma  and
//           f = alpha T.  They go to a per-warp shared-memory table [slot][pixel].
//   phase B (lane = Gaussian) runs when kSlots Gaussians are tabled: each lane owns one of them (and
//           one part of the warp's 64 pixels), streams its table rows with 128-bit loads and
//           accumulates the nine sums  sum vs u, sum vs dy, sum vs dx^2, ...  sum f v_rgb  privately in
//           packed registers — no cross-lane reduction except one xor-shuffle per pixel part — then
//           scales them by the Gaussian's constants and issues two float4 REDs (+ one scalar).
// Issued instructions per (warp, Gaussian) drop from ~147 to ~87 at equal arithmetic (ncu: 5.39 G ->
// 3.43 G warp instructions per frame of c3).  A three-phase variant that first tables the alphas and
// then runs the sequential part as a separate branch-free loop was tried and lost (4.45 G
// instructions: the extra table round trip and loop bookkeeping cost more than the ILP gained).
// ---------------------------------------------------------------------------------------------
constexpr int kRowPad = 68;  // 64 pixels + 4: rows stay 16-byte aligned and 128-bit loads of different slots hit different banks

template <int kSlots>
struct __align__(16) BwdWarpSmem {
  float vs[kSlots][kRowPad];  // dL/dsigma of (slot, pixel)
  float f[kSlots][kRowPad];   // alpha * T of (slot, pixel)
  float vh[3][64];            // dL/dH of the warp's pixels (r, g, b planes)
  int ent[kSlots];            // staged-batch index of the slot's Gaussian
};

// phase B for the n_slots tabled Gaussians of this warp
template <int kSlots, class Smem>
__device__ __forceinline__ void bwd_round(const Smem& sm, const BwdWarpSmem<kSlots>& ws, int n_slots, int lane, float bxc, float byc,
                                          const BlendBwdArgs& a) {
  __syncwarp();
  constexpr int kW = kSlots <= 8 ? 8 : (kSlots <= 16 ? 16 : 32);  // lanes per pixel part (slots padded to a power of two)
  constexpr int kParts = 32 / kW;                                   // lanes per Gaussian, each covering kRows rows of 8 pixels
  constexpr int kRows = 8 / kParts;
  static_assert(kSlots >= 1 && kSlots <= 32, "at most one lane group per slot");
  const int k = lane % kW, part = lane / kW;
  const bool active = k < n_slots;
  P2 S_u = p2s(0.f), S_xx = p2s(0.f), S_dy = p2s(0.f), S_xy = p2s(0.f), S_yy = p2s(0.f), S_v = p2s(0.f);
  P2 G6 = p2s(0.f), G7 = p2s(0.f), G8 = p2s(0.f);
  float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sb = sa;
  float inv_opac = 0.f;
  if (active) {
    const int jj = ws.ent[k];
    sa = sm.a[jj];  // mx, my, qa, r
    sb = sm.b[jj];  // kc, log2(opacity), val, rbc
    inv_opac = sm.c[jj].w;
    const float dxb = sa.x - bxc;  // mean - centre of pixel column 0
    P2 dx2[4];
#pragma unroll
    for (int xp = 0; xp < 4; ++xp) dx2[xp] = p2(dxb - (float)(2 * xp), dxb - (float)(2 * xp + 1));
#pragma unroll
    for (int yy = 0; yy < kRows; ++yy) {
      const int y = part * kRows + yy;
      const float dy = sa.y - (byc + (float)y);
      const P2 rdy = p2s(sa.w * dy);
      P2 rowvs = p2s(0.f), rowt = p2s(0.f);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int q = y * 8 + 4 * c;
        const float4 vs4 = *reinterpret_cast<const float4*>(&ws.vs[k][q]);
        const float4 f4 = *reinterpret_cast<const float4*>(&ws.f[k][q]);
        const float4 r4 = *reinterpret_cast<const float4*>(&ws.vh[0][q]);
        const float4 g4 = *reinterpret_cast<const float4*>(&ws.vh[1][q]);
        const float4 b4 = *reinterpret_cast<const float4*>(&ws.vh[2][q]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const P2 vs2 = h ? p2(vs4.z, vs4.w) : p2(vs4.x, vs4.y);
          const P2 f2 = h ? p2(f4.z, f4.w) : p2(f4.x, f4.y);
          const P2 dxp = dx2[2 * c + h];
          const P2 u2 = dxp + rdy;
          S_u = fma2(vs2, u2, S_u);
          const P2 t2 = vs2 * dxp;
          S_xx = fma2(t2, dxp, S_xx);
          rowvs = rowvs + vs2;
          rowt = rowt + t2;
          G6 = fma2(f2, h ? p2(r4.z, r4.w) : p2(r4.x, r4.y), G6);
          G7 = fma2(f2, h ? p2(g4.z, g4.w) : p2(g4.x, g4.y), G7);
          G8 = fma2(f2, h ? p2(b4.z, b4.w) : p2(b4.x, b4.y), G8);
        }
      }
      const P2 dy2 = p2s(dy);
      S_dy = fma2(rowvs, dy2, S_dy);
      S_xy = fma2(rowt, dy2, S_xy);
      S_yy = fma2(rowvs * dy2, dy2, S_yy);
      S_v = S_v + rowvs;
    }
  }
  float t[9] = {p2sum(S_u), p2sum(S_dy), p2sum(S_xx), p2sum(S_xy), p2sum(S_yy), p2sum(S_v), p2sum(G6), p2sum(G7), p2sum(G8)};
#pragma unroll
  for (int o = kW; o < 32; o <<= 1)
#pragma unroll
    for (int i = 0; i < 9; ++i) t[i] += __shfl_xor_sync(CHS_FULL_MASK, t[i], o);
  if (active) {
    ChsSplat<float> sp;
    sp.qa = sa.z; sp.r = sa.w; sp.kc = sb.x; sp.inv_opac = inv_opac;
    float g[9];
    chs_moments_to_grads(sp, t, g);  // csrc/chs_math.cuh, checked on the host against chs_pair_bwd and the oracle
    const uint32_t val = (uint32_t)__float_as_int(sb.z);
    if (part == 0) {
      atomicAdd(a.v_geom + val, make_float4(g[0], g[1], g[2], g[3]));
      if (kParts < 3) atomicAdd(a.v_blue + val, g[8]);
    }
    if (part == (kParts > 1 ? 1 : 0)) atomicAdd(a.v_cogr + val, make_float4(g[4], g[5], g[6], g[7]));
    if (kParts >= 3 && part == 2) atomicAdd(a.v_blue + val, g[8]);
  }
  __syncwarp();  // the table is rewritten by the next round
}

template <int kSlots, int kB, int kMinBlocks, bool kPipe>
__global__ void __launch_bounds__(kThreads, kMinBlocks) blend_bwd2_kernel(BlendBwdArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using Smem = SplatSmemT<kB>;
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  __shared__ int s_max_last;

  const int tile = blockIdx.x, c = blockIdx.y;
  const int frame = c / a.n_virtual;
  const int cam_base = a.rgbo_per_camera ? 0 : c * a.N;
  const int tx = tile % a.tile_w, ty = tile / a.tile_w;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  BwdWarpSmem<kSlots>& ws = reinterpret_cast<BwdWarpSmem<kSlots>*>(smem_raw + sizeof(Smem))[warp];
  const int bx = tx * CHS_TILE + (warp & 1) * 8, by = ty * CHS_TILE + (warp >> 1) * 8;
  const int ix = bx + (lane & 7), iy0 = by + (lane >> 3);
  const float px = ix + 0.5f;
  const float bx0 = bx + 0.5f, bx1 = bx + 7.5f, by0 = by + 0.5f, by1 = by + 7.5f;
  const int64_t P = (int64_t)a.W * a.H;
  const uint32_t start = a.tile_offsets[(int64_t)c * a.tiles + tile];
  const uint32_t end = a.tile_offsets[(int64_t)c * a.tiles + tile + 1];
  if (end <= start) return;

  const float inv_nv = 1.f / (float)a.n_virtual;
  const int vimg = a.v_hdr_per_camera ? c : frame;
  int last[2];
  int warp_last = 0;
  float Tf[2] = {1.f, 1.f}, v[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}}, vat[2] = {0.f, 0.f};
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int iy = iy0 + 4 * h;
    last[h] = 0;
    if (ix < a.W && iy < a.H) {
      const int64_t pix = (int64_t)iy * a.W + ix;
      Tf[h] = a.final_T[(int64_t)c * P + pix];
      last[h] = a.last_id[(int64_t)c * P + pix];
      const int64_t o = ((int64_t)vimg * P + pix) * 3;
      v[h][0] = a.v_hdr[o]; v[h][1] = a.v_hdr[o + 1]; v[h][2] = a.v_hdr[o + 2];
      const float v_al = a.v_alpha ? a.v_alpha[(int64_t)frame * P + pix] * inv_nv : 0.f;
      vat[h] = Tf[h] * (v_al - (a.bg[0] * v[h][0] + a.bg[1] * v[h][1] + a.bg[2] * v[h][2]));
    }
    warp_last = max(warp_last, last[h]);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) ws.vh[ch][lane + 32 * h] = v[h][ch];  // pixel id = lane + 32 h  <->  (x, y) = (id & 7, id >> 3)
  }
  const P2 py2 = p2(iy0 + 0.5f, iy0 + 4.5f);
  P2 Tr2 = p2(Tf[0], Tf[1]);
  const P2 vh_r2 = p2(v[0][0], v[1][0]), vh_g2 = p2(v[0][1], v[1][1]), vh_b2 = p2(v[0][2], v[1][2]);
  const P2 vat2 = p2(vat[0], vat[1]);
  P2 buf_r2 = p2s(0.f), buf_g2 = p2s(0.f), buf_b2 = p2s(0.f);
  if (tid == 0) s_max_last = 0;
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) warp_last = max(warp_last, __shfl_xor_sync(CHS_FULL_MASK, warp_last, o));
  if (lane == 0 && warp_last > 0) atomicMax(&s_max_last, warp_last);
  __syncthreads();
  const int n_walk = s_max_last;

  int n_slots = 0;
  for (int hi = n_walk; hi > 0; hi -= kB) {
    const int lo = max(0, hi - kB);
    const int cnt = hi - lo;
    __syncthreads();
    for (int i = tid; i < cnt; i += kThreads) stage_splat(sm, i, a.vals[start + lo + i], cam_base, a.geom, a.conic_c, a.rgbo);
    __syncthreads();
    if (warp_last <= lo) continue;
    for (int sub_hi = cnt; sub_hi > 0; sub_hi -= 32) {
      const int sub_lo = max(0, sub_hi - 32);
      if (warp_last <= lo + sub_lo) continue;
      const int j = sub_lo + lane;
      const bool hit = (j < sub_hi) && (lo + j < warp_last) && splat_hits_block(sm, j, bx0, bx1, by0, by1);
      unsigned mask = __ballot_sync(CHS_FULL_MASK, hit);
      // ---- phase A over the survivors, back to front.  kPipe: the record and exponent of the NEXT survivor are fetched
      // before the current one's sequential chain (they do not depend on it), which roughly doubles the ILP of the loop ----
      int jj = 0;
      float4 sa = make_float4(0.f, 0.f, 0.f, 0.f);
      float2 sb = make_float2(0.f, 0.f);
      P2 pw2 = p2s(0.f);
      float dx;
      P2 dy2, u2;
      bool have = mask != 0;
      if (kPipe && have) {
        const int bit = 31 - __clz(mask);
        mask &= ~(1u << bit);
        jj = sub_lo + bit;
        sa = sm.a[jj];
        sb = *reinterpret_cast<const float2*>(&sm.b[jj]);
        pw2 = pair_power2(sa, sb.x, sb.y, px, py2, dx, dy2, u2);
      }
      while (kPipe ? have : mask != 0) {
        int jn = 0;
        P2 pwn = p2s(0.f);
        if (kPipe) {
          have = mask != 0;
          if (have) {
            const int bit = 31 - __clz(mask);
            mask &= ~(1u << bit);
            jn = sub_lo + bit;
            const float4 san = sm.a[jn];
            const float2 sbn = *reinterpret_cast<const float2*>(&sm.b[jn]);
            pwn = pair_power2(san, sbn.x, sbn.y, px, py2, dx, dy2, u2);
          }
        } else {
          const int bit = 31 - __clz(mask);
          mask &= ~(1u << bit);
          jj = sub_lo + bit;
          sa = sm.a[jj];
          sb = *reinterpret_cast<const float2*>(&sm.b[jj]);  // kc, log2(opacity)
          pw2 = pair_power2(sa, sb.x, sb.y, px, py2, dx, dy2, u2);
        }
        const int rel = lo + jj + 1;
        const float pA = p2lo(pw2), pB = p2hi(pw2);
        const bool validA = (rel <= last[0]) && pA >= CHS_LOG2_ALPHA_MIN;
        const bool validB = (rel <= last[1]) && pB >= CHS_LOG2_ALPHA_MIN;
        const int jcur = jj;
        if (kPipe) {
          jj = jn;
          pw2 = pwn;
        }
        if (!__any_sync(CHS_FULL_MASK, validA || validB)) continue;
        const float4 col = sm.c[jcur];
        const float auA = validA ? chs_exp2_fast(pA) : 0.f;
        const float auB = validB ? chs_exp2_fast(pB) : 0.f;
        const P2 al2 = p2(fminf(CHS_ALPHA_MAX, auA), fminf(CHS_ALPHA_MAX, auB));
        const P2 om2 = p2s(1.f) - al2;
        const P2 ra2 = p2(chs_rcp_fast(p2lo(om2)), chs_rcp_fast(p2hi(om2)));
        Tr2 = Tr2 * ra2;  // transmittance before this Gaussian
        const P2 f2 = al2 * Tr2;
        const P2 nra2 = neg2(ra2);
        const P2 c0 = fma2(p2s(col.x), Tr2, buf_r2 * nra2);
        const P2 c1 = fma2(p2s(col.y), Tr2, buf_g2 * nra2);
        const P2 c2 = fma2(p2s(col.z), Tr2, buf_b2 * nra2);
        const P2 v_al2 = fma2(c0, vh_r2, fma2(c1, vh_g2, fma2(c2, vh_b2, vat2 * ra2)));
        buf_r2 = fma2(p2s(col.x), f2, buf_r2);
        buf_g2 = fma2(p2s(col.y), f2, buf_g2);
        buf_b2 = fma2(p2s(col.z), f2, buf_b2);
        // v_sigma = -alpha_unclamped * v_alpha, and no gradient through the 0.999 clamp
        const P2 vs2 = p2(auA <= CHS_ALPHA_MAX ? -auA : 0.f, auB <= CHS_ALPHA_MAX ? -auB : 0.f) * v_al2;
        ws.vs[n_slots][lane] = p2lo(vs2);
        ws.vs[n_slots][lane + 32] = p2hi(vs2);
        ws.f[n_slots][lane] = p2lo(f2);
        ws.f[n_slots][lane + 32] = p2hi(f2);
        if (lane == 0) ws.ent[n_slots] = jcur;
        if (++n_slots == kSlots) {
          bwd_round<kSlots>(sm, ws, n_slots, lane, bx0, by0, a);
          n_slots = 0;
        }
      }
    }
    if (n_slots > 0) {  // the staged batch is about to be replaced
      bwd_round<kSlots>(sm, ws, n_slots, lane, bx0, by0, a);
      n_slots = 0;
    }
  }
}


// =================================================================================================
// Round-2 kernels (defaults).  Same tile / warp / pixel-pair layout and the same exact culling as above; what changes is the
// instruction count of the inner loops (ncu source pages of the r1g kernels, profiles/r1g_ncu_blend.csv):
//   * SURVIVOR LIST.  The lanes that test 32 staged Gaussians against the warp's block write the survivors' staged indices
//     compactly to a per-warp shared-memory list (one STS per survivor); the pair loop then reads "next index" with one
//     broadcast LDS instead of finding and clearing the next set bit of the ballot (FLO/BREV, shifts, masks, address
//     arithmetic: 6-8 instructions per (warp, Gaussian) in r1).
//   * Staged record repacked as A = (mx, my, qa, r), B = (kc, lo, cr, cg), C = (cb, 1/opacity, val, rbc): the hot loops read A,
//     B and one scalar of C.
//   * Forward: the transmittance is updated as T -= w with w = alpha T (or 0), which replaces the per-pixel select between
//     "T (1 - alpha)" and "T" (four register moves per pair in r1) and reuses the product the colour sums need anyway; the last
//     accumulated index is tracked batch-relative.
//   * Backward, phase A: the three "colour behind" accumulators become ONE scalar R, the colour behind per unit of
//     transmittance dotted with the pixel's upstream gradient (chs_pair_bwd_scalars_r in chs_math.cuh: R <- R - alpha (R - c.v),
//     no division, dL/dalpha = T (c.v - R)): 11 packed operations instead of 19.  The table row of a slot is reached through a
//     running pointer (r1 recomputed the address from the thread id every iteration), a lane's two pixels are adjacent in the
//     row (two STS.64 instead of four STS.32), and the staged index of the slot rides in the row's padding.
//   * Backward, phase B: a lane's 128-bit table loads now hold pixel pairs that differ only in the row (y, y + 4), so the row
//     sums stay packed to the end; sum vs.u is derived as sum vs.dx + r sum vs.dy instead of being accumulated per pixel.
// =================================================================================================
template <int kB>
struct SplatSmem3 {  // what the block culling reads (a, b.xyz) is two conflict-free 128-bit loads per lane
  float4 a[kB];  // mx, my, qa, r
  float4 b[kB];  // kc, lo, rbc, cr
  float4 c[kB];  // cg, cb, 1/opacity, val (int bits; c * N + g)
};

template <class Smem>
__device__ __forceinline__ void stage_splat3(Smem& sm, int slot, int32_t val, int cam_base, const float4* __restrict__ geom,
                                             const float* __restrict__ conic_c, const float4* __restrict__ rgbo) {
  const float4 gm = __ldg(geom + val);
  const float cc = __ldg(conic_c + val);
  const float4 col = __ldg(rgbo + (val - cam_base));
  ChsSplat<float> s;
  chs_make_splat(gm.x, gm.y, gm.z, gm.w, cc, col.w, col.x, col.y, col.z, s);
  sm.a[slot] = make_float4(s.mx, s.my, s.qa, s.r);
  sm.b[slot] = make_float4(s.kc, s.lo, s.rbc, s.cr);
  sm.c[slot] = make_float4(s.cg, s.cb, s.inv_opac, __int_as_float(val));
}

// two entries of one thread: both gathers are in flight before either record is transformed
template <class Smem>
__device__ __forceinline__ void stage_splat3_pair(Smem& sm, int slot0, bool on0, int32_t val0, int slot1, bool on1, int32_t val1, int cam_base,
                                                  const float4* __restrict__ geom, const float* __restrict__ conic_c,
                                                  const float4* __restrict__ rgbo) {
  float4 gm0 = make_float4(0.f, 0.f, 1.f, 0.f), gm1 = gm0, col0 = make_float4(0.f, 0.f, 0.f, 1.f), col1 = col0;
  float cc0 = 1.f, cc1 = 1.f;
  if (on0) {
    gm0 = __ldg(geom + val0);
    cc0 = __ldg(conic_c + val0);
    col0 = __ldg(rgbo + (val0 - cam_base));
  }
  if (on1) {
    gm1 = __ldg(geom + val1);
    cc1 = __ldg(conic_c + val1);
    col1 = __ldg(rgbo + (val1 - cam_base));
  }
  ChsSplat<float> s;
  if (on0) {
    chs_make_splat(gm0.x, gm0.y, gm0.z, gm0.w, cc0, col0.w, col0.x, col0.y, col0.z, s);
    sm.a[slot0] = make_float4(s.mx, s.my, s.qa, s.r);
    sm.b[slot0] = make_float4(s.kc, s.lo, s.rbc, s.cr);
    sm.c[slot0] = make_float4(s.cg, s.cb, s.inv_opac, __int_as_float(val0));
  }
  if (on1) {
    chs_make_splat(gm1.x, gm1.y, gm1.z, gm1.w, cc1, col1.w, col1.x, col1.y, col1.z, s);
    sm.a[slot1] = make_float4(s.mx, s.my, s.qa, s.r);
    sm.b[slot1] = make_float4(s.kc, s.lo, s.rbc, s.cr);
    sm.c[slot1] = make_float4(s.cg, s.cb, s.inv_opac, __int_as_float(val1));
  }
}

template <class Smem>
__device__ __forceinline__ bool splat_hits_block3(const Smem& sm, int slot, float bx0, float bx1, float by0, float by1) {
  const float4 a = sm.a[slot];
  const float4 b = sm.b[slot];
  ChsSplat<float> s;
  s.mx = a.x; s.my = a.y; s.qa = a.z; s.r = a.w;
  s.kc = b.x; s.lo = b.y; s.rbc = b.z;
  return chs_block_max_power(s, bx0, bx1, by0, by1) >= CHS_LOG2_ALPHA_MIN - 1e-3f;
}

// ---- asynchronous staging (cp.async = LDGSTS): the gathers of the NEXT batch's raw records travel while the current batch
// is processed; each thread later transforms the record it fetched itself, so only its own copies have to be complete ----
template <int kB>
struct RawSmem {  // tile-list entries as gathered (36 bytes each); the record index stays in a register of the fetching thread
  float4 gm[kB];   // geom: mean2d.x, mean2d.y, conic A, conic B
  float4 col[kB];  // rgbo: r, g, b, opacity
  float cc[kB];    // conic C
};
__device__ __forceinline__ void cp_async_16(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <class Raw>
__device__ __forceinline__ void gather_raw_async(Raw& r, int slot, int32_t val, int cam_base, const float4* __restrict__ geom,
                                                 const float* __restrict__ conic_c, const float4* __restrict__ rgbo) {
  cp_async_16(&r.gm[slot], geom + val);
  cp_async_16(&r.col[slot], rgbo + (val - cam_base));
  cp_async_4(&r.cc[slot], conic_c + val);
}
template <class Smem, class Raw>
__device__ __forceinline__ void stage_from_raw(Smem& sm, int slot, const Raw& r, int32_t val) {
  const float4 gm = r.gm[slot], col = r.col[slot];
  ChsSplat<float> s;
  chs_make_splat(gm.x, gm.y, gm.z, gm.w, r.cc[slot], col.w, col.x, col.y, col.z, s);
  sm.a[slot] = make_float4(s.mx, s.my, s.qa, s.r);
  sm.b[slot] = make_float4(s.kc, s.lo, s.rbc, s.cr);
  sm.c[slot] = make_float4(s.cg, s.cb, s.inv_opac, __int_as_float(val));
}

__device__ __forceinline__ unsigned lanemask_lt_() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}
__device__ __forceinline__ unsigned lanemask_gt_() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_gt;" : "=r"(m));
  return m;
}

// ---- forward ----
// kGroup (round 2, second half; the default): the warp first culls the WHOLE staged batch into its survivor list, pads the list
// to a multiple of four with the index of a null record (log2 opacity = -inf: alpha = 0 everywhere), and then walks it in
// groups of four.  A group evaluates the four alphas (independent of the transmittance), runs the transmittance chain
// T <- T - alpha T speculatively and takes ONE vote "did any pixel of the warp cross the 1e-4 stop threshold in this group?".
// Almost always nobody did: the colour sums take the four products as they are and nothing per Gaussian is spent on the
// stop logic (r2r source page of the ungrouped loop: 2 FSETP + 2 FSEL + 2 MOV + the reconvergence of a per-lane branch that
// 96 % of the iterations took anyway, per Gaussian).  If somebody did, the same four alphas go through the exact sequential
// code of the ungrouped loop, so results are bit-identical to it.
template <int kMinBlocks, bool kPerPoseCrf, bool kAsync, bool kGroup = false, int kFB = kBatch>
__global__ void __launch_bounds__(kThreads, kMinBlocks) blend_fwd2_kernel(BlendFwdArgs a) {
  static_assert(kFB == 2 * kThreads || (kFB == kThreads && kGroup && !kAsync), "two tile-list entries per thread and batch (one: grouped kernel only)");
  __shared__ SplatSmem3<kFB + (kGroup ? 1 : 0)> sm;  // kGroup: entry kFB is the null record
  __shared__ RawSmem<kAsync ? kFB : 1> raw;  // kAsync: the next batch's raw records, gathered with cp.async
  __shared__ __align__(16) int s_list[kThreads / 32][kGroup ? kFB + 4 : 32];
  extern __shared__ float s_crf[];  // the CRF parameters [3, stride] when the CRF is learned

  const int tile = blockIdx.x, frame = blockIdx.y;
  const int tx = tile % a.tile_w, ty = tile / a.tile_w;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bx = tx * CHS_TILE + (warp & 1) * 8, by = ty * CHS_TILE + (warp >> 1) * 8;
  const int ix = bx + (lane & 7), iyA = by + (lane >> 3), iyB = iyA + 4;
  const bool insideA = ix < a.W && iyA < a.H, insideB = ix < a.W && iyB < a.H;
  const float px = ix + 0.5f;
  const P2 py2 = p2(iyA + 0.5f, iyB + 0.5f);
  const float bx0 = bx + 0.5f, bx1 = bx + 7.5f, by0 = by + 0.5f, by1 = by + 7.5f;
  const int64_t P = (int64_t)a.W * a.H;
  const int64_t pixA = (int64_t)iyA * a.W + ix, pixB = (int64_t)iyB * a.W + ix;
  const float kInf = __int_as_float(0x7f800000);
  const unsigned lt = lanemask_lt_();
  int* list = s_list[warp];

  const int crf_stride = chs_crf_stride(a.crf_kind, a.crf_hidden);  // 0 for the identity CRF
  for (int i = tid; i < 3 * crf_stride; i += kThreads) s_crf[i] = a.crf_params[i];
  if constexpr (kGroup) if (tid == 0) {  // the null record (made visible by the first barrier of the batch loop)
    sm.a[kFB] = make_float4(0.f, 0.f, 0.f, 0.f);
    sm.b[kFB] = make_float4(0.f, -kInf, 0.f, 0.f);
    sm.c[kFB] = make_float4(0.f, 0.f, 0.f, 0.f);
  }

  constexpr bool per_pose_crf = kPerPoseCrf;
  const float dt = a.exposure[frame];
  if (per_pose_crf) __syncthreads();  // s_crf is read inside the pose loop
  P2 sum_r2 = p2s(0.f), sum_g2 = p2s(0.f), sum_b2 = p2s(0.f), sum_al2 = p2s(0.f);
  for (int k = 0; k < a.n_virtual; ++k) {
    const int c = frame * a.n_virtual + k;
    const int cam_base = a.rgbo_per_camera ? 0 : c * a.N;  // index of the camera's first record in rgbo
    const int lc = a.fused ? frame : c;                     // whose tile list this camera walks
    const int rec_shift = a.fused ? (c - frame) * a.N : 0;  // list entry frame * N + g -> record c * N + g
    const uint32_t start = a.tile_offsets[(int64_t)lc * a.tiles + tile];
    const uint32_t end = a.tile_offsets[(int64_t)lc * a.tiles + tile + 1];
    P2 T2 = p2s(1.f), acc_r2 = p2s(0.f), acc_g2 = p2s(0.f), acc_b2 = p2s(0.f);
    int lastA = 0, lastB = 0;
    // "done" is folded into the pixel's alpha threshold: a finished pixel has threshold +inf
    float thrA = insideA ? CHS_LOG2_ALPHA_MIN : kInf, thrB = insideB ? CHS_LOG2_ALPHA_MIN : kInf;
    bool warp_done = __all_sync(CHS_FULL_MASK, thrA == kInf && thrB == kInf);
    // kAsync pipeline: this thread's two entries of batch b + 1 are gathered (cp.async) while batch b is processed, and the
    // tile-list values of batch b + 2 are already on their way in registers
    int32_t val_cur[2] = {0, 0}, val_next[2] = {0, 0};
    if (kAsync) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t i0 = start + h * kThreads + tid, i1 = i0 + kFB;
        if (i0 < end) {
          val_cur[h] = a.vals[i0] + rec_shift;
          gather_raw_async(raw, h * kThreads + tid, val_cur[h], cam_base, a.geom, a.conic_c, a.rgbo);
        }
        if (i1 < end) val_next[h] = a.vals[i1] + rec_shift;
      }
      cp_async_commit();
    }
    int32_t vpre0 = 0, vpre1 = 0;  // kGroup: this thread's two tile-list entries of the batch about to be staged
    if (kGroup && !kAsync) {
      if (start + tid < end) vpre0 = a.vals[start + tid];
      if (kFB == 2 * kThreads && start + tid + kThreads < end) vpre1 = a.vals[start + tid + kThreads];
    }
    for (uint32_t base = start; base < end; base += kFB) {
      if (kAsync) cp_async_wait_all();  // this thread's own gathers of the batch have landed
      // barrier + CTA-wide early exit; also protects the staged batch of the previous iteration
      if (__syncthreads_and(thrA == kInf && thrB == kInf)) break;
      const int cnt = min((uint32_t)kFB, end - base);
      if (kAsync) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int slot = h * kThreads + tid;
          if (slot < cnt) stage_from_raw(sm, slot, raw, val_cur[h]);
          val_cur[h] = val_next[h];
          if (base + kFB + slot < end) gather_raw_async(raw, slot, val_cur[h], cam_base, a.geom, a.conic_c, a.rgbo);
          const uint32_t i2 = base + 2 * kFB + slot;
          if (i2 < end) val_next[h] = a.vals[i2] + rec_shift;
        }
        cp_async_commit();
      } else if (kGroup) {
        // the thread's two list entries were requested during the previous batch; the next batch's are requested now
        if (kFB == 2 * kThreads) {
          stage_splat3_pair(sm, tid, tid < cnt, vpre0 + rec_shift, tid + kThreads, tid + kThreads < cnt, vpre1 + rec_shift, cam_base, a.geom,
                            a.conic_c, a.rgbo);
        } else if (tid < cnt) {
          stage_splat3(sm, tid, vpre0 + rec_shift, cam_base, a.geom, a.conic_c, a.rgbo);
        }
        const uint32_t nb = base + kFB + tid;
        if (nb < end) vpre0 = a.vals[nb];
        if (kFB == 2 * kThreads && nb + kThreads < end) vpre1 = a.vals[nb + kThreads];
      } else {
        for (int i = tid; i < cnt; i += kThreads) stage_splat3(sm, i, a.vals[base + i] + rec_shift, cam_base, a.geom, a.conic_c, a.rgbo);
      }
      __syncthreads();
      if (warp_done) continue;
      const int idx0 = (int)(base - start) + 1;
      int relA = -1, relB = -1;  // staged index of the last Gaussian accumulated from this batch
      if constexpr (kGroup) {
        int n_surv = 0;
        for (int sub = 0; sub < cnt; sub += 32) {
          const int j = sub + lane;
          const bool hit = (j < cnt) && splat_hits_block3(sm, j, bx0, bx1, by0, by1);
          const unsigned mask = __ballot_sync(CHS_FULL_MASK, hit);
          if (hit) list[n_surv + __popc(mask & lt)] = j;  // ascending: front to back
          n_surv += __popc(mask);
        }
        if (lane < 3) list[n_surv + lane] = kFB;  // padding: the null record
        __syncwarp();
        for (int i = 0; i < n_surv; i += 4) {
          const int4 j4 = *reinterpret_cast<const int4*>(list + i);
          const int js[4] = {j4.x, j4.y, j4.z, j4.w};
          P2 al2[4], wq2[4];
          float cr[4];
          int srelA = relA, srelB = relB;  // speculative: committed when nobody stops in this group
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float4 sa = sm.a[js[u]];
            const float4 sb = sm.b[js[u]];  // kc, log2(opacity), rbc, cr
            float dx;
            P2 dy2, u2;
            const P2 pw2 = pair_power2(sa, sb.x, sb.y, px, py2, dx, dy2, u2);
            const float pA = p2lo(pw2), pB = p2hi(pw2);
            const bool actA = pA >= thrA, actB = pB >= thrB;
            al2[u] = p2(actA ? fminf(CHS_ALPHA_MAX, chs_exp2_fast(pA)) : 0.f, actB ? fminf(CHS_ALPHA_MAX, chs_exp2_fast(pB)) : 0.f);
            cr[u] = sb.w;
            srelA = actA ? js[u] : srelA;
            srelB = actB ? js[u] : srelB;
          }
          P2 Tq2 = T2;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            wq2[u] = al2[u] * Tq2;
            Tq2 = Tq2 - wq2[u];  // = T (1 - alpha)
          }
          const bool cross = fminf(p2lo(Tq2), p2hi(Tq2)) <= CHS_T_STOP;
          if (!__any_sync(CHS_FULL_MASK, cross)) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float2 cgb = *reinterpret_cast<const float2*>(&sm.c[js[u]]);  // cg, cb
              acc_r2 = fma2(p2s(cr[u]), wq2[u], acc_r2);
              acc_g2 = fma2(p2s(cgb.x), wq2[u], acc_g2);
              acc_b2 = fma2(p2s(cgb.y), wq2[u], acc_b2);
            }
            T2 = Tq2;
            relA = srelA;
            relB = srelB;
          } else {  // somebody stops inside this group: the ungrouped loop's exact sequence on the same alphas
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              // a pixel that stopped earlier in the group ignores the rest of it (its threshold is +inf from then on)
              const float alA = thrA == kInf ? 0.f : p2lo(al2[u]), alB = thrB == kInf ? 0.f : p2hi(al2[u]);
              const bool actA = alA != 0.f, actB = alB != 0.f;
              const P2 wq = p2(alA, alB) * T2;
              const P2 Tn2 = T2 - wq;
              const bool accA = actA && p2lo(Tn2) > CHS_T_STOP, accB = actB && p2hi(Tn2) > CHS_T_STOP;
              thrA = (actA && !accA) ? kInf : thrA;
              thrB = (actB && !accB) ? kInf : thrB;
              const P2 w2 = p2(accA ? p2lo(wq) : 0.f, accB ? p2hi(wq) : 0.f);
              const float2 cgb = *reinterpret_cast<const float2*>(&sm.c[js[u]]);  // cg, cb
              acc_r2 = fma2(p2s(cr[u]), w2, acc_r2);
              acc_g2 = fma2(p2s(cgb.x), w2, acc_g2);
              acc_b2 = fma2(p2s(cgb.y), w2, acc_b2);
              T2 = T2 - w2;
              relA = accA ? js[u] : relA;
              relB = accB ? js[u] : relB;
            }
            if (__all_sync(CHS_FULL_MASK, thrA == kInf && thrB == kInf)) {
              warp_done = true;
              break;
            }
          }
        }
      } else
      for (int sub = 0; sub < cnt; sub += 32) {
        const int j = sub + lane;
        const bool hit = (j < cnt) && splat_hits_block3(sm, j, bx0, bx1, by0, by1);
        const unsigned mask = __ballot_sync(CHS_FULL_MASK, hit);
        if (mask == 0u) continue;
        const int n_surv = __popc(mask);
        if (hit) list[__popc(mask & lt)] = j;  // ascending: front to back
        __syncwarp();
        for (int i = 0; i < n_surv; ++i) {
          const int jj = list[i];
          const float4 sa = sm.a[jj];
          const float4 sb = sm.b[jj];  // kc, log2(opacity), rbc, cr
          float dx;
          P2 dy2, u2;
          const P2 pw2 = pair_power2(sa, sb.x, sb.y, px, py2, dx, dy2, u2);
          const float pA = p2lo(pw2), pB = p2hi(pw2);
          const bool actA = pA >= thrA, actB = pB >= thrB;
          if (actA || actB) {
            const float alA = actA ? fminf(CHS_ALPHA_MAX, chs_exp2_fast(pA)) : 0.f;
            const float alB = actB ? fminf(CHS_ALPHA_MAX, chs_exp2_fast(pB)) : 0.f;
            const P2 wq2 = p2(alA, alB) * T2;
            const P2 Tn2 = T2 - wq2;  // = T (1 - alpha)
            // stop (this Gaussian is not accumulated) when the transmittance would drop to <= 1e-4
            const bool accA = actA && p2lo(Tn2) > CHS_T_STOP, accB = actB && p2hi(Tn2) > CHS_T_STOP;
            thrA = (actA && !accA) ? kInf : thrA;
            thrB = (actB && !accB) ? kInf : thrB;
            const P2 w2 = p2(accA ? p2lo(wq2) : 0.f, accB ? p2hi(wq2) : 0.f);
            const float2 cgb = *reinterpret_cast<const float2*>(&sm.c[jj]);  // cg, cb
            acc_r2 = fma2(p2s(sb.w), w2, acc_r2);
            acc_g2 = fma2(p2s(cgb.x), w2, acc_g2);
            acc_b2 = fma2(p2s(cgb.y), w2, acc_b2);
            T2 = T2 - w2;
            relA = accA ? jj : relA;
            relB = accB ? jj : relB;
          }
        }
        __syncwarp();  // the list is rewritten by the next sub-batch
        if (__all_sync(CHS_FULL_MASK, thrA == kInf && thrB == kInf)) {
          warp_done = true;
          break;
        }
      }
      lastA = relA >= 0 ? idx0 + relA : lastA;
      lastB = relB >= 0 ? idx0 + relB : lastB;
    }
    if (kAsync) cp_async_wait_all();  // (early exit) gathers of a batch nobody will read must land before the buffer is reused
    __syncthreads();  // the next pose restages shared memory
    if (insideA) {
      a.final_T[(int64_t)c * P + pixA] = p2lo(T2);
      a.last_id[(int64_t)c * P + pixA] = lastA;
    }
    if (insideB) {
      a.final_T[(int64_t)c * P + pixB] = p2hi(T2);
      a.last_id[(int64_t)c * P + pixB] = lastB;
    }
    const P2 h_r2 = fma2(T2, p2s(a.bg[0]), acc_r2), h_g2 = fma2(T2, p2s(a.bg[1]), acc_g2), h_b2 = fma2(T2, p2s(a.bg[2]), acc_b2);
    sum_al2 = sum_al2 + (p2s(1.f) - T2);
    if (!per_pose_crf) {
      sum_r2 = sum_r2 + h_r2;
      sum_g2 = sum_g2 + h_g2;
      sum_b2 = sum_b2 + h_b2;
    } else {
      float y[2][3];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float hr = h == 0 ? p2lo(h_r2) : p2hi(h_r2), hg = h == 0 ? p2lo(h_g2) : p2hi(h_g2), hb = h == 0 ? p2lo(h_b2) : p2hi(h_b2);
        y[h][0] = dt * hr; y[h][1] = dt * hg; y[h][2] = dt * hb;
        if (a.crf_kind != CHS_CRF_IDENTITY) {
          y[h][0] = chs_crf_fwd(a.crf_kind, y[h][0], s_crf, a.crf_hidden);
          y[h][1] = chs_crf_fwd(a.crf_kind, y[h][1], s_crf + crf_stride, a.crf_hidden);
          y[h][2] = chs_crf_fwd(a.crf_kind, y[h][2], s_crf + 2 * crf_stride, a.crf_hidden);
        }
        if (h == 0 ? insideA : insideB) {
          const int64_t o = ((int64_t)c * P + (h == 0 ? pixA : pixB)) * 3;
          a.hdr_mean[o] = hr; a.hdr_mean[o + 1] = hg; a.hdr_mean[o + 2] = hb;
        }
      }
      sum_r2 = sum_r2 + p2(y[0][0], y[1][0]);
      sum_g2 = sum_g2 + p2(y[0][1], y[1][1]);
      sum_b2 = sum_b2 + p2(y[0][2], y[1][2]);
    }
  }
  // formation epilogue (A.7, decision D0): mean over poses, x exposure, CRF
  const float inv_n = 1.f / (float)a.n_virtual;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (!(h == 0 ? insideA : insideB)) continue;
    const float hr = (h == 0 ? p2lo(sum_r2) : p2hi(sum_r2)) * inv_n;
    const float hg = (h == 0 ? p2lo(sum_g2) : p2hi(sum_g2)) * inv_n;
    const float hb = (h == 0 ? p2lo(sum_b2) : p2hi(sum_b2)) * inv_n;
    const float al = (h == 0 ? p2lo(sum_al2) : p2hi(sum_al2)) * inv_n;
    const int64_t pix = h == 0 ? pixA : pixB;
    if (per_pose_crf) {  // the sums already hold F(dt * H_k)
      const int64_t o = ((int64_t)frame * P + pix) * 3;
      a.ldr[o] = hr; a.ldr[o + 1] = hg; a.ldr[o + 2] = hb;
      a.alpha[(int64_t)frame * P + pix] = al;
      continue;
    }
    float o0 = dt * hr, o1 = dt * hg, o2 = dt * hb;
    if (a.crf_kind != CHS_CRF_IDENTITY) {
      o0 = chs_crf_fwd(a.crf_kind, o0, s_crf, a.crf_hidden);
      o1 = chs_crf_fwd(a.crf_kind, o1, s_crf + crf_stride, a.crf_hidden);
      o2 = chs_crf_fwd(a.crf_kind, o2, s_crf + 2 * crf_stride, a.crf_hidden);
    }
    const int64_t o = ((int64_t)frame * P + pix) * 3;
    a.ldr[o] = o0; a.ldr[o + 1] = o1; a.ldr[o + 2] = o2;
    a.hdr_mean[o] = hr; a.hdr_mean[o + 1] = hg; a.hdr_mean[o + 2] = hb;
    a.alpha[(int64_t)frame * P + pix] = al;
  }
}

// ---- backward ----
constexpr int kRow3 = 68;  // floats per table row: 64 pixel ids (id = 2 lane + h), the slot's staged index, padding to 16 bytes

template <int kSlots>
struct __align__(16) BwdWarp3 {
  float nvs[kSlots][kRow3];  // -dL/dsigma of (slot, pixel id); column 64: staged-batch index of the slot's Gaussian (int bits)
  float nf[kSlots][kRow3];   // -alpha * T of (slot, pixel id)
  float vh[3][64];           // dL/dH of the warp's pixels by pixel id (r, g, b planes)
  int list[32];              // survivors of the current 32-entry sub-batch, back to front
  int ent[kSlots];           // staged-batch index of every tabled slot's Gaussian
  int pad_[4 - kSlots % 4];
};

// phase B for the n_slots tabled Gaussians of this warp.  Lane = (slot k = lane % 8, part = lane / 8); a part covers the two
// pixel rows y0 = part and y0 + 4 of the warp's 8x8 block: pixel ids 16 part .. 16 part + 15, i.e. four 128-bit chunks
// [(x, y0), (x, y0 + 4), (x + 1, y0), (x + 1, y0 + 4)] for x = 0, 2, 4, 6.
template <int kSlots, class Smem, class Warp, class EntT>
__device__ __forceinline__ void bwd_round3(const Smem& sm, const Warp& ws, const EntT* __restrict__ ents, int null_ent, int n_slots,
                                           int lane, float bxc, float byc, const BlendBwdArgs& a) {
  static_assert(kSlots == 8, "phase B assigns one lane per (slot, row pair): 8 slots x 4 parts");
  __syncwarp();
  const int k = lane & 7, part = lane >> 3;
  const bool active = k < n_slots && (int)ents[k] != null_ent;
  P2 rowvs = p2s(0.f), rowt = p2s(0.f), S_xx = p2s(0.f), G6 = p2s(0.f), G7 = p2s(0.f), G8 = p2s(0.f);
  float4 sa = make_float4(0.f, 0.f, 0.f, 0.f);
  float kc = 0.f, inv_opac = 0.f, dy_lo = 0.f;
  uint32_t val = 0;
  if (active) {
    const int jj = (int)ents[k];
    sa = sm.a[jj];  // mx, my, qa, r
    kc = sm.b[jj].x;
    const float2 sc = *reinterpret_cast<const float2*>(&sm.c[jj].z);  // 1/opacity, val
    inv_opac = sc.x;
    val = (uint32_t)__float_as_int(sc.y);
    const float dxb = sa.x - bxc;  // mean - centre of pixel column 0
    dy_lo = sa.y - (byc + (float)part);
    const float* rv = &ws.nvs[k][16 * part];
    const float* rf = &ws.nf[k][16 * part];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float4 vs4 = *reinterpret_cast<const float4*>(rv + 4 * c);
      const float4 f4 = *reinterpret_cast<const float4*>(rf + 4 * c);
      const float4 r4 = *reinterpret_cast<const float4*>(&ws.vh[0][16 * part + 4 * c]);
      const float4 g4 = *reinterpret_cast<const float4*>(&ws.vh[1][16 * part + 4 * c]);
      const float4 b4 = *reinterpret_cast<const float4*>(&ws.vh[2][16 * part + 4 * c]);
      const P2 vA = p2(vs4.x, vs4.y), vB = p2(vs4.z, vs4.w);  // columns 2c and 2c + 1, rows (y0, y0 + 4)
      const P2 dxa = p2s(dxb - (float)(2 * c)), dxc = p2s(dxb - (float)(2 * c + 1));
      const P2 tA = vA * dxa, tB = vB * dxc;
      rowvs = rowvs + (vA + vB);
      rowt = rowt + (tA + tB);
      S_xx = fma2(tA, dxa, S_xx);
      S_xx = fma2(tB, dxc, S_xx);
      const P2 fA = p2(f4.x, f4.y), fB = p2(f4.z, f4.w);
      G6 = fma2(fA, p2(r4.x, r4.y), G6); G6 = fma2(fB, p2(r4.z, r4.w), G6);
      G7 = fma2(fA, p2(g4.x, g4.y), G7); G7 = fma2(fB, p2(g4.z, g4.w), G7);
      G8 = fma2(fA, p2(b4.x, b4.y), G8); G8 = fma2(fB, p2(b4.z, b4.w), G8);
    }
  }
  // the two halves of every packed sum are the rows y0 (dy = dy_lo) and y0 + 4 (dy = dy_lo - 4)
  const P2 dyv = p2(dy_lo, dy_lo - 4.f);
  const P2 S_dy = rowvs * dyv, S_xy = rowt * dyv;
  const P2 S_yy = S_dy * dyv;
  const float s_dy = p2sum(S_dy);
  float t[9] = {fmaf(sa.w, s_dy, p2sum(rowt)), s_dy, p2sum(S_xx), p2sum(S_xy), p2sum(S_yy), p2sum(rowvs), p2sum(G6), p2sum(G7), p2sum(G8)};
#pragma unroll
  for (int o = 8; o < 32; o <<= 1)
#pragma unroll
    for (int i = 0; i < 9; ++i) t[i] += __shfl_xor_sync(CHS_FULL_MASK, t[i], o);
  if (active) {
    ChsSplat<float> sp;
    sp.qa = sa.z; sp.r = sa.w; sp.kc = kc; sp.inv_opac = inv_opac;
    float g[9];
    chs_moments_to_grads_neg(sp, t, g);  // csrc/chs_math.cuh, checked on the host against chs_pair_bwd and the oracle
    if (part == 0) atomicAdd(a.v_geom + val, make_float4(g[0], g[1], g[2], g[3]));
    if (part == 1) atomicAdd(a.v_cogr + val, make_float4(g[4], g[5], g[6], g[7]));
    if (part == 2) atomicAdd(a.v_blue + val, g[8]);
  }
  __syncwarp();  // the table is rewritten by the next round
}

template <int kSlots, int kB, int kMinBlocks, bool kAsync>
__global__ void __launch_bounds__(kThreads, kMinBlocks) blend_bwd3_kernel(BlendBwdArgs a) {
  static_assert(!kAsync || kB == kThreads, "asynchronous staging: one tile-list entry per thread and batch");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using Smem = SplatSmem3<kB>;
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  // kAsync: raw records of the next batch (one per thread), behind the per-warp tables
  RawSmem<kB>& raw = *reinterpret_cast<RawSmem<kB>*>(smem_raw + sizeof(Smem) + 4 * sizeof(BwdWarp3<kSlots>));
  __shared__ int s_max_last;

  const int tile = blockIdx.x, c = blockIdx.y;
  const int frame = c / a.n_virtual;
  const int cam_base = a.rgbo_per_camera ? 0 : c * a.N;
  const int tx = tile % a.tile_w, ty = tile / a.tile_w;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  BwdWarp3<kSlots>& ws = reinterpret_cast<BwdWarp3<kSlots>*>(smem_raw + sizeof(Smem))[warp];
  const int bx = tx * CHS_TILE + (warp & 1) * 8, by = ty * CHS_TILE + (warp >> 1) * 8;
  const int ix = bx + (lane & 7), iy0 = by + (lane >> 3);
  const float px = ix + 0.5f;
  const float bx0 = bx + 0.5f, bx1 = bx + 7.5f, by0 = by + 0.5f, by1 = by + 7.5f;
  const int64_t P = (int64_t)a.W * a.H;
  const int lc = a.fused ? frame : c;                          // whose tile list this camera walks
  const int rec_shift = a.fused ? (c - frame) * a.N : 0;       // list entry frame * N + g -> record c * N + g
  const uint32_t start = a.tile_offsets[(int64_t)lc * a.tiles + tile];
  const uint32_t end = a.tile_offsets[(int64_t)lc * a.tiles + tile + 1];
  if (end <= start) return;

  const float inv_nv = 1.f / (float)a.n_virtual;
  const int vimg = a.v_hdr_per_camera ? c : frame;
  int last[2];
  int warp_last = 0;
  float Tf[2] = {1.f, 1.f}, v[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}}, r0[2] = {0.f, 0.f};
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int iy = iy0 + 4 * h;
    last[h] = 0;
    if (ix < a.W && iy < a.H) {
      const int64_t pix = (int64_t)iy * a.W + ix;
      Tf[h] = a.final_T[(int64_t)c * P + pix];
      last[h] = a.last_id[(int64_t)c * P + pix];
      const int64_t o = ((int64_t)vimg * P + pix) * 3;
      v[h][0] = a.v_hdr[o]; v[h][1] = a.v_hdr[o + 1]; v[h][2] = a.v_hdr[o + 2];
      const float v_al = a.v_alpha ? a.v_alpha[(int64_t)frame * P + pix] * inv_nv : 0.f;
      // behind the last accumulated Gaussian lies the background: R = bg . v_H - v_alpha (chs_pair_bwd_scalars_r)
      r0[h] = (a.bg[0] * v[h][0] + a.bg[1] * v[h][1] + a.bg[2] * v[h][2]) - v_al;
    }
    warp_last = max(warp_last, last[h]);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) ws.vh[ch][2 * lane + h] = v[h][ch];  // pixel id = 2 lane + h  <->  (x, y) = (lane & 7, (lane >> 3) + 4 h)
  }
  const P2 py2 = p2(iy0 + 0.5f, iy0 + 4.5f);
  P2 Tr2 = p2(Tf[0], Tf[1]);
  P2 R2 = p2(r0[0], r0[1]);
  const P2 vh_r2 = p2(v[0][0], v[1][0]), vh_g2 = p2(v[0][1], v[1][1]), vh_b2 = p2(v[0][2], v[1][2]);
  if (tid == 0) s_max_last = 0;
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) warp_last = max(warp_last, __shfl_xor_sync(CHS_FULL_MASK, warp_last, o));
  if (lane == 0 && warp_last > 0) atomicMax(&s_max_last, warp_last);
  __syncthreads();
  const int n_walk = s_max_last;

  const unsigned gt = lanemask_gt_();
  // The table is written through a 32-bit shared-memory address that is carried across iterations (the lane's float2 slot in the
  // next free row): one add per tabled Gaussian, no address arithmetic from the thread id inside the loop.
  const uint32_t tab0 = (uint32_t)__cvta_generic_to_shared(&ws.nvs[0][0]) + 8u * (uint32_t)lane;
  constexpr uint32_t kRowBytes = kRow3 * 4, kFOffBytes = kSlots * kRow3 * 4;
  uint32_t tab = tab0;
  int slots_left = kSlots;  // free rows of the table (a down-counter: one subtract-and-test per tabled Gaussian)
  // kAsync pipeline: this thread's entry of batch b + 1 is gathered (cp.async) while batch b is processed, and the tile-list
  // value of batch b + 2 is already on its way in a register
  int32_t val_cur = 0, val_next = 0;  // record index whose gather is in flight / of the batch after it
  if (kAsync) {
    const int lo0 = max(0, n_walk - kB);
    if (tid < n_walk - lo0) {
      val_cur = a.vals[start + lo0 + tid] + rec_shift;
      gather_raw_async(raw, tid, val_cur, cam_base, a.geom, a.conic_c, a.rgbo);
    }
    cp_async_commit();
    const int hi1 = n_walk - kB, lo1 = max(0, hi1 - kB);
    if (hi1 > 0 && tid < hi1 - lo1) val_next = a.vals[start + lo1 + tid] + rec_shift;
  }
  for (int hi = n_walk; hi > 0; hi -= kB) {
    const int lo = max(0, hi - kB);
    const int cnt = hi - lo;
    if (kAsync) {
      cp_async_wait_all();  // this thread's own gathers of the batch have landed
      __syncthreads();      // every warp is done with the previous batch's records
      if (tid < cnt) stage_from_raw(sm, tid, raw, val_cur);
      const int hi1 = hi - kB, lo1 = max(0, hi1 - kB);
      val_cur = val_next;
      if (hi1 > 0 && tid < hi1 - lo1) gather_raw_async(raw, tid, val_cur, cam_base, a.geom, a.conic_c, a.rgbo);
      cp_async_commit();
      const int hi2 = hi1 - kB, lo2 = max(0, hi2 - kB);
      if (hi2 > 0 && tid < hi2 - lo2) val_next = a.vals[start + lo2 + tid] + rec_shift;
      __syncthreads();
    } else {
      __syncthreads();
      for (int i = tid; i < cnt; i += kThreads) stage_splat3(sm, i, a.vals[start + lo + i] + rec_shift, cam_base, a.geom, a.conic_c, a.rgbo);
      __syncthreads();
    }
    if (warp_last <= lo) continue;
    const int lrA = last[0] - lo - 1, lrB = last[1] - lo - 1;  // staged indices <= lr are inside the pixel's accumulated prefix
    for (int sub_hi = cnt; sub_hi > 0; sub_hi -= 32) {
      const int sub_lo = max(0, sub_hi - 32);
      if (warp_last <= lo + sub_lo) continue;
      const int j = sub_lo + lane;
      const bool hit = (j < sub_hi) && (lo + j < warp_last) && splat_hits_block3(sm, j, bx0, bx1, by0, by1);
      const unsigned mask = __ballot_sync(CHS_FULL_MASK, hit);
      if (mask == 0u) continue;
      const int n_surv = __popc(mask);
      if (hit) ws.list[__popc(mask & gt)] = j;  // descending: back to front
      __syncwarp();
      // ---- phase A over the survivors, in chunks that fit the free rows of the table: the inner loop has no table-full test,
      // no call and no vote (r2r source page: 99 % of the survivors have a contributing pixel in the warp, so the "any valid"
      // vote + branch cost more than the iterations they skipped), which lets the compiler unroll it ----
      int i = 0;
      while (i < n_surv) {
        const int m = min(n_surv - i, slots_left);
        if (lane < m) ws.ent[kSlots - slots_left + lane] = ws.list[i + lane];  // staged indices of the chunk's Gaussians, for phase B
#pragma unroll 4
        for (int q = 0; q < m; ++q) {
          const int jj = ws.list[i + q];
          const float4 sa = sm.a[jj];  // mx, my, qa, r
          const float4 sb = sm.b[jj];  // kc, log2(opacity), rbc, cr
          float dx;
          P2 dy2, u2;
          const P2 pw2 = pair_power2(sa, sb.x, sb.y, px, py2, dx, dy2, u2);
          const float pA = p2lo(pw2), pB = p2hi(pw2);
          const bool validA = (jj <= lrA) && pA >= CHS_LOG2_ALPHA_MIN;
          const bool validB = (jj <= lrB) && pB >= CHS_LOG2_ALPHA_MIN;
          const float2 cgb = *reinterpret_cast<const float2*>(&sm.c[jj]);  // cg, cb
          const float auA = validA ? chs_exp2_fast(pA) : 0.f;
          const float auB = validB ? chs_exp2_fast(pB) : 0.f;
          // packed chs_pair_bwd_scalars_r with na = -alpha: a pixel that does not contribute runs with alpha = 0, which leaves
          // T and R untouched and tables zeros
          const P2 na2 = p2(fmaxf(-CHS_ALPHA_MAX, -auA), fmaxf(-CHS_ALPHA_MAX, -auB));
          const P2 om2 = p2s(1.f) + na2;
          Tr2 = Tr2 * p2(chs_rcp_fast(p2lo(om2)), chs_rcp_fast(p2hi(om2)));  // transmittance before this Gaussian
          const P2 nf2 = na2 * Tr2;
          const P2 s2 = fma2(p2s(cgb.y), vh_b2, fma2(p2s(cgb.x), vh_g2, p2s(sb.w) * vh_r2));
          const P2 e2 = R2 - s2;
          R2 = fma2(na2, e2, R2);
          // no gradient through the 0.999 clamp
          const P2 gate2 = p2(auA <= CHS_ALPHA_MAX ? 1.f : 0.f, auB <= CHS_ALPHA_MAX ? 1.f : 0.f);
          const P2 nvs2 = (nf2 * e2) * gate2;
          asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(tab), "f"(p2lo(nvs2)), "f"(p2hi(nvs2)) : "memory");
          asm volatile("st.shared.v2.f32 [%0+%3], {%1, %2};" ::"r"(tab), "f"(p2lo(nf2)), "f"(p2hi(nf2)), "n"(kFOffBytes) : "memory");
          tab += kRowBytes;
        }
        i += m;
        slots_left -= m;
        if (slots_left == 0) {
          bwd_round3<kSlots>(sm, ws, ws.ent, -1, kSlots, lane, bx0, by0, a);
          tab = tab0;
          slots_left = kSlots;
        }
      }
      __syncwarp();  // the list is rewritten by the next sub-batch
    }
    if (slots_left != kSlots) {  // the staged batch is about to be replaced
      bwd_round3<kSlots>(sm, ws, ws.ent, -1, kSlots - slots_left, lane, bx0, by0, a);
      tab = tab0;
      slots_left = kSlots;
    }
  }
}


// ---------------------------------------------------------------------------------------------
// blend_bwd5_kernel (r3c): the fp32-table backward with phase A written for instruction-level parallelism.  ncu of
// blend_bwd3_kernel: every issued instruction waits 2.0 cycles on a fixed-latency dependency and 2.4 on a shared-memory /
// MUFU result, because one Gaussian's chain (record load -> exponent -> ex2 -> rcp -> T -> table store) runs start to finish
// before the next one's begins.  Here the warp culls the whole staged batch into a byte list padded with the null record
// (as the grouped forward does) and takes the survivors four at a time in two stages: stage 1 evaluates everything that does
// not depend on the running state for all four (alpha, 1 / (1 - alpha), c . v, the clamp gate: four independent streams),
// stage 2 runs the two short recurrences (T: one multiply, R: subtract + fma) and stores the table rows.  Phase B is
// bwd_round3 after every second group.
// ---------------------------------------------------------------------------------------------
template <int kB, int kRows>
struct __align__(16) BwdWarp5 {
  float nvs[kRows][kRow3];  // -dL/dsigma of (slot, pixel id)
  float nf[kRows][kRow3];   // -alpha * T of (slot, pixel id)
  float vh[3][64];          // dL/dH of the warp's pixels by pixel id (r, g, b planes)
  typedef typename std::conditional<(kB < 255), uint8_t, uint16_t>::type ent_t;
  ent_t list[kB + 4];       // survivors of the staged batch, back to front, padded to a multiple of four with kB (the null record)
};

// phase B over kSets x 8 tabled slots: lane = (k = lane % 8, part = lane / 8) sums the slots k, k + 8, ... over the part's two pixel
// rows.  With two sets the warp's dL/dH (a quarter of phase B's shared-memory wavefronts) is read once for two Gaussians.
template <int kSets, class Smem, class Warp, class EntT>
__device__ __forceinline__ void bwd_round5(const Smem& sm, const Warp& ws, const EntT* __restrict__ ents, int null_ent, int n_slots,
                                           int lane, float bxc, float byc, const BlendBwdArgs& a) {
  __syncwarp();
  const int k = lane & 7, part = lane >> 3;
  bool active[kSets];
  P2 rowvs[kSets], rowt[kSets], S_xx[kSets], G6[kSets], G7[kSets], G8[kSets];
  float dxb[kSets];
  int jj[kSets];
#pragma unroll
  for (int s = 0; s < kSets; ++s) {
    jj[s] = k + 8 * s < n_slots ? (int)ents[k + 8 * s] : null_ent;
    active[s] = jj[s] != null_ent;
    rowvs[s] = rowt[s] = S_xx[s] = G6[s] = G7[s] = G8[s] = p2s(0.f);
    dxb[s] = sm.a[jj[s]].x - bxc;  // (the null record is a valid staged entry) mean - centre of pixel column 0
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 r4 = *reinterpret_cast<const float4*>(&ws.vh[0][16 * part + 4 * c]);
    const float4 g4 = *reinterpret_cast<const float4*>(&ws.vh[1][16 * part + 4 * c]);
    const float4 b4 = *reinterpret_cast<const float4*>(&ws.vh[2][16 * part + 4 * c]);
#pragma unroll
    for (int s = 0; s < kSets; ++s) {
      const float4 vs4 = *reinterpret_cast<const float4*>(&ws.nvs[k + 8 * s][16 * part + 4 * c]);
      const float4 f4 = *reinterpret_cast<const float4*>(&ws.nf[k + 8 * s][16 * part + 4 * c]);
      const P2 vA = p2(vs4.x, vs4.y), vB = p2(vs4.z, vs4.w);  // columns 2c and 2c + 1, rows (y0, y0 + 4)
      const P2 dxa = p2s(dxb[s] - (float)(2 * c)), dxc = p2s(dxb[s] - (float)(2 * c + 1));
      const P2 tA = vA * dxa, tB = vB * dxc;
      rowvs[s] = rowvs[s] + (vA + vB);
      rowt[s] = rowt[s] + (tA + tB);
      S_xx[s] = fma2(tA, dxa, S_xx[s]);
      S_xx[s] = fma2(tB, dxc, S_xx[s]);
      const P2 fA = p2(f4.x, f4.y), fB = p2(f4.z, f4.w);
      G6[s] = fma2(fA, p2(r4.x, r4.y), G6[s]); G6[s] = fma2(fB, p2(r4.z, r4.w), G6[s]);
      G7[s] = fma2(fA, p2(g4.x, g4.y), G7[s]); G7[s] = fma2(fB, p2(g4.z, g4.w), G7[s]);
      G8[s] = fma2(fA, p2(b4.x, b4.y), G8[s]); G8[s] = fma2(fB, p2(b4.z, b4.w), G8[s]);
    }
  }
#pragma unroll
  for (int s = 0; s < kSets; ++s) {
    const float4 sa = sm.a[jj[s]];  // mx, my, qa, r
    const float dy_lo = sa.y - (byc + (float)part);
    // the two halves of every packed sum are the rows y0 (dy = dy_lo) and y0 + 4 (dy = dy_lo - 4)
    const P2 dyv = p2(dy_lo, dy_lo - 4.f);
    const P2 S_dy = rowvs[s] * dyv, S_xy = rowt[s] * dyv;
    const P2 S_yy = S_dy * dyv;
    const float s_dy = p2sum(S_dy);
    float t[9] = {fmaf(sa.w, s_dy, p2sum(rowt[s])), s_dy, p2sum(S_xx[s]), p2sum(S_xy), p2sum(S_yy), p2sum(rowvs[s]), p2sum(G6[s]),
                  p2sum(G7[s]), p2sum(G8[s])};
#pragma unroll
    for (int o = 8; o < 32; o <<= 1)
#pragma unroll
      for (int i = 0; i < 9; ++i) t[i] += __shfl_xor_sync(CHS_FULL_MASK, t[i], o);
    if (active[s]) {
      const float2 sc = *reinterpret_cast<const float2*>(&sm.c[jj[s]].z);  // 1/opacity, val
      const uint32_t val = (uint32_t)__float_as_int(sc.y);
      ChsSplat<float> sp;
      sp.qa = sa.z; sp.r = sa.w; sp.kc = sm.b[jj[s]].x; sp.inv_opac = sc.x;
      float g[9];
      chs_moments_to_grads_neg(sp, t, g);  // csrc/chs_math.cuh, checked on the host against chs_pair_bwd and the oracle
      if (part == 0) atomicAdd(a.v_geom + val, make_float4(g[0], g[1], g[2], g[3]));
      if (part == 1) atomicAdd(a.v_cogr + val, make_float4(g[4], g[5], g[6], g[7]));
      if (part == 2) atomicAdd(a.v_blue + val, g[8]);
    }
  }
  __syncwarp();  // the table is rewritten by the next round
}

template <int kB, int kMinBlocks, int kSets>
__global__ void __launch_bounds__(kThreads, kMinBlocks) blend_bwd5_kernel(BlendBwdArgs a) {
  constexpr int kRows = 8 * kSets;  // table rows = Gaussians per phase-B round
  static_assert(kB == kThreads || kB == 2 * kThreads, "one or two tile-list entries per thread and batch");
  constexpr int kPer = kB / kThreads;  // tile-list entries a thread stages per batch
  typedef typename BwdWarp5<kB, kRows>::ent_t ent_t;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using Smem = SplatSmem3<kB + 1>;  // entry kB: the null record
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  __shared__ int s_max_last;

  const int tile = blockIdx.x, c = blockIdx.y;
  const int frame = c / a.n_virtual;
  const int cam_base = a.rgbo_per_camera ? 0 : c * a.N;
  const int tx = tile % a.tile_w, ty = tile / a.tile_w;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  BwdWarp5<kB, kRows>& ws = reinterpret_cast<BwdWarp5<kB, kRows>*>(smem_raw + sizeof(Smem))[warp];
  const int bx = tx * CHS_TILE + (warp & 1) * 8, by = ty * CHS_TILE + (warp >> 1) * 8;
  const int ix = bx + (lane & 7), iy0 = by + (lane >> 3);
  const float px = ix + 0.5f;
  const float bx0 = bx + 0.5f, bx1 = bx + 7.5f, by0 = by + 0.5f, by1 = by + 7.5f;
  const int64_t P = (int64_t)a.W * a.H;
  const int lc = a.fused ? frame : c;
  const int rec_shift = a.fused ? (c - frame) * a.N : 0;
  const uint32_t start = a.tile_offsets[(int64_t)lc * a.tiles + tile];
  const uint32_t end = a.tile_offsets[(int64_t)lc * a.tiles + tile + 1];
  if (end <= start) return;

  if (tid == 0) {
    s_max_last = 0;
    sm.a[kB] = make_float4(0.f, 0.f, 0.f, 0.f);
    sm.b[kB] = make_float4(0.f, -__int_as_float(0x7f800000), 0.f, 0.f);
    sm.c[kB] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float inv_nv = 1.f / (float)a.n_virtual;
  const int vimg = a.v_hdr_per_camera ? c : frame;
  int last[2];
  int warp_last = 0;
  float Tf[2] = {1.f, 1.f}, v[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}}, r0[2] = {0.f, 0.f};
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int iy = iy0 + 4 * h;
    last[h] = 0;
    if (ix < a.W && iy < a.H) {
      const int64_t pix = (int64_t)iy * a.W + ix;
      Tf[h] = a.final_T[(int64_t)c * P + pix];
      last[h] = a.last_id[(int64_t)c * P + pix];
      const int64_t o = ((int64_t)vimg * P + pix) * 3;
      v[h][0] = a.v_hdr[o]; v[h][1] = a.v_hdr[o + 1]; v[h][2] = a.v_hdr[o + 2];
      const float v_al = a.v_alpha ? a.v_alpha[(int64_t)frame * P + pix] * inv_nv : 0.f;
      r0[h] = (a.bg[0] * v[h][0] + a.bg[1] * v[h][1] + a.bg[2] * v[h][2]) - v_al;
    }
    warp_last = max(warp_last, last[h]);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) ws.vh[ch][2 * lane + h] = v[h][ch];  // pixel id = 2 lane + h
  }
  const P2 py2 = p2(iy0 + 0.5f, iy0 + 4.5f);
  P2 Tr2 = p2(Tf[0], Tf[1]);
  P2 R2 = p2(r0[0], r0[1]);
  const P2 vh_r2 = p2(v[0][0], v[1][0]), vh_g2 = p2(v[0][1], v[1][1]), vh_b2 = p2(v[0][2], v[1][2]);
  __syncthreads();
  warp_last = __reduce_max_sync(CHS_FULL_MASK, warp_last);
  if (lane == 0 && warp_last > 0) atomicMax(&s_max_last, warp_last);
  __syncthreads();
  const int n_walk = s_max_last;

  const unsigned gt = lanemask_gt_();
  const uint32_t tab0 = (uint32_t)__cvta_generic_to_shared(&ws.nvs[0][0]) + 8u * (uint32_t)lane;
  constexpr uint32_t kRowBytes = kRow3 * 4, kFOffBytes = kRows * kRow3 * 4;
  // this thread's tile-list entry of the NEXT batch is requested while the current batch is processed: the staging of a batch
  // then starts with the record gathers instead of a dependent list read
  int32_t val_next[kPer];
  {
    const int lo0 = max(0, n_walk - kB);
#pragma unroll
    for (int e = 0; e < kPer; ++e) {
      val_next[e] = 0;
      if (tid + e * kThreads < n_walk - lo0) val_next[e] = a.vals[start + lo0 + tid + e * kThreads];
    }
  }
  for (int hi = n_walk; hi > 0; hi -= kB) {
    const int lo = max(0, hi - kB);
    const int cnt = hi - lo;
    __syncthreads();
    if (kPer == 1) {
      if (tid < cnt) stage_splat3(sm, tid, val_next[0] + rec_shift, cam_base, a.geom, a.conic_c, a.rgbo);
    } else {
      stage_splat3_pair(sm, tid, tid < cnt, val_next[0] + rec_shift, tid + kThreads, tid + kThreads < cnt, val_next[kPer - 1] + rec_shift,
                        cam_base, a.geom, a.conic_c, a.rgbo);
    }
    {
      const int hi1 = lo, lo1 = max(0, hi1 - kB);
#pragma unroll
      for (int e = 0; e < kPer; ++e)
        if (hi1 > 0 && tid + e * kThreads < hi1 - lo1) val_next[e] = a.vals[start + lo1 + tid + e * kThreads];
    }
    __syncthreads();
    if (warp_last <= lo) continue;
    const int lrA = last[0] - lo - 1, lrB = last[1] - lo - 1;
    int n_surv = 0;
    for (int sub_hi = cnt; sub_hi > 0; sub_hi -= 32) {
      const int sub_lo = max(0, sub_hi - 32);
      if (warp_last <= lo + sub_lo) continue;
      const int j = sub_lo + lane;
      const bool hit = (j < sub_hi) && (lo + j < warp_last) && splat_hits_block3(sm, j, bx0, bx1, by0, by1);
      const unsigned mask = __ballot_sync(CHS_FULL_MASK, hit);
      if (hit) ws.list[n_surv + __popc(mask & gt)] = (ent_t)j;  // descending: back to front
      n_surv += __popc(mask);
    }
    if (lane < 3) ws.list[n_surv + lane] = (ent_t)kB;
    __syncwarp();
    for (int i = 0; i < n_surv; i += 4) {
      int js[4];
      if (sizeof(ent_t) == 1) {
        const uint32_t j4 = *reinterpret_cast<const uint32_t*>(ws.list + i);
        js[0] = j4 & 0xffu; js[1] = (j4 >> 8) & 0xffu; js[2] = (j4 >> 16) & 0xffu; js[3] = j4 >> 24;
      } else {
        const uint2 j4 = *reinterpret_cast<const uint2*>(ws.list + i);
        js[0] = j4.x & 0xffffu; js[1] = j4.x >> 16; js[2] = j4.y & 0xffffu; js[3] = j4.y >> 16;
      }
      const uint32_t tab = tab0 + (uint32_t)(i & (kRows - 4)) * kRowBytes;
      // ---- stage 1: what does not depend on T or R, for the four Gaussians ----
      P2 na2[4], ra2[4], s2[4], gate2[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int jj = js[u];
        const float4 sa = sm.a[jj];  // mx, my, qa, r
        const float4 sb = sm.b[jj];  // kc, log2(opacity), rbc, cr
        const float2 cgb = *reinterpret_cast<const float2*>(&sm.c[jj]);  // cg, cb
        float dx;
        P2 dy2, u2;
        const P2 pw2 = pair_power2(sa, sb.x, sb.y, px, py2, dx, dy2, u2);
        const float pA = p2lo(pw2), pB = p2hi(pw2);
        const bool validA = (jj <= lrA) && pA >= CHS_LOG2_ALPHA_MIN;
        const bool validB = (jj <= lrB) && pB >= CHS_LOG2_ALPHA_MIN;
        const float auA = validA ? chs_exp2_fast(pA) : 0.f;
        const float auB = validB ? chs_exp2_fast(pB) : 0.f;
        na2[u] = p2(fmaxf(-CHS_ALPHA_MAX, -auA), fmaxf(-CHS_ALPHA_MAX, -auB));
        const P2 om2 = p2s(1.f) + na2[u];
        ra2[u] = p2(chs_rcp_fast(p2lo(om2)), chs_rcp_fast(p2hi(om2)));
        s2[u] = fma2(p2s(cgb.y), vh_b2, fma2(p2s(cgb.x), vh_g2, p2s(sb.w) * vh_r2));
        gate2[u] = p2(auA <= CHS_ALPHA_MAX ? 1.f : 0.f, auB <= CHS_ALPHA_MAX ? 1.f : 0.f);  // no gradient through the 0.999 clamp
      }
      // ---- stage 2: the two recurrences (chs_pair_bwd_scalars_r) and the table rows ----
      auto row = [&](auto uc) {
        constexpr int u = decltype(uc)::value;
        Tr2 = Tr2 * ra2[u];  // transmittance before this Gaussian
        const P2 nf2 = na2[u] * Tr2;
        const P2 e2 = R2 - s2[u];
        R2 = fma2(na2[u], e2, R2);
        const P2 nvs2 = (nf2 * e2) * gate2[u];
        asm volatile("st.shared.v2.f32 [%0+%3], {%1, %2};" ::"r"(tab), "f"(p2lo(nvs2)), "f"(p2hi(nvs2)), "n"(u * kRowBytes));
        asm volatile("st.shared.v2.f32 [%0+%3], {%1, %2};" ::"r"(tab), "f"(p2lo(nf2)), "f"(p2hi(nf2)), "n"(u * kRowBytes + kFOffBytes));
      };
      row(std::integral_constant<int, 0>{});
      row(std::integral_constant<int, 1>{});
      row(std::integral_constant<int, 2>{});
      row(std::integral_constant<int, 3>{});
      if ((i & (kRows - 4)) == kRows - 4) bwd_round5<kSets>(sm, ws, ws.list + (i + 4 - kRows), kB, kRows, lane, bx0, by0, a);
    }
    const int n_pad = (n_surv + 3) & ~3;
    if (n_pad & (kRows - 4))  // groups left in the table when the batch ends
      bwd_round5<kSets>(sm, ws, ws.list + (n_pad & ~(kRows - 1)), kB, n_pad & (kRows - 1), lane, bx0, by0, a);
  }
}

// =================================================================================================
// Backward with phase B on the tensor cores (blend_bwd4_kernel; chs_config.tune_blend_bwd = 47).
//
// ncu of blend_bwd3_kernel (profiles/r2_ncu_blend.csv): l1tex__data_pipe_lsu_wavefronts 85 % of peak, shared-memory
// wavefronts 75 % — the kernel is bound by the shared-memory pipe, not by instruction issue (cutting phase A from 59 to 40
// instructions per pair moved it from 3.90 to 3.80 ms).  Of the ~22 wavefronts per (warp, Gaussian) pair, 6 were phase B
// re-reading the warp's dL/dH for every tabled Gaussian and 2.3 its 18 reduction shuffles.  Phase B is a contraction over
// the warp's 64 pixels,
//     S[slot][m] = sum_pix nvs[slot][pix] * mono_m(pix),  mono = (1, x, y, x^2, x y, y^2) relative to the block origin,
//     G[slot][c] = sum_pix nf[slot][pix] * v_c[pix],
// i.e. [8 slots x 64 pixels] x [64 pixels x 8 columns] products with a constant right-hand side: mma.sync.m16n8k16 (fp16 in,
// fp32 accumulate), four k-steps of 16 pixels.  fp32 accuracy is kept by splitting every tabled value into fp16 hi + lo, which
// ride in rows 0-7 / 8-15 of the SAME A operand (the two halves of the accumulator are added afterwards), and v_c into hi + lo
// in neighbouring COLUMNS of the same B operand; the monomials are small integers, exact in fp16.  fp16 has five exponent bits,
// so everything is scaled by powers of two into its range: v by the warp's max |v|, alpha T by 2^14 (carried in the
// transmittance register), and dL/dsigma by a bound that cannot be exceeded: |nvs| <= |R - s| <= 2 max(|R0|, max_j |s_j|) with
// |s_j| <= (|cr| + |cg| + |cb|) max|v| over every Gaussian staged so far (a running maximum kept by the staging code; R is a
// convex combination of R0 and earlier s_j).  hi + lo then resolves 2^-25 of that bound: tests/test_hostsim_math.py
// (test_tensor_core_phase_b_precision) pins geometry gradients at the accuracy of the fp32 table and colours at 1e-6.
// The sums come out per (slot, column pair) on the lanes of a quad, five shuffles gather what the three reducing lanes need,
// chs_shift_moments moves the origin-relative sums to the Gaussian's mean, and the REDs go out as before.
// Per pair: shared-memory wavefronts ~22 -> ~16, phase B instructions 24 -> 9, phase A 40 -> 52 (the four conversions).
// =================================================================================================
constexpr int kRow4 = 72;  // halves per table row: 64 pixel ids + 8 (144-byte rows: the 8 rows of an ldmatrix tile hit 8 different 16-byte bank groups)

template <int kB>
struct __align__(16) BwdWarp4 {
  __half tbl[4][8][kRow4];  // [nvs hi | nvs lo | nf hi | nf lo][slot][pixel id = 2 lane + h]
  uint2 vfrag[4 * 32];      // B operand of the colour product per k-step and lane: columns 2..7 = (r hi, r lo, g hi, g lo, b hi, b lo)
  uint8_t list[kB + 4];     // survivors of the staged batch, back to front, padded to a multiple of four with kB (the null record)
};

__device__ __forceinline__ uint32_t f16x2_of(float lo, float hi) {  // {hi, lo} -> packed halves, lo in bits 0..15
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ P2 f16x2_to_p2(uint32_t v) {
  float a, b;
  asm("{ .reg .b16 l, h; mov.b32 {l, h}, %2; cvt.f32.f16 %0, l; cvt.f32.f16 %1, h; }" : "=f"(a), "=f"(b) : "r"(v));
  return p2(a, b);
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
__device__ __forceinline__ void mma_f16(float d[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// 2^(k) as a float for an integer exponent already clamped to the normal range
__device__ __forceinline__ float pow2i(int k) { return __int_as_float((k + 127) << 23); }
__device__ __forceinline__ int expo_of(float x) { return (int)((__float_as_uint(x) >> 23) & 0xffu) - 126; }  // x = f 2^e, f in [0.5, 1)

// phase B for the slots [0, n_slots) of this warp's table; slot g's Gaussian is staged record ents[g].
// Lane = (slot g = lane >> 2, column pair t = lane & 3).  Geometry columns: (x, x^2 | 1, y | x y, y^2 | 1, y); colour columns:
// (0, 0 | r hi, r lo | g hi, g lo | b hi, b lo).  Lane t = 0 sends v_geom, t = 1 v_cogr, t = 3 v_blue.
template <int kB, class Smem>
__device__ __forceinline__ void bwd_round4(const Smem& sm, uint32_t a_addr, const uint2* __restrict__ mono, const uint2* __restrict__ vfrag,
                                           const uint8_t* __restrict__ ents, int n_slots, int lane, float bxc, float byc, float inv_sn,
                                           float inv_sc, const BlendBwdArgs& a) {
  __syncwarp();
  float dg[4] = {0.f, 0.f, 0.f, 0.f}, dc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    uint32_t a0, a1, a2, a3;
    ldmatrix_x4(a_addr + 32 * s, a0, a1, a2, a3);  // rows: nvs hi (slots) | nvs lo; 16 pixel ids of the k-step
    const uint2 bm = mono[s * 32 + lane];
    mma_f16(dg, a0, a1, a2, a3, bm.x, bm.y);
    ldmatrix_x4(a_addr + 2 * 8 * kRow4 * 2 + 32 * s, a0, a1, a2, a3);  // rows: nf hi | nf lo
    const uint2 bv = vfrag[s * 32 + lane];
    mma_f16(dc, a0, a1, a2, a3, bv.x, bv.y);
  }
  const int g = lane >> 2, t = lane & 3;
  const float P = dg[0] + dg[2], Q = dg[1] + dg[3];          // hi rows + lo rows
  const float C = (dc[0] + dc[2]) + (dc[1] + dc[3]);         // t = 1: sum nf v_r, t = 2: v_g, t = 3: v_b (hi + lo columns)
  const float P1 = __shfl_xor_sync(CHS_FULL_MASK, P, 1), Q1 = __shfl_xor_sync(CHS_FULL_MASK, Q, 1);
  const float P2x = __shfl_xor_sync(CHS_FULL_MASK, P, 2);
  const float Q3 = __shfl_xor_sync(CHS_FULL_MASK, Q, 3), C3 = __shfl_xor_sync(CHS_FULL_MASK, C, 3);
  const int jj = ents[g];
  if (g < n_slots && jj != kB && t != 2) {
    const float4 sa = sm.a[jj];  // mx, my, qa, r
    const float X = sa.x - bxc, Y = sa.y - byc;
    const float2 sc = *reinterpret_cast<const float2*>(&sm.c[jj].z);  // 1/opacity, val
    const uint32_t val = (uint32_t)__float_as_int(sc.y);
    if (t == 0) {  // P, Q = Sx, Sxx;  P1, Q1 = S1, Sy;  P2x = Sxy
      float m[6];
      chs_shift_moments(P1, P, Q1, Q, P2x, 0.f, X, Y, sa.w, m);
      ChsSplat<float> sp;
      sp.qa = sa.z; sp.r = sa.w; sp.kc = sm.b[jj].x; sp.inv_opac = 0.f;
      float t9[9] = {m[0] * inv_sn, m[1] * inv_sn, m[2] * inv_sn, m[3] * inv_sn, 0.f, 0.f, 0.f, 0.f, 0.f}, gr[9];
      chs_moments_to_grads_neg(sp, t9, gr);
      atomicAdd(a.v_geom + val, make_float4(gr[0], gr[1], gr[2], gr[3]));
    } else if (t == 1) {  // P, Q = S1, Sy;  Q3 = Syy;  C = sum nf v_r, C3 = sum nf v_g
      const float sdy = Y * P - Q;
      const float m4 = (Y * (sdy - Q) + Q3) * inv_sn, m5 = P * inv_sn;
      atomicAdd(a.v_cogr + val, make_float4(-0.5f * m4, m5 * sc.x, -C * inv_sc, -C3 * inv_sc));
    } else {  // t == 3: C = sum nf v_b
      atomicAdd(a.v_blue + val, -C * inv_sc);
    }
  }
  __syncwarp();  // the table is rewritten by the next round
}

template <int kB, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks) blend_bwd4_kernel(BlendBwdArgs a) {
  static_assert(kB == kThreads && kB < 255, "one tile-list entry per thread and batch; staged indices fit a byte");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using Smem = SplatSmem3<kB + 1>;  // entry kB: the null record
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  uint2* mono = reinterpret_cast<uint2*>(smem_raw + sizeof(Smem));  // [4 k-steps][32 lanes] B operand of the geometry product
  __shared__ int s_max_last;
  __shared__ unsigned s_cbnd;  // running max of |cr| + |cg| + |cb| over every staged Gaussian (float bits; non-negative)

  const int tile = blockIdx.x, c = blockIdx.y;
  const int frame = c / a.n_virtual;
  const int cam_base = a.rgbo_per_camera ? 0 : c * a.N;
  const int tx = tile % a.tile_w, ty = tile / a.tile_w;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  BwdWarp4<kB>& ws = reinterpret_cast<BwdWarp4<kB>*>(smem_raw + sizeof(Smem) + 4 * 32 * sizeof(uint2))[warp];
  const int bx = tx * CHS_TILE + (warp & 1) * 8, by = ty * CHS_TILE + (warp >> 1) * 8;
  const int ix = bx + (lane & 7), iy0 = by + (lane >> 3);
  const float px = ix + 0.5f;
  const float bx0 = bx + 0.5f, bx1 = bx + 7.5f, by0 = by + 0.5f, by1 = by + 7.5f;
  const int64_t P = (int64_t)a.W * a.H;
  const int lc = a.fused ? frame : c;                     // whose tile list this camera walks
  const int rec_shift = a.fused ? (c - frame) * a.N : 0;  // list entry frame * N + g -> record c * N + g
  const uint32_t start = a.tile_offsets[(int64_t)lc * a.tiles + tile];
  const uint32_t end = a.tile_offsets[(int64_t)lc * a.tiles + tile + 1];
  if (end <= start) return;

  const float kInf = __int_as_float(0x7f800000);
  if (tid == 0) {
    s_max_last = 0;
    s_cbnd = 0u;
    sm.a[kB] = make_float4(0.f, 0.f, 0.f, 0.f);
    sm.b[kB] = make_float4(0.f, -kInf, 0.f, 0.f);
    sm.c[kB] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  {  // geometry B operand: k-step s, lane (g, t): b0 = column g at pixels (x = t, y = s), (t, s + 4); b1 = the same at x = t + 4
    const int s = tid >> 5, g = lane >> 2, t = lane & 3;
    float v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float x = (float)(t + 4 * (q >> 1)), y = (float)(s + 4 * (q & 1));
      v[q] = g == 0 ? x : g == 1 ? x * x : (g == 2 || g == 6) ? 1.f : (g == 3 || g == 7) ? y : g == 4 ? x * y : y * y;
    }
    mono[tid] = make_uint2(f16x2_of(v[0], v[1]), f16x2_of(v[2], v[3]));
  }

  const float inv_nv = 1.f / (float)a.n_virtual;
  const int vimg = a.v_hdr_per_camera ? c : frame;
  int last[2];
  int warp_last = 0;
  float Tf[2] = {1.f, 1.f}, v[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}}, r0[2] = {0.f, 0.f};
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int iy = iy0 + 4 * h;
    last[h] = 0;
    if (ix < a.W && iy < a.H) {
      const int64_t pix = (int64_t)iy * a.W + ix;
      Tf[h] = a.final_T[(int64_t)c * P + pix];
      last[h] = a.last_id[(int64_t)c * P + pix];
      const int64_t o = ((int64_t)vimg * P + pix) * 3;
      v[h][0] = a.v_hdr[o]; v[h][1] = a.v_hdr[o + 1]; v[h][2] = a.v_hdr[o + 2];
      const float v_al = a.v_alpha ? a.v_alpha[(int64_t)frame * P + pix] * inv_nv : 0.f;
      // behind the last accumulated Gaussian lies the background: R = bg . v_H - v_alpha (chs_pair_bwd_scalars_r)
      r0[h] = (a.bg[0] * v[h][0] + a.bg[1] * v[h][1] + a.bg[2] * v[h][2]) - v_al;
    }
    warp_last = max(warp_last, last[h]);
  }
  // power-of-two scales of this warp: v_H into [0, 1), and the two magnitudes the dL/dsigma bound is made of
  float vmax = 0.f, r0max = fmaxf(fabsf(r0[0]), fabsf(r0[1]));
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) vmax = fmaxf(vmax, fabsf(v[h][ch]));
  vmax = __uint_as_float(__reduce_max_sync(CHS_FULL_MASK, __float_as_uint(vmax)));
  r0max = __uint_as_float(__reduce_max_sync(CHS_FULL_MASK, __float_as_uint(r0max)));
  const int ev = min(max(expo_of(vmax), -100), 100);
  const float sv = pow2i(-ev);
  const float inv_sc = pow2i(ev - 14);  // undoes sv and the 2^14 carried by alpha T
  {  // colour B operand: this lane's pixel pair is k = (2 t, 2 t + 1) of register q >> 2 in k-step lane >> 3, t = lane & 3
    const int s = lane >> 3, q = lane & 7, t = q & 3;
    uint32_t* vf = reinterpret_cast<uint32_t*>(ws.vfrag) + (q >> 2);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (lane < 8) ws.vfrag[k * 32 + lane] = make_uint2(0u, 0u);  // columns 0 and 1
    __syncwarp();
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const uint32_t hi = f16x2_of(v[0][ch] * sv, v[1][ch] * sv);
      const P2 res = p2(v[0][ch] * sv, v[1][ch] * sv) - f16x2_to_p2(hi);
      vf[2 * (s * 32 + (2 + 2 * ch) * 4 + t)] = hi;
      vf[2 * (s * 32 + (3 + 2 * ch) * 4 + t)] = f16x2_of(p2lo(res), p2hi(res));
    }
  }
  const P2 py2 = p2(iy0 + 0.5f, iy0 + 4.5f);
  P2 Tr2 = p2(Tf[0] * 16384.f, Tf[1] * 16384.f);  // 2^14 x transmittance: alpha T lands in fp16's normal range
  P2 R2 = p2(r0[0], r0[1]);
  const P2 vh_r2 = p2(v[0][0], v[1][0]), vh_g2 = p2(v[0][1], v[1][1]), vh_b2 = p2(v[0][2], v[1][2]);
  __syncthreads();
  warp_last = __reduce_max_sync(CHS_FULL_MASK, warp_last);
  if (lane == 0 && warp_last > 0) atomicMax(&s_max_last, warp_last);
  __syncthreads();
  const int n_walk = s_max_last;

  const unsigned gt = lanemask_gt_();
  const uint32_t tbl0 = (uint32_t)__cvta_generic_to_shared(&ws.tbl[0][0][0]);
  const uint32_t tab0 = tbl0 + 4u * (uint32_t)lane;  // this lane's pixel pair in row 0
  // ldmatrix row address of this lane: matrix i = lane >> 3 = (hi | lo) x (k 0-7 | k 8-15), row = slot lane & 7
  const uint32_t a_addr = tbl0 + (uint32_t)(((lane >> 3) & 1) * 8 * kRow4 * 2 + (lane & 7) * kRow4 * 2 + (lane >> 4) * 16);
  constexpr uint32_t kRowB = kRow4 * 2, kArrB = 8 * kRow4 * 2;
  for (int hi = n_walk; hi > 0; hi -= kB) {
    const int lo = max(0, hi - kB);
    const int cnt = hi - lo;
    __syncthreads();  // every warp is done with the previous batch's records
    {
      float csum = 0.f;
      if (tid < cnt) {
        const int32_t val = a.vals[start + lo + tid] + rec_shift;
        stage_splat3(sm, tid, val, cam_base, a.geom, a.conic_c, a.rgbo);
        const float4 sb = sm.b[tid];
        const float2 sc = *reinterpret_cast<const float2*>(&sm.c[tid]);
        csum = fabsf(sb.w) + fabsf(sc.x) + fabsf(sc.y);
      }
      const unsigned cm = __reduce_max_sync(CHS_FULL_MASK, __float_as_uint(csum));
      if (lane == 0) atomicMax(&s_cbnd, cm);
    }
    __syncthreads();
    if (warp_last <= lo) continue;
    // |nvs| <= 2 max(|R0|, cbnd vmax): scaled into [0, 2^14)
    const float bound = 2.f * fmaxf(r0max, __uint_as_float(s_cbnd) * vmax);
    const int eb = min(max(expo_of(bound), -100), 100);
    const float gate_on = pow2i(-eb);  // = sn / 2^14 with sn = 2^(14 - eb): alpha T already carries the 2^14
    const float inv_sn = pow2i(eb - 14);
    const int lrA = last[0] - lo - 1, lrB = last[1] - lo - 1;  // staged indices <= lr are inside the pixel's accumulated prefix
    // ---- cull the whole batch, back to front ----
    int n_surv = 0;
    for (int sub_hi = cnt; sub_hi > 0; sub_hi -= 32) {
      const int sub_lo = max(0, sub_hi - 32);
      if (warp_last <= lo + sub_lo) continue;
      const int j = sub_lo + lane;
      const bool hit = (j < sub_hi) && (lo + j < warp_last) && splat_hits_block3(sm, j, bx0, bx1, by0, by1);
      const unsigned mask = __ballot_sync(CHS_FULL_MASK, hit);
      if (hit) ws.list[n_surv + __popc(mask & gt)] = (uint8_t)j;  // descending
      n_surv += __popc(mask);
    }
    if (lane < 3) ws.list[n_surv + lane] = (uint8_t)kB;  // padding: the null record (alpha = 0: tables zeros)
    __syncwarp();
    // ---- phase A in groups of four table rows; phase B every eight ----
    for (int i = 0; i < n_surv; i += 4) {
      const uint32_t j4 = *reinterpret_cast<const uint32_t*>(ws.list + i);
      const uint32_t tab = tab0 + (uint32_t)(i & 4) * kRowB;
      // ---- stage 1: what does not depend on T or R, for the four Gaussians (independent streams: blend_bwd5_kernel) ----
      P2 na2[4], ra2[4], s2[4], gate2[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int jj = (int)((j4 >> (8 * u)) & 0xffu);
        const float4 sa = sm.a[jj];  // mx, my, qa, r
        const float4 sb = sm.b[jj];  // kc, log2(opacity), rbc, cr
        const float2 cgb = *reinterpret_cast<const float2*>(&sm.c[jj]);  // cg, cb
        float dx;
        P2 dy2, u2;
        const P2 pw2 = pair_power2(sa, sb.x, sb.y, px, py2, dx, dy2, u2);
        const float pA = p2lo(pw2), pB = p2hi(pw2);
        const bool validA = (jj <= lrA) && pA >= CHS_LOG2_ALPHA_MIN;
        const bool validB = (jj <= lrB) && pB >= CHS_LOG2_ALPHA_MIN;
        const float auA = validA ? chs_exp2_fast(pA) : 0.f;
        const float auB = validB ? chs_exp2_fast(pB) : 0.f;
        na2[u] = p2(fmaxf(-CHS_ALPHA_MAX, -auA), fmaxf(-CHS_ALPHA_MAX, -auB));
        const P2 om2 = p2s(1.f) + na2[u];
        ra2[u] = p2(chs_rcp_fast(p2lo(om2)), chs_rcp_fast(p2hi(om2)));
        s2[u] = fma2(p2s(cgb.y), vh_b2, fma2(p2s(cgb.x), vh_g2, p2s(sb.w) * vh_r2));
        // no gradient through the 0.999 clamp; the gate carries the scale of the dL/dsigma table
        gate2[u] = p2(auA <= CHS_ALPHA_MAX ? gate_on : 0.f, auB <= CHS_ALPHA_MAX ? gate_on : 0.f);
      }
      // ---- stage 2: the recurrences (T carries 2^14), the fp16 hi + lo splits and the table rows ----
      auto pair = [&](auto uc) {
        constexpr int u = decltype(uc)::value;
        Tr2 = Tr2 * ra2[u];  // transmittance before this Gaussian
        const P2 nf2 = na2[u] * Tr2;
        const P2 e2 = R2 - s2[u];
        R2 = fma2(na2[u], e2, R2);
        const P2 nvs2 = (nf2 * e2) * gate2[u];
        const uint32_t vhi = f16x2_of(p2lo(nvs2), p2hi(nvs2));
        const P2 vres = nvs2 - f16x2_to_p2(vhi);
        const uint32_t vlo = f16x2_of(p2lo(vres), p2hi(vres));
        const uint32_t fhi = f16x2_of(p2lo(nf2), p2hi(nf2));
        const P2 fres = nf2 - f16x2_to_p2(fhi);
        const uint32_t flo = f16x2_of(p2lo(fres), p2hi(fres));
        asm volatile("st.shared.b32 [%0+%2], %1;" ::"r"(tab), "r"(vhi), "n"(u * kRowB));
        asm volatile("st.shared.b32 [%0+%2], %1;" ::"r"(tab), "r"(vlo), "n"(u * kRowB + kArrB));
        asm volatile("st.shared.b32 [%0+%2], %1;" ::"r"(tab), "r"(fhi), "n"(u * kRowB + 2 * kArrB));
        asm volatile("st.shared.b32 [%0+%2], %1;" ::"r"(tab), "r"(flo), "n"(u * kRowB + 3 * kArrB));
      };
      pair(std::integral_constant<int, 0>{});
      pair(std::integral_constant<int, 1>{});
      pair(std::integral_constant<int, 2>{});
      pair(std::integral_constant<int, 3>{});
      if (i & 4) bwd_round4<kB>(sm, a_addr, mono, ws.vfrag, ws.list + (i - 4), 8, lane, bx0, by0, inv_sn, inv_sc, a);
    }
    if (((n_surv + 3) >> 2) & 1)  // an odd number of groups: the last one sits alone in rows 0-3
      bwd_round4<kB>(sm, a_addr, mono, ws.vfrag, ws.list + ((n_surv - 1) & ~3), 4, lane, bx0, by0, inv_sn, inv_sc, a);
  }
}

}  // namespace

static size_t crf_smem_bytes(const chs_config* cfg) {
  return (size_t)3 * chs_crf_stride(cfg->crf_kind, cfg->crf_hidden) * sizeof(float);
}

extern "C" int chs_blend_fwd(const chs_config* cfg, const float* geom, const float* conic_c, const float* rgbo,
                             const int32_t* vals_sorted, const uint32_t* tile_offsets, const float* exposure,
                             const float* crf_params, float* ldr, float* alpha, float* hdr_mean, float* final_T, int32_t* last_id,
                             void* stream) {
  ChsDims d;
  int st = chs_make_dims(cfg, &d);
  if (st) return st;
  CHS_REQUIRE(geom && conic_c && rgbo && tile_offsets && exposure, "chs_blend_fwd: null input");
  CHS_REQUIRE(ldr && alpha && hdr_mean && final_T && last_id, "chs_blend_fwd: null output");
  CHS_REQUIRE(cfg->crf_kind == CHS_CRF_IDENTITY || crf_params, "chs_blend_fwd: crf_params required for a learned CRF");
  if (d.B == 0 || d.P == 0) return CHS_OK;
  BlendFwdArgs a;
  a.N = d.N; a.n_virtual = d.n; a.W = d.W; a.H = d.H; a.tile_w = d.tile_w; a.tiles = d.tiles;
  a.crf_kind = cfg->crf_kind; a.crf_hidden = cfg->crf_hidden; a.crf_before_average = cfg->crf_before_average;
  a.rgbo_per_camera = cfg->rgbo_per_camera;
  a.fused = cfg->pose_fused != 0;
  CHS_REQUIRE(!a.fused || (cfg->tune_blend_fwd != 1 && cfg->tune_blend_fwd != 3 && cfg->tune_blend_fwd != 8), "chs_blend_fwd: pose_fused needs the round-2 kernel");
  a.bg[0] = cfg->background[0]; a.bg[1] = cfg->background[1]; a.bg[2] = cfg->background[2];
  a.geom = (const float4*)geom; a.conic_c = conic_c; a.rgbo = (const float4*)rgbo; a.vals = vals_sorted; a.tile_offsets = tile_offsets;
  a.exposure = exposure; a.crf_params = crf_params;
  a.ldr = ldr; a.alpha = alpha; a.hdr_mean = hdr_mean; a.final_T = final_T; a.last_id = last_id;
  dim3 grid(d.tiles, d.B);
  const size_t dyn = crf_smem_bytes(cfg);
  cudaStream_t s = (cudaStream_t)stream;
  if (cfg->crf_before_average) {
    if (cfg->tune_blend_fwd == 8)
      blend_fwd_kernel<8, true><<<grid, kThreads, dyn, s>>>(a);
    else if (cfg->tune_blend_fwd == 27)
      blend_fwd2_kernel<8, true, false><<<grid, kThreads, dyn, s>>>(a);
    else
      blend_fwd2_kernel<7, true, false, true><<<grid, kThreads, dyn, s>>>(a);
  } else {
    switch (cfg->tune_blend_fwd) {  // development knob (chs_config)
      case 1: blend_fwd_kernel<6, false><<<grid, kThreads, dyn, s>>>(a); break;
      case 3: blend_fwd_kernel<10, false><<<grid, kThreads, dyn, s>>>(a); break;
      case 8: blend_fwd_kernel<8, false><<<grid, kThreads, dyn, s>>>(a); break;  // the round-1 kernel (64 registers, 32 warps/SM)
      case 26: blend_fwd2_kernel<6, false, false><<<grid, kThreads, dyn, s>>>(a); break;
      case 28: blend_fwd2_kernel<8, false, false><<<grid, kThreads, dyn, s>>>(a); break;
      case 29: blend_fwd2_kernel<7, false, true><<<grid, kThreads, dyn, s>>>(a); break;  // + cp.async staging
      case 30: blend_fwd2_kernel<6, false, true><<<grid, kThreads, dyn, s>>>(a); break;
      // round 2: survivor list, T -= w.  r2e, c3 (ms per frame): 6 CTAs/SM 2.33 | 7 (72 registers) 2.21 | 8 (64 registers, spills) 2.26;
      // round-1 kernel 2.50
      case 27: blend_fwd2_kernel<7, false, false><<<grid, kThreads, dyn, s>>>(a); break;  // the ungrouped round-2 loop
      case 46: blend_fwd2_kernel<6, false, false, true><<<grid, kThreads, dyn, s>>>(a); break;
      // 128-entry batches (one list entry per thread): r3m 1.997 (7 CTAs per SM) / 1.984 (8) vs 2.007 ms, inside the run-to-run noise
      case 57: blend_fwd2_kernel<7, false, false, true, 128><<<grid, kThreads, dyn, s>>>(a); break;
      case 58: blend_fwd2_kernel<8, false, false, true, 128><<<grid, kThreads, dyn, s>>>(a); break;
      case 48: blend_fwd2_kernel<8, false, false, true><<<grid, kThreads, dyn, s>>>(a); break;
      // grouped pair loop (speculative transmittance chain, one stop vote per four Gaussians)
      default: blend_fwd2_kernel<7, false, false, true><<<grid, kThreads, dyn, s>>>(a); break;
    }
  }
  CHS_LAUNCH_CHECK();
  return CHS_OK;
}

extern "C" int chs_blend_bwd(const chs_config* cfg, const float* geom, const float* conic_c, const float* rgbo,
                             const int32_t* vals_sorted, const uint32_t* tile_offsets, const float* final_T, const int32_t* last_id,
                             const float* v_hdr, const float* v_alpha, float* v_geom, float* v_cogr, float* v_blue, void* stream) {
  ChsDims d;
  int st = chs_make_dims(cfg, &d);
  if (st) return st;
  CHS_REQUIRE(geom && conic_c && rgbo && tile_offsets && final_T && last_id && v_hdr, "chs_blend_bwd: null input");
  CHS_REQUIRE(v_geom && v_cogr && v_blue, "chs_blend_bwd: null output");
  cudaStream_t s = (cudaStream_t)stream;
  CHS_CUDA(cudaMemsetAsync(v_geom, 0, (size_t)d.CN * 16, s));
  CHS_CUDA(cudaMemsetAsync(v_cogr, 0, (size_t)d.CN * 16, s));
  CHS_CUDA(cudaMemsetAsync(v_blue, 0, (size_t)d.CN * 4, s));
  if (d.C == 0 || d.P == 0) return CHS_OK;
  BlendBwdArgs a;
  a.N = d.N; a.n_virtual = d.n; a.W = d.W; a.H = d.H; a.tile_w = d.tile_w; a.tiles = d.tiles;
  a.bg[0] = cfg->background[0]; a.bg[1] = cfg->background[1]; a.bg[2] = cfg->background[2];
  a.geom = (const float4*)geom; a.conic_c = conic_c; a.rgbo = (const float4*)rgbo; a.vals = vals_sorted; a.tile_offsets = tile_offsets;
  a.final_T = final_T; a.last_id = last_id; a.v_hdr = v_hdr; a.v_alpha = v_alpha;
  a.v_hdr_per_camera = cfg->crf_before_average != 0;
  a.rgbo_per_camera = cfg->rgbo_per_camera;
  a.fused = cfg->pose_fused != 0;
  {
    const int tb = cfg->tune_blend_bwd;
    CHS_REQUIRE(!a.fused || tb == 0 || tb == 3 || (tb >= 35 && tb <= 39) || (tb >= 46 && tb <= 48) || (tb >= 56 && tb <= 58) || tb == 64 || tb == 65 || tb == 76 || tb == 77, "chs_blend_bwd: pose_fused needs a round-2 kernel");
  }
  a.v_geom = (float4*)v_geom; a.v_cogr = (float4*)v_cogr; a.v_blue = v_blue;
  dim3 grid(d.tiles, d.C);
#define CHS_BWD2_SMEM(S, B) (sizeof(SplatSmemT<B>) + 4 * sizeof(BwdWarpSmem<S>))
#define CHS_BWD2_LAUNCH(S, B, MB, PIPE) blend_bwd2_kernel<S, B, MB, PIPE><<<grid, kThreads, CHS_BWD2_SMEM(S, B), s>>>(a)
#define CHS_BWD2_ATTR(S, B, MB, PIPE) \
  CHS_CUDA(cudaFuncSetAttribute(blend_bwd2_kernel<S, B, MB, PIPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CHS_BWD2_SMEM(S, B)))
  // r1g sweep on c3 (ms per frame of 8 poses, tight bounds): direct kernel 6.07 | tabled 16 slots / batch 256 / 4 CTAs per SM 5.36 |
  // 16 / 128 / 5: 5.21 | 8 / 128 / 7: 5.01 (default) | 8 / 128 / 8 (64 registers, spills) 5.18 | 8 / 64 / 8: 5.12 | 8 / 256 / 6: 5.51 |
  // 12 / 128 / 6: 5.64 | 10 / 128 / 6: 5.69 (idle lanes in phase B) | software-pipelined phase A (kPipe): +0.5 ms in every configuration
  switch (cfg->tune_blend_bwd) {  // development knob (chs_config)
    case 1: blend_bwd_kernel<1, 8><<<grid, kThreads, 0, s>>>(a); break;          // the direct (r1d-f) kernel, two pixels per thread
    case 22: blend_bwd_kernel<2, 12><<<grid, kThreads / 2, 0, s>>>(a); break;    // direct, four pixels per thread
    case 40: CHS_BWD2_ATTR(16, 256, 4, false); CHS_BWD2_LAUNCH(16, 256, 4, false); break;  // 50 KB of dynamic shared memory: opt in
    case 41: CHS_BWD2_LAUNCH(16, 128, 5, false); break;
    case 45: CHS_BWD2_LAUNCH(8, 128, 7, true); break;
    case 2: CHS_BWD2_LAUNCH(8, 128, 7, false); break;  // the round-1 default: tabled, 8 slots, 128-entry batches, 72 registers
#define CHS_BWD3_SMEM(B) (sizeof(SplatSmem3<B>) + 4 * sizeof(BwdWarp3<8>))
#define CHS_BWD3_SMEM_ASYNC(B) (CHS_BWD3_SMEM(B) + sizeof(RawSmem<B>))
#define CHS_BWD3_LAUNCH(B, MB) blend_bwd3_kernel<8, B, MB, false><<<grid, kThreads, CHS_BWD3_SMEM(B), s>>>(a)
    case 36: CHS_BWD3_LAUNCH(128, 6); break;
    case 38: CHS_BWD3_LAUNCH(128, 8); break;
    case 37: CHS_BWD3_LAUNCH(64, 7); break;
    case 39: blend_bwd3_kernel<8, 128, 7, true><<<grid, kThreads, CHS_BWD3_SMEM_ASYNC(128), s>>>(a); break;  // + cp.async staging
    case 35: blend_bwd3_kernel<8, 128, 6, true><<<grid, kThreads, CHS_BWD3_SMEM_ASYNC(128), s>>>(a); break;
    // round 2: division-free colour state, survivor list, running table pointer.  r2e, c3 (ms per frame of 8 poses): batch 128 /
    // 7 CTAs per SM 4.03 | 128 / 6: 4.19 | 128 / 8 (64 registers): 4.24 | 64 / 7: 4.17; round-1 tabled kernel 4.97
#define CHS_BWD4_SMEM (sizeof(SplatSmem3<129>) + 4 * 32 * sizeof(uint2) + 4 * sizeof(BwdWarp4<128>))
#define CHS_BWD4_LAUNCH(MB)                                                                                                   \
  CHS_CUDA(cudaFuncSetAttribute(blend_bwd4_kernel<128, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CHS_BWD4_SMEM)); \
  blend_bwd4_kernel<128, MB><<<grid, kThreads, CHS_BWD4_SMEM, s>>>(a)
    // phase B on the tensor cores (mma.sync m16n8k16, fp16 hi + lo tables): parity-green, measured r2v on c3: 7 CTAs per SM 4.00 ms |
    // 6: 3.92 | 8: 3.98 against 3.83 for the default below
#define CHS_BWD5_SMEM(SETS) (sizeof(SplatSmem3<129>) + 4 * sizeof(BwdWarp5<128, 8 * SETS>))
#define CHS_BWD5_LAUNCH(MB) blend_bwd5_kernel<128, MB, 1><<<grid, kThreads, CHS_BWD5_SMEM(1), s>>>(a)
#define CHS_BWD5_LAUNCH2(MB)                                                                                                          \
  CHS_CUDA(cudaFuncSetAttribute(blend_bwd5_kernel<128, MB, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CHS_BWD5_SMEM(2))); \
  blend_bwd5_kernel<128, MB, 2><<<grid, kThreads, CHS_BWD5_SMEM(2), s>>>(a)
#define CHS_BWD5_SMEM_B(B, SETS) (sizeof(SplatSmem3<B + 1>) + 4 * sizeof(BwdWarp5<B, 8 * SETS>))
#define CHS_BWD5_LAUNCH_B256(MB) blend_bwd5_kernel<256, MB, 1><<<grid, kThreads, CHS_BWD5_SMEM_B(256, 1), s>>>(a)
    case 76: CHS_BWD5_LAUNCH_B256(6); break;  // 256-entry batches (two list entries per thread), half as many CTA barriers: r3j 3.65-3.69 vs 3.43 ms
    case 77: CHS_BWD5_LAUNCH_B256(7); break;
    case 64: CHS_BWD5_LAUNCH2(4); break;  // 16 table rows, two Gaussians per phase-B lane
    case 65: CHS_BWD5_LAUNCH2(5); break;
    // r3c, c3 (ms per frame of 8 poses): 6 CTAs per SM 3.45 | 7: 3.45 (default) | 8 (64 registers): 3.65; blend_bwd3_kernel 3.87
    case 56: CHS_BWD5_LAUNCH(6); break;
    case 58: CHS_BWD5_LAUNCH(8); break;
    case 3: CHS_BWD3_LAUNCH(128, 7); break;  // the chunked, unstaged phase A (the default until r3c)
    case 46: CHS_BWD4_LAUNCH(6); break;
    case 47: CHS_BWD4_LAUNCH(7); break;
    case 48: CHS_BWD4_LAUNCH(8); break;
    default: CHS_BWD5_LAUNCH(7); break;  // fp32 table, phase A staged for instruction-level parallelism
  }
  CHS_LAUNCH_CHECK();
  return CHS_OK;
}
