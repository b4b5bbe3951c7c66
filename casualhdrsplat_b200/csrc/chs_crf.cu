// chs_crf.cu — K7 crf_bwd: gradient of the formation epilogue B = F_theta(dt * mean_k H_k)
// (SURVEY.md Appendix A.7).  Streaming per-pixel kernel; the CRF parameter gradients are reduced
// warp -> shared memory -> fp64 global accumulators.
#include "chs_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kPix = 4;  // pixels per thread

struct CrfBwdArgs {
  int64_t P;
  int n_virtual, crf_kind, hd;
  int imgs_per_frame;  // 1: one HDR image per frame (default order); n: one per virtual pose (crf_before_average)
  float vy_scale;      // gradient reaching each image's CRF output: 1 (default) or 1/n
  float vh_div;        // v_hdr = dt * v_X / vh_div: n (default: mean over poses follows) or 1
  const float *hdr_mean, *exposure, *crf_params, *v_ldr;
  float* v_hdr;
  double* acc_crf;       // [3 * (3 hd + 1)]
  double* acc_exposure;  // [B]
};

// Phase 1 (pixel-parallel): each thread recomputes the CRF of kPix pixels, writes v_hdr, and leaves
// z = ln(X + eps) and gy = v_y * y * (1 - y) of every (pixel, channel) in shared memory.
// Phase 2 (hidden-unit-parallel): lane <-> hidden unit.  A warp walks its 128 pixels reading (z, gy)
// by shared-memory broadcast; every lane accumulates the three parameter gradients of its own
// units in registers, so no cross-lane reduction is needed at all.
constexpr int kMaxUnitsPerLane = 4;  // crf_hidden <= 128

template <int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks) crf_bwd_kernel(CrfBwdArgs a) {
  extern __shared__ float smem[];
  const int stride = chs_crf_stride(a.crf_kind, a.hd);
  float* s_p = smem;                         // parameters [3, stride]
  float* s_g = s_p + 3 * stride;             // block-partial parameter gradients [3, stride]
  float* s_z = s_g + 3 * stride;             // [3][kThreads * kPix]
  float* s_gy = s_z + 3 * kThreads * kPix;   // [3][kThreads * kPix]
  const int img = blockIdx.y;
  const int frame = img / a.imgs_per_frame;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool mlp = a.crf_kind == CHS_CRF_MLP, lut = a.crf_kind == CHS_CRF_LUT;
  for (int i = tid; i < 3 * stride; i += kThreads) {  // stride = 0 for the identity CRF
    s_p[i] = a.crf_params[i];
    s_g[i] = 0.f;
  }
  __syncthreads();
  const float dt = a.exposure[frame];
  const float scale = dt / a.vh_div;
  float v_dt = 0.f;
  // persistent over pixel chunks: the block-partial parameter gradients stay in shared memory across chunks, so the
  // contended fp64 atomics happen once per block, not once per 1024 pixels
  const int64_t n_chunks = (a.P + (int64_t)kPix * kThreads - 1) / ((int64_t)kPix * kThreads);
  for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
  int64_t o_img[kPix], o_frm[kPix];
  bool live[kPix];
#pragma unroll
  for (int q = 0; q < kPix; ++q) {
    const int64_t pix = chunk * (kPix * kThreads) + q * kThreads + tid;
    live[q] = pix < a.P;
    o_img[q] = ((int64_t)img * a.P + (live[q] ? pix : 0)) * 3;
    o_frm[q] = ((int64_t)frame * a.P + (live[q] ? pix : 0)) * 3;
  }
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    float h[kPix], vy[kPix], vx[kPix];
#pragma unroll
    for (int q = 0; q < kPix; ++q) {
      h[q] = live[q] ? a.hdr_mean[o_img[q] + ch] : 0.f;
      vy[q] = live[q] ? a.v_ldr[o_frm[q] + ch] * a.vy_scale : 0.f;
    }
    if (mlp) {
      // the unit's three weights are fetched once for the thread's kPix pixels
      const float* p = s_p + ch * stride;
      float xe[kPix], zz[kPix], acc[kPix], dz[kPix];
#pragma unroll
      for (int q = 0; q < kPix; ++q) {
        xe[q] = fmaxf(dt * h[q], 0.f) + CHS_CRF_EPS;  // X clamped to >= 0 (chs_crf_mlp_fwd)
        zz[q] = logf(xe[q]);
        acc[q] = p[3 * a.hd];
        dz[q] = 0.f;
      }
      for (int j = 0; j < a.hd; ++j) {
        const float w1 = p[j], b1 = p[a.hd + j], w2 = p[2 * a.hd + j];
        const float w21 = w2 * w1;
#pragma unroll
        for (int q = 0; q < kPix; ++q) {
          const float pre = fmaf(w1, zz[q], b1);
          if (pre > 0.f) {
            acc[q] = fmaf(w2, pre, acc[q]);
            dz[q] += w21;
          }
        }
      }
#pragma unroll
      for (int q = 0; q < kPix; ++q) {
        const float y = 1.f / (1.f + expf(-acc[q]));
        const float g = vy[q] * y * (1.f - y);  // 0 for pixels past the end: they add nothing below
        vx[q] = dt * h[q] >= 0.f ? g * dz[q] / xe[q] : 0.f;  // zero gradient below the clamp
        s_z[ch * (kThreads * kPix) + q * kThreads + tid] = zz[q];
        s_gy[ch * (kThreads * kPix) + q * kThreads + tid] = g;
      }
    } else if (lut) {
      // piecewise-linear table: the two knot gradients go straight to the block's shared accumulators
      const float* p = s_p + ch * stride;
#pragma unroll
      for (int q = 0; q < kPix; ++q) {
        vx[q] = 0.f;
        if (live[q]) {
          const ChsLutPos<float> pos = chs_crf_lut_pos(dt * h[q], p, a.hd);
          atomicAdd(&s_g[ch * stride + 2 + pos.i], vy[q] * (1.f - pos.f));
          atomicAdd(&s_g[ch * stride + 3 + pos.i], vy[q] * pos.f);
          vx[q] = vy[q] * (p[3 + pos.i] - p[2 + pos.i]) * pos.du_dz / pos.xe;
        }
      }
    } else {
#pragma unroll
      for (int q = 0; q < kPix; ++q) vx[q] = vy[q];
    }
#pragma unroll
    for (int q = 0; q < kPix; ++q) {
      if (live[q]) {
        v_dt = fmaf(vx[q], h[q], v_dt);
        a.v_hdr[o_img[q] + ch] = vx[q] * scale;
      }
    }
  }
  __syncthreads();
  if (mlp) {
    const int upl = (a.hd + 31) / 32;  // units per lane
    const int p0 = warp * (kThreads * kPix / 8), p1 = p0 + kThreads * kPix / 8;
    for (int ch = 0; ch < 3; ++ch) {
      const float* p = s_p + ch * stride;
      float w1[kMaxUnitsPerLane], b1[kMaxUnitsPerLane], w2[kMaxUnitsPerLane];
      float g_w1[kMaxUnitsPerLane], g_b1[kMaxUnitsPerLane], g_w2[kMaxUnitsPerLane];
#pragma unroll
      for (int u = 0; u < kMaxUnitsPerLane; ++u) {
        const int j = u * 32 + lane;
        const bool on = u < upl && j < a.hd;
        w1[u] = on ? p[j] : 0.f;
        b1[u] = on ? p[a.hd + j] : -1.f;  // pre = -1 < 0: inactive unit contributes nothing
        w2[u] = on ? p[2 * a.hd + j] : 0.f;
        g_w1[u] = g_b1[u] = g_w2[u] = 0.f;
      }
      float g_b2 = 0.f;
      const float* zc = s_z + ch * (kThreads * kPix);
      const float* gc = s_gy + ch * (kThreads * kPix);
      for (int q = p0; q < p1; ++q) {
        const float zz = zc[q], g = gc[q];
        g_b2 += g;
#pragma unroll
        for (int u = 0; u < kMaxUnitsPerLane; ++u) {
          if (u < upl) {
            const float pre = fmaf(w1[u], zz, b1[u]);
            const float gate = pre > 0.f ? g : 0.f;
            g_w2[u] = fmaf(gate, pre, g_w2[u]);
            const float dh = gate * w2[u];
            g_w1[u] = fmaf(dh, zz, g_w1[u]);
            g_b1[u] += dh;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kMaxUnitsPerLane; ++u) {
        const int j = u * 32 + lane;
        if (u < upl && j < a.hd) {
          atomicAdd(&s_g[ch * stride + j], g_w1[u]);
          atomicAdd(&s_g[ch * stride + a.hd + j], g_b1[u]);
          atomicAdd(&s_g[ch * stride + 2 * a.hd + j], g_w2[u]);
        }
      }
      if (lane == 0) atomicAdd(&s_g[ch * stride + 3 * a.hd], g_b2);
    }
  }
  __syncthreads();  // s_z / s_gy are rewritten by the next chunk
  }
  // exposure (brightness path): block reduce -> one fp64 atomic
  v_dt = chs_warp_sum(v_dt);
  __shared__ float s_dt[kThreads / 32];
  if (lane == 0) s_dt[warp] = v_dt;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) s += s_dt[w];
    atomicAdd(&a.acc_exposure[frame], (double)s);
  }
  for (int i = tid; i < 3 * stride; i += kThreads)
    if (s_g[i] != 0.f) atomicAdd(&a.acc_crf[i], (double)s_g[i]);
}

// ---------------------------------------------------------------------------------------------
// MLP CRF, interval form (default since r3b).  With a scalar input z the network  sum_j w2_j relu(w1_j z + b1_j) + b2  is
// piecewise linear in z: unit j switches at t_j = -b1_j / w1_j.  Per block the Hd breakpoints of a channel are ranked once
// (Hd^2 compares), and for each of the Hd + 1 intervals the slope A_I = sum_{j on} w2_j w1_j and offset B_I = b2 + sum_{j on}
// w2_j b1_j are tabulated.  A pixel then costs a binary search and one fma instead of Hd units (phase 1 of crf_bwd_kernel: 4 Hd
// instructions per pixel and channel), and the parameter gradients need only two sums per INTERVAL,
//     H_I = sum_{pix in I} g,   HZ_I = sum_{pix in I} g z          (g = v_y y (1 - y)),
// because "unit j is on at this pixel" is a statement about the pixel's interval:  S_j = sum_{I: j on} H_I,  SZ_j likewise, and
//     dL/dw2_j = w1_j SZ_j + b1_j S_j,   dL/dw1_j = w2_j SZ_j,   dL/db1_j = w2_j S_j,   dL/db2 = sum_I H_I.
// Neighbouring pixels fall into the same one to three intervals, so a warp adds its 32 values per distinct interval with one
// shuffle reduction and a plain store into a warp-private table (no atomics, no per-unit loop: phase 2 of crf_bwd_kernel was
// 15 instructions per pixel and channel).  A unit whose pre-activation is within rounding of zero may be classified differently
// from the forward's `pre > 0`; its contribution to the value is continuous there and the affected pixels have measure zero.
// ---------------------------------------------------------------------------------------------
template <int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks) crf_bwd_interval_kernel(CrfBwdArgs a) {
  extern __shared__ float smem[];
  const int hd = a.hd, stride = 3 * hd + 1, nI = hd + 1;
  float* s_p = smem;                      // parameters [3, stride]
  float* s_g = s_p + 3 * stride;          // block-partial parameter gradients [3, stride]
  float* s_bp = s_g + 3 * stride;         // sorted breakpoints [3][hd]
  float* s_A = s_bp + 3 * hd;             // [3][nI] slope of the interval
  float* s_B = s_A + 3 * nI;              // [3][nI] offset
  int* s_key = reinterpret_cast<int*>(s_B + 3 * nI);  // [3][hd] unit j is on in interval I  <=>  sgn_j * I + off_j >= 0 (key = off_j * 2 + (sgn_j < 0))
  float* s_H = reinterpret_cast<float*>(s_key + 3 * hd);  // [warps][3][nI][2] warp-private interval sums (g, g z)
  const int img = blockIdx.y;
  const int frame = img / a.imgs_per_frame;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 3 * stride; i += kThreads) {
    s_p[i] = a.crf_params[i];
    s_g[i] = 0.f;
  }
  for (int i = tid; i < (kThreads / 32) * 3 * nI * 2; i += kThreads) s_H[i] = 0.f;
  __syncthreads();
  // rank of every unit's breakpoint inside its channel (ties by unit index); units with w1 = 0 never switch: breakpoint +inf.
  // The breakpoints are computed once (s_A is free until the ranks are known) so that the Hd^2 ranking loop only compares.
  float* s_t = s_A;
  for (int i = tid; i < 3 * hd; i += kThreads) {
    const int ch = i / hd, j = i - ch * hd;
    const float* p = s_p + ch * stride;
    s_t[i] = chs_crf_breakpoint(p[j], p[hd + j]);
  }
  __syncthreads();
  for (int i = tid; i < 3 * hd; i += kThreads) {
    const int ch = i / hd, j = i - ch * hd;
    const float* p = s_p + ch * stride;
    const float* tc = s_t + ch * hd;
    const float t = tc[j];
    int r = 0;
    for (int k = 0; k < hd; ++k) {
      const float tk = tc[k];
      r += (tk < t || (tk == t && k < j)) ? 1 : 0;
    }
    s_bp[ch * hd + r] = t;
    s_key[i] = chs_crf_unit_key(p[j], p[hd + j], r, hd);
  }
  __syncthreads();
  for (int i = tid; i < 3 * nI; i += kThreads) {
    const int ch = i / nI, I = i - ch * nI;
    const float* p = s_p + ch * stride;
    float A = 0.f, B = p[3 * hd];
    for (int j = 0; j < hd; ++j) {
      if (chs_crf_key_on(s_key[ch * hd + j], I)) {
        A = fmaf(p[2 * hd + j], p[j], A);
        B = fmaf(p[2 * hd + j], p[hd + j], B);
      }
    }
    s_A[i] = A;
    s_B[i] = B;
  }
  __syncthreads();
  const float dt = a.exposure[frame];
  const float scale = dt / a.vh_div;
  float v_dt = 0.f;
  float* myH = s_H + warp * (3 * nI * 2);
  const int64_t n_chunks = (a.P + (int64_t)kPix * kThreads - 1) / ((int64_t)kPix * kThreads);
  for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
#pragma unroll
    for (int q = 0; q < kPix; ++q) {
      const int64_t pix = chunk * (kPix * kThreads) + q * kThreads + tid;
      const bool live = pix < a.P;
      const int64_t o_img = ((int64_t)img * a.P + (live ? pix : 0)) * 3, o_frm = ((int64_t)frame * a.P + (live ? pix : 0)) * 3;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const float h = live ? a.hdr_mean[o_img + ch] : 0.f;
        const float vy = live ? a.v_ldr[o_frm + ch] * a.vy_scale : 0.f;
        const float xe = fmaxf(dt * h, 0.f) + CHS_CRF_EPS;  // X clamped to >= 0 (chs_crf_mlp_fwd)
        const float z = logf(xe);
        const int I = chs_crf_interval_of(s_bp + ch * hd, hd, z);  // number of breakpoints below z
        const float A = s_A[ch * nI + I];
        const float acc = fmaf(A, z, s_B[ch * nI + I]);
        const float y = 1.f / (1.f + expf(-acc));
        const float g = vy * y * (1.f - y);  // 0 for pixels past the end
        const float vx = dt * h >= 0.f ? g * A / xe : 0.f;  // zero gradient below the clamp
        if (live) {
          v_dt = fmaf(vx, h, v_dt);
          a.v_hdr[o_img + ch] = vx * scale;
        }
        // the warp's 32 (I, g, g z) triples, one shuffle reduction per distinct interval
        const float gz = g * z;
        unsigned todo = __ballot_sync(CHS_FULL_MASK, g != 0.f);
        while (todo) {
          const int leader = __ffs(todo) - 1;
          const int Ib = __shfl_sync(CHS_FULL_MASK, I, leader);
          const bool mine = (I == Ib) && g != 0.f;
          const float sg = chs_warp_sum(mine ? g : 0.f), sgz = chs_warp_sum(mine ? gz : 0.f);
          if (lane == leader) {
            float* hrow = myH + (ch * nI + Ib) * 2;
            hrow[0] += sg;
            hrow[1] += sgz;
          }
          todo &= ~__ballot_sync(CHS_FULL_MASK, mine);
          __syncwarp();  // the next leader may be another lane adding to the same interval
        }
      }
    }
  }
  __syncthreads();
  // interval sums of the block, then the per-unit sums over the intervals where the unit is on
  for (int i = tid; i < 3 * nI * 2; i += kThreads) {
    float t = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) t += s_H[w * (3 * nI * 2) + i];
    s_H[i] = t;  // (warp 0's table becomes the block's: every thread rewrites only the entry it summed)
  }
  __syncthreads();
  for (int i = tid; i < 3 * hd; i += kThreads) {
    const int ch = i / hd, j = i - ch * hd;
    const float* p = s_p + ch * stride;
    const int key = s_key[i];
    float S = 0.f, SZ = 0.f;
    for (int I = 0; I < nI; ++I) {
      if (chs_crf_key_on(key, I)) {
        S += s_H[(ch * nI + I) * 2];
        SZ += s_H[(ch * nI + I) * 2 + 1];
      }
    }
    s_g[ch * stride + j] = p[2 * hd + j] * SZ;                        // dL/dw1
    s_g[ch * stride + hd + j] = p[2 * hd + j] * S;                    // dL/db1
    s_g[ch * stride + 2 * hd + j] = fmaf(p[j], SZ, p[hd + j] * S);    // dL/dw2
  }
  if (tid < 3) {
    float t = 0.f;
    for (int I = 0; I < nI; ++I) t += s_H[(tid * nI + I) * 2];
    s_g[tid * stride + 3 * hd] = t;  // dL/db2
  }
  __syncthreads();
  v_dt = chs_warp_sum(v_dt);
  __shared__ float s_dt[kThreads / 32];
  if (lane == 0) s_dt[warp] = v_dt;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) t += s_dt[w];
    atomicAdd(&a.acc_exposure[frame], (double)t);
  }
  for (int i = tid; i < 3 * stride; i += kThreads)
    if (s_g[i] != 0.f) atomicAdd(&a.acc_crf[i], (double)s_g[i]);
}

__global__ void finalize_f64_to_f32(const double* src, float* dst, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (float)src[i];
}

}  // namespace

extern "C" int chs_crf_bwd(const chs_config* cfg, const float* hdr_mean, const float* exposure, const float* crf_params,
                           const float* v_ldr, float* v_hdr, float* v_crf_params, float* v_exposure, void* workspace,
                           uint64_t workspace_bytes, void* stream) {
  ChsDims d;
  int st = chs_make_dims(cfg, &d);
  if (st) return st;
  CHS_REQUIRE(hdr_mean && exposure && v_ldr && v_hdr && v_exposure && workspace, "chs_crf_bwd: null pointer");
  const bool mlp = cfg->crf_kind == CHS_CRF_MLP;
  const bool learned = cfg->crf_kind != CHS_CRF_IDENTITY;
  CHS_REQUIRE(!learned || (crf_params && v_crf_params), "chs_crf_bwd: crf_params / v_crf_params required for a learned CRF");
  const int n_par = 3 * chs_crf_stride(cfg->crf_kind, cfg->crf_hidden);
  const uint64_t need = (uint64_t)(n_par + d.B) * sizeof(double);
  if (workspace_bytes < need) {
    chs_set_error("chs_crf_bwd: workspace too small (%llu < %llu)", (unsigned long long)workspace_bytes, (unsigned long long)need);
    return CHS_ERR_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t s = (cudaStream_t)stream;
  double* acc = (double*)workspace;
  CHS_CUDA(cudaMemsetAsync(acc, 0, need, s));
  if (d.B > 0 && d.P > 0) {
    CrfBwdArgs a;
    const bool per_pose = cfg->crf_before_average != 0;
    a.P = d.P; a.n_virtual = d.n; a.crf_kind = cfg->crf_kind; a.hd = learned ? cfg->crf_hidden : 0;
    a.imgs_per_frame = per_pose ? d.n : 1;
    a.vy_scale = per_pose ? 1.f / (float)d.n : 1.f;
    a.vh_div = per_pose ? 1.f : (float)d.n;
    a.hdr_mean = hdr_mean; a.exposure = exposure; a.crf_params = crf_params; a.v_ldr = v_ldr; a.v_hdr = v_hdr;
    a.acc_crf = acc; a.acc_exposure = acc + n_par;
    const int64_t n_chunks = (d.P + (int64_t)kThreads * kPix - 1) / ((int64_t)kThreads * kPix);
    const int n_img = per_pose ? d.C : d.B;
    const int mb = cfg->tune_crf_bwd ? cfg->tune_crf_bwd % 10 : 4;  // development knob (chs_config): resident blocks per SM  // r1g, c3: 2 -> 0.306 ms, 3 -> 0.296, 4 -> 0.280 (previous phase-1 loop order: 0.317)
    int gx = (148 * mb + n_img - 1) / n_img;  // one wave of resident blocks over all images
    if (gx > n_chunks) gx = (int)n_chunks;
    dim3 grid((unsigned)gx, n_img);
    size_t smem = mlp ? ((size_t)2 * n_par + (size_t)6 * kThreads * kPix) * sizeof(float) : (size_t)2 * n_par * sizeof(float) + 16;
    if (mlp && cfg->tune_crf_bwd < 10) {  // interval form; tune_crf_bwd = 12 / 13 / 14: the per-unit kernel at 2 / 3 / 4 blocks per SM
      const int hd = cfg->crf_hidden, nI = hd + 1;
      const size_t smem_i = ((size_t)2 * n_par + 3 * hd + 6 * nI + 3 * hd + (size_t)(kThreads / 32) * 3 * nI * 2) * sizeof(float);
      if (mb == 2)
        crf_bwd_interval_kernel<2><<<grid, kThreads, smem_i, s>>>(a);
      else if (mb == 3)
        crf_bwd_interval_kernel<3><<<grid, kThreads, smem_i, s>>>(a);
      else
        crf_bwd_interval_kernel<4><<<grid, kThreads, smem_i, s>>>(a);
    } else if (mb == 4)
      crf_bwd_kernel<4><<<grid, kThreads, smem, s>>>(a);
    else if (mb == 2)
      crf_bwd_kernel<2><<<grid, kThreads, smem, s>>>(a);
    else
      crf_bwd_kernel<3><<<grid, kThreads, smem, s>>>(a);
    CHS_LAUNCH_CHECK();
  }
  if (n_par > 0) {
    finalize_f64_to_f32<<<(n_par + 255) / 256, 256, 0, s>>>(acc, v_crf_params, n_par);
    CHS_LAUNCH_CHECK();
  }
  if (d.B > 0) {
    finalize_f64_to_f32<<<(d.B + 255) / 256, 256, 0, s>>>(acc + n_par, v_exposure, d.B);
    CHS_LAUNCH_CHECK();
  }
  return CHS_OK;
}
