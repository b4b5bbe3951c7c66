// chs_ssim.cu — the D-SSIM photometric loss of 3DGS trainers, fused with its gradient
// ("Loss ... L1/SSIM on B_i").
//
//   L = l1_weight * mean|x - y| + ssim_weight * (1 - mean SSIM(x, y))
//
// x = estimated blurred LDR frames, y = captured frames, both [n_img, H, W, 3] fp32; SSIM per channel
// with the usual 11x11 Gaussian window (sigma 1.5), zero padding, C1 = 0.01^2, C2 = 0.03^2; the means
// run over all n_img * H * W * 3 values.
//
// Two streaming kernels, one 16x16-pixel tile (all three channels) per CTA, separable 11-tap window in
// shared memory:
//   ssim_fwd_kernel  x, y (+5-pixel halo) -> window statistics -> SSIM; adds the loss terms to a device
//                    fp64 scalar and stores the three partial-derivative maps dS/dmu1, dS/dE[x^2],
//                    dS/dE[xy] (12 B per value) for the backward;
//   ssim_bwd_kernel  maps (+halo) -> the same window applied to each map ->
//                    dL/dx = -ssim_weight / n * (W*dmu + 2 x W*de11 + y W*de12) + l1_weight / n * sign(x - y).
// Algorithmic traffic: 8 B + 12 B forward, 20 B + 4 B backward per value — 44 B, i.e. ~0.27 GB for a
// 1080p frame: an HBM stream, nothing to reuse beyond the halo.
#include "chs_common.cuh"

namespace {

constexpr int kTile = 16;
constexpr int kHalo = 5;
constexpr int kWin = 2 * kHalo + 1;     // 11
constexpr int kExt = kTile + 2 * kHalo; // 26
constexpr int kThreads = kTile * kTile;
constexpr float kC1 = 0.01f * 0.01f, kC2 = 0.03f * 0.03f;

struct SsimWindow {
  float w[kWin];
};

struct SsimArgs {
  int n_img, H, W;
  float l1_scale, ssim_scale;  // l1_weight / n, ssim_weight / n
  const float *x, *y;
  float* maps;   // [3][n_img, H, W, 3]
  float* v_x;
  double* loss_acc;
  SsimWindow win;
};

// loads the (tile + halo) x 3-channel patch of `src` at the CTA's tile into dst[ch][row][col], zero outside the image
__device__ __forceinline__ void load_patch(const float* __restrict__ src, int img, int H, int W, int y0, int x0,
                                           float (*dst)[kExt][kExt + 1], int tid) {
  const float* base = src + (int64_t)img * H * W * 3;
  for (int i = tid; i < kExt * kExt * 3; i += kThreads) {
    const int row = i / (kExt * 3), rem = i % (kExt * 3);
    const int col = rem / 3, ch = rem % 3;
    const int yy = y0 - kHalo + row, xx = x0 - kHalo + col;
    float v = 0.f;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = base[((int64_t)yy * W + xx) * 3 + ch];
    dst[ch][row][col] = v;
  }
}

__global__ void __launch_bounds__(kThreads) ssim_fwd_kernel(SsimArgs a) {
  __shared__ float s_x[3][kExt][kExt + 1];
  __shared__ float s_y[3][kExt][kExt + 1];
  __shared__ float s_h[5][kExt][kTile + 1];  // horizontal pass of x, y, x^2, y^2, xy for one channel
  __shared__ float s_red[kThreads / 32];
  const int tid = threadIdx.x, tx = tid % kTile, ty = tid / kTile;
  const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile, img = blockIdx.z;
  load_patch(a.x, img, a.H, a.W, y0, x0, s_x, tid);
  load_patch(a.y, img, a.H, a.W, y0, x0, s_y, tid);
  __syncthreads();
  const int px = x0 + tx, py = y0 + ty;
  const bool inside = px < a.W && py < a.H;
  float part = 0.f;  // this thread's contribution to the loss
  float m_mu[3], m_e11[3], m_e12[3];
  for (int ch = 0; ch < 3; ++ch) {
    for (int i = tid; i < kExt * kTile; i += kThreads) {
      const int row = i / kTile, col = i % kTile;
      float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
#pragma unroll
      for (int k = 0; k < kWin; ++k) {
        const float w = a.win.w[k], xv = s_x[ch][row][col + k], yv = s_y[ch][row][col + k];
        sx = fmaf(w, xv, sx);
        sy = fmaf(w, yv, sy);
        sxx = fmaf(w * xv, xv, sxx);
        syy = fmaf(w * yv, yv, syy);
        sxy = fmaf(w * xv, yv, sxy);
      }
      s_h[0][row][col] = sx; s_h[1][row][col] = sy; s_h[2][row][col] = sxx; s_h[3][row][col] = syy; s_h[4][row][col] = sxy;
    }
    __syncthreads();
    float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
    for (int k = 0; k < kWin; ++k) {
      const float w = a.win.w[k];
      mu1 = fmaf(w, s_h[0][ty + k][tx], mu1);
      mu2 = fmaf(w, s_h[1][ty + k][tx], mu2);
      e11 = fmaf(w, s_h[2][ty + k][tx], e11);
      e22 = fmaf(w, s_h[3][ty + k][tx], e22);
      e12 = fmaf(w, s_h[4][ty + k][tx], e12);
    }
    __syncthreads();  // s_h is rewritten for the next channel
    const float s11 = e11 - mu1 * mu1, s22 = e22 - mu2 * mu2, s12 = e12 - mu1 * mu2;
    const float A1 = 2.f * mu1 * mu2 + kC1, A2 = 2.f * s12 + kC2, B1 = mu1 * mu1 + mu2 * mu2 + kC1, B2 = s11 + s22 + kC2;
    const float iB1 = 1.f / B1, iB2 = 1.f / B2;
    const float S = A1 * A2 * iB1 * iB2;
    // partial derivatives of S w.r.t. the x-side window statistics (mu1 total, through s11 and s12 as well)
    const float d_e11 = -S * iB2;
    const float d_e12 = 2.f * A1 * iB1 * iB2;
    const float d_mu = 2.f * mu2 * A2 * iB1 * iB2 - 2.f * mu1 * S * iB1 - 2.f * mu1 * d_e11 - mu2 * d_e12;
    m_mu[ch] = inside ? d_mu : 0.f;
    m_e11[ch] = inside ? d_e11 : 0.f;
    m_e12[ch] = inside ? d_e12 : 0.f;
    if (inside) {
      const float d = s_x[ch][ty + kHalo][tx + kHalo] - s_y[ch][ty + kHalo][tx + kHalo];
      part += a.l1_scale * fabsf(d) + a.ssim_scale * (1.f - S);
    }
  }
  if (inside) {
    const int64_t plane = (int64_t)a.n_img * a.H * a.W * 3;
    const int64_t o = (((int64_t)img * a.H + py) * a.W + px) * 3;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      a.maps[o + ch] = m_mu[ch];
      a.maps[plane + o + ch] = m_e11[ch];
      a.maps[2 * plane + o + ch] = m_e12[ch];
    }
  }
  part = chs_warp_sum(part);
  if ((tid & 31) == 0) s_red[tid >> 5] = part;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) s += s_red[w];
    atomicAdd(a.loss_acc, (double)s);
  }
}

__global__ void __launch_bounds__(kThreads) ssim_bwd_kernel(SsimArgs a) {
  __shared__ float s_m[3][kExt][kExt + 1];   // one derivative map, three channels, with halo
  __shared__ float s_h[3][kExt][kTile + 1];  // its horizontal pass
  const int tid = threadIdx.x, tx = tid % kTile, ty = tid / kTile;
  const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile, img = blockIdx.z;
  const int px = x0 + tx, py = y0 + ty;
  const bool inside = px < a.W && py < a.H;
  const int64_t plane = (int64_t)a.n_img * a.H * a.W * 3;
  const int64_t o = (((int64_t)img * a.H + py) * a.W + px) * 3;
  float acc[3] = {0.f, 0.f, 0.f};
  float xv[3] = {0.f, 0.f, 0.f}, yv[3] = {0.f, 0.f, 0.f};
  if (inside) {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      xv[ch] = a.x[o + ch];
      yv[ch] = a.y[o + ch];
    }
  }
  for (int m = 0; m < 3; ++m) {  // dS/dmu1, dS/dE[x^2], dS/dE[xy]
    __syncthreads();
    load_patch(a.maps + m * plane, img, a.H, a.W, y0, x0, s_m, tid);
    __syncthreads();
    for (int i = tid; i < 3 * kExt * kTile; i += kThreads) {
      const int ch = i / (kExt * kTile), rem = i % (kExt * kTile);
      const int row = rem / kTile, col = rem % kTile;
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < kWin; ++k) s = fmaf(a.win.w[k], s_m[ch][row][col + k], s);
      s_h[ch][row][col] = s;
    }
    __syncthreads();
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < kWin; ++k) s = fmaf(a.win.w[k], s_h[ch][ty + k][tx], s);
      acc[ch] += m == 0 ? s : (m == 1 ? 2.f * xv[ch] * s : yv[ch] * s);
    }
  }
  if (inside) {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const float d = xv[ch] - yv[ch];
      const float l1 = d > 0.f ? a.l1_scale : (d < 0.f ? -a.l1_scale : 0.f);
      a.v_x[o + ch] = l1 - a.ssim_scale * acc[ch];
    }
  }
}

}  // namespace

extern "C" int chs_ssim_loss(const float* ldr, const float* target, int32_t n_img, int32_t height, int32_t width, float l1_weight,
                             float ssim_weight, float* v_ldr, double* loss_acc, void* workspace, uint64_t workspace_bytes,
                             void* stream) {
  CHS_REQUIRE(n_img >= 0 && height >= 1 && width >= 1, "chs_ssim_loss: bad shape");
  if (n_img == 0) return CHS_OK;
  CHS_REQUIRE(ldr && target && v_ldr && loss_acc && workspace, "chs_ssim_loss: null pointer");
  CHS_REQUIRE(n_img <= 65535 && (height + kTile - 1) / kTile <= 65535, "chs_ssim_loss: too many images / rows for one launch");
  const uint64_t count = (uint64_t)n_img * height * width * 3;
  const uint64_t need = 3 * count * sizeof(float);
  if (workspace_bytes < need) {
    chs_set_error("chs_ssim_loss: workspace too small (%llu < %llu)", (unsigned long long)workspace_bytes, (unsigned long long)need);
    return CHS_ERR_WORKSPACE_TOO_SMALL;
  }
  SsimArgs a;
  a.n_img = n_img; a.H = height; a.W = width;
  a.l1_scale = (float)((double)l1_weight / (double)count);
  a.ssim_scale = (float)((double)ssim_weight / (double)count);
  a.x = ldr; a.y = target; a.maps = (float*)workspace; a.v_x = v_ldr; a.loss_acc = loss_acc;
  double w[kWin], sum = 0.0;
  for (int k = 0; k < kWin; ++k) {
    w[k] = exp(-(double)((k - kHalo) * (k - kHalo)) / (2.0 * 1.5 * 1.5));
    sum += w[k];
  }
  for (int k = 0; k < kWin; ++k) a.win.w[k] = (float)(w[k] / sum);
  dim3 grid((width + kTile - 1) / kTile, (height + kTile - 1) / kTile, n_img);
  cudaStream_t s = (cudaStream_t)stream;
  ssim_fwd_kernel<<<grid, kThreads, 0, s>>>(a);
  CHS_LAUNCH_CHECK();
  ssim_bwd_kernel<<<grid, kThreads, 0, s>>>(a);
  CHS_LAUNCH_CHECK();
  return CHS_OK;
}
