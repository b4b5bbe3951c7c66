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
// shared memory, register blocked (4 outputs per work item in both passes, 128-bit shared-memory loads),
// outputs staged through shared memory so that the channel-interleaved stores are coalesced:
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

constexpr int kPW = 28;  // padded patch row: 26 used columns, rows stay 16-byte aligned for 128-bit loads

// loads the (tile + halo) x 3-channel patch of `src` at the CTA's tile into dst[ch][row][col], zero outside the image.
// Three patch rows per sweep: a thread decodes its (row offset, column, channel) once, the loop only adds.
__device__ __forceinline__ void load_patch(const float* __restrict__ src, int img, int H, int W, int y0, int x0,
                                           float (*dst)[kExt][kPW], int tid) {
  constexpr int kRowF = kExt * 3;  // 78 consecutive floats per patch row
  if (tid >= 3 * kRowF) return;    // 234 of the 256 threads carry the loads
  const int r_off = tid / kRowF, cc = tid % kRowF;
  const int col = cc / 3, ch = cc % 3;
  const int xx = x0 - kHalo + col;
  const bool x_in = xx >= 0 && xx < W;
  const float* base = src + (int64_t)img * H * W * 3 + (int64_t)xx * 3 + ch;
#pragma unroll
  for (int row = r_off; row < kExt; row += 3) {
    const int yy = y0 - kHalo + row;
    float v = 0.f;
    if (x_in && yy >= 0 && yy < H) v = base[(int64_t)yy * W * 3];
    dst[ch][row][col] = v;
  }
}

// the tile's own 256 x 3 values in [pixel][channel] order are covered by three sweeps of the CTA: element tid + 256 s
struct TileElems {
  int64_t off[3];  // offset of the element in a [n_img, H, W, 3] array
  bool in[3];
};
__device__ __forceinline__ TileElems tile_elems(int img, int H, int W, int y0, int x0, int tid) {
  TileElems e;
#pragma unroll
  for (int sw = 0; sw < 3; ++sw) {
    const int i = tid + kThreads * sw;
    const int p = i / 3, c = i % 3;
    const int py = y0 + (p >> 4), px = x0 + (p & 15);
    e.in[sw] = px < W && py < H;
    e.off[sw] = (((int64_t)img * H + py) * W + px) * 3 + c;
  }
  return e;
}

// 16 consecutive patch values starting at a 16-byte aligned column
__device__ __forceinline__ void load16(const float* p, float v[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = reinterpret_cast<const float4*>(p)[i];
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}

// Both kernels are register blocked: a horizontal work item is (channel, patch row, group of 4 output columns) — 14 inputs
// fetched with four 128-bit loads feed 4 x 11 taps — and a vertical one is (channel, group of 4 output rows, column).
constexpr int kHItems = 3 * kExt * (kTile / 4);  // 312
constexpr int kVItems = 3 * (kTile / 4) * kTile; // 192 threads carry 4 outputs each

__global__ void __launch_bounds__(kThreads, 4) ssim_fwd_kernel(SsimArgs a) {
  __shared__ __align__(16) float s_xy[2][3][kExt][kPW];    // x and y patches; reused as the output staging area
  __shared__ __align__(16) float s_h[5][3][kExt][kTile];   // horizontal pass of x, y, x^2, y^2, xy
  __shared__ float s_red[kThreads / 32];
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile, img = blockIdx.z;
  load_patch(a.x, img, a.H, a.W, y0, x0, s_xy[0], tid);
  load_patch(a.y, img, a.H, a.W, y0, x0, s_xy[1], tid);
  __syncthreads();
  for (int item = tid; item < kHItems; item += kThreads) {
    const int ch = item / (kExt * 4), rem = item % (kExt * 4);
    const int row = rem / 4, c0 = (rem % 4) * 4;
    float xv[16], yv[16];
    load16(&s_xy[0][ch][row][c0], xv);
    load16(&s_xy[1][ch][row][c0], yv);
    float o[5][4];
#pragma unroll
    for (int q = 0; q < 5; ++q)
#pragma unroll
      for (int j = 0; j < 4; ++j) o[q][j] = 0.f;
#pragma unroll
    for (int i = 0; i < 14; ++i) {  // input i feeds output j through tap k = i - j
      const float xx = xv[i] * xv[i], yy = yv[i] * yv[i], xy = xv[i] * yv[i];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = i - j;
        if (k >= 0 && k < kWin) {
          const float w = a.win.w[k];
          o[0][j] = fmaf(w, xv[i], o[0][j]);
          o[1][j] = fmaf(w, yv[i], o[1][j]);
          o[2][j] = fmaf(w, xx, o[2][j]);
          o[3][j] = fmaf(w, yy, o[3][j]);
          o[4][j] = fmaf(w, xy, o[4][j]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) *reinterpret_cast<float4*>(&s_h[q][ch][row][c0]) = make_float4(o[q][0], o[q][1], o[q][2], o[q][3]);
  }
  __syncthreads();
  float part = 0.f;  // this thread's contribution to the loss
  float m_out[3][4];
  const int tx = tid % kTile, grp = tid / kTile;  // grp = ch * 4 + row group
  const int ch = grp / 4, r0 = (grp % 4) * 4;
  const bool vwork = tid < kVItems;
  if (vwork) {
    float st[5][4];
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      float v[14];
#pragma unroll
      for (int i = 0; i < 14; ++i) v[i] = s_h[q][ch][r0 + i][tx];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < kWin; ++k) acc = fmaf(a.win.w[k], v[j + k], acc);
        st[q][j] = acc;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float mu1 = st[0][j], mu2 = st[1][j];
      const float s11 = st[2][j] - mu1 * mu1, s22 = st[3][j] - mu2 * mu2, s12 = st[4][j] - mu1 * mu2;
      const float A1 = 2.f * mu1 * mu2 + kC1, A2 = 2.f * s12 + kC2, B1 = mu1 * mu1 + mu2 * mu2 + kC1, B2 = s11 + s22 + kC2;
      const float iB1 = 1.f / B1, iB2 = 1.f / B2;
      const float S = A1 * A2 * iB1 * iB2;
      // partial derivatives of S w.r.t. the x-side window statistics (mu1 total, through s11 and s12 as well)
      const float d_e11 = -S * iB2;
      const float d_e12 = 2.f * A1 * iB1 * iB2;
      const float d_mu = 2.f * mu2 * A2 * iB1 * iB2 - 2.f * mu1 * S * iB1 - 2.f * mu1 * d_e11 - mu2 * d_e12;
      const bool inside = x0 + tx < a.W && y0 + r0 + j < a.H;
      m_out[0][j] = inside ? d_mu : 0.f;
      m_out[1][j] = inside ? d_e11 : 0.f;
      m_out[2][j] = inside ? d_e12 : 0.f;
      if (inside) {
        const float d = s_xy[0][ch][r0 + j + kHalo][tx + kHalo] - s_xy[1][ch][r0 + j + kHalo][tx + kHalo];
        part += a.l1_scale * fabsf(d) + a.ssim_scale * (1.f - S);
      }
    }
  }
  __syncthreads();  // the patches are dead: stage the three maps [map][pixel][channel] for coalesced stores
  float* s_out = &s_xy[0][0][0][0];
  if (vwork) {
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
      for (int j = 0; j < 4; ++j) s_out[m * (kThreads * 3) + ((r0 + j) * kTile + tx) * 3 + ch] = m_out[m][j];
  }
  __syncthreads();
  const int64_t plane = (int64_t)a.n_img * a.H * a.W * 3;
  const TileElems te = tile_elems(img, a.H, a.W, y0, x0, tid);
#pragma unroll
  for (int m = 0; m < 3; ++m)
#pragma unroll
    for (int sw = 0; sw < 3; ++sw)
      if (te.in[sw]) a.maps[m * plane + te.off[sw]] = s_out[m * (kThreads * 3) + tid + kThreads * sw];
  part = chs_warp_sum(part);
  if ((tid & 31) == 0) s_red[tid >> 5] = part;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) s += s_red[w];
    atomicAdd(a.loss_acc, (double)s);
  }
}

__global__ void __launch_bounds__(kThreads, 4) ssim_bwd_kernel(SsimArgs a) {
  __shared__ __align__(16) float s_m[3][kExt][kPW];    // one derivative map, three channels, with halo; reused for the output
  __shared__ __align__(16) float s_h[3][kExt][kTile];  // its horizontal pass
  __shared__ float s_c[2][kThreads * 3];               // x and y of the tile's own pixels, [pixel][channel]
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile, img = blockIdx.z;
  const int64_t plane = (int64_t)a.n_img * a.H * a.W * 3;
  const TileElems te = tile_elems(img, a.H, a.W, y0, x0, tid);
#pragma unroll
  for (int sw = 0; sw < 3; ++sw) {
    s_c[0][tid + kThreads * sw] = te.in[sw] ? a.x[te.off[sw]] : 0.f;
    s_c[1][tid + kThreads * sw] = te.in[sw] ? a.y[te.off[sw]] : 0.f;
  }
  const int tx = tid % kTile, grp = tid / kTile;
  const int ch = grp / 4, r0 = (grp % 4) * 4;
  const bool vwork = tid < kVItems;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int m = 0; m < 3; ++m) {  // dS/dmu1, dS/dE[x^2], dS/dE[xy]
    __syncthreads();
    load_patch(a.maps + m * plane, img, a.H, a.W, y0, x0, s_m, tid);
    __syncthreads();
    for (int item = tid; item < kHItems; item += kThreads) {
      const int hc = item / (kExt * 4), rem = item % (kExt * 4);
      const int row = rem / 4, c0 = (rem % 4) * 4;
      float v[16];
      load16(&s_m[hc][row][c0], v);
      float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 14; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = i - j;
          if (k >= 0 && k < kWin) o[j] = fmaf(a.win.w[k], v[i], o[j]);
        }
      *reinterpret_cast<float4*>(&s_h[hc][row][c0]) = make_float4(o[0], o[1], o[2], o[3]);
    }
    __syncthreads();
    if (vwork) {
      float v[14];
#pragma unroll
      for (int i = 0; i < 14; ++i) v[i] = s_h[ch][r0 + i][tx];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < kWin; ++k) s = fmaf(a.win.w[k], v[j + k], s);
        const int pi = ((r0 + j) * kTile + tx) * 3 + ch;
        acc[j] += m == 0 ? s : (m == 1 ? 2.f * s_c[0][pi] * s : s_c[1][pi] * s);
      }
    }
  }
  __syncthreads();
  float* s_out = &s_m[0][0][0];
  if (vwork) {
#pragma unroll
    for (int j = 0; j < 4; ++j) s_out[((r0 + j) * kTile + tx) * 3 + ch] = acc[j];
  }
  __syncthreads();
#pragma unroll
  for (int sw = 0; sw < 3; ++sw) {
    if (te.in[sw]) {
      const int i = tid + kThreads * sw;
      const float d = s_c[0][i] - s_c[1][i];
      const float l1 = d > 0.f ? a.l1_scale : (d < 0.f ? -a.l1_scale : 0.f);
      a.v_x[te.off[sw]] = l1 - a.ssim_scale * s_out[i];
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
