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
  const float *hdr_mean, *exposure, *crf_params, *v_ldr;
  float* v_hdr;
  double* acc_crf;       // [3 * (3 hd + 1)]
  double* acc_exposure;  // [B]
};

__global__ void __launch_bounds__(kThreads) crf_bwd_kernel(CrfBwdArgs a) {
  extern __shared__ float smem[];
  const int stride = 3 * a.hd + 1;
  float* s_p = smem;                // parameters [3, stride]
  float* s_g = smem + 3 * stride;   // block-partial parameter gradients [3, stride]
  const int frame = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31;
  const bool mlp = a.crf_kind == CHS_CRF_MLP;
  if (mlp)
    for (int i = tid; i < 3 * stride; i += kThreads) {
      s_p[i] = a.crf_params[i];
      s_g[i] = 0.f;
    }
  __syncthreads();
  const float dt = a.exposure[frame];
  const float scale = dt / (float)a.n_virtual;
  float z[kPix][3], gy[kPix][3];
  float v_dt = 0.f;
#pragma unroll
  for (int q = 0; q < kPix; ++q) {
    const int64_t pix = ((int64_t)blockIdx.x * kPix + q) * kThreads + tid;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) z[q][ch] = gy[q][ch] = 0.f;
    if (pix >= a.P) continue;
    const int64_t o = ((int64_t)frame * a.P + pix) * 3;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const float h = a.hdr_mean[o + ch];
      const float vy = a.v_ldr[o + ch];
      float vx;
      if (mlp) {
        const float* p = s_p + ch * stride;
        const float xe = dt * h + CHS_CRF_EPS;
        const float zz = logf(xe);
        float acc = p[3 * a.hd], dz = 0.f;
        for (int j = 0; j < a.hd; ++j) {
          const float pre = fmaf(p[j], zz, p[a.hd + j]);
          if (pre > 0.f) {
            acc = fmaf(p[2 * a.hd + j], pre, acc);
            dz = fmaf(p[2 * a.hd + j], p[j], dz);
          }
        }
        const float y = 1.f / (1.f + expf(-acc));
        const float g = vy * y * (1.f - y);
        z[q][ch] = zz;
        gy[q][ch] = g;
        vx = g * dz / xe;
      } else {
        vx = vy;
      }
      v_dt = fmaf(vx, h, v_dt);
      a.v_hdr[o + ch] = vx * scale;
    }
  }
  // exposure (brightness path): block reduce -> one fp64 atomic
  v_dt = chs_warp_sum(v_dt);
  __shared__ float s_dt[kThreads / 32];
  if (lane == 0) s_dt[tid >> 5] = v_dt;
  if (mlp) {
    for (int ch = 0; ch < 3; ++ch) {
      const float* p = s_p + ch * stride;
      float g_b2 = 0.f;
#pragma unroll
      for (int q = 0; q < kPix; ++q) g_b2 += gy[q][ch];
      g_b2 = chs_warp_sum(g_b2);
      if (lane == 0) atomicAdd(&s_g[ch * stride + 3 * a.hd], g_b2);
      for (int j = 0; j < a.hd; ++j) {
        const float w1 = p[j], b1 = p[a.hd + j], w2 = p[2 * a.hd + j];
        float g_w1 = 0.f, g_b1 = 0.f, g_w2 = 0.f;
#pragma unroll
        for (int q = 0; q < kPix; ++q) {
          const float pre = fmaf(w1, z[q][ch], b1);
          if (pre > 0.f) {
            g_w2 = fmaf(gy[q][ch], pre, g_w2);
            const float dh = gy[q][ch] * w2;
            g_w1 = fmaf(dh, z[q][ch], g_w1);
            g_b1 += dh;
          }
        }
        g_w1 = chs_warp_sum(g_w1);
        g_b1 = chs_warp_sum(g_b1);
        g_w2 = chs_warp_sum(g_w2);
        if (lane == 0) {
          atomicAdd(&s_g[ch * stride + j], g_w1);
          atomicAdd(&s_g[ch * stride + a.hd + j], g_b1);
          atomicAdd(&s_g[ch * stride + 2 * a.hd + j], g_w2);
        }
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) s += s_dt[w];
    atomicAdd(&a.acc_exposure[frame], (double)s);
  }
  if (mlp)
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
  CHS_REQUIRE(!mlp || (crf_params && v_crf_params), "chs_crf_bwd: crf_params / v_crf_params required for the MLP CRF");
  if (cfg->crf_before_average) {
    chs_set_error("chs_crf_bwd: crf_before_average=1 is not implemented in CUDA yet");
    return CHS_ERR_UNSUPPORTED;
  }
  const int n_par = mlp ? 3 * (3 * cfg->crf_hidden + 1) : 0;
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
    a.P = d.P; a.n_virtual = d.n; a.crf_kind = cfg->crf_kind; a.hd = mlp ? cfg->crf_hidden : 0;
    a.hdr_mean = hdr_mean; a.exposure = exposure; a.crf_params = crf_params; a.v_ldr = v_ldr; a.v_hdr = v_hdr;
    a.acc_crf = acc; a.acc_exposure = acc + n_par;
    dim3 grid((unsigned)((d.P + (int64_t)kThreads * kPix - 1) / ((int64_t)kThreads * kPix)), d.B);
    size_t smem = mlp ? (size_t)2 * n_par * sizeof(float) : 16;
    crf_bwd_kernel<<<grid, kThreads, smem, s>>>(a);
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
