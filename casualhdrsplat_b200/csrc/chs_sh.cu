// chs_sh.cu — view-dependent HDR colour from spherical harmonics, the step immediately before the
// formation path in a 3DGS trainer.  One thread per Gaussian looping
// over the cameras in-thread (coefficients read once); forward writes the per-camera (r,g,b,opacity)
// records the blend kernels gather; backward accumulates the coefficient gradients in registers across
// cameras (no atomics) and reduces the camera-centre gradients warp -> shared memory -> fp64.
#include "chs_common.cuh"
#include "chs_sh.cuh"

namespace {

constexpr int kThreads = 128;

__device__ __forceinline__ void load_cam_pose(const float* __restrict__ viewmats, int c, float R[9], float t[3]) {
  const float* v = viewmats + (size_t)c * 16;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) R[i * 3 + j] = v[i * 4 + j];
    t[i] = v[i * 4 + 3];
  }
}

template <int DEG>
__global__ void __launch_bounds__(kThreads) sh_fwd_kernel(int N, int C, const float* __restrict__ sh, const float* __restrict__ means,
                                                          const float* __restrict__ opacities, const float* __restrict__ viewmats,
                                                          float4* __restrict__ rgbo_c) {
  constexpr int K = (DEG + 1) * (DEG + 1);
  extern __shared__ float s_cp[];  // [C,3] camera centres
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float R[9], t[3], cp[3];
    load_cam_pose(viewmats, c, R, t);
    chs_campos(R, t, cp);
    s_cp[c * 3] = cp[0]; s_cp[c * 3 + 1] = cp[1]; s_cp[c * 3 + 2] = cp[2];
  }
  __syncthreads();
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= N) return;
  float coef[K * 3];
#pragma unroll
  for (int i = 0; i < K * 3; ++i) coef[i] = sh[g * (K * 3) + i];
  const float mu[3] = {means[g * 3], means[g * 3 + 1], means[g * 3 + 2]};
  const float op = opacities[g];
  for (int c = 0; c < C; ++c) {
    const float cp[3] = {s_cp[c * 3], s_cp[c * 3 + 1], s_cp[c * 3 + 2]};
    float rgb[3];
    chs_sh_color<float>(DEG, coef, mu, cp, rgb);
    rgbo_c[(int64_t)c * N + g] = make_float4(rgb[0], rgb[1], rgb[2], op);
  }
}

template <int DEG>
__global__ void __launch_bounds__(kThreads) sh_bwd_kernel(int N, int C, const float* __restrict__ sh, const float* __restrict__ means,
                                                          const float* __restrict__ viewmats, const int32_t* __restrict__ radii,
                                                          const float4* __restrict__ v_cogr, const float* __restrict__ v_blue,
                                                          float* __restrict__ v_sh, float* __restrict__ v_means, double* __restrict__ acc_cp) {
  constexpr int K = (DEG + 1) * (DEG + 1);
  extern __shared__ float smem[];
  float* s_cp = smem;           // [C,3]
  float* s_vcp = smem + C * 3;  // [C,3] block partial of v_campos
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float R[9], t[3], cp[3];
    load_cam_pose(viewmats, c, R, t);
    chs_campos(R, t, cp);
    for (int i = 0; i < 3; ++i) {
      s_cp[c * 3 + i] = cp[i];
      s_vcp[c * 3 + i] = 0.f;
    }
  }
  __syncthreads();
  const int64_t g0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = g0 < N;
  const int64_t g = live ? g0 : 0;
  const int lane = threadIdx.x & 31;
  float coef[K * 3], vcoef[K * 3];
#pragma unroll
  for (int i = 0; i < K * 3; ++i) {
    coef[i] = sh[g * (K * 3) + i];
    vcoef[i] = 0.f;
  }
  const float mu[3] = {means[g * 3], means[g * 3 + 1], means[g * 3 + 2]};
  float v_mu[3] = {0.f, 0.f, 0.f};
  for (int c = 0; c < C; ++c) {
    const int64_t o = (int64_t)c * N + g;
    const bool hit = live && radii[o] != 0;  // packed tight radii may have bit 31 set
    float v_cp[3] = {0.f, 0.f, 0.f};
    if (hit) {
      const float4 vc = v_cogr[o];
      const float v_rgb[3] = {vc.z, vc.w, v_blue[o]};
      const float cp[3] = {s_cp[c * 3], s_cp[c * 3 + 1], s_cp[c * 3 + 2]};
      chs_sh_color_bwd<float>(DEG, coef, mu, cp, v_rgb, vcoef, v_mu, v_cp);
    }
    if (DEG > 0 && __any_sync(CHS_FULL_MASK, hit)) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float r = chs_warp_sum(v_cp[i]);
        if (lane == 0) atomicAdd(&s_vcp[c * 3 + i], r);
      }
    }
  }
  if (live) {
#pragma unroll
    for (int i = 0; i < K * 3; ++i) v_sh[g * (K * 3) + i] = vcoef[i];
    v_means[g * 3] = v_mu[0]; v_means[g * 3 + 1] = v_mu[1]; v_means[g * 3 + 2] = v_mu[2];
  }
  __syncthreads();
  if (DEG > 0)
    for (int i = threadIdx.x; i < C * 3; i += blockDim.x)
      if (s_vcp[i] != 0.f) atomicAdd(&acc_cp[i], (double)s_vcp[i]);
}

// v_viewmats [C,4,4] from the camera-centre gradients: campos = -R^T t
__global__ void sh_finalize_kernel(int C, const float* __restrict__ viewmats, const double* __restrict__ acc_cp, float* __restrict__ v_viewmats) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float R[9], t[3];
  load_cam_pose(viewmats, c, R, t);
  const float v_cp[3] = {(float)acc_cp[c * 3], (float)acc_cp[c * 3 + 1], (float)acc_cp[c * 3 + 2]};
  float vR[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, vt[3] = {0, 0, 0};
  chs_campos_bwd(R, t, v_cp, vR, vt);
  float* o = v_viewmats + (size_t)c * 16;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) o[i * 4 + j] = vR[i * 3 + j];
    o[i * 4 + 3] = vt[i];
  }
  o[12] = o[13] = o[14] = o[15] = 0.f;
}

}  // namespace

extern "C" int chs_sh_fwd(const chs_config* cfg, int32_t sh_degree, const float* sh_coeffs, const float* means, const float* opacities,
                          const float* viewmats, float* rgbo_c, void* stream) {
  ChsDims d;
  int st = chs_make_dims(cfg, &d);
  if (st) return st;
  CHS_REQUIRE(sh_degree >= 0 && sh_degree <= 3, "chs_sh_fwd: sh_degree must be 0..3");
  CHS_REQUIRE(sh_coeffs && means && opacities && viewmats && rgbo_c, "chs_sh_fwd: null pointer");
  if (d.N == 0 || d.C == 0) return CHS_OK;
  const int blocks = (d.N + kThreads - 1) / kThreads;
  const size_t smem = (size_t)d.C * 3 * sizeof(float);
  cudaStream_t s = (cudaStream_t)stream;
  float4* out = (float4*)rgbo_c;
  switch (sh_degree) {
    case 0: sh_fwd_kernel<0><<<blocks, kThreads, smem, s>>>(d.N, d.C, sh_coeffs, means, opacities, viewmats, out); break;
    case 1: sh_fwd_kernel<1><<<blocks, kThreads, smem, s>>>(d.N, d.C, sh_coeffs, means, opacities, viewmats, out); break;
    case 2: sh_fwd_kernel<2><<<blocks, kThreads, smem, s>>>(d.N, d.C, sh_coeffs, means, opacities, viewmats, out); break;
    default: sh_fwd_kernel<3><<<blocks, kThreads, smem, s>>>(d.N, d.C, sh_coeffs, means, opacities, viewmats, out); break;
  }
  CHS_LAUNCH_CHECK();
  return CHS_OK;
}

extern "C" int chs_sh_bwd(const chs_config* cfg, int32_t sh_degree, const float* sh_coeffs, const float* means, const float* viewmats,
                          const int32_t* radii, const float* v_cogr, const float* v_blue, float* v_sh, float* v_means,
                          float* v_viewmats, void* workspace, uint64_t workspace_bytes, void* stream) {
  ChsDims d;
  int st = chs_make_dims(cfg, &d);
  if (st) return st;
  CHS_REQUIRE(sh_degree >= 0 && sh_degree <= 3, "chs_sh_bwd: sh_degree must be 0..3");
  CHS_REQUIRE(sh_coeffs && means && viewmats && radii && v_cogr && v_blue && v_sh && v_means && v_viewmats && workspace,
              "chs_sh_bwd: null pointer");
  if (workspace_bytes < (uint64_t)d.C * 3 * sizeof(double)) {
    chs_set_error("chs_sh_bwd: workspace too small");
    return CHS_ERR_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t s = (cudaStream_t)stream;
  double* acc = (double*)workspace;
  CHS_CUDA(cudaMemsetAsync(acc, 0, (size_t)d.C * 3 * sizeof(double), s));
  if (d.N > 0 && d.C > 0) {
    const int blocks = (d.N + kThreads - 1) / kThreads;
    const size_t smem = (size_t)d.C * 6 * sizeof(float);
    const float4* vc = (const float4*)v_cogr;
    switch (sh_degree) {
      case 0: sh_bwd_kernel<0><<<blocks, kThreads, smem, s>>>(d.N, d.C, sh_coeffs, means, viewmats, radii, vc, v_blue, v_sh, v_means, acc); break;
      case 1: sh_bwd_kernel<1><<<blocks, kThreads, smem, s>>>(d.N, d.C, sh_coeffs, means, viewmats, radii, vc, v_blue, v_sh, v_means, acc); break;
      case 2: sh_bwd_kernel<2><<<blocks, kThreads, smem, s>>>(d.N, d.C, sh_coeffs, means, viewmats, radii, vc, v_blue, v_sh, v_means, acc); break;
      default: sh_bwd_kernel<3><<<blocks, kThreads, smem, s>>>(d.N, d.C, sh_coeffs, means, viewmats, radii, vc, v_blue, v_sh, v_means, acc); break;
    }
    CHS_LAUNCH_CHECK();
  }
  if (d.C > 0) {
    sh_finalize_kernel<<<(d.C + 63) / 64, 64, 0, s>>>(d.C, viewmats, acc, v_viewmats);
    CHS_LAUNCH_CHECK();
  }
  return CHS_OK;
}
