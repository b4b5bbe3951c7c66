// chs_project.cu — K1 project_fwd and K9 project_bwd.
//
// Both kernels are HBM streams: one thread per Gaussian, looping over the cameras in-thread so the
// Gaussian attributes are read once (44 B) and the 3D covariance is built once, while the per
// (camera, Gaussian) records are written/read fully coalesced ([C,N] camera-major planes).
// [N,3] attribute arrays are staged through shared memory with 128-bit streaming loads.
// project_bwd accumulates the Gaussian gradients in registers across cameras (no atomics on the
// Gaussian gradient buffer); the per-camera pose gradients are warp-reduced, block-reduced in shared
// memory and added to fp64 accumulators.
#include "chs_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kCamFloats = 16;  // R[9], t[3], fx, fy, cx, cy

__device__ __forceinline__ void load_cameras(float* s_cam, int C, int n_virtual, int ks_per_camera,
                                             const float* __restrict__ viewmats, const float* __restrict__ Ks) {
  for (int i = threadIdx.x; i < C * kCamFloats; i += blockDim.x) {
    int c = i / kCamFloats, k = i % kCamFloats;
    float v;
    if (k < 9) {
      v = viewmats[c * 16 + (k / 3) * 4 + (k % 3)];
    } else if (k < 12) {
      v = viewmats[c * 16 + (k - 9) * 4 + 3];
    } else {
      const float* K = Ks + (size_t)(ks_per_camera ? c : c / n_virtual) * 9;
      v = (k == 12) ? K[0] : (k == 13) ? K[4] : (k == 14) ? K[2] : K[5];
    }
    s_cam[i] = v;
  }
}

__device__ __forceinline__ void read_camera(const float* s_cam, int c, ChsCam<float>& cam) {
  const float* p = s_cam + c * kCamFloats;
#pragma unroll
  for (int i = 0; i < 9; ++i) cam.R[i] = p[i];
  cam.t[0] = p[9]; cam.t[1] = p[10]; cam.t[2] = p[11];
  cam.fx = p[12]; cam.fy = p[13]; cam.cx = p[14]; cam.cy = p[15];
}

struct ProjectFwdArgs {
  int N, C, n_virtual, ks_per_camera, tile_w, tile_h, tight_bounds, pose_fused;
  float width, height, near_plane, far_plane, eps2d;
  const float *means, *quats, *scales, *opacities, *colors, *viewmats, *Ks;
  float4* geom;
  float* conic_c;
  float* depths;
  int32_t* radii;
  int32_t* tiles_touched;
  float4* rgbo;
};

__global__ void __launch_bounds__(kThreads) project_fwd_kernel(ProjectFwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* s_cam = smem;
  float* s_means = s_cam + ((a.C * kCamFloats + 3) & ~3);
  float* s_scales = s_means + kThreads * 3;
  float* s_colors = s_scales + kThreads * 3;

  const int64_t row0 = (int64_t)blockIdx.x * kThreads;
  const int rows = (int)min((int64_t)kThreads, (int64_t)a.N - row0);
  load_cameras(s_cam, a.C, a.n_virtual, a.ks_per_camera, a.viewmats, a.Ks);
  chs_stage_rows3(a.means, row0, rows, s_means);
  chs_stage_rows3(a.scales, row0, rows, s_scales);
  chs_stage_rows3(a.colors, row0, rows, s_colors);
  __syncthreads();
  const int t = threadIdx.x;
  if (t >= rows) return;
  const int64_t g = row0 + t;

  const float4 q4 = chs_ldg_stream(reinterpret_cast<const float4*>(a.quats) + g);
  const float q[4] = {q4.x, q4.y, q4.z, q4.w};
  const float s[3] = {s_scales[t * 3], s_scales[t * 3 + 1], s_scales[t * 3 + 2]};
  const float mu[3] = {s_means[t * 3], s_means[t * 3 + 1], s_means[t * 3 + 2]};
  float S[6];
  chs_cov3d(q, s, S);
  const float opac = a.opacities[g];
  a.rgbo[g] = make_float4(s_colors[t * 3], s_colors[t * 3 + 1], s_colors[t * 3 + 2], opac);

  // pose_fused: the union of the frame's per-pose tile rectangles (poses where the Gaussian is live)
  int ux0 = 0, uy0 = 0, ux1 = 0, uy1 = 0, k = 0;
  for (int c = 0; c < a.C; ++c) {
    ChsCam<float> cam;
    read_camera(s_cam, c, cam);
    ChsProj<float> pr;
    int radius = chs_project_fwd(mu, S, cam, a.width, a.height, a.near_plane, a.far_plane, a.eps2d, pr);
    if (radius > 0 && a.tight_bounds) radius = chs_tight_radii(pr.sxx, pr.syy, opac, radius);  // packed rx | ry << 16, or 0
    int touched = 0;
    ChsTileRect r;
    r.x0 = r.y0 = r.x1 = r.y1 = 0;
    if (radius != 0) {  // packed tight radii are an unsigned pair: ry >= 32768 sets bit 31
      r = chs_tile_bounds_of(pr.mx, pr.my, radius, a.tight_bounds, a.tile_w, a.tile_h);
      touched = (r.x1 - r.x0) * (r.y1 - r.y0);
    }
    const int64_t o = (int64_t)c * a.N + g;
    a.depths[o] = pr.depth;
    a.radii[o] = radius;
    if (!a.pose_fused) {
      a.geom[o] = make_float4(pr.mx, pr.my, pr.ca, pr.cb);
      a.conic_c[o] = pr.cc;
      a.tiles_touched[o] = touched;
      continue;
    }
    if (touched > 0) {
      a.geom[o] = make_float4(pr.mx, pr.my, pr.ca, pr.cb);
      a.conic_c[o] = pr.cc;
      if (ux1 > ux0) {
        ux0 = min(ux0, r.x0); uy0 = min(uy0, r.y0); ux1 = max(ux1, r.x1); uy1 = max(uy1, r.y1);
      } else {
        ux0 = r.x0; uy0 = r.y0; ux1 = r.x1; uy1 = r.y1;
      }
    } else {
      // a pose that does not see the Gaussian still walks the frame's list: give it a record whose alpha is 0 everywhere
      // (mean far off screen under a unit conic: the exponent is ~ -1e12)
      a.geom[o] = make_float4(-1e6f, -1e6f, 1.f, 0.f);
      a.conic_c[o] = 1.f;
    }
    if (++k == a.n_virtual) {
      a.tiles_touched[(int64_t)(c / a.n_virtual) * a.N + g] = (ux1 - ux0) * (uy1 - uy0);
      ux0 = uy0 = ux1 = uy1 = 0;
      k = 0;
    }
  }
}

struct ProjectBwdArgs {
  int N, C, n_virtual, ks_per_camera, quat_section_aligned;
  float width, height, eps2d;
  const float *means, *quats, *scales, *viewmats, *Ks;
  const int32_t* radii;
  const float4* v_geom;
  const float4* v_cogr;
  const float* v_blue;
  float* grads_flat;
  double* v_cam_acc;  // [C, 12] = v_R (9) | v_t (3)
};

template <int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks) project_bwd_kernel(ProjectBwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* s_cam = smem;
  float* s_vcam = s_cam + ((a.C * kCamFloats + 3) & ~3);  // [C,12] block partial sums
  float* s_a = s_vcam + ((a.C * 12 + 3) & ~3);            // means in, v_means out
  float* s_b = s_a + kThreads * 3;                        // scales in, v_scales out
  float* s_c = s_b + kThreads * 3;                        // v_colors out

  const int64_t row0 = (int64_t)blockIdx.x * kThreads;
  const int rows = (int)min((int64_t)kThreads, (int64_t)a.N - row0);
  load_cameras(s_cam, a.C, a.n_virtual, a.ks_per_camera, a.viewmats, a.Ks);
  for (int i = threadIdx.x; i < a.C * 12; i += blockDim.x) s_vcam[i] = 0.0f;
  chs_stage_rows3(a.means, row0, rows, s_a);
  chs_stage_rows3(a.scales, row0, rows, s_b);
  __syncthreads();
  const int t = threadIdx.x;
  const int lane = t & 31;
  const bool live = t < rows;
  const int64_t g = row0 + (live ? t : 0);

  float q[4] = {1.f, 0.f, 0.f, 0.f}, s[3] = {1.f, 1.f, 1.f}, mu[3] = {0.f, 0.f, 0.f};
  if (live) {
    const float4 q4 = chs_ldg_stream(reinterpret_cast<const float4*>(a.quats) + g);
    q[0] = q4.x; q[1] = q4.y; q[2] = q4.z; q[3] = q4.w;
    s[0] = s_b[t * 3]; s[1] = s_b[t * 3 + 1]; s[2] = s_b[t * 3 + 2];
    mu[0] = s_a[t * 3]; mu[1] = s_a[t * 3 + 1]; mu[2] = s_a[t * 3 + 2];
  }
  float S[6];
  chs_cov3d(q, s, S);
  float v_mu[3] = {0.f, 0.f, 0.f}, G[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float v_op = 0.f, v_rgb[3] = {0.f, 0.f, 0.f};

  for (int c = 0; c < a.C; ++c) {
    const int64_t o = (int64_t)c * a.N + g;
    const bool hit = live && a.radii[o] != 0;
    float vcam[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) vcam[i] = 0.f;
    if (hit) {
      const float4 vg = a.v_geom[o];
      const float4 vc = a.v_cogr[o];
      v_op += vc.y;
      v_rgb[0] += vc.z;
      v_rgb[1] += vc.w;
      v_rgb[2] += a.v_blue[o];
      ChsCam<float> cam;
      read_camera(s_cam, c, cam);
      chs_project_bwd(mu, S, cam, a.width, a.height, a.eps2d, vg.x, vg.y, vg.z, vg.w, vc.x, v_mu, G, vcam, vcam + 9);
    }
    if (__any_sync(CHS_FULL_MASK, hit)) {
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        float r = chs_warp_sum(vcam[i]);
        if (lane == 0) atomicAdd(&s_vcam[c * 12 + i], r);
      }
    }
  }
  float v_q[4] = {0.f, 0.f, 0.f, 0.f}, v_s[3] = {0.f, 0.f, 0.f};
  if (live) chs_cov3d_bwd(q, s, G, v_q, v_s);
  __syncthreads();  // everyone is done reading s_a / s_b as inputs; s_vcam complete
  if (live) {
    s_a[t * 3] = v_mu[0]; s_a[t * 3 + 1] = v_mu[1]; s_a[t * 3 + 2] = v_mu[2];
    s_b[t * 3] = v_s[0]; s_b[t * 3 + 1] = v_s[1]; s_b[t * 3 + 2] = v_s[2];
    s_c[t * 3] = v_rgb[0]; s_c[t * 3 + 1] = v_rgb[1]; s_c[t * 3 + 2] = v_rgb[2];
    const int64_t N = a.N;
    float* dq = a.grads_flat + 3 * N + 4 * g;
    if (a.quat_section_aligned) {
      *reinterpret_cast<float4*>(dq) = make_float4(v_q[0], v_q[1], v_q[2], v_q[3]);
    } else {
      dq[0] = v_q[0]; dq[1] = v_q[1]; dq[2] = v_q[2]; dq[3] = v_q[3];
    }
    a.grads_flat[10 * N + g] = v_op;
  }
  __syncthreads();
  {
    const int64_t N = a.N;
    float* d_means = a.grads_flat + row0 * 3;
    float* d_scales = a.grads_flat + 7 * N + row0 * 3;
    float* d_colors = a.grads_flat + 11 * N + row0 * 3;
    for (int i = t; i < rows * 3; i += kThreads) {
      d_means[i] = s_a[i];
      d_scales[i] = s_b[i];
      d_colors[i] = s_c[i];
    }
  }
  for (int i = t; i < a.C * 12; i += kThreads) {
    float v = s_vcam[i];
    if (v != 0.0f) atomicAdd(&a.v_cam_acc[i], (double)v);
  }
}

__global__ void finalize_viewmat_grads(const double* acc, float* v_viewmats, int C) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * 16) return;
  int c = i / 16, r = (i % 16) / 4, k = i % 4;
  float v = 0.f;
  if (r < 3) v = (float)(k < 3 ? acc[c * 12 + r * 3 + k] : acc[c * 12 + 9 + r]);
  v_viewmats[i] = v;
}

}  // namespace

extern "C" int chs_project_fwd(const chs_config* cfg, const float* means, const float* quats, const float* scales,
                               const float* opacities, const float* colors, const float* viewmats, const float* Ks,
                               float* geom, float* conic_c, float* depths, int32_t* radii, int32_t* tiles_touched,
                               float* rgbo, void* stream) {
  ChsDims d;
  int st = chs_make_dims(cfg, &d);
  if (st) return st;
  CHS_REQUIRE(means && quats && scales && opacities && colors && viewmats && Ks, "chs_project_fwd: null input");
  CHS_REQUIRE(geom && conic_c && depths && radii && tiles_touched && rgbo, "chs_project_fwd: null output");
  CHS_REQUIRE(((uintptr_t)means | (uintptr_t)quats | (uintptr_t)scales | (uintptr_t)colors | (uintptr_t)geom | (uintptr_t)rgbo) % 16 == 0,
              "chs_project_fwd: means/quats/scales/colors/geom/rgbo must be 16-byte aligned");
  if (d.N == 0) return CHS_OK;
  ProjectFwdArgs a;
  a.N = d.N; a.C = d.C; a.n_virtual = d.n; a.ks_per_camera = cfg->ks_per_camera; a.tile_w = d.tile_w; a.tile_h = d.tile_h;
  a.tight_bounds = cfg->tight_bounds != 0;
  a.pose_fused = cfg->pose_fused != 0;
  a.width = (float)d.W; a.height = (float)d.H; a.near_plane = cfg->near_plane; a.far_plane = cfg->far_plane; a.eps2d = cfg->eps2d;
  a.means = means; a.quats = quats; a.scales = scales; a.opacities = opacities; a.colors = colors; a.viewmats = viewmats; a.Ks = Ks;
  a.geom = (float4*)geom; a.conic_c = conic_c; a.depths = depths; a.radii = radii; a.tiles_touched = tiles_touched; a.rgbo = (float4*)rgbo;
  size_t smem = (((size_t)d.C * kCamFloats + 3) & ~(size_t)3) * 4 + 3 * kThreads * 3 * 4;
  int blocks = (d.N + kThreads - 1) / kThreads;
  // the camera table lives in dynamic shared memory (64 B per camera): above the 48 KB default the kernel must opt in
  if (smem > 48 * 1024) CHS_CUDA(cudaFuncSetAttribute(project_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  project_fwd_kernel<<<blocks, kThreads, smem, (cudaStream_t)stream>>>(a);
  CHS_LAUNCH_CHECK();
  return CHS_OK;
}

extern "C" int chs_project_bwd(const chs_config* cfg, const float* means, const float* quats, const float* scales,
                               const float* viewmats, const float* Ks, const int32_t* radii, const float* v_geom,
                               const float* v_cogr, const float* v_blue, float* grads_flat, float* v_viewmats,
                               void* workspace, uint64_t workspace_bytes, void* stream) {
  ChsDims d;
  int st = chs_make_dims(cfg, &d);
  if (st) return st;
  CHS_REQUIRE(means && quats && scales && viewmats && Ks && radii && v_geom && v_cogr && v_blue, "chs_project_bwd: null input");
  CHS_REQUIRE(grads_flat && v_viewmats && workspace, "chs_project_bwd: null output/workspace");
  CHS_REQUIRE(((uintptr_t)means | (uintptr_t)quats | (uintptr_t)scales | (uintptr_t)v_geom | (uintptr_t)v_cogr) % 16 == 0,
              "chs_project_bwd: vector inputs must be 16-byte aligned");
  if (workspace_bytes < (uint64_t)d.C * 12 * sizeof(double)) {
    chs_set_error("chs_project_bwd: workspace too small");
    return CHS_ERR_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t s = (cudaStream_t)stream;
  double* acc = (double*)workspace;
  CHS_CUDA(cudaMemsetAsync(acc, 0, (size_t)d.C * 12 * sizeof(double), s));
  if (d.N > 0) {
    ProjectBwdArgs a;
    a.N = d.N; a.C = d.C; a.n_virtual = d.n; a.ks_per_camera = cfg->ks_per_camera;
    a.width = (float)d.W; a.height = (float)d.H; a.eps2d = cfg->eps2d;
    a.means = means; a.quats = quats; a.scales = scales; a.viewmats = viewmats; a.Ks = Ks; a.radii = radii;
    a.v_geom = (const float4*)v_geom; a.v_cogr = (const float4*)v_cogr; a.v_blue = v_blue;
    a.grads_flat = grads_flat; a.v_cam_acc = acc;
    a.quat_section_aligned = ((uintptr_t)(grads_flat + 3 * (size_t)d.N)) % 16 == 0;
    size_t smem = ((((size_t)d.C * kCamFloats + 3) & ~(size_t)3) + (((size_t)d.C * 12 + 3) & ~(size_t)3)) * 4 + 3 * kThreads * 3 * 4;
    int blocks = (d.N + kThreads - 1) / kThreads;
    // resident blocks per SM (chs_config.tune_project_bwd).  The kernel is latency-bound, so occupancy beats spills (r2p, c3):
    // 2 (115 registers, no spills) 0.450 ms | 3 (80 registers) 0.346 | 4 (64 registers, 176 B of spills) 0.318 (default)
    // r2x negative result: requesting the next camera's 40 bytes before this camera's arithmetic (software prefetch) made it
    // slower at every occupancy (0.412 | 0.349 | 0.355): the extra live registers cost more than the overlap gains
    const int mb = cfg->tune_project_bwd ? cfg->tune_project_bwd : 4;
    if (smem > 48 * 1024) {
      CHS_CUDA(cudaFuncSetAttribute(project_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CHS_CUDA(cudaFuncSetAttribute(project_bwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CHS_CUDA(cudaFuncSetAttribute(project_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    if (mb == 3)
      project_bwd_kernel<3><<<blocks, kThreads, smem, s>>>(a);
    else if (mb == 4)
      project_bwd_kernel<4><<<blocks, kThreads, smem, s>>>(a);
    else
      project_bwd_kernel<2><<<blocks, kThreads, smem, s>>>(a);
    CHS_LAUNCH_CHECK();
  }
  finalize_viewmat_grads<<<(d.C * 16 + 255) / 256, 256, 0, s>>>(acc, v_viewmats, d.C);
  CHS_LAUNCH_CHECK();
  return CHS_OK;
}
