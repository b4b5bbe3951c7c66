// chs_spline.cu — K0 spline_sample_fwd / _bwd.
// fp64 on device: C <= 128 poses per step, correctness-critical, not performance-critical.
#include "chs_common.cuh"
#include "chs_spline.cuh"

namespace {

struct SplineArgs {
  int kind, n_knots, n_frames, n_virtual;
  double knot_t0, knot_dt;
  const float *knots, *frame_times, *exposure;
};

__device__ __forceinline__ void camera_segment(const SplineArgs& a, int c, int& s, double& u, double& w) {
  const int i = c / a.n_virtual, k = c % a.n_virtual;
  w = chs_sample_weight(k, a.n_virtual);
  const double time = (double)a.frame_times[i] + w * (double)a.exposure[i];
  chs_spline_segment(a.kind, a.n_knots, a.knot_t0, a.knot_dt, time, s, u);
}

__global__ void spline_fwd_kernel(SplineArgs a, float* viewmats) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.n_frames * a.n_virtual) return;
  int s;
  double u, w;
  camera_segment(a, c, s, u, w);
  const int first = a.kind == CHS_SPLINE_LINEAR ? s : s - 1;
  const int nk = a.kind == CHS_SPLINE_LINEAR ? 2 : 4;
  double k[4][7];
  for (int j = 0; j < nk; ++j)
    for (int e = 0; e < 7; ++e) k[j][e] = (double)a.knots[(first + j) * 7 + e];
  double vm[12];
  chs_spline_viewmat<double>(a.kind, k, u, vm);
  float* o = viewmats + (size_t)c * 16;
  for (int e = 0; e < 12; ++e) o[e] = (float)vm[e];
  o[12] = 0.f; o[13] = 0.f; o[14] = 0.f; o[15] = 1.f;
}

// one thread per (camera, input): forward-mode tangent of the pose w.r.t. that input, contracted
// with v_viewmats.  acc = [v_knots (K*7) | v_frame_times (B) | v_exposure (B)] in fp64.
__global__ void spline_bwd_kernel(SplineArgs a, const float* v_viewmats, double* acc) {
  const int nk = a.kind == CHS_SPLINE_LINEAR ? 2 : 4;
  const int n_in = nk * 7 + 1;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int C = a.n_frames * a.n_virtual;
  if (idx >= C * n_in) return;
  const int c = idx / n_in, in = idx % n_in;
  int s;
  double u, w;
  camera_segment(a, c, s, u, w);
  const int first = a.kind == CHS_SPLINE_LINEAR ? s : s - 1;
  ChsDual k[4][7];
  for (int j = 0; j < nk; ++j)
    for (int e = 0; e < 7; ++e) k[j][e] = ChsDual((double)a.knots[(first + j) * 7 + e], (j * 7 + e == in) ? 1.0 : 0.0);
  ChsDual ud(u, in == nk * 7 ? 1.0 : 0.0);
  ChsDual vm[12];
  chs_spline_viewmat<ChsDual>(a.kind, k, ud, vm);
  const float* v = v_viewmats + (size_t)c * 16;
  double dot = 0.0;
  for (int e = 0; e < 12; ++e) dot += (double)v[e] * vm[e].d;
  if (dot == 0.0) return;
  if (in < nk * 7) {
    atomicAdd(&acc[(first + in / 7) * 7 + in % 7], dot);
  } else {
    const int i = c / a.n_virtual;
    const double du = dot / a.knot_dt;  // u = (time - t0) / dt - s
    atomicAdd(&acc[a.n_knots * 7 + i], du);
    atomicAdd(&acc[a.n_knots * 7 + a.n_frames + i], du * w);
  }
}

__global__ void spline_finalize_kernel(const double* acc, int n_knots, int n_frames, float* v_knots, float* v_ft, float* v_ex) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int nk7 = n_knots * 7;
  if (i < nk7) v_knots[i] = (float)acc[i];
  else if (i < nk7 + n_frames) v_ft[i - nk7] = (float)acc[i];
  else if (i < nk7 + 2 * n_frames) v_ex[i - nk7 - n_frames] = (float)acc[i];
}

int check_spline_args(int32_t kind, const float* knots, int32_t n_knots, double knot_dt, const float* ft, const float* ex,
                      int32_t n_frames, int32_t n_virtual) {
  CHS_REQUIRE(kind == CHS_SPLINE_LINEAR || kind == CHS_SPLINE_CUBIC, "spline: unknown kind %d", kind);
  CHS_REQUIRE(knots && ft && ex, "spline: null input");
  CHS_REQUIRE(n_knots >= (kind == CHS_SPLINE_LINEAR ? 2 : 4), "spline: not enough knots (%d) for kind %d", n_knots, kind);
  CHS_REQUIRE(knot_dt > 0.0, "spline: knot_dt must be positive");
  CHS_REQUIRE(n_frames >= 0 && n_virtual >= 1, "spline: bad n_frames / n_virtual");
  return CHS_OK;
}

}  // namespace

extern "C" int chs_spline_fwd(int32_t kind, const float* knots, int32_t n_knots, double knot_t0, double knot_dt,
                              const float* frame_times, const float* exposure, int32_t n_frames, int32_t n_virtual,
                              float* viewmats, void* stream) {
  int st = check_spline_args(kind, knots, n_knots, knot_dt, frame_times, exposure, n_frames, n_virtual);
  if (st) return st;
  CHS_REQUIRE(viewmats, "chs_spline_fwd: null output");
  const int C = n_frames * n_virtual;
  if (C == 0) return CHS_OK;
  SplineArgs a{kind, n_knots, n_frames, n_virtual, knot_t0, knot_dt, knots, frame_times, exposure};
  spline_fwd_kernel<<<(C + 63) / 64, 64, 0, (cudaStream_t)stream>>>(a, viewmats);
  CHS_LAUNCH_CHECK();
  return CHS_OK;
}

extern "C" int chs_spline_bwd(int32_t kind, const float* knots, int32_t n_knots, double knot_t0, double knot_dt,
                              const float* frame_times, const float* exposure, int32_t n_frames, int32_t n_virtual,
                              const float* v_viewmats, float* v_knots, float* v_frame_times, float* v_exposure, void* workspace,
                              uint64_t workspace_bytes, void* stream) {
  int st = check_spline_args(kind, knots, n_knots, knot_dt, frame_times, exposure, n_frames, n_virtual);
  if (st) return st;
  CHS_REQUIRE(v_viewmats && v_knots && v_frame_times && v_exposure && workspace, "chs_spline_bwd: null pointer");
  const uint64_t n_acc = (uint64_t)n_knots * 7 + 2 * (uint64_t)n_frames;
  if (workspace_bytes < n_acc * sizeof(double)) {
    chs_set_error("chs_spline_bwd: workspace too small");
    return CHS_ERR_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t s = (cudaStream_t)stream;
  double* acc = (double*)workspace;
  CHS_CUDA(cudaMemsetAsync(acc, 0, n_acc * sizeof(double), s));
  const int C = n_frames * n_virtual;
  if (C > 0) {
    SplineArgs a{kind, n_knots, n_frames, n_virtual, knot_t0, knot_dt, knots, frame_times, exposure};
    const int n_in = (kind == CHS_SPLINE_LINEAR ? 2 : 4) * 7 + 1;
    spline_bwd_kernel<<<(C * n_in + 63) / 64, 64, 0, s>>>(a, v_viewmats, acc);
    CHS_LAUNCH_CHECK();
  }
  spline_finalize_kernel<<<((int)n_acc + 127) / 128, 128, 0, s>>>(acc, n_knots, n_frames, v_knots, v_frame_times, v_exposure);
  CHS_LAUNCH_CHECK();
  return CHS_OK;
}
