// chs_math.cuh — per-element arithmetic of the CasualHDRSplat image-formation path.
//
// Every formula the CUDA kernels evaluate lives here as a CHS_HD (host+device) function templated
// on the scalar type, so that tests/hostsim can compile the *same* code with g++ and check it
// against the float64 oracle without a GPU (tests only; the product never runs this on the host).
//
// What each block implements (the reference ships no code; the model is the one described in
// /root/reference/Readme.md:54 and /root/reference/assets/pipeline.png, completed by SURVEY.md
// Appendix A):
//   A.3  projection + EWA 2D covariance            -> chs_project_fwd / chs_project_bwd
//   A.4  tile bounds (bit-exact integer contract)  -> chs_tile_bounds
//   A.5  alpha of one (pixel, Gaussian) pair       -> chs_pair_power / chs_pair_alpha
//   A.6  blend backward of one pair                -> chs_pair_bwd
//   A.7  camera response curve F_theta             -> chs_crf_mlp_fwd / chs_crf_mlp_bwd, chs_crf_lut_fwd / chs_crf_lut_bwd
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define CHS_HD __host__ __device__ __forceinline__
#else
#define CHS_HD inline
#endif

#define CHS_TILE 16
#define CHS_ALPHA_MIN (1.0f / 255.0f)
#define CHS_ALPHA_MAX 0.999f
#define CHS_T_STOP 1e-4f
#define CHS_LOG2E 1.4426950408889634f
#define CHS_LN2 0.6931471805599453f
#define CHS_CRF_EPS 1e-5f

// model constants per scalar type (the double set exists so the host test harness can check the
// formulas against the float64 oracle to ~1e-12; kernels only ever instantiate <float>)
template <class T> struct ChsK;
template <> struct ChsK<float> {
  static constexpr float log2e = CHS_LOG2E, alpha_max = CHS_ALPHA_MAX, alpha_min = CHS_ALPHA_MIN, t_stop = CHS_T_STOP, crf_eps = CHS_CRF_EPS;
};
template <> struct ChsK<double> {
  static constexpr double log2e = 1.4426950408889634, alpha_max = 0.999, alpha_min = 1.0 / 255.0, t_stop = 1e-4, crf_eps = 1e-5;
};

// ---------------------------------------------------------------------------------------------
// MLP CRF in interval form (crf_bwd_interval_kernel).  With a scalar input z the network is piecewise linear: unit j switches
// at its breakpoint t_j = -b1_j / w1_j (+inf for w1_j = 0: it never switches).  Units are ranked by breakpoint inside a channel
// (ties by index); interval I in [0, Hd] = "I breakpoints lie below z".  "Unit j is on in interval I" is the integer test
// chs_crf_key_on(key_j, I) with key_j = chs_crf_unit_key(w1_j, b1_j, rank_j, Hd):
//   w1 > 0: on <=> z > t_j <=> I >= rank + 1;   w1 < 0: on <=> z < t_j <=> I <= rank;   w1 = 0: on <=> b1 > 0, in every interval.
// ---------------------------------------------------------------------------------------------
template <class T> CHS_HD T chs_crf_breakpoint(T w1, T b1) { return w1 != T(0) ? -b1 / w1 : T(INFINITY); }
template <class T> CHS_HD int chs_crf_unit_key(T w1, T b1, int rank, int hd) {
  int sgn = 1, off = -(rank + 1);
  if (w1 < T(0)) { sgn = -1; off = rank; }
  if (w1 == T(0)) { sgn = 1; off = b1 > T(0) ? 0 : -(hd + 1); }
  return off * 2 + (sgn < 0 ? 1 : 0);
}
CHS_HD bool chs_crf_key_on(int key, int I) { return ((key & 1) ? (key >> 1) - I : I + (key >> 1)) >= 0; }
// number of sorted breakpoints below z (branch-free bisection)
template <class T> CHS_HD int chs_crf_interval_of(const T* bp, int hd, T z) {
  int lo = 0, n = hd;
  while (n > 0) {
    const int half = n >> 1;
    const bool right = bp[lo + half] < z;
    lo = right ? lo + half + 1 : lo;
    n = right ? n - half - 1 : half;
  }
  return lo;
}

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
template <class T> CHS_HD T chs_min(T a, T b) { return a < b ? a : b; }
template <class T> CHS_HD T chs_max(T a, T b) { return a > b ? a : b; }

CHS_HD float chs_exp2_fast(float x) {
#if defined(__CUDA_ARCH__)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#else
  return exp2f(x);
#endif
}
CHS_HD double chs_exp2_fast(double x) { return exp2(x); }

CHS_HD float chs_log2_fast(float x) {
#if defined(__CUDA_ARCH__)
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#else
  return log2f(x);
#endif
}
CHS_HD double chs_log2_fast(double x) { return log2(x); }

CHS_HD float chs_fma(float a, float b, float c) { return fmaf(a, b, c); }
CHS_HD double chs_fma(double a, double b, double c) { return fma(a, b, c); }

CHS_HD float chs_rcp_fast(float x) {
#if defined(__CUDA_ARCH__)
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#else
  return 1.0f / x;
#endif
}
CHS_HD double chs_rcp_fast(double x) { return 1.0 / x; }

// ---------------------------------------------------------------------------------------------
// quaternion (wxyz, normalised inside) -> rotation; 3D covariance Sigma = R diag(s^2) R^T
// Sigma stored as 6 unique entries: [xx, xy, xz, yy, yz, zz]
// ---------------------------------------------------------------------------------------------
template <class T> CHS_HD void chs_quat_to_rotmat(const T q[4], T R[9]) {
  T n = T(1) / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  T w = q[0] * n, x = q[1] * n, y = q[2] * n, z = q[3] * n;
  R[0] = T(1) - T(2) * (y * y + z * z); R[1] = T(2) * (x * y - w * z); R[2] = T(2) * (x * z + w * y);
  R[3] = T(2) * (x * y + w * z); R[4] = T(1) - T(2) * (x * x + z * z); R[5] = T(2) * (y * z - w * x);
  R[6] = T(2) * (x * z - w * y); R[7] = T(2) * (y * z + w * x); R[8] = T(1) - T(2) * (x * x + y * y);
}

template <class T> CHS_HD void chs_cov3d(const T q[4], const T s[3], T S[6]) {
  T R[9];
  chs_quat_to_rotmat(q, R);
  T M[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) M[i * 3 + j] = R[i * 3 + j] * s[j];
  S[0] = M[0] * M[0] + M[1] * M[1] + M[2] * M[2];
  S[1] = M[0] * M[3] + M[1] * M[4] + M[2] * M[5];
  S[2] = M[0] * M[6] + M[1] * M[7] + M[2] * M[8];
  S[3] = M[3] * M[3] + M[4] * M[4] + M[5] * M[5];
  S[4] = M[3] * M[6] + M[4] * M[7] + M[5] * M[8];
  S[5] = M[6] * M[6] + M[7] * M[7] + M[8] * M[8];
}

// Backward of chs_cov3d. G[6] is the gradient w.r.t. the *full symmetric matrix* Sigma laid out as
// [Gxx, Gxy, Gxz, Gyy, Gyz, Gzz] where the off-diagonal numbers are the entries of the symmetric
// gradient matrix (i.e. dL = sum_ij Gfull_ij dSigma_ij).  Accumulates into v_q[4], v_s[3].
template <class T> CHS_HD void chs_cov3d_bwd(const T q[4], const T s[3], const T G[6], T v_q[4], T v_s[3]) {
  T qn2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  T inv = T(1) / sqrt(qn2);
  T w = q[0] * inv, x = q[1] * inv, y = q[2] * inv, z = q[3] * inv;
  T R[9];
  R[0] = T(1) - T(2) * (y * y + z * z); R[1] = T(2) * (x * y - w * z); R[2] = T(2) * (x * z + w * y);
  R[3] = T(2) * (x * y + w * z); R[4] = T(1) - T(2) * (x * x + z * z); R[5] = T(2) * (y * z - w * x);
  R[6] = T(2) * (x * z - w * y); R[7] = T(2) * (y * z + w * x); R[8] = T(1) - T(2) * (x * x + y * y);
  const T Gf[9] = {G[0], G[1], G[2], G[1], G[3], G[4], G[2], G[4], G[5]};
  // v_M = 2 Gf M, M = R diag(s)
  T vR[9];
  for (int j = 0; j < 3; ++j) {
    T vs = T(0);
    for (int i = 0; i < 3; ++i) {
      T gm = T(2) * (Gf[i * 3 + 0] * R[0 * 3 + j] + Gf[i * 3 + 1] * R[1 * 3 + j] + Gf[i * 3 + 2] * R[2 * 3 + j]) * s[j];
      vR[i * 3 + j] = gm * s[j];
      vs += R[i * 3 + j] * gm;
    }
    v_s[j] += vs;
  }
  T vw = T(2) * (-z * vR[1] + y * vR[2] + z * vR[3] - x * vR[5] - y * vR[6] + x * vR[7]);
  T vx = T(2) * (y * vR[1] + z * vR[2] + y * vR[3] - T(2) * x * vR[4] - w * vR[5] + z * vR[6] + w * vR[7] - T(2) * x * vR[8]);
  T vy = T(2) * (-T(2) * y * vR[0] + x * vR[1] + w * vR[2] + x * vR[3] + z * vR[5] - w * vR[6] + z * vR[7] - T(2) * y * vR[8]);
  T vz = T(2) * (-T(2) * z * vR[0] - w * vR[1] + x * vR[2] + w * vR[3] - T(2) * z * vR[4] + y * vR[5] + x * vR[6] + y * vR[7]);
  // through the normalisation q_hat = q / |q|
  T dot = vw * w + vx * x + vy * y + vz * z;
  v_q[0] += (vw - dot * w) * inv;
  v_q[1] += (vx - dot * x) * inv;
  v_q[2] += (vy - dot * y) * inv;
  v_q[3] += (vz - dot * z) * inv;
}

// ---------------------------------------------------------------------------------------------
// camera: world->camera rotation R (row major 3x3), translation t, intrinsics
// ---------------------------------------------------------------------------------------------
template <class T> struct ChsCam {
  T R[9];
  T t[3];
  T fx, fy, cx, cy;
};

template <class T> struct ChsProj {
  T mx, my;      // mean2d
  T depth;       // z in camera frame
  T ca, cb, cc;  // conic = inverse 2D covariance (A, B, C)
  T sxx, syy;    // diagonal of the 2D covariance (after the eps2d blur)
  int radius;    // ceil(3 sqrt(lambda_max)), 0 when culled
};

// A.3 forward for one (camera, Gaussian). Returns radius (0 = culled).
template <class T>
CHS_HD int chs_project_fwd(const T mu[3], const T S[6], const ChsCam<T>& cam, T width, T height, T near_plane,
                           T far_plane, T eps2d, ChsProj<T>& out) {
  const T* R = cam.R;
  T x = R[0] * mu[0] + R[1] * mu[1] + R[2] * mu[2] + cam.t[0];
  T y = R[3] * mu[0] + R[4] * mu[1] + R[5] * mu[2] + cam.t[1];
  T z = R[6] * mu[0] + R[7] * mu[1] + R[8] * mu[2] + cam.t[2];
  out.mx = out.my = out.ca = out.cb = out.cc = out.sxx = out.syy = T(0);
  out.depth = z;
  out.radius = 0;
  if (!(z >= near_plane) || !(z <= far_plane)) return 0;
  // Sigma_c = R Sigma R^T
  T RS[9];
  for (int i = 0; i < 3; ++i) {
    RS[i * 3 + 0] = R[i * 3 + 0] * S[0] + R[i * 3 + 1] * S[1] + R[i * 3 + 2] * S[2];
    RS[i * 3 + 1] = R[i * 3 + 0] * S[1] + R[i * 3 + 1] * S[3] + R[i * 3 + 2] * S[4];
    RS[i * 3 + 2] = R[i * 3 + 0] * S[2] + R[i * 3 + 1] * S[4] + R[i * 3 + 2] * S[5];
  }
  T c00 = RS[0] * R[0] + RS[1] * R[1] + RS[2] * R[2];
  T c01 = RS[0] * R[3] + RS[1] * R[4] + RS[2] * R[5];
  T c02 = RS[0] * R[6] + RS[1] * R[7] + RS[2] * R[8];
  T c11 = RS[3] * R[3] + RS[4] * R[4] + RS[5] * R[5];
  T c12 = RS[3] * R[6] + RS[4] * R[7] + RS[5] * R[8];
  T c22 = RS[6] * R[6] + RS[7] * R[7] + RS[8] * R[8];
  T rz = T(1) / z;
  T lim_xp = (width - cam.cx) / cam.fx + T(0.3) * (width / (T(2) * cam.fx));
  T lim_xm = cam.cx / cam.fx + T(0.3) * (width / (T(2) * cam.fx));
  T lim_yp = (height - cam.cy) / cam.fy + T(0.3) * (height / (T(2) * cam.fy));
  T lim_ym = cam.cy / cam.fy + T(0.3) * (height / (T(2) * cam.fy));
  T tx = z * chs_min(lim_xp, chs_max(-lim_xm, x * rz));
  T ty = z * chs_min(lim_yp, chs_max(-lim_ym, y * rz));
  T j00 = cam.fx * rz, j02 = -cam.fx * tx * rz * rz;
  T j11 = cam.fy * rz, j12 = -cam.fy * ty * rz * rz;
  T a = j00 * j00 * c00 + T(2) * j00 * j02 * c02 + j02 * j02 * c22 + eps2d;
  T b = j00 * j11 * c01 + j00 * j12 * c02 + j02 * j11 * c12 + j02 * j12 * c22;
  T c = j11 * j11 * c11 + T(2) * j11 * j12 * c12 + j12 * j12 * c22 + eps2d;
  T det = a * c - b * b;
  if (!(det > T(0))) return 0;
  T rdet = T(1) / det;
  T mx = cam.fx * x * rz + cam.cx;
  T my = cam.fy * y * rz + cam.cy;
  T m = T(0.5) * (a + c);
  T lam = m + sqrt(chs_max(T(0.01), m * m - det));
  T rad = ceil(T(3) * sqrt(lam));
  if (mx + rad <= T(0) || mx - rad >= width || my + rad <= T(0) || my - rad >= height) return 0;
  out.mx = mx;
  out.my = my;
  out.ca = c * rdet;
  out.cb = -b * rdet;
  out.cc = a * rdet;
  out.sxx = a;
  out.syy = c;
  out.radius = (int)rad;
  return out.radius;
}

// Opacity-aware per-axis bounds (chs_config.tight_bounds).  alpha >= 1/255 only inside the ellipse
// sigma(d) <= ln(255 o); its axis-aligned bounding box has half extents sqrt(tau Sxx), sqrt(tau Syy)
// with tau = 2 (ln(255 o) + margin).  Returns rx | ry << 16 with rx = min(radius, ceil(sqrt(tau Sxx)))
// (likewise ry), each in [1, 65535], or 0 when tau <= 0 (the Gaussian can never reach 1/255).  The
// tight rectangle only drops tiles whose every pixel fails the alpha test, so images are unchanged.
// 65535 means "unbounded along this axis" (a splat wider than 65534 pixels): chs_tile_bounds_of then
// takes every tile column / row, which is a superset of what the classic square would give.
#define CHS_TIGHT_MARGIN 2e-3
template <class T> CHS_HD int chs_tight_radii(T sxx, T syy, T opacity, int radius) {
  const T tau = T(2) * (log(T(255) * opacity) + T(CHS_TIGHT_MARGIN));
  if (!(tau > T(0))) return 0;
  const T cap = T(radius < 65535 ? radius : 65535);
  const T rx = chs_max(T(1), chs_min(cap, ceil(sqrt(tau * sxx))));
  const T ry = chs_max(T(1), chs_min(cap, ceil(sqrt(tau * syy))));
  return (int)((uint32_t)rx | ((uint32_t)ry << 16));  // an unsigned pair: consumers gate on != 0, never on > 0
}

// A.3 backward for one (camera, Gaussian): given v_mean2d and v_conic, accumulate
//   v_mu[3]       gradient w.r.t. the world mean
//   G[6]          gradient w.r.t. the full symmetric 3D covariance (see chs_cov3d_bwd)
//   v_R[9], v_t[3] gradient w.r.t. the camera's world->camera rotation / translation
// The forward quantities are recomputed (cheaper than storing them per (c, g)).
template <class T>
CHS_HD void chs_project_bwd(const T mu[3], const T S[6], const ChsCam<T>& cam, T width, T height, T eps2d,
                            T v_mx, T v_my, T v_ca, T v_cb, T v_cc, T v_mu[3], T G[6], T v_R[9], T v_t[3]) {
  const T* R = cam.R;
  T x = R[0] * mu[0] + R[1] * mu[1] + R[2] * mu[2] + cam.t[0];
  T y = R[3] * mu[0] + R[4] * mu[1] + R[5] * mu[2] + cam.t[1];
  T z = R[6] * mu[0] + R[7] * mu[1] + R[8] * mu[2] + cam.t[2];
  T RS[9];
  for (int i = 0; i < 3; ++i) {
    RS[i * 3 + 0] = R[i * 3 + 0] * S[0] + R[i * 3 + 1] * S[1] + R[i * 3 + 2] * S[2];
    RS[i * 3 + 1] = R[i * 3 + 0] * S[1] + R[i * 3 + 1] * S[3] + R[i * 3 + 2] * S[4];
    RS[i * 3 + 2] = R[i * 3 + 0] * S[2] + R[i * 3 + 1] * S[4] + R[i * 3 + 2] * S[5];
  }
  T Sc[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Sc[i * 3 + j] = RS[i * 3 + 0] * R[j * 3 + 0] + RS[i * 3 + 1] * R[j * 3 + 1] + RS[i * 3 + 2] * R[j * 3 + 2];
  T rz = T(1) / z, rz2 = rz * rz;
  T lim_xp = (width - cam.cx) / cam.fx + T(0.3) * (width / (T(2) * cam.fx));
  T lim_xm = cam.cx / cam.fx + T(0.3) * (width / (T(2) * cam.fx));
  T lim_yp = (height - cam.cy) / cam.fy + T(0.3) * (height / (T(2) * cam.fy));
  T lim_ym = cam.cy / cam.fy + T(0.3) * (height / (T(2) * cam.fy));
  T xz = x * rz, yz = y * rz;
  T cxz = chs_min(lim_xp, chs_max(-lim_xm, xz));
  T cyz = chs_min(lim_yp, chs_max(-lim_ym, yz));
  bool x_free = (cxz == xz), y_free = (cyz == yz);
  T tx = z * cxz, ty = z * cyz;
  T j00 = cam.fx * rz, j02 = -cam.fx * tx * rz2;
  T j11 = cam.fy * rz, j12 = -cam.fy * ty * rz2;
  T a = j00 * j00 * Sc[0] + T(2) * j00 * j02 * Sc[2] + j02 * j02 * Sc[8] + eps2d;
  T b = j00 * j11 * Sc[1] + j00 * j12 * Sc[2] + j02 * j11 * Sc[5] + j02 * j12 * Sc[8];
  T c = j11 * j11 * Sc[4] + T(2) * j11 * j12 * Sc[5] + j12 * j12 * Sc[8] + eps2d;
  T rdet = T(1) / (a * c - b * b);
  T A = c * rdet, B = -b * rdet, C = a * rdet;
  // conic -> 2D covariance: Gm = -X V X with V = [[vA, vB/2], [vB/2, vC]]
  T p = v_ca, q = T(0.5) * v_cb, r = v_cc;
  T g00 = -(A * A * p + T(2) * A * B * q + B * B * r);
  T g01 = -(A * B * p + (B * B + A * C) * q + B * C * r);
  T g11 = -(B * B * p + T(2) * B * C * q + C * C * r);
  // v_Sigma_c = J^T Gm J  (J = [[j00, 0, j02], [0, j11, j12]])
  T Gc[9];
  Gc[0] = j00 * j00 * g00;
  Gc[1] = j00 * j11 * g01;
  Gc[2] = j00 * (j02 * g00 + j12 * g01);
  Gc[4] = j11 * j11 * g11;
  Gc[5] = j11 * (j02 * g01 + j12 * g11);
  Gc[8] = j02 * j02 * g00 + T(2) * j02 * j12 * g01 + j12 * j12 * g11;
  Gc[3] = Gc[1]; Gc[6] = Gc[2]; Gc[7] = Gc[5];
  // v_J = 2 Gm J Sigma_c
  T JS0[3], JS1[3];  // rows of J Sigma_c
  for (int k = 0; k < 3; ++k) {
    JS0[k] = j00 * Sc[0 * 3 + k] + j02 * Sc[2 * 3 + k];
    JS1[k] = j11 * Sc[1 * 3 + k] + j12 * Sc[2 * 3 + k];
  }
  T vj00 = T(2) * (g00 * JS0[0] + g01 * JS1[0]);
  T vj02 = T(2) * (g00 * JS0[2] + g01 * JS1[2]);
  T vj11 = T(2) * (g01 * JS0[1] + g11 * JS1[1]);
  T vj12 = T(2) * (g01 * JS0[2] + g11 * JS1[2]);
  T v_x = cam.fx * rz * v_mx;
  T v_y = cam.fy * rz * v_my;
  T v_z = -(cam.fx * x * v_mx + cam.fy * y * v_my) * rz2;
  v_z += -vj00 * cam.fx * rz2 - vj11 * cam.fy * rz2 + T(2) * (vj02 * cam.fx * tx + vj12 * cam.fy * ty) * rz2 * rz;
  T v_tx = -vj02 * cam.fx * rz2, v_ty = -vj12 * cam.fy * rz2;
  if (x_free) v_x += v_tx; else v_z += v_tx * cxz;
  if (y_free) v_y += v_ty; else v_z += v_ty * cyz;
  const T vp[3] = {v_x, v_y, v_z};
  for (int k = 0; k < 3; ++k) v_mu[k] += R[0 * 3 + k] * vp[0] + R[1 * 3 + k] * vp[1] + R[2 * 3 + k] * vp[2];
  for (int i = 0; i < 3; ++i) {
    v_t[i] += vp[i];
    for (int k = 0; k < 3; ++k) v_R[i * 3 + k] += vp[i] * mu[k];
  }
  // v_R += 2 Gc R Sigma = 2 Gc RS ;  v_Sigma += R^T Gc R
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 3; ++k)
      v_R[i * 3 + k] += T(2) * (Gc[i * 3 + 0] * RS[0 * 3 + k] + Gc[i * 3 + 1] * RS[1 * 3 + k] + Gc[i * 3 + 2] * RS[2 * 3 + k]);
  T GR[9];  // Gc R
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 3; ++k) GR[i * 3 + k] = Gc[i * 3 + 0] * R[0 * 3 + k] + Gc[i * 3 + 1] * R[1 * 3 + k] + Gc[i * 3 + 2] * R[2 * 3 + k];
  // (R^T Gc R)_{ab} = sum_i R[i][a] GR[i][b]
  G[0] += R[0] * GR[0] + R[3] * GR[3] + R[6] * GR[6];
  G[1] += R[0] * GR[1] + R[3] * GR[4] + R[6] * GR[7];
  G[2] += R[0] * GR[2] + R[3] * GR[5] + R[6] * GR[8];
  G[3] += R[1] * GR[1] + R[4] * GR[4] + R[7] * GR[7];
  G[4] += R[1] * GR[2] + R[4] * GR[5] + R[7] * GR[8];
  G[5] += R[2] * GR[2] + R[5] * GR[5] + R[8] * GR[8];
}

// ---------------------------------------------------------------------------------------------
// A.4 tile bounds: integer function of fp32 (mean2d, radius). fp32, round-to-nearest, no FMA
// contraction: the two scalings by 1/16 are exact, then one rounding each for lo and hi.
// ---------------------------------------------------------------------------------------------
CHS_HD float chs_sub_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  volatile float r = a - b;
  return r;
#endif
}
CHS_HD float chs_add_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  volatile float r = a + b;
  return r;
#endif
}
CHS_HD float chs_mul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  volatile float r = a * b;
  return r;
#endif
}

struct ChsTileRect {
  int x0, y0, x1, y1;  // [x0, x1) x [y0, y1) in tile units
};

CHS_HD ChsTileRect chs_tile_bounds2(float mx, float my, int radius_x, int radius_y, int tile_w, int tile_h) {
  const float inv = 1.0f / CHS_TILE;
  float trx = chs_mul_rn((float)radius_x, inv), try_ = chs_mul_rn((float)radius_y, inv);
  float tx = chs_mul_rn(mx, inv), ty = chs_mul_rn(my, inv);
  ChsTileRect r;
  r.x0 = (int)fminf(fmaxf(floorf(chs_sub_rn(tx, trx)), 0.0f), (float)tile_w);
  r.x1 = (int)fminf(fmaxf(ceilf(chs_add_rn(tx, trx)), 0.0f), (float)tile_w);
  r.y0 = (int)fminf(fmaxf(floorf(chs_sub_rn(ty, try_)), 0.0f), (float)tile_h);
  r.y1 = (int)fminf(fmaxf(ceilf(chs_add_rn(ty, try_)), 0.0f), (float)tile_h);
  return r;
}
CHS_HD ChsTileRect chs_tile_bounds(float mx, float my, int radius, int tile_w, int tile_h) {
  return chs_tile_bounds2(mx, my, radius, radius, tile_w, tile_h);
}
// `radii` entry -> tile rectangle: a plain radius, or the packed per-axis radii of chs_config.tight_bounds
CHS_HD ChsTileRect chs_tile_bounds_of(float mx, float my, int radii_entry, int tight, int tile_w, int tile_h) {
  if (!tight) return chs_tile_bounds2(mx, my, radii_entry, radii_entry, tile_w, tile_h);
  const int rx = radii_entry & 0xffff, ry = (radii_entry >> 16) & 0xffff;
  ChsTileRect r = chs_tile_bounds2(mx, my, rx, ry, tile_w, tile_h);
  if (rx == 0xffff) { r.x0 = 0; r.x1 = tile_w; }  // unbounded axis (see chs_tight_radii)
  if (ry == 0xffff) { r.y0 = 0; r.y1 = tile_h; }
  return r;
}

// ---------------------------------------------------------------------------------------------
// A.5 / A.6 blend arithmetic of one (pixel, Gaussian) pair.
// A staged Gaussian carries its conic in completed-square form, pre-scaled for exp2:
//   log2(alpha) = qa u^2 + kc dy^2 + lo,   u = dx + r dy,   (dx, dy) = mean2d - pixel
//   qa = -0.5 log2(e) A (< 0),  r = B / A,  kc = -0.5 log2(e) (C - B^2 / A) (< 0),  lo = log2(opacity)
// which equals -log2(e) sigma + log2(o) with sigma = 0.5 (A dx^2 + C dy^2) + B dx dy.  Both squared
// terms have non-positive coefficients for every positive-definite conic, so the exponent can never
// round to a positive number and A.5's "skip if sigma < 0" needs no code.
// ---------------------------------------------------------------------------------------------
template <class T> struct ChsSplat {
  T mx, my, qa, r;
  T kc, lo, rbc;  // rbc = -B/C (culling only); -B/A = -r
  T cr, cg, cb, inv_opac;
};

template <class T>
CHS_HD void chs_make_splat(T mx, T my, T ca, T cb, T cc, T opac, T r, T g, T b, ChsSplat<T>& s) {
  // staged once per (tile, Gaussian) and shared by forward and backward: approximate log2 / reciprocal
  // (2^-22 relative) are ample; rbc only steers the conservative culling bound
  const T rca = ca > T(0) ? chs_rcp_fast(ca) : T(0);
  s.mx = mx; s.my = my;
  s.qa = T(-0.5) * ChsK<T>::log2e * ca;
  s.r = cb * rca;
  s.kc = T(-0.5) * ChsK<T>::log2e * (cc - cb * s.r);
  s.kc = chs_min(s.kc, T(0));
  s.lo = chs_log2_fast(opac);
  s.rbc = cc > T(0) ? -cb * chs_rcp_fast(cc) : T(0);
  s.cr = r; s.cg = g; s.cb = b;
  s.inv_opac = chs_rcp_fast(opac);
}

// log2(alpha) before the 0.999 clamp; also returns dx, dy and u = dx + r dy
template <class T> CHS_HD T chs_pair_power(const ChsSplat<T>& s, T px, T py, T& dx, T& dy, T& u) {
  dx = s.mx - px;
  dy = s.my - py;
  u = chs_fma(s.r, dy, dx);
  return chs_fma(s.qa * u, u, chs_fma(s.kc * dy, dy, s.lo));
}

// Sub-tile culling.  Upper bound of log2(alpha) over the axis-aligned rectangle of pixel centres
// [x0,x1] x [y0,y1]: the exponent is a concave quadratic of d = p - mean, so its maximum over the
// rectangle is attained either at the mean (inside the rectangle) or on one of the two edges that
// face the mean, at the 1-D maximiser clamped to the edge.  (With x or y at a far edge the KKT
// conditions would need B^2 >= AC, impossible for a positive-definite conic.)  A warp skips a
// staged Gaussian when this bound is below log2(1/255): exactly the pairs A.5 skips anyway.
template <class T> CHS_HD T chs_block_max_power(const ChsSplat<T>& s, T x0, T x1, T y0, T y1) {
  T ax0 = x0 - s.mx, ax1 = x1 - s.mx, ay0 = y0 - s.my, ay1 = y1 - s.my;
  T cx = chs_min(chs_max(T(0), ax0), ax1);
  T cy = chs_min(chs_max(T(0), ay0), ay1);
  T dy1 = chs_min(chs_max(s.rbc * cx, ay0), ay1);   // maximiser along the vertical line x = cx
  T dx2 = chs_min(chs_max(-s.r * cy, ax0), ax1);    // maximiser along the horizontal line y = cy
  T u1 = cx + s.r * dy1, u2 = dx2 + s.r * cy;
  T p1 = s.qa * u1 * u1 + s.kc * dy1 * dy1;
  T p2 = s.qa * u2 * u2 + s.kc * cy * cy;
  return chs_max(p1, p2) + s.lo;
}

// Backward of one pair, walking back to front (A.6). On entry T = transmittance *after* this
// Gaussian, buf = colour accumulated behind it. vh = v_H (3), va_t = T_final * (v_alpha - bg.v_H).
// Outputs the 9 per-Gaussian partials for this pixel in g[9] =
//   [v_mx, v_my, v_A, v_B, v_C, v_opacity, v_r, v_g, v_b]   (zero where the clamp is active).
// Branch-free: a pair that does not contribute is passed with alpha_unclamped = alpha = 0, which
// leaves T and buf untouched and yields g = 0.
template <class T>
CHS_HD void chs_pair_bwd(const ChsSplat<T>& s, T dx, T dy, T u, T alpha_unclamped, T alpha, T& Tr, T buf[3],
                         const T vh[3], T va_t, T g[9]) {
  T ra = chs_rcp_fast(T(1) - alpha);
  Tr = Tr * ra;  // transmittance before this Gaussian
  T f = alpha * Tr;
  g[6] = f * vh[0];
  g[7] = f * vh[1];
  g[8] = f * vh[2];
  T v_alpha = (s.cr * Tr - buf[0] * ra) * vh[0] + (s.cg * Tr - buf[1] * ra) * vh[1] + (s.cb * Tr - buf[2] * ra) * vh[2] + va_t * ra;
  buf[0] += s.cr * f;
  buf[1] += s.cg * f;
  buf[2] += s.cb * f;
  // no gradient through the 0.999 clamp
  T v_sigma = alpha_unclamped <= ChsK<T>::alpha_max ? -alpha_unclamped * v_alpha : T(0);
  // d sigma / d mean = (A dx + B dy, B dx + C dy) = (A u, r A u - (2 / log2e) kc dy),  A = -(2 / log2e) qa
  const T k2 = T(-2) / ChsK<T>::log2e;
  T vk = v_sigma * k2;
  g[0] = vk * s.qa * u;
  g[1] = s.r * g[0] + vk * s.kc * dy;
  T hx = T(0.5) * v_sigma * dx;
  g[2] = hx * dx;
  g[3] = T(2) * hx * dy;
  g[4] = T(0.5) * v_sigma * dy * dy;
  g[5] = -v_sigma * s.inv_opac;
}

// Tabled form of the pair backward (blend_bwd2_kernel in chs_blend.cu evaluates exactly these three steps, packed two
// pixels per lane).  All nine partials of a (pixel, Gaussian) pair are multiples of two scalars, vs = dL/dsigma and
// f = alpha * T_before:
//   phase A  chs_pair_bwd_scalars: the state update of chs_pair_bwd (same operations, same order), returning (vs, f);
//   phase B  chs_pair_moments:     m += (vs u, vs dy, vs dx^2, vs dx dy, vs dy^2, vs, f v_r, f v_g, f v_b), summed over pixels;
//            chs_moments_to_grads: the Gaussian's constants applied once to the sums.
template <class T>
CHS_HD void chs_pair_bwd_scalars(const ChsSplat<T>& s, T alpha_unclamped, T alpha, T& Tr, T buf[3], const T vh[3], T va_t, T& vs,
                                 T& f) {
  T ra = chs_rcp_fast(T(1) - alpha);
  Tr = Tr * ra;  // transmittance before this Gaussian
  f = alpha * Tr;
  T v_alpha = (s.cr * Tr - buf[0] * ra) * vh[0] + (s.cg * Tr - buf[1] * ra) * vh[1] + (s.cb * Tr - buf[2] * ra) * vh[2] + va_t * ra;
  buf[0] += s.cr * f;
  buf[1] += s.cg * f;
  buf[2] += s.cb * f;
  vs = alpha_unclamped <= ChsK<T>::alpha_max ? -alpha_unclamped * v_alpha : T(0);  // no gradient through the 0.999 clamp
}
template <class T> CHS_HD void chs_pair_moments(T vs, T f, T dx, T dy, T u, const T vh[3], T m[9]) {
  m[0] += vs * u;
  m[1] += vs * dy;
  m[2] += vs * dx * dx;
  m[3] += vs * dx * dy;
  m[4] += vs * dy * dy;
  m[5] += vs;
  m[6] += f * vh[0];
  m[7] += f * vh[1];
  m[8] += f * vh[2];
}
// g = [v_mx, v_my, v_A, v_B, v_C, v_opacity, v_r, v_g, v_b] from the moment sums (uses s.qa, s.r, s.kc, s.inv_opac)
template <class T> CHS_HD void chs_moments_to_grads(const ChsSplat<T>& s, const T m[9], T g[9]) {
  const T k2 = T(-2) / ChsK<T>::log2e;
  g[0] = k2 * s.qa * m[0];
  g[1] = s.r * g[0] + k2 * s.kc * m[1];
  g[2] = T(0.5) * m[2];
  g[3] = m[3];
  g[4] = T(0.5) * m[4];
  g[5] = -m[5] * s.inv_opac;
  g[6] = m[6];
  g[7] = m[7];
  g[8] = m[8];
}

// Division-free colour state (round 2, blend_bwd3_kernel).  Dotting the colours with the pixel's upstream gradient first turns
// the three "colour behind" accumulators into ONE scalar, and normalising it by the transmittance removes every division:
//   s_i = c_i . v_H,   R_i = (sum_{j>i} alpha_j T_j s_j + T_final (bg . v_H - v_alpha)) / T_{i+1}   ("what lies behind i", per unit
//   of transmittance),   dL/dalpha_i = T_i (s_i - R_i),   R_{i-1} = alpha_i s_i + (1 - alpha_i) R_i = R_i - alpha_i (R_i - s_i),
// starting behind the last accumulated Gaussian with R = bg . v_H - v_alpha.  Only T still needs its reciprocal.  The routine
// returns the NEGATED scalars nvs = -dL/dsigma, nf = -alpha T (the kernel's packed arithmetic produces these signs for free;
// chs_moments_to_grads_neg undoes them once per Gaussian).
template <class T>
CHS_HD void chs_pair_bwd_scalars_r(const ChsSplat<T>& s, T alpha_unclamped, T alpha, T& Tr, T& R, const T vh[3], T& nvs, T& nf) {
  T ra = chs_rcp_fast(T(1) - alpha);
  Tr = Tr * ra;  // transmittance before this Gaussian
  nf = -alpha * Tr;
  T sv = s.cb * vh[2] + (s.cg * vh[1] + s.cr * vh[0]);
  T e = R - sv;
  R = R - alpha * e;
  nvs = alpha_unclamped <= ChsK<T>::alpha_max ? nf * e : T(0);  // no gradient through the 0.999 clamp
}
// as chs_moments_to_grads for moment sums of the negated scalars (nvs, nf)
template <class T> CHS_HD void chs_moments_to_grads_neg(const ChsSplat<T>& s, const T m[9], T g[9]) {
  const T k2 = T(2) / ChsK<T>::log2e;
  g[0] = k2 * s.qa * m[0];
  g[1] = s.r * g[0] + k2 * s.kc * m[1];
  g[2] = T(-0.5) * m[2];
  g[3] = -m[3];
  g[4] = T(-0.5) * m[4];
  g[5] = m[5] * s.inv_opac;
  g[6] = -m[6];
  g[7] = -m[7];
  g[8] = -m[8];
}

// Tensor-core phase B (blend_bwd4_kernel): the sums over a warp's 8x8 pixel block are taken against pixel monomials relative to
// the block ORIGIN (x, y in 0 .. 7, exactly representable in fp16), S = sum nvs * (1, x, y, x^2, x y, y^2), as one small matrix
// product.  This shifts them to the mean-relative sums chs_pair_moments accumulates, with (X, Y) = mean2d - centre of the block's
// pixel (0, 0), i.e. dx = X - x, dy = Y - y, u = dx + r dy:
//   m = (sum nvs u, sum nvs dy, sum nvs dx^2, sum nvs dx dy, sum nvs dy^2, sum nvs)
template <class T> CHS_HD void chs_shift_moments(T S1, T Sx, T Sy, T Sxx, T Sxy, T Syy, T X, T Y, T r, T m[6]) {
  const T sdx = X * S1 - Sx, sdy = Y * S1 - Sy;
  m[0] = sdx + r * sdy;
  m[1] = sdy;
  m[2] = X * (sdx - Sx) + Sxx;         // X^2 S1 - 2 X Sx + Sxx
  m[3] = X * sdy - Y * Sx + Sxy;       // X Y S1 - X Sy - Y Sx + Sxy
  m[4] = Y * (sdy - Sy) + Syy;         // Y^2 S1 - 2 Y Sy + Syy
  m[5] = S1;
}

// ---------------------------------------------------------------------------------------------
// A.7 camera response curve, MLP kind [D6]: per channel z = ln(X + 1e-5), h = relu(w1 z + b1),
// y = sigmoid(w2 . h + b2).  params = [w1 (Hd) | b1 (Hd) | w2 (Hd) | b2].
// ---------------------------------------------------------------------------------------------
// Exposure X = dt * H is clamped to >= 0 before the logarithm (zero gradient w.r.t. X below the clamp): an optimiser stepping
// on unclamped per-Gaussian colours can drive H slightly negative, and ln(X + 1e-5) must not turn that into NaN.
template <class T> CHS_HD T chs_crf_mlp_fwd(T X, const T* p, int hd) {
  T z = log(chs_max(X, T(0)) + ChsK<T>::crf_eps);
  T acc = p[3 * hd];
  for (int j = 0; j < hd; ++j) acc += p[2 * hd + j] * chs_max(T(0), p[j] * z + p[hd + j]);
  return T(1) / (T(1) + exp(-acc));
}

// Returns dy/dX (for v_X = v_y * dy/dX) and, if v_p != nullptr, accumulates v_y * dy/dparams into v_p.
template <class T> CHS_HD T chs_crf_mlp_bwd(T X, const T* p, int hd, T v_y, T* v_p) {
  T xe = chs_max(X, T(0)) + ChsK<T>::crf_eps;
  T z = log(xe);
  T acc = p[3 * hd];
  T dz = T(0);
  for (int j = 0; j < hd; ++j) {
    T pre = p[j] * z + p[hd + j];
    if (pre > T(0)) {
      acc += p[2 * hd + j] * pre;
      dz += p[2 * hd + j] * p[j];
    }
  }
  T y = T(1) / (T(1) + exp(-acc));
  T gy = v_y * y * (T(1) - y);  // dL/dacc
  if (v_p) {
    for (int j = 0; j < hd; ++j) {
      T pre = p[j] * z + p[hd + j];
      if (pre > T(0)) {
        v_p[2 * hd + j] += gy * pre;
        T dh = gy * p[2 * hd + j];
        v_p[j] += dh * z;
        v_p[hd + j] += dh;
      }
    }
    v_p[3 * hd] += gy;
  }
  return X >= T(0) ? y * (T(1) - y) * dz / xe : T(0);
}

// ---------------------------------------------------------------------------------------------
// A.7 camera response curve, LUT kind: per channel a piecewise-linear
// curve over log exposure, the classical Debevec-style response table.
//   z = ln(X + 1e-5),  u = clamp((z - z_min) / (z_max - z_min), 0, 1) (L - 1),
//   i = min(floor(u), L - 2),  f = u - i,  y = v_i + f (v_{i+1} - v_i)
// params = [z_min | z_max | v_0 .. v_{L-1}] (L >= 2 knots).  The range (z_min, z_max) is a fixed
// calibration of the table and receives no gradient; outside it the curve is constant.
// ---------------------------------------------------------------------------------------------
template <class T> struct ChsLutPos {
  int i;      // left knot
  T f;        // position inside the segment, in [0, 1]
  T du_dz;    // (L - 1) / (z_max - z_min), or 0 where the clamp is active
  T xe;       // X + eps
};

template <class T> CHS_HD ChsLutPos<T> chs_crf_lut_pos(T X, const T* p, int L) {
  ChsLutPos<T> r;
  r.xe = chs_max(X, T(0)) + ChsK<T>::crf_eps;  // X clamped to >= 0, zero gradient below (see chs_crf_mlp_fwd)
  const T z = log(r.xe);
  const T scale = T(L - 1) / (p[1] - p[0]);
  T u = (z - p[0]) * scale;
  r.du_dz = (u > T(0) && u < T(L - 1) && X >= T(0)) ? scale : T(0);
  u = chs_min(chs_max(u, T(0)), T(L - 1));
  int i = (int)u;
  if (i > L - 2) i = L - 2;
  r.i = i;
  r.f = u - T(i);
  return r;
}

template <class T> CHS_HD T chs_crf_lut_fwd(T X, const T* p, int L) {
  const ChsLutPos<T> q = chs_crf_lut_pos(X, p, L);
  const T a = p[2 + q.i], b = p[3 + q.i];
  return a + q.f * (b - a);
}

// Returns dy/dX and, if v_p != nullptr, accumulates v_y * dy/dparams into v_p (knot values only).
template <class T> CHS_HD T chs_crf_lut_bwd(T X, const T* p, int L, T v_y, T* v_p) {
  const ChsLutPos<T> q = chs_crf_lut_pos(X, p, L);
  if (v_p) {
    v_p[2 + q.i] += v_y * (T(1) - q.f);
    v_p[3 + q.i] += v_y * q.f;
  }
  return (p[3 + q.i] - p[2 + q.i]) * q.du_dz / q.xe;
}

// number of CRF parameters per channel for a (kind, size) pair; kinds as in chs.h
CHS_HD int chs_crf_stride(int crf_kind, int crf_size) { return crf_kind == 1 ? 3 * crf_size + 1 : crf_kind == 2 ? crf_size + 2 : 0; }

// F_theta of either learned kind
template <class T> CHS_HD T chs_crf_fwd(int crf_kind, T X, const T* p, int crf_size) {
  return crf_kind == 1 ? chs_crf_mlp_fwd(X, p, crf_size) : chs_crf_lut_fwd(X, p, crf_size);
}
