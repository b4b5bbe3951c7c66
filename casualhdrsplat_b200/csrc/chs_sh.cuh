// chs_sh.cuh — real spherical harmonics (degree 0..3) for view-dependent HDR colour
// ("SH / view-dependent HDR colour evaluation fused into K1").
//   colour_ch = max(0, 0.5 + sum_k sh[k][ch] * Y_k(dir)),   dir = (mean - campos) / |mean - campos|
// Basis constants and ordering follow the convention every 3DGS code base uses (degree l occupies
// coefficients l^2 .. (l+1)^2 - 1).  Host+device templates: tests/hostsim checks value and gradients
// against the float64 oracle (oracle/sh.py).
#pragma once
#include "chs_math.cuh"

#define CHS_SH_MAX_COEFFS 16

// Y[k], k < (deg+1)^2, for a unit direction (x, y, z)
template <class T> CHS_HD void chs_sh_basis(int deg, T x, T y, T z, T Y[CHS_SH_MAX_COEFFS]) {
  Y[0] = T(0.28209479177387814);
  if (deg < 1) return;
  const T c1 = T(0.4886025119029199);
  Y[1] = -c1 * y; Y[2] = c1 * z; Y[3] = -c1 * x;
  if (deg < 2) return;
  const T xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
  Y[4] = T(1.0925484305920792) * xy;
  Y[5] = T(-1.0925484305920792) * yz;
  Y[6] = T(0.31539156525252005) * (T(2) * zz - xx - yy);
  Y[7] = T(-1.0925484305920792) * xz;
  Y[8] = T(0.5462742152960396) * (xx - yy);
  if (deg < 3) return;
  Y[9] = T(-0.5900435899266435) * y * (T(3) * xx - yy);
  Y[10] = T(2.890611442640554) * xy * z;
  Y[11] = T(-0.4570457994644658) * y * (T(4) * zz - xx - yy);
  Y[12] = T(0.3731763325901154) * z * (T(2) * zz - T(3) * xx - T(3) * yy);
  Y[13] = T(-0.4570457994644658) * x * (T(4) * zz - xx - yy);
  Y[14] = T(1.445305721320277) * z * (xx - yy);
  Y[15] = T(-0.5900435899266435) * x * (xx - T(3) * yy);
}

// gradient of sum_k w[k] Y_k with respect to (x, y, z) (treating them as free variables)
template <class T> CHS_HD void chs_sh_basis_bwd(int deg, T x, T y, T z, const T w[CHS_SH_MAX_COEFFS], T& gx, T& gy, T& gz) {
  gx = gy = gz = T(0);
  if (deg < 1) return;
  const T c1 = T(0.4886025119029199);
  gy += -c1 * w[1]; gz += c1 * w[2]; gx += -c1 * w[3];
  if (deg < 2) return;
  const T a = T(1.0925484305920792), b6 = T(0.31539156525252005), b8 = T(0.5462742152960396);
  gx += a * y * w[4];             gy += a * x * w[4];
  gy += -a * z * w[5];            gz += -a * y * w[5];
  gx += -T(2) * b6 * x * w[6];    gy += -T(2) * b6 * y * w[6];   gz += T(4) * b6 * z * w[6];
  gx += -a * z * w[7];            gz += -a * x * w[7];
  gx += T(2) * b8 * x * w[8];     gy += -T(2) * b8 * y * w[8];
  if (deg < 3) return;
  const T xx = x * x, yy = y * y, zz = z * z;
  const T c9 = T(-0.5900435899266435), c10 = T(2.890611442640554), c11 = T(-0.4570457994644658), c12 = T(0.3731763325901154),
          c14 = T(1.445305721320277);
  // Y9 = c9 y (3xx - yy)
  gx += c9 * T(6) * x * y * w[9];                       gy += c9 * (T(3) * xx - T(3) * yy) * w[9];
  // Y10 = c10 x y z
  gx += c10 * y * z * w[10];  gy += c10 * x * z * w[10];  gz += c10 * x * y * w[10];
  // Y11 = c11 y (4zz - xx - yy)
  gx += c11 * (-T(2) * x * y) * w[11];  gy += c11 * (T(4) * zz - xx - T(3) * yy) * w[11];  gz += c11 * T(8) * y * z * w[11];
  // Y12 = c12 z (2zz - 3xx - 3yy)
  gx += c12 * (-T(6) * x * z) * w[12];  gy += c12 * (-T(6) * y * z) * w[12];  gz += c12 * (T(6) * zz - T(3) * xx - T(3) * yy) * w[12];
  // Y13 = c11 x (4zz - xx - yy)
  gx += c11 * (T(4) * zz - T(3) * xx - yy) * w[13];  gy += c11 * (-T(2) * x * y) * w[13];  gz += c11 * T(8) * x * z * w[13];
  // Y14 = c14 z (xx - yy)
  gx += c14 * T(2) * x * z * w[14];  gy += -c14 * T(2) * y * z * w[14];  gz += c14 * (xx - yy) * w[14];
  // Y15 = c9 x (xx - 3yy)
  gx += c9 * (T(3) * xx - T(3) * yy) * w[15];  gy += c9 * (-T(6) * x * y) * w[15];
}

// Camera centre in world coordinates from the world->camera pose: campos = -R^T t
template <class T> CHS_HD void chs_campos(const T R[9], const T t[3], T cp[3]) {
  for (int i = 0; i < 3; ++i) cp[i] = -(R[i] * t[0] + R[3 + i] * t[1] + R[6 + i] * t[2]);
}

// colour of one (camera, Gaussian): sh is [K][3] row-major (K = (deg+1)^2). Returns rgb (post relu).
template <class T> CHS_HD void chs_sh_color(int deg, const T* sh, const T mu[3], const T cp[3], T rgb[3]) {
  T d[3] = {mu[0] - cp[0], mu[1] - cp[1], mu[2] - cp[2]};
  T inv = T(1) / sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  T Y[CHS_SH_MAX_COEFFS];
  chs_sh_basis(deg, d[0] * inv, d[1] * inv, d[2] * inv, Y);
  const int K = (deg + 1) * (deg + 1);
  for (int ch = 0; ch < 3; ++ch) {
    T acc = T(0.5);
    for (int k = 0; k < K; ++k) acc += sh[k * 3 + ch] * Y[k];
    rgb[ch] = chs_max(acc, T(0));
  }
}

// backward of chs_sh_color: accumulates v_sh [K][3], v_mu[3], v_cp[3] given v_rgb[3]
template <class T>
CHS_HD void chs_sh_color_bwd(int deg, const T* sh, const T mu[3], const T cp[3], const T v_rgb[3], T* v_sh, T v_mu[3], T v_cp[3]) {
  T d[3] = {mu[0] - cp[0], mu[1] - cp[1], mu[2] - cp[2]};
  T inv = T(1) / sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  T x = d[0] * inv, y = d[1] * inv, z = d[2] * inv;
  T Y[CHS_SH_MAX_COEFFS];
  chs_sh_basis(deg, x, y, z, Y);
  const int K = (deg + 1) * (deg + 1);
  T vc[3];
  for (int ch = 0; ch < 3; ++ch) {
    T acc = T(0.5);
    for (int k = 0; k < K; ++k) acc += sh[k * 3 + ch] * Y[k];
    vc[ch] = acc > T(0) ? v_rgb[ch] : T(0);  // relu
  }
  T w[CHS_SH_MAX_COEFFS];
  for (int k = 0; k < K; ++k) {
    w[k] = sh[k * 3] * vc[0] + sh[k * 3 + 1] * vc[1] + sh[k * 3 + 2] * vc[2];
    v_sh[k * 3] += Y[k] * vc[0];
    v_sh[k * 3 + 1] += Y[k] * vc[1];
    v_sh[k * 3 + 2] += Y[k] * vc[2];
  }
  T gx, gy, gz;
  chs_sh_basis_bwd(deg, x, y, z, w, gx, gy, gz);
  // through the normalisation dir = d / |d|
  T dot = gx * x + gy * y + gz * z;
  T vd[3] = {(gx - dot * x) * inv, (gy - dot * y) * inv, (gz - dot * z) * inv};
  for (int i = 0; i < 3; ++i) {
    v_mu[i] += vd[i];
    v_cp[i] -= vd[i];
  }
}

// campos = -R^T t: v_R[j][i] += -t_j v_cp_i, v_t[j] += -(R v_cp)_j
template <class T> CHS_HD void chs_campos_bwd(const T R[9], const T t[3], const T v_cp[3], T v_R[9], T v_t[3]) {
  for (int j = 0; j < 3; ++j) {
    v_t[j] += -(R[j * 3] * v_cp[0] + R[j * 3 + 1] * v_cp[1] + R[j * 3 + 2] * v_cp[2]);
    for (int i = 0; i < 3; ++i) v_R[j * 3 + i] += -t[j] * v_cp[i];
  }
}
