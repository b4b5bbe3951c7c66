// chs_spline.cuh — SE(3) trajectory spline (decisions D3/D4), evaluated in
// fp64 and templated on the scalar so the backward can use forward-mode dual numbers: the pose of
// one virtual camera depends on at most 4 knots x 7 numbers + the time parameter u = 29 inputs, so
// 29 one-tangent evaluations give its full Jacobian — exact, and negligible next to the rest of
// the step (C <= 128 cameras).  Host+device so tests/hostsim can check it against the oracle.
//
// Required by "Trajectory control knots", "Camera motion spline", "Virtual camera pose",
// "Exposure time range" in /root/reference/assets/pipeline.png (referenced at Readme.md:50).
#pragma once
#include "../../include/chs.h"
#include "chs_math.cuh"

#define CHS_SMALL_ANGLE 1e-4  // same threshold as oracle/se3.py SMALL_ANGLE

struct ChsDual {
  double v, d;
  CHS_HD ChsDual() : v(0), d(0) {}
  CHS_HD ChsDual(double v_) : v(v_), d(0) {}
  CHS_HD ChsDual(double v_, double d_) : v(v_), d(d_) {}
};
CHS_HD ChsDual operator+(ChsDual a, ChsDual b) { return ChsDual(a.v + b.v, a.d + b.d); }
CHS_HD ChsDual operator-(ChsDual a, ChsDual b) { return ChsDual(a.v - b.v, a.d - b.d); }
CHS_HD ChsDual operator-(ChsDual a) { return ChsDual(-a.v, -a.d); }
CHS_HD ChsDual operator*(ChsDual a, ChsDual b) { return ChsDual(a.v * b.v, a.d * b.v + a.v * b.d); }
CHS_HD ChsDual operator/(ChsDual a, ChsDual b) {
  double q = a.v / b.v;
  return ChsDual(q, (a.d - q * b.d) / b.v);
}
CHS_HD bool operator<(ChsDual a, ChsDual b) { return a.v < b.v; }
CHS_HD ChsDual chs_sqrt(ChsDual a) {
  double s = sqrt(a.v);
  return ChsDual(s, a.d / (2.0 * s));
}
CHS_HD ChsDual chs_sin(ChsDual a) { return ChsDual(sin(a.v), cos(a.v) * a.d); }
CHS_HD ChsDual chs_cos(ChsDual a) { return ChsDual(cos(a.v), -sin(a.v) * a.d); }
CHS_HD ChsDual chs_atan2(ChsDual y, ChsDual x) {
  double r2 = x.v * x.v + y.v * y.v;
  return ChsDual(atan2(y.v, x.v), (x.v * y.d - y.v * x.d) / r2);
}
CHS_HD double chs_sqrt(double a) { return sqrt(a); }
CHS_HD double chs_sin(double a) { return sin(a); }
CHS_HD double chs_cos(double a) { return cos(a); }
CHS_HD double chs_atan2(double y, double x) { return atan2(y, x); }
CHS_HD double chs_value(double a) { return a; }
CHS_HD double chs_value(ChsDual a) { return a.v; }

template <class S> CHS_HD void chs_sp_quat_normalize(const S q[4], S o[4]) {
  S n = S(1.0) / chs_sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; ++i) o[i] = q[i] * n;
}
template <class S> CHS_HD void chs_sp_rotmat(const S q[4], S R[9]) {  // q unit
  S w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = S(1.0) - S(2.0) * (y * y + z * z); R[1] = S(2.0) * (x * y - w * z); R[2] = S(2.0) * (x * z + w * y);
  R[3] = S(2.0) * (x * y + w * z); R[4] = S(1.0) - S(2.0) * (x * x + z * z); R[5] = S(2.0) * (y * z - w * x);
  R[6] = S(2.0) * (x * z - w * y); R[7] = S(2.0) * (y * z + w * x); R[8] = S(1.0) - S(2.0) * (x * x + y * y);
}
template <class S> CHS_HD void chs_sp_matmul(const S A[9], const S B[9], S C[9]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
template <class S> CHS_HD void chs_sp_matvec(const S A[9], const S v[3], S o[3]) {
  for (int i = 0; i < 3; ++i) o[i] = A[i * 3] * v[0] + A[i * 3 + 1] * v[1] + A[i * 3 + 2] * v[2];
}
template <class S> CHS_HD void chs_sp_hat2(const S p[3], S K[9], S K2[9]) {
  S o = S(0.0);
  K[0] = o; K[1] = -p[2]; K[2] = p[1];
  K[3] = p[2]; K[4] = o; K[5] = -p[0];
  K[6] = -p[1]; K[7] = p[0]; K[8] = o;
  chs_sp_matmul(K, K, K2);
}

// A = sin t / t, B = (1 - cos t) / t^2, C = (t - sin t) / t^3, Taylor below CHS_SMALL_ANGLE
template <class S> CHS_HD void chs_sp_abc(S theta2, S& A, S& B, S& C) {
  if (chs_value(theta2) < CHS_SMALL_ANGLE * CHS_SMALL_ANGLE) {
    A = S(1.0) - theta2 / S(6.0);
    B = S(0.5) - theta2 / S(24.0);
    C = S(1.0 / 6.0) - theta2 / S(120.0);
  } else {
    S t = chs_sqrt(theta2);
    A = chs_sin(t) / t;
    B = (S(1.0) - chs_cos(t)) / theta2;
    C = (t - chs_sin(t)) / (theta2 * t);
  }
}

// (R, t) = Exp(rho, phi)
template <class S> CHS_HD void chs_sp_se3_exp(const S rho[3], const S phi[3], S R[9], S t[3]) {
  S th2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
  S A, B, C;
  chs_sp_abc(th2, A, B, C);
  S K[9], K2[9], V[9];
  chs_sp_hat2(phi, K, K2);
  for (int i = 0; i < 9; ++i) {
    S eye = (i % 4 == 0) ? S(1.0) : S(0.0);
    R[i] = eye + A * K[i] + B * K2[i];
    V[i] = eye + B * K[i] + C * K2[i];
  }
  chs_sp_matvec(V, rho, t);
}

// (rho, phi) = Log(T_a^-1 T_b); knots as (t, unit q)
template <class S> CHS_HD void chs_sp_rel_log(const S ta[3], const S qa[4], const S tb[3], const S qb[4], S rho[3], S phi[3]) {
  S Ra[9];
  chs_sp_rotmat(qa, Ra);
  // q_rel = conj(qa) * qb
  S aw = qa[0], ax = -qa[1], ay = -qa[2], az = -qa[3];
  S bw = qb[0], bx = qb[1], by = qb[2], bz = qb[3];
  S q[4] = {aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
            aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw};
  if (chs_value(q[0]) < 0.0)
    for (int i = 0; i < 4; ++i) q[i] = -q[i];
  S d[3] = {tb[0] - ta[0], tb[1] - ta[1], tb[2] - ta[2]};
  S t_rel[3];
  for (int i = 0; i < 3; ++i) t_rel[i] = Ra[i] * d[0] + Ra[3 + i] * d[1] + Ra[6 + i] * d[2];  // Ra^T d
  S s2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  S k;
  if (chs_value(s2) < CHS_SMALL_ANGLE * CHS_SMALL_ANGLE) {
    k = (S(2.0) / q[0]) * (S(1.0) - s2 / (S(3.0) * q[0] * q[0]));
  } else {
    S s = chs_sqrt(s2);
    k = S(2.0) * chs_atan2(s, q[0]) / s;
  }
  for (int i = 0; i < 3; ++i) phi[i] = q[1 + i] * k;
  S th2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
  S D;
  if (chs_value(th2) < CHS_SMALL_ANGLE * CHS_SMALL_ANGLE) {
    D = S(1.0 / 12.0) + th2 / S(720.0);
  } else {
    S t = chs_sqrt(th2);
    D = S(1.0) / th2 - (S(1.0) + chs_cos(t)) / (S(2.0) * t * chs_sin(t));
  }
  S K[9], K2[9], Vi[9];
  chs_sp_hat2(phi, K, K2);
  for (int i = 0; i < 9; ++i) {
    S eye = (i % 4 == 0) ? S(1.0) : S(0.0);
    Vi[i] = eye - S(0.5) * K[i] + D * K2[i];
  }
  chs_sp_matvec(Vi, t_rel, rho);
}

// World-to-camera pose (rows 0..2 of the view matrix, row major 3x4) of the spline at local
// parameter u.  kind 0 (linear): knots k[0], k[1].  kind 1 (cubic cumulative B-spline): k[0..3] =
// knots s-1 .. s+2.  Knot layout: (tx, ty, tz, qw, qx, qy, qz), quaternions normalised here.
template <class S> CHS_HD void chs_spline_viewmat(int kind, const S k[4][7], S u, S vm[12]) {
  const int nk = kind == CHS_SPLINE_LINEAR ? 2 : 4;
  S q[4][4];
  for (int i = 0; i < nk; ++i) chs_sp_quat_normalize(&k[i][3], q[i]);
  S R[9], t[3];
  chs_sp_rotmat(q[0], R);
  for (int i = 0; i < 3; ++i) t[i] = k[0][i];
  S w[3];
  int nseg;
  if (kind == CHS_SPLINE_LINEAR) {
    w[0] = u;
    nseg = 1;
  } else {
    S u2 = u * u, u3 = u2 * u;
    w[0] = (S(5.0) + S(3.0) * u - S(3.0) * u2 + u3) / S(6.0);
    w[1] = (S(1.0) + S(3.0) * u + S(3.0) * u2 - S(2.0) * u3) / S(6.0);
    w[2] = u3 / S(6.0);
    nseg = 3;
  }
  for (int j = 0; j < nseg; ++j) {
    S rho[3], phi[3];
    chs_sp_rel_log(k[j], q[j], k[j + 1], q[j + 1], rho, phi);
    for (int i = 0; i < 3; ++i) {
      rho[i] = rho[i] * w[j];
      phi[i] = phi[i] * w[j];
    }
    S Rd[9], td[3], Rn[9], tn[3];
    chs_sp_se3_exp(rho, phi, Rd, td);
    chs_sp_matmul(R, Rd, Rn);
    chs_sp_matvec(R, td, tn);
    for (int i = 0; i < 9; ++i) R[i] = Rn[i];
    for (int i = 0; i < 3; ++i) t[i] = tn[i] + t[i];
  }
  // invert: Rv = R^T, tv = -R^T t
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) vm[i * 4 + j] = R[j * 3 + i];
    vm[i * 4 + 3] = -(R[i] * t[0] + R[3 + i] * t[1] + R[6 + i] * t[2]);
  }
}

// segment index (clamped) and local parameter for a sample time
CHS_HD void chs_spline_segment(int kind, int n_knots, double knot_t0, double knot_dt, double time, int& s, double& u) {
  double x = (time - knot_t0) / knot_dt;
  int lo = kind == CHS_SPLINE_LINEAR ? 0 : 1;
  int hi = kind == CHS_SPLINE_LINEAR ? n_knots - 2 : n_knots - 3;
  double f = floor(x);
  s = f < (double)lo ? lo : (f > (double)hi ? hi : (int)f);
  u = x - (double)s;
}

CHS_HD double chs_sample_weight(int k, int n_virtual) { return n_virtual > 1 ? (double)k / (double)(n_virtual - 1) - 0.5 : 0.0; }
