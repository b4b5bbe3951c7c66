// hostsim.cpp — TEST-ONLY host harness around casualhdrsplat_b200/csrc/chs_math.cuh and
// chs_spline.cuh.  It compiles the exact per-element arithmetic the CUDA kernels use with g++ and
// runs it in plain serial loops so tests/test_hostsim_*.py can compare it with the float64 oracle
// in this GPU-less container.  It is NOT part of the product: libchs.so never contains or calls
// this code, and casualhdrsplat_b200 has no CPU path.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

#include "../../casualhdrsplat_b200/csrc/chs_math.cuh"
#include "../../casualhdrsplat_b200/csrc/chs_sh.cuh"
#include "../../casualhdrsplat_b200/csrc/chs_spline.cuh"

template <class T> static void load_cam(const T* viewmats, const T* Ks, int c, ChsCam<T>& cam) {
  const T* v = viewmats + c * 16;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) cam.R[i * 3 + j] = v[i * 4 + j];
    cam.t[i] = v[i * 4 + 3];
  }
  const T* K = Ks + c * 9;
  cam.fx = K[0]; cam.fy = K[4]; cam.cx = K[2]; cam.cy = K[5];
}

template <class T>
static void project_fwd_impl(int N, int C, int W, int H, T near_p, T far_p, T eps2d, const T* means, const T* quats, const T* scales,
                             const T* viewmats, const T* Ks, T* means2d, T* depths, T* conics, int32_t* radii, int32_t* touched) {
  int tile_w = (W + 15) / 16, tile_h = (H + 15) / 16;
  for (int g = 0; g < N; ++g) {
    T S[6];
    chs_cov3d(quats + g * 4, scales + g * 3, S);
    for (int c = 0; c < C; ++c) {
      ChsCam<T> cam;
      load_cam(viewmats, Ks, c, cam);
      ChsProj<T> pr;
      int r = chs_project_fwd(means + g * 3, S, cam, (T)W, (T)H, near_p, far_p, eps2d, pr);
      size_t o = (size_t)c * N + g;
      means2d[o * 2] = pr.mx; means2d[o * 2 + 1] = pr.my;
      depths[o] = pr.depth;
      conics[o * 3] = pr.ca; conics[o * 3 + 1] = pr.cb; conics[o * 3 + 2] = pr.cc;
      radii[o] = r;
      int t = 0;
      if (r > 0) {
        ChsTileRect tr = chs_tile_bounds((float)pr.mx, (float)pr.my, r, tile_w, tile_h);
        t = (tr.x1 - tr.x0) * (tr.y1 - tr.y0);
      }
      touched[o] = t;
    }
  }
}

template <class T>
static void project_bwd_impl(int N, int C, int W, int H, T eps2d, const T* means, const T* quats, const T* scales, const T* viewmats,
                             const T* Ks, const int32_t* radii, const T* v_means2d, const T* v_conics, T* v_means, T* v_quats,
                             T* v_scales, T* v_viewmats /* [C,12]: R(9), t(3) */) {
  std::memset(v_viewmats, 0, sizeof(T) * C * 12);
  for (int g = 0; g < N; ++g) {
    T S[6];
    chs_cov3d(quats + g * 4, scales + g * 3, S);
    T v_mu[3] = {0, 0, 0}, G[6] = {0, 0, 0, 0, 0, 0};
    for (int c = 0; c < C; ++c) {
      size_t o = (size_t)c * N + g;
      if (radii[o] <= 0) continue;
      ChsCam<T> cam;
      load_cam(viewmats, Ks, c, cam);
      chs_project_bwd(means + g * 3, S, cam, (T)W, (T)H, eps2d, v_means2d[o * 2], v_means2d[o * 2 + 1], v_conics[o * 3],
                      v_conics[o * 3 + 1], v_conics[o * 3 + 2], v_mu, G, v_viewmats + c * 12, v_viewmats + c * 12 + 9);
    }
    T vq[4] = {0, 0, 0, 0}, vs[3] = {0, 0, 0};
    chs_cov3d_bwd(quats + g * 4, scales + g * 3, G, vq, vs);
    for (int k = 0; k < 3; ++k) { v_means[g * 3 + k] = v_mu[k]; v_scales[g * 3 + k] = vs[k]; }
    for (int k = 0; k < 4; ++k) v_quats[g * 4 + k] = vq[k];
  }
}

// One tile list blended over a set of pixels, forward then backward, the way the kernels walk it.
// splat params per list entry: mean2d (2), conic (3), opacity, rgb (3) = 9 numbers.
// tabled != 0: the backward in the form blend_bwd2_kernel uses (per-pair scalars -> moment sums -> gradients);
// tabled == 4: the division-free form of blend_bwd3_kernel (chs_pair_bwd_scalars_r, negated scalars, chs_moments_to_grads_neg)
template <class T>
static void blend_impl(int n_list, const T* params, int n_pix, const T* pix_xy, const T* bg, const T* v_hdr, const T* v_alpha,
                       T* out_hdr, T* out_alpha, int32_t* out_last, T* v_params, int tabled = 0) {
  std::vector<T> moments((size_t)n_list * 9, T(0));
  std::vector<ChsSplat<T>> sp(n_list);
  for (int j = 0; j < n_list; ++j) {
    const T* p = params + j * 9;
    chs_make_splat(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], sp[j]);
  }
  const T thr = (T)log2(1.0 / 255.0);
  std::memset(v_params, 0, sizeof(T) * n_list * 9);
  // tabled == 5: blend_bwd4_kernel's tensor-core phase B.  Per 8x8 pixel block (a warp): v_H scaled by a power of two into fp16
  // range and split hi + lo, nvs scaled and split hi + lo in fp16, nf rounded to fp16; origin-relative monomial sums and colour
  // sums accumulated per (Gaussian, block), shifted to the mean (chs_shift_moments) and unscaled.
  auto h16 = [](T v) { return (T)(float)(_Float16)(float)v; };
  auto block_of = [&](int p) {
    const int bx = (int)floor((pix_xy[p * 2] - T(0.5)) / 8), by = (int)floor((pix_xy[p * 2 + 1] - T(0.5)) / 8);
    return by * 4096 + bx;
  };
  std::map<int, T> blk_vmax, blk_ebound;
  std::map<int, std::vector<T>> blk_sums;  // block -> [n_list][9] (S1, Sx, Sy, Sxx, Sxy, Syy, Gr, Gg, Gb), scaled
  if (tabled >= 5) {
    T cbnd = 0;  // the kernel keeps a running bound over the staged batches; one batch here
    for (int j = 0; j < n_list; ++j) cbnd = chs_max(cbnd, (T)(fabs((double)sp[j].cr) + fabs((double)sp[j].cg) + fabs((double)sp[j].cb)));
    for (int p = 0; p < n_pix; ++p) {
      T& m = blk_vmax[block_of(p)];
      for (int ch = 0; ch < 3; ++ch) m = chs_max(m, (T)fabs((double)v_hdr[p * 3 + ch]));
    }
    for (int p = 0; p < n_pix; ++p) {  // |e| = |R - s| <= 2 max(|R0|, max_j |s_j|),  |s_j| <= (|cr| + |cg| + |cb|) vmax
      const int b = block_of(p);
      const T r0 = (T)fabs((double)((bg[0] * v_hdr[p * 3] + bg[1] * v_hdr[p * 3 + 1] + bg[2] * v_hdr[p * 3 + 2]) - v_alpha[p]));
      T& eb = blk_ebound[b];
      eb = chs_max(eb, 2 * chs_max(r0, cbnd * blk_vmax[b]));
    }
  }
  for (int p = 0; p < n_pix; ++p) {
    T px = pix_xy[p * 2], py = pix_xy[p * 2 + 1];
    T Tr = 1, acc[3] = {0, 0, 0};
    int last = 0;
    for (int j = 0; j < n_list; ++j) {
      // sub-tile cull exactly as the kernels apply it (8x4 pixel block of this pixel's warp): must never
      // drop a contributing pair
      T bx0 = floor((px - T(0.5)) / 8) * 8 + T(0.5), by0 = floor((py - T(0.5)) / 8) * 8 + T(0.5);
      bool reach = chs_block_max_power(sp[j], bx0, bx0 + 7, by0, by0 + 7) >= thr - T(1e-3);
      T dx, dy, u;
      T power = chs_pair_power(sp[j], px, py, dx, dy, u);
      if (!(power >= thr)) continue;
      if (!reach) { out_last[p] = -1000000; }  // flag: the cull would have dropped a live pair
      T alpha = chs_min(ChsK<T>::alpha_max, chs_exp2_fast(power));
      T Tn = Tr * (1 - alpha);
      if (Tn <= ChsK<T>::t_stop) break;
      T w = alpha * Tr;
      acc[0] += w * sp[j].cr; acc[1] += w * sp[j].cg; acc[2] += w * sp[j].cb;
      Tr = Tn;
      last = j + 1;
    }
    for (int ch = 0; ch < 3; ++ch) out_hdr[p * 3 + ch] = acc[ch] + Tr * bg[ch];
    out_alpha[p] = 1 - Tr;
    if (out_last[p] != -1000000) out_last[p] = last;
    // backward
    const T vh[3] = {v_hdr[p * 3], v_hdr[p * 3 + 1], v_hdr[p * 3 + 2]};
    T va_t = Tr * (v_alpha[p] - (bg[0] * vh[0] + bg[1] * vh[1] + bg[2] * vh[2]));
    T buf[3] = {0, 0, 0};
    T Rn = (bg[0] * vh[0] + bg[1] * vh[1] + bg[2] * vh[2]) - v_alpha[p];  // behind the last Gaussian: the background
    for (int j = last - 1; j >= 0; --j) {
      T dx, dy, u;
      T power = chs_pair_power(sp[j], px, py, dx, dy, u);
      if (!(power >= thr)) continue;
      T au = chs_exp2_fast(power);
      if (tabled == 4) {
        T nvs, nf;
        chs_pair_bwd_scalars_r(sp[j], au, chs_min(ChsK<T>::alpha_max, au), Tr, Rn, vh, nvs, nf);
        chs_pair_moments(nvs, nf, dx, dy, u, vh, &moments[(size_t)j * 9]);
        continue;
      }
      if (tabled >= 5) {  // 6: the kernel (alpha T as fp16 hi + lo); what-ifs: 5: alpha T single fp16; 7: alpha T and v_H single
        T nvs, nf;
        chs_pair_bwd_scalars_r(sp[j], au, chs_min(ChsK<T>::alpha_max, au), Tr, Rn, vh, nvs, nf);
        const int b = block_of(p);
        int e;
        frexp((double)blk_vmax[b], &e);              // vmax = f 2^e, f in [0.5, 1)
        const T sv = (T)ldexp(1.0, -e);              // scale of v_H: |v_H| sv < 1
        frexp((double)blk_ebound[b], &e);
        const T sn = (T)ldexp(1.0, 14 - e);          // scale of nvs: |nvs| sn <= |e| sn < 2^14
        std::vector<T>& S = blk_sums[b];
        if (S.empty()) S.assign((size_t)n_list * 9, T(0));
        const T x = px - (floor((px - T(0.5)) / 8) * 8 + T(0.5)), y = py - (floor((py - T(0.5)) / 8) * 8 + T(0.5));
        const T hi = h16(nvs * sn), lo = h16(nvs * sn - hi), q = hi + lo;
        const T mono[6] = {T(1), x, y, x * x, x * y, y * y};
        for (int k = 0; k < 6; ++k) S[(size_t)j * 9 + k] += q * mono[k];
        const T fs = nf * T(16384);
        const T fhi = h16(fs), flo = tabled == 6 ? h16(fs - fhi) : T(0);
        for (int ch = 0; ch < 3; ++ch) {
          const T vhi = h16(vh[ch] * sv), vlo = tabled == 7 ? T(0) : h16(vh[ch] * sv - vhi);
          S[(size_t)j * 9 + 6 + ch] += (fhi * vhi + fhi * vlo) + (flo * vhi + flo * vlo);
        }
        continue;
      }
      if (tabled) {
        T vs, f;
        chs_pair_bwd_scalars(sp[j], au, chs_min(ChsK<T>::alpha_max, au), Tr, buf, vh, va_t, vs, f);
        if (tabled == 2) {  // what-if: the table held in IEEE half precision
          vs = (T)(float)(_Float16)(float)vs;
          f = (T)(float)(_Float16)(float)f;
        } else if (tabled == 3) {  // what-if: bfloat16 (round to nearest even on the top 16 bits)
          auto bf = [](T v) {
            float x = (float)v;
            uint32_t b;
            std::memcpy(&b, &x, 4);
            b += 0x7fffu + ((b >> 16) & 1u);
            b &= 0xffff0000u;
            std::memcpy(&x, &b, 4);
            return (T)x;
          };
          vs = bf(vs);
          f = bf(f);
        }
        chs_pair_moments(vs, f, dx, dy, u, vh, &moments[(size_t)j * 9]);
        continue;
      }
      T g[9];
      chs_pair_bwd(sp[j], dx, dy, u, au, chs_min(ChsK<T>::alpha_max, au), Tr, buf, vh, va_t, g);
      // g = [v_mx, v_my, v_A, v_B, v_C, v_o, v_r, v_g, v_b] -> params order (mx,my,A,B,C,o,r,g,b)
      for (int k = 0; k < 9; ++k) v_params[j * 9 + k] += g[k];
    }
  }
  if (tabled >= 5) {
    for (auto& kv : blk_sums) {
      const int b = kv.first;
      const T bxc = (T)(b % 4096) * 8 + T(0.5), byc = (T)(b / 4096) * 8 + T(0.5);
      int e;
      frexp((double)blk_vmax[b], &e);
      const T inv_sv = (T)ldexp(1.0, e - 14);  // also undoes the 2^14 of nf
      frexp((double)blk_ebound[b], &e);
      const T inv_sn = (T)ldexp(1.0, e - 14);
      for (int j = 0; j < n_list; ++j) {
        const T* S = &kv.second[(size_t)j * 9];
        T m[6];
        chs_shift_moments(S[0], S[1], S[2], S[3], S[4], S[5], sp[j].mx - bxc, sp[j].my - byc, sp[j].r, m);
        for (int k = 0; k < 6; ++k) moments[(size_t)j * 9 + k] += m[k] * inv_sn;
        for (int ch = 0; ch < 3; ++ch) moments[(size_t)j * 9 + 6 + ch] += S[6 + ch] * inv_sv;
      }
    }
    for (int j = 0; j < n_list; ++j) chs_moments_to_grads_neg(sp[j], &moments[(size_t)j * 9], v_params + (size_t)j * 9);
  } else if (tabled == 4)
    for (int j = 0; j < n_list; ++j) chs_moments_to_grads_neg(sp[j], &moments[(size_t)j * 9], v_params + (size_t)j * 9);
  else if (tabled)
    for (int j = 0; j < n_list; ++j) chs_moments_to_grads(sp[j], &moments[(size_t)j * 9], v_params + (size_t)j * 9);
}

// CRF MLP in interval form, step by step as crf_bwd_interval_kernel does it (ranking, per-interval slope / offset, bisection,
// two sums per interval, per-unit sums over the intervals where the unit is on).  Same outputs as hs_crf_f64.
template <class T>
static void crf_interval_impl(int n, const T* X, const T* p, int hd, const T* v_y, T* y, T* dydx, T* v_params) {
  const int nI = hd + 1;
  std::vector<T> bp(hd), A(nI), B(nI), H(nI, T(0)), HZ(nI, T(0));
  std::vector<int> key(hd);
  for (int j = 0; j < hd; ++j) {
    const T t = chs_crf_breakpoint(p[j], p[hd + j]);
    int r = 0;
    for (int k = 0; k < hd; ++k) {
      const T tk = chs_crf_breakpoint(p[k], p[hd + k]);
      r += (tk < t || (tk == t && k < j)) ? 1 : 0;
    }
    bp[r] = t;
    key[j] = chs_crf_unit_key(p[j], p[hd + j], r, hd);
  }
  for (int I = 0; I < nI; ++I) {
    T a = 0, b = p[3 * hd];
    for (int j = 0; j < hd; ++j)
      if (chs_crf_key_on(key[j], I)) {
        a = chs_fma(p[2 * hd + j], p[j], a);
        b = chs_fma(p[2 * hd + j], p[hd + j], b);
      }
    A[I] = a;
    B[I] = b;
  }
  for (int i = 0; i < n; ++i) {
    const T xe = chs_max(X[i], T(0)) + ChsK<T>::crf_eps;
    const T z = log(xe);
    const int I = chs_crf_interval_of(bp.data(), hd, z);
    const T acc = chs_fma(A[I], z, B[I]);
    y[i] = T(1) / (T(1) + exp(-acc));
    const T g = v_y[i] * y[i] * (T(1) - y[i]);
    dydx[i] = X[i] >= T(0) ? y[i] * (T(1) - y[i]) * A[I] / xe : T(0);
    H[I] += g;
    HZ[I] += g * z;
  }
  T gb2 = 0;
  for (int I = 0; I < nI; ++I) gb2 += H[I];
  for (int j = 0; j < hd; ++j) {
    T S = 0, SZ = 0;
    for (int I = 0; I < nI; ++I)
      if (chs_crf_key_on(key[j], I)) {
        S += H[I];
        SZ += HZ[I];
      }
    v_params[j] = p[2 * hd + j] * SZ;
    v_params[hd + j] = p[2 * hd + j] * S;
    v_params[2 * hd + j] = chs_fma(p[j], SZ, p[hd + j] * S);
  }
  v_params[3 * hd] = gb2;
}

extern "C" {

void hs_project_fwd_f64(int N, int C, int W, int H, double near_p, double far_p, double eps2d, const double* means, const double* quats,
                        const double* scales, const double* viewmats, const double* Ks, double* means2d, double* depths, double* conics,
                        int32_t* radii, int32_t* touched) {
  project_fwd_impl<double>(N, C, W, H, near_p, far_p, eps2d, means, quats, scales, viewmats, Ks, means2d, depths, conics, radii, touched);
}
void hs_project_fwd_f32(int N, int C, int W, int H, float near_p, float far_p, float eps2d, const float* means, const float* quats,
                        const float* scales, const float* viewmats, const float* Ks, float* means2d, float* depths, float* conics,
                        int32_t* radii, int32_t* touched) {
  project_fwd_impl<float>(N, C, W, H, near_p, far_p, eps2d, means, quats, scales, viewmats, Ks, means2d, depths, conics, radii, touched);
}
void hs_project_bwd_f64(int N, int C, int W, int H, double eps2d, const double* means, const double* quats, const double* scales,
                        const double* viewmats, const double* Ks, const int32_t* radii, const double* v_means2d, const double* v_conics,
                        double* v_means, double* v_quats, double* v_scales, double* v_viewmats) {
  project_bwd_impl<double>(N, C, W, H, eps2d, means, quats, scales, viewmats, Ks, radii, v_means2d, v_conics, v_means, v_quats, v_scales,
                           v_viewmats);
}
void hs_project_bwd_f32(int N, int C, int W, int H, float eps2d, const float* means, const float* quats, const float* scales,
                        const float* viewmats, const float* Ks, const int32_t* radii, const float* v_means2d, const float* v_conics,
                        float* v_means, float* v_quats, float* v_scales, float* v_viewmats) {
  project_bwd_impl<float>(N, C, W, H, eps2d, means, quats, scales, viewmats, Ks, radii, v_means2d, v_conics, v_means, v_quats, v_scales,
                          v_viewmats);
}
void hs_blend_f64(int n_list, const double* params, int n_pix, const double* pix_xy, const double* bg, const double* v_hdr,
                  const double* v_alpha, double* out_hdr, double* out_alpha, int32_t* out_last, double* v_params) {
  blend_impl<double>(n_list, params, n_pix, pix_xy, bg, v_hdr, v_alpha, out_hdr, out_alpha, out_last, v_params);
}
void hs_blend_f32(int n_list, const float* params, int n_pix, const float* pix_xy, const float* bg, const float* v_hdr,
                  const float* v_alpha, float* out_hdr, float* out_alpha, int32_t* out_last, float* v_params) {
  blend_impl<float>(n_list, params, n_pix, pix_xy, bg, v_hdr, v_alpha, out_hdr, out_alpha, out_last, v_params);
}
void hs_blend_tabled_f64(int n_list, const double* params, int n_pix, const double* pix_xy, const double* bg, const double* v_hdr,
                         const double* v_alpha, double* out_hdr, double* out_alpha, int32_t* out_last, double* v_params) {
  blend_impl<double>(n_list, params, n_pix, pix_xy, bg, v_hdr, v_alpha, out_hdr, out_alpha, out_last, v_params, 1);
}
void hs_blend_tabled_f32(int n_list, const float* params, int n_pix, const float* pix_xy, const float* bg, const float* v_hdr,
                         const float* v_alpha, float* out_hdr, float* out_alpha, int32_t* out_last, float* v_params) {
  blend_impl<float>(n_list, params, n_pix, pix_xy, bg, v_hdr, v_alpha, out_hdr, out_alpha, out_last, v_params, 1);
}
void hs_blend_tabled_r_f64(int n_list, const double* params, int n_pix, const double* pix_xy, const double* bg, const double* v_hdr,
                           const double* v_alpha, double* out_hdr, double* out_alpha, int32_t* out_last, double* v_params) {
  blend_impl<double>(n_list, params, n_pix, pix_xy, bg, v_hdr, v_alpha, out_hdr, out_alpha, out_last, v_params, 4);
}
void hs_blend_tabled_mma_f32(int n_list, const float* params, int n_pix, const float* pix_xy, const float* bg, const float* v_hdr,
                             const float* v_alpha, float* out_hdr, float* out_alpha, int32_t* out_last, float* v_params) {
  blend_impl<float>(n_list, params, n_pix, pix_xy, bg, v_hdr, v_alpha, out_hdr, out_alpha, out_last, v_params, 6);  // the kernel's form
}
void hs_blend_tabled_mma_variant_f32(int mode, int n_list, const float* params, int n_pix, const float* pix_xy, const float* bg,
                                     const float* v_hdr, const float* v_alpha, float* out_hdr, float* out_alpha, int32_t* out_last,
                                     float* v_params) {
  blend_impl<float>(n_list, params, n_pix, pix_xy, bg, v_hdr, v_alpha, out_hdr, out_alpha, out_last, v_params, mode);
}
void hs_blend_tabled_r_f32(int n_list, const float* params, int n_pix, const float* pix_xy, const float* bg, const float* v_hdr,
                           const float* v_alpha, float* out_hdr, float* out_alpha, int32_t* out_last, float* v_params) {
  blend_impl<float>(n_list, params, n_pix, pix_xy, bg, v_hdr, v_alpha, out_hdr, out_alpha, out_last, v_params, 4);
}
// design study (DESIGN.md "next levers"): mode 2 = table in fp16, 3 = bf16
void hs_blend_tabled_lowp_f32(int mode, int n_list, const float* params, int n_pix, const float* pix_xy, const float* bg, const float* v_hdr,
                              const float* v_alpha, float* out_hdr, float* out_alpha, int32_t* out_last, float* v_params) {
  blend_impl<float>(n_list, params, n_pix, pix_xy, bg, v_hdr, v_alpha, out_hdr, out_alpha, out_last, v_params, mode);
}
// brute-force check helper: the culling bound for a block vs the true maximum over its pixel centres
void hs_block_bound_f32(int n, const float* params /* mx,my,A,B,C,o */, const float* rect /* x0,x1,y0,y1 */, float* bound, float* brute) {
  for (int i = 0; i < n; ++i) {
    ChsSplat<float> s;
    const float* p = params + i * 6;
    chs_make_splat(p[0], p[1], p[2], p[3], p[4], p[5], 0.f, 0.f, 0.f, s);
    const float* r = rect + i * 4;
    bound[i] = chs_block_max_power(s, r[0], r[1], r[2], r[3]);
    float best = -1e30f;
    for (float y = r[2]; y <= r[3] + 1e-3f; y += 1.f)
      for (float x = r[0]; x <= r[1] + 1e-3f; x += 1.f) {
        float dx, dy, u;
        best = fmaxf(best, chs_pair_power(s, x, y, dx, dy, u));
      }
    brute[i] = best;
  }
}
// SH colours for [C] cameras x [N] Gaussians and their backward (chs_sh.cuh), fp64
void hs_sh_fwd_bwd(int N, int C, int deg, const double* sh, const double* means, const double* viewmats, const double* v_rgb,
                   double* rgb, double* v_sh, double* v_means, double* v_viewmats /* [C,12] R|t */) {
  const int K = (deg + 1) * (deg + 1);
  std::memset(v_sh, 0, sizeof(double) * N * K * 3);
  std::memset(v_means, 0, sizeof(double) * N * 3);
  std::memset(v_viewmats, 0, sizeof(double) * C * 12);
  for (int c = 0; c < C; ++c) {
    double R[9], t[3], cp[3], v_cp[3] = {0, 0, 0};
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) R[i * 3 + j] = viewmats[c * 16 + i * 4 + j]; t[i] = viewmats[c * 16 + i * 4 + 3]; }
    chs_campos(R, t, cp);
    for (int g = 0; g < N; ++g) {
      size_t o = ((size_t)c * N + g) * 3;
      chs_sh_color(deg, sh + (size_t)g * K * 3, means + g * 3, cp, rgb + o);
      chs_sh_color_bwd(deg, sh + (size_t)g * K * 3, means + g * 3, cp, v_rgb + o, v_sh + (size_t)g * K * 3, v_means + g * 3, v_cp);
    }
    chs_campos_bwd(R, t, v_cp, v_viewmats + c * 12, v_viewmats + c * 12 + 9);
  }
}
void hs_tile_bounds(int n, const float* mx, const float* my, const int32_t* radius, int tile_w, int tile_h, int32_t* rect) {
  for (int i = 0; i < n; ++i) {
    ChsTileRect r = chs_tile_bounds(mx[i], my[i], radius[i], tile_w, tile_h);
    rect[i * 4] = r.x0; rect[i * 4 + 1] = r.y0; rect[i * 4 + 2] = r.x1; rect[i * 4 + 3] = r.y1;
  }
}
// tight bounds: packed per-axis radii and the rectangle they decode to
void hs_tight_radii(int n, const double* sxx, const double* syy, const double* opacity, const int* radius, int* packed64, int* packed32) {
  for (int i = 0; i < n; ++i) {
    packed64[i] = chs_tight_radii(sxx[i], syy[i], opacity[i], radius[i]);
    packed32[i] = chs_tight_radii((float)sxx[i], (float)syy[i], (float)opacity[i], radius[i]);
  }
}
void hs_tile_bounds_packed(int n, const float* mx, const float* my, const int* packed, int tile_w, int tile_h, int* rect) {
  for (int i = 0; i < n; ++i) {
    ChsTileRect r = chs_tile_bounds_of(mx[i], my[i], packed[i], 1, tile_w, tile_h);
    rect[i * 4] = r.x0; rect[i * 4 + 1] = r.y0; rect[i * 4 + 2] = r.x1; rect[i * 4 + 3] = r.y1;
  }
}
// CRF MLP: y and dy/dX per value, parameter gradient accumulated with weights v_y
void hs_crf_f64(int n, const double* X, const double* params, int hd, const double* v_y, double* y, double* dydx, double* v_params) {
  std::memset(v_params, 0, sizeof(double) * (3 * hd + 1));
  for (int i = 0; i < n; ++i) {
    y[i] = chs_crf_mlp_fwd(X[i], params, hd);
    dydx[i] = chs_crf_mlp_bwd(X[i], params, hd, v_y[i], v_params);
  }
}
void hs_crf_interval_f64(int n, const double* X, const double* params, int hd, const double* v_y, double* y, double* dydx, double* v_params) {
  crf_interval_impl<double>(n, X, params, hd, v_y, y, dydx, v_params);
}
void hs_crf_interval_f32(int n, const float* X, const float* params, int hd, const float* v_y, float* y, float* dydx, float* v_params) {
  crf_interval_impl<float>(n, X, params, hd, v_y, y, dydx, v_params);
}
// CRF LUT: same contract; v_params has L + 2 entries (the two range entries stay 0)
void hs_crf_lut_f64(int n, const double* X, const double* params, int L, const double* v_y, double* y, double* dydx, double* v_params) {
  std::memset(v_params, 0, sizeof(double) * (L + 2));
  for (int i = 0; i < n; ++i) {
    y[i] = chs_crf_lut_fwd(X[i], params, L);
    dydx[i] = chs_crf_lut_bwd(X[i], params, L, v_y[i], v_params);
  }
}
void hs_crf_lut_f32(int n, const float* X, const float* params, int L, float* y) {
  for (int i = 0; i < n; ++i) y[i] = chs_crf_fwd(2, X[i], params, L);
}
// spline: viewmats [C,16] fp64 and the backward contraction, mirroring chs_spline.cu
void hs_spline_fwd(int kind, const float* knots, int n_knots, double t0, double dt, const float* frame_times, const float* exposure,
                   int B, int n, double* viewmats) {
  for (int c = 0; c < B * n; ++c) {
    int i = c / n, kk = c % n;
    double w = chs_sample_weight(kk, n);
    double time = (double)frame_times[i] + w * (double)exposure[i];
    int s; double u;
    chs_spline_segment(kind, n_knots, t0, dt, time, s, u);
    int first = kind == CHS_SPLINE_LINEAR ? s : s - 1, nk = kind == CHS_SPLINE_LINEAR ? 2 : 4;
    double k[4][7];
    for (int j = 0; j < nk; ++j) for (int e = 0; e < 7; ++e) k[j][e] = knots[(first + j) * 7 + e];
    double vm[12];
    chs_spline_viewmat<double>(kind, k, u, vm);
    for (int e = 0; e < 12; ++e) viewmats[c * 16 + e] = vm[e];
    viewmats[c * 16 + 12] = viewmats[c * 16 + 13] = viewmats[c * 16 + 14] = 0; viewmats[c * 16 + 15] = 1;
  }
}
void hs_spline_bwd(int kind, const float* knots, int n_knots, double t0, double dt, const float* frame_times, const float* exposure,
                   int B, int n, const double* v_viewmats, double* v_knots, double* v_ft, double* v_ex) {
  std::memset(v_knots, 0, sizeof(double) * n_knots * 7);
  std::memset(v_ft, 0, sizeof(double) * B);
  std::memset(v_ex, 0, sizeof(double) * B);
  int nk = kind == CHS_SPLINE_LINEAR ? 2 : 4, n_in = nk * 7 + 1;
  for (int c = 0; c < B * n; ++c) {
    int i = c / n, kk = c % n;
    double w = chs_sample_weight(kk, n);
    double time = (double)frame_times[i] + w * (double)exposure[i];
    int s; double u;
    chs_spline_segment(kind, n_knots, t0, dt, time, s, u);
    int first = kind == CHS_SPLINE_LINEAR ? s : s - 1;
    for (int in = 0; in < n_in; ++in) {
      ChsDual k[4][7];
      for (int j = 0; j < nk; ++j) for (int e = 0; e < 7; ++e) k[j][e] = ChsDual(knots[(first + j) * 7 + e], (j * 7 + e == in) ? 1.0 : 0.0);
      ChsDual ud(u, in == nk * 7 ? 1.0 : 0.0);
      ChsDual vm[12];
      chs_spline_viewmat<ChsDual>(kind, k, ud, vm);
      double dot = 0;
      for (int e = 0; e < 12; ++e) dot += v_viewmats[c * 16 + e] * vm[e].d;
      if (in < nk * 7) v_knots[(first + in / 7) * 7 + in % 7] += dot;
      else { v_ft[i] += dot / dt; v_ex[i] += dot / dt * w; }
    }
  }
}
}
