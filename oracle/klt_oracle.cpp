/*
 * klt_oracle.cpp -- CPU restatement of the reference's pyramidal KLT tracker (TEST INFRASTRUCTURE,
 * see oracle.h).  PARITY UNPINNED (no reference tests/fixtures exist; the reference path is Cg
 * shaders over OpenGL and has no CPU implementation).
 *
 * Follows, function by function (paths relative to /root/reference/src/tracking/CGKLT):
 *   build_pyramid        Shaders/pyramid_with_derivative_pass1v.cg:63-83 (PRESMOOTHING==1),
 *                        Shaders/pyramid_with_derivative_pass1h.cg:96-128,
 *                        Shaders/pyramid_with_derivative_pass2.cg:1-15,
 *                        v3d_gpupyramid.cpp:16-52 (tap offsets), :366-429 (pass order, level sizes)
 *   sample               GL_LINEAR + GL_CLAMP_TO_EDGE, v3d_gpuklt.cpp:129-141,231-241
 *   track_2x2            Shaders/klt_tracker.cg:24-132, uniforms v3d_gpuklt.cpp:143-147
 *   track_gain           Shaders/klt_tracker_with_gain.cg:42-148, host loop v3d_gpuklt.cpp:205-305
 *   cornerness/nonmax    Shaders/klt_detector_pass1.cg, klt_detector_pass2.cg, klt_detector_nonmax.cg,
 *                        host v3d_gpuklt.cpp:423-545
 *   detect/redetect/track/feed/advance   v3d_gpuklt.cpp:650-889, v3d_gpuklt.h:252-259
 *
 * Conventions fixed where GL semantics are ambiguous or hardware defined ([choice] in SURVEY.md
 * Appendix A): fp32 pyramid storage (reference: RGB16F), exact fp32 two-stage lerp (hardware: 8-bit
 * weights), centred [1 3 3 1] taps {2j-1..2j+2}, candidates = all survivors, strongest-first with
 * the deterministic order (cornerness desc, y asc, x asc), dead slots provided as (-1,-1), feature
 * double buffer with copy semantics on advance.
 *
 * Compiled with -ffp-contract=off: every fp32 operation below is an individually rounded IEEE op.
 */
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

static int g_threads = 1;
extern "C" void orc_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
extern "C" int orc_get_max_threads(void) {
#ifdef _OPENMP
  return omp_get_num_procs();
#else
  return 1;
#endif
}
int orc_threads() { return g_threads; }

namespace {

struct Level {
  int w = 0, h = 0;
  std::vector<float> d;  // interleaved (I, Ix, Iy)
};

struct Cand {
  float x, y, c;  // normalised pixel-centre position, cornerness
  int px, py;
};

}  // namespace

struct orc_klt {
  cosl_klt_config cfg;
  int W, H, L, fw, fh, F, plCap;
  int levelSkip, halfWidth;
  float trackMargin, conv, ssd, detectMargin;
  std::vector<Level> pyr[2];
  int cur;                        // index of pyr1 ("current"); pyr0 = pyr[1-cur]
  std::vector<float> src, dst;    // F x (x, y, gain): features for the next track / last provided
  std::vector<float> corn, tmp;   // W*H cornerness work buffers
  std::vector<float> conv3;       // W*H*3 vertical structure-tensor sums
  std::vector<int> nbr;           // F x 8 neighbour slots for the gain smoothness term
  int lastNumCand = 0;
  bool havePrev = false;
};

namespace {

inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ---------------------------------------------------------------- pyramid */
void build_pyramid(orc_klt* h, const uint8_t* img, size_t pitch, std::vector<Level>& P) {
  const int W = h->W, H = h->H;
  // pass1v: vertical [1 2 1]/4 and [-1 -2 0 2 1]/8 on the 0..255 image (pass1v.cg:63-83)
  std::vector<float> v((size_t)W * H), dv((size_t)W * H);
#pragma omp parallel for num_threads(g_threads) schedule(static)
  for (int y = 0; y < H; ++y) {
    const uint8_t* rm2 = img + (size_t)clampi(y - 2, 0, H - 1) * pitch;
    const uint8_t* rm1 = img + (size_t)clampi(y - 1, 0, H - 1) * pitch;
    const uint8_t* r0 = img + (size_t)y * pitch;
    const uint8_t* rp1 = img + (size_t)clampi(y + 1, 0, H - 1) * pitch;
    const uint8_t* rp2 = img + (size_t)clampi(y + 2, 0, H - 1) * pitch;
    for (int x = 0; x < W; ++x) {
      float gm2 = rm2[x], gm1 = rm1[x], g0 = r0[x], gp1 = rp1[x], gp2 = rp2[x];
      // dot(f1,g1): 0*g(-2) + .25 g(-1) + .5 g(0) + .25 g(+1)
      float vv = 0.25f * gm1;
      vv = vv + 0.5f * g0;
      vv = vv + 0.25f * gp1;
      // dot(df1,g1) + df2*g2: -1/8 g(-2) - 2/8 g(-1) + 0 g(0) + 2/8 g(+1) + 1/8 g(+2)
      float dd = -0.125f * gm2;
      dd = dd + -0.25f * gm1;
      dd = dd + 0.25f * gp1;
      dd = dd + 0.125f * gp2;
      v[(size_t)y * W + x] = vv;
      dv[(size_t)y * W + x] = dd;
    }
  }
  // pass1h: I = smooth_h(v), Ix = deriv_h(v), Iy = smooth_h(dv)   (pass1h.cg:96-128)
  Level& L0 = P[0];
#pragma omp parallel for num_threads(g_threads) schedule(static)
  for (int y = 0; y < H; ++y) {
    const float* rv = &v[(size_t)y * W];
    const float* rd = &dv[(size_t)y * W];
    for (int x = 0; x < W; ++x) {
      int xm2 = clampi(x - 2, 0, W - 1), xm1 = clampi(x - 1, 0, W - 1);
      int xp1 = clampi(x + 1, 0, W - 1), xp2 = clampi(x + 2, 0, W - 1);
      float I = 0.25f * rv[xm1];
      I = I + 0.5f * rv[x];
      I = I + 0.25f * rv[xp1];
      float Ix = -0.125f * rv[xm2];
      Ix = Ix + -0.25f * rv[xm1];
      Ix = Ix + 0.25f * rv[xp1];
      Ix = Ix + 0.125f * rv[xp2];
      float Iy = 0.25f * rd[xm1];
      Iy = Iy + 0.5f * rd[x];
      Iy = Iy + 0.25f * rd[xp1];
      float* o = &L0.d[((size_t)y * W + x) * 3];
      o[0] = I;
      o[1] = Ix;
      o[2] = Iy;
    }
  }
  // pass2: [1 3 3 1]/8 vertical then horizontal, decimate by 2, all three channels
  // (pass2.cg:1-15; host v3d_gpupyramid.cpp:402-420; centred taps)
  for (int l = 1; l < h->L; ++l) {
    const Level& S = P[l - 1];
    Level& D = P[l];
    const int sw = S.w, sh = S.h, dw = D.w, dh = D.h;
    std::vector<float> t((size_t)sw * dh * 3);
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int j = 0; j < dh; ++j) {
      const float* a = &S.d[(size_t)clampi(2 * j - 1, 0, sh - 1) * sw * 3];
      const float* b = &S.d[(size_t)clampi(2 * j, 0, sh - 1) * sw * 3];
      const float* c = &S.d[(size_t)clampi(2 * j + 1, 0, sh - 1) * sw * 3];
      const float* d = &S.d[(size_t)clampi(2 * j + 2, 0, sh - 1) * sw * 3];
      float* o = &t[(size_t)j * sw * 3];
      for (int k = 0; k < sw * 3; ++k) {
        float r = a[k] + 3.0f * b[k];
        r = r + 3.0f * c[k];
        r = r + d[k];
        o[k] = r / 8.0f;
      }
    }
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int j = 0; j < dh; ++j) {
      const float* r = &t[(size_t)j * sw * 3];
      float* o = &D.d[(size_t)j * dw * 3];
      for (int i = 0; i < dw; ++i) {
        int xa = clampi(2 * i - 1, 0, sw - 1), xb = clampi(2 * i, 0, sw - 1);
        int xc = clampi(2 * i + 1, 0, sw - 1), xd = clampi(2 * i + 2, 0, sw - 1);
        for (int ch = 0; ch < 3; ++ch) {
          float q = r[xa * 3 + ch] + 3.0f * r[xb * 3 + ch];
          q = q + 3.0f * r[xc * 3 + ch];
          q = q + r[xd * 3 + ch];
          o[i * 3 + ch] = q / 8.0f;
        }
      }
    }
  }
}

/* ---------------------------------------------------------------- bilinear sampling (A.3) */
inline void sample(const Level& lv, float s, float t, float out[3]) {
  float u = s * (float)lv.w - 0.5f;
  float v = t * (float)lv.h - 0.5f;
  // keep the cast defined for wild/NaN positions; exact for all finite positions because
  // everything left of -1 / right of w clamps to the edge texel anyway
  u = fminf(fmaxf(u, -2.0f), (float)lv.w + 1.0f);
  v = fminf(fmaxf(v, -2.0f), (float)lv.h + 1.0f);
  float fu = floorf(u), fv = floorf(v);
  float ax = u - fu, ay = v - fv;
  int x0 = clampi((int)fu, 0, lv.w - 1), x1 = clampi((int)fu + 1, 0, lv.w - 1);
  int y0 = clampi((int)fv, 0, lv.h - 1), y1 = clampi((int)fv + 1, 0, lv.h - 1);
  const float* p00 = &lv.d[((size_t)y0 * lv.w + x0) * 3];
  const float* p10 = &lv.d[((size_t)y0 * lv.w + x1) * 3];
  const float* p01 = &lv.d[((size_t)y1 * lv.w + x0) * 3];
  const float* p11 = &lv.d[((size_t)y1 * lv.w + x1) * 3];
  for (int c = 0; c < 3; ++c) {
    float top = p00[c] + ax * (p10[c] - p00[c]);
    float bot = p01[c] + ax * (p11[c] - p01[c]);
    out[c] = top + ay * (bot - top);
  }
}

/* ---------------------------------------------------------------- 2x2 LK (klt_tracker.cg) */
void track_2x2(orc_klt* h, const std::vector<Level>& P0, const std::vector<Level>& P1,
               const float* src, float* out) {
  const int F = h->F, hw = h->halfWidth, L = h->L, skip = h->levelSkip;
  const float Wf = (float)h->W, Hf = (float)h->H;
  const float dsx = 1.0f / Wf, dsy = 1.0f / Hf;  // v3d_gpuklt.cpp:125-126
  const int nIter = (h->cfg.compat & COSL_KLT_COMPAT_ITER5) ? 5 : h->cfg.nIterations;
  const float sqrConv = h->conv * h->conv, ssdThr = h->ssd;
  const float m = h->trackMargin;
  const float vr[4] = {m / Wf, m / Hf, 1.0f - m / Wf, 1.0f - m / Hf};  // :147
#pragma omp parallel for num_threads(g_threads) schedule(dynamic, 16)
  for (int i = 0; i < F; ++i) {
    const float X0x = src[3 * i], X0y = src[3 * i + 1];
    float X1x = X0x, X1y = X0y;
    bool invalid = (X1x < 0);
    float sqrLen = 0, SSD = 0;
    float mult = (float)(1 << (L - 1));
    for (int level = L - 1; level >= 0; level -= skip) {
      const float dx = dsx * mult, dy = dsy * mult;
      for (int it = 0; it < nIter; ++it) {
        float a = 0, b = 0, c = 0, r0 = 0, r1 = 0;
        SSD = 0;
        for (int y = -hw; y <= hw; ++y) {
          float sy0 = X0y + (float)y * dy, sy1 = X1y + (float)y * dy;
          for (int x = -hw; x <= hw; ++x) {
            float sx0 = X0x + (float)x * dx, sx1 = X1x + (float)x * dx;
            float I0[3], I1[3];
            sample(P0[level], sx0, sy0, I0);
            sample(P1[level], sx1, sy1, I1);
            float e = I0[0] - I1[0];
            float Jx = (I0[1] + I1[1]) * Wf / 2;
            float Jy = (I0[2] + I1[2]) * Hf / 2;
            a += Jx * Jx;
            b += Jx * Jy;
            c += Jy * Jy;
            r0 += e * Jx;
            r1 += e * Jy;
            SSD += e * e;
          }
        }
        float det = a * c - b * b;
        invalid = invalid || (det < 0.00001f);
        float rdet = 1.0f / det;
        float ux = rdet * (c * r0 - b * r1);
        float uy = rdet * (-b * r0 + a * r1);
        X1x += ux;
        X1y += uy;
        ux *= Wf;
        uy *= Hf;
        sqrLen = ux * ux + uy * uy;
      }
      invalid = invalid || (sqrLen > sqrConv);
      invalid = invalid || (SSD > ssdThr);
      mult /= (float)(1 << skip);
    }
    invalid = invalid || (X1x < vr[0] || X1y < vr[1]) || (X1x > vr[2] || X1y > vr[3]);
    if (invalid) {
      out[3 * i] = out[3 * i + 1] = out[3 * i + 2] = -1.0f;
    } else {
      out[3 * i] = X1x;
      out[3 * i + 1] = X1y;
      out[3 * i + 2] = X0x;
    }
  }
}

/* ---------------------------------------------------------------- 3x3 LK with gain */
inline float det3(const float abc[3], const float def[3]) {
  const float a = abc[0], b = abc[1], c = abc[2], d = def[0], e = def[1], f = def[2];
  float res = a * d * f + 2 * b * c * e;
  res -= a * e * e + b * b * f + c * c * d;
  return res;
}

// one pass over all slots (one GL draw of klt_tracker_with_gain.cg)
void gain_pass(orc_klt* h, const Level& L0, const Level& L1, const float* X0buf, const float* in,
               float* out, float dsx, float dsy, float sqrConv, float ssdThr, const float vr[4],
               float lambda, float delta) {
  const int F = h->F, hw = h->halfWidth;
  const float Wf = (float)h->W, Hf = (float)h->H;
#pragma omp parallel for num_threads(g_threads) schedule(dynamic, 16)
  for (int i = 0; i < F; ++i) {
    const float X0x = X0buf[3 * i], X0y = X0buf[3 * i + 1];
    float X1x = in[3 * i], X1y = in[3 * i + 1];
    const float beta = in[3 * i + 2];
    float bn[8];
    for (int k = 0; k < 8; ++k) {
      float b = in[3 * h->nbr[8 * i + k] + 2];
      bn[k] = (b < 0) ? beta : b;
    }
    // dot(float4(1), betaN1 + betaN2 - 2*beta)   (klt_tracker_with_gain.cg:111)
    float s4[4];
    for (int k = 0; k < 4; ++k) s4[k] = bn[k] + bn[4 + k] - 2 * beta;
    const float nbterm = ((s4[0] + s4[1]) + s4[2]) + s4[3];
    bool invalid = (X1x < 0) || (X0x < 0);
    float abc[3] = {0, 0, 0}, def[3] = {0, 0, 0}, rhs[3] = {0, 0, 0};
    float SSD = 0;
    for (int y = -hw; y <= hw; ++y) {
      float sy0 = X0y + (float)y * dsy, sy1 = X1y + (float)y * dsy;
      for (int x = -hw; x <= hw; ++x) {
        float sx0 = X0x + (float)x * dsx, sx1 = X1x + (float)x * dsx;
        float I0[3], I1[3];
        sample(L0, sx0, sy0, I0);
        sample(L1, sx1, sy1, I1);
        float e = beta * I0[0] - I1[0];
        float Jx = (beta * I0[1] + I1[1]) * Wf / 2;
        float Jy = (beta * I0[2] + I1[2]) * Hf / 2;
        float g0 = sqrtf(I0[1] * I0[1] + I0[2] * I0[2]);
        float g1 = sqrtf(I1[1] * I1[1] + I1[2] * I1[2]);
        abc[0] += Jx * Jx;
        abc[1] += Jx * Jy;
        abc[2] += Jx * -I0[0];
        def[0] += Jy * Jy;
        def[1] += Jy * -I0[0];
        def[2] += I0[0] * I0[0] + lambda * g0 * g0 + delta * 8;
        rhs[0] += e * Jx;
        rhs[1] += e * Jy;
        rhs[2] += -e * I0[0] + lambda * g0 * (g1 - beta * g0) + delta * nbterm;
        SSD += e * e;
      }
    }
    const float det = det3(abc, def);
    const float rdet = 1.0f / det;
    const float a = abc[0], b = abc[1], c = abc[2], d = def[0], e = def[1], f = def[2];
    const float A = d * f - e * e, B = c * e - b * f, C = b * e - c * d;
    const float D = a * f - c * c, E = b * c - a * e, Fq = a * d - b * b;
    float ux = (A * rhs[0] + B * rhs[1] + C * rhs[2]) * rdet;
    float uy = (B * rhs[0] + D * rhs[1] + E * rhs[2]) * rdet;
    float ub = (C * rhs[0] + E * rhs[1] + Fq * rhs[2]) * rdet;
    X1x += ux;
    X1y += uy;
    ux *= Wf;
    uy *= Hf;
    const float sqrLen = ux * ux + uy * uy;
    invalid = invalid || (det < 0.00001f);
    invalid = invalid || (SSD > ssdThr);
    invalid = invalid || (sqrLen > sqrConv);
    invalid = invalid || (X1x < vr[0] || X1y < vr[1]) || (X1x > vr[2] || X1y > vr[3]);
    if (invalid) {
      out[3 * i] = out[3 * i + 1] = out[3 * i + 2] = -1.0f;
    } else {
      out[3 * i] = X1x;
      out[3 * i + 1] = X1y;
      out[3 * i + 2] = beta + ub;
    }
  }
}

void track_gain(orc_klt* h, const std::vector<Level>& P0, const std::vector<Level>& P1,
                const float* src, float* out) {
  const int F = h->F;
  const float Wf = (float)h->W, Hf = (float)h->H;
  std::vector<float> A(src, src + 3 * F), B(3 * (size_t)F);
  for (int i = 0; i < F; ++i) A[3 * i + 2] = 1.0f;  // gain cleared to 1 (v3d_gpuklt.cpp:223-227)
  const float lax[4] = {-1.0f, -1.0f, 2.0f, 2.0f};
  const float m = h->trackMargin;
  const float strict[4] = {m / Wf, m / Hf, 1.0f - m / Wf, 1.0f - m / Hf};
  float delta = 200.0f;  // :243 (tau = 1 -> constant)
  float sqrConv = 1000000.0f, ssdThr = 1000000.0f;
  const float* vr = lax;
  float* in = A.data();
  float* o = B.data();
  for (int level = h->L - 1; level >= 0; level -= h->levelSkip) {
    const int w = h->W >> level, ht = h->H >> level;
    const float dsx = 1.0f / (float)w, dsy = 1.0f / (float)ht;
    for (int iter = 1; iter <= h->cfg.nIterations; ++iter) {
      if (iter == 1) {  // :266-270
        sqrConv = 1000000.0f;
        ssdThr = 1000000.0f;
        vr = lax;
      } else if (iter == h->cfg.nIterations) {  // :271-279
        sqrConv = h->conv * h->conv;
        ssdThr = h->ssd;
        vr = strict;
      }
      gain_pass(h, P0[level], P1[level], src, in, o, dsx, dsy, sqrConv, ssdThr, vr, 1.0f, delta);
      std::swap(in, o);
    }
  }
  std::memcpy(out, in, sizeof(float) * 3 * F);
}

/* ---------------------------------------------------------------- detector */
// cornerness map incl. margin mask (pass1 + pass2), present suppression, separable non-max.
// Result: h->corn holds the final signed map (value > 0 == surviving corner).
void detect_corners(orc_klt* h, const Level& L0, int nPresent, const float* present3,
                    std::vector<Cand>& cands) {
  const int W = h->W, H = h->H;
  const float minC = h->cfg.minCornerness;
  const float mg = h->detectMargin;
  const float Wf = (float)W, Hf = (float)H;
  const float vr[4] = {mg / Wf, mg / Hf, 1.0f - mg / Wf, 1.0f - mg / Hf};
  std::vector<float>& conv = h->conv3;
  // pass1: vertical 7-tap sums of (Ix^2, IxIy, Iy^2), taps -3..+3 in order (klt_detector_pass1.cg)
#pragma omp parallel for num_threads(g_threads) schedule(static)
  for (int y = 0; y < H; ++y) {
    for (int x = 0; x < W; ++x) {
      float r0 = 0, r1 = 0, r2 = 0;
      for (int k = -3; k <= 3; ++k) {
        const float* p = &L0.d[((size_t)clampi(y + k, 0, H - 1) * W + x) * 3];
        r0 += p[1] * p[1];
        r1 += p[1] * p[2];
        r2 += p[2] * p[2];
      }
      float* o = &conv[((size_t)y * W + x) * 3];
      o[0] = r0;
      o[1] = r1;
      o[2] = r2;
    }
  }
  // pass2: horizontal 7-tap sums, min-eigenvalue minus threshold, margin (klt_detector_pass2.cg)
  std::vector<float>& corn = h->corn;
#pragma omp parallel for num_threads(g_threads) schedule(static)
  for (int y = 0; y < H; ++y) {
    for (int x = 0; x < W; ++x) {
      float a = 0, b = 0, c = 0;
      for (int k = -3; k <= 3; ++k) {
        const float* p = &conv[((size_t)y * W + clampi(x + k, 0, W - 1)) * 3];
        a += p[0];
        b += p[1];
        c += p[2];
      }
      float amc = a - c;
      float cs = 0.5f * (a + c - sqrtf(amc * amc + 4 * (b * b)));
      cs = fmaxf(cs - minC, 0.0f);
      float sx = ((float)x + 0.5f) / Wf, sy = ((float)y + 0.5f) / Hf;
      bool inside = (sx >= vr[0] && sy >= vr[1]) && (sx <= vr[2] && sy <= vr[3]);
      corn[(size_t)y * W + x] = inside ? cs : 0.0f;
    }
  }
  // suppress around present tracks: -1e30 at the pixel containing each live point (:475-500)
  for (int i = 0; i < nPresent; ++i) {
    float px = present3[3 * i], py = present3[3 * i + 1];
    if (!(px >= 0.0f && px < 1.0f && py >= 0.0f && py < 1.0f)) continue;  // clipped by GL
    int ix = (int)floorf(px * Wf), iy = (int)floorf(py * Hf);
    if (ix >= 0 && ix < W && iy >= 0 && iy < H) corn[(size_t)iy * W + ix] = -1e30f;
  }
  // separable non-max, horizontal then vertical (klt_detector_nonmax.cg:12-26)
  const int r = h->cfg.minDistance;
  std::vector<float>& t = h->tmp;
#pragma omp parallel for num_threads(g_threads) schedule(static)
  for (int y = 0; y < H; ++y) {
    const float* row = &corn[(size_t)y * W];
    for (int x = 0; x < W; ++x) {
      float mx = row[x];
      for (int i = -r; i < 0; ++i) {
        float cc = fabsf(row[clampi(x + i, 0, W - 1)]);
        mx = (cc >= fabsf(mx)) ? -cc : mx;
      }
      for (int i = 1; i <= r; ++i) {
        float cc = fabsf(row[clampi(x + i, 0, W - 1)]);
        mx = (cc >= fabsf(mx)) ? -cc : mx;
      }
      t[(size_t)y * W + x] = mx;
    }
  }
  std::vector<float> fin((size_t)W * H);
#pragma omp parallel for num_threads(g_threads) schedule(static)
  for (int y = 0; y < H; ++y) {
    for (int x = 0; x < W; ++x) {
      float mx = t[(size_t)y * W + x];
      for (int i = -r; i < 0; ++i) {
        float cc = fabsf(t[(size_t)clampi(y + i, 0, H - 1) * W + x]);
        mx = (cc >= fabsf(mx)) ? -cc : mx;
      }
      for (int i = 1; i <= r; ++i) {
        float cc = fabsf(t[(size_t)clampi(y + i, 0, H - 1) * W + x]);
        mx = (cc >= fabsf(mx)) ? -cc : mx;
      }
      fin[(size_t)y * W + x] = mx;
    }
  }
  cands.clear();
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      float v = fin[(size_t)y * W + x];
      if (v > 0) {
        Cand c;
        c.px = x;
        c.py = y;
        c.c = v;
        c.x = ((float)x + 0.5f) / Wf;  // traverse_histpyr.cg:84-86 -> pixel centre
        c.y = ((float)y + 0.5f) / Hf;
        cands.push_back(c);
      }
    }
  // deterministic strongest-first order [choice]
  std::sort(cands.begin(), cands.end(), [](const Cand& a, const Cand& b) {
    if (a.c != b.c) return a.c > b.c;
    if (a.py != b.py) return a.py < b.py;
    return a.px < b.px;
  });
  h->lastNumCand = (int)cands.size();
}

// COSL_KLT_COMPAT_HISTOPYR: the candidate list the reference actually reads back.
//  * the discriminator pass renders W/2 x H/2 texels of 2x2 pixels (v3d_gpuklt.cpp:519-523): a last odd
//    row / column never becomes a candidate;
//  * extractCorners walks the HistoPyramid top-down visiting the children (x,y), (x+1,y), (x,y+1),
//    (x+1,y+1) in that order (klt_detector_traverse_histpyr.cg:33-50; base-level taps in the same
//    order, v3d_gpuklt.cpp:42-57): candidate k of the list is the k-th pixel in MORTON order;
//  * only the first pointListWidth*pointListHeight entries are read back (:660-661, :755-757);
//  * they are ranked by cornerness only if they exceed the free slots (:662-665, :761-768) -- the
//    reference's std::nth_element keeps the same SET and leaves the order unspecified, the oracle
//    (and the CUDA path) order that case strongest-first; otherwise extraction order is kept.
unsigned part1by1(unsigned v) {
  v &= 0xffffu;
  v = (v | (v << 8)) & 0x00ff00ffu;
  v = (v | (v << 4)) & 0x0f0f0f0fu;
  v = (v | (v << 2)) & 0x33333333u;
  v = (v | (v << 1)) & 0x55555555u;
  return v;
}
void histopyr_list(const orc_klt* h, std::vector<Cand>& cands, int freeSlots) {
  if (!(h->cfg.compat & COSL_KLT_COMPAT_HISTOPYR)) return;
  const int Wc = 2 * (h->W / 2), Hc = 2 * (h->H / 2);
  std::vector<Cand> in;
  for (const Cand& c : cands)
    if (c.px < Wc && c.py < Hc) in.push_back(c);
  auto morton = [](const Cand& c) { return part1by1((unsigned)c.px) | (part1by1((unsigned)c.py) << 1); };
  std::sort(in.begin(), in.end(), [&](const Cand& a, const Cand& b) { return morton(a) < morton(b); });
  if ((int)in.size() > h->plCap) in.resize(h->plCap);
  if ((int)in.size() > freeSlots)
    std::sort(in.begin(), in.end(), [](const Cand& a, const Cand& b) {
      if (a.c != b.c) return a.c > b.c;
      if (a.py != b.py) return a.py < b.py;
      return a.px < b.px;
    });
  cands.swap(in);
}

void run_tracker(orc_klt* h, float* out) {
  const std::vector<Level>& P0 = h->pyr[1 - h->cur];
  const std::vector<Level>& P1 = h->pyr[h->cur];
  if (h->cfg.trackWithGain)
    track_gain(h, P0, P1, h->src.data(), out);
  else
    track_2x2(h, P0, P1, h->src.data(), out);
}

}  // namespace

/* ================================================================ C API */
extern "C" {

orc_klt* orc_klt_create(const cosl_klt_config* cfg, int width, int height, int nLevels, int featW,
                        int featH, int plW, int plH) {
  if (!cfg || width < 8 || height < 8 || nLevels < 1 || nLevels > 8 || featW < 1 || featH < 1)
    return nullptr;
  orc_klt* h = new orc_klt();
  h->cfg = *cfg;
  h->cfg.nLevels = nLevels;
  h->W = width;
  h->H = height;
  h->L = nLevels;
  h->fw = featW;
  h->fh = featH;
  h->F = featW * featH;
  if (plW <= 0) plW = 2 * featW;
  if (plH <= 0) plH = 2 * featH;
  h->plCap = plW * plH;
  h->levelSkip = cfg->levelSkip > 0 ? cfg->levelSkip : (nLevels - 1);  // v3d_gpuklt.h:14
  if (h->levelSkip < 1) h->levelSkip = 1;
  h->halfWidth = cfg->windowWidth / 2;  // v3d_gpuklt.cpp:114,212
  h->trackMargin = cfg->trackBorderMargin;
  h->conv = cfg->convergenceThreshold;
  h->ssd = cfg->SSD_Threshold;
  h->detectMargin = 10.0f;  // KLT_Detector::_margin default, v3d_gpuklt.h:113-114
  for (int b = 0; b < 2; ++b) {
    h->pyr[b].resize(nLevels);
    for (int l = 0; l < nLevels; ++l) {
      h->pyr[b][l].w = width >> l;
      h->pyr[b][l].h = height >> l;
      h->pyr[b][l].d.assign((size_t)(width >> l) * (height >> l) * 3, 0.0f);
    }
  }
  h->cur = 1;
  h->src.assign((size_t)h->F * 3, -1.0f);
  h->dst.assign((size_t)h->F * 3, -1.0f);
  h->corn.assign((size_t)width * height, 0.0f);
  h->tmp.assign((size_t)width * height, 0.0f);
  h->conv3.assign((size_t)width * height * 3, 0.0f);
  // neighbour slots of the gain smoothness term (klt_tracker_with_gain.cg:64-75): NEAREST,
  // CLAMP_TO_EDGE lookups in the featW x featH slot texture at st0 +- ds0.  betaN1 adds the SCALAR
  // ds0.x (resp. ds0.y) to both coordinates.
  h->nbr.resize((size_t)h->F * 8);
  const double fw = featW, fh = featH;
  for (int sy = 0; sy < featH; ++sy)
    for (int sx = 0; sx < featW; ++sx) {
      const double cx = sx + 0.5, cy = sy + 0.5;  // in slot units
      // offsets in slot units: ds0.x = 1/fw normalised = 1 col = fh/fw rows; ds0.y = 1/fh = fw/fh cols = 1 row
      const double off[8][2] = {{+1.0, +fh / fw}, {-1.0, -fh / fw}, {+fw / fh, +1.0}, {-fw / fh, -1.0},
                                {+1.0, 0.0},      {-1.0, 0.0},      {0.0, +1.0},      {0.0, -1.0}};
      for (int k = 0; k < 8; ++k) {
        int nx = clampi((int)std::floor(cx + off[k][0]), 0, featW - 1);
        int ny = clampi((int)std::floor(cy + off[k][1]), 0, featH - 1);
        h->nbr[(size_t)(sy * featW + sx) * 8 + k] = ny * featW + nx;
      }
    }
  return h;
}

void orc_klt_destroy(orc_klt* h) { delete h; }
void orc_klt_set_margin(orc_klt* h, float m) {  // v3d_gpuklt.h:217-224: tracker AND detector
  h->trackMargin = m;
  h->detectMargin = m;
}
void orc_klt_set_conv(orc_klt* h, float t) { h->conv = t; }
void orc_klt_set_ssd(orc_klt* h, float t) { h->ssd = t; }
int orc_klt_last_num_candidates(orc_klt* h) { return h->lastNumCand; }

// KLT_SequenceTracker::detect, both overloads (v3d_gpuklt.cpp:651-738)
int orc_klt_detect(orc_klt* h, const uint8_t* img, size_t pitch, int nPresent,
                   const float* present3, cosl_klt_feature* dest, int* nDetected) {
  const int F = h->F;
  if (nPresent < 0 || nPresent > F) return COSL_E_INVALID;
  build_pyramid(h, img, pitch, h->pyr[h->cur]);
  std::vector<Cand> cands;
  detect_corners(h, h->pyr[h->cur][0], nPresent, present3, cands);
  histopyr_list(h, cands, F - nPresent);
  int nDet = std::min((int)cands.size(), h->plCap);
  nDet = std::min(nDet, F - nPresent);
  for (int i = 0; i < F; ++i) {
    h->dst[3 * i] = h->dst[3 * i + 1] = -1.0f;
    h->dst[3 * i + 2] = 1.0f;
    dest[i].status = -1;
    dest[i].fed = -1;
    dest[i].pos[0] = dest[i].pos[1] = -1.0f;
    dest[i].gain = 1.0f;
  }
  for (int i = 0; i < nDet; ++i) {
    h->dst[3 * i] = cands[i].x;
    h->dst[3 * i + 1] = cands[i].y;
    // the 2x2 path provides the cornerness in the third channel, the gain path 1.0 (:668-673)
    h->dst[3 * i + 2] = h->cfg.trackWithGain ? 1.0f : cands[i].c;
    dest[i].status = 1;
    dest[i].pos[0] = cands[i].x;
    dest[i].pos[1] = cands[i].y;
    dest[i].gain = h->dst[3 * i + 2];
    dest[i].fed = -1;
  }
  for (int i = 0; i < nPresent; ++i) {
    int s = nDet + i;
    h->dst[3 * s] = present3[3 * i];
    h->dst[3 * s + 1] = present3[3 * i + 1];
    h->dst[3 * s + 2] = 1.0f;
    dest[s].status = 1;
    dest[s].pos[0] = present3[3 * i];
    dest[s].pos[1] = present3[3 * i + 1];
    dest[s].gain = 1.0f;
    dest[s].fed = i;
  }
  *nDetected = nDet + nPresent;
  return COSL_OK;
}

// KLT_SequenceTracker::track (v3d_gpuklt.cpp:857-889)
int orc_klt_track(orc_klt* h, const uint8_t* img, size_t pitch, cosl_klt_feature* dest,
                  int* nPresent) {
  const int F = h->F;
  build_pyramid(h, img, pitch, h->pyr[h->cur]);
  std::vector<float> res((size_t)F * 3);
  run_tracker(h, res.data());
  int n = 0;
  for (int i = 0; i < F; ++i) {
    float X = res[3 * i], Y = res[3 * i + 1], g = res[3 * i + 2];
    if (X >= 0) {
      dest[i].status = 0;
      dest[i].pos[0] = X;
      dest[i].pos[1] = Y;
      dest[i].gain = g;
      dest[i].fed = -1;
      ++n;
    } else {
      dest[i].status = -1;
      dest[i].fed = -1;
      dest[i].pos[0] = dest[i].pos[1] = -1.0f;
      dest[i].gain = 1.0f;
    }
  }
  h->dst = res;
  *nPresent = n;
  return COSL_OK;
}

// KLT_SequenceTracker::redetect (v3d_gpuklt.cpp:740-805)
int orc_klt_redetect(orc_klt* h, const uint8_t* img, size_t pitch, cosl_klt_feature* dest,
                     int* nNewFeatures) {
  const int F = h->F;
  int nPresent = 0;
  orc_klt_track(h, img, pitch, dest, &nPresent);
  std::vector<float> present((size_t)F * 3);
  for (int i = 0; i < F; ++i) {
    if (dest[i].status >= 0) {
      present[3 * i] = dest[i].pos[0];
      present[3 * i + 1] = dest[i].pos[1];
    } else {
      present[3 * i] = present[3 * i + 1] = -1.0f;
    }
    present[3 * i + 2] = 0;
  }
  std::vector<Cand> cands;
  detect_corners(h, h->pyr[h->cur][0], F, present.data(), cands);
  histopyr_list(h, cands, F - nPresent);
  int nNew = std::min((int)cands.size(), h->plCap);
  nNew = std::min(nNew, F - nPresent);
  int k = 0;
  for (int i = 0; i < F && k < nNew; ++i) {
    if (dest[i].status < 0) {
      dest[i].status = 1;
      dest[i].pos[0] = cands[k].x;
      dest[i].pos[1] = cands[k].y;
      dest[i].gain = cands[k].c;  // :780 copies data[2] (= cornerness) into gain
      dest[i].fed = -1;
      ++k;
    }
  }
  for (int i = 0; i < F; ++i) {
    if (dest[i].status >= 0) {
      h->dst[3 * i] = dest[i].pos[0];
      h->dst[3 * i + 1] = dest[i].pos[1];
    } else {
      h->dst[3 * i] = h->dst[3 * i + 1] = -1.0f;
    }
    h->dst[3 * i + 2] = 1.0f;
  }
  *nNewFeatures = k + nPresent;
  return COSL_OK;
}

// KLT_SequenceTracker::feedExternFeaturePoints (v3d_gpuklt.cpp:808-855): stride 3 throughout by
// default; with COSL_KLT_COMPAT_FEED_STRIDE2 the proximity test reads featPts[2k], featPts[2k+1]
// exactly as :826-827 do.
int orc_klt_feed(orc_klt* h, int npts, const float* pts3, int* trackIds, int* nFed) {
  const int F = h->F;
  const int st = (h->cfg.compat & COSL_KLT_COMPAT_FEED_STRIDE2) ? 2 : 3;
  const double radius2 = 1e-4;
  std::vector<float>& c = h->dst;
  for (int k = 0; k < npts; ++k)
    for (int i = 0; i < F; ++i) {
      if (c[3 * i] < 0) continue;
      double dx = (double)(pts3[st * k] - c[3 * i]);
      double dy = (double)(pts3[st * k + 1] - c[3 * i + 1]);
      if (dx * dx + dy * dy < radius2) c[3 * i] = -1.0f;
    }
  int k = 0;
  for (int i = 0; i < F && k < npts; ++i) {
    if (c[3 * i] < 0) {
      c[3 * i] = pts3[3 * k];
      c[3 * i + 1] = pts3[3 * k + 1];
      c[3 * i + 2] = 1.0f;
      trackIds[k] = i;
      ++k;
    }
  }
  *nFed = k;
  return COSL_OK;
}

// KLT_SequenceTracker::advanceFrame (v3d_gpuklt.h:252-259)
int orc_klt_advance(orc_klt* h) {
  h->src = h->dst;
  h->cur = 1 - h->cur;
  return COSL_OK;
}

int orc_klt_debug_pyramid(orc_klt* h, int which, int level, float* out3, int* w, int* ht) {
  if (level < 0 || level >= h->L) return COSL_E_INVALID;
  const Level& lv = h->pyr[which ? h->cur : 1 - h->cur][level];
  if (w) *w = lv.w;
  if (ht) *ht = lv.h;
  if (out3) std::memcpy(out3, lv.d.data(), lv.d.size() * sizeof(float));
  return COSL_OK;
}

int orc_klt_debug_cornerness(orc_klt* h, float* out) {
  std::memcpy(out, h->corn.data(), h->corn.size() * sizeof(float));
  return COSL_OK;
}

}  // extern "C"
