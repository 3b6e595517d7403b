// klt_front.cuh -- fused front end of a KLT frame: from the u8 image, ONE kernel produces
//   * pyramid level 0 (I, Ix, Iy)          [pyramid_with_derivative_pass1v.cg + pass1h.cg]
//   * pyramid level 1 ([1 3 3 1]^2 / 64)   [pyramid_with_derivative_pass2.cg]
//   * the detector's cornerness map        [klt_detector_pass1.cg + klt_detector_pass2.cg]
// so the 16 B/px level-0 texels are written once and never re-read from HBM by the pyramid or the
// detector (the three separate kernels re-read them twice: 32 B/px of avoidable traffic).
// Every output is computed with exactly the same non-fused operations, in the same order, as the
// stand-alone kernels (klt_pyr_level0, klt_pyr_down, klt_cornerness) and the oracle: bit-exact.
//
// Tile: 64 x 32 level-0 pixels per 256-thread CTA.  Level 0 is evaluated on the tile plus a halo
// of 3 (needed by the 7x7 structure tensor; the level-1 taps need only 1), from a u8 tile with a
// halo of 5.  All tiles are indexed by UNCLAMPED coordinates and hold the values of the CLAMPED
// coordinates, which reproduces CLAMP_TO_EDGE of every pass.
#pragma once
#include "klt_kernels.cuh"

namespace coslam {

constexpr int FR_TW = 64, FR_TH = 32;
constexpr int FR_LW = FR_TW + 6, FR_LH = FR_TH + 6;    // level-0 tile incl. halo 3
constexpr int FR_UW = FR_TW + 10, FR_UH = FR_TH + 10;  // u8 tile incl. halo 5
constexpr int FR_UWP = FR_UW + 2;                      // padded row (bytes), multiple of 4

struct FrontSmem {
  float l0[3][FR_LH][FR_LW];       // I, Ix, Iy planes
  union {
    struct {
      unsigned char u8[FR_UH][FR_UWP];
      float v[FR_LH][FR_UW];       // vertical smooth   (rows of the level-0 tile, all u8 columns)
      float dv[FR_LH][FR_UW];      // vertical derivative
    } a;
    struct {
      float st[3][FR_TH][FR_LW];   // vertical 7-sums of Ix^2, IxIy, Iy^2
      float t1[3][FR_TH / 2][FR_TW + 2];  // level-1 vertical pass
    } b;
  } u;
};

__global__ void __launch_bounds__(256)
klt_front(const uint8_t* __restrict__ img, size_t imgPitch, size_t imgStride,
          float4* __restrict__ pyr, long long pyrStride, long long lv1Off, float* __restrict__ corn,
          int W, int H, int wantL1, int wantCorn, float minC, float vr0, float vr1, float vr2,
          float vr3) {
  extern __shared__ unsigned char s_raw[];
  FrontSmem& S = *reinterpret_cast<FrontSmem*>(s_raw);
  const int cam = blockIdx.z;
  const int x0 = blockIdx.x * FR_TW, y0 = blockIdx.y * FR_TH;
  const uint8_t* im = img + (size_t)cam * imgStride;
  const int tid = threadIdx.x;
  // ---- u8 tile, halo 5.  Interior tiles fetch aligned 32-bit words (x0 is a multiple of 64 and
  // the pitch a multiple of 16), all loads of a thread in flight before the first shared store;
  // tiles that touch the left/right image border clamp per byte.
  if (x0 - 8 >= 0 && x0 + FR_TW + 8 <= W) {
    constexpr int WPR = (FR_TW + 16) / 4;  // 20 words per row cover columns x0-8 .. x0+71
    unsigned int wv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = tid + 256 * u;
      if (i < FR_UH * WPR) {
        const int uy = i / WPR, wx = i - uy * WPR;
        const int gy = clampi(y0 - 5 + uy, 0, H - 1);
        wv[u] = __ldg(reinterpret_cast<const unsigned int*>(im + (size_t)gy * imgPitch + (x0 - 8)) + wx);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = tid + 256 * u;
      if (i < FR_UH * WPR) {
        const int uy = i / WPR, wx = i - uy * WPR;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int ux = 4 * wx + b - 3;  // column x0-8+4wx+b  ->  tile index relative to x0-5
          if (ux >= 0 && ux < FR_UW) S.u.a.u8[uy][ux] = (unsigned char)((wv[u] >> (8 * b)) & 0xff);
        }
      }
    }
  } else {
    for (int i = tid; i < FR_UH * FR_UW; i += 256) {
      const int uy = i / FR_UW, ux = i - uy * FR_UW;
      const int gy = clampi(y0 - 5 + uy, 0, H - 1), gx = clampi(x0 - 5 + ux, 0, W - 1);
      S.u.a.u8[uy][ux] = im[(size_t)gy * imgPitch + gx];
    }
  }
  __syncthreads();
  // ---- vertical pass of level 0 at the rows of the level-0 tile (clamped row coordinate)
  for (int i = tid; i < FR_LH * FR_UW; i += 256) {
    const int ty = i / FR_UW, ux = i - ty * FR_UW;
    const int cgy = clampi(y0 - 3 + ty, 0, H - 1);
    const int uy = cgy - (y0 - 5);  // u8-tile row of the clamped coordinate; +-2 stay inside
    const float gm2 = S.u.a.u8[uy - 2][ux], gm1 = S.u.a.u8[uy - 1][ux], g0 = S.u.a.u8[uy][ux];
    const float gp1 = S.u.a.u8[uy + 1][ux], gp2 = S.u.a.u8[uy + 2][ux];
    float vv = __fmul_rn(0.25f, gm1);
    vv = __fadd_rn(vv, __fmul_rn(0.5f, g0));
    vv = __fadd_rn(vv, __fmul_rn(0.25f, gp1));
    float dd = __fmul_rn(-0.125f, gm2);
    dd = __fadd_rn(dd, __fmul_rn(-0.25f, gm1));
    dd = __fadd_rn(dd, __fmul_rn(0.25f, gp1));
    dd = __fadd_rn(dd, __fmul_rn(0.125f, gp2));
    S.u.a.v[ty][ux] = vv;
    S.u.a.dv[ty][ux] = dd;
  }
  __syncthreads();
  // ---- horizontal pass -> level-0 tile (halo 3), and the store of the interior
  float4* out0 = pyr + (size_t)cam * pyrStride;
  for (int i = tid; i < FR_LH * FR_LW; i += 256) {
    const int ty = i / FR_LW, tx = i - ty * FR_LW;
    const int gx = x0 - 3 + tx, gy = y0 - 3 + ty;
    const int cgx = clampi(gx, 0, W - 1);
    const int ux = cgx - (x0 - 5);
    const float* rv = &S.u.a.v[ty][ux];
    const float* rd = &S.u.a.dv[ty][ux];
    float I = __fmul_rn(0.25f, rv[-1]);
    I = __fadd_rn(I, __fmul_rn(0.5f, rv[0]));
    I = __fadd_rn(I, __fmul_rn(0.25f, rv[1]));
    float Ix = __fmul_rn(-0.125f, rv[-2]);
    Ix = __fadd_rn(Ix, __fmul_rn(-0.25f, rv[-1]));
    Ix = __fadd_rn(Ix, __fmul_rn(0.25f, rv[1]));
    Ix = __fadd_rn(Ix, __fmul_rn(0.125f, rv[2]));
    float Iy = __fmul_rn(0.25f, rd[-1]);
    Iy = __fadd_rn(Iy, __fmul_rn(0.5f, rd[0]));
    Iy = __fadd_rn(Iy, __fmul_rn(0.25f, rd[1]));
    S.l0[0][ty][tx] = I;
    S.l0[1][ty][tx] = Ix;
    S.l0[2][ty][tx] = Iy;
    if (tx >= 3 && tx < 3 + FR_TW && ty >= 3 && ty < 3 + FR_TH && gx < W && gy < H)
      out0[(size_t)gy * W + gx] = make_float4(I, Ix, Iy, 0.0f);
  }
  __syncthreads();  // u.a is dead from here on, u.b may be written
  // ---- level 1: vertical [1 3 3 1] at rows 2j-1..2j+2, all tile columns -1..+64
  const int dw = W >> 1, dh = H >> 1;
  if (wantL1) {
    for (int i = tid; i < 3 * (FR_TH / 2) * (FR_TW + 2); i += 256) {
      const int ch = i / ((FR_TH / 2) * (FR_TW + 2));
      const int rem = i - ch * ((FR_TH / 2) * (FR_TW + 2));
      const int jj = rem / (FR_TW + 2), cx = rem - jj * (FR_TW + 2);  // column x0 - 1 + cx
      const int ty = 2 * jj - 1 + 3;  // tile row of source row 2j - 1 (j = y0/2 + jj)
      const int tx = cx - 1 + 3;
      S.u.b.t1[ch][jj][cx] = tap1331(S.l0[ch][ty][tx], S.l0[ch][ty + 1][tx], S.l0[ch][ty + 2][tx],
                                     S.l0[ch][ty + 3][tx]);
    }
  }
  // ---- cornerness: vertical 7-sums of the gradient products (rows of the interior, halo-3 columns)
  if (wantCorn) {
    for (int i = tid; i < FR_TH * FR_LW; i += 256) {
      const int ty = i / FR_LW, tx = i - ty * FR_LW;
      float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        const float gx = S.l0[1][ty + k][tx], gy = S.l0[2][ty + k][tx];
        r0 = __fadd_rn(r0, __fmul_rn(gx, gx));
        r1 = __fadd_rn(r1, __fmul_rn(gx, gy));
        r2 = __fadd_rn(r2, __fmul_rn(gy, gy));
      }
      S.u.b.st[0][ty][tx] = r0;
      S.u.b.st[1][ty][tx] = r1;
      S.u.b.st[2][ty][tx] = r2;
    }
  }
  __syncthreads();
  if (wantL1) {
    float4* out1 = pyr + (size_t)cam * pyrStride + lv1Off;
    for (int i = tid; i < (FR_TH / 2) * (FR_TW / 2); i += 256) {
      const int jj = i / (FR_TW / 2), ii = i - jj * (FR_TW / 2);
      const int gi = x0 / 2 + ii, gj = y0 / 2 + jj;
      if (gi < dw && gj < dh) {
        float r[3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const float* t = &S.u.b.t1[ch][jj][2 * ii];  // columns 2i-1 .. 2i+2 -> cx = 2ii .. 2ii+3
          r[ch] = tap1331(t[0], t[1], t[2], t[3]);
        }
        out1[(size_t)gj * dw + gi] = make_float4(r[0], r[1], r[2], 0.0f);
      }
    }
  }
  if (wantCorn) {
    float* outc = corn + (size_t)cam * W * H;
    const float Wf = (float)W, Hf = (float)H;
    for (int i = tid; i < FR_TH * FR_TW; i += 256) {
      const int ty = i / FR_TW, tx = i - ty * FR_TW;
      const int gx = x0 + tx, gy = y0 + ty;
      if (gx >= W || gy >= H) continue;
      float a = 0.f, b = 0.f, c = 0.f;
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        a = __fadd_rn(a, S.u.b.st[0][ty][tx + k]);
        b = __fadd_rn(b, S.u.b.st[1][ty][tx + k]);
        c = __fadd_rn(c, S.u.b.st[2][ty][tx + k]);
      }
      const float amc = __fsub_rn(a, c);
      const float rad = __fadd_rn(__fmul_rn(amc, amc), __fmul_rn(4.0f, __fmul_rn(b, b)));
      float cs = __fmul_rn(0.5f, __fsub_rn(__fadd_rn(a, c), __fsqrt_rn(rad)));
      cs = fmaxf(__fsub_rn(cs, minC), 0.0f);
      const float sx = __fdiv_rn(__fadd_rn((float)gx, 0.5f), Wf);
      const float sy = __fdiv_rn(__fadd_rn((float)gy, 0.5f), Hf);
      const bool inside = (sx >= vr0 && sy >= vr1) && (sx <= vr2 && sy <= vr3);
      outc[(size_t)gy * W + gx] = inside ? cs : 0.0f;
    }
  }
}

}  // namespace coslam
