// klt_track.cuh -- the 3x3 Lucas-Kanade solve with gain (klt_tracker_with_gain.cg:42-148, host loop
// v3d_gpuklt.cpp:205-305) for sm_100a.
//
// Mapping: one HALF-WARP (16 lanes) per feature slot, two slots per warp.  The (2hw+1)^2 window
// pixels are spread over the 16 lanes in ROUNDS rounds (49 pixels -> 4 rounds), the ten running
// sums are reduced with width-16 butterfly shuffles and every lane solves the 3x3 system in closed
// form (adjugate), so control flow is uniform.
//
// Two drivers share the same arithmetic (bit-identical results, checked by
// tests/test_gpu_klt.py::test_fused_gain_tracker_equals_pass_kernels):
//   klt_gain_pass   one launch per (level, iteration) pass, like the reference's draw calls
//   klt_gain_fused  ALL passes in one persistent cooperative launch:
//     * per level, the I0 window samples (constant over the iterations of a level) are computed
//       once and kept in registers, and a 12x12-texel tile of the current-frame pyramid around the
//       feature is staged in shared memory (clamped 128-bit loads; TMA box loads cannot reproduce
//       CLAMP_TO_EDGE, and at the coarse levels most windows straddle the border) -- the
//       iterations then sample from shared memory, with a global-memory fallback for samples that
//       drift outside the tile;
//     * the only coupling between slots, the gain-smoothness term that reads beta of <= 8
//       neighbour slots from the PREVIOUS pass, is synchronised without grid barriers: each slot
//       publishes (x, y, beta) of pass p into a parity double buffer followed by a version number;
//       a slot starts pass p once every slot in its wait set (neighbours + reverse neighbours) has
//       published p-1.  Waiting also on the reverse neighbours makes overwriting the p-1 record
//       (when publishing p+1) safe on non-square slot grids whose neighbour relation is not
//       symmetric.  Warps own their items for the whole launch and walk the passes in order, so the
//       least advanced item can always run (cooperative launch guarantees co-residency).
#pragma once
#include "klt_kernels.cuh"

namespace coslam {

constexpr int KLT_TW = 12;      // staged I1 tile side (texels): 2*hw + 2 + 2*margin with hw = 3
constexpr int KLT_ROUNDS = 4;   // window pixels per lane on the fast path (<= 64 pixels)

struct KltSamplePos {
  float ax, ay;
  int xi, yi;  // unclamped integer texel coordinates of the top-left tap
};

__device__ __forceinline__ KltSamplePos klt_sample_pos(int w, int h, float s, float t) {
  float u = s * (float)w - 0.5f;
  float v = t * (float)h - 0.5f;
  u = fminf(fmaxf(u, -2.0f), (float)w + 1.0f);
  v = fminf(fmaxf(v, -2.0f), (float)h + 1.0f);
  const float fu = floorf(u), fv = floorf(v);
  KltSamplePos p;
  p.ax = u - fu;
  p.ay = v - fv;
  p.xi = (int)fu;
  p.yi = (int)fv;
  return p;
}

__device__ __forceinline__ float3 klt_lerp4(const float4 p00, const float4 p10, const float4 p01,
                                            const float4 p11, float ax, float ay) {
  float3 r;
  {
    const float top = p00.x + ax * (p10.x - p00.x), bot = p01.x + ax * (p11.x - p01.x);
    r.x = top + ay * (bot - top);
  }
  {
    const float top = p00.y + ax * (p10.y - p00.y), bot = p01.y + ax * (p11.y - p01.y);
    r.y = top + ay * (bot - top);
  }
  {
    const float top = p00.z + ax * (p10.z - p00.z), bot = p01.z + ax * (p11.z - p01.z);
    r.z = top + ay * (bot - top);
  }
  return r;
}

__device__ __forceinline__ float3 klt_fetch_global(const float4* __restrict__ lv, int w, int h,
                                                   const KltSamplePos& p) {
  const int x0 = clampi(p.xi, 0, w - 1), x1 = clampi(p.xi + 1, 0, w - 1);
  const int y0 = clampi(p.yi, 0, h - 1), y1 = clampi(p.yi + 1, 0, h - 1);
  return klt_lerp4(__ldg(&lv[(size_t)y0 * w + x0]), __ldg(&lv[(size_t)y0 * w + x1]),
                   __ldg(&lv[(size_t)y1 * w + x0]), __ldg(&lv[(size_t)y1 * w + x1]), p.ax, p.ay);
}

// tile[b * KLT_TW + a] == level[clamp(ty0 + b)][clamp(tx0 + a)], so indexing it with the UNCLAMPED
// tap coordinates reproduces the clamped fetches exactly
__device__ __forceinline__ float3 klt_fetch_tile(const float4* __restrict__ tile, int tx0, int ty0,
                                                 const float4* __restrict__ lv, int w, int h,
                                                 const KltSamplePos& p) {
  const int a = p.xi - tx0, b = p.yi - ty0;
  if ((unsigned)a < (unsigned)(KLT_TW - 1) && (unsigned)b < (unsigned)(KLT_TW - 1)) {
    const float4* q = tile + b * KLT_TW + a;
    return klt_lerp4(q[0], q[1], q[KLT_TW], q[KLT_TW + 1], p.ax, p.ay);
  }
  return klt_fetch_global(lv, w, h, p);
}

__device__ __forceinline__ float half_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, 16);
  return v;
}

// accumulate one window pixel
struct KltAcc {
  float a0, a1, a2, d0, d1, d2, r0, r1, r2, ssd;
};

__device__ __forceinline__ void klt_acc_pixel(KltAcc& A, const float3 I0, const float3 I1, float beta,
                                              float nbterm, float Wf, float Hf, float lambda,
                                              float delta) {
  const float e = beta * I0.x - I1.x;
  const float Jx = (beta * I0.y + I1.y) * Wf * 0.5f;
  const float Jy = (beta * I0.z + I1.z) * Hf * 0.5f;
  const float g0 = sqrtf(I0.y * I0.y + I0.z * I0.z);
  const float g1 = sqrtf(I1.y * I1.y + I1.z * I1.z);
  A.a0 += Jx * Jx;
  A.a1 += Jx * Jy;
  A.a2 += Jx * -I0.x;
  A.d0 += Jy * Jy;
  A.d1 += Jy * -I0.x;
  A.d2 += I0.x * I0.x + lambda * g0 * g0 + delta * 8.0f;
  A.r0 += e * Jx;
  A.r1 += e * Jy;
  A.r2 += -e * I0.x + lambda * g0 * (g1 - beta * g0) + delta * nbterm;
  A.ssd += e * e;
}

// reduce over the half-warp, solve, test (klt_tracker_with_gain.cg:12-40,124-147)
__device__ __forceinline__ float4 klt_gain_finish(KltAcc A, float X1x, float X1y, float beta,
                                                  const KltTrackParams& P) {
  const float a = half_sum(A.a0), b = half_sum(A.a1), c = half_sum(A.a2);
  const float d = half_sum(A.d0), e = half_sum(A.d1), f = half_sum(A.d2);
  const float r0 = half_sum(A.r0), r1 = half_sum(A.r1), r2 = half_sum(A.r2);
  const float ssd = half_sum(A.ssd);
  float det = a * d * f + 2 * b * c * e;
  det -= a * e * e + b * b * f + c * c * d;
  const float rdet = 1.0f / det;
  const float Aa = d * f - e * e, Bb = c * e - b * f, Cc = b * e - c * d;
  const float Dd = a * f - c * c, Ee = b * c - a * e, Ff = a * d - b * b;
  float ux = (Aa * r0 + Bb * r1 + Cc * r2) * rdet;
  float uy = (Bb * r0 + Dd * r1 + Ee * r2) * rdet;
  const float ub = (Cc * r0 + Ee * r1 + Ff * r2) * rdet;
  X1x += ux;
  X1y += uy;
  ux *= (float)P.W;
  uy *= (float)P.H;
  const float sqrLen = ux * ux + uy * uy;
  bool invalid = (det < 0.00001f);
  invalid = invalid || (ssd > P.ssdThr);
  invalid = invalid || (sqrLen > P.sqrConv);
  invalid = invalid || (X1x < P.vr0 || X1y < P.vr1) || (X1x > P.vr2 || X1y > P.vr3);
  return invalid ? make_float4(-1.f, -1.f, -1.f, 0.f) : make_float4(X1x, X1y, beta + ub, 0.f);
}

// dot(float4(1), betaN1 + betaN2 - 2*beta) of klt_tracker_with_gain.cg:111 from the eight
// neighbour gains held by half-lanes 0..7 (bn); invalid (< 0) neighbours count as own beta
__device__ __forceinline__ float klt_nbterm(float bn, float beta, int hl, int halfBase) {
  bn = (bn < 0.f) ? beta : bn;
  const float hi = __shfl_sync(0xffffffffu, bn, halfBase + (hl & 3) + 4);
  const float s4 = bn + hi - 2.0f * beta;
  const float s0 = __shfl_sync(0xffffffffu, s4, halfBase + 0);
  const float s1 = __shfl_sync(0xffffffffu, s4, halfBase + 1);
  const float s2 = __shfl_sync(0xffffffffu, s4, halfBase + 2);
  const float s3 = __shfl_sync(0xffffffffu, s4, halfBase + 3);
  return ((s0 + s1) + s2) + s3;
}

// ------------------------------------------------------------------------------------------
// One pass per launch (diagnostic / fallback).  Grid: x = ceil(F / 16) blocks of 256 threads
// (16 half-warps), y = camera.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
klt_gain_pass(const float4* __restrict__ pyr0, const float4* __restrict__ pyr1, long long pyrStride,
              long long lvOff, int w, int h, const float4* __restrict__ X0buf,
              const float4* __restrict__ in, float4* __restrict__ out,
              const int* __restrict__ nbr, float dsx, float dsy, KltTrackParams P, int firstPass) {
  const int cam = blockIdx.y;
  const int hl = threadIdx.x & 15, halfBase = threadIdx.x & 16;
  int slot = blockIdx.x * (blockDim.x >> 4) + (threadIdx.x >> 4);
  const bool active = slot < P.F;
  if (!active) slot = P.F - 1;  // keep the whole warp in the shuffles; result is discarded
  const float4* L0 = pyr0 + (size_t)cam * pyrStride + lvOff;
  const float4* L1 = pyr1 + (size_t)cam * pyrStride + lvOff;
  const size_t fb = (size_t)cam * P.F;
  const float4 x0 = X0buf[fb + slot];
  float4 cur = in[fb + slot];
  if (firstPass) cur.z = 1.0f;  // gain cleared to 1 before the first pass (v3d_gpuklt.cpp:223-227)
  const float beta = cur.z;
  float bn = 0.f;
  if (hl < 8) bn = firstPass ? 1.0f : in[fb + nbr[slot * 8 + hl]].z;
  const float nbterm = klt_nbterm(bn, beta, hl, halfBase);
  const bool pre_invalid = (cur.x < 0.f) || (x0.x < 0.f);
  const int hw = P.halfWidth, fwid = 2 * hw + 1, npx = fwid * fwid;
  const float Wf = (float)P.W, Hf = (float)P.H;
  KltAcc A = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int p = hl; p < npx; p += 16) {
    const int py = p / fwid, px = p - py * fwid;
    const float fx = (float)(px - hw), fy = (float)(py - hw);
    const float3 I0 = klt_fetch_global(L0, w, h, klt_sample_pos(w, h, x0.x + fx * dsx, x0.y + fy * dsy));
    const float3 I1 = klt_fetch_global(L1, w, h, klt_sample_pos(w, h, cur.x + fx * dsx, cur.y + fy * dsy));
    klt_acc_pixel(A, I0, I1, beta, nbterm, Wf, Hf, P.lambda, P.delta);
  }
  float4 res = klt_gain_finish(A, cur.x, cur.y, beta, P);
  if (pre_invalid) res = make_float4(-1.f, -1.f, -1.f, 0.f);
  if (hl == 0 && active) out[fb + slot] = res;
}

// ------------------------------------------------------------------------------------------
// All passes in one persistent cooperative launch.
//   rec[2][T] u64 = (pass number << 32) | float_bits(beta): value and version travel in ONE naturally
//   atomic 8-byte word, so publishing needs no fence and a reader gets beta with the poll itself;
//   state[T] float4 holds (x, y, beta) of items that are not register-resident
//   waitset[F][16]: first 8 = neighbours (values + versions), last 8 = reverse neighbours or -1
// Work item = (camera, slot); half-warp q (global index) owns items q, q + Q, ...; the first item
// of a half-warp uses the shared-memory tile, further items (only when T > Q) the global path.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

constexpr int KLT_FUSED_THREADS = 128;  // 8 half-warps per CTA

__global__ void __launch_bounds__(KLT_FUSED_THREADS, 7)
klt_gain_fused(const float4* __restrict__ pyr0, const float4* __restrict__ pyr1,
               long long pyrStride, KltLevels LV, int nIter, const float4* __restrict__ X0buf,
               float4* __restrict__ state, unsigned long long* __restrict__ rec,
               const int* __restrict__ waitset,
               float4* __restrict__ out, int C, KltTrackParams Plax, KltTrackParams Pstrict,
               int verBase) {
  __shared__ float4 s_tile[KLT_FUSED_THREADS / 16][KLT_TW * KLT_TW];  // one tile per half-warp
  const int hl = threadIdx.x & 15, halfBase = threadIdx.x & 16;
  const int halfInBlock = threadIdx.x >> 4;
  const int q = blockIdx.x * (KLT_FUSED_THREADS / 16) + halfInBlock;
  const int Q = gridDim.x * (KLT_FUSED_THREADS / 16);
  const int F = Plax.F;
  const int T = C * F;
  const int hw = Plax.halfWidth, fwid = 2 * hw + 1, npx = fwid * fwid;
  const bool fast = (npx <= 16 * KLT_ROUNDS) && (2 * hw + 2 + 2 <= KLT_TW);
  const float Wf = (float)Plax.W, Hf = (float)Plax.H;
  float4* tile = s_tile[halfInBlock];
  // both halves of a warp must execute the same number of outer iterations (full-mask shuffles)
  const int nOwn = (T + Q - 1) / Q;

  // state of the FIRST owned item lives in registers across passes
  float3 I0r[KLT_ROUNDS];
  // window offsets of this lane's pixels, hoisted out of every loop (no integer division inside)
  float fxr[KLT_ROUNDS], fyr[KLT_ROUNDS];
#pragma unroll
  for (int r = 0; r < KLT_ROUNDS; ++r) {
    const int p = min(hl + 16 * r, npx - 1);
    const int py = p / fwid, px = p - py * fwid;
    fxr[r] = (float)(px - hw);
    fyr[r] = (float)(py - hw);
  }
  float4 cur0 = make_float4(-1.f, -1.f, -1.f, 0.f);
  int tx0 = 0, ty0 = 0;

  int pass = 0;
  for (int li = 0; li < LV.n; ++li) {
    const int w = LV.w[li], h = LV.h[li];
    const float dsx = 1.0f / (float)w, dsy = 1.0f / (float)h;
    for (int it = 1; it <= nIter; ++it) {
      ++pass;
      // thresholds are lax except on the last iteration of a level (v3d_gpuklt.cpp:266-279)
      const bool strict = (it == nIter) && (it != 1);
      const int rd = (pass - 1) & 1, wr = pass & 1;
      for (int own = 0; own < nOwn; ++own) {
        int item = q + own * Q;
        const bool active = item < T;
        if (!active) item = T - 1;
        const int cam = item / F, slot = item - cam * F;
        const float4* L0 = pyr0 + (size_t)cam * pyrStride + LV.off[li];
        const float4* L1 = pyr1 + (size_t)cam * pyrStride + LV.off[li];
        const float4 x0 = X0buf[item];
        const bool staged = fast && (own == 0);
        float4 cur;
        float bn = 0.f;
        if (pass == 1) {
          cur = make_float4(x0.x, x0.y, 1.0f, 0.f);  // X1 <- X0, gain cleared to 1 (:223-227)
          if (hl < 8) bn = 1.0f;
        } else {
          // wait for the pass-(p-1) records of everything this slot reads or is read by; the
          // first 8 entries also deliver the neighbour gains
          if (active) {
            const int nb = waitset[slot * 16 + hl];
            if (nb >= 0) {
              const unsigned long long* rp = rec + (size_t)rd * T + (size_t)cam * F + nb;
              const int need = verBase + pass - 1;
              unsigned long long v = ld_relaxed_u64(rp);
              while ((int)(v >> 32) < need) {
                __nanosleep(20);
                v = ld_relaxed_u64(rp);
              }
              bn = __uint_as_float((unsigned)v);
            }
          }
          __syncwarp();
          cur = (staged) ? cur0 : __ldcg(&state[item]);
        }
        const float beta = cur.z;
        const float nbterm = klt_nbterm(bn, beta, hl, halfBase);
        const bool pre_invalid = (cur.x < 0.f) || (x0.x < 0.f);
        // ---- per-level staging (first iteration of a level): I0 samples + I1 tile
        if (staged && it == 1) {
#pragma unroll
          for (int r = 0; r < KLT_ROUNDS; ++r) {
            const int p = hl + 16 * r;
            const float fx = fxr[r], fy = fyr[r];
            I0r[r] = (p < npx && !pre_invalid)
                         ? klt_fetch_global(L0, w, h, klt_sample_pos(w, h, x0.x + fx * dsx, x0.y + fy * dsy))
                         : make_float3(0.f, 0.f, 0.f);
          }
          const KltSamplePos c = klt_sample_pos(w, h, cur.x, cur.y);
          tx0 = c.xi - hw - 2;
          ty0 = c.yi - hw - 2;
          if (!pre_invalid) {
            for (int i = hl; i < KLT_TW * KLT_TW; i += 16) {
              const int b = i / KLT_TW, a = i - b * KLT_TW;
              tile[i] = __ldg(&L1[(size_t)clampi(ty0 + b, 0, h - 1) * w + clampi(tx0 + a, 0, w - 1)]);
            }
          }
          __syncwarp();
        }
        // ---- the iteration
        KltAcc A = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (!pre_invalid) {
          if (staged) {
#pragma unroll
            for (int r = 0; r < KLT_ROUNDS; ++r) {
              const int p = hl + 16 * r;
              if (p < npx) {
                const float fx = fxr[r], fy = fyr[r];
                const float3 I1 = klt_fetch_tile(tile, tx0, ty0, L1, w, h,
                                                 klt_sample_pos(w, h, cur.x + fx * dsx, cur.y + fy * dsy));
                klt_acc_pixel(A, I0r[r], I1, beta, nbterm, Wf, Hf, Plax.lambda, Plax.delta);
              }
            }
          } else {
            for (int p = hl; p < npx; p += 16) {
              const int py = p / fwid, px = p - py * fwid;
              const float fx = (float)(px - hw), fy = (float)(py - hw);
              const float3 I0 = klt_fetch_global(L0, w, h, klt_sample_pos(w, h, x0.x + fx * dsx, x0.y + fy * dsy));
              const float3 I1 = klt_fetch_global(L1, w, h, klt_sample_pos(w, h, cur.x + fx * dsx, cur.y + fy * dsy));
              klt_acc_pixel(A, I0, I1, beta, nbterm, Wf, Hf, Plax.lambda, Plax.delta);
            }
          }
        }
        float4 res = klt_gain_finish(A, cur.x, cur.y, beta, strict ? Pstrict : Plax);
        if (pre_invalid) res = make_float4(-1.f, -1.f, -1.f, 0.f);
        if (staged) cur0 = res;
        if (hl == 0 && active) {
          if (pass == LV.n * nIter) out[item] = res;
          if (!staged) __stcg(&state[item], res);
          st_relaxed_u64(rec + (size_t)wr * T + item,
                         ((unsigned long long)(unsigned)(verBase + pass) << 32) |
                             (unsigned long long)__float_as_uint(res.z));
        }
        __syncwarp();
      }
    }
  }
}

}  // namespace coslam
