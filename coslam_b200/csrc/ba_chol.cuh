// ba_chol.cuh -- dense / skyline Cholesky solve of the reduced camera system (fp64, sm_100a).
//
// Storage: S is column-major LOWER with leading dimension ld >= ns + 1: L(i, j), i >= j, lives at
// S[j * ld + i].  Row index ns of every column holds the right-hand side, i.e. the matrix is the
// (ns+1) x ns lower trapezoid [S ; b^T]: factorising it performs the forward substitution for free
// (after the last panel, row ns holds y^T = (L^-1 b)^T).
//
// Blocked right-looking algorithm with block size CB = 64 and a block-row ENVELOPE: block row I has
// structural non-zeros only from block column firstBlk[I] on (camera co-visibility is banded for
// sequential key frames), so panel k touches block rows (k, lastBlk[k]] plus the block that holds
// the right-hand-side row.  With dense co-visibility the envelope is full and this is the plain
// dense algorithm.  Per panel:
//   ba_chol_potf2_inv : factor the 64x64 diagonal block and invert the factor (one CTA, 1024 thr)
//   ba_chol_trsm      : X = A_panel * Linv_k^T          (a GEMM, no triangular dependency chain)
//   ba_chol_syrk      : A_IJ -= X_I X_J^T on the active tiles
// then ba_chol_backward: x = L^-T y, one persistent CTA walking the block columns backwards.
// The whole sequence is captured once into a CUDA graph per solver (fixed structure).
#pragma once
#include <cuda_runtime.h>

namespace coslam {

constexpr int CB = 64;
constexpr int BA_CHOL_SMEM = 2 * CB * (CB + 1) * (int)sizeof(double);

// ------------------------------------------------------------------------------------------
// Small systems: ns*ns doubles fit in shared memory -> factor + both substitutions in one CTA.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ba_chol_small(double* __restrict__ S, int ld, int ns, double* __restrict__ xout,
              double* __restrict__ sc, int scFail) {
  extern __shared__ double sL[];  // column-major lower, leading dimension ns
  __shared__ int s_fail;
  __shared__ double sx[1024];
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int t = tid; t < ns * ns; t += nt) {
    const int j = t / ns, i = t - j * ns;
    sL[t] = (i >= j) ? S[(size_t)j * ld + i] : 0.0;
  }
  if (tid == 0) s_fail = 0;
  __syncthreads();
  for (int k = 0; k < ns; ++k) {
    const double dkk = sL[k * ns + k];
    if (!(dkk > 0) || !isfinite(dkk)) {
      if (tid == 0) s_fail = 1;
      break;  // uniform: every thread reads the same dkk
    }
    const double rs = 1.0 / sqrt(dkk);
    __syncthreads();
    for (int i = k + tid; i < ns; i += nt) sL[k * ns + i] = (i == k) ? sqrt(dkk) : sL[k * ns + i] * rs;
    __syncthreads();
    // trailing update: L(i, j) -= L(i,k) L(j,k), k < j <= i; thread grid 16 (j) x 16 (i)
    const int tj = tid >> 4, ti = tid & 15;
    for (int j = k + 1 + tj; j < ns; j += 16) {
      const double ljk = sL[k * ns + j];
      for (int i = j + ti; i < ns; i += 16) sL[j * ns + i] -= sL[k * ns + i] * ljk;
    }
    __syncthreads();
  }
  __syncthreads();
  if (s_fail) {
    if (tid == 0) sc[scFail] = 1.0;
    return;
  }
  if (tid < 32) {
    for (int i = tid; i < ns; i += 32) sx[i] = S[(size_t)i * ld + ns];
    __syncwarp();
    for (int k = 0; k < ns; ++k) {
      const double xk = sx[k] / sL[k * ns + k];
      __syncwarp();
      if (tid == 0) sx[k] = xk;
      for (int i = k + 1 + tid; i < ns; i += 32) sx[i] -= sL[k * ns + i] * xk;
      __syncwarp();
    }
    for (int k = ns - 1; k >= 0; --k) {
      double part = 0;
      for (int i = k + 1 + tid; i < ns; i += 32) part += sL[k * ns + i] * sx[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      __syncwarp();
      if (tid == 0) sx[k] = (sx[k] - part) / sL[k * ns + k];
      __syncwarp();
    }
    for (int i = tid; i < ns; i += 32) xout[i] = sx[i];
  }
}

// ------------------------------------------------------------------------------------------
// Diagonal block: Cholesky factor (written back into S) and its inverse Linv (dense, row-major
// Linv[r * CB + c], lower triangular).  256 threads, register resident:
//   factor : thread (i = tid & 63, g = tid >> 6) owns row i, columns j == g (mod 4) in 16
//            registers; per pivot the raw column goes through a double-buffered shared vector,
//            every thread scales it itself (one rsqrt per thread, no second barrier) and applies
//            the rank-1 update to its registers: ONE barrier per pivot, all loops fully unrolled
//            so the register arrays are statically indexed;
//   inverse: thread (c = tid >> 2, gg = tid & 3) owns X[p][c], p == gg (mod 4); row by row forward
//            substitution with a 4-lane shuffle reduction.
// ------------------------------------------------------------------------------------------
// 1/sqrt(d) for the pivots: single-precision MUFU seed + two Newton steps in fp64 (full double
// accuracy for d inside the float range, which Gauss-Newton diagonals always are).  The library
// rsqrt() is ~3x longer, and this value sits on the pivot-to-pivot critical path of the whole solve.
__device__ __forceinline__ double ba_rsqrt(double d) {
  if (d < 1e-30 || d > 1e30) return rsqrt(d);  // outside the float range of the seed
  double y = (double)rsqrtf((float)d);
  const double h = 0.5 * d;
  y = __fma_rn(y, __fma_rn(-h * y, y, 0.5), y);
  y = __fma_rn(y, __fma_rn(-h * y, y, 0.5), y);
  return y;
}

__global__ void __launch_bounds__(256)
ba_chol_potf2_inv(double* __restrict__ S, int ld, int k0, int bs, double* __restrict__ Linv,
                  double* __restrict__ sc, int scFail) {
  __shared__ double colraw[2][CB];
  __shared__ double rdiag[CB];
  __shared__ double Lm[CB][CB + 1];
  __shared__ int s_fail;
  const int tid = threadIdx.x;
  if (tid == 0) s_fail = 0;
  {
    const int i = tid & 63, g = tid >> 6;
    double a[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int j = g + 4 * q;
      a[q] = (i < bs && j < bs && j <= i) ? S[(size_t)(k0 + j) * ld + k0 + i] : ((i == j) ? 1.0 : 0.0);
    }
#pragma unroll
    for (int k = 0; k < CB; ++k) {
      const int qk = k >> 2, gk = k & 3, buf = k & 1;
      if (g == gk) colraw[buf][i] = a[qk];  // (a[] stays in registers: static indices only)
      __syncthreads();
      double d = colraw[buf][k];
      if (!(d > 0) || !isfinite(d)) {
        if (tid == 0) s_fail = 1;
        d = 1.0;
      }
      const double rs = ba_rsqrt(d);
      const double li = colraw[buf][i] * rs;  // L(i, k) for i > k; sqrt(d) for i == k
      if (g == gk) Lm[i][k] = (i >= k) ? li : 0.0;
      if (tid == k) rdiag[k] = rs;
#pragma unroll
      for (int q = qk; q < 16; ++q) {
        const int j = g + 4 * q;
        if (j > k && j <= i) a[q] -= li * (colraw[buf][j] * rs);
      }
    }
  }
  __syncthreads();
  // write the factor back (coalesced along the rows of a column)
  for (int t = tid; t < CB * CB; t += 256) {
    const int j = t >> 6, i = t & 63;
    if (i < bs && j < bs && i >= j) S[(size_t)(k0 + j) * ld + k0 + i] = Lm[i][j];
  }
  if (s_fail && tid == 0) sc[scFail] = 1.0;
  {
    const int c = tid >> 2, gg = tid & 3;
    double x[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) x[q] = 0.0;
#pragma unroll
    for (int r = 0; r < CB; ++r) {
      // four independent partial sums: the dot product is on the row-to-row critical path
      double sp[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int q = 0; q <= ((r - 1) >> 2); ++q) {
        const int p = gg + 4 * q;
        if (r > 0 && p < r) sp[q & 3] = __fma_rn(Lm[r][p], x[q], sp[q & 3]);
      }
      double s = (sp[0] + sp[1]) + (sp[2] + sp[3]);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      const double xr = (((r == c) ? 1.0 : 0.0) - s) * rdiag[r];
      x[r >> 2] = (gg == (r & 3)) ? ((r >= c) ? xr : 0.0) : x[r >> 2];
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int r = gg + 4 * q;
      Linv[r * CB + c] = (r < bs && c < bs) ? x[q] : 0.0;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Panel: rows [r0, r0 + 64) of block column k:  X(r, c) = sum_{p <= c} A(r, p) Linv(c, p).
// blockIdx.x < nAct addresses active block row k + 1 + blockIdx.x, blockIdx.x == nAct the extra
// block row `extraBlk` (the one holding the right-hand-side row).  nrows = ns + 1.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ba_chol_trsm(double* __restrict__ S, int ld, int nrows, int k0, int bs, int kblk, int nAct,
             int extraBlk, const double* __restrict__ Linv) {
  const int rowMin = k0 + bs;  // rows of the diagonal block itself are never touched
  extern __shared__ double s_dyn[];
  double (*sA)[CB + 1] = reinterpret_cast<double (*)[CB + 1]>(s_dyn);                  // [p][row]
  double (*sI)[CB + 1] = reinterpret_cast<double (*)[CB + 1]>(s_dyn + CB * (CB + 1));  // [p][c]
  const int blk = (blockIdx.x < nAct) ? (kblk + 1 + blockIdx.x) : extraBlk;
  const int r0 = blk * CB;
  const int tid = threadIdx.x;
  {
    // stage with all 32 global loads in flight before the first shared store (one L2 round trip)
    double ra[16], ri[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int t = tid + 256 * u;
      const int p = t >> 6, r = t & 63;
      ra[u] = (p < bs && r0 + r < nrows && r0 + r >= rowMin) ? S[(size_t)(k0 + p) * ld + r0 + r] : 0.0;
      ri[u] = Linv[t];  // Linv(c = t >> 6, pp = t & 63)
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int t = tid + 256 * u;
      sA[t >> 6][t & 63] = ra[u];
      sI[t & 63][t >> 6] = ri[u];
    }
  }
  __syncthreads();
  const int tx = tid & 15, ty = tid >> 4;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0;
  for (int p = 0; p < bs; ++p) {
    double ar[4], ic[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) ar[a] = sA[p][tx + 16 * a];
#pragma unroll
    for (int b = 0; b < 4; ++b) ic[b] = sI[p][ty + 16 * b];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] += ar[a] * ic[b];
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int r = r0 + tx + 16 * a, c = ty + 16 * b;
      if (r < nrows && r >= rowMin && c < bs) S[(size_t)(k0 + c) * ld + r] = acc[a][b];
    }
}

// ------------------------------------------------------------------------------------------
// Trailing update on the active tiles: A(i, j) -= sum_p X(i, p) X(j, p), i >= j, j < ns.
// Tile rows/cols index the list {k+1 .. k+nAct, extraBlk}; grid = (nT, nT), lower tiles only.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ba_chol_syrk(double* __restrict__ S, int ld, int ns, int nrows, int k0, int bs, int kblk,
             int nAct, int extraBlk) {
  const int tjI = blockIdx.x, tiI = blockIdx.y;
  if (tiI < tjI) return;
  const int bi = (tiI < nAct) ? (kblk + 1 + tiI) : extraBlk;
  const int bj = (tjI < nAct) ? (kblk + 1 + tjI) : extraBlk;
  extern __shared__ double s_dyn[];
  double (*sAi)[CB + 1] = reinterpret_cast<double (*)[CB + 1]>(s_dyn);                  // [p][row]
  double (*sAj)[CB + 1] = reinterpret_cast<double (*)[CB + 1]>(s_dyn + CB * (CB + 1));
  const int i0 = bi * CB, j0 = bj * CB;
  const int rowMin = k0 + bs;  // only the trailing part (rows/cols beyond the panel) is updated
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  // prefetch the C tile (independent of the panel loads -> one memory round trip for everything)
  double cold[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = i0 + tx + 16 * a, j = j0 + ty + 16 * b;
      cold[a][b] = (i < nrows && j < ns && i >= j && j >= rowMin) ? S[(size_t)j * ld + i] : 0.0;
    }
  {
    double ri[16], rj[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int t = tid + 256 * u;
      const int p = t >> 6, r = t & 63;
      ri[u] = (p < bs && i0 + r < nrows && i0 + r >= rowMin) ? S[(size_t)(k0 + p) * ld + i0 + r] : 0.0;
      rj[u] = (p < bs && j0 + r < nrows && j0 + r >= rowMin) ? S[(size_t)(k0 + p) * ld + j0 + r] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int t = tid + 256 * u;
      sAi[t >> 6][t & 63] = ri[u];
      sAj[t >> 6][t & 63] = rj[u];
    }
  }
  __syncthreads();
  double c[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) c[a][b] = 0;
  for (int p = 0; p < bs; ++p) {
    double ai[4], aj[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) ai[a] = sAi[p][tx + 16 * a];
#pragma unroll
    for (int b = 0; b < 4; ++b) aj[b] = sAj[p][ty + 16 * b];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) c[a][b] += ai[a] * aj[b];
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = i0 + tx + 16 * a, j = j0 + ty + 16 * b;
      if (i < nrows && j < ns && i >= j && j >= rowMin) S[(size_t)j * ld + i] = cold[a][b] - c[a][b];
    }
}

// ------------------------------------------------------------------------------------------
// Envelope helpers.  Column c of block column J holds structural non-zeros in rows
// [J*CB, rowEnd[J]) plus the right-hand-side row ns.  envOff[J] = packed offset (in doubles) of
// block column J in the communication buffer (CB columns x (rowEnd[J] - J*CB) rows, column-major),
// the rhs row follows at envOff[nb].
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ba_env_pack(const double* __restrict__ S, int ld, int ns, const int* __restrict__ rowEnd,
            const long long* __restrict__ envOff, int nb, double* __restrict__ buf, int unpack,
            double* __restrict__ Sout) {
  const int J = blockIdx.x;
  if (J == nb) {  // rhs row
    for (int c = threadIdx.x; c < ns; c += 256) {
      if (unpack) Sout[(size_t)c * ld + ns] = buf[envOff[nb] + c];
      else buf[envOff[nb] + c] = S[(size_t)c * ld + ns];
    }
    return;
  }
  const int r0 = J * CB, r1 = rowEnd[J], len = r1 - r0;
  const int ncol = min(CB, ns - r0);
  double* b = buf + envOff[J];
  for (int t = threadIdx.x; t < ncol * len; t += 256) {
    const int cc = t / len, rr = t - cc * len;
    const size_t si = (size_t)(r0 + cc) * ld + r0 + rr;
    if (unpack) Sout[si] = b[t];
    else b[t] = S[si];
  }
}

// ------------------------------------------------------------------------------------------
// Backward substitution L^T x = y, one persistent 1024-thread CTA walking the block columns
// backwards: y (row ns of S after the factorisation) lives in shared memory for the whole solve;
// per block k: x_k = Linv_k^T y_k (Linv_k staged in shared memory with coalesced loads), then
// y_c -= sum_p L(k0 + p, c) x_k[p] for the columns c of block row k's envelope -- one warp per
// column, lanes over p (coalesced 512-byte column segments), shuffle reduction.
// Dynamic shared memory: (ns + CB*CB + CB) doubles.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
ba_chol_backward(const double* __restrict__ S, int ld, int ns, int nb,
                 const double* __restrict__ Linv, const int* __restrict__ firstBlk,
                 double* __restrict__ x) {
  extern __shared__ double s_dyn[];
  double* sy = s_dyn;                 // [ns]
  double* sLi = s_dyn + ns;           // [CB*CB] row-major Linv_k
  double* sx = sLi + CB * CB;         // [CB]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int c = tid; c < ns; c += 1024) sy[c] = S[(size_t)c * ld + ns];
  // Linv of the next block to process is prefetched into registers one step ahead
  double nxt[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) nxt[u] = Linv[(size_t)(nb - 1) * CB * CB + tid + 1024 * u];
  for (int k = nb - 1; k >= 0; --k) {
    const int k0 = k * CB, bs = min(CB, ns - k0);
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 4; ++u) sLi[tid + 1024 * u] = nxt[u];
    if (k > 0) {
#pragma unroll
      for (int u = 0; u < 4; ++u) nxt[u] = Linv[(size_t)(k - 1) * CB * CB + tid + 1024 * u];
    }
    __syncthreads();
    // x_k[c] = sum_{r >= c} Linv(r, c) y_k[r]: 16 threads per output column
    {
      const int c = tid >> 4, q = tid & 15;
      double part = 0;
      for (int r = c + q; r < bs; r += 16) part += sLi[r * CB + c] * sy[k0 + r];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o, 16);
      if (q == 0) sx[c] = (c < bs) ? part : 0.0;
    }
    __syncthreads();
    if (tid < bs) x[k0 + tid] = sx[tid];
    const int c0 = firstBlk[k] * CB;
    const double x0 = sx[lane], x1 = sx[lane + 32];
    // one warp per column, 8 columns per warp in flight (all loads issued before the reductions)
    for (int cb = c0 + warp; cb < k0; cb += 32 * 8) {
      double v0[8], v1[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int c = cb + 32 * u;
        const double* col = S + (size_t)c * ld + k0;
        v0[u] = (c < k0 && lane < bs) ? col[lane] : 0.0;
        v1[u] = (c < k0 && lane + 32 < bs) ? col[lane + 32] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int c = cb + 32 * u;
        double s = v0[u] * x0 + v1[u] * x1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0 && c < k0) sy[c] -= s;
      }
    }
  }
}

}  // namespace coslam
