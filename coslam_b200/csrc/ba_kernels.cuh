// ba_kernels.cuh -- device code of the Schur-complement Levenberg-Marquardt bundle adjustment
// (sm_100a, fp64).  Algorithm: sba_motstr_levmar_x of Lourakis & Argyros with the KRTS camera
// parameterisation of app/SL_CoSLAMBA.cpp:290-378.
//
// Data layout in HBM (per solver / per rank):
//   cameras (replicated): camK[m][5] (K0,K1,K2,K4,K5), camR0[m][9] (rotation of q0), pa[m][6] (v,t)
//   points  (this rank's shard): pb[n][3]
//   observations, point-major CSR: cam[N], pt[N], xy[N][2], wgt[N]
//   camera-major permutation of the free-camera observations: cobs[Nc], ccam[Nc]
//   W[N][18]   W_ij = A_ij^T B_ij (6x3 row-major) of the current linearisation
//   V[n][6], eb[n][3], U[m][21] (packed upper), ea[m][6]
//   pair work list for the Schur contraction: items {rowCam, colCam, begin, end} over entry pairs
//   (obsA, obsB) sorted by camera pair
//   S: block-sparse 64x64 tiles of the permuted reduced camera system + block-padded rhs, see
//   ba_tile.cuh / ba_plan.h
//
// One-observation-per-thread kernels evaluate residuals and the 2x6 / 2x3 Jacobian blocks; per-point
// (resp. per-camera) sums are collapsed with segmented warp shuffles and only segment heads touch
// memory with atomics.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace coslam {

struct BaDev {
  int m, n, mcon, ncon, mf, ns;
  long long N, Nc;
  const double* camK;
  const double* camR0;
  const int* cam;
  const int* pt;
  const double* xy;
  double* wgt;
  const long long* ptr;
  const int* cobs;
  const int* ccam;
  double* W;
  double* V;
  double* eb;
  double* U;
  double* ea;
  double* tiles;      // reduced camera system, tile t at tiles + 4096 t (column-major 64x64)
  double* rhs;        // block padded: block k at rhs + 64 k
  const int* solIdx;  // per free camera: index of its first row in rhs / x
  double* sc;         // scalars, see BaScalar
};

enum BaScalar {
  SC_COST = 0,     // sum w |e|^2 of the evaluated parameter set
  SC_DP_L2 = 1,    // |dp|^2
  SC_DL = 2,       // dp . (mu dp + g)
  SC_P_L2 = 3,     // |p|^2 over free parameters
  SC_NONFINITE = 4,
  SC_NSUM = 8,     // [0, SC_NSUM) are sums (all-reduce sum)
  SC_GINF = 8,     // |g|_inf                 (all-reduce max)
  SC_MAXDIAG = 9,  // max diag(U, V)
  SC_FAIL = 10,    // Cholesky break-down flag
  SC_NTOT = 16
};

__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
  // valid for non-negative doubles: the bit pattern is monotone
  atomicMax(reinterpret_cast<unsigned long long*>(addr),
            static_cast<unsigned long long>(__double_as_longlong(v)));
}

__device__ __forceinline__ void cross3(const double a[3], const double b[3], double c[3]) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

// x = pi(K (R(dq(v)) R0 X + t)); optionally A = dx/d(v,t) (2x6), B = dx/dX (2x3).
// Same formulas as oracle/ba_oracle.cpp:orc_ba_project.
template <bool WANT_A, bool WANT_B>
__device__ __forceinline__ void ba_project(const double* __restrict__ Kc,
                                           const double* __restrict__ R0,
                                           const double* __restrict__ p6, const double X[3],
                                           double xy[2], double A[12], double B[6]) {
  const double v[3] = {p6[0], p6[1], p6[2]};
  const double Y[3] = {R0[0] * X[0] + R0[1] * X[1] + R0[2] * X[2],
                       R0[3] * X[0] + R0[4] * X[1] + R0[5] * X[2],
                       R0[6] * X[0] + R0[7] * X[1] + R0[8] * X[2]};
  const double w = sqrt(1.0 - (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]));
  double vxY[3], vxvxY[3];
  cross3(v, Y, vxY);
  cross3(v, vxY, vxvxY);
  const double P[3] = {Y[0] + 2 * w * vxY[0] + 2 * vxvxY[0] + p6[3],
                       Y[1] + 2 * w * vxY[1] + 2 * vxvxY[1] + p6[4],
                       Y[2] + 2 * w * vxY[2] + 2 * vxvxY[2] + p6[5]};
  const double iz = 1.0 / P[2];
  const double un = Kc[0] * P[0] + Kc[1] * P[1];
  xy[0] = un * iz + Kc[2];
  xy[1] = Kc[3] * P[1] * iz + Kc[4];
  if (!WANT_A && !WANT_B) return;
  const double Jp[6] = {Kc[0] * iz, Kc[1] * iz, -un * iz * iz, 0.0, Kc[3] * iz,
                        -Kc[3] * P[1] * iz * iz};
  if (WANT_A) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double e[3] = {0, 0, 0};
      e[k] = 1.0;
      double exY[3], exvxY[3], vxexY[3];
      cross3(e, Y, exY);
      cross3(e, vxY, exvxY);
      cross3(v, exY, vxexY);
      const double dw = -v[k] / w;
      double dP[3];
#pragma unroll
      for (int c = 0; c < 3; ++c)
        dP[c] = 2 * dw * vxY[c] + 2 * w * exY[c] + 2 * exvxY[c] + 2 * vxexY[c];
      A[k] = Jp[0] * dP[0] + Jp[1] * dP[1] + Jp[2] * dP[2];
      A[6 + k] = Jp[3] * dP[0] + Jp[4] * dP[1] + Jp[5] * dP[2];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      A[3 + k] = Jp[k];
      A[9 + k] = Jp[3 + k];
    }
  }
  if (WANT_B) {
    // R(dq) from the unit quaternion (w, v), then Rf = R(dq) R0
    const double x = v[0], y = v[1], z = v[2];
    const double Rd[9] = {w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y),
                          2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x),
                          2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z};
    double Rf[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        Rf[3 * i + j] = Rd[3 * i] * R0[j] + Rd[3 * i + 1] * R0[3 + j] + Rd[3 * i + 2] * R0[6 + j];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      B[k] = Jp[0] * Rf[k] + Jp[1] * Rf[3 + k] + Jp[2] * Rf[6 + k];
      B[3 + k] = Jp[3] * Rf[k] + Jp[4] * Rf[3 + k] + Jp[5] * Rf[6 + k];
    }
  }
}

// Segmented warp reduction by key for contiguous segments: after the call the FIRST lane of every
// run of equal keys holds the run's sum.
template <int NV>
__device__ __forceinline__ void seg_reduce(double* v, int key, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int ko = __shfl_down_sync(0xffffffffu, key, o);
    const bool take = (lane + o < 32) && (ko == key);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const double t = __shfl_down_sync(0xffffffffu, v[k], o);
      if (take) v[k] += t;
    }
  }
}

__device__ __forceinline__ double block_sum_1(double v, double* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) s_red[wid] = v;
  __syncthreads();
  double t = 0;
  if (wid == 0) {
    t = (lane < (int)(blockDim.x >> 5)) ? s_red[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;  // valid in warp 0
}

__device__ __forceinline__ double tukey_w(double e, double s) {
  if (e >= s) return 0;
  e /= s;
  e = 1 - e * e;
  return e * e;
}

// ------------------------------------------------------------------------------------------
// cost / weights / outliers: one observation per thread
// mode 0: sc[SC_COST] += w |e|^2 ; mode 1: wgt = tukey(|e|, maxErr) ; mode 2: outlier = |e|>=maxErr,
// sc[SC_COST] += number of outliers
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ba_residual_kernel(BaDev d, const double* __restrict__ pa, const double* __restrict__ pb, int mode,
                   double maxErr, unsigned char* __restrict__ outlier) {
  __shared__ double s_red[8];
  const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double c = 0;
  if (o < d.N) {
    const int j = d.cam[o], i = d.pt[o];
    const double X[3] = {pb[3 * (size_t)i], pb[3 * (size_t)i + 1], pb[3 * (size_t)i + 2]};
    double h[2];
    ba_project<false, false>(d.camK + 5 * j, d.camR0 + 9 * j, pa + 6 * j, X, h, nullptr, nullptr);
    const double dx = d.xy[2 * o] - h[0], dy = d.xy[2 * o + 1] - h[1];
    const double e2 = dx * dx + dy * dy;
    if (mode == 0) {
      c = d.wgt[o] * e2;
    } else if (mode == 1) {
      d.wgt[o] = tukey_w(sqrt(e2), maxErr);
    } else {
      const int isOut = (sqrt(e2) >= maxErr) ? 1 : 0;
      outlier[o] = (unsigned char)isOut;
      c = (double)isOut;
    }
  }
  if (mode == 2) {  // number of outliers -> sc[SC_COST] (info[13] of the solver entry points)
    const double t = block_sum_1(c, s_red);
    if (threadIdx.x == 0 && t != 0) atomicAdd(&d.sc[SC_COST], t);
  }
  if (mode == 0) {
    const double t = block_sum_1(c, s_red);
    if (threadIdx.x == 0) {
      if (isfinite(t))
        atomicAdd(&d.sc[SC_COST], t);
      else
        d.sc[SC_NONFINITE] = 1.0;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Linearisation, point-major: per observation e, A (2x6), B (2x3) (scaled by sqrt(w)); stores
// W_ij = A^T B; collapses V_i = sum B^T B and eb_i = sum B^T e per point with segmented shuffles.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2)
ba_linearize_points(BaDev d, const double* __restrict__ pa, const double* __restrict__ pb) {
  if (blockIdx.x == 0 && threadIdx.x < 2) d.sc[SC_GINF + threadIdx.x] = 0.0;  // |g|_inf, max diag: set by ba_stats_kernel
  const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  double acc[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = 0;
  int key = -1 - lane;  // inactive lanes never merge
  if (o < d.N) {
    const int j = d.cam[o], i = d.pt[o];
    key = i;
    const double X[3] = {pb[3 * (size_t)i], pb[3 * (size_t)i + 1], pb[3 * (size_t)i + 2]};
    double h[2], A[12], B[6];
    ba_project<true, true>(d.camK + 5 * j, d.camR0 + 9 * j, pa + 6 * j, X, h, A, B);
    const double sw = sqrt(d.wgt[o]);
    const double e0 = (d.xy[2 * o] - h[0]) * sw, e1 = (d.xy[2 * o + 1] - h[1]) * sw;
#pragma unroll
    for (int k = 0; k < 6; ++k) B[k] *= sw;
    if (i >= d.ncon) {
      acc[0] = B[0] * B[0] + B[3] * B[3];
      acc[1] = B[0] * B[1] + B[3] * B[4];
      acc[2] = B[0] * B[2] + B[3] * B[5];
      acc[3] = B[1] * B[1] + B[4] * B[4];
      acc[4] = B[1] * B[2] + B[4] * B[5];
      acc[5] = B[2] * B[2] + B[5] * B[5];
      acc[6] = B[0] * e0 + B[3] * e1;
      acc[7] = B[1] * e0 + B[4] * e1;
      acc[8] = B[2] * e0 + B[5] * e1;
      if (j >= d.mcon) {
#pragma unroll
        for (int k = 0; k < 12; ++k) A[k] *= sw;
        double* Wp = d.W + 18 * (size_t)o;
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) Wp[3 * r + c] = A[r] * B[c] + A[6 + r] * B[3 + c];
      }
    }
  }
  seg_reduce<9>(acc, key, lane);
  const int kprev = __shfl_up_sync(0xffffffffu, key, 1);
  const bool head = (lane == 0) || (kprev != key);
  if (head && key >= d.ncon) {
#pragma unroll
    for (int k = 0; k < 6; ++k) atomicAdd(&d.V[6 * (size_t)key + k], acc[k]);
#pragma unroll
    for (int k = 0; k < 3; ++k) atomicAdd(&d.eb[3 * (size_t)key + k], acc[6 + k]);
  }
}

// Linearisation, camera-major over the free-camera observations: U_j = sum A^T A (packed upper,
// 21) and ea_j = sum A^T e (6), collapsed per camera with segmented shuffles.
__global__ void __launch_bounds__(128)
ba_linearize_cams(BaDev d, const double* __restrict__ pa, const double* __restrict__ pb) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  double acc[27];
#pragma unroll
  for (int k = 0; k < 27; ++k) acc[k] = 0;
  int key = -1 - lane;
  if (q < d.Nc) {
    const int o = d.cobs[q], j = d.ccam[q], i = d.pt[o];
    key = j;
    const double X[3] = {pb[3 * (size_t)i], pb[3 * (size_t)i + 1], pb[3 * (size_t)i + 2]};
    double h[2], A[12];
    ba_project<true, false>(d.camK + 5 * j, d.camR0 + 9 * j, pa + 6 * j, X, h, A, nullptr);
    const double sw = sqrt(d.wgt[o]);
    const double e0 = (d.xy[2 * (size_t)o] - h[0]) * sw, e1 = (d.xy[2 * (size_t)o + 1] - h[1]) * sw;
#pragma unroll
    for (int k = 0; k < 12; ++k) A[k] *= sw;
    int t = 0;
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = r; c < 6; ++c) acc[t++] = A[r] * A[c] + A[6 + r] * A[6 + c];
#pragma unroll
    for (int r = 0; r < 6; ++r) acc[21 + r] = A[r] * e0 + A[6 + r] * e1;
  }
  seg_reduce<27>(acc, key, lane);
  const int kprev = __shfl_up_sync(0xffffffffu, key, 1);
  const bool head = (lane == 0) || (kprev != key);
  if (head && key >= 0) {
#pragma unroll
    for (int k = 0; k < 21; ++k) atomicAdd(&d.U[21 * (size_t)key + k], acc[k]);
#pragma unroll
    for (int k = 0; k < 6; ++k) atomicAdd(&d.ea[6 * (size_t)key + k], acc[21 + k]);
  }
}

// |g|_inf and max diag over (U, ea) of the free cameras [global after all-reduce] and (V, eb) of
// this rank's free points.
__global__ void __launch_bounds__(256) ba_stats_kernel(BaDev d) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double g = 0, dg = 0;
  if (t < d.n) {
    if (t >= d.ncon) {
      const double* V = d.V + 6 * t;
      const double* e = d.eb + 3 * t;
      g = fmax(fabs(e[0]), fmax(fabs(e[1]), fabs(e[2])));
      dg = fmax(V[0], fmax(V[3], V[5]));
    }
  } else if (t < (long long)d.n + d.m) {
    const int j = (int)(t - d.n);
    if (j >= d.mcon) {
      const double* U = d.U + 21 * (size_t)j;
      const double* e = d.ea + 6 * (size_t)j;
      // packed upper: diagonal entries at offsets 0, 6, 11, 15, 18, 20
      dg = fmax(fmax(U[0], U[6]), fmax(fmax(U[11], U[15]), fmax(U[18], U[20])));
#pragma unroll
      for (int k = 0; k < 6; ++k) g = fmax(g, fabs(e[k]));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    g = fmax(g, __shfl_xor_sync(0xffffffffu, g, o));
    dg = fmax(dg, __shfl_xor_sync(0xffffffffu, dg, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (g > 0) atomic_max_nonneg(&d.sc[SC_GINF], g);
    if (dg > 0) atomic_max_nonneg(&d.sc[SC_MAXDIAG], dg);
  }
}

// S <- 0 with the damped camera blocks U*_j on the diagonal tiles, rhs <- ea.  One CTA per tile
// (blockIdx.x < nTiles) zeroes it and, if it is the diagonal tile of block k (diagBlk[t] = k), writes
// the lower part of U_j + mu I for the cameras order[blkCam0[k] .. blkCam0[k+1]) of that block and
// their slice of rhs.  addU: this rank contributes U/ea/mu (rank 0 only in the multi-GPU case, where
// U/ea are already all-reduced).  Padding rows of a block stay zero (the factorisation treats them
// as identity).
__device__ __forceinline__ void ba_tile_init_body(const BaDev& d, double mu, int addU, const int* __restrict__ diagBlk,
                                                  const int* __restrict__ blkCam0, const int* __restrict__ order) {
  const int t = blockIdx.x;
  double2* tile = reinterpret_cast<double2*>(d.tiles + (size_t)t * 4096);
  const double2 z = make_double2(0.0, 0.0);
#pragma unroll
  for (int u = 0; u < 8; ++u) tile[threadIdx.x + 256 * u] = z;
  const int k = diagBlk[t];
  if (k < 0) return;
  if (threadIdx.x < 64) d.rhs[(size_t)k * 64 + threadIdx.x] = 0.0;
  __syncthreads();
  if (!addU) return;
  const int p0 = blkCam0[k], ncam = blkCam0[k + 1] - p0;
  for (int e = threadIdx.x; e < ncam * 27; e += 256) {
    const int c = e / 27, q = e - 27 * c;
    const int cam = order[p0 + c] + d.mcon;
    const int off = 6 * c;
    if (q < 21) {
      // packed upper index q -> (a <= b); stored at row off+b, column off+a of the lower tile
      int a = 0, rem = q;
      while (rem >= 6 - a) {
        rem -= 6 - a;
        ++a;
      }
      const int b = a + rem;
      double v = d.U[21 * (size_t)cam + q];
      if (a == b) v += mu;
      d.tiles[(size_t)t * 4096 + (off + a) * 64 + off + b] = v;
    } else {
      d.rhs[(size_t)k * 64 + off + (q - 21)] = d.ea[6 * (size_t)cam + (q - 21)];
    }
  }
}

// ------------------------------------------------------------------------------------------
// Schur contraction over camera pairs.  Work item = a run of (obsA, obsB) entries that belong to
// one camera pair (j <= k): block(j,k) -= sum_i W_ij V*_i^-1 W_ik^T (and, for j == k,
// rhs_j -= sum_i W_ij V*_i^-1 eb_i).  One warp per item; two lanes share an entry (3 block columns
// each), 18 accumulators per lane, butterfly reduction, 18 atomics per lane-parity per item.
// ------------------------------------------------------------------------------------------
struct BaPairItem {
  int rowCam, colCam;  // free-camera indices (0-based in the reduced system)
  int begin, end;      // entry range
  // destination of the 6x6 block acc(r, c) (r: row camera's parameter, c: column camera's):
  // tiles[dst + (cOff + (trans ? c : r)) * 64 + rOff + (trans ? r : c)] -- the block lands in the
  // LOWER triangle of the permuted system whichever of the two cameras comes first in the ordering
  int dst, rOff, cOff, trans;
  int rhsIdx, pad0, pad1, pad2;  // rhs index of the row camera (diagonal items only)
};

__device__ __forceinline__ void inv3sym_mu(const double* __restrict__ V, double mu, double I[6]) {
  const double a = V[0] + mu, b = V[1], c = V[2], d = V[3] + mu, e = V[4], f = V[5] + mu;
  const double A = d * f - e * e, B = c * e - b * f, C = b * e - c * d;
  const double r = 1.0 / (a * A + b * B + c * C);
  I[0] = A * r;
  I[1] = B * r;
  I[2] = C * r;
  I[3] = (a * f - c * c) * r;
  I[4] = (b * c - a * e) * r;
  I[5] = (a * d - b * b) * r;
}

// One launch in front of every Schur contraction ("prep"): blocks [0, nTiles) initialise the tiles
// (ba_tile_init_body), the following blocks compute the damped inverses V*_i^-1 (one point per thread) and --
// when a linearisation preceded this trial (doStats) -- the gradient / diagonal statistics |g|_inf and
// max diag(U, V) that the host needs for the eps1 test (it reads them with the trial's scalars).  Block 0
// also clears the scalars of the trial and the grid clears the task counters of the solve.
__global__ void __launch_bounds__(256)
ba_prep_kernel(BaDev d, double mu, int addU, const int* __restrict__ diagBlk, const int* __restrict__ blkCam0,
               const int* __restrict__ order, int* __restrict__ cnt, int nCnt, int nTiles,
               double* __restrict__ Vinv, int doStats, double* __restrict__ zeroBuf, long long zeroCount) {
  for (int q = blockIdx.x * 256 + threadIdx.x; q < nCnt; q += gridDim.x * 256) cnt[q] = 0;
  // accumulator of the point back-substitution (ba_back_cams_points)
  for (long long q = (long long)blockIdx.x * 256 + threadIdx.x; q < zeroCount; q += (long long)gridDim.x * 256)
    zeroBuf[q] = 0.0;
  if (blockIdx.x == 0 && threadIdx.x < 6) {
    const int slot[6] = {SC_DP_L2, SC_DL, SC_P_L2, SC_FAIL, SC_COST, SC_NONFINITE};
    d.sc[slot[threadIdx.x]] = 0.0;
  }
  if ((int)blockIdx.x < nTiles) {
    ba_tile_init_body(d, mu, addU, diagBlk, blkCam0, order);
    return;
  }
  const long long t = (long long)(blockIdx.x - nTiles) * 256 + threadIdx.x;
  double g = 0, dg = 0;
  if (t < d.n) {
    double I[6] = {0, 0, 0, 0, 0, 0};
    if (t >= d.ncon) {
      const double* V = d.V + 6 * t;
      inv3sym_mu(V, mu, I);
      if (doStats) {
        const double* e = d.eb + 3 * t;
        g = fmax(fabs(e[0]), fmax(fabs(e[1]), fabs(e[2])));
        dg = fmax(V[0], fmax(V[3], V[5]));
      }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) Vinv[6 * t + k] = I[k];
  } else if (doStats && t < (long long)d.n + d.m) {
    const int j = (int)(t - d.n);
    if (j >= d.mcon) {
      const double* U = d.U + 21 * (size_t)j;
      const double* e = d.ea + 6 * (size_t)j;
      dg = fmax(fmax(U[0], U[6]), fmax(fmax(U[11], U[15]), fmax(U[18], U[20])));
#pragma unroll
      for (int k = 0; k < 6; ++k) g = fmax(g, fabs(e[k]));
    }
  }
  if (doStats) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      g = fmax(g, __shfl_xor_sync(0xffffffffu, g, o));
      dg = fmax(dg, __shfl_xor_sync(0xffffffffu, dg, o));
    }
    if ((threadIdx.x & 31) == 0) {
      if (g > 0) atomic_max_nonneg(&d.sc[SC_GINF], g);
      if (dg > 0) atomic_max_nonneg(&d.sc[SC_MAXDIAG], dg);
    }
  }
}

// Entries carry the point index ({obsA, obsB, point, -}) and the damped inverses V*^-1 are computed
// once per trial (ba_vinv_kernel), so the loads of an entry are ONE dependent level below the entry
// itself; the next entry is fetched while the current one is processed.
template <int MINB>
__global__ void __launch_bounds__(128, MINB)
ba_schur_pairs_t(BaDev d, const BaPairItem* __restrict__ items, int nItems,
                 const int4* __restrict__ entries, const double* __restrict__ Vinv) {
  const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (item >= nItems) return;
  const int lane = threadIdx.x & 31;
  const int half = lane & 1;  // which 3 columns of the 6x6 block
  const BaPairItem it = items[item];
  double acc[18];
#pragma unroll
  for (int k = 0; k < 18; ++k) acc[k] = 0;
  double racc[3] = {0, 0, 0};
  const bool diag = (it.rowCam == it.colCam);
  int e = it.begin + (lane >> 1);
  int4 nxt = (e < it.end) ? __ldg(&entries[e]) : make_int4(0, 0, 0, 0);
  for (; e < it.end; e += 16) {
    const int4 ob = nxt;
    if (e + 16 < it.end) nxt = __ldg(&entries[e + 16]);
    const double* Wa = d.W + 18 * (size_t)ob.x;
    const double* Wb = d.W + 18 * (size_t)ob.y;
    const int i = ob.z;
    const double* Vi = Vinv + 6 * (size_t)i;
    double Iv[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) Iv[k] = Vi[k];
    double wa[18];
#pragma unroll
    for (int k = 0; k < 18; ++k) wa[k] = Wa[k];
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) {
      const int c = 3 * half + cc;
      const double b0 = Wb[3 * c], b1 = Wb[3 * c + 1], b2 = Wb[3 * c + 2];
      const double t0 = Iv[0] * b0 + Iv[1] * b1 + Iv[2] * b2;
      const double t1 = Iv[1] * b0 + Iv[3] * b1 + Iv[4] * b2;
      const double t2 = Iv[2] * b0 + Iv[4] * b1 + Iv[5] * b2;
#pragma unroll
      for (int r = 0; r < 6; ++r)
        acc[3 * r + cc] += wa[3 * r] * t0 + wa[3 * r + 1] * t1 + wa[3 * r + 2] * t2;
    }
    if (diag) {
      const double* eb = d.eb + 3 * (size_t)i;
      const double t0 = Iv[0] * eb[0] + Iv[1] * eb[1] + Iv[2] * eb[2];
      const double t1 = Iv[1] * eb[0] + Iv[3] * eb[1] + Iv[4] * eb[2];
      const double t2 = Iv[2] * eb[0] + Iv[4] * eb[1] + Iv[5] * eb[2];
#pragma unroll
      for (int rr = 0; rr < 3; ++rr) {
        const double* wr = Wa + 3 * (3 * half + rr);  // row (3*half + rr) of W_a
        racc[rr] += wr[0] * t0 + wr[1] * t1 + wr[2] * t2;
      }
    }
  }
  // butterfly over lanes of equal parity
#pragma unroll
  for (int o = 2; o < 32; o <<= 1) {
#pragma unroll
    for (int k = 0; k < 18; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
#pragma unroll
    for (int k = 0; k < 3; ++k) racc[k] += __shfl_xor_sync(0xffffffffu, racc[k], o);
  }
  if (lane < 2) {
    double* T = d.tiles + it.dst;
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        const int c = 3 * half + cc;
        // only the lower part (row >= col of the tile, i.e. c >= r) of diagonal blocks is kept
        if (!diag || c >= r) {
          const int col = it.cOff + (it.trans ? c : r), row = it.rOff + (it.trans ? r : c);
          atomicAdd(&T[col * 64 + row], -acc[3 * r + cc]);
        }
      }
    if (diag) {
#pragma unroll
      for (int rr = 0; rr < 3; ++rr) atomicAdd(&d.rhs[it.rhsIdx + 3 * half + rr], -racc[rr]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Staged variant of the pair-list contraction (default).  ba_schur_pairs_t gathers the rows of an entry
// with scalar 8-byte loads issued by the lanes that consume them: 33 load instructions per 16 entries,
// each touching ~16 different 128-byte lines -> ~33 L1TEX wavefronts per entry, which is what bounds it
// (0.47 ms at c4 for 1.7 GFLOP).  Here a warp works on 32 entries at a time:
//   1. STAGE: the 21 16-byte segments of every entry -- W_a (9), W_b (9), V*^-1 (3) -- are copied with
//      cp.async in SEGMENT-major order (consecutive lanes copy consecutive segments of the same row, so
//      one instruction touches ~8 lines instead of 32: ~5 wavefronts per entry), into a row of 42
//      doubles per entry (336 B = 21 x 16 B, an odd multiple of 16 B: conflict-free 16-byte reads);
//   2. CONTRACT: lane j owns entry j, reads its row with 21 LDS.128 and accumulates the whole 6x6 block
//      (+ the 6 right-hand-side terms of diagonal items) in registers: 162 DFMA per entry;
//   3. the next batch is in flight (second buffer) while the current one is contracted;
//   4. at the end of the item the 32 partial blocks are summed through the (now free) staging buffer:
//      lane k adds up element k of all lanes -- 36 + 6 elements, one atomic each.
// ------------------------------------------------------------------------------------------
constexpr int BA_ST_WARPS = 4;
constexpr int BA_ST_ROW = 42;                                   // doubles per staged entry
constexpr int BA_ST_BUF = 32 * BA_ST_ROW;                       // doubles per batch buffer (10752 B)
constexpr int BA_ST_SMEM = BA_ST_WARPS * 2 * BA_ST_BUF * 8;     // 86016 B per CTA (dynamic)

__device__ __forceinline__ void ba_cp_async16(double* smemDst, const double* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smemDst);
  // .cg: L2 only.  Allocating in L1 (.ca; the four warps of a CTA work on partner cameras of the same row
  // camera) was measured: no difference (0.359 vs 0.361 ms at c4).
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc) : "memory");
}

// copy the rows of the batch whose entry descriptors the lanes hold (lane j: entry j; ne valid entries)
__device__ __forceinline__ void ba_st_stage(const BaDev& d, const double* __restrict__ Vinv, int4 ob, int ne,
                                            int lane, double* buf) {
#pragma unroll
  for (int t = 0; t < 9; ++t) {  // W_a and W_b: 32 rows x 9 segments each
    const int sidx = t * 32 + lane, j = sidx / 9, part = sidx - 9 * j;
    const int oa = __shfl_sync(0xffffffffu, ob.x, j), obb = __shfl_sync(0xffffffffu, ob.y, j);
    if (j < ne) {
      ba_cp_async16(buf + j * BA_ST_ROW + 2 * part, d.W + 18 * (size_t)oa + 2 * part);
      ba_cp_async16(buf + j * BA_ST_ROW + 18 + 2 * part, d.W + 18 * (size_t)obb + 2 * part);
    }
  }
#pragma unroll
  for (int t = 0; t < 3; ++t) {  // V*^-1: 32 rows x 3 segments
    const int sidx = t * 32 + lane, j = sidx / 3, part = sidx - 3 * j;
    const int pi = __shfl_sync(0xffffffffu, ob.z, j);
    if (j < ne) ba_cp_async16(buf + j * BA_ST_ROW + 36 + 2 * part, Vinv + 6 * (size_t)pi + 2 * part);
  }
  asm volatile("cp.async.commit_group;\n" ::: "memory");
}

__global__ void __launch_bounds__(32 * BA_ST_WARPS, 2)
ba_schur_pairs_st(BaDev d, const BaPairItem* __restrict__ items, int nItems, const int4* __restrict__ entries,
                  const double* __restrict__ Vinv) {
  extern __shared__ __align__(16) double s_st[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * BA_ST_WARPS + w;
  if (item >= nItems) return;
  const BaPairItem it = items[item];
  const bool diag = (it.rowCam == it.colCam);
  double* buf0 = s_st + (size_t)w * 2 * BA_ST_BUF;
  double acc[36], racc[6];
#pragma unroll
  for (int k = 0; k < 36; ++k) acc[k] = 0.0;
#pragma unroll
  for (int k = 0; k < 6; ++k) racc[k] = 0.0;
  const int nb = (it.end - it.begin + 31) >> 5;
  const int4 zero4 = make_int4(0, 0, 0, 0);
  int4 cur = (it.begin + lane < it.end) ? __ldg(&entries[it.begin + lane]) : zero4;
  int4 nxt = (it.begin + 32 + lane < it.end) ? __ldg(&entries[it.begin + 32 + lane]) : zero4;
  ba_st_stage(d, Vinv, cur, min(32, it.end - it.begin), lane, buf0);
  for (int b = 0; b < nb; ++b) {
    const int e0 = it.begin + 32 * b;
    const int4 mine = cur;  // descriptor of this lane's entry in batch b (point index for e_b)
    cur = nxt;
    if (b + 1 < nb) {
      nxt = (e0 + 64 + lane < it.end) ? __ldg(&entries[e0 + 64 + lane]) : zero4;
      ba_st_stage(d, Vinv, cur, min(32, it.end - (e0 + 32)), lane, buf0 + ((b + 1) & 1) * BA_ST_BUF);
      asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    }
    __syncwarp();
    if (e0 + lane < it.end) {
      const double* row = buf0 + (b & 1) * BA_ST_BUF + lane * BA_ST_ROW;
      double wa[18], wb[18], Iv[6];
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const double2 v = *reinterpret_cast<const double2*>(row + 2 * k);
        wa[2 * k] = v.x;
        wa[2 * k + 1] = v.y;
      }
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const double2 v = *reinterpret_cast<const double2*>(row + 18 + 2 * k);
        wb[2 * k] = v.x;
        wb[2 * k + 1] = v.y;
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double2 v = *reinterpret_cast<const double2*>(row + 36 + 2 * k);
        Iv[2 * k] = v.x;
        Iv[2 * k + 1] = v.y;
      }
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        const double b0 = wb[3 * c], b1 = wb[3 * c + 1], b2 = wb[3 * c + 2];
        const double t0 = Iv[0] * b0 + Iv[1] * b1 + Iv[2] * b2;
        const double t1 = Iv[1] * b0 + Iv[3] * b1 + Iv[4] * b2;
        const double t2 = Iv[2] * b0 + Iv[4] * b1 + Iv[5] * b2;
#pragma unroll
        for (int r = 0; r < 6; ++r) acc[6 * r + c] += wa[3 * r] * t0 + wa[3 * r + 1] * t1 + wa[3 * r + 2] * t2;
      }
      if (diag) {
        const double* eb = d.eb + 3 * (size_t)mine.z;
        const double e0v = eb[0], e1v = eb[1], e2v = eb[2];
        const double t0 = Iv[0] * e0v + Iv[1] * e1v + Iv[2] * e2v;
        const double t1 = Iv[1] * e0v + Iv[3] * e1v + Iv[4] * e2v;
        const double t2 = Iv[2] * e0v + Iv[4] * e1v + Iv[5] * e2v;
#pragma unroll
        for (int r = 0; r < 6; ++r) racc[r] += wa[3 * r] * t0 + wa[3 * r + 1] * t1 + wa[3 * r + 2] * t2;
      }
    }
    __syncwarp();  // the buffer of batch b is free again (restaged at iteration b + 1)
  }
  // ---- sum over the 32 lanes through shared memory: red[k * 33 + lane]
  double* red = buf0;  // 42 x 33 doubles = 11088 B <= 2 buffers
#pragma unroll
  for (int k = 0; k < 36; ++k) red[k * 33 + lane] = acc[k];
  if (diag) {
#pragma unroll
    for (int k = 0; k < 6; ++k) red[(36 + k) * 33 + lane] = racc[k];
  }
  __syncwarp();
  double* T = d.tiles + it.dst;
  for (int k = lane; k < (diag ? 42 : 36); k += 32) {
    double sum = 0.0;
#pragma unroll 8
    for (int l = 0; l < 32; ++l) sum += red[k * 33 + l];
    if (k < 36) {
      const int r = k / 6, c = k - 6 * r;
      // only the lower part (row >= col of the tile, i.e. c >= r) of diagonal blocks is kept
      if (!diag || c >= r) {
        const int col = it.cOff + (it.trans ? c : r), row = it.rOff + (it.trans ? r : c);
        atomicAdd(&T[col * 64 + row], -sum);
      }
    } else {
      atomicAdd(&d.rhs[it.rhsIdx + (k - 36)], -sum);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Pair work lists built on the device (solver set-up).  For every free observation a = (camera ja,
// point i) of the camera-major list and every observation b of the same point with camera jb > ja,
// or jb == ja and b >= a: one entry in bucket (ja, jb).  Pass 1 counts per bucket, ba_scan_u32 turns
// the counts into offsets, pass 2 places {a, b, i} at offset + running index.  (The host version
// wrote 81 MB of entries at c4 and uploaded them: 57 ms of a 74 ms solver creation.)  The order
// inside a bucket depends on the atomics; the contraction sums the same terms in a different order.
// ------------------------------------------------------------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(256)
ba_pairs_build(BaDev d, unsigned* __restrict__ cnt, const unsigned* __restrict__ off,
               int4* __restrict__ entries) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= d.Nc) return;
  const int a = d.cobs[q], i = d.pt[a];
  if (i < d.ncon) return;
  const int ja = d.cam[a] - d.mcon;
  const long long o0 = d.ptr[i], o1 = d.ptr[i + 1];
  for (long long b = o0; b < o1; ++b) {
    const int jb = d.cam[b] - d.mcon;
    if (jb > ja || (jb == ja && b >= a)) {
      const size_t bucket = (size_t)ja * d.mf + jb;
      const unsigned k = atomicAdd(&cnt[bucket], 1u);
      if (FILL) entries[off[bucket] + k] = make_int4(a, (int)b, i, 0);
    }
  }
}

// exclusive prefix sum of n unsigned counts into off[0 .. n] (one CTA; n is the number of camera
// pairs, a few 10^5)
__global__ void __launch_bounds__(1024) ba_scan_u32(const unsigned* __restrict__ cnt, unsigned* __restrict__ off,
                                                     long long n) {
  __shared__ unsigned s_sum[1024];
  const int tid = threadIdx.x;
  const long long chunk = (n + 1023) / 1024, lo = tid * chunk, hi = min(n, lo + chunk);
  unsigned acc = 0;
  for (long long k = lo; k < hi; ++k) acc += cnt[k];
  s_sum[tid] = acc;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const unsigned v = (tid >= o) ? s_sum[tid - o] : 0u;
    __syncthreads();
    s_sum[tid] += v;
    __syncthreads();
  }
  unsigned run = (tid > 0) ? s_sum[tid - 1] : 0u;
  for (long long k = lo; k < hi; ++k) {
    off[k] = run;
    run += cnt[k];
  }
  if (tid == 1023) off[n] = s_sum[1023];
}

// ------------------------------------------------------------------------------------------
// Schur contraction on the fp64 tensor cores (DMMA m8n8k4).  Same work items as ba_schur_pairs_t
// (one warp per run of <= 512 entries of one camera pair); the 6x6 block
//     sum_i W_ia V*_i^-1 W_ib^T  =  [W_ia]_(6 x 3n) . [V*_i^-1 W_ib^T]_(3n x 6)
// is ONE dense contraction with the entries stacked along k.  Per chunk of 16 entries:
//   1. the rows W_a, W_b (144 B each), V*^-1 (48 B) and e_b of the entries are staged in shared
//      memory with 16-byte loads of contiguous segments (the scalar gathers of the SIMT kernel are
//      what bound it: L1TEX wavefronts, see profiles/);
//   2. T_i = V*_i^-1 W_ib^T (3 x 6) and the right-hand-side column V*_i^-1 e_b,i are formed in place;
//   3. 12 DMMAs accumulate C(8x8) += A(8x4) B(4x8): A(r, k) = W_ia(r, c) with k = 3 i + c (rows 6, 7
//      are zero), B(k, n) = T_i(c, n) for n < 6 and B(k, 6) = (V*^-1 e_b)(c): column 6 of the
//      accumulator is the right-hand side of diagonal items for free.
// ------------------------------------------------------------------------------------------
constexpr int BA_MMA_CHUNK = 16;                      // entries per chunk
constexpr int BA_MMA_ESTRIDE = 18 + 24 + 2;           // doubles per entry: W_a | T (3 x 8) | pad

__device__ __forceinline__ void ba_dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(128, 6)
ba_schur_mma(BaDev d, const BaPairItem* __restrict__ items, int nItems, const int4* __restrict__ entries,
             const double* __restrict__ Vinv) {
  __shared__ __align__(16) double s_e[4][BA_MMA_CHUNK * BA_MMA_ESTRIDE];  // per warp
  __shared__ __align__(16) double s_wb[4][BA_MMA_CHUNK * 28];             // W_b (18) | Vinv (6) | e_b (3) | pad
  __shared__ int4 s_ob[4][BA_MMA_CHUNK];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * 4 + w;
  if (item >= nItems) return;
  const BaPairItem it = items[item];
  const bool diag = (it.rowCam == it.colCam);
  double* se = s_e[w];
  double* sb = s_wb[w];
  double c0 = 0.0, c1 = 0.0;
  const int lr = lane >> 2, lk = lane & 3;
  for (int e0 = it.begin; e0 < it.end; e0 += BA_MMA_CHUNK) {
    const int ne = min(BA_MMA_CHUNK, it.end - e0);
    // ---- 1. stage: 21 (+2) 16-byte segments per entry: W_a (9), W_b (9), Vinv (3), e_b (1.5)
    if (lane < BA_MMA_CHUNK) s_ob[w][lane] = (lane < ne) ? __ldg(&entries[e0 + lane]) : make_int4(0, 0, 0, 0);
    __syncwarp();
#pragma unroll 4
    for (int t = lane; t < BA_MMA_CHUNK * 23; t += 32) {
      const int j = t / 23, part = t - 23 * j;
      double2 v = make_double2(0.0, 0.0);
      if (j < ne) {
        const int4 ob = s_ob[w][j];
        if (part < 9) v = *reinterpret_cast<const double2*>(d.W + 18 * (size_t)ob.x + 2 * part);
        else if (part < 18) v = *reinterpret_cast<const double2*>(d.W + 18 * (size_t)ob.y + 2 * (part - 9));
        else if (part < 21) v = *reinterpret_cast<const double2*>(Vinv + 6 * (size_t)ob.z + 2 * (part - 18));
        else if (diag) {
          const double* eb = d.eb + 3 * (size_t)ob.z;
          v = (part == 21) ? make_double2(eb[0], eb[1]) : make_double2(eb[2], 0.0);
        }
      }
      if (part < 9) *reinterpret_cast<double2*>(se + j * BA_MMA_ESTRIDE + 2 * part) = v;
      else *reinterpret_cast<double2*>(sb + j * 28 + 2 * (part - 9)) = v;
    }
    __syncwarp();
    // ---- 2. T_i(c, n) = sum_c' Vinv(c, c') W_b(n, c')  (n < 6),  T_i(c, 6) = (Vinv e_b)(c), T_i(c, 7) = 0
    for (int t = lane; t < BA_MMA_CHUNK * 24; t += 32) {
      const int j = t / 24, cn = t - 24 * j, c = cn >> 3, n = cn & 7;
      const double* q = sb + j * 28;
      const double* Vi = q + 18;
      const double v0 = (c == 0) ? Vi[0] : (c == 1) ? Vi[1] : Vi[2];
      const double v1 = (c == 0) ? Vi[1] : (c == 1) ? Vi[3] : Vi[4];
      const double v2 = (c == 0) ? Vi[2] : (c == 1) ? Vi[4] : Vi[5];
      double r = 0.0;
      if (n < 6) r = v0 * q[3 * n] + v1 * q[3 * n + 1] + v2 * q[3 * n + 2];
      else if (n == 6) r = v0 * q[24] + v1 * q[25] + v2 * q[26];
      se[j * BA_MMA_ESTRIDE + 18 + cn] = r;
    }
    __syncwarp();
    // ---- 3. 12 k-steps of 4: k = 3 j + c
#pragma unroll
    for (int ks = 0; ks < 3 * BA_MMA_CHUNK / 4; ++ks) {
      const int k = 4 * ks + lk, j = k / 3, c = k - 3 * j;
      const double a = (lr < 6) ? se[j * BA_MMA_ESTRIDE + 3 * lr + c] : 0.0;  // W_a(r = lr, c)
      const double b = se[j * BA_MMA_ESTRIDE + 18 + 8 * c + lr];             // T(c, n = lr)
      ba_dmma884(c0, c1, a, b);
    }
    __syncwarp();
  }
  // C(r = lr, n = 2 lk, 2 lk + 1): block element (r, n) for n < 6; column 6 = rhs
  if (lr < 6) {
    double* T = d.tiles + it.dst;
    const double cv[2] = {c0, c1};
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      const int n = 2 * lk + h2, r = lr;
      if (n < 6) {
        if (!diag || n >= r) {
          const int col = it.cOff + (it.trans ? n : r), row = it.rOff + (it.trans ? r : n);
          atomicAdd(&T[col * 64 + row], -cv[h2]);
        }
      } else if (n == 6 && diag) {
        atomicAdd(&d.rhs[it.rhsIdx + r], -cv[h2]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Schur contraction, camera-row form (the default): CTA = one free camera a.  It walks a's
// observations (camera-major list); for each point i it reads the point's W rows -- ONE contiguous
// segment of the point-major W array -- and accumulates Y_a(i) W_b(i)^T for every free camera
// b >= a that sees i into warp-private shared-memory accumulators indexed by b - a (the band of
// the reduced camera system), plus Y_a(i) eb_i for the right-hand side.  At the end the four warp
// copies are summed in a fixed order and subtracted from the tiles: every 6x6 block (a, b) has
// exactly one owner, so there are no atomics and the result is deterministic.  Traffic: W is read
// ~(k_i + 1)/2 times per row from L2 in contiguous 144-byte rows (the pair-list kernel gathered
// 288 B per pair entry through index lists).  Used when the band (nSlots = max(b - a) + 1) fits
// shared memory; ba_schur_pairs remains for arbitrary co-visibility.
// ------------------------------------------------------------------------------------------
struct BaRowDst {
  int dst, rOff, cOff, trans;  // as in BaPairItem; dst < 0: cameras (a, a + slot) share no point
};


constexpr int BA_ROWS_WARPS = 4;

// visit = one (camera a, point i) pair of the camera-major list: {observation o, point i, first
// observation of i, number of observations of i}.  The records and the camera lists of the NEXT
// visits are prefetched while the current one is processed, so the dependent chain per visit is
// one level of loads (W rows) instead of four (cobs -> pt -> ptr -> cam -> W).
__global__ void __launch_bounds__(32 * BA_ROWS_WARPS)
ba_schur_rows(BaDev d, const int* __restrict__ cptr, const int4* __restrict__ visit,
              const BaRowDst* __restrict__ rowDst, int nSlots, const double* __restrict__ Vinv,
              const int* __restrict__ rhsIdx, int splits) {
  extern __shared__ double sacc[];  // [BA_ROWS_WARPS][stride]
  __shared__ int s_list[BA_ROWS_WARPS][32], s_slot[BA_ROWS_WARPS][32];
  __shared__ double s_Y[BA_ROWS_WARPS][18];
  // splits > 1 (few cameras, many observations: local BA): several CTAs share a camera's
  // observation list and add their partial blocks with atomics
  const int af = blockIdx.x / splits, part = blockIdx.x - af * splits;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int stride = nSlots * 36 + 8;
  double* acc = sacc + (size_t)w * stride;
  for (int e = lane; e < stride; e += 32) acc[e] = 0.0;
  __syncwarp();
  const int qend = cptr[af + 1], step = BA_ROWS_WARPS * splits;
  int q = cptr[af] + part * BA_ROWS_WARPS + w;
  const int4 none = make_int4(0, -1, 0, 0);
  int4 r0 = (q < qend) ? visit[q] : none;
  int4 r1 = (q + step < qend) ? visit[q + step] : none;
  int camCur = (r0.y >= 0 && lane < r0.w) ? d.cam[r0.z + lane] : -1;
  for (; q < qend; q += step) {
    const int4 r2 = (q + 2 * step < qend) ? visit[q + 2 * step] : none;
    const int camNext = (r1.y >= 0 && lane < r1.w) ? d.cam[r1.z + lane] : -1;
    const int o = r0.x, i = r0.y;
    if (i >= d.ncon) {  // fixed points do not enter the reduced system (warp-uniform)
      const double* Vi = Vinv + 6 * (size_t)i;
      const double i0 = Vi[0], i1 = Vi[1], i2 = Vi[2], i3 = Vi[3], i4 = Vi[4], i5 = Vi[5];
      if (lane < 18) {  // Y = W_o Vinv_i, entry (r, c) = lane
        const int r = lane / 3, c = lane - 3 * r;
        const double* wr = d.W + 18 * (size_t)o + 3 * r;
        const double v0 = (c == 0) ? i0 : (c == 1) ? i1 : i2;
        const double v1 = (c == 0) ? i1 : (c == 1) ? i3 : i4;
        const double v2 = (c == 0) ? i2 : (c == 1) ? i4 : i5;
        s_Y[w][lane] = wr[0] * v0 + wr[1] * v1 + wr[2] * v2;
      }
      __syncwarp();
      double Y[18];
#pragma unroll
      for (int k = 0; k < 18; ++k) Y[k] = s_Y[w][k];
      if (lane < 6) {  // rhs_a -= Y eb_i
        const double* eb = d.eb + 3 * (size_t)i;
        const double* yr = &s_Y[w][3 * lane];
        acc[nSlots * 36 + lane] += yr[0] * eb[0] + yr[1] * eb[1] + yr[2] * eb[2];
      }
      const int p0 = r0.z, p1 = r0.z + r0.w;
      for (int base = p0; base < p1; base += 32) {
        const int ob = base + lane;
        const int cm = (base == p0) ? camCur : ((ob < p1) ? d.cam[ob] : -1);
        const int bf = (ob < p1) ? cm - d.mcon : -1;
        const bool ok = bf >= af;
        const unsigned mask = __ballot_sync(0xffffffffu, ok);
        const int nq = __popc(mask);
        if (ok) {
          const int pos = __popc(mask & ((1u << lane) - 1u));
          s_list[w][pos] = ob;
          s_slot[w][pos] = bf - af;
        }
        __syncwarp();
        for (int t = lane; t < 6 * nq; t += 32) {
          const int j = t / 6, sc = t - 6 * j;
          const double* wb = d.W + 18 * (size_t)s_list[w][j] + 3 * sc;
          const double w0 = wb[0], w1 = wb[1], w2 = wb[2];
          double* ap = acc + s_slot[w][j] * 36 + sc;
#pragma unroll
          for (int r = 0; r < 6; ++r) ap[6 * r] += Y[3 * r] * w0 + Y[3 * r + 1] * w1 + Y[3 * r + 2] * w2;
        }
        __syncwarp();
      }
    }
    r0 = r1;
    r1 = r2;
    camCur = camNext;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < nSlots * 36; e += 32 * BA_ROWS_WARPS) {
    const int slot = e / 36, rs = e - 36 * slot, r = rs / 6, sc = rs - 6 * r;
    const BaRowDst rd = rowDst[(size_t)af * nSlots + slot];
    if (rd.dst < 0) continue;
    if (slot == 0 && sc < r) continue;  // diagonal block: lower part of the tile only
    double v = 0;
#pragma unroll
    for (int ww = 0; ww < BA_ROWS_WARPS; ++ww) v += sacc[(size_t)ww * stride + e];
    const int col = rd.cOff + (rd.trans ? sc : r), row = rd.rOff + (rd.trans ? r : sc);
    if (splits > 1) atomicAdd(&d.tiles[rd.dst + col * 64 + row], -v);
    else d.tiles[rd.dst + col * 64 + row] -= v;
  }
  if (threadIdx.x < 6) {
    double v = 0;
#pragma unroll
    for (int ww = 0; ww < BA_ROWS_WARPS; ++ww) v += sacc[(size_t)ww * stride + nSlots * 36 + threadIdx.x];
    if (splits > 1) atomicAdd(&d.rhs[rhsIdx[af] + threadIdx.x], -v);
    else d.rhs[rhsIdx[af] + threadIdx.x] -= v;
  }
}

// ------------------------------------------------------------------------------------------
// Back substitution, step 1, point-major, one observation per thread: t_i = sum_j W_ij^T da_j,
// collapsed per point with segmented shuffles; run heads add into acc3[n][3] (zeroed before).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ba_back_subst(BaDev d, const double* __restrict__ dpa, double* __restrict__ acc3) {
  const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  double acc[3] = {0, 0, 0};
  int key = -1 - lane;
  if (o < d.N) {
    const int j = d.cam[o], i = d.pt[o];
    key = i;
    if (i >= d.ncon && j >= d.mcon) {
      const double* Wp = d.W + 18 * (size_t)o;
      const double* da = dpa + 6 * (size_t)j;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double s = 0;
#pragma unroll
        for (int r = 0; r < 6; ++r) s += Wp[3 * r + c] * da[r];
        acc[c] = s;
      }
    }
  }
  seg_reduce<3>(acc, key, lane);
  const int kprev = __shfl_up_sync(0xffffffffu, key, 1);
  const bool head = (lane == 0) || (kprev != key);
  if (head && key >= d.ncon) {
#pragma unroll
    for (int c = 0; c < 3; ++c) atomicAdd(&acc3[3 * (size_t)key + c], acc[c]);
  }
}

// Back substitution, step 2, one point per thread: nb holds t_i on entry; db_i = V*_i^-1 (eb_i - t_i),
// nb <- pb + db, plus the point-side parts of |dp|^2, dL, |p|^2.
__global__ void __launch_bounds__(256)
ba_back_finish(BaDev d, const double* __restrict__ pb, double* __restrict__ nb,
               double* __restrict__ dpb, double mu) {
  __shared__ double s_red[8];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double dp2 = 0, dl = 0, p2 = 0;
  if (i < d.n) {
    double x[3] = {pb[3 * i], pb[3 * i + 1], pb[3 * i + 2]};
    double db[3] = {0, 0, 0};
    if (i >= d.ncon) {
      double Iv[6];
      inv3sym_mu(d.V + 6 * i, mu, Iv);
      const double* eb = d.eb + 3 * i;
      const double t0 = eb[0] - nb[3 * i], t1 = eb[1] - nb[3 * i + 1], t2 = eb[2] - nb[3 * i + 2];
      db[0] = Iv[0] * t0 + Iv[1] * t1 + Iv[2] * t2;
      db[1] = Iv[1] * t0 + Iv[3] * t1 + Iv[4] * t2;
      db[2] = Iv[2] * t0 + Iv[4] * t1 + Iv[5] * t2;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        dp2 += db[c] * db[c];
        dl += db[c] * (mu * db[c] + eb[c]);
        p2 += x[c] * x[c];
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      dpb[3 * i + c] = db[c];
      nb[3 * i + c] = x[c] + db[c];
    }
  }
  double t = block_sum_1(dp2, s_red);
  if (threadIdx.x == 0 && t != 0) atomicAdd(&d.sc[SC_DP_L2], t);
  t = block_sum_1(dl, s_red);
  if (threadIdx.x == 0 && t != 0) atomicAdd(&d.sc[SC_DL], t);
  t = block_sum_1(p2, s_red);
  if (threadIdx.x == 0 && t != 0) atomicAdd(&d.sc[SC_P_L2], t);
}

// Trial camera parameters na = pa + dpa and the camera-side parts of |dp|^2, dL, |p|^2.
// addSums: only one rank contributes the (replicated) camera-side sums.
__global__ void __launch_bounds__(256)
ba_cam_update(BaDev d, const double* __restrict__ pa, const double* __restrict__ sol,
              double* __restrict__ dpa, double* __restrict__ na, double mu, int addSums,
              double* __restrict__ zeroBuf, long long zeroCount) {
  __shared__ double s_red[8];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  for (long long q = t; q < zeroCount; q += (long long)gridDim.x * blockDim.x) zeroBuf[q] = 0.0;
  double dp2 = 0, dl = 0, p2 = 0;
  if (t < 6 * d.m) {
    const int j = t / 6, r = t - 6 * j;
    double dlt = 0;
    if (j >= d.mcon) {
      dlt = sol[d.solIdx[j - d.mcon] + r];
      dp2 = dlt * dlt;
      dl = dlt * (mu * dlt + d.ea[t]);
      p2 = pa[t] * pa[t];
    }
    dpa[t] = dlt;
    na[t] = pa[t] + dlt;
  }
  double s = block_sum_1(dp2, s_red);
  if (threadIdx.x == 0 && addSums && s != 0) atomicAdd(&d.sc[SC_DP_L2], s);
  s = block_sum_1(dl, s_red);
  if (threadIdx.x == 0 && addSums && s != 0) atomicAdd(&d.sc[SC_DL], s);
  s = block_sum_1(p2, s_red);
  if (threadIdx.x == 0 && addSums && s != 0) atomicAdd(&d.sc[SC_P_L2], s);
}

// ------------------------------------------------------------------------------------------
// Back substitution in two launches (instead of cam update | memset | point accumulate | point finish | cost):
//  ba_back_cams_points: blocks [0, nCamBlocks) form the trial camera parameters na = pa + da and the camera-side
//    parts of |dp|^2, dL, |p|^2 (ba_cam_update); the other blocks accumulate t_i = sum_j W_ij^T da_j per point
//    (one observation per thread, segmented shuffles, heads add into acc3, which the prep kernel zeroed),
//    reading da straight from the solution vector;
//  ba_finish_cost: one observation per thread: db_i = V*_i^-1 (e_b,i - t_i) recomputed from the point's
//    accumulator (read-only here; the FIRST observation of a point stores nb_i = pb_i + db_i and adds the
//    point-side sums), then the weighted squared residual at (na, nb_i) -> cost.  Extra threads finish the
//    points that have no observation.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ba_back_cams_points(BaDev d, const double* __restrict__ pa, const double* __restrict__ sol,
                    double* __restrict__ dpa, double* __restrict__ na, double mu, int addSums, int nCamBlocks,
                    double* __restrict__ acc3) {
  __shared__ double s_red[8];
  if ((int)blockIdx.x < nCamBlocks) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    double dp2 = 0, dl = 0, p2 = 0;
    if (t < 6 * d.m) {
      const int j = t / 6, r = t - 6 * j;
      double dlt = 0;
      if (j >= d.mcon) {
        dlt = sol[d.solIdx[j - d.mcon] + r];
        dp2 = dlt * dlt;
        dl = dlt * (mu * dlt + d.ea[t]);
        p2 = pa[t] * pa[t];
      }
      dpa[t] = dlt;
      na[t] = pa[t] + dlt;
    }
    double s = block_sum_1(dp2, s_red);
    if (threadIdx.x == 0 && addSums && s != 0) atomicAdd(&d.sc[SC_DP_L2], s);
    s = block_sum_1(dl, s_red);
    if (threadIdx.x == 0 && addSums && s != 0) atomicAdd(&d.sc[SC_DL], s);
    s = block_sum_1(p2, s_red);
    if (threadIdx.x == 0 && addSums && s != 0) atomicAdd(&d.sc[SC_P_L2], s);
    return;
  }
  const long long o = (long long)(blockIdx.x - nCamBlocks) * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  double acc[3] = {0, 0, 0};
  int key = -1 - lane;
  if (o < d.N) {
    const int j = d.cam[o], i = d.pt[o];
    key = i;
    if (i >= d.ncon && j >= d.mcon) {
      const double* Wp = d.W + 18 * (size_t)o;
      const double* da = sol + d.solIdx[j - d.mcon];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double s = 0;
#pragma unroll
        for (int r = 0; r < 6; ++r) s += Wp[3 * r + c] * da[r];
        acc[c] = s;
      }
    }
  }
  seg_reduce<3>(acc, key, lane);
  const int kprev = __shfl_up_sync(0xffffffffu, key, 1);
  const bool head = (lane == 0) || (kprev != key);
  if (head && key >= d.ncon) {
#pragma unroll
    for (int c = 0; c < 3; ++c) atomicAdd(&acc3[3 * (size_t)key + c], acc[c]);
  }
}

__global__ void __launch_bounds__(256)
ba_finish_cost(BaDev d, const double* __restrict__ na, const double* __restrict__ pb, double* __restrict__ nb,
               const double* __restrict__ acc3, const double* __restrict__ Vinv, double mu) {
  __shared__ double s_red[8];
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double dp2 = 0, dl = 0, p2 = 0, c = 0;
  long long o = -1;
  int i = -1;
  bool store = false;
  if (t < d.N) {
    o = t;
    i = d.pt[o];
    store = (d.ptr[i] == o);
  } else if (t < d.N + d.n) {
    i = (int)(t - d.N);
    store = (d.ptr[i + 1] == d.ptr[i]);  // a point nobody observes is finished here
    if (!store) i = -1;
  }
  if (i >= 0) {
    double x[3] = {pb[3 * (size_t)i], pb[3 * (size_t)i + 1], pb[3 * (size_t)i + 2]};
    double db[3] = {0, 0, 0};
    if (i >= d.ncon) {
      const double* Iv = Vinv + 6 * (size_t)i;
      const double* eb = d.eb + 3 * (size_t)i;
      const double t0 = eb[0] - acc3[3 * (size_t)i], t1 = eb[1] - acc3[3 * (size_t)i + 1],
                   t2 = eb[2] - acc3[3 * (size_t)i + 2];
      db[0] = Iv[0] * t0 + Iv[1] * t1 + Iv[2] * t2;
      db[1] = Iv[1] * t0 + Iv[3] * t1 + Iv[4] * t2;
      db[2] = Iv[2] * t0 + Iv[4] * t1 + Iv[5] * t2;
      if (store) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          dp2 += db[k] * db[k];
          dl += db[k] * (mu * db[k] + eb[k]);
          p2 += x[k] * x[k];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) x[k] += db[k];
    if (store) {
#pragma unroll
      for (int k = 0; k < 3; ++k) nb[3 * (size_t)i + k] = x[k];
    }
    if (o >= 0) {
      const int j = d.cam[o];
      double h[2];
      ba_project<false, false>(d.camK + 5 * j, d.camR0 + 9 * j, na + 6 * j, x, h, nullptr, nullptr);
      const double dx = d.xy[2 * o] - h[0], dy = d.xy[2 * o + 1] - h[1];
      c = d.wgt[o] * (dx * dx + dy * dy);
    }
  }
  double s = block_sum_1(c, s_red);
  if (threadIdx.x == 0) {
    if (isfinite(s)) {
      if (s != 0) atomicAdd(&d.sc[SC_COST], s);
    } else {
      d.sc[SC_NONFINITE] = 1.0;
    }
  }
  s = block_sum_1(dp2, s_red);
  if (threadIdx.x == 0 && s != 0) atomicAdd(&d.sc[SC_DP_L2], s);
  s = block_sum_1(dl, s_red);
  if (threadIdx.x == 0 && s != 0) atomicAdd(&d.sc[SC_DL], s);
  s = block_sum_1(p2, s_red);
  if (threadIdx.x == 0 && s != 0) atomicAdd(&d.sc[SC_P_L2], s);
}

}  // namespace coslam
