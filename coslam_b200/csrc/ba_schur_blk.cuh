// ba_schur_blk.cuh -- camera-BLOCK form of the Schur contraction on the fp64 tensor cores (sm_100a).
//
// The reduced camera system S = U* - sum_i W_i V*_i^-1 W_i^T (SBA's Schur complement; reference call
// sba_motstr_levmar_x, app/SL_CoSLAMBA.cpp:360-363 / bundleAdjustRobust, app/SL_CoSLAMRobustBA.cpp:174)
// is formed per PAIR OF CAMERA BLOCKS instead of per camera pair.  Free cameras are grouped in blocks
// of BA_CB = 4 (24 rows) by co-visibility: a greedy pass over the camera-pair counts puts the cameras
// that share the most points into one block (ba.cu), so that a visit is dense.  A VISIT is one free point seen from block A and block B
// (A <= B): up to 4 + 4 observations.  Its contribution to the 24 x 24 block (A, B) is the rank-3 product
//        [W_a]_(24 x 3) . ( V*^-1 . [W_b]^T )_(3 x 24)          (rows of absent cameras are zero)
// so a run of visits of one block pair is ONE dense contraction C(24 x 24) += A(24 x 3n) . T(3n x 24)
// on DMMA m8n8k4: 4 visits fill three k-steps exactly, 27 DMMAs per 4 visits (36 with the right-hand
// side column of diagonal block pairs).  Against the per-camera-pair lists (ba_schur_pairs_st) a point
// seen by 8 cameras of two blocks loads 8 rows of W for 16 pair entries instead of 32: the L2 -> SM
// traffic, which bounds the pair kernels (profiles/r2n), drops by ~3x, and the multiply-adds move from
// 162 DFMA per entry to 6.75 DMMA per visit.
//
// Work lists (device-built at solver set-up, like the pair lists): ba_visits_build counts / fills the
// visits per block-pair bucket, thread per point.
#pragma once

namespace coslam {

constexpr int BA_CB = 4;                 // cameras per block
constexpr int BA_VIS_MAXB = 24;          // blocks one point may touch (else the pair kernel is used)
constexpr int BA_BLK_NV = 4;             // visits per batch (K = 12 = 3 DMMA k-steps)
constexpr int BA_BLK_VSTRIDE = 154;      // doubles per staged visit: A rows 72 | B rows 72 | V*^-1 6 | e_b 3 | pad
constexpr int BA_BLK_TNS = 36;           // leading dimension of T (conflict-free B fragments)
constexpr int BA_BLK_WARPS = 4;
// per warp: 2 staging buffers + T + 2 descriptor buffers (as doubles)
constexpr int BA_BLK_RAW = BA_BLK_NV * BA_BLK_VSTRIDE;                       // 616 doubles
constexpr int BA_BLK_TSZ = 3 * BA_BLK_NV * BA_BLK_TNS;                       // 432 doubles
constexpr int BA_BLK_DESC = BA_BLK_NV * 12 / 2;                              // 24 doubles (48 ints)
constexpr int BA_BLK_WARP_DOUBLES = 2 * BA_BLK_RAW + BA_BLK_TSZ + 2 * BA_BLK_DESC;  // 1712
constexpr int BA_BLK_SMEM = BA_BLK_WARPS * BA_BLK_WARP_DOUBLES * 8;          // 54784 B per CTA

struct BaVisit {  // 48 B = 3 x int4
  int pt;
  int a[4];  // observation index of camera 4A + s, or -1
  int b[4];  // observation index of camera 4B + s, or -1 (unused when A == B)
  int pad[3];
};

struct BaBlkItem {
  int A, B;        // block pair, A <= B
  int begin, end;  // visit range
  int tab;         // index of the destination table of this block pair
  int pad[3];
};

// destinations of the 16 6x6 sub-blocks (a, b) of a block pair: tiles[dst + (cOff + (trans ? c : r)) * 64
// + rOff + (trans ? r : c)] as in BaPairItem; dst < 0: the two cameras share no point (sub-block is zero)
struct BaBlkDst {
  int dst[16], rOff[16], cOff[16], trans[16];
  int rhsIdx[4];  // rhs index of camera 4A + a (diagonal block pairs)
};

// ---- work-list construction: thread per free point ----
template <bool FILL>
__global__ void __launch_bounds__(128)
ba_visits_build(BaDev d, int nB, const int* __restrict__ camBlock, const int* __restrict__ camSlot,
                unsigned* __restrict__ cnt, const unsigned* __restrict__ off, BaVisit* __restrict__ visits,
                int* __restrict__ overflow) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + d.ncon;
  if (i >= d.n) return;
  int blk[BA_VIS_MAXB];
  int slot[BA_VIS_MAXB][BA_CB];
  int nblk = 0;
  const long long o0 = d.ptr[i], o1 = d.ptr[i + 1];
  for (long long o = o0; o < o1; ++o) {
    const int jf = d.cam[o] - d.mcon;
    if (jf < 0) continue;
    const int B = camBlock[jf], sl = camSlot[jf];
    int x = 0;
    while (x < nblk && blk[x] != B) ++x;
    if (x == nblk) {
      if (nblk == BA_VIS_MAXB) {
        atomicExch(overflow, 1);
        return;
      }
      blk[nblk] = B;
#pragma unroll
      for (int q = 0; q < BA_CB; ++q) slot[nblk][q] = -1;
      ++nblk;
    }
    slot[x][sl] = (int)o;
  }
  for (int x = 0; x < nblk; ++x)
    for (int y = 0; y < nblk; ++y) {
      if (blk[x] > blk[y]) continue;
      const size_t bucket = (size_t)blk[x] * nB + blk[y];
      const unsigned k = atomicAdd(&cnt[bucket], 1u);
      if (FILL) {
        BaVisit v;
        v.pt = i;
#pragma unroll
        for (int q = 0; q < BA_CB; ++q) {
          v.a[q] = slot[x][q];
          v.b[q] = (x == y) ? -1 : slot[y][q];
        }
        v.pad[0] = v.pad[1] = v.pad[2] = 0;
        int4* dst = reinterpret_cast<int4*>(visits + off[bucket] + k);
        const int4* src = reinterpret_cast<const int4*>(&v);
        dst[0] = src[0];
        dst[1] = src[1];
        dst[2] = src[2];
      }
    }
}

__device__ __forceinline__ void ba_blk_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void ba_blk_cp16(double* smemDst, const double* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smemDst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void ba_blk_cp8(double* smemDst, const double* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smemDst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gsrc) : "memory");
}

// stage the rows of the NV visits whose descriptors sit in desc[] (48 ints; pt < 0: no visit)
__device__ __forceinline__ void ba_blk_stage(const BaDev& d, const double* __restrict__ Vinv, const int* desc,
                                             bool diag, int lane, double* raw) {
  const int segsPerVisit = diag ? 39 : 75;  // A rows 36 (+ B rows 36) + V*^-1 3
  const int total = BA_BLK_NV * segsPerVisit;
  for (int sidx = lane; sidx < total; sidx += 32) {
    const int v = sidx / segsPerVisit, p = sidx - v * segsPerVisit;
    const int pt = desc[12 * v];
    double* base = raw + v * BA_BLK_VSTRIDE;
    const int nrow = diag ? 36 : 72;
    if (p < nrow) {
      const int row = p / 9, q = p - 9 * row;  // row 0..3: block A, 4..7: block B
      const int obs = (pt >= 0) ? desc[12 * v + 1 + row] : -1;
      double* dst = base + row * 18 + 2 * q;
      if (obs >= 0)
        ba_blk_cp16(dst, d.W + 18 * (size_t)obs + 2 * q);
      else
        *reinterpret_cast<double2*>(dst) = make_double2(0.0, 0.0);
    } else {
      const int q = p - nrow;
      double* dst = base + 144 + 2 * q;
      if (pt >= 0)
        ba_blk_cp16(dst, Vinv + 6 * (size_t)pt + 2 * q);
      else
        *reinterpret_cast<double2*>(dst) = make_double2(0.0, 0.0);
    }
  }
  if (diag && lane < 3 * BA_BLK_NV) {  // e_b: 3 doubles per visit (8-byte aligned only)
    const int v = lane / 3, q = lane - 3 * v;
    const int pt = desc[12 * v];
    double* dst = raw + v * BA_BLK_VSTRIDE + 150 + q;
    if (pt >= 0)
      ba_blk_cp8(dst, d.eb + 3 * (size_t)pt + q);
    else
      *dst = 0.0;
  }
  asm volatile("cp.async.commit_group;\n" ::: "memory");
}

__global__ void __launch_bounds__(32 * BA_BLK_WARPS, 4)
ba_schur_blk(BaDev d, const BaBlkItem* __restrict__ items, int nItems, const BaVisit* __restrict__ visits,
             const BaBlkDst* __restrict__ tabs, const double* __restrict__ Vinv) {
  extern __shared__ __align__(16) double s_blk[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * BA_BLK_WARPS + w;
  if (item >= nItems) return;
  const BaBlkItem it = items[item];
  const bool diag = (it.A == it.B);
  double* raw0 = s_blk + (size_t)w * BA_BLK_WARP_DOUBLES;
  double* sT = raw0 + 2 * BA_BLK_RAW;
  int* desc0 = reinterpret_cast<int*>(sT + BA_BLK_TSZ);
  const int g = lane >> 2, kk = lane & 3;
  // T columns 25..31 are never written: zero the whole T once
  for (int q = lane; q < BA_BLK_TSZ; q += 32) sT[q] = 0.0;
  // loop-invariant fragment offsets: A(m, k) = raw[v * VSTRIDE + m * 3 + c], k = 3 v + c
  int offA[3];
#pragma unroll
  for (int ks = 0; ks < 3; ++ks) {
    const int k = 4 * ks + kk, v = (k * 11) >> 5, c = k - 3 * v;
    offA[ks] = v * BA_BLK_VSTRIDE + c + 3 * g;
  }
  double acc[3][4][2];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

  const int nv = it.end - it.begin;
  const int nb = (nv + BA_BLK_NV - 1) / BA_BLK_NV;
  // descriptor pipeline: lanes 0..11 hold one int4 of the batch two ahead
  auto load_desc = [&](int batch) -> int4 {
    int4 r = make_int4(-1, -1, -1, -1);
    if (lane < 3 * BA_BLK_NV) {
      const int v = batch * BA_BLK_NV + lane / 3;
      if (v < nv) r = __ldg(reinterpret_cast<const int4*>(visits + it.begin + v) + (lane % 3));
    }
    return r;
  };
  auto put_desc = [&](int4 r, int buf) {
    if (lane < 3 * BA_BLK_NV) reinterpret_cast<int4*>(desc0 + 48 * buf)[lane] = r;
  };
  int4 dcur = load_desc(0);
  int4 dnxt = load_desc(1);
  put_desc(dcur, 0);
  __syncwarp();
  ba_blk_stage(d, Vinv, desc0, diag, lane, raw0);
  for (int b = 0; b < nb; ++b) {
    const int buf = b & 1;
    if (b + 1 < nb) {
      put_desc(dnxt, buf ^ 1);
      __syncwarp();
      ba_blk_stage(d, Vinv, desc0 + 48 * (buf ^ 1), diag, lane, raw0 + (buf ^ 1) * BA_BLK_RAW);
      dnxt = load_desc(b + 2);
      asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    }
    __syncwarp();
    const double* raw = raw0 + buf * BA_BLK_RAW;
    // ---- T(3 v + c, n) = sum_c' V*^-1(c, c') W_b(n, c'),  n < 24;  T(., 24) = V*^-1 e_b (diagonal pairs)
    const int ncol = diag ? 25 : 24;
    for (int q = lane; q < BA_BLK_NV * ncol; q += 32) {
      const int v = q / ncol, n = q - v * ncol;
      const double* rv = raw + v * BA_BLK_VSTRIDE;
      const double* wb = (n < 24) ? rv + (diag ? 0 : 72) + 3 * n : rv + 150;
      const double b0 = wb[0], b1 = wb[1], b2 = wb[2];
      const double i0 = rv[144], i1 = rv[145], i2 = rv[146], i3 = rv[147], i4 = rv[148], i5 = rv[149];
      sT[(3 * v) * BA_BLK_TNS + n] = i0 * b0 + i1 * b1 + i2 * b2;
      sT[(3 * v + 1) * BA_BLK_TNS + n] = i1 * b0 + i3 * b1 + i4 * b2;
      sT[(3 * v + 2) * BA_BLK_TNS + n] = i2 * b0 + i4 * b1 + i5 * b2;
    }
    __syncwarp();
    // ---- C(24 x 24|32) += A(24 x 12) T(12 x 24|32)
#pragma unroll
    for (int ks = 0; ks < 3; ++ks) {
      double af[3], bf[4];
#pragma unroll
      for (int mt = 0; mt < 3; ++mt) af[mt] = raw[offA[ks] + 24 * mt];
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) bf[nt] = sT[(4 * ks + kk) * BA_BLK_TNS + 8 * nt + g];
      bf[3] = diag ? sT[(4 * ks + kk) * BA_BLK_TNS + 24 + g] : 0.0;
#pragma unroll
      for (int mt = 0; mt < 3; ++mt) {
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) ba_blk_dmma(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
        if (diag) ba_blk_dmma(acc[mt][3][0], acc[mt][3][1], af[mt], bf[3]);
      }
    }
    __syncwarp();  // raw[buf] and T are free again
  }
  // ---- flush: lane holds C(8 mt + g, 8 nt + 2 kk + {0, 1})
  const BaBlkDst* tab = tabs + it.tab;
#pragma unroll
  for (int mt = 0; mt < 3; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int M = 8 * mt + g, N = 8 * nt + 2 * kk + h;
        const double val = acc[mt][nt][h];
        if (val == 0.0) continue;  // absent cameras, unused columns
        const int a = M / 6, r = M - 6 * a;
        if (N < 24) {
          const int bb = N / 6, c = N - 6 * bb;
          if (diag && (a > bb || (a == bb && c < r))) continue;
          const int t = a * 4 + bb;
          const int dst = __ldg(&tab->dst[t]);
          if (dst < 0) continue;
          const int tr = __ldg(&tab->trans[t]);
          const int col = __ldg(&tab->cOff[t]) + (tr ? c : r), row = __ldg(&tab->rOff[t]) + (tr ? r : c);
          atomicAdd(&d.tiles[(size_t)dst + col * 64 + row], -val);
        } else if (diag && N == 24) {
          const int ri = __ldg(&tab->rhsIdx[a]);
          if (ri >= 0) atomicAdd(&d.rhs[ri + r], -val);
        }
      }
}

}  // namespace coslam
