// posegraph.cu -- post-BA pose-graph spreading for chain graphs, sm_100a.
//
// Replaces, for the graphs RobustBundleRTS::constructCameraGraphs builds (one chain per camera, node k
// -> node k+1, key-frame nodes fixed; reference app/SL_CoSLAMRobustBA.cpp:182-232), the two sparse
// least-squares solves of RobustBundleRTS::updateNonKeyCameraPoses (:233-250):
//   GlobalPoseGraph::computeNewCameraRotations     slam/SL_GlobalPoseEstimation.cpp:52-218
//       min sum_k |X_{k+1} - R_k X_k|_F^2 over the free nodes, then nearest rotation (:205-211)
//   GlobalPoseGraph::computeNewCameraTranslations  slam/SL_GlobalPoseEstimation.cpp:220-359
//       min sum_k |t_{k+1} - R_k t_k - t_k,k+1|^2
// R_k is orthogonal, so with P_k = R_{k-1} ... R_a (product of the edge rotations from the fixed node a on
// the left of a run of free nodes) the substitution Y_k = P_k^T X_k turns every residual into
// Y_{k+1} - Y_k: the least-squares solution is the LINEAR INTERPOLATION of Y between the two fixed ends
// (constant when only one end is fixed), and the same substitution spreads the translation residual of a
// run uniformly over its edges.  No linear system is left: the work is one segmented scan of rigid
// transforms (P_k, q_k) per chain plus a 3x3 polar projection per free node.
//
// Data layout (device, fp64): nodes of all chains concatenated in frame order; R row-major 3x3 / node,
// t 3 / node, edge k (node k -> k+1) stored at index k (the slot of a chain's last node is unused).
// Kernel: one CTA per chain, each thread owns a contiguous chunk of nodes; chunk composites are scanned
// across the CTA in shared memory; newR/newt hold (P_k, q_k) between the two passes.
// Algorithmic bytes / node: 216 read (R, t, edge R, edge t, flag) + 96 written.
#include <cmath>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace coslam {
namespace {

constexpr int PG_THREADS = 256;

struct Rigid {
  double R[9];
  double t[3];
};

__device__ __forceinline__ void rigid_identity(Rigid& a) {
#pragma unroll
  for (int i = 0; i < 9; ++i) a.R[i] = (i % 4 == 0) ? 1.0 : 0.0;
  a.t[0] = a.t[1] = a.t[2] = 0.0;
}
// c = b after a :  x -> b.R (a.R x + a.t) + b.t
__device__ __forceinline__ void rigid_compose(const Rigid& a, const Rigid& b, Rigid& c) {
  Rigid o;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j)
      o.R[3 * i + j] = b.R[3 * i] * a.R[j] + b.R[3 * i + 1] * a.R[3 + j] + b.R[3 * i + 2] * a.R[6 + j];
    o.t[i] = b.R[3 * i] * a.t[0] + b.R[3 * i + 1] * a.t[1] + b.R[3 * i + 2] * a.t[2] + b.t[i];
  }
  c = o;
}
__device__ __forceinline__ void load_edge(const double* eR, const double* et, int k, Rigid& e) {
#pragma unroll
  for (int i = 0; i < 9; ++i) e.R[i] = eR[9 * (size_t)k + i];
#pragma unroll
  for (int i = 0; i < 3; ++i) e.t[i] = et[3 * (size_t)k + i];
}

// nearest rotation (Frobenius) of a 3x3: one-sided Jacobi SVD M = U S V^T, R = U diag(1,1,det) V^T with
// the sign flip on the smallest singular direction.
__device__ void nearest_rotation(const double* M, double* Rout) {
  double A[9], W[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    A[i] = M[i];
    W[i] = (i % 4 == 0) ? 1.0 : 0.0;
  }
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
      double a = 0, b = 0, c = 0;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        a += A[3 * i + p] * A[3 * i + p];
        b += A[3 * i + q] * A[3 * i + q];
        c += A[3 * i + p] * A[3 * i + q];
      }
      const double rel = fabs(c) / sqrt(a * b + 1e-300);
      off = fmax(off, rel);
      if (rel < 1e-18) continue;
      const double zeta = (b - a) / (2.0 * c);
      const double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
      const double cs = 1.0 / sqrt(1.0 + tt * tt), sn = cs * tt;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double x = A[3 * i + p], y = A[3 * i + q];
        A[3 * i + p] = cs * x - sn * y;
        A[3 * i + q] = sn * x + cs * y;
        x = W[3 * i + p];
        y = W[3 * i + q];
        W[3 * i + p] = cs * x - sn * y;
        W[3 * i + q] = sn * x + cs * y;
      }
    }
    if (off < 1e-17) break;
  }
  double S[3], U[9];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double s = sqrt(A[j] * A[j] + A[3 + j] * A[3 + j] + A[6 + j] * A[6 + j]);
    S[j] = s;
    const double inv = s > 0 ? 1.0 / s : 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) U[3 * i + j] = A[3 * i + j] * inv;
  }
  double Q[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) Q[3 * i + j] = U[3 * i] * W[3 * j] + U[3 * i + 1] * W[3 * j + 1] + U[3 * i + 2] * W[3 * j + 2];
  const double det = Q[0] * (Q[4] * Q[8] - Q[5] * Q[7]) - Q[1] * (Q[3] * Q[8] - Q[5] * Q[6]) +
                     Q[2] * (Q[3] * Q[7] - Q[4] * Q[6]);
  if (det < 0) {
    int sm = 0;
    if (S[1] < S[sm]) sm = 1;
    if (S[2] < S[sm]) sm = 2;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const double u = sm == 0 ? U[3 * i] : (sm == 1 ? U[3 * i + 1] : U[3 * i + 2]);
        const double w = sm == 0 ? W[3 * j] : (sm == 1 ? W[3 * j + 1] : W[3 * j + 2]);
        Q[3 * i + j] -= 2.0 * u * w;
      }
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) Rout[i] = Q[i];
}

// One CTA per chain.  Element k of the scan is the edge (k-1 -> k); a run restarts at the element that
// FOLLOWS a fixed node, so the composite stored for node k maps the coordinates of the nearest fixed node
// on its left (or of the chain's first node) to node k, and a fixed node on the right end of a run
// carries the composite of the whole run.
__global__ void __launch_bounds__(PG_THREADS)
posegraph_chain_kernel(const int* __restrict__ chainOff, const unsigned char* __restrict__ fixed,
                       const double* __restrict__ R, const double* __restrict__ t,
                       const double* __restrict__ eR, const double* __restrict__ et,
                       double* __restrict__ newR, double* __restrict__ newt, int* __restrict__ anchor,
                       int* __restrict__ status) {
  __shared__ Rigid s_agg[PG_THREADS];
  __shared__ int s_head[PG_THREADS];   // chunk contains a run head
  __shared__ int s_lastFix[PG_THREADS];
  __shared__ int s_firstFix[PG_THREADS];
  const int c = blockIdx.x, tid = threadIdx.x;
  const int n0 = chainOff[c], n = chainOff[c + 1] - n0;
  if (n <= 0) return;
  const int per = (n + PG_THREADS - 1) / PG_THREADS;
  const int lo = min(n, tid * per), hi = min(n, lo + per);

  // pass 1: composite of the chunk (from its last run head on), fixed-node positions
  Rigid acc;
  rigid_identity(acc);
  int head = 0, lastFix = -1, firstFix = n;
  for (int k = lo; k < hi; ++k) {
    const int g = n0 + k;
    if (k == 0 || fixed[g - 1]) {
      head = 1;
      rigid_identity(acc);
    }
    if (k > 0) {
      Rigid e;
      load_edge(eR, et, g - 1, e);
      if (!fixed[g - 1]) rigid_compose(acc, e, acc);
      else acc = e;
    }
    if (fixed[g]) {
      lastFix = k;
      if (firstFix == n) firstFix = k;
    }
  }
  s_agg[tid] = acc;
  s_head[tid] = head;
  s_lastFix[tid] = lastFix;
  s_firstFix[tid] = firstFix;
  __syncthreads();
  // segmented inclusive scan over the chunk composites (Hillis-Steele), max / min scans of the fixed ids
  for (int d = 1; d < PG_THREADS; d <<= 1) {
    Rigid mine = s_agg[tid], left;
    int mh = s_head[tid], lh = 0, lf = s_lastFix[tid];
    const bool has = tid >= d;
    if (has) {
      left = s_agg[tid - d];
      lh = s_head[tid - d];
      lf = max(lf, s_lastFix[tid - d]);
    }
    int ff = s_firstFix[tid];
    if (tid + d < PG_THREADS) ff = min(ff, s_firstFix[tid + d]);
    __syncthreads();
    if (has && !mh) {
      rigid_compose(left, mine, mine);
      s_agg[tid] = mine;
      s_head[tid] = lh;
    }
    s_lastFix[tid] = lf;
    s_firstFix[tid] = ff;
    __syncthreads();
  }
  // pass 2: per-node composite and anchors
  Rigid pre;
  if (tid > 0) pre = s_agg[tid - 1];
  else rigid_identity(pre);
  int a = tid > 0 ? s_lastFix[tid - 1] : -1;
  for (int k = lo; k < hi; ++k) {
    const int g = n0 + k;
    if (k == 0) rigid_identity(pre);
    else {
      Rigid e;
      load_edge(eR, et, g - 1, e);
      if (!fixed[g - 1]) rigid_compose(pre, e, pre);
      else pre = e;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) newR[9 * (size_t)g + i] = pre.R[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) newt[3 * (size_t)g + i] = pre.t[i];
    anchor[2 * (size_t)g] = a;
    if (fixed[g]) a = k;
  }
  // right anchors: walk the chunk backwards
  int b = tid + 1 < PG_THREADS ? s_firstFix[tid + 1] : n;
  for (int k = hi - 1; k >= lo; --k) {
    const int g = n0 + k;
    anchor[2 * (size_t)g + 1] = b;
    if (fixed[g]) b = k;
  }
  __syncthreads();  // (P,q) of the whole chain are in newR/newt (global, same CTA -> visible after the barrier)
  __threadfence_block();

  // pass 3: the closed-form least-squares solution per free node, in place.  A node only reads the
  // composites of ITSELF and of its right anchor (a fixed node, overwritten only by its own thread with
  // its input pose afterwards), so fixed nodes are finalised after a second barrier.
  for (int k = tid; k < n; k += PG_THREADS) {
    const int g = n0 + k;
    if (fixed[g]) continue;
    const int la = anchor[2 * (size_t)g], rb = anchor[2 * (size_t)g + 1];
    if (la < 0 && rb >= n) {
      atomicExch(status, 1 + c);  // no fixed node in this chain: the system is rank deficient
      continue;
    }
    double P[9], q[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) P[i] = newR[9 * (size_t)g + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) q[i] = newt[3 * (size_t)g + i];
    double X[9], tk[3];
    if (rb >= n) {  // only the left end is fixed: exact propagation
      const double* Ra = R + 9 * (size_t)(n0 + la);
      const double* ta = t + 3 * (size_t)(n0 + la);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) X[3 * i + j] = P[3 * i] * Ra[j] + P[3 * i + 1] * Ra[3 + j] + P[3 * i + 2] * Ra[6 + j];
        tk[i] = P[3 * i] * ta[0] + P[3 * i + 1] * ta[1] + P[3 * i + 2] * ta[2] + q[i];
      }
    } else {
      const int gb = n0 + rb;
      double Pb[9], qb[3];
#pragma unroll
      for (int i = 0; i < 9; ++i) Pb[i] = newR[9 * (size_t)gb + i];
#pragma unroll
      for (int i = 0; i < 3; ++i) qb[i] = newt[3 * (size_t)gb + i];
      const double* Rb = R + 9 * (size_t)gb;
      const double* tb = t + 3 * (size_t)gb;
      // Yb = Pb^T Rb, ub = Pb^T (tb - qb)
      double Yb[9], ub[3], d[3] = {tb[0] - qb[0], tb[1] - qb[1], tb[2] - qb[2]};
#pragma unroll
      for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) Yb[3 * i + j] = Pb[i] * Rb[j] + Pb[3 + i] * Rb[3 + j] + Pb[6 + i] * Rb[6 + j];
        ub[i] = Pb[i] * d[0] + Pb[3 + i] * d[1] + Pb[6 + i] * d[2];
      }
      double Y[9], u[3];
      if (la >= 0) {
        const double* Ra = R + 9 * (size_t)(n0 + la);
        const double* ta = t + 3 * (size_t)(n0 + la);
        const double w = (double)(k - la) / (double)(rb - la);
#pragma unroll
        for (int i = 0; i < 9; ++i) Y[i] = Ra[i] + w * (Yb[i] - Ra[i]);
#pragma unroll
        for (int i = 0; i < 3; ++i) u[i] = ta[i] + w * (ub[i] - ta[i]);
      } else {  // only the right end is fixed
#pragma unroll
        for (int i = 0; i < 9; ++i) Y[i] = Yb[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) u[i] = ub[i];
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) X[3 * i + j] = P[3 * i] * Y[j] + P[3 * i + 1] * Y[3 + j] + P[3 * i + 2] * Y[6 + j];
        tk[i] = P[3 * i] * u[0] + P[3 * i + 1] * u[1] + P[3 * i + 2] * u[2] + q[i];
      }
    }
    double Rn[9];
    nearest_rotation(X, Rn);
    // the right anchor's composite must stay intact until every node of the run has read it: free
    // nodes only overwrite their OWN slot
#pragma unroll
    for (int i = 0; i < 9; ++i) newR[9 * (size_t)g + i] = Rn[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) newt[3 * (size_t)g + i] = tk[i];
  }
  __syncthreads();
  for (int k = tid; k < n; k += PG_THREADS) {
    const int g = n0 + k;
    if (!fixed[g]) continue;
#pragma unroll
    for (int i = 0; i < 9; ++i) newR[9 * (size_t)g + i] = R[9 * (size_t)g + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) newt[3 * (size_t)g + i] = t[3 * (size_t)g + i];
  }
}

}  // namespace
}  // namespace coslam

using namespace coslam;

extern "C" int cosl_posegraph_spread_chains(int nChains, const int* chainOff, const uint8_t* fixed,
                                            const double* R, const double* t, const double* eR,
                                            const double* et, double* newR, double* newt, int device) {
  if (nChains < 0 || !chainOff) return set_error(COSL_E_INVALID, "cosl_posegraph_spread_chains: bad chain table");
  if (nChains == 0) return COSL_OK;
  const int N = chainOff[nChains];
  for (int c = 0; c < nChains; ++c)
    if (chainOff[c + 1] < chainOff[c] || chainOff[0] != 0)
      return set_error(COSL_E_INVALID, "cosl_posegraph_spread_chains: chainOff must start at 0 and be non-decreasing");
  if (N == 0) return COSL_OK;
  if (!fixed || !R || !t || !eR || !et || !newR || !newt)
    return set_error(COSL_E_INVALID, "cosl_posegraph_spread_chains: null array");
  COSL_CUDA(cudaSetDevice(device));
  // one device slab: [R 9N | t 3N | eR 9N | et 3N | newR 9N | newt 3N] doubles, [anchor 2N | off C+1 | status] ints, flags
  const size_t nd = (size_t)36 * N, ni = (size_t)2 * N + nChains + 2;
  const size_t bytes = nd * 8 + ni * 4 + (size_t)N;
  // stream-ordered allocation from the device's default pool: after the first call the slab comes out of
  // the pool without a device synchronisation (cudaMalloc / cudaFree cost more than the kernel)
  {
    static std::mutex mu;
    static bool kept[64] = {};
    std::lock_guard<std::mutex> lk(mu);
    if (device >= 0 && device < 64 && !kept[device]) {
      kept[device] = true;
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
    }
  }
  cudaStream_t s = cudaStreamPerThread;
  char* slab = nullptr;
  COSL_CUDA(cudaMallocAsync(&slab, bytes, s));
  double* dR = reinterpret_cast<double*>(slab);
  double *dt = dR + 9 * (size_t)N, *deR = dt + 3 * (size_t)N, *det = deR + 9 * (size_t)N;
  double *dnR = det + 3 * (size_t)N, *dnt = dnR + 9 * (size_t)N;
  int* dAnchor = reinterpret_cast<int*>(dnt + 3 * (size_t)N);
  int *dOff = dAnchor + 2 * (size_t)N, *dStatus = dOff + nChains + 1;
  unsigned char* dFixed = reinterpret_cast<unsigned char*>(dStatus + 1);
  int rc = COSL_OK;
  auto fail = [&](cudaError_t e, const char* what) {
    rc = set_error(COSL_E_CUDA, "cosl_posegraph_spread_chains: %s: %s", what, cudaGetErrorString(e));
  };
  cudaError_t e;
#define PG_CP(dst, src, n)                                                              \
  if (rc == COSL_OK && (e = cudaMemcpyAsync(dst, src, (n), cudaMemcpyHostToDevice, s)) != cudaSuccess) fail(e, "H2D")
  PG_CP(dR, R, 72 * (size_t)N);
  PG_CP(dt, t, 24 * (size_t)N);
  PG_CP(deR, eR, 72 * (size_t)N);
  PG_CP(det, et, 24 * (size_t)N);
  PG_CP(dOff, chainOff, 4 * (size_t)(nChains + 1));
  PG_CP(dFixed, fixed, (size_t)N);
#undef PG_CP
  if (rc == COSL_OK && (e = cudaMemsetAsync(dStatus, 0, 4, s)) != cudaSuccess) fail(e, "memset");
  int status = 0;
  if (rc == COSL_OK) {
    COSL_LAUNCH(posegraph_chain_kernel, nChains, PG_THREADS, 0, s, dOff, dFixed, dR, dt, deR, det, dnR, dnt,
                dAnchor, dStatus);
    if ((e = cudaGetLastError()) != cudaSuccess) fail(e, "launch");
  }
  if (rc == COSL_OK && (e = cudaMemcpyAsync(newR, dnR, 72 * (size_t)N, cudaMemcpyDeviceToHost, s)) != cudaSuccess) fail(e, "D2H");
  if (rc == COSL_OK && (e = cudaMemcpyAsync(newt, dnt, 24 * (size_t)N, cudaMemcpyDeviceToHost, s)) != cudaSuccess) fail(e, "D2H");
  if (rc == COSL_OK && (e = cudaMemcpyAsync(&status, dStatus, 4, cudaMemcpyDeviceToHost, s)) != cudaSuccess) fail(e, "D2H");
  if (rc == COSL_OK && (e = cudaStreamSynchronize(s)) != cudaSuccess) fail(e, "sync");
  cudaFreeAsync(slab, s);
  if (rc != COSL_OK) return rc;
  if (status != 0)
    return set_error(COSL_E_INVALID, "cosl_posegraph_spread_chains: chain %d has no fixed node (rank-deficient system)",
                     status - 1);
  return COSL_OK;
}
