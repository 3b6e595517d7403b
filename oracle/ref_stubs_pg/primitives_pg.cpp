/* TEST INFRASTRUCTURE.  Definitions behind the stand-in headers of oracle/ref_stubs_pg/ for the
 * oracle/_ref build of the UNMODIFIED reference file slam/SL_GlobalPoseEstimation.cpp
 * (`make -C oracle ref`), plus the C entry tests bind with ctypes.
 *
 * The five primitives on the post-BA path (computeNewCameraRotations :52-218,
 * computeNewCameraTranslations :220-359) live in LibVisualSLAM, which is NOT in /root/reference
 * (un-vendored dependency).  Their meaning is INFERRED from the call sites and stated here:
 *   mat33AB(A,B,C)            C = A B                  (:147, row-major 3x3)
 *   mat33Trans(A,At)          At = A^T                 (:148,:187,:209)
 *   mat33ProdVec(R,x,y,o,a,b) o = a R x + b y          (:333, comment "T_j = T_{ij} + R_{ij} T_i")
 *   sparseSolveLin(T,b,x)     x = argmin |T x - b|_2   (:200,:335; m >= n, full column rank)
 *   approxRotationMat(M,R)    R = nearest rotation to M in Frobenius norm = U diag(1,1,det(UV^T)) V^T
 *                             (:210 "get the approximated rotations")
 * sparseSolveLin is a dense Householder QR here (numerically independent of the closed form the
 * product uses); the off-path variants (constraint systems, :361-1281) only have to link. */
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "SL_GlobalPoseEstimation.h"
#include "geometry/SL_RigidTransform.h"
#include "math/SL_LinAlgWarper.h"
#include "math/SL_SparseLinearSystem.h"

void mat33AB(const double* A, const double* B, double* C) {
  double o[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) o[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  std::memcpy(C, o, sizeof(o));
}
void mat33Trans(const double* A, double* At) {
  double o[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) o[3 * j + i] = A[3 * i + j];
  std::memcpy(At, o, sizeof(o));
}
void mat33ProdVec(const double* R, const double* x, const double* y, double* out, double a, double b) {
  double o[3];
  for (int i = 0; i < 3; ++i) o[i] = a * (R[3 * i] * x[0] + R[3 * i + 1] * x[1] + R[3 * i + 2] * x[2]) + b * y[i];
  std::memcpy(out, o, sizeof(o));
}

/* dense Householder QR least squares */
void sparseSolveLin(const Triplets& T, const double* b, double* x) {
  const int m = T.m, n = T.n;
  if (n == 0) return;
  std::vector<double> A((size_t)m * n, 0.0), r(b, b + m);
  for (size_t k = 0; k < T.v.size(); ++k) A[(size_t)T.ri[k] * n + T.ci[k]] += T.v[k];
  for (int j = 0; j < n; ++j) {
    double s = 0;
    for (int i = j; i < m; ++i) s += A[(size_t)i * n + j] * A[(size_t)i * n + j];
    double nrm = std::sqrt(s);
    if (nrm == 0) { std::fprintf(stderr, "ref sparseSolveLin: rank deficient\n"); std::abort(); }
    double a0 = A[(size_t)j * n + j], alpha = a0 > 0 ? -nrm : nrm;
    std::vector<double> v(m - j);
    for (int i = j; i < m; ++i) v[i - j] = A[(size_t)i * n + j];
    v[0] -= alpha;
    double vv = 0;
    for (double e : v) vv += e * e;
    if (vv > 0) {
      for (int c = j; c < n; ++c) {
        double d = 0;
        for (int i = j; i < m; ++i) d += v[i - j] * A[(size_t)i * n + c];
        if (d == 0) continue;
        d = 2 * d / vv;
        for (int i = j; i < m; ++i) A[(size_t)i * n + c] -= d * v[i - j];
      }
      double d = 0;
      for (int i = j; i < m; ++i) d += v[i - j] * r[i];
      d = 2 * d / vv;
      for (int i = j; i < m; ++i) r[i] -= d * v[i - j];
    }
  }
  for (int j = n - 1; j >= 0; --j) {
    double s = r[j];
    for (int c = j + 1; c < n; ++c) s -= A[(size_t)j * n + c] * x[c];
    x[j] = s / A[(size_t)j * n + j];
  }
}

/* one-sided Jacobi SVD of a 3x3: M = U S V^T */
static void svd33(const double* M, double* U, double* S, double* V) {
  double A[9], W[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  std::memcpy(A, M, sizeof(A));
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double a = 0, b = 0, c = 0;
        for (int i = 0; i < 3; ++i) { a += A[3 * i + p] * A[3 * i + p]; b += A[3 * i + q] * A[3 * i + q]; c += A[3 * i + p] * A[3 * i + q]; }
        off = std::fmax(off, std::fabs(c) / std::sqrt(a * b + 1e-300));
        if (std::fabs(c) < 1e-300) continue;
        double zeta = (b - a) / (2 * c), t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1 + zeta * zeta));
        double cs = 1 / std::sqrt(1 + t * t), sn = cs * t;
        for (int i = 0; i < 3; ++i) {
          double x = A[3 * i + p], y = A[3 * i + q];
          A[3 * i + p] = cs * x - sn * y; A[3 * i + q] = sn * x + cs * y;
          x = W[3 * i + p]; y = W[3 * i + q];
          W[3 * i + p] = cs * x - sn * y; W[3 * i + q] = sn * x + cs * y;
        }
      }
    if (off < 1e-17) break;
  }
  for (int j = 0; j < 3; ++j) {
    double s = 0;
    for (int i = 0; i < 3; ++i) s += A[3 * i + j] * A[3 * i + j];
    S[j] = std::sqrt(s);
    for (int i = 0; i < 3; ++i) U[3 * i + j] = S[j] > 0 ? A[3 * i + j] / S[j] : 0;
  }
  std::memcpy(V, W, sizeof(W));
}
static double det33(const double* A) {
  return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
}
void approxRotationMat(const double* M, double* R) {
  double U[9], S[3], V[9];
  svd33(M, U, S, V);
  int smin = 0;
  for (int j = 1; j < 3; ++j) if (S[j] < S[smin]) smin = j;
  double UVt[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) UVt[3 * i + j] = U[3 * i] * V[3 * j] + U[3 * i + 1] * V[3 * j + 1] + U[3 * i + 2] * V[3 * j + 2];
  if (det33(UVt) < 0) { /* flip the direction of the smallest singular value */
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) UVt[3 * i + j] -= 2 * U[3 * i + smin] * V[3 * j + smin];
  }
  std::memcpy(R, UVt, sizeof(UVt));
}

static void off_path(const char* w) { std::fprintf(stderr, "oracle/_ref: %s is off the post-BA path\n", w); std::abort(); }
void matTrans(const Mat_d&, Mat_d&) { off_path("matTrans"); }
void matAx(int, int, const double*, const double*, double*) { off_path("matAx"); }
void matQR(const Mat_d&, Mat_d&, Mat_d&) { off_path("matQR"); }
void triplets2Sparse(const Triplets&, SparseMat&) { off_path("triplets2Sparse"); }
void tripletsSplitCol(const Triplets&, int, Triplets&, Triplets&) { off_path("tripletsSplitCol"); }
void sparseSplitCol(const SparseMat&, int, SparseMat&, bool) { off_path("sparseSplitCol"); }
void sparseMatMul(const SparseMat&, const SparseMat&, SparseMat&) { off_path("sparseMatMul"); }
void dense2Sparse(const Mat_d&, SparseMat&) { off_path("dense2Sparse"); }
void sparseSolveLin(const SparseMat&, const SparseMat&, const double*, double*, double*) { off_path("sparseSolveLin(constrained)"); }

/* C entry: builds the graph the way RobustBundleRTS::constructCameraGraphs does
 * (app/SL_CoSLAMRobustBA.cpp:182-232) and runs updateNonKeyCameraPoses' two solves (:233-250). */
extern "C" void ref_posegraph_spread(int nNodes, const int* fixed, const double* R, const double* t, int nEdges,
                                     const int* id1, const int* id2, const double* eR, const double* et,
                                     double* newR, double* newt) {
  GlobalPoseGraph g;
  g.reserve(nNodes, nEdges > 0 ? nEdges : 1);
  for (int k = 0; k < nNodes; ++k) {
    CamPoseNode* nd = g.newNode();
    nd->set(k, 0, R + 9 * k, t + 3 * k);
    nd->fixed = fixed[k] != 0;
  }
  for (int e = 0; e < nEdges; ++e) g.addEdge()->set(id1[e], id2[e], eR + 9 * e, et + 3 * e);
  g.computeNewCameraRotations();
  g.computeNewCameraTranslations();
  for (int k = 0; k < nNodes; ++k) {
    std::memcpy(newR + 9 * k, g.poseNodes[k].newR, 72);
    std::memcpy(newt + 3 * k, g.poseNodes[k].newt, 24);
  }
}
