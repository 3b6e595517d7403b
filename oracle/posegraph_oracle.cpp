/* posegraph_oracle.cpp -- CPU restatement of the post-BA pose-graph spreading (TEST INFRASTRUCTURE
 * ONLY; nothing under coslam_b200/ may call it).
 *
 * Follows GlobalPoseGraph::computeNewCameraRotations (reference slam/SL_GlobalPoseEstimation.cpp:52-218)
 * and ::computeNewCameraTranslations (:220-359) for graphs without uncertain-scale edges
 * (nConstraintEdge = 0, the only kind RobustBundleRTS::constructCameraGraphs builds,
 * app/SL_CoSLAMRobustBA.cpp:182-232): assemble the same over-determined linear systems, row for row,
 * solve them in the least-squares sense and project each rotation block onto SO(3).  Unlike the CUDA
 * path it handles ANY edge set, not only chains.
 *
 * PINNED: tests/test_posegraph.py compares it with oracle/_ref/libposegraph_ref.so = the unmodified
 * reference file compiled against the stand-ins of oracle/ref_stubs_pg/ (live when the .so is
 * present) and with tests/golden/posegraph_ref.npz generated from it.  The reference's
 * sparseSolveLin / approxRotationMat live in LibVisualSLAM (absent): least squares and the
 * Frobenius-nearest rotation are what their call sites state (:200-211).  The least-squares solve
 * here is normal equations + Cholesky with an iterative-refinement step -- a different algorithm
 * from the Householder QR of the stand-in, so agreement is a real check. */
#include <cmath>
#include <cstring>
#include <vector>

#include "oracle.h"

namespace {

struct Dense {
  int m, n;
  std::vector<double> a; /* row-major m x n */
  Dense(int m_, int n_) : m(m_), n(n_), a((size_t)m_ * n_, 0.0) {}
  double& at(int r, int c) { return a[(size_t)r * n + c]; }
};

/* x = argmin |A x - b| : A^T A x = A^T b by Cholesky, one refinement step on the residual */
bool lstsq(Dense& A, const std::vector<double>& b, std::vector<double>& x) {
  const int m = A.m, n = A.n;
  x.assign(n, 0.0);
  if (n == 0) return true;
  std::vector<double> G((size_t)n * n, 0.0);
  for (int r = 0; r < m; ++r) {
    const double* row = &A.a[(size_t)r * n];
    for (int i = 0; i < n; ++i) {
      if (row[i] == 0.0) continue;
      for (int j = 0; j <= i; ++j) G[(size_t)i * n + j] += row[i] * row[j];
    }
  }
  for (int j = 0; j < n; ++j) { /* in-place lower Cholesky */
    double d = G[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) d -= G[(size_t)j * n + k] * G[(size_t)j * n + k];
    if (!(d > 0)) return false;
    d = std::sqrt(d);
    G[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = G[(size_t)i * n + j];
      for (int k = 0; k < j; ++k) s -= G[(size_t)i * n + k] * G[(size_t)j * n + k];
      G[(size_t)i * n + j] = s / d;
    }
  }
  std::vector<double> res(b), g(n), y(n);
  for (int pass = 0; pass < 2; ++pass) {
    for (int i = 0; i < n; ++i) g[i] = 0;
    for (int r = 0; r < m; ++r)
      for (int i = 0; i < n; ++i) g[i] += A.a[(size_t)r * n + i] * res[r];
    for (int i = 0; i < n; ++i) {
      double s = g[i];
      for (int k = 0; k < i; ++k) s -= G[(size_t)i * n + k] * y[k];
      y[i] = s / G[(size_t)i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = y[i];
      for (int k = i + 1; k < n; ++k) s -= G[(size_t)k * n + i] * y[k];
      y[i] = s / G[(size_t)i * n + i];
    }
    for (int i = 0; i < n; ++i) x[i] += y[i];
    for (int r = 0; r < m; ++r) {
      double s = b[r];
      for (int i = 0; i < n; ++i) s -= A.a[(size_t)r * n + i] * x[i];
      res[r] = s;
    }
  }
  return true;
}

/* Frobenius-nearest rotation through the polar factor: Newton iteration X <- (X + X^-T)/2 gives the
 * orthogonal polar factor Q of M; for det(M) < 0 the nearest ROTATION flips the direction of the
 * smallest singular value: R = Q (I - 2 v v^T), v = eigenvector of M^T M with the smallest eigenvalue. */
double det3(const double* A) {
  return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
}
void nearest_rotation(const double* M, double* R) {
  double X[9];
  std::memcpy(X, M, sizeof(X));
  for (int it = 0; it < 100; ++it) {
    const double d = det3(X);
    double iT[9] = {(X[4] * X[8] - X[5] * X[7]) / d, (X[5] * X[6] - X[3] * X[8]) / d, (X[3] * X[7] - X[4] * X[6]) / d,
                    (X[2] * X[7] - X[1] * X[8]) / d, (X[0] * X[8] - X[2] * X[6]) / d, (X[1] * X[6] - X[0] * X[7]) / d,
                    (X[1] * X[5] - X[2] * X[4]) / d, (X[2] * X[3] - X[0] * X[5]) / d, (X[0] * X[4] - X[1] * X[3]) / d};
    double diff = 0;
    for (int i = 0; i < 9; ++i) {
      const double nx = 0.5 * (X[i] + iT[i]);
      diff = std::fmax(diff, std::fabs(nx - X[i]));
      X[i] = nx;
    }
    if (diff < 1e-16) break;
  }
  if (det3(X) < 0) {
    /* smallest eigenvector of S = M^T M by inverse power iteration on S - (lambda_min guess) */
    double S[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) S[3 * i + j] = M[i] * M[j] + M[3 + i] * M[3 + j] + M[6 + i] * M[6 + j];
    /* Jacobi eigen-decomposition of the symmetric 3x3 */
    double V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int sweep = 0; sweep < 60; ++sweep) {
      double off = std::fabs(S[1]) + std::fabs(S[2]) + std::fabs(S[5]);
      if (off < 1e-300) break;
      for (int p = 0; p < 2; ++p)
        for (int q = p + 1; q < 3; ++q) {
          const double apq = S[3 * p + q];
          if (std::fabs(apq) < 1e-300) continue;
          const double th = (S[3 * q + q] - S[3 * p + p]) / (2 * apq);
          const double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1));
          const double c = 1 / std::sqrt(t * t + 1), s = t * c;
          for (int k = 0; k < 3; ++k) { /* S <- S J */
            const double x = S[3 * k + p], y = S[3 * k + q];
            S[3 * k + p] = c * x - s * y; S[3 * k + q] = s * x + c * y;
          }
          for (int k = 0; k < 3; ++k) { /* S <- J^T S */
            const double x = S[3 * p + k], y = S[3 * q + k];
            S[3 * p + k] = c * x - s * y; S[3 * q + k] = s * x + c * y;
          }
          for (int k = 0; k < 3; ++k) {
            const double x = V[3 * k + p], y = V[3 * k + q];
            V[3 * k + p] = c * x - s * y; V[3 * k + q] = s * x + c * y;
          }
        }
    }
    int sm = 0;
    if (S[4] < S[3 * sm + sm]) sm = 1;
    if (S[8] < S[3 * sm + sm]) sm = 2;
    const double v[3] = {V[sm], V[3 + sm], V[6 + sm]};
    double Xv[3];
    for (int i = 0; i < 3; ++i) Xv[i] = X[3 * i] * v[0] + X[3 * i + 1] * v[1] + X[3 * i + 2] * v[2];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) X[3 * i + j] -= 2 * Xv[i] * v[j];
  }
  std::memcpy(R, X, sizeof(X));
}

}  // namespace

/* Returns 0, or 1 when a least-squares system is rank deficient (free nodes not tied to a fixed one). */
extern "C" int orc_posegraph_spread(int nNodes, const int* fixed, const double* R, const double* t, int nEdges,
                                    const int* id1, const int* id2, const double* eR, const double* et,
                                    double* newR, double* newt) {
  /* row / column numbering of :53-69 and :221-237 */
  std::vector<int> rowOf(nEdges, -1), colOf(nNodes, -1);
  int nRows = 0, nCols = 0;
  for (int e = 0; e < nEdges; ++e)
    if (!fixed[id1[e]] || !fixed[id2[e]]) rowOf[e] = nRows++;
  for (int k = 0; k < nNodes; ++k)
    if (!fixed[k]) colOf[k] = nCols++;

  /* ---- rotations (:70-200): unknown block of node k = R_k^T in row-major, i.e. the three columns of R_k;
   * edge (i -> j):  col_a(R_j) - R_ij col_a(R_i) = 0  for a = 0..2 ---- */
  {
    Dense A(9 * nRows, 9 * nCols);
    std::vector<double> b(9 * (size_t)nRows, 0.0), x;
    for (int e = 0; e < nEdges; ++e) {
      const int r = rowOf[e];
      if (r < 0) continue;
      const int i = id1[e], j = id2[e];
      const double* Rij = eR + 9 * (size_t)e;
      for (int a = 0; a < 3; ++a)
        for (int u = 0; u < 3; ++u) {
          const int row = 9 * r + 3 * a + u;
          double rhs = 0.0;
          if (!fixed[j]) A.at(row, 9 * colOf[j] + 3 * a + u) = 1.0;
          else rhs -= R[9 * (size_t)j + 3 * u + a]; /* b = -R_j^T (:187-198) */
          if (!fixed[i]) {
            for (int v = 0; v < 3; ++v) A.at(row, 9 * colOf[i] + 3 * a + v) = -Rij[3 * u + v];
          } else { /* b = (R_ij R_i)^T (:146-148) */
            for (int v = 0; v < 3; ++v) rhs += Rij[3 * u + v] * R[9 * (size_t)i + 3 * v + a];
          }
          b[row] = rhs;
        }
    }
    if (!lstsq(A, b, x)) return 1;
    for (int k = 0; k < nNodes; ++k) {
      if (colOf[k] < 0) {
        std::memcpy(newR + 9 * (size_t)k, R + 9 * (size_t)k, 72);
      } else {
        double M[9];
        for (int a = 0; a < 3; ++a)
          for (int u = 0; u < 3; ++u) M[3 * u + a] = x[9 * (size_t)colOf[k] + 3 * a + u]; /* transpose (:209) */
        nearest_rotation(M, newR + 9 * (size_t)k);
      }
    }
  }
  /* ---- translations (:238-335):  t_j - R_ij t_i = t_ij ---- */
  {
    Dense A(3 * nRows, 3 * nCols);
    std::vector<double> b(3 * (size_t)nRows, 0.0), x;
    for (int e = 0; e < nEdges; ++e) {
      const int r = rowOf[e];
      if (r < 0) continue;
      const int i = id1[e], j = id2[e];
      const double* Rij = eR + 9 * (size_t)e;
      const double* tij = et + 3 * (size_t)e;
      for (int u = 0; u < 3; ++u) {
        const int row = 3 * r + u;
        double rhs = tij[u];
        if (!fixed[j]) A.at(row, 3 * colOf[j] + u) = 1.0;
        else rhs -= t[3 * (size_t)j + u];
        if (!fixed[i]) {
          for (int v = 0; v < 3; ++v) A.at(row, 3 * colOf[i] + v) = -Rij[3 * u + v];
        } else {
          for (int v = 0; v < 3; ++v) rhs += Rij[3 * u + v] * t[3 * (size_t)i + v];
        }
        b[row] = rhs;
      }
    }
    if (!lstsq(A, b, x)) return 1;
    for (int k = 0; k < nNodes; ++k)
      std::memcpy(newt + 3 * (size_t)k, colOf[k] < 0 ? t + 3 * (size_t)k : &x[3 * (size_t)colOf[k]], 24);
  }
  return 0;
}
