/* primitives.cpp -- definitions of the LibVisualSLAM primitives that the UNMODIFIED reference file
 * /root/reference/src/slam/SL_IntraCamPose.cpp calls on the intraCamEstimate path, plus an
 * extern "C" entry so that tests can call the reference's own intraCamEstimate.
 * TEST INFRASTRUCTURE: this file + the reference source are compiled into
 * oracle/_ref/libintracam_ref.so by oracle/Makefile; nothing under coslam_b200/ links it.
 *
 * Semantics (LibVisualSLAM is not in the container; inferred from the call sites):
 *   matATB(ma,na,mb,nb,A,B,C)  C (na x nb) = A^T B, A is ma x na, B is mb x nb, row-major
 *                              (SL_IntraCamPose.cpp:239 builds the 6x6 J^T J from a 2x6 J)
 *   matAB(ma,na,mb,nb,A,B,C)   C (ma x nb) = A B                      (:257, param = inv(A) b)
 *   matInv(n,A,invA)           general inverse (LAPACK dgetrf/dgetri there; Gauss-Jordan with
 *                              partial pivoting here: same result up to rounding)       (:256)
 *   mat33AB(A,B,C)             3x3 product                                              (:59)
 *   project(K,R,t,M,m)         m = pi(K (R M + t))                                      (:60)
 *   reprojError2               sum |m_i - project(M_i)|^2                               (:400)
 *   doubleArrCopy(d,i,s,n)     copies n doubles from s to d + i                         (:406)
 * Everything else the file references (covariance / epipolar variants) aborts if reached. */
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>

#include "math/SL_LinAlg.h"
#include "geometry/SL_Geometry.h"
#include "geometry/SL_5point.h"
#include "geometry/SL_FundamentalMatrix.h"
#include "SL_IntraCamPose.h"

void matATB(int ma, int na, int mb, int nb, const double* A, const double* B, double* C) {
  (void)mb;
  for (int i = 0; i < na; ++i)
    for (int j = 0; j < nb; ++j) {
      double s = 0;
      for (int k = 0; k < ma; ++k) s += A[k * na + i] * B[k * nb + j];
      C[i * nb + j] = s;
    }
}
void matAB(int ma, int na, int mb, int nb, const double* A, const double* B, double* C) {
  (void)mb;
  for (int i = 0; i < ma; ++i)
    for (int j = 0; j < nb; ++j) {
      double s = 0;
      for (int k = 0; k < na; ++k) s += A[i * na + k] * B[k * nb + j];
      C[i * nb + j] = s;
    }
}
bool matInv(int n, const double* A, double* invA) {
  double* a = new double[(size_t)n * 2 * n];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      a[i * 2 * n + j] = A[i * n + j];
      a[i * 2 * n + n + j] = (i == j) ? 1.0 : 0.0;
    }
  bool ok = true;
  for (int c = 0; c < n && ok; ++c) {
    int p = c;
    double best = std::fabs(a[c * 2 * n + c]);
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(a[r * 2 * n + c]) > best) {
        best = std::fabs(a[r * 2 * n + c]);
        p = r;
      }
    if (best == 0.0) {
      ok = false;
      break;
    }
    if (p != c)
      for (int j = 0; j < 2 * n; ++j) std::swap(a[c * 2 * n + j], a[p * 2 * n + j]);
    const double d = 1.0 / a[c * 2 * n + c];
    for (int j = 0; j < 2 * n; ++j) a[c * 2 * n + j] *= d;
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      const double f = a[r * 2 * n + c];
      for (int j = 0; j < 2 * n; ++j) a[r * 2 * n + j] -= f * a[c * 2 * n + j];
    }
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) invA[i * n + j] = ok ? a[i * 2 * n + n + j] : 0.0;
  delete[] a;
  return ok;
}
void mat33AB(const double* A, const double* B, double* C) {
  double T[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      T[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  std::memcpy(C, T, sizeof(T));
}
void doubleArrCopy(double* dst, int dstStart, const double* src, int len) {
  std::memcpy(dst + dstStart, src, sizeof(double) * len);
}
void project(const double* K, const double* R, const double* t, const double* M, double* m) {
  const double c0 = R[0] * M[0] + R[1] * M[1] + R[2] * M[2] + t[0];
  const double c1 = R[3] * M[0] + R[4] * M[1] + R[5] * M[2] + t[1];
  const double c2 = R[6] * M[0] + R[7] * M[1] + R[8] * M[2] + t[2];
  const double u = K[0] * c0 + K[1] * c1 + K[2] * c2;
  const double v = K[3] * c0 + K[4] * c1 + K[5] * c2;
  const double w = K[6] * c0 + K[7] * c1 + K[8] * c2;
  m[0] = u / w;
  m[1] = v / w;
}
double reprojError2(const double* K, const double* R, const double* t, int npts, const double* Ms,
                    const double* ms) {
  double e = 0;
  for (int i = 0; i < npts; ++i) {
    double rm[2];
    project(K, R, t, Ms + 3 * i, rm);
    const double dx = rm[0] - ms[2 * i], dy = rm[1] - ms[2 * i + 1];
    e += dx * dx + dy * dy;
  }
  return e;
}
static void off_path(const char* what) {
  std::fprintf(stderr, "ref_stubs: %s is not on the intraCamEstimate path\n", what);
  std::abort();
}
void mat22Inv(const double*, double*) { off_path("mat22Inv"); }
void mat33Inv(const double*, double*) { off_path("mat33Inv"); }
void getProjectionCovMat(const double*, const double*, const double*, const double*, const double*,
                         double*, double) { off_path("getProjectionCovMat"); }
double mahaDist2(const double*, const double*, const double*) { off_path("mahaDist2"); return 0; }
void formEMat(const double*, const double*, const double*, const double*, double*) { off_path("formEMat"); }
void getFMat(const double*, const double*, const double*, double*) { off_path("getFMat"); }
double epipolarError(const double*, const double*, const double*) { off_path("epipolarError"); return 0; }
void computeEpipolarLine(const double*, double, double, double*) { off_path("computeEpipolarLine"); }

/* C entry: the reference's own intraCamEstimate with the option block flattened to
 * {maxIterLM, maxIterRW, epsErrorChangeLM, epsParamChangeLM, epsErrorChangeRW, lambda0} in and
 * {lambda, lambda0, err0, err, errRW, retTypeLM, nIterLM, nIterRW} out. */
extern "C" int ref_intraCamEstimate(const double* K, const double* R0, const double* t0, int npts,
                                    const double* prevErrs, const double* Ms, const double* ms,
                                    double tau, double* R_opt, double* t_opt, const double* optIn,
                                    double* optOut) {
  IntraCamPoseOption opt;
  if (optIn) {
    opt.maxIterLM = (int)optIn[0];
    opt.maxIterRW = (int)optIn[1];
    opt.epsErrorChangeLM = optIn[2];
    opt.epsParamChangeLM = optIn[3];
    opt.epsErrorChangeRW = optIn[4];
    opt.lambda0 = optIn[5];
  }
  const bool ok = intraCamEstimate(K, R0, t0, npts, prevErrs, Ms, ms, tau, R_opt, t_opt, &opt);
  if (optOut) {
    optOut[0] = opt.lambda;
    optOut[1] = opt.lambda0;
    optOut[2] = opt.err0;
    optOut[3] = opt.err;
    optOut[4] = opt.errRW;
    optOut[5] = opt.retTypeLM;
    optOut[6] = opt.nIterLM;
    optOut[7] = opt.nIterRW;
  }
  return ok ? 1 : 0;
}
extern "C" void ref_getSO3ExpMap(const double* w, double* R) { getSO3ExpMap(w, R); }
