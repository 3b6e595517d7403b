// Test infrastructure (CPU): the BA shim's host logic with a MOCK of the C-ABI behind it.  The mock
// cosl_ba_solve checks that bundleAdjustRobust flattened the STL containers exactly as
// RobustBundleRTS::parseInputs orders them (CSR by point, app/SL_CoSLAMRobustBA.cpp:109-165), then
// writes recognisable values back; main() checks the write-back and the failure path.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "math/SL_Matrix.h"
#include "geometry/SL_Point.h"

struct Meas2D {  // members CoSLAM relies on
  int viewId;
  double x, y;
  int outlier;
  Meas2D(int v, double x_, double y_) : viewId(v), x(x_), y(y_), outlier(0) {}
};

#include "SL_BundleAdjust.h"

static int g_fail = 0, g_checked = 0;

extern "C" void cosl_ba_options_default(cosl_ba_options* o) {
  std::memset(o, 0, sizeof(*o));
  o->max_err = 6.0;
  o->outer_iters = 5;
  o->inner_iters = 10;
}
extern "C" const char* cosl_last_error(void) { return "mock failure"; }

#define EXPECT(c)                                              \
  do {                                                         \
    if (!(c)) {                                                \
      std::printf("mock: expectation failed: %s\n", #c);       \
      g_fail = 1;                                              \
    }                                                          \
  } while (0)

extern "C" int cosl_ba_solve(cosl_ba_problem* p, const cosl_ba_options* o, double* info);
// the shim's nGpus > 1 path: same checks, device list ignored by the mock
extern "C" int cosl_ba_solve_multi(cosl_ba_problem* p, const cosl_ba_options* o, int, const int*, double* info) {
  return cosl_ba_solve(p, o, info);
}
extern "C" int cosl_ba_solve(cosl_ba_problem* p, const cosl_ba_options* o, double* info) {
  if (o->device == 7) return COSL_E_CUDA;  // failure path
  EXPECT(p->m == 3 && p->n == 4 && p->nobs == 7 && p->m_con == 1 && p->n_con == 2);
  EXPECT(o->max_err == 4.5 && o->outer_iters == 3 && o->inner_iters == 9 && o->device == 0);
  const long long ptr[5] = {0, 2, 3, 6, 7};
  for (int i = 0; i <= 4; ++i) EXPECT(p->ptr[i] == ptr[i]);
  const int cam[7] = {0, 2, 1, 0, 1, 2, 2};
  for (int k = 0; k < 7; ++k) {
    EXPECT(p->cam[k] == cam[k]);
    EXPECT(p->xy[2 * k] == 10.0 * k + 1 && p->xy[2 * k + 1] == 10.0 * k + 2);
  }
  for (int j = 0; j < 3; ++j)
    for (int k = 0; k < 9; ++k) {
      EXPECT(p->K[9 * j + k] == 100 * j + k);
      EXPECT(p->R[9 * j + k] == 1000 * j + k);
    }
  for (int j = 0; j < 3; ++j)
    for (int k = 0; k < 3; ++k) EXPECT(p->t[3 * j + k] == -(10 * j + k));
  for (int i = 0; i < 4; ++i)
    for (int k = 0; k < 3; ++k) EXPECT(p->X[3 * i + k] == 0.5 * i + k);
  g_checked = 1;
  for (int j = 0; j < 3; ++j) {
    for (int k = 0; k < 9; ++k) p->R[9 * j + k] += 0.25;
    for (int k = 0; k < 3; ++k) p->t[3 * j + k] -= 0.5;
  }
  for (int i = 0; i < 4; ++i)
    for (int k = 0; k < 3; ++k) p->X[3 * i + k] += 7.0;
  p->outlier[1] = 1;
  p->outlier[5] = 1;
  if (info) info[0] = 1.0;
  return COSL_OK;
}

int main() {
  std::vector<Mat_d> Ks(3), Rs(3), Ts(3);
  for (int j = 0; j < 3; ++j) {
    Ks[j].resize(3, 3);
    Rs[j].resize(3, 3);
    Ts[j].resize(3, 1);
    for (int k = 0; k < 9; ++k) {
      Ks[j].data[k] = 100 * j + k;
      Rs[j].data[k] = 1000 * j + k;
    }
    for (int k = 0; k < 3; ++k) Ts[j].data[k] = -(10 * j + k);
  }
  std::vector<Point3d> pts;
  for (int i = 0; i < 4; ++i) pts.push_back(Point3d(0.5 * i, 0.5 * i + 1, 0.5 * i + 2));
  std::vector<std::vector<Meas2D> > meas(4);
  const int views[4][3] = {{0, 2, -1}, {1, -1, -1}, {0, 1, 2}, {2, -1, -1}};
  int k = 0;
  for (int i = 0; i < 4; ++i)
    for (int q = 0; q < 3 && views[i][q] >= 0; ++q, ++k)
      meas[i].push_back(Meas2D(views[i][q], 10.0 * k + 1, 10.0 * k + 2));
  bundleAdjustRobust(1, Ks, Rs, Ts, 2, pts, meas, 4.5, 3, 9);
  EXPECT(g_checked == 1);
  for (int j = 0; j < 3; ++j) {
    for (int q = 0; q < 9; ++q) EXPECT(Rs[j].data[q] == 1000 * j + q + 0.25);
    for (int q = 0; q < 3; ++q) EXPECT(Ts[j].data[q] == -(10 * j + q) - 0.5);
    for (int q = 0; q < 9; ++q) EXPECT(Ks[j].data[q] == 100 * j + q);  // intrinsics untouched
  }
  for (int i = 0; i < 4; ++i) {
    EXPECT(pts[i].x == 0.5 * i + 7.0 && pts[i].y == 0.5 * i + 8.0 && pts[i].z == 0.5 * i + 9.0);
  }
  EXPECT(meas[0][0].outlier == 0 && meas[0][1].outlier == 1 && meas[1][0].outlier == 0);
  EXPECT(meas[2][0].outlier == 0 && meas[2][1].outlier == 0 && meas[2][2].outlier == 1);
  EXPECT(meas[3][0].outlier == 0);
  bool threw = false;
  try {
    bundleAdjustRobust(1, Ks, Rs, Ts, 2, pts, meas, 4.5, 3, 9, /*device=*/7);
  } catch (const std::runtime_error& e) {
    threw = std::strstr(e.what(), "mock failure") != 0;
  }
  EXPECT(threw);
  std::printf(g_fail ? "MOCK_BA_SHIM_FAILED\n" : "MOCK_BA_SHIM_OK\n");
  return g_fail;
}
