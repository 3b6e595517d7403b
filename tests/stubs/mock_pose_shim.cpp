// Test infrastructure (CPU): intraCamEstimate of coslam_b200/shim/SL_IntraCamPose.h against a MOCK
// of cosl_pose_intracam: options in, diagnostics out, return value, failure path.
#include <cstdio>
#include <cstring>

#include "SL_IntraCamPose.h"

static int g_fail = 0, g_mode = 0;
#define EXPECT(c)                                        \
  do {                                                   \
    if (!(c)) {                                          \
      std::printf("mock: expectation failed: %s\n", #c); \
      g_fail = 1;                                        \
    }                                                    \
  } while (0)

static const double Kc[9] = {1, 0, 2, 0, 3, 4, 0, 0, 1}, R0c[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1},
                    t0c[3] = {5, 6, 7}, Msc[6] = {1, 2, 3, 4, 5, 6}, msc[4] = {7, 8, 9, 10},
                    prevc[2] = {0.5, 1.5};

extern "C" {
void cosl_pose_opt_default(cosl_pose_opt* o) {
  std::memset(o, 0, sizeof(*o));
  o->maxIterLM = 100;
  o->maxIterRW = 5;
  o->lambda0 = 1e-3;
}
const char* cosl_last_error(void) { return "mock"; }
int cosl_pose_intracam(const double K[9], const double R0[9], const double t0[3], int npts,
                       const double* prevErrs, const double* Ms, const double* ms, double tau,
                       double R_opt[9], double t_opt[3], cosl_pose_opt* o, int* ok) {
  if (g_mode == 2) return COSL_E_CUDA;
  EXPECT(K == Kc && R0 == R0c && t0 == t0c && Ms == Msc && ms == msc && npts == 2 && tau == 10.0);
  EXPECT(prevErrs == (g_mode == 0 ? (const double*)0 : prevc));
  EXPECT(o->maxIterLM == 77 && o->maxIterRW == 3 && o->epsErrorChangeLM == 1e-5);
  EXPECT(o->epsParamChangeLM == 1e-4 && o->epsErrorChangeRW == 1e-3 && o->lambda0 == 0.5);
  for (int k = 0; k < 9; ++k) R_opt[k] = 10 + k;
  for (int k = 0; k < 3; ++k) t_opt[k] = 20 + k;
  o->lambda0 = 0.25;
  o->lambda = 0.125;
  o->err0 = 3;
  o->err = 2;
  o->errRW = 1;
  o->retTypeLM = 0;
  o->npts = 2;
  o->nIterLM = 9;
  o->nIterRW = 4;
  *ok = (g_mode == 0) ? 1 : 0;
  return COSL_OK;
}
}

int main() {
  double R[9], t[3];
  IntraCamPoseOption opt;
  EXPECT(opt.maxIterLM == 100 && opt.maxIterRW == 5 && opt.lambda0 == 1e-3);  // reference defaults
  opt.maxIterLM = 77;
  opt.maxIterRW = 3;
  opt.epsErrorChangeLM = 1e-5;
  opt.epsParamChangeLM = 1e-4;
  opt.epsErrorChangeRW = 1e-3;
  opt.lambda0 = 0.5;
  g_mode = 0;
  EXPECT(intraCamEstimate(Kc, R0c, t0c, 2, 0, Msc, msc, 10.0, R, t, &opt) == true);
  EXPECT(R[0] == 10 && R[8] == 18 && t[0] == 20 && t[2] == 22);
  EXPECT(opt.lambda0 == 0.25 && opt.lambda == 0.125 && opt.err0 == 3 && opt.err == 2 && opt.errRW == 1);
  EXPECT(opt.retTypeLM == 0 && opt.npts == 2 && opt.nIterLM == 9 && opt.nIterRW == 4);
  opt.lambda0 = 0.5;
  g_mode = 1;  // LM failure reported through *ok
  EXPECT(intraCamEstimate(Kc, R0c, t0c, 2, prevc, Msc, msc, 10.0, R, t, &opt) == false);
  g_mode = 2;  // C-ABI error: false, options untouched
  opt.nIterRW = -5;
  EXPECT(intraCamEstimate(Kc, R0c, t0c, 2, prevc, Msc, msc, 10.0, R, t, &opt) == false);
  EXPECT(opt.nIterRW == -5);
  std::printf(g_fail ? "MOCK_POSE_SHIM_FAILED\n" : "MOCK_POSE_SHIM_OK\n");
  return g_fail;
}
