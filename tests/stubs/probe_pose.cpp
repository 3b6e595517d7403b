// Test infrastructure: compiled once against the reference's slam/SL_IntraCamPose.h and once against
// coslam_b200/shim/SL_IntraCamPose.h (tests/test_shim_reference_callers.py).  Both must accept the
// same function-pointer type and the same option members: the shim is a drop-in for the header.
#include "slam/SL_IntraCamPose.h"

typedef bool (*intraCamEstimate_t)(const double*, const double*, const double*, int, const double*,
                                   const double*, const double*, const double, double*, double*,
                                   IntraCamPoseOption*);
intraCamEstimate_t probe = &intraCamEstimate;

double use_every_member() {
  IntraCamPoseOption o;  // default constructible, as SingleSLAM::poseUpdate3D uses it
  o.maxIterLM = 100;
  o.maxIterRW = 5;
  o.epsErrorChangeLM = 1e-7;
  o.epsParamChangeLM = 1e-6;
  o.epsErrorChangeRW = 1e-6;
  o.verboseLM = 0;
  o.verboseRW = 0;
  o.lambda0 = 1e-3;
  return o.lambda + o.err0 + o.err + o.errRW + o.retTypeLM + o.npts + o.nIterLM + o.nIterRW;
}
