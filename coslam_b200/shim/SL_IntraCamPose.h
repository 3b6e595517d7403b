// Drop-in for slam/SL_IntraCamPose.h of danping/CoSLAM (the declarations CoSLAM's live pipeline
// uses): IntraCamPoseOption (:19-57) and intraCamEstimate (:92-95) with the reference's exact
// signature, forwarding to cosl_pose_intracam.  app/SL_SingleSLAM.cpp:664 compiles unchanged.
#ifndef SL_INTRACAMPOSE_H_
#define SL_INTRACAMPOSE_H_
#include <cmath>
#include <cstdio>

#include "coslam_b200.h"

class IntraCamPoseOption {
public:
	int maxIterLM;
	int maxIterRW;
	double epsErrorChangeLM;
	double epsParamChangeLM;
	double epsErrorChangeRW;
	int verboseLM;
	int verboseRW;
public:
	double lambda0;
	double lambda;
	double err0;
	double err;
	double errRW;
	int retTypeLM;
	int npts;
	int nIterLM;
	int nIterRW;
public:
	IntraCamPoseOption() :
			maxIterLM(100), maxIterRW(5), epsErrorChangeLM(1e-7), epsParamChangeLM(1e-6), epsErrorChangeRW(1e-6),
			verboseLM(0), verboseRW(0), lambda0(1e-3) {
	}
	void printLM() {
		printf("lambda:%lf -> %lf\n", lambda0, lambda);
		printf("ssd: %lf -> %lf\n", err0, err);
		printf("err: %lf -> %lf\n", sqrt(err0 / npts), sqrt(err / npts));
		printf("npts:%d\nreturn type:%d\nnumber of LM interation:%d\n", npts, retTypeLM, nIterLM);
	}
	void printRW() {
	}
};

inline bool intraCamEstimate(const double K[9], const double R0[9], const double t0[3], int npts,
		const double prevErrs[], const double Ms[], const double ms[], const double tau, double R_opt[9],
		double t_opt[3], IntraCamPoseOption* opt) {
	cosl_pose_opt o;
	cosl_pose_opt_default(&o);
	o.maxIterLM = opt->maxIterLM;
	o.maxIterRW = opt->maxIterRW;
	o.epsErrorChangeLM = opt->epsErrorChangeLM;
	o.epsParamChangeLM = opt->epsParamChangeLM;
	o.epsErrorChangeRW = opt->epsErrorChangeRW;
	o.lambda0 = opt->lambda0;
	int ok = 0;
	const int rc = cosl_pose_intracam(K, R0, t0, npts, prevErrs, Ms, ms, tau, R_opt, t_opt, &o, &ok);
	if (rc != COSL_OK) {
		std::fprintf(stderr, "intraCamEstimate: %s\n", cosl_last_error());
		return false;
	}
	opt->lambda0 = o.lambda0;
	opt->lambda = o.lambda;
	opt->err0 = o.err0;
	opt->err = o.err;
	opt->errRW = o.errRW;
	opt->retTypeLM = o.retTypeLM;
	opt->npts = o.npts;
	opt->nIterLM = o.nIterLM;
	opt->nIterRW = o.nIterRW;
	return ok != 0;
}
#endif /* SL_INTRACAMPOSE_H_ */
