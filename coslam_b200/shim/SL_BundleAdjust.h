// Drop-in for LibVisualSLAM's geometry/SL_BundleAdjust.h as CoSLAM uses it:
//   void bundleAdjustRobust(int nCamsCon, vector<Mat_d>& Ks, vector<Mat_d>& Rs, vector<Mat_d>& Ts,
//                           int nPtsCon, vector<Point3d>& pts, vector<vector<Meas2D> >& meas,
//                           double maxErr, int maxIter, int nInnerMaxIter)
// (argument order from app/SL_CoSLAMRobustBA.cpp:174-175, app/SL_InterCamPoseEstimator.cpp:95,
// app/SL_MergeCameraGroup.cpp:646-647).  The LibVisualSLAM types are not redefined here: the
// function is a template over them and only touches the members CoSLAM itself relies on
// (SURVEY.md Appendix D): Mat_d::data, Point3d::{x,y,z} (aliased M[3]), Meas2D::{viewId,x,y,outlier}.
// Include AFTER the LibVisualSLAM headers that define those types.  Throws std::runtime_error on
// failure, like the original (callers wrap the call in try/catch, SL_CoSLAMRobustBA.cpp:173-179).
#ifndef SL_BUNDLEADJUST_H_
#define SL_BUNDLEADJUST_H_
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "coslam_b200.h"

template<class MatT, class Pt3T, class MeasT>
void bundleAdjustRobust(int nCamsCon, std::vector<MatT>& Ks, std::vector<MatT>& Rs, std::vector<MatT>& Ts,
		int nPtsCon, std::vector<Pt3T>& pts, std::vector<std::vector<MeasT> >& meas, double maxErr, int maxIter,
		int nInnerMaxIter, int device = 0, int nGpus = 1) {
	const int m = (int) Ks.size(), n = (int) pts.size();
	std::vector<double> K(9 * (size_t) m), R(9 * (size_t) m), t(3 * (size_t) m), X(3 * (size_t) n);
	for (int j = 0; j < m; ++j) {
		for (int k = 0; k < 9; ++k) {
			K[9 * j + k] = Ks[j].data[k];
			R[9 * j + k] = Rs[j].data[k];
		}
		for (int k = 0; k < 3; ++k)
			t[3 * j + k] = Ts[j].data[k];
	}
	std::vector<int64_t> ptr(n + 1, 0);
	for (int i = 0; i < n; ++i) {
		X[3 * i] = pts[i].x;
		X[3 * i + 1] = pts[i].y;
		X[3 * i + 2] = pts[i].z;
		ptr[i + 1] = ptr[i] + (int64_t) meas[i].size();
	}
	const int64_t N = ptr[n];
	std::vector<int32_t> cam((size_t) N);
	std::vector<double> xy(2 * (size_t) N);
	std::vector<uint8_t> outl((size_t) N, 0);
	for (int i = 0; i < n; ++i)
		for (size_t k = 0; k < meas[i].size(); ++k) {
			const int64_t o = ptr[i] + (int64_t) k;
			cam[o] = meas[i][k].viewId;
			xy[2 * o] = meas[i][k].x;
			xy[2 * o + 1] = meas[i][k].y;
		}
	cosl_ba_problem p;
	p.m = m;
	p.n = n;
	p.nobs = N;
	p.m_con = nCamsCon;
	p.n_con = nPtsCon;
	p.K = K.data();
	p.R = R.data();
	p.t = t.data();
	p.X = X.data();
	p.ptr = ptr.data();
	p.cam = cam.data();
	p.xy = xy.data();
	p.outlier = outl.data();
	cosl_ba_options o;
	cosl_ba_options_default(&o);
	o.max_err = maxErr;
	o.outer_iters = maxIter;
	o.inner_iters = nInnerMaxIter;
	o.device = device;
	double info[COSL_BA_INFOSZ];
	// nGpus > 1: devices 0 .. nGpus-1 of this process, map points sharded over them (global BA)
	const int rc = (nGpus > 1) ? cosl_ba_solve_multi(&p, &o, nGpus, 0, info) : cosl_ba_solve(&p, &o, info);
	if (rc != COSL_OK)
		throw std::runtime_error(std::string("bundleAdjustRobust: ") + cosl_last_error());
	for (int j = 0; j < m; ++j) {
		for (int k = 0; k < 9; ++k)
			Rs[j].data[k] = R[9 * j + k];
		for (int k = 0; k < 3; ++k)
			Ts[j].data[k] = t[3 * j + k];
	}
	for (int i = 0; i < n; ++i) {
		pts[i].x = X[3 * i];
		pts[i].y = X[3 * i + 1];
		pts[i].z = X[3 * i + 2];
		for (size_t k = 0; k < meas[i].size(); ++k)
			meas[i][k].outlier = outl[ptr[i] + (int64_t) k];
	}
}
#endif /* SL_BUNDLEADJUST_H_ */
