// -*- C++ -*-
// Drop-in replacement for tracking/CGKLT/v3d_gpuklt.h of danping/CoSLAM: the same class and POD
// names, fields and method signatures (V3D_GPU::KLT_SequenceTrackerConfig, KLT_TrackedFeature,
// KLT_SequenceTracker -- v3d_gpuklt.h:166-294), forwarding to the C-ABI of libcoslam_b200.so.
// tracking/GPUKLT.cpp and tracking/test_klt_for_video.cpp compile against this header unchanged;
// no OpenGL / Cg / GLEW context is needed any more.  Errors (which the reference printed to cerr
// through checkGLErrorsHere0) are printed to stderr with the C-ABI's message.
#ifndef V3D_GPU_KLT_H
#define V3D_GPU_KLT_H

#include <cstdio>
#include <cstring>

#include "coslam_b200.h"

namespace V3D_GPU {

struct KLT_TrackedFeature {
	KLT_TrackedFeature() :
			status(-1), gain(1.0f), fed(-1) {
	}
	//! 0 means tracked from previous frame, 1 is newly created and -1 means invalidated track.
	int status;
	float pos[2];
	float gain;
	int fed; // >=0 is the id of a feature point fed to tracking
};

struct KLT_SequenceTrackerConfig {
	KLT_SequenceTrackerConfig() :
			nIterations(12), nLevels(3), levelSkip(2), windowWidth(5), trackBorderMargin(4.0f),
			convergenceThreshold(0.1f), SSD_Threshold(5000.0f), trackWithGain(false), minDistance(8),
			minCornerness(1000.0f), detectBorderMargin(4.0f) {
	}
	int nIterations, nLevels, levelSkip, windowWidth;
	float trackBorderMargin, convergenceThreshold, SSD_Threshold;
	bool trackWithGain;
	int minDistance;
	float minCornerness, detectBorderMargin;
};

struct KLT_SequenceTracker {
	KLT_SequenceTracker(KLT_SequenceTrackerConfig const& config) :
			_config(config), _h(0), _device(0) {
	}
	~KLT_SequenceTracker() {
	}
	//! extension: CUDA ordinal used by allocate() (default 0)
	void setDevice(int device) {
		_device = device;
	}
	void allocate(int width, int height, int nLevels, int featuresWidth, int featuresHeight) {
		this->allocate(width, height, nLevels, featuresWidth, featuresHeight, 2 * featuresWidth, 2 * featuresHeight);
	}
	void allocate(int width, int height, int nLevels, int featuresWidth, int featuresHeight, int pointListWidth,
			int pointListHeight) {
		cosl_klt_config c;
		cosl_klt_config_default(&c);
		c.nIterations = _config.nIterations;
		c.nLevels = nLevels;
		c.levelSkip = _config.levelSkip;
		c.windowWidth = _config.windowWidth;
		c.trackBorderMargin = _config.trackBorderMargin;
		c.convergenceThreshold = _config.convergenceThreshold;
		c.SSD_Threshold = _config.SSD_Threshold;
		c.trackWithGain = _config.trackWithGain ? 1 : 0;
		c.minDistance = _config.minDistance;
		c.minCornerness = _config.minCornerness;
		c.detectBorderMargin = _config.detectBorderMargin;
		_width = width;
		check(cosl_klt_create(&c, width, height, nLevels, featuresWidth, featuresHeight, pointListWidth,
				pointListHeight, _device, &_h), "allocate");
	}
	void deallocate() {
		if (_h)
			cosl_klt_destroy(_h);
		_h = 0;
	}
	void setBorderMargin(float margin) {
		check(cosl_klt_set_margin(_h, margin), "setBorderMargin");
	}
	void setConvergenceThreshold(float thr) {
		check(cosl_klt_set_conv(_h, thr), "setConvergenceThreshold");
	}
	void setSSD_Threshold(float thr) {
		check(cosl_klt_set_ssd(_h, thr), "setSSD_Threshold");
	}
	void detect(unsigned char const * image, int& nDetectedFeatures, KLT_TrackedFeature * dest) {
		check(cosl_klt_detect(_h, image, (size_t) _width, 0, 0, cast(dest), &nDetectedFeatures), "detect");
	}
	void detect(unsigned char const * image, int& nDetectedFeatures, KLT_TrackedFeature* dest, int nPresent,
			float* present) {
		check(cosl_klt_detect(_h, image, (size_t) _width, nPresent, present, cast(dest), &nDetectedFeatures), "detect");
	}
	void redetect(unsigned char const * image, int& nNewFeatures, KLT_TrackedFeature * dest) {
		check(cosl_klt_redetect(_h, image, (size_t) _width, cast(dest), &nNewFeatures), "redetect");
	}
	void feedExternFeaturePoints(int npts, float* featPts, int * trackIds, int& nFed) {
		check(cosl_klt_feed(_h, npts, featPts, trackIds, &nFed), "feedExternFeaturePoints");
	}
	void track(unsigned char const * image, int& nPresentFeatures, KLT_TrackedFeature * dest) {
		check(cosl_klt_track(_h, image, (size_t) _width, cast(dest), &nPresentFeatures), "track");
	}
	void advanceFrame() {
		check(cosl_klt_advance(_h), "advanceFrame");
	}
	//! GUI only in the reference (GL texture of the current frame); there is no GL texture here.
	unsigned int getCurrentFrameTextureID() const {
		return 0;
	}

protected:
	static cosl_klt_feature* cast(KLT_TrackedFeature* p) {
		// identical layout: {int status; float pos[2]; float gain; int fed;}
		return reinterpret_cast<cosl_klt_feature*>(p);
	}
	static void check(int rc, const char* what) {
		if (rc != COSL_OK)
			std::fprintf(stderr, "KLT_SequenceTracker::%s: %s\n", what, cosl_last_error());
	}
	KLT_SequenceTrackerConfig const _config;
	cosl_klt* _h;
	int _device, _width;
};

static_assert(sizeof(KLT_TrackedFeature) == sizeof(cosl_klt_feature), "KLT_TrackedFeature layout");

} // end namespace V3D_GPU

#endif
