// Test infrastructure: stand-in for LibVisualSLAM's geometry/SL_Distortion.h (declarations only).
#pragma once
void undistorPoint(const double* K, const double* kud, const double* in, double* out);
void invDistorParam(int w, int h, const double* iK, const double* kc, double* kud);
