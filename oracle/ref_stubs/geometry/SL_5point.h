/* Stand-in for LibVisualSLAM geometry/SL_5point.h (epipolar variants, off the intraCamEstimate path). */
#pragma once
void formEMat(const double* R1, const double* t1, const double* R2, const double* t2, double* E);
