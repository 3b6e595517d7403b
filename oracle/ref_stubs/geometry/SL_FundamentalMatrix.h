/* Stand-in for LibVisualSLAM geometry/SL_FundamentalMatrix.h (off the intraCamEstimate path). */
#pragma once
void getFMat(const double* invK1, const double* invK2, const double* E, double* F);
double epipolarError(const double* F, const double* m1, const double* m2);
void computeEpipolarLine(const double* F, double x, double y, double* l);
