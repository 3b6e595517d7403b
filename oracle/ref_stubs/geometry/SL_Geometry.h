/* Stand-in for LibVisualSLAM geometry/SL_Geometry.h -- see math/SL_LinAlg.h in this directory. */
#pragma once
void project(const double* K, const double* R, const double* t, const double* M, double* m);
double reprojError2(const double* K, const double* R, const double* t, int npts, const double* Ms,
                    const double* ms);
/* off the intraCamEstimate path (covariance-weighted variants): declared so the file compiles,
 * defined as aborting stubs */
void getProjectionCovMat(const double* K, const double* R, const double* t, const double* M,
                         const double* cov, double* var, double sigma);
double mahaDist2(const double* a, const double* b, const double* ivar);
