/* Stand-in for LibVisualSLAM geometry/SL_RigidTransform.h. */
#pragma once
/* nearest rotation matrix (Frobenius norm) of a 3x3 matrix */
void approxRotationMat(const double* M, double* R);
