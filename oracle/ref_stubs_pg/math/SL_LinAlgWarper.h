/* Stand-in for LibVisualSLAM math/SL_LinAlgWarper.h: declarations used by SL_GlobalPoseEstimation.cpp. */
#pragma once
#include "math/SL_Matrix.h"
void mat33AB(const double* A, const double* B, double* C);
void mat33Trans(const double* A, double* At);
void mat33ProdVec(const double* R, const double* x, const double* y, double* out, double a, double b);
/* off the post-BA path (constraint variants); declared to match the call sites, defined as aborting stubs */
void matTrans(const Mat_d& A, Mat_d& At);
void matAx(int m, int n, const double* A, const double* x, double* y);
void matQR(const Mat_d& A, Mat_d& Q, Mat_d& R);
