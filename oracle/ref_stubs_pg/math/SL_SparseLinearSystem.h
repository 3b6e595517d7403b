/* Stand-in for LibVisualSLAM math/SL_SparseLinearSystem.h. */
#pragma once
#include "math/SL_SparseMat.h"
/* least-squares solution of T x = b (T is m x n, m >= n) */
void sparseSolveLin(const Triplets& T, const double* b, double* x);
void sparseSolveLin(const SparseMat& A1, const SparseMat& A2, const double* b, double* x, double* y);
