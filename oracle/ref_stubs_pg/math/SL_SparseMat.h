/* Stand-in for LibVisualSLAM math/SL_SparseMat.h: triplet container on the path, the rest only
 * declared (used by variants that are off the post-BA path). */
#pragma once
#include "math/SL_Matrix.h"
#include <vector>
class Triplets {
 public:
  int m, n;
  std::vector<int> ri, ci;
  std::vector<double> v;
  Triplets() : m(0), n(0) {}
  void reserve(int rows, int cols, int nnz) { m = rows; n = cols; ri.clear(); ci.clear(); v.clear(); ri.reserve(nnz); ci.reserve(nnz); v.reserve(nnz); }
  void add(int r, int c, double val) { ri.push_back(r); ci.push_back(c); v.push_back(val); }
};
class SparseMat {
 public:
  int m, n;
  SparseMat() : m(0), n(0) {}
};
void triplets2Sparse(const Triplets& T, SparseMat& A);
void tripletsSplitCol(const Triplets& T, int c, Triplets& T1, Triplets& T2);
void sparseSplitCol(const SparseMat& A, int c, SparseMat& A1, bool left);
void sparseMatMul(const SparseMat& A, const SparseMat& B, SparseMat& C);
void dense2Sparse(const Mat_d& A, SparseMat& S);
