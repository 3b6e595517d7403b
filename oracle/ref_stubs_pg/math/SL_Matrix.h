/* Stand-in for LibVisualSLAM math/SL_Matrix.h for the oracle/_ref build of the reference's
 * slam/SL_GlobalPoseEstimation.cpp (TEST INFRASTRUCTURE).  Only what that file uses. */
#pragma once
#include <cassert>
#include <cstring>
#include <vector>
template <class T>
class MyMat {
 public:
  int rows, cols, m, n;
  T* data;
  MyMat() : rows(0), cols(0), m(0), n(0), data(0) {}
  MyMat(int r, int c) : rows(0), cols(0), m(0), n(0), data(0) { resize(r, c); }
  MyMat(const MyMat& o) : rows(0), cols(0), m(0), n(0), data(0) { resize(o.rows, o.cols); if (o.data) std::memcpy(data, o.data, sizeof(T) * (size_t)rows * cols); }
  ~MyMat() { delete[] data; }
  void resize(int r, int c) { delete[] data; data = new T[(size_t)r * c](); rows = m = r; cols = n = c; }
  void clear() { delete[] data; data = 0; rows = cols = m = n = 0; }
  void fill(T v) { for (int i = 0; i < rows * cols; ++i) data[i] = v; }
  T& operator[](int i) { return data[i]; }
  T& operator()(int r, int c) { return data[(size_t)r * cols + c]; }
  const T& operator()(int r, int c) const { return data[(size_t)r * cols + c]; }
  operator T*() { return data; }
  operator const T*() const { return data; }
};
typedef MyMat<double> Mat_d;
typedef MyMat<int> Mat_i;
