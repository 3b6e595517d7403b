// Test infrastructure: minimal stand-in for LibVisualSLAM's math/SL_Matrix.h (the library is not
// part of the reference tree).  Only what the reference's KLT / pose / BA callers need to COMPILE in
// the syntax-only drop-in checks of tests/test_shim_reference_callers.py.
#pragma once
#include <cstring>
#include <vector>
using std::vector;  // LibVisualSLAM headers leak std names; the reference relies on it (slam/SL_MapPoint.h:131)
template <class T>
class MyMat {
 public:
  int rows, cols, m, n;
  T* data;
  MyMat() : rows(0), cols(0), m(0), n(0), data(0) {}
  MyMat(int r, int c) : rows(0), cols(0), m(0), n(0), data(0) { resize(r, c); }
  MyMat(const MyMat& o) : rows(0), cols(0), m(0), n(0), data(0) { cloneFrom(o.data, o.rows, o.cols); }
  ~MyMat() { delete[] data; }
  MyMat& operator=(const MyMat& o) {
    if (this != &o) cloneFrom(o.data, o.rows, o.cols);
    return *this;
  }
  void resize(int r, int c) {
    delete[] data;
    data = new T[(size_t)r * c]();
    rows = m = r;
    cols = n = c;
  }
  void cloneFrom(const T* src, int r, int c) {
    resize(r, c);
    if (src) std::memcpy(data, src, sizeof(T) * (size_t)r * c);
  }
  void fill(T v) {
    for (int i = 0; i < rows * cols; ++i) data[i] = v;
  }
  bool empty() const { return data == 0; }
  void clear() {
    delete[] data;
    data = 0;
    rows = cols = m = n = 0;
  }
  T& operator[](int i) { return data[i]; }
  const T& operator[](int i) const { return data[i]; }
  T& operator()(int i, int j) { return data[i * cols + j]; }
  operator T*() { return data; }
  operator const T*() const { return data; }
};
typedef MyMat<double> Mat_d;
typedef MyMat<float> Mat_f;
typedef MyMat<int> Mat_i;
typedef MyMat<unsigned char> Mat_uc;
