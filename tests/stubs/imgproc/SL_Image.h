// Test infrastructure: stand-in for LibVisualSLAM's imgproc/SL_Image.h (see math/SL_Matrix.h).
#pragma once
#include "math/SL_Matrix.h"
typedef unsigned char uchar;
typedef MyMat<unsigned char> ImgG;
class ImgRGB {
 public:
  int w, h;
  unsigned char* data;
  ImgRGB() : w(0), h(0), data(0) {}
};
