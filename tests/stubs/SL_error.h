// Test infrastructure: stand-in for LibVisualSLAM's SL_error.h (see math/SL_Matrix.h).
#pragma once
#include <cassert>
#include <cstdarg>
#include <cstdio>
#include <stdexcept>
inline void repErr(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  std::vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  throw std::runtime_error(buf);
}
inline void logInfo(const char*, ...) {}
