/* Stand-in for LibVisualSLAM SL_error.h (oracle/_ref build). */
#pragma once
#include <cassert>
#include <cstdarg>
#include <cstdio>
#include <stdexcept>
inline void repErr(const char* fmt, ...) { char b[1024]; va_list ap; va_start(ap, fmt); std::vsnprintf(b, sizeof(b), fmt, ap); va_end(ap); throw std::runtime_error(b); }
inline void logInfo(const char*, ...) {}
#define GET_FMT_STR(fmt, buf) { va_list ap_; va_start(ap_, fmt); std::vsnprintf(buf, sizeof(buf), fmt, ap_); va_end(ap_); }
