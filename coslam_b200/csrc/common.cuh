// common.cuh -- shared host/device helpers of libcoslam_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/coslam_b200.h"

namespace coslam {

// ---- error plumbing: the C-ABI never throws, every failure sets a thread-local message ----
std::string& last_error_ref();
int set_error(int code, const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define COSL_CUDA(expr)                                                                       \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return coslam::set_error(COSL_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                             __FILE__, __LINE__);                                             \
  } while (0)

#define COSL_TRY(expr)            \
  do {                            \
    int _rc = (expr);             \
    if (_rc != COSL_OK) return _rc; \
  } while (0)

// every kernel launch of this library goes through this macro so that gpu_launches is a count,
// not an estimate
#define COSL_LAUNCH(kernel, grid, block, smem, stream, ...)            \
  do {                                                                 \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);        \
    coslam::g_launches.fetch_add(1, std::memory_order_relaxed);          \
  } while (0)

inline int div_up(int a, int b) { return (a + b - 1) / b; }
inline int64_t div_up64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// simple per-stream section timer (CUDA events), used to report per-kernel-class device time
struct SectionTimer {
  static const int kMaxSections = 16;
  static const int kMaxPending = 4096;
  const char* names[kMaxSections];
  double ms[kMaxSections];
  int calls[kMaxSections];
  int nsec = 0;
  bool enabled = false;
  cudaEvent_t ev[kMaxPending][2];
  int pend_sec[kMaxPending];
  int npend = 0, nalloc = 0;
  int open_sec = -1;

  int section(const char* name) {
    for (int i = 0; i < nsec; ++i)
      if (!std::strcmp(names[i], name)) return i;
    if (nsec >= kMaxSections) return nsec - 1;
    names[nsec] = name;
    ms[nsec] = 0;
    calls[nsec] = 0;
    return nsec++;
  }
  void begin(int sec, cudaStream_t s) {
    if (!enabled) return;
    if (npend >= kMaxPending) flush();
    if (npend >= nalloc) {
      cudaEventCreate(&ev[nalloc][0]);
      cudaEventCreate(&ev[nalloc][1]);
      ++nalloc;
    }
    cudaEventRecord(ev[npend][0], s);
    pend_sec[npend] = sec;
    open_sec = sec;
  }
  void end(cudaStream_t s) {
    if (!enabled || open_sec < 0) return;
    cudaEventRecord(ev[npend][1], s);
    ++npend;
    open_sec = -1;
  }
  void flush() {  // requires the stream to be idle or will block on the last event
    for (int i = 0; i < npend; ++i) {
      cudaEventSynchronize(ev[i][1]);
      float t = 0;
      cudaEventElapsedTime(&t, ev[i][0], ev[i][1]);
      ms[pend_sec[i]] += t;
      calls[pend_sec[i]] += 1;
    }
    npend = 0;
  }
  void reset() {
    flush();
    for (int i = 0; i < nsec; ++i) {
      ms[i] = 0;
      calls[i] = 0;
    }
  }
  void destroy() {
    for (int i = 0; i < nalloc; ++i) {
      cudaEventDestroy(ev[i][0]);
      cudaEventDestroy(ev[i][1]);
    }
    nalloc = 0;
    npend = 0;
  }
};

}  // namespace coslam
