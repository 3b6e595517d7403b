// Test infrastructure (CPU): V3D_GPU::KLT_SequenceTracker (coslam_b200/shim/v3d_gpuklt.h) against a
// MOCK of the C-ABI: every method must forward exactly the arguments the reference's class receives
// (v3d_gpuklt.h:166-294, call order of tracking/GPUKLT.cpp:98-190).
#include <cstdio>
#include <cstring>
#include <string>

#include "v3d_gpuklt.h"

static std::string g_log;
static int g_fail = 0;
static cosl_klt* const HANDLE = reinterpret_cast<cosl_klt*>(0x1234);
#define EXPECT(c)                                        \
  do {                                                   \
    if (!(c)) {                                          \
      std::printf("mock: expectation failed: %s\n", #c); \
      g_fail = 1;                                        \
    }                                                    \
  } while (0)

extern "C" {
void cosl_klt_config_default(cosl_klt_config* c) {
  std::memset(c, 0, sizeof(*c));
  c->compat = 1;  // the default carries the ITER5 quirk: the shim must keep it
}
const char* cosl_last_error(void) { return "mock"; }
int cosl_klt_create(const cosl_klt_config* c, int w, int h, int nl, int fw, int fh, int plw, int plh,
                    int device, cosl_klt** out) {
  EXPECT(c->nIterations == 12 && c->nLevels == 6 && c->levelSkip == 2 && c->windowWidth == 6);
  EXPECT(c->trackBorderMargin == 4.0f && c->convergenceThreshold == 1.0f && c->SSD_Threshold == 20000.0f);
  EXPECT(c->trackWithGain == 1 && c->minDistance == 8 && c->minCornerness == 3000.0f);
  EXPECT(c->compat == 1);
  EXPECT(w == 640 && h == 480 && nl == 6 && fw == 32 && fh == 32 && plw == 64 && plh == 64 && device == 2);
  *out = HANDLE;
  g_log += "create;";
  return COSL_OK;
}
int cosl_klt_destroy(cosl_klt* h) {
  EXPECT(h == HANDLE);
  g_log += "destroy;";
  return COSL_OK;
}
int cosl_klt_set_margin(cosl_klt* h, float m) {
  EXPECT(h == HANDLE && m == 7.0f);
  g_log += "margin;";
  return COSL_OK;
}
int cosl_klt_set_conv(cosl_klt* h, float t) {
  EXPECT(h == HANDLE && t == 0.5f);
  g_log += "conv;";
  return COSL_OK;
}
int cosl_klt_set_ssd(cosl_klt* h, float t) {
  EXPECT(h == HANDLE && t == 123.0f);
  g_log += "ssd;";
  return COSL_OK;
}
int cosl_klt_detect(cosl_klt* h, const uint8_t* img, size_t pitch, int nPresent, const float* present,
                    cosl_klt_feature* dest, int* n) {
  EXPECT(h == HANDLE && img != 0 && pitch == 640 && dest != 0);
  EXPECT((nPresent == 0 && present == 0) || (nPresent == 3 && present != 0));
  dest[0].status = 1;
  dest[0].pos[0] = 0.25f;
  dest[0].pos[1] = 0.75f;
  dest[0].gain = 2.0f;
  dest[0].fed = nPresent ? 0 : -1;
  *n = nPresent ? 5 : 4;
  g_log += nPresent ? "detect3;" : "detect;";
  return COSL_OK;
}
int cosl_klt_redetect(cosl_klt* h, const uint8_t* img, size_t pitch, cosl_klt_feature* dest, int* n) {
  EXPECT(h == HANDLE && img != 0 && pitch == 640 && dest != 0);
  *n = 6;
  g_log += "redetect;";
  return COSL_OK;
}
int cosl_klt_track(cosl_klt* h, const uint8_t* img, size_t pitch, cosl_klt_feature* dest, int* n) {
  EXPECT(h == HANDLE && img != 0 && pitch == 640 && dest != 0);
  *n = 7;
  g_log += "track;";
  return COSL_E_CUDA;  // error path: the shim reports and carries on, like the reference's GL errors
}
int cosl_klt_feed(cosl_klt* h, int npts, const float* pts, int* ids, int* nFed) {
  EXPECT(h == HANDLE && npts == 2 && pts != 0 && ids != 0);
  ids[0] = 11;
  ids[1] = 12;
  *nFed = 2;
  g_log += "feed;";
  return COSL_OK;
}
int cosl_klt_advance(cosl_klt* h) {
  EXPECT(h == HANDLE);
  g_log += "advance;";
  return COSL_OK;
}
}

int main() {
  V3D_GPU::KLT_SequenceTrackerConfig cfg;  // what SingleSLAM::initTracker sets (SL_SingleSLAM.cpp:291-298)
  cfg.nLevels = 6;
  cfg.windowWidth = 6;
  cfg.convergenceThreshold = 1.0f;
  cfg.SSD_Threshold = 20000.0f;
  cfg.trackWithGain = true;
  cfg.minDistance = 8;
  cfg.minCornerness = 3000.0f;
  V3D_GPU::KLT_SequenceTracker trk(cfg);
  trk.setDevice(2);
  trk.allocate(640, 480, 6, 32, 32);
  static unsigned char img[640 * 480];
  V3D_GPU::KLT_TrackedFeature feats[1024];
  EXPECT(feats[5].status == -1 && feats[5].gain == 1.0f && feats[5].fed == -1);
  int n = 0;
  trk.detect(img, n, feats);
  EXPECT(n == 4 && feats[0].status == 1 && feats[0].pos[0] == 0.25f && feats[0].pos[1] == 0.75f);
  EXPECT(feats[0].gain == 2.0f && feats[0].fed == -1);
  float present[9] = {0};
  trk.detect(img, n, feats, 3, present);
  EXPECT(n == 5 && feats[0].fed == 0);
  trk.advanceFrame();
  trk.redetect(img, n, feats);
  EXPECT(n == 6);
  trk.track(img, n, feats);
  EXPECT(n == 7);
  float pts[6] = {0};
  int ids[2] = {0, 0}, nFed = 0;
  trk.feedExternFeaturePoints(2, pts, ids, nFed);
  EXPECT(nFed == 2 && ids[0] == 11 && ids[1] == 12);
  trk.setBorderMargin(7.0f);
  trk.setConvergenceThreshold(0.5f);
  trk.setSSD_Threshold(123.0f);
  EXPECT(trk.getCurrentFrameTextureID() == 0);
  trk.deallocate();
  trk.deallocate();  // idempotent
  EXPECT(g_log == "create;detect;detect3;advance;redetect;track;feed;margin;conv;ssd;destroy;");
  std::printf(g_fail ? "MOCK_KLT_SHIM_FAILED %s\n" : "MOCK_KLT_SHIM_OK %s\n", g_log.c_str());
  return g_fail;
}
