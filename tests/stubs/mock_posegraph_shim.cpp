// Test infrastructure (CPU): GlobalPoseGraph of coslam_b200/shim/SL_GlobalPoseEstimation.h against a
// MOCK of cosl_posegraph_spread_chains: flattening of nodes / edges, the newR/newt semantics of the two
// reference methods (Rotations: newt = t; Translations: newt = solution), one ABI call for both halves,
// the batch helper, and the exceptions for unsupported graphs and ABI failures.
#include <cstdio>
#include <cstring>
#include <stdexcept>

#include "SL_GlobalPoseEstimation.h"

static int g_fail = 0, g_calls = 0, g_rc = 0, g_expect_chains = 1;
#define EXPECT(c)                                        \
  do {                                                   \
    if (!(c)) {                                          \
      std::printf("mock: expectation failed: %s\n", #c); \
      g_fail = 1;                                        \
    }                                                    \
  } while (0)

extern "C" {
const char* cosl_last_error(void) { return "mock failure"; }
int cosl_posegraph_spread_chains(int nChains, const int* off, const uint8_t* fixed, const double* R,
                                 const double* t, const double* eR, const double* et, double* newR,
                                 double* newt, int device) {
  ++g_calls;
  if (g_rc) return g_rc;
  EXPECT(nChains == g_expect_chains && off[0] == 0 && device == 3);
  const int N = off[nChains];
  for (int c = 0; c < nChains; ++c) {
    const int a = off[c], n = off[c + 1] - a;
    for (int k = 0; k < n; ++k) {
      const int g = a + k;
      EXPECT(fixed[g] == (k % 3 == 0 ? 1 : 0));
      EXPECT(R[9 * g] == 100 * c + k && R[9 * g + 8] == 100 * c + k + 0.5 && t[3 * g + 2] == -(100 * c + k));
      if (k + 1 < n) EXPECT(eR[9 * g + 4] == 1000 + 100 * c + k && et[3 * g + 1] == 2000 + 100 * c + k);
    }
  }
  for (int g = 0; g < N; ++g) {
    for (int i = 0; i < 9; ++i) newR[9 * g + i] = 7000 + g + 0.01 * i;
    for (int i = 0; i < 3; ++i) newt[3 * g + i] = 8000 + g + 0.01 * i;
  }
  return COSL_OK;
}
}

static void fill(GlobalPoseGraph& g, int c, int n) {
  g.device = 3;
  g.reserve(n, n);
  for (int k = 0; k < n; ++k) {
    double R[9] = {0}, t[3] = {0};
    R[0] = 100 * c + k;
    R[8] = 100 * c + k + 0.5;
    t[2] = -(100 * c + k);
    CamPoseNode* nd = g.newNode();
    nd->set(k, c, R, t);
    nd->fixed = k % 3 == 0;
    EXPECT(nd->id == k);
  }
  for (int k = 0; k + 1 < n; ++k) {
    double R[9] = {0}, t[3] = {0};
    R[4] = 1000 + 100 * c + k;
    t[1] = 2000 + 100 * c + k;
    g.addEdge()->set(k, k + 1, R, t);
  }
}

int main() {
  GlobalPoseGraph g;
  fill(g, 0, 7);
  g.computeNewCameraRotations();
  EXPECT(g_calls == 1);
  EXPECT(g.poseNodes[2].newR[3] == 7002.03 && g.poseNodes[2].newt[2] == -2.0);  // newt = t after the rotation half
  g.computeNewCameraTranslations();
  EXPECT(g_calls == 1);                                                          // both halves came from ONE call
  EXPECT(g.poseNodes[2].newR[3] == 7002.03 && g.poseNodes[2].newt[2] == 8002.02);
  EXPECT(g.poseNodes[3].fixed && g.poseNodes[3].newR[0] == 3.0);                // fixed: newR = R (:339-343)
  g.computeNewCameraTranslations();  // translations alone: a fresh solve
  EXPECT(g_calls == 2);
  // a non-chain edge set is refused
  g.poseEdges[2].id2 = 5;
  bool threw = false;
  try { g.computeNewCameraRotations(); } catch (const std::runtime_error&) { threw = true; }
  EXPECT(threw && g_calls == 2);
  g.poseEdges[2].id2 = 3;
  g.poseEdges[1].uncertainScale = true;
  threw = false;
  try { g.computeNewCameraRotations(); } catch (const std::runtime_error&) { threw = true; }
  EXPECT(threw);
  g.poseEdges[1].uncertainScale = false;
  // ABI failure -> exception carrying cosl_last_error()
  g_rc = COSL_E_CUDA;
  threw = false;
  try { g.computeNewCameraRotations(); } catch (const std::runtime_error& e) { threw = std::strstr(e.what(), "mock failure") != 0; }
  EXPECT(threw);
  g_rc = 0;
  // batch helper: all cameras in one call
  GlobalPoseGraph gs[3];
  fill(gs[0], 0, 5);
  fill(gs[1], 1, 1);
  fill(gs[2], 2, 9);
  g_expect_chains = 3;
  const int before = g_calls;
  computeNewCameraPosesBatch(gs, 3, 3);
  EXPECT(g_calls == before + 1);
  EXPECT(gs[2].poseNodes[4].newR[0] == 7000 + 5 + 1 + 4 && gs[2].poseNodes[4].newt[1] == 8000 + 10 + 0.01);
  // empty graph: nothing to do, no call
  GlobalPoseGraph e;
  e.computeNewCameraRotations();
  e.computeNewCameraTranslations();
  EXPECT(g_calls == before + 1);
  if (!g_fail) std::printf("MOCK_POSEGRAPH_SHIM_OK\n");
  return g_fail;
}
