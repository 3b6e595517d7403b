// klt.cu -- host side of the KLT tracker group + its C-ABI (see include/coslam_b200.h).
// Replaces V3D_GPU::KLT_SequenceTracker (tracking/CGKLT/v3d_gpuklt.{h,cpp}) and its GL/Cg runtime.
// All sequencing (status mapping, top-N selection, slot refill, provide) runs on the device; a
// frame costs one H2D copy per camera, one D2H copy of the feature table and one synchronise.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "common.cuh"
#include "klt_kernels.cuh"
#include "klt_track.cuh"
#include "klt_front.cuh"
#include "klt_pyr_tma.cuh"

using namespace coslam;

struct cosl_klt {
  cosl_klt_config cfg;
  int C = 1, W = 0, H = 0, L = 0, fw = 0, fh = 0, F = 0, plCap = 0, device = 0;
  int levelSkip = 1, halfWidth = 3;
  float trackMargin = 4.f, conv = 0.1f, ssd = 5000.f, detectMargin = 10.f;
  cudaStream_t stream = nullptr;
  // host uploads run on their own stream, one event per camera, so that klt_front of camera c
  // overlaps the copy of camera c+1 (cosl_klt_group_next path)
  cudaStream_t copyStream = nullptr;
  cudaEvent_t evImg[64] = {};
  cudaEvent_t evFront = nullptr;
  bool imgPerCam = false, frontRecorded = false;
  // the detector's prefilter needs only the cornerness map: on the re-detect path it runs on a side
  // stream next to the coarse pyramid levels and the tracker (valid on the unsuppressed map: the
  // suppression marks can only remove candidates, and klt_nm_verify reads the marked map)
  cudaStream_t auxStream = nullptr;
  cudaEvent_t evPre = nullptr;
  bool earlyPrefilter = false, prefilterDone = false;
  // geometry
  int lvW[8], lvH[8];
  long long lvOff[8];
  long long pyrStride = 0;  // float4 elements per camera
  size_t imgPitch = 0, imgStride = 0;
  int candCap = 0, smemKeys = 0;
  // device memory
  uint8_t* d_img = nullptr;
  float4* d_pyr[2] = {nullptr, nullptr};
  int cur = 1;  // pyr1 == d_pyr[cur]
  float* d_corn = nullptr;
  float4 *d_src = nullptr, *d_dst = nullptr, *d_ping = nullptr, *d_pong = nullptr, *d_res = nullptr;
  float4* d_present = nullptr;
  int* d_nbr = nullptr;
  unsigned long long* d_cand = nullptr;
  unsigned* d_prelim = nullptr;  // [C][prelimCap] packed (y << 16 | x) of 3x3 local maxima
  int prelimCap = 0;
  int* d_counters = nullptr;
  cosl_klt_feature* d_feat = nullptr;
  float* d_feedpts = nullptr;
  int* d_feedids = nullptr;
  int feedCap = 0;
  // pinned host mirrors
  cosl_klt_feature* h_feat = nullptr;
  int* h_counters = nullptr;
  float4* h_present = nullptr;
  // frame-pipelined ingest (cosl_klt_group_submit / _collect): two result slots, completion events
  cosl_klt_feature* h_featQ[2] = {nullptr, nullptr};
  int* h_cntQ[2] = {nullptr, nullptr};
  cudaEvent_t evDone[2] = {nullptr, nullptr};
  unsigned subSeq = 0;
  int nInFlight = 0;
  bool havePrev = false;
  bool cornValid = false;  // cornerness of the current frame already produced by klt_front
  // persistent fused gain tracker (klt_gain_fused)
  float4* d_state = nullptr;             // [C*F]
  unsigned long long* d_ver = nullptr;   // [2][C*F] (pass << 32 | beta bits)
  int* d_waitset = nullptr;   // [F][16]
  bool fusedOK = false;
  int fusedBlocks = 0;
  int numSM = 148;
  // launch folding on the combined next() path
  bool wantTail = false, statusDone = false, suppressDone = false;
  bool foldAdvance = false, advanceDone = false;
  int detXlo = 0, detXhi = -1, detYlo = 0, detYhi = -1;  // detector window in pixels
  int verBase = 0;
  // TMA descriptors of every pyramid level of both buffers: [buffer][level], load box (source of
  // level + 1) and store box (destination of level - 1 -> level); klt_pyr_tma.cuh
  CUtensorMap tmLoad[2][8], tmStore[2][8];
  bool tmaOK = false;
  SectionTimer timer;
  int secPyr = 0, secTrack = 0, secDetect = 0, secSelect = 0;
};

namespace {

// Detector window (klt_detector_pass2.cg: the pixel centre (i + 0.5) / n must lie inside
// [margin / n, 1 - margin / n]); evaluated with the same fp32 expressions the stand-alone
// cornerness kernel evaluates per pixel, so the integer bounds select exactly the same pixels.
static void update_detect_window(cosl_klt* g) {
  const float mg = g->detectMargin;
  auto bounds = [](int n, float lo, float hi, int& i0, int& i1) {
    const float nf = (float)n;
    i0 = n;
    i1 = -1;
    for (int i = 0; i < n; ++i) {
      const float sc = ((float)i + 0.5f) / nf;
      if (sc >= lo && sc <= hi) {
        i0 = std::min(i0, i);
        i1 = std::max(i1, i);
      }
    }
  };
  const float Wf = (float)g->W, Hf = (float)g->H;
  bounds(g->W, mg / Wf, 1.0f - mg / Wf, g->detXlo, g->detXhi);
  bounds(g->H, mg / Hf, 1.0f - mg / Hf, g->detYlo, g->detYhi);
}

int alloc_group(cosl_klt* g) {
  const int C = g->C, W = g->W, H = g->H, F = g->F;
  long long off = 0;
  for (int l = 0; l < g->L; ++l) {
    g->lvW[l] = W >> l;
    g->lvH[l] = H >> l;
    g->lvOff[l] = off;
    off += (long long)g->lvW[l] * g->lvH[l];
  }
  g->pyrStride = off;
  g->imgPitch = ((size_t)W + 15) & ~(size_t)15;
  g->imgStride = g->imgPitch * H;
  // candidate capacity: a strict (2r+1)^2 maximum admits at most one survivor per (r+1)^2 pixels
  const int r = std::max(1, g->cfg.minDistance);
  long long maxCand = ((long long)(W + r) / (r + 1) + 1) * ((long long)(H + r) / (r + 1) + 1);
  int cap = 1024;
  while (cap < maxCand) cap <<= 1;
  g->candCap = cap;
  g->smemKeys = std::min(cap, 16384);
  COSL_CUDA(cudaMalloc(&g->d_img, g->imgStride * C));
  for (int b = 0; b < 2; ++b) {
    COSL_CUDA(cudaMalloc(&g->d_pyr[b], sizeof(float4) * g->pyrStride * C));
    COSL_CUDA(cudaMemsetAsync(g->d_pyr[b], 0, sizeof(float4) * g->pyrStride * C, g->stream));
  }
  COSL_CUDA(cudaMalloc(&g->d_corn, sizeof(float) * (size_t)W * H * C));
  const size_t fbytes = sizeof(float4) * (size_t)F * C;
  float4** bufs[] = {&g->d_src, &g->d_dst, &g->d_ping, &g->d_pong, &g->d_res, &g->d_present};
  std::vector<float4> init((size_t)F * C, make_float4(-1.f, -1.f, 1.f, 0.f));
  for (auto pb : bufs) {
    COSL_CUDA(cudaMalloc(pb, fbytes));
    COSL_CUDA(cudaMemcpy(*pb, init.data(), fbytes, cudaMemcpyHostToDevice));
  }
  COSL_CUDA(cudaMalloc(&g->d_cand, sizeof(unsigned long long) * (size_t)cap * C));
  // strict 3x3 maxima cannot be 8-adjacent: at most one per 2x2 block
  g->prelimCap = ((W + 1) / 2) * ((H + 1) / 2);
  COSL_CUDA(cudaMalloc(&g->d_prelim, sizeof(unsigned) * (size_t)g->prelimCap * C));
  COSL_CUDA(cudaMalloc(&g->d_counters, sizeof(int) * 8 * C));
  COSL_CUDA(cudaMemset(g->d_counters, 0, sizeof(int) * 8 * C));
  COSL_CUDA(cudaMalloc(&g->d_feat, sizeof(cosl_klt_feature) * (size_t)F * C));
  COSL_CUDA(cudaMallocHost(&g->h_feat, sizeof(cosl_klt_feature) * (size_t)F * C));
  COSL_CUDA(cudaMallocHost(&g->h_counters, sizeof(int) * 8 * C));
  COSL_CUDA(cudaMallocHost(&g->h_present, fbytes));
  // neighbour slots of the gain smoothness term (klt_tracker_with_gain.cg:64-75): NEAREST,
  // CLAMP_TO_EDGE lookups in the featW x featH slot texture at st0 +- ds0; betaN1 adds the SCALAR
  // ds0.x (resp. ds0.y) to both coordinates.
  std::vector<int> nbr((size_t)F * 8);
  const double fw = g->fw, fh = g->fh;
  for (int sy = 0; sy < g->fh; ++sy)
    for (int sx = 0; sx < g->fw; ++sx) {
      const double cx = sx + 0.5, cy = sy + 0.5;
      const double offs[8][2] = {{+1.0, +fh / fw}, {-1.0, -fh / fw}, {+fw / fh, +1.0},
                                 {-fw / fh, -1.0}, {+1.0, 0.0},      {-1.0, 0.0},
                                 {0.0, +1.0},      {0.0, -1.0}};
      for (int k = 0; k < 8; ++k) {
        const int nx = std::min(std::max((int)std::floor(cx + offs[k][0]), 0), g->fw - 1);
        const int ny = std::min(std::max((int)std::floor(cy + offs[k][1]), 0), g->fh - 1);
        nbr[(size_t)(sy * g->fw + sx) * 8 + k] = ny * g->fw + nx;
      }
    }
  COSL_CUDA(cudaMalloc(&g->d_nbr, sizeof(int) * nbr.size()));
  COSL_CUDA(cudaMemcpy(g->d_nbr, nbr.data(), sizeof(int) * nbr.size(), cudaMemcpyHostToDevice));
  // wait sets of the persistent gain tracker: 8 neighbours + up to 8 reverse neighbours (slots
  // that read this slot but are not among its neighbours are enough; duplicates are harmless)
  {
    std::vector<int> ws((size_t)F * 16, -1);
    std::vector<int> nrev(F, 0);
    bool ok = true;
    for (int i = 0; i < F; ++i)
      for (int k = 0; k < 8; ++k) ws[(size_t)i * 16 + k] = nbr[(size_t)i * 8 + k];
    for (int i = 0; i < F && ok; ++i)
      for (int k = 0; k < 8 && ok; ++k) {
        const int j = nbr[(size_t)i * 8 + k];  // i reads j  ->  j must wait for i
        bool have = false;
        for (int q = 0; q < 8 + nrev[j]; ++q)
          if (ws[(size_t)j * 16 + q] == i) have = true;
        if (have) continue;
        if (nrev[j] >= 8) {
          ok = false;
          break;
        }
        ws[(size_t)j * 16 + 8 + nrev[j]++] = i;
      }
    int dev = 0, coop = 0, nsm = 0, perSM = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    g->numSM = std::max(1, nsm);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, klt_gain_fused, KLT_FUSED_THREADS, 0);
    g->fusedOK = ok && coop && perSM > 0 && !(g->cfg.compat & COSL_KLT_PASS_KERNELS);
    const int T = F * C;
    g->fusedBlocks = std::max(1, std::min(div_up(T, KLT_GPB), nsm * perSM));
    COSL_CUDA(cudaMalloc(&g->d_waitset, sizeof(int) * ws.size()));
    COSL_CUDA(cudaMemcpy(g->d_waitset, ws.data(), sizeof(int) * ws.size(), cudaMemcpyHostToDevice));
    COSL_CUDA(cudaMalloc(&g->d_state, sizeof(float4) * 2 * (size_t)T));
    COSL_CUDA(cudaMalloc(&g->d_ver, sizeof(unsigned long long) * 2 * (size_t)T));
    COSL_CUDA(cudaMemset(g->d_ver, 0, sizeof(unsigned long long) * 2 * (size_t)T));
    g->verBase = 0;
  }
  // tensor maps for the TMA pyramid kernel (driver entry point through the runtime: no -lcuda)
  g->tmaOK = false;
  if (!std::getenv("COSL_KLT_NO_TMA") && g->L > 2) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) == cudaSuccess && fn &&
        qr == cudaDriverEntryPointSuccess) {
      bool ok = true;
      for (int b = 0; b < 2 && ok; ++b)
        for (int l = 1; l < g->L && ok; ++l) {
          // a level of the whole camera group: x = texels as pairs of 8-byte elements, y, camera
          const cuuint64_t dims[3] = {2ull * g->lvW[l], (cuuint64_t)g->lvH[l], (cuuint64_t)C};
          const cuuint64_t strides[2] = {16ull * g->lvW[l], 16ull * (cuuint64_t)g->pyrStride};
          const cuuint32_t es[3] = {1, 1, 1};
          const cuuint32_t boxL[3] = {2 * PT_SW, PT_SH, 1}, boxS[3] = {2 * PT_TW, PT_TH, 1};
          void* base = (void*)(g->d_pyr[b] + g->lvOff[l]);
          ok = ok && ((EncodeFn)fn)(&g->tmLoad[b][l], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, boxL,
                                    es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
          ok = ok && ((EncodeFn)fn)(&g->tmStore[b][l], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, boxS,
                                    es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        }
      g->tmaOK = ok;
    }
  }
  // dynamic shared memory opt-ins
  COSL_CUDA(cudaFuncSetAttribute(klt_select_refill, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 g->smemKeys * (int)sizeof(unsigned long long)));
  update_detect_window(g);
  g->secPyr = g->timer.section("klt_pyramid");
  g->secTrack = g->timer.section("klt_track");
  g->secDetect = g->timer.section("klt_detect");
  g->secSelect = g->timer.section("klt_select");
  COSL_CUDA(cudaStreamSynchronize(g->stream));
  return COSL_OK;
}

void free_group(cosl_klt* g) {
  if (!g) return;
  cudaSetDevice(g->device);
  if (g->stream) cudaStreamSynchronize(g->stream);
  g->timer.destroy();
  cudaFree(g->d_img);
  cudaFree(g->d_pyr[0]);
  cudaFree(g->d_pyr[1]);
  cudaFree(g->d_corn);
  cudaFree(g->d_src);
  cudaFree(g->d_dst);
  cudaFree(g->d_ping);
  cudaFree(g->d_pong);
  cudaFree(g->d_res);
  cudaFree(g->d_present);
  cudaFree(g->d_nbr);
  if (g->auxStream) cudaStreamDestroy(g->auxStream);
  if (g->evPre) cudaEventDestroy(g->evPre);
  if (g->copyStream) cudaStreamDestroy(g->copyStream);
  if (g->evFront) cudaEventDestroy(g->evFront);
  for (int c = 0; c < 64; ++c)
    if (g->evImg[c]) cudaEventDestroy(g->evImg[c]);
  cudaFree(g->d_cand);
  cudaFree(g->d_prelim);
  cudaFree(g->d_counters);
  cudaFree(g->d_feat);
  cudaFree(g->d_feedpts);
  cudaFree(g->d_feedids);
  cudaFree(g->d_state);
  cudaFree(g->d_ver);
  cudaFree(g->d_waitset);
  cudaFreeHost(g->h_feat);
  cudaFreeHost(g->h_counters);
  cudaFreeHost(g->h_present);
  for (int i = 0; i < 2; ++i) {
    if (g->h_featQ[i]) cudaFreeHost(g->h_featQ[i]);
    if (g->h_cntQ[i]) cudaFreeHost(g->h_cntQ[i]);
    if (g->evDone[i]) cudaEventDestroy(g->evDone[i]);
  }
  if (g->stream) cudaStreamDestroy(g->stream);
  delete g;
}

// ---- stages (all asynchronous on g->stream) ----

int upload_images(cosl_klt* g, const uint8_t* const* imgs, size_t pitch, cudaMemcpyKind kind) {
  if (kind == cudaMemcpyHostToDevice && g->C > 1 && g->copyStream) {
    // the previous frame's front pass must have consumed d_img before it is overwritten
    if (g->frontRecorded) COSL_CUDA(cudaStreamWaitEvent(g->copyStream, g->evFront, 0));
    const bool flat = (pitch == (size_t)g->W) && (g->imgPitch == (size_t)g->W);  // one contiguous block
    for (int c = 0; c < g->C; ++c) {
      if (flat)
        COSL_CUDA(cudaMemcpyAsync(g->d_img + (size_t)c * g->imgStride, imgs[c], (size_t)g->W * g->H,
                                  kind, g->copyStream));
      else
        COSL_CUDA(cudaMemcpy2DAsync(g->d_img + (size_t)c * g->imgStride, g->imgPitch, imgs[c],
                                    pitch, g->W, g->H, kind, g->copyStream));
      COSL_CUDA(cudaEventRecord(g->evImg[c], g->copyStream));
    }
    g->imgPerCam = true;
    return COSL_OK;
  }
  for (int c = 0; c < g->C; ++c)
    COSL_CUDA(cudaMemcpy2DAsync(g->d_img + (size_t)c * g->imgStride, g->imgPitch, imgs[c], pitch,
                                g->W, g->H, kind, g->stream));
  return COSL_OK;
}

// klt_nm_prefilter on `st` (prefilter radius = min(minDistance, 3))
static void launch_prefilter(cosl_klt* g, cudaStream_t st) {
  const int pr = std::min(std::max(1, g->cfg.minDistance), 3);
  dim3 gp(div_up(g->W, 32 - 2 * pr), div_up(g->H, 8 * NP_ROWS), g->C);
  if (pr == 3)
    COSL_LAUNCH(klt_nm_prefilter<3>, gp, 256, 0, st, g->d_corn, g->W, g->H, g->d_prelim,
                g->prelimCap, g->d_counters);
  else if (pr == 2)
    COSL_LAUNCH(klt_nm_prefilter<2>, gp, 256, 0, st, g->d_corn, g->W, g->H, g->d_prelim,
                g->prelimCap, g->d_counters);
  else
    COSL_LAUNCH(klt_nm_prefilter<1>, gp, 256, 0, st, g->d_corn, g->W, g->H, g->d_prelim,
                g->prelimCap, g->d_counters);
}

// PyramidWithDerivativesCreator::buildPyramidForGrayscaleImage (v3d_gpupyramid.cpp:366-429)
int build_pyramid(cosl_klt* g, bool wantCorn) {
  float4* P = g->d_pyr[g->cur];
  g->timer.begin(g->secPyr, g->stream);
  // level 0 + level 1 + (optionally) the detector's cornerness map in one pass over the image
  FrontParams fp;
  fp.img = g->d_img;
  fp.imgPitch = g->imgPitch;
  fp.imgStride = g->imgStride;
  fp.pyr = P;
  fp.pyrStride = g->pyrStride;
  fp.lv1Off = g->L > 1 ? g->lvOff[1] : 0;
  fp.corn = g->d_corn;
  fp.W = g->W;
  fp.H = g->H;
  fp.nStrips = div_up(g->W, FS_SW);
  fp.nChunks = div_up(g->H, FS_R);
  fp.wantL1 = g->L > 1 ? 1 : 0;
  fp.wantCorn = wantCorn ? 1 : 0;
  fp.minC = g->cfg.minCornerness;
  fp.ixlo = g->detXlo;
  fp.ixhi = g->detXhi;
  fp.iylo = g->detYlo;
  fp.iyhi = g->detYhi;
  if (g->imgPerCam) {
    // two half-groups: the front pass of the first half overlaps the upload of the second (a launch
    // per camera would leave the SMs mostly empty: 500 warps per 1280x720 camera)
    g->imgPerCam = false;
    const int half = (g->C + 1) / 2;
    for (int c0 = 0; c0 < g->C; c0 += half) {
      const int nc = std::min(half, g->C - c0);
      COSL_CUDA(cudaStreamWaitEvent(g->stream, g->evImg[c0 + nc - 1], 0));
      fp.camBase = c0;
      dim3 g1(div_up(fp.nStrips * fp.nChunks, FS_WARPS), nc);
      COSL_LAUNCH(klt_front, g1, 32 * FS_WARPS, 0, g->stream, fp);
    }
  } else {
    fp.camBase = 0;
    dim3 g0(div_up(fp.nStrips * fp.nChunks, FS_WARPS), g->C);
    COSL_LAUNCH(klt_front, g0, 32 * FS_WARPS, 0, g->stream, fp);
  }
  if (g->evFront) {
    COSL_CUDA(cudaEventRecord(g->evFront, g->stream));
    g->frontRecorded = true;
  }
  if (g->earlyPrefilter && wantCorn && g->auxStream && g->evFront && g->evPre) {
    COSL_CUDA(cudaStreamWaitEvent(g->auxStream, g->evFront, 0));
    launch_prefilter(g, g->auxStream);
    COSL_CUDA(cudaEventRecord(g->evPre, g->auxStream));
    g->prefilterDone = true;
  }
  g->cornValid = wantCorn;
  for (int l = 2; l < g->L; ++l) {
    dim3 gl(div_up(g->lvW[l], PD_TW), div_up(g->lvH[l], PD_TH), g->C);
    if (g->tmaOK)  // source tile by TMA load, destination tile by TMA store (klt_pyr_tma.cuh)
      COSL_LAUNCH(klt_pyr_down_tma, gl, 256, 0, g->stream, g->tmLoad[g->cur][l - 1], g->tmStore[g->cur][l],
                  g->lvW[l - 1], g->lvH[l - 1]);
    else
      COSL_LAUNCH(klt_pyr_down, gl, 256, 0, g->stream, P + g->lvOff[l - 1], P + g->lvOff[l],
                  g->pyrStride, g->lvW[l - 1], g->lvH[l - 1], g->lvW[l], g->lvH[l]);
  }
  g->timer.end(g->stream);
  COSL_CUDA(cudaGetLastError());
  return COSL_OK;
}

KltTrackParams track_params(const cosl_klt* g, bool strict) {
  KltTrackParams P;
  P.W = g->W;
  P.H = g->H;
  P.F = g->F;
  P.halfWidth = g->halfWidth;
  const float Wf = (float)g->W, Hf = (float)g->H, m = g->trackMargin;
  if (strict) {
    P.sqrConv = g->conv * g->conv;
    P.ssdThr = g->ssd;
    P.vr0 = m / Wf;
    P.vr1 = m / Hf;
    P.vr2 = 1.0f - m / Wf;
    P.vr3 = 1.0f - m / Hf;
  } else {  // v3d_gpuklt.cpp:246-248, 266-270
    P.sqrConv = 1000000.0f;
    P.ssdThr = 1000000.0f;
    P.vr0 = -1.0f;
    P.vr1 = -1.0f;
    P.vr2 = 2.0f;
    P.vr3 = 2.0f;
  }
  P.lambda = 1.0f;
  P.delta = 200.0f;
  return P;
}

// KLT_TrackerWithGain::trackFeaturesAndGain (v3d_gpuklt.cpp:205-305) /
// KLT_Tracker::trackFeatures (:99-161); result -> d_res
int run_tracker(cosl_klt* g) {
  const float4* P0 = g->d_pyr[1 - g->cur];
  const float4* P1 = g->d_pyr[g->cur];
  const int wpb = 8;  // warps per block (2x2 tracker: one warp per slot)
  dim3 grid(div_up(g->F, wpb), g->C);
  dim3 gridHalf(div_up(g->F, 256 / KLT_G), g->C);  // gain tracker: one lane group per slot
  g->timer.begin(g->secTrack, g->stream);
  if (g->cfg.trackWithGain && g->fusedOK) {
    KltLevels LV;
    LV.n = 0;
    for (int level = g->L - 1; level >= 0; level -= g->levelSkip) {
      LV.level[LV.n] = level;
      LV.w[LV.n] = g->lvW[level];
      LV.h[LV.n] = g->lvH[level];
      LV.off[LV.n] = g->lvOff[level];
      LV.mult[LV.n] = 1.0f;
      ++LV.n;
    }
    int nIter = g->cfg.nIterations, C = g->C;
    const int nPass = LV.n * nIter;
    if (g->verBase > 2000000000 - 2 * nPass) {  // version wrap: restart the counters
      COSL_CUDA(cudaMemsetAsync(g->d_ver, 0, sizeof(unsigned long long) * 2 * (size_t)g->F * g->C,
                                g->stream));
      g->verBase = 0;
    }
    long long pyrStride = g->pyrStride;
    KltTrackParams Plax = track_params(g, false), Pstrict = track_params(g, true);
    int verBase = g->verBase;
    KltTail tail = {nullptr, nullptr, nullptr, 0, 0, 0};
    if (g->wantTail) {
      tail.dest = g->d_feat;
      tail.counters = g->d_counters;
      tail.corn = g->d_corn;
      tail.W = g->W;
      tail.H = g->H;
      tail.suppress = g->cornValid ? 1 : 0;  // only a map of THIS frame may be marked
      g->statusDone = true;
      g->suppressDone = tail.suppress != 0;
    }
    void* args[] = {(void*)&P0,          (void*)&P1,     (void*)&pyrStride, (void*)&LV,
                    (void*)&nIter,       (void*)&g->d_src, (void*)&g->d_state, (void*)&g->d_ver,
                    (void*)&g->d_waitset, (void*)&g->d_res, (void*)&C,        (void*)&Plax,
                    (void*)&Pstrict,     (void*)&verBase, (void*)&tail};
    COSL_CUDA(cudaLaunchCooperativeKernel((const void*)klt_gain_fused, dim3(g->fusedBlocks),
                                          dim3(KLT_FUSED_THREADS), args, 0, g->stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    g->verBase += nPass;
#if KLT_EXP_CLOCK
    {
      static int calls = 0;
      if (++calls == 50) {
        cudaStreamSynchronize(g->stream);
        const int nw = g->fusedBlocks * KLT_FUSED_THREADS / 32;
        std::vector<float> hs((size_t)nw * 12);
        cudaMemcpy(hs.data(), g->d_state, sizeof(float) * hs.size(), cudaMemcpyDeviceToHost);
        double acc[6] = {0, 0, 0, 0, 0, 0}, mx = 0;
        unsigned s0 = ~0u, s1 = 0, e0 = ~0u, e1 = 0;
        for (int i = 0; i < nw; ++i) {
          for (int k = 0; k < 6; ++k) acc[k] += hs[(size_t)i * 12 + k];
          mx = std::max(mx, (double)hs[(size_t)i * 12 + 5]);
          unsigned a, b;
          memcpy(&a, &hs[(size_t)i * 12 + 6], 4);
          memcpy(&b, &hs[(size_t)i * 12 + 7], 4);
          s0 = std::min(s0, a); s1 = std::max(s1, a); e0 = std::min(e0, b); e1 = std::max(e1, b);
        }
        if (const char* dump = std::getenv("COSL_KLT_CLOCK_DUMP")) {
          if (FILE* f = fopen(dump, "wb")) {
            fwrite(hs.data(), sizeof(float), hs.size(), f);
            fclose(f);
          }
        }
        fprintf(stderr, "[klt clock] globaltimer ns: start spread %u, end spread %u, first start -> last end %u, last start -> first end %u\n", s1 - s0, e1 - e0, e1 - s0, e0 - s1);
        fprintf(stderr, "[klt clock] warps %d  mean cycles: poll %.0f  pre+stage %.0f  loop %.0f  finish %.0f  (stage-only %.0f)  total %.0f max %.0f\n", nw,
                acc[0] / nw, acc[1] / nw, acc[2] / nw, acc[3] / nw, acc[4] / nw, acc[5] / nw, mx);
      }
    }
#endif
  } else if (g->cfg.trackWithGain) {
    const float4* in = g->d_src;
    float4* bufs[2] = {g->d_ping, g->d_pong};
    int which = 0, first = 1;
    bool strict = false;
    for (int level = g->L - 1; level >= 0; level -= g->levelSkip) {
      const int w = g->lvW[level], h = g->lvH[level];
      const float dsx = 1.0f / (float)w, dsy = 1.0f / (float)h;
      for (int iter = 1; iter <= g->cfg.nIterations; ++iter) {
        if (iter == 1)
          strict = false;
        else if (iter == g->cfg.nIterations)
          strict = true;
        const bool last =
            (level - g->levelSkip < 0) && (iter == g->cfg.nIterations);
        float4* out = last ? g->d_res : bufs[which];
        COSL_LAUNCH(klt_gain_pass, gridHalf, 256, 0, g->stream, P0, P1, g->pyrStride,
                    g->lvOff[level], w, h, g->d_src, in, out, g->d_nbr, dsx, dsy,
                    track_params(g, strict), first);
        in = out;
        which ^= 1;
        first = 0;
      }
    }
  } else {
    KltLevels LV;
    LV.n = 0;
    float mult = (float)(1 << (g->L - 1));
    for (int level = g->L - 1; level >= 0; level -= g->levelSkip) {
      LV.level[LV.n] = level;
      LV.w[LV.n] = g->lvW[level];
      LV.h[LV.n] = g->lvH[level];
      LV.off[LV.n] = g->lvOff[level];
      LV.mult[LV.n] = mult;
      mult /= (float)(1 << g->levelSkip);
      ++LV.n;
    }
    const int nIter = (g->cfg.compat & COSL_KLT_COMPAT_ITER5) ? 5 : g->cfg.nIterations;
    const int hw2 = g->halfWidth, npx2 = (2 * hw2 + 1) * (2 * hw2 + 1);
    const bool tiled = (npx2 <= KLT_G * KLT_ROUNDS) && (2 * hw2 + 4 <= KLT_TW) && (2 * hw2 + 2 <= 8) &&
                       !(g->cfg.compat & COSL_KLT_PASS_KERNELS);
    if (tiled) {
      dim3 gt(div_up(g->F, 16), g->C);
      COSL_LAUNCH(klt_track_2x2_tiled, gt, 128, 0, g->stream, P0, P1, g->pyrStride, LV, g->d_src,
                  g->d_res, track_params(g, true), nIter);
    } else {
      COSL_LAUNCH(klt_track_2x2, grid, wpb * 32, 0, g->stream, P0, P1, g->pyrStride, LV, g->d_src,
                  g->d_res, track_params(g, true), nIter);
    }
  }
  g->timer.end(g->stream);
  COSL_CUDA(cudaGetLastError());
  return COSL_OK;
}

int zero_counters(cosl_klt* g) {
  COSL_CUDA(cudaMemsetAsync(g->d_counters, 0, sizeof(int) * 8 * g->C, g->stream));
  return COSL_OK;
}

int run_status(cosl_klt* g) {
  dim3 grid(div_up(g->F, 256), g->C);
  COSL_LAUNCH(klt_status, grid, 256, 0, g->stream, g->d_res, g->d_feat, g->d_counters, g->F);
  return COSL_OK;
}

// KLT_Detector::detectCorners + extractCorners (v3d_gpuklt.cpp:423-588) and the slot logic of
// detect/redetect (:651-805).  mode 0: detect (present = external points), 1: redetect.
int run_detector(cosl_klt* g, int mode, int nPresentExt) {
  const float Wf = (float)g->W, Hf = (float)g->H, mg = g->detectMargin;
  g->timer.begin(g->secDetect, g->stream);
  if (!g->cornValid) {
    dim3 gc(div_up(g->W, DC_TW), div_up(g->H, DC_TH), g->C);
    COSL_LAUNCH(klt_cornerness, gc, 256, 0, g->stream, g->d_pyr[g->cur], g->pyrStride, g->d_corn,
                g->W, g->H, g->cfg.minCornerness, mg / Wf, mg / Hf, 1.0f - mg / Wf, 1.0f - mg / Hf);
  }
  g->cornValid = false;  // the suppression below modifies the map
  if (mode == 1 && g->suppressDone) {
    g->suppressDone = false;  // the tracker's last pass already marked the live tracks
  } else if (mode == 1) {
    dim3 gs(div_up(g->F, 256), g->C);
    COSL_LAUNCH(klt_suppress, gs, 256, 0, g->stream, g->d_res, g->F, g->F, g->d_corn, g->W, g->H);
  } else if (nPresentExt > 0) {
    dim3 gs(div_up(nPresentExt, 256), g->C);
    COSL_LAUNCH(klt_suppress, gs, 256, 0, g->stream, g->d_present, nPresentExt, g->F, g->d_corn,
                g->W, g->H);
  }
  const int r = std::max(1, g->cfg.minDistance);
  if (g->prefilterDone) {
    g->prefilterDone = false;
    COSL_CUDA(cudaStreamWaitEvent(g->stream, g->evPre, 0));
  } else {
    launch_prefilter(g, g->stream);
  }
  dim3 gv(2 * g->numSM, g->C);
  COSL_LAUNCH(klt_nm_verify, gv, 256, 0, g->stream, g->d_corn, g->W, g->H, r, g->d_prelim,
              g->prelimCap, g->d_cand, g->candCap, g->d_counters);
  g->timer.end(g->stream);
  g->timer.begin(g->secSelect, g->stream);
  COSL_LAUNCH(klt_select_refill, g->C, 1024, g->smemKeys * sizeof(unsigned long long), g->stream,
              g->d_cand, g->candCap, g->plCap, g->d_counters, g->d_feat, g->d_dst,
              g->foldAdvance ? g->d_src : (float4*)nullptr, g->d_present,
              nPresentExt, g->F, g->W, g->H, mode, g->cfg.trackWithGain ? 1 : 0, g->smemKeys,
              (g->cfg.compat & COSL_KLT_COMPAT_HISTOPYR) ? 1 : 0);
  g->timer.end(g->stream);
  g->advanceDone = g->foldAdvance;
  COSL_CUDA(cudaGetLastError());
  return COSL_OK;
}

int fetch_results(cosl_klt* g) {
  COSL_CUDA(cudaMemcpyAsync(g->h_feat, g->d_feat, sizeof(cosl_klt_feature) * (size_t)g->F * g->C,
                            cudaMemcpyDeviceToHost, g->stream));
  COSL_CUDA(cudaMemcpyAsync(g->h_counters, g->d_counters, sizeof(int) * 8 * g->C,
                            cudaMemcpyDeviceToHost, g->stream));
  COSL_CUDA(cudaStreamSynchronize(g->stream));
  return COSL_OK;
}

int advance(cosl_klt* g) {
  // swapFeatureBuffers + swap pyramids (v3d_gpuklt.h:252-259); the feature buffer is copied, not
  // swapped, so a later feed() still sees the provided set (see DESIGN.md, quirk list)
  const size_t n = (size_t)g->F * g->C;
  if (g->advanceDone)
    g->advanceDone = false;  // klt_select_refill already wrote the provided set to both buffers
  else
    COSL_LAUNCH(klt_copy_f4, (unsigned)div_up64(n, 256), 256, 0, g->stream, g->d_dst, g->d_src, n);
  g->cur = 1 - g->cur;
  return COSL_OK;
}

int do_track(cosl_klt* g, bool wantCorn = false) {
  COSL_TRY(zero_counters(g));  // before the front pass: the early prefilter counts into them
  COSL_TRY(build_pyramid(g, wantCorn));
  g->statusDone = g->suppressDone = false;
  g->wantTail = true;  // the persistent gain tracker writes the feature table itself
  COSL_TRY(run_tracker(g));
  g->wantTail = false;
  if (!g->statusDone) COSL_TRY(run_status(g));
  return COSL_OK;
}

int do_redetect(cosl_klt* g) {
  g->earlyPrefilter = true;
  const int rcTrack = do_track(g, true);
  g->earlyPrefilter = false;
  if (rcTrack != COSL_OK) return rcTrack;
  COSL_TRY(run_detector(g, 1, 0));
  return COSL_OK;
}

int do_detect(cosl_klt* g, int nPresentExt) {
  COSL_TRY(build_pyramid(g, true));
  COSL_TRY(zero_counters(g));
  COSL_TRY(run_detector(g, 0, nPresentExt));
  return COSL_OK;
}

void copy_out(cosl_klt* g, int cam, cosl_klt_feature* dest) {
  std::memcpy(dest, g->h_feat + (size_t)cam * g->F, sizeof(cosl_klt_feature) * g->F);
}

}  // namespace

/* ================================================================ C-ABI */
extern "C" {

void cosl_klt_config_default(cosl_klt_config* c) {
  if (!c) return;
  c->nIterations = 12;
  c->nLevels = 3;
  c->levelSkip = 2;
  c->windowWidth = 5;
  c->trackBorderMargin = 4.0f;
  c->convergenceThreshold = 0.1f;
  c->SSD_Threshold = 5000.0f;
  c->trackWithGain = 0;
  c->minDistance = 8;
  c->minCornerness = 1000.0f;
  c->detectBorderMargin = 4.0f;
  c->compat = COSL_KLT_COMPAT_ITER5;
}

int cosl_klt_group_create(const cosl_klt_config* cfg, int nCams, int width, int height,
                          int nLevels, int featW, int featH, int plW, int plH, int device,
                          cosl_klt** out) {
  if (!cfg || !out) return set_error(COSL_E_INVALID, "null argument");
  *out = nullptr;
  if (nCams < 1 || nCams > 64 || width < 16 || height < 16 || width > 65535 || height > 65535 ||
      nLevels < 1 || nLevels > 8 || featW < 1 || featH < 1 || (width >> (nLevels - 1)) < 2 ||
      (height >> (nLevels - 1)) < 2)
    return set_error(COSL_E_INVALID, "bad tracker geometry %dx%d L=%d feat=%dx%d C=%d", width,
                     height, nLevels, featW, featH, nCams);
  if (cfg->windowWidth < 1 || cfg->windowWidth > 31 || cfg->minDistance < 1 ||
      cfg->minDistance > 32 || cfg->nIterations < 1)
    return set_error(COSL_E_INVALID, "bad tracker config");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    return set_error(COSL_E_CUDA, "CUDA device %d not available (%d devices)", device, ndev);
  COSL_CUDA(cudaSetDevice(device));
  cosl_klt* g = new (std::nothrow) cosl_klt();
  if (!g) return set_error(COSL_E_NOMEM, "host allocation failed");
  g->cfg = *cfg;
  g->cfg.nLevels = nLevels;
  g->C = nCams;
  g->W = width;
  g->H = height;
  g->L = nLevels;
  g->fw = featW;
  g->fh = featH;
  g->F = featW * featH;
  if (plW <= 0) plW = 2 * featW;
  if (plH <= 0) plH = 2 * featH;
  g->plCap = plW * plH;
  g->device = device;
  g->levelSkip = cfg->levelSkip > 0 ? cfg->levelSkip : (nLevels - 1);
  if (g->levelSkip < 1) g->levelSkip = 1;
  g->halfWidth = cfg->windowWidth / 2;
  g->trackMargin = cfg->trackBorderMargin;
  g->conv = cfg->convergenceThreshold;
  g->ssd = cfg->SSD_Threshold;
  g->detectMargin = 10.0f;  // KLT_Detector::_margin default (v3d_gpuklt.h:113-114)
  cudaError_t e = cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete g;
    return set_error(COSL_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
  }
  if (cudaStreamCreateWithFlags(&g->copyStream, cudaStreamNonBlocking) == cudaSuccess) {
    bool ok = cudaEventCreateWithFlags(&g->evFront, cudaEventDisableTiming) == cudaSuccess;
    for (int c = 0; c < g->C && ok; ++c)
      ok = cudaEventCreateWithFlags(&g->evImg[c], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {  // fall back to uploads on the compute stream
      cudaStreamDestroy(g->copyStream);
      g->copyStream = nullptr;
    }
  } else {
    g->copyStream = nullptr;
  }
  if (cudaStreamCreateWithFlags(&g->auxStream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&g->evPre, cudaEventDisableTiming) != cudaSuccess) {
    if (g->auxStream) cudaStreamDestroy(g->auxStream);
    g->auxStream = nullptr;
  }
  int rc = alloc_group(g);
  if (rc != COSL_OK) {
    free_group(g);
    return rc;
  }
  *out = g;
  return COSL_OK;
}

int cosl_klt_create(const cosl_klt_config* cfg, int width, int height, int nLevels, int featW,
                    int featH, int plW, int plH, int device, cosl_klt** out) {
  return cosl_klt_group_create(cfg, 1, width, height, nLevels, featW, featH, plW, plH, device, out);
}

int cosl_klt_destroy(cosl_klt* h) {
  free_group(h);
  return COSL_OK;
}

#define KLT_ENTER(h)                                              \
  if (!(h)) return set_error(COSL_E_INVALID, "null tracker handle"); \
  COSL_CUDA(cudaSetDevice((h)->device));

#define KLT_SINGLE(h) \
  KLT_ENTER(h)        \
  if ((h)->C != 1) return set_error(COSL_E_STATE, "single-camera call on a %d-camera group", (h)->C);

int cosl_klt_detect(cosl_klt* h, const uint8_t* img, size_t pitch, int nPresent,
                    const float* present3, cosl_klt_feature* dest, int* nDetected) {
  KLT_SINGLE(h)
  if (!img || !dest || !nDetected || nPresent < 0 || nPresent > h->F || (nPresent && !present3) ||
      pitch < (size_t)h->W)
    return set_error(COSL_E_INVALID, "cosl_klt_detect: bad argument");
  for (int i = 0; i < nPresent; ++i)
    h->h_present[i] = make_float4(present3[3 * i], present3[3 * i + 1], 0.f, 0.f);
  if (nPresent)
    COSL_CUDA(cudaMemcpyAsync(h->d_present, h->h_present, sizeof(float4) * nPresent,
                              cudaMemcpyHostToDevice, h->stream));
  const uint8_t* imgs[1] = {img};
  COSL_TRY(upload_images(h, imgs, pitch, cudaMemcpyHostToDevice));
  COSL_TRY(do_detect(h, nPresent));
  COSL_TRY(fetch_results(h));
  copy_out(h, 0, dest);
  *nDetected = h->h_counters[2];
  return COSL_OK;
}

int cosl_klt_redetect(cosl_klt* h, const uint8_t* img, size_t pitch, cosl_klt_feature* dest,
                      int* nNewFeatures) {
  KLT_SINGLE(h)
  if (!img || !dest || !nNewFeatures || pitch < (size_t)h->W)
    return set_error(COSL_E_INVALID, "cosl_klt_redetect: bad argument");
  const uint8_t* imgs[1] = {img};
  COSL_TRY(upload_images(h, imgs, pitch, cudaMemcpyHostToDevice));
  COSL_TRY(do_redetect(h));
  COSL_TRY(fetch_results(h));
  copy_out(h, 0, dest);
  *nNewFeatures = h->h_counters[2];
  return COSL_OK;
}

int cosl_klt_track(cosl_klt* h, const uint8_t* img, size_t pitch, cosl_klt_feature* dest,
                   int* nPresent) {
  KLT_SINGLE(h)
  if (!img || !dest || !nPresent || pitch < (size_t)h->W)
    return set_error(COSL_E_INVALID, "cosl_klt_track: bad argument");
  const uint8_t* imgs[1] = {img};
  COSL_TRY(upload_images(h, imgs, pitch, cudaMemcpyHostToDevice));
  COSL_TRY(do_track(h));
  // provide: the tracked set becomes the feature table of the next frame
  const size_t n = (size_t)h->F;
  COSL_LAUNCH(klt_copy_f4, (unsigned)div_up64(n, 256), 256, 0, h->stream, h->d_res, h->d_dst, n);
  COSL_TRY(fetch_results(h));
  copy_out(h, 0, dest);
  *nPresent = h->h_counters[0];
  return COSL_OK;
}

int cosl_klt_feed(cosl_klt* h, int npts, const float* pts3, int* trackIds, int* nFed) {
  KLT_SINGLE(h)
  if (npts < 0 || (npts && (!pts3 || !trackIds)) || !nFed)
    return set_error(COSL_E_INVALID, "cosl_klt_feed: bad argument");
  *nFed = 0;
  if (npts == 0) return COSL_OK;
  if (npts > h->feedCap) {
    cudaFree(h->d_feedpts);
    cudaFree(h->d_feedids);
    h->d_feedpts = nullptr;
    h->d_feedids = nullptr;
    h->feedCap = 0;
    COSL_CUDA(cudaMalloc(&h->d_feedpts, sizeof(float) * 3 * npts));
    COSL_CUDA(cudaMalloc(&h->d_feedids, sizeof(int) * (npts + 1)));
    h->feedCap = npts;
  }
  COSL_CUDA(cudaMemcpyAsync(h->d_feedpts, pts3, sizeof(float) * 3 * npts, cudaMemcpyHostToDevice,
                            h->stream));
  COSL_CUDA(cudaMemsetAsync(h->d_feedids, 0xff, sizeof(int) * (npts + 1), h->stream));
  COSL_LAUNCH(klt_feed_kill, div_up(h->F, 256), 256, 0, h->stream, h->d_dst, h->F, h->d_feedpts,
              npts, (h->cfg.compat & COSL_KLT_COMPAT_FEED_STRIDE2) ? 2 : 3);
  COSL_LAUNCH(klt_feed_place, 1, 32, 0, h->stream, h->d_dst, h->F, h->d_feedpts, npts,
              h->d_feedids, h->d_feedids + npts);
  std::vector<int> ids(npts + 1);
  COSL_CUDA(cudaMemcpyAsync(ids.data(), h->d_feedids, sizeof(int) * (npts + 1),
                            cudaMemcpyDeviceToHost, h->stream));
  COSL_CUDA(cudaStreamSynchronize(h->stream));
  std::memcpy(trackIds, ids.data(), sizeof(int) * npts);
  *nFed = ids[npts];
  return COSL_OK;
}

int cosl_klt_advance(cosl_klt* h) {
  KLT_ENTER(h)
  COSL_TRY(advance(h));
  COSL_CUDA(cudaGetLastError());
  return COSL_OK;
}

int cosl_klt_set_margin(cosl_klt* h, float m) {  // tracker AND detector (v3d_gpuklt.h:217-224)
  KLT_ENTER(h)
  h->trackMargin = m;
  h->detectMargin = m;
  update_detect_window(h);
  return COSL_OK;
}
int cosl_klt_set_conv(cosl_klt* h, float t) {
  KLT_ENTER(h)
  h->conv = t;
  return COSL_OK;
}
int cosl_klt_set_ssd(cosl_klt* h, float t) {
  KLT_ENTER(h)
  h->ssd = t;
  return COSL_OK;
}

int cosl_klt_group_first(cosl_klt* h, const uint8_t* const* imgs, size_t pitch,
                         cosl_klt_feature* const* dest, int* nDetected) {
  KLT_ENTER(h)
  if (!imgs || pitch < (size_t)h->W) return set_error(COSL_E_INVALID, "group_first: bad argument");
  COSL_TRY(upload_images(h, imgs, pitch, cudaMemcpyHostToDevice));
  h->foldAdvance = true;  // the provided set goes to both feature buffers in one launch
  const int rcDet = do_detect(h, 0);
  h->foldAdvance = false;
  if (rcDet != COSL_OK) return rcDet;
  COSL_TRY(advance(h));
  COSL_TRY(fetch_results(h));
  for (int c = 0; c < h->C; ++c) {
    if (dest && dest[c]) copy_out(h, c, dest[c]);
    if (nDetected) nDetected[c] = h->h_counters[8 * c + 2];
  }
  return COSL_OK;
}

int cosl_klt_group_next(cosl_klt* h, const uint8_t* const* imgs, size_t pitch,
                        cosl_klt_feature* const* dest, int* nNew) {
  KLT_ENTER(h)
  if (!imgs || pitch < (size_t)h->W) return set_error(COSL_E_INVALID, "group_next: bad argument");
  COSL_TRY(upload_images(h, imgs, pitch, cudaMemcpyHostToDevice));
  h->foldAdvance = true;  // the provided set goes to both feature buffers in one launch
  const int rcDet = do_redetect(h);
  h->foldAdvance = false;
  if (rcDet != COSL_OK) return rcDet;
  COSL_TRY(advance(h));
  COSL_TRY(fetch_results(h));
  for (int c = 0; c < h->C; ++c) {
    if (dest && dest[c]) copy_out(h, c, dest[c]);
    if (nNew) nNew[c] = h->h_counters[8 * c + 2];
  }
  return COSL_OK;
}

int cosl_klt_group_next_dev(cosl_klt* h, const uint8_t* const* dimgs, size_t pitch) {
  KLT_ENTER(h)
  if (!dimgs || pitch < (size_t)h->W)
    return set_error(COSL_E_INVALID, "group_next_dev: bad argument");
  COSL_TRY(upload_images(h, dimgs, pitch, cudaMemcpyDeviceToDevice));
  h->foldAdvance = true;  // the provided set goes to both feature buffers in one launch
  const int rcDet = do_redetect(h);
  h->foldAdvance = false;
  if (rcDet != COSL_OK) return rcDet;
  COSL_TRY(advance(h));
  return COSL_OK;
}

int cosl_klt_group_fetch(cosl_klt* h, cosl_klt_feature* const* dest, int* nNew) {
  KLT_ENTER(h)
  COSL_TRY(fetch_results(h));
  for (int c = 0; c < h->C; ++c) {
    if (dest && dest[c]) copy_out(h, c, dest[c]);
    if (nNew) nNew[c] = h->h_counters[8 * c + 2];
  }
  return COSL_OK;
}

// Frame-pipelined ingest (the upload half): submit() enqueues the H2D copy of frame
// n + 1, its kernels and the D2H copy of its feature table and returns; collect() waits for the OLDEST
// submitted frame.  With submit(n+1) called before collect(n), the image upload of frame n + 1 runs on
// the copy stream under the tracker / detector kernels of frame n (it only waits for frame n's front
// pass, which is the last reader of the image buffer).
int cosl_klt_group_submit(cosl_klt* h, const uint8_t* const* imgs, size_t pitch) {
  KLT_ENTER(h)
  if (!imgs || pitch < (size_t)h->W) return set_error(COSL_E_INVALID, "group_submit: bad argument");
  if (h->nInFlight >= 2) return set_error(COSL_E_STATE, "group_submit: two frames in flight, collect one first");
  const int slot = (int)(h->subSeq & 1u);
  if (!h->h_featQ[slot]) {
    COSL_CUDA(cudaMallocHost(&h->h_featQ[slot], sizeof(cosl_klt_feature) * (size_t)h->F * h->C));
    COSL_CUDA(cudaMallocHost(&h->h_cntQ[slot], sizeof(int) * 8 * h->C));
    COSL_CUDA(cudaEventCreateWithFlags(&h->evDone[slot], cudaEventDisableTiming));
  }
  COSL_TRY(upload_images(h, imgs, pitch, cudaMemcpyHostToDevice));
  h->foldAdvance = true;
  const int rcDet = do_redetect(h);
  h->foldAdvance = false;
  if (rcDet != COSL_OK) return rcDet;
  COSL_TRY(advance(h));
  COSL_CUDA(cudaMemcpyAsync(h->h_featQ[slot], h->d_feat, sizeof(cosl_klt_feature) * (size_t)h->F * h->C,
                            cudaMemcpyDeviceToHost, h->stream));
  COSL_CUDA(cudaMemcpyAsync(h->h_cntQ[slot], h->d_counters, sizeof(int) * 8 * h->C, cudaMemcpyDeviceToHost,
                            h->stream));
  COSL_CUDA(cudaEventRecord(h->evDone[slot], h->stream));
  ++h->subSeq;
  ++h->nInFlight;
  return COSL_OK;
}

int cosl_klt_group_collect(cosl_klt* h, cosl_klt_feature* const* dest, int* nNew) {
  KLT_ENTER(h)
  if (h->nInFlight <= 0) return set_error(COSL_E_STATE, "group_collect: nothing submitted");
  const int slot = (int)((h->subSeq - (unsigned)h->nInFlight) & 1u);
  COSL_CUDA(cudaEventSynchronize(h->evDone[slot]));
  --h->nInFlight;
  for (int c = 0; c < h->C; ++c) {
    if (dest && dest[c])
      std::memcpy(dest[c], h->h_featQ[slot] + (size_t)c * h->F, sizeof(cosl_klt_feature) * h->F);
    if (nNew) nNew[c] = h->h_cntQ[slot][8 * c + 2];
  }
  return COSL_OK;
}

int cosl_klt_group_sync(cosl_klt* h) {
  KLT_ENTER(h)
  COSL_CUDA(cudaStreamSynchronize(h->stream));
  return COSL_OK;
}

void* cosl_klt_stream(cosl_klt* h) { return h ? (void*)h->stream : nullptr; }

int cosl_klt_debug_pyramid(cosl_klt* h, int cam, int which, int level, float* out3, int* w,
                           int* ht) {
  KLT_ENTER(h)
  if (cam < 0 || cam >= h->C || level < 0 || level >= h->L)
    return set_error(COSL_E_INVALID, "debug_pyramid: bad argument");
  const int lw = h->lvW[level], lh = h->lvH[level];
  if (w) *w = lw;
  if (ht) *ht = lh;
  if (!out3) return COSL_OK;
  std::vector<float4> tmp((size_t)lw * lh);
  const float4* src =
      h->d_pyr[which ? h->cur : 1 - h->cur] + (size_t)cam * h->pyrStride + h->lvOff[level];
  COSL_CUDA(cudaStreamSynchronize(h->stream));
  COSL_CUDA(cudaMemcpy(tmp.data(), src, sizeof(float4) * tmp.size(), cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < tmp.size(); ++i) {
    out3[3 * i] = tmp[i].x;
    out3[3 * i + 1] = tmp[i].y;
    out3[3 * i + 2] = tmp[i].z;
  }
  return COSL_OK;
}

int cosl_klt_debug_cornerness(cosl_klt* h, int cam, float* out) {
  KLT_ENTER(h)
  if (cam < 0 || cam >= h->C || !out) return set_error(COSL_E_INVALID, "debug_cornerness");
  COSL_CUDA(cudaStreamSynchronize(h->stream));
  COSL_CUDA(cudaMemcpy(out, h->d_corn + (size_t)cam * h->W * h->H,
                       sizeof(float) * (size_t)h->W * h->H, cudaMemcpyDeviceToHost));
  return COSL_OK;
}

int cosl_klt_debug_num_candidates(cosl_klt* h, int cam) {
  if (!h || cam < 0 || cam >= h->C) return -1;
  return h->h_counters[8 * cam + 3];
}

double cosl_klt_algorithmic_bytes(cosl_klt* h) {
  if (!h) return 0;
  // B_frame = C * (25 W H + 4640 F); the 4640 B/feature generalise to
  // n_lvl * 2 * (w+1)^2 * 12 + 32 for other window sizes / level sets
  int nlv = 0;
  for (int level = h->L - 1; level >= 0; level -= h->levelSkip) ++nlv;
  const double w1 = 2 * h->halfWidth + 2;
  double pyr = 1.0;
  for (int l = 0; l < h->L; ++l) pyr += 12.0 / (double)(1 << (2 * l));
  const double perFeat = nlv * 2.0 * w1 * w1 * 12.0 + 32.0;
  return (double)h->C * ((pyr + 8.0) * h->W * h->H + perFeat * h->F);
}

int cosl_klt_profile_enable(cosl_klt* h, int on) {
  KLT_ENTER(h)
  COSL_CUDA(cudaStreamSynchronize(h->stream));
  h->timer.reset();
  h->timer.enabled = on != 0;
  return COSL_OK;
}

const char* cosl_klt_profile_get(cosl_klt* h, int idx, double* ms, int* calls) {
  if (!h || idx < 0 || idx >= h->timer.nsec) return nullptr;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  h->timer.flush();
  if (ms) *ms = h->timer.ms[idx];
  if (calls) *calls = h->timer.calls[idx];
  return h->timer.names[idx];
}

}  // extern "C"
