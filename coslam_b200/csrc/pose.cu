// pose.cu -- per-frame 6-DoF pose refinement (Tukey-IRLS around Levenberg-Marquardt) on the GPU.
// Replaces intraCamEstimate / intraCamWeightedLMProc / intraCamWeightedLMStep
// (slam/SL_IntraCamPose.cpp:626-709 / 475-549 / 259-303).  One 128-thread CTA per camera runs ALL
// IRLS rounds and LM iterations inside one kernel launch (no host round trips): points are spread
// over the threads, the 6x6 normal equations are reduced with warp shuffles + shared memory, and
// every thread redundantly solves the 6x6 system so that control flow stays uniform.
//
// The Jacobian is the reference's forward difference (eps = 1e-8, :43-117) evaluated with non-fused
// fp64 operations (this file is compiled with -fmad=false) so that the differences see the same
// roundings as the oracle.
#include <algorithm>
#include <cmath>
#include <mutex>
#include <vector>

#include "common.cuh"

using namespace coslam;

namespace {

constexpr int POSE_THREADS = 128;

struct PoseCam {
  double K[9], R0[9], t0[3];
  int npts;
  int hasPrev;
  long long off;  // offset of this camera's points in the packed arrays
};

struct PoseOut {
  double R[9], t[3];
  double lambda0, lambda, err0, err, errRW;
  int retTypeLM, npts, nIterLM, nIterRW, ok;
};

struct PoseOut;
struct PoseCam;
struct PoseWorkspace {
  cudaStream_t stream = nullptr;
  PoseCam* d_cams = nullptr;
  double *d_M = nullptr, *d_m = nullptr, *d_p = nullptr, *d_W = nullptr;
  PoseOut* d_out = nullptr;
  int capC = 0;
  long long capPts = 0;
};

struct PoseCfg {
  int maxIterLM, maxIterRW;
  double epsErr, epsParam, epsRW, lambda0, tau;
};

__device__ __forceinline__ void so3_exp(const double w[3], double R[9]) {
  const double th = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  if (th == 0) {
    R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
    return;
  }
  const double a0 = w[0] / th, a1 = w[1] / th, a2 = w[2] / th;
  double s, c;
  sincos(th, &s, &c);
  const double c1 = 1 - c;
  const double a00 = a0 * a0, a01 = a0 * a1, a02 = a0 * a2, a11 = a1 * a1, a12 = a1 * a2,
               a22 = a2 * a2;
  R[0] = -c1 * a11 - c1 * a22 + 1;
  R[1] = c1 * a01 - s * a2;
  R[2] = s * a1 + c1 * a02;
  R[3] = s * a2 + c1 * a01;
  R[4] = -c1 * a00 - c1 * a22 + 1;
  R[5] = c1 * a12 - s * a0;
  R[6] = c1 * a02 - s * a1;
  R[7] = s * a0 + c1 * a12;
  R[8] = -c1 * a00 - c1 * a11 + 1;
}

__device__ __forceinline__ void project(const double K[9], const double R[9], const double t[3],
                                        const double M[3], double m[2]) {
  const double c0 = R[0] * M[0] + R[1] * M[1] + R[2] * M[2] + t[0];
  const double c1 = R[3] * M[0] + R[4] * M[1] + R[5] * M[2] + t[1];
  const double c2 = R[6] * M[0] + R[7] * M[1] + R[8] * M[2] + t[2];
  const double u = K[0] * c0 + K[1] * c1 + K[2] * c2;
  const double v = K[3] * c0 + K[4] * c1 + K[5] * c2;
  const double w = K[6] * c0 + K[7] * c1 + K[8] * c2;
  m[0] = u / w;
  m[1] = v / w;
}

__device__ __forceinline__ void mat33(const double A[9], const double B[9], double C[9]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

__device__ __forceinline__ double tukey(double e, double tau) {
  if (e >= tau) return 0;
  e /= tau;
  e = 1 - e * e;
  return e * e;
}

// block-wide sum of NV doubles; result broadcast to every thread (bitwise identical everywhere)
template <int NV>
__device__ __forceinline__ void block_sum(double* v, double* s_red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double x = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) s_red[wid * NV + k] = x;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double x = s_red[k];
#pragma unroll
    for (int w = 1; w < POSE_THREADS / 32; ++w) x += s_red[w * NV + k];
    v[k] = x;
  }
  __syncthreads();
}

// inverse by Gauss-Jordan elimination with partial pivoting (same algorithm as the oracle)
__device__ bool inv6(const double A[36], double Ainv[36]) {
  double a[6][12];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      a[i][j] = A[6 * i + j];
      a[i][6 + j] = (i == j) ? 1.0 : 0.0;
    }
  for (int c = 0; c < 6; ++c) {
    int p = c;
    double best = fabs(a[c][c]);
    for (int r = c + 1; r < 6; ++r)
      if (fabs(a[r][c]) > best) {
        best = fabs(a[r][c]);
        p = r;
      }
    if (best == 0.0) return false;
    if (p != c)
      for (int j = 0; j < 12; ++j) {
        const double tmp = a[c][j];
        a[c][j] = a[p][j];
        a[p][j] = tmp;
      }
    const double d = 1.0 / a[c][c];
    for (int j = 0; j < 12; ++j) a[c][j] *= d;
    for (int r = 0; r < 6; ++r) {
      if (r == c) continue;
      const double f = a[r][c];
      for (int j = 0; j < 12; ++j) a[r][j] -= f * a[c][j];
    }
  }
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) Ainv[6 * i + j] = a[i][6 + j];
  return true;
}

__device__ double weighted_cost(const double K[9], const double R[9], const double t[3], int n,
                                const double* Ws, const double* Ms, const double* ms,
                                double* s_red) {
  double err[1] = {0};
  for (int i = threadIdx.x; i < n; i += POSE_THREADS) {
    double rm[2];
    project(K, R, t, Ms + 3 * i, rm);
    const double dx = ms[2 * i] - rm[0], dy = ms[2 * i + 1] - rm[1];
    err[0] += (dx * dx + dy * dy) * Ws[i];
  }
  block_sum<1>(err, s_red);
  return err[0];
}

// intraCamWeightedLMStep (:259-303)
__device__ void weighted_step(const double K[9], const double R[9], const double t[3], int n,
                              const double* Ws, const double* Ms, const double* ms,
                              double param[6], double lambda, const double dRk[3][9],
                              double* s_red) {
  const double eps = 1e-8;
  double acc[27];  // 21 upper-triangle entries of J^T J + 6 of J^T r
#pragma unroll
  for (int k = 0; k < 27; ++k) acc[k] = 0;
  for (int i = threadIdx.x; i < n; i += POSE_THREADS) {
    const double* M = Ms + 3 * i;
    double rm0[2], rm[2], J[12];
    project(K, R, t, M, rm0);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double R1[9];
      mat33(R, dRk[k], R1);
      project(K, R1, t, M, rm);
      J[k] = (rm[0] - rm0[0]) / eps;
      J[6 + k] = (rm[1] - rm0[1]) / eps;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double t1[3] = {t[0], t[1], t[2]};
      t1[k] = t[k] + eps;
      project(K, R, t1, M, rm);
      J[3 + k] = (rm[0] - rm0[0]) / eps;
      J[9 + k] = (rm[1] - rm0[1]) / eps;
    }
    const double w = Ws[i];
#pragma unroll
    for (int k = 0; k < 12; ++k) J[k] = w * J[k];
    const double r0 = (-rm0[0] + ms[2 * i]) * w;
    const double r1 = (-rm0[1] + ms[2 * i + 1]) * w;
    int q = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
#pragma unroll
      for (int b = a; b < 6; ++b) acc[q++] += J[a] * J[b] + J[6 + a] * J[6 + b];
    }
#pragma unroll
    for (int a = 0; a < 6; ++a) acc[21 + a] += J[a] * r0 + J[6 + a] * r1;
  }
  block_sum<27>(acc, s_red);
  double sA[36], inv[36];
  int q = 0;
  for (int a = 0; a < 6; ++a)
    for (int b = a; b < 6; ++b) {
      sA[6 * a + b] = acc[q];
      sA[6 * b + a] = acc[q];
      ++q;
    }
  for (int a = 0; a < 6; ++a) sA[7 * a] += lambda;
  if (!inv6(sA, inv)) {
    for (int a = 0; a < 6; ++a) param[a] = 0;
    return;
  }
  for (int a = 0; a < 6; ++a) {
    double s = 0;
    for (int b = 0; b < 6; ++b) s += inv[6 * a + b] * acc[21 + b];
    param[a] = s;
  }
}

__global__ void __launch_bounds__(POSE_THREADS)
pose_intracam_kernel(const PoseCam* __restrict__ cams, const double* __restrict__ MsAll,
                     const double* __restrict__ msAll, const double* __restrict__ prevAll,
                     double* __restrict__ WsAll, PoseCfg cfg, PoseOut* __restrict__ outs) {
  __shared__ double s_red[(POSE_THREADS / 32) * 27];
  const PoseCam cam = cams[blockIdx.x];
  const int n = cam.npts;
  const double* Ms = MsAll + 3 * cam.off;
  const double* ms = msAll + 2 * cam.off;
  double* Ws = WsAll + cam.off;
  const double tau = cfg.tau;
  double K[9], R[9], t[3], R_opt[9], t_opt[3], R_tmp[9], t_tmp[3], Rlm[9], tlm[3];
  for (int k = 0; k < 9; ++k) {
    K[k] = cam.K[k];
    R[k] = cam.R0[k];
    R_opt[k] = cam.R0[k];
  }
  for (int k = 0; k < 3; ++k) {
    t[k] = cam.t0[k];
    t_opt[k] = cam.t0[k];
  }
  // rotation perturbations exp(eps e_k) of the numeric Jacobian (:56-84)
  double dRk[3][9];
  for (int k = 0; k < 3; ++k) {
    double w[3] = {0, 0, 0};
    w[k] = 1e-8;
    so3_exp(w, dRk[k]);
  }
  for (int i = threadIdx.x; i < n; i += POSE_THREADS)
    Ws[i] = cam.hasPrev ? tukey(fabs(prevAll[cam.off + i]), tau) : 1.0;
  __syncthreads();

  double lambda0 = cfg.lambda0, lambda = cfg.lambda0, err0 = 0, errLM = 0, errRW = -1;
  int retType = 1, nIterLM = 0, k = 0;
  bool ret = true;
  for (; k < cfg.maxIterRW; ++k) {
    // ---- intraCamWeightedLMProc (:475-549) from (R, t) ----
    double param[6];
    lambda = lambda0;
    err0 = weighted_cost(K, R, t, n, Ws, Ms, ms, s_red);
    errLM = err0;
    for (int q = 0; q < 9; ++q) {
      Rlm[q] = R[q];
      R_tmp[q] = R[q];
    }
    for (int q = 0; q < 3; ++q) {
      tlm[q] = t[q];
      t_tmp[q] = t[q];
    }
    retType = 1;
    int i = 0;
    double err = err0;
    for (; i < cfg.maxIterLM; ++i) {
      weighted_step(K, Rlm, tlm, n, Ws, Ms, ms, param, lambda, dRk, s_red);
      {  // intraCamUpdatePose (:367-380)
        double dR[9];
        so3_exp(param, dR);
        mat33(Rlm, dR, R_opt);
        t_opt[0] = tlm[0] + param[3];
        t_opt[1] = tlm[1] + param[4];
        t_opt[2] = tlm[2] + param[5];
      }
      double p2 = 0;
      for (int q = 0; q < 6; ++q) p2 += param[q] * param[q];
      if (p2 < cfg.epsParam) {
        for (int q = 0; q < 9; ++q) Rlm[q] = R_opt[q];
        for (int q = 0; q < 3; ++q) tlm[q] = t_opt[q];
        retType = 0;
        break;
      }
      err = weighted_cost(K, R_opt, t_opt, n, Ws, Ms, ms, s_red);
      if (fabs(err - errLM) < cfg.epsErr) {
        retType = 0;
        break;
      }
      if (err <= errLM) {
        for (int q = 0; q < 9; ++q) {
          Rlm[q] = R_opt[q];
          R_tmp[q] = R_opt[q];
        }
        for (int q = 0; q < 3; ++q) {
          tlm[q] = t_opt[q];
          t_tmp[q] = t_opt[q];
        }
        errLM = err;
        lambda /= 10;
      } else {
        lambda *= 10;
        if (lambda > 1e+18) {
          retType = -1;
          break;
        }
      }
    }
    if (retType == -1) {
      for (int q = 0; q < 9; ++q) R_opt[q] = R_tmp[q];
      for (int q = 0; q < 3; ++q) t_opt[q] = t_tmp[q];
    }
    errLM = err;
    nIterLM = i;
    if (retType < 0) {
      ret = false;
      break;
    }
    // ---- IRLS bookkeeping (:664-702) ----
    lambda0 = lambda;
    if (errRW < 0)
      errRW = errLM;
    else {
      if (fabs(errLM - errRW) < cfg.epsRW) {
        ret = true;
        break;
      }
      errRW = errLM;
    }
    for (int q = 0; q < 9; ++q) R[q] = R_opt[q];
    for (int q = 0; q < 3; ++q) t[q] = t_opt[q];
    for (int p = threadIdx.x; p < n; p += POSE_THREADS) {
      double rm[2];
      project(K, R, t, Ms + 3 * p, rm);
      const double dx = rm[0] - ms[2 * p], dy = rm[1] - ms[2 * p + 1];
      Ws[p] = tukey(sqrt(dx * dx + dy * dy), tau);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    PoseOut o;
    for (int q = 0; q < 9; ++q) o.R[q] = R_opt[q];
    for (int q = 0; q < 3; ++q) o.t[q] = t_opt[q];
    o.lambda0 = lambda0;
    o.lambda = lambda;
    o.err0 = err0;
    o.err = errLM;
    o.errRW = errRW;
    o.retTypeLM = retType;
    o.npts = n;
    o.nIterLM = nIterLM;
    o.nIterRW = k;
    o.ok = ret ? 1 : 0;
    outs[blockIdx.x] = o;
  }
}

}  // namespace

extern "C" {

void cosl_pose_opt_default(cosl_pose_opt* o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  o->maxIterLM = 100;
  o->maxIterRW = 5;
  o->epsErrorChangeLM = 1e-7;
  o->epsParamChangeLM = 1e-6;
  o->epsErrorChangeRW = 1e-6;
  o->lambda0 = 1e-3;
}

int cosl_pose_intracam_batch(int C, const double* K9, const double* R0, const double* t0,
                             const int* npts, const double* const* Ms, const double* const* ms,
                             const double* const* prevErrs, double tau, double* R_opt,
                             double* t_opt, cosl_pose_opt* opts, int* ok, int device) {
  if (C < 1 || !K9 || !R0 || !t0 || !npts || !Ms || !ms || !R_opt || !t_opt || !opts)
    return set_error(COSL_E_INVALID, "cosl_pose_intracam_batch: null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    return set_error(COSL_E_CUDA, "CUDA device %d not available", device);
  COSL_CUDA(cudaSetDevice(device));
  std::vector<PoseCam> cams(C);
  long long tot = 0;
  for (int c = 0; c < C; ++c) {
    if (npts[c] < 0 || (npts[c] && (!Ms[c] || !ms[c])))
      return set_error(COSL_E_INVALID, "cosl_pose_intracam_batch: bad camera %d", c);
    std::memcpy(cams[c].K, K9 + 9 * c, sizeof(double) * 9);
    std::memcpy(cams[c].R0, R0 + 9 * c, sizeof(double) * 9);
    std::memcpy(cams[c].t0, t0 + 3 * c, sizeof(double) * 3);
    cams[c].npts = npts[c];
    cams[c].hasPrev = (prevErrs && prevErrs[c]) ? 1 : 0;
    cams[c].off = tot;
    tot += npts[c];
  }
  const long long totA = tot > 0 ? tot : 1;
  std::vector<double> hM(3 * totA), hm(2 * totA), hp(totA, 0.0);
  for (int c = 0; c < C; ++c) {
    if (!npts[c]) continue;
    std::memcpy(&hM[3 * cams[c].off], Ms[c], sizeof(double) * 3 * npts[c]);
    std::memcpy(&hm[2 * cams[c].off], ms[c], sizeof(double) * 2 * npts[c]);
    if (cams[c].hasPrev) std::memcpy(&hp[cams[c].off], prevErrs[c], sizeof(double) * npts[c]);
  }
  // all opts share the solver settings of opts[0] except lambda0, which is per camera in the
  // reference (carried across calls by the caller); kernel cfg takes opts[0], lambda0 is uniform
  PoseCfg cfg;
  cfg.maxIterLM = opts[0].maxIterLM;
  cfg.maxIterRW = opts[0].maxIterRW;
  cfg.epsErr = opts[0].epsErrorChangeLM;
  cfg.epsParam = opts[0].epsParamChangeLM;
  cfg.epsRW = opts[0].epsErrorChangeRW;
  cfg.lambda0 = opts[0].lambda0;
  cfg.tau = tau;
  for (int c = 1; c < C; ++c)
    if (opts[c].lambda0 != opts[0].lambda0 || opts[c].maxIterLM != opts[0].maxIterLM ||
        opts[c].maxIterRW != opts[0].maxIterRW || opts[c].epsErrorChangeLM != opts[0].epsErrorChangeLM ||
        opts[c].epsParamChangeLM != opts[0].epsParamChangeLM ||
        opts[c].epsErrorChangeRW != opts[0].epsErrorChangeRW)
      return set_error(COSL_E_INVALID, "batched pose solve needs identical options per camera");
  // per-device workspace, reused across calls: the solve is latency-bound, a cudaMalloc/cudaFree
  // pair per call would cost more than the kernel
  static std::mutex mtx;
  static PoseWorkspace ws[64];
  std::lock_guard<std::mutex> lock(mtx);
  PoseWorkspace& W = ws[device & 63];
  int rc = COSL_OK;
  std::vector<PoseOut> outs(C);
#define POSE_CK(expr)                                                               \
  if (rc == COSL_OK) {                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) rc = set_error(COSL_E_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  }
  if (!W.stream) POSE_CK(cudaStreamCreateWithFlags(&W.stream, cudaStreamNonBlocking));
  if (rc == COSL_OK && (W.capC < C || W.capPts < totA)) {
    cudaFree(W.d_cams);
    cudaFree(W.d_M);
    cudaFree(W.d_m);
    cudaFree(W.d_p);
    cudaFree(W.d_W);
    cudaFree(W.d_out);
    W = PoseWorkspace{W.stream};
    const int capC = std::max(C, 16);
    const long long capP = std::max<long long>(totA, 4096);
    POSE_CK(cudaMalloc(&W.d_cams, sizeof(PoseCam) * capC));
    POSE_CK(cudaMalloc(&W.d_M, sizeof(double) * 3 * capP));
    POSE_CK(cudaMalloc(&W.d_m, sizeof(double) * 2 * capP));
    POSE_CK(cudaMalloc(&W.d_p, sizeof(double) * capP));
    POSE_CK(cudaMalloc(&W.d_W, sizeof(double) * capP));
    POSE_CK(cudaMalloc(&W.d_out, sizeof(PoseOut) * capC));
    if (rc == COSL_OK) {
      W.capC = capC;
      W.capPts = capP;
    }
  }
  cudaStream_t st = W.stream;
  POSE_CK(cudaMemcpyAsync(W.d_cams, cams.data(), sizeof(PoseCam) * C, cudaMemcpyHostToDevice, st));
  POSE_CK(cudaMemcpyAsync(W.d_M, hM.data(), sizeof(double) * 3 * totA, cudaMemcpyHostToDevice, st));
  POSE_CK(cudaMemcpyAsync(W.d_m, hm.data(), sizeof(double) * 2 * totA, cudaMemcpyHostToDevice, st));
  POSE_CK(cudaMemcpyAsync(W.d_p, hp.data(), sizeof(double) * totA, cudaMemcpyHostToDevice, st));
  if (rc == COSL_OK) {
    COSL_LAUNCH(pose_intracam_kernel, C, POSE_THREADS, 0, st, W.d_cams, W.d_M, W.d_m, W.d_p, W.d_W,
                cfg, W.d_out);
  }
  POSE_CK(cudaGetLastError());
  POSE_CK(cudaMemcpyAsync(outs.data(), W.d_out, sizeof(PoseOut) * C, cudaMemcpyDeviceToHost, st));
  POSE_CK(cudaStreamSynchronize(st));
#undef POSE_CK
  if (rc != COSL_OK) return rc;
  for (int c = 0; c < C; ++c) {
    std::memcpy(R_opt + 9 * c, outs[c].R, sizeof(double) * 9);
    std::memcpy(t_opt + 3 * c, outs[c].t, sizeof(double) * 3);
    opts[c].lambda0 = outs[c].lambda0;
    opts[c].lambda = outs[c].lambda;
    opts[c].err0 = outs[c].err0;
    opts[c].err = outs[c].err;
    opts[c].errRW = outs[c].errRW;
    opts[c].retTypeLM = outs[c].retTypeLM;
    opts[c].npts = outs[c].npts;
    opts[c].nIterLM = outs[c].nIterLM;
    opts[c].nIterRW = outs[c].nIterRW;
    if (ok) ok[c] = outs[c].ok;
  }
  return COSL_OK;
}

int cosl_pose_intracam(const double K[9], const double R0[9], const double t0[3], int npts,
                       const double* prevErrs, const double* Ms, const double* ms, double tau,
                       double R_opt[9], double t_opt[3], cosl_pose_opt* opt, int* ok) {
  const double* Msp[1] = {Ms};
  const double* msp[1] = {ms};
  const double* pep[1] = {prevErrs};
  int device = 0;
  cudaGetDevice(&device);
  return cosl_pose_intracam_batch(1, K, R0, t0, &npts, Msp, msp, prevErrs ? pep : nullptr, tau,
                                  R_opt, t_opt, opt, ok, device);
}

}  // extern "C"
