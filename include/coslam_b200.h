/*
 * coslam_b200.h -- C-ABI of the B200-native CoSLAM hot paths (KLT tracker, 6-DoF pose solver,
 * multi-camera bundle adjustment).  Plain pointers and sizes only; no C++/torch types.
 *
 * Every entry point returns an int status (COSL_OK == 0, negative == error) and never throws.
 * All reference citations are relative to the upstream tree (danping/CoSLAM, src/...).
 *
 * Seams replaced:
 *   KLT engine  : V3D_GPU::KLT_SequenceTracker            tracking/CGKLT/v3d_gpuklt.h:202-262
 *   Pose solver : intraCamEstimate                        slam/SL_IntraCamPose.h:92-95
 *   BA solver   : bundleAdjustRobust (LibVisualSLAM)      call site app/SL_CoSLAMRobustBA.cpp:174
 *                 sba_motstr_levmar_x  (sba-1.6)           call site app/SL_CoSLAMBA.cpp:360-363
 */
#ifndef COSLAM_B200_H_
#define COSLAM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- status codes */
enum {
  COSL_OK = 0,
  COSL_E_INVALID = -1,   /* bad argument                                          */
  COSL_E_CUDA = -2,      /* CUDA runtime error (see cosl_last_error())            */
  COSL_E_NOMEM = -3,     /* host or device allocation failed                      */
  COSL_E_NUMERIC = -4,   /* solver break-down (non-SPD reduced system, NaN, ...)  */
  COSL_E_NCCL = -5,      /* NCCL not loadable / collective failed                 */
  COSL_E_STATE = -6      /* call sequence error (e.g. track before detect)        */
};

/* Thread-local, human readable description of the last error on the calling thread. */
const char* cosl_last_error(void);
/* Library version string "coslam_b200 x.y (sm_100a)". */
const char* cosl_version(void);
/* Number of this library's kernel launches issued so far by the calling process (for bench.py). */
uint64_t cosl_kernel_launch_count(void);

/* ================================================================ KLT tracker */

/* Mirrors V3D_GPU::KLT_SequenceTrackerConfig (v3d_gpuklt.h:180-199), same field order and
 * defaults (use cosl_klt_config_default()).  `trackWithGain` is an int instead of bool.
 * `compat` is an extension: bit 0 (COSL_KLT_COMPAT_ITER5) reproduces the reference's macro-name
 * bug that makes the 2x2 tracker run 5 iterations per level whatever nIterations says
 * (klt_tracker.cg:16-18 vs v3d_gpuklt.cpp:108). */
typedef struct cosl_klt_config {
  int nIterations;            /* 12  */
  int nLevels;                /* 3   (CoSLAM live: 6, app/SL_GlobParam.cpp:30)   */
  int levelSkip;              /* 2   (<=0 means nLevels-1, v3d_gpuklt.h:14)      */
  int windowWidth;            /* 5   (CoSLAM live: 6 -> half width 3)            */
  float trackBorderMargin;    /* 4.0 px */
  float convergenceThreshold; /* 0.1 px  (live 1.0)   */
  float SSD_Threshold;        /* 5000    (live 20000) */
  int trackWithGain;          /* 0       (live 1)     */
  int minDistance;            /* 8  */
  float minCornerness;        /* 1000    (live 3000)  */
  float detectBorderMargin;   /* 4.0 -- dead in the reference: the detector always uses 10 px
                                 (v3d_gpuklt.h:113-114); honoured only via cosl_klt_set_margin */
  int compat;                 /* extension, see above; default COSL_KLT_COMPAT_ITER5 */
} cosl_klt_config;

/* bit 1 (COSL_KLT_PASS_KERNELS, diagnostic): run the gain tracker as one kernel launch per
 * (level, iteration) pass like the reference's draw calls, instead of the single persistent
 * kernel; results are bit-identical, only slower. */
enum { COSL_KLT_COMPAT_ITER5 = 1, COSL_KLT_PASS_KERNELS = 2, COSL_KLT_COMPAT_FEED_STRIDE2 = 4,
       COSL_KLT_COMPAT_HISTOPYR = 8 };
/* bit 2 (COSL_KLT_COMPAT_FEED_STRIDE2, parity runs): feedExternFeaturePoints' proximity test reads
 * the fed points as featPts[2k], featPts[2k+1] although the array holds 3 floats per point
 * (v3d_gpuklt.cpp:826-827 vs :841-842).  The default uses stride 3 in both loops (the evident
 * intent); with this bit the kill test reproduces the reference's stride-2 reads, so the same
 * tracks die and the same trackIds come back as from the reference for the same input.
 * bit 3 (COSL_KLT_COMPAT_HISTOPYR, parity runs): the reference's candidate list instead of "all
 * non-max survivors, strongest first": the corners are read back in HistoPyramid extraction order
 * (Morton order of the pixel, x in the even bits; klt_detector_traverse_histpyr.cg:33-50) and
 * TRUNCATED to pointListWidth*pointListHeight entries before anything is ranked
 * (v3d_gpuklt.cpp:660-661, 755-757), a last odd image row / column is never examined (:519-523),
 * and the list is ordered by cornerness only when it exceeds the free slots (:662-665, 761-768),
 * otherwise slots are filled in extraction order.  With the bit set the SET of new corners equals
 * the reference's; when the list was cut by cornerness their ORDER is strongest-first, whereas the
 * reference's std::nth_element leaves an implementation-defined order. */

/* Mirrors V3D_GPU::KLT_TrackedFeature (v3d_gpuklt.h:166-176): 20 bytes. */
typedef struct cosl_klt_feature {
  int status;   /* 0 tracked from previous frame, 1 newly created, -1 invalid */
  float pos[2]; /* normalised [0,1] image coordinates, texel centres at (i+0.5)/W */
  float gain;
  int fed;      /* >=0: index of the externally fed point */
} cosl_klt_feature;

typedef struct cosl_klt cosl_klt; /* opaque: one tracker group = C cameras of equal geometry */

void cosl_klt_config_default(cosl_klt_config* cfg);

/* KLT_SequenceTracker ctor + allocate(w,h,nLevels,featW,featH[,plW,plH]) (v3d_gpuklt.h:212-215,
 * v3d_gpuklt.cpp:592-624).  plW/plH <= 0 selects the reference default 2*featW x 2*featH
 * (capacity of the candidate list).  `device` is the CUDA ordinal.  One camera. */
int cosl_klt_create(const cosl_klt_config* cfg, int width, int height, int nLevels, int featW,
                    int featH, int plW, int plH, int device, cosl_klt** out);
/* Extension: C cameras of identical geometry that are always advanced together; every kernel is
 * launched once for the whole group (camera = grid.z).  Single-camera calls below address
 * camera 0 of a group. */
int cosl_klt_group_create(const cosl_klt_config* cfg, int nCams, int width, int height,
                          int nLevels, int featW, int featH, int plW, int plH, int device,
                          cosl_klt** out);
/* KLT_SequenceTracker::deallocate (v3d_gpuklt.cpp:626-648). */
int cosl_klt_destroy(cosl_klt* h);

/* KLT_SequenceTracker::detect(image, n, dest) and detect(image, n, dest, nPresent, present)
 * (v3d_gpuklt.cpp:651-738).  `img` is a HOST pointer, rows `pitch` bytes apart.  present3 =
 * nPresent float triples (x, y, ignored), may be NULL when nPresent == 0.  dest has featW*featH
 * entries. */
int cosl_klt_detect(cosl_klt* h, const uint8_t* img, size_t pitch, int nPresent,
                    const float* present3, cosl_klt_feature* dest, int* nDetected);
/* KLT_SequenceTracker::redetect (v3d_gpuklt.cpp:740-805): track + detect around live tracks +
 * refill dead slots in slot order with the strongest new corners. */
int cosl_klt_redetect(cosl_klt* h, const uint8_t* img, size_t pitch, cosl_klt_feature* dest,
                      int* nNewFeatures);
/* KLT_SequenceTracker::track (v3d_gpuklt.cpp:857-889). */
int cosl_klt_track(cosl_klt* h, const uint8_t* img, size_t pitch, cosl_klt_feature* dest,
                   int* nPresent);
/* KLT_SequenceTracker::feedExternFeaturePoints (v3d_gpuklt.cpp:808-855).  pts3 = npts float
 * triples (stride 3 everywhere; the reference reads stride 2 in its distance test -- a bug the
 * survey documents).  trackIds receives the slot of each fed point (npts entries). */
int cosl_klt_feed(cosl_klt* h, int npts, const float* pts3, int* trackIds, int* nFed);
/* KLT_SequenceTracker::advanceFrame (v3d_gpuklt.h:252-259). */
int cosl_klt_advance(cosl_klt* h);
/* setBorderMargin / setConvergenceThreshold / setSSD_Threshold (v3d_gpuklt.h:217-239). */
int cosl_klt_set_margin(cosl_klt* h, float margin);
int cosl_klt_set_conv(cosl_klt* h, float thr);
int cosl_klt_set_ssd(cosl_klt* h, float thr);

/* Extension (batched GPUKLT::next, tracking/GPUKLT.cpp:144-161): redetect + advanceFrame for all
 * cameras of the group.  imgs[c] are HOST pointers (pinned memory recommended); dest[c] has
 * featW*featH entries; nNew[c] as in redetect.  One H2D copy per camera, one D2H copy of the
 * whole feature table, one stream synchronise. */
int cosl_klt_group_next(cosl_klt* h, const uint8_t* const* imgs, size_t pitch,
                        cosl_klt_feature* const* dest, int* nNew);
/* Same, but images already resident in device memory (dimgs[c] = device pointers, tightly packed
 * rows of `pitch` bytes) and the feature table left on the device: no host transfer, no
 * synchronise; used for the kernel-only throughput number.  Results can be fetched later with
 * cosl_klt_group_fetch. */
int cosl_klt_group_next_dev(cosl_klt* h, const uint8_t* const* dimgs, size_t pitch);
int cosl_klt_group_first(cosl_klt* h, const uint8_t* const* imgs, size_t pitch,
                         cosl_klt_feature* const* dest, int* nDetected);
int cosl_klt_group_fetch(cosl_klt* h, cosl_klt_feature* const* dest, int* nNew);
int cosl_klt_group_sync(cosl_klt* h);
/* Frame-pipelined form of cosl_klt_group_next (ingest half of SURVEY.md 8f-4; what
 * SingleSLAM::grabReadFrame + GPUKLT::next become when the next frame is already decoded,
 * app/SL_SingleSLAM.cpp:317-327, tracking/GPUKLT.cpp:144-161): submit() enqueues upload, kernels and
 * the read-back of one frame and returns at once; collect() blocks for the OLDEST submitted frame
 * and hands out its feature tables.  At most two frames may be in flight; calling submit(n+1)
 * before collect(n) overlaps the upload of frame n+1 with the kernels of frame n.  The image
 * buffers of a submitted frame must stay untouched until that frame has been collected. */
int cosl_klt_group_submit(cosl_klt* h, const uint8_t* const* imgs, size_t pitch);
int cosl_klt_group_collect(cosl_klt* h, cosl_klt_feature* const* dest, int* nNew);
/* CUDA stream (cudaStream_t as void*) the group launches on; lets callers time with events. */
void* cosl_klt_stream(cosl_klt* h);

/* Test/inspection hooks (not part of the reference surface). */
/* Copy pyramid level `level` of camera `cam` to host as interleaved (I, Ix, Iy) floats.
 * which: 0 = previous frame (pyr0), 1 = current frame (pyr1). */
int cosl_klt_debug_pyramid(cosl_klt* h, int cam, int which, int level, float* out3, int* w,
                           int* ht);
/* Copy the cornerness map after non-max suppression input stage (W*H floats) of camera cam. */
int cosl_klt_debug_cornerness(cosl_klt* h, int cam, float* out);
/* Raw number of detector candidates of the last detect/redetect fetched to the host. */
int cosl_klt_debug_num_candidates(cosl_klt* h, int cam);
/* Per-kernel-class device time (CUDA events on the group's stream).  enable(1) resets and starts
 * accumulating; get(idx) returns the class name ("klt_pyramid", "klt_track", "klt_detect",
 * "klt_select") or NULL past the end. */
int cosl_klt_profile_enable(cosl_klt* h, int on);
const char* cosl_klt_profile_get(cosl_klt* h, int idx, double* ms, int* calls);
/* Algorithmic HBM bytes of one group_next call (SURVEY.md 8d contract): C*(25*W*H + 4640*F). */
double cosl_klt_algorithmic_bytes(cosl_klt* h);

/* ================================================================ pose solver */

/* Mirrors IntraCamPoseOption (slam/SL_IntraCamPose.h:19-57). */
typedef struct cosl_pose_opt {
  int maxIterLM;           /* 100  */
  int maxIterRW;           /* 5    */
  double epsErrorChangeLM; /* 1e-7 */
  double epsParamChangeLM; /* 1e-6 */
  double epsErrorChangeRW; /* 1e-6 */
  int verboseLM, verboseRW;
  double lambda0;          /* 1e-3 */
  double lambda;
  double err0, err, errRW;
  int retTypeLM;
  int npts;
  int nIterLM, nIterRW;
} cosl_pose_opt;

void cosl_pose_opt_default(cosl_pose_opt* opt);

/* intraCamEstimate (slam/SL_IntraCamPose.cpp:626-709), exact reference argument list; returns
 * 1 (true) / 0 (false = LM failure) through *ok.  Arrays are HOST pointers. */
int cosl_pose_intracam(const double K[9], const double R0[9], const double t0[3], int npts,
                       const double* prevErrs /* nullable */, const double* Ms, const double* ms,
                       double tau, double R_opt[9], double t_opt[3], cosl_pose_opt* opt, int* ok);
/* Batched over C cameras in one launch (one 128-thread CTA per camera; the option
 * fields maxIterLM, maxIterRW, lambda0 and the three eps* must be equal for all cameras, else
 * COSL_E_INVALID).  K9/R0/t0/R_opt/t_opt are C
 * consecutive blocks; Ms[c]/ms[c]/prevErrs[c] per-camera HOST arrays (prevErrs or prevErrs[c]
 * may be NULL); opts has C entries; ok has C entries. */
int cosl_pose_intracam_batch(int C, const double* K9, const double* R0, const double* t0,
                             const int* npts, const double* const* Ms, const double* const* ms,
                             const double* const* prevErrs, double tau, double* R_opt,
                             double* t_opt, cosl_pose_opt* opts, int* ok, int device);

/* ================================================================ bundle adjustment */

/* Flat form of the arguments of
 *   bundleAdjustRobust(nCamsCon, Ks, Rs, Ts, nPtsCon, pts, meas, maxErr, maxIter, nInnerMaxIter)
 * (LibVisualSLAM geometry/SL_BundleAdjust.h; call app/SL_CoSLAMRobustBA.cpp:174-175).
 * Observations are CSR by point: point i owns obs [ptr[i], ptr[i+1]), sorted by camera index,
 * exactly the order of vector<vector<Meas2D>> built at app/SL_CoSLAMRobustBA.cpp:109-165. */
typedef struct cosl_ba_problem {
  int m;              /* cameras (poses)                                   */
  int n;              /* points                                            */
  int64_t nobs;       /* observations                                      */
  int m_con;          /* first m_con cameras fixed (nCamsCon)              */
  int n_con;          /* first n_con points fixed  (nPtsCon)               */
  double* K;          /* [m][9] row-major intrinsics, read only            */
  double* R;          /* [m][9] row-major, x_cam = R X + t; updated in place */
  double* t;          /* [m][3] updated in place                           */
  double* X;          /* [n][3] updated in place                           */
  const int64_t* ptr; /* [n+1]                                             */
  const int32_t* cam; /* [nobs] Meas2D::viewId                             */
  const double* xy;   /* [nobs][2] Meas2D::x,y                             */
  uint8_t* outlier;   /* [nobs] out: Meas2D::outlier (may be NULL)         */
} cosl_ba_problem;

typedef struct cosl_ba_options {
  double max_err;     /* maxErr: Tukey scale / outlier threshold in px; <= 0 disables robust
                         weighting (plain sba_motstr_levmar_x behaviour)              */
  int outer_iters;    /* maxIter: robust re-weighting rounds                            */
  int inner_iters;    /* nInnerMaxIter: LM iterations (linear solves) per round         */
  double opts[5];     /* sba opts: {tau for mu0, eps1 |J^T e|_inf, eps2 |dp|, eps3 |e|^2,
                         eps4 relative reduction}; BundleRTS uses {1e-7,1e-12,1e-12,0,1e-16}
                         (app/SL_CoSLAMBA.cpp:323-328)                                  */
  int device;         /* CUDA ordinal of this rank                                      */
  int verbose;
} cosl_ba_options;

/* info[0..9] as sba: {|e0|^2, |e|^2, |J^T e|_inf, |dp|^2, mu/max diag, #iterations, stop reason,
 * #fevals, #jevals, #linear systems}; [10] = total LM trials over all rounds, [11] = seconds spent
 * in the device solve loop (excl. upload), [12] = final weighted cost, [13] = #outliers. */
#define COSL_BA_INFOSZ 16

void cosl_ba_options_default(cosl_ba_options* o);

/* Single-GPU solve, HOST arrays in/out (the drop-in for bundleAdjustRobust). */
int cosl_ba_solve(cosl_ba_problem* prob, const cosl_ba_options* opt, double info[COSL_BA_INFOSZ]);
/* The same call on n_gpus devices of THIS process (devices = CUDA ordinals, NULL = 0..n_gpus-1):
 * one host thread per device, map points sharded over them, one NCCL all-reduce of the reduced
 * camera system per LM trial.  n_gpus <= 1 is cosl_ba_solve.  opt->device is
 * ignored when devices is given. */
int cosl_ba_solve_multi(cosl_ba_problem* prob, const cosl_ba_options* opt, int n_gpus,
                        const int* devices, double info[COSL_BA_INFOSZ]);

/* sba_motstr_levmar_x specialised to the KRTS projection (app/SL_CoSLAMBA.cpp:360-363).  Same
 * argument meaning as sba-1.6: n points, ncon fixed, m cameras, mcon fixed, vmask[n*m],
 * p = [m*cnp camera params | n*pnp points] with cnp=11 (fx,cx,cy,ar,s, q1,q2,q3, t1,t2,t3),
 * x = measurements in point-major order, rot0params = m unit quaternions (sbaGlobs::rot0params).
 * func/fjac pointers of the original are not part of this signature (the projection is fixed).
 * Returns number of iterations (>= 0) or a negative COSL_E_*. */
int cosl_sba_motstr_levmar_x(int n, int ncon, int m, int mcon, const char* vmask, double* p,
                             int cnp, int pnp, const double* x, int mnp,
                             const double* rot0params, int itmax, int verbose,
                             const double opts[5], double info[10], int device);

/* ---- multi-GPU (points sharded across ranks, cameras replicated; one NCCL all-reduce of the
 * packed reduced camera system per LM trial + one small all-reduce of scalars). ---- */
typedef struct cosl_ba_comm cosl_ba_comm; /* opaque */
/* Fill 128 bytes with an ncclUniqueId (rank 0 calls this, then broadcasts the bytes out of band,
 * e.g. through torch.distributed). */
int cosl_nccl_unique_id(uint8_t id[128]);
int cosl_ba_comm_create(const uint8_t id[128], int rank, int nranks, int device,
                        cosl_ba_comm** out);
int cosl_ba_comm_destroy(cosl_ba_comm* c);

/* A solver handle keeps the problem resident on the device so that bench.py can time LM
 * iterations without the upload.  `prob` holds THIS RANK's shard of points/observations (n, nobs,
 * ptr, cam, xy, X local; m, m_con, K, R, t global and identical on all ranks).  n_con applies to
 * the global point order: pass the number of fixed points that fall into this shard. */
typedef struct cosl_ba_solver cosl_ba_solver;
int cosl_ba_solver_create(const cosl_ba_problem* prob, const cosl_ba_options* opt,
                          cosl_ba_comm* comm /* NULL = single GPU */, cosl_ba_solver** out);
/* Re-upload parameters (R, t, X) from the problem struct -- resets the solver to its start. */
int cosl_ba_solver_reset(cosl_ba_solver* s, const cosl_ba_problem* prob);
/* Run the robust loop (outer_iters x inner_iters) on the resident problem. */
int cosl_ba_solver_run(cosl_ba_solver* s, double info[COSL_BA_INFOSZ]);
/* Run exactly `trials` LM trials (linear solves) of one weighted inner loop without any stop test
 * (benchmark mode: LM-iterations/s).  */
int cosl_ba_solver_run_fixed(cosl_ba_solver* s, int trials, double info[COSL_BA_INFOSZ]);
/* Download R, t, X (and outlier flags) into the problem struct's arrays. */
int cosl_ba_solver_download(cosl_ba_solver* s, cosl_ba_problem* prob);
int cosl_ba_solver_destroy(cosl_ba_solver* s);
void* cosl_ba_solver_stream(cosl_ba_solver* s);
/* Per-kernel accumulated device time in ms since create/reset, by name (NULL-terminated list via
 * index): returns name or NULL when idx is out of range. */
const char* cosl_ba_solver_timer(cosl_ba_solver* s, int idx, double* ms, int* calls);
/* enable(1) resets and starts accumulating the per-kernel-class timers above. */
int cosl_ba_solver_profile_enable(cosl_ba_solver* s, int on);
/* Structure of the reduced camera system, for the roofline arithmetic of bench.py:
 * out[0] = order ns, out[1] = doubles in the all-reduce payload [rhs | Schur-structure tiles],
 * out[2] = multiply-adds of the tile Cholesky + substitutions on the used rows (ns^3/3 for the
 * single-CTA small solve), out[3] = Schur pair entries, out[4] = pair work items, out[5] = blocks,
 * out[6] = tiles incl. fill, out[7] = tasks of the solve DAG. */
int cosl_ba_solver_stats(cosl_ba_solver* s, double out[8]);
/* Plan of the reduced-system solve: out = {blocks, Schur-structure tiles, tiles incl. fill, tasks,
 * nested-dissection depth, longest dependency chain (tasks), persistent grid size, small-solve}. */
int cosl_ba_solver_plan_info(cosl_ba_solver* s, int out[8]);
/* Diagnostic timeline of the persistent solve kernel.  enable != 0: arm tracing for the following
 * solves (returns the task count).  enable == 0: copy the timeline of the LAST solve into
 * out[cap][8] = {SM id, globaltimer ns at ticket, when operands were ready, when done, up to 4
 * phase stamps inside the task} and
 * meta[cap][3] = {task type 0 POTRF / 1 TRSM / 2 UPDATE / 3 BACKWARD / 4 SUM, pivot block, row block},
 * then disarm. */
int cosl_ba_solver_trace(cosl_ba_solver* s, int enable, uint64_t* out, int32_t* meta, int cap);
/* Diagnostic: the task list of the solve plan (16 int32 per task, layout of BaTask in
 * coslam_b200/csrc/ba_plan.h) and the wait lists of the BACKWARD / SUM tasks as (counter, value)
 * pairs; returns the number of tasks. */
int cosl_ba_solver_tasks(cosl_ba_solver* s, int32_t* tasks, int cap, int32_t* lists, int listCap);

/* ================= post-BA pose-graph spreading =================
 * Replaces, for chain graphs, GlobalPoseGraph::computeNewCameraRotations
 * (slam/SL_GlobalPoseEstimation.cpp:52-218) followed by ::computeNewCameraTranslations (:220-359) as
 * RobustBundleRTS::updateNonKeyCameraPoses calls them once per camera
 * (app/SL_CoSLAMRobustBA.cpp:233-250) on the graphs constructCameraGraphs builds (:182-232): one
 * chain per camera, node k -> node k+1, key-frame nodes fixed.  All chains are solved by ONE launch.
 *   chainOff[nChains+1]  first node of each chain (chainOff[0] = 0), N = chainOff[nChains]
 *   fixed[N]             1 = fixed node (CamPoseNode::fixed)
 *   R[9N], t[3N]         node poses, row-major (CamPoseNode::R, ::t)
 *   eR[9N], et[3N]       edge k = rigid transform from node k to node k+1 (CamPoseEdge::R, ::t);
 *                        the slot of a chain's last node is not read
 *   newR[9N], newt[3N]   out: CamPoseNode::newR, ::newt (fixed nodes: copies of R, t)
 * Edges with uncertain scale and the constrained variants (:361-1281) are not on the post-BA path
 * and not provided.  COSL_E_INVALID when a chain with free nodes has no fixed node (the reference's
 * system is rank deficient there). */
int cosl_posegraph_spread_chains(int nChains, const int* chainOff, const uint8_t* fixed,
                                 const double* R, const double* t, const double* eR,
                                 const double* et, double* newR, double* newt, int device);

#ifdef __cplusplus
}
#endif
#endif /* COSLAM_B200_H_ */
