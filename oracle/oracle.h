/*
 * oracle.h -- C API of the CPU oracle (TEST INFRASTRUCTURE ONLY).
 *
 * The oracle is a plain C++ restatement of the reference arithmetic for the KLT, pose and BA hot
 * paths.  It exists to check the CUDA product path; nothing under coslam_b200/ may call, link or
 * import it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it.
 *
 * PINNING: pose (orc_pose_intracam) and pose-graph spreading (orc_posegraph_spread) are PINNED to the
 * reference's own code: slam/SL_IntraCamPose.cpp and slam/SL_GlobalPoseEstimation.cpp are compiled
 * unmodified into oracle/_ref/ (Makefile target `ref`) and compared live and through the vectors
 * tests/golden/pose_ref.npz / posegraph_ref.npz they produced (tests/test_pose_ref.py,
 * tests/test_posegraph.py).  KLT and BA are PARITY UNPINNED: the reference's KLT exists only as Cg
 * shaders and its BA arithmetic lives in LibVisualSLAM/sba-1.6, which is not in the tree
 * (SURVEY.md 8c); they are held by closed-form known-answer cases and independent cross-checks
 * (scipy / numpy restatements / numeric differentiation), see tests/test_oracle_*.py.
 *
 * Struct layouts are shared with include/coslam_b200.h so the same ctypes structures drive both.
 */
#ifndef COSLAM_ORACLE_H_
#define COSLAM_ORACLE_H_
#include "../include/coslam_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

void orc_set_threads(int n); /* OpenMP threads used by every oracle entry point (default 1) */
int orc_get_max_threads(void);

/* ---- KLT (mirrors cosl_klt_*) ---- */
typedef struct orc_klt orc_klt;
orc_klt* orc_klt_create(const cosl_klt_config* cfg, int width, int height, int nLevels, int featW,
                        int featH, int plW, int plH);
void orc_klt_destroy(orc_klt* h);
int orc_klt_detect(orc_klt* h, const uint8_t* img, size_t pitch, int nPresent,
                   const float* present3, cosl_klt_feature* dest, int* nDetected);
int orc_klt_redetect(orc_klt* h, const uint8_t* img, size_t pitch, cosl_klt_feature* dest,
                     int* nNewFeatures);
int orc_klt_track(orc_klt* h, const uint8_t* img, size_t pitch, cosl_klt_feature* dest,
                  int* nPresent);
int orc_klt_feed(orc_klt* h, int npts, const float* pts3, int* trackIds, int* nFed);
int orc_klt_advance(orc_klt* h);
void orc_klt_set_margin(orc_klt* h, float m);
void orc_klt_set_conv(orc_klt* h, float t);
void orc_klt_set_ssd(orc_klt* h, float t);
int orc_klt_debug_pyramid(orc_klt* h, int which, int level, float* out3, int* w, int* ht);
int orc_klt_debug_cornerness(orc_klt* h, float* out);
/* number of detector candidates found by the last detect/redetect (before top-N selection) */
int orc_klt_last_num_candidates(orc_klt* h);

/* ---- pose (mirrors cosl_pose_intracam) ---- */
int orc_pose_intracam(const double K[9], const double R0[9], const double t0[3], int npts,
                      const double* prevErrs, const double* Ms, const double* ms, double tau,
                      double R_opt[9], double t_opt[3], cosl_pose_opt* opt);
void orc_so3_exp(const double w[3], double R[9]);
void orc_project(const double K[9], const double R[9], const double t[3], const double M[3],
                 double m[2]);

/* ---- post-BA pose-graph spreading (general edge set; mirrors cosl_posegraph_spread_chains) ---- */
int orc_posegraph_spread(int nNodes, const int* fixed, const double* R, const double* t, int nEdges,
                         const int* id1, const int* id2, const double* eR, const double* et,
                         double* newR, double* newt);

/* ---- BA (mirrors cosl_ba_solve / cosl_sba_motstr_levmar_x) ---- */
int orc_ba_solve(cosl_ba_problem* prob, const cosl_ba_options* opt, double info[COSL_BA_INFOSZ]);
/* Exactly `trials` LM trials without stop tests (the bench unit), weights from the start point. */
int orc_ba_run_fixed(cosl_ba_problem* prob, const cosl_ba_options* opt, int trials,
                     double info[COSL_BA_INFOSZ]);
int orc_sba_motstr_levmar_x(int n, int ncon, int m, int mcon, const char* vmask, double* p, int cnp,
                            int pnp, const double* x, int mnp, const double* rot0params, int itmax,
                            int verbose, const double opts[5], double info[10]);
/* projection + analytic Jacobians of the KRTS camera model (for the numeric cross-check test):
 * cam = (fx,cx,cy,ar,s), q0 = unit quaternion (w,x,y,z), v = local quaternion vector part. */
void orc_ba_project(const double K[9], const double q0[4], const double v[3], const double t[3],
                    const double X[3], double xy[2], double A[12] /*2x6 wrt (v,t)*/,
                    double B[6] /*2x3 wrt X*/);
void orc_mat2quat(const double R[9], double q[4]);
void orc_quat2mat(const double q[4], double R[9]);
/* total weighted / unweighted squared reprojection error of a problem at its current parameters */
double orc_ba_cost(const cosl_ba_problem* prob);

#ifdef __cplusplus
}
#endif
#endif
