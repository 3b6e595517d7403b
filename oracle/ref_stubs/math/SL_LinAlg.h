/* Stand-in for LibVisualSLAM math/SL_LinAlg.h (danping/LibVisualSLAM, not vendored by the
 * reference): ONLY the declarations /root/reference/src/slam/SL_IntraCamPose.cpp needs to compile
 * unmodified.  Definitions and the semantics inferred from the call sites are in primitives.cpp.
 * TEST INFRASTRUCTURE (oracle/_ref build), never part of the product. */
#pragma once
#include <cmath>
#include <cstring>
void matATB(int ma, int na, int mb, int nb, const double* A, const double* B, double* C);
void matAB(int ma, int na, int mb, int nb, const double* A, const double* B, double* C);
bool matInv(int n, const double* A, double* invA);
void mat33AB(const double* A, const double* B, double* C);
void mat22Inv(const double* A, double* invA);
void mat33Inv(const double* A, double* invA);
void doubleArrCopy(double* dst, int dstStart, const double* src, int len);
