// Drop-in for reference image_compression/include/PowerMethod.hpp (image_compression/src/PowerMethod.cpp:3-128):
// powerMethod(A, B, sigma, u, v) and its MPI twin -- the older spelling of PM (include/PM.hpp).  SURVEY.md 8(f) rank 1.
#ifndef POWER_METHOD_H
#define POWER_METHOD_H

#include "../PM.hpp"

inline void powerMethod(Mat_m& A, Mat_m& B, double& sigma, Vec_v& u, Vec_v& v) { PM(A, B, sigma, u, v); }
// the reference's MPI variant splits the B*x mat-vec over ranks (PowerMethod.cpp:45-128); same result
inline void powerMethod_mpi(Mat_m& A, Mat_m& B, double& sigma, Vec_v& u, Vec_v& v) { PM(A, B, sigma, u, v); }

#endif  // POWER_METHOD_H
