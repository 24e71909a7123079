// g++ -O2 -shared -fPIC -ffp-contract=off solvers_host.cpp -o libsolvers_host.so
#include "cuda_host_shim.h"
#include "../../ros2_mono_vo_b200/csrc/solvers.cuh"
extern "C" {
int host_solve_h4(const float* M, const float* m, double* H) { return mvo::solve_h4((const float2*)M, (const float2*)m, H); }
int host_solve_f7(const float* m1, const float* m2, double* F) { return mvo::solve_f7((const float2*)m1, (const float2*)m2, F); }
int host_solve_e5(const double* q1, const double* q2, double* E) { return mvo::solve_e5((const double2*)q1, (const double2*)q2, E); }
float host_h_error(const float* Hf, const float* M, const float* m) { return mvo::h_error(Hf, *(const float2*)M, *(const float2*)m); }
float host_f_error(const double* F, const float* a, const float* b) { return mvo::f_error(F, *(const float2*)a, *(const float2*)b); }
float host_e_error(const double* E, const double* a, const double* b) { return mvo::e_error(E, *(const double2*)a, *(const double2*)b); }
int host_real_roots10(const double* c, double* r) { return mvo::real_roots<10>(c, r); }
}
