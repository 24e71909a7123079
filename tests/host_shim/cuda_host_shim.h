// Host-side shim so that the per-thread device helpers (linalg.cuh, solvers.cuh) can be compiled with g++
// and unit-tested on a CPU-only box (tests/test_host_solvers.py).  Test infrastructure only.
#pragma once
#include <algorithm>
#include <cmath>
#define __device__
#define __host__
#define __forceinline__ inline
#define __constant__ static const
struct float2 { float x, y; };
struct double2 { double x, y; };
using std::fmax;
using std::isfinite;
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
