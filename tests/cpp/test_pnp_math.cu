// Host harness for csrc/pnp_math.cuh: the per-hypothesis arithmetic of pnp.cu compiled for the CPU (same source the
// kernels run), so that tests/test_pnp_math_host.py can compare it with oracle/pnp_oracle.py without a GPU.
//   stdin : n, then n rows "X Y Z u v" (u, v = normalised image coordinates), n == 5
//   argv  : [--device [impl]] -- run epnp_solve_warp<5, impl> on the GPU instead of epnp_solve<5> on the host
//   stdout: R (9) t (3) of epnp_solve<5>, then rvec of rotation_to_rvec, then R again from rvec_to_rotation
#include <cstdio>
#include <cstdlib>
#include <string>
#include <cuda_runtime.h>
#include "../../ros2_mono_vo_b200/csrc/pnp_math.cuh"

// one warp: the form pnp_epnp_kernel runs (epnp_solve_warp: warp-cooperative 12 x 12 Jacobi, the three beta
// approximations on three lanes)
template <int IMPL>
__global__ void epnp_device(const double* in, double* out) {
  __shared__ double sA[144], sV[144];
  double pw[mvo::kPnpK][3], us[mvo::kPnpK][2];
  for (int i = 0; i < mvo::kPnpK; ++i) {
    for (int k = 0; k < 3; ++k) pw[i][k] = in[i * 5 + k];
    us[i][0] = in[i * 5 + 3];
    us[i][1] = in[i * 5 + 4];
  }
  double R[9], t[3];
  const bool ok = mvo::epnp_solve_warp<mvo::kPnpK, IMPL>(pw, us, sA, sV, threadIdx.x, R, t);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 9; ++i) out[i] = R[i];
    for (int i = 0; i < 3; ++i) out[9 + i] = t[i];
    out[12] = ok ? 1.0 : 0.0;
  }
}

int main(int argc, char** argv) {
  if (argc > 1 && std::string(argv[1]) == "--cs") {
    // rotation parameters of a Jacobi step, both forms (linalg.cuh: jacobi_cs): stdin "app aqq apq" per line
    double app, aqq, apq;
    while (scanf("%lf %lf %lf", &app, &aqq, &apq) == 3) {
      double c0, s0, c1, s1;
      mvo::jacobi_cs<false>(app, aqq, apq, c0, s0);
      mvo::jacobi_cs<true>(app, aqq, apq, c1, s1);
      printf("%.17g %.17g %.17g %.17g\n", c0, s0, c1, s1);
    }
    return 0;
  }
  int n = 0;
  if (scanf("%d", &n) != 1 || n != mvo::kPnpK) return 2;
  double pw[mvo::kPnpK][3], us[mvo::kPnpK][2];
  for (int i = 0; i < n; ++i)
    if (scanf("%lf %lf %lf %lf %lf", &pw[i][0], &pw[i][1], &pw[i][2], &us[i][0], &us[i][1]) != 5) return 2;
  double R[9], t[3], r[3], R2[9], J[27];
  bool ok;
  if (argc > 1) {   // --device: the warp form the kernels run
    double hin[25], hout[13], *din, *dout;
    for (int i = 0; i < 5; ++i) { hin[i * 5] = pw[i][0]; hin[i * 5 + 1] = pw[i][1]; hin[i * 5 + 2] = pw[i][2]; hin[i * 5 + 3] = us[i][0]; hin[i * 5 + 4] = us[i][1]; }
    cudaMalloc(&din, sizeof(hin)); cudaMalloc(&dout, sizeof(hout));
    cudaMemcpy(din, hin, sizeof(hin), cudaMemcpyHostToDevice);
    const int impl = argc > 2 ? atoi(argv[2]) : 1;   // the 12 x 12 Jacobi form (pnp_math.cuh: epnp_solve_warp)
    if (impl == 0) epnp_device<0><<<1, 32>>>(din, dout);
    else epnp_device<1><<<1, 32>>>(din, dout);
    if (cudaMemcpy(hout, dout, sizeof(hout), cudaMemcpyDeviceToHost) != cudaSuccess) return 3;
    for (int i = 0; i < 9; ++i) R[i] = hout[i];
    for (int i = 0; i < 3; ++i) t[i] = hout[9 + i];
    ok = hout[12] != 0;
  } else {
    ok = mvo::epnp_solve<mvo::kPnpK>(pw, us, R, t);
  }
  mvo::rotation_to_rvec(R, r);
  mvo::rvec_to_rotation(r, R2, J);
  printf("%d\n", ok ? 1 : 0);
  for (int i = 0; i < 9; ++i) printf("%.17g ", R[i]);
  for (int i = 0; i < 3; ++i) printf("%.17g ", t[i]);
  for (int i = 0; i < 3; ++i) printf("%.17g ", r[i]);
  for (int i = 0; i < 9; ++i) printf("%.17g ", R2[i]);
  for (int i = 0; i < 27; ++i) printf("%.17g ", J[i]);
  printf("\n");
  return 0;
}
