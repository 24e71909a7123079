// known-good pattern from the CUDA programming guide (libcu++ wrappers), 2-D and 3-D u8 tiles
#include <cuda.h>
#include <cuda/barrier>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;

__global__ void k3(const __grid_constant__ CUtensorMap tmap, int x, int y, int b, int rows, uint32_t* out) {
  __shared__ alignas(128) uint32_t buf[256];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_3d_global_to_shared(buf, &tmap, x, y, b, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, 32 * rows);
  } else {
    token = bar.arrive();
  }
  bar.wait(std::move(token));
  for (int i = threadIdx.x; i < 8 * rows; i += blockDim.x) out[i] = buf[i];
}

int main(int argc, char** argv) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeFn encode = (EncodeFn)fn;
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int pitch = 1280, rows = 376, planes = 3, boxh = 29;
  const long long fs = (long long)pitch * rows + 256;
  std::vector<uint8_t> h(fs * planes);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 7 + (i >> 9));
  uint8_t* d; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  CUtensorMap tm;
  cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)planes};
  cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)fs};
  cuuint32_t box[3] = {32u, (cuuint32_t)boxh, 1u}; cuuint32_t es[3] = {1u, 1u, 1u};
  if (variant == 1) { dims[0] = 1241; }                       // inner extent not a multiple of 16
  if (variant == 2) { box[0] = 64; }
  CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, variant == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("variant %d encode %d\n", variant, (int)r);
  uint32_t* out; cudaMalloc(&out, 4096); cudaMemset(out, 0, 4096);
  const int x = argc > 2 ? atoi(argv[2]) : 37, y = 11, b = 2;
  k3<<<1, 64>>>(tm, x, y, b, boxh * (variant == 2 ? 2 : 1), out);
  cudaError_t e = cudaDeviceSynchronize();
  uint8_t ho[4096]; cudaMemcpy(ho, out, 4096, cudaMemcpyDeviceToHost);
  int bad = 0; const int bw = variant == 2 ? 64 : 32;
  for (int rr = 0; rr < boxh; ++rr) for (int c = 0; c < bw; ++c) bad += ho[rr * bw + c] != h[(size_t)b * fs + (size_t)(y + rr) * pitch + x + c];
  printf("variant %d: %s mismatches %d\n", variant, cudaGetErrorString(e), bad);
  return 0;
}
