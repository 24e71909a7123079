// probe: which way of handing a CUtensorMap to cp.async.bulk.tensor works on this box
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <vector>

struct Maps { CUtensorMap m[2][4]; };
struct __align__(128) Sm { uint32_t a[256]; uint32_t b[256]; unsigned long long bar[2]; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tile3(void* dst, const CUtensorMap* map, int x, int y, int b, unsigned long long* bar, uint32_t bytes) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(b), "r"(smem_u32(bar)) : "memory");
}
__device__ __noinline__ void tile3_noinline(void* dst, const CUtensorMap* map, int x, int y, int b, unsigned long long* bar, uint32_t bytes) {
  tile3(dst, map, x, y, b, bar, bytes);
}

__global__ void probe(const __grid_constant__ Maps tm, int mode, int L, int x, int y, int b, int rows, uint32_t* out, int* status) {
  __shared__ Sm sm[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Sm& s = sm[warp];
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s.bar[0])), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (lane == 0) {
    if (mode == 0) tile3(s.a, &tm.m[1][L], x, y, b, &s.bar[0], 32 * rows);
    else tile3_noinline(s.a, &tm.m[1][L], x, y, b, &s.bar[0], 32 * rows);
  }
  int spins = 0;
  while (!try_wait(&s.bar[0], 0)) if (++spins > (1 << 22)) break;
  if (lane == 0) status[warp] = spins;
  for (int i = lane; i < 8 * rows; i += 32) out[warp * 256 + i] = s.a[i];
}

__global__ void probe1(const __grid_constant__ CUtensorMap tmap, int mode, int x, int y, int b, int rows, uint32_t* out, int* status) {
  __shared__ Sm sm[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Sm& s = sm[warp];
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s.bar[0])), "r"(1) : "memory");
    if (mode == 4) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    else asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (lane == 0) tile3(s.a, &tmap, x, y, b, &s.bar[0], 32 * rows);
  int spins = 0;
  while (!try_wait(&s.bar[0], 0)) if (++spins > (1 << 22)) break;
  if (lane == 0) status[warp] = spins;
  for (int i = lane; i < 8 * rows; i += 32) out[warp * 256 + i] = s.a[i];
}
__global__ void probe3(const __grid_constant__ Maps tm, int x, int y, int b, int rows, uint32_t* out, int* status) {
  __shared__ Sm sm[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Sm& s = sm[warp];
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s.bar[0])), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (lane == 0) tile3(s.a, &tm.m[1][1], x, y, b, &s.bar[0], 32 * rows);
  int spins = 0;
  while (!try_wait(&s.bar[0], 0)) if (++spins > (1 << 22)) break;
  if (lane == 0) status[warp] = spins;
  for (int i = lane; i < 8 * rows; i += 32) out[warp * 256 + i] = s.a[i];
}

int main(int argc, char** argv) {
  const int want_mode = argc > 1 ? atoi(argv[1]) : 0;
  const int want_boxh = argc > 2 ? atoi(argv[2]) : 29;
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  printf("entry %p q=%d\n", fn, (int)q);
  EncodeFn encode = (EncodeFn)fn;
  const int pitch = 1280, rows = 376, planes = 3;
  const long long fs = (long long)pitch * rows + 256;
  std::vector<uint8_t> h(fs * planes);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 7 + (i >> 9));
  uint8_t* d; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  for (int boxh : {want_boxh}) {
    Maps tm; memset(&tm, 0, sizeof(tm));
    for (int k = 0; k < 2; ++k) for (int l = 0; l < 4; ++l) {
      cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)planes};
      cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)fs};
      cuuint32_t box[3] = {32u, (cuuint32_t)boxh, 1u}; cuuint32_t es[3] = {1u, 1u, 1u};
      CUresult r = encode(&tm.m[k][l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) printf("encode failed %d (boxh %d)\n", (int)r, boxh);
    }
    uint32_t* out; int* st; cudaMalloc(&out, 2 * 1024); cudaMalloc(&st, 8);
    for (int mode : {want_mode}) {
      cudaMemset(out, 0, 2048); cudaMemset(st, 0xff, 8);
      const int x = 37, y = 11, b = 2;
      if (mode <= 1) probe<<<1, 64>>>(tm, mode, 1, x, y, b, boxh, out, st);
      else if (mode == 3) probe3<<<1, 64>>>(tm, x, y, b, boxh, out, st);
      else probe1<<<1, 64>>>(tm.m[1][1], mode, x, y, b, boxh, out, st);
      cudaError_t e = cudaDeviceSynchronize();
      uint32_t ho[512]; int hs[2];
      cudaMemcpy(ho, out, 2048, cudaMemcpyDeviceToHost); cudaMemcpy(hs, st, 8, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int w = 0; w < 2; ++w) for (int r = 0; r < boxh; ++r) for (int c = 0; c < 32; ++c) {
        const uint8_t got = ((uint8_t*)ho)[w * 1024 + r * 32 + c];
        const uint8_t exp = h[(size_t)b * fs + (size_t)(y + r) * pitch + x + c];
        bad += got != exp;
      }
      printf("boxh %d mode %d: %s spins %d %d mismatches %d\n", boxh, mode, cudaGetErrorString(e), hs[0], hs[1], bad);
      if (e != cudaSuccess) return 1;
    }
  }
  return 0;
}
