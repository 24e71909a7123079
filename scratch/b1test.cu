#include <stdint.h>
__global__ void k(const uint32_t* a, const uint32_t* b, int* c) {
  uint32_t a0=a[threadIdx.x], a1=a[threadIdx.x+32], a2=a[threadIdx.x+64], a3=a[threadIdx.x+96], b0=b[threadIdx.x], b1=b[threadIdx.x+32];
  int c0=0,c1=0,c2=0,c3=0;
  asm volatile("mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.and.popc {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
    : "+r"(c0),"+r"(c1),"+r"(c2),"+r"(c3) : "r"(a0),"r"(a1),"r"(a2),"r"(a3),"r"(b0),"r"(b1));
  c[threadIdx.x]=c0+c1+c2+c3;
}
