// Issue-rate microbenchmark of the ALU / FMA pipe instructions the dequant path uses (development tool).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP 64
template <int OP>
__global__ void k(uint32_t* out, long long* cyc, uint32_t seed) {
  uint32_t a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 8 + i;
  uint32_t b = seed * 3 + 1, c = seed * 7 + 5;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 64; ++it) {
#pragma unroll
    for (int r = 0; r < REP / 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if constexpr (OP == 0) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
        if constexpr (OP == 1) asm volatile("add.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
        if constexpr (OP == 2) asm volatile("mul.rn.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
        if constexpr (OP == 3) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float*)&a[i]) : "f"(__uint_as_float(b)), "f"(__uint_as_float(c)));
        if constexpr (OP == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0xea;" : "+r"(a[i]) : "r"(b), "r"(c));
        if constexpr (OP == 5) asm volatile("shf.r.clamp.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
        if constexpr (OP == 6) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
        if constexpr (OP == 7) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
        if constexpr (OP == 8) asm volatile("fma.rn.bf16x2 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
        if constexpr (OP == 9) asm volatile("sub.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
        if constexpr (OP == 10) {   // lop3 + hfma2 alternating (independent chains)
          if (i & 1) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
          else asm volatile("lop3.b32 %0, %0, %1, %2, 0xea;" : "+r"(a[i]) : "r"(b), "r"(c));
        }
        if constexpr (OP == 11) asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(a[i]) : "r"(b));
        if constexpr (OP == 12) {   // ffma + hfma2 alternating
          if (i & 1) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
          else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float*)&a[i]) : "f"(__uint_as_float(b)), "f"(__uint_as_float(c)));
        }
        if constexpr (OP == 13) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
        if constexpr (OP == 14) asm volatile("fma.rn.f16 %0, %0, %1, %2;" : "+h"(*(unsigned short*)&a[i]) : "h"((unsigned short)b), "h"((unsigned short)c));
        if constexpr (OP == 15) asm volatile("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(a[i]) : "h"((unsigned short)(a[i] ^ b)));
      }
    }
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name) {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  for (int warps : {4, 8, 16, 32}) {
    k<OP><<<1, warps * 32>>>(out, cyc, 12345);
    k<OP><<<1, warps * 32>>>(out, cyc, 12345);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)h / (64.0 * REP);           // cycles per instruction per warp
    double rt = per / (warps / 4.0);                 // cycles per warp-instruction per SMSP
    printf("%-28s warps/SMSP=%d  cyc/instr/warp=%.2f  rt_SMSP=%.2f\n", name, warps / 4, per, rt);
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("fma.f16x2 (3 src)");
  run<11>("fma.f16x2 (b,b)");
  run<1>("add.f16x2");
  run<9>("sub.f16x2");
  run<2>("mul.f16x2");
  run<8>("fma.bf16x2");
  run<14>("fma.f16 scalar");
  run<3>("fma.f32");
  run<4>("lop3");
  run<5>("shf");
  run<6>("prmt");
  run<7>("mad.lo.u32");
  run<13>("add.u32");
  run<10>("lop3 + fma.f16x2 alternating");
  run<12>("fma.f32 + fma.f16x2 alternating");
  run<15>("cvt f16x2<-e4m3x2");
  cudaError_t e = cudaGetLastError();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
