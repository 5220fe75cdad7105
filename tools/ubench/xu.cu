// XU-pipe probe: MUFU.EX2, F2FP (cvt.rn.f16x2.f32), and both together, per SMSP (sm_100a)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2000
template <int OP>
__global__ void k(float* out, long long* cyc, float seed) {
  float a[8]; uint32_t u[8];
  for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 0.001f + i; u[i] = i; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) { uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[(i + 1) & 7])); u[i] ^= r; }
      if (OP == 2) { uint32_t r; asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[(i + 1) & 7])); u[i] ^= r; }
      if (OP == 3) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[(i + 1) & 7])); u[i] ^= r; }
      if (OP == 4) { float r; asm volatile("cvt.f32.f16 %0, %1;" : "=f"(r) : "h"((unsigned short)u[i])); a[i] += r; }
      if (OP == 5) { u[i] = __byte_perm(u[i], u[(i+1)&7], 0x7632) ^ u[i]; }
      if (OP == 6) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i+1)&7]), "f"(a[(i+2)&7])); }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i] + u[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int OP>
void run(const char* name, int w, int extra) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  k<OP><<<148, w * 128>>>(out, cyc, 0.5f);
  k<OP><<<148, w * 128>>>(out, cyc, 0.5f);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-30s warps/SMSP=%d  cycles per iteration-op per SMSP = %.2f\n", name, w, (double)c / (ITERS * 8.0 * w));
}
int main() {
  for (int w : {1, 2, 4}) {
    run<0>("MUFU.EX2", w, 1); run<1>("F2FP.f16 (+LOP3)", w, 1); run<2>("EX2 + F2FP (+LOP3)", w, 1);
    run<3>("F2FP.bf16 (+LOP3)", w, 1); run<4>("cvt f16->f32 + FADD", w, 1); run<5>("PRMT+LOP3", w, 1); run<6>("FFMA", w, 1);
  }
}
