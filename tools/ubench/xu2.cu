// Pipe probe 2: F2FP (cvt.rn.f16x2.f32) with a loop-carried dependence, HFMA2 with immediates, HMNMX2, LDTM round trip (sm_100a)
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#define ITERS 2000
template <int OP>
__global__ void k(float* out, long long* cyc, float seed) {
  float a[8]; uint32_t u[8];
  for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 0.001f + i; u[i] = 0x3c003c00u + i; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) { uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[(i + 1) & 7])); a[i] = __uint_as_float((r & 0x007fffffu) | 0x3f800000u); }
      if (OP == 1) { a[i] = __uint_as_float((__float_as_uint(a[i]) & 0x007fffffu) | 0x3f800000u); }
      if (OP == 2) { asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(0x3c003c00u), "r"(0x00010001u)); }
      if (OP == 3) { asm volatile("max.f16x2 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 7])); }
      if (OP == 4) { asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]), "r"(u[(i + 2) & 7])); }
      if (OP == 5) { uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[(i + 1) & 7])); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); a[i] = __uint_as_float((r & 0x007fffffu) | (__float_as_uint(a[i]) & 0x3f800000u)); }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i] + u[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int OP>
void run(const char* name, int w) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  k<OP><<<148, w * 128>>>(out, cyc, 0.5f);
  k<OP><<<148, w * 128>>>(out, cyc, 0.5f);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-34s warps/SMSP=%d  cycles per iteration-op per SMSP = %.2f\n", name, w, (double)c / (ITERS * 8.0 * w));
}
int main() {
  for (int w : {1, 2, 4}) {
    run<0>("F2FP + LOP3 (dependent)", w); run<1>("LOP3 (dependent)", w); run<2>("HFMA2 imm operands", w);
    run<3>("HMNMX2", w); run<4>("HFMA2 3-reg", w); run<5>("F2FP + EX2 + 2 LOP3", w);
  }
}
