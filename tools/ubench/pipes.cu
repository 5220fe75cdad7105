// Micro-benchmark: per-SMSP throughput of the instruction classes the fused epilogues are made of (sm_100a).
// One CTA per SM, W warps per SMSP; every warp runs N iterations of an unrolled block of independent instructions.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define ITERS 2000

template <int OP>
__global__ void k(float* out, long long* cyc, float seed) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 0.001f + i;
  __half2 h[8];
  for (int i = 0; i < 8; ++i) h[i] = __floats2half2_rn(a[i], a[i] * 0.5f);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) a[i] = fmaf(a[i], 1.0001f, 0.5f);
      if (OP == 2) h[i] = __hfma2(h[i], h[(i + 1) & 7], h[i]);
      if (OP == 3) { __half2 t = __floats2half2_rn(a[i], a[(i + 1) & 7]); a[i] += __low2float(t); }   // F2FP + cvt back + add
      if (OP == 4) { uint32_t u = *reinterpret_cast<uint32_t*>(&h[i]); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u)); h[i] = *reinterpret_cast<__half2*>(&u); }
      if (OP == 5) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); h[i] = __hfma2(h[i], h[(i + 1) & 7], h[i]); h[i] = __hfma2(h[i], h[(i + 2) & 7], h[i]); h[i] = __hfma2(h[i], h[(i + 3) & 7], h[i]); }
      if (OP == 6) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 7) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 8) { uint32_t u = *reinterpret_cast<uint32_t*>(&h[i]); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u)); h[i] = *reinterpret_cast<__half2*>(&u); }
      if (OP == 9) a[i] = __fmul_rn(a[i], 1.0001f);
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i] + __low2float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int OP>
void run(const char* name, int warps_per_smsp, int extra) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 8);
  int threads = warps_per_smsp * 4 * 32;
  k<OP><<<148, threads>>>(out, cyc, 0.5f);
  k<OP><<<148, threads>>>(out, cyc, 0.5f);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  double per = (double)c / (ITERS * 8.0 * warps_per_smsp * extra);
  printf("%-28s warps/SMSP=%d  cycles per warp-instruction per SMSP = %.2f\n", name, warps_per_smsp, per);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int w : {1, 2, 4}) {
    run<0>("MUFU.EX2 f32", w, 1);
    run<6>("MUFU.LG2 f32", w, 1);
    run<7>("MUFU.RSQ f32", w, 1);
    run<4>("MUFU.EX2 f16x2", w, 1);
    run<8>("MUFU.TANH f16x2", w, 1);
    run<1>("FFMA", w, 1);
    run<9>("FMUL", w, 1);
    run<2>("HFMA2", w, 1);
    run<3>("F2FP+cvt+FADD (3 instr)", w, 3);
    run<5>("EX2 + 3 HFMA2 (4 instr)", w, 4);
  }
  return 0;
}
