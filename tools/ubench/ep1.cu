// Micro-benchmark of the epilogue-1 inner step (16 MUFU + packed-half Horner of the previous chunk) in isolation.
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float fast_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) { const __half2 h = __floats2half2_rn(lo, hi); return *reinterpret_cast<const uint32_t*>(&h); }
__device__ __forceinline__ __half2 as_h2(uint32_t u) { return *reinterpret_cast<const __half2*>(&u); }
__device__ __forceinline__ uint32_t as_u32(__half2 h) { return *reinterpret_cast<const uint32_t*>(&h); }
struct Ep1Chunk { uint32_t th[8]; uint32_t xh[8]; };

template <bool DO_A, bool DO_B>
__device__ __forceinline__ void ep1_step(const float (&v)[16], Ep1Chunk& nxt, const Ep1Chunk& cur, const uint32_t (&cw)[8], uint32_t (&o)[8]) {
  const __half2 k3 = __float2half2_rn(-0.08479055f), k2 = __float2half2_rn(0.32563294f), k1 = __float2half2_rn(-0.67996303f), k0 = __float2half2_rn(1.43901745f), zero = __float2half2_rn(0.0f);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float t0 = 0.0f, t1 = 0.0f;
    if (DO_A) { t0 = fast_ex2(-fabsf(v[2 * j])); t1 = fast_ex2(-fabsf(v[2 * j + 1])); nxt.xh[j] = pack_f16x2(v[2 * j], v[2 * j + 1]); }
    if (DO_B) {
      const __half2 t = as_h2(cur.th[j]);
      __half2 q = __hfma2(k3, t, k2); q = __hfma2(q, t, k1); q = __hfma2(q, t, k0);
      q = __hfma2(t, q, __hmax2(as_h2(cur.xh[j]), zero));
      o[j] = as_u32(__hfma2(q, as_h2(cw[j]), __hneg2(as_h2(cw[j]))));
    }
    if (DO_A) nxt.th[j] = pack_f16x2(t0, t1);
  }
}

template <int MODE>
__global__ void k(uint32_t* out, long long* cyc, float seed, int iters) {
  __shared__ uint4 sm[1024];
  float v[16];
  for (int i = 0; i < 16; ++i) v[i] = seed + threadIdx.x * 0.01f + i;
  Ep1Chunk s0, s1;
  uint32_t cw[8], o[8];
  for (int i = 0; i < 8; ++i) { cw[i] = 0x3c003c00u; o[i] = 0; s0.th[i] = 0x38003800u; s0.xh[i] = 0x3c003c00u; }
  __syncthreads();
  long long t0 = clock64();
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {          // both stages interleaved (steady state of the kernel loop)
      ep1_step<true, true>(v, s1, s0, cw, o);
      for (int i = 0; i < 16; ++i) v[i] += __uint_as_float(o[i & 7] & 0x3f800000u);   // keep a dependence, cheap
      sm[(threadIdx.x + it) & 1023] = make_uint4(o[0], o[1], o[2], o[3]);
      sm[(threadIdx.x + it + 512) & 1023] = make_uint4(o[4], o[5], o[6], o[7]);
      ep1_step<true, true>(v, s0, s1, cw, o);
      sm[(threadIdx.x + it + 7) & 1023] = make_uint4(o[0], o[1], o[2], o[3]);
      sm[(threadIdx.x + it + 519) & 1023] = make_uint4(o[4], o[5], o[6], o[7]);
    } else if (MODE == 1) {   // stage A only
      ep1_step<true, false>(v, s1, s0, cw, o);
      for (int i = 0; i < 8; ++i) acc += s1.th[i] ^ s1.xh[i];
      for (int i = 0; i < 16; ++i) v[i] += 1e-3f;
      ep1_step<true, false>(v, s0, s1, cw, o);
      for (int i = 0; i < 8; ++i) acc += s0.th[i] ^ s0.xh[i];
    } else {                  // stage B only
      ep1_step<false, true>(v, s1, s0, cw, o);
      for (int i = 0; i < 8; ++i) { s0.th[i] ^= (o[i] & 1u); }
      ep1_step<false, true>(v, s1, s0, cw, o);
      for (int i = 0; i < 8; ++i) { s0.xh[i] ^= (o[i] & 1u); }
    }
  }
  long long t1 = clock64();
  for (int i = 0; i < 8; ++i) acc += o[i] + s0.th[i] + s1.xh[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (uint32_t)v[3];
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps_per_smsp) {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 500;
  k<MODE><<<148, warps_per_smsp * 128>>>(out, cyc, 0.5f, iters);
  k<MODE><<<148, warps_per_smsp * 128>>>(out, cyc, 0.5f, iters);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-34s warps/SMSP=%d  cycles per 16-column chunk per warp = %.1f  (SMSP throughput: %.1f cycles/chunk)\n", name, warps_per_smsp,
         c / (2.0 * iters), c / (2.0 * iters * warps_per_smsp));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {1, 2, 3, 4}) {
    run<0>("ep1 A(k+1) + B(k) interleaved", w);
    run<1>("ep1 stage A only (16 MUFU)", w);
    run<2>("ep1 stage B only (Horner)", w);
  }
  return 0;
}
