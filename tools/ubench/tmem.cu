// Micro-benchmark: tcgen05.ld latency / throughput (sm_100a).  One CTA, 4 warps (one per TMEM lane quarter).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int NCOL>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t* r);
template <>
__device__ __forceinline__ void ld<16>(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void ld<32>(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// MODE 0: dependent chain (address of the next load depends on the previous data): latency
// MODE 1: independent back-to-back loads, one wait per load: throughput with waits
// MODE 2: 4 loads in flight then one wait
template <int NCOL, int MODE>
__global__ void k(long long* cyc, uint32_t* sink, int active_warps) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tbase + ((uint32_t)(warp * 32) << 16);
  uint32_t r[32];
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < active_warps) {
    // zero-ish init not possible without st; garbage is fine (masked to keep addresses valid)
    t0 = clock64();
    uint32_t off = 0;
    for (int it = 0; it < 256; ++it) {
      if (MODE == 0) {
        ld<NCOL>(base + off, r);
        wait_ld();
        off = (r[0] & 0u) + ((it * NCOL) & 255);   // data dependence, value forced to a valid column
        acc += r[1];
      } else if (MODE == 1) {
        ld<NCOL>(base + ((it * NCOL) & 255), r);
        wait_ld();
        acc += r[0] + r[NCOL - 1];
      } else {
        uint32_t r2[16], r3[16], r4[16];
        ld<16>(base + ((it * 64) & 255), r);
        ld<16>(base + ((it * 64 + 16) & 255), r2);
        ld<16>(base + ((it * 64 + 32) & 255), r3);
        ld<16>(base + ((it * 64 + 48) & 255), r4);
        wait_ld();
        acc += r[0] + r2[0] + r3[0] + r4[15];
      }
    }
    t1 = clock64();
  }
  sink[threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512) : "memory");
}

template <int NCOL, int MODE>
void run(const char* name, int warps) {
  long long* cyc; uint32_t* sink;
  cudaMalloc(&cyc, 8); cudaMalloc(&sink, 4096);
  k<NCOL, MODE><<<1, 128>>>(cyc, sink, warps);
  k<NCOL, MODE><<<1, 128>>>(cyc, sink, warps);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-40s active warps=%d  cycles per iteration = %.1f  (%s)\n", name, warps, c / 256.0, cudaGetErrorString(cudaGetLastError()));
  cudaFree(cyc); cudaFree(sink);
}

int main() {
  for (int w : {1, 4}) {
    run<16, 0>("x16 dependent chain (latency)", w);
    run<32, 0>("x32 dependent chain (latency)", w);
    run<16, 1>("x16 load + wait", w);
    run<32, 1>("x32 load + wait", w);
    run<16, 2>("4 x x16 in flight + one wait", w);
  }
  return 0;
}
