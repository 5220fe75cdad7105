// Single-tile UMMA probe: D[128, N] = A * B with operands handed over as ready-made shared-memory
// images.  tests/test_gpu_umma.py packs A/B with the canonical-layout formulas of tc_common.cuh and
// checks D against PyTorch, which pins descriptor encoding, instruction descriptor and the
// TMEM lane/column mapping independently of the fused kernels built on them.
#include "common.cuh"
#include "tc_common.cuh"

namespace cmp {
namespace {

__global__ void __launch_bounds__(128)
umma_probe_kernel(const uint8_t* __restrict__ a_img, uint32_t a_bytes, const uint8_t* __restrict__ b_img,
                  uint32_t b_bytes, float* __restrict__ D, int N, int K, int fmt, int a_mn, int b_mn, uint32_t a_lbo,
                  uint32_t a_sbo, uint32_t a_kstep, uint32_t b_lbo, uint32_t b_sbo, uint32_t b_kstep) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((a_bytes + 1023) / 1024) * 1024;
  for (uint32_t i = threadIdx.x * 16; i < a_bytes; i += blockDim.x * 16)
    *reinterpret_cast<uint4*>(sa + i) = *reinterpret_cast<const uint4*>(a_img + i);
  for (uint32_t i = threadIdx.x * 16; i < b_bytes; i += blockDim.x * 16)
    *reinterpret_cast<uint4*>(sb + i) = *reinterpret_cast<const uint4*>(b_img + i);
  tc::fence_proxy_async();
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_fence_init();
  }
  if (threadIdx.x < 32) tc::tmem_alloc(&tmem_base_s, 256);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (threadIdx.x == 0) {
    const uint32_t idesc = tc::umma_idesc_f16(128, N, fmt, a_mn, b_mn);
    for (int ks = 0; ks < K / 16; ++ks) {
      uint64_t ad = tc::umma_smem_desc(tc::smem_u32(sa) + ks * a_kstep, a_lbo, a_sbo);
      uint64_t bd = tc::umma_smem_desc(tc::smem_u32(sb) + ks * b_kstep, b_lbo, b_sbo);
      tc::umma_f16(tmem_base, ad, bd, idesc, ks > 0 ? 1u : 0u);
    }
    tc::umma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tc::tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
    tc::tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c0 + j < N) D[(int64_t)row * N + c0 + j] = v[j];
  }
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(tmem_base, 256);
}

}  // namespace
}  // namespace cmp

using namespace cmp;

extern "C" int cmp_debug_umma_gemm(const void* a_img, int64_t a_bytes, const void* b_img, int64_t b_bytes, float* D,
                                   int N, int K, int fmt, int a_mn, int b_mn, int a_lbo, int a_sbo, int a_kstep,
                                   int b_lbo, int b_sbo, int b_kstep, cmp_stream_t stream) {
  CMP_REQUIRE(a_img && b_img && D, CMP_EINVAL, "cmp_debug_umma_gemm: null pointer");
  CMP_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0 && K >= 16 && K % 16 == 0, CMP_EINVAL,
              "cmp_debug_umma_gemm: N must be a multiple of 16 in [16,256], K a multiple of 16");
  CMP_REQUIRE(a_bytes % 16 == 0 && b_bytes % 16 == 0, CMP_EINVAL, "cmp_debug_umma_gemm: images must be 16-byte multiples");
  CMP_REQUIRE(cmp_device_is_sm100(), CMP_EUNSUPPORTED, "cmp_debug_umma_gemm: needs an sm_100 device");
  size_t smem = ((size_t)(a_bytes + 1023) / 1024) * 1024 + (size_t)b_bytes + 1024;
  CMP_REQUIRE(smem <= 200 * 1024, CMP_EUNSUPPORTED, "cmp_debug_umma_gemm: operands too large");
  if (cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("cmp_debug_umma_gemm: cannot opt in to %zu bytes of shared memory", smem);
    return CMP_ECUDA;
  }
  umma_probe_kernel<<<1, 128, smem, as_stream(stream)>>>(
      reinterpret_cast<const uint8_t*>(a_img), (uint32_t)a_bytes, reinterpret_cast<const uint8_t*>(b_img),
      (uint32_t)b_bytes, D, N, K, fmt, a_mn, b_mn, (uint32_t)a_lbo, (uint32_t)a_sbo, (uint32_t)a_kstep, (uint32_t)b_lbo,
      (uint32_t)b_sbo, (uint32_t)b_kstep);
  CMP_LAUNCH_CHECK("cmp_debug_umma_gemm");
  return CMP_OK;
}
