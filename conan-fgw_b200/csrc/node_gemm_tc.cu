// Node-level linears on tcgen05 with split-bf16 operands (hi + lo, three MMA passes: hi*hi + hi*lo +
// lo*hi, fp32 accumulation in TMEM), i.e. ~16 mantissa bits per operand: the node features stay
// fp32-grade while the GEMMs run on the tensor pipe.  Replaces the torch.nn.Linear calls around the
// message passing (PyG CFConv.lin1/lin2, InteractionBlock.lin, ConAN heads sns.py:177-179,225-231)
// and their autograd GEMMs in the bf16 mode; the exact-fp32 mode keeps cmp_gemm_f32.
//
//   forward / dX:  Y[M, Nout] = act(X'[M, K] * W[Nout, K]^T + b) + R      (dX: W := W^T image)
//                  orientation D[out channel (TMEM lane), atom (column)] so the epilogue thread that
//                  owns a channel writes coalesced rows and holds its bias in a register;
//   dW / db:       D[Nout, K (+ ones column)] += dY'^T X over 128-atom tiles (K of the MMA = atoms), both
//                  operand images are re-read as MN-major views (LBO/SBO exchanged), per-CTA partial sums
//                  are reduced in a fixed order.
//   X' / dY' = input * (1 - exp(-y)/2) when saved_y is given: the ShiftedSoftplus backward fused into the
//   operand load.
#include "common.cuh"
#include "tc_common.cuh"

namespace cmp {
namespace {

constexpr int TM = 64;             // atoms per tile of the weight-gradient kernels (UMMA K runs over atoms); two CTAs
                                   // per SM overlap each other's operand loads, conversions and MMAs
constexpr int TMG = TM / 8;        // 8-atom groups per tile
constexpr int TMF = 64;            // atoms per tile of the forward / dX kernel: 2 CTAs per SM hide each other's loads
constexpr int CW = 8;              // compute warps
constexpr int NT = CW * 32 + 32;   // + 1 MMA warp
constexpr int MAXC = 128;          // max channels (K or Nout)

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// fp32 tile [TM atoms, C channels] -> bf16 hi / lo K-major images (rows = atoms):
//   byte(a, c) = (a%8)*16 + (c%8)*2 + (a/8)*sbo + (c/8)*128
// when ones_chunk >= 0 that 16-byte chunk of every row is set to {1, 0, 0, ...} (hi) / 0 (lo).
struct NoWait {
  __device__ __forceinline__ void operator()() const {}
};
// `after_loads` runs once, after the first batch of global loads has been issued and before anything is written to the
// images: the place to wait for the previous reader of the images, so that the loads overlap its tail.
template <int ROWS, int NW = CW, class AfterLoads = NoWait>
__device__ __forceinline__ void convert_tile(const float* __restrict__ X, int64_t ld, const float* __restrict__ Ysaved,
                                             int64_t ldys, int64_t m0, int64_t M, int C, uint32_t sbo, uint8_t* hi,
                                             uint8_t* lo, int ones_chunk, int warp, int lane,
                                             AfterLoads after_loads = AfterLoads()) {
  const int nchunk = C >> 3;
  const int cbs = (nchunk + 3) >> 2;
  const int al = lane & 7, cl = lane >> 3;
  constexpr int UN = 4;   // warp-iterations whose global loads are issued back to back (memory-level parallelism)
  const int total = (ROWS / 8) * cbs;
  for (int base = warp; base < total; base += NW * UN) {
   float4 P0[UN], P1[UN], Y0[UN], Y1[UN];
   const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
   for (int u = 0; u < UN; ++u) {
     const int wi = base + u * NW;
     const int ab = wi / cbs, cb = wi - ab * cbs;
     const int a = ab * 8 + al, chunk = cb * 4 + cl;
     const int64_t row = m0 + a;
     P0[u] = P1[u] = zero4;
     Y0[u] = Y1[u] = make_float4(1e30f, 1e30f, 1e30f, 1e30f);   // exp(-y) = 0: factor 1
     if (wi < total && chunk < nchunk && row < M) {
       const float4* src = reinterpret_cast<const float4*>(X + row * ld + chunk * 8);
       P0[u] = __ldg(src);
       P1[u] = __ldg(src + 1);
       if (Ysaved) {
         const float4* ys = reinterpret_cast<const float4*>(Ysaved + row * ldys + chunk * 8);
         Y0[u] = __ldg(ys);
         Y1[u] = __ldg(ys + 1);
       }
     }
   }
   if (base == warp) after_loads();
#pragma unroll
   for (int u = 0; u < UN; ++u) {
    const int wi = base + u * NW;
    const int ab = wi / cbs, cb = wi - ab * cbs;
    const int a = ab * 8 + al, chunk = cb * 4 + cl;
    if (wi >= total || chunk >= nchunk) continue;
    float v[8] = {P0[u].x, P0[u].y, P0[u].z, P0[u].w, P1[u].x, P1[u].y, P1[u].z, P1[u].w};
    if (Ysaved) {
      const float y[8] = {Y0[u].x, Y0[u].y, Y0[u].z, Y0[u].w, Y1[u].x, Y1[u].y, Y1[u].z, Y1[u].w};
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= 1.0f - 0.5f * ex2_approx(-1.4426950408889634f * y[j]);
    }
    float l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float h = __bfloat162float(__float2bfloat16_rn(v[j]));
      l[j] = v[j] - h;
      v[j] = h;
    }
    const uint32_t off = (a & 7) * 16 + (a >> 3) * sbo + chunk * 128;
    *reinterpret_cast<uint4*>(hi + off) = make_uint4(tc::pack_bf16x2(v[0], v[1]), tc::pack_bf16x2(v[2], v[3]),
                                                      tc::pack_bf16x2(v[4], v[5]), tc::pack_bf16x2(v[6], v[7]));
    *reinterpret_cast<uint4*>(lo + off) = make_uint4(tc::pack_bf16x2(l[0], l[1]), tc::pack_bf16x2(l[2], l[3]),
                                                      tc::pack_bf16x2(l[4], l[5]), tc::pack_bf16x2(l[6], l[7]));
   }
  }
  if (warp >= total) after_loads();   // narrow tiles leave some warps without work: they still owe the one call
  if (ones_chunk >= 0) {
    // two extra chunks (16 channels): chunk ones_chunk = {1,0,...}, ones_chunk + 1 = 0
    const int t = warp * 32 + lane;
    for (int item = t; item < ROWS * 2; item += NW * 32) {
      const int a = item >> 1, which = item & 1;
      const uint32_t off = (a & 7) * 16 + (a >> 3) * sbo + (ones_chunk + which) * 128;
      const bool live = (m0 + a) < M && which == 0;
      *reinterpret_cast<uint4*>(hi + off) = make_uint4(live ? 0x00003F80u : 0u, 0u, 0u, 0u);   // bf16(1.0) = 0x3F80
      *reinterpret_cast<uint4*>(lo + off) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}

struct FwdParams {
  const float* X;
  int64_t ldx;
  const float* saved_y;   // optional, same shape as X
  int64_t ldys;
  const uint8_t* w_img;   // hi image | lo image, each TM*K*2 bytes (rows = Nout padded to 128)
  const float* bias;
  const float* residual;
  int64_t ldr;
  float* Y;
  int64_t ldy;
  int64_t M;
  int K, Nout, act;
};

__global__ void __launch_bounds__(NT, 2) node_gemm_fwd_kernel(const FwdParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bars[3];   // wbar, xready, dready
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = p.K;
  const uint32_t w_bytes = 128 * K * 2, x_bytes = TMF * K * 2;
  const uint32_t sbo = (K >> 3) * 128;
  uint8_t* sWh = smem;
  uint8_t* sWl = smem + w_bytes;
  uint8_t* sXh = smem + 2 * w_bytes;
  uint8_t* sXl = sXh + x_bytes;

  if (tid == 0) {
    tc::mbar_init(&bars[0], 1);
    tc::mbar_init(&bars[1], CW * 32);
    tc::mbar_init(&bars[2], 1);
    tc::mbar_fence_init();
  }
  __syncwarp();
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 64);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int64_t ntiles = (p.M + TMF - 1) / TMF;
  pdl_launch_dependents();

  if (warp == CW) {
    if (lane == 0) {
      pdl_wait();   // the weight image may have been packed by the launch right before this one
      tc::mbar_arrive_expect_tx(&bars[0], 2 * w_bytes);
      tc::bulk_g2s(sWh, p.w_img, 2 * w_bytes, &bars[0]);
      tc::mbar_wait(&bars[0], 0);
      const uint32_t aWh = tc::smem_u32(sWh), aWl = tc::smem_u32(sWl), aXh = tc::smem_u32(sXh), aXl = tc::smem_u32(sXl);
      const uint32_t idesc = tc::umma_idesc_f16(128, TMF, 1, 0, 0);
      uint32_t it = 0;
      for (int64_t ti = blockIdx.x; ti < ntiles; ti += gridDim.x, ++it) {
        tc::mbar_wait(&bars[1], it & 1);
        tc::tc_fence_after();
        const int ks_n = K >> 4;
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t a = (pass == 2) ? aWl : aWh;
          const uint32_t b = (pass == 1) ? aXl : aXh;
          for (int ks = 0; ks < ks_n; ++ks)
            tc::umma_f16(tmem_base, tc::umma_smem_desc(a + ks * 256, 128, sbo), tc::umma_smem_desc(b + ks * 256, 128, sbo),
                         idesc, (pass | ks) != 0);
        }
        tc::umma_commit(&bars[2]);
      }
    }
    __syncwarp();
  } else {
    const int wq = warp & 3, h = warp >> 2;
    const int chan = wq * 32 + lane;
    pdl_wait();
    const float bias = (p.bias && chan < p.Nout) ? p.bias[chan] : 0.0f;
    const uint32_t tD = tmem_base + ((uint32_t)(wq * 32) << 16);
    uint32_t it = 0;
    for (int64_t ti = blockIdx.x; ti < ntiles; ti += gridDim.x, ++it) {
      const int64_t m0 = ti * TMF;
      // everyone finished reading TMEM / previous images consumed - checked between this tile's loads and its image writes
      auto images_free = [&]() {
        if (it > 0) tc::named_bar_sync(1, CW * 32);
      };
      convert_tile<TMF, CW>(p.X, p.ldx, p.saved_y, p.ldys, m0, p.M, K, sbo, sXh, sXl, -1, warp, lane, images_free);
      tc::fence_proxy_async();
      tc::mbar_arrive(&bars[1]);
      tc::mbar_wait(&bars[2], it & 1);
      tc::tc_fence_after();
      for (int c0 = h * (TMF / 2); c0 < (h + 1) * (TMF / 2); c0 += 16) {
        if (m0 + c0 >= p.M) break;
        float v[16], r[16];
        tc::tmem_ld16(tD + c0, v);
        const bool live = chan < p.Nout;
#pragma unroll
        for (int j = 0; j < 16; ++j) {   // residual rows are fetched while the TMEM load is in flight
          const int64_t row = m0 + c0 + j;
          r[j] = (p.residual && live && row < p.M) ? __ldg(p.residual + row * p.ldr + chan) : 0.0f;
        }
        tc::tmem_wait_ld();
        if (live) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int64_t row = m0 + c0 + j;
            if (row < p.M) p.Y[row * p.ldy + chan] = apply_act(v[j] + bias, p.act) + r[j];
          }
        }
      }
      tc::tc_fence_before();
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, 64);
}

struct DwParams {
  const float* dY;      // [M, Nout]
  int64_t lddy;
  const float* saved_y; // optional [M, Nout]
  int64_t ldys;
  const float* X;       // [M, K]
  int64_t ldx;
  float* partial;       // [gridDim.x][128][K + 16]
  int64_t M;
  int K, Nout;
};

// One CTA's share of a weight-gradient problem: tiles [rank, rank+1) * ntiles / nranks, accumulated in TMEM, written
// as a [128][K + 16] partial block to `part`.
__device__ __forceinline__ void node_dw_body(const DwParams& p, int rank, int nranks, float* __restrict__ part) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bars[2];   // ready (images written), done (MMAs finished)
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = p.K, Nout = p.Nout;
  const int KX = K + 16;                             // + ones column block (bias gradient)
  const uint32_t sbo_y = (Nout >> 3) * 128, sbo_x = (KX >> 3) * 128;
  const uint32_t y_bytes = TMG * sbo_y + 2048, x_bytes = TMG * sbo_x;   // slack: M = 128 view of a 64-channel image
  uint8_t* sYh = smem;
  uint8_t* sYl = sYh + y_bytes;
  uint8_t* sXh = sYl + y_bytes;
  uint8_t* sXl = sXh + x_bytes;

  if (tid == 0) {
    tc::mbar_init(&bars[0], CW * 32);
    tc::mbar_init(&bars[1], 1);
    tc::mbar_fence_init();
  }
  __syncwarp();
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 256);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int64_t ntiles = (p.M + TM - 1) / TM;
  // contiguous tile range per CTA
  const int64_t t0 = (int64_t)rank * ntiles / nranks, t1 = (int64_t)(rank + 1) * ntiles / nranks;

  if (warp == CW) {
    if (lane == 0) {
      const uint32_t aYh = tc::smem_u32(sYh), aYl = tc::smem_u32(sYl), aXh = tc::smem_u32(sXh), aXl = tc::smem_u32(sXl);
      const uint32_t idesc = tc::umma_idesc_f16(128, KX, 1, 1, 1);
      uint32_t it = 0;
      for (int64_t ti = t0; ti < t1; ++ti, ++it) {
        tc::mbar_wait(&bars[0], it & 1);
        tc::tc_fence_after();
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t a = (pass == 2) ? aYl : aYh;
          const uint32_t b = (pass == 1) ? aXl : aXh;
          for (int ks = 0; ks < TM / 16; ++ks)
            // MN-major views: LBO = stride between 8-atom groups, SBO = stride between 8-channel groups
            tc::umma_f16(tmem_base, tc::umma_smem_desc(a + ks * 2 * sbo_y, sbo_y, 128),
                         tc::umma_smem_desc(b + ks * 2 * sbo_x, sbo_x, 128), idesc, (it | pass | ks) != 0);
        }
        tc::umma_commit(&bars[1]);
      }
    }
    __syncwarp();
  } else {
    uint32_t it = 0;
    for (int64_t ti = t0; ti < t1; ++ti, ++it) {
      const int64_t m0 = ti * TM;
      // the previous tile's MMAs must be done reading the images before they are rewritten - but not before the loads of
      // this tile are issued: they are in flight while those MMAs finish
      auto images_free = [&]() {
        if (it > 0) tc::mbar_wait(&bars[1], (it - 1) & 1);
      };
      convert_tile<TM, CW>(p.dY, p.lddy, p.saved_y, p.ldys, m0, p.M, Nout, sbo_y, sYh, sYl, -1, warp, lane, images_free);
      convert_tile<TM>(p.X, p.ldx, nullptr, 0, m0, p.M, K, sbo_x, sXh, sXl, K >> 3, warp, lane);
      tc::fence_proxy_async();
      tc::mbar_arrive(&bars[0]);
    }
    const int wq = warp & 3, h = warp >> 2;
    const int chan = wq * 32 + lane;
    const uint32_t tD = tmem_base + ((uint32_t)(wq * 32) << 16);
    const bool any = t0 < t1;
    if (any) {
      tc::mbar_wait(&bars[1], (it - 1) & 1);
      tc::tc_fence_after();
    }
    // columns [0, KX) split between the two warp halves in 16-column blocks
    const int nblk = KX >> 4;
    for (int blk = h; blk < nblk; blk += 2) {
      float v[16];
      if (any) {
        tc::tmem_ld16(tD + blk * 16, v);
        tc::tmem_wait_ld();
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.0f;
      }
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(part + chan * KX + blk * 16 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, 256);
}

__global__ void __launch_bounds__(NT, 2) node_gemm_dw_kernel(const DwParams p) {
  node_dw_body(p, blockIdx.x, gridDim.x, p.partial + (int64_t)blockIdx.x * 128 * (p.K + 16));
}

// ---- chained node linears: up to three Linear stages on one 64-atom tile without leaving the SM ---------------------
// Stage s:  V_s = act_s(U_s W_s^T + b_s) * (1 - exp(-Ys)/2)  + R_s ,   U_0 = X (global),  U_(s+1) = V_s (on chip).
//   forward  of an interaction-block tail (PyG CFConv.lin2 -> ssp -> InteractionBlock.lin (+ h) -> next block's CFConv.lin1,
//            sns.py:163-164):  agg -> [lin2, ssp] -> y -> [lin, + h] -> h' -> [lin1'] -> x''
//   backward of the same tail: dx'' -> [lin1'^T, + dh'] -> dh -> [lin^T, * ssp'(y)] -> dpre -> [lin2^T] -> dagg
// Every stage is the split-bf16 GEMM of node_gemm_fwd_kernel (hi + lo images, three passes, fp32 accumulate).  The
// epilogue thread that owns output channel k holds V_s[k, atoms] and writes 8 consecutive atoms as one 16-byte chunk of
// the NEXT stage's operand, which is therefore read as an MN-major B operand (atoms contiguous): no transposition.
// `out` of a stage may be null (value only feeds the next stage).  Weights of all stages stay in shared memory.
constexpr int MAXS = 3;
constexpr int CWC = 16;            // compute warps of the chain kernel: 4 per TMEM lane quarter, 16 atoms each
constexpr int NTC = CWC * 32 + 32;
struct ChainStage {
  const uint8_t* w_img;    // hi | lo image, rows = Nout padded to 128, K columns
  const float* bias;       // [Nout] or null
  const float* residual;   // [M, Nout] or null (added last)
  const float* scale_y;    // [M, Nout] or null: saved ShiftedSoftplus OUTPUT y, V *= 1 - exp(-y) / 2  (ssp' on the output side)
  float* out;              // [M, Nout] or null
  int64_t ldr, lds, ldo;
  int K, Nout, act;
};
struct ChainParams {
  const float* X;
  int64_t ldx;
  int64_t M;
  int nstages;
  ChainStage st[MAXS];
};

__global__ void __launch_bounds__(NTC, 1) node_chain_kernel(const __grid_constant__ ChainParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bars[MAXS + 2];   // wbar[s], xready, dready
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t w_off[MAXS + 1];
  w_off[0] = 0;
#pragma unroll
  for (int s = 0; s < MAXS; ++s) w_off[s + 1] = w_off[s] + (s < p.nstages ? 2u * 128u * (uint32_t)p.st[s].K * 2u : 0u);
  uint8_t* sXh = smem + w_off[MAXS];
  uint8_t* sXl = sXh + TMF * MAXC * 2;

  if (tid == 0) {
    for (int s = 0; s < MAXS; ++s) tc::mbar_init(&bars[s], 1);
    tc::mbar_init(&bars[MAXS], CWC * 32);
    tc::mbar_init(&bars[MAXS + 1], 1);
    tc::mbar_fence_init();
  }
  __syncwarp();
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 64);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int64_t ntiles = (p.M + TMF - 1) / TMF;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == CWC) {
    if (lane == 0) {
      for (int s = 0; s < p.nstages; ++s) {
        const uint32_t bytes = w_off[s + 1] - w_off[s];
        tc::mbar_arrive_expect_tx(&bars[s], bytes);
        tc::bulk_g2s(smem + w_off[s], p.st[s].w_img, bytes, &bars[s]);
      }
      const uint32_t aXh = tc::smem_u32(sXh), aXl = tc::smem_u32(sXl);
      uint32_t n = 0;   // operand images consumed so far
      for (int64_t ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
        for (int s = 0; s < p.nstages; ++s, ++n) {
          const int K = p.st[s].K;
          const uint32_t sbo = (uint32_t)(K >> 3) * 128;
          const uint32_t aWh = tc::smem_u32(smem + w_off[s]), aWl = aWh + 128u * (uint32_t)K * 2u;
          if (ti == (int64_t)blockIdx.x) tc::mbar_wait(&bars[s], 0);
          tc::mbar_wait(&bars[MAXS], n & 1);
          tc::tc_fence_after();
          // stage 0 reads the K-major image of convert_tile, later stages the MN-major image the previous epilogue wrote
          const uint32_t idesc = tc::umma_idesc_f16(128, TMF, 1, 0, s == 0 ? 0 : 1);
          const int ks_n = K >> 4;
          for (int pass = 0; pass < 3; ++pass) {
            const uint32_t a = (pass == 2) ? aWl : aWh;
            const uint32_t b = (pass == 1) ? aXl : aXh;
            for (int ks = 0; ks < ks_n; ++ks)
              tc::umma_f16(tmem_base, tc::umma_smem_desc(a + ks * 256, 128, sbo), tc::umma_smem_desc(b + ks * 256, 128, sbo),
                           idesc, (pass | ks) != 0);
          }
          tc::umma_commit(&bars[MAXS + 1]);
        }
      }
    }
    __syncwarp();
  } else {
    const int wq = warp & 3, h = warp >> 2;
    const int chan = wq * 32 + lane;
    const uint32_t tD = tmem_base + ((uint32_t)(wq * 32) << 16);
    uint32_t n = 0;
    for (int64_t ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
      const int64_t m0 = ti * TMF;
      // the last stage of the previous tile was read out of TMEM by every thread before anyone arrives on xready below;
      // its MMAs (the only readers of the images) completed before that read-out
      // (the barrier sits between the global loads of this tile and the first write to the images)
      auto images_free = [&]() {
        if (n > 0) tc::named_bar_sync(1, CWC * 32);
      };
      convert_tile<TMF, CWC>(p.X, p.ldx, nullptr, 0, m0, p.M, p.st[0].K, (uint32_t)(p.st[0].K >> 3) * 128, sXh, sXl, -1, warp, lane,
                             images_free);
      tc::fence_proxy_async();
      tc::mbar_arrive(&bars[MAXS]);
      for (int s = 0; s < p.nstages; ++s, ++n) {
        // stage parameters into registers (p.st[s] with a run-time s would otherwise be re-read from the parameter bank
        // through local memory inside the element loops)
        const int Nout = p.st[s].Nout, act = p.st[s].act;
        const bool live = chan < Nout;
        const float bias = (p.st[s].bias && live) ? __ldg(p.st[s].bias + chan) : 0.0f;
        const bool feeds = s + 1 < p.nstages;
        const uint32_t sbo_n = (uint32_t)(Nout >> 3) * 128;   // K of the next stage = Nout of this one
        const int64_t ldr = p.st[s].ldr, lds = p.st[s].lds, ldo = p.st[s].ldo;
        const float* res_p = p.st[s].residual ? p.st[s].residual + m0 * ldr + chan : nullptr;
        const float* ys_p = p.st[s].scale_y ? p.st[s].scale_y + m0 * lds + chan : nullptr;
        float* out_p = p.st[s].out ? p.st[s].out + m0 * ldo + chan : nullptr;
        const int rows = (int)min((int64_t)TMF, p.M - m0);    // atoms of this tile that exist
        const int c0 = h * 16;             // this warp's 16 atoms of the tile
        // full tile and a real channel: no per-element predicates (the common case: one ragged tile per launch at most)
        const bool full = live && rows == TMF;
        const int ir = (int)ldr, is = (int)lds, io = (int)ldo;   // row strides fit 32 bits (a tile spans 64 rows)
        float v[16], r[16], y[16];
        if (full) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {   // residual / saved-activation rows are fetched while the MMAs of the stage run
            r[j] = res_p ? __ldg(res_p + (c0 + j) * ir) : 0.0f;
            y[j] = ys_p ? __ldg(ys_p + (c0 + j) * is) : 1e30f;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const bool ok = live && (c0 + j) < rows;
            r[j] = (res_p && ok) ? __ldg(res_p + (c0 + j) * ir) : 0.0f;
            y[j] = (ys_p && ok) ? __ldg(ys_p + (c0 + j) * is) : 1e30f;
          }
        }
        tc::mbar_wait(&bars[MAXS + 1], n & 1);
        tc::tc_fence_after();
        tc::tmem_ld16(tD + c0, v);
        tc::tmem_wait_ld();
        if (act == CMP_ACT_SSP) {
          // softplus(x) - ln 2 = max(x, 0) + ln 2 (log2(1 + 2^(-|x| log2 e)) - 1): two MUFU ops, ~1e-6 relative
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float x = v[j] + bias;
            const float t = ex2_approx(-1.4426950408889634f * fabsf(x));
            v[j] = fmaf(tc::fast_lg2(1.0f + t) - 1.0f, kLn2, fmaxf(x, 0.0f));
          }
        } else if (act == CMP_ACT_SILU) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = silu(v[j] + bias);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += bias;
        }
        if (ys_p) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] *= 1.0f - 0.5f * ex2_approx(-1.4426950408889634f * y[j]);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += r[j];
        if (out_p && live) {
          if (full) {
#pragma unroll
            for (int j = 0; j < 16; ++j) out_p[(c0 + j) * io] = v[j];
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < rows) out_p[(c0 + j) * io] = v[j];
          }
        }
        if (feeds && live) {
          // this thread's channel is K row `chan` of the next operand: 8 atoms = one 16-byte chunk (MN-major)
#pragma unroll
          for (int g8 = 0; g8 < 2; ++g8) {
            float hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float x = (full || c0 + g8 * 8 + j < rows) ? v[g8 * 8 + j] : 0.0f;
              hi[j] = __bfloat162float(__float2bfloat16_rn(x));
              lo[j] = x - hi[j];
            }
            const uint32_t off = (uint32_t)chan * 16 + (uint32_t)((c0 >> 3) + g8) * sbo_n;
            *reinterpret_cast<uint4*>(sXh + off) = make_uint4(tc::pack_bf16x2(hi[0], hi[1]), tc::pack_bf16x2(hi[2], hi[3]),
                                                              tc::pack_bf16x2(hi[4], hi[5]), tc::pack_bf16x2(hi[6], hi[7]));
            *reinterpret_cast<uint4*>(sXl + off) = make_uint4(tc::pack_bf16x2(lo[0], lo[1]), tc::pack_bf16x2(lo[2], lo[3]),
                                                              tc::pack_bf16x2(lo[4], lo[5]), tc::pack_bf16x2(lo[6], lo[7]));
          }
        }
        tc::tc_fence_before();
        if (feeds) {
          tc::fence_proxy_async();
          tc::mbar_arrive(&bars[MAXS]);
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, 64);
}

// Two-stream variant (CMP_CHAIN_TWO_STREAMS=1; NOT the default): the 16 compute warps form two groups of 8, each with its own MMA-issuing warp, its
// own operand images (32 atoms) and 32 TMEM columns.  Group g walks tiles 2 * blockIdx.x + g, + 2 * gridDim.x, ... of 32
// atoms; the groups share the weight images and never synchronise with each other, so the global loads / conversions /
// epilogues of one stream overlap the MMA round trips of the other.  Same arithmetic per element: results are
// bit-identical to the one-stream kernel.  Measured at cfg 2: 27.5 us against 26.7 us per launch under ncu, i.e. NO gain
// (profiles/r02_node_chain_kernels.md): the kernel is not bound by the serial phases of a tile but by its instruction
// count (8.7 M warp instructions per launch = ~120 per element: the channel-per-thread epilogue issues one 4-byte access
// per atom for residual / saved activation / output, plus the hi + lo conversions) at 40 % issue utilisation with
// `long_scoreboard` as the dominant stall.  Kept for cross-checking; the next step is an epilogue that moves 16-byte rows.
constexpr int TMC = 32;            // atoms per tile and stream
constexpr int CWG = 8;             // compute warps per stream: 4 TMEM lane quarters x 2 halves of 16 atoms
constexpr int NTC2 = 2 * CWG * 32 + 64;
__global__ void __launch_bounds__(NTC2, 1) node_chain2_kernel(const __grid_constant__ ChainParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bars[MAXS + 4];   // wbar[s] | xready[g] | dready[g]
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t w_off[MAXS + 1];
  w_off[0] = 0;
#pragma unroll
  for (int s = 0; s < MAXS; ++s) w_off[s + 1] = w_off[s] + (s < p.nstages ? 2u * 128u * (uint32_t)p.st[s].K * 2u : 0u);

  if (tid == 0) {
    for (int s = 0; s < MAXS; ++s) tc::mbar_init(&bars[s], 1);
    for (int g = 0; g < 2; ++g) {
      tc::mbar_init(&bars[MAXS + g], CWG * 32);
      tc::mbar_init(&bars[MAXS + 2 + g], 1);
    }
    tc::mbar_fence_init();
  }
  __syncwarp();
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 64);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int64_t ntiles = (p.M + TMC - 1) / TMC;
  pdl_launch_dependents();
  pdl_wait();

  if (warp >= 2 * CWG) {
    const int g = warp - 2 * CWG;
    if (lane == 0) {
      if (g == 0) {
        for (int s = 0; s < p.nstages; ++s) {
          const uint32_t bytes = w_off[s + 1] - w_off[s];
          tc::mbar_arrive_expect_tx(&bars[s], bytes);
          tc::bulk_g2s(smem + w_off[s], p.st[s].w_img, bytes, &bars[s]);
        }
      }
      uint8_t* sXh = smem + w_off[MAXS] + (uint32_t)g * (2u * TMC * MAXC * 2u);
      const uint32_t aXh = tc::smem_u32(sXh), aXl = aXh + TMC * MAXC * 2;
      const uint32_t tD = tmem_base + (uint32_t)g * TMC;
      uint32_t n = 0;   // operand images consumed so far
      bool waited = false;
      for (int64_t ti = 2 * (int64_t)blockIdx.x + g; ti < ntiles; ti += 2 * (int64_t)gridDim.x) {
        for (int s = 0; s < p.nstages; ++s, ++n) {
          const int K = p.st[s].K;
          const uint32_t sbo = (uint32_t)(K >> 3) * 128;
          const uint32_t aWh = tc::smem_u32(smem + w_off[s]), aWl = aWh + 128u * (uint32_t)K * 2u;
          if (!waited) tc::mbar_wait(&bars[s], 0);
          tc::mbar_wait(&bars[MAXS + g], n & 1);
          tc::tc_fence_after();
          const uint32_t idesc = tc::umma_idesc_f16(128, TMC, 1, 0, s == 0 ? 0 : 1);
          const int ks_n = K >> 4;
          for (int pass = 0; pass < 3; ++pass) {
            const uint32_t a = (pass == 2) ? aWl : aWh;
            const uint32_t b = (pass == 1) ? aXl : aXh;
            for (int ks = 0; ks < ks_n; ++ks)
              tc::umma_f16(tD, tc::umma_smem_desc(a + ks * 256, 128, sbo), tc::umma_smem_desc(b + ks * 256, 128, sbo), idesc,
                           (pass | ks) != 0);
          }
          tc::umma_commit(&bars[MAXS + 2 + g]);
        }
        waited = true;
      }
      // the thread that issued the weight copies must not leave before they have landed
      if (g == 0 && !waited)
        for (int s = 0; s < p.nstages; ++s) tc::mbar_wait(&bars[s], 0);
    }
    __syncwarp();
  } else {
    const int g = warp / CWG, wg = warp % CWG;
    const int wq = wg & 3, h = wg >> 2;
    const int chan = wq * 32 + lane;
    uint8_t* sXh = smem + w_off[MAXS] + (uint32_t)g * (2u * TMC * MAXC * 2u);
    uint8_t* sXl = sXh + TMC * MAXC * 2;
    const uint32_t tD = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)g * TMC;
    uint32_t n = 0;
    for (int64_t ti = 2 * (int64_t)blockIdx.x + g; ti < ntiles; ti += 2 * (int64_t)gridDim.x) {
      const int64_t m0 = ti * TMC;
      // as in the one-stream kernel: the previous tile's last read-out of TMEM (after its last MMAs) precedes this barrier
      auto images_free = [&]() {
        if (n > 0) tc::named_bar_sync(1 + g, CWG * 32);
      };
      convert_tile<TMC, CWG>(p.X, p.ldx, nullptr, 0, m0, p.M, p.st[0].K, (uint32_t)(p.st[0].K >> 3) * 128, sXh, sXl, -1, wg, lane,
                             images_free);
      tc::fence_proxy_async();
      tc::mbar_arrive(&bars[MAXS + g]);
      for (int s = 0; s < p.nstages; ++s, ++n) {
        const int Nout = p.st[s].Nout, act = p.st[s].act;
        const bool live = chan < Nout;
        const float bias = (p.st[s].bias && live) ? __ldg(p.st[s].bias + chan) : 0.0f;
        const bool feeds = s + 1 < p.nstages;
        const uint32_t sbo_n = (uint32_t)(Nout >> 3) * 128;   // K of the next stage = Nout of this one
        const int64_t ldr = p.st[s].ldr, lds = p.st[s].lds, ldo = p.st[s].ldo;
        const float* res_p = p.st[s].residual ? p.st[s].residual + m0 * ldr + chan : nullptr;
        const float* ys_p = p.st[s].scale_y ? p.st[s].scale_y + m0 * lds + chan : nullptr;
        float* out_p = p.st[s].out ? p.st[s].out + m0 * ldo + chan : nullptr;
        const int rows = (int)min((int64_t)TMC, p.M - m0);    // atoms of this tile that exist
        const int c0 = h * 16;             // this warp's 16 atoms of the tile
        const bool full = live && rows == TMC;
        const int ir = (int)ldr, is = (int)lds, io = (int)ldo;
        float v[16], r[16], y[16];
        if (full) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            r[j] = res_p ? __ldg(res_p + (c0 + j) * ir) : 0.0f;
            y[j] = ys_p ? __ldg(ys_p + (c0 + j) * is) : 1e30f;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const bool ok = live && (c0 + j) < rows;
            r[j] = (res_p && ok) ? __ldg(res_p + (c0 + j) * ir) : 0.0f;
            y[j] = (ys_p && ok) ? __ldg(ys_p + (c0 + j) * is) : 1e30f;
          }
        }
        tc::mbar_wait(&bars[MAXS + 2 + g], n & 1);
        tc::tc_fence_after();
        tc::tmem_ld16(tD + c0, v);
        tc::tmem_wait_ld();
        if (act == CMP_ACT_SSP) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float x = v[j] + bias;
            const float t = ex2_approx(-1.4426950408889634f * fabsf(x));
            v[j] = fmaf(tc::fast_lg2(1.0f + t) - 1.0f, kLn2, fmaxf(x, 0.0f));
          }
        } else if (act == CMP_ACT_SILU) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = silu(v[j] + bias);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += bias;
        }
        if (ys_p) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] *= 1.0f - 0.5f * ex2_approx(-1.4426950408889634f * y[j]);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += r[j];
        if (out_p && live) {
          if (full) {
#pragma unroll
            for (int j = 0; j < 16; ++j) out_p[(c0 + j) * io] = v[j];
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < rows) out_p[(c0 + j) * io] = v[j];
          }
        }
        if (feeds && live) {
#pragma unroll
          for (int g8 = 0; g8 < 2; ++g8) {
            float hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float x = (full || c0 + g8 * 8 + j < rows) ? v[g8 * 8 + j] : 0.0f;
              hi[j] = __bfloat162float(__float2bfloat16_rn(x));
              lo[j] = x - hi[j];
            }
            const uint32_t off = (uint32_t)chan * 16 + (uint32_t)((c0 >> 3) + g8) * sbo_n;
            *reinterpret_cast<uint4*>(sXh + off) = make_uint4(tc::pack_bf16x2(hi[0], hi[1]), tc::pack_bf16x2(hi[2], hi[3]),
                                                              tc::pack_bf16x2(hi[4], hi[5]), tc::pack_bf16x2(hi[6], hi[7]));
            *reinterpret_cast<uint4*>(sXl + off) = make_uint4(tc::pack_bf16x2(lo[0], lo[1]), tc::pack_bf16x2(lo[2], lo[3]),
                                                              tc::pack_bf16x2(lo[4], lo[5]), tc::pack_bf16x2(lo[6], lo[7]));
          }
        }
        tc::tc_fence_before();
        if (feeds) {
          tc::fence_proxy_async();
          tc::mbar_arrive(&bars[MAXS + g]);
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, 64);
}

// ---- grouped launch: many weight-gradient problems share one grid (the node linears of a whole backward pass) ----
constexpr int MAX_GROUP = 32;
constexpr int PART_STRIDE = 128 * (MAXC + 16);      // floats per CTA in the grouped workspace

struct DwGroupEntry {
  DwParams p;        // p.partial unused
  float* dW;
  float* db;
  int64_t lddw;      // leading dimension of dW (a block of a wider weight gradient)
  int cta_begin, cta_count;
};

struct DwGroup {
  DwGroupEntry e[MAX_GROUP];
  float* partial;    // [gridDim.x][PART_STRIDE]
  int count;
};

__global__ void __launch_bounds__(NT, 2) node_gemm_dw_grouped_kernel(const __grid_constant__ DwGroup g) {
  int idx = 0;
  for (int i = 1; i < g.count; ++i)
    if ((int)blockIdx.x >= g.e[i].cta_begin) idx = i;
  const DwGroupEntry& en = g.e[idx];
  node_dw_body(en.p, (int)blockIdx.x - en.cta_begin, en.cta_count, g.partial + (int64_t)blockIdx.x * PART_STRIDE);
}

// dW / db of problem blockIdx.y = sum of its CTAs' partial blocks in CTA order (fixed order: deterministic)
__global__ void node_dw_grouped_reduce_kernel(const __grid_constant__ DwGroup g) {
  const DwGroupEntry& en = g.e[blockIdx.y];
  const int K = en.p.K, Nout = en.p.Nout, KX = K + 16;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Nout * (K + 1)) return;
  const int n = i / (K + 1), k = i % (K + 1);
  const float* src = g.partial + (int64_t)en.cta_begin * PART_STRIDE + n * KX + k;
  float s = 0.0f;
  for (int c = 0; c < en.cta_count; ++c) s += src[(int64_t)c * PART_STRIDE];
  if (k < K) en.dW[(int64_t)n * en.lddw + k] = s;
  else if (en.db) en.db[n] = s;
}

__global__ void node_dw_reduce_kernel(const float* __restrict__ partial, int P, int K, int Nout, float* __restrict__ dW,
                                      float* __restrict__ db) {
  // block = 32 outputs x 8 segments of the partial list: each thread sums its contiguous segment in order, the 8
  // segment sums are combined in segment order (fixed summation tree: deterministic)
  __shared__ float seg[8][33];
  const int KX = K + 16;
  const int i = blockIdx.x * 32 + threadIdx.x;   // over Nout * (K + 1)
  const bool live = i < Nout * (K + 1);
  const int n = live ? i / (K + 1) : 0, k = live ? i % (K + 1) : 0;
  const int per = (P + 7) / 8;
  const int c_lo = threadIdx.y * per, c_hi = min(P, c_lo + per);
  float s = 0.0f;
  if (live)
    for (int c = c_lo; c < c_hi; ++c) s += partial[((int64_t)c * 128 + n) * KX + k];
  seg[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y != 0 || !live) return;
  s = 0.0f;
#pragma unroll
  for (int y = 0; y < 8; ++y) s += seg[y][threadIdx.x];
  if (k < K) dW[n * K + k] = s;
  else if (db) db[n] = s;
}

// W[rows, cols] (ld) fp32 -> hi | lo K-major images with 128 rows (zero padded): rows = MMA M.
// transpose = 1 packs W^T (rows of the image = columns of W).
__device__ __forceinline__ void node_pack_weight_body(const float* __restrict__ W, int rows, int cols, int transpose,
                                                      uint8_t* __restrict__ out, int idx) {
  if (transpose == 2) {   // both images in one launch: [hi|lo of W] followed by [hi|lo of W^T]
    const int n_norm = 128 * cols;
    if (idx < n_norm) {
      transpose = 0;
    } else {
      transpose = 1;
      idx -= n_norm;
      out += (size_t)2 * n_norm * 2;
    }
  }
  const int R = transpose ? cols : rows;   // image rows that carry data
  const int C = transpose ? rows : cols;   // image K extent
  if (idx >= 128 * C) return;
  const int r = idx / C, c = idx % C;
  float v = 0.0f;
  if (r < R) v = transpose ? W[c * cols + r] : W[r * cols + c];
  const float h = __bfloat162float(__float2bfloat16_rn(v));
  const uint32_t off = (r & 7) * 16 + (c & 7) * 2 + (r >> 3) * ((C >> 3) * 128) + (c >> 3) * 128;
  *reinterpret_cast<__nv_bfloat16*>(out + off) = __float2bfloat16_rn(h);
  *reinterpret_cast<__nv_bfloat16*>(out + 128 * C * 2 + off) = __float2bfloat16_rn(v - h);
}

__global__ void node_pack_weight_kernel(const float* __restrict__ W, int rows, int cols, int transpose,
                                        uint8_t* __restrict__ out) {
  node_pack_weight_body(W, rows, cols, transpose, out, blockIdx.x * blockDim.x + threadIdx.x);
}

// grouped: job blockIdx.y, both images of every weight (the node linears of a model in one launch)
struct PackNodeJob {
  const float* W;
  int32_t rows, cols;
  uint8_t* packed;
};
struct PackNodeGroup {
  PackNodeJob j[MAX_GROUP];
};
__global__ void node_pack_weight_grouped_kernel(const __grid_constant__ PackNodeGroup g) {
  const PackNodeJob& j = g.j[blockIdx.y];
  node_pack_weight_body(j.W, j.rows, j.cols, 2, j.packed, blockIdx.x * blockDim.x + threadIdx.x);
}

bool dims_ok(int K, int Nout) {
  return K >= 16 && K <= MAXC && K % 16 == 0 && Nout >= 16 && Nout <= MAXC && Nout % 16 == 0;
}

}  // namespace
}  // namespace cmp

using namespace cmp;

extern "C" int cmp_node_gemm_tc_supported(int K, int Nout) { return dims_ok(K, Nout) ? 1 : 0; }

extern "C" size_t cmp_node_gemm_weight_bytes(int image_K) { return (size_t)2 * 128 * image_K * 2; }

extern "C" int cmp_node_gemm_pack_weight(const float* W, int rows, int cols, int transpose, void* packed,
                                         cmp_stream_t stream) {
  CMP_REQUIRE(W && packed, CMP_EINVAL, "cmp_node_gemm_pack_weight: null pointer");
  CMP_REQUIRE(dims_ok(cols, rows), CMP_EUNSUPPORTED, "cmp_node_gemm_pack_weight: dims must be multiples of 16 in [16,128]");
  CMP_REQUIRE(transpose >= 0 && transpose <= 2, CMP_EINVAL, "cmp_node_gemm_pack_weight: transpose must be 0, 1 or 2 (both)");
  const int C = transpose == 2 ? rows + cols : (transpose ? rows : cols);
  node_pack_weight_kernel<<<(128 * C + 255) / 256, 256, 0, as_stream(stream)>>>(W, rows, cols, transpose,
                                                                                reinterpret_cast<uint8_t*>(packed));
  CMP_LAUNCH_CHECK("cmp_node_gemm_pack_weight");
  return CMP_OK;
}

extern "C" int cmp_node_gemm_fwd(const float* X, int64_t ldx, const float* saved_y, int64_t ldys, const void* w_img,
                                 const float* bias, int act, const float* residual, int64_t ldr, float* Y, int64_t ldy,
                                 int64_t M, int K, int Nout, cmp_stream_t stream) {
  CMP_REQUIRE(M >= 0, CMP_EINVAL, "cmp_node_gemm_fwd: negative size");
  if (M == 0) return CMP_OK;
  CMP_REQUIRE(dims_ok(K, Nout), CMP_EUNSUPPORTED, "cmp_node_gemm_fwd: K / Nout must be multiples of 16 in [16,128]");
  CMP_REQUIRE(X && w_img && Y, CMP_EINVAL, "cmp_node_gemm_fwd: null pointer");
  CMP_REQUIRE(ldx % 4 == 0 && (uintptr_t)X % 16 == 0 && (uintptr_t)w_img % 16 == 0 &&
                  (!saved_y || (ldys % 4 == 0 && (uintptr_t)saved_y % 16 == 0)),
              CMP_EINVAL, "cmp_node_gemm_fwd: operands must be 16-byte aligned with ld % 4 == 0");
  CMP_REQUIRE(act == CMP_ACT_NONE || act == CMP_ACT_SSP || act == CMP_ACT_SILU, CMP_EINVAL, "cmp_node_gemm_fwd: bad act");
  CMP_REQUIRE(cmp_device_is_sm100(), CMP_EUNSUPPORTED, "cmp_node_gemm_fwd: needs an sm_100 device (tcgen05)");
  const size_t smem = (size_t)2 * 128 * K * 2 + (size_t)2 * TMF * K * 2;
  if (cudaFuncSetAttribute(node_gemm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           2 * 128 * MAXC * 2 + 2 * TMF * MAXC * 2) !=
      cudaSuccess) {
    (void)cudaGetLastError();
    set_error("cmp_node_gemm_fwd: cannot opt in to shared memory");
    return CMP_ECUDA;
  }
  FwdParams p{X, ldx, saved_y, ldys, reinterpret_cast<const uint8_t*>(w_img), bias, residual, ldr, Y, ldy, M, K, Nout, act};
  const int64_t ntiles = ceil_div(M, TMF);
  const int grid = (int)(ntiles < 2 * sm_count() ? ntiles : 2 * sm_count());
  CMP_REQUIRE(launch_pdl(node_gemm_fwd_kernel, dim3(grid), dim3(NT), smem, as_stream(stream), p) == cudaSuccess, CMP_ECUDA,
              "cmp_node_gemm_fwd: launch failed");
  CMP_LAUNCH_CHECK("cmp_node_gemm_fwd");
  return CMP_OK;
}

extern "C" int cmp_node_chain_max_stages(void) { return MAXS; }

extern "C" int cmp_node_chain_fwd(const float* X, int64_t ldx, int64_t M, const void* stages, int nstages,
                                  cmp_stream_t stream) {
  CMP_REQUIRE(M >= 0 && nstages >= 1 && nstages <= MAXS, CMP_EINVAL, "cmp_node_chain_fwd: bad size");
  if (M == 0) return CMP_OK;
  CMP_REQUIRE(X && stages, CMP_EINVAL, "cmp_node_chain_fwd: null pointer");
  CMP_REQUIRE(ldx % 4 == 0 && (uintptr_t)X % 16 == 0, CMP_EINVAL, "cmp_node_chain_fwd: X must be 16-byte aligned, ldx % 4 == 0");
  CMP_REQUIRE(cmp_device_is_sm100(), CMP_EUNSUPPORTED, "cmp_node_chain_fwd: needs an sm_100 device (tcgen05)");
  const cmp_chain_stage_t* in = reinterpret_cast<const cmp_chain_stage_t*>(stages);
  ChainParams p;
  p.X = X;
  p.ldx = ldx;
  p.M = M;
  p.nstages = nstages;
  size_t w_bytes = 0;
  for (int s = 0; s < nstages; ++s) {
    CMP_REQUIRE(dims_ok(in[s].K, in[s].Nout), CMP_EUNSUPPORTED,
                "cmp_node_chain_fwd: K / Nout must be multiples of 16 in [16,128]");
    CMP_REQUIRE(in[s].w_img && (uintptr_t)in[s].w_img % 16 == 0, CMP_EINVAL, "cmp_node_chain_fwd: weight image missing / misaligned");
    CMP_REQUIRE(s == 0 || in[s].K == in[s - 1].Nout, CMP_EINVAL, "cmp_node_chain_fwd: stage K must equal the previous Nout");
    CMP_REQUIRE(in[s].act == CMP_ACT_NONE || in[s].act == CMP_ACT_SSP || in[s].act == CMP_ACT_SILU, CMP_EINVAL,
                "cmp_node_chain_fwd: bad act");
    ChainStage& d = p.st[s];
    d.w_img = reinterpret_cast<const uint8_t*>(in[s].w_img);
    d.bias = in[s].bias;
    d.residual = in[s].residual;
    d.scale_y = in[s].scale_y;
    d.out = in[s].out;
    d.ldr = in[s].ldr;
    d.lds = in[s].lds;
    d.ldo = in[s].ldo;
    d.K = in[s].K;
    d.Nout = in[s].Nout;
    d.act = in[s].act;
    w_bytes += (size_t)2 * 128 * in[s].K * 2;
  }
  CMP_REQUIRE(p.st[nstages - 1].out != nullptr, CMP_EINVAL, "cmp_node_chain_fwd: the last stage needs an output");
  for (int s = nstages; s < MAXS; ++s) p.st[s] = ChainStage{};
  const size_t smem = w_bytes + (size_t)2 * TMF * MAXC * 2;       // = 2 streams x 2 images x TMC atoms for the two-stream kernel
  static_assert(2 * TMC == TMF, "both chain kernels use the same operand-image bytes");
  static bool attr_set = false;
  if (!attr_set) {
    const int max_smem = MAXS * 2 * 128 * MAXC * 2 + 2 * TMF * MAXC * 2;
    if (cudaFuncSetAttribute(node_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem) != cudaSuccess ||
        cudaFuncSetAttribute(node_chain2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem) != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("cmp_node_chain_fwd: cannot opt in to shared memory");
      return CMP_ECUDA;
    }
    attr_set = true;
  }
  // CMP_CHAIN_TWO_STREAMS=1: the two-stream kernel (two independent streams of 32-atom tiles per CTA), same results
  static const bool two_streams = [] { const char* e = getenv("CMP_CHAIN_TWO_STREAMS"); return e && e[0] && e[0] != '0'; }();
  if (!two_streams) {
    const int64_t ntiles = ceil_div(M, TMF);
    const int grid = (int)(ntiles < sm_count() ? ntiles : sm_count());
    CMP_REQUIRE(launch_pdl(node_chain_kernel, dim3(grid), dim3(NTC), smem, as_stream(stream), p) == cudaSuccess, CMP_ECUDA,
                "cmp_node_chain_fwd: launch failed");
  } else {
    const int64_t npairs = ceil_div(ceil_div(M, TMC), 2);
    const int grid = (int)(npairs < sm_count() ? npairs : sm_count());
    CMP_REQUIRE(launch_pdl(node_chain2_kernel, dim3(grid), dim3(NTC2), smem, as_stream(stream), p) == cudaSuccess, CMP_ECUDA,
                "cmp_node_chain_fwd: launch failed");
  }
  CMP_LAUNCH_CHECK("cmp_node_chain_fwd");
  return CMP_OK;
}

extern "C" size_t cmp_node_gemm_dw_workspace(int K) { return align_up((size_t)sm_count() * 128 * (K + 16) * sizeof(float), 256); }

extern "C" int cmp_node_gemm_dw(const float* dY, int64_t lddy, const float* saved_y, int64_t ldys, const float* X,
                                int64_t ldx, int64_t M, int K, int Nout, float* dW, float* db, void* workspace,
                                size_t workspace_bytes, cmp_stream_t stream) {
  CMP_REQUIRE(M >= 0, CMP_EINVAL, "cmp_node_gemm_dw: negative size");
  CMP_REQUIRE(dims_ok(K, Nout), CMP_EUNSUPPORTED, "cmp_node_gemm_dw: K / Nout must be multiples of 16 in [16,128]");
  CMP_REQUIRE(dW && (M == 0 || (dY && X)), CMP_EINVAL, "cmp_node_gemm_dw: null pointer");
  CMP_REQUIRE(lddy % 4 == 0 && ldx % 4 == 0 && (uintptr_t)dY % 16 == 0 && (uintptr_t)X % 16 == 0 &&
                  (!saved_y || (ldys % 4 == 0 && (uintptr_t)saved_y % 16 == 0)),
              CMP_EINVAL, "cmp_node_gemm_dw: operands must be 16-byte aligned with ld % 4 == 0");
  CMP_REQUIRE(workspace && workspace_bytes >= cmp_node_gemm_dw_workspace(K), CMP_EWORKSPACE,
              "cmp_node_gemm_dw: workspace too small");
  CMP_REQUIRE(cmp_device_is_sm100(), CMP_EUNSUPPORTED, "cmp_node_gemm_dw: needs an sm_100 device (tcgen05)");
  const int KX = K + 16;
  const size_t smem = (size_t)2 * (TMG * (Nout / 8) * 128 + 2048) + (size_t)2 * TMG * (KX / 8) * 128;
  if (cudaFuncSetAttribute(node_gemm_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           2 * (TMG * (MAXC / 8) * 128 + 2048) + 2 * TMG * ((MAXC + 16) / 8) * 128) != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("cmp_node_gemm_dw: cannot opt in to shared memory");
    return CMP_ECUDA;
  }
  DwParams p{dY, lddy, saved_y, ldys, X, ldx, reinterpret_cast<float*>(workspace), M, K, Nout};
  const int64_t ntiles = ceil_div(M, TM);
  int grid = (int)(ntiles < sm_count() ? ntiles : sm_count());
  if (grid < 1) grid = 1;
  node_gemm_dw_kernel<<<grid, NT, smem, as_stream(stream)>>>(p);
  CMP_LAUNCH_CHECK("cmp_node_gemm_dw");
  node_dw_reduce_kernel<<<(Nout * (K + 1) + 31) / 32, dim3(32, 8), 0, as_stream(stream)>>>(p.partial, grid, K, Nout, dW,
                                                                                          db);
  CMP_LAUNCH_CHECK("cmp_node_gemm_dw(reduce)");
  return CMP_OK;
}

// ---- grouped weight gradients ------------------------------------------------------------------------
// Mirrors cmp_dw_problem_t of include/conanmp.h.
struct cmp_dw_problem_host {
  const float* dY;
  int64_t lddy;
  const float* saved_y;
  int64_t ldys;
  const float* X;
  int64_t ldx;
  int64_t M;
  int32_t K;
  int32_t Nout;
  float* dW;
  int64_t lddw;
  float* db;
};

extern "C" int cmp_node_gemm_dw_group_max(void) { return MAX_GROUP; }

extern "C" size_t cmp_node_gemm_dw_grouped_workspace(void) {
  return align_up((size_t)2 * sm_count() * PART_STRIDE * sizeof(float), 256);
}

extern "C" int cmp_node_gemm_dw_grouped(const void* problems, int count, void* workspace, size_t workspace_bytes,
                                        cmp_stream_t stream) {
  CMP_REQUIRE(count >= 0 && count <= MAX_GROUP, CMP_EINVAL, "cmp_node_gemm_dw_grouped: count must be in [0, %d]", MAX_GROUP);
  if (count == 0) return CMP_OK;
  CMP_REQUIRE(problems, CMP_EINVAL, "cmp_node_gemm_dw_grouped: null pointer");
  CMP_REQUIRE(workspace && workspace_bytes >= cmp_node_gemm_dw_grouped_workspace(), CMP_EWORKSPACE,
              "cmp_node_gemm_dw_grouped: workspace too small");
  CMP_REQUIRE(cmp_device_is_sm100(), CMP_EUNSUPPORTED, "cmp_node_gemm_dw_grouped: needs an sm_100 device (tcgen05)");
  const cmp_dw_problem_host* pr = reinterpret_cast<const cmp_dw_problem_host*>(problems);
  const int G = 2 * sm_count();   // two resident CTAs per SM
  CMP_REQUIRE(count <= G, CMP_EUNSUPPORTED, "cmp_node_gemm_dw_grouped: more problems than CTA slots");
  DwGroup g;
  int64_t tiles[MAX_GROUP];
  int64_t total = 0;
  for (int i = 0; i < count; ++i) {
    const cmp_dw_problem_host& q = pr[i];
    CMP_REQUIRE(q.M >= 0, CMP_EINVAL, "cmp_node_gemm_dw_grouped: negative size");
    CMP_REQUIRE(dims_ok(q.K, q.Nout), CMP_EUNSUPPORTED,
                "cmp_node_gemm_dw_grouped: K / Nout must be multiples of 16 in [16,128]");
    CMP_REQUIRE(q.dW && (q.M == 0 || (q.dY && q.X)), CMP_EINVAL, "cmp_node_gemm_dw_grouped: null pointer");
    CMP_REQUIRE(q.lddy % 4 == 0 && q.ldx % 4 == 0 && (uintptr_t)q.dY % 16 == 0 && (uintptr_t)q.X % 16 == 0 &&
                    (!q.saved_y || (q.ldys % 4 == 0 && (uintptr_t)q.saved_y % 16 == 0)),
                CMP_EINVAL, "cmp_node_gemm_dw_grouped: operands must be 16-byte aligned with ld % 4 == 0");
    tiles[i] = q.M > 0 ? ceil_div(q.M, (int64_t)TM) : 1;
    total += tiles[i];
    g.e[i].p = DwParams{q.dY, q.lddy, q.saved_y, q.ldys, q.X, q.ldx, nullptr, q.M, q.K, q.Nout};
    CMP_REQUIRE(q.lddw >= q.K, CMP_EINVAL, "cmp_node_gemm_dw_grouped: lddw must be >= K");
    g.e[i].dW = q.dW;
    g.e[i].db = q.db;
    g.e[i].lddw = q.lddw;
  }
  // CTAs per problem: proportional to its tiles, at least one, never more than its tiles; spare CTAs go to the
  // problems with the most tiles per CTA
  int cta[MAX_GROUP];
  int used = 0;
  for (int i = 0; i < count; ++i) {
    int64_t c = (int64_t)G * tiles[i] / total;
    if (c < 1) c = 1;
    if (c > tiles[i]) c = tiles[i];
    cta[i] = (int)c;
    used += cta[i];
  }
  while (used > G) {
    int big = 0;
    for (int i = 1; i < count; ++i)
      if (cta[i] > cta[big]) big = i;
    --cta[big];
    --used;
  }
  while (used < G) {
    int best = -1;
    double load = 1.0;   // only problems with more than one tile per CTA can use another CTA
    for (int i = 0; i < count; ++i) {
      const double l = (double)tiles[i] / cta[i];
      if (cta[i] < tiles[i] && l > load) {
        load = l;
        best = i;
      }
    }
    if (best < 0) break;
    ++cta[best];
    ++used;
  }
  int begin = 0;
  for (int i = 0; i < count; ++i) {
    g.e[i].cta_begin = begin;
    g.e[i].cta_count = cta[i];
    begin += cta[i];
  }
  g.partial = reinterpret_cast<float*>(workspace);
  g.count = count;
  static bool attr_set = false;
  const int smem = 2 * (TMG * (MAXC / 8) * 128 + 2048) + 2 * TMG * ((MAXC + 16) / 8) * 128;
  if (!attr_set) {
    if (cudaFuncSetAttribute(node_gemm_dw_grouped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
        cudaSuccess) {
      (void)cudaGetLastError();
      set_error("cmp_node_gemm_dw_grouped: cannot opt in to shared memory");
      return CMP_ECUDA;
    }
    attr_set = true;
  }
  cudaStream_t st = as_stream(stream);
  node_gemm_dw_grouped_kernel<<<begin, NT, smem, st>>>(g);
  CMP_LAUNCH_CHECK("cmp_node_gemm_dw_grouped");
  node_dw_grouped_reduce_kernel<<<dim3((MAXC * (MAXC + 1) + 127) / 128, count), 128, 0, st>>>(g);
  CMP_LAUNCH_CHECK("cmp_node_gemm_dw_grouped(reduce)");
  return CMP_OK;
}

extern "C" int cmp_node_gemm_pack_weights_grouped(const void* jobs, int count, cmp_stream_t stream) {
  CMP_REQUIRE(count >= 0 && count <= MAX_GROUP, CMP_EINVAL, "cmp_node_gemm_pack_weights_grouped: count must be in [0, %d]",
              MAX_GROUP);
  if (count == 0) return CMP_OK;
  CMP_REQUIRE(jobs, CMP_EINVAL, "cmp_node_gemm_pack_weights_grouped: null pointer");
  const PackNodeJob* in = reinterpret_cast<const PackNodeJob*>(jobs);
  PackNodeGroup g;
  for (int i = 0; i < count; ++i) {
    CMP_REQUIRE(in[i].W && in[i].packed, CMP_EINVAL, "cmp_node_gemm_pack_weights_grouped: null pointer");
    CMP_REQUIRE(dims_ok(in[i].cols, in[i].rows), CMP_EUNSUPPORTED,
                "cmp_node_gemm_pack_weights_grouped: dims must be multiples of 16 in [16,128]");
    g.j[i] = in[i];
  }
  node_pack_weight_grouped_kernel<<<dim3((128 * 2 * MAXC + 255) / 256, count), 256, 0, as_stream(stream)>>>(g);
  CMP_LAUNCH_CHECK("cmp_node_gemm_pack_weights_grouped");
  return CMP_OK;
}
