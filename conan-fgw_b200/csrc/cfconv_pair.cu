// Fused CFConv over UNDIRECTED PAIRS for small conformers (<= 30 atoms), sm_100a.
//
// The filter of PyG's CFConv (W_ij = (W2 ssp(W1 rbf(d_ij) + b1) + b2) C(d_ij)) depends on the distance
// only, so both directions of a pair share it.  This kernel evaluates the filter MLP ONCE per pair on tcgen05 / TMEM
// (half the Gaussian expansions, MMAs and softplus evaluations of the per-edge kernel in cfconv_tc.cu) and applies it to
// both directions:
//       agg[i] += W_p * x[j]          (the primary edge j -> i of cmp_build_pair_list)
//       agg[j] += W_p * x[i]          (its reverse, when it exists)
// Work unit of a CTA = one conformer: its x rows are staged in shared memory once (1-D TMA bulk copy), its pair list is
// cut into 128-column tiles that the CTA's NG pipelines take round-robin (a 27-atom conformer = 351 pairs = one
// full round of three tiles), and every pipeline accumulates into its OWN
// [atoms, 128] fp32 accumulator in shared memory (the thread that owns channel f is the only one that touches column f:
// no atomics).  When the conformer is finished the accumulators are summed in pipeline order and the rows written with
// plain stores: deterministic.  Conformers with more than NCAP atoms are left to cfconv_tc.cu (tile list filtered by
// cmp_build_tiles_min_atoms).
//
// With `transposed` the roles of the two directions swap (d x'[j] += W_p g[i]; d x'[i] += W_p g[j] when the reverse
// exists): the backward pass with respect to x' is this kernel again.
//
// Orientation as in cfconv_tc.cu: filter channels on the 128 TMEM lanes, pairs on the columns,
//   D1[128, p] = W1aug[128, 64] * rbf_aug[64, p],   D2[128, p] = W2aug[128, 144] * a'[144, p],  a' = C ssp(D1).
#include "common.cuh"
#include "tc_common.cuh"

namespace cmp {
namespace {

constexpr int F = 128;
constexpr int TE = 128;           // pair columns per tile = UMMA N
constexpr int K1 = 64;            // Gaussians padded (+ bias column)
constexpr int K2 = 144;           // hidden channels + (cutoff, bias) row, padded to 16
constexpr int NCAP = 30;          // atoms of a conformer held in shared memory
constexpr int NG = 3;             // pipelines per CTA
constexpr int GT = 128;           // compute threads per pipeline: one per filter channel (= TMEM lane)
constexpr int NCOMP = NG * GT;
constexpr int CTA_THREADS = NCOMP + NG * 32;
constexpr int CTA_BAR = 1 + NG;   // named barrier of all compute threads

constexpr uint32_t W1_BYTES = F * K1 * 2;         // 16384
constexpr uint32_t W2_BYTES = F * K2 * 2;         // 36864
constexpr uint32_t B1_SBO = (K1 / 8) * 128;       // 1024
constexpr uint32_t A2_SBO = (K2 / 8) * 128;       // 2304
constexpr uint32_t B2_BYTES = K2 * TE * 2;        // 18432 (the rbf image, 8192 B, aliases its head)
constexpr uint32_t XS_BYTES = NCAP * F * 4;       // 16384
constexpr uint32_t OFF_ACC = B2_BYTES;
constexpr uint32_t OFF_OJ = OFF_ACC + XS_BYTES;   // int[64]   x / accumulator offset (floats) of the pair's source j
constexpr uint32_t OFF_OI = OFF_OJ + TE * 4;      // int[64]   ... of its target i
constexpr uint32_t OFF_FLAGS = OFF_OI + TE * 4;   // uint32[8]: row-end masks (32 columns per word), then unpaired masks
constexpr uint32_t OFF_CH = OFF_FLAGS + 32;       // half[64]  cosine cutoff again, as f16 (packed epilogue 1)
constexpr uint32_t GROUP_BYTES = OFF_CH + TE * 2;
constexpr uint32_t SMEM_BYTES = W1_BYTES + W2_BYTES + XS_BYTES + NG * GROUP_BYTES;
static_assert(GROUP_BYTES % 16 == 0, "pipeline blocks must stay 16-byte aligned");
static_assert(SMEM_BYTES <= 232448 - 2048, "shared memory budget (dynamic + ~2 KB static)");
static_assert(GT % TE == 0 && TE % 32 == 0 && NG * TE <= 512, "tile geometry");

struct PairParams {
  const float* x;                  // [N, F]  x' (forward) or dL/dagg (transposed pass)
  const int32_t* seg_ptr;          // [G + 1] first atom of every conformer
  const int32_t* conf_pair_ptr;    // [G + 1] first primary edge of every conformer
  const int32_t* psrc;
  const int32_t* pdst;
  const float* pdist;
  const int32_t* prev;
  const uint8_t* weights;          // W1 image followed by W2 image (cmp_cfconv_tc_pack_weights)
  const float* offset;
  float* out;                      // [N, F]
  float coeff_log2e;
  float cutoff;
  int Ng;
  int G;
  int transposed;
  long long* dbg;                  // optional clock64 timeline of CTA 0, pipeline 0 (10 stamps per tile, 24 tiles)
};

__global__ void __launch_bounds__(CTA_THREADS, 1) cfconv_pair_kernel(const PairParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bars[2 + NG * 4];  // wbar, xbar | per pipeline: b1ready, d1ready, b2ready, d2ready
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_offset[K1];
  __shared__ __align__(16) float s_c2[K1];

  uint8_t* sW1 = smem;
  uint8_t* sW2 = smem + W1_BYTES;
  float* sXs = reinterpret_cast<float*>(smem + W1_BYTES + W2_BYTES);
  uint8_t* groups = smem + W1_BYTES + W2_BYTES + XS_BYTES;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    tc::mbar_init(&bars[0], 1);
    tc::mbar_init(&bars[1], 1);
    for (int g = 0; g < NG; ++g) {
      tc::mbar_init(&bars[2 + g * 4 + 0], GT);
      tc::mbar_init(&bars[2 + g * 4 + 1], 1);
      tc::mbar_init(&bars[2 + g * 4 + 2], GT);
      tc::mbar_init(&bars[2 + g * 4 + 3], 1);
    }
    tc::mbar_fence_init();
  }
  if (tid < K1) {
    s_offset[tid] = (tid < p.Ng) ? p.offset[tid] : 0.0f;
    s_c2[tid] = (tid < p.Ng) ? p.coeff_log2e : 0.0f;
  }
  __syncwarp();
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int k1steps = (p.Ng + 1 + 15) >> 4;
  const int G = p.G;

  if (warp >= NCOMP / 32) {
    // ======================= MMA-issuing warp of pipeline g =======================
    const int g = warp - NCOMP / 32;
    if (lane == 0) {
      uint64_t* wbar = &bars[0];
      uint64_t* b1ready = &bars[2 + g * 4 + 0];
      uint64_t* d1ready = &bars[2 + g * 4 + 1];
      uint64_t* b2ready = &bars[2 + g * 4 + 2];
      uint64_t* d2ready = &bars[2 + g * 4 + 3];
      if (g == 0) {
        tc::mbar_arrive_expect_tx(wbar, W1_BYTES + W2_BYTES);
        tc::bulk_g2s(sW1, p.weights, W1_BYTES, wbar);
        tc::bulk_g2s(sW2, p.weights + W1_BYTES, W2_BYTES, wbar);
      }
      const uint32_t d = tmem_base + g * TE;
      const uint32_t aW1 = tc::smem_u32(sW1), aW2 = tc::smem_u32(sW2), aB = tc::smem_u32(groups + g * GROUP_BYTES);
      tc::mbar_wait(wbar, 0);
      uint32_t it = 0;
      for (int conf = blockIdx.x; conf < G; conf += gridDim.x) {
        const int cn = __ldg(p.seg_ptr + conf + 1) - __ldg(p.seg_ptr + conf);
        const int np = __ldg(p.conf_pair_ptr + conf + 1) - __ldg(p.conf_pair_ptr + conf);
        if (cn > NCAP || np <= 0) continue;
        const int ntiles = (np + TE - 1) / TE;
        for (int k = g; k < ntiles; k += NG, ++it) {
          const int ne = min(TE, np - k * TE);
          const int npad = (ne + 15) & ~15;
          const uint32_t par = it & 1;
          tc::mbar_wait_spin(b1ready, par);
          tc::tc_fence_after();
          const uint32_t idesc1 = tc::umma_idesc_f16(F, npad, 1, 0, 0);
          for (int ks = 0; ks < k1steps; ++ks)
            tc::umma_f16(d, tc::umma_smem_desc(aW1 + ks * 256, 128, B1_SBO), tc::umma_smem_desc(aB + ks * 256, 128, B1_SBO),
                         idesc1, ks > 0);
          tc::umma_commit(d1ready);
          tc::mbar_wait_spin(b2ready, par);
          tc::tc_fence_after();
          const uint32_t idesc2 = tc::umma_idesc_f16(F, npad, 0, 0, 1);   // W2 image and a' are f16
#pragma unroll
          for (int ks = 0; ks < K2 / 16; ++ks)
            tc::umma_f16(d, tc::umma_smem_desc(aW2 + ks * 256, 128, A2_SBO), tc::umma_smem_desc(aB + ks * 256, 128, A2_SBO),
                         idesc2, ks > 0);
          tc::umma_commit(d2ready);
        }
      }
    }
    __syncwarp();
  } else {
    // ======================= compute warps of pipeline g =======================
    const int g = warp >> 2;
    const int tt = tid - g * GT;          // = filter channel owned in the epilogues (TMEM lane)
    const int chan = tt;
    const int e = tt % TE;                // pair column of this thread in the metadata / rbf phase
    const int q = tt / TE;                // GT / TE threads share a column in the rbf phase
    uint64_t* xbar = &bars[1];
    uint64_t* b1ready = &bars[2 + g * 4 + 0];
    uint64_t* d1ready = &bars[2 + g * 4 + 1];
    uint64_t* b2ready = &bars[2 + g * 4 + 2];
    uint64_t* d2ready = &bars[2 + g * 4 + 3];
    uint8_t* sB = groups + g * GROUP_BYTES;
    float* acc = reinterpret_cast<float*>(sB + OFF_ACC);
    int* sOJ = reinterpret_cast<int*>(sB + OFF_OJ);
    int* sOI = reinterpret_cast<int*>(sB + OFF_OI);
    uint32_t* sFlags = reinterpret_cast<uint32_t*>(sB + OFF_FLAGS);
    __half* sCh = reinterpret_cast<__half*>(sB + OFF_CH);
    const uint32_t d = tmem_base + g * TE + ((uint32_t)((warp & 3) * 32) << 16);
    const float cutoff = p.cutoff;
    const bool transposed = p.transposed != 0;

    uint32_t it = 0, xphase = 0;
    for (int conf = blockIdx.x; conf < G; conf += gridDim.x) {
      const int cs = __ldg(p.seg_ptr + conf);
      const int cn = __ldg(p.seg_ptr + conf + 1) - cs;
      const int p0 = __ldg(p.conf_pair_ptr + conf);
      const int np = __ldg(p.conf_pair_ptr + conf + 1) - p0;
      if (cn > NCAP || np <= 0) continue;   // same rule in the MMA warps: large conformers belong to cfconv_tc.cu
      const int ntiles = (np + TE - 1) / TE;

      // every pipeline is done with the previous conformer (x rows, accumulators)
      const bool rec = p.dbg && blockIdx.x == 0 && tid == 0 && it < 24;
      long long t_conf = 0;
      if (rec) t_conf = clock64();
      tc::named_bar_sync(CTA_BAR, NCOMP);
      if (tid == 0) {
        tc::fence_proxy_async();
        const uint32_t bytes = (uint32_t)cn * F * 4;
        tc::mbar_arrive_expect_tx(xbar, bytes);
        tc::bulk_g2s(sXs, p.x + (int64_t)cs * F, bytes, xbar);
      }
      for (int item = tt; item < cn * (F / 4); item += GT)
        reinterpret_cast<float4*>(acc)[item] = make_float4(0.f, 0.f, 0.f, 0.f);
      bool x_ready = false;

      // global operands of this pipeline's first tile
      int pre_src = 0, pre_dst = 0, pre_nd = -1, pre_rev = 0;
      float pre_d = 0.0f;
      auto prefetch = [&](int k) {
        const int e0 = p0 + k * TE;
        const int ne = min(TE, np - k * TE);
        if (e < ne) {
          pre_d = __ldg(p.pdist + e0 + e);     // both threads of a column need the distance (rbf phase)
          if (tt < TE) {
            pre_src = __ldg(p.psrc + e0 + e);
            pre_dst = __ldg(p.pdst + e0 + e);
            pre_rev = __ldg(p.prev + e0 + e);
            pre_nd = (e + 1 < ne) ? __ldg(p.pdst + e0 + e + 1) : -1;
          }
        }
      };
      if (g < ntiles) prefetch(g);

      for (int k = g; k < ntiles; k += NG, ++it) {
        const int ne = min(TE, np - k * TE);
        const int npad = (ne + 15) & ~15;
        const uint32_t par = it & 1;
        if (rec) { p.dbg[it * 10 + 0] = t_conf; p.dbg[it * 10 + 1] = clock64(); }
        tc::named_bar_sync(1 + g, GT);   // previous tile of this pipeline consumed (images, metadata); acc zeroed
        if (rec) p.dbg[it * 10 + 2] = clock64();

        // ---- per-pair metadata (warps 0, 1 of the pipeline) ----
        if (tt < TE) {
          const bool live = e < ne;
          const bool rev = live && pre_rev != 0;
          sOJ[e] = live ? (pre_src - cs) * F : 0;
          sOI[e] = live ? (pre_dst - cs) * F : 0;
          const float cval = live ? 0.5f * (__cosf(pre_d * kPi / cutoff) + 1.0f) : 0.0f;
          sCh[e] = __float2half_rn(cval);
          const bool last = live && (pre_nd != pre_dst);
          const unsigned ends = __ballot_sync(0xffffffffu, last);
          const unsigned unp = __ballot_sync(0xffffffffu, live && !rev);
          if (lane == 0) {
            sFlags[e >> 5] = ends;
            sFlags[TE / 32 + (e >> 5)] = unp;
          }
        }
        // ---- Gaussian expansion -> B1 (K-major [pair, 64]); two threads per column ----
        if (e < npad) {
          const float dist = (e < ne) ? pre_d : 0.0f;
          uint8_t* rowp = sB + (e >> 3) * B1_SBO + (e & 7) * 16;
          for (int jc = q; jc < 2 * k1steps; jc += GT / TE) {
            float v[8];
            const float4* op = reinterpret_cast<const float4*>(s_offset + jc * 8);
            const float4* cp2 = reinterpret_cast<const float4*>(s_c2 + jc * 8);
            const float4 o0 = op[0], o1 = op[1], k0 = cp2[0], k1 = cp2[1];
            const float off[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
            const float ck[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float x = dist - off[j];
              v[j] = tc::fast_ex2(ck[j] * (x * x));
            }
            *reinterpret_cast<uint4*>(rowp + jc * 128) =
                make_uint4(tc::pack_bf16x2(v[0], v[1]), tc::pack_bf16x2(v[2], v[3]), tc::pack_bf16x2(v[4], v[5]),
                           tc::pack_bf16x2(v[6], v[7]));
          }
        }
        if (rec) p.dbg[it * 10 + 3] = clock64();
        tc::fence_proxy_async();
        tc::mbar_arrive(b1ready);
        if (k + NG < ntiles) prefetch(k + NG);

        // ---- epilogue 1: a' = C * ssp(D1) -> B2 (MN-major [144, pair]) ----
        tc::mbar_wait(d1ready, par);
        tc::tc_fence_after();
        if (rec) p.dbg[it * 10 + 4] = clock64();
        {
          uint8_t* colp = sB + chan * 16;
          for (int c0 = 0; c0 < npad; c0 += 16) {
            float v[16];
            tc::tmem_ld16(d + c0, v);
            const uint4 ca = *reinterpret_cast<const uint4*>(sCh + c0), cb4 = *reinterpret_cast<const uint4*>(sCh + c0 + 8);
            const uint32_t cw[8] = {ca.x, ca.y, ca.z, ca.w, cb4.x, cb4.y, cb4.z, cb4.w};
            tc::tmem_wait_ld();
            uint32_t o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              o[j] = tc::ssp_cutoff_f16x2(v[2 * j], v[2 * j + 1], *reinterpret_cast<const __half2*>(&cw[j]));
            *reinterpret_cast<uint4*>(colp + (c0 >> 3) * A2_SBO) = make_uint4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<uint4*>(colp + ((c0 >> 3) + 1) * A2_SBO) = make_uint4(o[4], o[5], o[6], o[7]);
          }
          // rows 128..143: row 128 = C_p (multiplies the b2 column of W2aug), rows 129..143 = 0
          for (int item = tt; item < (npad >> 3) * 16; item += GT) {
            const int ec = item >> 4, kr = item & 15;
            uint4 w = make_uint4(0, 0, 0, 0);
            if (kr == 0) w = *reinterpret_cast<const uint4*>(sCh + ec * 8);
            *reinterpret_cast<uint4*>(sB + ec * A2_SBO + (128 + kr) * 16) = w;
          }
        }
        if (rec) p.dbg[it * 10 + 5] = clock64();
        tc::tc_fence_before();
        tc::fence_proxy_async();
        tc::mbar_arrive(b2ready);

        // ---- epilogue 2: both directions of every pair, accumulated in shared memory ----
        tc::mbar_wait(d2ready, par);
        tc::tc_fence_after();
        if (rec) p.dbg[it * 10 + 6] = clock64();
        if (!x_ready) {
          tc::mbar_wait(xbar, xphase & 1);
          x_ready = true;
        }
        {
          const uint32_t xs_addr = tc::smem_u32(sXs + chan);   // x rows are read-only here: explicit shared addresses
          float* accb = acc + chan;
          uint32_t unp_any = 0u;
#pragma unroll
          for (int w = 0; w < TE / 32; ++w) unp_any |= sFlags[TE / 32 + w];
          const bool fast = unp_any == 0u;   // every live column has both directions
          int offI = sOI[0];
          float xi = tc::lds_f32(xs_addr + 4u * (uint32_t)(offI));
          float accA = 0.0f;
          long long t_ld = 0;
          for (int c0 = 0; c0 < npad; c0 += 16) {
            float v[16];
            const long long t_a = rec ? clock64() : 0;
            tc::tmem_ld16(d + c0, v);
            int oj[16];
            const int4* jp = reinterpret_cast<const int4*>(sOJ + c0);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const int4 o = jp[r];
              oj[4 * r + 0] = o.x; oj[4 * r + 1] = o.y; oj[4 * r + 2] = o.z; oj[4 * r + 3] = o.w;
            }
            const uint32_t ends = (sFlags[c0 >> 5] >> (c0 & 31)) & 0xffffu;
            if (fast) {
              tc::tmem_wait_ld();
              if (rec) t_ld += clock64() - t_a;
              // Groups of four columns.  The sources of one target row are strictly ascending, so when no row ends
              // INSIDE a group its four read-modify-write targets are distinct: their loads are issued together instead
              // of a dependent LDS -> FFMA -> STS chain per column (the compiler cannot prove the stores do not alias).
#pragma unroll
              for (int k = 0; k < 16; k += 4) {
                if (((ends >> k) & 0x7u) == 0u) {
                  const float x0 = tc::lds_f32(xs_addr + 4u * (uint32_t)oj[k]), x1 = tc::lds_f32(xs_addr + 4u * (uint32_t)oj[k + 1]), x2 = tc::lds_f32(xs_addr + 4u * (uint32_t)oj[k + 2]), x3 = tc::lds_f32(xs_addr + 4u * (uint32_t)oj[k + 3]);
                  float a0 = accb[oj[k]], a1 = accb[oj[k + 1]], a2 = accb[oj[k + 2]], a3 = accb[oj[k + 3]];
                  accA = fmaf(v[k], x0, accA);
                  accA = fmaf(v[k + 1], x1, accA);
                  accA = fmaf(v[k + 2], x2, accA);
                  accA = fmaf(v[k + 3], x3, accA);
                  a0 = fmaf(v[k], xi, a0);
                  a1 = fmaf(v[k + 1], xi, a1);
                  a2 = fmaf(v[k + 2], xi, a2);
                  a3 = fmaf(v[k + 3], xi, a3);
                  accb[oj[k]] = a0;
                  accb[oj[k + 1]] = a1;
                  accb[oj[k + 2]] = a2;
                  accb[oj[k + 3]] = a3;
                  if ((ends >> (k + 3)) & 1u) {
                    accb[offI] += accA;
                    accA = 0.0f;
                    offI = sOI[min(c0 + k + 4, TE - 1)];
                    xi = tc::lds_f32(xs_addr + 4u * (uint32_t)(offI));
                  }
                } else {
#pragma unroll
                  for (int j = k; j < k + 4; ++j) {
                    const float xj = tc::lds_f32(xs_addr + 4u * (uint32_t)oj[j]);
                    float aj = accb[oj[j]];
                    accA = fmaf(v[j], xj, accA);
                    aj = fmaf(v[j], xi, aj);
                    accb[oj[j]] = aj;
                    if ((ends >> j) & 1u) {
                      accb[offI] += accA;
                      accA = 0.0f;
                      offI = sOI[min(c0 + j + 1, TE - 1)];
                      xi = tc::lds_f32(xs_addr + 4u * (uint32_t)(offI));
                    }
                  }
                }
              }
            } else {
              // direction into i (the primary edge j -> i) and into j (its reverse); an unpaired edge keeps only its
              // own direction, which is the one into i - or into j in the transposed pass
              const uint32_t unp = (sFlags[TE / 32 + (c0 >> 5)] >> (c0 & 31)) & 0xffffu;
              float fa[16], fb[16];
#pragma unroll
              for (int r = 0; r < 16; ++r) {
                const float paired = ((unp >> r) & 1u) ? 0.0f : 1.0f;
                fa[r] = transposed ? paired : 1.0f;
                fb[r] = transposed ? 1.0f : paired;
              }
              tc::tmem_wait_ld();
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float xj = tc::lds_f32(xs_addr + 4u * (uint32_t)oj[j]);
                float aj = accb[oj[j]];
                accA = fmaf(v[j] * fa[j], xj, accA);
                aj = fmaf(v[j] * fb[j], xi, aj);
                accb[oj[j]] = aj;
                if ((ends >> j) & 1u) {
                  accb[offI] += accA;
                  accA = 0.0f;
                  offI = sOI[min(c0 + j + 1, TE - 1)];
                  xi = tc::lds_f32(xs_addr + 4u * (uint32_t)(offI));
                }
              }
            }
          }
          if (rec) p.dbg[it * 10 + 9] = t_ld;
        }
        if (rec) p.dbg[it * 10 + 7] = clock64();
        tc::tc_fence_before();
      }
      ++xphase;

      // ---- all pipelines finished this conformer: sum their accumulators in pipeline order, write the rows ----
      tc::named_bar_sync(CTA_BAR, NCOMP);
      {
        float4* outp = reinterpret_cast<float4*>(p.out + (int64_t)cs * F);
        for (int item = tid; item < cn * (F / 4); item += NCOMP) {
          float4 s = reinterpret_cast<const float4*>(groups + OFF_ACC)[item];
#pragma unroll
          for (int gg = 1; gg < NG; ++gg) {
            const float4 t = reinterpret_cast<const float4*>(groups + gg * GROUP_BYTES + OFF_ACC)[item];
            s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
          }
          outp[item] = s;
        }
      }
      if (rec && it > 0) p.dbg[(it - 1) * 10 + 8] = clock64();   // finalize of the conformer this tile belonged to
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace
}  // namespace cmp

using namespace cmp;

static long long* g_pair_dbg = nullptr;
extern "C" void cmp_debug_set_pair_timestamps(void* buf) { g_pair_dbg = reinterpret_cast<long long*>(buf); }

extern "C" int cmp_cfconv_pair_max_atoms(void) { return NCAP; }

extern "C" int cmp_cfconv_pair_fwd(const float* x, const int32_t* seg_ptr, const int32_t* conf_pair_ptr,
                                   const int32_t* pair_src, const int32_t* pair_dst, const float* pair_dist,
                                   const int32_t* pair_rev, int64_t G, const void* packed_weights, const float* offset,
                                   int num_gaussians, float coeff, float cutoff, int num_filters, int transposed,
                                   float* out, cmp_stream_t stream) {
  CMP_REQUIRE(num_filters == F && num_gaussians >= 1 && num_gaussians < K1, CMP_EUNSUPPORTED,
              "cmp_cfconv_pair_fwd: needs num_filters == 128 and num_gaussians < 64 (got %d, %d)", num_filters,
              num_gaussians);
  CMP_REQUIRE(G >= 0 && G < ((int64_t)1 << 31) && cutoff > 0.0f, CMP_EINVAL, "cmp_cfconv_pair_fwd: bad size");
  if (G == 0) return CMP_OK;
  CMP_REQUIRE(x && seg_ptr && conf_pair_ptr && pair_src && pair_dst && pair_dist && pair_rev && packed_weights && offset &&
                  out,
              CMP_EINVAL, "cmp_cfconv_pair_fwd: null pointer");
  CMP_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)packed_weights % 16 == 0) && ((uintptr_t)out % 16 == 0), CMP_EINVAL,
              "cmp_cfconv_pair_fwd: x / packed_weights / out must be 16-byte aligned");
  CMP_REQUIRE(cmp_device_is_sm100(), CMP_EUNSUPPORTED, "cmp_cfconv_pair_fwd: needs an sm_100 device (tcgen05)");
  cudaStream_t st = as_stream(stream);
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(cfconv_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES) !=
        cudaSuccess) {
      (void)cudaGetLastError();
      set_error("cmp_cfconv_pair_fwd: cannot opt in to %u bytes of shared memory", SMEM_BYTES);
      return CMP_ECUDA;
    }
    attr_set = true;
  }
  PairParams p;
  p.x = x;
  p.seg_ptr = seg_ptr;
  p.conf_pair_ptr = conf_pair_ptr;
  p.psrc = pair_src;
  p.pdst = pair_dst;
  p.pdist = pair_dist;
  p.prev = pair_rev;
  p.weights = reinterpret_cast<const uint8_t*>(packed_weights);
  p.offset = offset;
  p.out = out;
  p.coeff_log2e = coeff * 1.4426950408889634f;
  p.cutoff = cutoff;
  p.Ng = num_gaussians;
  p.G = (int)G;
  p.transposed = transposed;
  p.dbg = g_pair_dbg;
  const int grid = (int)std::min<int64_t>(G, sm_count());
  cfconv_pair_kernel<<<grid, CTA_THREADS, SMEM_BYTES, st>>>(p);
  CMP_LAUNCH_CHECK("cmp_cfconv_pair_fwd");
  return CMP_OK;
}
