// Fused CFConv for sm_100a: one pass per interaction block, the E x F filter never leaves the SM.
//
// Replaces PyG CFConv.forward / message for num_filters = 128:
//   agg[i,:] = sum_{j->i} x'[j,:] * ( W2 * ssp(W1 * rbf(d_ij) + b1) + b2 ) * C(d_ij)
//
// Work unit = an edge tile: up to 128 edges = a run of whole target rows of one conformer (built by
// cmp_build_tiles).  Orientation: filter channels live on the 128 TMEM lanes, edges on the columns,
//   D1[128 hid, e] = W1aug[128, 64]  * rbf_aug[64, e]     (A, B K-major;   bias b1 rides in column Ng)
//   D2[128 out, e] = W2aug[128, 144] * a'[144, e]         (A K-major, B MN-major; both f16: the softplus epilogue
//                                                          runs in packed f16x2, tc_common.cuh)
// with a'[k, e] = C_e * ssp(D1[k, e]) and the extra row a'[128, e] = C_e carrying b2 * C_e, so the
// epilogue thread that owns channel f walks the tile's edges in CSR order doing one FMA per edge,
//   acc += D2[f, e] * x'[src_e, f],
// and stores agg[dst, f] when a target row ends: deterministic, no shuffles, no atomics, same
// summation order as the oracle's index_add_.
//
// CTA = NG independent pipelines ("groups": 4 compute warps + 1 MMA-issuing warp each) that alternate
// between SIMT phases and tensor phases so one group's epilogue overlaps the other's MMAs.  x' rows of
// the tile's conformer are staged in shared memory by 1-D TMA bulk copies (re-used by consecutive
// tiles of the same conformer); weights arrive as ready-made UMMA shared-memory images.
#include "common.cuh"
#include "tc_common.cuh"

namespace cmp {
namespace {

constexpr int F = 128;           // filter channels = UMMA M
constexpr int TILE_E = 128;      // edges per tile   = max UMMA N
constexpr int K1 = 64;           // Gaussians padded (+ bias column)
constexpr int K2 = 144;          // hidden channels + (cutoff, bias) row, padded to 16
constexpr int XP_CAP = 39;       // atoms of a conformer staged in shared memory
constexpr int NG = 3;            // pipelines ("groups") per CTA
constexpr int GT = 256;          // compute threads per group (8 warps: 2 per TMEM lane quarter)
constexpr int CTA_THREADS = NG * GT + NG * 32;

constexpr uint32_t W1_BYTES = F * K1 * 2;         // 16384
constexpr uint32_t W2_BYTES = F * K2 * 2;         // 36864
constexpr uint32_t B1_SBO = (K1 / 8) * 128;       // 1024: 8-row group stride of a K-major [rows, 64] image
constexpr uint32_t A2_SBO = (K2 / 8) * 128;       // 2304: 8-row group stride of a K-major [rows, 144] image
constexpr uint32_t B2_BYTES = K2 * TILE_E * 2;    // 36864 (the rbf image, 16384 B, aliases its head)
constexpr uint32_t XP_BYTES = XP_CAP * F * 4;     // 40960
constexpr uint32_t OFF_X = B2_BYTES;
constexpr uint32_t OFF_META = OFF_X + XP_BYTES;          // int[128] x' row offset (floats) | int[128] target row
constexpr uint32_t OFF_ROW = OFF_META + TILE_E * 8;      // int[136]  row offsets of the tile (relative)
constexpr uint32_t OFF_END = OFF_ROW + 544;              // uint32[8] row-end bit mask of each 16-edge chunk
constexpr uint32_t OFF_CH = OFF_END + 32;                // half[128] cosine cutoff again, as f16 (packed epilogue 1)
constexpr uint32_t GROUP_BYTES = OFF_CH + TILE_E * 2;
static_assert(W1_BYTES + W2_BYTES + NG * GROUP_BYTES <= 232448 - 2048, "shared memory budget (dynamic + ~2 KB static)");
constexpr uint32_t SMEM_BYTES = W1_BYTES + W2_BYTES + NG * GROUP_BYTES;

struct FwdParams {
  const float* xprime;
  const float* dist;
  const int32_t* rowptr;
  const int32_t* col;
  const int4* tiles;       // 2 x int4 per tile
  const int32_t* num_tiles;
  const uint8_t* weights;  // W1 image followed by W2 image
  const float* offset;     // Gaussian centres [Ng]
  float* agg;
  float coeff_log2e;       // coeff * log2(e)
  float cutoff;
  int Ng;
  long long* dbg;          // optional phase timestamps (CTA 0, pipeline 0): 8 x clock64 per tile
};

struct TileInfo {
  int row_begin, row_end, cs, cn, e0, ne, partial;   // partial: a chunk of ONE oversized row (sum is added, not stored)
};

__device__ __forceinline__ TileInfo load_tile(const int4* __restrict__ tiles, int64_t ti) {
  const int4 a = __ldg(tiles + 2 * ti), b = __ldg(tiles + 2 * ti + 1);
  TileInfo t;
  t.row_begin = a.x; t.row_end = a.y; t.cs = a.z; t.cn = a.w; t.e0 = b.x; t.ne = b.y; t.partial = b.z;
  return t;
}

__global__ void __launch_bounds__(CTA_THREADS, 1) cfconv_fused_fwd_kernel(const FwdParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bars[1 + NG * 5];  // wbar | per group: b1ready, d1ready, b2ready, d2ready, xbar
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_offset[K1];
  __shared__ __align__(16) float s_c2[K1];   // coeff*log2(e) for the Gaussians, 0 for the bias column and the padding

  uint8_t* sW1 = smem;
  uint8_t* sW2 = smem + W1_BYTES;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  // nothing to do (every conformer went to the pair kernel): leave before touching TMEM or the weight images
  if (*p.num_tiles == 0) return;

  if (tid == 0) {
    tc::mbar_init(&bars[0], 1);
    for (int g = 0; g < NG; ++g) {
      tc::mbar_init(&bars[1 + g * 5 + 0], GT);  // b1ready: every compute thread arrives
      tc::mbar_init(&bars[1 + g * 5 + 1], 1);   // d1ready: tcgen05.commit
      tc::mbar_init(&bars[1 + g * 5 + 2], GT);  // b2ready
      tc::mbar_init(&bars[1 + g * 5 + 3], 1);   // d2ready
      tc::mbar_init(&bars[1 + g * 5 + 4], 1);   // xbar: TMA bulk copy of x'
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

  const int64_t T = *p.num_tiles;
  const int64_t U = (int64_t)gridDim.x * NG;
  const int k1steps = (p.Ng + 1 + 15) >> 4;   // UMMA K-steps that hold Gaussians + the bias column

  if (warp >= NG * (GT / 32)) {
    // ======================= MMA-issuing warp of group g =======================
    const int g = warp - NG * (GT / 32);
    if (lane == 0) {
      uint64_t* wbar = &bars[0];
      uint64_t* b1ready = &bars[1 + g * 5 + 0];
      uint64_t* d1ready = &bars[1 + g * 5 + 1];
      uint64_t* b2ready = &bars[1 + g * 5 + 2];
      uint64_t* d2ready = &bars[1 + g * 5 + 3];
      if (g == 0) {
        tc::mbar_arrive_expect_tx(wbar, W1_BYTES + W2_BYTES);
        tc::bulk_g2s(sW1, p.weights, W1_BYTES, wbar);
        tc::bulk_g2s(sW2, p.weights + W1_BYTES, W2_BYTES, wbar);
      }
      uint8_t* sB = smem + W1_BYTES + W2_BYTES + g * GROUP_BYTES;
      const uint32_t d1 = tmem_base + g * 128, d2 = d1;
      const uint32_t aW1 = tc::smem_u32(sW1), aW2 = tc::smem_u32(sW2), aB = tc::smem_u32(sB);
      const int64_t u = (int64_t)blockIdx.x * NG + g;
      const int64_t t0 = u * T / U, t1 = (u + 1) * T / U;
      tc::mbar_wait(wbar, 0);
      uint32_t it = 0;
      int ne = (t0 < t1) ? load_tile(p.tiles, t0).ne : 0;
      for (int64_t ti = t0; ti < t1; ++ti, ++it) {
        const int npad = (ne + 15) & ~15;
        if (ti + 1 < t1) ne = load_tile(p.tiles, ti + 1).ne;
        const uint32_t par = it & 1;
        tc::mbar_wait_spin(b1ready, par);
        tc::tc_fence_after();
        const uint32_t idesc1 = tc::umma_idesc_f16(F, npad, 1, 0, 0);
        for (int ks = 0; ks < k1steps; ++ks)
          tc::umma_f16(d1, tc::umma_smem_desc(aW1 + ks * 256, 128, B1_SBO), tc::umma_smem_desc(aB + ks * 256, 128, B1_SBO),
                       idesc1, ks > 0);
        tc::umma_commit(d1ready);
        tc::mbar_wait_spin(b2ready, par);
        tc::tc_fence_after();
        const uint32_t idesc2 = tc::umma_idesc_f16(F, npad, 0, 0, 1);   // W2 image and a' are f16
#pragma unroll
        for (int ks = 0; ks < K2 / 16; ++ks)
          tc::umma_f16(d2, tc::umma_smem_desc(aW2 + ks * 256, 128, A2_SBO), tc::umma_smem_desc(aB + ks * 256, 128, A2_SBO),
                       idesc2, ks > 0);
        tc::umma_commit(d2ready);
      }
    }
    __syncwarp();
  } else {
    // ======================= compute warps of group g =======================
    const int g = warp / (GT / 32);
    const int tt = tid - g * GT;           // 0..255 within the group
    const int e = tt & 127;                // edge slot of this thread in the rbf phase
    const int h = tt >> 7;                 // which half of the work this thread takes
    const int wq = warp & 3;               // TMEM lane quarter this warp may touch
    uint64_t* b1ready = &bars[1 + g * 5 + 0];
    uint64_t* d1ready = &bars[1 + g * 5 + 1];
    uint64_t* b2ready = &bars[1 + g * 5 + 2];
    uint64_t* d2ready = &bars[1 + g * 5 + 3];
    uint64_t* xbar = &bars[1 + g * 5 + 4];
    uint8_t* sB = smem + W1_BYTES + W2_BYTES + g * GROUP_BYTES;
    float* sX = reinterpret_cast<float*>(sB + OFF_X);
    int* sSrc = reinterpret_cast<int*>(sB + OFF_META);
    int* sDst = sSrc + TILE_E;
    int* sRow = reinterpret_cast<int*>(sB + OFF_ROW);
    uint32_t* sEnd = reinterpret_cast<uint32_t*>(sB + OFF_END);
    __half* sCh = reinterpret_cast<__half*>(sB + OFF_CH);
    const uint32_t d1 = tmem_base + g * 128 + ((uint32_t)(wq * 32) << 16);
    const uint32_t d2 = d1;
    const int64_t u = (int64_t)blockIdx.x * NG + g;
    const int64_t t0 = u * T / U, t1 = (u + 1) * T / U;
    const float c2 = p.coeff_log2e, cutoff = p.cutoff;
    const int Ng = p.Ng;
    const int chan = wq * 32 + lane;       // channel owned in the epilogues (TMEM lane)

    int staged_conf = -1;
    uint32_t xloads = 0;
    uint32_t it = 0;

    // software prefetch of everything the next tile needs from global memory
    TileInfo cur;
    int pre_row = 0, pre_src = 0;
    float pre_d = 0.0f;
    if (t0 < t1) {
      cur = load_tile(p.tiles, t0);
      if (tt <= cur.row_end - cur.row_begin) pre_row = __ldg(p.rowptr + cur.row_begin + tt);
      if (e < cur.ne) {
        pre_d = __ldg(p.dist + cur.e0 + e);
        pre_src = __ldg(p.col + cur.e0 + e);
      }
    }

    for (int64_t ti = t0; ti < t1; ++ti, ++it) {
      const TileInfo tile = cur;
      const bool have_next = ti + 1 < t1;
      TileInfo nxt = tile;
      if (have_next) nxt = load_tile(p.tiles, ti + 1);
      const int ne = tile.ne;
      const int npad = (ne + 15) & ~15;
      const int nrows = tile.row_end - tile.row_begin;
      const uint32_t par = it & 1;
      const int cs = tile.cs, cn = tile.cn;
      const bool staged = cn <= XP_CAP;

      const bool rec = p.dbg && blockIdx.x == 0 && g == 0 && tt == 0 && it < 32;
      if (rec) p.dbg[it * 8 + 0] = clock64();
      tc::named_bar_sync(1 + g, GT);  // previous tile of this group fully consumed (sB, sX, sMeta, TMEM)

      bool x_wait = false;
      if (staged && cs != staged_conf) {
        if (tt == 0) {
          const uint32_t bytes = (uint32_t)cn * F * 4;
          tc::mbar_arrive_expect_tx(xbar, bytes);
          tc::bulk_g2s(sX, p.xprime + (int64_t)cs * F, bytes, xbar);
        }
        staged_conf = cs;
        x_wait = true;
      }
      if (tt <= nrows) sRow[tt] = tile.partial ? (tt == 0 ? 0 : ne) : pre_row - tile.e0;
      tc::named_bar_sync(1 + g, GT);

      if (rec) p.dbg[it * 8 + 1] = clock64();
      // ---- per-edge metadata + Gaussian expansion -> B1 (K-major [edge, 64]) ----
      if (h == 0) {   // warps 0..3 of the group: one edge slot per thread
        bool last = false;
        int srcoff = 0, dst = tile.row_begin;
        float cval = 0.0f;
        if (e < ne) {
          int r = 0;
          while (sRow[r + 1] <= e) ++r;
          last = (e + 1 == sRow[r + 1]);
          srcoff = (staged ? (pre_src - cs) : pre_src) * F;
          dst = tile.row_begin + r;
          cval = 0.5f * (__cosf(pre_d * kPi / cutoff) + 1.0f);
        }
        sSrc[e] = srcoff;
        sDst[e] = dst;
        sCh[e] = __float2half_rn(cval);
        const unsigned ends = __ballot_sync(0xffffffffu, last);
        if (lane == 0) {
          sEnd[e >> 4] = ends & 0xffffu;
          sEnd[(e >> 4) + 1] = ends >> 16;
        }
      }
      if (e < npad) {
        const float d = (e < ne) ? pre_d : 0.0f;
        uint8_t* rowp = sB + (e >> 3) * B1_SBO + (e & 7) * 16;
        const int jc0 = h * k1steps, jc1 = jc0 + k1steps;   // each half writes k1steps of the 2*k1steps chunks
        // exp2(c2_k (d - mu_k)^2) with c2_k = 0 beyond the Gaussians: the bias column (k = Ng) becomes exactly 1, the
        // padding columns meet zero weights, and rows of padded edges only have to be finite (their C is 0)
        for (int jc = jc0; jc < jc1; ++jc) {
          float v[8];
          const float4* op = reinterpret_cast<const float4*>(s_offset + jc * 8);
          const float4* cp2 = reinterpret_cast<const float4*>(s_c2 + jc * 8);
          const float4 o0 = op[0], o1 = op[1], k0 = cp2[0], k1 = cp2[1];
          const float off[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
          const float ck[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float x = d - off[j];
            v[j] = tc::fast_ex2(ck[j] * (x * x));
          }
          *reinterpret_cast<uint4*>(rowp + jc * 128) =
              make_uint4(tc::pack_bf16x2(v[0], v[1]), tc::pack_bf16x2(v[2], v[3]), tc::pack_bf16x2(v[4], v[5]),
                         tc::pack_bf16x2(v[6], v[7]));
        }
      }
      if (rec) p.dbg[it * 8 + 2] = clock64();
      tc::fence_proxy_async();
      tc::mbar_arrive(b1ready);

      // prefetch the next tile's global data while the tensor core and the epilogues run
      if (have_next) {
        if (tt <= nxt.row_end - nxt.row_begin) pre_row = __ldg(p.rowptr + nxt.row_begin + tt);
        if (e < nxt.ne) {
          pre_d = __ldg(p.dist + nxt.e0 + e);
          pre_src = __ldg(p.col + nxt.e0 + e);
        }
      }
      const int csplit = (((npad >> 4) + 1) >> 1) << 4;    // ep1 column split (multiple of 16)

      // ---- epilogue 1: a' = C * ssp(D1) -> B2 (MN-major [144, edge]) ----
      tc::mbar_wait(d1ready, par);
      tc::tc_fence_after();
      if (rec) p.dbg[it * 8 + 3] = clock64();
      {
        uint8_t* colp = sB + chan * 16;  // k = chan: (k/8)*128 + (k%8)*16 = k*16
        const int cb = h ? csplit : 0, ce = h ? npad : csplit;
        for (int c0 = cb; c0 < ce; c0 += 16) {
          float v[16];
          tc::tmem_ld16(d1 + c0, v);
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
        // rows 128..143: row 128 = C_e (multiplies the b2 column of W2aug), rows 129..143 = 0
        for (int item = tt; item < (npad >> 3) * 16; item += GT) {
          const int ec = item >> 4, kr = item & 15;
          uint4 q = make_uint4(0, 0, 0, 0);
          if (kr == 0) q = *reinterpret_cast<const uint4*>(sCh + ec * 8);
          *reinterpret_cast<uint4*>(sB + ec * A2_SBO + (128 + kr) * 16) = q;
        }
      }
      if (rec) p.dbg[it * 8 + 4] = clock64();
      tc::tc_fence_before();
      tc::fence_proxy_async();
      tc::mbar_arrive(b2ready);

      // ---- epilogue 2: gather x'_src, multiply, reduce per target row (CSR order) ----
      tc::mbar_wait(d2ready, par);
      tc::tc_fence_after();
      if (rec) p.dbg[it * 8 + 5] = clock64();
      if (x_wait) {
        tc::mbar_wait(xbar, xloads & 1);
        ++xloads;
      }
      {
        // each warp half takes a 16-aligned half of the columns; a target row cut by the split is completed by two
        // atomic adds onto the pre-zeroed output (two addends: commutative, hence still bitwise deterministic)
        const int cb = h ? csplit : 0, ce = h ? npad : csplit;
        float acc = 0.0f;
        float* aggc = p.agg + chan;
        bool cont = h && cb < ne && !((sEnd[(cb - 1) >> 4] >> ((cb - 1) & 15)) & 1u);
        const bool partial = tile.partial != 0;
        const float* xs_smem = sX + chan;
        const float* xs_gmem = p.xprime + chan;
        for (int c0 = cb; c0 < ce; c0 += 16) {
          float v[16];
          tc::tmem_ld16(d2 + c0, v);
          const uint32_t ends = sEnd[c0 >> 4];
          float xs[16];
          const int4* sp = reinterpret_cast<const int4*>(sSrc + c0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int4 o = sp[q];
            if (staged) {
              xs[4 * q + 0] = xs_smem[o.x]; xs[4 * q + 1] = xs_smem[o.y];
              xs[4 * q + 2] = xs_smem[o.z]; xs[4 * q + 3] = xs_smem[o.w];
            } else {
              xs[4 * q + 0] = __ldg(xs_gmem + o.x); xs[4 * q + 1] = __ldg(xs_gmem + o.y);
              xs[4 * q + 2] = __ldg(xs_gmem + o.z); xs[4 * q + 3] = __ldg(xs_gmem + o.w);
            }
          }
          tc::tmem_wait_ld();
          if (ends == 0u) {
#pragma unroll
            for (int j = 0; j < 16; ++j) acc = fmaf(v[j], xs[j], acc);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              acc = fmaf(v[j], xs[j], acc);
              if ((ends >> j) & 1u) {
                float* dstp = aggc + (int64_t)sDst[c0 + j] * F;
                if (cont || partial) {
                  atomicAdd(dstp, acc);
                  cont = false;
                } else {
                  *dstp = acc;
                }
                acc = 0.0f;
              }
            }
          }
        }
        // half 0 ends inside a target row: hand its partial sum over
        if (!h && ce > 0 && ce <= ne && ce < npad + 1 && !((sEnd[(ce - 1) >> 4] >> ((ce - 1) & 15)) & 1u) && (ce - 1) < ne)
          atomicAdd(aggc + (int64_t)sDst[ce - 1] * F, acc);
      }
      if (rec) p.dbg[it * 8 + 6] = clock64();
      tc::tc_fence_before();
      cur = nxt;
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

// ---- weight images ----------------------------------------------------------------------------------
__device__ __forceinline__ void pack_weights_body(const float* __restrict__ W1, const float* __restrict__ b1,
                                                  const float* __restrict__ W2, const float* __restrict__ b2, int Ng,
                                                  uint8_t* __restrict__ out, int idx) {
  if (idx < F * K1) {
    const int m = idx / K1, k = idx % K1;
    float v = (k < Ng) ? W1[m * Ng + k] : (k == Ng ? b1[m] : 0.0f);
    uint32_t off = (m & 7) * 16 + (k & 7) * 2 + (m >> 3) * B1_SBO + (k >> 3) * 128;
    *reinterpret_cast<__nv_bfloat16*>(out + off) = __float2bfloat16_rn(v);
  } else if (idx < F * K1 + F * K2) {
    const int j = idx - F * K1;
    const int m = j / K2, k = j % K2;
    float v = 0.0f;
    if (k < F) {
      v = W2[m * F + k];
    } else if (k == F) {
      v = b2[m];
    }
    uint32_t off = (m & 7) * 16 + (k & 7) * 2 + (m >> 3) * A2_SBO + (k >> 3) * 128;
    *reinterpret_cast<__half*>(out + W1_BYTES + off) = __float2half_rn(v);   // f16: the operand of the f16 a' image
  }
}

__global__ void pack_weights_kernel(const float* __restrict__ W1, const float* __restrict__ b1,
                                    const float* __restrict__ W2, const float* __restrict__ b2, int Ng,
                                    uint8_t* __restrict__ out) {
  pack_weights_body(W1, b1, W2, b2, Ng, out, blockIdx.x * blockDim.x + threadIdx.x);
}

// grouped: the filter MLPs of all interaction blocks in one launch (job = blockIdx.y)
constexpr int MAX_PACK_JOBS = 32;
struct PackFilterJob {
  const float* W1;
  const float* b1;
  const float* W2;
  const float* b2;
  uint8_t* packed_fwd;
  uint8_t* packed_bwd;
};
struct PackFilterGroup {
  PackFilterJob j[MAX_PACK_JOBS];
};
__global__ void pack_weights_grouped_kernel(const __grid_constant__ PackFilterGroup g, int Ng) {
  const PackFilterJob& j = g.j[blockIdx.y];
  pack_weights_body(j.W1, j.b1, j.W2, j.b2, Ng, j.packed_fwd, blockIdx.x * blockDim.x + threadIdx.x);
}

}  // namespace
}  // namespace cmp

using namespace cmp;

// debug hook: device buffer of 32 x 8 int64 phase timestamps filled by CTA 0 / pipeline 0 of the next launches
static long long* g_fwd_dbg = nullptr;
extern "C" void cmp_debug_set_fwd_timestamps(void* buf) { g_fwd_dbg = reinterpret_cast<long long*>(buf); }

extern "C" int cmp_cfconv_tc_supported(int num_filters, int num_gaussians) {
  return num_filters == F && num_gaussians >= 1 && num_gaussians < K1;
}

extern "C" size_t cmp_cfconv_tc_weights_bytes(void) { return W1_BYTES + W2_BYTES; }

extern "C" int cmp_cfconv_tc_tile_edges(void) { return TILE_E; }

extern "C" int cmp_cfconv_tc_pack_weights(const float* W1, const float* b1, const float* W2, const float* b2,
                                          int num_filters, int num_gaussians, void* packed, cmp_stream_t stream) {
  CMP_REQUIRE(cmp_cfconv_tc_supported(num_filters, num_gaussians), CMP_EUNSUPPORTED,
              "cmp_cfconv_tc_pack_weights: needs num_filters == 128 and num_gaussians < 64 (got %d, %d)", num_filters,
              num_gaussians);
  CMP_REQUIRE(W1 && b1 && W2 && b2 && packed, CMP_EINVAL, "cmp_cfconv_tc_pack_weights: null pointer");
  const int total = F * K1 + F * K2;
  pack_weights_kernel<<<(total + 255) / 256, 256, 0, as_stream(stream)>>>(W1, b1, W2, b2, num_gaussians,
                                                                         reinterpret_cast<uint8_t*>(packed));
  CMP_LAUNCH_CHECK("cmp_cfconv_tc_pack_weights");
  return CMP_OK;
}

extern "C" int cmp_cfconv_fused_fwd(const float* xprime, const float* dist, const int32_t* rowptr, const int32_t* col,
                                    const void* tiles, const int32_t* num_tiles, const void* packed_weights,
                                    const float* offset, int num_gaussians, float coeff, float cutoff, int64_t N,
                                    int num_filters, float* agg, cmp_stream_t stream) {
  CMP_REQUIRE(cmp_cfconv_tc_supported(num_filters, num_gaussians), CMP_EUNSUPPORTED,
              "cmp_cfconv_fused_fwd: needs num_filters == 128 and num_gaussians < 64 (got %d, %d)", num_filters,
              num_gaussians);
  CMP_REQUIRE(N >= 0 && cutoff > 0.0f, CMP_EINVAL, "cmp_cfconv_fused_fwd: bad size");
  if (N == 0) return CMP_OK;
  CMP_REQUIRE(xprime && dist && rowptr && col && tiles && num_tiles && packed_weights && offset && agg, CMP_EINVAL,
              "cmp_cfconv_fused_fwd: null pointer");
  CMP_REQUIRE(((uintptr_t)xprime % 16 == 0) && ((uintptr_t)packed_weights % 16 == 0), CMP_EINVAL,
              "cmp_cfconv_fused_fwd: xprime / packed_weights must be 16-byte aligned (TMA bulk copy)");
  CMP_REQUIRE(cmp_device_is_sm100(), CMP_EUNSUPPORTED, "cmp_cfconv_fused_fwd: needs an sm_100 device (tcgen05)");
  cudaStream_t st = as_stream(stream);
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(cfconv_fused_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES) !=
        cudaSuccess) {
      (void)cudaGetLastError();
      set_error("cmp_cfconv_fused_fwd: cannot opt in to %u bytes of shared memory", SMEM_BYTES);
      return CMP_ECUDA;
    }
    attr_set = true;
  }
  // rows without edges are never visited by a tile: their aggregate is zero
  CMP_REQUIRE(cudaMemsetAsync(agg, 0, (size_t)N * F * sizeof(float), st) == cudaSuccess, CMP_ECUDA,
              "cmp_cfconv_fused_fwd: memset failed");
  FwdParams p;
  p.xprime = xprime;
  p.dist = dist;
  p.rowptr = rowptr;
  p.col = col;
  p.tiles = reinterpret_cast<const int4*>(tiles);
  p.num_tiles = num_tiles;
  p.weights = reinterpret_cast<const uint8_t*>(packed_weights);
  p.offset = offset;
  p.agg = agg;
  p.coeff_log2e = coeff * 1.4426950408889634f;
  p.cutoff = cutoff;
  p.Ng = num_gaussians;
  p.dbg = g_fwd_dbg;
  cfconv_fused_fwd_kernel<<<sm_count(), CTA_THREADS, SMEM_BYTES, st>>>(p);
  CMP_LAUNCH_CHECK("cmp_cfconv_fused_fwd");
  return CMP_OK;
}

extern "C" int cmp_cfconv_tc_pack_weights_grouped(const void* jobs, int count, int num_filters, int num_gaussians,
                                                  cmp_stream_t stream) {
  CMP_REQUIRE(cmp_cfconv_tc_supported(num_filters, num_gaussians), CMP_EUNSUPPORTED,
              "cmp_cfconv_tc_pack_weights_grouped: needs num_filters == 128 and num_gaussians < 64");
  CMP_REQUIRE(count >= 0 && count <= MAX_PACK_JOBS, CMP_EINVAL, "cmp_cfconv_tc_pack_weights_grouped: count must be in [0, %d]",
              MAX_PACK_JOBS);
  if (count == 0) return CMP_OK;
  CMP_REQUIRE(jobs, CMP_EINVAL, "cmp_cfconv_tc_pack_weights_grouped: null pointer");
  const PackFilterJob* in = reinterpret_cast<const PackFilterJob*>(jobs);
  PackFilterGroup g;
  for (int i = 0; i < count; ++i) {
    CMP_REQUIRE(in[i].W1 && in[i].b1 && in[i].W2 && in[i].b2 && in[i].packed_fwd, CMP_EINVAL,
                "cmp_cfconv_tc_pack_weights_grouped: null pointer");
    g.j[i] = in[i];
  }
  const int total = F * K1 + F * K2;
  pack_weights_grouped_kernel<<<dim3((total + 255) / 256, count), 256, 0, as_stream(stream)>>>(g, num_gaussians);
  CMP_LAUNCH_CHECK("cmp_cfconv_tc_pack_weights_grouped");
  return CMP_OK;
}
