// Fused CFConv backward (filter-MLP weight gradients) for sm_100a.
//
// Forward (cfconv_tc.cu):  W_e = W2 a'_e + b2 C_e,  a'_e = C_e ssp(h_e),  h_e = W1 rbf_e + b1,
//                          agg_i = sum_{e: dst(e)=i} x'_{src(e)} * W_e.
// Given g = dL/dagg this kernel produces dW1, db1, dW2, db2 in ONE pass over the edges, recomputing
// rbf / h / a' on chip (nothing of size E x F is ever stored):
//   dF[f,e]  = g[dst_e,f] * x'[src_e,f]                       (SIMT, thread = channel f)
//   da'[k,e] = sum_f W2[f,k] dF[f,e]                          (UMMA:  W2^T image  x  dF as MN-major B)
//   dh[k,e]  = da'[k,e] * C_e * sigmoid(h[k,e])               (SIMT epilogue on TMEM)
//   dW2[f,k] += sum_e dF[f,e] a'[k,e]                          (UMMA, K = edges, accumulates in TMEM)
//   dW1[k,j] += sum_e dh[k,e] rbf[e,j]                         (UMMA, K = edges, accumulates in TMEM)
//   db2[f]   += sum_e dF[f,e] C_e ,  db1[k] += sum_e dh[k,e]   (registers)
// (d x' is NOT computed here: it is the forward kernel run over the transposed neighbour list.)
// Every operand image is written once and read under two descriptor views (K-major / MN-major with
// LBO and SBO exchanged), so no transpose is ever materialised.
//
// Work unit = 64 consecutive edges of one conformer (no row alignment needed: edges are independent
// here).  CTA = 2 pipelines of 8 compute warps + 1 MMA warp; each pipeline owns 256 TMEM columns:
// [0,64) h / da', [64,192) dW2 accumulator, [192,256) dW1 accumulator.  Per-pipeline partial sums go
// to global memory and are reduced in a fixed order by a second kernel (deterministic).
#include "common.cuh"
#include "tc_common.cuh"

namespace cmp {
namespace {

constexpr int F = 128;
constexpr int TE = 64;            // edges per tile
constexpr int K1 = 64;            // padded Gaussians (+ bias column)
constexpr int XP_CAP = 80;        // atoms staged (bf16 x')
constexpr int GROWS = 8;          // target rows of g staged per tile
constexpr int NG = 2;
constexpr int GT = 256;
constexpr int CTA_THREADS = NG * GT + NG * 32;

constexpr uint32_t W1_BYTES = F * K1 * 2;       // 16384, K-major [rows=k, K=j], SBO 1024, LBO 128
constexpr uint32_t W2T_BYTES = F * F * 2;       // 32768, K-major [rows=k, K=f], SBO 2048, LBO 128
constexpr uint32_t R_BYTES = TE * K1 * 2;       // 8192   rbf image  (e%8)*16 + (j%8)*2 + (e/8)*1024 + (j/8)*128
constexpr uint32_t CH_BYTES = F * TE * 2;       // 16384  channel-major images: c*16 + (e/8)*2048 + (e%8)*2
constexpr uint32_t OFF_A = R_BYTES;
constexpr uint32_t OFF_F = OFF_A + CH_BYTES;
constexpr uint32_t OFF_S = OFF_F + CH_BYTES;
constexpr uint32_t OFF_X = OFF_S + CH_BYTES;                // bf16 x' rows of the conformer
constexpr uint32_t XB_BYTES = XP_CAP * F * 2;               // 20480
constexpr uint32_t OFF_G = OFF_X + XB_BYTES;                // fp32 g rows [GROWS][128]
constexpr uint32_t OFF_META = OFF_G + GROWS * F * 4;        // int2[64] {src (local or global), dst}
constexpr uint32_t OFF_C = OFF_META + TE * 8;               // float[64]
constexpr uint32_t OFF_REV = OFF_C + TE * 4;                // float[64]  (pair mode: 1 when the reverse edge exists)
constexpr uint32_t GROUP_BYTES = OFF_REV + TE * 4;
// pair mode stages bf16 x' AND bf16 g of the conformer in the x' + g-row area: 2 x XP_PAIR rows
constexpr int XP_PAIR = 48;
static_assert(2u * XP_PAIR * F * 2 <= XB_BYTES + GROWS * F * 4, "pair-mode staging must fit the x' + g-row area");
constexpr uint32_t SMEM_BYTES = W1_BYTES + W2T_BYTES + NG * GROUP_BYTES;

constexpr int PART_FLOATS = F * F + F * K1 + 2 * F + 2 * F;   // dW2 | dW1 | db2[2 halves] | db1[2 halves]

struct BwdParams {
  const float* g;                 // [N, F] dL/dagg                      (edge mode)
  const __nv_bfloat16* gb;        // [N, F] bf16 copy of dL/dagg          (pair mode)
  const int32_t* rev;             // [P] reverse-edge flags               (pair mode)
  const __nv_bfloat16* xprime;    // [N, F] bf16 copy of x'
  const float* dist;
  const int32_t* col;
  const int32_t* erow;            // [E] target row of every edge
  const int4* tiles;              // 2 x int4: {first_row, end_row, conf_first_atom, conf_atoms}, {first_edge, num_edges,0,0}
  const int32_t* num_tiles;
  const uint8_t* weights;         // W1aug image | W2^T image
  const float* offset;
  float* partial;                 // [gridDim.x * NG][PART_FLOATS]
  float coeff_log2e;
  float cutoff;
  int Ng;
  long long* dbg;                 // optional phase timestamps (CTA 0, pipeline 0): 12 x clock64 per tile
};

struct TileInfo {
  int row_begin, row_end, cs, cn, e0, ne;
};

__device__ __forceinline__ TileInfo load_tile(const int4* __restrict__ tiles, int64_t ti) {
  const int4 a = __ldg(tiles + 2 * ti), b = __ldg(tiles + 2 * ti + 1);
  TileInfo t;
  t.row_begin = a.x; t.row_end = a.y; t.cs = a.z; t.cn = a.w; t.e0 = b.x; t.ne = b.y;
  return t;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void unpack_bf16x8(const uint4 q, float* v) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}

__device__ __forceinline__ uint4 pack_bf16x8(const float* v) {
  return make_uint4(tc::pack_bf16x2(v[0], v[1]), tc::pack_bf16x2(v[2], v[3]), tc::pack_bf16x2(v[4], v[5]),
                    tc::pack_bf16x2(v[6], v[7]));
}

// PAIR = false: one column per directed edge, dF = g[dst] x'[src].
// PAIR = true : one column per primary edge (graph.cu: one representative of {j->i, i->j}); both directions share the
//               filter, so dF = g[dst] x'[src] + rev * g[src] x'[dst] and the tile count halves.
template <bool PAIR>
__global__ void __launch_bounds__(CTA_THREADS, 1) cfconv_fused_bwd_kernel(const BwdParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  // wbar | per group: r_ready, d1_ready, f_ready, dda_ready, h_ready, w_done, xbar
  __shared__ uint64_t bars[1 + NG * 7];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_offset[K1];
  __shared__ __align__(16) float s_c2[K1];

  uint8_t* sW1 = smem;
  uint8_t* sW2T = smem + W1_BYTES;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    tc::mbar_init(&bars[0], 1);
    for (int g = 0; g < NG; ++g) {
      uint64_t* b = &bars[1 + g * 7];
      tc::mbar_init(b + 0, GT);  // r_ready   (rbf image + metadata written)
      tc::mbar_init(b + 1, 1);   // d1_ready  (h in TMEM)
      tc::mbar_init(b + 2, GT);  // f_ready   (a', S, dF images written; h consumed)
      tc::mbar_init(b + 3, 1);   // dda_ready (da' in TMEM)
      tc::mbar_init(b + 4, GT);  // h_ready   (dh image written)
      tc::mbar_init(b + 5, 1);   // w_done    (weight-gradient MMAs finished reading the images)
      tc::mbar_init(b + 6, 1);   // xbar
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
  const int k1steps = (p.Ng + 1 + 15) >> 4;

  if (warp >= NG * (GT / 32)) {
    // ======================= MMA-issuing warp of group g =======================
    const int g = warp - NG * (GT / 32);
    if (lane == 0) {
      uint64_t* wbar = &bars[0];
      uint64_t* b = &bars[1 + g * 7];
      if (g == 0) {
        tc::mbar_arrive_expect_tx(wbar, W1_BYTES + W2T_BYTES);
        tc::bulk_g2s(sW1, p.weights, W1_BYTES, wbar);
        tc::bulk_g2s(sW2T, p.weights + W1_BYTES, W2T_BYTES, wbar);
      }
      uint8_t* sG0 = smem + W1_BYTES + W2T_BYTES + g * GROUP_BYTES;
      const uint32_t aW1 = tc::smem_u32(sW1), aW2T = tc::smem_u32(sW2T);
      const uint32_t aR = tc::smem_u32(sG0), aA = aR + OFF_A, aF = aR + OFF_F, aS = aR + OFF_S;
      const uint32_t tD = tmem_base + g * 256, tW2 = tD + 64, tW1 = tD + 192;
      const int64_t u = (int64_t)blockIdx.x * NG + g;
      const int64_t t0 = u * T / U, t1 = (u + 1) * T / U;
      tc::mbar_wait(wbar, 0);
      uint32_t it = 0;
      int ne = (t0 < t1) ? load_tile(p.tiles, t0).ne : 0;
      for (int64_t ti = t0; ti < t1; ++ti, ++it) {
        const int npad = (ne + 15) & ~15;
        if (ti + 1 < t1) ne = load_tile(p.tiles, ti + 1).ne;
        const uint32_t par = it & 1;
        // h = W1aug * rbf^T
        tc::mbar_wait_spin(b + 0, par);
        tc::tc_fence_after();
        const uint32_t id1 = tc::umma_idesc_f16(F, npad, 1, 0, 0);
        for (int ks = 0; ks < k1steps; ++ks)
          tc::umma_f16(tD, tc::umma_smem_desc(aW1 + ks * 256, 128, 1024), tc::umma_smem_desc(aR + ks * 256, 128, 1024), id1,
                       ks > 0);
        tc::umma_commit(b + 1);
        // da' = W2^T * dF      (dF image read as MN-major [K=f, N=e]: LBO 128, SBO 2048)
        tc::mbar_wait_spin(b + 2, par);
        tc::tc_fence_after();
        const uint32_t id2 = tc::umma_idesc_f16(F, npad, 1, 0, 1);
#pragma unroll
        for (int ks = 0; ks < F / 16; ++ks)
          tc::umma_f16(tD, tc::umma_smem_desc(aW2T + ks * 256, 128, 2048), tc::umma_smem_desc(aF + ks * 256, 128, 2048), id2,
                       ks > 0);
        tc::umma_commit(b + 3);
        // weight gradients, K = edges of the tile (images read as K-major [rows=channel, K=e]: SBO 128, LBO 2048)
        tc::mbar_wait_spin(b + 4, par);
        tc::tc_fence_after();
        const uint32_t id3 = tc::umma_idesc_f16(F, F, 1, 0, 0);
        const uint32_t id4 = tc::umma_idesc_f16(F, K1, 1, 0, 1);
        for (int ks = 0; ks < (npad >> 4); ++ks) {
          const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
          tc::umma_f16(tW2, tc::umma_smem_desc(aF + ks * 4096, 2048, 128), tc::umma_smem_desc(aA + ks * 4096, 2048, 128), id3,
                       acc);
          // rbf image read as MN-major [K=e, N=j]: LBO (e-group stride) 1024, SBO (j-group stride) 128
          tc::umma_f16(tW1, tc::umma_smem_desc(aS + ks * 4096, 2048, 128), tc::umma_smem_desc(aR + ks * 2048, 1024, 128), id4,
                       acc);
        }
        tc::umma_commit(b + 5);
      }
    }
    __syncwarp();
  } else {
    // ======================= compute warps of group g =======================
    const int g = warp / (GT / 32);
    const int tt = tid - g * GT;
    const int wq = warp & 3;                  // TMEM lane quarter
    const int h = (warp >> 2) & 1;            // column (edge) half handled in the channel-major phases
    const int chan = wq * 32 + lane;
    const int e = tt & 63;                    // edge slot in the rbf phase
    const int q = tt >> 6;                    // 4 threads share an edge in the rbf phase
    uint64_t* b = &bars[1 + g * 7];
    uint64_t* xbar = b + 6;
    uint8_t* sR = smem + W1_BYTES + W2T_BYTES + g * GROUP_BYTES;
    uint8_t* sA = sR + OFF_A;
    uint8_t* sF = sR + OFF_F;
    uint8_t* sS = sR + OFF_S;
    const __nv_bfloat16* sXb = reinterpret_cast<const __nv_bfloat16*>(sR + OFF_X);
    float* sGr = reinterpret_cast<float*>(sR + OFF_G);
    int2* sMeta = reinterpret_cast<int2*>(sR + OFF_META);
    float* sC = reinterpret_cast<float*>(sR + OFF_C);
    float* sRev = reinterpret_cast<float*>(sR + OFF_REV);
    const __nv_bfloat16* sGb = sXb + XP_PAIR * F;            // pair mode: bf16 g rows of the conformer
    const uint32_t tD = tmem_base + g * 256 + ((uint32_t)(wq * 32) << 16);
    const uint32_t tW2 = tD + 64, tW1 = tD + 192;
    const int64_t u = (int64_t)blockIdx.x * NG + g;
    const int64_t t0 = u * T / U, t1 = (u + 1) * T / U;
    const float c2 = p.coeff_log2e, cutoff = p.cutoff;
    const int Ng = p.Ng;

    float db1 = 0.0f, db2 = 0.0f;
    int staged_conf = -1;
    uint32_t xloads = 0;
    uint32_t it = 0;

    TileInfo cur;
    int pre_src = 0, pre_dst = 0, pre_rev = 0;
    float pre_d = 0.0f;
    float4 pre_g = make_float4(0.f, 0.f, 0.f, 0.f);
    auto prefetch = [&](const TileInfo& t) {
      if (e < t.ne) {
        pre_d = __ldg(p.dist + t.e0 + e);
        pre_src = __ldg(p.col + t.e0 + e);
        pre_dst = __ldg(p.erow + t.e0 + e);
        if (PAIR) pre_rev = __ldg(p.rev + t.e0 + e);
      }
      if (!PAIR) {
        const int r = t.row_begin + (tt >> 5);
        if (t.row_end - t.row_begin <= GROWS && r < t.row_end)
          pre_g = __ldg(reinterpret_cast<const float4*>(p.g + (int64_t)r * F) + (tt & 31));
      }
    };
    if (t0 < t1) {
      cur = load_tile(p.tiles, t0);
      prefetch(cur);
    }

    for (int64_t ti = t0; ti < t1; ++ti, ++it) {
      const TileInfo tile = cur;
      const bool have_next = ti + 1 < t1;
      TileInfo nxt = tile;
      if (have_next) nxt = load_tile(p.tiles, ti + 1);
      const int ne = tile.ne;
      const int npad = (ne + 15) & ~15;
      const uint32_t par = it & 1;
      const int cs = tile.cs, cn = tile.cn;
      const bool staged = cn <= (PAIR ? XP_PAIR : XP_CAP);
      const int r0 = tile.row_begin;
      const bool g_staged = !PAIR && (tile.row_end - tile.row_begin) <= GROWS;

      const bool rec = p.dbg && blockIdx.x == 0 && g == 0 && tt == 0 && it < 20;
      if (rec) p.dbg[it * 12 + 0] = clock64();
      // the previous tile's weight-gradient MMAs must be done reading the images before they are rewritten
      if (it > 0) tc::mbar_wait(b + 5, (it - 1) & 1);
      if (rec) p.dbg[it * 12 + 1] = clock64();

      bool x_wait = false;
      if (staged && cs != staged_conf) {
        if (tt == 0) {
          const uint32_t bytes = (uint32_t)cn * F * 2;
          tc::mbar_arrive_expect_tx(xbar, PAIR ? 2 * bytes : bytes);
          tc::bulk_g2s(sR + OFF_X, p.xprime + (int64_t)cs * F, bytes, xbar);
          if (PAIR) tc::bulk_g2s(sR + OFF_X + XP_PAIR * F * 2, p.gb + (int64_t)cs * F, bytes, xbar);
        }
        staged_conf = cs;
        x_wait = true;
      }

      // ---- metadata, g rows, Gaussian expansion -> rbf image ----
      if (g_staged && r0 + (tt >> 5) < tile.row_end)
        reinterpret_cast<float4*>(sGr + (tt >> 5) * F)[tt & 31] = pre_g;
      if (e < npad) {
        const bool live = e < ne;
        const float d = pre_d;
        if (q == 0) {
          // pre-multiplied element offsets of the x' row and the g row this edge reads
          if (live) {
            if (PAIR)
              sMeta[e] = make_int2((staged ? (pre_src - cs) : pre_src) * F, (staged ? (pre_dst - cs) : pre_dst) * F);
            else
              sMeta[e] = make_int2((staged ? (pre_src - cs) : pre_src) * F, (g_staged ? (pre_dst - r0) : pre_dst) * F);
            sC[e] = 0.5f * (__cosf(d * kPi / cutoff) + 1.0f);
          } else {
            sMeta[e] = make_int2((staged ? 0 : cs) * F, ((PAIR ? staged : g_staged) ? 0 : (PAIR ? cs : r0)) * F);
            sC[e] = 0.0f;
          }
          if (PAIR) sRev[e] = (live && pre_rev) ? 1.0f : 0.0f;
        }
        uint8_t* rowp = sR + (e >> 3) * 1024 + (e & 7) * 16;
        // exp2(c2_k (d - mu_k)^2), c2_k = 0 beyond the Gaussians (bias column = 1; padding columns meet zero weights
        // in h and are discarded in dW1).  Rows of padded edges must be ZERO here: they enter dW1 through K = edges.
        for (int jc = q; jc < 2 * k1steps; jc += 4) {
          float v[8];
          const float4* op = reinterpret_cast<const float4*>(s_offset + jc * 8);
          const float4* cp2 = reinterpret_cast<const float4*>(s_c2 + jc * 8);
          const float4 o0 = op[0], o1 = op[1], k0 = cp2[0], k1 = cp2[1];
          const float off[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
          const float ck[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float x = d - off[j];
            v[j] = live ? tc::fast_ex2(ck[j] * (x * x)) : 0.0f;
          }
          *reinterpret_cast<uint4*>(rowp + jc * 128) = pack_bf16x8(v);
        }
      }
      if (rec) p.dbg[it * 12 + 2] = clock64();
      tc::fence_proxy_async();
      tc::mbar_arrive(b + 0);
      if (have_next) prefetch(nxt);
      tc::named_bar_sync(1 + g, GT);   // sMeta / sC / sGr visible to the whole group
      if (x_wait) {
        tc::mbar_wait(xbar, xloads & 1);
        ++xloads;
      }

      if (rec) p.dbg[it * 12 + 3] = clock64();
      // ---- dF[f, e] = g[dst_e, f] * x'[src_e, f]  ->  dF image (runs while the tensor core computes h) ----
      {
        // all shared-memory operands of 8 edges are fetched before any arithmetic (no dependent-load chains), and the
        // staged / fallback decision is hoisted out of the loop
        const int cb = h * 32, ce = min(npad, h * 32 + 32);
        const __nv_bfloat16* xs_s = sXb + chan;
        const float* gs_s = sGr + chan;
        const __nv_bfloat16* xs_g = p.xprime + chan;
        const float* gs_g = p.g + chan;
        for (int c0 = cb; c0 < ce; c0 += 8) {
          int2 m[8];
          const int4* mp = reinterpret_cast<const int4*>(sMeta + c0);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const int4 mm = mp[k4];
            m[2 * k4] = make_int2(mm.x, mm.y);
            m[2 * k4 + 1] = make_int2(mm.z, mm.w);
          }
          const float4 c0v = *reinterpret_cast<const float4*>(sC + c0), c1v = *reinterpret_cast<const float4*>(sC + c0 + 4);
          const float cc[8] = {c0v.x, c0v.y, c0v.z, c0v.w, c1v.x, c1v.y, c1v.z, c1v.w};
          float v[8];
          if (PAIR) {
            // both directions of the pair: g[dst] x'[src] + rev * g[src] x'[dst]  (all four rows from the bf16 copies)
            const float4 r0v = *reinterpret_cast<const float4*>(sRev + c0), r1v = *reinterpret_cast<const float4*>(sRev + c0 + 4);
            const float rv[8] = {r0v.x, r0v.y, r0v.z, r0v.w, r1v.x, r1v.y, r1v.z, r1v.w};
            const __nv_bfloat16* xs = staged ? xs_s : xs_g;
            const __nv_bfloat16* gs = staged ? (sGb + chan) : (p.gb + chan);
            float xs_[8], gd_[8], xd_[8], gs_[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              xs_[j] = __bfloat162float(xs[m[j].x]);
              gd_[j] = __bfloat162float(gs[m[j].y]);
              xd_[j] = __bfloat162float(xs[m[j].y]);
              gs_[j] = __bfloat162float(gs[m[j].x]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              v[j] = (c0 + j < ne) ? fmaf(rv[j] * gs_[j], xd_[j], gd_[j] * xs_[j]) : 0.0f;
              db2 = fmaf(v[j], cc[j], db2);
            }
          } else {
            float xv[8], gv[8];
            if (staged) {
#pragma unroll
              for (int j = 0; j < 8; ++j) xv[j] = __bfloat162float(xs_s[m[j].x]);
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) xv[j] = __bfloat162float(xs_g[m[j].x]);
            }
            if (g_staged) {
#pragma unroll
              for (int j = 0; j < 8; ++j) gv[j] = gs_s[m[j].y];
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) gv[j] = __ldg(gs_g + m[j].y);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              // padded edges carry C = 0 and valid (row 0) offsets; their dF must still be exactly 0 (K = edges)
              v[j] = (c0 + j < ne) ? gv[j] * xv[j] : 0.0f;
              db2 = fmaf(v[j], cc[j], db2);
            }
          }
          *reinterpret_cast<uint4*>(sF + chan * 16 + (c0 >> 3) * 2048) = pack_bf16x8(v);
        }
      }

      if (rec) p.dbg[it * 12 + 4] = clock64();
      // ---- epilogue 1: a' = C ssp(h), S = C sigmoid(h) -> images ----
      tc::mbar_wait(b + 1, par);
      tc::tc_fence_after();
      if (rec) p.dbg[it * 12 + 5] = clock64();
      {
        const int cb = h * 32, ce = min(npad, h * 32 + 32);
        for (int c0 = cb; c0 < ce; c0 += 16) {
          float v[16];
          tc::tmem_ld16(tD + c0, v);
          float c[16];
          const float4* cp = reinterpret_cast<const float4*>(sC + c0);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const float4 cc = cp[k4];
            c[k4 * 4 + 0] = cc.x; c[k4 * 4 + 1] = cc.y; c[k4 * 4 + 2] = cc.z; c[k4 * 4 + 3] = cc.w;
          }
          tc::tmem_wait_ld();
          float a[16], s[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float x = v[j];
            const float t = tc::fast_ex2(-1.4426950408889634f * fabsf(x));
            const float inv = __fdividef(1.0f, 1.0f + t);
            a[j] = c[j] * fmaf(tc::fast_lg2(1.0f + t) - 1.0f, kLn2, fmaxf(x, 0.0f));
            s[j] = c[j] * (x >= 0.0f ? inv : t * inv);
          }
          *reinterpret_cast<uint4*>(sA + chan * 16 + (c0 >> 3) * 2048) = pack_bf16x8(a);
          *reinterpret_cast<uint4*>(sA + chan * 16 + ((c0 >> 3) + 1) * 2048) = pack_bf16x8(a + 8);
          *reinterpret_cast<uint4*>(sS + chan * 16 + (c0 >> 3) * 2048) = pack_bf16x8(s);
          *reinterpret_cast<uint4*>(sS + chan * 16 + ((c0 >> 3) + 1) * 2048) = pack_bf16x8(s + 8);
        }
      }
      if (rec) p.dbg[it * 12 + 6] = clock64();
      tc::tc_fence_before();
      tc::fence_proxy_async();
      tc::mbar_arrive(b + 2);

      // ---- epilogue 3: dh = da' * S -> dh image (in place over S) ----
      tc::mbar_wait(b + 3, par);
      tc::tc_fence_after();
      if (rec) p.dbg[it * 12 + 7] = clock64();
      {
        const int cb = h * 32, ce = min(npad, h * 32 + 32);
        for (int c0 = cb; c0 < ce; c0 += 16) {
          float v[16];
          tc::tmem_ld16(tD + c0, v);
          float s[16];
          uint4* sp0 = reinterpret_cast<uint4*>(sS + chan * 16 + (c0 >> 3) * 2048);
          uint4* sp1 = reinterpret_cast<uint4*>(sS + chan * 16 + ((c0 >> 3) + 1) * 2048);
          unpack_bf16x8(*sp0, s);
          unpack_bf16x8(*sp1, s + 8);
          tc::tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            v[j] *= s[j];
            db1 += v[j];
          }
          *sp0 = pack_bf16x8(v);
          *sp1 = pack_bf16x8(v + 8);
        }
      }
      if (rec) p.dbg[it * 12 + 8] = clock64();
      tc::tc_fence_before();
      tc::fence_proxy_async();
      tc::mbar_arrive(b + 4);
      cur = nxt;
    }

    // ---- drain: accumulators -> this pipeline's partial block ----
    float* part = p.partial + u * (int64_t)PART_FLOATS;
    const bool any = t0 < t1;
    if (any) {
      tc::mbar_wait(b + 5, (it - 1) & 1);
      tc::tc_fence_after();
    }
    // dW2[f = chan][k]: columns [h*64, h*64+64)
    for (int c0 = h * 64; c0 < h * 64 + 64; c0 += 16) {
      float v[16];
      if (any) {
        tc::tmem_ld16(tW2 + c0, v);
        tc::tmem_wait_ld();
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.0f;
      }
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(part + chan * F + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    // dW1[k = chan][j]: columns [h*32, h*32+32)
    for (int c0 = h * 32; c0 < h * 32 + 32; c0 += 16) {
      float v[16];
      if (any) {
        tc::tmem_ld16(tW1 + c0, v);
        tc::tmem_wait_ld();
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.0f;
      }
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(part + F * F + chan * K1 + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    part[F * F + F * K1 + h * F + chan] = db2;
    part[F * F + F * K1 + 2 * F + h * F + chan] = db1;
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

// out = sum over pipelines of the partial blocks, scattered into the parameter layouts.  block = 32 outputs x 8
// segments of the pipeline list: each thread sums its contiguous segment in order, the segment sums are combined in
// segment order (a fixed summation tree: deterministic)
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int U, int Ng, float* __restrict__ dW1,
                                       float* __restrict__ db1, float* __restrict__ dW2, float* __restrict__ db2) {
  __shared__ float seg[8][33];
  const int total = F * F + F * K1 + 2 * F;
  const int i = blockIdx.x * 32 + threadIdx.x;
  const bool live = i < total;
  const int per = (U + 7) / 8;
  const int u_lo = threadIdx.y * per, u_hi = min(U, u_lo + per);
  float s = 0.0f;
  if (live) {
    if (i < F * F + F * K1) {
      for (int u = u_lo; u < u_hi; ++u) s += partial[(int64_t)u * PART_FLOATS + i];
    } else {
      const int r = i - (F * F + F * K1);      // [0,128): db2, [128,256): db1; two warp halves each
      const int base = F * F + F * K1 + (r / F) * 2 * F + (r % F);
      for (int u = u_lo; u < u_hi; ++u)
        s += partial[(int64_t)u * PART_FLOATS + base] + partial[(int64_t)u * PART_FLOATS + base + F];
    }
  }
  seg[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y != 0 || !live) return;
  s = 0.0f;
#pragma unroll
  for (int y = 0; y < 8; ++y) s += seg[y][threadIdx.x];
  if (i < F * F) {
    dW2[i] = s;
  } else if (i < F * F + F * K1) {
    const int k = (i - F * F) / K1, j = (i - F * F) % K1;
    if (j < Ng) dW1[k * Ng + j] = s;
  } else {
    const int r = i - (F * F + F * K1);
    (r / F == 0 ? db2 : db1)[r % F] = s;
  }
}

__device__ __forceinline__ void pack_bwd_weights_body(const float* __restrict__ W1, const float* __restrict__ b1,
                                                      const float* __restrict__ W2, int Ng, uint8_t* __restrict__ out,
                                                      int idx) {
  if (idx < F * K1) {
    const int m = idx / K1, k = idx % K1;
    const float v = (k < Ng) ? W1[m * Ng + k] : (k == Ng ? b1[m] : 0.0f);
    const uint32_t off = (m & 7) * 16 + (k & 7) * 2 + (m >> 3) * 1024 + (k >> 3) * 128;
    *reinterpret_cast<__nv_bfloat16*>(out + off) = __float2bfloat16_rn(v);
  } else if (idx < F * K1 + F * F) {
    const int j = idx - F * K1;
    const int k = j / F, f = j % F;           // image rows = k (hidden), K = f (output channel): W2^T
    const float v = W2[f * F + k];
    const uint32_t off = (k & 7) * 16 + (f & 7) * 2 + (k >> 3) * 2048 + (f >> 3) * 128;
    *reinterpret_cast<__nv_bfloat16*>(out + W1_BYTES + off) = __float2bfloat16_rn(v);
  }
}

__global__ void pack_bwd_weights_kernel(const float* __restrict__ W1, const float* __restrict__ b1,
                                        const float* __restrict__ W2, int Ng, uint8_t* __restrict__ out) {
  pack_bwd_weights_body(W1, b1, W2, Ng, out, blockIdx.x * blockDim.x + threadIdx.x);
}

// grouped: same job layout as cfconv_tc.cu's PackFilterJob (cmp_pack_filter_job_t of the header)
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
__global__ void pack_bwd_weights_grouped_kernel(const __grid_constant__ PackFilterGroup g, int Ng) {
  const PackFilterJob& j = g.j[blockIdx.y];
  pack_bwd_weights_body(j.W1, j.b1, j.W2, Ng, j.packed_bwd, blockIdx.x * blockDim.x + threadIdx.x);
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ src, int64_t n, __nv_bfloat16* __restrict__ dst) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (; i + 3 < n; i += stride) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    uint2 o = make_uint2(tc::pack_bf16x2(v.x, v.y), tc::pack_bf16x2(v.z, v.w));
    *reinterpret_cast<uint2*>(dst + i) = o;
  }
}

__global__ void expand_rows_kernel(const int32_t* __restrict__ rowptr, int64_t N, int32_t* __restrict__ erow) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) erow[k] = (int32_t)i;
}

// flat tiles: consecutive chunks of <= tile_edges edges of ONE conformer
__global__ void flat_tiles_count_kernel(const int32_t* __restrict__ conf_edge_ptr, int64_t G, int tile_edges,
                                        int32_t* __restrict__ counts) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const int n = conf_edge_ptr[g + 1] - conf_edge_ptr[g];
  counts[g] = (n + tile_edges - 1) / tile_edges;
}

__global__ void flat_tiles_scan_kernel(const int32_t* __restrict__ counts, int64_t G, int32_t* __restrict__ ptr) {
  // single thread block; G is at most a few 10^4 conformers per GPU
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < G; base += blockDim.x) {
    const int64_t i = base + threadIdx.x;
    const int v = (i < G) ? counts[i] : 0;
    // inclusive scan through shared memory (blockDim.x = 256)
    __shared__ int buf[256];
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
      const int t = (threadIdx.x >= o) ? buf[threadIdx.x - o] : 0;
      __syncthreads();
      buf[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < G) ptr[i] = carry + buf[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 255) carry += buf[255];
    __syncthreads();
  }
  if (threadIdx.x == 0) ptr[G] = carry;
}

__global__ void flat_tiles_fill_kernel(const int32_t* __restrict__ conf_edge_ptr, const int32_t* __restrict__ seg_ptr,
                                       const int32_t* __restrict__ erow, int64_t G, int tile_edges,
                                       const int32_t* __restrict__ ptr, int64_t cap, int4* __restrict__ tiles,
                                       int32_t* __restrict__ num_tiles, int* status) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool fits = (int64_t)ptr[G] <= cap;
  if (g == 0) {
    *num_tiles = fits ? ptr[G] : 0;
    if (!fits) atomicOr(status, CMP_STATUS_EDGE_OVERFLOW);
  }
  if (g >= G || !fits) return;
  const int eb = conf_edge_ptr[g], ee = conf_edge_ptr[g + 1];
  int4* out = tiles + 2 * (int64_t)ptr[g];
  int k = 0;
  for (int e0 = eb; e0 < ee; e0 += tile_edges, ++k) {
    const int ne = min(tile_edges, ee - e0);
    out[2 * k] = make_int4(erow[e0], erow[e0 + ne - 1] + 1, seg_ptr[g], seg_ptr[g + 1] - seg_ptr[g]);
    out[2 * k + 1] = make_int4(e0, ne, 0, 0);
  }
}

}  // namespace
}  // namespace cmp

using namespace cmp;

static long long* g_bwd_dbg = nullptr;
extern "C" void cmp_debug_set_bwd_timestamps(void* buf) { g_bwd_dbg = reinterpret_cast<long long*>(buf); }

extern "C" int cmp_csr_expand_rows(const int32_t* rowptr, int64_t N, int32_t* erow, cmp_stream_t stream) {
  CMP_REQUIRE(N >= 0, CMP_EINVAL, "cmp_csr_expand_rows: negative size");
  if (N == 0) return CMP_OK;
  CMP_REQUIRE(rowptr && erow, CMP_EINVAL, "cmp_csr_expand_rows: null pointer");
  expand_rows_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, as_stream(stream)>>>(rowptr, N, erow);
  CMP_LAUNCH_CHECK("cmp_csr_expand_rows");
  return CMP_OK;
}

extern "C" int cmp_cfconv_tc_bwd_tile_edges(void) { return TE; }

extern "C" size_t cmp_build_flat_tiles_workspace(int64_t G) {
  return align_up((size_t)(2 * G + 8) * sizeof(int32_t), 256);
}

extern "C" int cmp_build_flat_tiles(const int32_t* conf_edge_ptr, const int32_t* seg_ptr, const int32_t* erow, int64_t G,
                                    int tile_edges, void* tiles, int64_t cap_tiles, int32_t* num_tiles, void* workspace,
                                    size_t workspace_bytes, int* status, cmp_stream_t stream) {
  CMP_REQUIRE(G >= 0 && tile_edges >= 16 && cap_tiles >= 0, CMP_EINVAL, "cmp_build_flat_tiles: bad size");
  CMP_REQUIRE(num_tiles && status, CMP_EINVAL, "cmp_build_flat_tiles: null pointer");
  cudaStream_t st = as_stream(stream);
  if (G == 0) {
    CMP_REQUIRE(cudaMemsetAsync(num_tiles, 0, sizeof(int32_t), st) == cudaSuccess, CMP_ECUDA,
                "cmp_build_flat_tiles: memset failed");
    return CMP_OK;
  }
  CMP_REQUIRE(conf_edge_ptr && seg_ptr && erow && tiles, CMP_EINVAL, "cmp_build_flat_tiles: null pointer");
  CMP_REQUIRE(workspace && workspace_bytes >= cmp_build_flat_tiles_workspace(G), CMP_EWORKSPACE,
              "cmp_build_flat_tiles: workspace too small");
  int32_t* counts = reinterpret_cast<int32_t*>(workspace);
  int32_t* ptr = counts + G;
  flat_tiles_count_kernel<<<(unsigned)ceil_div(G, 128), 128, 0, st>>>(conf_edge_ptr, G, tile_edges, counts);
  CMP_LAUNCH_CHECK("cmp_build_flat_tiles(count)");
  flat_tiles_scan_kernel<<<1, 256, 0, st>>>(counts, G, ptr);
  CMP_LAUNCH_CHECK("cmp_build_flat_tiles(scan)");
  flat_tiles_fill_kernel<<<(unsigned)ceil_div(G, 128), 128, 0, st>>>(conf_edge_ptr, seg_ptr, erow, G, tile_edges, ptr,
                                                                    cap_tiles, reinterpret_cast<int4*>(tiles), num_tiles,
                                                                    status);
  CMP_LAUNCH_CHECK("cmp_build_flat_tiles(fill)");
  return CMP_OK;
}

extern "C" int cmp_f32_to_bf16(const float* src, int64_t n, void* dst, cmp_stream_t stream) {
  CMP_REQUIRE(n >= 0 && n % 4 == 0, CMP_EINVAL, "cmp_f32_to_bf16: n must be a non-negative multiple of 4");
  if (n == 0) return CMP_OK;
  CMP_REQUIRE(src && dst && ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 8 == 0), CMP_EINVAL,
              "cmp_f32_to_bf16: null or misaligned pointer");
  int64_t blocks = ceil_div(n / 4, 256);
  if (blocks > 2048) blocks = 2048;
  f32_to_bf16_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(src, n, reinterpret_cast<__nv_bfloat16*>(dst));
  CMP_LAUNCH_CHECK("cmp_f32_to_bf16");
  return CMP_OK;
}

extern "C" size_t cmp_cfconv_tc_bwd_weights_bytes(void) { return W1_BYTES + W2T_BYTES; }

extern "C" size_t cmp_cfconv_fused_bwd_workspace(void) {
  return align_up((size_t)sm_count() * NG * PART_FLOATS * sizeof(float), 256);
}

extern "C" int cmp_cfconv_tc_pack_bwd_weights(const float* W1, const float* b1, const float* W2, int num_filters,
                                              int num_gaussians, void* packed, cmp_stream_t stream) {
  CMP_REQUIRE(num_filters == F && num_gaussians >= 1 && num_gaussians < K1, CMP_EUNSUPPORTED,
              "cmp_cfconv_tc_pack_bwd_weights: needs num_filters == 128 and num_gaussians < 64");
  CMP_REQUIRE(W1 && b1 && W2 && packed, CMP_EINVAL, "cmp_cfconv_tc_pack_bwd_weights: null pointer");
  const int total = F * K1 + F * F;
  pack_bwd_weights_kernel<<<(total + 255) / 256, 256, 0, as_stream(stream)>>>(W1, b1, W2, num_gaussians,
                                                                             reinterpret_cast<uint8_t*>(packed));
  CMP_LAUNCH_CHECK("cmp_cfconv_tc_pack_bwd_weights");
  return CMP_OK;
}

template <bool PAIR>
static int launch_fused_bwd(BwdParams p, int num_gaussians, float* dW1, float* db1, float* dW2, float* db2,
                            cudaStream_t st, const char* what) {
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(cfconv_fused_bwd_kernel<PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)SMEM_BYTES) != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("%s: cannot opt in to %u bytes of shared memory", what, SMEM_BYTES);
      return CMP_ECUDA;
    }
    attr_set = true;
  }
  const int grid = sm_count();
  cfconv_fused_bwd_kernel<PAIR><<<grid, CTA_THREADS, SMEM_BYTES, st>>>(p);
  CMP_LAUNCH_CHECK(what);
  const int total = F * F + F * K1 + 2 * F;
  reduce_partials_kernel<<<(total + 31) / 32, dim3(32, 8), 0, st>>>(p.partial, grid * NG, num_gaussians, dW1, db1, dW2,
                                                                    db2);
  CMP_LAUNCH_CHECK(what);
  return CMP_OK;
}

extern "C" int cmp_cfconv_fused_bwd_weights(const float* g, const void* xprime_bf16, const float* dist,
                                            const int32_t* col, const int32_t* erow, const void* flat_tiles,
                                            const int32_t* num_tiles, const void* packed_bwd_weights,
                                            const float* offset, int num_gaussians, float coeff, float cutoff,
                                            int num_filters, float* dW1, float* db1, float* dW2, float* db2,
                                            void* workspace, size_t workspace_bytes, cmp_stream_t stream) {
  CMP_REQUIRE(num_filters == F && num_gaussians >= 1 && num_gaussians < K1, CMP_EUNSUPPORTED,
              "cmp_cfconv_fused_bwd_weights: needs num_filters == 128 and num_gaussians < 64");
  CMP_REQUIRE(g && xprime_bf16 && dist && col && erow && flat_tiles && num_tiles && packed_bwd_weights && offset && dW1 &&
                  db1 && dW2 && db2,
              CMP_EINVAL, "cmp_cfconv_fused_bwd_weights: null pointer");
  CMP_REQUIRE(((uintptr_t)xprime_bf16 % 16 == 0) && ((uintptr_t)packed_bwd_weights % 16 == 0) && ((uintptr_t)g % 16 == 0),
              CMP_EINVAL, "cmp_cfconv_fused_bwd_weights: pointers must be 16-byte aligned");
  CMP_REQUIRE(workspace && workspace_bytes >= cmp_cfconv_fused_bwd_workspace(), CMP_EWORKSPACE,
              "cmp_cfconv_fused_bwd_weights: workspace too small");
  CMP_REQUIRE(cmp_device_is_sm100(), CMP_EUNSUPPORTED, "cmp_cfconv_fused_bwd_weights: needs an sm_100 device (tcgen05)");
  BwdParams p;
  p.g = g;
  p.gb = nullptr;
  p.rev = nullptr;
  p.xprime = reinterpret_cast<const __nv_bfloat16*>(xprime_bf16);
  p.dist = dist;
  p.col = col;
  p.erow = erow;
  p.tiles = reinterpret_cast<const int4*>(flat_tiles);
  p.num_tiles = num_tiles;
  p.weights = reinterpret_cast<const uint8_t*>(packed_bwd_weights);
  p.offset = offset;
  p.partial = reinterpret_cast<float*>(workspace);
  p.coeff_log2e = coeff * 1.4426950408889634f;
  p.cutoff = cutoff;
  p.Ng = num_gaussians;
  p.dbg = g_bwd_dbg;
  return launch_fused_bwd<false>(p, num_gaussians, dW1, db1, dW2, db2, as_stream(stream), "cmp_cfconv_fused_bwd_weights");
}

extern "C" int cmp_cfconv_fused_bwd_weights_pairs(const void* g_bf16, const void* xprime_bf16, const float* pair_dist,
                                                  const int32_t* pair_src, const int32_t* pair_dst,
                                                  const int32_t* pair_rev, const void* pair_tiles,
                                                  const int32_t* num_tiles, const void* packed_bwd_weights,
                                                  const float* offset, int num_gaussians, float coeff, float cutoff,
                                                  int num_filters, float* dW1, float* db1, float* dW2, float* db2,
                                                  void* workspace, size_t workspace_bytes, cmp_stream_t stream) {
  CMP_REQUIRE(num_filters == F && num_gaussians >= 1 && num_gaussians < K1, CMP_EUNSUPPORTED,
              "cmp_cfconv_fused_bwd_weights_pairs: needs num_filters == 128 and num_gaussians < 64");
  CMP_REQUIRE(g_bf16 && xprime_bf16 && pair_dist && pair_src && pair_dst && pair_rev && pair_tiles && num_tiles &&
                  packed_bwd_weights && offset && dW1 && db1 && dW2 && db2,
              CMP_EINVAL, "cmp_cfconv_fused_bwd_weights_pairs: null pointer");
  CMP_REQUIRE(((uintptr_t)xprime_bf16 % 16 == 0) && ((uintptr_t)packed_bwd_weights % 16 == 0) &&
                  ((uintptr_t)g_bf16 % 16 == 0),
              CMP_EINVAL, "cmp_cfconv_fused_bwd_weights_pairs: pointers must be 16-byte aligned");
  CMP_REQUIRE(workspace && workspace_bytes >= cmp_cfconv_fused_bwd_workspace(), CMP_EWORKSPACE,
              "cmp_cfconv_fused_bwd_weights_pairs: workspace too small");
  CMP_REQUIRE(cmp_device_is_sm100(), CMP_EUNSUPPORTED,
              "cmp_cfconv_fused_bwd_weights_pairs: needs an sm_100 device (tcgen05)");
  BwdParams p;
  p.g = nullptr;
  p.gb = reinterpret_cast<const __nv_bfloat16*>(g_bf16);
  p.rev = pair_rev;
  p.xprime = reinterpret_cast<const __nv_bfloat16*>(xprime_bf16);
  p.dist = pair_dist;
  p.col = pair_src;
  p.erow = pair_dst;
  p.tiles = reinterpret_cast<const int4*>(pair_tiles);
  p.num_tiles = num_tiles;
  p.weights = reinterpret_cast<const uint8_t*>(packed_bwd_weights);
  p.offset = offset;
  p.partial = reinterpret_cast<float*>(workspace);
  p.coeff_log2e = coeff * 1.4426950408889634f;
  p.cutoff = cutoff;
  p.Ng = num_gaussians;
  p.dbg = g_bwd_dbg;
  return launch_fused_bwd<true>(p, num_gaussians, dW1, db1, dW2, db2, as_stream(stream),
                                "cmp_cfconv_fused_bwd_weights_pairs");
}

extern "C" int cmp_cfconv_tc_pack_bwd_weights_grouped(const void* jobs, int count, int num_filters, int num_gaussians,
                                                      cmp_stream_t stream) {
  CMP_REQUIRE(num_filters == F && num_gaussians >= 1 && num_gaussians < K1, CMP_EUNSUPPORTED,
              "cmp_cfconv_tc_pack_bwd_weights_grouped: needs num_filters == 128 and num_gaussians < 64");
  CMP_REQUIRE(count >= 0 && count <= MAX_PACK_JOBS, CMP_EINVAL,
              "cmp_cfconv_tc_pack_bwd_weights_grouped: count must be in [0, %d]", MAX_PACK_JOBS);
  if (count == 0) return CMP_OK;
  CMP_REQUIRE(jobs, CMP_EINVAL, "cmp_cfconv_tc_pack_bwd_weights_grouped: null pointer");
  const PackFilterJob* in = reinterpret_cast<const PackFilterJob*>(jobs);
  PackFilterGroup g;
  for (int i = 0; i < count; ++i) {
    CMP_REQUIRE(in[i].W1 && in[i].b1 && in[i].W2 && in[i].packed_bwd, CMP_EINVAL,
                "cmp_cfconv_tc_pack_bwd_weights_grouped: null pointer");
    g.j[i] = in[i];
  }
  const int total = F * K1 + F * F;
  pack_bwd_weights_grouped_kernel<<<dim3((total + 255) / 256, count), 256, 0, as_stream(stream)>>>(g, num_gaussians);
  CMP_LAUNCH_CHECK("cmp_cfconv_tc_pack_bwd_weights_grouped");
  return CMP_OK;
}
