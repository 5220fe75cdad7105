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

// fp32-grade images: W1 hi | W1 lo | W2^T hi | W2^T lo (lo = bf16(v - hi)), layouts of pack_bwd_weights_body
__device__ __forceinline__ void pack_bwd_weights_x3_body(const float* __restrict__ W1, const float* __restrict__ b1,
                                                         const float* __restrict__ W2, int Ng,
                                                         uint8_t* __restrict__ out, int idx) {
  float v;
  uint32_t off_hi, off_lo;
  if (idx < F * K1) {
    const int m = idx / K1, k = idx % K1;
    v = (k < Ng) ? W1[m * Ng + k] : (k == Ng ? b1[m] : 0.0f);
    off_hi = (m & 7) * 16 + (k & 7) * 2 + (m >> 3) * 1024 + (k >> 3) * 128;
    off_lo = off_hi + W1_BYTES;
  } else if (idx < F * K1 + F * F) {
    const int j = idx - F * K1;
    const int k = j / F, f = j % F;
    v = W2[f * F + k];
    off_hi = 2 * W1_BYTES + (k & 7) * 16 + (f & 7) * 2 + (k >> 3) * 2048 + (f >> 3) * 128;
    off_lo = off_hi + W2T_BYTES;
  } else {
    return;
  }
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  *reinterpret_cast<__nv_bfloat16*>(out + off_hi) = hi;
  *reinterpret_cast<__nv_bfloat16*>(out + off_lo) = __float2bfloat16_rn(v - __bfloat162float(hi));
}

__global__ void pack_bwd_weights_kernel(const float* __restrict__ W1, const float* __restrict__ b1,
                                        const float* __restrict__ W2, int Ng, uint8_t* __restrict__ out) {
  pack_bwd_weights_body(W1, b1, W2, Ng, out, blockIdx.x * blockDim.x + threadIdx.x);
}

struct PackBwdX3Job {
  const float* W1;
  const float* b1;
  const float* W2;
  uint8_t* packed;
};
struct PackBwdX3Group {
  PackBwdX3Job j[32];
};
__global__ void pack_bwd_weights_x3_kernel(const __grid_constant__ PackBwdX3Group g, int Ng) {
  const PackBwdX3Job& j = g.j[blockIdx.y];
  pack_bwd_weights_x3_body(j.W1, j.b1, j.W2, Ng, j.packed, blockIdx.x * blockDim.x + threadIdx.x);
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


// =================================================================================================================
// Dense-block weight gradients (conformers of <= 128 atoms): same phases, columns = the pairs of 16 x 16 atom blocks
// =================================================================================================================
// cfconv_fused_bwd_kernel<pairs> reads a pair list: per column and channel four gathers from bf16 copies of g and x'
// staged per conformer (two extra conversion launches per layer), plus per-column metadata.  Here the columns of a tile
// are the pairs of the dense blocks of cfconv_dense.cu, cut into tiles of 64: DIAG(b) columns [0, 64) and [64, 120),
// RECT(b, j0) = block b x 4 later atoms.  The pair of a column is a compile-time function of its index, so the thread
// that owns channel f keeps g[16 rows][f] and x'[16 rows][f] of the row block in fp32 registers and builds
//     dF[f, (i, j)] = [j -> i] g[i] x'[j] + [i -> j] g[j] x'[i]
// with two multiply-adds per column - no gathers, no staging, no bf16 copies (the products are rounded once, to the
// bf16 dF image).  Distances come from the positions, the two directions of a pair from the adjacency bit matrix
// (cmp_build_adjacency).  Everything after dF (the MMAs, epilogues 1 and 3, TMEM accumulators, partial sums) is the
// pipeline of cfconv_fused_bwd_kernel.
constexpr int DN_MAX = 128;        // atoms per conformer
constexpr int DN_AW = 4;           // adjacency words per atom
// the rbf image R and the dF image F are double buffered (tile parity): the Gaussians and dF of tile k + 1 are written
// while the weight-gradient MMAs of tile k still read R / F of tile k
constexpr uint32_t D_OFF_A = 2 * R_BYTES;
constexpr uint32_t D_OFF_F = D_OFF_A + CH_BYTES;
constexpr uint32_t D_OFF_S = D_OFF_F + 2 * CH_BYTES;
constexpr uint32_t D_OFF_POS = D_OFF_S + CH_BYTES;                 // float[128][3]
constexpr uint32_t D_OFF_ADJ = D_OFF_POS + DN_MAX * 12;            // uint32[128][4]
constexpr uint32_t D_OFF_C = D_OFF_ADJ + DN_MAX * DN_AW * 4;       // float[64]
constexpr uint32_t D_OFF_MASK = D_OFF_C + TE * 4;                  // uint32[4]: [j -> i] of columns 0..31, 32..63; [i -> j] likewise
constexpr uint32_t D_GROUP_BYTES = (D_OFF_MASK + 16 + 127) / 128 * 128;
constexpr uint32_t D_SMEM_BYTES = W1_BYTES + W2T_BYTES + NG * D_GROUP_BYTES;
static_assert(D_SMEM_BYTES <= 232448 - 1024, "shared memory budget of the dense weight-gradient kernel");

// column c of the DIAG pairs of a block holds (il, jl), il < jl, c = jl (jl - 1) / 2 + il
__host__ __device__ constexpr int dg_j(int c) {
  return 1 + (c >= 1) + (c >= 3) + (c >= 6) + (c >= 10) + (c >= 15) + (c >= 21) + (c >= 28) + (c >= 36) + (c >= 45) +
         (c >= 55) + (c >= 66) + (c >= 78) + (c >= 91) + (c >= 105);
}
__host__ __device__ constexpr int dg_i(int c) { return c - dg_j(c) * (dg_j(c) - 1) / 2; }

__host__ __device__ inline int dense_block_tiles(int n, int a0) {
  const int m = (n - a0) < 16 ? (n - a0) : 16;
  const int nd = m * (m - 1) / 2;
  const int rest = n - a0 - 16;
  return (nd > 0) + (nd > 64) + (rest > 0 ? (rest + 3) / 4 : 0);
}
__host__ __device__ inline int dense_conf_tiles(int n) {
  if (n > DN_MAX || n <= 0) return 0;
  int t = 0;
  for (int a0 = 0; a0 < n; a0 += 16) t += dense_block_tiles(n, a0);
  return t;
}

// tile_ptr[g] = tiles of the conformers before g (single block; G is at most a few 10^4 conformers per GPU)
__global__ void dense_tile_ptr_kernel(const int32_t* __restrict__ seg_ptr, int64_t G, int32_t* __restrict__ ptr,
                                      int* status) {
  __shared__ int carry;
  __shared__ int buf[256];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < G; base += blockDim.x) {
    const int64_t i = base + threadIdx.x;
    int v = 0;
    if (i < G) {
      const int n = seg_ptr[i + 1] - seg_ptr[i];
      if (n > DN_MAX && status) atomicOr(status, CMP_STATUS_EDGE_OVERFLOW);
      v = dense_conf_tiles(n);
    }
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

// a' = c ssp(x) and S = c sigmoid(x) of two columns, in packed f16x2, returned as bf16x2 image words:
//   t = 2^(-|x| log2 e);  ssp(x) = max(x, 0) + ln2 (log2(1 + t) - 1),  log2(1 + t) = t Q4(t) (8e-5 on [0, 1]);
//   sigmoid(x) = 1/2 + tanh(x / 2) / 2.
__device__ __forceinline__ void ssp_sigmoid_f16x2(float x0, float x1, float c0, float c1, uint32_t& a_bf, uint32_t& s_bf) {
  const __half2 x = __floats2half2_rn(x0, x1);
  const __half2 c = __floats2half2_rn(c0, c1);
  uint32_t tu, au = *reinterpret_cast<const uint32_t*>(&x);
  const __half2 arg = __hmul2(__habs2(x), __float2half2_rn(-1.4426950408889634f));
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(tu) : "r"(*reinterpret_cast<const uint32_t*>(&arg)));
  const __half2 t = *reinterpret_cast<const __half2*>(&tu);
  __half2 q = __hfma2(__float2half2_rn(0.0599455865f), t, __float2half2_rn(-0.227712643f));
  q = __hfma2(q, t, __float2half2_rn(0.442274178f));
  q = __hfma2(q, t, __float2half2_rn(-0.717063932f));
  q = __hfma2(q, t, __float2half2_rn(1.44261568f));
  const __half2 l2m1 = __hfma2(q, t, __float2half2_rn(-1.0f));                         // log2(1 + t) - 1
  const __half2 sp = __hfma2(l2m1, __float2half2_rn(0.6931471805599453f), __hmax2(x, __float2half2_rn(0.0f)));
  const float2 af = __half22float2(__hmul2(sp, c));
  a_bf = tc::pack_bf16x2(af.x, af.y);
  const __half2 hx = __hmul2(x, __float2half2_rn(0.5f));
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(au) : "r"(*reinterpret_cast<const uint32_t*>(&hx)));
  const __half2 sg = __hfma2(*reinterpret_cast<const __half2*>(&au), __float2half2_rn(0.5f), __float2half2_rn(0.5f));
  const float2 sf = __half22float2(__hmul2(sg, c));
  s_bf = tc::pack_bf16x2(sf.x, sf.y);
}

struct DenseBwdParams {
  const float* g;                 // [N, F] dL/dagg (fp32)
  const float* xprime;            // [N, F] x' (fp32)
  const float* pos;               // [N, 3]
  const int32_t* seg_ptr;         // [G + 1]
  const uint32_t* adj;            // [N, 4]
  const int32_t* tile_ptr;        // [G + 1]
  const uint8_t* weights;         // W1aug image | W2^T image (cmp_cfconv_tc_pack_bwd_weights)
  const float* offset;
  float* partial;                 // [gridDim.x * NG][PART_FLOATS]
  float coeff_log2e;
  float cutoff;
  int Ng;
  int G;
};

// position of a pipeline in its tile sequence
struct DenseWalk {
  int conf, cs, n, a0, lt, nt;    // conformer, first atom, atoms; row block; local tile and tiles of the block
  __device__ __forceinline__ void load_conf(const DenseBwdParams& p) {
    cs = __ldg(p.seg_ptr + conf);
    n = __ldg(p.seg_ptr + conf + 1) - cs;
    if (n > DN_MAX || n <= 0) n = 0;
  }
  // position at global tile t (t < total)
  __device__ __forceinline__ void seek(const DenseBwdParams& p, int64_t t) {
    int lo = 0, hi = p.G;            // last conformer with tile_ptr[conf] <= t
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if ((int64_t)__ldg(p.tile_ptr + mid) <= t) lo = mid; else hi = mid;
    }
    conf = lo;
    load_conf(p);
    int local = (int)(t - __ldg(p.tile_ptr + conf));
    a0 = 0;
    nt = dense_block_tiles(n, 0);
    while (local >= nt) {
      local -= nt;
      a0 += 16;
      nt = dense_block_tiles(n, a0);
    }
    lt = local;
  }
  __device__ __forceinline__ void next(const DenseBwdParams& p) {
    if (++lt < nt) return;
    lt = 0;
    for (;;) {
      a0 += 16;
      if (a0 < n) {
        nt = dense_block_tiles(n, a0);
        if (nt > 0) return;
        continue;
      }
      if (++conf >= p.G) {        // past the end: never used (callers count tiles)
        n = 0; a0 = 0; nt = 1;
        return;
      }
      load_conf(p);
      a0 = -16;
    }
  }
  // geometry of the current tile: kind 0..3 = DIAG columns [32 kind', ...) pairs of 64: c_base = 64 * (kind >> 1); 4 = RECT
  __device__ __forceinline__ void tile(bool& diag, int& c_base, int& j0, int& ncols) const {
    const int m = min(16, n - a0);
    const int nd = m * (m - 1) / 2;
    const int ndt = (nd > 0) + (nd > 64);
    diag = lt < ndt;
    if (diag) {
      c_base = 64 * lt;
      j0 = a0;
      ncols = min(64, nd - c_base);
    } else {
      c_base = 0;
      j0 = a0 + 16 + 4 * (lt - ndt);
      ncols = 16 * min(4, n - j0);
    }
  }
};

// dF of 32 columns of a DIAG tile: columns C0 .. C0 + 31 of the block's pair list (compile-time pairs)
// bf16 hi + lo images of 8 values (lo = bf16(v - hi): 16 significant bits together)
__device__ __forceinline__ void split_bf16x8(const float* v, uint4& hi, uint4& lo) {
  hi = pack_bf16x8(v);
  float hf[8], r[8];
  unpack_bf16x8(hi, hf);
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = v[j] - hf[j];
  lo = pack_bf16x8(r);
}
// LO_OFF != 0: also write the lo image, LO_OFF bytes behind the hi image (fp32-grade kernel)
template <uint32_t LO_OFF>
__device__ __forceinline__ void store_df8(uint8_t* dst, const float* v) {
  if (LO_OFF == 0) {
    *reinterpret_cast<uint4*>(dst) = pack_bf16x8(v);
  } else {
    uint4 hi, lo;
    split_bf16x8(v, hi, lo);
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + LO_OFF) = lo;
  }
}

template <int C0, uint32_t LO_OFF = 0>
__device__ __forceinline__ void dense_df_diag(const float (&gr)[16], const float (&xr)[16], uint32_t mf, uint32_t mr,
                                              const float* __restrict__ sC32, uint8_t* __restrict__ dstF, float& db2) {
#pragma unroll
  for (int c8 = 0; c8 < 32; c8 += 8) {
    float v[8];
    const float4 ca = *reinterpret_cast<const float4*>(sC32 + c8), cb = *reinterpret_cast<const float4*>(sC32 + c8 + 4);
    const float cc[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      constexpr int dummy = 0;
      (void)dummy;
      const int c = C0 + c8 + j;                     // compile-time after unrolling
      const int il = c < 120 ? dg_i(c) : 0, jl = c < 120 ? dg_j(c) : 0;
      const float a = ((mf >> (c8 + j)) & 1u) ? gr[il] * xr[jl] : 0.0f;       // edge j -> i: g[i] x'[j]
      v[j] = ((mr >> (c8 + j)) & 1u) ? fmaf(gr[jl], xr[il], a) : a;           // edge i -> j: g[j] x'[i]
      db2 = fmaf(v[j], cc[j], db2);
    }
    store_df8<LO_OFF>(dstF + (c8 >> 3) * 2048, v);
  }
}

// dF of 32 columns of a RECT tile: rows il of the block x the two column atoms of this half (gj / xj)
template <uint32_t LO_OFF = 0>
__device__ __forceinline__ void dense_df_rect(const float (&gr)[16], const float (&xr)[16], const float (&gj)[2],
                                              const float (&xj)[2], uint32_t mf, uint32_t mr,
                                              const float* __restrict__ sC32, uint8_t* __restrict__ dstF, float& db2) {
#pragma unroll
  for (int c8 = 0; c8 < 32; c8 += 8) {
    float v[8];
    const float4 ca = *reinterpret_cast<const float4*>(sC32 + c8), cb = *reinterpret_cast<const float4*>(sC32 + c8 + 4);
    const float cc[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c8 + j, il = c & 15, jj = c >> 4;
      const float a = ((mf >> c) & 1u) ? gr[il] * xj[jj] : 0.0f;              // edge j -> i
      v[j] = ((mr >> c) & 1u) ? fmaf(gj[jj], xr[il], a) : a;                  // edge i -> j
      db2 = fmaf(v[j], cc[j], db2);
    }
    store_df8<LO_OFF>(dstF + (c8 >> 3) * 2048, v);
  }
}

__global__ void __launch_bounds__(CTA_THREADS, 1) cfconv_dense_bwd_kernel(const __grid_constant__ DenseBwdParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ float s_db[NG * 2 * 2 * F];     // db2 | db1 partials of (group, half), combined in the drain
  // wbar | per group: r_ready, d1_ready, f_ready, dda_ready, h_ready, w_done
  __shared__ uint64_t bars[1 + NG * 6];
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
      uint64_t* b = &bars[1 + g * 6];
      tc::mbar_init(b + 0, GT);  // r_ready   (rbf image, cutoffs and masks written)
      tc::mbar_init(b + 1, 1);   // d1_ready  (h in TMEM)
      tc::mbar_init(b + 2, GT);  // f_ready   (a', S, dF images written; h consumed)
      tc::mbar_init(b + 3, 1);   // dda_ready (da' in TMEM)
      tc::mbar_init(b + 4, GT);  // h_ready   (dh image written)
      tc::mbar_init(b + 5, 1);   // w_done    (weight-gradient MMAs finished reading the images)
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

  const int64_t T = __ldg(p.tile_ptr + p.G);
  const int64_t U = (int64_t)gridDim.x * NG;
  const int k1steps = (p.Ng + 1 + 15) >> 4;

  if (warp >= NG * (GT / 32)) {
    // ======================= MMA-issuing warp of group g =======================
    const int g = warp - NG * (GT / 32);
    if (lane == 0) {
      uint64_t* wbar = &bars[0];
      uint64_t* b = &bars[1 + g * 6];
      if (g == 0) {
        tc::mbar_arrive_expect_tx(wbar, W1_BYTES + W2T_BYTES);
        tc::bulk_g2s(sW1, p.weights, W1_BYTES, wbar);
        tc::bulk_g2s(sW2T, p.weights + W1_BYTES, W2T_BYTES, wbar);
      }
      uint8_t* sG0 = smem + W1_BYTES + W2T_BYTES + g * D_GROUP_BYTES;
      const uint32_t aW1 = tc::smem_u32(sW1), aW2T = tc::smem_u32(sW2T);
      const uint32_t aR0 = tc::smem_u32(sG0), aA = aR0 + D_OFF_A, aF0 = aR0 + D_OFF_F, aS = aR0 + D_OFF_S;
      const uint32_t tD = tmem_base + g * 256, tW2 = tD + 64, tW1 = tD + 192;
      const int64_t u = (int64_t)blockIdx.x * NG + g;
      const int64_t t0 = u * T / U, t1 = (u + 1) * T / U;
      tc::mbar_wait(wbar, 0);
      DenseWalk w;
      if (t0 < t1) w.seek(p, t0);
      uint32_t it = 0;
      for (int64_t ti = t0; ti < t1; ++ti, ++it) {
        bool diag;
        int c_base, j0, ncols;
        w.tile(diag, c_base, j0, ncols);
        const int npad = (ncols + 15) & ~15;
        if (ti + 1 < t1) w.next(p);
        const uint32_t par = it & 1;
        const uint32_t aR = aR0 + par * R_BYTES, aF = aF0 + par * CH_BYTES;
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
        // weight gradients, K = columns of the tile (images read as K-major [rows=channel, K=e]: SBO 128, LBO 2048)
        tc::mbar_wait_spin(b + 4, par);
        tc::tc_fence_after();
        const uint32_t id3 = tc::umma_idesc_f16(F, F, 1, 0, 0);
        const uint32_t id4 = tc::umma_idesc_f16(F, K1, 1, 0, 1);
        for (int ks = 0; ks < (npad >> 4); ++ks) {
          const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
          tc::umma_f16(tW2, tc::umma_smem_desc(aF + ks * 4096, 2048, 128), tc::umma_smem_desc(aA + ks * 4096, 2048, 128), id3,
                       acc);
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
    const int h = (warp >> 2) & 1;            // column half handled in the channel-major phases
    const int chan = wq * 32 + lane;
    const int e = tt & 63;                    // column in the rbf phase
    const int q = tt >> 6;                    // 4 threads share a column in the rbf phase
    uint64_t* b = &bars[1 + g * 6];
    uint8_t* sG = smem + W1_BYTES + W2T_BYTES + g * D_GROUP_BYTES;
    uint8_t* sA = sG + D_OFF_A;
    uint8_t* sS = sG + D_OFF_S;
    float* sPos = reinterpret_cast<float*>(sG + D_OFF_POS);
    uint32_t* sAdj = reinterpret_cast<uint32_t*>(sG + D_OFF_ADJ);
    float* sC = reinterpret_cast<float*>(sG + D_OFF_C);
    uint32_t* sMask = reinterpret_cast<uint32_t*>(sG + D_OFF_MASK);
    const uint32_t tD = tmem_base + g * 256 + ((uint32_t)(wq * 32) << 16);
    const uint32_t tW2 = tD + 64, tW1 = tD + 192;
    const int64_t u = (int64_t)blockIdx.x * NG + g;
    const int64_t t0 = u * T / U, t1 = (u + 1) * T / U;
    const float c2 = p.coeff_log2e, cutoff = p.cutoff;
    (void)c2;

    float db1 = 0.0f, db2 = 0.0f;
    float gr[16], xr[16];                     // g and x' rows of the current row block, this thread's channel
#pragma unroll
    for (int i = 0; i < 16; ++i) gr[i] = xr[i] = 0.0f;
    int rows_conf = -1, rows_a0 = -1, staged_conf = -1;
    uint32_t it = 0;
    DenseWalk w;
    if (t0 < t1) w.seek(p, t0);

    for (int64_t ti = t0; ti < t1; ++ti, ++it) {
      bool diag;
      int c_base, j0, ncols;
      w.tile(diag, c_base, j0, ncols);
      const int cs = w.cs, n = w.n, a0 = w.a0, conf = w.conf;
      const int m = min(16, n - a0);
      const int npad = (ncols + 15) & ~15;
      const uint32_t par = it & 1;
      uint8_t* sR = sG + par * R_BYTES;                  // this tile's rbf and dF images
      uint8_t* sF = sG + D_OFF_F + par * CH_BYTES;
      const int goff = (cs + a0) * F + chan;

      // rows of a new row block (fp32, straight from global memory: coalesced over the channel)
      if (conf != rows_conf || a0 != rows_a0) {
        rows_conf = conf;
        rows_a0 = a0;
#pragma unroll
        for (int il = 0; il < 16; ++il) {
          gr[il] = (il < m) ? __ldg(p.g + goff + il * F) : 0.0f;
          xr[il] = (il < m) ? __ldg(p.xprime + goff + il * F) : 0.0f;
        }
      }
      // the two column atoms of this half of a RECT tile
      float gj[2] = {0.0f, 0.0f}, xj[2] = {0.0f, 0.0f};
      if (!diag) {
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int a = j0 + 2 * h + jj;
          if (a < n) {
            gj[jj] = __ldg(p.g + (cs + a) * F + chan);
            xj[jj] = __ldg(p.xprime + (cs + a) * F + chan);
          }
        }
      }

      // positions and adjacency rows of a new conformer (the group's private copy)
      if (conf != staged_conf) {
        staged_conf = conf;
        tc::named_bar_sync(1 + g, GT);       // nobody still reads the previous conformer's copy
        if (tt < n) {
          const float* pp = p.pos + (int64_t)(cs + tt) * 3;
          sPos[3 * tt + 0] = __ldg(pp + 0);
          sPos[3 * tt + 1] = __ldg(pp + 1);
          sPos[3 * tt + 2] = __ldg(pp + 2);
          reinterpret_cast<uint4*>(sAdj)[tt] = __ldg(reinterpret_cast<const uint4*>(p.adj) + cs + tt);
        }
        tc::named_bar_sync(1 + g, GT);
      }

      // ---- column e: pair, directions, distance, cutoff, Gaussian expansion -> rbf image ----
      {
        int il, jl, i_loc, j_loc;
        if (diag) {
          const int c = min(c_base + e, 119);
          // jl = floor((1 + sqrt(1 + 8 c)) / 2): exact for c < 120 with one correction step each way
          jl = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)c)) * 0.5f);
          if (jl * (jl - 1) / 2 > c) --jl;
          if ((jl + 1) * jl / 2 <= c) ++jl;
          il = c - jl * (jl - 1) / 2;
          i_loc = a0 + il;
          j_loc = a0 + jl;
        } else {
          il = e & 15;
          i_loc = a0 + il;
          j_loc = j0 + (e >> 4);
        }
        bool ef = false, er = false;
        if (e < ncols) {
          ef = (sAdj[i_loc * DN_AW + (j_loc >> 5)] >> (j_loc & 31)) & 1u;     // edge j -> i
          er = (sAdj[j_loc * DN_AW + (i_loc >> 5)] >> (i_loc & 31)) & 1u;     // edge i -> j
        }
        const bool live = ef || er;
        float d = 0.0f;
        if (live) {
          const float dx = sPos[3 * j_loc] - sPos[3 * i_loc], dy = sPos[3 * j_loc + 1] - sPos[3 * i_loc + 1],
                      dz = sPos[3 * j_loc + 2] - sPos[3 * i_loc + 2];
          const float d2 = dx * dx + dy * dy + dz * dz;
          d = d2 * rsqrtf(fmaxf(d2, 1e-20f));
        }
        if (q == 0) {       // warps 0 and 1 of the group hold columns 0..31 and 32..63 (whole warps: the ballots are uniform)
          sC[e] = live ? 0.5f * (__cosf(d * kPi / cutoff) + 1.0f) : 0.0f;
          const unsigned bf = __ballot_sync(0xffffffffu, ef), br = __ballot_sync(0xffffffffu, er);
          if (lane == 0) {
            sMask[e >> 5] = bf;
            sMask[2 + (e >> 5)] = br;
          }
        }
        if (e < npad) {
          uint8_t* rowp = sR + (e >> 3) * 1024 + (e & 7) * 16;
          // exp2(c2_k (d - mu_k)^2), c2_k = 0 beyond the Gaussians (bias column = 1).  Rows of missing pairs and of padded
          // columns must be ZERO: they enter dW1 through K = columns.
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
      }
      tc::fence_proxy_async();
      tc::mbar_arrive(b + 0);
      tc::named_bar_sync(1 + g, GT);   // sC / sMask visible to the whole group

      // ---- dF image from the row registers (runs while the tensor core computes h) ----
      {
        const uint32_t mf = sMask[h], mr = sMask[2 + h];
        uint8_t* dstF = sF + chan * 16 + (h * 4) * 2048;
        const float* sC32 = sC + 32 * h;
        if (32 * h < npad) {
          if (!diag) {
            dense_df_rect(gr, xr, gj, xj, mf, mr, sC32, dstF, db2);
          } else {
            switch ((c_base >> 5) + h) {
              case 0: dense_df_diag<0>(gr, xr, mf, mr, sC32, dstF, db2); break;
              case 1: dense_df_diag<32>(gr, xr, mf, mr, sC32, dstF, db2); break;
              case 2: dense_df_diag<64>(gr, xr, mf, mr, sC32, dstF, db2); break;
              default: dense_df_diag<96>(gr, xr, mf, mr, sC32, dstF, db2); break;
            }
          }
        }
      }

      // ---- epilogue 1: a' = C ssp(h), S = C sigmoid(h) -> images (fp32 arithmetic: a packed f16x2 version with one
      // special-function op per element was 2 % faster but tripled the error of dW2: 3.2e-3 -> 8.6e-3 vs the oracle) ----
      // (the a' / S images are single: the previous tile's weight-gradient MMAs must be done reading them - by now they
      // have had the whole Gaussian + dF phase of this tile to finish)
      if (it > 0) tc::mbar_wait(b + 5, (it - 1) & 1);
      tc::mbar_wait(b + 1, par);
      tc::tc_fence_after();
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
          float a[16], sg[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float x = v[j];
            const float t = tc::fast_ex2(-1.4426950408889634f * fabsf(x));
            const float inv = __fdividef(1.0f, 1.0f + t);
            a[j] = c[j] * fmaf(tc::fast_lg2(1.0f + t) - 1.0f, kLn2, fmaxf(x, 0.0f));
            sg[j] = c[j] * (x >= 0.0f ? inv : t * inv);
          }
          *reinterpret_cast<uint4*>(sA + chan * 16 + (c0 >> 3) * 2048) = pack_bf16x8(a);
          *reinterpret_cast<uint4*>(sA + chan * 16 + ((c0 >> 3) + 1) * 2048) = pack_bf16x8(a + 8);
          *reinterpret_cast<uint4*>(sS + chan * 16 + (c0 >> 3) * 2048) = pack_bf16x8(sg);
          *reinterpret_cast<uint4*>(sS + chan * 16 + ((c0 >> 3) + 1) * 2048) = pack_bf16x8(sg + 8);
        }
      }
      tc::tc_fence_before();
      tc::fence_proxy_async();
      tc::mbar_arrive(b + 2);

      // ---- epilogue 3: dh = da' * S -> dh image (in place over S) ----
      tc::mbar_wait(b + 3, par);
      tc::tc_fence_after();
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
      tc::tc_fence_before();
      tc::fence_proxy_async();
      tc::mbar_arrive(b + 4);
      if (ti + 1 < t1) w.next(p);
    }

    // ---- drain: the accumulators of BOTH groups are added in registers (same TMEM lanes = same channel rows, a thread
    // reads the other group's columns as easily as its own) and leave the CTA as ONE partial block: half the partial
    // traffic and half the work of the reduction kernel.  Group g, half h owns a quarter of the columns.
    const bool any = t0 < t1;
    if (any) {
      tc::mbar_wait(b + 5, (it - 1) & 1);
      tc::tc_fence_after();
    }
    s_db[(g * 2 + h) * 2 * F + chan] = db2;
    s_db[(g * 2 + h) * 2 * F + F + chan] = db1;
    tc::tc_fence_before();
    tc::named_bar_sync(NG + 1, NG * GT);      // both groups: every MMA finished, every db partial visible
    tc::tc_fence_after();
    bool any_g[NG];
#pragma unroll
    for (int gg = 0; gg < NG; ++gg) {
      const int64_t uu = (int64_t)blockIdx.x * NG + gg;
      any_g[gg] = uu * T / U < (uu + 1) * T / U;
    }
    float* part = p.partial + (int64_t)blockIdx.x * PART_FLOATS;
    const uint32_t tLane = tmem_base + ((uint32_t)(wq * 32) << 16);
    const int qd = g * 2 + h;                                    // column quarter of this warp
    for (int c0 = qd * 32; c0 < qd * 32 + 32; c0 += 16) {        // dW2[f = chan][k]
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 0.0f;
#pragma unroll
      for (int gg = 0; gg < NG; ++gg) {
        if (any_g[gg]) {
          float t[16];
          tc::tmem_ld16(tLane + gg * 256 + 64 + c0, t);
          tc::tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += t[j];
        }
      }
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(part + chan * F + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    {                                                            // dW1[k = chan][j]: 16 of the 64 columns
      const int c0 = qd * 16;
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 0.0f;
#pragma unroll
      for (int gg = 0; gg < NG; ++gg) {
        if (any_g[gg]) {
          float t[16];
          tc::tmem_ld16(tLane + gg * 256 + 192 + c0, t);
          tc::tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += t[j];
        }
      }
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(part + F * F + chan * K1 + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    if (g == 0) {       // db2 / db1: the two halves stay separate slots (the reduction adds them), the groups are added here
      part[F * F + F * K1 + h * F + chan] = s_db[(0 * 2 + h) * 2 * F + chan] + s_db[(1 * 2 + h) * 2 * F + chan];
      part[F * F + F * K1 + 2 * F + h * F + chan] = s_db[(0 * 2 + h) * 2 * F + F + chan] + s_db[(1 * 2 + h) * 2 * F + F + chan];
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}


// =================================================================================================================
// fp32-grade dense weight gradients ("x3"): bf16 hi + lo operand images, three MMA passes per product
// =================================================================================================================
// Same tiles, column <-> pair maps, register-resident dF and TMEM accumulators as cfconv_dense_bwd_kernel, but every
// operand (Gaussians, W1, W2^T, a', dF, dh) exists as a bf16 hi and a bf16 lo image (lo = bf16(v - hi): 16 significant
// bits together; bf16 rather than f16 because dF / dh carry the dynamic range of the loss gradient) and every product
// runs as hi hi + lo hi + hi lo into the same accumulator.  The per-column rounding errors (2^-17 relative, unbiased)
// average out over the hundreds of thousands of columns a weight gradient sums, so dW1 / dW2 meet the 1e-5 bar of the
// exact path.  Epilogues in fp32; distances and cutoffs as the exact kernels compute them; the pre-activations h stay
// in TMEM next to da', so epilogue 3 recomputes sigmoid(h) instead of reading a rounded image.
// The doubled images leave room for ONE group of 256 threads per CTA, single buffered: the phases of a tile run in
// sequence and only the dF build overlaps the first product.
namespace bx3 {
constexpr int THREADS = GT + 32;
constexpr uint32_t OFF_W1L = W1_BYTES, OFF_W2H = 2 * W1_BYTES, OFF_W2L = 2 * W1_BYTES + W2T_BYTES;
constexpr uint32_t W_BYTES = 2 * (W1_BYTES + W2T_BYTES);          // W1 hi | W1 lo | W2^T hi | W2^T lo
constexpr uint32_t OFF_R = 0;                                       // rbf hi | lo
constexpr uint32_t OFF_A = OFF_R + 2 * R_BYTES;                     // a'  hi | lo
constexpr uint32_t OFF_F = OFF_A + 2 * CH_BYTES;                    // dF  hi | lo
constexpr uint32_t OFF_H = OFF_F + 2 * CH_BYTES;                    // dh  hi | lo
constexpr uint32_t OFF_POS = OFF_H + 2 * CH_BYTES;                  // float[128][3]
constexpr uint32_t OFF_ADJ = OFF_POS + DN_MAX * 12;                 // uint32[128][4]
constexpr uint32_t OFF_C = OFF_ADJ + DN_MAX * DN_AW * 4;            // float[64]
constexpr uint32_t OFF_MASK = OFF_C + TE * 4;                       // uint32[4]
constexpr uint32_t BODY = (OFF_MASK + 16 + 127) / 128 * 128;
constexpr uint32_t SMEM = W_BYTES + BODY;
static_assert(SMEM <= 232448 - 1024, "shared memory budget of the fp32-grade weight-gradient kernel");
}  // namespace bx3

__global__ void __launch_bounds__(bx3::THREADS, 1) cfconv_dense_bwd_x3_kernel(const __grid_constant__ DenseBwdParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  // wbar | r_ready, d1_ready, f_ready, dda_ready, h_ready, w_done
  __shared__ uint64_t bars[7];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_offset[K1];
  __shared__ __align__(16) float s_c2[K1];

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    tc::mbar_init(&bars[0], 1);
    tc::mbar_init(&bars[1], GT);  // r_ready   (rbf images, cutoffs and masks written)
    tc::mbar_init(&bars[2], 1);   // d1_ready  (h in TMEM)
    tc::mbar_init(&bars[3], GT);  // f_ready   (a' and dF images written)
    tc::mbar_init(&bars[4], 1);   // dda_ready (da' in TMEM)
    tc::mbar_init(&bars[5], GT);  // h_ready   (dh images written)
    tc::mbar_init(&bars[6], 1);   // w_done    (every MMA of the tile finished: all images may be overwritten)
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

  const int64_t T = __ldg(p.tile_ptr + p.G);
  const int64_t U = (int64_t)gridDim.x;
  const int k1steps = (p.Ng + 1 + 15) >> 4;
  const int64_t u = (int64_t)blockIdx.x;
  const int64_t t0 = u * T / U, t1 = (u + 1) * T / U;
  uint64_t* b = &bars[1];
  uint8_t* sG = smem + bx3::W_BYTES;

  if (warp >= GT / 32) {
    // ======================= MMA-issuing warp =======================
    if (lane == 0) {
      uint64_t* wbar = &bars[0];
      tc::mbar_arrive_expect_tx(wbar, bx3::W_BYTES);
      tc::bulk_g2s(smem, p.weights, bx3::W_BYTES / 2, wbar);
      tc::bulk_g2s(smem + bx3::W_BYTES / 2, p.weights + bx3::W_BYTES / 2, bx3::W_BYTES / 2, wbar);
      const uint32_t aW = tc::smem_u32(smem);
      const uint32_t aG = tc::smem_u32(sG);
      const uint32_t aR = aG + bx3::OFF_R, aA = aG + bx3::OFF_A, aF = aG + bx3::OFF_F, aH = aG + bx3::OFF_H;
      const uint32_t tH = tmem_base, tD = tmem_base + 64, tW2 = tmem_base + 128, tW1 = tmem_base + 256;
      tc::mbar_wait(wbar, 0);
      DenseWalk w;
      if (t0 < t1) w.seek(p, t0);
      uint32_t it = 0;
      for (int64_t ti = t0; ti < t1; ++ti, ++it) {
        bool diag;
        int c_base, j0, ncols;
        w.tile(diag, c_base, j0, ncols);
        const int npad = (ncols + 15) & ~15;
        if (ti + 1 < t1) w.next(p);
        const uint32_t par = it & 1;
        // h = W1aug * rbf^T
        tc::mbar_wait_spin(b + 0, par);
        tc::tc_fence_after();
        const uint32_t id1 = tc::umma_idesc_f16(F, npad, 1, 0, 0);
#pragma unroll 1
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t wa = aW + (pass == 1 ? bx3::OFF_W1L : 0u), ra = aR + (pass == 2 ? R_BYTES : 0u);
          for (int ks = 0; ks < k1steps; ++ks)
            tc::umma_f16(tH, tc::umma_smem_desc(wa + ks * 256, 128, 1024), tc::umma_smem_desc(ra + ks * 256, 128, 1024), id1,
                         (pass | ks) != 0);
        }
        tc::umma_commit(b + 1);
        // da' = W2^T * dF      (dF image read as MN-major [K=f, N=e]: LBO 128, SBO 2048)
        tc::mbar_wait_spin(b + 2, par);
        tc::tc_fence_after();
        const uint32_t id2 = tc::umma_idesc_f16(F, npad, 1, 0, 1);
#pragma unroll 1
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t wa = aW + (pass == 1 ? bx3::OFF_W2L : bx3::OFF_W2H), fa = aF + (pass == 2 ? CH_BYTES : 0u);
#pragma unroll
          for (int ks = 0; ks < F / 16; ++ks)
            tc::umma_f16(tD, tc::umma_smem_desc(wa + ks * 256, 128, 2048), tc::umma_smem_desc(fa + ks * 256, 128, 2048), id2,
                         (pass | ks) != 0);
        }
        tc::umma_commit(b + 3);
        // dW2 += dF a'^T, K = columns of the tile (images read as K-major [rows=channel, K=e]: SBO 128, LBO 2048)
        const uint32_t id3 = tc::umma_idesc_f16(F, F, 1, 0, 0);
#pragma unroll 1
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t fa = aF + (pass == 1 ? CH_BYTES : 0u), aa = aA + (pass == 2 ? CH_BYTES : 0u);
          for (int ks = 0; ks < (npad >> 4); ++ks)
            tc::umma_f16(tW2, tc::umma_smem_desc(fa + ks * 4096, 2048, 128), tc::umma_smem_desc(aa + ks * 4096, 2048, 128), id3,
                         (it | (uint32_t)pass | (uint32_t)ks) != 0u);
        }
        // dW1 += dh rbf^T
        tc::mbar_wait_spin(b + 4, par);
        tc::tc_fence_after();
        const uint32_t id4 = tc::umma_idesc_f16(F, K1, 1, 0, 1);
#pragma unroll 1
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t ha = aH + (pass == 1 ? CH_BYTES : 0u), ra = aR + (pass == 2 ? R_BYTES : 0u);
          for (int ks = 0; ks < (npad >> 4); ++ks)
            tc::umma_f16(tW1, tc::umma_smem_desc(ha + ks * 4096, 2048, 128), tc::umma_smem_desc(ra + ks * 2048, 1024, 128), id4,
                         (it | (uint32_t)pass | (uint32_t)ks) != 0u);
        }
        tc::umma_commit(b + 5);
      }
    }
    __syncwarp();
  } else {
    // ======================= compute warps =======================
    const int tt = tid;
    const int wq = warp & 3;                  // TMEM lane quarter
    const int h = (warp >> 2) & 1;            // column half handled in the channel-major phases
    const int chan = wq * 32 + lane;
    const int e = tt & 63;                    // column in the rbf phase
    const int q = tt >> 6;                    // 4 threads share a column in the rbf phase
    uint8_t* sR = sG + bx3::OFF_R;
    uint8_t* sA = sG + bx3::OFF_A;
    uint8_t* sF = sG + bx3::OFF_F;
    uint8_t* sH = sG + bx3::OFF_H;
    float* sPos = reinterpret_cast<float*>(sG + bx3::OFF_POS);
    uint32_t* sAdj = reinterpret_cast<uint32_t*>(sG + bx3::OFF_ADJ);
    float* sC = reinterpret_cast<float*>(sG + bx3::OFF_C);
    uint32_t* sMask = reinterpret_cast<uint32_t*>(sG + bx3::OFF_MASK);
    const uint32_t tH = tmem_base + ((uint32_t)(wq * 32) << 16);
    const uint32_t tD = tH + 64, tW2 = tH + 128, tW1 = tH + 256;
    const float pi_over_cutoff = kPi / p.cutoff;

    float db1 = 0.0f, db2 = 0.0f;
    float gr[16], xr[16];                     // g and x' rows of the current row block, this thread's channel
#pragma unroll
    for (int i = 0; i < 16; ++i) gr[i] = xr[i] = 0.0f;
    int rows_conf = -1, rows_a0 = -1, staged_conf = -1;
    uint32_t it = 0;
    DenseWalk w;
    if (t0 < t1) w.seek(p, t0);

    for (int64_t ti = t0; ti < t1; ++ti, ++it) {
      bool diag;
      int c_base, j0, ncols;
      w.tile(diag, c_base, j0, ncols);
      const int cs = w.cs, n = w.n, a0 = w.a0, conf = w.conf;
      const int m = min(16, n - a0);
      const int npad = (ncols + 15) & ~15;
      const uint32_t par = it & 1;
      const int goff = (cs + a0) * F + chan;

      if (conf != rows_conf || a0 != rows_a0) {
        rows_conf = conf;
        rows_a0 = a0;
#pragma unroll
        for (int il = 0; il < 16; ++il) {
          gr[il] = (il < m) ? __ldg(p.g + goff + il * F) : 0.0f;
          xr[il] = (il < m) ? __ldg(p.xprime + goff + il * F) : 0.0f;
        }
      }
      float gj[2] = {0.0f, 0.0f}, xj[2] = {0.0f, 0.0f};
      if (!diag) {
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int a = j0 + 2 * h + jj;
          if (a < n) {
            gj[jj] = __ldg(p.g + (cs + a) * F + chan);
            xj[jj] = __ldg(p.xprime + (cs + a) * F + chan);
          }
        }
      }

      if (conf != staged_conf) {
        staged_conf = conf;
        tc::named_bar_sync(1, GT);       // nobody still reads the previous conformer's copy
        if (tt < n) {
          const float* pp = p.pos + (int64_t)(cs + tt) * 3;
          sPos[3 * tt + 0] = __ldg(pp + 0);
          sPos[3 * tt + 1] = __ldg(pp + 1);
          sPos[3 * tt + 2] = __ldg(pp + 2);
          reinterpret_cast<uint4*>(sAdj)[tt] = __ldg(reinterpret_cast<const uint4*>(p.adj) + cs + tt);
        }
        tc::named_bar_sync(1, GT);
      }

      // every MMA of the previous tile has finished reading the images
      if (it > 0) tc::mbar_wait(b + 5, (it - 1) & 1);

      // ---- column e: pair, directions, distance, cutoff, Gaussian expansion -> rbf images ----
      {
        int il, jl, i_loc, j_loc;
        if (diag) {
          const int c = min(c_base + e, 119);
          jl = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)c)) * 0.5f);
          if (jl * (jl - 1) / 2 > c) --jl;
          if ((jl + 1) * jl / 2 <= c) ++jl;
          il = c - jl * (jl - 1) / 2;
          i_loc = a0 + il;
          j_loc = a0 + jl;
        } else {
          il = e & 15;
          i_loc = a0 + il;
          j_loc = j0 + (e >> 4);
        }
        bool ef = false, er = false;
        if (e < ncols) {
          ef = (sAdj[i_loc * DN_AW + (j_loc >> 5)] >> (j_loc & 31)) & 1u;     // edge j -> i
          er = (sAdj[j_loc * DN_AW + (i_loc >> 5)] >> (i_loc & 31)) & 1u;     // edge i -> j
        }
        const bool live = ef || er;
        float d = 0.0f;
        if (live) {
          const float dx = sPos[3 * j_loc] - sPos[3 * i_loc], dy = sPos[3 * j_loc + 1] - sPos[3 * i_loc + 1],
                      dz = sPos[3 * j_loc + 2] - sPos[3 * i_loc + 2];
          d = sqrtf(dx * dx + dy * dy + dz * dz);          // the expression of the neighbour search (graph.cu)
        }
        if (q == 0) {
          sC[e] = live ? cos_cutoff_nomask(d, pi_over_cutoff) : 0.0f;
          const unsigned bf = __ballot_sync(0xffffffffu, ef), br = __ballot_sync(0xffffffffu, er);
          if (lane == 0) {
            sMask[e >> 5] = bf;
            sMask[2 + (e >> 5)] = br;
          }
        }
        if (e < npad) {
          uint8_t* rowp = sR + (e >> 3) * 1024 + (e & 7) * 16;
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
            uint4 hi, lo;
            split_bf16x8(v, hi, lo);
            *reinterpret_cast<uint4*>(rowp + jc * 128) = hi;
            *reinterpret_cast<uint4*>(rowp + R_BYTES + jc * 128) = lo;
          }
        }
      }
      tc::fence_proxy_async();
      tc::mbar_arrive(b + 0);
      tc::named_bar_sync(1, GT);   // sC / sMask visible to the whole group

      // ---- dF images from the row registers (runs while the tensor core computes h) ----
      {
        const uint32_t mf = sMask[h], mr = sMask[2 + h];
        uint8_t* dstF = sF + chan * 16 + (h * 4) * 2048;
        const float* sC32 = sC + 32 * h;
        if (32 * h < npad) {
          if (!diag) {
            dense_df_rect<CH_BYTES>(gr, xr, gj, xj, mf, mr, sC32, dstF, db2);
          } else {
            switch ((c_base >> 5) + h) {
              case 0: dense_df_diag<0, CH_BYTES>(gr, xr, mf, mr, sC32, dstF, db2); break;
              case 1: dense_df_diag<32, CH_BYTES>(gr, xr, mf, mr, sC32, dstF, db2); break;
              case 2: dense_df_diag<64, CH_BYTES>(gr, xr, mf, mr, sC32, dstF, db2); break;
              default: dense_df_diag<96, CH_BYTES>(gr, xr, mf, mr, sC32, dstF, db2); break;
            }
          }
        }
      }

      // ---- epilogue 1: a' = C ssp(h) -> images (h stays in TMEM for epilogue 3) ----
      tc::mbar_wait(b + 1, par);
      tc::tc_fence_after();
      {
        const int cb = h * 32, ce = min(npad, h * 32 + 32);
        for (int c0 = cb; c0 < ce; c0 += 16) {
          float v[16];
          tc::tmem_ld16(tH + c0, v);
          float c[16];
          const float4* cp = reinterpret_cast<const float4*>(sC + c0);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const float4 cc = cp[k4];
            c[k4 * 4 + 0] = cc.x; c[k4 * 4 + 1] = cc.y; c[k4 * 4 + 2] = cc.z; c[k4 * 4 + 3] = cc.w;
          }
          tc::tmem_wait_ld();
          float a[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float x = v[j];
            const float t = tc::fast_ex2(-1.4426950408889634f * fabsf(x));
            a[j] = c[j] * fmaf(tc::fast_lg2(1.0f + t) - 1.0f, kLn2, fmaxf(x, 0.0f));
          }
          uint4 hi, lo;
          split_bf16x8(a, hi, lo);
          *reinterpret_cast<uint4*>(sA + chan * 16 + (c0 >> 3) * 2048) = hi;
          *reinterpret_cast<uint4*>(sA + CH_BYTES + chan * 16 + (c0 >> 3) * 2048) = lo;
          split_bf16x8(a + 8, hi, lo);
          *reinterpret_cast<uint4*>(sA + chan * 16 + ((c0 >> 3) + 1) * 2048) = hi;
          *reinterpret_cast<uint4*>(sA + CH_BYTES + chan * 16 + ((c0 >> 3) + 1) * 2048) = lo;
        }
      }
      tc::tc_fence_before();
      tc::fence_proxy_async();
      tc::mbar_arrive(b + 2);

      // ---- epilogue 3: dh = da' * C sigmoid(h) -> dh images ----
      tc::mbar_wait(b + 3, par);
      tc::tc_fence_after();
      {
        const int cb = h * 32, ce = min(npad, h * 32 + 32);
        for (int c0 = cb; c0 < ce; c0 += 16) {
          float v[16], hh[16];
          tc::tmem_ld16(tD + c0, v);
          tc::tmem_ld16(tH + c0, hh);
          float c[16];
          const float4* cp = reinterpret_cast<const float4*>(sC + c0);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const float4 cc = cp[k4];
            c[k4 * 4 + 0] = cc.x; c[k4 * 4 + 1] = cc.y; c[k4 * 4 + 2] = cc.z; c[k4 * 4 + 3] = cc.w;
          }
          tc::tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float x = hh[j];
            const float t = tc::fast_ex2(-1.4426950408889634f * fabsf(x));
            const float inv = __fdividef(1.0f, 1.0f + t);
            v[j] *= c[j] * (x >= 0.0f ? inv : t * inv);
            db1 += v[j];
          }
          uint4 hi, lo;
          split_bf16x8(v, hi, lo);
          *reinterpret_cast<uint4*>(sH + chan * 16 + (c0 >> 3) * 2048) = hi;
          *reinterpret_cast<uint4*>(sH + CH_BYTES + chan * 16 + (c0 >> 3) * 2048) = lo;
          split_bf16x8(v + 8, hi, lo);
          *reinterpret_cast<uint4*>(sH + chan * 16 + ((c0 >> 3) + 1) * 2048) = hi;
          *reinterpret_cast<uint4*>(sH + CH_BYTES + chan * 16 + ((c0 >> 3) + 1) * 2048) = lo;
        }
      }
      tc::tc_fence_before();
      tc::fence_proxy_async();
      tc::mbar_arrive(b + 4);
      if (ti + 1 < t1) w.next(p);
    }

    // ---- drain: accumulators -> this CTA's partial block ----
    float* part = p.partial + u * (int64_t)PART_FLOATS;
    const bool any = t0 < t1;
    if (any) {
      tc::mbar_wait(b + 5, (it - 1) & 1);
      tc::tc_fence_after();
    }
    for (int c0 = h * 64; c0 < h * 64 + 64; c0 += 16) {       // dW2[f = chan][k]: columns [h*64, h*64+64)
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
    for (int c0 = h * 32; c0 < h * 32 + 32; c0 += 16) {       // dW1[k = chan][j]: columns [h*32, h*32+32)
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


// =================================================================================================================
// Warp-specialised dense weight gradients: ONE stream of 64-column tiles per CTA, every phase on its own warps
// =================================================================================================================
// cfconv_dense_bwd_kernel runs the phases of a tile one after the other on the 256 threads of a group (two groups per
// CTA): ~13 K cycles per tile and group, issue slots 49 % busy, the da' MMA round trip exposed in every tile.  Here the
// tiles of a CTA flow through warp groups that only meet at mbarriers (the recipe of cfconv_dense_ws_kernel):
//   XG  (4 warps, 2 threads per column)  pair -> adjacency bits, distance, cutoff, Gaussians -> R[k & 1], cutoffs / masks
//   DF  (4 warps, thread = channel)      dF from the fp32 rows of g and x' in registers -> F[k & 1]; db2
//   MMA (1 thread)                       h[k & 1] = W1 R;  da'[k & 1] = W2^T F;  dW2 += F A^T,  dW1 += H R^T
//   EP1 (4 warps, thread = channel)      h -> a' = C ssp(h) -> A[k & 1]
//   EP3 (4 warps, thread = channel)      h, da' -> dh = da' C sigmoid(h) -> H[k & 1]; db1   (sigmoid recomputed from h: no
//                                        S image, which pays for double buffering every other image)
// Every image, h and da' are double buffered (TMEM: 2 x 64 + 2 x 64 + 128 + 64 = 448 columns), so a tile's Gaussians, dF,
// both epilogues and all three MMA groups overlap those of its neighbours.  All roles walk the same tile sequence; tile k
// of the CTA uses buffer k & 1 and completion number k >> 1 of that buffer's barriers.
// MEASURED SLOWER than cfconv_dense_bwd_kernel (cfg 2: 150 us against 89 us) and therefore NOT the default
// (cmp_debug_set_dense_bwd_variant(1) selects it; the tests run both): with two tiles in flight the chain
// XG -> DF -> da' MMA -> EP3 -> weight-gradient MMAs -> w_done -> XG(k + 2) is serial per buffer and every link runs on 4
// warps instead of 8.  A third set of images would fit in shared memory (3 x 56 KB), a third h / da' pair not in TMEM.
//   r_ready[s]  XG  -> MMA, DF         R[s], cutoffs, masks of tile k written                      (128 arrivals)
//   df_ready[s] DF  -> MMA             F[s] written                                               (128)
//   d1_ready[s] MMA -> EP1, EP3        h[s] holds tile k                                          (commit)
//   dda_ready[s] MMA -> EP3            da'[s] holds tile k                                        (commit)
//   a_ready[s]  EP1 -> MMA             A[s] written, h[s] read                                    (128)
//   h_ready[s]  EP3 -> MMA             H[s] written, h[s] and da'[s] read                         (128)
//   w_done[s]   MMA -> XG, DF, EP1, EP3  the weight-gradient MMAs of tile k have read R, F, A, H [s]   (commit)
namespace bws {
constexpr int W_XG = 0, W_DF = 4, W_EP1 = 8, W_EP3 = 12, W_MMA = 16, NWARPS = 17;
constexpr int THREADS = NWARPS * 32;
constexpr uint32_t OFF_R = 0;                              // 2 x R_BYTES
constexpr uint32_t OFF_F = OFF_R + 2 * R_BYTES;            // 2 x CH_BYTES
constexpr uint32_t OFF_A = OFF_F + 2 * CH_BYTES;           // 2 x CH_BYTES
constexpr uint32_t OFF_H = OFF_A + 2 * CH_BYTES;           // 2 x CH_BYTES
constexpr uint32_t OFF_POS = OFF_H + 2 * CH_BYTES;         // float[128][3]
constexpr uint32_t OFF_ADJ = OFF_POS + DN_MAX * 12;        // uint32[128][4]
constexpr uint32_t OFF_META = OFF_ADJ + DN_MAX * DN_AW * 4;   // 2 x { float C[64]; uint32 mask[4] }
constexpr uint32_t META_BYTES = TE * 4 + 16;
constexpr uint32_t BODY_BYTES = OFF_META + 2 * META_BYTES;
constexpr uint32_t SMEM = W1_BYTES + W2T_BYTES + (BODY_BYTES + 127) / 128 * 128;
static_assert(SMEM <= 232448 - 1024, "shared memory budget of the warp-specialised weight-gradient kernel");
// barrier slots
constexpr int B_W = 0, B_R = 1, B_DF = 3, B_D1 = 5, B_DDA = 7, B_A = 9, B_H = 11, B_WD = 13, NBARS = 15;
}  // namespace bws

// a' = c ssp(x) of two columns in packed f16x2 (see ssp_sigmoid_f16x2), returned as a bf16x2 image word
__device__ __forceinline__ uint32_t ssp_f16x2_bf16(float x0, float x1, float c0, float c1) {
  const __half2 x = __floats2half2_rn(x0, x1);
  const __half2 c = __floats2half2_rn(c0, c1);
  uint32_t tu;
  const __half2 arg = __hmul2(__habs2(x), __float2half2_rn(-1.4426950408889634f));
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(tu) : "r"(*reinterpret_cast<const uint32_t*>(&arg)));
  const __half2 t = *reinterpret_cast<const __half2*>(&tu);
  __half2 q = __hfma2(__float2half2_rn(0.0599455865f), t, __float2half2_rn(-0.227712643f));
  q = __hfma2(q, t, __float2half2_rn(0.442274178f));
  q = __hfma2(q, t, __float2half2_rn(-0.717063932f));
  q = __hfma2(q, t, __float2half2_rn(1.44261568f));
  const __half2 l2m1 = __hfma2(q, t, __float2half2_rn(-1.0f));                         // log2(1 + t) - 1
  const __half2 sp = __hfma2(l2m1, __float2half2_rn(0.6931471805599453f), __hmax2(x, __float2half2_rn(0.0f)));
  const float2 af = __half22float2(__hmul2(sp, c));
  return tc::pack_bf16x2(af.x, af.y);
}
// c sigmoid(x) of two columns (sigmoid = 1/2 + tanh(x / 2) / 2, one tanh.approx.f16x2), as two floats
__device__ __forceinline__ float2 csigmoid_f16x2(float x0, float x1, float c0, float c1) {
  const __half2 hx = __hmul2(__floats2half2_rn(x0, x1), __float2half2_rn(0.5f));
  uint32_t au;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(au) : "r"(*reinterpret_cast<const uint32_t*>(&hx)));
  const __half2 sg = __hfma2(*reinterpret_cast<const __half2*>(&au), __float2half2_rn(0.5f), __float2half2_rn(0.5f));
  const float2 sf = __half22float2(__hmul2(sg, __floats2half2_rn(c0, c1)));
  return sf;
}

__global__ void __launch_bounds__(bws::THREADS, 1) cfconv_dense_bwd_ws_kernel(const __grid_constant__ DenseBwdParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bars[bws::NBARS];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_offset[K1];
  __shared__ __align__(16) float s_c2[K1];

  uint8_t* sW1 = smem;
  uint8_t* sW2T = smem + W1_BYTES;
  uint8_t* sB = smem + W1_BYTES + W2T_BYTES;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int wq = warp & 3;

  if (tid == 0) {
    tc::mbar_init(&bars[bws::B_W], 1);
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&bars[bws::B_R + s], 128);
      tc::mbar_init(&bars[bws::B_DF + s], 128);
      tc::mbar_init(&bars[bws::B_D1 + s], 1);
      tc::mbar_init(&bars[bws::B_DDA + s], 1);
      tc::mbar_init(&bars[bws::B_A + s], 128);
      tc::mbar_init(&bars[bws::B_H + s], 128);
      tc::mbar_init(&bars[bws::B_WD + s], 1);
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
  const uint32_t tH0 = tmem_base, tDA0 = tmem_base + 128, tW2 = tmem_base + 256, tW1 = tmem_base + 384;

  const int64_t T = __ldg(p.tile_ptr + p.G);
  const int64_t U = gridDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * T / U, t1 = ((int64_t)blockIdx.x + 1) * T / U;
  const uint32_t nt = (uint32_t)(t1 - t0);           // tiles of this CTA
  const int k1steps = (p.Ng + 1 + 15) >> 4;
  // parity of completion number (k >> 1) of a buffer-k&1 barrier, and of the one two tiles earlier
  auto par_now = [](uint32_t k) { return (k >> 1) & 1u; };
  auto par_prev = [](uint32_t k) { return ((k >> 1) - 1u) & 1u; };

  if (warp < bws::W_DF) {
    // ================================ XG: Gaussians, cutoff, direction masks ================================
    const int t = tid;                       // 0..127
    const int e = t & 63, q = t >> 6;        // column, and which half of the K chunks this thread writes
    float* sPos = reinterpret_cast<float*>(sB + bws::OFF_POS);
    uint32_t* sAdj = reinterpret_cast<uint32_t*>(sB + bws::OFF_ADJ);
    const float cutoff = p.cutoff;
    int staged_conf = -1;
    DenseWalk w;
    if (nt > 0) w.seek(p, t0);
    for (uint32_t k = 0; k < nt; ++k) {
      const uint32_t s = k & 1u;
      bool diag;
      int c_base, j0, ncols;
      w.tile(diag, c_base, j0, ncols);
      const int cs = w.cs, n = w.n, a0 = w.a0, conf = w.conf;
      const int npad = (ncols + 15) & ~15;
      uint8_t* sR = sB + bws::OFF_R + s * R_BYTES;
      float* sC = reinterpret_cast<float*>(sB + bws::OFF_META + s * bws::META_BYTES);
      uint32_t* sMask = reinterpret_cast<uint32_t*>(sC + TE);
      if (conf != staged_conf) {
        staged_conf = conf;
        tc::named_bar_sync(1, 128);          // nobody still reads the previous conformer's copy
        if (t < n) {
          const float* pp = p.pos + (int64_t)(cs + t) * 3;
          sPos[3 * t + 0] = __ldg(pp + 0);
          sPos[3 * t + 1] = __ldg(pp + 1);
          sPos[3 * t + 2] = __ldg(pp + 2);
          reinterpret_cast<uint4*>(sAdj)[t] = __ldg(reinterpret_cast<const uint4*>(p.adj) + cs + t);
        }
        tc::named_bar_sync(1, 128);
      }
      int il, jl, i_loc, j_loc;
      if (diag) {
        const int c = min(c_base + e, 119);
        jl = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)c)) * 0.5f);
        if (jl * (jl - 1) / 2 > c) --jl;
        if ((jl + 1) * jl / 2 <= c) ++jl;
        il = c - jl * (jl - 1) / 2;
        i_loc = a0 + il;
        j_loc = a0 + jl;
      } else {
        il = e & 15;
        i_loc = a0 + il;
        j_loc = j0 + (e >> 4);
      }
      bool ef = false, er = false;
      if (e < ncols) {
        ef = (sAdj[i_loc * DN_AW + (j_loc >> 5)] >> (j_loc & 31)) & 1u;     // edge j -> i
        er = (sAdj[j_loc * DN_AW + (i_loc >> 5)] >> (i_loc & 31)) & 1u;     // edge i -> j
      }
      const bool live = ef || er;
      float d = 0.0f;
      if (live) {
        const float dx = sPos[3 * j_loc] - sPos[3 * i_loc], dy = sPos[3 * j_loc + 1] - sPos[3 * i_loc + 1],
                    dz = sPos[3 * j_loc + 2] - sPos[3 * i_loc + 2];
        const float d2 = dx * dx + dy * dy + dz * dz;
        d = d2 * rsqrtf(fmaxf(d2, 1e-20f));
      }
      // the weight-gradient MMAs of tile k - 2 have read R[s]; everybody has read its cutoffs and masks
      if (k >= 2) tc::mbar_wait(&bars[bws::B_WD + s], par_prev(k));
      if (q == 0) {       // warps 0 and 1 hold columns 0..31 and 32..63 (whole warps: the ballots are uniform)
        sC[e] = live ? 0.5f * (__cosf(d * kPi / cutoff) + 1.0f) : 0.0f;
        const unsigned bf = __ballot_sync(0xffffffffu, ef), br = __ballot_sync(0xffffffffu, er);
        if (lane == 0) {
          sMask[e >> 5] = bf;
          sMask[2 + (e >> 5)] = br;
        }
      }
      if (e < npad) {
        uint8_t* rowp = sR + (e >> 3) * 1024 + (e & 7) * 16;
        for (int jc = q; jc < 2 * k1steps; jc += 2) {
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
      tc::fence_proxy_async();
      tc::mbar_arrive(&bars[bws::B_R + s]);
      if (k + 1 < nt) w.next(p);
    }
  } else if (warp < bws::W_EP1) {
    // ================================ DF: dF image from the row registers ================================
    const int chan = tid - bws::W_DF * 32;
    float gr[16], xr[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) gr[i] = xr[i] = 0.0f;
    float db2 = 0.0f;
    int rows_conf = -1, rows_a0 = -1;
    DenseWalk w;
    if (nt > 0) w.seek(p, t0);
    for (uint32_t k = 0; k < nt; ++k) {
      const uint32_t s = k & 1u;
      bool diag;
      int c_base, j0, ncols;
      w.tile(diag, c_base, j0, ncols);
      const int cs = w.cs, n = w.n, a0 = w.a0, conf = w.conf;
      const int m = min(16, n - a0);
      const int npad = (ncols + 15) & ~15;
      const int goff = (cs + a0) * F + chan;
      if (conf != rows_conf || a0 != rows_a0) {
        rows_conf = conf;
        rows_a0 = a0;
#pragma unroll
        for (int il = 0; il < 16; ++il) {
          gr[il] = (il < m) ? __ldg(p.g + goff + il * F) : 0.0f;
          xr[il] = (il < m) ? __ldg(p.xprime + goff + il * F) : 0.0f;
        }
      }
      float gj[4] = {0.0f, 0.0f, 0.0f, 0.0f}, xj[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      if (!diag) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int a = j0 + jj;
          if (a < n) {
            gj[jj] = __ldg(p.g + (cs + a) * F + chan);
            xj[jj] = __ldg(p.xprime + (cs + a) * F + chan);
          }
        }
      }
      const float* sC = reinterpret_cast<const float*>(sB + bws::OFF_META + s * bws::META_BYTES);
      const uint32_t* sMask = reinterpret_cast<const uint32_t*>(sC + TE);
      uint8_t* dstF = sB + bws::OFF_F + s * CH_BYTES + chan * 16;
      tc::mbar_wait(&bars[bws::B_R + s], par_now(k));                    // cutoffs and masks of tile k
      if (k >= 2) tc::mbar_wait(&bars[bws::B_WD + s], par_prev(k));      // F[s] free
      const uint32_t mf0 = sMask[0], mf1 = sMask[1], mr0 = sMask[2], mr1 = sMask[3];
      if (!diag) {
        const float gj0[2] = {gj[0], gj[1]}, xj0[2] = {xj[0], xj[1]}, gj1[2] = {gj[2], gj[3]}, xj1[2] = {xj[2], xj[3]};
        dense_df_rect(gr, xr, gj0, xj0, mf0, mr0, sC, dstF, db2);
        if (npad > 32) dense_df_rect(gr, xr, gj1, xj1, mf1, mr1, sC + 32, dstF + 4 * 2048, db2);
      } else if (c_base == 0) {
        dense_df_diag<0>(gr, xr, mf0, mr0, sC, dstF, db2);
        if (npad > 32) dense_df_diag<32>(gr, xr, mf1, mr1, sC + 32, dstF + 4 * 2048, db2);
      } else {
        dense_df_diag<64>(gr, xr, mf0, mr0, sC, dstF, db2);
        if (npad > 32) dense_df_diag<96>(gr, xr, mf1, mr1, sC + 32, dstF + 4 * 2048, db2);
      }
      tc::fence_proxy_async();
      tc::mbar_arrive(&bars[bws::B_DF + s]);
      if (k + 1 < nt) w.next(p);
    }
    float* part = p.partial + (int64_t)blockIdx.x * PART_FLOATS;
    part[F * F + F * K1 + chan] = db2;
    part[F * F + F * K1 + F + chan] = 0.0f;
  } else if (warp < bws::W_EP3) {
    // ================================ EP1: h -> a' image; drains dW2 ================================
    const int chan = tid - bws::W_EP1 * 32;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    DenseWalk w;
    if (nt > 0) w.seek(p, t0);
    for (uint32_t k = 0; k < nt; ++k) {
      const uint32_t s = k & 1u;
      bool diag;
      int c_base, j0, ncols;
      w.tile(diag, c_base, j0, ncols);
      const int npad = (ncols + 15) & ~15;
      const float* sC = reinterpret_cast<const float*>(sB + bws::OFF_META + s * bws::META_BYTES);
      uint8_t* sA = sB + bws::OFF_A + s * CH_BYTES + chan * 16;
      const uint32_t tH = tH0 + s * 64 + lane_off;
      tc::mbar_wait(&bars[bws::B_D1 + s], par_now(k));                   // h[s] holds tile k
      if (k >= 2) tc::mbar_wait(&bars[bws::B_WD + s], par_prev(k));      // A[s] free
      tc::tc_fence_after();
      for (int c0 = 0; c0 < npad; c0 += 16) {
        float v[16];
        tc::tmem_ld16(tH + c0, v);
        float c[16];
        const float4* cp = reinterpret_cast<const float4*>(sC + c0);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const float4 cc = cp[k4];
          c[k4 * 4 + 0] = cc.x; c[k4 * 4 + 1] = cc.y; c[k4 * 4 + 2] = cc.z; c[k4 * 4 + 3] = cc.w;
        }
        tc::tmem_wait_ld();
        uint32_t a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = ssp_f16x2_bf16(v[2 * j], v[2 * j + 1], c[2 * j], c[2 * j + 1]);
        *reinterpret_cast<uint4*>(sA + (c0 >> 3) * 2048) = make_uint4(a[0], a[1], a[2], a[3]);
        *reinterpret_cast<uint4*>(sA + ((c0 >> 3) + 1) * 2048) = make_uint4(a[4], a[5], a[6], a[7]);
      }
      tc::tc_fence_before();
      tc::fence_proxy_async();
      tc::mbar_arrive(&bars[bws::B_A + s]);
      if (k + 1 < nt) w.next(p);
    }
    // drain dW2[f = chan][k = 0..127]
    float* part = p.partial + (int64_t)blockIdx.x * PART_FLOATS;
    if (nt > 0) {
      tc::mbar_wait(&bars[bws::B_WD + ((nt - 1) & 1u)], par_now(nt - 1));
      tc::tc_fence_after();
    }
    for (int c0 = 0; c0 < F; c0 += 16) {
      float v[16];
      if (nt > 0) {
        tc::tmem_ld16(tW2 + lane_off + c0, v);
        tc::tmem_wait_ld();
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.0f;
      }
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(part + chan * F + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
  } else if (warp < bws::W_MMA) {
    // ================================ EP3: h, da' -> dh image; db1; drains dW1 ================================
    const int chan = tid - bws::W_EP3 * 32;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    float db1 = 0.0f;
    DenseWalk w;
    if (nt > 0) w.seek(p, t0);
    for (uint32_t k = 0; k < nt; ++k) {
      const uint32_t s = k & 1u;
      bool diag;
      int c_base, j0, ncols;
      w.tile(diag, c_base, j0, ncols);
      const int npad = (ncols + 15) & ~15;
      const float* sC = reinterpret_cast<const float*>(sB + bws::OFF_META + s * bws::META_BYTES);
      uint8_t* sH = sB + bws::OFF_H + s * CH_BYTES + chan * 16;
      const uint32_t tH = tH0 + s * 64 + lane_off, tDA = tDA0 + s * 64 + lane_off;
      tc::mbar_wait(&bars[bws::B_D1 + s], par_now(k));                   // h[s]
      tc::mbar_wait(&bars[bws::B_DDA + s], par_now(k));                  // da'[s]
      if (k >= 2) tc::mbar_wait(&bars[bws::B_WD + s], par_prev(k));      // H[s] free
      tc::tc_fence_after();
      for (int c0 = 0; c0 < npad; c0 += 16) {
        float hv[16], dv[16];
        tc::tmem_ld16(tH + c0, hv);
        tc::tmem_ld16(tDA + c0, dv);
        float c[16];
        const float4* cp = reinterpret_cast<const float4*>(sC + c0);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const float4 cc = cp[k4];
          c[k4 * 4 + 0] = cc.x; c[k4 * 4 + 1] = cc.y; c[k4 * 4 + 2] = cc.z; c[k4 * 4 + 3] = cc.w;
        }
        tc::tmem_wait_ld();
        float o[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 sg = csigmoid_f16x2(hv[2 * j], hv[2 * j + 1], c[2 * j], c[2 * j + 1]);
          o[2 * j] = dv[2 * j] * sg.x;
          o[2 * j + 1] = dv[2 * j + 1] * sg.y;
          db1 += o[2 * j] + o[2 * j + 1];
        }
        *reinterpret_cast<uint4*>(sH + (c0 >> 3) * 2048) = pack_bf16x8(o);
        *reinterpret_cast<uint4*>(sH + ((c0 >> 3) + 1) * 2048) = pack_bf16x8(o + 8);
      }
      tc::tc_fence_before();
      tc::fence_proxy_async();
      tc::mbar_arrive(&bars[bws::B_H + s]);
      if (k + 1 < nt) w.next(p);
    }
    // drain dW1[k = chan][j = 0..63]
    float* part = p.partial + (int64_t)blockIdx.x * PART_FLOATS;
    if (nt > 0) {
      tc::mbar_wait(&bars[bws::B_WD + ((nt - 1) & 1u)], par_now(nt - 1));
      tc::tc_fence_after();
    }
    for (int c0 = 0; c0 < K1; c0 += 16) {
      float v[16];
      if (nt > 0) {
        tc::tmem_ld16(tW1 + lane_off + c0, v);
        tc::tmem_wait_ld();
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.0f;
      }
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(part + F * F + chan * K1 + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    part[F * F + F * K1 + 2 * F + chan] = db1;
    part[F * F + F * K1 + 3 * F + chan] = 0.0f;
  } else {
    // ================================ MMA: one thread issues everything ================================
    if (lane == 0) {
      tc::mbar_arrive_expect_tx(&bars[bws::B_W], W1_BYTES + W2T_BYTES);
      tc::bulk_g2s(sW1, p.weights, W1_BYTES, &bars[bws::B_W]);
      tc::bulk_g2s(sW2T, p.weights + W1_BYTES, W2T_BYTES, &bars[bws::B_W]);
      const uint32_t aW1 = tc::smem_u32(sW1), aW2T = tc::smem_u32(sW2T), aB = tc::smem_u32(sB);
      tc::mbar_wait(&bars[bws::B_W], 0);
      DenseWalk w;
      if (nt > 0) w.seek(p, t0);
      int np_prev = 0;                       // width of tile k - 1
      const uint32_t id3 = tc::umma_idesc_f16(F, F, 1, 0, 0);
      const uint32_t id4 = tc::umma_idesc_f16(F, K1, 1, 0, 1);
      auto weight_grads = [&](uint32_t j, int npad) {      // dW2 += F A^T, dW1 += H R^T of tile j
        const uint32_t s = j & 1u;
        tc::mbar_wait_spin(&bars[bws::B_A + s], par_now(j));
        tc::mbar_wait_spin(&bars[bws::B_H + s], par_now(j));
        tc::tc_fence_after();
        const uint32_t aR = aB + bws::OFF_R + s * R_BYTES, aF = aB + bws::OFF_F + s * CH_BYTES,
                       aA = aB + bws::OFF_A + s * CH_BYTES, aH = aB + bws::OFF_H + s * CH_BYTES;
        for (int ks = 0; ks < (npad >> 4); ++ks) {
          const uint32_t acc = (j > 0 || ks > 0) ? 1u : 0u;
          tc::umma_f16(tW2, tc::umma_smem_desc(aF + ks * 4096, 2048, 128), tc::umma_smem_desc(aA + ks * 4096, 2048, 128), id3,
                       acc);
          tc::umma_f16(tW1, tc::umma_smem_desc(aH + ks * 4096, 2048, 128), tc::umma_smem_desc(aR + ks * 2048, 1024, 128), id4,
                       acc);
        }
        tc::umma_commit(&bars[bws::B_WD + s]);
      };
      for (uint32_t k = 0; k < nt; ++k) {
        const uint32_t s = k & 1u;
        bool diag;
        int c_base, j0, ncols;
        w.tile(diag, c_base, j0, ncols);
        const int npad = (ncols + 15) & ~15;
        if (k + 1 < nt) w.next(p);
        const uint32_t aR = aB + bws::OFF_R + s * R_BYTES, aF = aB + bws::OFF_F + s * CH_BYTES;
        // h[s] = W1aug * rbf^T   (h[s] and da'[s] are free: the weight-gradient MMAs of tile k - 2, issued in the previous
        // iteration, waited for both epilogues of that tile)
        tc::mbar_wait_spin(&bars[bws::B_R + s], par_now(k));
        tc::tc_fence_after();
        const uint32_t id1 = tc::umma_idesc_f16(F, npad, 1, 0, 0);
        for (int ks = 0; ks < k1steps; ++ks)
          tc::umma_f16(tH0 + s * 64, tc::umma_smem_desc(aW1 + ks * 256, 128, 1024), tc::umma_smem_desc(aR + ks * 256, 128, 1024),
                       id1, ks > 0);
        tc::umma_commit(&bars[bws::B_D1 + s]);
        // da'[s] = W2^T * dF
        tc::mbar_wait_spin(&bars[bws::B_DF + s], par_now(k));
        tc::tc_fence_after();
        const uint32_t id2 = tc::umma_idesc_f16(F, npad, 1, 0, 1);
#pragma unroll
        for (int ks = 0; ks < F / 16; ++ks)
          tc::umma_f16(tDA0 + s * 64, tc::umma_smem_desc(aW2T + ks * 256, 128, 2048),
                       tc::umma_smem_desc(aF + ks * 256, 128, 2048), id2, ks > 0);
        tc::umma_commit(&bars[bws::B_DDA + s]);
        // weight gradients of the previous tile
        if (k >= 1) weight_grads(k - 1, np_prev);
        np_prev = npad;
      }
      if (nt >= 1) weight_grads(nt - 1, np_prev);
    }
    __syncwarp();
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
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

static int g_dense_bwd_variant = 0;   // 0: two groups of 256 threads (default)   1: warp-specialised tile pipeline
extern "C" void cmp_debug_set_dense_bwd_variant(int v) { g_dense_bwd_variant = v; }

extern "C" size_t cmp_cfconv_dense_bwd_workspace(void) { return cmp_cfconv_fused_bwd_workspace(); }

extern "C" int cmp_build_dense_bwd_tiles(const int32_t* seg_ptr, int64_t G, int32_t* tile_ptr, int32_t* status,
                                         cmp_stream_t stream) {
  CMP_REQUIRE(G >= 0 && G < ((int64_t)1 << 31), CMP_EINVAL, "cmp_build_dense_bwd_tiles: bad number of conformers");
  CMP_REQUIRE(seg_ptr && tile_ptr, CMP_EINVAL, "cmp_build_dense_bwd_tiles: null pointer");
  dense_tile_ptr_kernel<<<1, 256, 0, as_stream(stream)>>>(seg_ptr, G, tile_ptr, status);
  CMP_LAUNCH_CHECK("cmp_build_dense_bwd_tiles");
  return CMP_OK;
}

extern "C" int cmp_cfconv_dense_bwd_weights(const float* g, const float* xprime, const float* pos, const int32_t* seg_ptr,
                                            const uint32_t* adj, const int32_t* tile_ptr, int64_t G,
                                            const void* packed_bwd_weights,
                                            const float* offset, int num_gaussians, float coeff, float cutoff,
                                            int num_filters, float* dW1, float* db1, float* dW2, float* db2,
                                            void* workspace, size_t workspace_bytes, cmp_stream_t stream) {
  CMP_REQUIRE(num_filters == F && num_gaussians >= 1 && num_gaussians < K1, CMP_EUNSUPPORTED,
              "cmp_cfconv_dense_bwd_weights: needs num_filters == 128 and num_gaussians < 64");
  CMP_REQUIRE(G >= 1 && G < ((int64_t)1 << 31), CMP_EINVAL, "cmp_cfconv_dense_bwd_weights: bad number of conformers");
  CMP_REQUIRE(g && xprime && pos && seg_ptr && adj && tile_ptr && packed_bwd_weights && offset && dW1 && db1 && dW2 && db2,
              CMP_EINVAL, "cmp_cfconv_dense_bwd_weights: null pointer");
  CMP_REQUIRE(((uintptr_t)packed_bwd_weights % 16 == 0) && ((uintptr_t)adj % 16 == 0), CMP_EINVAL,
              "cmp_cfconv_dense_bwd_weights: packed_bwd_weights / adj must be 16-byte aligned");
  CMP_REQUIRE(workspace && workspace_bytes >= cmp_cfconv_dense_bwd_workspace(), CMP_EWORKSPACE,
              "cmp_cfconv_dense_bwd_weights: workspace too small");
  CMP_REQUIRE(cmp_device_is_sm100(), CMP_EUNSUPPORTED, "cmp_cfconv_dense_bwd_weights: needs an sm_100 device (tcgen05)");
  cudaStream_t st = as_stream(stream);
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(cfconv_dense_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D_SMEM_BYTES) !=
            cudaSuccess ||
        cudaFuncSetAttribute(cfconv_dense_bwd_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bws::SMEM) !=
            cudaSuccess) {
      (void)cudaGetLastError();
      set_error("cmp_cfconv_dense_bwd_weights: cannot opt in to %u bytes of shared memory", D_SMEM_BYTES);
      return CMP_ECUDA;
    }
    attr_set = true;
  }
  DenseBwdParams p;
  p.g = g;
  p.xprime = xprime;
  p.pos = pos;
  p.seg_ptr = seg_ptr;
  p.adj = adj;
  p.tile_ptr = tile_ptr;
  p.weights = reinterpret_cast<const uint8_t*>(packed_bwd_weights);
  p.offset = offset;
  p.partial = reinterpret_cast<float*>(workspace);
  p.coeff_log2e = coeff * 1.4426950408889634f;
  p.cutoff = cutoff;
  p.Ng = num_gaussians;
  p.G = (int)G;
  const int grid = sm_count();
  const bool ws = g_dense_bwd_variant == 1;
  if (ws)
    cfconv_dense_bwd_ws_kernel<<<grid, bws::THREADS, bws::SMEM, st>>>(p);
  else
    cfconv_dense_bwd_kernel<<<grid, CTA_THREADS, D_SMEM_BYTES, st>>>(p);
  CMP_LAUNCH_CHECK("cmp_cfconv_dense_bwd_weights");
  const int total = F * F + F * K1 + 2 * F;
  // one partial block per CTA in both variants (the default kernel adds its two groups in the drain)
  reduce_partials_kernel<<<(total + 31) / 32, dim3(32, 8), 0, st>>>(p.partial, grid, num_gaussians, dW1, db1, dW2, db2);
  CMP_LAUNCH_CHECK("cmp_cfconv_dense_bwd_weights");
  return CMP_OK;
}

extern "C" size_t cmp_cfconv_dense_bwd_x3_weights_bytes(void) { return 2 * (W1_BYTES + W2T_BYTES); }

extern "C" int cmp_cfconv_dense_bwd_x3_pack_weights_grouped(const void* jobs, int count, int num_filters,
                                                            int num_gaussians, cmp_stream_t stream) {
  CMP_REQUIRE(num_filters == F && num_gaussians >= 1 && num_gaussians < K1, CMP_EUNSUPPORTED,
              "cmp_cfconv_dense_bwd_x3_pack_weights_grouped: needs num_filters == 128 and num_gaussians < 64");
  CMP_REQUIRE(count >= 0 && count <= 32, CMP_EINVAL,
              "cmp_cfconv_dense_bwd_x3_pack_weights_grouped: count must be in [0, 32]");
  if (count == 0) return CMP_OK;
  CMP_REQUIRE(jobs, CMP_EINVAL, "cmp_cfconv_dense_bwd_x3_pack_weights_grouped: null pointer");
  const PackBwdX3Job* in = reinterpret_cast<const PackBwdX3Job*>(jobs);
  PackBwdX3Group g;
  for (int i = 0; i < count; ++i) {
    CMP_REQUIRE(in[i].W1 && in[i].b1 && in[i].W2 && in[i].packed, CMP_EINVAL,
                "cmp_cfconv_dense_bwd_x3_pack_weights_grouped: null pointer");
    g.j[i] = in[i];
  }
  const int total = F * K1 + F * F;
  pack_bwd_weights_x3_kernel<<<dim3((total + 255) / 256, count), 256, 0, as_stream(stream)>>>(g, num_gaussians);
  CMP_LAUNCH_CHECK("cmp_cfconv_dense_bwd_x3_pack_weights_grouped");
  return CMP_OK;
}

extern "C" int cmp_cfconv_dense_bwd_x3_weights(const float* g, const float* xprime, const float* pos,
                                               const int32_t* seg_ptr, const uint32_t* adj, const int32_t* tile_ptr,
                                               int64_t G, const void* packed_bwd_x3_weights, const float* offset,
                                               int num_gaussians, float coeff, float cutoff, int num_filters,
                                               float* dW1, float* db1, float* dW2, float* db2, void* workspace,
                                               size_t workspace_bytes, cmp_stream_t stream) {
  CMP_REQUIRE(num_filters == F && num_gaussians >= 1 && num_gaussians < K1, CMP_EUNSUPPORTED,
              "cmp_cfconv_dense_bwd_x3_weights: needs num_filters == 128 and num_gaussians < 64");
  CMP_REQUIRE(G >= 1 && G < ((int64_t)1 << 31), CMP_EINVAL, "cmp_cfconv_dense_bwd_x3_weights: bad number of conformers");
  CMP_REQUIRE(g && xprime && pos && seg_ptr && adj && tile_ptr && packed_bwd_x3_weights && offset && dW1 && db1 && dW2 &&
                  db2,
              CMP_EINVAL, "cmp_cfconv_dense_bwd_x3_weights: null pointer");
  CMP_REQUIRE(((uintptr_t)packed_bwd_x3_weights % 16 == 0) && ((uintptr_t)adj % 16 == 0), CMP_EINVAL,
              "cmp_cfconv_dense_bwd_x3_weights: packed weights / adj must be 16-byte aligned");
  CMP_REQUIRE(workspace && workspace_bytes >= cmp_cfconv_dense_bwd_workspace(), CMP_EWORKSPACE,
              "cmp_cfconv_dense_bwd_x3_weights: workspace too small");
  CMP_REQUIRE(cmp_device_is_sm100(), CMP_EUNSUPPORTED,
              "cmp_cfconv_dense_bwd_x3_weights: needs an sm_100 device (tcgen05)");
  cudaStream_t st = as_stream(stream);
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(cfconv_dense_bwd_x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bx3::SMEM) !=
        cudaSuccess) {
      (void)cudaGetLastError();
      set_error("cmp_cfconv_dense_bwd_x3_weights: cannot opt in to %u bytes of shared memory", bx3::SMEM);
      return CMP_ECUDA;
    }
    attr_set = true;
  }
  DenseBwdParams p;
  p.g = g;
  p.xprime = xprime;
  p.pos = pos;
  p.seg_ptr = seg_ptr;
  p.adj = adj;
  p.tile_ptr = tile_ptr;
  p.weights = reinterpret_cast<const uint8_t*>(packed_bwd_x3_weights);
  p.offset = offset;
  p.partial = reinterpret_cast<float*>(workspace);
  p.coeff_log2e = coeff * 1.4426950408889634f;
  p.cutoff = cutoff;
  p.Ng = num_gaussians;
  p.G = (int)G;
  const int grid = sm_count();
  cfconv_dense_bwd_x3_kernel<<<grid, bx3::THREADS, bx3::SMEM, st>>>(p);
  CMP_LAUNCH_CHECK("cmp_cfconv_dense_bwd_x3_weights");
  const int total = F * F + F * K1 + 2 * F;
  reduce_partials_kernel<<<(total + 31) / 32, dim3(32, 8), 0, st>>>(p.partial, grid, num_gaussians, dW1, db1, dW2, db2);
  CMP_LAUNCH_CHECK("cmp_cfconv_dense_bwd_x3_weights");
  return CMP_OK;
}
