// ViSNet edge-level kernels (exact fp32): the gather / message / aggregate / edge-update pieces of the vendored
// ViSNet (conan_fgw/src/model/graph_embeddings/torch_geometric_visnet.py), each with a hand-written backward.
// Every reduction runs over the destination-sorted CSR (or its source-sorted transpose) in a fixed order:
// deterministic, no atomics (the reference scatter-adds with atomics, tgv.py:671-672).
//
// Conventions: edge e = (src j = col[e]) -> (dst i = row of e);  `_i` tensors are gathered at dst, `_j` at src
// (SURVEY.md A.4).  One warp per CSR row; lanes stride the channel axis.
#include "common.cuh"

namespace cmp {
namespace {

__device__ __forceinline__ float cos_cutoff_masked(float d, float cutoff) {
  // tgv.py:44-46
  return d < cutoff ? 0.5f * (cosf(d * kPi / cutoff) + 1.0f) : 0.0f;
}

// ---- edge geometry: ExpNormalSmearing (tgv.py:100-111), unit vectors (tgv.py:864-866), cosine cutoff ----
__global__ void vis_edge_geometry_kernel(const float* __restrict__ evec, const float* __restrict__ dist,
                                         const int32_t* __restrict__ col, const int32_t* __restrict__ erow, int64_t E,
                                         float cutoff, float alpha, const float* __restrict__ means,
                                         const float* __restrict__ betas, int R, float* __restrict__ rbf,
                                         float* __restrict__ dhat, float* __restrict__ C) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = E * R;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; idx < total; idx += stride) {
    const int64_t e = idx / R;
    const int k = (int)(idx - e * R);
    const float d = dist[e];
    const float c = cos_cutoff_masked(d, cutoff);
    const float t = expf(alpha * (-d)) - means[k];
    rbf[idx] = c * expf(-betas[k] * (t * t));
    if (k == 0) {
      C[e] = c;
      const bool loop = col[e] == erow[e];
      const float x = evec[3 * e], y = evec[3 * e + 1], z = evec[3 * e + 2];
      if (loop) {
        dhat[3 * e] = x; dhat[3 * e + 1] = y; dhat[3 * e + 2] = z;     // zero vector, left untouched by the reference
      } else {
        const float n = sqrtf(x * x + y * y + z * z);
        dhat[3 * e] = x / n; dhat[3 * e + 1] = y / n; dhat[3 * e + 2] = z / n;
      }
    }
  }
}

// ---- LayerNorm over the last axis (torch.nn.LayerNorm, eps inside the sqrt) ----
__global__ void layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                     int64_t M, int H, float eps, float* __restrict__ y, float* __restrict__ mean,
                                     float* __restrict__ rstd) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* xr = x + row * H;
  float s = 0.0f;
  for (int c = lane; c < H; c += 32) s += xr[c];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mu = s / H;
  float v = 0.0f;
  for (int c = lane; c < H; c += 32) {
    const float t = xr[c] - mu;
    v += t * t;
  }
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const float rs = rsqrtf(v / H + eps);
  for (int c = lane; c < H; c += 32) y[row * H + c] = (xr[c] - mu) * rs * w[c] + b[c];
  if (lane == 0) {
    mean[row] = mu;
    rstd[row] = rs;
  }
}

// dx and the per-row product dy * xhat (column sums of it / of dy give dw / db)
__global__ void layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ w,
                                     const float* __restrict__ mean, const float* __restrict__ rstd, int64_t M, int H,
                                     float* __restrict__ dx, float* __restrict__ dyxhat) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float mu = mean[row], rs = rstd[row];
  float s1 = 0.0f, s2 = 0.0f;
  for (int c = lane; c < H; c += 32) {
    const float xh = (x[row * H + c] - mu) * rs;
    const float g = dy[row * H + c] * w[c];
    s1 += g;
    s2 += g * xh;
  }
  for (int o = 16; o; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  s1 /= H;
  s2 /= H;
  for (int c = lane; c < H; c += 32) {
    const float xh = (x[row * H + c] - mu) * rs;
    const float g = dy[row * H + c] * w[c];
    dx[row * H + c] = (g - s1 - xh * s2) * rs;
    dyxhat[row * H + c] = dy[row * H + c] * xh;
  }
}

// ---- generic CSR reductions / expansions over [E, C] <-> [N, C] ----
// out[i] = sum_{e in row i} x[perm ? perm[e] : e]      (perm = eid_t for the source-sorted transpose)
__global__ void csr_segment_sum_kernel(const float* __restrict__ x, const int32_t* __restrict__ rowptr,
                                       const int32_t* __restrict__ perm, int64_t N, int C, float* __restrict__ out) {
  // one warp per (row, 32-channel slice): C / 32 times the warps of a warp-per-row mapping (the rows of a small batch
  // do not fill the GPU, and the edge loop of a row is a serial chain)
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int slices = (C + 31) >> 5;
  const int64_t row = wid / slices;
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const int b = rowptr[row], e = rowptr[row + 1];
  for (int c = (int)(wid - row * slices) * 32 + lane; c < C; c += C) {     // a single iteration
    float acc = 0.0f;
    for (int k = b; k < e; ++k) acc += x[(int64_t)(perm ? perm[k] : k) * C + c];
    out[row * C + c] = acc;
  }
}

// out[e] = x[idx[e]]   (idx = col for `_j`, erow for `_i`)
__global__ void gather_rows_kernel(const float* __restrict__ x, const int32_t* __restrict__ idx, int64_t E, int C,
                                   float* __restrict__ out) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = E * C;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; t < total; t += stride) {
    const int64_t e = t / C;
    const int c = (int)(t - e * C);
    out[t] = x[(int64_t)idx[e] * C + c];
  }
}

// ---- EdgeEmbedding (tgv.py:463-465): f[e] = (x_i + x_j) * ep[e];  t[e] = g[e] * ep[e] feeds both node reductions ----
__global__ void vis_edge_embed_fwd_kernel(const float* __restrict__ x, const float* __restrict__ ep,
                                          const int32_t* __restrict__ col, const int32_t* __restrict__ erow, int64_t E,
                                          int H, float* __restrict__ f) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = E * H;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; t < total; t += stride) {
    const int64_t e = t / H;
    const int c = (int)(t - e * H);
    f[t] = (x[(int64_t)erow[e] * H + c] + x[(int64_t)col[e] * H + c]) * ep[t];
  }
}

__global__ void vis_edge_embed_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                          const float* __restrict__ ep, const int32_t* __restrict__ col,
                                          const int32_t* __restrict__ erow, int64_t E, int H, float* __restrict__ gep,
                                          float* __restrict__ t_out) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = E * H;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; t < total; t += stride) {
    const int64_t e = t / H;
    const int c = (int)(t - e * H);
    gep[t] = g[t] * (x[(int64_t)erow[e] * H + c] + x[(int64_t)col[e] * H + c]);
    t_out[t] = g[t] * ep[t];
  }
}

// ---- ViS_MP.message, scalar part (tgv.py:644-648): attn = silu(sum_d q_i k_j dk) * C;  m = v_j * dv * attn ----
// one warp per edge, channel c owned by lane c % 32 ... heads are contiguous channel blocks of size hd
__global__ void __launch_bounds__(256)
vis_message_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                       const float* __restrict__ dk, const float* __restrict__ dv, const float* __restrict__ C,
                       const int32_t* __restrict__ col, const int32_t* __restrict__ erow, int64_t E, int H, int heads,
                       int pre_act, float* __restrict__ m, float* __restrict__ attn_pre) {
  // pre_act: dk / dv hold the PRE-activations of tgv.py:622-629 (the outputs of dk_proj / dv_proj): SiLU is applied here,
  // so the activated [E, H] tensors never exist in HBM
  extern __shared__ float sh[];   // [warps][heads]
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t e = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  float* hs = sh + wib * heads;
  if (e >= E) return;
  const int hd = H / heads;
  const int64_t i = erow[e], j = col[e];
  if (hd == 16 && (H & 31) == 0 && H <= 256) {
    // ConAN's shape (H = 128, 8 heads): lane l owns channels r * 32 + l, i.e. head 2 r + (l >> 4); every head sum is a
    // butterfly inside a 16-lane half, all H / 32 of them in flight at once (the generic loop below walks the heads one
    // after the other with half the lanes idle).  Same butterfly order per head: bit-identical results.
    const int R = H >> 5;
    const float ce = C[e];
    float p[8], vv[8], dvc[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (r < R) {
        const int c = r * 32 + lane;
        const float dkp = dk[e * H + c], dvp = dv[e * H + c];
        p[r] = q[i * H + c] * k[j * H + c] * (pre_act ? silu(dkp) : dkp);
        vv[r] = v[j * H + c];
        dvc[r] = pre_act ? silu(dvp) : dvp;
      }
    }
#pragma unroll
    for (int o = 8; o; o >>= 1)
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if (r < R) p[r] += __shfl_xor_sync(0xffffffffu, p[r], o);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (r < R) {
        if ((lane & 15) == 0) attn_pre[e * heads + 2 * r + (lane >> 4)] = p[r];
        m[e * H + r * 32 + lane] = vv[r] * dvc[r] * (silu(p[r]) * ce);
      }
    }
    return;
  }
  for (int h = lane; h < heads; h += 32) hs[h] = 0.0f;
  __syncwarp();
  // head sums: each lane accumulates its channels, then adds into the head slot (fixed lane order via serialised loop)
  for (int h = 0; h < heads; ++h) {
    float s = 0.0f;
    for (int c = h * hd + lane; c < (h + 1) * hd; c += 32) {
      const float dkc = pre_act ? silu(dk[e * H + c]) : dk[e * H + c];
      s += q[i * H + c] * k[j * H + c] * dkc;
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) hs[h] = s;
  }
  __syncwarp();
  const float ce = C[e];
  for (int h = lane; h < heads; h += 32) attn_pre[e * heads + h] = hs[h];
  for (int c = lane; c < H; c += 32) {
    const float s = hs[c / hd];
    const float dvc = pre_act ? silu(dv[e * H + c]) : dv[e * H + c];
    m[e * H + c] = v[j * H + c] * dvc * (silu(s) * ce);
  }
}

__global__ void __launch_bounds__(256)
vis_message_bwd_kernel(const float* __restrict__ gm, const float* __restrict__ q, const float* __restrict__ k,
                       const float* __restrict__ v, const float* __restrict__ dk, const float* __restrict__ dv,
                       const float* __restrict__ C, const float* __restrict__ attn_pre, const int32_t* __restrict__ col,
                       const int32_t* __restrict__ erow, int64_t E, int H, int heads, int pre_act,
                       float* __restrict__ g_dk, float* __restrict__ g_dv, float* __restrict__ geq,
                       float* __restrict__ gek, float* __restrict__ gev) {
  extern __shared__ float sh[];   // [warps][heads]  d(loss)/d(pre-activation)
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t e = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  float* gs = sh + wib * heads;
  if (e >= E) return;
  const int hd = H / heads;
  const int64_t i = erow[e], j = col[e];
  const float ce = C[e];
  if (hd == 16 && (H & 31) == 0 && H <= 256) {     // as in the forward kernel: all head sums in flight at once
    const int R = H >> 5;
    float ga[8], g[8], qi[8], kj[8], vj[8], dkp[8], dvp[8], dkc[8], dvc[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (r < R) {
        const int c = r * 32 + lane;
        g[r] = gm[e * H + c];
        qi[r] = q[i * H + c]; kj[r] = k[j * H + c]; vj[r] = v[j * H + c];
        dkp[r] = dk[e * H + c]; dvp[r] = dv[e * H + c];
        dkc[r] = pre_act ? silu(dkp[r]) : dkp[r];
        dvc[r] = pre_act ? silu(dvp[r]) : dvp[r];
        ga[r] = g[r] * vj[r] * dvc[r];
      }
    }
#pragma unroll
    for (int o = 8; o; o >>= 1)
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if (r < R) ga[r] += __shfl_xor_sync(0xffffffffu, ga[r], o);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (r < R) {
        const int c = r * 32 + lane;
        const float ap = attn_pre[e * heads + 2 * r + (lane >> 4)];
        const float sg = ga[r] * ce * silu_grad(ap);
        const float a = silu(ap) * ce;
        gev[e * H + c] = g[r] * dvc[r] * a;
        g_dv[e * H + c] = g[r] * vj[r] * a * (pre_act ? silu_grad(dvp[r]) : 1.0f);
        geq[e * H + c] = sg * kj[r] * dkc[r];
        gek[e * H + c] = sg * qi[r] * dkc[r];
        g_dk[e * H + c] = sg * qi[r] * kj[r] * (pre_act ? silu_grad(dkp[r]) : 1.0f);
      }
    }
    return;
  }
  for (int h = 0; h < heads; ++h) {
    float ga = 0.0f;   // d/d attn[h]
    for (int c = h * hd + lane; c < (h + 1) * hd; c += 32)
      ga += gm[e * H + c] * v[j * H + c] * (pre_act ? silu(dv[e * H + c]) : dv[e * H + c]);
    for (int o = 16; o; o >>= 1) ga += __shfl_xor_sync(0xffffffffu, ga, o);
    if (lane == 0) gs[h] = ga * ce * silu_grad(attn_pre[e * heads + h]);
  }
  __syncwarp();
  for (int c = lane; c < H; c += 32) {
    const int h = c / hd;
    const float a = silu(attn_pre[e * heads + h]) * ce;
    const float g = gm[e * H + c];
    const float qi = q[i * H + c], kj = k[j * H + c], vj = v[j * H + c];
    const float dkp = dk[e * H + c], dvp = dv[e * H + c];
    const float dkc = pre_act ? silu(dkp) : dkp, dvc = pre_act ? silu(dvp) : dvp;
    gev[e * H + c] = g * dvc * a;
    g_dv[e * H + c] = g * vj * a * (pre_act ? silu_grad(dvp) : 1.0f);      // gradient of the pre-activation when pre_act
    const float s = gs[h];
    geq[e * H + c] = s * kj * dkc;
    gek[e * H + c] = s * qi * dkc;
    g_dk[e * H + c] = s * qi * kj * (pre_act ? silu_grad(dkp) : 1.0f);
  }
}

// ---- ViS_MP vector message + aggregation (tgv.py:650-651, 672): vagg[i] = sum_e vec[j] * s1[e] + s2[e] * dhat[e] ----
__global__ void __launch_bounds__(256)
vis_vecagg_fwd_kernel(const float* __restrict__ vec, const float* __restrict__ s12, const float* __restrict__ dhat,
                      const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t N, int H, int pre_act,
                      float* __restrict__ vagg) {
  // one warp per (row, 32-channel slice): H / 32 times the warps of a warp-per-row mapping (the rows of a small batch
  // do not fill the GPU, and the edge loop of a row is a serial chain)
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int slices = (H + 31) >> 5;
  const int64_t row = wid / slices;
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const int b = rowptr[row], e = rowptr[row + 1];
  for (int c = (int)(wid - row * slices) * 32 + lane; c < H; c += H) {     // a single iteration
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
    for (int kk = b; kk < e; ++kk) {
      const int64_t j = col[kk];
      float s1 = s12[(int64_t)kk * 2 * H + c], s2 = s12[(int64_t)kk * 2 * H + H + c];
      if (pre_act) { s1 = silu(s1); s2 = silu(s2); }     // s12 = the output of s_proj (tgv.py:649): SiLU applied here
      a0 += vec[(j * 3 + 0) * H + c] * s1 + s2 * dhat[3 * kk + 0];
      a1 += vec[(j * 3 + 1) * H + c] * s1 + s2 * dhat[3 * kk + 1];
      a2 += vec[(j * 3 + 2) * H + c] * s1 + s2 * dhat[3 * kk + 2];
    }
    vagg[(row * 3 + 0) * H + c] = a0;
    vagg[(row * 3 + 1) * H + c] = a1;
    vagg[(row * 3 + 2) * H + c] = a2;
  }
}

// per-edge gradients of s1 | s2 (warp per target row)
__global__ void __launch_bounds__(256)
vis_vecagg_bwd_s_kernel(const float* __restrict__ g, const float* __restrict__ vec, const float* __restrict__ dhat,
                        const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t N, int H,
                        const float* __restrict__ s12_pre, float* __restrict__ g_s12) {
  // one warp per (row, 32-channel slice): H / 32 times the warps of a warp-per-row mapping (the rows of a small batch
  // do not fill the GPU, and the edge loop of a row is a serial chain)
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int slices = (H + 31) >> 5;
  const int64_t row = wid / slices;
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const int b = rowptr[row], e = rowptr[row + 1];
  for (int c = (int)(wid - row * slices) * 32 + lane; c < H; c += H) {     // a single iteration
    const float g0 = g[(row * 3 + 0) * H + c], g1 = g[(row * 3 + 1) * H + c], g2 = g[(row * 3 + 2) * H + c];
    for (int kk = b; kk < e; ++kk) {
      const int64_t j = col[kk];
      float gs1 = g0 * vec[(j * 3 + 0) * H + c] + g1 * vec[(j * 3 + 1) * H + c] + g2 * vec[(j * 3 + 2) * H + c];
      float gs2 = g0 * dhat[3 * kk + 0] + g1 * dhat[3 * kk + 1] + g2 * dhat[3 * kk + 2];
      if (s12_pre) {     // gradients of the pre-activations
        gs1 *= silu_grad(s12_pre[(int64_t)kk * 2 * H + c]);
        gs2 *= silu_grad(s12_pre[(int64_t)kk * 2 * H + H + c]);
      }
      g_s12[(int64_t)kk * 2 * H + c] = gs1;
      g_s12[(int64_t)kk * 2 * H + H + c] = gs2;
    }
  }
}

// gradient of vec at the SOURCE atoms (warp per source row of the transpose)
__global__ void __launch_bounds__(256)
vis_vecagg_bwd_vec_kernel(const float* __restrict__ g, const float* __restrict__ s12, const int32_t* __restrict__ rowptr_t,
                          const int32_t* __restrict__ col_t, const int32_t* __restrict__ eid_t, int64_t N, int H, int pre_act,
                          float* __restrict__ g_vec) {
  // one warp per (row, 32-channel slice): H / 32 times the warps of a warp-per-row mapping (the rows of a small batch
  // do not fill the GPU, and the edge loop of a row is a serial chain)
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int slices = (H + 31) >> 5;
  const int64_t row = wid / slices;
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const int b = rowptr_t[row], e = rowptr_t[row + 1];
  for (int c = (int)(wid - row * slices) * 32 + lane; c < H; c += H) {     // a single iteration
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
    for (int kk = b; kk < e; ++kk) {
      const int64_t i = col_t[kk];
      float s1 = s12[(int64_t)eid_t[kk] * 2 * H + c];
      if (pre_act) s1 = silu(s1);
      a0 += g[(i * 3 + 0) * H + c] * s1;
      a1 += g[(i * 3 + 1) * H + c] * s1;
      a2 += g[(i * 3 + 2) * H + c] * s1;
    }
    g_vec[(row * 3 + 0) * H + c] = a0;
    g_vec[(row * 3 + 1) * H + c] = a1;
    g_vec[(row * 3 + 2) * H + c] = a2;
  }
}

// ---- ViS_MP.edge_update (tgv.py:655-661) with w_trg / w_src hoisted to the nodes (they are bias-free linears):
//   w1 = P wt_i, w2 = P ws_j, P = I - dhat dhat^T  =>  wdot = wt_i . ws_j - (wt_i . dhat)(ws_j . dhat)
__global__ void __launch_bounds__(256)
vis_edge_update_fwd_kernel(const float* __restrict__ wt, const float* __restrict__ ws, const float* __restrict__ dhat,
                           const float* __restrict__ fpa, const int32_t* __restrict__ col,
                           const int32_t* __restrict__ erow, int64_t E, int H, int pre_act, float* __restrict__ df,
                           float* __restrict__ wdot) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = E * H;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; t < total; t += stride) {
    const int64_t e = t / H;
    const int c = (int)(t - e * H);
    const int64_t i = erow[e], j = col[e];
    const float d0 = dhat[3 * e], d1 = dhat[3 * e + 1], d2 = dhat[3 * e + 2];
    const float t0 = wt[(i * 3 + 0) * H + c], t1 = wt[(i * 3 + 1) * H + c], t2 = wt[(i * 3 + 2) * H + c];
    const float s0 = ws[(j * 3 + 0) * H + c], s1 = ws[(j * 3 + 1) * H + c], s2 = ws[(j * 3 + 2) * H + c];
    const float a = t0 * d0 + t1 * d1 + t2 * d2, b = s0 * d0 + s1 * d1 + s2 * d2;
    const float w = t0 * s0 + t1 * s1 + t2 * s2 - a * b;
    wdot[t] = w;
    df[t] = (pre_act ? silu(fpa[t]) : fpa[t]) * w;     // pre_act: fpa = the output of f_proj (tgv.py:659), SiLU applied here
  }
}

// backward of df = act(fpa) * wdot w.r.t. its two factors:  g_fpa = g * wdot [* silu'(fpa)],  gw = g * act(fpa)
__global__ void vis_edge_update_bwd_prep_kernel(const float* __restrict__ g, const float* __restrict__ fpa,
                                                const float* __restrict__ wdot, int64_t n, int pre_act,
                                                float* __restrict__ g_fpa, float* __restrict__ gw) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; t < n; t += stride) {
    const float gv = g[t], f = fpa[t];
    g_fpa[t] = gv * wdot[t] * (pre_act ? silu_grad(f) : 1.0f);
    gw[t] = gv * (pre_act ? silu(f) : f);
  }
}

// d wt[i] = sum_{e in row i} gw[e] * (ws_j - b dhat)   (by_src = 0, CSR rows, neighbour = col)
// d ws[j] = sum_{e: src = j} gw[e] * (wt_i - a dhat)   (by_src = 1, transposed rows, neighbour = col_t, edge = eid_t)
__global__ void __launch_bounds__(256)
vis_edge_update_bwd_kernel(const float* __restrict__ gw, const float* __restrict__ other, const float* __restrict__ dhat,
                           const int32_t* __restrict__ rowptr, const int32_t* __restrict__ nbr,
                           const int32_t* __restrict__ eid, int64_t N, int H, float* __restrict__ out) {
  // one warp per (row, 32-channel slice): H / 32 times the warps of a warp-per-row mapping (the rows of a small batch
  // do not fill the GPU, and the edge loop of a row is a serial chain)
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int slices = (H + 31) >> 5;
  const int64_t row = wid / slices;
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const int b = rowptr[row], e = rowptr[row + 1];
  for (int c = (int)(wid - row * slices) * 32 + lane; c < H; c += H) {     // a single iteration
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
    for (int kk = b; kk < e; ++kk) {
      const int64_t o = nbr[kk];
      const int64_t ed = eid ? eid[kk] : kk;
      const float d0 = dhat[3 * ed], d1 = dhat[3 * ed + 1], d2 = dhat[3 * ed + 2];
      const float o0 = other[(o * 3 + 0) * H + c], o1 = other[(o * 3 + 1) * H + c], o2 = other[(o * 3 + 2) * H + c];
      const float proj = o0 * d0 + o1 * d1 + o2 * d2;
      const float g = gw[ed * H + c];
      a0 += g * (o0 - proj * d0);
      a1 += g * (o1 - proj * d1);
      a2 += g * (o2 - proj * d2);
    }
    out[(row * 3 + 0) * H + c] = a0;
    out[(row * 3 + 1) * H + c] = a1;
    out[(row * 3 + 2) * H + c] = a2;
  }
}

int grid1d(int64_t n) {
  int64_t b = ceil_div(n, 256);
  const int64_t cap = (int64_t)sm_count() * 32;
  if (b < 1) b = 1;
  return (int)(b < cap ? b : cap);
}

// ---- ViS_MP node update (tgv.py:616-627): the element-wise tail of a layer in one kernel per direction ----
//   vp[N, 3, 3H] = vec_proj(vec) = [vec1 | vec2 | vec3],  o[N, 3H] = o_proj(x_agg) = [o1 | o2 | o3]
//   vec_dot = sum_d vec1 vec2;   dx = vec_dot o2 + o3;   dvec[d] = vec3[d] o1 + vec_agg[d]
// (the reference spells this as two splits, a product, a reduction and four more element-wise ops: ~8 launches forward
// and ~14 in the backward pass of the same expressions)
__global__ void vis_node_update_fwd_kernel(const float* __restrict__ vp, const float* __restrict__ o,
                                           const float* __restrict__ vagg, int64_t N, int H, float* __restrict__ dx,
                                           float* __restrict__ dvec) {
  const int64_t total = N * H;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = t / H;
    const int c = (int)(t - n * H);
    const float o1 = o[n * 3 * H + c], o2 = o[n * 3 * H + H + c], o3 = o[n * 3 * H + 2 * H + c];
    float dot = 0.0f;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float* r = vp + (n * 3 + d) * 3 * H;
      dot += r[c] * r[H + c];
      dvec[(n * 3 + d) * H + c] = r[2 * H + c] * o1 + vagg[(n * 3 + d) * H + c];
    }
    dx[t] = dot * o2 + o3;
  }
}

__global__ void vis_node_update_bwd_kernel(const float* __restrict__ g_dx, const float* __restrict__ g_dvec,
                                           const float* __restrict__ vp, const float* __restrict__ o, int64_t N, int H,
                                           float* __restrict__ g_vp, float* __restrict__ g_o) {
  const int64_t total = N * H;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = t / H;
    const int c = (int)(t - n * H);
    const float o1 = o[n * 3 * H + c], o2 = o[n * 3 * H + H + c];
    const float gx = g_dx[t];
    const float gdot = gx * o2;
    float dot = 0.0f, go1 = 0.0f;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float* r = vp + (n * 3 + d) * 3 * H;
      float* gr = g_vp + (n * 3 + d) * 3 * H;
      const float v1 = r[c], v2 = r[H + c], v3 = r[2 * H + c];
      const float gd = g_dvec[(n * 3 + d) * H + c];
      dot += v1 * v2;
      gr[c] = gdot * v2;
      gr[H + c] = gdot * v1;
      gr[2 * H + c] = gd * o1;
      go1 += gd * v3;
    }
    g_o[n * 3 * H + c] = go1;
    g_o[n * 3 * H + H + c] = gx * dot;
    g_o[n * 3 * H + 2 * H + c] = gx;
  }
}

unsigned warp_rows_grid(int64_t rows) { return (unsigned)ceil_div(rows * 32, 256); }

}  // namespace
}  // namespace cmp

using namespace cmp;

extern "C" int cmp_vis_edge_geometry(const float* evec, const float* dist, const int32_t* col, const int32_t* erow,
                                     int64_t E, float cutoff, float alpha, const float* means, const float* betas,
                                     int num_rbf, float* rbf, float* dhat, float* C, cmp_stream_t stream) {
  CMP_REQUIRE(E >= 0 && num_rbf >= 1 && cutoff > 0.0f, CMP_EINVAL, "cmp_vis_edge_geometry: bad size");
  if (E == 0) return CMP_OK;
  CMP_REQUIRE(evec && dist && col && erow && means && betas && rbf && dhat && C, CMP_EINVAL,
              "cmp_vis_edge_geometry: null pointer");
  vis_edge_geometry_kernel<<<grid1d(E * num_rbf), 256, 0, as_stream(stream)>>>(evec, dist, col, erow, E, cutoff, alpha,
                                                                              means, betas, num_rbf, rbf, dhat, C);
  CMP_LAUNCH_CHECK("cmp_vis_edge_geometry");
  return CMP_OK;
}

extern "C" int cmp_layernorm_fwd(const float* x, const float* w, const float* b, int64_t M, int H, float eps, float* y,
                                 float* mean, float* rstd, cmp_stream_t stream) {
  CMP_REQUIRE(M >= 0 && H >= 1, CMP_EINVAL, "cmp_layernorm_fwd: bad size");
  if (M == 0) return CMP_OK;
  CMP_REQUIRE(x && w && b && y && mean && rstd, CMP_EINVAL, "cmp_layernorm_fwd: null pointer");
  layernorm_fwd_kernel<<<warp_rows_grid(M), 256, 0, as_stream(stream)>>>(x, w, b, M, H, eps, y, mean, rstd);
  CMP_LAUNCH_CHECK("cmp_layernorm_fwd");
  return CMP_OK;
}

extern "C" int cmp_layernorm_bwd(const float* dy, const float* x, const float* w, const float* mean, const float* rstd,
                                 int64_t M, int H, float* dx, float* dyxhat, cmp_stream_t stream) {
  CMP_REQUIRE(M >= 0 && H >= 1, CMP_EINVAL, "cmp_layernorm_bwd: bad size");
  if (M == 0) return CMP_OK;
  CMP_REQUIRE(dy && x && w && mean && rstd && dx && dyxhat, CMP_EINVAL, "cmp_layernorm_bwd: null pointer");
  layernorm_bwd_kernel<<<warp_rows_grid(M), 256, 0, as_stream(stream)>>>(dy, x, w, mean, rstd, M, H, dx, dyxhat);
  CMP_LAUNCH_CHECK("cmp_layernorm_bwd");
  return CMP_OK;
}

extern "C" int cmp_csr_segment_sum(const float* x, const int32_t* rowptr, const int32_t* perm, int64_t N, int C,
                                   float* out, cmp_stream_t stream) {
  CMP_REQUIRE(N >= 0 && C >= 1, CMP_EINVAL, "cmp_csr_segment_sum: bad size");
  if (N == 0) return CMP_OK;
  CMP_REQUIRE(rowptr && out, CMP_EINVAL, "cmp_csr_segment_sum: null pointer");
  csr_segment_sum_kernel<<<warp_rows_grid(N * ((C + 31) / 32)), 256, 0, as_stream(stream)>>>(x, rowptr, perm, N, C, out);
  CMP_LAUNCH_CHECK("cmp_csr_segment_sum");
  return CMP_OK;
}

extern "C" int cmp_gather_rows(const float* x, const int32_t* idx, int64_t E, int C, float* out, cmp_stream_t stream) {
  CMP_REQUIRE(E >= 0 && C >= 1, CMP_EINVAL, "cmp_gather_rows: bad size");
  if (E == 0) return CMP_OK;
  CMP_REQUIRE(x && idx && out, CMP_EINVAL, "cmp_gather_rows: null pointer");
  gather_rows_kernel<<<grid1d(E * C), 256, 0, as_stream(stream)>>>(x, idx, E, C, out);
  CMP_LAUNCH_CHECK("cmp_gather_rows");
  return CMP_OK;
}

extern "C" int cmp_vis_edge_embed_fwd(const float* x, const float* ep, const int32_t* col, const int32_t* erow, int64_t E,
                                      int H, float* f, cmp_stream_t stream) {
  CMP_REQUIRE(E >= 0 && H >= 1, CMP_EINVAL, "cmp_vis_edge_embed_fwd: bad size");
  if (E == 0) return CMP_OK;
  CMP_REQUIRE(x && ep && col && erow && f, CMP_EINVAL, "cmp_vis_edge_embed_fwd: null pointer");
  vis_edge_embed_fwd_kernel<<<grid1d(E * H), 256, 0, as_stream(stream)>>>(x, ep, col, erow, E, H, f);
  CMP_LAUNCH_CHECK("cmp_vis_edge_embed_fwd");
  return CMP_OK;
}

extern "C" int cmp_vis_edge_embed_bwd(const float* g, const float* x, const float* ep, const int32_t* col,
                                      const int32_t* erow, int64_t E, int H, float* gep, float* t_out,
                                      cmp_stream_t stream) {
  CMP_REQUIRE(E >= 0 && H >= 1, CMP_EINVAL, "cmp_vis_edge_embed_bwd: bad size");
  if (E == 0) return CMP_OK;
  CMP_REQUIRE(g && x && ep && col && erow && gep && t_out, CMP_EINVAL, "cmp_vis_edge_embed_bwd: null pointer");
  vis_edge_embed_bwd_kernel<<<grid1d(E * H), 256, 0, as_stream(stream)>>>(g, x, ep, col, erow, E, H, gep, t_out);
  CMP_LAUNCH_CHECK("cmp_vis_edge_embed_bwd");
  return CMP_OK;
}

extern "C" int cmp_vis_message_fwd(const float* q, const float* k, const float* v, const float* dk, const float* dv,
                                   const float* C, const int32_t* col, const int32_t* erow, int64_t E, int H, int heads,
                                   int pre_act, float* m, float* attn_pre, cmp_stream_t stream) {
  CMP_REQUIRE(E >= 0 && H >= 1 && heads >= 1 && H % heads == 0 && heads <= 64, CMP_EINVAL, "cmp_vis_message_fwd: bad size");
  if (E == 0) return CMP_OK;
  CMP_REQUIRE(q && k && v && dk && dv && C && col && erow && m && attn_pre, CMP_EINVAL, "cmp_vis_message_fwd: null pointer");
  vis_message_fwd_kernel<<<(unsigned)ceil_div(E, 8), 256, 8 * heads * sizeof(float), as_stream(stream)>>>(
      q, k, v, dk, dv, C, col, erow, E, H, heads, pre_act, m, attn_pre);
  CMP_LAUNCH_CHECK("cmp_vis_message_fwd");
  return CMP_OK;
}

extern "C" int cmp_vis_message_bwd(const float* gm, const float* q, const float* k, const float* v, const float* dk,
                                   const float* dv, const float* C, const float* attn_pre, const int32_t* col,
                                   const int32_t* erow, int64_t E, int H, int heads, int pre_act, float* g_dk, float* g_dv,
                                   float* geq, float* gek, float* gev, cmp_stream_t stream) {
  CMP_REQUIRE(E >= 0 && H >= 1 && heads >= 1 && H % heads == 0 && heads <= 64, CMP_EINVAL, "cmp_vis_message_bwd: bad size");
  if (E == 0) return CMP_OK;
  CMP_REQUIRE(gm && q && k && v && dk && dv && C && attn_pre && col && erow && g_dk && g_dv && geq && gek && gev,
              CMP_EINVAL, "cmp_vis_message_bwd: null pointer");
  vis_message_bwd_kernel<<<(unsigned)ceil_div(E, 8), 256, 8 * heads * sizeof(float), as_stream(stream)>>>(
      gm, q, k, v, dk, dv, C, attn_pre, col, erow, E, H, heads, pre_act, g_dk, g_dv, geq, gek, gev);
  CMP_LAUNCH_CHECK("cmp_vis_message_bwd");
  return CMP_OK;
}

extern "C" int cmp_vis_vecagg_fwd(const float* vec, const float* s12, const float* dhat, const int32_t* rowptr,
                                  const int32_t* col, int64_t N, int H, int pre_act, float* vagg, cmp_stream_t stream) {
  CMP_REQUIRE(N >= 0 && H >= 1, CMP_EINVAL, "cmp_vis_vecagg_fwd: bad size");
  if (N == 0) return CMP_OK;
  CMP_REQUIRE(vec && rowptr && vagg, CMP_EINVAL, "cmp_vis_vecagg_fwd: null pointer");
  vis_vecagg_fwd_kernel<<<warp_rows_grid(N * ((H + 31) / 32)), 256, 0, as_stream(stream)>>>(vec, s12, dhat, rowptr, col, N, H, pre_act, vagg);
  CMP_LAUNCH_CHECK("cmp_vis_vecagg_fwd");
  return CMP_OK;
}

extern "C" int cmp_vis_vecagg_bwd(const float* g, const float* vec, const float* s12, const float* dhat,
                                  const int32_t* rowptr, const int32_t* col, const int32_t* rowptr_t, const int32_t* col_t,
                                  const int32_t* eid_t, int64_t N, int H, int pre_act, float* g_s12, float* g_vec,
                                  cmp_stream_t stream) {
  CMP_REQUIRE(N >= 0 && H >= 1, CMP_EINVAL, "cmp_vis_vecagg_bwd: bad size");
  if (N == 0) return CMP_OK;
  CMP_REQUIRE(g && vec && rowptr && rowptr_t && col_t && eid_t, CMP_EINVAL, "cmp_vis_vecagg_bwd: null pointer");
  if (g_s12) {
    vis_vecagg_bwd_s_kernel<<<warp_rows_grid(N * ((H + 31) / 32)), 256, 0, as_stream(stream)>>>(g, vec, dhat, rowptr, col, N, H,
                                                                              pre_act ? s12 : nullptr, g_s12);
    CMP_LAUNCH_CHECK("cmp_vis_vecagg_bwd(s)");
  }
  if (g_vec) {
    vis_vecagg_bwd_vec_kernel<<<warp_rows_grid(N * ((H + 31) / 32)), 256, 0, as_stream(stream)>>>(g, s12, rowptr_t, col_t, eid_t, N, H,
                                                                                pre_act, g_vec);
    CMP_LAUNCH_CHECK("cmp_vis_vecagg_bwd(vec)");
  }
  return CMP_OK;
}

extern "C" int cmp_vis_edge_update_fwd(const float* wt, const float* ws, const float* dhat, const float* fpa,
                                       const int32_t* col, const int32_t* erow, int64_t E, int H, int pre_act, float* df,
                                       float* wdot, cmp_stream_t stream) {
  CMP_REQUIRE(E >= 0 && H >= 1, CMP_EINVAL, "cmp_vis_edge_update_fwd: bad size");
  if (E == 0) return CMP_OK;
  CMP_REQUIRE(wt && ws && dhat && fpa && col && erow && df && wdot, CMP_EINVAL, "cmp_vis_edge_update_fwd: null pointer");
  vis_edge_update_fwd_kernel<<<grid1d(E * H), 256, 0, as_stream(stream)>>>(wt, ws, dhat, fpa, col, erow, E, H, pre_act, df,
                                                                           wdot);
  CMP_LAUNCH_CHECK("cmp_vis_edge_update_fwd");
  return CMP_OK;
}

extern "C" int cmp_vis_edge_update_bwd_prep(const float* g, const float* fpa, const float* wdot, int64_t n, int pre_act,
                                            float* g_fpa, float* gw, cmp_stream_t stream) {
  CMP_REQUIRE(n >= 0, CMP_EINVAL, "cmp_vis_edge_update_bwd_prep: bad size");
  if (n == 0) return CMP_OK;
  CMP_REQUIRE(g && fpa && wdot && g_fpa && gw, CMP_EINVAL, "cmp_vis_edge_update_bwd_prep: null pointer");
  vis_edge_update_bwd_prep_kernel<<<grid1d(n), 256, 0, as_stream(stream)>>>(g, fpa, wdot, n, pre_act, g_fpa, gw);
  CMP_LAUNCH_CHECK("cmp_vis_edge_update_bwd_prep");
  return CMP_OK;
}

extern "C" int cmp_vis_edge_update_bwd(const float* gw, const float* wt, const float* ws, const float* dhat,
                                       const int32_t* rowptr, const int32_t* col, const int32_t* rowptr_t,
                                       const int32_t* col_t, const int32_t* eid_t, int64_t N, int H, float* g_wt,
                                       float* g_ws, cmp_stream_t stream) {
  CMP_REQUIRE(N >= 0 && H >= 1, CMP_EINVAL, "cmp_vis_edge_update_bwd: bad size");
  if (N == 0) return CMP_OK;
  CMP_REQUIRE(gw && wt && ws && dhat && rowptr && col && rowptr_t && col_t && eid_t && g_wt && g_ws, CMP_EINVAL,
              "cmp_vis_edge_update_bwd: null pointer");
  vis_edge_update_bwd_kernel<<<warp_rows_grid(N * ((H + 31) / 32)), 256, 0, as_stream(stream)>>>(gw, ws, dhat, rowptr, col, nullptr, N, H,
                                                                              g_wt);
  CMP_LAUNCH_CHECK("cmp_vis_edge_update_bwd(wt)");
  vis_edge_update_bwd_kernel<<<warp_rows_grid(N * ((H + 31) / 32)), 256, 0, as_stream(stream)>>>(gw, wt, dhat, rowptr_t, col_t, eid_t, N, H,
                                                                              g_ws);
  CMP_LAUNCH_CHECK("cmp_vis_edge_update_bwd(ws)");
  return CMP_OK;
}

extern "C" int cmp_vis_node_update_fwd(const float* vp, const float* o, const float* vagg, int64_t N, int H, float* dx,
                                       float* dvec, cmp_stream_t stream) {
  CMP_REQUIRE(N >= 0 && H >= 1, CMP_EINVAL, "cmp_vis_node_update_fwd: bad size");
  if (N == 0) return CMP_OK;
  CMP_REQUIRE(vp && o && vagg && dx && dvec, CMP_EINVAL, "cmp_vis_node_update_fwd: null pointer");
  vis_node_update_fwd_kernel<<<grid1d(N * H), 256, 0, as_stream(stream)>>>(vp, o, vagg, N, H, dx, dvec);
  CMP_LAUNCH_CHECK("cmp_vis_node_update_fwd");
  return CMP_OK;
}

extern "C" int cmp_vis_node_update_bwd(const float* g_dx, const float* g_dvec, const float* vp, const float* o, int64_t N,
                                       int H, float* g_vp, float* g_o, cmp_stream_t stream) {
  CMP_REQUIRE(N >= 0 && H >= 1, CMP_EINVAL, "cmp_vis_node_update_bwd: bad size");
  if (N == 0) return CMP_OK;
  CMP_REQUIRE(g_dx && g_dvec && vp && o && g_vp && g_o, CMP_EINVAL, "cmp_vis_node_update_bwd: null pointer");
  vis_node_update_bwd_kernel<<<grid1d(N * H), 256, 0, as_stream(stream)>>>(g_dx, g_dvec, vp, o, N, H, g_vp, g_o);
  CMP_LAUNCH_CHECK("cmp_vis_node_update_bwd");
  return CMP_OK;
}
