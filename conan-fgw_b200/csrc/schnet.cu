// SchNet pieces of the exact-fp32 path: Gaussian RBF, embedding, CFConv message/aggregate (CSR
// gather-reduce, deterministic), sum readout.  The E x F filter is materialised only on this
// path; the fused tcgen05 kernel (cfconv_tc.cu) keeps it on chip.
#include "common.cuh"

namespace cmp {
namespace {

__global__ void rbf_gaussian_kernel(const float* __restrict__ d, int64_t E, const float* __restrict__ offset, int Ng,
                                    float coeff, float* __restrict__ out, int64_t ldo) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = E * Ng;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; idx < total; idx += stride) {
    int64_t e = idx / Ng;
    int k = (int)(idx - e * Ng);
    float t = d[e] - offset[k];
    out[e * ldo + k] = expf(coeff * (t * t));
  }
}

__global__ void embedding_fwd_kernel(const int64_t* __restrict__ z, int64_t N, const float* __restrict__ w, int V,
                                     int H, float* __restrict__ out, int* status) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = N * H;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; idx < total; idx += stride) {
    int64_t i = idx / H;
    int c = (int)(idx - i * H);
    int64_t zi = z[i];
    if (zi < 0 || zi >= V) {
      if (c == 0) atomicOr(status, CMP_STATUS_BAD_ATOMIC_NUMBER);
      out[idx] = 0.0f;
    } else {
      out[idx] = w[zi * H + c];
    }
  }
}

// H % 4 == 0: one float4 per thread
__global__ void embedding_fwd_vec4_kernel(const int64_t* __restrict__ z, int64_t N, const float4* __restrict__ w, int V,
                                          int H4, float4* __restrict__ out, int* status) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * H4) return;
  const int64_t i = idx / H4;
  const int c = (int)(idx - i * H4);
  const int64_t zi = __ldg(z + i);
  if (zi < 0 || zi >= V) {
    if (c == 0) atomicOr(status, CMP_STATUS_BAD_ATOMIC_NUMBER);
    out[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    out[idx] = __ldg(w + zi * H4 + c);
  }
}

// stage 1: each CTA owns a contiguous chunk of atoms and accumulates dweight rows in shared memory
// sequentially (fixed order); stage 2 sums the chunk partials in a fixed order.
// The chunk's atomic numbers are staged in shared memory once, and the gradient rows of 32 atoms are loaded together
// (the loop used to expose one global-memory latency per 8 atoms: 29 us at cfg 2); the accumulation keeps the atom order.
__global__ void embedding_bwd_stage1(const int64_t* __restrict__ z, int64_t N, const float* __restrict__ dout, int V,
                                     int H, int chunk, float* __restrict__ partial) {
  extern __shared__ __align__(16) float acc[];  // [V][H] | int zs[chunk]
  int* zs = reinterpret_cast<int*>(acc + (size_t)V * H);
  const int total = V * H;
  const int64_t i0 = (int64_t)blockIdx.x * chunk;
  const int cnt = (int)((i0 + chunk > N ? N : i0 + chunk) - i0);
  for (int t = threadIdx.x; t < total; t += blockDim.x) acc[t] = 0.0f;
  for (int t = threadIdx.x; t < cnt; t += blockDim.x) {
    const int64_t zz = __ldg(z + i0 + t);
    zs[t] = (zz >= 0 && zz < V) ? (int)zz : -1;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    const float* src = dout + i0 * H + c;
    for (int i = 0; i < cnt; i += 32) {
      float dd[32];
#pragma unroll
      for (int u = 0; u < 32; ++u) dd[u] = (i + u < cnt) ? __ldg(src + (int64_t)(i + u) * H) : 0.0f;
#pragma unroll
      for (int u = 0; u < 32; ++u) {
        const int zq = (i + u < cnt) ? zs[i + u] : -1;
        if (zq >= 0) acc[zq * H + c] += dd[u];
      }
    }
  }
  __syncthreads();
  float* P = partial + (int64_t)blockIdx.x * total;
  for (int t = threadIdx.x; t < total; t += blockDim.x) P[t] = acc[t];
}

__global__ void embedding_bwd_stage2(const float* __restrict__ partial, int chunks, int V, int H, int padding_idx,
                                     float* __restrict__ dweight) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= V * H) return;
  float s = 0.0f;
  for (int c = 0; c < chunks; ++c) s += partial[(int64_t)c * V * H + t];
  if (t / H == padding_idx) s = 0.0f;
  dweight[t] = s;
}

// V * H % 4 == 0: a CTA sums 64 consecutive values (16 float4) with 16 chunk lanes per value - lane l adds chunks l, l + 16,
// ... in order, then the 16 lane sums are added in lane order (a fixed order: deterministic)
__global__ void __launch_bounds__(256)
embedding_bwd_stage2_vec4(const float4* __restrict__ partial, int chunks, int total4, int H, int padding_idx,
                          float4* __restrict__ dweight) {
  __shared__ float4 part[16][16];
  const int f = threadIdx.x & 15, l = threadIdx.x >> 4;
  const int t4 = blockIdx.x * 16 + f;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t4 < total4) {
#pragma unroll 4
    for (int c = l; c < chunks; c += 16) {
      const float4 v = __ldg(partial + (int64_t)c * total4 + t4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  part[l][f] = s;
  __syncthreads();
  if (l == 0 && t4 < total4) {
    float4 r = part[0][f];
#pragma unroll
    for (int k = 1; k < 16; ++k) {
      const float4 v = part[k][f];
      r.x += v.x; r.y += v.y; r.z += v.z; r.w += v.w;
    }
    if ((t4 * 4) / H == padding_idx) r = make_float4(0.f, 0.f, 0.f, 0.f);   // H % 4 == 0: a float4 never straddles rows
    dweight[t4] = r;
  }
}

int embedding_chunk(int64_t N) {
  int64_t c = ceil_div(N, 1024);   // at most 1024 chunks; short chunks keep the sequential loops short
  return (int)(c < 128 ? 128 : c);
}

// One warp per target row; lanes stride the channel axis (float4 when F % 4 == 0).
template <int VEC>
__global__ void __launch_bounds__(256)
cfconv_message_fwd_kernel(const float* __restrict__ xprime, const float* __restrict__ filt,
                          const float* __restrict__ dist, const int32_t* __restrict__ rowptr,
                          const int32_t* __restrict__ col, int64_t N, int F, float cutoff, float* __restrict__ agg,
                          const float* __restrict__ scale) {
  int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= N) return;
  int b = rowptr[row], e = rowptr[row + 1];
  for (int c = lane * VEC; c < F; c += 32 * VEC) {
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.0f;
    for (int k = b; k < e; ++k) {
      int j = col[k];
      float C = scale ? scale[k] : 0.5f * (cosf(dist[k] * kPi / cutoff) + 1.0f);
      if (VEC == 4) {
        float4 x = *reinterpret_cast<const float4*>(xprime + (int64_t)j * F + c);
        float4 w = *reinterpret_cast<const float4*>(filt + (int64_t)k * F + c);
        acc[0] += x.x * (w.x * C);
        acc[1 % VEC] += x.y * (w.y * C);
        acc[2 % VEC] += x.z * (w.z * C);
        acc[3 % VEC] += x.w * (w.w * C);
      } else {
        acc[0] += xprime[(int64_t)j * F + c] * (filt[(int64_t)k * F + c] * C);
      }
    }
    if (VEC == 4) {
      *reinterpret_cast<float4*>(agg + row * F + c) = make_float4(acc[0], acc[1 % VEC], acc[2 % VEC], acc[3 % VEC]);
    } else {
      agg[row * F + c] = acc[0];
    }
  }
}

// dfilt[e] = g[dst(e)] * xprime[col[e]] * C(e)   (warp per target row)
__global__ void __launch_bounds__(256)
cfconv_dfilt_kernel(const float* __restrict__ g, const float* __restrict__ xprime, const float* __restrict__ dist,
                    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t N, int F,
                    float cutoff, float* __restrict__ dfilt, const float* __restrict__ scale) {
  int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= N) return;
  int b = rowptr[row], e = rowptr[row + 1];
  for (int k = b; k < e; ++k) {
    int j = col[k];
    float C = scale ? scale[k] : 0.5f * (cosf(dist[k] * kPi / cutoff) + 1.0f);
    for (int c = lane; c < F; c += 32)
      dfilt[(int64_t)k * F + c] = g[row * F + c] * xprime[(int64_t)j * F + c] * C;
  }
}

// dxprime[j] = sum over the transposed row of j   (warp per source atom; fixed order)
__global__ void __launch_bounds__(256)
cfconv_dx_kernel(const float* __restrict__ g, const float* __restrict__ filt, const float* __restrict__ dist,
                 const int32_t* __restrict__ rowptr_t, const int32_t* __restrict__ col_t,
                 const int32_t* __restrict__ eid_t, int64_t N, int F, float cutoff, float* __restrict__ dx,
                 const float* __restrict__ scale) {
  int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= N) return;
  int b = rowptr_t[row], e = rowptr_t[row + 1];
  for (int c = lane; c < F; c += 32) {
    float acc = 0.0f;
    for (int k = b; k < e; ++k) {
      int i = col_t[k];
      int eid = eid_t[k];
      float C = scale ? scale[eid] : 0.5f * (cosf(dist[eid] * kPi / cutoff) + 1.0f);
      acc += g[(int64_t)i * F + c] * (filt[(int64_t)eid * F + c] * C);
    }
    dx[row * F + c] = acc;
  }
}

__global__ void segment_sum_fwd_kernel(const float* __restrict__ x, const int32_t* __restrict__ seg_ptr, int C,
                                       float* __restrict__ out) {
  int g = blockIdx.x;
  int s = seg_ptr[g], e = seg_ptr[g + 1];
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.0f;
    for (int i = s; i < e; ++i) acc += x[(int64_t)i * C + c];
    out[(int64_t)g * C + c] = acc;
  }
}

__global__ void segment_sum_bwd_kernel(const float* __restrict__ dout, const int32_t* __restrict__ seg_ptr, int C,
                                       float* __restrict__ dx) {
  int g = blockIdx.x;
  int s = seg_ptr[g], e = seg_ptr[g + 1];
  int64_t total = (int64_t)(e - s) * C;
  for (int64_t t = threadIdx.x; t < total; t += blockDim.x) {
    int c = (int)(t % C);
    dx[(int64_t)s * C + t] = dout[(int64_t)g * C + c];
  }
}

int grid1d(int64_t n, int per_block = 256) {
  int64_t b = ceil_div(n, per_block);
  int64_t cap = (int64_t)sm_count() * 32;
  if (b < 1) b = 1;
  return (int)(b < cap ? b : cap);
}

}  // namespace

// ---- regression head of a ConAN training step in two launches (schnet_based_models.py:242 conformer mean,
// :17-29 linear head, model/common.py:288 MSE):  mol[b] = mean_k emb[b K + k];  pred[b] = w . mol[b] + bias;
// loss = mean_b (pred[b] - target[b])^2.  Replaces ~20 launches on tensors of a few KB (segment mean, tiny GEMMs and
// their gradients, MSE forward / backward, fills).  One CTA: sums run in a fixed order (deterministic).
constexpr int HEAD_THREADS = 1024;      // one CTA of 32 warps: a warp per molecule, lanes over the channels (coalesced rows)
constexpr int HEAD_WARPS = HEAD_THREADS / 32;
constexpr int HEAD_MAXC = 512;          // channels per lane: HEAD_MAXC / 32

__global__ void __launch_bounds__(HEAD_THREADS)
regression_head_fwd_kernel(const float* __restrict__ emb, int64_t ld, int B, int K, int C, const float* __restrict__ w,
                           const float* __restrict__ bias, const float* __restrict__ target, float* __restrict__ err,
                           float* __restrict__ loss) {
  __shared__ float red[HEAD_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float invK = 1.0f / (float)K;
  const float b0 = bias ? bias[0] : 0.0f;
  float sq = 0.0f;     // lane 0 of every warp: squared errors of its molecules, in molecule order
  for (int b = warp; b < B; b += HEAD_WARPS) {
    float part = 0.0f;
    for (int c = lane; c < C; c += 32) {
      float m = 0.0f;
      for (int k = 0; k < K; ++k) m += emb[((int64_t)b * K + k) * ld + c];
      part = fmaf(m * invK, w[c], part);
    }
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    const float e = part + b0 - target[b];
    if (lane == 0) err[b] = e;
    sq += e * e;
  }
  if (lane == 0) red[warp] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int i = 0; i < HEAD_WARPS; ++i) t += red[i];     // fixed order: deterministic
    loss[0] = t / (float)B;
  }
}

// d loss / d pred[b] = gscale * 2 err[b] / B.  A warp per molecule writes its d emb rows and keeps its share of dw in
// registers (lane c); the per-warp shares are added in warp order.
__global__ void __launch_bounds__(HEAD_THREADS)
regression_head_bwd_kernel(const float* __restrict__ emb, int64_t ld, int B, int K, int C, const float* __restrict__ w,
                           const float* __restrict__ err, const float* __restrict__ gscale, float* __restrict__ d_emb,
                           int64_t ldd, float* __restrict__ dw, float* __restrict__ db) {
  extern __shared__ float part[];      // [HEAD_WARPS][C] | [HEAD_WARPS]
  float* pdb = part + HEAD_WARPS * C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float gs = (gscale ? gscale[0] : 1.0f) * 2.0f / (float)B;
  const float invK = 1.0f / (float)K;
  float acc[HEAD_MAXC / 32];
#pragma unroll
  for (int r = 0; r < HEAD_MAXC / 32; ++r) acc[r] = 0.0f;
  float accb = 0.0f;
  for (int b = warp; b < B; b += HEAD_WARPS) {
    const float gl = gs * err[b];
    accb += gl;
#pragma unroll
    for (int r = 0; r < HEAD_MAXC / 32; ++r) {
      const int c = r * 32 + lane;
      if (c < C) {
        float m = 0.0f;
        for (int k = 0; k < K; ++k) m += emb[((int64_t)b * K + k) * ld + c];
        acc[r] = fmaf(gl, m * invK, acc[r]);
        const float d = gl * invK * w[c];
        for (int k = 0; k < K; ++k) d_emb[((int64_t)b * K + k) * ldd + c] = d;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < HEAD_MAXC / 32; ++r) {
    const int c = r * 32 + lane;
    if (c < C) part[warp * C + c] = acc[r];
  }
  if (lane == 0) pdb[warp] = accb;
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += HEAD_THREADS) {
    float t = 0.0f;
    for (int i = 0; i < HEAD_WARPS; ++i) t += part[i * C + c];
    dw[c] = t;
  }
  if (db && threadIdx.x == HEAD_THREADS - 1) {
    float t = 0.0f;
    for (int i = 0; i < HEAD_WARPS; ++i) t += pdb[i];
    db[0] = t;
  }
}

}  // namespace cmp

using namespace cmp;

extern "C" int cmp_rbf_gaussian_fwd(const float* d, int64_t E, const float* offset, int Ng, float coeff, float* out,
                                    int64_t ldo, cmp_stream_t stream) {
  CMP_REQUIRE(E >= 0 && Ng >= 1 && ldo >= Ng, CMP_EINVAL, "cmp_rbf_gaussian_fwd: bad size");
  if (E == 0) return CMP_OK;
  CMP_REQUIRE(d && offset && out, CMP_EINVAL, "cmp_rbf_gaussian_fwd: null pointer");
  rbf_gaussian_kernel<<<grid1d(E * Ng), 256, 0, as_stream(stream)>>>(d, E, offset, Ng, coeff, out, ldo);
  CMP_LAUNCH_CHECK("cmp_rbf_gaussian_fwd");
  return CMP_OK;
}

extern "C" int cmp_embedding_fwd(const int64_t* z, int64_t N, const float* weight, int V, int H, float* out,
                                 int* status, cmp_stream_t stream) {
  CMP_REQUIRE(N >= 0 && V >= 1 && H >= 1, CMP_EINVAL, "cmp_embedding_fwd: bad size");
  if (N == 0) return CMP_OK;
  CMP_REQUIRE(z && weight && out && status, CMP_EINVAL, "cmp_embedding_fwd: null pointer");
  if (H % 4 == 0 && ((reinterpret_cast<uintptr_t>(weight) | reinterpret_cast<uintptr_t>(out)) & 15) == 0)
    embedding_fwd_vec4_kernel<<<(unsigned)ceil_div(N * (H / 4), 256), 256, 0, as_stream(stream)>>>(
        z, N, reinterpret_cast<const float4*>(weight), V, H / 4, reinterpret_cast<float4*>(out), status);
  else
    embedding_fwd_kernel<<<grid1d(N * H), 256, 0, as_stream(stream)>>>(z, N, weight, V, H, out, status);
  CMP_LAUNCH_CHECK("cmp_embedding_fwd");
  return CMP_OK;
}

extern "C" size_t cmp_embedding_bwd_workspace(int64_t N, int V, int H) {
  if (N <= 0) return 256;
  int64_t chunks = ceil_div(N, embedding_chunk(N));
  return align_up((size_t)chunks * V * H * sizeof(float), 256);
}

extern "C" int cmp_embedding_bwd(const int64_t* z, int64_t N, const float* dout, int V, int H, int padding_idx,
                                 float* dweight, void* workspace, size_t workspace_bytes, cmp_stream_t stream) {
  CMP_REQUIRE(N >= 0 && V >= 1 && H >= 1, CMP_EINVAL, "cmp_embedding_bwd: bad size");
  CMP_REQUIRE(dweight, CMP_EINVAL, "cmp_embedding_bwd: null pointer");
  cudaStream_t st = as_stream(stream);
  if (N == 0) {
    CMP_REQUIRE(cudaMemsetAsync(dweight, 0, (size_t)V * H * sizeof(float), st) == cudaSuccess, CMP_ECUDA,
                "cmp_embedding_bwd: memset failed");
    return CMP_OK;
  }
  CMP_REQUIRE(z && dout, CMP_EINVAL, "cmp_embedding_bwd: null pointer");
  int chunk = embedding_chunk(N);
  size_t smem = (size_t)V * H * sizeof(float) + (size_t)chunk * sizeof(int);
  CMP_REQUIRE(smem <= 227 * 1024, CMP_EUNSUPPORTED, "cmp_embedding_bwd: V*H*4 = %zu bytes exceeds shared memory", smem);
  int chunks = (int)ceil_div(N, chunk);
  CMP_REQUIRE(workspace && workspace_bytes >= (size_t)chunks * V * H * sizeof(float), CMP_EWORKSPACE,
              "cmp_embedding_bwd: workspace too small");
  if (cudaFuncSetAttribute(embedding_bwd_stage1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
      cudaSuccess) {
    (void)cudaGetLastError();
    set_error("cmp_embedding_bwd: cannot opt in to %zu bytes of shared memory", smem);
    return CMP_ECUDA;
  }
  int threads = H < 256 ? ((H + 31) / 32) * 32 : 256;
  embedding_bwd_stage1<<<chunks, threads, smem, st>>>(z, N, dout, V, H, chunk, reinterpret_cast<float*>(workspace));
  CMP_LAUNCH_CHECK("cmp_embedding_bwd(stage1)");
  if (H % 4 == 0 && (reinterpret_cast<uintptr_t>(dweight) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0)
    embedding_bwd_stage2_vec4<<<(unsigned)ceil_div((int64_t)V * H / 4, 16), 256, 0, st>>>(
        reinterpret_cast<const float4*>(workspace), chunks, V * H / 4, H, padding_idx, reinterpret_cast<float4*>(dweight));
  else
    embedding_bwd_stage2<<<(unsigned)ceil_div((int64_t)V * H, 256), 256, 0, st>>>(reinterpret_cast<float*>(workspace),
                                                                                 chunks, V, H, padding_idx, dweight);
  CMP_LAUNCH_CHECK("cmp_embedding_bwd(stage2)");
  return CMP_OK;
}

extern "C" int cmp_cfconv_message_fwd(const float* xprime, const float* filt, const float* dist, const int32_t* rowptr,
                                      const int32_t* col, int64_t N, int F, float cutoff, float* agg,
                                      cmp_stream_t stream) {
  CMP_REQUIRE(N >= 0 && F >= 1 && cutoff > 0.0f, CMP_EINVAL, "cmp_cfconv_message_fwd: bad size");
  if (N == 0) return CMP_OK;
  CMP_REQUIRE(xprime && rowptr && agg, CMP_EINVAL, "cmp_cfconv_message_fwd: null pointer");
  unsigned blocks = (unsigned)ceil_div(N * 32, 256);
  bool vec = (F % 4 == 0) && (((uintptr_t)xprime | (uintptr_t)filt | (uintptr_t)agg) % 16 == 0);
  if (vec)
    cfconv_message_fwd_kernel<4><<<blocks, 256, 0, as_stream(stream)>>>(xprime, filt, dist, rowptr, col, N, F, cutoff,
                                                                        agg, nullptr);
  else
    cfconv_message_fwd_kernel<1><<<blocks, 256, 0, as_stream(stream)>>>(xprime, filt, dist, rowptr, col, N, F, cutoff,
                                                                        agg, nullptr);
  CMP_LAUNCH_CHECK("cmp_cfconv_message_fwd");
  return CMP_OK;
}

extern "C" int cmp_cfconv_message_bwd(const float* g, const float* xprime, const float* filt, const float* dist,
                                      const int32_t* rowptr, const int32_t* col, const int32_t* rowptr_t,
                                      const int32_t* col_t, const int32_t* eid_t, int64_t N, int F, float cutoff,
                                      float* dfilt, float* dxprime, cmp_stream_t stream) {
  CMP_REQUIRE(N >= 0 && F >= 1 && cutoff > 0.0f, CMP_EINVAL, "cmp_cfconv_message_bwd: bad size");
  if (N == 0) return CMP_OK;
  CMP_REQUIRE(g && rowptr, CMP_EINVAL, "cmp_cfconv_message_bwd: null pointer");
  unsigned blocks = (unsigned)ceil_div(N * 32, 256);
  if (dfilt) {
    CMP_REQUIRE(xprime && col && dist, CMP_EINVAL, "cmp_cfconv_message_bwd: null pointer (dfilt inputs)");
    cfconv_dfilt_kernel<<<blocks, 256, 0, as_stream(stream)>>>(g, xprime, dist, rowptr, col, N, F, cutoff, dfilt,
                                                               nullptr);
    CMP_LAUNCH_CHECK("cmp_cfconv_message_bwd(dfilt)");
  }
  if (dxprime) {
    CMP_REQUIRE(rowptr_t && col_t && eid_t && filt && dist, CMP_EINVAL,
                "cmp_cfconv_message_bwd: null pointer (dxprime inputs)");
    cfconv_dx_kernel<<<blocks, 256, 0, as_stream(stream)>>>(g, filt, dist, rowptr_t, col_t, eid_t, N, F, cutoff,
                                                            dxprime, nullptr);
    CMP_LAUNCH_CHECK("cmp_cfconv_message_bwd(dx)");
  }
  return CMP_OK;
}

extern "C" int cmp_segment_sum_fwd(const float* x, const int32_t* seg_ptr, int64_t G, int C, float* out,
                                   cmp_stream_t stream) {
  CMP_REQUIRE(G >= 0 && C >= 1, CMP_EINVAL, "cmp_segment_sum_fwd: bad size");
  if (G == 0) return CMP_OK;
  CMP_REQUIRE(seg_ptr && out, CMP_EINVAL, "cmp_segment_sum_fwd: null pointer");
  int threads = C < 256 ? ((C + 31) / 32) * 32 : 256;
  segment_sum_fwd_kernel<<<(unsigned)G, threads, 0, as_stream(stream)>>>(x, seg_ptr, C, out);
  CMP_LAUNCH_CHECK("cmp_segment_sum_fwd");
  return CMP_OK;
}

extern "C" int cmp_segment_sum_bwd(const float* dout, const int32_t* seg_ptr, int64_t G, int C, float* dx,
                                   cmp_stream_t stream) {
  CMP_REQUIRE(G >= 0 && C >= 1, CMP_EINVAL, "cmp_segment_sum_bwd: bad size");
  if (G == 0) return CMP_OK;
  CMP_REQUIRE(dout && seg_ptr, CMP_EINVAL, "cmp_segment_sum_bwd: null pointer");
  segment_sum_bwd_kernel<<<(unsigned)G, 256, 0, as_stream(stream)>>>(dout, seg_ptr, C, dx);
  CMP_LAUNCH_CHECK("cmp_segment_sum_bwd");
  return CMP_OK;
}

// Same gather-multiply-reduce with an explicit per-edge scale instead of the SchNet cosine cutoff:
//   out[i] = sum_{e in row i} x[col[e]] * filt[e] * scale[e]
// (ViSNet NeighborEmbedding, tgv.py:408-423: scale = masked cosine cutoff, 0 on self loops.)
extern "C" int cmp_edge_message_fwd(const float* x, const float* filt, const float* scale, const int32_t* rowptr,
                                    const int32_t* col, int64_t N, int F, float* out, cmp_stream_t stream) {
  CMP_REQUIRE(N >= 0 && F >= 1, CMP_EINVAL, "cmp_edge_message_fwd: bad size");
  if (N == 0) return CMP_OK;
  CMP_REQUIRE(x && rowptr && out && scale, CMP_EINVAL, "cmp_edge_message_fwd: null pointer");
  unsigned blocks = (unsigned)ceil_div(N * 32, 256);
  cfconv_message_fwd_kernel<1><<<blocks, 256, 0, as_stream(stream)>>>(x, filt, nullptr, rowptr, col, N, F, 1.0f, out,
                                                                      scale);
  CMP_LAUNCH_CHECK("cmp_edge_message_fwd");
  return CMP_OK;
}

extern "C" int cmp_edge_message_bwd(const float* g, const float* x, const float* filt, const float* scale,
                                    const int32_t* rowptr, const int32_t* col, const int32_t* rowptr_t,
                                    const int32_t* col_t, const int32_t* eid_t, int64_t N, int F, float* dfilt, float* dx,
                                    cmp_stream_t stream) {
  CMP_REQUIRE(N >= 0 && F >= 1, CMP_EINVAL, "cmp_edge_message_bwd: bad size");
  if (N == 0) return CMP_OK;
  CMP_REQUIRE(g && rowptr && scale, CMP_EINVAL, "cmp_edge_message_bwd: null pointer");
  unsigned blocks = (unsigned)ceil_div(N * 32, 256);
  if (dfilt) {
    CMP_REQUIRE(x && col, CMP_EINVAL, "cmp_edge_message_bwd: null pointer (dfilt inputs)");
    cfconv_dfilt_kernel<<<blocks, 256, 0, as_stream(stream)>>>(g, x, nullptr, rowptr, col, N, F, 1.0f, dfilt, scale);
    CMP_LAUNCH_CHECK("cmp_edge_message_bwd(dfilt)");
  }
  if (dx) {
    CMP_REQUIRE(rowptr_t && col_t && eid_t && filt, CMP_EINVAL, "cmp_edge_message_bwd: null pointer (dx inputs)");
    cfconv_dx_kernel<<<blocks, 256, 0, as_stream(stream)>>>(g, filt, nullptr, rowptr_t, col_t, eid_t, N, F, 1.0f, dx,
                                                            scale);
    CMP_LAUNCH_CHECK("cmp_edge_message_bwd(dx)");
  }
  return CMP_OK;
}

extern "C" int cmp_regression_head_max_channels(void) { return cmp::HEAD_MAXC; }

extern "C" int cmp_regression_head_fwd(const float* emb, int64_t ld, int64_t B, int K, int C, const float* w,
                                       const float* bias, const float* target, float* err, float* loss,
                                       cmp_stream_t stream) {
  CMP_REQUIRE(B >= 1 && K >= 1 && C >= 1 && B <= (1 << 20) && C <= cmp::HEAD_MAXC, CMP_EINVAL,
              "cmp_regression_head_fwd: bad size");
  CMP_REQUIRE(emb && w && target && err && loss, CMP_EINVAL, "cmp_regression_head_fwd: null pointer");
  cmp::regression_head_fwd_kernel<<<1, cmp::HEAD_THREADS, 0, cmp::as_stream(stream)>>>(emb, ld, (int)B, K, C, w, bias,
                                                                                    target, err, loss);
  CMP_LAUNCH_CHECK("cmp_regression_head_fwd");
  return CMP_OK;
}

extern "C" int cmp_regression_head_bwd(const float* emb, int64_t ld, int64_t B, int K, int C, const float* w,
                                       const float* err, const float* gscale, float* d_emb, int64_t ldd, float* dw,
                                       float* db, cmp_stream_t stream) {
  CMP_REQUIRE(B >= 1 && K >= 1 && C >= 1 && B <= (1 << 20) && C <= cmp::HEAD_MAXC, CMP_EINVAL,
              "cmp_regression_head_bwd: bad size");
  CMP_REQUIRE(emb && w && err && d_emb && dw, CMP_EINVAL, "cmp_regression_head_bwd: null pointer");
  const size_t smem = (size_t)cmp::HEAD_WARPS * (C + 1) * sizeof(float);
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(cmp::regression_head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
          cudaSuccess) {
    (void)cudaGetLastError();
    cmp::set_error("cmp_regression_head_bwd: cannot opt in to %zu bytes of shared memory", smem);
    return CMP_ECUDA;
  }
  cmp::regression_head_bwd_kernel<<<1, cmp::HEAD_THREADS, smem, cmp::as_stream(stream)>>>(emb, ld, (int)B, K, C, w, err,
                                                                                       gscale, d_emb, ldd, dw, db);
  CMP_LAUNCH_CHECK("cmp_regression_head_bwd");
  return CMP_OK;
}
