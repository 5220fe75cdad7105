// Dense fp32 building blocks: tiled SIMT GEMM with fused epilogue + deterministic split-K,
// column sums, activations, Adam.  These are the exact-fp32 ("parity") node/edge GEMMs and the
// generic-shape path; the fused tcgen05 kernels (cfconv_tc.cu) take over for the hot shapes.
#include <stdarg.h>

#include <atomic>

#include <stdlib.h>

#include "common.cuh"

namespace cmp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

bool pdl_enabled() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("CMP_NO_PDL");
    cached = (e && e[0] && e[0] != '0') ? 0 : 1;
  }
  return cached != 0;
}

int sm_count() {
  static int cached = 0;
  if (cached) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) == cudaSuccess &&
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) {
    cached = n;
  } else {
    (void)cudaGetLastError();
    return 148;  // B200; not cached so a later call with a device present can correct it
  }
  return cached;
}

namespace {

constexpr int BM = 128, BN = 64, BK = 16, GEMM_THREADS = 256;

template <bool TA, bool TB>
__global__ void __launch_bounds__(GEMM_THREADS)
gemm_f32_kernel(int M, int N, int K, const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
                int64_t ldb, float* __restrict__ C, int64_t ldc, const float* __restrict__ bias, int act,
                const float* __restrict__ residual, int64_t ldr, float* __restrict__ partial, int kchunk) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * kchunk;
  const int kend = min(K, kbeg + kchunk);

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- stage A tile (BM x BK) ----
#pragma unroll
    for (int r = 0; r < (BM * BK) / GEMM_THREADS; ++r) {
      int e = tid + r * GEMM_THREADS;
      int m, k;
      if (TA) { m = e & (BM - 1); k = e >> 7; } else { k = e & (BK - 1); m = e >> 4; }
      int gm = m0 + m, gk = k0 + k;
      float v = 0.0f;
      if (gm < M && gk < kend) v = TA ? A[(int64_t)gk * lda + gm] : A[(int64_t)gm * lda + gk];
      As[k][m] = v;
    }
    // ---- stage B tile (BK x BN) ----
#pragma unroll
    for (int r = 0; r < (BN * BK) / GEMM_THREADS; ++r) {
      int e = tid + r * GEMM_THREADS;
      int n, k;
      if (TB) { k = e & (BK - 1); n = e >> 4; } else { n = e & (BN - 1); k = e >> 6; }
      int gn = n0 + n, gk = k0 + k;
      float v = 0.0f;
      if (gn < N && gk < kend) v = TB ? B[(int64_t)gn * ldb + gk] : B[(int64_t)gk * ldb + gn];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }

  if (gridDim.z > 1) {
    float* P = partial + (int64_t)blockIdx.z * M * N;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int gm = m0 + ty * 8 + i;
      if (gm >= M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int gn = n0 + tx * 4 + j;
        if (gn < N) P[(int64_t)gm * N + gn] = acc[i][j];
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int gm = m0 + ty * 8 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[gn];
      v = apply_act(v, act);
      if (residual) v += residual[(int64_t)gm * ldr + gn];
      C[(int64_t)gm * ldc + gn] = v;
    }
  }
}

// Vectorised variant of gemm_f32_kernel for 16-byte aligned operands (every node linear and its gradients): the tiles
// are fetched as float4 along the contiguous dimension, and the global loads of tile k + 1 are issued into registers
// before the FMAs of tile k, so their latency hides behind the arithmetic (the scalar kernel above exposes it in every
// 16-wide k step).  Same tiling, same per-thread accumulation order over k: results equal the scalar kernel's bit for bit.
// Needs lda % 4 == ldb % 4 == 0, K % 4 == 0, 16-byte aligned A / B, and M % 4 == 0 (TA) / N % 4 == 0 (!TB).
template <bool TA, bool TB>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_f32_vec_kernel(int M, int N, int K, const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
                    int64_t ldb, float* __restrict__ C, int64_t ldc, const float* __restrict__ bias, int act,
                    const float* __restrict__ residual, int64_t ldr, float* __restrict__ partial, int kchunk) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * kchunk;
  const int kend = min(K, kbeg + kchunk);

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  float4 ra[2], rb;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto fetch = [&](int k0) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int e = tid + r * GEMM_THREADS;
      ra[r] = zero4;
      if (TA) {
        const int m = (e & 31) * 4, k = e >> 5;
        if (m0 + m < M && k0 + k < kend) ra[r] = __ldg(reinterpret_cast<const float4*>(A + (int64_t)(k0 + k) * lda + m0 + m));
      } else {
        const int k = (e & 3) * 4, m = e >> 2;
        if (m0 + m < M && k0 + k < kend) ra[r] = __ldg(reinterpret_cast<const float4*>(A + (int64_t)(m0 + m) * lda + k0 + k));
      }
    }
    rb = zero4;
    if (TB) {
      const int k = (tid & 3) * 4, n = tid >> 2;
      if (n0 + n < N && k0 + k < kend) rb = __ldg(reinterpret_cast<const float4*>(B + (int64_t)(n0 + n) * ldb + k0 + k));
    } else {
      const int n = (tid & 15) * 4, k = tid >> 4;
      if (n0 + n < N && k0 + k < kend) rb = __ldg(reinterpret_cast<const float4*>(B + (int64_t)(k0 + k) * ldb + n0 + n));
    }
  };
  auto stage = [&]() {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int e = tid + r * GEMM_THREADS;
      if (TA) {
        const int m = (e & 31) * 4, k = e >> 5;
        *reinterpret_cast<float4*>(&As[k][m]) = ra[r];
      } else {
        const int k = (e & 3) * 4, m = e >> 2;
        As[k + 0][m] = ra[r].x; As[k + 1][m] = ra[r].y; As[k + 2][m] = ra[r].z; As[k + 3][m] = ra[r].w;
      }
    }
    if (TB) {
      const int k = (tid & 3) * 4, n = tid >> 2;
      Bs[k + 0][n] = rb.x; Bs[k + 1][n] = rb.y; Bs[k + 2][n] = rb.z; Bs[k + 3][n] = rb.w;
    } else {
      const int n = (tid & 15) * 4, k = tid >> 4;
      *reinterpret_cast<float4*>(&Bs[k][n]) = rb;
    }
  };

  if (kbeg < kend) fetch(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    stage();
    __syncthreads();
    if (k0 + BK < kend) fetch(k0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }

  const int gn0 = n0 + tx * 4;
  if (gridDim.z > 1) {
    float* P = partial + (int64_t)blockIdx.z * M * N;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int gm = m0 + ty * 8 + i;
      if (gm >= M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (gn0 + j < N) P[(int64_t)gm * N + gn0 + j] = acc[i][j];
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + ty * 8 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = gn0 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[gn];
      v = apply_act(v, act);
      if (residual) v += residual[(int64_t)gm * ldr + gn];
      C[(int64_t)gm * ldc + gn] = v;
    }
  }
}

// Tiny problems (the regression head of a training step: [128, 64] x [64, 1] and its two gradient GEMMs): one thread per
// output element walks K with eight loads in flight - the tiled kernels above spend 10-20 us on 16-wide k steps with two
// block barriers each for a few thousand multiply-adds.  Same ascending-k summation order per output.
__global__ void __launch_bounds__(256)
gemm_f32_small_kernel(int M, int N, int K, const float* __restrict__ A, int64_t sam, int64_t sak,
                      const float* __restrict__ B, int64_t sbk, int64_t sbn, float* __restrict__ C, int64_t ldc,
                      const float* __restrict__ bias, int act, const float* __restrict__ residual, int64_t ldr) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * N) return;
  const int m = idx / N, n = idx - m * N;
  const float* a = A + (int64_t)m * sam;
  const float* b = B + (int64_t)n * sbn;
  float acc = 0.0f;
  int k = 0;
  for (; k + 8 <= K; k += 8) {
    float av[8], bv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      av[u] = __ldg(a + (int64_t)(k + u) * sak);
      bv[u] = __ldg(b + (int64_t)(k + u) * sbk);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) acc = fmaf(av[u], bv[u], acc);
  }
  for (; k < K; ++k) acc = fmaf(__ldg(a + (int64_t)k * sak), __ldg(b + (int64_t)k * sbk), acc);
  if (bias) acc += bias[n];
  acc = apply_act(acc, act);
  if (residual) acc += residual[(int64_t)m * ldr + n];
  C[(int64_t)m * ldc + n] = acc;
}

__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int S, int64_t M, int64_t N,
                                     float* __restrict__ C, int64_t ldc, const float* __restrict__ bias, int act,
                                     const float* __restrict__ residual, int64_t ldr) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * N) return;
  int64_t m = idx / N, n = idx - m * N;
  float v = 0.0f;
  for (int s = 0; s < S; ++s) v += partial[(int64_t)s * M * N + idx];  // fixed order: deterministic
  if (bias) v += bias[n];
  v = apply_act(v, act);
  if (residual) v += residual[m * ldr + n];
  C[m * ldc + n] = v;
}

// M * N % 4 == 0, N % 4 == 0 and 16-byte aligned rows: 16 lanes per float4 of the output - lane l adds partials l, l + 16,
// ... in order, the 16 lane sums are then added in lane order (fixed order: deterministic; the scalar kernel above walks
// all S partials in one dependent chain per thread, 18 us for the node-linear weight gradients of cfg 2)
__global__ void __launch_bounds__(256)
splitk_reduce_vec4_kernel(const float4* __restrict__ partial, int S, int64_t total4, int N4, float* __restrict__ C,
                          int64_t ldc, const float* __restrict__ bias, int act, const float* __restrict__ residual,
                          int64_t ldr) {
  __shared__ float4 part[16][16];
  const int f = threadIdx.x & 15, l = threadIdx.x >> 4;
  const int64_t t4 = (int64_t)blockIdx.x * 16 + f;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t4 < total4) {
#pragma unroll 4
    for (int c = l; c < S; c += 16) {
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
    const int64_t m = t4 / N4;
    const int n = (int)(t4 - m * N4) * 4;
    float o[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (bias) o[j] += bias[n + j];
      o[j] = apply_act(o[j], act);
      if (residual) o[j] += residual[m * ldr + n + j];
    }
    *reinterpret_cast<float4*>(C + m * ldc + n) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

struct SplitPlan {
  int S;
  int kchunk;
};

SplitPlan plan_split(int64_t M, int64_t N, int64_t K) {
  int64_t tiles = ceil_div(M, BM) * ceil_div(N, BN);
  int sms = sm_count();
  SplitPlan p{1, (int)(ceil_div(K, BK) * BK)};
  if (tiles >= sms || K <= 512) return p;
  int64_t want = ceil_div(2 * (int64_t)sms, tiles);
  int64_t maxS = ceil_div(K, 256);
  int64_t S = want < maxS ? want : maxS;
  if (S <= 1) return p;
  int64_t kchunk = ceil_div(ceil_div(K, S), BK) * BK;
  p.kchunk = (int)kchunk;
  p.S = (int)ceil_div(K, kchunk);
  return p;
}

// ---- column sums ------------------------------------------------------------------------
constexpr int CS_ROWS = 8;  // blockDim.y

__global__ void colsum_stage1_kernel(const float* __restrict__ X, int64_t ldx, int64_t M, int N, int rows_per_chunk,
                                     float* __restrict__ partial) {
  __shared__ float red[CS_ROWS][33];
  int n = blockIdx.x * 32 + threadIdx.x;
  int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
  int64_t r1 = r0 + rows_per_chunk;
  if (r1 > M) r1 = M;
  float s = 0.0f;
  if (n < N)
    for (int64_t r = r0 + threadIdx.y; r < r1; r += CS_ROWS) s += X[r * ldx + n];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float t = 0.0f;
#pragma unroll
    for (int y = 0; y < CS_ROWS; ++y) t += red[y][threadIdx.x];
    partial[(int64_t)blockIdx.y * N + n] = t;
  }
}

// N % 4 == 0, ldx % 4 == 0, 16-byte aligned X: a warp reads 512 contiguous bytes of a row, 8 rows in flight per thread
__global__ void __launch_bounds__(32 * CS_ROWS)
colsum_stage1_vec4_kernel(const float* __restrict__ X, int64_t ldx, int64_t M, int N4, int rows_per_chunk,
                          float4* __restrict__ partial) {
  __shared__ float4 red[CS_ROWS][32];
  const int n4 = blockIdx.x * 32 + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
  const int64_t r1 = (r0 + rows_per_chunk > M) ? M : r0 + rows_per_chunk;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n4 < N4) {
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8 * CS_ROWS) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int64_t rr = r + (int64_t)u * CS_ROWS;
        v[u] = (rr < r1) ? __ldg(reinterpret_cast<const float4*>(X + rr * ldx) + n4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
    }
  }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && n4 < N4) {
    float4 t = red[0][threadIdx.x];
#pragma unroll
    for (int y = 1; y < CS_ROWS; ++y) {
      const float4 v = red[y][threadIdx.x];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    partial[(int64_t)blockIdx.y * N4 + n4] = t;
  }
}

// 8 lanes per column: lane l adds chunks l, l + 8, ... in order, then the lane sums in lane order
__global__ void __launch_bounds__(256)
colsum_stage2_lanes_kernel(const float* __restrict__ partial, int chunks, int N, float* __restrict__ out) {
  __shared__ float part[8][32];
  const int f = threadIdx.x & 31, l = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + f;
  float t = 0.0f;
  if (n < N)
#pragma unroll 4
    for (int c = l; c < chunks; c += 8) t += __ldg(partial + (int64_t)c * N + n);
  part[l][f] = t;
  __syncthreads();
  if (l == 0 && n < N) {
    float r = part[0][f];
#pragma unroll
    for (int k = 1; k < 8; ++k) r += part[k][f];
    out[n] = r;
  }
}

int colsum_chunks(int64_t M) {
  int64_t c = ceil_div(M, 128);
  if (c < 1) c = 1;
  if (c > 1024) c = 1024;
  return (int)c;
}

// ---- elementwise --------------------------------------------------------------------------
__global__ void act_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, int act) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) y[i] = apply_act(x[i], act);
}

__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ saved, float* __restrict__ dx,
                               int64_t n, int act) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float g = dy[i];
    if (act == CMP_ACT_SSP) {
      // saved = y = softplus(x) - ln2  =>  sigmoid(x) = 1 - exp(-(y + ln2)) = 1 - 0.5 * exp(-y)
      g *= 1.0f - 0.5f * expf(-saved[i]);
    } else if (act == CMP_ACT_SILU) {
      g *= silu_grad(saved[i]);
    }
    dx[i] = g;
  }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps, float wd,
                            float bc1, float bc2_sqrt, float grad_scale) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float gi = g[i] * grad_scale;
    float pi = p[i];
    if (wd != 0.0f) gi += wd * pi;
    float mi = b1 * m[i] + (1.0f - b1) * gi;
    float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

int ew_blocks(int64_t n) {
  int64_t b = ceil_div(n, 256);
  int64_t cap = (int64_t)sm_count() * 16;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace
}  // namespace cmp

using namespace cmp;

extern "C" const char* cmp_last_error_string(void) { return g_err; }
extern "C" int cmp_version(void) { return 100; }
extern "C" long long cmp_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" void cmp_launch_count_reset(void) { g_launches.store(0, std::memory_order_relaxed); }

extern "C" int cmp_device_is_sm100(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return major == 10;
}

extern "C" size_t cmp_gemm_workspace(int64_t M, int64_t N, int64_t K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  SplitPlan p = plan_split(M, N, K);
  return p.S > 1 ? align_up((size_t)p.S * M * N * sizeof(float), 256) : 0;
}

extern "C" int cmp_gemm_f32(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                            const float* B, int64_t ldb, float* C, int64_t ldc, const float* bias, int act,
                            const float* residual, int64_t ldr, void* workspace, size_t workspace_bytes,
                            cmp_stream_t stream) {
  CMP_REQUIRE(M >= 0 && N >= 0 && K >= 0, CMP_EINVAL, "cmp_gemm_f32: negative size");
  if (M == 0 || N == 0) return CMP_OK;
  CMP_REQUIRE(M < ((int64_t)1 << 31) && N < ((int64_t)1 << 31) && K < ((int64_t)1 << 31), CMP_EUNSUPPORTED,
              "cmp_gemm_f32: sizes must fit int32");
  CMP_REQUIRE(C && (K == 0 || (A && B)), CMP_EINVAL, "cmp_gemm_f32: null pointer");
  CMP_REQUIRE(!(transA && transB), CMP_EUNSUPPORTED, "cmp_gemm_f32: transA && transB is not provided");
  CMP_REQUIRE(act >= CMP_ACT_NONE && act <= CMP_ACT_SILU, CMP_EINVAL, "cmp_gemm_f32: unknown activation %d", act);
  cudaStream_t st0 = as_stream(stream);
  if (M * N <= 16384 && K <= 1024) {
    // op(A)[m, k] = A[m * sam + k * sak], op(B)[k, n] = B[k * sbk + n * sbn]
    const int64_t sam = transA ? 1 : lda, sak = transA ? lda : 1;
    const int64_t sbk = transB ? 1 : ldb, sbn = transB ? ldb : 1;
    gemm_f32_small_kernel<<<(unsigned)ceil_div(M * N, 256), 256, 0, st0>>>((int)M, (int)N, (int)K, A, sam, sak, B, sbk, sbn, C,
                                                                           ldc, bias, act, residual, ldr);
    CMP_LAUNCH_CHECK("cmp_gemm_f32(small)");
    return CMP_OK;
  }
  SplitPlan p = (K > 0) ? plan_split(M, N, K) : SplitPlan{1, BK};
  float* partial = nullptr;
  if (p.S > 1) {
    size_t need = (size_t)p.S * M * N * sizeof(float);
    CMP_REQUIRE(workspace && workspace_bytes >= need, CMP_EWORKSPACE, "cmp_gemm_f32: workspace too small (%zu < %zu)",
                workspace_bytes, need);
    partial = reinterpret_cast<float*>(workspace);
  }
  dim3 grid((unsigned)ceil_div(N, BN), (unsigned)ceil_div(M, BM), (unsigned)p.S);
  CMP_REQUIRE(grid.y <= 65535u * 1024u, CMP_EUNSUPPORTED, "cmp_gemm_f32: M too large");
  cudaStream_t st = as_stream(stream);
  // grid.y limit is 65535: fold very tall problems by looping the launch over row bands
  const int64_t band_rows = (int64_t)65535 * BM;
  for (int64_t mb = 0; mb < M; mb += band_rows) {
    int64_t Mb = (M - mb < band_rows) ? (M - mb) : band_rows;
    dim3 g((unsigned)ceil_div(N, BN), (unsigned)ceil_div(Mb, BM), (unsigned)p.S);
    const float* Ab = transA ? A + mb : A + mb * lda;
    float* Cb = C + mb * ldc;
    const float* Rb = residual ? residual + mb * ldr : nullptr;
    float* Pb = partial;  // split-K is only planned for short problems (single band)
    static const bool force_scalar = [] { const char* e = getenv("CMP_GEMM_SCALAR"); return e && e[0] && e[0] != '0'; }();
    const bool vec = !force_scalar && (lda % 4 == 0) && (ldb % 4 == 0) && (K % 4 == 0) && ((reinterpret_cast<uintptr_t>(Ab) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && (transA ? (Mb % 4 == 0 && mb % 4 == 0) : true) &&
                     (transB ? true : N % 4 == 0);
    if (vec && !transA && transB)
      gemm_f32_vec_kernel<false, true><<<g, GEMM_THREADS, 0, st>>>((int)Mb, (int)N, (int)K, Ab, lda, B, ldb, Cb, ldc,
                                                                  bias, act, Rb, ldr, Pb, p.kchunk);
    else if (vec && !transA && !transB)
      gemm_f32_vec_kernel<false, false><<<g, GEMM_THREADS, 0, st>>>((int)Mb, (int)N, (int)K, Ab, lda, B, ldb, Cb, ldc,
                                                                   bias, act, Rb, ldr, Pb, p.kchunk);
    else if (vec)
      gemm_f32_vec_kernel<true, false><<<g, GEMM_THREADS, 0, st>>>((int)Mb, (int)N, (int)K, Ab, lda, B, ldb, Cb, ldc,
                                                                  bias, act, Rb, ldr, Pb, p.kchunk);
    else if (!transA && transB)
      gemm_f32_kernel<false, true><<<g, GEMM_THREADS, 0, st>>>((int)Mb, (int)N, (int)K, Ab, lda, B, ldb, Cb, ldc, bias,
                                                              act, Rb, ldr, Pb, p.kchunk);
    else if (!transA && !transB)
      gemm_f32_kernel<false, false><<<g, GEMM_THREADS, 0, st>>>((int)Mb, (int)N, (int)K, Ab, lda, B, ldb, Cb, ldc,
                                                               bias, act, Rb, ldr, Pb, p.kchunk);
    else
      gemm_f32_kernel<true, false><<<g, GEMM_THREADS, 0, st>>>((int)Mb, (int)N, (int)K, Ab, lda, B, ldb, Cb, ldc, bias,
                                                              act, Rb, ldr, Pb, p.kchunk);
    CMP_LAUNCH_CHECK("cmp_gemm_f32");
  }
  if (p.S > 1) {
    int64_t total = M * N;
    if (N % 4 == 0 && ldc % 4 == 0 && ((reinterpret_cast<uintptr_t>(C) | reinterpret_cast<uintptr_t>(partial)) & 15) == 0)
      splitk_reduce_vec4_kernel<<<(unsigned)ceil_div(total / 4, 16), 256, 0, st>>>(
          reinterpret_cast<const float4*>(partial), p.S, total / 4, (int)(N / 4), C, ldc, bias, act, residual, ldr);
    else
      splitk_reduce_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(partial, p.S, M, N, C, ldc, bias, act,
                                                                          residual, ldr);
    CMP_LAUNCH_CHECK("cmp_gemm_f32(split-K reduce)");
  }
  return CMP_OK;
}

extern "C" size_t cmp_colsum_workspace(int64_t M, int64_t N) {
  if (M <= 0 || N <= 0) return 0;
  return align_up((size_t)colsum_chunks(M) * N * sizeof(float), 256);
}

extern "C" int cmp_colsum_f32(const float* X, int64_t ldx, int64_t M, int64_t N, float* out, void* workspace,
                              size_t workspace_bytes, cmp_stream_t stream) {
  CMP_REQUIRE(M >= 0 && N >= 0, CMP_EINVAL, "cmp_colsum_f32: negative size");
  if (N == 0) return CMP_OK;
  CMP_REQUIRE(out, CMP_EINVAL, "cmp_colsum_f32: null pointer");
  cudaStream_t st = as_stream(stream);
  if (M == 0) {
    CMP_REQUIRE(cudaMemsetAsync(out, 0, N * sizeof(float), st) == cudaSuccess, CMP_ECUDA, "cmp_colsum_f32: memset");
    return CMP_OK;
  }
  CMP_REQUIRE(X, CMP_EINVAL, "cmp_colsum_f32: null pointer");
  int chunks = colsum_chunks(M);
  CMP_REQUIRE(workspace && workspace_bytes >= (size_t)chunks * N * sizeof(float), CMP_EWORKSPACE,
              "cmp_colsum_f32: workspace too small");
  int rows_per_chunk = (int)ceil_div(M, chunks);
  if (N % 4 == 0 && ldx % 4 == 0 && ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(workspace)) & 15) == 0) {
    dim3 grid((unsigned)ceil_div(N / 4, 32), (unsigned)chunks);
    colsum_stage1_vec4_kernel<<<grid, dim3(32, CS_ROWS), 0, st>>>(X, ldx, M, (int)(N / 4), rows_per_chunk,
                                                                 reinterpret_cast<float4*>(workspace));
  } else {
    dim3 grid((unsigned)ceil_div(N, 32), (unsigned)chunks);
    colsum_stage1_kernel<<<grid, dim3(32, CS_ROWS), 0, st>>>(X, ldx, M, (int)N, rows_per_chunk,
                                                            reinterpret_cast<float*>(workspace));
  }
  CMP_LAUNCH_CHECK("cmp_colsum_f32(stage1)");
  colsum_stage2_lanes_kernel<<<(unsigned)ceil_div(N, 32), 256, 0, st>>>(reinterpret_cast<float*>(workspace), chunks,
                                                                       (int)N, out);
  CMP_LAUNCH_CHECK("cmp_colsum_f32(stage2)");
  return CMP_OK;
}

extern "C" int cmp_act_fwd(const float* x, float* y, int64_t n, int act, cmp_stream_t stream) {
  CMP_REQUIRE(n >= 0, CMP_EINVAL, "cmp_act_fwd: negative size");
  if (n == 0) return CMP_OK;
  CMP_REQUIRE(x && y, CMP_EINVAL, "cmp_act_fwd: null pointer");
  CMP_REQUIRE(act >= CMP_ACT_NONE && act <= CMP_ACT_SILU, CMP_EINVAL, "cmp_act_fwd: unknown activation %d", act);
  act_fwd_kernel<<<ew_blocks(n), 256, 0, as_stream(stream)>>>(x, y, n, act);
  CMP_LAUNCH_CHECK("cmp_act_fwd");
  return CMP_OK;
}

extern "C" int cmp_act_bwd(const float* dy, const float* saved, float* dx, int64_t n, int act, cmp_stream_t stream) {
  CMP_REQUIRE(n >= 0, CMP_EINVAL, "cmp_act_bwd: negative size");
  if (n == 0) return CMP_OK;
  CMP_REQUIRE(dy && dx && (saved || act == CMP_ACT_NONE), CMP_EINVAL, "cmp_act_bwd: null pointer");
  CMP_REQUIRE(act >= CMP_ACT_NONE && act <= CMP_ACT_SILU, CMP_EINVAL, "cmp_act_bwd: unknown activation %d", act);
  act_bwd_kernel<<<ew_blocks(n), 256, 0, as_stream(stream)>>>(dy, saved, dx, n, act);
  CMP_LAUNCH_CHECK("cmp_act_bwd");
  return CMP_OK;
}

extern "C" int cmp_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                             cmp_stream_t stream) {
  CMP_REQUIRE(n >= 0 && step >= 1, CMP_EINVAL, "cmp_adam_step: bad size/step");
  if (n == 0) return CMP_OK;
  CMP_REQUIRE(param && grad && exp_avg && exp_avg_sq, CMP_EINVAL, "cmp_adam_step: null pointer");
  float bc1 = 1.0f - powf(beta1, (float)step);
  float bc2 = sqrtf(1.0f - powf(beta2, (float)step));
  adam_kernel<<<ew_blocks(n), 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                          weight_decay, bc1, bc2, grad_scale);
  CMP_LAUNCH_CHECK("cmp_adam_step");
  return CMP_OK;
}
