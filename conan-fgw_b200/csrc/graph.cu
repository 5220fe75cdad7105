// Neighbour-list kernels: sorted batch -> conformer segments -> destination-sorted CSR
// (+ source-sorted transpose).  One CTA per conformer; a conformer's coordinates live in
// shared memory while its n*n pair tests run.  Integer outputs are bit-exact against
// oracle/radius.py (torch-cluster CUDA truncation rule).
#include "common.cuh"

namespace cmp {
namespace {

constexpr int kGraphThreads = 128;
constexpr int kSmemAtoms = 2048;  // conformers up to this size are staged in shared memory

__global__ void batch_to_segments_kernel(const int64_t* __restrict__ batch, int64_t N, int64_t G,
                                         int32_t* __restrict__ seg_ptr, int* status) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (N == 0) {
    if (i <= G) seg_ptr[i] = 0;
    return;
  }
  if (i >= N) return;
  int64_t b = batch[i];
  int64_t prev = (i == 0) ? -1 : batch[i - 1];
  if (b < prev || b < 0) atomicOr(status, CMP_STATUS_UNSORTED_BATCH);
  if (b >= G) {
    atomicOr(status, CMP_STATUS_UNSORTED_BATCH);
    b = G - 1;
  }
  for (int64_t g = prev + 1; g <= b; ++g)
    if (g >= 0 && g < G) seg_ptr[g] = (int32_t)i;
  if (i == N - 1)
    for (int64_t g = b + 1; g <= G; ++g) seg_ptr[g] = (int32_t)N;
}

// d2 exactly as the oracle: each product and each sum individually rounded.
__device__ __forceinline__ float dist2_nofma(float ax, float ay, float az, float bx, float by, float bz) {
  float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

struct PosView {
  const float* x;
  const float* y;
  const float* z;
  int stride;  // 1 for shared SoA, 3 for global AoS
  __device__ __forceinline__ void get(int j, float& px, float& py, float& pz) const {
    px = x[j * stride];
    py = y[j * stride];
    pz = z[j * stride];
  }
};

__device__ __forceinline__ PosView stage_positions(const float* __restrict__ pos, int s, int n, float* smem) {
  PosView v;
  if (n <= kSmemAtoms) {
    for (int t = threadIdx.x; t < n * 3; t += blockDim.x) {
      int a = t / 3, c = t - a * 3;
      smem[c * kSmemAtoms + a] = pos[(int64_t)s * 3 + t];
    }
    __syncthreads();
    v.x = smem;
    v.y = smem + kSmemAtoms;
    v.z = smem + 2 * kSmemAtoms;
    v.stride = 1;
  } else {
    v.x = pos + (int64_t)s * 3;
    v.y = v.x + 1;
    v.z = v.x + 2;
    v.stride = 3;
  }
  return v;
}

// pass 1: out-degree (as a target) of each atom + per-conformer edge totals
__global__ void __launch_bounds__(kGraphThreads)
radius_count_kernel(const float* __restrict__ pos, const int32_t* __restrict__ seg_ptr, float r2, int cap,
                    int loop, int32_t* __restrict__ deg, int32_t* __restrict__ conf_edges) {
  __shared__ float spos[3 * kSmemAtoms];
  __shared__ int warp_sums[kGraphThreads / 32];
  const int g = blockIdx.x;
  const int s = seg_ptr[g], e = seg_ptr[g + 1], n = e - s;
  PosView P = stage_positions(pos, s, n, spos);
  int local = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float xi, yi, zi;
    P.get(i, xi, yi, zi);
    int count = 0, d = 0;
    for (int j = 0; j < n; ++j) {
      float xj, yj, zj;
      P.get(j, xj, yj, zj);
      if (dist2_nofma(xj, yj, zj, xi, yi, zi) < r2) {
        if (loop || j != i) ++d;
        if (++count >= cap) break;
      }
    }
    deg[s + i] = d;
    local += d;
  }
  for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < kGraphThreads / 32; ++w) t += warp_sums[w];
    conf_edges[g] = t;
  }
}

// pass 2: exclusive scan of per-conformer totals (single CTA, any G)
__global__ void __launch_bounds__(1024)
scan_conformers_kernel(const int32_t* __restrict__ conf_edges, int64_t G, int32_t* __restrict__ conf_edge_ptr) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int64_t base = 0; base < G; base += 1024) {
    int64_t i = base + threadIdx.x;
    int v = (i < G) ? conf_edges[i] : 0;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    if (w == 0) {
      int t = warp_tot[lane];
      int ti = t;
      for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, ti, o);
        if (lane >= o) ti += u;
      }
      warp_tot[lane] = ti - t;  // exclusive offset of each warp
    }
    __syncthreads();
    int carry = carry_s;
    if (i < G) conf_edge_ptr[i] = carry + warp_tot[w] + inc - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + warp_tot[w] + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) conf_edge_ptr[G] = carry_s;
}

// pass 3: fill col/dist(/evec) in CSR order, then the source-sorted transpose
__global__ void __launch_bounds__(kGraphThreads)
radius_fill_kernel(const float* __restrict__ pos, const int32_t* __restrict__ seg_ptr, float r2, int cap, int loop,
                   int64_t cap_E, int64_t N, int64_t G, const int32_t* __restrict__ deg,
                   const int32_t* __restrict__ conf_edge_ptr, int32_t* __restrict__ rowptr,
                   int32_t* __restrict__ col, float* __restrict__ dist, float* __restrict__ evec,
                   int32_t* __restrict__ rowptr_t, int32_t* __restrict__ col_t, int32_t* __restrict__ eid_t,
                   int32_t* __restrict__ tcount, int* status) {
  __shared__ float spos[3 * kSmemAtoms];
  const int g = blockIdx.x;
  const int s = seg_ptr[g], e = seg_ptr[g + 1], n = e - s;
  const int ebase = conf_edge_ptr[g];
  PosView P = stage_positions(pos, s, n, spos);

  // row offsets: warp 0 scans the degrees of this conformer
  if (threadIdx.x < 32) {
    int carry = ebase;
    for (int base = 0; base < n; base += 32) {
      int i = base + threadIdx.x;
      int v = (i < n) ? deg[s + i] : 0;
      int inc = v;
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if ((int)threadIdx.x >= o) inc += t;
      }
      if (i < n) rowptr[s + i] = carry + inc - v;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (g == G - 1 && threadIdx.x == 0) rowptr[N] = conf_edge_ptr[G];
  }
  __syncthreads();
  if ((int64_t)conf_edge_ptr[g + 1] > cap_E) {
    if (threadIdx.x == 0) atomicOr(status, CMP_STATUS_EDGE_OVERFLOW);
    return;  // capacity exceeded: this conformer's edges are not written
  }

  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float xi, yi, zi;
    P.get(i, xi, yi, zi);
    int count = 0;
    int w = rowptr[s + i];
    for (int j = 0; j < n; ++j) {
      float xj, yj, zj;
      P.get(j, xj, yj, zj);
      if (dist2_nofma(xj, yj, zj, xi, yi, zi) < r2) {
        if (loop || j != i) {
          col[w] = s + j;
          float dx = xj - xi, dy = yj - yi, dz = zj - zi;
          dist[w] = (j == i) ? 0.0f : sqrtf(dx * dx + dy * dy + dz * dz);
          if (evec) {
            evec[3 * (int64_t)w + 0] = dx;
            evec[3 * (int64_t)w + 1] = dy;
            evec[3 * (int64_t)w + 2] = dz;
          }
          ++w;
        }
        if (++count >= cap) break;
      }
    }
  }
  if (!rowptr_t) return;
  __syncthreads();  // this CTA's col[] rows are now visible to all of its threads

  // transpose: for source j, the rows i (ascending) whose sorted col list contains j
  auto find_in_row = [&](int i, int j) -> int {
    int lo = rowptr[s + i];
    int hi = (i + 1 < n) ? rowptr[s + i + 1] : conf_edge_ptr[g + 1];
    int target = s + j;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      int c = col[mid];
      if (c == target) return mid;
      if (c < target) lo = mid + 1; else hi = mid;
    }
    return -1;
  };
  // A conformer with at most `cap` atoms cannot have been truncated, so its neighbour list is symmetric: the transposed
  // row of j holds the same atoms as row j, and only the edge ids need a search - one thread per EDGE instead of a
  // serial sweep over all (i, j) per atom.
  if (n <= cap) {
    const int eend = conf_edge_ptr[g + 1];
    for (int i = threadIdx.x; i < n; i += blockDim.x) rowptr_t[s + i] = rowptr[s + i];
    if (g == G - 1 && threadIdx.x == 0) rowptr_t[N] = conf_edge_ptr[G];
    for (int e = ebase + threadIdx.x; e < eend; e += blockDim.x) {
      // row j of edge e: the last row whose offset is <= e
      int lo = 0, hi = n - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (rowptr[s + mid] <= e) lo = mid; else hi = mid - 1;
      }
      const int j = lo, i = col[e] - s;
      col_t[e] = s + i;                    // target of the transposed edge j -> i
      eid_t[e] = find_in_row(i, j);        // ... which is stored in row i of the target-sorted list
    }
    return;
  }
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    int c = 0;
    for (int i = 0; i < n; ++i) c += (find_in_row(i, j) >= 0);
    tcount[s + j] = c;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    int carry = ebase;
    for (int base = 0; base < n; base += 32) {
      int j = base + threadIdx.x;
      int v = (j < n) ? tcount[s + j] : 0;
      int inc = v;
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if ((int)threadIdx.x >= o) inc += t;
      }
      if (j < n) rowptr_t[s + j] = carry + inc - v;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (g == G - 1 && threadIdx.x == 0) rowptr_t[N] = conf_edge_ptr[G];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    int w = rowptr_t[s + j];
    for (int i = 0; i < n; ++i) {
      int p = find_in_row(i, j);
      if (p >= 0) {
        col_t[w] = s + i;
        eid_t[w] = p;
        ++w;
      }
    }
  }
}

__global__ void csr_to_edge_index_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                         int64_t N, int64_t E, int64_t* __restrict__ ei) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int b = rowptr[i], e = rowptr[i + 1];
  for (int k = b; k < e && k < E; ++k) {
    ei[k] = col[k];
    ei[E + k] = i;
  }
}

// ---- edge tiles: runs of whole target rows of ONE conformer holding <= tile_edges edges ----------
// (the unit of work of the fused tensor-core kernels: a tile is one UMMA N-extent)
// descriptor = 2 x int4: {first_row, end_row, conformer_first_atom, conformer_atoms}, {first_edge, num_edges, partial, 0}
// A row with more than tile_edges edges (only possible in the source-sorted transpose of a conformer above 128 atoms:
// a low-index atom is a source for every in-range target) is cut into chunks of tile_edges edges; each chunk is its own
// tile with partial = 1 and the kernel ADDS its sum to the (zero-filled) output row instead of storing it.
__device__ __forceinline__ int walk_tiles(const int32_t* __restrict__ rowptr, int s, int e, int tile_edges,
                                          int4* __restrict__ out, int* status) {
  (void)status;
  int count = 0, first = s, cur = 0;
  for (int r = s; r < e; ++r) {
    const int d = rowptr[r + 1] - rowptr[r];
    if (d > tile_edges) {
      if (cur > 0) {
        if (out) {
          out[2 * count] = make_int4(first, r, s, e - s);
          out[2 * count + 1] = make_int4(rowptr[first], rowptr[r] - rowptr[first], 0, 0);
        }
        ++count;
      }
      for (int k0 = 0; k0 < d; k0 += tile_edges) {
        if (out) {
          out[2 * count] = make_int4(r, r + 1, s, e - s);
          out[2 * count + 1] = make_int4(rowptr[r] + k0, min(tile_edges, d - k0), 1, 0);
        }
        ++count;
      }
      first = r + 1;
      cur = 0;
      continue;
    }
    if (cur + d > tile_edges && cur > 0) {
      if (out) {
        out[2 * count] = make_int4(first, r, s, e - s);
        out[2 * count + 1] = make_int4(rowptr[first], rowptr[r] - rowptr[first], 0, 0);
      }
      ++count;
      first = r;
      cur = 0;
    }
    cur += d;
  }
  if (cur > 0) {
    if (out) {
      out[2 * count] = make_int4(first, e, s, e - s);
      out[2 * count + 1] = make_int4(rowptr[first], rowptr[e] - rowptr[first], 0, 0);
    }
    ++count;
  }
  return count;
}

// conformers with fewer than min_atoms atoms get no tiles (they belong to the pair kernel, cfconv_pair.cu)
__global__ void tiles_count_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ seg_ptr, int64_t G,
                                   int tile_edges, int min_atoms, int32_t* __restrict__ counts, int* status) {
  int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const int s = seg_ptr[g], e = seg_ptr[g + 1];
  counts[g] = (e - s < min_atoms) ? 0 : walk_tiles(rowptr, s, e, tile_edges, nullptr, status);
}

__global__ void tiles_fill_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ seg_ptr, int64_t G,
                                  int tile_edges, int min_atoms, const int32_t* __restrict__ tile_ptr, int64_t cap,
                                  int4* __restrict__ tiles, int32_t* __restrict__ num_tiles, int* status) {
  int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g == 0) *num_tiles = (tile_ptr[G] <= cap) ? tile_ptr[G] : 0;
  if (g >= G) return;
  if ((int64_t)tile_ptr[G] > cap) {
    if (g == 0) atomicOr(status, CMP_STATUS_EDGE_OVERFLOW);
    return;
  }
  if (seg_ptr[g + 1] - seg_ptr[g] < min_atoms) return;
  walk_tiles(rowptr, seg_ptr[g], seg_ptr[g + 1], tile_edges, tiles + 2 * (int64_t)tile_ptr[g], status);
}

// ---- primary edges: one representative per undirected pair -----------------------------------------
// The filter of an edge depends on the distance only, so (j -> i) and (i -> j) share it.  Edge (j -> i) (row i,
// column j) is PRIMARY when j > i, or when j < i and the reverse edge is missing (the neighbour cap may keep one
// direction only); rev = 1 when a primary edge's reverse exists.  Order: rows ascending, columns ascending.
__device__ __forceinline__ bool csr_has(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int row,
                                        int target) {
  int lo = rowptr[row], hi = rowptr[row + 1];
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const int c = col[mid];
    if (c == target) return true;
    if (c < target) lo = mid + 1; else hi = mid;
  }
  return false;
}

__global__ void __launch_bounds__(kGraphThreads)
pair_count_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                  const int32_t* __restrict__ seg_ptr, int sym_atoms, int32_t* __restrict__ pcount,
                  int32_t* __restrict__ conf_pairs) {
  __shared__ int warp_sums[kGraphThreads / 32];
  const int g = blockIdx.x;
  const int s = seg_ptr[g], n = seg_ptr[g + 1] - s;
  const bool sym = n <= sym_atoms;    // too small to have been truncated: every edge has its reverse
  int local = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int row = s + i;
    int c = 0;
    for (int k = rowptr[row]; k < rowptr[row + 1]; ++k) {
      const int j = col[k];
      if (j >= row || (!sym && !csr_has(rowptr, col, j, row))) ++c;   // a self loop is its own (unpaired) representative
    }
    pcount[row] = c;
    local += c;
  }
  for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < kGraphThreads / 32; ++w) t += warp_sums[w];
    conf_pairs[g] = t;
  }
}

__global__ void __launch_bounds__(kGraphThreads)
pair_fill_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ dist,
                 const int32_t* __restrict__ seg_ptr, int sym_atoms, const int32_t* __restrict__ pcount,
                 const int32_t* __restrict__ conf_pair_ptr, int64_t cap_P, int32_t* __restrict__ p_src,
                 int32_t* __restrict__ p_dst, float* __restrict__ p_dist, int32_t* __restrict__ p_rev, int* status) {
  const int g = blockIdx.x;
  const int s = seg_ptr[g], n = seg_ptr[g + 1] - s;
  if ((int64_t)conf_pair_ptr[g + 1] > cap_P) {
    if (threadIdx.x == 0) atomicOr(status, CMP_STATUS_EDGE_OVERFLOW);
    return;
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int row = s + i;
    int w = conf_pair_ptr[g];
    for (int r = 0; r < i; ++r) w += pcount[s + r];
    for (int k = rowptr[row]; k < rowptr[row + 1]; ++k) {
      const int j = col[k];
      const bool has = (j != row) && (n <= sym_atoms || csr_has(rowptr, col, j, row));
      if (j >= row || !has) {
        p_src[w] = j;
        p_dst[w] = row;
        p_dist[w] = dist[k];
        p_rev[w] = has ? 1 : 0;
        ++w;
      }
    }
  }
}

// ---- dense views for the FGW input preparation (to_dense_batch / to_dense_adj of PyG) ------------------------
// out[g, a, :] = x[seg_ptr[g] + a, :] (fill beyond the conformer), mask[g, a] = a < atoms(g)
__global__ void dense_batch_kernel(const float* __restrict__ x, const int32_t* __restrict__ seg_ptr, int64_t G,
                                   int64_t n_max, int C, float fill, float* __restrict__ out,
                                   uint8_t* __restrict__ mask) {
  const int64_t row = blockIdx.x;            // g * n_max + a
  const int64_t g = row / n_max, a = row - g * n_max;
  const int s = seg_ptr[g], n = seg_ptr[g + 1] - s;
  const bool live = a < n;
  if (threadIdx.x == 0 && mask) mask[row] = live ? 1 : 0;
  const float* src = x + (int64_t)(s + a) * C;
  float* dst = out + row * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) dst[c] = live ? src[c] : fill;
}

// adj[g, src - s, dst - s] += 1 for every edge whose two ends lie inside conformer g and below n_max
// (integer-valued sums: the order of the atomic adds does not change the result)
__global__ void dense_adj_kernel(const int64_t* __restrict__ ei_src, const int64_t* __restrict__ ei_dst, int64_t E,
                                 const int64_t* __restrict__ batch, const int32_t* __restrict__ seg_ptr, int64_t G,
                                 int64_t n_max, float* __restrict__ adj) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t j = ei_src[e], i = ei_dst[e];
  const int64_t g = batch ? batch[j] : 0;
  if (g < 0 || g >= G) return;
  const int64_t s = seg_ptr[g];
  const int64_t a = j - s, b = i - s;
  if (a < 0 || b < 0 || a >= n_max || b >= n_max) return;
  atomicAdd(adj + (g * n_max + a) * n_max + b, 1.0f);
}

__global__ void gather_f32_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx,
                                  const int32_t* __restrict__ count_ptr, float* __restrict__ dst) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t n = *count_ptr;
  for (; i < n; i += stride) dst[i] = src[idx[i]];
}

__global__ void set_i32_kernel(int32_t* p, int64_t n, int32_t v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// the caller sized its [E, *] tensors by a promised edge count: report a geometry that breaks the promise
__global__ void check_edge_count_kernel(const int32_t* __restrict__ rowptr, int64_t N, int64_t expected, int* status) {
  if (threadIdx.x == 0 && blockIdx.x == 0 && (int64_t)rowptr[N] != expected) atomicOr(status, CMP_STATUS_EDGE_OVERFLOW);
}

}  // namespace
}  // namespace cmp

using namespace cmp;

extern "C" int cmp_batch_to_segments(const int64_t* batch, int64_t N, int64_t G, int32_t* seg_ptr, int* status,
                                     cmp_stream_t stream) {
  CMP_REQUIRE(N >= 0 && G >= 0, CMP_EINVAL, "cmp_batch_to_segments: negative size");
  CMP_REQUIRE(seg_ptr && status && (batch || N == 0), CMP_EINVAL, "cmp_batch_to_segments: null pointer");
  CMP_REQUIRE(N < (int64_t)1 << 31, CMP_EUNSUPPORTED, "cmp_batch_to_segments: N must fit int32");
  int64_t work = (N == 0) ? G + 1 : N;
  int blocks = (int)ceil_div(work, 256);
  batch_to_segments_kernel<<<blocks, 256, 0, as_stream(stream)>>>(batch, N, G, seg_ptr, status);
  CMP_LAUNCH_CHECK("cmp_batch_to_segments");
  return CMP_OK;
}

extern "C" size_t cmp_radius_csr_workspace(int64_t N, int64_t G) {
  // deg[N] + tcount[N] + conf_edges[G]
  return align_up((size_t)(2 * N + G + 8) * sizeof(int32_t), 256);
}

extern "C" int cmp_radius_csr(const float* pos, const int32_t* seg_ptr, int64_t N, int64_t G, double r,
                              int max_num_neighbors, int loop, int64_t cap_E, int32_t* rowptr, int32_t* col,
                              float* dist, float* evec, int32_t* rowptr_t, int32_t* col_t, int32_t* eid_t,
                              int32_t* conf_edge_ptr, void* workspace, size_t workspace_bytes, int* status,
                              cmp_stream_t stream) {
  CMP_REQUIRE(N >= 0 && G >= 0 && cap_E >= 0, CMP_EINVAL, "cmp_radius_csr: negative size");
  CMP_REQUIRE(seg_ptr && rowptr && conf_edge_ptr && status, CMP_EINVAL, "cmp_radius_csr: null pointer");
  CMP_REQUIRE((pos && col && dist) || N == 0, CMP_EINVAL, "cmp_radius_csr: null pointer");
  CMP_REQUIRE((rowptr_t != nullptr) == (col_t != nullptr) && (col_t != nullptr) == (eid_t != nullptr), CMP_EINVAL,
              "cmp_radius_csr: rowptr_t/col_t/eid_t must be given together");
  CMP_REQUIRE(max_num_neighbors >= 1, CMP_EINVAL, "cmp_radius_csr: max_num_neighbors must be >= 1");
  CMP_REQUIRE(workspace_bytes >= cmp_radius_csr_workspace(N, G) && workspace, CMP_EWORKSPACE,
              "cmp_radius_csr: workspace too small (%zu < %zu)", workspace_bytes, cmp_radius_csr_workspace(N, G));
  CMP_REQUIRE(N * (int64_t)(max_num_neighbors + 1) < ((int64_t)1 << 31), CMP_EUNSUPPORTED,
              "cmp_radius_csr: edge count must fit int32");
  cudaStream_t st = as_stream(stream);
  if (N == 0 || G == 0) {
    set_i32_kernel<<<1, 32, 0, st>>>(rowptr, 1, 0);
    if (rowptr_t) set_i32_kernel<<<1, 32, 0, st>>>(rowptr_t, 1, 0);
    set_i32_kernel<<<(int)ceil_div(G + 1, 256), 256, 0, st>>>(conf_edge_ptr, G + 1, 0);
    CMP_LAUNCH_CHECK("cmp_radius_csr(empty)");
    return CMP_OK;
  }
  int32_t* deg = reinterpret_cast<int32_t*>(workspace);
  int32_t* tcount = deg + N;
  int32_t* conf_edges = tcount + N;
  const float r2 = (float)(r * r);
  const int cap = loop ? max_num_neighbors : max_num_neighbors + 1;
  radius_count_kernel<<<(unsigned)G, kGraphThreads, 0, st>>>(pos, seg_ptr, r2, cap, loop, deg, conf_edges);
  CMP_LAUNCH_CHECK("cmp_radius_csr(count)");
  scan_conformers_kernel<<<1, 1024, 0, st>>>(conf_edges, G, conf_edge_ptr);
  CMP_LAUNCH_CHECK("cmp_radius_csr(scan)");
  radius_fill_kernel<<<(unsigned)G, kGraphThreads, 0, st>>>(pos, seg_ptr, r2, cap, loop, cap_E, N, G, deg,
                                                            conf_edge_ptr, rowptr, col, dist, evec, rowptr_t,
                                                            col_t, eid_t, tcount, status);
  CMP_LAUNCH_CHECK("cmp_radius_csr(fill)");
  return CMP_OK;
}

extern "C" int cmp_csr_to_edge_index(const int32_t* rowptr, const int32_t* col, int64_t N, int64_t E,
                                     int64_t* edge_index, cmp_stream_t stream) {
  CMP_REQUIRE(N >= 0 && E >= 0, CMP_EINVAL, "cmp_csr_to_edge_index: negative size");
  if (N == 0 || E == 0) return CMP_OK;
  CMP_REQUIRE(rowptr && col && edge_index, CMP_EINVAL, "cmp_csr_to_edge_index: null pointer");
  csr_to_edge_index_kernel<<<(int)ceil_div(N, 256), 256, 0, as_stream(stream)>>>(rowptr, col, N, E, edge_index);
  CMP_LAUNCH_CHECK("cmp_csr_to_edge_index");
  return CMP_OK;
}

extern "C" size_t cmp_build_tiles_workspace(int64_t G) {
  return align_up((size_t)(2 * G + 8) * sizeof(int32_t), 256);  // counts[G] + tile_ptr[G+1]
}

extern "C" int cmp_build_tiles_min_atoms(const int32_t* rowptr, const int32_t* seg_ptr, int64_t G, int tile_edges,
                                         int min_atoms, void* tiles, int64_t cap_tiles, int32_t* num_tiles,
                                         void* workspace, size_t workspace_bytes, int* status, cmp_stream_t stream);

extern "C" int cmp_build_tiles(const int32_t* rowptr, const int32_t* seg_ptr, int64_t G, int tile_edges, void* tiles,
                               int64_t cap_tiles, int32_t* num_tiles, void* workspace, size_t workspace_bytes,
                               int* status, cmp_stream_t stream) {
  return cmp_build_tiles_min_atoms(rowptr, seg_ptr, G, tile_edges, 0, tiles, cap_tiles, num_tiles, workspace,
                                   workspace_bytes, status, stream);
}

extern "C" int cmp_build_tiles_min_atoms(const int32_t* rowptr, const int32_t* seg_ptr, int64_t G, int tile_edges,
                                         int min_atoms, void* tiles, int64_t cap_tiles, int32_t* num_tiles,
                                         void* workspace, size_t workspace_bytes, int* status, cmp_stream_t stream) {
  CMP_REQUIRE(G >= 0 && tile_edges >= 16 && cap_tiles >= 0, CMP_EINVAL, "cmp_build_tiles: bad size");
  CMP_REQUIRE(num_tiles && status, CMP_EINVAL, "cmp_build_tiles: null pointer");
  cudaStream_t st = as_stream(stream);
  if (G == 0) {
    set_i32_kernel<<<1, 32, 0, st>>>(num_tiles, 1, 0);
    CMP_LAUNCH_CHECK("cmp_build_tiles(empty)");
    return CMP_OK;
  }
  CMP_REQUIRE(rowptr && seg_ptr && tiles, CMP_EINVAL, "cmp_build_tiles: null pointer");
  CMP_REQUIRE(workspace && workspace_bytes >= cmp_build_tiles_workspace(G), CMP_EWORKSPACE,
              "cmp_build_tiles: workspace too small");
  int32_t* counts = reinterpret_cast<int32_t*>(workspace);
  int32_t* tile_ptr = counts + G;
  tiles_count_kernel<<<(unsigned)ceil_div(G, 128), 128, 0, st>>>(rowptr, seg_ptr, G, tile_edges, min_atoms, counts,
                                                                status);
  CMP_LAUNCH_CHECK("cmp_build_tiles(count)");
  scan_conformers_kernel<<<1, 1024, 0, st>>>(counts, G, tile_ptr);
  CMP_LAUNCH_CHECK("cmp_build_tiles(scan)");
  tiles_fill_kernel<<<(unsigned)ceil_div(G, 128), 128, 0, st>>>(rowptr, seg_ptr, G, tile_edges, min_atoms, tile_ptr,
                                                               cap_tiles,
                                                               reinterpret_cast<int4*>(tiles), num_tiles, status);
  CMP_LAUNCH_CHECK("cmp_build_tiles(fill)");
  return CMP_OK;
}

extern "C" size_t cmp_build_pair_list_workspace(int64_t N, int64_t G) {
  return align_up((size_t)(N + G + 8) * sizeof(int32_t), 256);   // pcount[N] + conf_pairs[G]
}

extern "C" int cmp_build_pair_list(const int32_t* rowptr, const int32_t* col, const float* dist, const int32_t* seg_ptr,
                                   int64_t N, int64_t G, int sym_atoms, int64_t cap_P, int32_t* p_src, int32_t* p_dst,
                                   float* p_dist,
                                   int32_t* p_rev, int32_t* conf_pair_ptr, void* workspace, size_t workspace_bytes,
                                   int* status, cmp_stream_t stream) {
  CMP_REQUIRE(N >= 0 && G >= 0 && cap_P >= 0, CMP_EINVAL, "cmp_build_pair_list: negative size");
  CMP_REQUIRE(conf_pair_ptr && status, CMP_EINVAL, "cmp_build_pair_list: null pointer");
  cudaStream_t st = as_stream(stream);
  if (N == 0 || G == 0) {
    set_i32_kernel<<<(int)ceil_div(G + 1, 256), 256, 0, st>>>(conf_pair_ptr, G + 1, 0);
    CMP_LAUNCH_CHECK("cmp_build_pair_list(empty)");
    return CMP_OK;
  }
  CMP_REQUIRE(rowptr && col && dist && seg_ptr && p_src && p_dst && p_dist && p_rev, CMP_EINVAL,
              "cmp_build_pair_list: null pointer");
  CMP_REQUIRE(workspace && workspace_bytes >= cmp_build_pair_list_workspace(N, G), CMP_EWORKSPACE,
              "cmp_build_pair_list: workspace too small");
  int32_t* pcount = reinterpret_cast<int32_t*>(workspace);
  int32_t* conf_pairs = pcount + N;
  pair_count_kernel<<<(unsigned)G, kGraphThreads, 0, st>>>(rowptr, col, seg_ptr, sym_atoms, pcount, conf_pairs);
  CMP_LAUNCH_CHECK("cmp_build_pair_list(count)");
  scan_conformers_kernel<<<1, 1024, 0, st>>>(conf_pairs, G, conf_pair_ptr);
  CMP_LAUNCH_CHECK("cmp_build_pair_list(scan)");
  pair_fill_kernel<<<(unsigned)G, kGraphThreads, 0, st>>>(rowptr, col, dist, seg_ptr, sym_atoms, pcount, conf_pair_ptr,
                                                          cap_P, p_src,
                                                          p_dst, p_dist, p_rev, status);
  CMP_LAUNCH_CHECK("cmp_build_pair_list(fill)");
  return CMP_OK;
}

extern "C" int cmp_dense_batch(const float* x, const int32_t* seg_ptr, int64_t G, int64_t n_max, int C, float fill,
                               float* out, uint8_t* mask, cmp_stream_t stream) {
  CMP_REQUIRE(G >= 0 && n_max >= 0 && C >= 0, CMP_EINVAL, "cmp_dense_batch: negative size");
  if (G == 0 || n_max == 0) return CMP_OK;
  CMP_REQUIRE(seg_ptr && out && (x || C == 0), CMP_EINVAL, "cmp_dense_batch: null pointer");
  CMP_REQUIRE(G * n_max < ((int64_t)1 << 31), CMP_EUNSUPPORTED, "cmp_dense_batch: G * n_max must fit int32");
  dense_batch_kernel<<<(unsigned)(G * n_max), 128, 0, as_stream(stream)>>>(x, seg_ptr, G, n_max, C, fill, out, mask);
  CMP_LAUNCH_CHECK("cmp_dense_batch");
  return CMP_OK;
}

extern "C" int cmp_dense_adj(const int64_t* edge_index, int64_t E, const int64_t* batch, const int32_t* seg_ptr,
                             int64_t G, int64_t n_max, float* adj, cmp_stream_t stream) {
  CMP_REQUIRE(E >= 0 && G >= 0 && n_max >= 0, CMP_EINVAL, "cmp_dense_adj: negative size");
  if (G == 0 || n_max == 0) return CMP_OK;
  CMP_REQUIRE(seg_ptr && adj && (edge_index || E == 0), CMP_EINVAL, "cmp_dense_adj: null pointer");
  cudaStream_t st = as_stream(stream);
  CMP_REQUIRE(cudaMemsetAsync(adj, 0, (size_t)G * n_max * n_max * sizeof(float), st) == cudaSuccess, CMP_ECUDA,
              "cmp_dense_adj: memset failed");
  if (E == 0) return CMP_OK;
  dense_adj_kernel<<<(unsigned)ceil_div(E, 256), 256, 0, st>>>(edge_index, edge_index + E, E, batch, seg_ptr, G, n_max,
                                                                 adj);
  CMP_LAUNCH_CHECK("cmp_dense_adj");
  return CMP_OK;
}

extern "C" int cmp_gather_f32(const float* src, const int32_t* idx, const int32_t* count_ptr, int64_t max_count,
                              float* dst, cmp_stream_t stream) {
  CMP_REQUIRE(max_count >= 0, CMP_EINVAL, "cmp_gather_f32: negative size");
  if (max_count == 0) return CMP_OK;
  CMP_REQUIRE(src && idx && count_ptr && dst, CMP_EINVAL, "cmp_gather_f32: null pointer");
  int64_t blocks = ceil_div(max_count, 256);
  if (blocks > 4096) blocks = 4096;
  gather_f32_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(src, idx, count_ptr, dst);
  CMP_LAUNCH_CHECK("cmp_gather_f32");
  return CMP_OK;
}

extern "C" int cmp_check_edge_count(const int32_t* rowptr, int64_t N, int64_t expected_edges, int* status,
                                    cmp_stream_t stream) {
  CMP_REQUIRE(N >= 0 && expected_edges >= 0, CMP_EINVAL, "cmp_check_edge_count: negative size");
  CMP_REQUIRE(rowptr && status, CMP_EINVAL, "cmp_check_edge_count: null pointer");
  check_edge_count_kernel<<<1, 32, 0, as_stream(stream)>>>(rowptr, N, expected_edges, status);
  CMP_LAUNCH_CHECK("cmp_check_edge_count");
  return CMP_OK;
}
