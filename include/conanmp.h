/*
 * conanmp.h - C ABI of libconanmp.so, the sm_100a kernels behind the ConAN
 * message-passing backbone (SchNet / ViSNet + radius graph).
 *
 * The reference (duyhominhnguyen/conan-fgw) has no native boundary of its own:
 * its hot path is a chain of Python calls into torch-geometric 2.3.0 /
 * torch-cluster 1.6.1 / ATen.  Each entry point below therefore cites the
 * reference-side Python call it replaces (paths relative to the reference root,
 * "sns.py" = conan_fgw/src/model/graph_embeddings/schnet_no_sum.py,
 * "tgv.py" = conan_fgw/src/model/graph_embeddings/torch_geometric_visnet.py).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *     its name ends in _host.  The caller (PyTorch) owns all memory; the library
 *     never allocates, frees or retains device memory.
 *   - every function enqueues on `stream` and returns immediately: 0 on success,
 *     a negative CMP_E* code on a rejected call (see cmp_last_error_string()).
 *     No call synchronises the device, so all of them are CUDA-graph capturable.
 *   - data-dependent errors that only the device can see (unsorted batch,
 *     atomic number out of range) are reported by setting bits in a caller
 *     provided `int* status` word (CMP_STATUS_*), checked by the host lazily.
 *   - row-major dense matrices with an explicit leading dimension (elements).
 *   - there is no CPU fallback anywhere in this library.
 */
#ifndef CONANMP_H_
#define CONANMP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* cmp_stream_t; /* a cudaStream_t */

/* return codes */
#define CMP_OK 0
#define CMP_EINVAL (-1)     /* bad size / null pointer / misaligned pointer */
#define CMP_EUNSUPPORTED (-2) /* shape not supported by this kernel family */
#define CMP_EWORKSPACE (-3) /* workspace too small */
#define CMP_ECUDA (-4)      /* a CUDA runtime call failed (launch configuration ...) */

/* device-side status bits */
#define CMP_STATUS_UNSORTED_BATCH 1
#define CMP_STATUS_BAD_ATOMIC_NUMBER 2
#define CMP_STATUS_EDGE_OVERFLOW 4

/* activation selector of the dense epilogues */
#define CMP_ACT_NONE 0
#define CMP_ACT_SSP 1   /* shifted softplus: softplus(x) - ln 2  (PyG ShiftedSoftplus, sns.py:179) */
#define CMP_ACT_SILU 2  /* x * sigmoid(x)                         (tgv.py:518-519) */

/* numerics of the fused tensor-core kernels */
#define CMP_PREC_FP32 0  /* split-precision tensor-core GEMMs, fp32-grade (parity mode) */
#define CMP_PREC_BF16 1  /* single-pass bf16 operands, fp32 accumulate (throughput mode) */

const char* cmp_last_error_string(void);
int cmp_version(void);
/* number of kernels this library has launched in this process (bench.py reports it) */
long long cmp_launch_count(void);
void cmp_launch_count_reset(void);
/* 1 when the running device is compute capability 10.x (tcgen05 kernels usable) */
int cmp_device_is_sm100(void);

/* ------------------------------------------------------------------------- *
 * Neighbour lists
 * ------------------------------------------------------------------------- */

/* seg_ptr[g] = first atom of conformer g (g = 0..G), from a sorted batch vector.
 * Replaces the bucketize() inside torch_cluster.radius; sets
 * CMP_STATUS_UNSORTED_BATCH when batch decreases anywhere. */
int cmp_batch_to_segments(const int64_t* batch, int64_t N, int64_t G, int32_t* seg_ptr,
                          int* status, cmp_stream_t stream);

/* Destination-sorted CSR radius graph per conformer.
 * Replaces radius_graph()/RadiusInteractionGraph.forward (sns.py:160,208,342;
 * tgv.py:331-347; conan_fgw/src/model/graph_embeddings/visnet.py:90,276) with
 * torch-cluster's CUDA truncation rule: for target i keep the first `cap`
 * candidates j (ascending, self included) with
 *   ((dx*dx)+(dy*dy))+(dz*dz) < (float)((double)r*r)      (fp32, no FMA),
 * cap = max_num_neighbors + (loop ? 0 : 1); self pairs dropped when !loop.
 *
 * Three launches, no host sync:
 *   rowptr int32[N+1]   CSR offsets by target atom (rowptr[N] = E)
 *   col    int32[cap_E] source atom of each edge, ascending inside a row
 *   dist   float[cap_E] |pos[src]-pos[dst]|  (0 on self loops)
 *   evec   float[cap_E*3] or NULL: pos[src]-pos[dst]          (ViSNet Distance)
 *   rowptr_t/col_t/eid_t (each may be NULL together): the same edges grouped by
 *     SOURCE atom (col_t = target, ascending; eid_t = index into col/dist) - the
 *     adjacency transpose the backward pass gathers through.
 *   conf_edge_ptr int32[G+1]: first edge of each conformer.
 * cap_E = capacity of col/dist; N*cap always suffices.  Overflow sets
 * CMP_STATUS_EDGE_OVERFLOW and truncates.
 * workspace: cmp_radius_csr_workspace(N, G) bytes. */
size_t cmp_radius_csr_workspace(int64_t N, int64_t G);
int cmp_radius_csr(const float* pos, const int32_t* seg_ptr, int64_t N, int64_t G, double r,
                   int max_num_neighbors, int loop, int64_t cap_E, int32_t* rowptr, int32_t* col,
                   float* dist, float* evec, int32_t* rowptr_t, int32_t* col_t, int32_t* eid_t,
                   int32_t* conf_edge_ptr, void* workspace, size_t workspace_bytes, int* status,
                   cmp_stream_t stream);

/* Sync-free callers size their [E, *] tensors by an edge count they vouch for (a CUDA-graph replay of the same
 * geometry): sets CMP_STATUS_EDGE_OVERFLOW in *status when rowptr[N] differs from expected_edges. */
int cmp_check_edge_count(const int32_t* rowptr, int64_t N, int64_t expected_edges, int* status,
                         cmp_stream_t stream);

/* CSR -> PyG edge_index int64[2, E] (row 0 = source j, row 1 = target i). */
int cmp_csr_to_edge_index(const int32_t* rowptr, const int32_t* col, int64_t N, int64_t E,
                          int64_t* edge_index, cmp_stream_t stream);

/* ------------------------------------------------------------------------- *
 * Dense building blocks (fp32)
 * ------------------------------------------------------------------------- */

/* C[M,N] = act(opA(A) * opB(B) + bias[N]) + residual[M,N]
 *   transA = 0: A is [M,K] (lda);  1: A is stored [K,M] (lda)
 *   transB = 1: B is stored [N,K] (ldb) - a torch Linear weight; 0: B is [K,N]
 * Replaces every torch.nn.Linear on the path (PyG CFConv.lin1/lin2/nn,
 * InteractionBlock.lin, sns.py:177-179,225-231) and their autograd GEMMs.
 * bias / residual may be NULL.  workspace is used for split-K partial sums:
 * cmp_gemm_workspace(M,N,K) bytes. */
size_t cmp_gemm_workspace(int64_t M, int64_t N, int64_t K);
int cmp_gemm_f32(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A,
                 int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, const float* bias,
                 int act, const float* residual, int64_t ldr, void* workspace,
                 size_t workspace_bytes, cmp_stream_t stream);

/* out[N] = sum_m X[m, :]  (bias gradients); deterministic two-stage reduction.
 * workspace: cmp_colsum_workspace(M, N). */
size_t cmp_colsum_workspace(int64_t M, int64_t N);
int cmp_colsum_f32(const float* X, int64_t ldx, int64_t M, int64_t N, float* out, void* workspace,
                   size_t workspace_bytes, cmp_stream_t stream);

/* y = act(x) elementwise, and dx = dy * act'(.) computed from the forward
 * OUTPUT y for SSP (sigmoid(x) = 1 - exp(-y)/2) or the forward INPUT x for SILU. */
int cmp_act_fwd(const float* x, float* y, int64_t n, int act, cmp_stream_t stream);
int cmp_act_bwd(const float* dy, const float* saved, float* dx, int64_t n, int act,
                cmp_stream_t stream);

/* ------------------------------------------------------------------------- *
 * SchNet pieces
 * ------------------------------------------------------------------------- */

/* GaussianSmearing.forward (PyG; sns.py:161): out[e,k] = exp(coeff*(d[e]-offset[k])^2). */
int cmp_rbf_gaussian_fwd(const float* d, int64_t E, const float* offset, int Ng, float coeff,
                         float* out, int64_t ldo, cmp_stream_t stream);

/* Embedding(100, H, padding_idx=0) forward / weight gradient (sns.py:159).
 * bwd is deterministic; workspace cmp_embedding_bwd_workspace(N, V, H). */
int cmp_embedding_fwd(const int64_t* z, int64_t N, const float* weight, int V, int H, float* out,
                      int* status, cmp_stream_t stream);
size_t cmp_embedding_bwd_workspace(int64_t N, int V, int H);
int cmp_embedding_bwd(const int64_t* z, int64_t N, const float* dout, int V, int H, int padding_idx,
                      float* dweight, void* workspace, size_t workspace_bytes, cmp_stream_t stream);

/* CFConv message + aggregation (PyG CFConv.forward/message):
 *   agg[i,:] = sum_{e in row i} xprime[col[e],:] * filt[e,:] * C(dist[e]),
 *   C(d) = 0.5*(cos(d*pi/cutoff)+1)            (no d<cutoff mask in SchNet)
 * deterministic: edges of a row are summed in CSR order. */
int cmp_cfconv_message_fwd(const float* xprime, const float* filt, const float* dist,
                           const int32_t* rowptr, const int32_t* col, int64_t N, int F,
                           float cutoff, float* agg, cmp_stream_t stream);
/* Backward of the above for upstream gradient g[N,F]:
 *   dfilt[e,:]  = g[dst(e),:] * xprime[col[e],:] * C(dist[e])
 *   dxprime[j,:] = sum_{e' in row_t j} g[col_t[e'],:] * filt[eid_t[e'],:] * C(dist[eid_t[e']]) */
int cmp_cfconv_message_bwd(const float* g, const float* xprime, const float* filt,
                           const float* dist, const int32_t* rowptr, const int32_t* col,
                           const int32_t* rowptr_t, const int32_t* col_t, const int32_t* eid_t,
                           int64_t N, int F, float cutoff, float* dfilt, float* dxprime,
                           cmp_stream_t stream);

/* Regression head of a ConAN training step in two launches instead of ~20 on KB-sized tensors
 * (schnet_based_models.py:242 conformer mean, :17-29 linear head, model/common.py:288 MSE):
 *   mol[b] = mean_k emb[b K + k];  pred[b] = w . mol[b] + bias;  loss = mean_b (pred[b] - target[b])^2
 * fwd: err[b] = pred[b] - target[b] (kept for the backward), loss[0].  bwd: d_emb[B K, C], dw[C], db[1] for an upstream
 * gradient gscale[0] of the loss (device scalar; NULL = 1).  emb has B * K rows (conformers of a molecule consecutive). */
int cmp_regression_head_max_channels(void);   /* C <= this (512) */
int cmp_regression_head_fwd(const float* emb, int64_t ld, int64_t B, int K, int C, const float* w,
                            const float* bias, const float* target, float* err, float* loss,
                            cmp_stream_t stream);
int cmp_regression_head_bwd(const float* emb, int64_t ld, int64_t B, int K, int C, const float* w,
                            const float* err, const float* gscale, float* d_emb, int64_t ldd, float* dw,
                            float* db, cmp_stream_t stream);
/* Sum readout over sorted segments (PyG SumAggregation; sns.py:184,353) and its
 * backward (broadcast of dout[g] to the atoms of g). */
int cmp_segment_sum_fwd(const float* x, const int32_t* seg_ptr, int64_t G, int C, float* out,
                        cmp_stream_t stream);
int cmp_segment_sum_bwd(const float* dout, const int32_t* seg_ptr, int64_t G, int C, float* dx,
                        cmp_stream_t stream);

/* ------------------------------------------------------------------------- *
 * Fused tensor-core CFConv (tcgen05 / TMEM / TMA bulk copies), num_filters = 128
 * ------------------------------------------------------------------------- */

/* Edge tiles = the work units of the fused kernels: runs of whole target rows of ONE conformer
 * holding <= tile_edges edges.  tiles is an array of 8 x int32 descriptors {first_row, end_row,
 * conformer_first_atom, conformer_atom_count, first_edge, num_edges, 0, 0}; *num_tiles receives the
 * count (device memory, no host sync).
 * Works on either orientation of the CSR (rowptr or rowptr_t). */
size_t cmp_build_tiles_workspace(int64_t G);
int cmp_build_tiles(const int32_t* rowptr, const int32_t* seg_ptr, int64_t G, int tile_edges,
                    void* tiles, int64_t cap_tiles, int32_t* num_tiles, void* workspace,
                    size_t workspace_bytes, int* status, cmp_stream_t stream);
/* Same, skipping conformers with fewer than min_atoms atoms (those are served by cmp_cfconv_pair_fwd). */
int cmp_build_tiles_min_atoms(const int32_t* rowptr, const int32_t* seg_ptr, int64_t G, int tile_edges,
                              int min_atoms, void* tiles, int64_t cap_tiles, int32_t* num_tiles,
                              void* workspace, size_t workspace_bytes, int* status, cmp_stream_t stream);

/* Dense views for the FGW input preparation: torch_geometric.utils.to_dense_batch /
 * to_dense_adj as schnet_no_sum.py:242-253 and visnet.py:168-177 call them.
 * cmp_dense_batch: out[g, a, :] = x[seg_ptr[g] + a, :] for a < atoms(g), `fill` beyond; mask[g, a] (uint8, may be NULL).
 * cmp_dense_adj:  adj[g, src - first(g), dst - first(g)] += 1 per edge of edge_index (int64 [2, E], row 0 = source),
 *                 g = batch[src]; adj is zero-filled first; ends at or beyond n_max are dropped. */
int cmp_dense_batch(const float* x, const int32_t* seg_ptr, int64_t G, int64_t n_max, int C, float fill,
                    float* out, uint8_t* mask, cmp_stream_t stream);
int cmp_dense_adj(const int64_t* edge_index, int64_t E, const int64_t* batch, const int32_t* seg_ptr,
                  int64_t G, int64_t n_max, float* adj, cmp_stream_t stream);

/* dst[i] = src[idx[i]] for i < *count_ptr (per-edge data in transposed order, no host sync). */
int cmp_gather_f32(const float* src, const int32_t* idx, const int32_t* count_ptr,
                   int64_t max_count, float* dst, cmp_stream_t stream);

/* 1 when (num_filters, num_gaussians) is served by the fused kernels (128, < 64). */
int cmp_cfconv_tc_supported(int num_filters, int num_gaussians);
size_t cmp_cfconv_tc_weights_bytes(void);
int cmp_cfconv_tc_tile_edges(void);
/* Filter-MLP weights (PyG InteractionBlock.mlp: Linear(Ng,F), ssp, Linear(F,F)) -> bf16 UMMA
 * shared-memory images with the biases folded in as an extra K column. */
int cmp_cfconv_tc_pack_weights(const float* W1, const float* b1, const float* W2, const float* b2,
                               int num_filters, int num_gaussians, void* packed,
                               cmp_stream_t stream);
/* agg[i,:] = sum_{j->i} xprime[j,:] * (W2 ssp(W1 rbf(d_ij) + b1) + b2) * C(d_ij) in ONE kernel:
 * Gaussian expansion, filter MLP (tcgen05, bf16 operands, fp32 accumulate), cosine cutoff,
 * gather-multiply and the per-target reduction (CSR order, deterministic).  Replaces
 * GaussianSmearing + CFConv.nn + CFConv.propagate (sns.py:161-164 through PyG CFConv.forward). */
int cmp_cfconv_fused_fwd(const float* xprime, const float* dist, const int32_t* rowptr,
                         const int32_t* col, const void* tiles, const int32_t* num_tiles,
                         const void* packed_weights, const float* offset, int num_gaussians,
                         float coeff, float cutoff, int64_t N, int num_filters, float* agg,
                         cmp_stream_t stream);

/* The same aggregation for conformers of at most cmp_cfconv_pair_max_atoms() (30) atoms, one filter evaluation per
 * UNDIRECTED pair (cmp_build_pair_list): the filter depends on d_ij only, so j -> i and i -> j share it.  One CTA per
 * conformer at a time (x rows staged in shared memory by a TMA bulk copy, per-pipeline [atoms, F] accumulators in
 * shared memory, summed in a fixed order: deterministic, no atomics).  Rows of larger conformers (and of conformers
 * without edges) are NOT written: run cmp_cfconv_fused_fwd first with tiles from
 * cmp_build_tiles_min_atoms(min_atoms = cmp_cfconv_pair_max_atoms() + 1) - it zero-fills `agg` and serves the large conformers.
 * transposed = 1 exchanges the two directions of every pair (the d x' pass of the backward, x = dL/dagg). */
int cmp_cfconv_pair_max_atoms(void);
int cmp_cfconv_pair_fwd(const float* x, const int32_t* seg_ptr, const int32_t* conf_pair_ptr,
                        const int32_t* pair_src, const int32_t* pair_dst, const float* pair_dist,
                        const int32_t* pair_rev, int64_t G, const void* packed_weights,
                        const float* offset, int num_gaussians, float coeff, float cutoff,
                        int num_filters, int transposed, float* agg, cmp_stream_t stream);

/* The same aggregation over the DENSE 16 x 16 atom blocks of every conformer of at most cmp_cfconv_dense_max_atoms()
 * (128) atoms (cfconv_dense.cu): one filter evaluation per undirected pair, the pair (i, j) of a TMEM column is a
 * compile-time function of the column, and the thread that owns a filter channel applies the column to both directions
 * with x' / agg of the row block held in registers (no shared-memory gathers, no atomics, fixed summation order).
 * Replaces CFConv.message / propagate as called from sns.py:161-164.
 *   adj [N, 4] uint32: bit j of row i = the graph has the (conformer-local) edge j -> i; from cmp_build_adjacency
 *                      (rows of conformers above the atom limit are not written and not read);
 *   pos [N, 3], seg_ptr [G + 1]: as given to cmp_radius_csr;
 *   offset_host: the num_gaussians Gaussian centres in HOST memory (they become kernel parameters);
 *   packed_weights: cmp_cfconv_dense_pack_weights (f16 images, log2 e folded into W1 / b1 and ln 2 into W2);
 *   counter: one int32 of device scratch (work queue head; the call zeroes it on the stream);
 *   skip_large = 1: conformers above the atom limit are left untouched (serve them with cmp_cfconv_fused_fwd on tiles
 *                   from cmp_build_tiles_min_atoms(limit + 1), issued BEFORE this call); 0: they set
 *                   CMP_STATUS_EDGE_OVERFLOW in *status.
 *   max_atoms_hint: an upper bound of the atoms per conformer known to the caller (0 = unknown).  It only selects the
 *                   kernel: up to 32 atoms the warp-specialised tile pipeline (cfconv_dense_ws_kernel: Gaussians + softplus
 *                   epilogue, MMA issue and the two-direction epilogue on separate warp groups, the conformer's rows in
 *                   registers), otherwise one pipeline of 128 threads per conformer.  Results are bit-identical.
 * Every row of `agg` that belongs to a conformer within the limit is written (zeros where an atom has no neighbour).
 * transposed = 1 exchanges the two directions of every pair (the d x' pass of the backward, x = dL/dagg). */
int cmp_cfconv_dense_max_atoms(void);
int cmp_cfconv_dense_supported(int num_filters, int num_gaussians);
size_t cmp_cfconv_dense_weights_bytes(void);
int cmp_build_adjacency(const int32_t* rowptr, const int32_t* col, const int32_t* seg_ptr, int64_t N,
                        int64_t G, uint32_t* adj, cmp_stream_t stream);
int cmp_cfconv_dense_pack_weights(const float* W1, const float* b1, const float* W2, const float* b2,
                                  int num_filters, int num_gaussians, void* packed, cmp_stream_t stream);
/* jobs: array of `count` cmp_dense_pack_job_t in HOST memory (count <= 32) */
typedef struct {
  const float* W1;
  const float* b1;
  const float* W2;
  const float* b2;
  void* packed;
} cmp_dense_pack_job_t;
int cmp_cfconv_dense_pack_weights_grouped(const void* jobs, int count, int num_filters, int num_gaussians,
                                          cmp_stream_t stream);
int cmp_cfconv_dense_fwd(const float* x, const float* pos, const int32_t* seg_ptr, const uint32_t* adj,
                         int64_t G, const void* packed_weights, const float* offset_host,
                         int num_gaussians, float coeff, float cutoff, int num_filters, int transposed,
                         int skip_large, int max_atoms_hint, float* agg, int32_t* counter, int32_t* status,
                         cmp_stream_t stream);

/* fp32-grade variant of cmp_cfconv_dense_fwd ("x3", cfconv_dense_x3_kernel): same tiles, pair <-> column maps and
 * register-resident two-direction epilogue, but every MMA operand (Gaussians, a', W1, W2) is split into f16 hi + lo
 * images and every product runs as three tcgen05 passes (hi hi + lo hi + hi lo, 22 significant bits), the softplus
 * epilogue is evaluated in fp32, distances / cutoffs exactly as the exact kernels compute them.  Meets the 1e-5 parity
 * bar of the exact path (sns.py:161-164 semantics) without materialising any [E, *] tensor.
 *   packed_weights: cmp_cfconv_dense_x3_pack_weights (cmp_cfconv_dense_x3_weights_bytes() bytes: W1 hi | W1 lo | W2 hi |
 *   W2 lo).  Other arguments as cmp_cfconv_dense_fwd (one kernel for every conformer size up to the atom limit). */
size_t cmp_cfconv_dense_x3_weights_bytes(void);
int cmp_cfconv_dense_x3_pack_weights(const float* W1, const float* b1, const float* W2, const float* b2,
                                     int num_filters, int num_gaussians, void* packed, cmp_stream_t stream);
int cmp_cfconv_dense_x3_pack_weights_grouped(const void* jobs, int count, int num_filters, int num_gaussians,
                                             cmp_stream_t stream);
int cmp_cfconv_dense_x3_fwd(const float* x, const float* pos, const int32_t* seg_ptr, const uint32_t* adj,
                            int64_t G, const void* packed_weights, const float* offset_host,
                            int num_gaussians, float coeff, float cutoff, int num_filters, int transposed,
                            int skip_large, float* agg, int32_t* counter, int32_t* status, cmp_stream_t stream);

/* Filter-MLP weight gradients over the DENSE blocks of cmp_cfconv_dense_fwd (conformers of at most 128 atoms; larger
 * ones contribute nothing (cmp_build_dense_bwd_tiles sets CMP_STATUS_EDGE_OVERFLOW) - serve such batches with
 * cmp_cfconv_fused_bwd_weights_pairs): one column per undirected pair in tiles of 64, the pair of a column is a
 * compile-time function of its index, so dF[f, (i, j)] = [j -> i] g[i] x'[j] + [i -> j] g[j] x'[i] is built from fp32 rows
 * of g = dL/dagg and x' held in registers (no pair list, no bf16 copies of g / x', no gathers); distances from `pos`, the
 * directions of a pair from `adj` (cmp_build_adjacency).  Same MMAs, epilogues and TMEM accumulators as
 * cmp_cfconv_fused_bwd_weights; packed_bwd_weights from cmp_cfconv_tc_pack_bwd_weights; offset in DEVICE memory.
 * tile_ptr [G + 1]: tiles of the conformers before g, from cmp_build_dense_bwd_tiles (once per neighbour list; it also
 * raises the status bit for conformers above the limit).  workspace: cmp_cfconv_dense_bwd_workspace() bytes.
 * Replaces the weight-gradient half of CFConv's backward (PyG autograd through CFConv.message / nn, sns.py:161-164). */
size_t cmp_cfconv_dense_bwd_workspace(void);
int cmp_build_dense_bwd_tiles(const int32_t* seg_ptr, int64_t G, int32_t* tile_ptr, int32_t* status,
                              cmp_stream_t stream);
int cmp_cfconv_dense_bwd_weights(const float* g, const float* xprime, const float* pos, const int32_t* seg_ptr,
                                 const uint32_t* adj, const int32_t* tile_ptr, int64_t G,
                                 const void* packed_bwd_weights, const float* offset, int num_gaussians, float coeff,
                                 float cutoff, int num_filters, float* dW1, float* db1, float* dW2, float* db2,
                                 void* workspace, size_t workspace_bytes, cmp_stream_t stream);

/* fp32-grade variant of cmp_cfconv_dense_bwd_weights ("x3", cfconv_dense_bwd_x3_kernel): every operand (Gaussians,
 * W1, W2^T, a', dF, dh) as bf16 hi + lo images, three tcgen05 passes per product, fp32 epilogues, sigmoid recomputed from
 * the TMEM-resident pre-activations.  dW1 / db1 / dW2 / db2 meet the 1e-5 bar of the exact path.
 *   packed_bwd_x3_weights: cmp_cfconv_dense_bwd_x3_pack_weights_grouped (cmp_cfconv_dense_bwd_x3_weights_bytes() bytes:
 *   W1aug hi | lo | W2^T hi | lo); jobs = array of cmp_bwd_x3_pack_job_t in HOST memory (count <= 32).
 * Other arguments, workspace and tile_ptr as cmp_cfconv_dense_bwd_weights. */
typedef struct {
  const float* W1;
  const float* b1;
  const float* W2;
  void* packed;
} cmp_bwd_x3_pack_job_t;
size_t cmp_cfconv_dense_bwd_x3_weights_bytes(void);
int cmp_cfconv_dense_bwd_x3_pack_weights_grouped(const void* jobs, int count, int num_filters, int num_gaussians,
                                                 cmp_stream_t stream);
int cmp_cfconv_dense_bwd_x3_weights(const float* g, const float* xprime, const float* pos, const int32_t* seg_ptr,
                                    const uint32_t* adj, const int32_t* tile_ptr, int64_t G,
                                    const void* packed_bwd_x3_weights, const float* offset, int num_gaussians,
                                    float coeff, float cutoff, int num_filters, float* dW1, float* db1, float* dW2,
                                    float* db2, void* workspace, size_t workspace_bytes, cmp_stream_t stream);

/* Filter-MLP weight gradients of the fused CFConv in ONE kernel (+ a fixed-order reduction of the
 * per-pipeline partial sums): recomputes rbf / hidden / a' per 64-edge tile on chip and accumulates
 * dW2 = sum_e dF_e a'_e^T and dW1 = sum_e dh_e rbf_e^T in TMEM (tcgen05, bf16 operands, fp32
 * accumulate).  g = dL/dagg [N,F]; xprime_bf16 = bf16 copy of x' [N,F]; erow = target row per edge;
 * flat_tiles from cmp_build_flat_tiles.  Replaces autograd through CFConv.nn (PyG) - the GEMMs with
 * K = E that dominate the reference's backward pass.  d x' is cmp_cfconv_fused_fwd over the
 * transposed neighbour list with g as its input. */
int cmp_csr_expand_rows(const int32_t* rowptr, int64_t N, int32_t* erow, cmp_stream_t stream);
int cmp_cfconv_tc_bwd_tile_edges(void);
size_t cmp_build_flat_tiles_workspace(int64_t G);
int cmp_build_flat_tiles(const int32_t* conf_edge_ptr, const int32_t* seg_ptr, const int32_t* erow,
                         int64_t G, int tile_edges, void* tiles, int64_t cap_tiles,
                         int32_t* num_tiles, void* workspace, size_t workspace_bytes, int* status,
                         cmp_stream_t stream);
int cmp_f32_to_bf16(const float* src, int64_t n, void* dst, cmp_stream_t stream);
size_t cmp_cfconv_tc_bwd_weights_bytes(void);
size_t cmp_cfconv_fused_bwd_workspace(void);
int cmp_cfconv_tc_pack_bwd_weights(const float* W1, const float* b1, const float* W2,
                                   int num_filters, int num_gaussians, void* packed,
                                   cmp_stream_t stream);
int cmp_cfconv_fused_bwd_weights(const float* g, const void* xprime_bf16, const float* dist,
                                 const int32_t* col, const int32_t* erow, const void* flat_tiles,
                                 const int32_t* num_tiles, const void* packed_bwd_weights,
                                 const float* offset, int num_gaussians, float coeff, float cutoff,
                                 int num_filters, float* dW1, float* db1, float* dW2, float* db2,
                                 void* workspace, size_t workspace_bytes, cmp_stream_t stream);

/* Pair mode of the same pass.  The filter W(d_ij) of PyG CFConv (schnet.py CFConv.forward: W = nn(edge_attr) * C)
 * depends on the distance only, so the two directions of an undirected pair share it and their weight-gradient
 * terms add: dF = g[dst] x'[src] + rev * g[src] x'[dst].  cmp_build_pair_list emits one PRIMARY edge per pair
 * (j -> i with j > i; or the surviving direction when the neighbour cap of radius_graph dropped the other one,
 * rev = 0; rows and columns ascending) with per-conformer offsets conf_pair_ptr[G+1]; tiles for it come from
 * cmp_build_flat_tiles(conf_pair_ptr, seg_ptr, pair_dst, ...).  Halves the filter-MLP recomputation of the backward.
 * sym_atoms: conformers of at most this many atoms are known to be untruncated (max_num_neighbors + 1 for a graph
 * without self loops), hence symmetric - their reverse-edge searches are skipped; 0 = search always. */
size_t cmp_build_pair_list_workspace(int64_t N, int64_t G);
int cmp_build_pair_list(const int32_t* rowptr, const int32_t* col, const float* dist,
                        const int32_t* seg_ptr, int64_t N, int64_t G, int sym_atoms, int64_t cap_P,
                        int32_t* pair_src,
                        int32_t* pair_dst, float* pair_dist, int32_t* pair_rev, int32_t* conf_pair_ptr,
                        void* workspace, size_t workspace_bytes, int* status, cmp_stream_t stream);
int cmp_cfconv_fused_bwd_weights_pairs(const void* g_bf16, const void* xprime_bf16,
                                       const float* pair_dist, const int32_t* pair_src,
                                       const int32_t* pair_dst, const int32_t* pair_rev,
                                       const void* pair_tiles, const int32_t* num_tiles,
                                       const void* packed_bwd_weights, const float* offset,
                                       int num_gaussians, float coeff, float cutoff, int num_filters,
                                       float* dW1, float* db1, float* dW2, float* db2, void* workspace,
                                       size_t workspace_bytes, cmp_stream_t stream);

/* Node-level linears on tcgen05 with split-bf16 operands (hi + lo images, three MMA passes, fp32
 * accumulate): Y = act(X' W^T + b) + R with X' = X * (1 - exp(-saved_y)/2) when saved_y is given (the
 * ShiftedSoftplus backward fused into the operand load).  K, Nout multiples of 16 in [16, 128].
 * dX is the same kernel with the weight packed transposed; cmp_node_gemm_dw returns dW = dY'^T X
 * and db = column sums of dY' (tile partials reduced in a fixed order).  Replaces torch.nn.Linear and
 * its autograd GEMMs around the message passing (PyG CFConv.lin1/lin2, InteractionBlock.lin, ConAN
 * heads sns.py:177-179,225-231) in the bf16 mode. */
int cmp_node_gemm_tc_supported(int K, int Nout);

/* Up to cmp_node_chain_max_stages() (3) of those linears CHAINED on one 64-atom tile, intermediates never leaving the SM:
 *   V_s = act_s(U_s W_s^T + b_s) * (1 - exp(-scale_y_s) / 2) + R_s,   U_0 = X,  U_(s+1) = V_s.
 * Forward of an interaction-block tail (PyG CFConv.lin2 -> ShiftedSoftplus -> InteractionBlock.lin (+ h) -> the next
 * block's CFConv.lin1; sns.py:163-164) and, with transposed weight images, its backward
 * (dx'' -> lin1'^T (+ dh') -> lin^T * ssp'(y) -> lin2^T).  scale_y is the saved ShiftedSoftplus OUTPUT.
 * stages: array of nstages cmp_chain_stage_t in HOST memory; K of stage s must equal Nout of stage s - 1; `out` may be
 * NULL for a stage whose value only feeds the next one (not for the last). */
typedef struct {
  const void* w_img;     /* cmp_node_gemm_pack_weight image (hi | lo) of W [Nout, K] */
  const float* bias;     /* [Nout] or NULL */
  const float* residual; /* [M, Nout] or NULL */
  const float* scale_y;  /* [M, Nout] or NULL */
  float* out;            /* [M, Nout] or NULL */
  int64_t ldr, lds, ldo;
  int32_t K, Nout, act;
} cmp_chain_stage_t;
int cmp_node_chain_max_stages(void);
int cmp_node_chain_fwd(const float* X, int64_t ldx, int64_t M, const void* stages, int nstages,
                       cmp_stream_t stream);
size_t cmp_node_gemm_weight_bytes(int image_K);
int cmp_node_gemm_pack_weight(const float* W, int rows, int cols, int transpose, void* packed,
                              cmp_stream_t stream);
int cmp_node_gemm_fwd(const float* X, int64_t ldx, const float* saved_y, int64_t ldys,
                      const void* w_img, const float* bias, int act, const float* residual,
                      int64_t ldr, float* Y, int64_t ldy, int64_t M, int K, int Nout,
                      cmp_stream_t stream);
size_t cmp_node_gemm_dw_workspace(int K);
int cmp_node_gemm_dw(const float* dY, int64_t lddy, const float* saved_y, int64_t ldys,
                     const float* X, int64_t ldx, int64_t M, int K, int Nout, float* dW, float* db,
                     void* workspace, size_t workspace_bytes, cmp_stream_t stream);

/* Grouped form: up to cmp_node_gemm_dw_group_max() weight-gradient problems in ONE launch (the node linears of a
 * whole backward pass, issued once at its end).  The SMs are divided between the problems in proportion to their
 * tile counts; every CTA accumulates its tiles in TMEM and writes one partial block; a second kernel sums the
 * partial blocks of each problem in CTA order (deterministic).  `problems` is a HOST array. */
typedef struct cmp_dw_problem {
  const float* dY;      /* [M, Nout] upstream gradient */
  int64_t lddy;
  const float* saved_y; /* optional [M, Nout]: forward output of a fused ShiftedSoftplus */
  int64_t ldys;
  const float* X;       /* [M, K] forward input */
  int64_t ldx;
  int64_t M;
  int32_t K;
  int32_t Nout;
  float* dW;            /* [Nout, K] with leading dimension lddw (a block of a wider gradient when lddw > K) */
  int64_t lddw;
  float* db;            /* [Nout] or NULL */
} cmp_dw_problem_t;
int cmp_node_gemm_dw_group_max(void);
size_t cmp_node_gemm_dw_grouped_workspace(void);
int cmp_node_gemm_dw_grouped(const void* problems /* const cmp_dw_problem_t[count], host */, int count,
                             void* workspace, size_t workspace_bytes, cmp_stream_t stream);

/* Grouped weight packing: every weight image a training step needs in three launches instead of one per layer
 * (weights change once per step, in cmp_adam_step).  `jobs` are HOST arrays of at most 32 entries.
 * node job: both images of W[rows, cols] (normal, then transposed) into `packed`
 *           (cmp_node_gemm_weight_bytes(cols) + cmp_node_gemm_weight_bytes(rows) bytes);
 * filter job: forward images (cmp_cfconv_tc_weights_bytes()) into packed_fwd and/or backward images
 *           (cmp_cfconv_tc_bwd_weights_bytes()) into packed_bwd of one interaction block's filter MLP. */
typedef struct cmp_pack_node_job {
  const float* W;
  int32_t rows, cols;
  void* packed;
} cmp_pack_node_job_t;
typedef struct cmp_pack_filter_job {
  const float* W1;
  const float* b1;
  const float* W2;
  const float* b2;
  void* packed_fwd;
  void* packed_bwd;
} cmp_pack_filter_job_t;
int cmp_node_gemm_pack_weights_grouped(const void* jobs /* cmp_pack_node_job_t[count] */, int count,
                                       cmp_stream_t stream);
int cmp_cfconv_tc_pack_weights_grouped(const void* jobs /* cmp_pack_filter_job_t[count] */, int count,
                                       int num_filters, int num_gaussians, cmp_stream_t stream);
int cmp_cfconv_tc_pack_bwd_weights_grouped(const void* jobs /* cmp_pack_filter_job_t[count] */, int count,
                                           int num_filters, int num_gaussians, cmp_stream_t stream);

/* ------------------------------------------------------------------------- *
 * ViSNet edge-level kernels (exact fp32; tgv.py = torch_geometric_visnet.py)
 * Every reduction runs over the CSR (or its transpose) in a fixed order.
 * ------------------------------------------------------------------------- */

/* out[i] = sum_{e in row i} x[col[e]] * filt[e] * scale[e]  and its backward: the CFConv message
 * kernels with an explicit per-edge scale (NeighborEmbedding.forward/message, tgv.py:408-423:
 * scale = masked cosine cutoff, 0 on self loops). */
int cmp_edge_message_fwd(const float* x, const float* filt, const float* scale,
                         const int32_t* rowptr, const int32_t* col, int64_t N, int F, float* out,
                         cmp_stream_t stream);
int cmp_edge_message_bwd(const float* g, const float* x, const float* filt, const float* scale,
                         const int32_t* rowptr, const int32_t* col, const int32_t* rowptr_t,
                         const int32_t* col_t, const int32_t* eid_t, int64_t N, int F, float* dfilt,
                         float* dx, cmp_stream_t stream);
/* ExpNormalSmearing (tgv.py:100-111), CosineCutoff with the d < cutoff mask (tgv.py:44-46) and the
 * unit edge vectors of ViSNetBlock.forward (tgv.py:864-866; self loops keep their zero vector). */
int cmp_vis_edge_geometry(const float* evec, const float* dist, const int32_t* col,
                          const int32_t* erow, int64_t E, float cutoff, float alpha,
                          const float* means, const float* betas, int num_rbf, float* rbf,
                          float* dhat, float* C, cmp_stream_t stream);
/* torch.nn.LayerNorm over the last axis (tgv.py:605,883).  bwd returns dx and dy*xhat (its column
 * sums and those of dy are dweight / dbias). */
int cmp_layernorm_fwd(const float* x, const float* w, const float* b, int64_t M, int H, float eps,
                      float* y, float* mean, float* rstd, cmp_stream_t stream);
int cmp_layernorm_bwd(const float* dy, const float* x, const float* w, const float* mean,
                      const float* rstd, int64_t M, int H, float* dx, float* dyxhat,
                      cmp_stream_t stream);
/* out[i] = sum_{e in row i} x[perm ? perm[e] : e] (scatter(..., reduce='sum') of tgv.py:671; with
 * rowptr_t / eid_t the reduction at the SOURCE atoms), and out[e] = x[idx[e]] (the `_i` / `_j`
 * gathers of MessagePassing). */
int cmp_csr_segment_sum(const float* x, const int32_t* rowptr, const int32_t* perm, int64_t N, int C,
                        float* out, cmp_stream_t stream);
int cmp_gather_rows(const float* x, const int32_t* idx, int64_t E, int C, float* out,
                    cmp_stream_t stream);
/* EdgeEmbedding.forward (tgv.py:463-465): f[e] = (x_i + x_j) * ep[e]; bwd gives d ep and t = g*ep. */
int cmp_vis_edge_embed_fwd(const float* x, const float* ep, const int32_t* col, const int32_t* erow,
                           int64_t E, int H, float* f, cmp_stream_t stream);
int cmp_vis_edge_embed_bwd(const float* g, const float* x, const float* ep, const int32_t* col,
                           const int32_t* erow, int64_t E, int H, float* gep, float* t_out,
                           cmp_stream_t stream);
/* ViS_MP.message scalar part (tgv.py:644-648): attn = silu(sum_d q_i k_j dk) * C, m = v_j dv attn.
 * pre_act != 0: dk / dv are the OUTPUTS of dk_proj / dv_proj (tgv.py:622-629); their SiLU is applied inside the kernel
 * and g_dk / g_dv are the gradients of those pre-activations (no activated [E, H] tensor in HBM, no activation launch). */
int cmp_vis_message_fwd(const float* q, const float* k, const float* v, const float* dk,
                        const float* dv, const float* C, const int32_t* col, const int32_t* erow,
                        int64_t E, int H, int heads, int pre_act, float* m, float* attn_pre,
                        cmp_stream_t stream);
int cmp_vis_message_bwd(const float* gm, const float* q, const float* k, const float* v,
                        const float* dk, const float* dv, const float* C, const float* attn_pre,
                        const int32_t* col, const int32_t* erow, int64_t E, int H, int heads, int pre_act,
                        float* g_dk, float* g_dv, float* geq, float* gek, float* gev,
                        cmp_stream_t stream);
/* ViS_MP vector message + aggregation (tgv.py:650-651,672):
 * vagg[i] = sum_e vec[j] * s1[e] + s2[e] * dhat[e], s12 = [s1 | s2] per edge.
 * pre_act != 0: s12 is the OUTPUT of s_proj (tgv.py:649), SiLU applied inside; g_s12 = gradient of the pre-activation. */
int cmp_vis_vecagg_fwd(const float* vec, const float* s12, const float* dhat, const int32_t* rowptr,
                       const int32_t* col, int64_t N, int H, int pre_act, float* vagg, cmp_stream_t stream);
int cmp_vis_vecagg_bwd(const float* g, const float* vec, const float* s12, const float* dhat,
                       const int32_t* rowptr, const int32_t* col, const int32_t* rowptr_t,
                       const int32_t* col_t, const int32_t* eid_t, int64_t N, int H, int pre_act, float* g_s12,
                       float* g_vec, cmp_stream_t stream);
/* ViS_MP.edge_update (tgv.py:655-661) with w_trg / w_src applied at the nodes (bias-free linears):
 * wdot = wt_i . ws_j - (wt_i . dhat)(ws_j . dhat);  df = fpa * wdot. */
/* pre_act != 0: fpa is the OUTPUT of f_proj (tgv.py:659), SiLU applied inside.
 * cmp_vis_edge_update_bwd_prep: the two per-edge factors of the backward in one pass over n = E * H values,
 * g_fpa = g * wdot [* silu'(fpa)] and gw = g * act(fpa) (the input of cmp_vis_edge_update_bwd). */
int cmp_vis_edge_update_fwd(const float* wt, const float* ws, const float* dhat, const float* fpa,
                            const int32_t* col, const int32_t* erow, int64_t E, int H, int pre_act, float* df,
                            float* wdot, cmp_stream_t stream);
int cmp_vis_edge_update_bwd_prep(const float* g, const float* fpa, const float* wdot, int64_t n, int pre_act,
                                 float* g_fpa, float* gw, cmp_stream_t stream);
int cmp_vis_edge_update_bwd(const float* gw, const float* wt, const float* ws, const float* dhat,
                            const int32_t* rowptr, const int32_t* col, const int32_t* rowptr_t,
                            const int32_t* col_t, const int32_t* eid_t, int64_t N, int H, float* g_wt,
                            float* g_ws, cmp_stream_t stream);

/* ViS_MP node update (tgv.py:616-627), the element-wise tail of a layer in one kernel per direction:
 *   vp[N, 3, 3H] = vec_proj(vec) = [vec1 | vec2 | vec3],  o[N, 3H] = o_proj(x_agg) = [o1 | o2 | o3]
 *   dx = (sum_d vec1 vec2) o2 + o3,   dvec[d] = vec3[d] o1 + vagg[d]
 * bwd: g_vp[N, 3, 3H], g_o[N, 3H] from g_dx[N, H], g_dvec[N, 3, H] (the gradient of vagg is g_dvec itself). */
int cmp_vis_node_update_fwd(const float* vp, const float* o, const float* vagg, int64_t N, int H, float* dx,
                            float* dvec, cmp_stream_t stream);
int cmp_vis_node_update_bwd(const float* g_dx, const float* g_dvec, const float* vp, const float* o, int64_t N,
                            int H, float* g_vp, float* g_o, cmp_stream_t stream);

/* Debug hook: when non-NULL, CTA 0 / pipeline 0 of cmp_cfconv_fused_fwd stores 8 clock64() phase
 * timestamps per tile (first 32 tiles) into this device buffer of 256 int64. */
void cmp_debug_set_fwd_timestamps(void* buf);
/* Same for cmp_cfconv_pair_fwd: 10 timestamps per tile of pipeline 0 (first 24 tiles), 240 int64. */
void cmp_debug_set_pair_timestamps(void* buf);
/* Same for cmp_cfconv_dense_fwd: 7 timestamps + the tile width per executed tile of pipeline 0 (first 32), 256 int64. */
void cmp_debug_set_dense_timestamps(void* buf);
/* Tuning knob: start-up delay (ns) between the four pipelines of a cmp_cfconv_dense_fwd CTA (default 1500). */
void cmp_debug_set_dense_stagger(int ns);
/* Debug knob: only the first n pipelines of every cmp_cfconv_dense_fwd CTA take work (default 4). */
void cmp_debug_set_dense_pipes(int n);
/* Debug knob of the timestamped build of cmp_cfconv_dense_fwd: bit 0 = skip the a' stores, bit 1 = skip the cutoff loads,
 * bit 2 = skip the TMEM loads of epilogue 1 (results are then meaningless; for phase timing only). */
void cmp_debug_set_dense_mode(int mode);
/* Kernel variant of cmp_cfconv_dense_fwd: -1 = by max_atoms_hint (default: the warp-specialised tile pipeline when no
 * conformer exceeds 32 atoms - they then stay in registers -, the per-pipeline kernel otherwise), 0 = warp-specialised,
 * 1 = per-pipeline. */
void cmp_debug_set_dense_variant(int variant);
/* Kernel variant of cmp_cfconv_dense_bwd_weights: 0 = two groups of 256 threads that run the phases of a tile in sequence
 * (default), 1 = warp-specialised tile pipeline (measured slower; kept for comparison). */
void cmp_debug_set_dense_bwd_variant(int variant);
/* Same for cmp_cfconv_fused_bwd_weights: 12 timestamps per tile (first 20 tiles), 240 int64. */
void cmp_debug_set_bwd_timestamps(void* buf);

/* Single-tile UMMA probe used by the tests to pin descriptor / TMEM conventions. */
int cmp_debug_umma_gemm(const void* a_img, int64_t a_bytes, const void* b_img, int64_t b_bytes,
                        float* D, int N, int K, int fmt, int a_mn, int b_mn, int a_lbo, int a_sbo,
                        int a_kstep, int b_lbo, int b_sbo, int b_kstep, cmp_stream_t stream);

/* Fused Adam step on flat fp32 buffers (the optimiser of model/common.py:368-370). */
int cmp_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                  float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                  float grad_scale, cmp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CONANMP_H_ */
