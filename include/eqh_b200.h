/*
 * eqh_b200.h -- C ABI of libeqh_b200.so: the B200 (sm_100a) kernels behind EquiHGNN's
 * message-passing hot path.
 *
 * The reference (HySonLab/EquiHGNN) is pure Python and has no FFI of its own; the seam it
 * offers is the nn.Module surface (MHNNConv / MHNNSConv / EGNN) and the third-party
 * `torch_scatter.scatter` call.  Every entry point below replaces a specific group of ATen /
 * torch_scatter calls inside those modules; the reference lines are cited per function.
 * INTEGRATION.md shows the Python (ctypes) binding a maintainer adds on the reference side.
 *
 * Conventions (all functions):
 *   - plain C types only: device pointers, explicit sizes, a CUDA stream handle;
 *   - the caller owns every buffer (inputs, outputs, workspaces); nothing is allocated here;
 *   - launches are asynchronous on `stream`; no function synchronises or throws;
 *   - inputs are never modified;
 *   - return value: 0 = EQH_OK, <0 = argument/support error (EQH_ERR_*), >0 = cudaError_t;
 *   - feature matrices are row-major fp32 with row stride == C (contiguous), 16-byte aligned;
 *   - device index arrays are int32 unless the name says i64;
 *   - re-entrant: no global mutable state.
 */
#ifndef EQH_B200_H_
#define EQH_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* eqh_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define EQH_API __attribute__((visibility("default")))
#else
#define EQH_API
#endif

#define EQH_OK 0
#define EQH_ERR_ARG (-1)         /* null pointer / negative size / misaligned buffer            */
#define EQH_ERR_UNSUPPORTED (-2) /* valid request outside the implemented set (e.g. C % 32 != 0) */
#define EQH_ERR_WORKSPACE (-3)   /* workspace too small                                          */

#define EQH_REDUCE_SUM 0
#define EQH_REDUCE_MEAN 1

#define EQH_KNN_K 16 /* neighbours per atom (equihnn_egnn.py:128 num_nearest_neighbors=16) */
#define EQH_M_DIM 16 /* EGNN message width (egnn_layer.py:150 m_dim=16)                     */

EQH_API int eqh_version(void);
EQH_API const char* eqh_error_string(int rc);
/* number of CTAs the persistent kernels launch (= partial-buffer rows the caller must provide) */
EQH_API int eqh_partial_rows(void);

/* ------------------------------------------------------------------------------------------
 * Incidence CSR (stable sort by destination).
 * Replaces: the implicit grouping done by torch_scatter's scatter_add_ over `index`
 * (conv.py:91-93,97,173,177) and fixes the batched numbering of HData.__inc__
 * (data/utils.py:172-178).  For dst[z] in [0,n_dst):
 *     rowptr[s]            = #{z : dst[z] < s}                       (n_dst+1 entries)
 *     perm[rowptr[s] + r]  = the r-th smallest z with dst[z] == s     (stable)
 *     dst32[z]             = (int32) dst[z]
 *     other_sorted[k]      = (int32) other[perm[k]]    (optional, other may be NULL)
 * Bit-exact against torch.sort(stable=True) / bincount / cumsum.
 * status (optional, 1 int32, caller zeroes it): set to 1 if any dst is outside [0,n_dst).
 * ---------------------------------------------------------------------------------------- */
EQH_API size_t eqh_incidence_workspace_bytes(int64_t Z, int64_t n_dst);
EQH_API int eqh_incidence_build_i64(const int64_t* dst, const int64_t* other, int64_t Z, int64_t n_dst, int32_t* rowptr,
                            int32_t* perm, int32_t* dst32, int32_t* other_sorted, int32_t* status, void* ws,
                            size_t ws_bytes, eqh_stream_t stream);
EQH_API int eqh_incidence_build_i32(const int32_t* dst, const int32_t* other, int64_t Z, int64_t n_dst, int32_t* rowptr,
                            int32_t* perm, int32_t* dst32, int32_t* other_sorted, int32_t* status, void* ws,
                            size_t ws_bytes, eqh_stream_t stream);

/* Pooling index of a PyG batch (global_add_pool, mhnn.py:216, equihnn_egnn.py:167): `batch` [n] is sorted, so the CSR
 * of the pooling is `ptr` [num_graphs+1] with the identity permutation; this only narrows both to int32. */
EQH_API int eqh_pool_index(const int64_t* ptr, int64_t num_graphs, const int64_t* batch, int64_t n, int32_t* rowptr32,
                   int32_t* batch32, eqh_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Deterministic segmented reduction (no atomics), the replacement of
 * torch_scatter.scatter(src, index, dim=-2, reduce="sum"|"mean") (conv.py:91-93,97,173,177)
 * and of torch_geometric global_add_pool (mhnn.py:216, equihnn_egnn.py:167).
 *     out[s,:] = reduce_{k in [rowptr[s],rowptr[s+1])}  w(k) * table[row(k),:]
 *     row(k) = idx ? idx[k] : k ;  w(k) = 1 / max(1, src_rowptr[row+1]-src_rowptr[row]) if src_rowptr else 1
 *     mean divides by max(1,count) (empty segments give 0, as torch_scatter does).
 * Rows are added in ascending k, i.e. in original incidence order when idx = perm: the fp32
 * result is bit-identical to the sequential CPU scatter_add_ of the reference.
 * The same kernel is the backward of a gather+mean (src_rowptr carries the forward counts).
 * ---------------------------------------------------------------------------------------- */
EQH_API int eqh_gather_reduce(const float* table, const int32_t* idx, const int32_t* rowptr, const int32_t* src_rowptr,
                      int64_t n_out, int64_t C, int mode, float* out, eqh_stream_t stream);
/* out[r,:] = w(r) * table[idx[r],:]  for r in [0,n_rows): the backward of a segmented reduction
 * (w = 1/max(1,count of segment idx[r]) when scale_rowptr != NULL) and the row gather X[vertex]. */
EQH_API int eqh_gather_rows(const float* table, const int32_t* idx, const int32_t* scale_rowptr, int64_t n_rows, int64_t C,
                    float* out, eqh_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused per-incidence message + segmented reduction:
 *     t_k   = ReLU( A[ia(k),:] + B[ib(k),:] )            ia/ib == NULL -> the segment id
 *     u_k   = gamma ? LayerNorm(t_k; gamma, beta, eps) : t_k
 *     U[s,:] = reduce_{k in segment s} u_k               (sum | mean, empty -> 0)
 * Replaces, per MLP(2 layers) applied to cat(X[vertex], E[edges]) and scattered
 * (conv.py:90-93, 96-97, 175-177 with mlp.py:91-99): the gathers, the concat, ReLU, LayerNorm
 * and the scatter.  The first Linear is split over the concat and applied to the N / E rows
 * beforehand (A, B; bias folded into one of them), the last Linear commutes with the
 * reduction and is applied to U afterwards -- both by the host module.
 * ---------------------------------------------------------------------------------------- */
EQH_API int eqh_fused_msg_fwd(const float* A, const float* B, const int32_t* ia, const int32_t* ib, const int32_t* rowptr,
                      int64_t n_out, int64_t C, const float* gamma, const float* beta, float eps, int mode, float* U,
                      eqh_stream_t stream);
/* Backward of the above, evaluated per grouping (call once per operand):
 *     g_k    = su(k) * dU[iu(k),:]      su = 1/max(1,count) from u_rowptr (mean) or 1 (NULL)
 *     dpre_k = ReLU'( . ) * LayerNormBackward(g_k)   (t_k recomputed from A, B)
 *     dOp[s,:] = sum_{k in segment s} dpre_k         (gradient of the operand that is constant
 *                                                     over this grouping's segments)
 *     dgb_partial[cta, 0:C] += sum g_k * xhat_k ; dgb_partial[cta, C:2C] += sum g_k
 *         (optional; eqh_partial_rows() x 2C floats, then eqh_colsum)
 * ia / ib / iu == NULL mean "the segment id".
 */
EQH_API int eqh_fused_msg_bwd(const float* A, const float* B, const int32_t* ia, const int32_t* ib, const int32_t* iu,
                      const int32_t* rowptr, int64_t n_seg, int64_t C, const float* gamma, const float* beta,
                      float eps, const float* dU, const int32_t* u_rowptr, float* dOp, float* dgb_partial,
                      eqh_stream_t stream);
/* Row LayerNorm of the dense MLP hidden layers (mlp.py:94-96: relu_ followed by LayerNorm), with the
 * ReLU optionally folded in: y = LN(relu ? max(x,0) : x).  Backward: dx (ReLU mask applied when relu),
 * dgb_partial [eqh_partial_rows(), 2C] = per-CTA partials of dgamma | dbeta (finish with eqh_colsum). */
EQH_API int eqh_rowln_fwd(const float* x, const float* gamma, const float* beta, float eps, int relu, int64_t rows, int64_t C,
                  float* y, eqh_stream_t stream);
EQH_API int eqh_rowln_bwd(const float* x, const float* gamma, float eps, int relu, const float* dy, int64_t rows, int64_t C,
                  float* dx, float* dgb_partial, eqh_stream_t stream);
/* Column sums of a tall row-major [rows, cols] matrix with row stride ld (bias gradients of nn.Linear), optionally
 * row-weighted: out (+)= sum_r roww[r] * x[r,:]; two fixed-order stages through ws (eqh_colsum_workspace_bytes,
 * always required when roww != NULL or rows > 64).  Deterministic. */
EQH_API size_t eqh_colsum_workspace_bytes(int64_t rows, int64_t cols);
EQH_API int eqh_colsum_tall(const float* x, int64_t rows, int64_t cols, int64_t ld, const float* roww /* optional [rows] */,
                    float* out, int accumulate, float* ws, eqh_stream_t stream);
/* out[c] (+)= scale * sum_r partial[r*ld + c] for r in [0,rows): fixed-order (deterministic) column sum */
EQH_API int eqh_colsum(const float* partial, int64_t rows, int64_t cols, int64_t ld, float* out, float scale, int accumulate,
               eqh_stream_t stream);

/* Batched form: every bias gradient and per-CTA partial buffer of one layer's backward in ONE launch.
 *     out_j[c] (+)= scale_j * sum_r roww_j[r] * x_j[r*ld_j + c]      (roww_j optional)
 * Row chunks are reduced by separate blocks into ws; the block arriving last at a column tile's counter adds the
 * chunk partials in ascending chunk order (deterministic) and resets the counter: `counters`
 * (eqh_colsum_batch_counters() int32) must be zero on first use and is left zero.  Two launches may share
 * counters / ws only if they are ordered on one stream. */
#define EQH_COLSUM_MAX_JOBS 32
typedef struct eqh_colsum_job {
  const float* x;      /* [rows, ld] row-major */
  const float* roww;   /* optional [rows] */
  float* out;          /* [cols] */
  int64_t rows, cols, ld;
  float scale;
  int32_t accumulate;  /* 0: out = ..., 1: out += ... */
} eqh_colsum_job;
EQH_API size_t eqh_colsum_batch_workspace_bytes(const eqh_colsum_job* jobs, int n_jobs);
EQH_API int64_t eqh_colsum_batch_counters(const eqh_colsum_job* jobs, int n_jobs);
EQH_API int eqh_colsum_batch(const eqh_colsum_job* jobs /* host array */, int n_jobs, float* ws, int32_t* counters,
                     eqh_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Batch-wide 16-nearest-neighbour search on squared distance.
 * Replaces egnn_layer.py:253-259,281-283: the [1,N,N,3] rel_coors, [1,N,N] ranking and topk(16,
 * largest=False).  d2 = (dx*dx + dy*dy) + dz*dz evaluated with separate fp32 multiplies and adds
 * (no FMA contraction), identical to the reference's (rel**2).sum(-1), so the neighbour set is
 * bit-identical for tie-free inputs.  Output ascending by (d2, index); self is rank 0.
 * N < 16 is an argument error (the reference's topk raises).  O(N*K) memory.
 * ---------------------------------------------------------------------------------------- */
EQH_API int eqh_knn16(const float* pos /*[N,3]*/, int64_t N, int32_t* idx /*[N,16]*/, float* d2 /*[N,16]*/,
              eqh_stream_t stream);
/* Same contract and same output bits through a uniform cell grid built on device (robust box, counting sort by
 * cell, ring-by-ring search with an exact stopping bound): O(N * neighbourhood) instead of O(N^2), for the
 * configurations with tens of thousands of atoms per batch.  ws: eqh_knn16_grid_workspace_bytes(N), 16-byte aligned. */
EQH_API size_t eqh_knn16_grid_workspace_bytes(int64_t N);
EQH_API int eqh_knn16_grid(const float* pos /*[N,3]*/, int64_t N, int32_t* idx /*[N,16]*/, float* d2 /*[N,16]*/, void* ws,
                   size_t ws_bytes, eqh_stream_t stream);
/* Same, and `order` [N] receives the atoms sorted by grid cell (stable): visiting atoms in this order makes the
 * neighbour-row gathers of the edge kernels local (the *_ordered entry points below). */
EQH_API int eqh_knn16_grid_ordered(const float* pos, int64_t N, int32_t* idx, float* d2, int32_t* order /*[N]*/, void* ws,
                           size_t ws_bytes, eqh_stream_t stream);
/* The cell visiting order alone: order [N] = atoms sorted by grid cell (stable), same grid as eqh_knn16_grid; workspace
 * of eqh_knn16_grid_workspace_bytes(N). */
EQH_API int eqh_cell_order(const float* pos, int64_t N, int32_t* order, void* ws, size_t ws_bytes, eqh_stream_t stream);


/* ------------------------------------------------------------------------------------------
 * EGNN edge model phi_e + neighbour pooling (egnn_layer.py:298-310,357-358).
 *   With W1 = edge_mlp.0.weight [H, 2C+1] (H = 2(2C+1), zero-padded by the host to a multiple of 64) split by columns into W1a|W1b|wc, the
 *   host computes P = h W1a^T + b1 and Q = h W1b^T on the N atom rows; per kNN edge (i,k), j=idx[i,k]:
 *     pre  = P[i,:] + Q[j,:] + d2[i,k] * wc
 *     s2   = W2 silu(pre) + b2                 (W2 = edge_mlp.3.weight [16,H])
 *     m_ij = silu(s2) ;  m_i = sum_k m_ij      (m_pool_method="sum")
 *   S2 [N*16,16] (pre-activations) is kept for the backward; Mij [N*16,16] (optional) feeds phi_x.
 * ---------------------------------------------------------------------------------------- */
/* Re-lay phi_e's parameters for the kernels in one launch: edge_mlp.0.weight W1 [H0, 2C+1] -> W1a | W1b [Hp, C] and the
 * d2 column wc [Hp]; edge_mlp.0.bias -> b1p [Hp]; edge_mlp.3.weight W2 [16, H0] -> W2p [16, Hp] (Hp = H0 rounded up to
 * 64; the caller zeroes the destination buffers once, pad entries are never written). */
EQH_API int eqh_egnn_pack(const float* W1, const float* b1, const float* W2, int64_t H0, int64_t C, int64_t Hp, float* W1a,
                  float* W1b, float* wc, float* b1p, float* W2p, eqh_stream_t stream);
EQH_API int eqh_egnn_edge_fwd(const float* P, const float* Q, const float* wc, const float* W2, const float* b2,
                      const int32_t* idx, const float* d2, int64_t N, int64_t H, float* S2, float* Mij /* optional */,
                      float* Mi, eqh_stream_t stream);
EQH_API size_t eqh_egnn_edge_fwd_smem_bytes(int64_t H);
/* Same contract, tensor-core implementation: the 16-wide contraction runs on tcgen05 (3xTF32, fp32-faithful,
 * [128 edges x 32] hid tiles produced into 128B-swizzled shared memory by the gather/SiLU warps, accumulators
 * double-buffered in TMEM).  eqh_egnn_edge_fwd is the CUDA-core version of the same math. */
EQH_API int eqh_egnn_edge_fwd_tc(const float* P, const float* Q, const float* wc, const float* W2, const float* b2,
                         const int32_t* idx, const float* d2, int64_t N, int64_t H, float* S2, float* Mij /* optional */,
                         float* Mi, eqh_stream_t stream);
/* ------------------------------------------------------------------------------------------
 * Shared neighbour-search primitive for the other geometric encoders.
 *
 * eqh_knn: exact k nearest neighbours on squared distance, k in [1, 32], ascending, ties by smaller index.
 *   seg_ptr / seg_of NULL: the whole cloud is one range (EGNN, egnn_layer.py:281-283; self is rank 0 unless
 *   exclude_self).  Otherwise atoms are grouped in contiguous ranges [seg_ptr[s], seg_ptr[s+1]) (a collated batch:
 *   PyG Batch.ptr) and seg_of[i] is atom i's range: each atom searches its own molecule only -- the dense-batch
 *   semantics of FAFormer's _build_graph (fa_former_layer.py:650-667, exclude_self = 1), Equiformer
 *   (equiformer_layer.py:1326) and the SE(3)-Transformer (se3_transformer_layer.py:1577).  Slots beyond the number of
 *   candidates hold idx = -1, d2 = +inf (the references fill them with masked 1e9 / max_value distances).
 * eqh_radius: torch_cluster.radius_graph semantics (visnet_layer.py:336): for atom i the first max_num_neighbors atoms
 *   j of its range, in ascending j, with d2(i,j) < r*r (strict); loop != 0 keeps j == i.  idx [N, max_num_neighbors]
 *   (-1 padded), count [N], optional d2 [N, max_num_neighbors].
 * ---------------------------------------------------------------------------------------- */
EQH_API int eqh_knn(const float* pos /*[N,3]*/, int64_t N, int k, const int32_t* seg_ptr, const int32_t* seg_of,
            int exclude_self, int32_t* idx /*[N,k]*/, float* d2 /*[N,k]*/, eqh_stream_t stream);
EQH_API int eqh_radius(const float* pos /*[N,3]*/, int64_t N, float r, const int32_t* seg_ptr, const int32_t* seg_of, int loop,
               int max_num_neighbors, int32_t* idx, int32_t* count /*[N]*/, float* d2 /*optional*/, eqh_stream_t stream);

/* eqh_egnn_edge_fwd_tc visiting the atoms in `order` (a permutation of [0,N), e.g. from eqh_knn16_grid_ordered; NULL =
 * natural order).  Outputs are identical; only the memory access pattern changes. */
EQH_API int eqh_egnn_edge_fwd_tc_ordered(const float* P, const float* Q, const float* wc, const float* W2, const float* b2,
                                 const int32_t* idx, const float* d2, int64_t N, int64_t H, float* S2,
                                 float* Mij /* optional */, float* Mi, const int32_t* order, eqh_stream_t stream);
/* Backward.  dMi [N,16] and/or dMij [N*16,16] (either may be NULL).  Outputs:
 *   dP [N,H]; dPre [N*16,H] (per-edge gradient of `pre`: dQ[j] = sum over incoming edges, done by
 *   eqh_gather_reduce over the transposed kNN incidence); dd2 [N,16] (optional);
 *   part [eqh_partial_rows(), 17*H + 16]: per-CTA partial sums of dW2 (16*H), dwc (H), db2 (16),
 *   to be finished with eqh_colsum.
 */
EQH_API int eqh_egnn_edge_bwd(const float* P, const float* Q, const float* wc, const float* W2, const float* b2,
                      const int32_t* idx, const float* d2, const float* S2, const float* dMi, const float* dMij,
                      int64_t N, int64_t H, float* dP, float* dPre, float* dd2, float* part, eqh_stream_t stream);
/* Same, atoms visited in `order` (dP / dPre / dd2 identical; the per-CTA partials in `part` are summed in another
 * order, i.e. the weight gradients agree to rounding). */
EQH_API int eqh_egnn_edge_bwd_ordered(const float* P, const float* Q, const float* wc, const float* W2, const float* b2,
                              const int32_t* idx, const float* d2, const float* S2, const float* dMi, const float* dMij,
                              int64_t N, int64_t H, float* dP, float* dPre, float* dd2, float* part, const int32_t* order,
                              eqh_stream_t stream);

/* Backward on the tcgen05 tensor cores (egnn_edge_bwd_tc.cu) for the case egnn_equihnns trains: only dMi given
 * (no dMij), no gradient w.r.t. d2.  Both 16-wide contractions (dH = dS2.W2, dW2 = dS2^T.hid) run as 3xTF32 MMAs; a
 * CTA owns 128 hidden columns (the TMEM lanes) for a slice of the edges, hid goes back to tensor memory as the A
 * operand of the weight-gradient MMA.  Same outputs as eqh_egnn_edge_bwd_ordered (dP, dPre identical up to fp32
 * rounding of the contraction), except that `part` has eqh_egnn_edge_bwd_tc_rows(H) rows (<= eqh_partial_rows()),
 * every one of them fully written.  order: optional visiting order of the atoms (NULL = natural). */
EQH_API int eqh_egnn_edge_bwd_tc_rows(int64_t H);
/* edge slices a launch over N atoms really uses (<= eqh_egnn_edge_bwd_tc_rows(H): small batches leave a few SMs to the
 * parallel branches of the step; the rows of `part` it does not write are zeroed by the launch) */
EQH_API int eqh_egnn_edge_bwd_tc_slices(int64_t H, int64_t N);
EQH_API int eqh_egnn_edge_bwd_tc(const float* P, const float* Q, const float* wc, const float* W2, const int32_t* idx,
                         const float* d2, const float* S2, const float* dMi, int64_t N, int64_t H, float* dP, float* dPre,
                         float* part, const int32_t* order, eqh_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Readout head + loss: the last Linear of the readout MLP with ONE target
 * (mhnn.py:216-218, equihnn_egnn.py:167-169) and the MSE loss of main.py:49-51 over the first B rows
 * (row B.. of a ghost-padded batch are not in the loss), forward and backward, one launch each.
 *   fwd: pred [R] = z [R,K] . w [K] + b ; resid [R] = 2 (pred - y) / B (0 for r >= B) ; loss [1] = mean_{r<B} (pred - y)^2
 *   bwd: dz [R,K] = g resid w ; dw [K] (+)= g sum_r resid[r] z[r,:] ; db [1] (+)= g sum_r resid[r]
 *        g: device scalar (d loss), NULL = 1; accumulate != 0 adds into dw / db.   K <= 1024. */
EQH_API int eqh_head_mse_fwd(const float* z, const float* w, const float* b /* optional */, const float* y, int64_t R,
                     int64_t K, int64_t B, float* pred, float* resid, float* loss, eqh_stream_t stream);
EQH_API int eqh_head_mse_bwd(const float* z, const float* w, const float* resid, const float* g /* optional */, int64_t R,
                     int64_t K, float* dz, float* dw, float* db /* optional */, int accumulate, eqh_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * AtomEncoder (third-party ogb==1.3.6; call sites mhnn.py:201, equihnn_egnn.py:157):
 *     out[n,:] = sum_f tables[f][x[n,f], :]            x [N,F] int64, tables[f] [rows_per_table[f], C]
 * (tables / rows_per_table are HOST arrays of F device pointers / row counts, F <= 16).
 * Backward: per-slice partial table gradients, partial [eqh_atom_encoder_slices(N), sum(rows), C],
 * to be finished with eqh_colsum(partial, slices, sum(rows)*C); table f's gradient is rows
 * [off_f, off_f + rows_f) of the result.  Deterministic (ascending atom order, no atomics).
 * status (optional): set to 1 if a feature value is outside its table.
 * ---------------------------------------------------------------------------------------- */
EQH_API int eqh_atom_encoder_slices(int64_t N);
EQH_API int eqh_atom_encoder_fwd(const int64_t* x, const float* const* tables, const int32_t* rows_per_table, int F, int64_t N,
                         int64_t C, float* out, int32_t* status, eqh_stream_t stream);
EQH_API int eqh_atom_encoder_bwd(const int64_t* x, const float* dh, const int32_t* rows_per_table, int F, int64_t N, int64_t C,
                         float* partial, eqh_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Device-side batch collation (the caller of the path: main.py:227-229 DataLoader + PyG Batch.from_data_list,
 * data/utils.py:150-178).  The dataset lives in HBM as a packed per-molecule store; a batch is gathered by molecule id:
 *     x / pos / y copied; edge_index0 = local atom id + atom offset of the molecule in the batch, edge_index1 = local
 *     hyperedge id + hyperedge offset (HData.__inc__); batch / ptr / n_e as PyG builds them.
 * `offsets` (device, int32 [3][B+1]): exclusive prefix sums of the selected molecules' atom / hyperedge / incidence
 * counts (computed by the host from its copy of the store's offsets -- 3(B+1) integers per step instead of the batch).
 * N_pad > N appends the inert ghost molecule of the CUDA-graph trainer (graphed.pad_batch): ghost atoms on a 1.5 A
 * lattice of side `ghost_side` at `ghost_origin`, ghost hyperedges (edge_attr 5) and incidences that only connect
 * ghosts.  Bit-exact against the host collation.  Outputs hold N_pad / E_pad / Z_pad / B(+1) entries.
 * ---------------------------------------------------------------------------------------- */
typedef struct eqh_store {
  const int64_t* x;         /* [atoms, 9] */
  const float* pos;         /* [atoms, 3] */
  const float* y;           /* [molecules] */
  const int32_t* v_loc;     /* [incidences] atom id inside the molecule */
  const int32_t* e_loc;     /* [incidences] hyperedge id inside the molecule */
  const int64_t* edge_attr; /* [hyperedges] */
  const int64_t* e_order;   /* [hyperedges] */
  const int64_t* atom_off;  /* [molecules+1] */
  const int64_t* he_off;    /* [molecules+1] */
  const int64_t* inc_off;   /* [molecules+1] */
} eqh_store;
typedef struct eqh_batch {
  int64_t* x; float* pos; int64_t* edge_index0; int64_t* edge_index1; int64_t* edge_attr; int64_t* n_e; int64_t* e_order;
  int64_t* batch; int64_t* ptr; float* y;
} eqh_batch;
EQH_API int eqh_collate(const eqh_store* store, const int32_t* ids /*dev [B]*/, const int32_t* offsets /*dev [3][B+1]*/,
                int64_t B, int64_t N, int64_t E, int64_t Z, int64_t N_pad, int64_t E_pad, int64_t Z_pad, int64_t ghost_side,
                float ghost_origin, const eqh_batch* out, eqh_stream_t stream);

/* The incidence CSR of the collated batch (both groupings, what eqh_incidence_build_* would compute from the batch's
 * edge_index0 / edge_index1: data/utils.py:172-178 numbering, conv.py:173/177 scatters) straight out of the collation:
 * a batch is a block-diagonal concatenation of molecules in id order, so its stable by-hyperedge / by-vertex orders are
 * the molecules' own LOCAL orders (computed once per dataset, kept in the store) shifted by the molecule's offsets --
 * one gather + offset per element instead of histogram / scan / placement / ordering passes every step; the ghost
 * molecule's segments (incidence zi -> atom N + zi % ga, hyperedge E + zi % ge) are closed-form.
 * Bit-exact against eqh_incidence_build_i64 on the collated index lists.  All outputs int32. */
typedef struct eqh_store_csr {
  const int32_t* rowptr_e_loc; /* per molecule n_he+1 entries, molecule m starts at he_off[m] + m */
  const int32_t* perm_e_loc;   /* [incidences] local position of the k-th incidence in by-hyperedge order */
  const int32_t* rowptr_v_loc; /* per molecule n_atoms+1 entries, molecule m starts at atom_off[m] + m */
  const int32_t* perm_v_loc;   /* [incidences] */
} eqh_store_csr;
typedef struct eqh_incidence_out {
  int32_t* rowptr_e;  /* [E_pad+1] */
  int32_t* perm_e;    /* [Z_pad] */
  int32_t* vtx_by_e;  /* [Z_pad] */
  int32_t* rowptr_v;  /* [N_pad+1] */
  int32_t* perm_v;    /* [Z_pad] */
  int32_t* edge_by_v; /* [Z_pad] */
  int32_t* vertex32;  /* [Z_pad] edge_index0 as int32 */
  int32_t* edges32;   /* [Z_pad] edge_index1 as int32 */
} eqh_incidence_out;
EQH_API int eqh_collate_incidence(const eqh_store* store, const eqh_store_csr* csr, const int32_t* ids /*dev [B]*/,
                          const int32_t* offsets /*dev [3][B+1]*/, int64_t B, int64_t N, int64_t E, int64_t Z, int64_t N_pad,
                          int64_t E_pad, int64_t Z_pad, const eqh_incidence_out* out, eqh_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * fp32-faithful row GEMM on the tcgen05 tensor cores (3xTF32 operand splitting, fp32 accumulate):
 *     D[M,N] = act( A . B^T + bias ) + residual          A logical [M,K], B logical [N,K]
 * Replaces the cuBLAS SIMT sgemm calls behind nn.Linear in mlp.py:91-99, egnn_layer.py:180-208 and
 * the P/Q GEMMs of the phi_e split, forward (A, B K-major), dX = dY.W (B MN-major) and
 * dW = dY^T.X (both MN-major; split-K over the rows).
 *   a_kmajor / b_kmajor: 1 = stored row-major [rows, K] (K contiguous), 0 = stored [K, rows].
 *   lda / ldb: row stride of the STORED matrix in floats (multiple of 4, base 16-byte aligned).
 *   Full epilogue: D = out_scale * act(A.B^T + bias[n] * rowscale[m]) + res_scale * residual
 *   D row-major [M, ldd]; bias [N], rowscale [M], residual [M, ldd] optional (residual may alias D: in-place
 *   accumulation); act: 0 none, 1 ReLU, 2 SiLU.
 *   splits > 1 (split-K over the reduction, bias/act must be null/0): partials go to ws [splits, M, N]
 *   (eqh_gemm_workspace_bytes); the CTA arriving last at an output tile's counter adds them in split order
 *   (deterministic) and applies out_scale / residual.  `counters`: one zeroed int32 per output tile (count from
 *   eqh_gemm_plan), left zero.  `splits` must be a value returned by eqh_gemm_plan.
 *   eqh_gemm_plan(M,N,K, splits_req, flags): splits_req 1 = no split, 0 = choose, >1 = at most that many;
 *   returns the tile width the kernel will use, the effective split count and the number of output tiles.
 *   act flags: bits 0-7 activation; 0x100 store the masked hi part (verification); bits 9-10 force the tile
 *   width (1/2/3 = 32/64/128); 0x1000 first-generation kernel (both operands staged and split in shared memory;
 *   the default kernel moves A from shared memory through registers into tensor memory, gemm_ts.cu); 0x4000
 *   never use the persistent variant (gemm_ps.cu: chosen automatically for multi-wave problems without split-K);
 *   0x8000 (eqh_gemm_plan with splits_req 0) background GEMM: minimise CTAs x duration instead of latency (weight
 *   gradients issued on a side stream next to the critical path).
 * Ragged M / N / K are handled (TMA zero fill, bounds-checked stores).
 * ---------------------------------------------------------------------------------------- */
EQH_API int eqh_gemm_plan(int64_t M, int64_t N, int64_t K, int splits_req, int flags, int* bn, int* splits, int64_t* tiles);
EQH_API size_t eqh_gemm_workspace_bytes(int64_t M, int64_t N, int splits);
EQH_API int eqh_gemm_tf32x3(const float* A, int a_kmajor, int64_t lda, const float* B, int b_kmajor, int64_t ldb, float* D,
                    int64_t ldd, int64_t M, int64_t N, int64_t K, const float* bias, const float* rowscale, int act,
                    const float* residual, float out_scale, float res_scale, int splits, float* ws, int32_t* counters,
                    eqh_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Flat fused Adam step over one contiguous fp32 parameter buffer (main.py:137-140:
 * torch.optim.Adam(lr, weight_decay)); grad is multiplied by grad_scale first (1/world after the
 * summed all-reduce).  L2 weight decay as torch.optim.Adam (not AdamW).
 * ---------------------------------------------------------------------------------------- */
EQH_API int eqh_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                  float beta1, float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                  eqh_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * EGNN coordinate branch phi_x (egnn_layer.py:325-340 + CoorsNorm :71-81):
 *   coors_out[i] = pos[i] + sum_k w_ik * (pos[i] - pos[j]) * scale / max(|pos[i] - pos[j]|, 1e-8),  j = idx[i,k],
 *   w_ik = W4 . SiLU(W3 . m_ij[i,k,:] + b3) + b4      (W3 [64,16], b3 [64], W4 [64], b4 [1]; scale NULL = no CoorsNorm)
 * forward: Mij [N,16,16] (from the edge kernels) -> coors_out [N,3] (+ optional wij [N,16]).
 * backward (per kNN edge e): dMij [N*16,16], dpre3 / thid [N*16,64], dw / dscale_e [N*16], grel [N*16,3]; the
 * parameter gradients are dW3 = dpre3^T . Mij, db3 = colsum(dpre3), dW4 = thid^T . dw, db4 = sum(dw),
 * dscale = sum(dscale_e) (library GEMM / column sums).  eqh_egnn_coors_bwd_pos: dpos = dcoors_out + own-edge minus
 * incoming-edge sums of grel over the transposed kNN incidence (rowptr_j / perm_j), deterministic.
 * ---------------------------------------------------------------------------------------- */
EQH_API int eqh_egnn_coors_fwd(const float* Mij, const float* pos, const int32_t* idx, const float* W3, const float* b3,
                       const float* W4, const float* b4, const float* scale, int64_t N, float* coors_out, float* wij,
                       eqh_stream_t stream);
EQH_API int eqh_egnn_coors_bwd(const float* Mij, const float* pos, const int32_t* idx, const float* W3, const float* b3,
                       const float* W4, const float* b4, const float* scale, const float* dcoors_out, int64_t N, float* dMij,
                       float* dpre3, float* thid, float* dw, float* dscale_e, float* grel, eqh_stream_t stream);
EQH_API int eqh_egnn_coors_bwd_pos(const float* dcoors_out, const float* grel, const int32_t* rowptr_j, const int32_t* perm_j,
                           int64_t N, float* dpos, eqh_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * nn.BatchNorm1d over atom rows in training mode (mhnn.py:182,206; equihnn_egnn.py:226,246), forward and backward.
 * Batch statistics over the first n_stat rows (ghost-padded batches keep their padding atoms after the real ones and
 * out of the statistics), all `rows` rows normalised: y = (x - mean) * rstd * gamma + beta with the biased variance;
 * running_mean / running_var (optional) are updated with `momentum` and the unbiased variance, num_batches_tracked
 * (optional int64 scalar) is incremented.  n_stat_dev (optional, device int64) overrides n_stat: the number of real atoms
 * of a ghost-padded batch changes between replays of one captured graph.  partial: eqh_batchnorm_partial_floats(C) floats.
 * backward: dx, and dgamma / dbeta (optional; accumulate != 0 adds to their contents).  Deterministic.
 * ---------------------------------------------------------------------------------------- */
EQH_API int eqh_batchnorm_partial_floats(int64_t C);
EQH_API int eqh_batchnorm_fwd(const float* x, int64_t rows, int64_t n_stat, int64_t C, const float* gamma, const float* beta,
                      float eps, float momentum, float* y, float* save_mean, float* save_rstd, float* running_mean,
                      float* running_var, int64_t* num_batches_tracked, const int64_t* n_stat_dev, float* partial,
                      eqh_stream_t stream);
EQH_API int eqh_batchnorm_bwd(const float* x, const float* dy, int64_t rows, int64_t n_stat, int64_t C, const float* gamma,
                      const float* save_mean, const float* save_rstd, float* dx, float* dgamma, float* dbeta, int accumulate,
                      const int64_t* n_stat_dev, float* partial, eqh_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Strided / fused-epilogue variants used by the hand-scheduled layer Functions (layers/fused_*.py), so that concat,
 * slice, activation and residual-add launches of the reference's autograd graph (egnn_layer.py:357-362, mlp.py:94-96)
 * never exist as separate kernels.  Row strides are in floats and must be multiples of 4.
 * ---------------------------------------------------------------------------------------- */
EQH_API int eqh_rowln_fwd_ld(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, int relu, int64_t rows,
                     int64_t C, float* y, int64_t ldy, eqh_stream_t stream);
/* dx = (add ? add : 0) + LayerNorm/ReLU backward; add has the row stride of dx and may alias it */
EQH_API int eqh_rowln_bwd_ld(const float* x, int64_t ldx, const float* gamma, float eps, int relu, const float* dy, int64_t lddy,
                     int64_t rows, int64_t C, float* dx, int64_t lddx, const float* add, float* dgb_partial,
                     eqh_stream_t stream);
EQH_API int eqh_gather_reduce_ld(const float* table, int64_t ld_table, const int32_t* idx, const int32_t* rowptr,
                         const int32_t* src_rowptr, int64_t n_out, int64_t C, int mode, float* out, int64_t ld_out,
                         eqh_stream_t stream);
EQH_API int eqh_egnn_edge_fwd_tc_ld(const float* P, const float* Q, int64_t ldpq, const float* wc, const float* W2,
                            const float* b2, const int32_t* idx, const float* d2, int64_t N, int64_t H, float* S2, float* Mij,
                            float* Mi, int64_t ldmi, const int32_t* order, eqh_stream_t stream);
EQH_API int eqh_egnn_edge_bwd_tc_ld(const float* P, const float* Q, int64_t ldpq, const float* wc, const float* W2,
                            const int32_t* idx, const float* d2, const float* S2, const float* dMi, int64_t lddmi, int64_t N,
                            int64_t H, float* dP, int64_t lddp, float* dPre, float* part, const int32_t* order,
                            eqh_stream_t stream);
/* phi_e backward with the dQ reduction started INSIDE the kernel ("slot mode"; egnn_layer.py:298-310 autograd, the
 * scatter of d h_j): instead of one dPre row per kNN edge (16 N rows written, then read back by the segmented sum over the
 * transposed neighbour lists) every (sub-slice of the visiting order, neighbour j) pair owns ONE row that its edges
 * accumulate into in program order -- deterministic, no atomics.  With the atoms in grid-cell order a sub-slice meets each
 * neighbour 5-6 times (measured on Molecule3D- / PCQM4Mv2-shaped batches), so the buffer and the reduction shrink by that
 * factor.  Slots are numbered j-major: dQ[j] = sum of the consecutive rows slot_ext[j*U] .. slot_ext[(j+1)*U].
 *   U = eqh_egnn_edge_slot_keys(H);  mark [N*U] int32 filled with INT32_MAX by the caller;
 *   eqh_egnn_edge_slots_mark -> mark;  slot_ext [N*U+1] = exclusive prefix sum of (mark != INT32_MAX) (caller);
 *   eqh_egnn_edge_slots_assign -> slot_edge [N*16];  eqh_egnn_edge_bwd_tc_slots(..., slot_edge): dPre receives the slot
 *   rows (at most N*16), everything else as eqh_egnn_edge_bwd_tc_ld.  slot_edge == NULL: one row per edge. */
EQH_API int eqh_egnn_edge_slot_keys(int64_t H);
EQH_API int eqh_egnn_edge_slots_mark(const int32_t* idx, const int32_t* order, int64_t N, int64_t H, int32_t* mark,
                             eqh_stream_t stream);
EQH_API int eqh_egnn_edge_slots_assign(const int32_t* idx, const int32_t* order, int64_t N, int64_t H, const int32_t* mark,
                               const int32_t* slot_ext, int32_t* slot_edge, eqh_stream_t stream);
EQH_API int eqh_egnn_edge_bwd_tc_slots(const float* P, const float* Q, int64_t ldpq, const float* wc, const float* W2,
                               const int32_t* idx, const float* d2, const float* S2, const float* dMi, int64_t lddmi,
                               int64_t N, int64_t H, float* dP, int64_t lddp, float* dPre, float* part, const int32_t* order,
                               const int32_t* slot_edge, eqh_stream_t stream);

/* gradients of phi_e's packed weights back into the parameters' own layout, accumulating (inverse of eqh_egnn_pack):
 * T [2 Hp, C] = [dP | dQ]^T . h, red = (dW2 [16, Hp] | dwc [Hp] | db2 [16]) */
EQH_API int eqh_egnn_unpack_grads(const float* T, const float* red, int64_t H0, int64_t C, int64_t Hp, float* gW1, float* gW2,
                          float* gb2, eqh_stream_t stream);
/* eqh_gemm_tf32x3 with two optional epilogue extras (splits must be 1): stash [M, ldd] receives the pre-activation
 * A.B^T + bias * rowscale (act = SiLU keeps what its backward needs); dsilu_aux [M, ld_aux]: the result is multiplied by
 * SiLU'(dsilu_aux[m, n]) before the residual is added (SiLU backward fused into the dgrad GEMM). */
EQH_API int eqh_gemm_tf32x3_ex(const float* A, int a_kmajor, int64_t lda, const float* B, int b_kmajor, int64_t ldb, float* D,
                       int64_t ldd, int64_t M, int64_t N, int64_t K, const float* bias, const float* rowscale, int act,
                       const float* residual, float out_scale, float res_scale, int splits, float* ws, int32_t* counters,
                       float* stash, const float* dsilu_aux, int64_t ld_aux, eqh_stream_t stream);

/* ... and with a second bias [N] added unscaled after the activation and out_scale:
 *   D = out_scale * act(A.B^T + bias * rowscale) + bias2 + res_scale * residual
 * (two Linears merged into one GEMM: the inner bias carries the row scale / mix factor, the outer one does not). */
EQH_API int eqh_gemm_tf32x3_ex2(const float* A, int a_kmajor, int64_t lda, const float* B, int b_kmajor, int64_t ldb, float* D,
                        int64_t ldd, int64_t M, int64_t N, int64_t K, const float* bias, const float* rowscale, int act,
                        const float* residual, float out_scale, float res_scale, int splits, float* ws, int32_t* counters,
                        float* stash, const float* dsilu_aux, int64_t ld_aux, const float* bias2, eqh_stream_t stream);
/* ... and with the ReLU backward of the consumer fused in (the activation between two MHNNSConv applications,
 * equihnn_egnn.py:164-166 / mhnn.py:128-130): elements whose relu_mask[m, n] (the ReLU's OUTPUT, [M, ld_mask]) is not
 * positive are zeroed after the residual is added:  D = relu_mask > 0 ? (... + res_scale * residual) : 0. */
EQH_API int eqh_gemm_tf32x3_ex3(const float* A, int a_kmajor, int64_t lda, const float* B, int b_kmajor, int64_t ldb, float* D,
                        int64_t ldd, int64_t M, int64_t N, int64_t K, const float* bias, const float* rowscale, int act,
                        const float* residual, float out_scale, float res_scale, int splits, float* ws, int32_t* counters,
                        float* stash, const float* dsilu_aux, int64_t ld_aux, const float* bias2, const float* relu_mask,
                        int64_t ld_mask, eqh_stream_t stream);
/* out[m, n] += u[m] * v[n]  (rank-1 update of a row-strided [M, N] matrix: the bias term of a merged weight product) */
EQH_API int eqh_rank1_update(float* out, int64_t ld, const float* u, const float* v, int64_t M, int64_t N, float scale,
                     eqh_stream_t stream);

/* Several independent GEMMs with the same operand majors as ONE launch (the weight gradients dW = dY^T . X of all the
 * Linears of a layer, mlp.py:91-99 / conv.py:169-182 backward): D_i = out_scale_i * (A_i . B_i^T [+ bias_i * rowscale_i])
 * + res_scale_i * residual_i, operand conventions of eqh_gemm_tf32x3.  The planner picks one tile width for the group and
 * a split-K count per problem (partials in ws, finished in fixed split order by the last-arriving CTA of a tile:
 * deterministic); a problem that is split must not carry a bias.  eqh_gemm_grouped_plan returns the tile width, the
 * per-problem split counts (splits[n]) and the workspace (floats) / zeroed int32 counters the launch needs.
 * n <= 8. */
typedef struct eqh_gemm_problem {
  const float* A; int64_t lda;
  const float* B; int64_t ldb;
  float* D; int64_t ldd;
  int64_t M, N, K;
  const float* bias; const float* rowscale; const float* residual;
  float out_scale, res_scale;
} eqh_gemm_problem;
EQH_API int eqh_gemm_grouped_plan(const eqh_gemm_problem* problems, int n, int a_kmajor, int b_kmajor, int flags, int* bn,
                          int* splits, int64_t* ws_floats, int64_t* n_counters);
EQH_API int eqh_gemm_grouped(const eqh_gemm_problem* problems, int n, int a_kmajor, int b_kmajor, int flags, float* ws,
                     int32_t* counters, eqh_stream_t stream);

/* The same step with the step count t-1 held in DEVICE memory (int64), for an optimizer that lives inside a captured
 * CUDA graph: every block derives the bias corrections of step t = *step_dev + 1 in double precision; with advance != 0
 * the last block to finish stores t to *step_dev (counter: one zeroed int32, left zero).  An optimizer step split over
 * several launches (one per gradient bucket, each after its own all-reduce; main.py:281 DDP buckets) passes advance = 0
 * for all but the last launch. */
EQH_API int eqh_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                      float beta1, float beta2, float eps, float weight_decay, int64_t* step_dev, int advance,
                      int32_t* counter, float grad_scale, eqh_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* EQH_B200_H_ */
