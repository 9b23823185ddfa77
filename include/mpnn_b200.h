/* mpnn_b200.h -- C ABI of libmpnn_b200.so: hochshi/mpnn's message-passing hot path on B200 (sm_100a).
 *
 * The reference has no native code and no FFI: its hot path is a set of PyTorch nn.Modules
 * (mpnn_functions/..., models/mask_batch_norm.py) that call aten ops.  This ABI is the boundary a
 * maintainer binds those modules' forward/backward to (INTEGRATION.md shows the ctypes stubs); each entry
 * point below cites the reference lines it replaces.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to contiguous row-major fp32 (float) or int32 (int) data, except the
 *    `float* const*` weight-pointer arrays, which are HOST arrays of device pointers;
 *  - the caller owns all memory (inputs, outputs, saved-for-backward buffers, workspaces); the library never
 *    allocates device memory and keeps no global mutable state beyond a thread-local error string;
 *  - every call only enqueues work on `stream` and returns; it never synchronises;
 *  - return value 0 = success, < 0 = error (text via mpnn_last_error()); nothing aborts or exits;
 *  - gradients are WRITTEN (not accumulated) unless stated otherwise;
 *  - all reductions have a fixed order: results are bit-reproducible run to run.
 *  - `*_workspace_bytes` give the scratch size for BOTH the _fwd and _bwd call of an op.
 */
#ifndef MPNN_B200_H
#define MPNN_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* mpnn_stream_t; /* == cudaStream_t */

#define MPNN_OK 0
#define MPNN_ERR_ARG -1
#define MPNN_ERR_UNSUPPORTED -2
#define MPNN_ERR_CUDA -3
#define MPNN_ERR_WORKSPACE -4

int mpnn_version(void);            /* major*10000 + minor*100 + patch */
/* p[0:bytes] <- 0 on `stream` (cudaMemsetAsync: a memset node under graph capture, not a kernel) */
int mpnn_zero_bytes(void* p, size_t bytes, cudaStream_t stream);
/* Precision of the widths 33..256: 1 (default) = tcgen05 kernels with TF32 operands and fp32 accumulation (SURVEY 8c:
 * <= 2e-2 relative after the GRU / readout); 0 = the fp32 kernels everywhere (fp32 accuracy, a fraction of the speed).
 * Process-wide; returns the previous setting. */
int mpnn_set_tensor_cores(int enabled);
int mpnn_tensor_cores_enabled(void);
const char* mpnn_last_error(void); /* thread-local, valid until the next failing call on this thread */

/* ---- generic building blocks ------------------------------------------------------------------------ */
/* out[r,:] (+)= scale * sum_{k in ptr[r]:ptr[r+1]} src[idx ? idx[k] : k, :]   (CSR/CSC gather-sum) */
int mpnn_segment_sum(const float* src, const int* ptr, const int* idx, int rows, int width, long long lds, float* out,
                     long long ldo, int accumulate, float scale, mpnn_stream_t stream);
size_t mpnn_colsum_workspace_bytes(long long rows, int width);
/* out[c] (+)= sum_r X[r,c] * (Y ? Y[r,c] : 1) */
int mpnn_colsum(const float* X, const float* Y, long long rows, int width, long long ldx, long long ldy, float* out,
                int accumulate, void* workspace, size_t workspace_bytes, mpnn_stream_t stream);
size_t mpnn_gemm_workspace_bytes(int M, int N, int K);
/* C[m,n] = sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn] (+ bias[n]); flags: 1 = ReLU, 2 = accumulate into C.
 * Replaces the small aten::mm/addmm calls of the path (gru_update.py:27-28, graph_level_output.py:36,
 * set2vec.py:69-72,128, the growth layers of edge_network.py:15-19). */
int mpnn_gemm(const float* A, const float* B, float* C, int M, int N, int K, long long sam, long long sak,
              long long sbk, long long sbn, long long ldc, const float* bias, int flags, void* workspace,
              size_t workspace_bytes, mpnn_stream_t stream);

/* ---- a0: padded batch -> edge list (layout produced by pre_process/data_loader.py:50-70) -------------- */
size_t mpnn_compact_workspace_bytes(int B, int N);
/* Pass 1: edge predicate (adj != 0 or any bfm != 0; adj may be NULL), per-row / per-column counts and their
 * exclusive scans.  row_ptr[B*N] (device) = number of edges E.  Order = torch.nonzero order (bit-exact). */
int mpnn_compact_count(const float* bfm, const float* adj, int B, int N, int ef, int* row_ptr, int* col_ptr,
                       void* workspace, size_t workspace_bytes, mpnn_stream_t stream);
/* Pass 2: edge_src/edge_dst (flat node ids), edge_w (adj value), edge_x [capacity(+1), ef] bond rows,
 * csc_eid (edge ids grouped by sender).  `workspace` must be the buffer pass 1 filled. */
int mpnn_compact_fill(const float* bfm, const float* adj, int B, int N, int ef, const int* row_ptr, const int* col_ptr,
                      int capacity, int* edge_src, int* edge_dst, float* edge_w, float* edge_x, int* csc_eid,
                      const void* workspace, mpnn_stream_t stream);
/* capacity mode: clamp row_ptr / col_ptr [n_rows+1] to the allocated edge slots in place (consumers walk these ranges);
 * e_true [2 ints]: e_true[0] receives the true edge count */
int mpnn_compact_clamp(int* row_ptr, int* col_ptr, int n_rows, int capacity, int* e_true, mpnn_stream_t stream);
/* backward of the bond-row gather: dense[b,i,j,:] = d_edge_x[e,:] (dense is pre-zeroed by the caller) */
int mpnn_scatter_edge_rows(const float* d_edge_x, const int* edge_dst, const int* edge_src, int E, int N, int ef,
                           float* dense, mpnn_stream_t stream);

/* a0, the step before the path: device-side collate.  The host ships the batch ragged (the real
 * atoms' feature rows + the edge list, ~6x..12x fewer bytes than the padded tensors) and this call writes the
 * reference's padded layout (pre_process/data_loader.py:50-70).  atom_row[a] = b*N+i, edge_dst[e] = b*N+i, edge_j[e] = j;
 * the outputs are zero-filled here. */
int mpnn_collate_ragged(const int* atom_row, const float* afm_cat, long long n, int Fa, const int* edge_dst,
                        const int* edge_j, const float* edge_w, const float* edge_x, long long E, int ef, int B, int N,
                        float* afm, float* bfm, float* adj, float* mask, mpnn_stream_t stream);

/* ---- a0 (cont.): exact de-duplication of the compacted bond rows + stable grouping of edges by distinct row.
 * Bond features are categorical (reference mol_graph/mol_graph.py:74-90), and edge_map is a pure function of the
 * row (edge_network.py:14-21,36-37), so evaluating it once per distinct row is exact.  Rows are compared by bit
 * pattern; distinct rows are numbered by first occurrence.  Nothing is read back to the host. */
size_t mpnn_dedup_workspace_bytes(int edge_capacity, int unique_capacity);
/* n_edges_ptr: DEVICE pointer to the edge count (row_ptr + B*N).  urows [unique_capacity+1, ef] must be pre-zeroed
 * (rows >= U stay zero; row unique_capacity is the zero bond row x_0).  counts [4] = {E, U, overflow, 0}. */
int mpnn_dedup_rows(const float* rows, const int* n_edges_ptr, int edge_capacity, int ef, int unique_capacity,
                    int* uid, float* urows, int* counts, int sort, int* type_ptr, int* type_eid, int* type_pos,
                    void* workspace, size_t workspace_bytes, mpnn_stream_t stream);
size_t mpnn_type_sort_workspace_bytes(int edge_capacity, int unique_capacity);
int mpnn_type_sort(const int* uid, const int* counts, int edge_capacity, int unique_capacity, int* type_ptr,
                   int* type_eid, int* type_pos, void* workspace, size_t workspace_bytes, mpnn_stream_t stream);

/* ---- a1-a3, a5 on the distinct rows ("typed" path, mpnn_b200/csrc/typed.cu) -----------------------------------
 * table T[u][l][k] = edge_map(distinct row u).view(mf, nf)[k][l]  (edge_network.py:21,37), padded to DP x DP;
 * tableT is its transpose.  DP = mpnn_typed_dp(nf, mf) (power of two >= 8; -1: widths > 32, not served here). */
int mpnn_typed_dp(int nf, int mf);
int mpnn_graph_sum(const float* X, int B, int N, int width, float* out, mpnn_stream_t stream);
int mpnn_table_from_flat(const float* flat, int R, int nf, int mf, float* table, float* tableT, mpnn_stream_t stream);
int mpnn_table_to_flat(const float* dT, int R, int nf, int mf, float* dflat, mpnn_stream_t stream);
/* fused growth layers + n_tied tied layers + last Linear on R distinct rows (trunk width P <= 64) */
int mpnn_enet_supported(int ef, int n_growth, int P);
int mpnn_enet_max_dp(void); /* widest padded feature width (64) the fused kernel serves */
long long mpnn_enet_saved_floats(int R, int n_growth, int n_tied);
size_t mpnn_enet_workspace_bytes(int R, int ef, int n_growth, int P);
int mpnn_enet_fwd(const float* rows, int R, int ef, int n_growth, const float* const* growth_w,
                  const float* const* growth_b, const float* w_tied, int P, int n_tied, const float* w_last,
                  const float* b_last, int nf, int mf, float* saved, float* table, float* tableT,
                  mpnn_stream_t stream);
int mpnn_enet_bwd(const float* rows, int R, int ef, int n_growth, const float* const* growth_w, const float* w_tied,
                  int P, int n_tied, const float* w_last, int nf, int mf, const float* saved, const float* dT,
                  float* const* d_growth_w, float* const* d_growth_b, float* d_w_tied, float* d_w_last,
                  float* d_b_last, float* d_rows, void* workspace, size_t workspace_bytes, mpnn_stream_t stream);
/* message function + aggregation as a gather over the CSR (edge_network.py:42-52 + adjacent_message_agg.py:18).
 * S != NULL ([B, nf] per-graph sums of H) selects the HEAD form (edge_network.py:50-51: all pairs + beta). */
size_t mpnn_tmsg_bwd_workspace_bytes(int edge_capacity, int unique_capacity, int nf, int mf, int B);
int mpnn_tmsg_fwd(const int* row_ptr, const int* edge_src, const int* uid, const float* alpha, const float* H,
                  const float* table, const float* S, const float* beta, int n_rows, int N, int nf, int mf,
                  int zero_type, float* M, mpnn_stream_t stream);
/* n_src_rows: rows of H (0 = n_rows: node states gathered through edge_src; the edge count when H holds one explicit
 * sender vector per edge -- the gated states of AttEdgeNetwork, att_edge_network.py:26 -- with identity edge_src /
 * col_ptr / csc_eid) */
int mpnn_tmsg_bwd(const int* row_ptr, const int* col_ptr, const int* csc_eid, const int* edge_src, const int* edge_dst,
                  const int* uid, const int* type_ptr, const int* type_eid, const int* counts, const float* alpha,
                  const float* H, const float* table, const float* tableT, const float* S, int n_rows, int n_src_rows,
                  int B, int N, int nf, int mf, int edge_capacity, int unique_capacity, const float* dM, float* dH,
                  float* dT, void* workspace, size_t workspace_bytes, mpnn_stream_t stream);

/* ---- a1-a3, a5 on the distinct rows, feature widths 33..256: tcgen05 tensor cores (mpnn_b200/csrc/tc_message.cu) ----
 * The type-sorted edge list is cut into single-type tiles of <= 128 edges (the plan, device side, no host read);
 * the message function (edge_network.py:42-52) is then a grouped TF32 GEMM with the accumulator in TMEM:
 *   Y[e, 0:N] = alpha_e * Bm[uid_e] (N x K, K contiguous, zero padded to DP x DP) . A[gidx[e], 0:K]
 * forward: A = H, rows = edge_src, Bm = tableT, K = nf, N = mf; backward (d sender states): A = dM, rows = edge_dst,
 * Bm = table, K = mf, N = nf.  The aggregation (adjacent_message_agg.py:18) is mpnn_segment_sum over row_ptr /
 * (col_ptr, csc_eid).  mpnn_tc_table_grad: dT[u][l][k] = sum_{e of type u} alpha_e H[src_e, l] dM[dst_e, k].
 * mpnn_tc_dp: 64 / 128 / 256, or -1 when the shape is not served (widths <= 32, > 256 or not multiples of 4). */
int mpnn_tc_dp(int nf, int mf);
size_t mpnn_tc_plan_bytes(int edge_capacity, int unique_capacity);
/* plan = tiles + per-position gather rows (edge_src / edge_dst) and weights (edge_w, NULL = 1) of the sorted list */
int mpnn_tc_plan(const int* type_ptr, const int* type_eid, const int* edge_src, const int* edge_dst,
                 const float* edge_w, int edge_capacity, int unique_capacity, void* plan, size_t plan_bytes,
                 mpnn_stream_t stream);
size_t mpnn_tc_edge_gemm_workspace_bytes(int unique_capacity, int DP);
/* use_dst: 0 = rows of A are the senders (forward), 1 = the receivers (backward); use_alpha: scale by the plan's
 * edge weights; workspace holds the pre-swizzled shared-memory image of Bm (one bulk copy per pipeline stage) */
int mpnn_tc_edge_gemm(const void* plan, int edge_capacity, int unique_capacity, const int* type_eid, int use_dst,
                      const float* A, int lda, int K, const float* Bm, int DP, int use_alpha, float* Y, int ldy,
                      int N, void* workspace, size_t workspace_bytes, mpnn_stream_t stream);
size_t mpnn_tc_table_grad_workspace_bytes(int unique_capacity, int DP);
int mpnn_tc_table_grad(const void* plan, int edge_capacity, int unique_capacity, const float* H, int nf,
                       const float* dM, int mf, int DP, int use_alpha, float* dT, void* workspace,
                       size_t workspace_bytes, mpnn_stream_t stream);

/* Dense GEMMs on the same tcgen05 kernels (identity plan): the GRU gate products (gru_update.py:27-28) and the
 * readout projections (graph_level_output.py:36) at widths 33..256.  TF32 operands, fp32 accumulate.
 *   Y[r, g*ycol + n] (+)= sum_{s<kseg} sum_{k<K} A[r, s*acol + k] * W[n*w_sn + k*w_sk + g*w_sg + s*w_ss] + bias[g*N + n]
 *   out[g*o_sg + l*o_sl + k] = sum_r X[r, l] * D[r, g*dcol + k]                 (weight gradients X^T D) */
size_t mpnn_tc_dense_workspace_bytes(int n_blocks, int DP);
int mpnn_tc_dense_gemm(const float* A, long long rows, int lda, int K, int kseg, int acol, const float* W,
                       long long w_sn, long long w_sk, long long w_sg, long long w_ss, int G, int N, const float* bias,
                       float* Y, int ldy, int ycol, int accumulate, int DP, void* workspace, size_t workspace_bytes,
                       mpnn_stream_t stream);
/* mpnn_tc_dense_gemm with a 64-bit offset (floats, multiple of 4) between output blocks that may be different buffers
 * (the GRU backward writes dm and dh): the G blocks, or -- nsplit > 0, G = 1 -- every nsplit columns of one product */
int mpnn_tc_dense_gemm_ll(const float* A, long long rows, int lda, int K, int kseg, int acol, const float* W,
                          long long w_sn, long long w_sk, long long w_sg, long long w_ss, int G, int N, const float* bias,
                          float* Y, int ldy, long long ycol, int nsplit, int accumulate, int DP, void* workspace,
                          size_t workspace_bytes, cudaStream_t stream);
/* GRU weight gradients for widths <= 64 in ONE pass over the gate gradients (gru_update.py:27-28 backward):
 * dW_ih [d,3d] = m^T (dar|daz|dan), dW_hh [d,3d] = h^T (dar|daz|dnh); dg [rows, ldg] holds the blocks dar|daz|dan|dnh. */
size_t mpnn_tc_gru_param_workspace_bytes(void);
int mpnn_tc_gru_param_grad(const float* m, const float* h, const float* dg, int ldg, long long rows, int d,
                           float* dW_ih, float* dW_hh, void* workspace, size_t workspace_bytes, cudaStream_t stream);
/* The same with the pointwise GRU backward folded in: reads the saved gates [rows, 4d], m, h, dh', mask once; writes
 * dg [rows, 6d] when non-NULL (dar | daz | dan | dnh | hi(go z) | lo(go z)), the bias partials
 * [mpnn_tc_gru_param_bias_parts()][4d] and dW_ih, dW_hh. */
/* GRU data gradients for widths <= 64 straight from the saved gates (no gate-gradient array in HBM):
 * (dm | dh) = (dar|daz|dan|dnh|hi(go z)|lo(go z)) x Wc, Wc = combined weights [6][2d][d] (device). */
/* profiling aid: DEVICE buffer of 64 x uint64 receiving %globaltimer stamps of k_tc_gru_data_grad's producer; NULL = off */
void mpnn_tc_debug(unsigned long long* buf);
size_t mpnn_tc_gru_data_workspace_bytes(void);
int mpnn_tc_gru_data_grad(const float* gates, const float* h, const float* dh_out, const float* mask, const float* Wc,
                          long long rows, int d, float* dm, float* dh, void* workspace, size_t workspace_bytes,
                          mpnn_stream_t stream);
int mpnn_tc_gru_param_bias_parts(void);
int mpnn_tc_gru_param_point(const float* m, const float* h, const float* mask, const float* gates, const float* dh_out,
                            long long rows, int d, float* dg, float* bias_part, float* dW_ih, float* dW_hh,
                            void* workspace, size_t workspace_bytes, mpnn_stream_t stream);
size_t mpnn_tc_dense_grad_workspace_bytes(int G, int DP);
int mpnn_tc_dense_gemm_tn(const float* X, long long rows, int ldx, int M, const float* D, int ldd, int dcol, int G,
                          int N, int DP, float* out, long long o_sg, long long o_sl, void* workspace,
                          size_t workspace_bytes, mpnn_stream_t stream);

/* Fused masked GRU forward (gru_update.py:26-35,66-68) for widths 33..256 (256: two 128-column blocks): both gate products accumulate in TMEM and
 * the gate arithmetic runs in the epilogue, so the [rows, 3d] pre-activations never reach HBM.  mpnn_gru_fwd uses it. */
int mpnn_tc_gru_supported(int d);
size_t mpnn_tc_gru_workspace_bytes(int d);
int mpnn_tc_gru_fwd(const float* m, const float* h, const float* mask, const float* W_ih, const float* W_hh,
                    const float* b_ih, const float* b_hh, long long rows, int d, float* h_out, float* gates,
                    void* workspace, size_t workspace_bytes, mpnn_stream_t stream);
int mpnn_tc_gru_fwd_agg(const float* Y, const int* row_ptr, const float* m, const float* h, const float* mask,
                        const float* W_ih, const float* W_hh, const float* b_ih, const float* b_hh, long long rows,
                        int d, float* m_out, float* h_out, float* gates, void* workspace, size_t workspace_bytes,
                        mpnn_stream_t stream);

/* nn.Linear-shaped wrappers (W [N, K] row-major; K, N multiples of 4 up to 1024, cut into blocks of <= 256) */
int mpnn_tc_linear_supported(int K, int N);
size_t mpnn_tc_linear_workspace_bytes(int K, int N);
int mpnn_tc_linear_fwd(const float* X, long long rows, int ldx, int K, const float* W, int N, const float* bias,
                       float* Y, int ldy, int accumulate, void* workspace, size_t workspace_bytes,
                       mpnn_stream_t stream);
int mpnn_tc_linear_bwd_data(const float* dY, long long rows, int ldd, int N, const float* W, int K, float* dX, int ldx,
                            int accumulate, void* workspace, size_t workspace_bytes, mpnn_stream_t stream);
int mpnn_tc_linear_bwd_weight(const float* dY, long long rows, int ldd, int N, const float* X, int ldx, int K,
                              float* dW, void* workspace, size_t workspace_bytes, mpnn_stream_t stream);

/* ---- a1/a2: edge-network trunk = edge_map[:-1] (edge_network.py:14-21,36-37) on compacted rows -------- */
long long mpnn_edge_trunk_saved_floats(int R, int ef, int n_growth, int P, int n_tied, long long* x_offset, int* ldx);
size_t mpnn_edge_trunk_workspace_bytes(int R, int ef, int n_growth, int P);
/* growth_w[g] [in^2, in], growth_b[g] [in^2] (nn.Linear layout), w_tied [P, P]; x = saved + *x_offset, [R, *ldx] */
int mpnn_edge_trunk_fwd(const float* rows_in, int R, int ef, int n_growth, const float* const* growth_w,
                        const float* const* growth_b, const float* w_tied, int P, int n_tied, float* saved,
                        void* workspace, size_t workspace_bytes, mpnn_stream_t stream);
int mpnn_edge_trunk_bwd(const float* rows_in, int R, int ef, int n_growth, const float* const* growth_w,
                        const float* w_tied, int P, int n_tied, const float* saved, const float* dx, int lddx,
                        float* const* d_growth_w, float* const* d_growth_b, float* d_w_tied, float* d_rows_in,
                        void* workspace, size_t workspace_bytes, mpnn_stream_t stream);

/* ---- a3/a3'/a5-a7: message function fused with the neighbour aggregation ---------------------------------
 * (edge_network.py:42-52 + adjacent_message_agg.py:18 / weighted_adjacent_message_agg.py:20 /
 *  attention_message_agg.py:24).  See mpnn_b200/csrc/message.cu for the algebra. */
long long mpnn_message_wt_floats(int nf, int mf, int P);
/* Wt <- transposed + bias-augmented copy of edge_map[-1] (W_last [mf*nf, P], B_last [mf*nf]) */
int mpnn_message_prepare(const float* W_last, const float* B_last, int nf, int mf, int P, float* Wt,
                         mpnn_stream_t stream);
int mpnn_message_fwd(const int* row_ptr, const int* edge_dst, const int* gidx, const int* xid, const float* alpha,
                     const float* X, int ldx, int x0_row, const float* Gsrc, int ldg, const float* Q, const float* Wt,
                     const float* beta, int nrows, int nf, int mf, int P, float* M, mpnn_stream_t stream);
size_t mpnn_message_bwd_workspace_bytes(int nrows, int nf, int mf, int P);
int mpnn_message_bwd(const int* row_ptr, const int* edge_dst, const int* gidx, const int* xid, const float* alpha,
                     const float* X, int ldx, int x0_row, const float* Gsrc, int ldg, const float* Q, const float* Wt,
                     const float* beta, int nrows, int n_edges, int nf, int mf, int P, const float* dM, float* T,
                     int ldt, float* dG, float* dQ, float* dalpha, float* dW_last, float* dB_last, float* dbeta,
                     void* workspace, size_t workspace_bytes, mpnn_stream_t stream);

/* ---- BiLiniearEdgeNetwork (message/bilinear_edge_network.py:25-37): parameter-free,
 * Y[e, p] = sum_{a,q} H[src_e, a] * X_e.view(nf,nf,nf)[a, p, q] * H[dst_e, q] on the compacted pairs (ef == nf^3). */
int mpnn_bilinear_fwd(const int* edge_src, const int* edge_dst, const float* X, long long ldx, const float* H,
                      long long E, int nf, float* Y, mpnn_stream_t stream);
int mpnn_bilinear_bwd(const int* row_ptr, const int* col_ptr, const int* csc_eid, const int* edge_src,
                      const int* edge_dst, const float* X, long long ldx, const float* H, const float* dY, int n_rows,
                      long long E, int nf, float* dH, float* dX, long long lddx, mpnn_stream_t stream);

/* ---- a4: attention gate of AttEdgeNetwork (att_edge_network.py:18-26) / row softmax -------------------- */
int mpnn_softmax_mul_fwd(const float* logits, const float* V, long long rows, int n, float* gate, float* out,
                         mpnn_stream_t stream);
int mpnn_softmax_mul_bwd(const float* gate, const float* V, const float* dout, long long rows, int n, float* dlogits,
                         float* dV, mpnn_stream_t stream);

/* ---- a5-a7 on dense [B,N,N,mf] messages (the aggregators' stand-alone contract) ------------------------ */
int mpnn_dense_agg_fwd(const float* messages, const float* weights, long long R, int N, int mf, float* out,
                       mpnn_stream_t stream);
int mpnn_dense_agg_bwd(const float* messages, const float* weights, const float* dout, long long R, int N, int mf,
                       float* dmessages, float* dweights, mpnn_stream_t stream);

/* ---- a8/a9: masked GRU update (gru_update.py:26-35,66-68); gates [rows, 4d] saved for backward --------- */
size_t mpnn_gru_workspace_bytes(long long rows, int d);
int mpnn_gru_fwd(const float* m, const float* h, const float* mask, const float* W_ih, const float* W_hh,
                 const float* b_ih, const float* b_hh, long long rows, int d, float* h_out, float* gates,
                 void* workspace, size_t workspace_bytes, mpnn_stream_t stream);
/* 1 (default): at widths 33..64 mpnn_gru_bwd runs the pointwise pass inside the weight-gradient kernel's producers
 * (mpnn_tc_gru_param_point); 0: as a separate launch, like the wider tensor-core widths.  Returns the previous setting. */
int mpnn_gru_bwd_one_pass(int enabled);
/* GRU forward with the aggregation (adjacent_message_agg.py:18) folded into the kernel's operand producer, tensor-core
 * widths only: message row i = sum of Y[row_ptr[i] : row_ptr[i+1], :] (per-edge messages in CSR order, summed in edge
 * order); the sums are also written to m_out [rows, d], which mpnn_gru_bwd reads as `m`. */
int mpnn_gru_agg_supported(int d);
int mpnn_gru_fwd_agg(const float* Y, const int* row_ptr, const float* h, const float* mask, const float* W_ih,
                     const float* W_hh, const float* b_ih, const float* b_hh, long long rows, int d, float* m_out,
                     float* h_out, float* gates, void* workspace, size_t workspace_bytes, mpnn_stream_t stream);
int mpnn_gru_bwd(const float* m, const float* h, const float* mask, const float* W_ih, const float* W_hh,
                 const float* gates, const float* dh_out, long long rows, int d, float* dm, float* dh, float* dW_ih,
                 float* dW_hh, float* db_ih, float* db_hh, void* workspace, size_t workspace_bytes,
                 mpnn_stream_t stream);
/* Shared-parameter form (one GRUCell applied at every step, basic_model.py:50-58): each step's backward leaves its
 * per-CTA weight-gradient partials in a caller-provided slab (mpnn_gru_bwd_partial_bytes each; 0 = width not served),
 * ONE reduction over all slabs gives the four parameter gradients of the whole T-step loop. */
size_t mpnn_gru_bwd_partial_bytes(long long rows, int d);
int mpnn_gru_bwd_data(const float* m, const float* h, const float* mask, const float* W_ih, const float* W_hh,
                      const float* gates, const float* dh_out, long long rows, int d, float* dm, float* dh,
                      float* partial, mpnn_stream_t stream);
int mpnn_gru_bwd_params(const float* partial, int slabs, long long rows, int d, float* dW_ih, float* dW_hh,
                        float* db_ih, float* db_hh, mpnn_stream_t stream);

/* ---- a10/a11: masked batch norms (models/mask_batch_norm.py:9-15, 20-38); stats [2C+1] saved ------------
 * Workspace contract: zero-filled ONCE by the caller; every call leaves it reusable (its completion counter is
 * reset by the last block), so consecutive calls on one stream need no memset. */
size_t mpnn_bn_workspace_bytes(long long rows, int C);
int mpnn_mask_bn_fwd(const float* x, const float* mask, long long rows, int C, float eps, float* y, float* stats,
                     void* workspace, size_t workspace_bytes, mpnn_stream_t stream);
int mpnn_mask_bn_bwd(const float* x, const float* mask, const float* dy, const float* stats, long long rows, int C,
                     float* dx, void* workspace, size_t workspace_bytes, mpnn_stream_t stream);
/* Masked batch norms of a categorical bond tensor in row space (graph.TypedBonds): x [R,F] distinct rows, a [R] their mask
 * (adjacency) values, cnt [R] their multiplicities.  masked_mean 1 / eps_inside 0 = MaskBatchNorm1d (mask_batch_norm.py:20-38),
 * masked_mean 0 / eps_inside 1 = MaskBatchNorm (:9-15).  stats [3F+1] is saved for the backward; gamma, beta, running_*,
 * dx, dgamma, dbeta may be NULL. */
int mpnn_row_bn_fwd(const float* x, const float* a, const float* cnt, int R, int F, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, int masked_mean, int eps_inside, int training,
                    float momentum, float eps, float* y, float* stats, mpnn_stream_t stream);
int mpnn_row_bn_bwd(const float* x, const float* a, const float* cnt, int R, int F, const float* gamma, const float* stats,
                    const float* dy, int masked_mean, int eps_inside, int training, float* dx, float* dgamma,
                    float* dbeta, mpnn_stream_t stream);
int mpnn_mask_bn1d_fwd(const float* x, const float* mask, const float* weight, const float* bias, float* running_mean,
                       float* running_var, long long rows, int C, int training, float momentum, float eps, float* y,
                       float* stats, void* workspace, size_t workspace_bytes, mpnn_stream_t stream);
int mpnn_mask_bn1d_bwd(const float* x, const float* mask, const float* dy, const float* weight, const float* stats,
                       const float* running_mean, const float* running_var, long long rows, int C, int training,
                       float eps, float* dx, float* dweight, float* dbias, void* workspace, size_t workspace_bytes,
                       mpnn_stream_t stream);

/* ---- a12: GraphLevelOutput (readout/graph_level_output.py:30-47); mask NULL = the unmasked branch (:39) - */
size_t mpnn_glo_workspace_bytes(int B, int N, int F2, int O);
int mpnn_glo_fwd(const float* x, const float* mask, const float* Wi, const float* bi, const float* Wj, const float* bj,
                 int B, int N, int F2, int O, float* out, float* u, float* v, float* UV, void* workspace,
                 size_t workspace_bytes, mpnn_stream_t stream);
/* The fused masked backward (O <= 64, F2 <= 64) as two calls: the data half writes dx and per-CTA partials into the
 * workspace, the parameter half reduces them (fixed order); the caller may enqueue the second on another stream. */
int mpnn_glo_bwd_split_supported(int has_mask, int F2, int O);
int mpnn_glo_bwd_data(const float* x, const float* mask, const float* Wi, const float* Wj, const float* u, const float* v,
                      const float* dout, int B, int N, int F2, int O, float* dx, void* workspace, size_t workspace_bytes,
                      mpnn_stream_t stream);
int mpnn_glo_bwd_params(const void* workspace, int B, int N, int F2, int O, float* dWi, float* dbi, float* dWj, float* dbj,
                        mpnn_stream_t stream);
int mpnn_glo_bwd(const float* x, const float* mask, const float* Wi, const float* Wj, const float* u, const float* v,
                 const float* UV, const float* dout, int B, int N, int F2, int O, float* dx, float* dWi, float* dbi,
                 float* dWj, float* dbj, void* workspace, size_t workspace_bytes, mpnn_stream_t stream);

/* K sibling edge networks (one per message-passing step, normed_basic_model.py:24-27: same layer plan, same distinct
 * rows, different weights) in ONE launch each way.  Per-network arguments are HOST arrays of device pointers:
 * growth_w / growth_b / d_growth_* [K * n_growth] (network-major), the others [K]. */
int mpnn_enet_max_nets(void);
int mpnn_enet_fwd_multi(int K, const float* rows, int R, int ef, int n_growth, const float* const* growth_w,
                        const float* const* growth_b, const float* const* w_tied, int P, int n_tied,
                        const float* const* w_last, const float* const* b_last, int nf, int mf, float* const* saved,
                        float* const* table, float* const* tableT, mpnn_stream_t stream);
int mpnn_enet_bwd_multi(int K, const float* rows, int R, int ef, int n_growth, const float* const* growth_w,
                        const float* const* w_tied, int P, int n_tied, const float* const* w_last, int nf, int mf,
                        const float* const* saved, const float* const* dT, float* const* d_growth_w,
                        float* const* d_growth_b, float* const* d_w_tied, float* const* d_w_last,
                        float* const* d_b_last, float* const* d_rows, void* workspace, size_t workspace_bytes,
                        mpnn_stream_t stream);
/* table gradients of K steps that share the edge list and the sender states: dM [K][n_rows][mf] ->
 * dT [K][unique_capacity+1][DP][DP]; workspace K x mpnn_tmsg_bwd_workspace_bytes(.., B = 1) */
int mpnn_tmsg_bwd_table_multi(int K, const int* edge_src, const int* edge_dst, const int* uid, const int* type_ptr,
                              const int* type_eid, const int* counts, const float* alpha, const float* H, int n_rows,
                              int nf, int mf, int edge_capacity, int unique_capacity, const float* dM, float* dT,
                              void* workspace, size_t workspace_bytes, mpnn_stream_t stream);

/* ---- a0 (small batches): compaction + exact de-duplication as ONE cooperative launch (csrc/prep.cu) ------------------
 * Same outputs as mpnn_compact_count/_fill + mpnn_dedup_rows in capacity mode: CSR/CSC of the edge set (row-major =
 * torch.nonzero order), edge_w = adj value, uid = distinct-row id by first occurrence (clamped to the zero type
 * `unique_capacity` on overflow), urows [unique_capacity+1, ef], counts = {E, U, overflow, 0} (the overflow flag is only
 * ever SET: sticky across replays).  The workspace must be zero on first use and is left zero. */
int mpnn_prep_supported(int B, int N, int ef, int unique_capacity);
size_t mpnn_prep_workspace_bytes(int B, int unique_capacity);
int mpnn_prep_edges(const float* bfm, const float* adj, int B, int N, int ef, int edge_capacity, int unique_capacity,
                    int* row_ptr, int* col_ptr, int* edge_src, int* edge_dst, float* edge_w, int* csc_eid, int* uid,
                    float* urows, int* counts, void* workspace, size_t workspace_bytes, mpnn_stream_t stream);

/* ---- x1: the whole T-step message-passing loop as one persistent kernel each way (feature widths <= 32) ---------
 * h <- bn_t(GRU(sum_{e in E(i)} alpha_e T_t[uid_e]^T H0[src_e], h) * mask) for t = 0..T-1: the loops of
 * models/normed_basic_model.py:56-59, basic_model.py:50-58, normed_encoded_basic_model_ecfp.py:67-69 on the typed
 * edge list (csrc/chain.cu).  bn_kind[t]: 0 none, 1 MaskBatchNorm (mask_batch_norm.py:9-15), 2 MaskBatchNorm1d (:20-38);
 * bn_kind / bn_training / bn_eps / bn_momentum are HOST arrays [T]; bn_ptrs is a HOST array [4T] of device pointers
 * (gamma, beta, running_mean, running_var per step, NULL allowed), tables a HOST array [T] of device pointers to the
 * per-type matrices T[u][l][k] of each step (equal pointers = shared edge network).  The first 256 bytes of the workspace
 * (barrier words) and the 2048 bytes behind them (mailboxes) must be zero on entry and are zero again on exit.
 * saved: mpnn_chain_saved_floats floats, read by the backward.
 * bwd writes dM [T][rows][d], dh_init [rows][d] (or NULL), the GRU cell's gradients and bn_grads[2t], [2t+1]
 * (d gamma, d beta of step t's MaskBatchNorm1d; HOST array of device pointers, NULL allowed). */
int mpnn_chain_supported(int d, int T);
int mpnn_chain_debug(long long* out64);   /* profiling aid: phase timestamps of the last forward (MPNN_B200_CHAIN_DEBUG=1) */
long long mpnn_chain_saved_floats(long long rows, int d, int T);
size_t mpnn_chain_workspace_bytes(long long rows, int d, int T);
int mpnn_chain_fwd(const int* row_ptr, const int* edge_src, const int* uid, const float* alpha, int ecap, int zero_type,
                   const float* H0, const float* h_init, const float* mask, const float* const* tables, int T,
                   const float* W_ih, const float* W_hh, const float* b_ih, const float* b_hh, const int* bn_kind,
                   const int* bn_training, const float* bn_eps, const float* bn_momentum, float* const* bn_ptrs,
                   long long rows, int d, const int* real_list, float* saved, float* out, void* workspace,
                   size_t workspace_bytes, mpnn_stream_t stream);
int mpnn_chain_bwd(const int* row_ptr, const int* edge_src, const int* uid, const float* alpha, int ecap, int zero_type,
                   const float* H0, const float* h_init, const float* mask, const float* const* tables, int T,
                   const float* W_ih, const float* W_hh, const float* b_ih, const float* b_hh, const int* bn_kind,
                   const int* bn_training, const float* bn_eps, const float* bn_momentum, float* const* bn_ptrs,
                   long long rows, int d, const int* real_list, float* saved, const float* dout, long long dout_ld,
                   float* dM, float* dh_init, float* dW_ih, float* dW_hh, float* db_ih, float* db_hh,
                   float* const* bn_grads, void* workspace, size_t workspace_bytes, mpnn_stream_t stream);
/* dout: gradient w.r.t. the last step's output, [rows, d] with a row stride of dout_ld >= d floats (the reference models
 * concatenate the final state with afm before the readout, normed_basic_model.py:59: the gradient arrives as a column
 * slice of a [rows, 2d] array and is read in place). */
/* real_list (optional, NULL allowed): [rows + 1] ints from mpnn_real_rows = the rows with mask != 0 in increasing order
 * and, in the last slot, their number; lets every CTA of the step kernels own the same number of real rows.  `out`,
 * `dM` and `dh_init` must be zero-filled by the caller: rows with mask == 0 are skipped (their values are exact zeros). */
int mpnn_real_rows_max(void);
int mpnn_real_rows(const float* mask, long long rows, int* list, void* workspace, size_t workspace_bytes,
                   mpnn_stream_t stream);

/* ---- a13/a14: Set2Vec with its input-less LSTM (readout/set2vec.py:68-75, 93-151) ----------------------- */
long long mpnn_set2vec_saved_floats(int B, int N, int F, int steps);
size_t mpnn_set2vec_workspace_bytes(int B, int N, int F);
size_t mpnn_set2vec_bwd_workspace_bytes(int B, int N, int F, int steps);
/* m0 [B,2F] / c0 [B,F]: caller-supplied initial LSTM state (set2vec.py:111-117: m0 = cat(mprev, 0)); NULL = zeros.
 * bwd writes dm0 [B,2F] / dc0 [B,F] when non-NULL. */
int mpnn_set2vec_fwd(const float* X, const float* mask, const float* Wcat, const float* bcat, const float* Wq,
                     const float* we, const float* m0, const float* c0, int B, int N, int F, int steps, float* out,
                     float* saved, void* workspace, size_t workspace_bytes, mpnn_stream_t stream);
int mpnn_set2vec_bwd(const float* X, const float* mask, const float* Wcat, const float* Wq, const float* we,
                     const float* m0, const float* c0, const float* saved, const float* dout, int B, int N, int F,
                     int steps, float* dX, float* dWcat, float* dbcat, float* dWq, float* dwe, float* dm0, float* dc0,
                     void* workspace, size_t workspace_bytes, mpnn_stream_t stream);
/* The loop of `steps` iterations runs in two persistent cooperative kernels (csrc/s2v_persist.cu: one CTA per group of
 * graphs, the batch-wide softmax statistics exchanged through tagged slots in L2) when F <= 64 and the batch fits;
 * other shapes, or after mpnn_set2vec_set_persistent(0), run six launches per iteration.  Returns the previous setting. */
int mpnn_set2vec_set_persistent(int enabled);
/* Profiling aid: CTA 0 of the persistent kernels writes %globaltimer stamps of the phases of its first four iterations
 * into buf (448 x uint64 of DEVICE memory: forward [0,64), backward [64,128), 16 per iteration; [128,448): post / gather-done time of every CTA in forward iteration 2); NULL = off. */
void mpnn_set2vec_debug(unsigned long long* buf);
/* LSTMCellHidden.forward alone (set2vec.py:68-75) on pre [B,4F] = hprev [w_hi|w_hf|w_hg|w_ho] + [b_*] (the caller's
 * mpnn_gemm): activated gates [B,4F], c' [B,F], tanh(c') [B,F], h' [B,F]; bwd: dh, dc' -> dpre [B,4F], dc_prev. */
int mpnn_lstm_hidden_fwd(const float* pre, const float* cprev, int B, int F, float* gates, float* c, float* tc,
                         float* h, mpnn_stream_t stream);
int mpnn_lstm_hidden_bwd(const float* gates, const float* tc, const float* cprev, const float* dh, const float* dc_next,
                         int B, int F, float* dpre, float* dc_prev, mpnn_stream_t stream);

/* ---- 8f rank 4: prediction head + loss of the drivers (test_graph_norm.py:86-90 nn.BatchNorm1d(out) ->
 * nn.Linear(out, targets) + nn.MSELoss), one single-CTA launch each way for problems that fit a CTA's shared memory
 * (mpnn_head_supported).  gamma/beta may be NULL (no affine); running_* / num_batches_tracked may be NULL (training).
 * fwd writes y [B,T], loss [1], stats [2C]; bwd takes the loss gradient as a DEVICE scalar and writes (not
 * accumulates) dx [B,C], dgamma/dbeta [C] (may be NULL), dW [T,C], db [T]. */
int mpnn_head_supported(int B, int C, int T);
int mpnn_head_bn_linear_mse_fwd(const float* x, const float* target, const float* gamma, const float* beta,
                                float* running_mean, float* running_var, long long* num_batches_tracked,
                                const float* W, const float* b, int B, int C, int T, int training, float momentum,
                                float eps, float* y, float* loss, float* stats, mpnn_stream_t stream);
int mpnn_head_bn_linear_mse_bwd(const float* x, const float* target, const float* gamma, const float* beta,
                                const float* W, const float* y, const float* stats, const float* gloss, int B, int C,
                                int T, int training, float* dx, float* dgamma, float* dbeta, float* dW, float* db,
                                mpnn_stream_t stream);

/* ---- 8f rank 4: the drivers' optimizer (torch.optim.Adam, test_lipo.py:138-139) over a list of small tensors as one
 * launch.  Host arrays of n device pointers / element counts; `step` = device float (steps taken so far, incremented
 * by the call: CUDA-graph replayable); `ticket` = device uint32, zero before the first call. */
int mpnn_adam_step(int n, float* const* params, const float* const* grads, float* const* exp_avg,
                   float* const* exp_avg_sq, const long long* numel, float* step, unsigned int* ticket, float lr,
                   float beta1, float beta2, float eps, float weight_decay, const int* const* guards, int n_guards,
                   mpnn_stream_t stream);
/* 8e: the gradient all-reduce fused into the Adam step over NVLink peer memory (one launch, csrc/optim.cu k_adam_ddp).
 * flat / flags: HOST arrays of `world` peer-mapped device pointers into every rank's symmetric buffer
 * ([2][region_floats] gradient regions + [world][n_chunks] flags, zero on first use); goff: offsets of the tensors inside
 * a region.  params == NULL: returns the number of chunks (sizing query). */
int mpnn_adam_step_ddp(int n, float* const* params, const float* const* grads, float* const* exp_avg,
                       float* const* exp_avg_sq, const long long* numel, const long long* goff, float* step,
                       unsigned int* ticket, float lr, float beta1, float beta2, float eps, float weight_decay,
                       float* const* flat, unsigned* const* flags, long long region_floats, int world, int rank,
                       mpnn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MPNN_B200_H */
