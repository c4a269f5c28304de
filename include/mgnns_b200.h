/*
 * mgnns_b200.h — C-ABI of the B200-native MGNNS forward/backward hot path.
 *
 * One shared object (mgnns_b200/csrc/libmgnns_b200.so, nvcc sm_100a).  Every
 * entry point takes plain device pointers + sizes + a cudaStream_t passed as
 * void*; nothing here knows about torch.  The Python host (mgnns_b200/ops.py)
 * binds these with ctypes and wraps them as torch.library ops ("mgnns::*")
 * with autograd formulas; see INTEGRATION.md for the reference-side stub.
 *
 * Conventions
 *   - all tensors are dense row-major fp32 unless stated, inputs are borrowed
 *     and never mutated, outputs are caller-allocated;
 *   - every function returns 0 on success, non-zero on error;
 *     mgnns_last_error() returns a thread-local message for the last failure;
 *   - no global mutable state, no hidden synchronisation: kernels are only
 *     enqueued on `stream`;
 *   - there is no CPU fallback: a call without a CUDA device fails.
 *
 * "ref:" comments cite the reference (YangXiaocui1215/MGNNS) file:line that the
 * entry point replaces.
 */
#ifndef MGNNS_B200_H
#define MGNNS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGNNS_ABI_VERSION 1

/* activation codes for fused epilogues */
#define MGNNS_ACT_NONE  0
#define MGNNS_ACT_RELU  1
#define MGNNS_ACT_LEAKY 2

int         mgnns_abi_version(void);
const char* mgnns_last_error(void);
/* number of kernels this library has enqueued in this process (for bench.py's gpu_launches) */
int64_t     mgnns_launch_count(void);

/* ---------------------------------------------------------------------------
 * Dense contraction (fp32 CUDA-core path, exact fp32 accumulate).
 *   C[z] (+)= act( sum_{r<reduce} op(A[z*reduce+r]) * op(B[z*reduce+r]) + bias )
 *   op(A) is M x K:  transA==0 -> A[m*lda+k],  transA!=0 -> A[k*lda+m]
 *   op(B) is K x N:  transB==0 -> B[k*ldb+n],  transB!=0 -> B[n*ldb+k]
 *   batch     : number of (A,B) pairs;  reduce : how many consecutive pairs are
 *               summed into one C (batch % reduce == 0); C count = batch/reduce
 *   accumulate: 0 = overwrite C, 1 = atomicAdd into C (C pre-initialised)
 *   bias      : N floats or NULL;  act: MGNNS_ACT_* (applied after bias;
 *               not allowed together with accumulate)
 * ref: torch.matmul at models/Multi_GCN_Multihead_att.py:53,:474,:500,
 *      nn.Linear at :412,:426,:477-479,:504-506,:563-566,
 *      models/submodules.py:68-70,:89,:135 (w_qs/w_ks/w_vs/fc/w_1/w_2)
 * ------------------------------------------------------------------------- */
int mgnns_gemm_f32(int transA, int transB, int M, int N, int K,
                   const float* A, int64_t lda, int64_t strideA,
                   const float* B, int64_t ldb, int64_t strideB,
                   float* C, int64_t ldc, int64_t strideC,
                   int batch, int reduce, int accumulate,
                   const float* bias, int act, float slope, void* stream);

/* Single-pair variant with a caller-provided workspace: small products (fewer than 148 CTAs) are
 * split along K into partial tiles in `workspace` and summed in a fixed order by a second kernel,
 * so results stay bitwise run-to-run deterministic.  mgnns_gemm_splitk_workspace() returns the
 * number of floats the split would need for a shape (0 = no split). */
int64_t mgnns_gemm_splitk_workspace(int M, int N, int K);
int mgnns_gemm_f32_ws(int transA, int transB, int M, int N, int K,
                      const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                      const float* bias, int act, float slope, float* workspace, int64_t workspace_floats,
                      void* stream);

/* y = act'(y_saved) * g   (elementwise; backward of a fused activation) */
int mgnns_act_bwd_f32(const float* y, const float* g, float* out, int64_t n,
                      int act, float slope, void* stream);
/* out[n] += sum_m x[m*ld + n]   (bias gradients), out pre-initialised */
int mgnns_colsum_f32(const float* x, int64_t M, int N, int64_t ld, float* out, void* stream);

/* ---------------------------------------------------------------------------
 * CSR SpMM, sum semiring:  Y[b,i,:] = sum_e val[e] * X[b,col[e],:]
 *   X: [batch, n_cols, F] with row stride ldx and batch stride strideX
 *   Y: [batch, n_rows, F] with row stride ldy and batch stride strideY
 * ref: torch.matmul(adj, support) models/Multi_GCN_Multihead_att.py:54
 * ------------------------------------------------------------------------- */
int mgnns_spmm_csr_f32(int n_rows, const int32_t* rowptr, const int32_t* col, const float* val,
                       const float* X, int64_t ldx, int64_t strideX,
                       float* Y, int64_t ldy, int64_t strideY,
                       int F, int batch, void* stream);

/* The same product for batched features on a big graph, with the hub neighbour rows staged in shared memory
 * (persistent 1024-thread CTAs; a work item = (sample, chunk of row segments); the `n_hub` most referenced columns of
 * X[b] live in a shared-memory table that is reloaded only when the sample changes).  The matrix comes as a host-built
 * plan (mgnns_b200.api.graph_util.hub_plan_arrays):
 *   hub_cols      int32 [n_hub]            column staged in each table slot (n_hub <= mgnns_spmm_hub_capacity(F))
 *   chunk_seg_ptr int32 [n_chunks + 1]     segment range of a chunk
 *   segs          int32 [n_segs, 4]        {first edge, hub edges, edges, row | 1<<31 if the row's only segment}
 *   edges         int32 [nnz, 2]           hub edges first within a row: {byte offset in the table | in X[b], bits of val}
 *   multi_rows    int32 [n_multi]          rows spanning several segments (zeroed first, then accumulated atomically)
 * ref: torch.matmul(adj, support) models/Multi_GCN_Multihead_att.py:54 at BASELINE.json configs[1] size */
int mgnns_spmm_hub_capacity(int F);
int mgnns_spmm_hub_f32(const float* X, int64_t ldx, int64_t strideX, float* Y, int64_t ldy, int64_t strideY,
                       int F, int batch, const int32_t* hub_cols, int n_hub,
                       const int32_t* chunk_seg_ptr, int n_chunks, const int32_t* segs, const int32_t* edges,
                       const int32_t* multi_rows, int n_multi, void* stream);

/* dense [n,n] fp32 adjacency -> CSR (entries != 0), three steps so the caller
 * can allocate col/val after reading nnz = rowptr[n]                        */
int mgnns_dense_row_nnz_f32(const float* A, int n_rows, int n_cols, int64_t ld, int32_t* row_nnz, void* stream);
int mgnns_exclusive_scan_i32(const int32_t* in, int32_t* out /* n+1 */, int n, void* stream);
int mgnns_dense_fill_csr_f32(const float* A, int n_rows, int n_cols, int64_t ld,
                             const int32_t* rowptr, int32_t* col, float* val, void* stream);

/* ---------------------------------------------------------------------------
 * TextLevelGCN channel: per document windowed max-aggregation + sum readout.
 *   doc_ids  int64 [B, L]   (0 = PAD; only the first max_length tokens count)
 *   node_hidden [V, F], edge_w [n_edge_w] (row 0 = shared "no PMI edge" weight)
 *   PMI edge-id map in CSR: row s holds sorted dst ids pmi_col[...] and their
 *   edge ids pmi_eid[...] (NULL -> id = 1 + CSR position, the reference order)
 *   out [B, F] = (relu?) sum_{v in unique non-PAD words} max_{(u->v)} w_e h[u]
 * ref: models/Text_GCN.py:142-166 (edges), :168-211 (graph), :242-249 (max
 *      aggregation), :268-271 (sum readout, ReLU)
 * ------------------------------------------------------------------------- */
int mgnns_text_maxagg_fwd(const int64_t* doc_ids, int B, int L, int max_length, int ngram,
                          const float* node_hidden, int V, int F,
                          const float* edge_w, int64_t n_edge_w,
                          const int32_t* pmi_rowptr, const int32_t* pmi_col, const int32_t* pmi_eid,
                          int apply_relu, float* out, void* stream);
/* grad_out [B,F]; out = forward output (ReLU mask when apply_relu);
 * grad_node_hidden [V,F] and grad_edge_w [n_edge_w] are accumulated (atomicAdd) */
int mgnns_text_maxagg_bwd(const int64_t* doc_ids, int B, int L, int max_length, int ngram,
                          const float* node_hidden, int V, int F,
                          const float* edge_w, int64_t n_edge_w,
                          const int32_t* pmi_rowptr, const int32_t* pmi_col, const int32_t* pmi_eid,
                          int apply_relu, const float* out, const float* grad_out,
                          float* grad_node_hidden, float* grad_edge_w, void* stream);

/* ---------------------------------------------------------------------------
 * Single-query multi-head attention core (re-associated form):
 *   s[b,h,l] = scale * <u[b,h,:], bank[b,l,:]>,  masked (mask[b,l]==0) -> -inf
 *   p = softmax_l(s);  pt = dropout(p, p_drop, seed)
 *   ctx[b,h,:] = sum_l pt[b,h,l] * bank[b,l,:];  psum[b,h] = sum_l pt[b,h,l]
 *   attn[h*B+b, l] = pt   (the reference's head-major [H*B,1,L] layout)
 *   lse[b,h] = logsumexp_l(s)  (saved for backward)
 *   seed_offset: optional device pointer whose value is added to `seed` inside the kernel, so a
 *   captured CUDA graph draws fresh dropout masks on every replay (NULL = use `seed` as is)
 * D % 4 == 0, D <= 512.
 * ref: models/submodules.py:68-78 (projections folded by the caller), :106-119
 * ------------------------------------------------------------------------- */
int mgnns_attn_q1_fwd(const float* u, const float* bank, const float* mask,
                      int B, int H, int L, int D, float scale, float p_drop, uint64_t seed,
                      const uint64_t* seed_offset, float* ctx, float* attn, float* psum, float* lse, void* stream);
int mgnns_attn_q1_bwd(const float* u, const float* bank, const float* mask, const float* lse,
                      const float* grad_ctx, const float* grad_psum /* may be NULL */,
                      int B, int H, int L, int D, float scale, float p_drop, uint64_t seed,
                      const uint64_t* seed_offset, float* grad_u, float* grad_bank, void* stream);

/* The same two entry points on tensor-core fragments (mma.sync m16n8k8 TF32 with the 3xTF32 split: fp32-class accuracy;
 * bank rows streamed through shared memory by bulk asynchronous copies).  Identical arguments and results;
 * mgnns_attn_q1_tc_supported() says whether a shape is covered (H <= 16, D % 4 == 0, D <= 512, tables fit in shared
 * memory) — callers use mgnns_attn_q1_fwd/bwd otherwise.  ref: models/submodules.py:106-119 */
int mgnns_attn_q1_tc_supported(int H, int L, int D);
int mgnns_attn_q1_tc_fwd(const float* u, const float* bank, const float* mask,
                         int B, int H, int L, int D, float scale, float p_drop, uint64_t seed,
                         const uint64_t* seed_offset, float* ctx, float* attn, float* psum, float* lse, void* stream);
int mgnns_attn_q1_tc_bwd(const float* u, const float* bank, const float* mask, const float* lse,
                         const float* grad_ctx, const float* grad_psum,
                         int B, int H, int L, int D, float scale, float p_drop, uint64_t seed,
                         const uint64_t* seed_offset, float* grad_u, float* grad_bank, void* stream);

/* ---------------------------------------------------------------------------
 * Label-query element-wise attention:
 *   out[b,c,h*dh+d] = dropout(softmax_d(Q[c,h,d]*K[b,h,d]*inv_scale))[d] * V[b,h,d]
 *   Q [C, heads*dh];  K, V [B, heads*dh] with row stride ldkv
 * ref: models/Multi_GCN_Multihead_att.py:97-131
 * ------------------------------------------------------------------------- */
int mgnns_label_attn_fwd(const float* Q, const float* K, const float* V, int64_t ldkv,
                         int B, int C, int heads, int dh, float inv_scale,
                         float p_drop, uint64_t seed, const uint64_t* seed_offset, float* out, void* stream);
/* grad_Q [C,heads*dh] is accumulated (atomicAdd, pre-zeroed by caller);
 * grad_K / grad_V [B, heads*dh] with row stride ldg are overwritten          */
int mgnns_label_attn_bwd(const float* Q, const float* K, const float* V, int64_t ldkv,
                         int B, int C, int heads, int dh, float inv_scale,
                         float p_drop, uint64_t seed, const uint64_t* seed_offset, const float* grad_out,
                         float* grad_Q, float* grad_K, float* grad_V, int64_t ldg, void* stream);

/* ---------------------------------------------------------------------------
 * Residual add + the reference's custom LayerNorm:
 *   z = x (+ res);  y = gamma * (z - mean) / (std_unbiased + eps) + beta
 * ref: models/submodules.py:142-156, call sites :90, :138
 * ------------------------------------------------------------------------- */
int mgnns_add_layernorm_fwd(const float* x, const float* res /* may be NULL */,
                            const float* gamma, const float* beta,
                            int64_t rows, int D, float eps, float* y, void* stream);
/* grad_z [rows,D] overwritten (gradient w.r.t. x and res alike);
 * grad_gamma / grad_beta [D] accumulated (atomicAdd, pre-zeroed by caller)   */
int mgnns_add_layernorm_bwd(const float* x, const float* res, const float* gamma,
                            const float* grad_y, int64_t rows, int D, float eps,
                            float* grad_z, float* grad_gamma, float* grad_beta, void* stream);

/* ---------------------------------------------------------------------------
 * Global spatial max over the trunk feature map  F[B*C, P] -> pooled[B*C],
 * argmax[B*C] (first maximum, the torch MaxPool2d convention).
 * ref: nn.MaxPool2d(14,14) models/Multi_GCN_Multihead_att.py:302,:454,:486
 * ------------------------------------------------------------------------- */
int mgnns_rowmax_f32(const float* F, int64_t rows, int P, float* pooled, int32_t* argmax, void* stream);
/* grad_F[r, argmax[r]] += grad_pooled[r]  (grad_F pre-initialised) */
int mgnns_rowmax_bwd_f32(const float* grad_pooled, const int32_t* argmax, int64_t rows, int P,
                         float* grad_F, void* stream);

/* ---------------------------------------------------------------------------
 * Dense layer on the tcgen05 tensor cores (TMA-fed, double-buffered TMEM accumulators):
 *   C[M,N] = act(A[M,K] . W + bias)        A row-major (lda), C row-major (ldc)
 *   w_is_kn: 1 = W is [K,N] (GraphConvolution layout), 0 = W is [N,K] (nn.Linear layout)
 *   precision: 0 = TF32 operands, 1 = 3xTF32 split (fp32-class accuracy; needs
 *              mgnns_linear_tc_workspace() floats of 16-byte aligned workspace for the split weight)
 *   constraints: lda % 4 == 0, ldw % 4 == 0, A and W 16-byte aligned
 * ref: support = torch.matmul(input, weight) + the activation applied by the caller,
 *      models/Multi_GCN_Multihead_att.py:52-58, :470-472; nn.Linear call sites with many rows
 * ------------------------------------------------------------------------- */
int64_t mgnns_linear_tc_workspace(int N, int K, int64_t ldw, int w_is_kn, int precision);
int mgnns_linear_tc(const float* A, int64_t lda, const float* W, int64_t ldw, int w_is_kn,
                    const float* bias, int act, float slope, int M, int N, int K, int precision,
                    float* workspace, int64_t workspace_floats, float* C, int64_t ldc, void* stream);

/* ---------------------------------------------------------------------------
 * Fused graph-convolution layer for batched node features, one kernel:
 *   Y[b] = act((A_hat . X[b]) . W + bias)      X [batch, n_cols, K] (row stride K), Y [batch, n_rows, N] (row stride ldy)
 * The aggregated rows A_hat.X[b] are gathered straight into the shared-memory operand of the tcgen05 tensor core
 * (they never exist in HBM); W [K,N] is streamed by TMA.  A_hat comes as a host-built execution plan of the CSR
 * matrix (mgnns_b200.api.graph_util.CSRAdjacency.fused_plan): 128-row tiles dealt by degree rank, each row cut
 * into segments of at most 128 edges.
 *   tile_seg_ptr   int32 [n_tiles + 1]     segment range of a tile
 *   segs           int32 [n_segs, 4]       {first edge, edges (<= 128), row in tile, 1 = the row's only segment}
 *   edges          int32 [nnz, 2]          {byte offset of the neighbour row inside X[b] (= col * K * 4), bits of val}
 *   tile_rows      int32 [n_tiles * 128]   output row of each tile slot, -1 = padding
 *   tile_multi_ptr int32 [n_tiles + 1], multi_rows int32 [...]   tile slots whose row spans several segments
 *   precision: 0 = TF32 operands, 1 = 3xTF32 split (fp32-class); workspace: mgnns_gcn_fused_workspace() floats
 *   constraints: K % 4 == 0, N % 32 == 0, N <= 512, X 16-byte aligned, strideX % 4 == 0
 * ref: GraphConvolution.forward, models/Multi_GCN_Multihead_att.py:52-58 (+ the caller's activation, :470-472);
 *      BASELINE.json configs[1] (10k-node word graph, 300 -> 512, batch 256)
 * ------------------------------------------------------------------------- */
int64_t mgnns_gcn_fused_workspace(int N, int K, int precision);
int mgnns_gcn_fused_tc(const float* X, int64_t strideX, int batch,
                       const int32_t* tile_seg_ptr, const int32_t* segs, const int32_t* edges,
                       const int32_t* tile_rows, const int32_t* tile_multi_ptr, const int32_t* multi_rows,
                       int n_tiles, const float* W, int64_t ldw, const float* bias, int act, float slope,
                       int K, int N, int precision, float* workspace, int64_t workspace_floats,
                       float* Y, int64_t ldy, int64_t strideY, void* stream);

/* Weight-gradient product on the tensor cores:  C[M,N] = A[K,M]^T . B[K,N]  with a long reduction (K = rows of
 * a batch / tokens).  A and B are row-major with the reduction index as the row (lda, ldb multiples of 4, 16-byte
 * aligned); C is overwritten.  K slices are accumulated with fp32 atomics (summation order varies run to run).
 * ref: the dW that autograd computes for nn.LSTM's projections (model:179-184) and the many-row nn.Linear layers */
int mgnns_wgrad_tc(const float* A, int64_t lda, const float* B, int64_t ldb, int M, int N, int K, int precision,
                   float* C, int64_t ldc, void* stream);

/* ---------------------------------------------------------------------------
 * Image-bank contraction on the tcgen05 tensor cores (TMA-fed, TMEM accumulators).
 *   fwd: bank[b,p,o] = sum_c fmap[b,c,p] * weight[o,c] + bias[o]
 *   dw : gW[o,c]    += sum_{b,p} gbank[b,p,o] * fmap[b,c,p]      (gW initialised by the caller)
 *   precision: 0 = TF32 operands, 1 = 3xTF32 split (fp32-class accuracy)
 *   constraints: C % 32 == 0, P % 4 == 0, O % 4 == 0, O <= 304 (fwd) / 320 (dw)
 * ref: self.liner_img_object / liner_img_place, models/Multi_GCN_Multihead_att.py:400-428
 * ------------------------------------------------------------------------- */
int mgnns_imgbank_fwd_tc(const float* fmap, const float* weight, const float* bias,
                         int B, int C, int P, int O, int precision,
                         float* workspace /* 2*O*C floats when precision == 1, else may be NULL */,
                         float* pooled /* optional [B,C]: global spatial max (ref: nn.MaxPool2d(14,14), model:302) fused
                                          into the operand pass; precision == 1 only, else NULL */,
                         float* bank, void* stream);
int mgnns_imgbank_fwd_tc_capped(const float* fmap, const float* weight, const float* bias,
                                int B, int C, int P, int O, int precision, float* workspace, float* pooled,
                                float* bank, int max_ctas /* 0 = every SM */, void* stream);
int mgnns_imgbank_dw_tc(const float* fmap, const float* gbank, int B, int C, int P, int O,
                        int precision, float* gW, void* stream);
/* same, on at most max_ctas SMs (0 = all): the weight gradient is not needed before the optimizer (engine:850-851), so
 * the training step runs it beside the latency-bound backward chain instead of in front of it */
int mgnns_imgbank_dw_tc_capped(const float* fmap, const float* gbank, int B, int C, int P, int O,
                               int precision, float* gW, int max_ctas, void* stream);

/* ---------------------------------------------------------------------------
 * Length-aware bidirectional LSTM recurrence over compacted tokens (PyTorch gate order i,f,g,o).
 *   offsets int32 [B+1] (first compact row of each sequence), lens int32 [B],
 *   tiles int32 [n_tiles*8] (sequence ids grouped by similar length, -1 = empty slot)
 *   G [N, 2*4H]: input projections x W_ih^T + b_ih + b_hh of both directions (forward | reverse)
 *   wt4_* [H,H,4]: gate-interleaved transpose of W_hh (mgnns_lstm_prep_whh);  H even, <= 160
 *   fwd writes Y [N,2H] (forward | reverse hidden states), gates [N,2,4,H] (post-activation),
 *   csave [N,2,H], hprev [N,2,H];  bwd writes dG [N, 2*4H] (pre-activation gradients)
 * ref: nn.LSTM over pack_padded_sequence, models/Multi_GCN_Multihead_att.py:376-384
 * ------------------------------------------------------------------------- */
int mgnns_lstm_prep_whh(const float* whh /* [4H,H] */, float* wt4, int H, void* stream);
int mgnns_lstm_rec_fwd(const int32_t* offsets, const int32_t* lens, const int32_t* tiles, int n_tiles,
                       int H, const float* G, const float* wt4_f, const float* wt4_r, float* Y,
                       float* gates, float* csave, float* hprev, void* stream);
int mgnns_lstm_rec_bwd(const int32_t* offsets, const int32_t* lens, const int32_t* tiles, int n_tiles,
                       int H, const float* dY, const float* gates, const float* csave,
                       const float* whh_f, const float* whh_r, float* dG, void* stream);

/* ---------------------------------------------------------------------------
 * PMI co-occurrence counting (integer, bit-exact, order independent).
 *   tokens int32 [D, L]: vocab index, or -1 for out-of-vocabulary; pad_id is the
 *   vocab index of the literal 'PAD' token (centre positions equal to pad_id
 *   are skipped; as a *target* it is counted, as in the reference).
 *   window: centre i pairs with j in [max(0,i-window), min(L,i+window)), j != i.
 *   pair_count int32 [V,V] and word_count int64 [V] must be zeroed by the caller.
 * ref: utils/pmi.py:40-58
 * ------------------------------------------------------------------------- */
int mgnns_pmi_count(const int32_t* tokens, int64_t D, int L, int V, int window, int pad_id,
                    int32_t* pair_count, int64_t* word_count, void* stream);
/* keep cells with count >= min_count (ref: utils/pmi.py:60-66), row-major order */
int mgnns_count_row_nnz_i32(const int32_t* M, int n_rows, int n_cols, int min_count,
                            int32_t* row_nnz, void* stream);
int mgnns_count_fill_csr_i32(const int32_t* M, int n_rows, int n_cols, int min_count,
                             const int32_t* rowptr, int32_t* col, int32_t* cnt, void* stream);

/* ---------------------------------------------------------------------------
 * PMI co-occurrence counts without the dense [V,V] table (the product path; the dense-table entry points above
 * remain for small vocabularies and as a cross-check).  Steps, all on `stream`:
 *   mgnns_pmi_row_emissions   row_emit[c] = number of (centre c -> target) pairs, word_count[c] (both zeroed here)
 *   mgnns_exclusive_scan_i64  row_start[V+1]
 *   mgnns_pmi_scatter_targets targets[row_start[c] ...] = the row's targets in arrival order (cursor zeroed here)
 *   mgnns_pmi_row_reduce      per centre row: shared-memory column counters, cells >= min_count written in column
 *                             order to tmp_col/tmp_cnt at row_start[c] (at most as many as emissions), row_nnz[c]
 *   mgnns_pmi_compact         tmp -> CSR col/cnt at rowptr (int32 scan of row_nnz by mgnns_exclusive_scan_i32)
 * Only centres in [row_lo,row_hi) are counted (row sharding across ranks: disjoint row ranges need no reduction).
 * A cell's count is bounded by its row's emission count; callers check max(row_emit) < 2^31 (cells are int32).
 * ref: utils/pmi.py:37-66 (pair_count / word_count loops and the min_cooccurence filter), :89-97 (row-major ids)
 * ------------------------------------------------------------------------- */
int mgnns_pmi_row_emissions(const int32_t* tokens, int64_t D, int L, int V, int window, int pad_id,
                            int row_lo, int row_hi, int64_t* row_emit, int64_t* word_count, void* stream);
int mgnns_exclusive_scan_i64(const int64_t* in, int64_t* out /* n+1 */, int n, void* stream);
int mgnns_pmi_scatter_targets(const int32_t* tokens, int64_t D, int L, int V, int window, int pad_id,
                              int row_lo, int row_hi, const int64_t* row_start, int64_t* cursor,
                              int32_t* targets, void* stream);
int mgnns_pmi_row_reduce(const int32_t* targets, const int64_t* row_start, int V, int min_count,
                         int32_t* tmp_col, int32_t* tmp_cnt, int32_t* row_nnz, void* stream);
int mgnns_pmi_compact(const int32_t* tmp_col, const int32_t* tmp_cnt, const int64_t* row_start,
                      const int32_t* rowptr, int V, int32_t* col, int32_t* cnt, void* stream);

/* ---------------------------------------------------------------------------
 * Training-loop bookkeeping on the device.
 *   mgnns_confusion_count: pred[b] = argmax_c scores[b,c]; conf[target[b]*C + pred[b]] += 1 (conf int32 [C,C],
 *     accumulated, caller zeroes it); pred_out (optional) int64 [B].
 *     ref: engine/Multi_GCN_Multihead_Att_engine.py:829-838 (argmax -> .cpu() -> sklearn accuracy/F1 every batch)
 *   mgnns_label_cooccurrence: labels int32 [n_images, max_len] (lens int32 [n_images] valid entries per image):
 *     nums[j] += images containing j, adj[a*C+b] += images containing both a and b, a != b (int64, accumulated).
 *     ref: utils/util.py:336-357 (generate_nums, generate_Adj)
 * ------------------------------------------------------------------------- */
int mgnns_confusion_count(const float* scores, int64_t ld, const int64_t* target, int B, int C,
                          int32_t* conf, int64_t* pred_out, void* stream);
int mgnns_label_cooccurrence(const int32_t* labels, const int32_t* lens, int64_t n_images, int max_len,
                             int C, int64_t* nums, int64_t* adj, void* stream);

/* ---------------------------------------------------------------------------
 * clip_grad_norm_ + Adam over flat buffers (two launches; mgnns_b200.optim.FlatClipAdam is the host side).
 *   mgnns_sqnorm_f32:    *out = sum of squares of g[0..n) (double; zeroed here)
 *   mgnns_clip_adam_f32: g *= min(1, max_norm/(sqrt(*sqnorm) + 1e-6)) in place; for every segment s (gradient range
 *     [seg_g[s], seg_g[s+1]), all multiples of 4) with seg_p[s] >= 0, torch.optim.Adam's update (L2 weight decay,
 *     bias correction from the device step counter *step, which the caller increments first) on p/m/v at seg_p[s].
 * ref: engine/Multi_GCN_Multihead_Att_engine.py:850-851, Tumblr_Multi_GCN_Multihead_Att.py:164
 * ------------------------------------------------------------------------- */
int mgnns_sqnorm_f32(const float* g, int64_t n, double* out, void* stream);
int mgnns_clip_adam_f32(float* g, int64_t n, const int64_t* seg_g, const int64_t* seg_p, const float* seg_lr,
                        const float* seg_wd, int n_seg, float* p, float* m, float* v, const double* sqnorm,
                        double max_norm, double beta1, double beta2, double eps, const int64_t* step, void* stream);


/* ---------------------------------------------------------------------------
 * Text memory bank glue around the compacted-token LSTM.
 * ref: models/Multi_GCN_Multihead_att.py:366-398 (self.embedding; pad_packed_sequence(total_length) -> zero rows past
 * each text's length)
 *   pad_rows_fwd   bank[b,t,:] = t < lens[b] ? y[offsets[b]+t,:] : 0          (bank [B,L,F], y [N,F], F % 4 == 0)
 *   pad_rows_bwd   gy[offsets[b]+t,:] = gbank[b,t,:] for t < lens[b]           (other rows of gy untouched)
 *   embedding_bwd  gw[tokens[i],:] += g[i,:] for tokens[i] != padding_idx      (gw [V,E] initialised by the caller)
 * ------------------------------------------------------------------------- */
int mgnns_pad_rows_fwd(const float* y, const int32_t* offsets, const int32_t* lens, int B, int L, int F,
                       float* bank, void* stream);
int mgnns_pad_rows_bwd(const float* gbank, const int32_t* offsets, const int32_t* lens, int B, int L, int F,
                       float* gy, void* stream);
int mgnns_embedding_bwd(const int64_t* tokens, const float* g, int64_t n, int E, int64_t padding_idx, int64_t V,
                        float* gw, void* stream);

/* ---------------------------------------------------------------------------
 * Gradient all-reduce over NVLink peer memory (one kernel, capturable in a CUDA graph).
 * ref: SURVEY 8e / engine/Multi_GCN_Multihead_Att_engine.py:847-851 — the one collective of the path sits between
 * loss.backward() and clip_grad_norm_; the reference itself is single-GPU.
 *   mgnns_p2p_alloc / _free       cudaMalloc'd, zeroed memory that CUDA IPC can export (the flat gradient buffer and a
 *                                 flag block of mgnns_p2p_flag_bytes() bytes per rank)
 *   mgnns_p2p_export / _import    64-byte cudaIpcMemHandle_t of an allocation / peer mapping of another rank's handle
 *   mgnns_allreduce_p2p_f32       buf_r[i] = scale * sum_q buf_q[i] on every rank r (bit-identical results on all
 *                                 ranks); bufs / flags are HOST arrays of `world` device pointers, own allocation at
 *                                 index `rank`; n % 4 == 0; all ranks must launch it the same number of times
 *   mgnns_p2p_error               1 if a cross-GPU barrier of this rank ever timed out (20 s), -1 on a CUDA error
 * ------------------------------------------------------------------------- */
int mgnns_p2p_flag_bytes(void);
int mgnns_p2p_alloc(int64_t bytes, void** out);
int mgnns_p2p_free(void* p);
int mgnns_p2p_export(void* p, void* handle64);
int mgnns_p2p_import(const void* handle64, void** out);
int mgnns_p2p_close(void* p);
int mgnns_allreduce_p2p_f32(const uint64_t* bufs, const uint64_t* flags, int rank, int world, int64_t n,
                            float scale, int ctas, void* stream);
int mgnns_p2p_error(const void* own_flags, void* stream);

/* A single-thread kernel that completes `ns` nanoseconds (<= 1 ms) after it starts: a timed dependency edge for the
 * multi-stream training step (mgnns_b200/ops.py gates the image-bank weight-gradient kernels behind the launch of the
 * LSTM recurrence, engine:847 loss.backward()). No reference counterpart. */
int mgnns_delay_ns(int ns, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MGNNS_B200_H */
