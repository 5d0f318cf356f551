/*
 * cgnn.h - C ABI of the B200-native batched message-passing path of connectome-gnn-suite.
 *
 * The reference (danieleschmidt/connectome-gnn-suite) is pure Python and has no FFI of its
 * own: its "operators" are ATen call sequences inside connectome_gnn/graph.py, models.py
 * and train.py.  Each entry point below replaces one such sequence (cited per function as
 * reference file:line).  A maintainer binds these with ctypes; the
 * host package under connectome-gnn-suite_b200/connectome_gnn does exactly that.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name
 *     ends in _host.  No torch types.  All tensors are dense, row-major, fp32 / int32 /
 *     int64 exactly as documented.
 *   - every function enqueues work on `stream` (a cudaStream_t passed as void*) and
 *     returns immediately; 0 = CGNN_OK, otherwise a cgnn_status code (never throws).
 *   - functions are re-entrant; the only state is caller-owned (`workspace`).
 *   - "rows" are the nodes (brain regions) of all subjects of a batch, packed subject
 *     after subject; `ptr[B+1]` (int64) delimits subjects, as ConnectomeBatch.ptr does.
 */
#ifndef CGNN_H_
#define CGNN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CGNN_ABI_VERSION 11

typedef void* cgnn_stream_t; /* cudaStream_t */

typedef enum {
  CGNN_OK = 0,
  CGNN_ERR_INVALID_ARG = 1,     /* null pointer / negative size / unsupported combination */
  CGNN_ERR_TILE_TOO_LARGE = 2,  /* one subject's [N_g, C] fp32 tile (or W) exceeds shared memory */
  CGNN_ERR_WORKSPACE = 3,       /* workspace too small: call cgnn_workspace_bytes()          */
  CGNN_ERR_CUDA = 4,            /* a CUDA runtime call failed: see cgnn_last_cuda_error()     */
  CGNN_ERR_NEED_CSR = 5,        /* a lean batch (blobs only) met a code path that reads the CSR arrays: materialise
                                 * them (full cgnn_collate_csr / cgnn_csr_from_coo) and call again                */
  CGNN_ERR_UNSUPPORTED = 6      /* cgnn_eval_fused_fwd: shape not covered - use the per-layer entry points         */
} cgnn_status;

const char* cgnn_status_string(int status);
int cgnn_abi_version(void);
int cgnn_last_cuda_error(void);          /* cudaError_t of the last CGNN_ERR_CUDA on this thread */
/* Upper bound on the scratch a single call needs (partials of BN statistics / weight
 * gradients).  The caller allocates it once per device+stream and passes it to every call. */
size_t cgnn_workspace_bytes(void);
/* Number of CUDA kernels this library has launched in this process so far (monotonic). */
uint64_t cgnn_kernel_launches(void);
/* Process-wide options.  CGNN_OPT_TENSOR_CORES: 1 (default) = eligible layer shapes run the tcgen05/TMEM
 * kernels, 0 = every shape runs the generic SIMT kernels (same results to fp32 round-off; used to
 * cross-check the two paths on the device). */
#define CGNN_OPT_TENSOR_CORES 1
int cgnn_set_option(int32_t key, int32_t value);

/* ---------------------------------------------------------------------------------------
 * Device CSR of a batch.  Built once per batch by cgnn_collate_csr / cgnn_csr_from_coo and
 * read by every layer (max_nodes / max_edges = largest subject of the batch, they size the
 * shared-memory tiles); replaces the per-layer recomputation in reference models.py:94-108
 * (self-loops, D^, d^-1/2, w^) and models.py:147-148 (w_sum).
 *
 *   in_*   rows = destination node, entries in COO order (stable), col = source node id
 *   out_*  rows = source node,      entries in COO order (stable), col = destination id
 *   *_w    raw edge weight;  *_wn = (dinv[src] * w) * dinv[dst]   (GCN normalisation)
 *   deg    D^_i = ((0 + w_e1) + w_e2 ... ) + 1.0f over edges with src == i in COO order
 *          (bit-exact with the CPU scatter_add_ of models.py:103-104, self-loop last)
 *   wsum   sum_{e: dst == i} w_e, same sequential order (models.py:147-148)
 *   dinv   (deg + 1e-8f)^-1/2  (models.py:105)
 * Self-loops are NOT stored; kernels add (dinv_i*dinv_i) * P_i last, as the reference's
 * concatenation order implies (models.py:98-100).
 * ------------------------------------------------------------------------------------- */
typedef struct {
  const int32_t* in_rowptr;  /* [rows+1] positions into in_col/in_w/in_wn   */
  const int32_t* in_col;     /* [E] global row id of the source node        */
  const float*   in_w;       /* [E] */
  const float*   in_wn;      /* [E] */
  const int32_t* out_rowptr; /* [rows+1] */
  const int32_t* out_col;    /* [E] global row id of the destination node   */
  const float*   out_w;      /* [E] */
  const float*   out_wn;     /* [E] */
  const float*   deg;        /* [rows] */
  const float*   dinv;       /* [rows] */
  const float*   wsum;       /* [rows] */
  const int32_t* graph_meta; /* [B][4] per subject {first row, rows, first edge, edges}: one 16-byte load per subject */
  /* Optional packed aggregation blobs built by cgnn_build_agg or by cgnn_collate_csr itself (NULL = not built: the
   * layer entry points then run their generic kernels).  One 16-byte aligned blob per subject and direction, at
   * word offset 8*first_row + 2*(first_edge rounded up to even) + 4*subject inside the buffer:
   *   [rows x int4 {rec_begin, rec_end, aux bits, row}] [records int2 {neighbour, weight bits}]
   * Descriptor k stands for row k.  Every row's record list is padded to an even length with
   * zero-weight records; agg_kind 0 (GCN) carries the normalised weights w^ and ends every row with the self-loop
   * record {row, dinv^2} (reference models.py:98-100 puts self-loops last); agg_kind 1 (GraphSAGE) carries w with
   * aux = w_sum by destination, and w / (w_sum[dst] + 1e-8) by source (the adjoint of models.py:146-149).
   * `neighbour` is the local row index in agg_out and (row << 8) | ((row & 7) << 4) in agg_in (the byte offset of the
   * row in a 256-byte-pitch shared-memory tile with its swizzle key).
   * A LEAN batch carries graph_meta, row_graph and the blobs only - every array pointer above is NULL: all the
   * tensor-core layer kernels read nothing else.  Entry points that would need the arrays return CGNN_ERR_NEED_CSR. */
  const int32_t* agg_in;     /* rows = destination nodes, neighbours = sources       */
  const int32_t* agg_out;    /* rows = source nodes,      neighbours = destinations (NULL on inference-only batches) */
  const int32_t* row_graph;  /* [rows] subject index of every row (ConnectomeBatch.batch as int32) */
  int32_t agg_kind;          /* 0 = GCN, 1 = GraphSAGE, -1 = none */
} cgnn_csr_t;

/* How a stored activation tensor t [rows, C] is turned into the layer input u on load:
 *     u = dropout( relu?( scale * t + shift ) )
 * This is the previous layer's BatchNorm1d + ReLU + dropout (reference models.py:208-210 for
 * GCN, :260-261 for SAGE) fused into the consumer.  scale == NULL means identity (first
 * layer reads node_features as is).  Dropout masks are a pure function of
 * (seed, site, global row id, channel) so forward and backward regenerate the same mask
 * and the result does not depend on how subjects are sharded over GPUs. */
typedef struct {
  const float* scale;  /* [C] gamma * rstd, or NULL */
  const float* shift;  /* [C] beta - mean * scale   */
  int32_t  relu;       /* 1: max(.,0) after the affine map */
  float    p_drop;     /* drop probability, 0 disables (eval mode or dropout=0) */
  uint64_t seed;
  uint32_t site;       /* which dropout site of the network (layer index)  */
  int64_t  row_base;   /* global row id of local row 0 (rank offset under data parallelism) */
  const uint32_t* salt; /* NULL, or two DEVICE words folded into the mask stream at kernel start: a CUDA-graphed step keeps
                         * them in its state block (cgnn_step_tick) so that every replay draws new masks */
} cgnn_act_t;

/* BatchNorm1d backward coefficients of one layer (reference: autograd of models.py:208/260):
 *   dz = scale * (dy - s1/count - xhat * s2/count),  xhat = (z - mean) * rstd   [train]
 *   dz = scale * dy                                                              [eval]
 * s1 = sum dy, s2 = sum dy*xhat over all rows of the GLOBAL batch (all ranks). */
typedef struct {
  const float* scale;  /* [C] gamma * rstd */
  const float* mean;   /* [C] */
  const float* rstd;   /* [C] */
  const float* s1;     /* [C] (unused when train == 0) */
  const float* s2;     /* [C] */
  double  count;       /* rows of the global batch */
  int32_t train;       /* 1: batch statistics were used in forward */
  const double* sums64; /* NULL, or [2, C] = the same s1, s2 in double (what cgnn_bn_bwd_sums / prev_sums64 wrote): read instead
                         * of s1 / s2.  The backward subtracts s/count from every row (cancellation): with the sums carried in
                         * double a data-parallel run and the single-process batch agree to fp32 round-off of the inputs */
} cgnn_bn_bwd_t;

/* ---- K0: collate (reference graph.py:143-167) ---------------------------------------- */

/* Subject store ("arena"): all subjects of a dataset packed once on the device.
 * Edge endpoints are subject-local int32 ids; node_ptr/edge_ptr are int64 prefix sums. */
typedef struct {
  const float*   x;         /* [sum N_s, F] */
  const int32_t* src;       /* [sum E_s] local source id      (edge_index[0]) */
  const int32_t* dst;       /* [sum E_s] local destination id (edge_index[1]); NULL = compact store: src[e] holds
                             * both endpoints as src | dst << 16 (every subject has < 65536 nodes)            */
  const float*   w;         /* [sum E_s] */
  const int64_t* node_ptr;  /* [S+1] */
  const int64_t* edge_ptr;  /* [S+1] */
  const int64_t* label;     /* [S] or NULL */
  int32_t num_features;
  int32_t edge_pairs;       /* 1 (compact store only, dst == NULL): src / w hold one entry per UNDIRECTED edge; entry k of a
                             * subject stands for the two adjacent COO edges 2k = (s -> d) and 2k + 1 = (d -> s) of equal
                             * weight - the layout the reference generator (synthetic.py:137-158) and its real-data recipe
                             * (README.md:145-179) emit.  edge_ptr keeps counting directed edges (even per subject); the
                             * entries of subject s start at edge_ptr[s] / 2.  Halves the edge bytes that cross PCIe. */
} cgnn_store_t;

/* Mutable view of the CSR arrays the collate kernels fill (same field order as cgnn_csr_t). */
typedef struct {
  int32_t* in_rowptr;  int32_t* in_col;  float* in_w;  float* in_wn;
  int32_t* out_rowptr; int32_t* out_col; float* out_w; float* out_wn;
  float* deg; float* dinv; float* wsum;
  int32_t* graph_meta;
  /* optional: have the collate kernel also emit the packed aggregation blobs of one model family (what
   * cgnn_build_agg would produce from the arrays above, bit for bit) while the sorted lists are still in shared
   * memory.  agg_kind = -1 (or NULL pointers): not requested; agg_out alone may be NULL (forward-only batches).
   * LEAN request: every array pointer above except graph_meta is NULL and the blobs are requested - the kernel then
   * writes nothing but node_features, labels, graph_meta, row_graph and the blobs (a quarter of the bytes). */
  int32_t* agg_in; int32_t* agg_out; int32_t* row_graph;
  int32_t agg_kind;
} cgnn_csr_out_t;

/* Gather `num_graphs` subjects (ids into the store, device int64) into one batch.
 * Writes the six ConnectomeBatch fields (graph.py:117-122) bit-exactly as collate_graphs
 * does - node_features [rows,F], edge_index [2,E] int64 with node offsets added
 * (graph.py:152), edge_weight [E], batch [rows] int64 (graph.py:154), labels [B] int64
 * (NULL to skip), ptr [B+1] int64 (graph.py:158,166) - plus eptr [B+1] int64 (edge prefix
 * sums) and the device CSR.  total_rows / total_edges are the host-computed sums used to
 * size the outputs; max_nodes / max_edges are the largest N_s / E_s among the selected subjects
 * (they size the per-CTA shared memory; max_edges < 0 = unknown).  edge_index + edge_weight and
 * batch may be NULL: the field is then not materialised (the host mirror fills it on first access
 * by collating again).  A lean request whose largest subject cannot be sorted in shared memory
 * returns CGNN_ERR_NEED_CSR. */
int cgnn_collate_csr(const cgnn_store_t* store, const int64_t* subject_ids, int64_t num_graphs,
                     int64_t total_rows, int64_t total_edges, int32_t max_nodes, int32_t max_edges,
                     float* node_features, int64_t* edge_index, float* edge_weight,
                     int64_t* batch, int64_t* labels, int64_t* ptr, int64_t* eptr,
                     const cgnn_csr_out_t* csr, cgnn_stream_t stream);

/* Dense connectivity matrices -> thresholded subjects (reference README.md:145-179, `hcp_matrix_to_graph`: keep the
 * entries above the matrix's own q-quantile, list every kept (i, j) in row-major order as i -> j and then again as
 * j -> i, feature = weighted row sum / its maximum).  matrices [S, N, N] f32.
 *   cgnn_ingest_threshold: threshold[s] = torch.quantile(A_s.flatten(), q) (linear interpolation, fp32 rank arithmetic),
 *                          selected[s] = number of entries with A > threshold and A > 0.
 *   cgnn_ingest_emit:      edge_ptr [S+1] = prefix sums of 2 * selected (the caller's scan); writes subject-local
 *                          src / dst / weight [edge_ptr[S]] and node_features [S * N, 1].
 * Edges and weights are bit-exact against the recipe (with `A_thresh[src, dst]` for its `A_thresh[src]`). */
int cgnn_ingest_threshold(const float* matrices, int64_t num_subjects, int32_t N, float q, float* threshold, int32_t* selected,
                          cgnn_stream_t stream);
int cgnn_ingest_emit(const float* matrices, int64_t num_subjects, int32_t N, const float* threshold, const int64_t* edge_ptr,
                     int32_t* src, int32_t* dst, float* weight, float* node_features, cgnn_stream_t stream);

/* ids[i] = pinned_host_ids[i]: brings the subject indices of a batch from PINNED (device-mapped) host memory into a
 * device buffer with a kernel instead of a DMA, so that they never queue behind a dataset upload on the copy engine. */
int cgnn_fetch_ids(const int64_t* pinned_host_ids, int64_t n, int64_t* ids, cgnn_stream_t stream);

/* Same CSR, from an already collated batch (COO with global ids, grouped by subject as
 * collate_graphs emits it).  Used for ConnectomeBatch objects built by hand / moved from
 * the host.  eptr [B+1] is an output. */
int cgnn_csr_from_coo(const int64_t* edge_index, const float* edge_weight, const int64_t* ptr,
                      int64_t num_graphs, int64_t total_rows, int64_t total_edges, int32_t max_nodes,
                      int64_t* eptr, const cgnn_csr_out_t* csr, cgnn_stream_t stream);

/* Packed aggregation blobs for one model family (kind 0 = GCN, 1 = GraphSAGE) from a built CSR: fills
 * agg_in / agg_out (each cgnn_agg_words(...) int32 words) and row_graph [rows].  Pure re-layout of the CSR
 * arrays above plus the self-loop / mean weights; nothing the reference computes is changed. */
size_t cgnn_agg_words(int64_t rows, int64_t edges, int64_t num_graphs);
int cgnn_build_agg(const cgnn_csr_t* csr, int32_t kind, int64_t num_graphs, int64_t rows, int64_t edges,
                   int32_t max_nodes, int32_t* agg_in, int32_t* agg_out, int32_t* row_graph, cgnn_stream_t stream);

/* ---- K1/K2: layer forward ------------------------------------------------------------ */

/* GCN layer (reference models.py:84-114):  z = A^ (u W^T) + bias,  u = act(t_in).
 * t_in [rows, d_in], W [H, d_in], bias [H], z [rows, H].  With the GCN blobs of cgnn_build_agg attached to `csr` the
 * layer runs as ONE kernel that evaluates the same product as (A^ u) W^T: the gather happens in the input width and
 * the gathered rows feed the tensor cores from registers (fp32 round-off differs from the reference order by ~1e-7).
 * If bn_stats != NULL also returns the BatchNorm batch statistics of z as doubles
 * [1 + 2H] = {count, mean[H], M2[H]} (Welford/Chan merged, deterministic). */
int cgnn_gcn_layer_fwd(const float* t_in, const cgnn_act_t* act, const float* W, const float* bias,
                       const cgnn_csr_t* csr, const int64_t* ptr, int64_t num_graphs, int64_t rows,
                       int32_t d_in, int32_t H, int32_t max_nodes, int32_t max_edges, float* z, double* bn_stats,
                       void* workspace, size_t workspace_bytes, cgnn_stream_t stream);

/* GraphSAGE layer (reference models.py:136-152):
 *   agg_i = sum_{e: dst=i} w_e u_src / (wsum_i + 1e-8);  z = relu([u || agg] W^T + b)
 * W [H, 2*d_in], b [H].  agg [rows, d_in] (may be NULL) receives the aggregated neighbourhood (every code path fills
 * it when given): with it and the blobs of cgnn_build_agg the layer runs as a gather kernel plus a tensor-core
 * contraction, and backward reuses it instead of gathering again. */
/* GCN layer forward with the readout folded in (eval mode, last layer; reference models.py:208-211 in eval):
 *   emb[g] = mean over the rows i of subject g of act_out(z_i),  z = A^ (act(t_in) W^T) + b
 * act_out = the BatchNorm affine (running statistics) + ReLU of this layer's output, p_drop = 0.  z is not written.
 * Hidden layers 64 -> 64 with the GCN blobs only; CGNN_ERR_UNSUPPORTED otherwise (the caller runs cgnn_gcn_layer_fwd +
 * cgnn_pool_fwd).  A subject's sum is formed in the same order in every batch. */
int cgnn_gcn_layer_fwd_pool(const float* t_in, const cgnn_act_t* act, const float* W, const float* bias,
                            const cgnn_csr_t* csr, const int64_t* ptr, int64_t num_graphs, int64_t rows,
                            int32_t d_in, int32_t H, int32_t max_nodes, int32_t max_edges,
                            const cgnn_act_t* act_out, float* emb, cgnn_stream_t stream);

int cgnn_sage_layer_fwd(const float* t_in, const cgnn_act_t* act, const float* W, const float* bias,
                        const cgnn_csr_t* csr, const int64_t* ptr, int64_t num_graphs, int64_t rows,
                        int32_t d_in, int32_t H, int32_t max_nodes, int32_t max_edges, float* z, double* bn_stats,
                        float* agg, void* workspace, size_t workspace_bytes, cgnn_stream_t stream);

/* ---- K9: the whole eval-mode network in one kernel (reference train.py:56-68 forward; models.py:203-216 with
 * BatchNorm1d in eval mode and dropout off).  A unit - one subject, or several small ones - goes through all layers,
 * the mean-pool readout (models.py:57-59) and the MLP head (models.py:196-201) without its activations leaving the SM:
 * per subject HBM sees 4 N F bytes of node features, the packed in-edge records of csr->agg_in and the outputs.
 * kind 0 = GCN (hidden width 64, 2..4 layers, F <= 8, subjects of <= 384 nodes); anything else returns
 * CGNN_ERR_UNSUPPORTED and the caller runs the per-layer entry points.  emb [B, H] and / or logits [B, K] (either may be
 * NULL).  Results agree with the per-layer path to fp32 round-off and do not depend on how subjects are batched.
 * workspace: (L - 1) * 32768 + 1024 L bytes, 16-byte aligned. */
typedef struct {
  const float* W;             /* GCN [H, d_in] (layer 0: d_in = num_features) */
  const float* bias;          /* [H] */
  const float* gamma;         /* BatchNorm1d after the layer: weight, bias, running_mean, running_var [H], eps */
  const float* beta;
  const float* running_mean;
  const float* running_var;
  float eps;
} cgnn_eval_layer_t;
int cgnn_eval_fused_fwd(int32_t kind, const float* x, int32_t num_features, const cgnn_eval_layer_t* layers,
                        int32_t num_layers, int32_t H, const float* W0, const float* b0, const float* W1, const float* b1,
                        int32_t M, int32_t K, const cgnn_csr_t* csr, const int64_t* ptr, int64_t num_graphs, int64_t rows,
                        int32_t max_nodes, int32_t max_edges, float* emb, float* logits, void* workspace,
                        size_t workspace_bytes, cgnn_stream_t stream);

/* ---- K10: the dense contraction on tensor cores ---------------------------------------------
 * P[rows, N] = X[rows, K] W[N, K]^T (reference models.py:111 `self.linear(x)`), fp32-grade: three TF32
 * tcgen05 MMAs per K step on hi/lo splits of both operands, fp32 accumulators in tensor memory.
 * K and N multiples of 32, <= 256; pointers 16-byte aligned.  The layer kernels embed the same sequence;
 * this entry exposes it on its own (and is how the descriptor plumbing is tested). */
int cgnn_project_tf32x3(const float* X, const float* W, int64_t rows, int32_t K, int32_t N, float* P,
                        cgnn_stream_t stream);

/* ---- K3: BatchNorm1d bookkeeping (reference models.py:191-193, 208, 260) -------------- */

/* Merge `parts` statistic records [parts, 1+2C] (one per rank) into one (Chan, fp64). */
int cgnn_bn_merge_stats(const double* stats_parts, int32_t parts, int32_t C, double* stats,
                        cgnn_stream_t stream);

/* Training mode: stats -> scale/shift (+ mean, rstd for backward) and the running-stat
 * update  r = (1-m) r + m * stat  with the unbiased variance, num_batches_tracked += 1. */
int cgnn_bn_finalize(const double* stats, const float* gamma, const float* beta, int32_t C,
                     float eps, float momentum, float* running_mean, float* running_var,
                     int64_t* num_batches_tracked, float* scale, float* shift, float* mean,
                     float* rstd, cgnn_stream_t stream);

/* Eval mode: scale = gamma / sqrt(running_var + eps), shift = beta - running_mean * scale. */
int cgnn_bn_eval_affine(const float* gamma, const float* beta, const float* running_mean,
                        const float* running_var, int32_t C, float eps, float* scale, float* shift,
                        float* mean, float* rstd, cgnn_stream_t stream);

/* ---- K4: readout, head, loss (reference models.py:57-59, 196-201; train.py:49,64-66) -- */

/* emb[g] = sum_{i in g} act(t)[i] / (N_g + 1e-8)    -> emb [B, C] */
int cgnn_pool_fwd(const float* t_in, const cgnn_act_t* act, const int64_t* ptr, int64_t num_graphs,
                  int64_t rows, int32_t C, float* emb, cgnn_stream_t stream);

/* logits = W1 drop(relu(W0 emb + b0)) + b1.  W0 [M, C], W1 [K, M].  hidden [B, M] receives the
 * post-ReLU, post-dropout activations (kept for backward). */
int cgnn_head_fwd(const float* emb, const float* W0, const float* b0, const float* W1, const float* b1,
                  int64_t num_graphs, int32_t C, int32_t M, int32_t K, float p_drop, uint64_t seed,
                  int64_t graph_base, const uint32_t* salt, float* hidden, float* logits, cgnn_stream_t stream);

/* Cross entropy over graphs: loss_sum[0] = sum_g nll_g * inv_count (inv_count = 1/B_global gives
 * nn.CrossEntropyLoss()'s mean), correct[0] = #argmax == label (int64), nll [B] per graph.
 * labels [B] int64, one per graph (the host wrapper rejects shorter label vectors); a label outside
 * [0, K) - where the reference's nn.CrossEntropyLoss raises - makes nll_g and the loss NaN. */
int cgnn_ce_fwd(const float* logits, const int64_t* labels, int64_t num_graphs, int32_t K,
                float inv_count, float* nll, float* loss, int64_t* correct, cgnn_stream_t stream);

/* ---- K5-K7: backward ------------------------------------------------------------------ */

/* dlogits = (softmax - onehot) * inv_count * gout[0] */
int cgnn_ce_bwd(const float* logits, const int64_t* labels, int64_t num_graphs, int32_t K,
                float inv_count, const float* gout, float* dlogits, cgnn_stream_t stream);

/* Head backward: given dlogits [B,K] -> demb [B,C] and parameter grads (overwritten). */
int cgnn_head_bwd(const float* emb, const float* hidden, const float* dlogits, const float* W0,
                  const float* W1, int64_t num_graphs, int32_t C, int32_t M, int32_t K,
                  float p_drop, float* demb, float* dW0, float* db0, float* dW1, float* db1,
                  void* workspace, size_t workspace_bytes, cgnn_stream_t stream);

/* BatchNorm backward sums of one layer:  s[0:C] = sum dy, s[C:2C] = sum dy * xhat where
 * dy = dact(du) and du is either a per-row tensor du [rows, C] (demb == NULL) or the pooled
 * gradient demb [B, C] broadcast as demb[g]/(N_g + 1e-8) (du == NULL).  z is the layer's raw
 * (pre-BN) output; act describes the BN+ReLU+dropout applied to it; post_relu = 1 when z
 * itself is a ReLU output (SAGE) - it does not change the sums, only documents the caller. */
int cgnn_bn_bwd_sums(const float* z, const cgnn_act_t* act, const float* mean, const float* rstd,
                     const float* du, const float* demb, const int64_t* ptr, int64_t num_graphs,
                     int64_t rows, int32_t C, float* sums, double* sums64, void* workspace, size_t workspace_bytes,
                     cgnn_stream_t stream);   /* sums64 (optional): the same [2, C] in double */

/* GCN layer backward (autograd of models.py:84-114 plus the BN/ReLU/dropout that follows).
 *   upstream: du [rows,H] or pooled demb [B,H] (exactly one non-NULL), w.r.t. act_out(z).
 *   z [rows,H]       this layer's forward output, act_out / bn its BN+ReLU+dropout
 *   t_in [rows,d_in] this layer's stored input, act_in how it was transformed on load
 * Outputs: dW [H,d_in], dbias [H]; du_in [rows,d_in] = gradient w.r.t. act_in(t_in) (NULL for
 * the first layer); prev_sums [2*d_in] = BN backward sums of the previous layer computed
 * from du_in on the fly (NULL to skip; needs prev_mean/prev_rstd), prev_sums64 [2*d_in] (optional) the same in double.
 * scratch [rows, H] fp32 (may be NULL) holds dP = A^^T dz between the gather kernel and the tensor-core
 * contraction; without it (or without the blobs of cgnn_build_agg) the generic kernel runs. */
int cgnn_gcn_layer_bwd(const float* du, const float* demb, const float* z, const cgnn_act_t* act_out,
                       const cgnn_bn_bwd_t* bn, const float* t_in, const cgnn_act_t* act_in,
                       const float* W, const cgnn_csr_t* csr, const int64_t* ptr, int64_t num_graphs,
                       int64_t rows, int32_t d_in, int32_t H, int32_t max_nodes, int32_t max_edges,
                       float* dW, float* dbias, float* du_in,
                       const float* prev_mean, const float* prev_rstd, float* prev_sums, double* prev_sums64,
                       float* scratch, void* workspace, size_t workspace_bytes, cgnn_stream_t stream);

/* GraphSAGE layer backward (autograd of models.py:136-152 plus the BN/dropout that follows).
 * Same contract as cgnn_gcn_layer_bwd; W [H, 2*d_in].  scratch [2, rows, d_in] fp32 ([3, rows, 256] for the wide
 * layers, H = d_in = 256: the third plane carries dz) holds the
 * direct and neighbour parts of the input gradient between the two kernels of this call.  agg [rows, d_in] is the
 * aggregate cgnn_sage_layer_fwd stored: with it, the blobs of cgnn_build_agg and row_graph the call runs as tensor-core
 * contractions (dz on load, [d_u || d_agg] = dz W, dW, dbias) followed by the transposed gather kernel; NULL: the
 * generic kernels gather again. */
int cgnn_sage_layer_bwd(const float* du, const float* demb, const float* z, const cgnn_act_t* act_out,
                        const cgnn_bn_bwd_t* bn, const float* t_in, const float* agg, const cgnn_act_t* act_in,
                        const float* W, const cgnn_csr_t* csr, const int64_t* ptr, int64_t num_graphs,
                        int64_t rows, int32_t d_in, int32_t H, int32_t max_nodes, int32_t max_edges,
                        float* dW, float* dbias, float* du_in,
                        const float* prev_mean, const float* prev_rstd, float* prev_sums, double* prev_sums64,
                        float* scratch, void* workspace, size_t workspace_bytes, cgnn_stream_t stream);

/* SyncBN exchange over peer memory (one process per GPU of one NVLink / NVSwitch box): this rank's record of n doubles is
 * written into every peer's symmetric buffer, flags are exchanged, and the world's records are merged in rank order - one
 * kernel instead of an NCCL collective plus a merge kernel.  peer_buffers: DEVICE array of `world` pointers to the ranks'
 * symmetric buffers (torch.distributed._symmetric_memory: handle.buffer_ptrs_dev), each
 *   slots * world * rec_cap doubles of records followed by slots * world uint32 flags, zero-initialised once.
 * seq: a number > 0 that grows by one with every use of `slot` on every rank; use two physical slots alternately for one
 * logical exchange.  mode 0: records are BatchNorm statistics {count, mean[C], M2[C]} (n = 1 + 2C), out = their exact
 * merge (as cgnn_bn_merge_stats); mode 1: out = the sum of the records.  *error is set to 1 if a peer's flag does not
 * arrive within a few seconds (the kernel then returns instead of hanging the device). */
int cgnn_peer_exchange(const uint64_t* peer_buffers, int32_t rank, int32_t world, int32_t slot, int32_t slots, int32_t rec_cap,
                       uint32_t seq, int32_t mode, const double* mine, int32_t n, int32_t C, double* out, int32_t* error,
                       cgnn_stream_t stream);

/* ---- the step after backward (reference train.py:51) -----------------------------------------
 * A training step that is replayed as a CUDA graph cannot take anything that changes from step to step as a launch
 * argument, so the step counter and the dropout salt live in a device state block of four uint64:
 *   [0] steps taken   [1] seed (set once by the host)   [2] = the two 32-bit salt words cgnn_act_t.salt points at   [3] spare
 * cgnn_step_tick: [0] += 1 and new salt words from (seed, step) - the first node of a graphed step.
 * cgnn_adam_step: torch.optim.Adam's update (weight decay added to the gradient, bias-corrected first / second moments)
 * over a FLAT parameter buffer of n floats, its flat gradient and moment buffers; the step number is read from state[0].
 * Same arithmetic, op for op in fp32, as torch's single-tensor implementation. */
int cgnn_step_tick(uint64_t* state, cgnn_stream_t stream);
int cgnn_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, const uint64_t* state, cgnn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CGNN_H_ */
