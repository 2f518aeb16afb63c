/*
 * lgc_b200.h -- C ABI of the B200-native LightGCN hot path (liblgc_b200.so).
 *
 * The reference (happykygo/GNN-eCommerce) has no FFI of its own: its boundary for this path is
 * the Python operator call `LGConv.forward(x, edge_index, edge_weight)` (src/lightgcn.py:96)
 * inside the `LightGCN` module (src/lightgcn.py:13-231), driven by
 * `TrainLightGCN.mini_batch_loop` (src/train_lightgcn.py:123-153) and `LightGCN.recommendK`
 * (src/lightgcn.py:169-182). Each entry point below names the reference lines it replaces.
 * The binding a maintainer adds on the reference side is the ctypes stub shown in
 * INTEGRATION.md (shipped as gnn_ecommerce_b200/_capi.py).
 *
 * Conventions
 *   - plain C: pointers + sizes; every pointer is a DEVICE pointer unless its name starts
 *     with `h_`; no torch types, no C++ exceptions cross this boundary;
 *   - every call takes an explicit `cudaStream_t` (passed as void*) and is asynchronous on it
 *     unless stated otherwise; no call allocates device memory except `lgc_graph_build*`
 *     (owned by the handle) -- workspaces are sized by `*_workspace_bytes` and passed in;
 *   - return value: 0 on success, negative `lgc_status` on failure; `lgc_last_error()` returns
 *     a thread-local message;
 *   - embedding tables are row-major fp32 `[num_nodes, ld]` with `ld % 4 == 0`, `ld >= d`,
 *     16-byte aligned base, padding columns (d..ld-1) zero. Node ids follow the reference:
 *     users 0..n_users-1, items n_users..N-1 (src/utils_v2.py:128).
 */
#ifndef LGC_B200_H
#define LGC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LGC_ABI_VERSION 1

typedef enum {
  LGC_OK = 0,
  LGC_ERR_INVALID = -1,      /* bad argument (null pointer, size, unsupported ld)           */
  LGC_ERR_CUDA = -2,         /* a CUDA runtime call or kernel launch failed                  */
  LGC_ERR_INDEX_RANGE = -3,  /* a node id is outside [0, num_nodes) (reference: device assert) */
  LGC_ERR_WORKSPACE = -4,    /* workspace too small                                          */
  LGC_ERR_UNSUPPORTED = -5
} lgc_status;

typedef struct lgc_graph lgc_graph_t; /* opaque, owns its device arrays */

int lgc_abi_version(void);
const char* lgc_last_error(void);
/* 1 if `ld` (floats per table row) has a kernel instantiation, else 0. */
int lgc_ld_supported(int ld);

/* Measurement hooks (bench.py): number of kernels this library has launched so far, and optional
 * per-kernel-class timing with CUDA events recorded on the launching stream around each launch.
 * Tags: 0-3 light-row SpMM by epilogue {plain, fwd-init, fwd-rmw, adam}, 4-7 heavy-row SpMM,
 * 8-11 split-row finish, 12 BPR, 13 misc, 14 peer-memory item exchange, 16-22 scoring {convert, gemm, threshold, rescore,
 * select, exhaustive, scan}.
 * lgc_profile_read synchronises the recorded events, sums milliseconds and launch counts per tag
 * into the HOST arrays and clears the record. */
long long lgc_launch_count(void);
int lgc_profile_enable(int on);
int lgc_profile_read(double* h_ms, long long* h_count, int n_tags);

/* ------------------------------------------------------------------ graph build (once per graph)
 * Replaces what PyG `gcn_norm` recomputes inside EVERY LGConv.forward call (K times per step;
 * call site src/lightgcn.py:96) on the edge list produced by `df_to_graph`
 * (src/utils_v2.py:146-165):
 *   deg[v]  = sum of w_e over edges whose target is v, accumulated in edge-list order (bit-exact
 *             to the CPU `scatter_add_`), dis = deg^-1/2 with inf -> 0,
 *   w_hat_e = (dis[src_e] * w_e) * dis[dst_e],
 * and a destination-major CSR (stable: edges of a row keep their list order).
 * `d_edge_index` is int64 [2, nnz] row-major (row 0 = source, row 1 = target);
 * `d_edge_weight` may be NULL (all ones). `normalize == 0` uses the weights as given
 * (already-normalised operator, e.g. the transpose of a non-symmetric graph).
 * Synchronises `stream` before returning (it reads back a few counters). */
int lgc_graph_build(int64_t num_nodes, int64_t nnz, const int64_t* d_edge_index,
                    const float* d_edge_weight, int normalize, void* stream,
                    lgc_graph_t** out_graph);
int lgc_graph_destroy(lgc_graph_t* graph);

/* Same graph straight from the E interaction pairs (the frame's `user_id_idx` / offset `item_id_idx`
 * columns, reference src/utils_v2.py:128,146-165): edge e is a[e] -> b[e], edge E + e is
 * b[e] -> a[e] with the same weight, i.e. the order in which df_to_graph concatenates the two
 * directions, so the result is bit-identical to lgc_graph_build on df_to_graph's output -- without
 * the [2, 2E] int64 COO intermediate. d_a / d_b: device int64 [n_pairs]. */
int lgc_graph_build_pairs(int64_t num_nodes, int64_t n_pairs, const int64_t* d_a, const int64_t* d_b,
                          const float* d_edge_weight, int normalize, void* stream,
                          lgc_graph_t** out_graph);

/* Rectangular operator for the row-partitioned multi-GPU path: rows = targets in [0, num_rows)
 * (a GPU's shard of destination nodes, ids shifted to the shard), sources in [0, num_cols) (ids
 * into the all-gathered table). Weights are used as given -- pass the w_hat of the global graph
 * (lgc_graph_get_info().w_hat / eid) so every shard carries the global normalisation. */
int lgc_graph_build_rect(int64_t num_rows, int64_t num_cols, int64_t nnz, const int64_t* d_edge_index,
                         const float* d_edge_weight, void* stream, lgc_graph_t** out_graph);

typedef struct {
  int64_t num_nodes;
  int64_t nnz;
  int32_t is_symmetric;      /* A_hat == A_hat^T (both directions, equal weights): backward may
                                reuse the forward operator                                     */
  int32_t light_max_degree;  /* rows up to this in-degree take the sub-warp path               */
  int64_t num_heavy_rows;    /* rows above it                                                  */
  int64_t num_chunks;        /* warp-sized work items the heavy rows were cut into             */
  int64_t num_split_rows;    /* heavy rows spanning more than one chunk (two-phase reduce)     */
  /* device arrays owned by the handle (valid until destroy) */
  const int32_t* rowptr;     /* [num_nodes+1]                                                  */
  const int32_t* src;        /* [nnz] source node of every CSR entry                           */
  const int32_t* eid;        /* [nnz] position of the entry in the input edge list             */
  const float* w_hat;        /* [nnz] normalised weight, CSR order                             */
  const float* deg;          /* [num_nodes] fp32 weighted in-degree                            */
  const float* dis;          /* [num_nodes] deg^-1/2, 0 for isolated nodes                     */
} lgc_graph_info;
int lgc_graph_get_info(const lgc_graph_t* graph, lgc_graph_info* info);

/* How the SpMM (the LGConv operator, src/lightgcn.py:96) splits the rows of this graph for row width
 * `ld` (built on the first query for that width): the high-degree rows go to the sweep kernel
 * (output rows pinned in shared memory, all CTAs walk the source table window by window), the rest
 * to the rows kernel (one sub-warp per row). `*_sources` = distinct source rows the class gathers
 * -- with the edge and row counts, the algorithmic bytes of one launch (bench.py). has_plan == 0:
 * no schedule for this width (ld % 16 != 0, or >= 2^25 source rows): the round-1 chunked kernels run. */
typedef struct {
  int32_t has_plan;
  int32_t sweep_slots_per_unit;
  int64_t sweep_units;
  int64_t sweep_rows, sweep_edges, sweep_sources;
  int64_t sweep_pieces_split_rows, sweep_partial_slots, sweep_iterations, sweep_windows;
  int64_t rows_rows, rows_edges, rows_sources;
} lgc_plan_info;
int lgc_graph_plan_info(const lgc_graph_t* graph, int ld, lgc_plan_info* info);

/* ------------------------------------------------------------------ LGConv (src/lightgcn.py:96)
 * y = A_hat x. x, y: [num_nodes, ld]; must not alias. Workspace holds the partial rows of the
 * hub rows that are split over several warps (usually a few MB; may be 0 bytes). */
size_t lgc_spmm_workspace_bytes(const lgc_graph_t* graph, int ld);
int lgc_spmm(const lgc_graph_t* graph, int ld, const float* x, float* y, void* workspace,
             size_t workspace_bytes, void* stream);

/* One LGConv layer with a fused epilogue (the building block of lgc_propagate / lgc_train_step,
 * exported for the sharded multi-GPU driver, which interleaves layers with NCCL all-gathers):
 *   s = (A_hat x)[r] for every row r of the graph (x: [num_cols, ld]; all other tables: [rows, ld])
 *   LGC_EPI_PLAIN    y = scale * s + beta * addend            (addend may be NULL)
 *   LGC_EPI_FWD_INIT acc = a0 * xrow + a1 * s;  y = s if y != NULL
 *   LGC_EPI_FWD_RMW  acc = acc + a1 * s;        y = s if y != NULL
 *   LGC_EPI_ADAM     g = scale * s + addend; torch.optim.Adam update of p, m, v with g
 *   LGC_EPI_FWD_FINAL acc = (((hist[0]*ah[0] + hist[1]*ah[1]) + ...) + a1 * s): the whole layer mean of
 *                    get_embedding (src/lightgcn.py:91-99) in the LAST layer's epilogue, from the
 *                    n_hist stored layer tables (E0, x_1, ..., x_{K-1}), same rounding sequence */
typedef enum { LGC_EPI_PLAIN = 0, LGC_EPI_FWD_INIT = 1, LGC_EPI_FWD_RMW = 2, LGC_EPI_ADAM = 3,
               LGC_EPI_FWD_FINAL = 4 } lgc_epilogue_mode;
typedef struct {
  int32_t mode;
  float a0, a1, scale, beta;
  float* y;
  float* acc;
  const float* xrow;
  const float* addend;
  float* p;
  float* m;
  float* v;
  double lr, beta1, beta2, eps;
  int64_t step;
  const float* adam_scalars;   /* optional DEVICE pointer to the 6 floats of lgc_adam_scalars(): lets a
                                  captured CUDA graph replay the launch with a new step count */
  const float* hist[6];        /* LGC_EPI_FWD_FINAL: layer tables x_0 .. x_{n_hist-1} ([rows, ld])        */
  float ah[6];                 /*                    their weights alpha_0 .. alpha_{n_hist-1}            */
  int32_t n_hist;
} lgc_spmm_epilogue;
int lgc_spmm_ex(const lgc_graph_t* graph, int ld, const float* x, const lgc_spmm_epilogue* epilogue,
                void* workspace, size_t workspace_bytes, void* stream);

/* The same fused epilogues applied to row sums that already exist: `sums` is a dense [n_rows, ld] table
 * (row r of `sums` plays the role of (A_hat x)[r]). The multi-GPU step uses it for the replicated item
 * rows, whose sums arrive through the all-reduce of the ranks' partial sums. */
int lgc_epilogue_apply(int64_t n_rows, int ld, const float* sums, const lgc_spmm_epilogue* epilogue, void* stream);

/* ------------------------------------------------------------------ get_embedding (src/lightgcn.py:91-99)
 * out = sum_{l=0..K} alpha_l A_hat^l x0, evaluated like the reference as a running sum but
 * fused into the SpMM epilogue. `h_alpha` is a HOST array of K+1 floats. Because A_hat is
 * symmetric for the reference's graphs the same call is the backward pass (x0 := dL/dout).
 * Workspace: lgc_propagate_workspace_bytes(). out must not alias x0. */
size_t lgc_propagate_workspace_bytes(const lgc_graph_t* graph, int ld, int num_layers);
int lgc_propagate(const lgc_graph_t* graph, int ld, int num_layers, const float* h_alpha,
                  const float* x0, float* out, void* workspace, size_t workspace_bytes,
                  void* stream);

/* ------------------------------------------------------------------ pair scores (src/lightgcn.py:123-125)
 * score[j] = <out[src_j], out[dst_j]>, j < n_pairs; `pairs` is int64 [2, n_pairs]. */
int lgc_pair_scores(int ld, const float* out, const int64_t* pairs, int64_t n_pairs, float* score,
                    void* stream);

/* ------------------------------------------------------------------ BPR loss + gradients
 * Replaces, for pre-sampled triples (users, pos, neg: int64 [batch]):
 *   forward gathers + dot (src/lightgcn.py:123-125), mean softplus(-(s+ - s-))
 *   (src/lightgcn.py:279-286, src/train_lightgcn.py:141), the layer-0 L2 term
 *   (src/utils_v2.py:193-211) and their backward (src/train_lightgcn.py:146).
 * loss3 <- {bpr, reg, bpr+reg}.
 * grad_out  [N, ld] += dL/d out          (rows of u, p, n; duplicates accumulate)
 * grad_e0   [N, ld] += alpha0 * dL/d out + (decay/batch) * multiplicity * e0   (may be NULL)
 * Both gradient buffers must be zero on entry at the rows touched. `touched` (int32
 * [3*batch], may be NULL) receives the touched row ids so a caller can re-zero sparsely. */
size_t lgc_bpr_workspace_bytes(int64_t batch);
int lgc_bpr_loss_grad(int64_t num_nodes, int ld, int64_t batch, const int64_t* users,
                      const int64_t* pos, const int64_t* neg, const float* out, const float* e0,
                      double decay, float alpha0, float* grad_out, float* grad_e0, int32_t* touched,
                      float* loss3, void* workspace, size_t workspace_bytes, void* stream);

/* table[idx[j], :] += rows[j, :] for j < n; duplicates are summed in input order (deterministic,
 * no atomics), entries with idx[j] < 0 are skipped. The multi-GPU step uses it to scatter the
 * <= 3*batch gradient rows into its (replicated or sharded) tables. rows: [n, ld]. */
int lgc_scatter_add_rows(int64_t n, int ld, const int64_t* idx, const float* rows, float* table, void* stream);

/* ------------------------------------------------------------------ Adam (src/train_lightgcn.py:58,147)
 * torch.optim.Adam defaults (no weight decay, no amsgrad), dense over `n` contiguous floats,
 * single pass: reads p, g, m, v; writes p, m, v. `step` is the 1-based step count. Hyper-
 * parameters are doubles (Python floats in the reference) and rounded to fp32 where torch does. */
int lgc_adam_step(int64_t n, float* p, const float* g, float* m, float* v, double lr, double beta1,
                  double beta2, double eps, int64_t step, void* stream);
/* The step-dependent scalars of that update as 6 HOST floats {1-beta1, beta2, 1-beta2,
 * sqrt(1-beta2^t), eps, -lr/(1-beta1^t)}, and the same update reading them from DEVICE memory
 * (launch arguments of a captured CUDA graph are frozen; the scalars are refreshed by a copy). */
int lgc_adam_scalars(double lr, double beta1, double beta2, double eps, int64_t step, float* h_out6);
int lgc_adam_step_dev(int64_t n, float* p, const float* g, float* m, float* v, const float* d_scalars6,
                      void* stream);

/* ------------------------------------------------------------------ fused training step
 * One iteration of `mini_batch_loop` (src/train_lightgcn.py:129-151) for pre-sampled triples:
 * K-layer forward, BPR + L2, K-layer backward (Horner form on the symmetric operator) with the
 * dense Adam update fused into the epilogue of the last backward layer. Requires a symmetric
 * graph. e0, m, v: [N, ld], updated in place. */
typedef struct {
  int32_t ld;
  int32_t num_layers;
  const float* h_alpha;   /* host, K+1 */
  int64_t batch;
  const int64_t* users;
  const int64_t* pos;
  const int64_t* neg;
  double decay;
  double lr, beta1, beta2, eps;
  int64_t step;           /* 1-based Adam step */
  float* e0;
  float* m;
  float* v;
  float* loss3;           /* device, 3 floats */
  void* workspace;
  size_t workspace_bytes;
} lgc_train_step_args;
size_t lgc_train_step_workspace_bytes(const lgc_graph_t* graph, int ld, int num_layers,
                                      int64_t batch);
/* Call once after allocating the workspace (zeroes the two sparse-gradient tables; every step
 * leaves them zero again). */
int lgc_train_workspace_init(const lgc_graph_t* graph, int ld, int num_layers, int64_t batch,
                             void* workspace, size_t workspace_bytes, void* stream);
int lgc_train_step(const lgc_graph_t* graph, const lgc_train_step_args* args, void* stream);

/* ------------------------------------------------------------------ batch_loader (src/utils_v2.py:168-181)
 * Device-side BPR sampler: `batch` triples (user, pos, neg), all outputs int64 on the device.
 *   users  distinct purchasers, uniform without replacement        (random.sample, :174)
 *   pos    uniform over the user's train purchases                 (random.choice, :178)
 *   neg    uniform item id + n_users outside the user's ignore list (rejection, :169-173,179)
 * Inputs are the CSR form of `train_pos_list_df` (src/utils_v2.py:64-89): purchasers [P];
 * pos_ptr [P+1] / pos_items (offset item ids); ign_ptr [P+1] / ign_items (offset item ids,
 * SORTED within a user). Draws are a pure function of (seed, step, triple index): the same call
 * reproduces the same batch. batch > P fails like random.sample ("Sample larger than population"). */
size_t lgc_sample_triples_workspace_bytes(int64_t n_purchasers, int64_t batch);
int lgc_sample_triples(int64_t n_purchasers, const int64_t* purchasers, const int64_t* pos_ptr,
                       const int64_t* pos_items, const int64_t* ign_ptr, const int64_t* ign_items,
                       int64_t n_users, int64_t n_items, int64_t batch, uint64_t seed, uint64_t step,
                       int64_t* users, int64_t* pos, int64_t* neg, void* workspace, size_t workspace_bytes,
                       void* stream);

/* ------------------------------------------------------------------ recommendK (src/lightgcn.py:169-182)
 * Top-k items for a list of users from the final embeddings:
 *   pred = user_emb[user_ids] @ item_emb^T                      (src/lightgcn.py:173)
 *   masked = pred * (1 - seen)   -- MULTIPLICATIVE mask: a seen item scores 0.0, not -inf (:175)
 *   top-k indices of every row                                   (:177)
 * The score matrix never reaches HBM: a tcgen05/TMEM fp16 GEMM (fp32 accumulate) keeps only the
 * maximum of every group of 16 consecutive items; the candidate groups that can hold a top-k
 * item (with a proven error margin) are re-scored exactly in fp32 with the mask applied, and a
 * user whose margin cannot be proven is re-scored against all items. Results equal the fp32
 * reference apart from exact ties / fp32 summation-order near-ties.
 * seen_ptr/seen_items: CSR over the scored users (row i belongs to user_ids[i]) of un-offset item
 * ids -- the sparse form of the dense `interactions_t` rows (src/utils_v2.py:92-103,137-138);
 * NULL = nothing seen. user_ids NULL = users 0..n_users-1.
 * stats (device, 4 x int64, may be NULL): {users re-scored exhaustively, candidate groups total,
 * 0, 0}. */
typedef struct {
  int32_t d;               /* embedding dim actually used (<= ld_user, ld_item)              */
  int32_t ld_user;         /* floats per row of user_emb                                     */
  int32_t ld_item;         /* floats per row of item_emb                                     */
  int32_t k;               /* 1..32                                                          */
  int64_t n_users;         /* users to score                                                 */
  int64_t n_items;
  const float* user_emb;   /* table the user ids index into                                  */
  const float* item_emb;   /* [n_items, ld_item]                                             */
  const int64_t* user_ids; /* [n_users] or NULL                                              */
  const int64_t* seen_ptr; /* [n_users + 1] or NULL                                          */
  const int64_t* seen_items;
  int64_t* topk_items;     /* [n_users, k] out, item ids 0..n_items-1, best first            */
  float* topk_scores;      /* [n_users, k] out, masked fp32 scores (may be NULL)             */
  int64_t* stats;          /* [4] out or NULL                                                */
  void* workspace;
  size_t workspace_bytes;
} lgc_score_topk_args;
size_t lgc_score_topk_workspace_bytes(int64_t n_users, int64_t n_items, int d, int k);
int lgc_score_topk(const lgc_score_topk_args* args, void* stream);

/* ------------------------------------------------------------------ MARK_MAPK (src/lightgcn.py:184-189)
 * Mean precision@k and recall@k over the evaluated users: overlap = |set(held) & set(top-k)|,
 * recall = overlap / len(held list), precision = overlap / k. held_ptr [n_users+1] / held_items:
 * CSR of the held-out purchases (row i belongs to the user of topk row i), un-offset item ids.
 * per_user (device, float [n_users, 2]) receives {precision, recall}; out2 (device, double[2])
 * the two means. */
int lgc_mark_mapk(int64_t n_users, int k, const int64_t* topk_items, const int64_t* held_ptr,
                  const int64_t* held_items, float* per_user, double* out2, void* stream);

/* ------------------------------------------------------------------ multi-GPU item-row exchange over peer memory
 * (SURVEY.md 8(b) last bullet / 8(e): the communicating variant of the LGConv layer; the reference
 * itself is single-device, src/train_lightgcn.py:35-37.) In the bipartite-sharded step every rank holds
 * PARTIAL sums of (A_hat x)[items] over its own users; lgc_item_exchange completes the aggregate of
 * PyG's LGConv (call site src/lightgcn.py:96) for the item rows and applies the fused epilogue in ONE
 * kernel over NVLink peer memory -- reduce-scatter by P2P loads in fixed rank order (deterministic),
 * epilogue on the owned slice, all-gather by P2P stores -- instead of ncclAllReduce followed by
 * lgc_epilogue_apply. Every rank must issue the same sequence of lgc_item_exchange calls.
 *
 * Memory: one ARENA per rank, allocated by this library (cudaMalloc, zero-filled) and exported as a
 * CUDA IPC handle (LGC_PEER_HANDLE_BYTES host bytes the caller passes to the other ranks however it
 * likes -- the Python driver all-gathers them over torch.distributed). lgc_peer_arena_open maps a
 * peer's arena into this process. The partial-sum table and every table the epilogue WRITES
 * (PLAIN: y; ADAM: p, m, v; FWD_FINAL: acc) must lie inside the arena at the same offset on every
 * rank; operands that are only read (addend, hist, adam_scalars) are local and may live anywhere.
 * ADAM stores the new weights p into every arena; the moments m, v of a row are only updated on the
 * rank that owns the row (rank r: rows [r * ceil(n_rows / world), ...)) -- they are sharded, not replicated. A
 * LGC_PEER_CTRL_BYTES control block inside the arena (zero at start, 256-byte aligned offset) carries
 * the barrier tickets; they are counted on the device, so a captured CUDA graph may replay the call.
 * Waits are bounded by timeout_ms (0: ~20 s): a timeout sets the error word read by
 * lgc_peer_exchange_status (synchronous) instead of hanging the GPU. */
#define LGC_PEER_MAX 8
#define LGC_PEER_HANDLE_BYTES 64
#define LGC_PEER_CTRL_BYTES 256
int lgc_peer_arena_alloc(size_t bytes, void** d_base, void* h_handle);
int lgc_peer_arena_open(const void* h_handle, void** d_peer_base);
int lgc_peer_arena_close(void* d_peer_base);
int lgc_peer_arena_free(void* d_base);
typedef struct {
  int32_t world, rank;
  void* bases[LGC_PEER_MAX];   /* arena of every rank as mapped in THIS process (bases[rank]: own)   */
  size_t arena_bytes;
  size_t ctrl_off;             /* byte offset of the control block inside every arena               */
  const float* part;           /* [n_rows, ld] partial sums inside the own arena                    */
  int64_t n_rows;              /* rank r reduces rows [r * ceil(n_rows / world), ...)               */
  int32_t ld;
  int32_t timeout_ms;
} lgc_peer_exchange;
int lgc_item_exchange(const lgc_peer_exchange* x, const lgc_spmm_epilogue* epilogue, void* stream);
int lgc_peer_exchange_status(const lgc_peer_exchange* x, int32_t* h_error, int64_t* h_epoch);

/* Diagnostics: clock64() cycles per phase of the light-row SpMM kernel since the last call, summed
 * over warps (only when the process runs with LGC_LIGHT_PHASES=1; zeros otherwise). out8[0..4] =
 * wait for CSR slices | gathers | wait for operand tiles | epilogue | store + refill; out8[5] =
 * warp-tiles. Synchronises the device. */
int lgc_debug_light_phases(unsigned long long* out8);

#ifdef __cplusplus
}
#endif
#endif /* LGC_B200_H */
