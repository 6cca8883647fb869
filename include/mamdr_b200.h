/*
 * mamdr_b200.h -- C-ABI of libmamdr_b200.so: the B200-native (sm_100a) replacement for the
 * TensorFlow-1.12 graph that the reference executes below its meta-learning wrappers.
 *
 * The reference (RManLuo/MAMDR) has no FFI; the seam is the wrapper protocol
 * (model_zoo/maml.py:153-194) plus one Keras train/eval function per mini-batch.  Each entry
 * point below cites the reference interface it replaces (paths relative to /root/reference).
 *
 * Conventions
 *   - extern "C"; every call returns int: 0 = MAMDR_OK, <0 = MAMDR_E_*; no exception crosses.
 *   - all pointers named *_dev / documented "device" are device pointers owned by the caller
 *     (PyTorch tensors' data_ptr()); the library never frees or retains them past the call.
 *   - all work is enqueued on the caller's stream (a cudaStream_t passed as void*) and is
 *     asynchronous; no host synchronisation and no allocation on hot calls (workspaces are sized
 *     by *_workspace_bytes and passed in), so every hot call is CUDA-graph capturable.
 *   - a ctx is bound to one device and is not thread-safe (one ctx per device per host thread).
 *   - no float atomics on any training path: results are bit-reproducible run to run and rank to
 *     rank (replicated DN phases on several GPUs stay bit-identical without communication).
 */
#ifndef MAMDR_B200_H
#define MAMDR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAMDR_ABI_VERSION 1

#define MAMDR_OK             0
#define MAMDR_E_INVALID     -1 /* bad argument (NULL, misaligned, out of range) */
#define MAMDR_E_CUDA        -2 /* CUDA runtime error; text in mamdr_last_error */
#define MAMDR_E_WORKSPACE   -3 /* workspace too small */
#define MAMDR_E_UNSUPPORTED -4 /* shape / mode not supported by this build */

#define MAMDR_MAX_LAYERS     8
#define MAMDR_AUC_ROWS       4 /* accumulator rows: tp, fp, fn, tn */

/* precision_mode of the tower GEMMs */
#define MAMDR_PREC_FP32      0 /* fp32 FFMA (SIMT) -- the parity mode */
#define MAMDR_PREC_TF32      1 /* tcgen05 kind::tf32, fp32 accumulate in TMEM -- speed mode */
#define MAMDR_PREC_TF32X3    2 /* tcgen05 3xTF32 error-compensated split -- ~fp32 accuracy */

/* merged_method (model_zoo/specific_base_model.py:164-172) */
#define MAMDR_MERGE_PLUS     0
#define MAMDR_MERGE_TIMES    1

typedef struct mamdr_ctx mamdr_ctx;
typedef void* mamdr_stream; /* cudaStream_t */

/* ---- context ------------------------------------------------------------------------------ */
int         mamdr_abi_version(void);
int         mamdr_ctx_create(mamdr_ctx** out, int device);
void        mamdr_ctx_destroy(mamdr_ctx* ctx);
const char* mamdr_last_error(const mamdr_ctx* ctx); /* ctx may be NULL: last create error */
int         mamdr_sm_count(const mamdr_ctx* ctx);
/* Number of CTAs the persistent pass kernel of THIS context launches (0 = one per SM, the default).  With n < SM count
 * several contexts can run their passes CONCURRENTLY on disjoint SMs of one GPU, each on its own stream -- the opt-in
 * "virtual ranks" mode (SURVEY.md 7.3 hard part 1(c): independent DR chains side by side; the semantics are those of the
 * multi-GPU sharded schedule, model_zoo/mamdr.py:59-108 with the chains of different query domains run in parallel). */
int         mamdr_ctx_set_pass_ctas(mamdr_ctx* ctx, int32_t n_ctas);

/* ---- K1: embedding gather  (replaces tf.gather under Embedding, DeepCTR/deepctr.py:125-128) --
 * out[i, 0:dim] = table[ids[i], 0:dim], i < n.  dim % 4 == 0, rows 16-byte aligned, out_stride in
 * floats (>= dim, % 4 == 0).  Bit-exact.  ids outside [0, rows) -> row of zeros is NOT produced:
 * the call is undefined for such ids (the reference would raise inside tf.gather). */
int mamdr_gather_f32(mamdr_ctx* ctx, const float* table_dev, int64_t rows, int32_t dim,
                     const int32_t* ids_dev, int64_t n, float* out_dev, int64_t out_stride,
                     mamdr_stream stream);

/* ---- K6: sparse-gradient de-duplication (replaces _deduplicate_indexed_slices = tf.unique +
 * unsorted_segment_sum in TF's optimizer).
 * uniq_ids = ascending unique ids (bit-exact); uniq_rows[k] = sum of grad rows whose id ==
 * uniq_ids[k], added sequentially in batch order (== numpy add.at order, deterministic).
 * n <= mamdr_scatter_max_n(); *n_uniq_dev receives the count (device int32).  Negative ids are padding and are
 * skipped together with their rows (fixed-capacity exchange buffers of the row-sharded tables). */
int64_t mamdr_scatter_max_n(void);
size_t  mamdr_scatter_workspace_bytes(int64_t n);
int mamdr_scatter_dedup_f32(mamdr_ctx* ctx, const int32_t* ids_dev, const float* grad_rows_dev,
                            int64_t grad_stride, int64_t n, int32_t dim, int32_t* uniq_ids_dev,
                            float* uniq_rows_dev, int32_t* n_uniq_dev, void* ws_dev, size_t ws_bytes,
                            mamdr_stream stream);

/* ---- routing for ROW-SHARDED tables (north_star: "tables that exceed a single GPU are row-sharded with NCCL all-to-all";
 * row r lives on rank r % world at local index r / world; the reference keeps whole tables in one TF variable,
 * DeepCTR/deepctr.py:105-126).  mamdr_route_plan: for the n local ids of one (or two: ids_b != NULL) id columns,
 * slot[i] = owner * cap + (number of earlier local ids with the same owner), send[slot[i]] = id / world and every other
 * entry of the cap-entry blocks = -1: fixed-capacity blocks, so the all-to-all splits are static.  n <= cap, world <= 64.
 * `block` is the distance between consecutive owner blocks of a send buffer: cap (separate buffers), or 2 * cap with
 * send_b = send_a + cap: both columns share ONE exchange buffer (owner block r = [cap entries of a | cap entries of b]) and
 * every slot indexes that shared buffer (slot_b values are offset by send_b - send_a): one all-to-all per direction for
 * both tables.  mamdr_route_pack_rows: dst[slot[i], 0:dim] = src[i, 0:dim] * scale (gradient rows into the exchange buffer).
 * mamdr_route_gather2: the owners' gather over a shared buffer -- out[e, :] = (entry e belongs to column b ? table_b :
 * table_a)[recv_idx[e], :] for recv_idx[e] >= 0, plus the received ids split per table for the de-duplication
 * (ids_a[e] = recv_idx[e] for a's entries, -1 elsewhere; ids_b likewise). */
int mamdr_route_plan(mamdr_ctx* ctx, const int32_t* ids_a_dev, const int32_t* ids_b_dev, int32_t n, int32_t world,
                     int32_t cap, int32_t block, int32_t* slot_a_dev, int32_t* slot_b_dev, int32_t* send_a_dev,
                     int32_t* send_b_dev, mamdr_stream stream);
int mamdr_route_gather2(mamdr_ctx* ctx, const float* table_a_dev, const float* table_b_dev, const int32_t* recv_idx_dev,
                        int32_t world, int32_t cap, int32_t dim, float* out_dev, int32_t* ids_a_dev, int32_t* ids_b_dev,
                        mamdr_stream stream);
int mamdr_route_pack_rows(mamdr_ctx* ctx, const float* src_dev, int64_t src_stride, const int32_t* slot_dev, int32_t n,
                          int32_t dim, float scale, float* dst_dev, mamdr_stream stream);

/* The same de-duplication for n beyond mamdr_scatter_max_n() (any n < 2^31), multi-CTA: a stable radix sort of
 * (id, position) + head scan, then ONE pass over the n gradient rows at HBM bandwidth.  uniq_ids: ascending, bit-exact.
 * Summation order (deterministic): windows of 256 consecutive SORTED positions; inside a window the rows of an id are
 * added sequentially in batch order; the window partials of an id are added in window order inside groups of 32
 * windows (counted from the id's first window), then the group sums in order (an id whose rows lie inside one
 * window: exactly the numpy add.at order). */
size_t mamdr_scatter_large_workspace_bytes(int64_t n, int32_t dim);
int mamdr_scatter_dedup_large_f32(mamdr_ctx* ctx, const int32_t* ids_dev, const float* grad_rows_dev,
                                  int64_t grad_stride, int64_t n, int32_t dim, int32_t* uniq_ids_dev,
                                  float* uniq_rows_dev, int32_t* n_uniq_dev, void* ws_dev, size_t ws_bytes,
                                  mamdr_stream stream);

/* ---- model description (replaces DeepCTR.build_inputs/build_emb/build_mlp,
 * model_zoo/DeepCTR/deepctr.py:95-136).  Offsets are in floats into a parameter "arena": one flat
 * fp32 buffer holding every trainable tensor in model.trainable_weights order
 * [user_emb?, item_emb?, domain_emb, kernel0.., bias0.., dense_kernel, global_bias], each tensor
 * start aligned to 32 floats, padding kept zero.  theta, theta_d, the live model, Adam m/v and
 * best-snapshots all use the same layout, so every meta op is one coalesced sweep. */
typedef struct {
    int32_t  n_layers;                       /* hidden layers, 1..MAMDR_MAX_LAYERS */
    int32_t  emb_dim[3];                     /* user, item, domain (each % 4 == 0) */
    int32_t  hidden[MAMDR_MAX_LAYERS];       /* widths (each % 4 == 0) */
    int32_t  n_domain;
    int32_t  emb_trainable;                  /* 0: user/item tables frozen, outside the arena */
    int64_t  n_uid, n_pid;
    float    dropout_rate;                   /* 0 disables dropout */
    uint32_t dropout_seed;                   /* layer l uses dropout_seed + l (DNN seed=1024) */
    float    l2_emb;                         /* embeddings_regularizer l2 (1e-5) */
    float    frozen_reg;                     /* l2_emb * (|E_u|^2 + |E_i|^2) for frozen tables */
    int64_t  off_user_emb, off_item_emb;     /* -1 when frozen */
    int64_t  off_domain_emb;
    int64_t  off_kernel[MAMDR_MAX_LAYERS];
    int64_t  off_bias[MAMDR_MAX_LAYERS];
    int64_t  off_dense_kernel, off_global_bias;
    int64_t  arena_floats;                   /* total arena length (multiple of 32) */
} mamdr_mlp_desc;

/* one mini-batch = a window of a domain-resident column store (utils/dataset.py:12-38: batch of
 * uid, pid, domain, label with one domain id per batch, ragged tail kept) */
typedef struct {
    const int32_t* uid_dev;     /* [n_d] */
    const int32_t* pid_dev;     /* [n_d] */
    const float*   label_dev;   /* [n_d] */
    const int32_t* order_dev;   /* [>= offset+rows] sample order of this pass, or NULL = identity */
    int64_t        offset;      /* first position of the batch inside order */
    int32_t        rows;        /* b, 1..max_batch */
    int32_t        domain;      /* domain id shared by the whole batch */
    int32_t        row0;        /* index of row 0 inside the (global) mini-batch: 0, or a rank's slice start when the batch is
                                   split data-parallel (row-sharded tables) -- the dropout masks are indexed by the global row */
    int32_t        reserved;
} mamdr_batch;

/* ---- optimizer state (replaces tf.train.AdamOptimizer's non-slot variables beta1_power /
 * beta2_power and Keras' iteration counter; DeepCTR/deepctr.py:54-55).  Device-resident so that
 * captured graphs can be replayed: {int64 step; float b1pow; float b2pow; ...}. */
size_t mamdr_opt_state_bytes(void);
int    mamdr_opt_state_init(mamdr_ctx* ctx, void* state_dev, float beta1, float beta2,
                            mamdr_stream stream);
int    mamdr_opt_state_read(mamdr_ctx* ctx, const void* state_dev, int64_t* step, float* b1pow,
                            float* b2pow, mamdr_stream stream); /* synchronises the stream */

/* ---- K2-K5,K8: one training mini-batch, forward + backward (replaces the Keras train function
 * behind Model.train_on_batch / Model.fit as driven by model_zoo/mamdr.py:54,85-97 and
 * model_zoo/domain_negotiation.py:71-72), WITHOUT the optimizer apply.
 *   grads_dev  : arena-shaped; every dense segment is overwritten.  With emb_trainable the table
 *                segments are not written; the sparse part is returned de-duplicated in the
 *                workspace (see mamdr_mlp_sparse_grads) and consumed by mamdr_adam_table_step.
 *   loss_dev   : 1 float, mean BCE + L2 penalties of this batch (Keras 'loss' output)
 *   probs_dev  : optional [rows] sigmoid outputs
 *   auc_acc_dev: optional [4, num_thresholds] fp32 accumulators updated with this batch
 *                (utils/metrics_utils.py:297-354); thresholds_dev = the fp32 threshold table.
 * Dropout masks: Philox4x32-10, key=(dropout_seed+layer, step), see oracle/philox.py. */
size_t mamdr_mlp_workspace_bytes(const mamdr_mlp_desc* desc, int32_t max_batch);
int mamdr_mlp_train_step(mamdr_ctx* ctx, const mamdr_mlp_desc* desc, const mamdr_batch* batch,
                         const float* user_table_dev, const float* item_table_dev,
                         const float* params_dev, float* grads_dev, void* ws_dev, size_t ws_bytes,
                         const void* opt_state_dev, float* loss_dev, float* probs_dev,
                         float* auc_acc_dev, const float* thresholds_dev, int32_t num_thresholds,
                         int32_t precision_mode, mamdr_stream stream);

/* With emb_trainable the sparse user (table 0) / item (table 1) gradients of the last mamdr_mlp_train_step are left
 * de-duplicated in the workspace (sorted unique ids, summed rows [n_uniq, emb_dim], device count); `rows` must be
 * the batch's row count.  They feed mamdr_adam_table_step. */
int mamdr_mlp_sparse_grads(const mamdr_mlp_desc* desc, int32_t rows, void* ws_dev, int32_t table,
                           const int32_t** uniq_ids_dev, const float** uniq_rows_dev,
                           const int32_t** n_uniq_dev);

/* Gradient rows of the gathered user / item embeddings of the LAST mamdr_mlp_train_step (`rows` rows):
 * dX_out[r, 0:du+di] = dZ_0[r, :] . W_0[0:du+di, :]^T.  For tables that live outside the arena -- row-sharded across
 * GPUs (mamdr_b200/sharded.py): the rows travel back to their owners by NCCL all-to-all, are de-duplicated with
 * mamdr_scatter_dedup_f32 and applied with mamdr_adam_table_step on the owner's shard. */
int mamdr_mlp_input_grads(mamdr_ctx* ctx, const mamdr_mlp_desc* desc, int32_t rows, const float* params_dev,
                          void* ws_dev, size_t ws_bytes, float* dX_out_dev, mamdr_stream stream);

/* ---- inference mini-batch (replaces the Keras test function behind Model.evaluate,
 * model_zoo/specific_base_model.py:82-85, model_zoo/base_model.py:130-133): no dropout. */
int mamdr_mlp_eval_step(mamdr_ctx* ctx, const mamdr_mlp_desc* desc, const mamdr_batch* batch,
                        const float* user_table_dev, const float* item_table_dev,
                        const float* params_dev, void* ws_dev, size_t ws_bytes, float* loss_dev,
                        float* probs_dev, float* auc_acc_dev, const float* thresholds_dev,
                        int32_t num_thresholds, int32_t precision_mode, mamdr_stream stream);

/* ---- STAR tower (BASELINE config #4; replaces the Keras train / test function of model_zoo/Star/star.py:70-113:
 * PartitionedNorm -> StarFCN x L -> Dense(1, sigmoid); Star/partitioned_norm.py:102-203, Star/star_fcn.py:105-139).
 * Arena order = Keras creation order: [domain_emb, gamma_specific [D,n], beta_specific [D,n], gamma_shared [n],
 * beta_shared [n], (kernel_specific [D,in,out], bias_specific [D,out], kernel_shared [in,out], bias_shared [out]) x L,
 * out_kernel [h_L,1], out_bias [1]], n = sum(emb_dim); user / item tables frozen, outside the arena.  pn_state_dev
 * (mamdr_star_state_bytes, non-trainable): moving mean / variance and their zero-debiased accumulators per domain;
 * initialise moving_var (second [D,n] block) to 1 and everything else to 0.  fp32 only.  The train step overwrites
 * the WHOLE gradient arena (untouched domain slices are zero) and is followed by mamdr_adam_step over the arena. */
typedef struct {
    int32_t n_layers;
    int32_t emb_dim[3];
    int32_t hidden[MAMDR_MAX_LAYERS];
    int32_t n_domain;
    int64_t n_uid, n_pid;
    float   pn_eps, pn_momentum;             /* 1e-3, 0.99 (Keras BatchNormalization defaults) */
    int64_t off_domain_emb, off_gamma_sp, off_beta_sp, off_gamma_sh, off_beta_sh;
    int64_t off_ksp[MAMDR_MAX_LAYERS], off_bsp[MAMDR_MAX_LAYERS], off_ksh[MAMDR_MAX_LAYERS], off_bsh[MAMDR_MAX_LAYERS];
    int64_t off_out_kernel, off_out_bias;
    int64_t arena_floats;
} mamdr_star_desc;

size_t mamdr_star_workspace_bytes(const mamdr_star_desc* desc, int32_t max_batch);
size_t mamdr_star_state_bytes(const mamdr_star_desc* desc);
/* debug hook: workspace byte offsets [X, xhat, dY, H_0..H_L, dZ_0..dZ_{L-1}] for a batch of `rows` */
int    mamdr_star_debug_offsets(const mamdr_star_desc* desc, int32_t rows, int64_t* out);
int mamdr_star_train_step(mamdr_ctx* ctx, const mamdr_star_desc* desc, const mamdr_batch* batch,
                          const float* user_table_dev, const float* item_table_dev, const float* params_dev,
                          float* grads_dev, void* pn_state_dev, void* ws_dev, size_t ws_bytes, float* loss_dev,
                          float* probs_dev, float* auc_acc_dev, const float* thresholds_dev,
                          int32_t num_thresholds, mamdr_stream stream);
int mamdr_star_eval_step(mamdr_ctx* ctx, const mamdr_star_desc* desc, const mamdr_batch* batch,
                         const float* user_table_dev, const float* item_table_dev, const float* params_dev,
                         void* pn_state_dev, void* ws_dev, size_t ws_bytes, float* loss_dev, float* probs_dev,
                         float* auc_acc_dev, const float* thresholds_dev, int32_t num_thresholds,
                         mamdr_stream stream);

/* ---- multi-task towers (BASELINE config #5; replace the Keras train / test function of the per-domain sub-models
 * `Model(inputs, outputs[t])` that model_zoo/DeepMTLCTR/deep_mtl_ctr.py:57-65 compiles over deepctr's MMOE / PLE
 * (num_levels = 1) / SharedBottom (:25-48), all sharing one AdamOptimizer (:53)).  fp32.
 *   X -> the k experts domain t mixes (each DNN(expert_hidden)) ; gate t: DNN(gate_hidden)(X) . G_t -> softmax [k] ;
 *   mix = sum_j a_j * expert_j ; tower t: DNN(tower_hidden)(mix) . w_t + g_t -> sigmoid ; BCE (+ embedding l2).
 * mamdr_mtl_desc holds the dimensions; mamdr_mtl_domain the arena offsets of ONE domain's sub-model (its experts in
 * gate-column order, gate, tower) -- the host builds one per domain.  has_gate = 0 (SharedBottom): k = 1, no gate.
 * The train step writes the gradients of the sub-model's dense variables into grads_dev (positions of variables that
 * output t does not reach are left untouched; the optimizer never reads them) and, with emb_trainable, leaves the
 * de-duplicated sparse user / item gradients in the workspace (mamdr_mtl_sparse_grads).  Follow it with
 * mamdr_adam_table_step per table (trainable tables) and ONE mamdr_adam_ranges_step over the sub-model's arena
 * ranges: TF applies Adam to the variables of sub-model t only -- all other variables keep their value and slots --
 * while the optimizer's beta powers advance once per step.
 * Dropout streams (philox.cuh): seed = dropout_seed + 8*e + l (expert e), + 4096 + 8*t + l (gate t),
 * + 8192 + 8*t + l (tower t). */
#define MAMDR_MTL_MAX_K 8
typedef struct {
    int32_t emb_dim[3];
    int32_t n_domain;
    int64_t n_uid, n_pid;
    int32_t emb_trainable;
    int32_t has_gate;
    int32_t k;                               /* experts mixed per domain, 1..MAMDR_MTL_MAX_K */
    int32_t n_expert_layers, n_gate_layers, n_tower_layers;        /* each 1..MAMDR_MAX_LAYERS (gate: if has_gate) */
    int32_t expert_hidden[MAMDR_MAX_LAYERS], gate_hidden[MAMDR_MAX_LAYERS], tower_hidden[MAMDR_MAX_LAYERS];
    float   dropout_rate;
    uint32_t dropout_seed;
    float   l2_emb, frozen_reg;
    int64_t off_user_emb, off_item_emb, off_domain_emb;
    int64_t arena_floats;
} mamdr_mtl_desc;

typedef struct {
    int32_t domain;
    int32_t expert_id[MAMDR_MTL_MAX_K];                            /* global expert index (dropout stream) */
    int64_t off_expert_kernel[MAMDR_MTL_MAX_K][MAMDR_MAX_LAYERS], off_expert_bias[MAMDR_MTL_MAX_K][MAMDR_MAX_LAYERS];
    int64_t off_gate_kernel[MAMDR_MAX_LAYERS], off_gate_bias[MAMDR_MAX_LAYERS], off_gate_out;   /* gate_out [g_last, k] */
    int64_t off_tower_kernel[MAMDR_MAX_LAYERS], off_tower_bias[MAMDR_MAX_LAYERS], off_tower_out, off_bias;
} mamdr_mtl_domain;

size_t mamdr_mtl_workspace_bytes(const mamdr_mtl_desc* desc, int32_t max_batch);
int mamdr_mtl_train_step(mamdr_ctx* ctx, const mamdr_mtl_desc* desc, const mamdr_mtl_domain* dom,
                         const mamdr_batch* batch, const float* user_table_dev, const float* item_table_dev,
                         const float* params_dev, float* grads_dev, void* ws_dev, size_t ws_bytes,
                         const void* opt_state_dev, float* loss_dev, float* probs_dev, float* auc_acc_dev,
                         const float* thresholds_dev, int32_t num_thresholds, mamdr_stream stream);
int mamdr_mtl_eval_step(mamdr_ctx* ctx, const mamdr_mtl_desc* desc, const mamdr_mtl_domain* dom,
                        const mamdr_batch* batch, const float* user_table_dev, const float* item_table_dev,
                        const float* params_dev, void* ws_dev, size_t ws_bytes, float* loss_dev, float* probs_dev,
                        float* auc_acc_dev, const float* thresholds_dev, int32_t num_thresholds,
                        mamdr_stream stream);
/* views of the de-duplicated sparse gradients left by the last train step of `rows` rows (table 0 = user, 1 = item) */
int mamdr_mtl_sparse_grads(const mamdr_mtl_desc* desc, int32_t rows, void* ws_dev, int32_t table,
                           const int32_t** uniq_ids_dev, const float** uniq_rows_dev, const int32_t** n_uniq_dev);
/* Gradient rows of the gathered user / item embeddings of the LAST mamdr_mtl_train_step of sub-model `dom` (`rows` rows):
 * dX_out[r, 0:du+di].  For tables that live outside the arena -- row-sharded across GPUs (mamdr_b200/sharded.py); the
 * counterpart of mamdr_mlp_input_grads. */
int mamdr_mtl_input_grads(mamdr_ctx* ctx, const mamdr_mtl_desc* desc, const mamdr_mtl_domain* dom, int32_t rows,
                          const float* params_dev, void* ws_dev, size_t ws_bytes, float* dX_out_dev, mamdr_stream stream);
/* TF ApplyAdam over n_ranges (<= 16) arena ranges [begin, begin + len) (floats, multiples of 4) of one arena; the beta
 * powers / step advance once.  begin / len are HOST arrays. */
int mamdr_adam_ranges_step(mamdr_ctx* ctx, float* params_dev, float* m_dev, float* v_dev, const float* grads_dev,
                           const int64_t* begin, const int64_t* len, int32_t n_ranges, void* opt_state_dev,
                           float lr, float beta1, float beta2, float eps, mamdr_stream stream);

/* ---- one whole domain pass in ONE persistent cooperative launch (tcgen05 modes, frozen tables) ----
 * mamdr_mlp_train_pass replaces `model.fit(train_iter, steps_per_epoch=S)` (model_zoo/mamdr.py:54) and
 * the `for step in range(train_step): model.train_on_batch(train_iter)` loops (mamdr.py:85-97,
 * model_zoo/domain_negotiation.py:71-72): `steps` consecutive mini-batches (batch s = positions
 * [s*batch_size, min((s+1)*batch_size, n_data)) of order_dev, or of the split itself when order_dev is
 * NULL) of forward + BCE head + backward + optimizer apply (optimizer 0 = Adam, TF ApplyAdam order of
 * operations; 1 = plain SGD of the finetune stage, model_zoo/specific_base_model.py:120).
 * mamdr_mlp_eval_pass replaces `model.evaluate(dataset, steps)` (specific_base_model.py:82-85,
 * base_model.py:130-133): inference forward, per-batch losses, optional probabilities [n rows].
 *   losses_dev  : [steps] Keras loss value of every mini-batch
 *   auc_acc_dev : optional [4, num_thresholds] accumulators, updated once with the whole pass
 *   grads_dev   : optional arena-shaped; receives the gradients of the LAST mini-batch (test hook)
 *   ws_dev      : mamdr_mlp_pass_workspace_bytes(desc, batch_size) bytes, zero-initialised once by the
 *                 caller (rows past a ragged batch are multiplied by zeros, so they must be finite)
 * precision_mode must be MAMDR_PREC_TF32 or MAMDR_PREC_TF32X3.  Both calls enqueue a 64-byte memset and
 * one cooperative kernel launch on `stream`; the device must be otherwise idle enough for one CTA per
 * SM to be co-resident (cudaLaunchCooperativeKernel fails otherwise -> MAMDR_E_CUDA).
 * mamdr_mlp_pass_supported returns MAMDR_OK when the descriptor fits the pass kernel (frozen tables,
 * hidden widths 32 or multiples of 64, last width 32/64, user+item width a multiple of 32). */
typedef struct {
    const int32_t* uid_dev;     /* [n_data] */
    const int32_t* pid_dev;     /* [n_data] */
    const float*   label_dev;   /* [n_data] */
    const int32_t* order_dev;   /* [n_data] sample order of this pass, or NULL = identity */
    int64_t        n_data;
    int32_t        batch_size;
    int32_t        steps;       /* 1 .. ceil(n_data / batch_size) */
    int32_t        domain;
} mamdr_pass;

size_t mamdr_mlp_pass_workspace_bytes(const mamdr_mlp_desc* desc, int32_t max_batch);
int    mamdr_mlp_pass_supported(mamdr_ctx* ctx, const mamdr_mlp_desc* desc, int32_t max_batch);
int mamdr_mlp_train_pass(mamdr_ctx* ctx, const mamdr_mlp_desc* desc, const mamdr_pass* pass,
                         const float* user_table_dev, const float* item_table_dev, float* params_dev,
                         float* m_dev, float* v_dev, float* grads_dev, void* ws_dev, size_t ws_bytes,
                         void* opt_state_dev, float* losses_dev, float* auc_acc_dev,
                         const float* thresholds_dev, int32_t num_thresholds, int32_t optimizer, float lr,
                         float beta1, float beta2, float eps, int32_t precision_mode, mamdr_stream stream);
int mamdr_mlp_eval_pass(mamdr_ctx* ctx, const mamdr_mlp_desc* desc, const mamdr_pass* pass,
                        const float* user_table_dev, const float* item_table_dev, const float* params_dev,
                        void* ws_dev, size_t ws_bytes, void* opt_state_dev, float* losses_dev,
                        float* probs_dev, float* auc_acc_dev, const float* thresholds_dev,
                        int32_t num_thresholds, int32_t precision_mode, mamdr_stream stream);

/* Trainable user / item tables (desc->emb_trainable, config #2) through the pass kernel: ONE mini-batch per launch
 * (steps == 1, not recordable).  The tables are read from the arena, the tcgen05 chain also produces dX = dZ_0 .
 * W_0[0:K0]^T, the dense variables take their optimizer apply in-kernel, and the launch is followed by the
 * de-duplication of the two sparse gradients; mamdr_mlp_pass_sparse_grads returns them (sorted unique ids, summed
 * rows, device count) for mamdr_adam_table_step, which must read the beta powers of BEFORE the launch (pass a copy
 * of the optimizer state taken before it). */
int mamdr_mlp_pass_sparse_grads(const mamdr_mlp_desc* desc, int32_t batch_size, void* ws_dev, int32_t table,
                                const int32_t** uniq_ids_dev, const float** uniq_rows_dev,
                                const int32_t** n_uniq_dev);

/* ---- a whole meta-step in ONE launch: deferred execution of passes and meta sweeps ------------------------------
 * Between mamdr_program_begin and mamdr_program_end, mamdr_mlp_train_pass and the K9/K10 meta sweeps (mamdr_copy,
 * mamdr_merge, mamdr_dn_update, mamdr_dr_update, mamdr_dr_accumulate, mamdr_dr_apply_accum, mamdr_sub,
 * mamdr_axpy_diff) are RECORDED in call order instead of launched; mamdr_program_end uploads the op list into
 * ops_dev (>= n_ops * mamdr_program_op_bytes() bytes of device memory, 16-byte aligned) and runs it as one
 * persistent cooperative kernel (the body of a MAMDR / DN meta-step, model_zoo/mamdr.py:48-108, without the ~130
 * launch + prologue round trips).  All recorded passes must share model, workspace, batch size, optimizer and
 * precision; every pointer handed to a recorded call must stay valid until the launch has run.  Other entry points
 * return MAMDR_E_INVALID while recording.  A program without a pass replays its sweeps as ordinary launches. */
int     mamdr_program_begin(mamdr_ctx* ctx);
int     mamdr_program_end(mamdr_ctx* ctx, void* ops_dev, size_t ops_dev_bytes, int32_t* n_ops_out, mamdr_stream stream);
void    mamdr_program_abort(mamdr_ctx* ctx);
int64_t mamdr_program_op_bytes(void);

/* debug hook: when buf_dev != NULL the next pass launches write globaltimer stamps
 * [step][phase][cta][2] (phase start, jobs done) into buf_dev (capacity in uint64 words). */
int mamdr_debug_pass_timing(mamdr_ctx* ctx, void* buf_dev, int64_t capacity_u64);

/* ---- K7: optimizer apply over a flat arena (replaces AdamOptimizer.apply_gradients, TF
 * ApplyAdam kernel order of operations; bit-exact vs the oracle for equal g).
 * Increments state.step and the beta powers once per call. */
int mamdr_adam_step(mamdr_ctx* ctx, float* params_dev, float* m_dev, float* v_dev,
                    const float* grads_dev, int64_t n, void* opt_state_dev, float lr, float beta1,
                    float beta2, float eps, mamdr_stream stream);
/* ---- K6 + K7 fused for a trainable embedding table (replaces TF's aggregation of the IndexedSlices gather gradient
 * with the dense l2 regulariser gradient and the NON-lazy Adam apply over the whole variable,
 * DeepCTR/deepctr.py:54-55,104-126):
 *   g[r,:] = 2*l2*E[r,:] (+ uniq_rows[k,:] if r == uniq_ids[k]);  Adam on every row;  *loss_dev += l2*sum(E^2)
 * (pre-update values; loss_dev optional).  Reads the beta powers of opt_state_dev without advancing them: call it
 * BEFORE the mamdr_adam_step of the same mini-batch.  slot_map_dev: int32 [rows], all -1 between calls (the call
 * fills and clears it).  max_uniq bounds *n_uniq_dev (0 = no sparse part).  24 B per element: HBM-bound. */
size_t mamdr_adam_table_workspace_bytes(void);
int mamdr_adam_table_step(mamdr_ctx* ctx, float* table_dev, float* m_dev, float* v_dev, int64_t rows, int32_t dim,
                          const int32_t* uniq_ids_dev, const float* uniq_rows_dev, const int32_t* n_uniq_dev,
                          int64_t max_uniq, int32_t* slot_map_dev, float l2, const void* opt_state_dev, float lr,
                          float beta1, float beta2, float eps, float* loss_dev, void* ws_dev, size_t ws_bytes,
                          mamdr_stream stream);

/* sum(x^2) in double with a fixed reduction order (the l2 penalty of a trainable table in `evaluate`); ws as for
 * mamdr_adam_table_step; *out_dev is a device double */
/* The same sweep for the finetune stage's plain SGD (GradientDescentOptimizer, specific_base_model.py:120,
 * base_model.py:69) on a trainable table: table[r,:] -= (2*l2*table[r,:] (+ uniq_rows[k,:] if r == uniq_ids[k])) * lr on
 * every row; *loss_dev += l2*sum(E^2) of the pre-update values.  8 B per element. */
int mamdr_sgd_table_step(mamdr_ctx* ctx, float* table_dev, int64_t rows, int32_t dim, const int32_t* uniq_ids_dev,
                         const float* uniq_rows_dev, const int32_t* n_uniq_dev, int64_t max_uniq, int32_t* slot_map_dev,
                         float l2, float lr, float* loss_dev, void* ws_dev, size_t ws_bytes, mamdr_stream stream);
int mamdr_sum_squares_f64(mamdr_ctx* ctx, const float* x_dev, int64_t n, double* out_dev, void* ws_dev,
                          size_t ws_bytes, mamdr_stream stream);

/* plain SGD of the finetune stage (train.GradientDescentOptimizer,
 * model_zoo/specific_base_model.py:120, model_zoo/base_model.py:69); also bumps state.step */
int mamdr_sgd_step(mamdr_ctx* ctx, float* params_dev, const float* grads_dev, int64_t n,
                   void* opt_state_dev, float lr, mamdr_stream stream);

/* ---- K9/K10: meta ops over arenas of identical layout (n floats) ------------------------------
 * mamdr_copy       : dst <- src                     (SetVarOp / K.batch_get_value round trips,
 *                                                    utils/tool.py:36-45, model_zoo/maml.py:181-194)
 * mamdr_merge      : out <- theta (+|*) theta_i     (specific_base_model.py:164-172)
 * mamdr_dn_update  : theta += (model - theta)*beta; model_out <- theta  (model_out may be NULL)
 *                                                   (domain_negotiation.py:118-123, mamdr.py:57)
 * mamdr_dr_update  : merged = theta (+|*) theta_i; theta_i += (model - merged)*beta;
 *                    model_out <- theta (+|*) theta_i(new)               (mamdr.py:103-105,173-180)
 * mamdr_dr_accumulate: accum += (model - merged) [* theta for 'times']   (mamdr.py:182-191)
 * mamdr_dr_apply_accum: theta_i += accum / sample_num * beta; accum <- 0 (mamdr.py:193-196)
 * mamdr_sub        : out <- a - b                   (mamdr.py:168-171, finetune_every_epoch)
 * mamdr_axpy_diff  : out += (a - b) * alpha         (generic form of mamdr.py:173-180 with an explicit
 *                                                    `merged_weights`; out may alias b)
 * All are bit-exact vs numpy fp32 (explicit round-to-nearest mul/add, no FMA contraction). */
int mamdr_copy(mamdr_ctx* ctx, float* dst_dev, const float* src_dev, int64_t n, mamdr_stream stream);
int mamdr_merge(mamdr_ctx* ctx, float* out_dev, const float* theta_dev, const float* theta_i_dev,
                int64_t n, int32_t merged_method, mamdr_stream stream);
int mamdr_dn_update(mamdr_ctx* ctx, float* theta_dev, const float* model_dev, float beta, int64_t n,
                    float* model_out_dev, mamdr_stream stream);
int mamdr_dr_update(mamdr_ctx* ctx, float* theta_i_dev, const float* theta_dev,
                    const float* model_dev, float beta, int64_t n, int32_t merged_method,
                    float* model_out_dev, mamdr_stream stream);
int mamdr_dr_accumulate(mamdr_ctx* ctx, float* accum_dev, const float* model_dev,
                        const float* theta_dev, const float* theta_i_dev, int64_t n,
                        int32_t merged_method, mamdr_stream stream);
int mamdr_dr_apply_accum(mamdr_ctx* ctx, float* theta_i_dev, float* accum_dev, float sample_num,
                         float beta, int64_t n, mamdr_stream stream);
int mamdr_sub(mamdr_ctx* ctx, float* out_dev, const float* a_dev, const float* b_dev, int64_t n,
              mamdr_stream stream);
int mamdr_axpy_diff(mamdr_ctx* ctx, float* out_dev, const float* a_dev, const float* b_dev, float alpha,
                    int64_t n, mamdr_stream stream);

/* PCGrad's gradient projection (model_zoo/pcgrad.py:152-160, host numpy in the reference) for ONE variable viewed as
 * [rows, cols] (a 1-D variable is one row): per row, dot = <final, aux>; where dot > 0 (the reference's test)
 * aux' = aux - dot / ||final_row||_2 * final_row; final_row += aux'.  `final_grads` aliases `current_grads` in the
 * reference (pcgrad.py:104), hence one in/out buffer.  Deterministic (one warp per row, fixed-order reductions). */
int mamdr_pcgrad_project(mamdr_ctx* ctx, float* final_grads_dev, const float* aux_grads_dev, int64_t rows,
                         int32_t cols, mamdr_stream stream);

/* ---- K8: streaming AUC (replaces utils/auc.py AUC.update_state / result / reset_states) ------ */
int mamdr_auc_update(mamdr_ctx* ctx, const float* probs_dev, const float* labels_dev, int64_t n,
                     float* acc_dev, const float* thresholds_dev, int32_t num_thresholds,
                     mamdr_stream stream);
int mamdr_auc_result(mamdr_ctx* ctx, const float* acc_dev, int32_t num_thresholds, float* auc_dev,
                     mamdr_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* MAMDR_B200_H */
