/*
 * nvsm_b200.h — C ABI of the B200-native NVSM/LSE training step.
 *
 * Drop-in boundary for the per-batch hot path of cvangysel/cuNVSM. The reference
 * has no FFI of its own: its boundary is the C++ class surface of
 * include/cuNVSM/{model,objective,params,storage,updates}.h. Every entry point
 * below names the reference interface it replaces (file:line in the reference
 * tree); the C++ façade in include/cuNVSM/ (this repo) and the Python mirror in
 * cunvsm_b200/ are thin layers over exactly these symbols.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types.
 *   - index type is `long` (the reference's `int32` typedef is `long`,
 *     include/cuNVSM/base.h:28), floats are float32 (release build,
 *     cpp/CMakeLists.txt:17).
 *   - tensors cross the boundary row-major [objects, dim] — the memory image of the
 *     reference's column-major dim x objects device_matrix (cpp/storage.cu:6-10).
 *   - every function returns 0 on success, non-zero on error; nvsm_last_error()
 *     holds the message. (The reference aborts via glog CHECK / LOG(FATAL); the C++
 *     façade restores that behaviour on a non-zero return.)
 *   - there is no CPU fallback: every compute entry point requires a CUDA device.
 */
#ifndef NVSM_B200_H
#define NVSM_B200_H

#if defined(__GNUC__)
#define NVSM_API __attribute__((visibility("default")))
#else
#define NVSM_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nvsm_model nvsm_model;

/* proto/nvsm.proto:11-14 */
enum { NVSM_TANH = 0, NVSM_HARD_TANH = 1 };
/* proto/nvsm.proto:41-45 */
enum { NVSM_SGD = 0, NVSM_ADAGRAD = 1, NVSM_ADAM = 2 };
/* proto/nvsm.proto:51-56; CLI names sparse_adam / dense_adam / full_adam (cpp/main.cu:479-485) */
enum { NVSM_ADAM_SPARSE = 1, NVSM_ADAM_DENSE_UPDATE = 2, NVSM_ADAM_DENSE_UPDATE_DENSE_VARIANCE = 3 };
/* projection GEMM arithmetic */
enum { NVSM_GEMM_FP32 = 0, NVSM_GEMM_TF32 = 1, NVSM_GEMM_3XTF32 = 2 };
/* Objectives of the reference (include/cuNVSM/objective.h): TextEntity (the NVSM/LSE step), the pairwise
 * RepresentationSimilarity objective over the entity or the word table, and the two mixtures. */
enum { NVSM_OBJECTIVE_TEXT_ENTITY = 0, NVSM_OBJECTIVE_ENTITY_ENTITY = 1, NVSM_OBJECTIVE_TERM_TERM = 2,
       NVSM_OBJECTIVE_TEXT_ENTITY_ENTITY_ENTITY = 3, NVSM_OBJECTIVE_TEXT_ENTITY_TERM_TERM = 4 };
/* Multi-GPU treatment of the sparse tables (nvsm_comm_set_sparse_mode). */
enum { NVSM_SPARSE_LOCAL = 0, NVSM_SPARSE_ALLGATHER = 1 };

/* lse::ModelDesc + lse::TrainConfig (proto/nvsm.proto:7-71) flattened; replaces the
 * arguments of Model::Model (include/cuNVSM/model.h:82-85). */
typedef struct nvsm_config {
    long num_words;               /* |V| */
    long num_entities;            /* |D| */
    int word_repr_size;           /* d_w */
    int entity_repr_size;         /* d_d */
    int nonlinearity;             /* NVSM_TANH | NVSM_HARD_TANH */
    int batch_normalization;      /* ModelDesc.TransformDesc.batch_normalization */
    int clip_sigmoid;             /* forced on by the CLI, cpp/main.cu:645 */
    int bias_negative_samples;
    int l2_normalize_phrase_reprs; /* Normalizer on the phrase representations, cpp/objective.cu:96-99,134-140,461-468 */
    int l2_normalize_entity_reprs; /* Normalizer on the gathered entity rows, cpp/objective.cu:101-104,170-176,403-409 */
    int update_method;            /* NVSM_SGD | NVSM_ADAGRAD | NVSM_ADAM */
    int adam_mode;                /* NVSM_ADAM_* when update_method == NVSM_ADAM */
    int num_random_entities;      /* z */
    int max_batch_size;           /* largest num_instances a step will see */
    int window_size;              /* n */
    float regularization_lambda;
    int device;                   /* CUDA device ordinal */
    int gemm_mode;                /* NVSM_GEMM_* */
    int num_batch_slots;          /* device-resident batch slots for nvsm_stage_batch (>= 1) */
    int objective;                /* NVSM_OBJECTIVE_*: which Model<...::Objective> this handle is (cpp/model.cu:222-228) */
    float text_entity_weight;     /* TrainConfig.text_entity_weight   (mixtures; cpp/main.cu:704-706) */
    float similarity_weight;      /* TrainConfig.entity_entity_weight or .term_term_weight */
    int max_similarity_batch_size; /* largest RepresentationSimilarity::Batch (pairs) a step will see */
    int reserved[3];
} nvsm_config;

NVSM_API const char* nvsm_last_error(void);
NVSM_API int nvsm_version(void);

/* Page-locked host memory for batches — TextEntity::Batch allocates its four arrays with
 * cudaHostAlloc (cpp/data.cu:8-30); callers without the CUDA runtime use these. */
NVSM_API int nvsm_host_alloc(void** ptr, unsigned long bytes);
NVSM_API int nvsm_host_free(void* ptr);

/* Model::Model / ~Model — include/cuNVSM/model.h:82-85, cpp/model.cu:6-35,95-103. */
NVSM_API int nvsm_create(const nvsm_config* config, nvsm_model** out);
NVSM_API void nvsm_destroy(nvsm_model* m);

/* Run all work of this model on an existing CUDA stream (a cudaStream_t). The
 * reference issues everything on the per-thread default stream (cpp/model.cu:13-14). */
NVSM_API int nvsm_set_stream(nvsm_model* m, void* cuda_stream);
NVSM_API int nvsm_synchronize(nvsm_model* m);

/* ModelBase::initialize — cpp/model.cu:37-43 + init_matrix_glorot,
 * include/cuNVSM/cuda_utils.h:35-56. *rng_state is the std::minstd_rand0 state
 * (the reference's RNG*, include/cuNVSM/base.h:36), advanced in place. */
NVSM_API int nvsm_initialize(nvsm_model* m, unsigned long* rng_state);

/* ModelBase::get_data — cpp/model.cu:64-93 (names from cpp/params.cu:29-33 +
 * cpp/storage.cu:115-121,242-250): "word_representations-representations" [V,d_w],
 * "entity_representations-representations" [D,d_d], "word_entity_mapping-transform"
 * [d_w,d_d], "word_entity_mapping-bias" [d_d]. Also readable (tests / gradient
 * checks): per-step tensors "phrase_reprs" [B,d_w], "word_projections" [B,d_d],
 * "similarity_probs" [B*R], "instance_multipliers" [B*R], "grad_transform" [d_w,d_d],
 * "grad_bias" [d_d], "grad_phrase_reprs" [B,d_w], "grad_entity_repr" [B*R,d_d]
 * and optimiser state "<param>-{m,v,acc}". */
NVSM_API long nvsm_tensor_size(nvsm_model* m, const char* name);
NVSM_API int nvsm_get_tensor(nvsm_model* m, const char* name, float* host_out, long n);
NVSM_API int nvsm_set_tensor(nvsm_model* m, const char* name, const float* host_in, long n);

/* Objective::generate_labels / UniformLabelGenerator::generate — cpp/objective.cu:5-28,
 * cpp/labels.cu:3-22: out[i*(z+1)] = labels[i]; out[i*(z+1)+1..z] ~ U{0..num_objects-1}
 * drawn serially from minstd_rand0 exactly like the reference (bit-exact ids). Host only. */
NVSM_API int nvsm_generate_labels(const long* labels, long num_labels, long num_negative_labels,
                         long num_objects, unsigned long* rng_state, long* out);

/* Model::compute_cost — include/cuNVSM/model.h:99-100, cpp/objective.cu:30-313, with the
 * sampled entity ids passed in (so the sampler stays swappable). All four arrays are HOST
 * buffers laid out as TextEntity::Batch (include/cuNVSM/data.h:114-177): features
 * [B*n], feature_weights [B*n], entity_ids [B*(z+1)] positive first, weights [B]. Copies
 * them to the device and enqueues the forward pass; does not synchronise.
 * feature_weights and / or weights may be NULL (here and in nvsm_train_step / nvsm_step_sampled / nvsm_stage_batch):
 * uniform weighting, i.e. all 1.0 -- what the reference's data sources write unless self-information / idf weighting is
 * selected (include/cuNVSM/data.h:465-467, cpp/data_indri.cpp). Nothing is transferred for a NULL array: the device copy
 * is filled with ones (once per batch slot). */
NVSM_API int nvsm_compute_cost(nvsm_model* m, const long* features, const float* feature_weights,
                      const long* entity_ids, const float* weights, long num_instances);

/* nvsm_compute_cost / nvsm_train_step / nvsm_step_sampled return once the H2D copies of their host arrays are
 * ENQUEUED. A caller that recycles those arrays (AsyncSource swaps pinned batches, cpp/data_async.cpp) waits here
 * first; the reference gets the same guarantee from the blocking get_cost() of every batch (cpp/main.cu:444). */
NVSM_API int nvsm_wait_upload(nvsm_model* m);

/* Model::compute_gradients — include/cuNVSM/model.h:109, cpp/objective.cu:315-481. */
NVSM_API int nvsm_compute_gradients(nvsm_model* m);

/* Model::update — include/cuNVSM/model.h:111-113, cpp/model.cu:187-220. */
NVSM_API int nvsm_update(nvsm_model* m, float learning_rate, float scaled_regularization_lambda);

/* ForwardResult::get_cost / scaled_regularization_lambda —
 * cpp/intermediate_results.cu:80-129. get_cost synchronises (device->host read). */
NVSM_API int nvsm_get_cost(nvsm_model* m, float* cost);
/* Same, for the forward pass `steps_back` calls ago (0 = latest, < 15): waits only for that
 * step's loss read-back, so a training loop can read step k-1 while step k runs. */
NVSM_API int nvsm_read_cost(nvsm_model* m, int steps_back, float* cost);
/* The same value before the final rounding to float (the device accumulates the loss in double): what the central
 * differences of GradientCheckFn (cpp/gradient_check.cu:3-133, --check_gradients) are taken over. */
NVSM_API int nvsm_read_cost_f64(nvsm_model* m, int steps_back, double* cost);
NVSM_API float nvsm_scaled_regularization_lambda(nvsm_model* m);

/* One whole training step on host buffers = compute_cost + compute_gradients + update
 * (the body of iterate_data, cpp/main.cu:405-431), scaled lambda = lambda / B. Does not
 * synchronise; call nvsm_get_cost to read the loss of the step. */
NVSM_API int nvsm_train_step(nvsm_model* m, const long* features, const float* feature_weights,
                    const long* entity_ids, const float* weights, long num_instances,
                    float learning_rate);

/* Device-side negative sampler, bit-exact with nvsm_generate_labels (same std::minstd_rand0 stream,
 * same libstdc++ uniform_int_distribution rejections): the engine state lives on the device between
 * steps. nvsm_step_sampled = upload features / feature_weights / labels / weights (HOST buffers),
 * draw the z negatives per instance on the device, compute_cost and — when train != 0 —
 * compute_gradients + update (lambda / B). No synchronisation. */
NVSM_API int nvsm_sampler_seed(nvsm_model* m, unsigned long rng_state);
NVSM_API int nvsm_sampler_state(nvsm_model* m, unsigned long* rng_state);   /* synchronises */
NVSM_API int nvsm_step_sampled(nvsm_model* m, const long* features, const float* feature_weights,
                               const long* labels, const float* weights, long num_instances,
                               float learning_rate, int train);
NVSM_API int nvsm_get_entity_ids(nvsm_model* m, long* out, long n);         /* ids of the running step */
/* nvsm_generate_labels on the device: host labels in, host ids out, *rng_state advanced. */
NVSM_API int nvsm_generate_labels_device(nvsm_model* m, const long* labels, long num_labels,
                                         long num_negative_labels, long num_objects,
                                         unsigned long* rng_state, long* out);

/* Skewed negatives (BASELINE.json configs[4]: Zipf-skewed sampling). The reference ships only
 * UniformLabelGenerator and leaves LabelGenerator (include/cuNVSM/labels.h:7-18) as the plug point; this is an
 * inverse-CDF generator on the same shared minstd_rand0: every negative consumes ONE engine output x,
 * u = (x - 1) / 2147483646 in [0, 1), id = min{k : cdf[k] > u}. cdf[0..num_objects-1] is non-decreasing and ends at
 * exactly 1.0. nvsm_generate_labels_cdf is the host loop; nvsm_sampler_set_cdf installs the distribution for the
 * device sampler (nvsm_step_sampled draws bit-identical ids); a NULL cdf restores the uniform generator. */
NVSM_API int nvsm_generate_labels_cdf(const long* labels, long num_labels, long num_negative_labels,
                                      const double* cdf, long num_objects, unsigned long* rng_state, long* out);
NVSM_API int nvsm_sampler_set_cdf(nvsm_model* m, const double* cdf, long num_objects);

/* Device-resident batches: copy a host batch into slot `slot` once, then run steps on it
 * without host traffic. */
NVSM_API int nvsm_stage_batch(nvsm_model* m, int slot, const long* features, const float* feature_weights,
                     const long* entity_ids, const float* weights, long num_instances);
NVSM_API int nvsm_compute_cost_staged(nvsm_model* m, int slot);
NVSM_API int nvsm_train_step_staged(nvsm_model* m, int slot, float learning_rate);

/* Model::infer — include/cuNVSM/model.h:95-97, cpp/model.cu:105-133 (no batch-norm at
 * inference). words: host [num_phrases*window]; out: host [num_phrases, d_d]. */
NVSM_API int nvsm_infer(nvsm_model* m, const long* words, long num_phrases, long window, float* out);

/* Storage::increment_parameter — cpp/storage.cu:123-131,259-272 (gradient checking). */
NVSM_API int nvsm_increment_parameter(nvsm_model* m, const char* name, long idx, float epsilon);

/* Instrumentation. Phase timing uses CUDA events on the model's stream. */
/* enabled: 0 off; 1 serial phase timing (the stream overlaps of the fused step are switched off so that phases add up);
 * 2 timeline (overlaps kept: nvsm_get_timeline reports every phase interval, on whatever stream it ran, in ms since
 * this call). The role nvprof / NVTX ranges play for the reference (cpp/main.cu:16,372-459). */
NVSM_API int nvsm_set_profiling(nvsm_model* m, int enabled);
/* Returns the number of intervals recorded since nvsm_set_profiling(m, 2) (-1 on error, message in nvsm_last_error);
 * fills at most `capacity` entries. */
NVSM_API int nvsm_get_timeline(nvsm_model* m, int* phases, float* start_ms, float* end_ms, int capacity);
NVSM_API int nvsm_num_phases(void);
NVSM_API const char* nvsm_phase_name(int phase);
NVSM_API int nvsm_get_phase_ms(nvsm_model* m, float* ms_out, int capacity); /* sums since last reset */
NVSM_API int nvsm_reset_phase_ms(nvsm_model* m);
NVSM_API long nvsm_kernel_launches(nvsm_model* m); /* kernels launched by this model so far */

/* Test hook: the tcgen05/TMA GEMM in isolation on host matrices. variant 0: A[M,K], B[N,K]
 * (both K-major) -> C[M,N] = alpha A.B^T + bias; variant 1: A[K,M], B[K,N] (both MN-major,
 * split-K) -> C[M,N] = alpha A^T.B. */
NVSM_API int nvsm_test_gemm_tc(nvsm_model* m, int variant, int M, int N, int K, const float* A, const float* B,
                               float* C, float alpha, const float* bias, int splits);

/* Micro-benchmark hook: average milliseconds of `iters` back-to-back launches of the
 * tensor-core GEMM on device-resident (zero) operands. */
NVSM_API int nvsm_bench_gemm_tc(nvsm_model* m, int variant, int M, int N, int K, int splits, int with_stats,
                                int iters, float* ms_out);

/* Micro-benchmark hook for the roofline denominators bench.py reports (csrc/microbench.cuh), measured on the device
 * the model lives on. kind 0 / 3 / 4: GB/s of a plain warp-per-item gather of `rows_per_item` pseudo-random rows of
 * `row_floats` floats out of a `table_bytes` table (L2-resident when it fits), 2 / 4 / 1 rows in flight per warp -- the
 * access pattern of the step's gather-type kernels, which re-read every embedding row ~10x per batch (reference:
 * average_repr_kernel, cpp/params.cu:75-95, and update_repr_kernel, cpp/storage.cu:37-49); kind 1: GB/s (read + write)
 * of a streaming copy of `table_bytes`; kind 2: GB/s of an L2-resident streaming read (`items` sweeps over a
 * `table_bytes` buffer per launch): the SM <-> L2 ceiling. */
NVSM_API int nvsm_bench_memory(nvsm_model* m, int kind, long table_bytes, int row_floats, int rows_per_item, long items,
                               int iters, float* gbs_out);

/* RepresentationSimilarity::Objective::compute_cost + compute_gradients — cpp/objective.cu:487-672 — on a
 * RepresentationSimilarity::Batch (include/cuNVSM/data.h:560-614): pair_ids [2*num_pairs] HOST (adjacent ids form a
 * pair, rows of the entity table for *_ENTITY_ENTITY objectives, of the word table for *_TERM_TERM), weights
 * [num_pairs] HOST. Only valid on handles created with such an objective. For the mixtures call it next to
 * nvsm_compute_cost, then nvsm_compute_gradients and nvsm_update as usual: gradients are merged with the weights
 * w_k / sum_k w_k (MergeGradientsFn, cpp/intermediate_results.cu:3-60), ForwardResult::get_cost and
 * scaled_regularization_lambda of the mixture are the plain averages of the constituents' (:200-235).
 * Readable afterwards: "similarity_pair_probs" [num_pairs], "similarity_multipliers" [num_pairs],
 * "grad_similarity" [2*num_pairs, dim]. */
NVSM_API int nvsm_similarity_compute_cost(nvsm_model* m, const long* pair_ids, const float* weights, long num_pairs);
NVSM_API int nvsm_similarity_get_cost(nvsm_model* m, float* cost);
NVSM_API float nvsm_similarity_scaled_regularization_lambda(nvsm_model* m);

/* Multi-GPU (one process per GPU). The batch is sharded by n-gram row; the library
 * all-reduces batch-norm statistics and the dense gradients with NCCL (new: the reference
 * is single-GPU). id: 128 bytes from nvsm_comm_unique_id on rank 0, broadcast by the
 * caller (torch.distributed). After init, the batch size used for 1/B and lambda/B is the
 * sum over ranks. */
NVSM_API int nvsm_comm_unique_id(char* id_out_128);
NVSM_API int nvsm_comm_init(nvsm_model* m, const char* id_128, int num_ranks, int rank);

/* Optional NVLink peer exchange for the latency-bound per-step reductions (batch-norm sums, backward column sums +
 * loss): a one-block kernel pushes each rank's few KB into the peers' inboxes over NVLink and sums them in rank
 * order (cunvsm_b200/csrc/peer_allreduce.cuh), ~3 us instead of ~17 us per small ncclAllReduce. After
 * nvsm_comm_init every rank exports 128 bytes (two CUDA IPC handles), the caller all-gathers the blobs in rank order
 * and passes num_ranks * 128 bytes to nvsm_comm_peer_import. Without it the reductions use NCCL. grad_transform
 * (d_w * d_d floats) always goes through ncclAllReduce, on a side stream under grad_phrase and the word update. */
NVSM_API int nvsm_comm_peer_export(nvsm_model* m, char* handles_out_128);
NVSM_API int nvsm_comm_peer_import(nvsm_model* m, const char* all_handles);
NVSM_API int nvsm_comm_peer_disable(nvsm_model* m);   /* back to NCCL (every rank must take the same decision) */
NVSM_API int nvsm_comm_peer_status(nvsm_model* m, int* ready, int* error);

/* How the replicated word / entity tables are updated when num_ranks > 1 (new: the reference is
 * single-GPU). NVSM_SPARSE_LOCAL (default): each rank applies only the updates of its
 * own rows, no exchange, replicas drift apart. NVSM_SPARSE_ALLGATHER: the ranks all-gather their rows of
 * (entity ids, multipliers, activations, word ids, word weights, grad_phrase) with ncclAllGather and every
 * replica applies the updates of the whole global batch, i.e. N GPUs follow the trajectory of
 * Model::update (cpp/model.cu:187-220) on the global batch. Call after nvsm_comm_init. */
NVSM_API int nvsm_comm_set_sparse_mode(nvsm_model* m, int mode);

/* ------------------------------------------------------------------------------------------------------------------
 * Stand-alone operators: the reference's Params / Storage / Updates / BatchNormalization classes, one operation per
 * call, on caller-owned DEVICE tensors (row-major [objects, dim]; ids are `long`). The header-only classes in
 * include/cuNVSM/{device_matrix,storage,updates,params,cudnn_utils}.h (this repo) bind exactly these symbols; the fused
 * training step above does not pass through them.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct nvsm_ops nvsm_ops;         /* launch context of the operators: one device, one stream */
typedef struct nvsm_updater nvsm_updater; /* optimiser state of one parameter group (GradientUpdater::storages_) */

/* RepresentationsStorage::SingleGradientType (include/cuNVSM/storage.h:64-69): grad [count, dim] (may be overwritten:
 * sparse Adam turns it into the step, like the reference), ids [count * window], weights [count * window] or NULL. */
typedef struct nvsm_grad_desc {
    float* grad;
    const long* ids;
    long count;
    int window;
    const float* weights;
} nvsm_grad_desc;

NVSM_API int nvsm_ops_create(int device, nvsm_ops** out);   /* Streams / DefaultStream (include/cuNVSM/cuda_utils.h) */
NVSM_API void nvsm_ops_destroy(nvsm_ops* ops);
NVSM_API int nvsm_ops_synchronize(nvsm_ops* ops);
NVSM_API long nvsm_ops_kernel_launches(nvsm_ops* ops);
/* device_matrix storage (the un-vendored device_matrix library's constructor / fillwith / copy / to_host, as used by
 * cpp/storage.cu:6-10, cpp/updates_tests.cu:34-60): zero-initialised allocation, host <-> device copies, fill. */
NVSM_API int nvsm_dev_malloc(nvsm_ops* ops, void** ptr, unsigned long bytes);
NVSM_API int nvsm_dev_free(nvsm_ops* ops, void* ptr);
NVSM_API int nvsm_dev_upload(nvsm_ops* ops, void* dst_dev, const void* src_host, unsigned long bytes);
NVSM_API int nvsm_dev_download(nvsm_ops* ops, void* dst_host, const void* src_dev, unsigned long bytes);
NVSM_API int nvsm_dev_copy(nvsm_ops* ops, void* dst_dev, const void* src_dev, unsigned long bytes);
NVSM_API int nvsm_dev_fill(nvsm_ops* ops, float* ptr, long n, float value);

/* Representations::get_average_representations (cpp/params.cu:138-172; average_repr_kernel :75-95): out[o, :] =
 * (1 / window) sum_w weights[o, w] * table[ids[o, w], :]; weights may be NULL. window 1 without weights =
 * Representations::get_representations (cpp/params.cu:97-117). */
NVSM_API int nvsm_op_average_representations(nvsm_ops* ops, const float* table, long num_objects, int dim, const long* ids,
                                             const float* weights, long num_out, int window, float* out);
/* The projection inside Transform::transform (cpp/params.cu:396-421): out[B, d_d] = P[B, d_w] . T[d_w, d_d] (+ bias when
 * not NULL: the reference pre-broadcasts it into the destination and runs the GEMM with beta = 1). Exact fp32. */
NVSM_API int nvsm_op_project(nvsm_ops* ops, const float* T, int word_repr_size, int entity_repr_size, const float* P,
                             long num_instances, const float* bias, float* out);
/* tanh / hard_tanh of Transform::transform (cpp/params.cu:425-446; func::clip, include/cuNVSM/cuda_utils.h:86-111). */
NVSM_API int nvsm_op_activation(nvsm_ops* ops, const float* x, long n, int nonlinearity, float* y);
/* BatchNormalization::forward / backward (cpp/cudnn_utils.cu:82-129, 143-183; include/cuNVSM/cudnn_utils.h:84-127):
 * per-activation statistics over the rows, biased variance, epsilon inside the sqrt, gamma == 1, beta = bias; mean and
 * invstd are the caches the backward pass consumes. y may alias x, dx may alias dy. */
NVSM_API int nvsm_op_batchnorm_forward(nvsm_ops* ops, const float* x, const float* bias, long rows, int num_features,
                                       float epsilon, float* y, float* mean, float* invstd);
NVSM_API int nvsm_op_batchnorm_backward(nvsm_ops* ops, const float* dy, const float* x, const float* mean, const float* invstd,
                                        long rows, int num_features, float* dx, float* dbias);
/* update_dense (include/cuNVSM/storage_inl.h:4-32): param = param (1 - lambda lr) + op(grad) lr, op = identity | square. */
NVSM_API int nvsm_op_update_dense(nvsm_ops* ops, float* param, const float* grad, long n, float lr, float lambda, int square);
/* RepresentationsStorage::update (cpp/storage.cu:51-102; update_repr_kernel :37-49): dense decay when lambda > 0, then
 * table[ids[x, y], :] += lr * weights[x, y] * grad[x, :] for every descriptor. */
NVSM_API int nvsm_op_representations_update(nvsm_ops* ops, float* table, long num_objects, int dim, const nvsm_grad_desc* descs,
                                            int num_descs, float lr, float lambda);
/* TransformStorage::update (cpp/storage.cu:198-228): T with decay, the bias without. */
NVSM_API int nvsm_op_transform_update(nvsm_ops* ops, float* T, float* b, const float* gT, const float* gb, long num_transform,
                                      int num_bias, float lr, float lambda);

/* GradientUpdater construction (include/cuNVSM/updates.h:87-203): kind 0 = *RepresentationsGradientUpdater over a
 * [num_objects, dim] table, kind 1 = *TransformGradientUpdater over T [dim, target_dim] + b [target_dim]. */
NVSM_API int nvsm_updater_create(nvsm_ops* ops, int kind, int update_method, int adam_mode, long num_objects, int dim,
                                 int target_dim, float beta1, float beta2, float epsilon, nvsm_updater** out);
NVSM_API void nvsm_updater_destroy(nvsm_updater* updater);
/* {SGD,Adagrad,Adam}RepresentationsGradientUpdater::update (cpp/updates.cu:37-48, cpp/updates_adagrad.cu:99-179,
 * cpp/updates_adam.cu:153-385). Adagrad and sparse Adam take a single descriptor (the reference aborts otherwise) and,
 * like the reference, rewrite its gradient in place (g / sqrt(mean acc + eps); the window-averaged Adam step). */
NVSM_API int nvsm_updater_update_representations(nvsm_updater* updater, float* table, const nvsm_grad_desc* descs, int num_descs,
                                                 float lr, float lambda);
/* {SGD,Adagrad,Adam}TransformGradientUpdater::update (cpp/updates.cu:24-35, cpp/updates_adagrad.cu:33-70,
 * cpp/updates_adam.cu:46-105). Adagrad and Adam leave the applied step direction in gT / gb, like the reference. */
NVSM_API int nvsm_updater_update_transform(nvsm_updater* updater, float* T, float* b, float* gT, float* gb, float lr,
                                           float lambda);
/* GradientUpdater::storages_ as the reference's tests read them (cpp/updates_tests.cu:299-425): "acc", "m", "v" and, for a
 * transform, "acc_bias" / "m_bias" / "v_bias". Borrowed device pointer + element count. */
NVSM_API int nvsm_updater_state(nvsm_updater* updater, const char* name, float** dev_ptr, long* count);

#ifdef __cplusplus
}
#endif
#endif /* NVSM_B200_H */
