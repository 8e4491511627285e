/* rltime_b200 — C ABI of the B200-native replay + learner engine.
 *
 * This header is the drop-in boundary for the hot path of opherlieber/rltime
 * (SURVEY.md section 8b).  Every entry point cites the reference interface it replaces
 * (paths relative to the reference tree).  Plain pointers and sizes only; no torch or
 * C++ types cross the boundary.  All functions return 0 on success, a positive
 * RT_NEED_MORE_DATA where noted, or a negative rt_status on error; the message for the
 * last error on the calling thread is available from rt_last_error().
 *
 * Threading: a handle is single-owner (the reference history buffers are documented
 * single-threaded, rltime/history/data_structures/cyclic_array.py:8).  Device work is
 * stream-ordered on the cudaStream_t passed as `stream` (void*, 0 = legacy default).
 * Ownership: the library owns all device storage; batch pointers handed out by
 * rt_replay_batch() are borrowed and stay valid until RT_BATCH_SLOTS further draws.
 */
#ifndef RLTIME_B200_H
#define RLTIME_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_OK 0
#define RT_NEED_MORE_DATA 1   /* the reference returns None: "feed me more samples" */
#define RT_ERR_INVALID (-1)
#define RT_ERR_CUDA (-2)
#define RT_ERR_STATE (-3)
#define RT_ERR_NCCL (-4)

#define RT_BATCH_SLOTS 3      /* mirrors StateStore's 3-deep history, general/backend.py:88,149-152 */
#define RT_MAX_FIELDS 16

#define RT_KIND_UNIFORM 0     /* rltime/history/replay_history.py:6 */
#define RT_KIND_PRIORITIZED 1 /* rltime/history/prioritized_replay_history.py:10 */

const char* rt_last_error(void);
int rt_version(void);
/* Number of kernel launches issued by this library on the calling process so far. */
int64_t rt_launch_count(void);

/* ------------------------------------------------------------------ replay buffer */
typedef struct rt_replay rt_replay;

/* Constructor arguments of ReplayHistoryBuffer / PrioritizedReplayHistoryBuffer
 * (replay_history.py:14-15, prioritized_replay_history.py:41-43) plus the History base
 * arguments (history.py:17-18).  `gamma` replaces discount_function, which the trainer
 * always builds as (gamma ** nstep) * reward (training/multi_step_trainer.py:70-74). */
typedef struct rt_replay_config {
  int64_t size;            /* capacity in transitions */
  int32_t kind;            /* RT_KIND_* */
  int32_t nstep_train;     /* T */
  int32_t prefix_steps;    /* P (burn-in) */
  int32_t nstep_target;    /* n */
  int32_t overlap;         /* resolved overlap in [0, T); PER only */
  int32_t global_importance_scaling;
  double gamma;
  double alpha, eps, max_weight_factor;
  int32_t max_envs;        /* dense env indices are in [0, max_envs) */
  int32_t device;          /* CUDA ordinal */
  int32_t num_state_fields;                 /* leaves of sample["next_state"] */
  int64_t state_field_bytes[RT_MAX_FIELDS]; /* bytes per transition per leaf */
  int32_t num_po_fields;                    /* leaves of sample["policy_output"] */
  int64_t po_field_bytes[RT_MAX_FIELDS];
  int32_t avoid_episode_crossing;           /* uniform replay: _refine_sample_range (replay_history.py:142-171) */
} rt_replay_config;

int rt_replay_create(const rt_replay_config* cfg, rt_replay** out);
void rt_replay_destroy(rt_replay* h);

/* History.update (history.py:123-176) + _sample_added (replay_history.py:77-91,
 * prioritized_replay_history.py:136-172) for m transitions in arrival order.
 * env[i] dense env index; env_ids[i] the caller's integer env id (echoed in loss_indices).
 * state_fields[f] / po_fields[f]: m contiguous items of the f-th leaf; host pointers, or
 * device pointers when fields_on_device != 0. */
int rt_replay_append(rt_replay* h, int64_t m, const int32_t* env, const int64_t* env_ids,
                     const double* reward, const uint8_t* done,
                     const void* const* state_fields, const void* const* po_fields,
                     int32_t fields_on_device, void* stream);

/* Train quota of the replay buffers (replay_history.py:62-75,173-184), for hosts that do not keep it themselves:
 * every appended transition adds train_frequency to the quota, every get_train_data takes mbatch * nstep_train.
 *   rt_replay_needed_feed  -> -1: the reference's None (train first, do not act); 0: take whatever is ready
 *                              (no train_frequency); > 0: samples wanted (at least num_envs)
 *   rt_replay_consume_quota: call once per get_train_data, BEFORE the draw (the reference subtracts even when it then
 *                              returns None); RT_ERR_STATE when the +-100 x mbatch x nstep_train sanity bound breaks. */
int rt_replay_set_train_frequency(rt_replay* h, double train_frequency);
int64_t rt_replay_needed_feed(const rt_replay* h, int32_t mbatch, int32_t num_envs);
int rt_replay_consume_quota(rt_replay* h, int32_t mbatch);
double rt_replay_train_quota(const rt_replay* h);

/* Bookkeeping queries (len(linear_history); active sequences = len(_index_data) -
 * len(_free_indexes), prioritized_replay_history.py:291; uniform `total_available`,
 * replay_history.py:98-107). */
int64_t rt_replay_len(const rt_replay* h);
int64_t rt_replay_active_sequences(const rt_replay* h);
int64_t rt_replay_uniform_available(rt_replay* h);

/* PrioritizedReplayHistoryBuffer._get_train_data (prioritized_replay_history.py:281-356):
 * stratified sum-tree draw from `uniforms` (B doubles from the caller's MT19937 stream,
 * :238), sequence lookup, n-step assembly (history.py:71-108,178-201), gather into the
 * (S+n, B) time-major batch (history.py:203-286), IS weights (:327,:347-354).
 * `beta` is the already-annealed exponent (:287-288).
 * Returns RT_NEED_MORE_DATA when fewer than B sequences are active (:295-299). */
int rt_replay_sample_prioritized(rt_replay* h, int32_t B, double beta, const double* uniforms,
                                 void* stream);

/* ReplayHistoryBuffer._get_train_data (replay_history.py:93-140): `choices` are the B
 * values of np.random.choice(total_available, B) (:118), mapped to (env, start) by
 * walking the envs in first-appearance order (:120-134). */
int rt_replay_sample_uniform(rt_replay* h, int32_t B, const int64_t* choices, void* stream);

/* Device-resident result of the last draw, all time-major.  Row r = t * B + b. */
typedef struct rt_batch {
  int32_t B, S, n;                 /* S = prefix_steps + nstep_train */
  int32_t num_state_fields, num_po_fields;
  /* all_states[f]: (S+n)*B items; states = rows [0, S*B), target_states = rows
   * [n*B, (S+n)*B) — the reference's overlapped stack (history.py:245-265). */
  void* all_states[RT_MAX_FIELDS];
  void* policy_outputs[RT_MAX_FIELDS]; /* S*B items each */
  double* returns;                 /* S*B */
  int64_t* nsteps;                 /* S*B */
  double* target_masks;            /* S*B */
  double* importance_weights;      /* S*B (PER) */
  int64_t* loss_indices;           /* S*B*2 (PER): (env_id, env_offset) or (-1,-1) */
  int32_t* idxes;                  /* B (PER): drawn prioritization indices */
  int32_t* slots;                  /* (S+n)*B storage slots behind all_states (debug/fusion) */
  /* Optional: separately stacked target states, S*B items per leaf (history.py:268-270, the layout of
   * batches whose n-step varies per row, e.g. OnlineHistoryBuffer with fixed_target).  When
   * target_states[0] is non-null the learner reads the bootstrap states from here and all_states only has
   * to hold the S*B training rows. */
  void* target_states[RT_MAX_FIELDS];
  /* PER: [1] the weight the importance weights of this batch were divided by (the batch maximum, or the
   * min-tree value with global_importance_scaling; prioritized_replay_history.py:347-354).  A sharded replay
   * rescales importance_weights by weight_max / max-over-ranks(weight_max) so every shard normalises by
   * the same global maximum (rt_comm_allreduce_max_f64). */
  double* weight_max;
} rt_batch;

int rt_replay_batch(rt_replay* h, rt_batch* out);

/* PrioritizedReplayHistoryBuffer.update_losses (prioritized_replay_history.py:243-279):
 * pairs = m x (dense env index, env offset); losses widened to fp64 by the caller. */
int rt_replay_update_losses(rt_replay* h, int64_t m, const int64_t* pairs, const double* losses,
                            void* stream);
/* Same, for the B*T trained rows of the last draw with |td| produced on the device
 * (fp32, time-major T*B); performs the D2H read-back itself. */
int rt_replay_update_losses_last(rt_replay* h, const float* td_abs_device, void* stream);

/* Measurement hook: when enabled, CUDA events bracket every gather-kernel launch on its own
 * stream; rt_replay_gather_time returns (and resets) the summed device time. */
int rt_replay_profile(rt_replay* h, int32_t enable);
int rt_replay_gather_time(rt_replay* h, double* total_ms, int64_t* launches);

/* Unit-test hooks on the fp64 sum/min trees (data_structures/segment_tree.py). */
int rt_replay_tree_sum(rt_replay* h, double* out, void* stream);
int rt_replay_tree_min(rt_replay* h, double* out, void* stream);
int rt_replay_tree_leaf(rt_replay* h, int32_t idx, double* out, void* stream);
/* Standalone tree for kernel tests/benchmarks: capacity must be a power of two. */
typedef struct rt_tree rt_tree;
int rt_tree_create(int32_t capacity, int32_t device, rt_tree** out);
void rt_tree_destroy(rt_tree* t);
int rt_tree_set(rt_tree* t, int32_t m, const int32_t* idx, const double* val, void* stream);
int rt_tree_sum(rt_tree* t, double* out, void* stream);
int rt_tree_find(rt_tree* t, int32_t m, const double* mass, int32_t* out_idx, void* stream);


/* ------------------------------------------------------------------------ learner */
typedef struct rt_learner rt_learner;
#define RT_MAX_CONV 8

/* Topology family of the hot path: CNN -> [LSTM] -> FC -> IQN quantile layer (injected
 * before the last layer, injection_layer=-1) -> out (+ dueling value branch).  Mirrors the
 * json model description consumed by SequentialModel (rltime/models/torch/sequential.py:15-66,
 * configs/models/nature_cnn_lstm512_fc512.json) and the IQNPolicy / DQNPolicy constructor
 * arguments (rltime/policies/torch/iqn.py:12-13, dqn.py:15-16). */
#define RT_MAX_PRE_FC 8
typedef struct rt_model_desc {
  int32_t in_c, in_h, in_w;            /* observation (C, H, W), uint8, channel first; num_conv == 0: a float32
                                        * vector of in_c * in_h * in_w values (configs/models/mlp_2x64.json) */
  int32_t num_conv;                    /* 0 = no CNN module */
  int32_t conv_filters[RT_MAX_CONV], conv_kernel[RT_MAX_CONV], conv_stride[RT_MAX_CONV];
  int32_t lstm_units;                  /* 0 = no recurrent layer */
  int32_t fc_size;
  int32_t num_actions;
  int32_t num_quantiles;               /* num_sampling_quantiles; 0 = plain DQNPolicy (no quantile layer,
                                        * rltime/policies/torch/dqn.py) trained by DQN._compute_grads */
  int32_t embedding_dim;
  int32_t dueling;
  /* Tuple observation (models/torch/torch_model.py:33-56): width of the extra 1-D feature vector that
   * SequentialModel._combine_extra_inputs concatenates to the input of the LSTM layer
   * (models/torch/sequential.py:146-165,193-195; env_wrappers/common.py:221-233).  0 = plain Box. */
  int32_t extra_dim;
  /* Linear + ReLU layers (models/torch/modules/fc.py:7-36, every layer of every FC module that is not the
   * last module) between the CNN / the raw observation and the LSTM / the last FC module, in forward order:
   * output width, index of the module in SequentialModel.layers, index of the layer inside its module. */
  int32_t num_pre_fc;
  int32_t pre_fc_size[RT_MAX_PRE_FC];
  int32_t pre_fc_module[RT_MAX_PRE_FC];
  int32_t pre_fc_sub[RT_MAX_PRE_FC];
} rt_model_desc;

/* Training arguments: union of IQN/DQN._train, TorchTrainer._train and MultiStepTrainer._train
 * (rltime/training/torch/dqn.py:15-17, torch_trainer.py:9-10, multi_step_trainer.py:152-156). */
typedef struct rt_train_desc {
  int32_t mbatch;         /* B */
  int32_t nstep_train;    /* T */
  int32_t burn_in;        /* burn_in_timesteps = prefix_steps P */
  int32_t nstep_target;   /* n */
  int32_t double_q;
  int32_t rnn_bootstrap;
  int32_t loss_sum;       /* loss_aggregation: 0 = mean, 1 = sum */
  int32_t gemm_mode;      /* RT_GEMM_FP32_SIMT or RT_GEMM_TF32_TCGEN05 */
  double gamma;
  double vf_scale_epsilon; /* <= 0: no value rescaling */
  double huber_kappa;
  double clip_grad;        /* <= 0: no clipping */
  double adam_epsilon;
  double lr;               /* NB the reference's train_init ignores lr (torch_trainer.py:80-83) */
  uint64_t seed;           /* device RNG for the quantile fractions when none are injected */
  int32_t loss_timestep_agg; /* loss_timestep_aggregation (dqn.py:116-124): 0 none, 1 mean, 2 sum */
  int32_t loss_mse;          /* DQN loss_mode (dqn.py:96-110): 0 huber, 1 mse */
  double clip_grad_dynamic_alpha; /* >= 0: clip to clip_grad x EMA(grad norm) (torch_trainer.py:153-175) */
  /* rnn_steps_train (multi_step_trainer.py:192-216,239,305): the LSTM views the T*B time-major rows of a
   * pass as (rnn_steps_train, T*B / rnn_steps_train) -- exactly the reference's x.view(timesteps, -1, ...)
   * (models/torch/modules/lstm.py:60-79) -- taking the stored state of its first T*B/rnn_steps_train rows.
   * 0 = nstep_train.  Must divide nstep_train. */
  int32_t rnn_steps_train;
} rt_train_desc;

/* Which leaves of the replay batch feed the learner. */
typedef struct rt_learner_io {
  int32_t field_x, field_hx, field_cx, field_initials; /* indices into rt_batch.all_states */
  int32_t po_field_actions;                            /* index into rt_batch.policy_outputs (int64) */
  int32_t field_extra;                                 /* extra feature vector (float32), read when extra_dim > 0 */
} rt_learner_io;

#define RT_GEMM_FP32_SIMT 0     /* fp32 CUDA-core GEMM (parity reference path) */
#define RT_GEMM_TF32_TCGEN05 1  /* tcgen05.mma kind::tf32, fp32 accumulate in TMEM; fp32 operands as they are in
                                 * memory: the tensor core drops their low 13 mantissa bits (truncation) */
#define RT_GEMM_TF32_RN 2       /* same kernels, but every forward operand is rounded to the NEAREST TF32 value by
                                 * its producer (activations in the producing kernel's epilogue, weights in a shadow
                                 * copy refreshed by the Adam pass): unbiased products, no extra pass over memory */

#define RT_BUF_ONLINE 0
#define RT_BUF_TARGET 1
#define RT_BUF_GRAD 2
#define RT_BUF_ADAM_M 3
#define RT_BUF_ADAM_V 4

/* PolicyTrainer.init_policies / IQN.create_policy + TorchTrainer.train_init
 * (rltime/training/policy_trainer.py:39-66, torch/iqn.py:11-13, torch_trainer.py:80-83). */
int rt_learner_create(const rt_model_desc* model, const rt_train_desc* train, int32_t device,
                      rt_learner** out);
void rt_learner_destroy(rt_learner* h);
/* state_dict surface (TorchPolicy.get_state / load_state, torch_policy.py:97-101): tensors are
 * enumerated in the reference's registration order with its parameter names and layouts. */
int32_t rt_learner_num_params(const rt_learner* h);
int64_t rt_learner_num_weights(const rt_learner* h);
int rt_learner_param_info(const rt_learner* h, int32_t i, char* name, int32_t name_cap,
                          int64_t* shape4, int32_t* ndim);
int rt_learner_load_params(rt_learner* h, int32_t which, const float* const* tensors);
int rt_learner_get_params(rt_learner* h, int32_t which, float* const* tensors);
/* PolicyTrainer.sync_target -> TorchPolicy.copy_from (policy_trainer.py:68-70). */
int rt_learner_sync_target(rt_learner* h, void* stream);
/* Call after writing the online / target flat buffers (rt_learner_flat_buffer) directly, e.g. after a
 * broadcast of the initial weights: refreshes the TF32 shadow copies of RT_GEMM_TF32_RN (no-op otherwise). */
int rt_learner_params_changed(rt_learner* h, void* stream);
/* TorchTrainer.set_lr (torch_trainer.py:149-151). */
int rt_learner_set_lr(rt_learner* h, double lr);
/* Optimizer step counter (Adam bias correction) and learning rate, for checkpoint / resume: the
 * reference checkpoint (policy_trainer.py:170-185) carries only the policy weights, so training
 * cannot resume there; with RT_BUF_ADAM_M / RT_BUF_ADAM_V through rt_learner_get/load_params these
 * complete the optimizer state. */
int rt_learner_get_opt_state(rt_learner* h, int64_t* adam_steps, double* lr);
int rt_learner_set_opt_state(rt_learner* h, int64_t adam_steps, double lr);
/* The rest of the state a bit-exact resume needs: the counter of the device RNG that draws the IQN quantile
 * fractions (policies/torch/iqn.py:88 draws them with torch.rand) and the moving average of the dynamic
 * gradient clip (torch_trainer.py:153-175). */
int rt_learner_get_aux_state(rt_learner* h, uint64_t* rng_counter, float* clip_ema, int32_t* clip_ema_init);
int rt_learner_set_aux_state(rt_learner* h, uint64_t rng_counter, float clip_ema, int32_t clip_ema_init);
/* One learner update on a replay batch: burn-in (multi_step_trainer.py:90-131), bootstrap
 * targets (torch/iqn.py:15-52, torch_trainer.py:101-147), training forward + quantile-Huber
 * loss + backward (torch/iqn.py:54-129), grad-norm clip + Adam (torch_trainer.py:177-199).
 * taus_host: NULL, or 3 host arrays of T*B*Nq fractions (target, action-selection, train
 * forward) replacing the reference's torch.rand draws (policies/torch/iqn.py:88). */
int rt_learner_step(rt_learner* h, const rt_batch* batch, const rt_learner_io* io,
                    const float* const* taus_host, void* stream);
/* Data-parallel split of rt_learner_step: everything up to and including the backward pass,
 * then (after the caller has all-reduced the flat gradient buffer, e.g. NCCL sum over NVLink)
 * grad-norm / clip / Adam with the gradient scaled by grad_scale (= 1/world_size). */
int rt_learner_compute_grads(rt_learner* h, const rt_batch* batch, const rt_learner_io* io,
                             const float* const* taus_host, void* stream);
int rt_learner_apply_grads(rt_learner* h, double grad_scale, void* stream);
/* Optional overlap hook: converts the uint8 frames of `batch` (cnn.py:44-45: float, x 1/255) on `stream` --
 * typically the replay buffer's stream, right behind the gather of the draw -- into a frame buffer private to
 * that batch slot, so the next rt_learner_step / compute_grads on this batch starts at the first convolution.
 * A no-op for model / training configurations whose passes do not share one frame conversion. */
int rt_learner_prefetch(rt_learner* h, const rt_batch* batch, const rt_learner_io* io, void* stream);
/* ---- data parallelism inside the library (SURVEY.md 8b export list: rt_comm_init).  The reference has no
 * multi-GPU path; these make the one exchange step of the sharded design (8e) host-language neutral: the
 * host only has to carry the 128-byte NCCL id from rank 0 to the other ranks (file, socket, MPI, a
 * torch.distributed broadcast ...).  NCCL is dlopen'ed at run time (libnccl.so.2). */
#define RT_COMM_ID_BYTES 128
int rt_comm_unique_id(uint8_t* id128);                    /* rank 0: ncclGetUniqueId */
int rt_comm_init(rt_learner* h, const uint8_t* id128, int32_t rank, int32_t world);
int rt_comm_destroy(rt_learner* h);
/* Every rank starts from rank `root`'s online / target weights. */
int rt_comm_broadcast_params(rt_learner* h, int32_t root, void* stream);
/* In-place max over the ranks of `count` device doubles (global importance-weight normalisation of the
 * sharded prioritized replay, prioritized_replay_history.py:347-354). */
int rt_comm_allreduce_max_f64(rt_learner* h, double* dev_values, int32_t count, void* stream);
/* rt_learner_step across the communicator: local gradients, NCCL sum (the non-convolution bucket overlaps
 * the convolution backward on an internal communication stream), identical clip + Adam with the gradient
 * mean on every rank. */
int rt_learner_step_dp(rt_learner* h, const rt_batch* batch, const rt_learner_io* io,
                       const float* const* taus_host, void* stream);
/* Flat fp32 buffers (RT_BUF_*): all tensors of one kind back to back, 256-byte aligned. */
int rt_learner_flat_buffer(rt_learner* h, int32_t which, float** dev_ptr, int64_t* count);
/* DQNPolicy.actor_predict / IQNPolicy._actor_predict_postprocess (policies/torch/dqn.py:132-148,
 * iqn.py:124-131) for E envs at timesteps = 1 with the ONLINE network: q-values averaged over
 * the sampled quantiles plus the new LSTM state (LSTM.last_state, models/torch/modules/lstm.py:
 * 118-120).  All pointers are device pointers except taus_host (E*Nq fractions or NULL); x is uint8
 * frames (float32 vectors when num_conv == 0), extra the E x extra_dim extra features or NULL.  Uses
 * the learner's activation buffers: call between updates, on the update stream. */
int rt_learner_act(rt_learner* h, int32_t E, const void* x, const float* extra, const float* hx,
                   const float* cx, const float* initials, const float* taus_host, float* qvalues,
                   float* h_out, float* c_out, void* stream);
/* Device pointer to the T*B reported |td| means (torch/iqn.py:112) of the last step. */
int rt_learner_td_abs(rt_learner* h, float** out_device);
/* The reported |td| / losses of a step are final BEFORE its backward pass (the reference reads
 * them after backward, training/torch/dqn.py:73-81, with the same values).  wait_loss makes
 * `stream` wait for that point of the last enqueued step, so the priority write-back
 * (rt_replay_update_losses_last) and the next draw can run on a second stream while the backward
 * pass and Adam still execute; read_loss does the same wait on `stream`, then reads qloss and
 * td_mean back (pinned) and synchronises `stream` only. */
int rt_learner_wait_loss(rt_learner* h, void* stream);
/* Data-parallel overlap: rt_learner_compute_grads finishes the gradients of every non-convolution
 * parameter (flat range [*first, *first + *count): LSTM, hidden, head and quantile layers, 99 % of
 * the bytes) before it starts the convolution backward.  This makes `stream` wait for that point
 * of the last enqueued compute_grads, so their all-reduce can run on `stream` while the
 * convolution backward still executes.  stream == (void*)-1: only report the range. */
int rt_learner_wait_late_grads(rt_learner* h, void* stream, int64_t* first, int64_t* count);
int rt_learner_read_loss(rt_learner* h, float* loss, float* td_mean, void* stream);
/* qloss, td_mean, grad_norm of the last step (torch/iqn.py:127-129, torch_trainer.py:187-190);
 * synchronises the stream. */
int rt_learner_read_stats(rt_learner* h, float* loss, float* td_mean, float* grad_norm, void* stream);
/* Test hook: C[M,N] = op(A) . op(B) on the selected GEMM path with host operands
 * (transA: A stored [K][M]; transB: B stored [N][K]); optional per-column bias and ReLU. */
int rt_gemm_test(int32_t mode, int32_t M, int32_t N, int32_t K, int32_t transA, int32_t transB,
                 const float* A, const float* B, const float* bias, int32_t relu, float* C,
                 int32_t device);
/* Measurement hook: when enabled, CUDA events bracket every GEMM-shaped launch (tcgen05 GEMM,
 * implicit conv, conv weight gradient, SIMT fallback) on the launch stream;
 * rt_learner_gemm_time returns and resets the summed device time and algorithmic flops (2MNK). */
int rt_learner_profile(rt_learner* h, int32_t enable);
int rt_learner_gemm_time(rt_learner* h, double* total_ms, double* total_flops, int64_t* launches);
/* Per-launch view of the same measurement (call before rt_learner_gemm_time, which resets it):
 * algorithmic flops and device milliseconds of up to `cap` timed launches, in launch order. */
int rt_learner_gemm_launches(rt_learner* h, int64_t cap, double* flops, double* ms, int64_t* count);
/* ... and their shapes, 6 ints per launch: kind (0 GEMM, 1 conv forward, 2 conv weight gradient,
 * 3 conv data gradient), M, N, K, transA, transB. */
int rt_learner_gemm_shapes(rt_learner* h, int64_t cap, int32_t* shapes6, int64_t* count);
/* Measurement hook: average device time (CUDA events) of `iters` back-to-back launches of one
 * GEMM shape; force_bn / force_stages (0 = heuristic) select the tcgen05 tile configuration. */
int rt_gemm_bench(int32_t mode, int32_t M, int32_t N, int32_t K, int32_t transA, int32_t transB,
                  int32_t force_bn, int32_t force_stages, int32_t iters, double* avg_us,
                  int32_t device);
/* Test hook: named intermediate activations / gradients of the last step. */
int rt_learner_debug_tensor(rt_learner* h, const char* name, void** dev_ptr, int64_t* count);

#ifdef __cplusplus
}
#endif
#endif /* RLTIME_B200_H */
