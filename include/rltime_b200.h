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
} rt_batch;

int rt_replay_batch(rt_replay* h, rt_batch* out);

/* PrioritizedReplayHistoryBuffer.update_losses (prioritized_replay_history.py:243-279):
 * pairs = m x (dense env index, env offset); losses widened to fp64 by the caller. */
int rt_replay_update_losses(rt_replay* h, int64_t m, const int64_t* pairs, const double* losses,
                            void* stream);
/* Same, for the B*T trained rows of the last draw with |td| produced on the device
 * (fp32, time-major T*B); performs the D2H read-back itself. */
int rt_replay_update_losses_last(rt_replay* h, const float* td_abs_device, void* stream);

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

#ifdef __cplusplus
}
#endif
#endif /* RLTIME_B200_H */
