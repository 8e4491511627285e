"""ctypes binding of librltime_b200.so (the C ABI declared in include/rltime_b200.h).

The product path has no CPU fallback: if the CUDA library is missing or fails to load,
importing a device-backed component raises immediately.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librltime_b200.so")

RT_OK = 0
RT_NEED_MORE_DATA = 1
RT_MAX_FIELDS = 16
RT_BATCH_SLOTS = 3
RT_KIND_UNIFORM = 0
RT_KIND_PRIORITIZED = 1


class RtError(RuntimeError):
    pass


class ReplayConfig(C.Structure):
    _fields_ = [
        ("size", C.c_int64), ("kind", C.c_int32), ("nstep_train", C.c_int32),
        ("prefix_steps", C.c_int32), ("nstep_target", C.c_int32), ("overlap", C.c_int32),
        ("global_importance_scaling", C.c_int32), ("gamma", C.c_double),
        ("alpha", C.c_double), ("eps", C.c_double), ("max_weight_factor", C.c_double),
        ("max_envs", C.c_int32), ("device", C.c_int32),
        ("num_state_fields", C.c_int32), ("state_field_bytes", C.c_int64 * RT_MAX_FIELDS),
        ("num_po_fields", C.c_int32), ("po_field_bytes", C.c_int64 * RT_MAX_FIELDS),
        ("avoid_episode_crossing", C.c_int32),
    ]


class Batch(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("S", C.c_int32), ("n", C.c_int32),
        ("num_state_fields", C.c_int32), ("num_po_fields", C.c_int32),
        ("all_states", C.c_void_p * RT_MAX_FIELDS),
        ("policy_outputs", C.c_void_p * RT_MAX_FIELDS),
        ("returns", C.c_void_p), ("nsteps", C.c_void_p), ("target_masks", C.c_void_p),
        ("importance_weights", C.c_void_p), ("loss_indices", C.c_void_p),
        ("idxes", C.c_void_p), ("slots", C.c_void_p),
        ("target_states", C.c_void_p * RT_MAX_FIELDS), ("weight_max", C.c_void_p),
    ]


RT_MAX_CONV = 8
RT_MAX_PRE_FC = 8
RT_BUF_ONLINE, RT_BUF_TARGET, RT_BUF_GRAD, RT_BUF_ADAM_M, RT_BUF_ADAM_V = range(5)
RT_GEMM_FP32_SIMT, RT_GEMM_TF32_TCGEN05, RT_GEMM_TF32_RN = 0, 1, 2


class ModelDesc(C.Structure):
    _fields_ = [
        ("in_c", C.c_int32), ("in_h", C.c_int32), ("in_w", C.c_int32), ("num_conv", C.c_int32),
        ("conv_filters", C.c_int32 * RT_MAX_CONV), ("conv_kernel", C.c_int32 * RT_MAX_CONV),
        ("conv_stride", C.c_int32 * RT_MAX_CONV), ("lstm_units", C.c_int32),
        ("fc_size", C.c_int32), ("num_actions", C.c_int32), ("num_quantiles", C.c_int32),
        ("embedding_dim", C.c_int32), ("dueling", C.c_int32),
        ("extra_dim", C.c_int32), ("num_pre_fc", C.c_int32),
        ("pre_fc_size", C.c_int32 * RT_MAX_PRE_FC), ("pre_fc_module", C.c_int32 * RT_MAX_PRE_FC),
        ("pre_fc_sub", C.c_int32 * RT_MAX_PRE_FC),
    ]


class TrainDesc(C.Structure):
    _fields_ = [
        ("mbatch", C.c_int32), ("nstep_train", C.c_int32), ("burn_in", C.c_int32),
        ("nstep_target", C.c_int32), ("double_q", C.c_int32), ("rnn_bootstrap", C.c_int32),
        ("loss_sum", C.c_int32), ("gemm_mode", C.c_int32), ("gamma", C.c_double),
        ("vf_scale_epsilon", C.c_double), ("huber_kappa", C.c_double), ("clip_grad", C.c_double),
        ("adam_epsilon", C.c_double), ("lr", C.c_double), ("seed", C.c_uint64),
        ("loss_timestep_agg", C.c_int32), ("loss_mse", C.c_int32),
        ("clip_grad_dynamic_alpha", C.c_double), ("rnn_steps_train", C.c_int32),
    ]


class LearnerIO(C.Structure):
    _fields_ = [("field_x", C.c_int32), ("field_hx", C.c_int32), ("field_cx", C.c_int32),
                ("field_initials", C.c_int32), ("po_field_actions", C.c_int32),
                ("field_extra", C.c_int32)]


# name -> (restype, argtypes); also the list the CPU test checks against the header
_VP = C.c_void_p
SIGNATURES = {
    "rt_last_error": (C.c_char_p, []),
    "rt_version": (C.c_int, []),
    "rt_launch_count": (C.c_int64, []),
    "rt_replay_create": (C.c_int, [C.POINTER(ReplayConfig), C.POINTER(_VP)]),
    "rt_replay_destroy": (None, [_VP]),
    "rt_replay_append": (C.c_int, [_VP, C.c_int64, _VP, _VP, _VP, _VP, _VP, _VP, C.c_int32, _VP]),
    "rt_replay_set_train_frequency": (C.c_int, [_VP, C.c_double]),
    "rt_replay_needed_feed": (C.c_int64, [_VP, C.c_int32, C.c_int32]),
    "rt_replay_consume_quota": (C.c_int, [_VP, C.c_int32]),
    "rt_replay_train_quota": (C.c_double, [_VP]),
    "rt_replay_len": (C.c_int64, [_VP]),
    "rt_replay_active_sequences": (C.c_int64, [_VP]),
    "rt_replay_uniform_available": (C.c_int64, [_VP]),
    "rt_replay_sample_prioritized": (C.c_int, [_VP, C.c_int32, C.c_double, _VP, _VP]),
    "rt_replay_sample_uniform": (C.c_int, [_VP, C.c_int32, _VP, _VP]),
    "rt_replay_batch": (C.c_int, [_VP, C.POINTER(Batch)]),
    "rt_replay_update_losses": (C.c_int, [_VP, C.c_int64, _VP, _VP, _VP]),
    "rt_replay_update_losses_last": (C.c_int, [_VP, _VP, _VP]),
    "rt_replay_profile": (C.c_int, [_VP, C.c_int32]),
    "rt_replay_gather_time": (C.c_int, [_VP, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "rt_replay_tree_sum": (C.c_int, [_VP, C.POINTER(C.c_double), _VP]),
    "rt_replay_tree_min": (C.c_int, [_VP, C.POINTER(C.c_double), _VP]),
    "rt_replay_tree_leaf": (C.c_int, [_VP, C.c_int32, C.POINTER(C.c_double), _VP]),
    "rt_tree_create": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(_VP)]),
    "rt_tree_destroy": (None, [_VP]),
    "rt_tree_set": (C.c_int, [_VP, C.c_int32, _VP, _VP, _VP]),
    "rt_tree_sum": (C.c_int, [_VP, C.POINTER(C.c_double), _VP]),
    "rt_tree_find": (C.c_int, [_VP, C.c_int32, _VP, _VP, _VP]),
    "rt_learner_create": (C.c_int, [C.POINTER(ModelDesc), C.POINTER(TrainDesc), C.c_int32,
                                    C.POINTER(_VP)]),
    "rt_learner_destroy": (None, [_VP]),
    "rt_learner_num_params": (C.c_int32, [_VP]),
    "rt_learner_num_weights": (C.c_int64, [_VP]),
    "rt_learner_param_info": (C.c_int, [_VP, C.c_int32, C.c_char_p, C.c_int32,
                                        C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "rt_learner_load_params": (C.c_int, [_VP, C.c_int32, _VP]),
    "rt_learner_get_params": (C.c_int, [_VP, C.c_int32, _VP]),
    "rt_learner_sync_target": (C.c_int, [_VP, _VP]),
    "rt_learner_set_lr": (C.c_int, [_VP, C.c_double]),
    "rt_learner_params_changed": (C.c_int, [_VP, _VP]),
    "rt_learner_step": (C.c_int, [_VP, C.POINTER(Batch), C.POINTER(LearnerIO), _VP, _VP]),
    "rt_learner_compute_grads": (C.c_int, [_VP, C.POINTER(Batch), C.POINTER(LearnerIO), _VP, _VP]),
    "rt_learner_apply_grads": (C.c_int, [_VP, C.c_double, _VP]),
    "rt_learner_prefetch": (C.c_int, [_VP, C.POINTER(Batch), C.POINTER(LearnerIO), _VP]),
    "rt_comm_unique_id": (C.c_int, [_VP]),
    "rt_comm_init": (C.c_int, [_VP, _VP, C.c_int32, C.c_int32]),
    "rt_comm_destroy": (C.c_int, [_VP]),
    "rt_comm_broadcast_params": (C.c_int, [_VP, C.c_int32, _VP]),
    "rt_comm_allreduce_max_f64": (C.c_int, [_VP, _VP, C.c_int32, _VP]),
    "rt_learner_step_dp": (C.c_int, [_VP, C.POINTER(Batch), C.POINTER(LearnerIO), _VP, _VP]),
    "rt_learner_flat_buffer": (C.c_int, [_VP, C.c_int32, C.POINTER(_VP), C.POINTER(C.c_int64)]),
    "rt_learner_act": (C.c_int, [_VP, C.c_int32] + [_VP] * 9 + [_VP]),
    "rt_learner_td_abs": (C.c_int, [_VP, C.POINTER(_VP)]),
    "rt_learner_wait_loss": (C.c_int, [_VP, _VP]),
    "rt_learner_wait_late_grads": (C.c_int, [_VP, _VP, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "rt_learner_read_loss": (C.c_int, [_VP, C.POINTER(C.c_float), C.POINTER(C.c_float), _VP]),
    "rt_learner_read_stats": (C.c_int, [_VP, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                        C.POINTER(C.c_float), _VP]),
    "rt_learner_debug_tensor": (C.c_int, [_VP, C.c_char_p, C.POINTER(_VP), C.POINTER(C.c_int64)]),
    "rt_gemm_test": (C.c_int, [C.c_int32] * 6 + [_VP, _VP, _VP, C.c_int32, _VP, C.c_int32]),
    "rt_learner_profile": (C.c_int, [_VP, C.c_int32]),
    "rt_learner_gemm_time": (C.c_int, [_VP, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                       C.POINTER(C.c_int64)]),
    "rt_learner_get_opt_state": (C.c_int, [_VP, C.POINTER(C.c_int64), C.POINTER(C.c_double)]),
    "rt_learner_set_opt_state": (C.c_int, [_VP, C.c_int64, C.c_double]),
    "rt_learner_get_aux_state": (C.c_int, [_VP, C.POINTER(C.c_uint64), C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "rt_learner_set_aux_state": (C.c_int, [_VP, C.c_uint64, C.c_float, C.c_int32]),
    "rt_learner_gemm_launches": (C.c_int, [_VP, C.c_int64, _VP, _VP, _VP]),
    "rt_learner_gemm_shapes": (C.c_int, [_VP, C.c_int64, _VP, C.POINTER(C.c_int64)]),
    "rt_gemm_bench": (C.c_int, [C.c_int32] * 9 + [C.POINTER(C.c_double), C.c_int32]),
}

_lib = None


def load():
    """Loads the shared library (once).  Fails loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RtError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    """Raises on a negative status; returns the (non-negative) status otherwise."""
    if rc < 0:
        raise RtError("rltime_b200: %s (status %d)" % (load().rt_last_error().decode(), rc))
    return rc


class DevPtr:
    """Borrowed device memory exposed through __cuda_array_interface__ (zero-copy into torch)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {
            "shape": tuple(int(s) for s in shape), "typestr": typestr,
            "data": (int(ptr), False), "version": 2, "strides": None,
        }


def as_tensor(ptr, shape, typestr, device):
    import torch
    if any(int(s) == 0 for s in shape):
        import numpy as np
        return torch.empty(tuple(shape), dtype=torch.from_numpy(np.empty(0, np.dtype(typestr))).dtype,
                           device=device)
    return torch.as_tensor(DevPtr(ptr, shape, typestr), device=device)
