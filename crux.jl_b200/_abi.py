"""ctypes binding of ``libcrux_cuda.so`` (include/crux_cuda.h).

This is the Python twin of the Julia ``ccall`` shim in INTEGRATION.md: plain pointers and
sizes only.  There is NO CPU fallback: if the library is missing, or a context cannot be
created because there is no CUDA device, the caller gets a loud error.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libcrux_cuda.so")

OK, ERR_INVALID, ERR_CUDA, ERR_NAN, ERR_OOM, ERR_NCCL, ERR_STATE = range(7)
ACT_IDENTITY, ACT_TANH, ACT_RELU = 0, 1, 2
U8, F32, I32, I64 = 0, 1, 2, 3
PPO_INFO_STRIDE = 8
STREAM_LEGACY = 1  # CRUX_STREAM_LEGACY == cudaStreamLegacy
PPO_LOSS, PPO_GRAD_NORM, PPO_ENTROPY, PPO_KL, PPO_CLIP_FRAC, PPO_AVG_ADV, PPO_AVG_RET, PPO_VALID = range(8)


class CruxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libcrux_cuda status {code}: {msg}")
        self.code = code


class NaNError(CruxError, FloatingPointError):
    """training.jl:20 ``error("NaN detected!")``."""


class PPOHp(C.Structure):
    _fields_ = [("eps_clip", C.c_float), ("lambda_p", C.c_float), ("lambda_e", C.c_float), ("target_kl", C.c_float),
                ("a2c", C.c_int32), ("actor_epochs", C.c_int32), ("actor_batch", C.c_int32),
                ("critic_epochs", C.c_int32), ("critic_batch", C.c_int32),
                ("actor_max_batches", C.c_int64), ("critic_max_batches", C.c_int64)]


class LagrangeHp(C.Structure):
    _fields_ = [("target_cost", C.c_float), ("penalty_max", C.c_float), ("Ki_max", C.c_float), ("Ki", C.c_float), ("Kp", C.c_float),
                ("Kd", C.c_float), ("ema_alpha", C.c_double), ("cost_epochs", C.c_int32), ("cost_batch", C.c_int64),
                ("cost_max_batches", C.c_int64)]


class RolloutCols(C.Structure):
    _fields_ = [("s", C.c_void_p), ("a", C.c_void_p), ("sp", C.c_void_p), ("r", C.c_void_p), ("done", C.c_void_p),
                ("episode_end", C.c_void_p), ("logprob", C.c_void_p)]


class ColDesc(C.Structure):
    _fields_ = [("id", C.c_int32), ("dtype", C.c_int32), ("rowlen", C.c_int64), ("init", C.c_double)]


_vp, _i32, _i64, _u64, _f32, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_double
_pp = C.POINTER(C.c_void_p)

# name -> argtypes (restype is int32 unless listed in _RESTYPE)
_SIGS = {
    "crux_abi_version": [],
    "crux_ctx_create": [_i32, _vp, _pp],
    "crux_ctx_destroy": [_vp],
    "crux_ctx_set_stream": [_vp, _vp],
    "crux_ctx_stream": [_vp, _pp],
    "crux_ctx_sync": [_vp],
    "crux_last_error": [_vp],
    "crux_ctx_launch_count": [_vp, C.POINTER(_i64)],
    "crux_ctx_check": [_vp],
    "crux_ctx_timing_begin": [_vp],
    "crux_ctx_timing_end": [_vp, _vp, _vp],
    "crux_dev_alloc": [_vp, C.c_size_t, _pp],
    "crux_dev_free": [_vp, _vp],
    "crux_pinned_alloc": [_vp, C.c_size_t, _pp],
    "crux_pinned_free": [_vp, _vp],
    "crux_memcpy_h2d": [_vp, _vp, _vp, C.c_size_t],
    "crux_memcpy_d2h": [_vp, _vp, _vp, C.c_size_t],
    "crux_memcpy_d2d": [_vp, _vp, _vp, C.c_size_t],
    "crux_memset": [_vp, _vp, _i32, C.c_size_t],
    "crux_graph_begin": [_vp],
    "crux_graph_end": [_vp, _pp],
    "crux_graph_launch": [_vp, _vp],
    "crux_graph_destroy": [_vp, _vp],
    "crux_mlp_create": [_vp, _i32, C.POINTER(_i32), C.POINTER(_i32), _pp],
    "crux_mlp_destroy": [_vp],
    "crux_mlp_num_params": [_vp, C.POINTER(_i64)],
    "crux_mlp_set_params": [_vp, _vp],
    "crux_mlp_get_params": [_vp, _vp],
    "crux_mlp_params_ptr": [_vp, _pp],
    "crux_mlp_grads_ptr": [_vp, _pp],
    "crux_mlp_set_adam": [_vp, _f64, _f64, _f64, _f64],
    "crux_mlp_forward": [_vp, _vp, _i64, _vp],
    "crux_mlp_forward_sa": [_vp, _vp, _i32, _vp, _i32, _i64, _vp],
    "crux_value_next": [_vp, _vp, _vp, _vp, _i64, _i64, _vp],
    "crux_mlp_copy": [_vp, _vp],
    "crux_mlp_polyak": [_vp, _vp, _f32],
    "crux_mlp_train_mse": [_vp, _vp, _vp, _i64, _vp],
    "crux_gaussian_create": [_vp, _vp, _i32, _vp, _i32, _f32, _pp],
    "crux_gaussian_destroy": [_vp],
    "crux_categorical_create": [_vp, _vp, _i32, _pp],
    "crux_gaussian_log_sigma_ptr": [_vp, _pp],
    "crux_gaussian_explore": [_vp, _vp, _i64, _vp, _u64, _u64, _vp, _vp],
    "crux_gaussian_action": [_vp, _vp, _i64, _vp],
    "crux_gaussian_logpdf": [_vp, _vp, _vp, _i64, _vp],
    "crux_gaussian_entropy": [_vp, _vp, _i64, _vp],
    "crux_discrete_argmax": [_vp, _vp, _i64, _i32, _vp],
    "crux_discrete_explore": [_vp, _vp, _i64, _i32, _vp, _u64, _u64, _vp, _vp],
    "crux_discrete_explore_t": [_vp, _vp, _i64, _i32, _f32, _vp, _u64, _u64, _vp, _vp],
    "crux_discrete_logpdf_t": [_vp, _vp, _vp, _i64, _i32, _f32, _vp],
    "crux_discrete_entropy_t": [_vp, _vp, _i64, _i32, _f32, _vp],
    "crux_softq_target": [_vp, _vp, _vp, _vp, _i64, _i32, _f32, _f32, _vp],
    "crux_discrete_logpdf": [_vp, _vp, _vp, _i64, _i32, _vp],
    "crux_discrete_entropy": [_vp, _vp, _i64, _i32, _vp],
    "crux_discrete_eps_greedy": [_vp, _vp, _i64, _i32, _f64, _vp, _u64, _u64, _vp, _vp, _vp],
    "crux_rollout_step": [_vp, _vp, _vp, _i64, _vp, _u64, _u64, _vp, _vp, _vp],
    "crux_rollout_host": [_vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, C.POINTER(RolloutCols), _u64, _u64],
    "crux_normalize_obs": [_vp, _vp, _i64, _f32, _f32, _vp],
    "crux_fill_gae_returns": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _f32, _f32, _vp, _vp],
    "crux_whiten": [_vp, _vp, _i64],
    "crux_noise_explore": [_vp, _vp, _i64, _i32, _f32, _f32, _f32, _vp, _i32, _vp, _i32, _vp, _u64, _u64],
    "crux_ddpg_create": [_vp, _vp, _vp, _vp, _vp, _vp, _f32, _pp],
    "crux_ddpg_destroy": [_vp],
    "crux_ddpg_train": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _f32, _i32, _f32, _f32, _f32, _vp, _i32, _vp, _i32, _vp, _u64, _u64, _i32, _i32,
                        _vp, _vp],
    "crux_dqn_target": [_vp, _vp, _vp, _vp, _i64, _i32, _f32, _vp],
    "crux_sac_target": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _f32, _vp, _vp],
    "crux_td_error": [_vp, _vp, _vp, _i64, _vp],
    "crux_discrete_q_sa": [_vp, _vp, _vp, _i64, _i32, _vp],
    "crux_ppo_update": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(PPOHp), _vp, _vp, _u64, _vp, _vp],
    "crux_ppo_update_async": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(PPOHp), _vp, _vp, _u64],
    "crux_ppo_info_ptrs": [_vp, _pp, _pp],
    "crux_lagrange_ppo_update": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(PPOHp), C.POINTER(LagrangeHp),
                                 _vp, _vp, _vp, _vp, _u64, _vp, _vp, _vp, _vp],
    "crux_dqn_train": [_vp, _vp, _vp, _vp, _vp, _i64, _vp],
    "crux_convq_create": [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _pp],
    "crux_convq_destroy": [_vp],
    "crux_convq_num_params": [_vp, C.POINTER(_i64)],
    "crux_convq_shape": [_vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32)],
    "crux_convq_set_params": [_vp, _vp],
    "crux_convq_get_params": [_vp, _vp],
    "crux_convq_grads": [_vp, _vp],
    "crux_convq_set_adam": [_vp, _f64, _f64, _f64, _f64, _f32],
    "crux_convq_forward": [_vp, _vp, _i32, _i64, _vp],
    "crux_convq_copy": [_vp, _vp],
    "crux_convq_polyak": [_vp, _vp, _f32],
    "crux_convq_dqn_train": [_vp, _vp, _i32, _vp, _vp, _vp, _i64, _vp],
    "crux_sac_create": [_vp, _vp, _vp, _vp, _vp, _f32, _f32, _f64, _f32, _pp],
    "crux_sac_destroy": [_vp],
    "crux_sac_log_alpha": [_vp, _vp],
    "crux_sac_train": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _f32, _vp, _vp, _vp, _u64, _u64, _vp, _vp],
    "crux_buffer_create": [_vp, _i64, _i32, C.POINTER(ColDesc), _i32, _f32, _pp],
    "crux_buffer_destroy": [_vp],
    "crux_buffer_col": [_vp, _i32, _pp, C.POINTER(_i64), C.POINTER(_i32)],
    "crux_buffer_state": [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)],
    "crux_buffer_clear": [_vp],
    "crux_buffer_push": [_vp, _i64, _i32, C.POINTER(_i32), _pp, _i32, _vp, C.POINTER(_i64)],
    "crux_buffer_push_from": [_vp, _vp, _i64, _vp],
    "crux_buffer_last_n_indices": [_vp, _i64, _vp, C.POINTER(_i64)],
    "crux_buffer_sample_uniform": [_vp, _vp, _i64, _vp, _u64, _u64],
    "crux_buffer_sample_prioritized": [_vp, _vp, _i64, _f32, _i32, _vp, _u64, _u64],
    "crux_buffer_indices": [_vp, _pp, C.POINTER(_i64)],
    "crux_buffer_update_priorities": [_vp, _vp, _vp, _i64],
    "crux_buffer_priorities": [_vp, _pp, _pp, C.POINTER(_f32), C.POINTER(_f32)],
    "crux_split_batches": [_i64, C.POINTER(_f64), _i32, C.POINTER(_i64)],
    "crux_gather_rows": [_vp, _vp, _vp, _vp, _i64, _i64],
    "crux_linquad_create": [_vp, _i32, _i32, _vp, _vp, _i64, _i32, _u64, _pp],
    "crux_linquad_destroy": [_vp],
    "crux_linquad_reset": [_vp, _vp],
    "crux_linquad_step": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32],
    "crux_linquad_rollout": [_vp, _vp, _i32, _i32, _vp, C.POINTER(RolloutCols), _u64, _u64],
    "crux_nccl_unique_id": [_vp],
    "crux_nccl_init": [_vp, _i32, _i32, _vp],
    "crux_nccl_allreduce_f32": [_vp, _vp, _i64],
    "crux_peer_handle": [_vp, _vp, _i64],
    "crux_peer_init": [_vp, _i32, _i32, _vp],
    "crux_peer_disable": [_vp],
}
_RESTYPE = {"crux_last_error": C.c_char_p}

_lib = None


def declared_symbols():
    return sorted(_SIGS)


def load():
    """dlopen the in-tree library (building is ``crux.jl_b200/build.py``'s job)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python crux.jl_b200/build.py` (nvcc, sm_100a). "
                          "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, args in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = _RESTYPE.get(name, C.c_int32)
    _lib = lib
    return lib


def check(rc, ctx=None):
    if rc == OK:
        return
    msg = load().crux_last_error(ctx)
    msg = msg.decode() if msg else ""
    raise (NaNError if rc == ERR_NAN else CruxError)(rc, msg)
