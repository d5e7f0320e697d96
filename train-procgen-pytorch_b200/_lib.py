"""ctypes binding of the C-ABI in include/tpp_b200.h.

The product path has NO fallback: if ``csrc/libtpp_b200.so`` is missing or a call returns a non-zero status,
a ``TppError`` is raised.  (Build with ``python -c "import __graft_entry__ as g; g.build()"``.)
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TPP_B200_LIB") or os.path.join(_HERE, "csrc", "libtpp_b200.so")
ABI_VERSION = 3
ENOTSUP = 10002          # TPP_ENOTSUP


class TppError(RuntimeError):
    pass


class EnvCfg(C.Structure):
    _fields_ = [("family", C.c_int32), ("n_envs", C.c_int32), ("max_steps", C.c_int32), ("n_state", C.c_int32),
                ("seed", C.c_uint64), ("start_low", C.c_float * 16), ("start_high", C.c_float * 16),
                ("p", C.c_float * 8)]


class BoxWorldState(C.Structure):
    _fields_ = [("n_envs", C.c_int32), ("n", C.c_int32), ("max_steps", C.c_int32), ("n_levels", C.c_int32),
                ("start_seed", C.c_int64), ("goal_length", C.c_int32), ("num_distractor", C.c_int32),
                ("distractor_length", C.c_int32), ("_pad", C.c_int32),
                ("world", C.c_void_p), ("world_dic", C.c_void_p), ("player_pos", C.c_void_p),
                ("owned_key", C.c_void_p), ("num_env_steps", C.c_void_p), ("episode_reward", C.c_void_p),
                ("seed_counter", C.c_void_p), ("bank_world", C.c_void_p), ("bank_dic", C.c_void_p),
                ("bank_pos", C.c_void_p), ("scratch", C.c_void_p)]


class LossCfg(C.Structure):
    _fields_ = [("eps_clip", C.c_float), ("value_coef", C.c_float), ("entropy_coef", C.c_float),
                ("entropy_multiplier", C.c_float), ("x_entropy_coef", C.c_float), ("n_actions", C.c_int32),
                ("mb", C.c_int32), ("_pad", C.c_int32), ("coef_dev", C.c_void_p)]


class AdamState(C.Structure):   # mirrors tpp_adam_state (lives in device memory; this is the host image)
    _fields_ = [("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
                ("max_grad_norm", C.c_float), ("grad_scale", C.c_float), ("step", C.c_int32), ("ticket", C.c_int32),
                ("sqnorm", C.c_double * 2)]


class TcGemm(C.Structure):   # mirrors tpp_tc_gemm
    _fields_ = [("a_hi", C.c_void_p), ("a_lo", C.c_void_p), ("lda", C.c_int64),
                ("b_hi", C.c_void_p), ("b_lo", C.c_void_p), ("ldb", C.c_int64),
                ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
                ("precision", C.c_int32), ("split_k", C.c_int32), ("flags", C.c_int32), ("block_n", C.c_int32),
                ("a_mn", C.c_int32), ("b_mn", C.c_int32), ("conv_wgrad", C.c_int32),
                ("bias", C.c_void_p), ("mask", C.c_void_p), ("ld_mask", C.c_int64),
                ("out", C.c_void_p), ("out_hi", C.c_void_p), ("out_lo", C.c_void_p), ("ldc", C.c_int64),
                ("colsum", C.c_void_p), ("dbg", C.c_void_p), ("addend", C.c_void_p), ("ld_add", C.c_int64),
                ("conv_B", C.c_int32), ("conv_H", C.c_int32), ("conv_W", C.c_int32), ("conv_C", C.c_int32),
                ("alpha", C.c_float), ("_reserved", C.c_int32), ("mask_bits_out", C.c_void_p), ("mask_bits", C.c_void_p)]


class FusedPolicy(C.Structure):   # mirrors tpp_fused_policy
    _fields_ = [("n_rows", C.c_int32), ("a1_mode", C.c_int32), ("x", C.c_void_p), ("ldx", C.c_int64),
                ("w_hi", C.c_void_p * 4), ("w_lo", C.c_void_p * 4), ("ldw", C.c_int64 * 4),
                ("k", C.c_int32 * 4), ("n", C.c_int32 * 4), ("bias", C.c_void_p * 4), ("relu", C.c_int32 * 4),
                ("head_w", C.c_void_p), ("head_b", C.c_void_p), ("n_actions", C.c_int32), ("ld_head", C.c_int32),
                ("act", C.c_void_p), ("logp", C.c_void_p), ("value", C.c_void_p), ("head_out", C.c_void_p),
                ("seed", C.c_uint64), ("tick", C.c_void_p), ("t_offset", C.c_uint64), ("greedy", C.c_int32),
                ("env_offset", C.c_int32), ("dbg", C.c_void_p), ("no_pdl", C.c_int32), ("_pad", C.c_int32),
                ("scratch", C.c_void_p), ("scratch_bytes", C.c_int64)]


FAMILY = {"cartpole": 0, "cartpole_swing": 1, "mountain_car": 2, "acrobot": 3, "lunar_lander": 4}
TC_A_EXACT, TC_B_EXACT = 16, 32
TC_A_SPLIT, TC_B_SPLIT = 64, 128      # the operand is plain fp32; its lo half is formed in shared memory
TC_TILE_PAIR, TC_TILE_PAIR_PERSISTENT, TC_TILE_PAIR64_PERSISTENT = 512, 513, 65     # tpp_tc_gemm.block_n codes
TC_TILE_PAIR_PERSISTENT_LEAN = 514          # 3 operand stages, 8 epilogue warps
EPI_BIAS, EPI_RELU, EPI_MASK, EPI_ACCUM, EPI_ADD, EPI_RELU_OUT, EPI_PAIR_RELU = 1, 2, 4, 8, 16, 32, 64

_vp, _i32, _i64, _u64, _f32, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_double

class WeightView(C.Structure):            # == tpp_weight_view
    _fields_ = [("offset", C.c_int64), ("rows", C.c_int32), ("cols", C.c_int32), ("hi", C.c_void_p), ("lo", C.c_void_p),
                ("ld", C.c_int64), ("scale", C.c_float), ("_pad", C.c_int32), ("col_of", C.c_void_p)]


MAX_WEIGHT_VIEWS = 8

# name -> argtypes (every function returns int status unless listed in _SPECIAL)
SIGNATURES = {
    "tpp_version": [],
    "tpp_device_sm_count": [C.POINTER(C.c_int)],
    "tpp_env_step": [C.POINTER(EnvCfg), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _u64, _i64, _vp],
    "tpp_env_reset": [C.POINTER(EnvCfg), _vp, _vp, _vp, _vp, _vp, _u64, _i64, _vp],
    "tpp_tick_advance": [_vp, _u64, _vp],
    "tpp_randperm_mt19937": [_vp, _vp, _vp, _i64, _vp],
    "tpp_randperm_mt19937_i32": [_vp, _vp, _vp, _i64, _vp],
    "tpp_boxworld_step": [C.POINTER(BoxWorldState), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp],
    "tpp_boxworld_gen_levels_host": [_i32, _i32, _i32, _i32, _i64, _i32, _vp, _vp, _vp],
    "tpp_boxworld_gen_levels_device": [C.POINTER(BoxWorldState), _vp, _vp, _i32, _vp],
    "tpp_boxworld_emit_frames": [C.POINTER(BoxWorldState), _vp, _vp],
    "tpp_vecnormalize_step": [_vp, _vp, _vp, C.c_int, _vp, _vp, _i32, _f64, _f64, _f64, _vp],
    "tpp_gae": [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i64, _f32, _f32, _vp],
    "tpp_gae_scan": [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i64, _f32, _f32, _i32, _vp],
    "tpp_adv_normalize": [_vp, _vp, _i32, _i32, _i64, _vp],
    "tpp_episode_scan": [_vp, _vp, _i32, _i32, _i64, _vp, _vp, _vp, _vp, _i32, _vp],
    "tpp_gather_vec": [_vp, _i32, _i32, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp,
                       _vp, _vp, _vp, _vp],
    "tpp_gather_img": [_vp, _i32, _i32, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp,
                       _vp, _vp, _vp, _vp, _vp, _i32, _vp],
    "tpp_frames_to_obs": [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _i32, _vp],
    "tpp_gemm_f32": [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp],
    "tpp_colsum_accum": [_vp, _i64, _i32, _i32, _vp, _vp],
    "tpp_gemm_tc": [C.POINTER(TcGemm), _vp],
    "tpp_split_tf32": [_vp, _i64, _i32, _i32, _vp, _vp, _i64, _vp, _vp, _i64, _vp],
    "tpp_im2col3x3": [_vp, _i32, _i32, _i32, _i32, _i32, _i64, _i64, _i64, _i64, _i32, _f32, _vp, _vp, _i32, _vp],
    "tpp_colsum_narrow": [_vp, _i64, _i32, _vp, _vp],
    "tpp_feature_sparsity": [_vp, _vp, _i32, _i32, _vp, _vp, _vp],
    "tpp_feature_sparsity_grad": [_vp, _i32, _f32, _vp, _vp, _vp, _vp],
    "tpp_bias_act_split": [_vp, _i64, _i32, _i32, _vp, _i32, _vp, _vp, _vp, _i64, _vp],
    "tpp_gru_mask_split": [_vp, _i64, _vp, _i32, _i32, _vp, _vp, _i64, _vp],
    "tpp_gru_gates": [_vp, _vp, _i64, _vp, _i64, _vp, _i32, _i32, _vp, _i64, _vp, _vp, _i64, _vp],
    "tpp_conv3x3_wgrad": [_vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp],
    "tpp_conv3x3_wgrad_first": [_vp, _i64, _i64, _i64, _vp, _vp, _i32, _i32, _i32, _i32, _vp],
    "tpp_conv3x3_fwd_first": [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp],
    "tpp_conv3x3_fma": [_vp, _i32, _vp, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp],
    "tpp_maxpool3x3s2_fwd": [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp],
    "tpp_maxpool3x3s2_bwd": [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp],
    "tpp_head_backward": [_vp, _i32, _vp, _vp, _i64, _vp, _i32, _i32, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i32, _vp],
    "tpp_sample_actions": [_vp, _i32, _i32, _i32, _vp, _vp, _vp, _u64, _vp, _u64, _i32, _i32, _vp],
    "tpp_mlp_tail_sample": [_vp, _i64, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _i32, _i32, _vp, _i32, _vp, _vp, _vp, _u64,
                            _vp, _u64, _i32, _i32, _vp],
    "tpp_policy_rollout_fused": [C.POINTER(FusedPolicy), _vp],
    "tpp_vecnormalize_rollout": [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i64, _f64, _f64, _f64, _vp, _i64, _vp],
    "tpp_ppo_loss_fwd_bwd": [C.POINTER(LossCfg), _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "tpp_ppo_pbar": [_vp, _i32, _i32, _i32, _vp, _vp],
    "tpp_ppo_loss_fwd_bwd_grouped": [C.POINTER(LossCfg), _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp],
    "tpp_grad_sqnorm": [_vp, _vp, _i64, _vp],
    "tpp_adam_clip_step": [_vp, _vp, _vp, _vp, _vp, _i64, _vp],
    "tpp_adam_clip_step_views": [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _i32, _vp],
    "tpp_peer_allreduce_sqnorm": [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp],
}
_NO_STATUS = {"tpp_version"}

_lib = None


def load():
    """Load the shared library once; raise loudly when it is absent (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TppError(f"{LIB_PATH} not found: the CUDA extension is not built. "
                       "Run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
                       "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.tpp_error_string.argtypes = [C.c_int]
    lib.tpp_error_string.restype = C.c_char_p
    if lib.tpp_version() != ABI_VERSION:
        raise TppError(f"ABI mismatch: library {lib.tpp_version()} vs binding {ABI_VERSION}")
    _lib = lib
    return lib


def call(name, *args):
    """Invoke an entry point and raise TppError on a non-zero status."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0 and name not in _NO_STATUS:
        msg = lib.tpp_error_string(rc)
        raise TppError(f"{name} failed with status {rc}: {msg.decode() if msg else '?'}")
    return rc


def try_call(name, *args):
    """``call`` for entry points that are built for a fixed set of shapes: False when the library answers TPP_ENOTSUP
    (the caller then takes its general path), True on success, TppError on anything else."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc == ENOTSUP:
        return False
    if rc != 0:
        msg = lib.tpp_error_string(rc)
        raise TppError(f"{name} failed with status {rc}: {msg.decode() if msg else '?'}")
    return True


def ptr(t):
    """Device (or host) address of a torch tensor / None -> NULL."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
