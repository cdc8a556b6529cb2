"""ctypes binding of libaread_sm100.so (the C ABI declared in include/aread_sm100.h).

There is deliberately no fallback: if the library is missing or fails to load, every op raises.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_double, c_float, c_int32, c_int64, c_size_t, c_uint32, c_uint64,
                    c_void_p)

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libaread_sm100.so")
ABI_VERSION = 12

AREAD_OK = 0
AREAD_ERR_INVALID = -1
AREAD_ERR_INDEX = -2
AREAD_ERR_CUDA = -3
AREAD_ERR_WORKSPACE = -4


class AreadError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libaread_sm100: {message} (status {code})")
        self.code = code


class EmbedPlan(Structure):
    _fields_ = [("n_cols", c_int32), ("n_fields", c_int32), ("max_src", c_int32), ("embed_dim", c_int32),
                ("n_rows", c_int64), ("col_offset", c_void_p), ("field_src", c_void_p),
                ("field_nsrc", c_void_p), ("field_div", c_void_p)]


class GatherArgs(Structure):
    _fields_ = [("plan", EmbedPlan), ("batch", c_int64), ("x", c_void_p), ("table", c_void_p),
                ("out", c_void_p), ("out_bf16", c_void_p), ("out_bf16_lo", c_void_p), ("status", c_void_p),
                ("shard_shift", c_int32), ("shards", c_void_p)]


class ScatterArgs(Structure):
    _fields_ = [("plan", EmbedPlan), ("batch", c_int64), ("x", c_void_p), ("d_out", c_void_p),
                ("d_table", c_void_p), ("zero_fill", c_int32), ("workspace", c_void_p),
                ("workspace_bytes", c_size_t), ("sorted_rows", c_void_p), ("sorted_pos", c_void_p),
                ("shard_shift", c_int32), ("shard_rows", c_int64), ("peer_grads", c_void_p)]


class GroupedLinearArgs(Structure):
    _fields_ = [("m", c_int64), ("n", c_int32), ("k", c_int32), ("groups", c_int32), ("a_group_cols", c_int32),
                ("group_mask", c_uint64), ("a", c_void_p), ("lda", c_int64), ("b", c_void_p), ("ldb", c_int64),
                ("bias", c_void_p), ("c_f32", c_void_p), ("c_bf16", c_void_p), ("ldc", c_int64),
                ("a_lo", c_void_p), ("b_lo", c_void_p), ("lo_lo", c_int32)]


class GroupedWgradArgs(Structure):
    _fields_ = [("m", c_int64), ("n", c_int32), ("k", c_int32), ("groups", c_int32), ("a_group_cols", c_int32),
                ("group_mask", c_uint64), ("dz", c_void_p), ("ldz", c_int64), ("a", c_void_p), ("lda", c_int64),
                ("dw", c_void_p), ("workspace", c_void_p), ("workspace_bytes", c_size_t),
                ("dz_lo", c_void_p), ("a_lo", c_void_p)]


class BnActArgs(Structure):
    _fields_ = [("m", c_int64), ("width", c_int32), ("training", c_int32), ("bn_skip", c_int32),
                ("momentum", c_float), ("eps", c_float), ("dropout_p", c_float), ("seed", c_uint64),
                ("salt", c_uint32), ("z", c_void_p), ("ldz", c_int64), ("gamma", c_void_p), ("beta", c_void_p),
                ("running_mean", c_void_p), ("running_var", c_void_p), ("mean", c_void_p), ("rstd", c_void_p),
                ("scale", c_void_p), ("shift", c_void_p), ("out_f32", c_void_p), ("out_bf16", c_void_p),
                ("ldo", c_int64), ("workspace", c_void_p), ("workspace_bytes", c_size_t), ("out_bf16_lo", c_void_p),
                ("seed_ptr", c_void_p)]


class BnActBwdArgs(Structure):
    _fields_ = [("m", c_int64), ("width", c_int32), ("bn_skip", c_int32), ("dropout_p", c_float),
                ("salt", c_uint32), ("seed", c_uint64), ("z", c_void_p), ("ldz", c_int64), ("d_out", c_void_p),
                ("ldd", c_int64), ("mean", c_void_p), ("rstd", c_void_p), ("scale", c_void_p), ("shift", c_void_p),
                ("d_gamma", c_void_p), ("d_beta", c_void_p), ("d_bias", c_void_p), ("dz_f32", c_void_p),
                ("dz_bf16", c_void_p), ("ldo", c_int64), ("workspace", c_void_p), ("workspace_bytes", c_size_t),
                ("dz_bf16_lo", c_void_p), ("seed_ptr", c_void_p), ("z_bf16", c_void_p)]


class MmoeMixArgs(Structure):
    _fields_ = [("m", c_int64), ("width", c_int32), ("n_expert", c_int32), ("n_gate", c_int32),
                ("dropout_p", c_float), ("seed", c_uint64), ("salt", c_uint32), ("z", c_void_p), ("ldz", c_int64),
                ("scale", c_void_p), ("shift", c_void_p), ("gate", c_void_p), ("out", c_void_p),
                ("d_out", c_void_p), ("d_h", c_void_p), ("d_gate", c_void_p), ("seed_ptr", c_void_p),
                ("z_bf16", c_void_p)]


class RowpassArgs(Structure):
    _fields_ = [("m", c_int64), ("e", c_int32), ("n_gate", c_int32), ("n_expert", c_int32), ("n_cross", c_int32),
                ("n_head", c_int32), ("ldp", c_int32), ("x", c_void_p), ("w", c_void_p), ("offset", c_void_p),
                ("p", c_void_p), ("lin", c_void_p), ("gate", c_void_p), ("alpha", c_void_p), ("head", c_void_p),
                ("d_lin", c_void_p), ("d_gate", c_void_p), ("d_head", c_void_p), ("d_p", c_void_p),
                ("d_c", c_void_p), ("d_x", c_void_p), ("d_w", c_void_p), ("workspace", c_void_p),
                ("workspace_bytes", c_size_t), ("dp16", c_void_p), ("ld16", c_int64), ("n_extra", c_int32),
                ("dp16_width", c_int32), ("d_c_sum", c_void_p), ("d_c_partial", c_void_p)]


class L2RegArgs(Structure):
    _fields_ = [("n_tensors", c_int32), ("n_chunks", c_int64), ("tensors", c_void_p), ("grads", c_void_p),
                ("sizes", c_void_p), ("l2", c_void_p), ("chunk_start", c_void_p), ("out", c_void_p),
                ("workspace", c_void_p), ("workspace_bytes", c_size_t)]


class HeiLayerFwdArgs(Structure):
    _fields_ = [("m", c_int64), ("groups", c_int32), ("k", c_int32), ("n", c_int32), ("training", c_int32),
                ("bn_skip", c_int32), ("momentum", c_float), ("eps", c_float), ("src", c_void_p), ("ld_src", c_int64),
                ("src_scale", c_void_p), ("src_shift", c_void_p), ("src_p", c_float), ("src_salt", c_uint32),
                ("seed", c_uint64), ("weight", c_void_p), ("bias", c_void_p), ("gamma", c_void_p), ("beta", c_void_p),
                ("running_mean", c_void_p), ("running_var", c_void_p), ("z", c_void_p), ("mean", c_void_p),
                ("rstd", c_void_p), ("scale", c_void_p), ("shift", c_void_p), ("workspace", c_void_p),
                ("workspace_bytes", c_size_t), ("seed_ptr", c_void_p)]


class HeiLayerBwdArgs(Structure):
    _fields_ = [("m", c_int64), ("groups", c_int32), ("k", c_int32), ("n", c_int32), ("bn_skip", c_int32),
                ("p", c_float), ("salt", c_uint32), ("seed", c_uint64), ("z", c_void_p), ("d_out", c_void_p),
                ("mean", c_void_p), ("rstd", c_void_p), ("scale", c_void_p), ("shift", c_void_p), ("coef", c_void_p),
                ("src", c_void_p), ("ld_src", c_int64), ("src_scale", c_void_p), ("src_shift", c_void_p),
                ("src_mean", c_void_p), ("src_rstd", c_void_p), ("src_p", c_float), ("src_salt", c_uint32),
                ("weight", c_void_p), ("d_in", c_void_p), ("d_w", c_void_p), ("src_coef", c_void_p),
                ("src_d_gamma", c_void_p), ("src_d_beta", c_void_p), ("src_d_bias", c_void_p), ("workspace", c_void_p),
                ("workspace_bytes", c_size_t), ("seed_ptr", c_void_p)]


class HeadArgs(Structure):
    _fields_ = [("m", c_int64), ("n_tower", c_int32), ("width", c_int32), ("head_cross", c_void_p), ("lin", c_void_p),
                ("h", c_void_p), ("w_tail", c_void_p), ("probs", c_void_p), ("d_probs", c_void_p), ("dz", c_void_p),
                ("d_lin", c_void_p), ("d_h", c_void_p), ("d_w_tail", c_void_p), ("workspace", c_void_p),
                ("workspace_bytes", c_size_t)]


class BaggingBceArgs(Structure):
    _fields_ = [("m", c_int64), ("n_tower", c_int32), ("probs", c_void_p), ("labels", c_void_p), ("loss", c_void_p),
                ("d_probs", c_void_p), ("workspace", c_void_p), ("workspace_bytes", c_size_t)]


class TowerLinearArgs(Structure):
    _fields_ = [("m", c_int64), ("groups", c_int32), ("in_width", c_int32), ("out_width", c_int32),
                ("weight_is_out_by_in", c_int32), ("in_", c_void_p), ("ld_in", c_int64),
                ("in_group_stride", c_int64), ("weight", c_void_p),
                ("bias", c_void_p), ("out", c_void_p), ("ld_out", c_int64)]


class TowerWgradArgs(Structure):
    _fields_ = [("m", c_int64), ("groups", c_int32), ("n", c_int32), ("k", c_int32), ("dz", c_void_p),
                ("ld_dz", c_int64), ("in_", c_void_p), ("ld_in", c_int64), ("in_group_stride", c_int64),
                ("d_w", c_void_p),
                ("workspace", c_void_p), ("workspace_bytes", c_size_t)]


class GateMixArgs(Structure):
    _fields_ = [("m", c_int64), ("n_tower", c_int32), ("n_prev", c_int32), ("n_prev_active", c_int32),
                ("width", c_int32), ("logits", c_void_p), ("edges", c_void_p), ("prev_slot", c_void_p),
                ("slot_tower", c_void_p), ("u_prev", c_void_p), ("out", c_void_p), ("sm", c_void_p),
                ("d_out", c_void_p), ("d_logits", c_void_p), ("d_u_prev", c_void_p), ("r_scratch", c_void_p),
                ("ld_logits", c_int64), ("logit_offset", c_void_p), ("ld_dlogits", c_int64)]


class AdamArgs(Structure):
    _fields_ = [("n_tensors", c_int32), ("n_chunks", c_int64), ("params", c_void_p), ("grads", c_void_p),
                ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p), ("sizes", c_void_p), ("chunk_start", c_void_p),
                ("step_size", c_void_p), ("bc2_sqrt", c_void_p), ("beta1", c_float), ("beta2", c_float),
                ("eps", c_float), ("weight_decay", c_float), ("l2_twice", c_void_p), ("step_counts", c_void_p),
                ("slot", c_void_p), ("lr", c_double), ("beta1_d", c_double), ("beta2_d", c_double)]


class ExpertGemmArgs(Structure):
    _fields_ = [("m", c_int64), ("n", c_int32), ("k", c_int32), ("groups", c_int32), ("a_group_cols", c_int32),
                ("a", c_void_p), ("lda", c_int64), ("b", c_void_p), ("ldb", c_int64), ("b_is_k_by_n", c_int32),
                ("epilogue", c_int32), ("bias", c_void_p), ("c_f32", c_void_p), ("c_bf16", c_void_p), ("ldc", c_int64),
                ("partial", c_void_p), ("scale", c_void_p), ("shift", c_void_p), ("mean", c_void_p), ("rstd", c_void_p),
                ("z", c_void_p), ("ldz", c_int64), ("dropout_p", c_float), ("salt", c_uint32), ("seed", c_uint64),
                ("seed_ptr", c_void_p)]


class ExpertBnFinalizeArgs(Structure):
    _fields_ = [("m", c_int64), ("width", c_int32), ("n_partial", c_int32), ("training", c_int32), ("bn_skip", c_int32),
                ("momentum", c_float), ("eps", c_float), ("partial", c_void_p), ("bias", c_void_p), ("gamma", c_void_p),
                ("beta", c_void_p), ("running_mean", c_void_p), ("running_var", c_void_p), ("mean", c_void_p),
                ("rstd", c_void_p), ("scale", c_void_p), ("shift", c_void_p)]


class ExpertBnBwdFinalizeArgs(Structure):
    _fields_ = [("m", c_int64), ("width", c_int32), ("n_partial", c_int32), ("bn_skip", c_int32), ("partial", c_void_p),
                ("d_gamma", c_void_p), ("d_beta", c_void_p), ("d_bias", c_void_p), ("coef", c_void_p)]


class Bn16Args(Structure):
    _fields_ = [("m", c_int64), ("width", c_int32), ("bn_skip", c_int32), ("z", c_void_p), ("ldz", c_int64),
                ("scale", c_void_p), ("shift", c_void_p), ("dropout_p", c_float), ("salt", c_uint32), ("seed", c_uint64),
                ("seed_ptr", c_void_p), ("out", c_void_p), ("ldo", c_int64), ("dy", c_void_p), ("ldd", c_int64),
                ("mean", c_void_p), ("rstd", c_void_p), ("coef", c_void_p), ("dy_is_raw", c_int32),
                ("pass_bits", c_void_p), ("keep_scale_bwd", c_float)]


_L, _J = 4, 3        # AREAD_MIXED_MAX_LEVEL, AREAD_MIXED_MAX_LAYER


class HeiMixedArgs(Structure):
    _fields_ = [("m", c_int64), ("n_level", c_int32), ("n_layer", c_int32), ("n_tower", c_int32 * _L),
                ("width_in", c_int32), ("dims", (c_int32 * _J) * _L), ("n_domain", c_int32), ("edge_words", c_int32),
                ("bn_skip", c_int32), ("eps", c_float), ("domain", c_void_p), ("domain_stride", c_int64),
                ("active", c_void_p), ("edges", c_void_p), ("t0", c_void_p), ("logits", c_void_p * _L),
                ("weight", (c_void_p * _J) * _L), ("bias", (c_void_p * _J) * _L), ("gamma", (c_void_p * _J) * _L),
                ("beta", (c_void_p * _J) * _L), ("running_mean", (c_void_p * _J) * _L),
                ("running_var", (c_void_p * _J) * _L), ("head_cross", c_void_p), ("lin", c_void_p), ("w_tail", c_void_p),
                ("y", c_void_p), ("y_stack", c_void_p)]


class DomainMeanArgs(Structure):
    _fields_ = [("m", c_int64), ("width", c_int32), ("n_domain", c_int32), ("values", c_void_p), ("ld", c_int64),
                ("domain", c_void_p), ("domain_stride", c_int64), ("mean", c_void_p), ("count", c_void_p)]


class MultiCopyArgs(Structure):
    _fields_ = [("n_tensors", c_int32), ("n_chunks", c_int64), ("dst", c_void_p), ("src", c_void_p),
                ("bytes", c_void_p), ("chunk_start", c_void_p)]


_SIGNATURES = {
    "aread_last_error": (c_char_p, []),
    "aread_abi_version": (c_int32, []),
    "aread_launch_count": (c_uint64, []),
    "aread_launch_count_add": (None, [c_uint64]),
    "aread_gather_fwd": (c_int32, [POINTER(GatherArgs), c_void_p]),
    "aread_scatter_workspace_bytes": (c_size_t, [c_int64, c_int32]),
    "aread_scatter_bwd": (c_int32, [POINTER(ScatterArgs), c_void_p]),
    "aread_grouped_linear_bf16": (c_int32, [POINTER(GroupedLinearArgs), c_void_p]),
    "aread_grouped_wgrad_workspace_bytes": (c_size_t, [POINTER(GroupedWgradArgs)]),
    "aread_grouped_wgrad_bf16": (c_int32, [POINTER(GroupedWgradArgs), c_void_p]),
    "aread_bn_workspace_bytes": (c_size_t, [c_int32]),
    "aread_bn_act_fwd": (c_int32, [POINTER(BnActArgs), c_void_p]),
    "aread_bn_act_bwd": (c_int32, [POINTER(BnActBwdArgs), c_void_p]),
    "aread_dropout_mask": (c_int32, [c_uint64, c_uint32, c_int64, c_float, c_void_p, c_void_p]),
    "aread_mmoe_mix": (c_int32, [POINTER(MmoeMixArgs), c_void_p]),
    "aread_rowpass_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32]),
    "aread_rowpass_fwd": (c_int32, [POINTER(RowpassArgs), c_void_p]),
    "aread_rowpass_bwd": (c_int32, [POINTER(RowpassArgs), c_void_p]),
    "aread_tower_linear": (c_int32, [POINTER(TowerLinearArgs), c_void_p]),
    "aread_tower_wgrad_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32, c_int32]),
    "aread_tower_wgrad": (c_int32, [POINTER(TowerWgradArgs), c_void_p]),
    "aread_gate_mix": (c_int32, [POINTER(GateMixArgs), c_void_p]),
    "aread_ipc_export": (c_int32, [c_void_p, c_char_p, POINTER(c_int64)]),
    "aread_ipc_open": (c_int32, [c_char_p, c_int64, c_int32, POINTER(c_void_p)]),
    "aread_ipc_close": (c_int32, [c_void_p, c_int64]),
    "aread_adam_chunk": (c_int64, []),
    "aread_adam_step": (c_int32, [POINTER(AdamArgs), c_void_p]),
    "aread_shard_grad_sum": (c_int32, [c_void_p, c_int32, c_int64, c_float, c_void_p, c_void_p]),
    "aread_multi_copy_chunk": (c_int64, []),
    "aread_multi_copy": (c_int32, [POINTER(MultiCopyArgs), c_void_p]),
    "aread_multi_cast_bf16": (c_int32, [POINTER(MultiCopyArgs), c_void_p]),
    "aread_expert_gemm_partials": (c_int32, [c_int64]),
    "aread_expert_gemm": (c_int32, [POINTER(ExpertGemmArgs), c_void_p]),
    "aread_expert_bn_finalize": (c_int32, [POINTER(ExpertBnFinalizeArgs), c_void_p]),
    "aread_expert_bn_bwd_finalize": (c_int32, [POINTER(ExpertBnBwdFinalizeArgs), c_void_p]),
    "aread_bn16": (c_int32, [POINTER(Bn16Args), c_void_p]),
    "aread_bn16_partials": (c_int32, [c_int64, c_int32]),
    "aread_bn16_bwd_stats": (c_int32, [POINTER(Bn16Args), c_void_p, c_void_p]),
    "aread_hei_mixed_eval_supported": (c_int32, [POINTER(HeiMixedArgs)]),
    "aread_hei_mixed_eval": (c_int32, [POINTER(HeiMixedArgs), c_void_p]),
    "aread_domain_mean": (c_int32, [POINTER(DomainMeanArgs), c_void_p]),
    "aread_l2_reg_chunk": (c_int64, []),
    "aread_bn_act_apply": (c_int32, [POINTER(BnActArgs), c_void_p]),
    "aread_bn_bwd_coef": (c_int32, [POINTER(BnActBwdArgs), c_void_p, c_void_p]),
    "aread_hei_layer_supported": (c_int32, [c_int32, c_int32, c_int32]),
    "aread_rowpass_prologue_ctas": (c_int32, [c_int64]),
    "aread_hei_set_path": (None, [c_int32, c_int32]),
    "aread_hei_layer_path": (c_int32, [c_int64, c_int32, c_int32, c_int32]),
    "aread_hei_layer_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32, c_int32]),
    "aread_hei_layer_fwd": (c_int32, [POINTER(HeiLayerFwdArgs), c_void_p]),
    "aread_hei_layer_bwd": (c_int32, [POINTER(HeiLayerBwdArgs), c_void_p]),
    "aread_head": (c_int32, [POINTER(HeadArgs), c_void_p]),
    "aread_head_workspace_bytes": (c_size_t, [c_int32, c_int32]),
    "aread_bagging_bce_workspace_bytes": (c_size_t, [c_int64, c_int32]),
    "aread_bagging_bce": (c_int32, [POINTER(BaggingBceArgs), c_void_p]),
    "aread_l2_reg_fwd": (c_int32, [POINTER(L2RegArgs), c_void_p]),
    "aread_l2_reg_bwd": (c_int32, [POINTER(L2RegArgs), c_void_p, c_void_p]),
}

_lib = None


def load():
    """Load the shared library once; raise (never fall back) when it is unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AreadError(AREAD_ERR_INVALID,
                         f"{LIB_PATH} not found -- build it with `python {os.path.join(PKG_DIR, 'build.py')}`; "
                         "this package has no CPU or eager fallback")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.aread_abi_version() != ABI_VERSION:
        raise AreadError(AREAD_ERR_INVALID, f"ABI version {lib.aread_abi_version()} != binding {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(code):
    if code != AREAD_OK:
        msg = load().aread_last_error().decode(errors="replace")
        if code == AREAD_ERR_INDEX:
            raise IndexError(msg or "index out of range in self")
        raise AreadError(code, msg)


def launch_count():
    return int(load().aread_launch_count())


def exported_symbols():
    return sorted(_SIGNATURES)
