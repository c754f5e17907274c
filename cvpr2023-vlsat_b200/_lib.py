"""ctypes binding of libvlsat_b200.so (the C ABI declared in include/vlsat_b200.h).

There is no CPU or PyTorch fallback: if the library is missing, or a call returns a non-zero status,
a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvlsat_b200.so")

i64, i32, f32, vp, sz = C.c_int64, C.c_int, C.c_float, C.c_void_p, C.c_size_t


class Epilogue(C.Structure):
    """Mirror of ``vlsat_epilogue`` (include/vlsat_b200.h)."""
    _fields_ = [("bias", vp), ("gather_a", vp), ("idx_a", vp), ("gather_b", vp), ("idx_b", vp),
                ("ld_gather", i64), ("residual", vp), ("ld_res", i64), ("alpha", f32), ("beta", f32),
                ("scale_ptr", vp), ("act", i32), ("bias_per_row", i32),
                ("split_hi", vp), ("split_lo", vp), ("ld_split", i64), ("split_fmt", i32)]


class AdamWTensor(C.Structure):
    """Mirror of ``vlsat_adamw_tensor``."""
    _fields_ = [("p", vp), ("g", vp), ("m", vp), ("v", vp), ("vmax", vp), ("n", i64), ("lr", f32), ("weight_decay", f32)]


class CopyTensor(C.Structure):
    """Mirror of ``vlsat_copy_tensor``."""
    _fields_ = [("dst", vp), ("src", vp), ("n", i64)]


class LinearOpts(C.Structure):
    """Mirror of ``vlsat_linear_opts``."""
    _fields_ = [("engine", i32), ("x_hi", vp), ("x_lo", vp), ("w_hi", vp), ("w_lo", vp), ("workspace", vp),
                ("workspace_bytes", sz)]


class Bf16Pair(C.Structure):
    """Mirror of ``vlsat_bf16_pair``."""
    _fields_ = [("hi", vp), ("lo", vp), ("ld", i64)]


class FlashBwdOperands(C.Structure):
    """Mirror of ``vlsat_flash_bwd_operands``."""
    _fields_ = [("q", Bf16Pair), ("k", Bf16Pair), ("v", Bf16Pair), ("dout", Bf16Pair),
                ("q_t", Bf16Pair), ("k_t", Bf16Pair), ("dout_t", Bf16Pair),
                ("lse2", vp), ("delta", vp), ("ld_stat", i64)]


# name -> argtypes ; every function returns int status unless listed in _RESTYPES
SIGNATURES = {
    "vlsat_version": [],
    "vlsat_error_string": [i32],
    "vlsat_gemm_engine": [],
    "vlsat_launch_count": [],
    "vlsat_set_precision": [i32],
    "vlsat_get_precision": [],
    "vlsat_pointnet_fwd": [vp, i64, i32, i64, vp, vp, i32, vp, vp, i32, vp, vp, i32, vp, vp, vp],
    "vlsat_pointnet_tc_fwd": [vp, i64, i32, i64, vp, vp, i32, vp, vp, i32, vp, vp, i32, vp, vp, vp],
    "vlsat_edge_descriptor_fwd": [vp, i64, vp, i64, vp, vp],
    "vlsat_linear_fwd": [vp, i64, vp, i64, vp, i64, i64, i64, i64, C.POINTER(Epilogue), C.POINTER(LinearOpts), vp],
    "vlsat_linear_workspace_bytes": [i64, i64, i64, i32, i32],
    "vlsat_gemm_pairs": [i32, vp, vp, i64, vp, vp, i64, vp, i64, i64, i64, i64, vp, sz, vp],
    "vlsat_gemm_pairs_workspace_bytes": [i32, i64, i64, i64],
    "vlsat_tf32_split": [vp, i64, i64, i64, vp, vp, vp],
    "vlsat_add_layernorm_fwd": [vp, i64, vp, i64, vp, vp, vp, i64, i64, i32, f32, i32, vp, vp, i64, vp],
    "vlsat_relu_fwd": [vp, vp, i64, vp],
    "vlsat_relu_pair_fwd": [vp, vp, vp, vp, i64, vp],
    "vlsat_row_l2norm_fwd": [vp, vp, i64, i32, vp],
    "vlsat_spatial_tail_fwd": [vp, vp, i64, i32, i64, vp],
    "vlsat_scene_ranges": [vp, i64, vp, vp, vp, vp],
    "vlsat_node_attn_fwd": [vp, i64, vp, i64, vp, i64, vp, i64, vp, vp, vp, i32, i32, vp, i64, i64, i32, vp],
    "vlsat_node_bias_table_max_scene": [],
    "vlsat_node_bias_table": [vp, i64, vp, vp, vp, i32, vp, i64, vp],
    "vlsat_node_attn_scene_fwd": [vp, i64, vp, i64, vp, i64, vp, vp, vp, i32, i32, vp, i64, i64, vp],
    "vlsat_dense_attn_fwd": [vp, i64, vp, i64, vp, i64, vp, i64, i32, vp, i64, vp, i64, i64, i64, i32, i32, vp],
    "vlsat_flash_attn_fwd": [vp, i64, vp, i64, vp, i64, vp, i64, vp, i64, i64, i32, i32, vp],
    "vlsat_flash_attn_tc_fwd": [vp, vp, i64, vp, vp, i64, vp, vp, i64, vp, i64, vp, i64, i64, i32, i32, vp],
    "vlsat_gat_edge_tc_fwd": [vp, vp, vp, i64, vp, i64, vp, vp, vp, vp, vp, vp, vp, i64, i64, i32, i32, i32, i32,
                              vp, i64, vp, vp, sz, i32, vp],
    "vlsat_permute_rows": [vp, i64, vp, i64, i32, vp, i64, i32, vp, vp, vp],
    "vlsat_permute_edges": [vp, vp, i64, vp, vp],
    "vlsat_bf16_split": [vp, i64, i64, i64, vp, vp, i64, vp],
    "vlsat_flash_attn_bf16x3_fwd": [vp, vp, i64, vp, vp, i64, vp, vp, i64, vp, i64, vp, i64, i64, i32, i32, vp, sz, vp],
    "vlsat_flash_attn_bf16x3_workspace_bytes": [i64, i64, i32],
    "vlsat_bf16_split_t": [vp, i64, i64, i64, vp, vp, i64, vp, vp, i64, vp],
    "vlsat_flash_attn_bwd_stats": [vp, i64, vp, i64, vp, i64, vp, vp, i64, i64, i32, i32, vp],
    "vlsat_flash_attn_bf16x3_bwd": [C.POINTER(FlashBwdOperands), vp, i64, vp, i64, vp, i64, i64, i64, i32, i32, vp, sz, vp],
    "vlsat_flash_attn_bf16x3_bwd_workspace_bytes": [i64, i64, i32],
    "vlsat_build_csr": [vp, i64, i64, vp, vp, vp, sz, vp],
    "vlsat_gat_edge_fwd": [vp, i64, vp, i64, vp, i64, vp, vp, vp, vp, vp, vp, vp, i64, i64,
                           i32, i32, i32, i32, i32, i32, i32, vp, i64, vp, vp, vp],
    "vlsat_transpose": [vp, i64, i64, vp, i64, i64, i64, i64, i64, vp],
    "vlsat_act_bwd": [vp, i64, vp, i64, i32, f32, vp, vp, i64, vp, i64, i64, vp],
    "vlsat_act_bwd_pair": [vp, i64, vp, i64, i32, f32, vp, vp, i64, vp, vp, vp, i64, i64, i64, vp],
    "vlsat_wgrad_small": [vp, i64, vp, i64, i64, i32, i32, vp, i64, vp],
    "vlsat_gather_rows": [vp, i64, vp, i32, i64, i32, vp, i64, vp],
    "vlsat_scatter_add_rows": [vp, i64, vp, i32, i64, i32, vp, i64, vp],
    "vlsat_add_layernorm_bwd": [vp, i64, vp, i64, vp, i64, vp, vp, vp, i64, vp, vp, i64, i32, f32, i32, vp],
    "vlsat_gat_softmax_aggr_fwd": [vp, vp, i64, vp, vp, i64, i64, i32, i32, i32, vp, i64, vp, vp, vp],
    "vlsat_gat_softmax_aggr_bwd": [vp, i64, vp, vp, i64, vp, vp, vp, i64, i64, i32, i32, i32, vp, vp, i64, vp],
    "vlsat_attn_prob_bwd": [vp, vp, i64, vp, vp, f32, vp, vp, vp, i64, i64, i64, vp],
    "vlsat_attn_prob_bwd_pairs": [vp, vp, i64, vp, vp, f32, vp, vp, i64, vp, vp, vp, vp, i64, i64, i64, vp],
    "vlsat_rowdot_heads": [vp, i64, vp, i64, vp, i64, i32, i32, vp],
    "vlsat_pair_features": [vp, i64, vp, vp, vp, i64, vp, vp],
    "vlsat_node_attn_bias_fwd": [vp, i64, vp, i64, vp, i64, vp, vp, vp, vp, i32, i32, i32, vp, i64, i64, vp],
    "vlsat_node_attn_bias_bwd": [vp, i64, vp, i64, vp, i64, vp, vp, vp, vp, vp, i64, i32, i32, i32, vp, i64, vp, i64,
                                 vp, i64, vp, i64, vp],
    "vlsat_pointnet_pool_bwd": [vp, vp, vp, vp, i64, i64, i32, i32, vp, vp, vp],
    "vlsat_dropout": [vp, i64, vp, i64, i64, i64, f32, C.c_uint64, C.c_uint64, vp, vp],
    "vlsat_dropout_pair": [vp, i64, vp, i64, i64, i64, f32, C.c_uint64, C.c_uint64, vp, vp, vp, i64, vp],
    "vlsat_batchnorm_fwd": [vp, i64, vp, vp, vp, vp, vp, vp, f32, f32, i32, i32, vp, i64, i64, i64, vp],
    "vlsat_batchnorm_bwd": [vp, i64, vp, i64, vp, vp, vp, vp, i32, i32, vp, i64, vp, vp, i64, i64, vp],
    "vlsat_row_l2norm_bwd": [vp, vp, vp, i64, i32, vp],
    "vlsat_dot_accum": [vp, vp, i64, vp, vp],
    "vlsat_cross_entropy_fwd": [vp, i64, vp, i64, i32, vp, vp, vp],
    "vlsat_cross_entropy_bwd": [vp, i64, vp, vp, vp, f32, vp, i64, i64, i32, vp],
    "vlsat_rel_class_weights": [vp, i64, i32, f32, vp, vp],
    "vlsat_bce_fwd": [vp, vp, vp, i64, i32, vp, vp],
    "vlsat_bce_bwd": [vp, vp, vp, vp, f32, vp, i64, i32, vp],
    "vlsat_cosine_margin_fwd": [vp, i64, vp, i64, i64, i32, f32, vp, vp],
    "vlsat_cosine_margin_bwd": [vp, i64, vp, i64, vp, f32, f32, vp, i64, vp, i64, i64, i32, vp],
    "vlsat_l1_unit_fwd": [vp, i64, vp, i64, i64, i32, vp, vp],
    "vlsat_l1_unit_bwd": [vp, i64, vp, i64, vp, f32, vp, i64, i64, i32, vp],
    "vlsat_sum_rows": [vp, i64, f32, vp, vp, f32, i32, vp],
    "vlsat_object_prep_fwd": [vp, i64, i64, i32, vp, i64, i64, vp, vp, vp],
    "vlsat_softmax_rows": [vp, i64, i64, i32, vp, vp],
    "vlsat_topk_object_ranks": [vp, i64, vp, i64, i32, i32, vp, vp],
    "vlsat_topk_predicate_ranks": [vp, vp, i64, i32, i32, f32, vp, vp],
    "vlsat_topk_triplet_ranks": [vp, i64, i32, vp, i32, vp, vp, vp, i64, i32, f32, vp, vp],
    "vlsat_recall_at": [vp, i64, i32, i32, i32, vp, vp],
    "vlsat_adamw_step": [vp, vp, vp, i64, i32, C.c_double, C.c_double, f32, vp, i64, vp],
    "vlsat_rel_text_embed": [vp, i32, i32, i32, vp, vp, i64, vp, i64, vp, i64, vp],
    "vlsat_pack_scale": [vp, vp, vp, i64, i32, f32, vp],
}
_RESTYPES = {"vlsat_linear_workspace_bytes": sz, "vlsat_gemm_pairs_workspace_bytes": sz, "vlsat_flash_attn_bf16x3_workspace_bytes": sz, "vlsat_flash_attn_bf16x3_bwd_workspace_bytes": sz, "vlsat_error_string": C.c_char_p, "vlsat_gemm_engine": C.c_char_p, "vlsat_launch_count": i64}

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). vlsat_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.argtypes = args
        fn.restype = _RESTYPES.get(name, i32)
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().vlsat_error_string(status).decode()
        raise RuntimeError(f"{what}: vlsat status {status} ({msg})")
