"""ctypes binding of libslotvps_b200.so (the C ABI declared in include/slotvps_b200.h).

The product path has NO CPU fallback: if the shared library is missing or a call fails, this
module raises.  ``build_library()`` compiles it in-tree with nvcc for sm_100a (the GPU box gets
the prebuilt ``.so`` with the repo snapshot).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libslotvps_b200.so")

MAX_LEVELS, MAX_STAGES, MAX_FRAMES = 4, 8, 8

STAGE_FIELDS = [
    "in_proj_w", "in_proj_b", "out_proj_w", "out_proj_b",
    "norm1_w", "norm1_b", "norm2_w", "norm2_b", "norm3_w", "norm3_b",
    "to_q_w", "to_q_b", "to_k_w", "to_k_b", "to_v_w", "to_v_b",
    "nq_w", "nq_b", "nk_w", "nk_b", "nv_w", "nv_b", "no_w", "no_b",
    "lin1_w", "lin1_b", "lin2_w", "lin2_b",
    "cls0_w", "cls0_nw", "cls0_nb", "cls1_w", "cls1_nw", "cls1_nb",
    "reg0_w", "reg0_nw", "reg0_nb", "reg1_w", "reg1_nw", "reg1_nb",
    "logit_w", "logit_b",
    "tq_to_q_w", "tq_to_q_b", "tq_to_k_w", "tq_to_k_b", "tq_to_v_w", "tq_to_v_b",
    "tq_nq_w", "tq_nq_b", "tq_nk_w", "tq_nk_b", "tq_nv_w", "tq_nv_b", "tq_no_w", "tq_no_b",
    "tq_lin1_w", "tq_lin1_b", "tq_lin2_w", "tq_lin2_b",
    "tq_norm2_w", "tq_norm2_b", "tq_norm3_w", "tq_norm3_b",
]

# state_dict suffix (relative to "head_series_{l}.{j}.") of every slotvps_stage_params field
STAGE_KEYS = {
    "in_proj_w": "self_attn.in_proj_weight", "in_proj_b": "self_attn.in_proj_bias",
    "out_proj_w": "self_attn.out_proj.weight", "out_proj_b": "self_attn.out_proj.bias",
    "norm1_w": "norm1.weight", "norm1_b": "norm1.bias", "norm2_w": "norm2.weight", "norm2_b": "norm2.bias",
    "norm3_w": "norm3.weight", "norm3_b": "norm3.bias",
    "to_q_w": "inst_interact.to_q.weight", "to_q_b": "inst_interact.to_q.bias",
    "to_k_w": "inst_interact.to_k.weight", "to_k_b": "inst_interact.to_k.bias",
    "to_v_w": "inst_interact.to_v.weight", "to_v_b": "inst_interact.to_v.bias",
    "nq_w": "inst_interact.norm_q.weight", "nq_b": "inst_interact.norm_q.bias",
    "nk_w": "inst_interact.norm_k.weight", "nk_b": "inst_interact.norm_k.bias",
    "nv_w": "inst_interact.norm_v.weight", "nv_b": "inst_interact.norm_v.bias",
    "no_w": "inst_interact.norm1.weight", "no_b": "inst_interact.norm1.bias",
    "lin1_w": "linear1.weight", "lin1_b": "linear1.bias", "lin2_w": "linear2.weight", "lin2_b": "linear2.bias",
    "cls0_w": "cls_module.0.weight", "cls0_nw": "cls_module.1.weight", "cls0_nb": "cls_module.1.bias",
    "cls1_w": "cls_module.3.weight", "cls1_nw": "cls_module.4.weight", "cls1_nb": "cls_module.4.bias",
    "reg0_w": "reg_module.0.weight", "reg0_nw": "reg_module.1.weight", "reg0_nb": "reg_module.1.bias",
    "reg1_w": "reg_module.3.weight", "reg1_nw": "reg_module.4.weight", "reg1_nb": "reg_module.4.bias",
    "logit_w": "class_logits.weight", "logit_b": "class_logits.bias",
}
_TQ = "temporal_query_head."
STAGE_KEYS.update({
    "tq_to_q_w": _TQ + "inst_interact.to_q.weight", "tq_to_q_b": _TQ + "inst_interact.to_q.bias",
    "tq_to_k_w": _TQ + "inst_interact.to_k.weight", "tq_to_k_b": _TQ + "inst_interact.to_k.bias",
    "tq_to_v_w": _TQ + "inst_interact.to_v.weight", "tq_to_v_b": _TQ + "inst_interact.to_v.bias",
    "tq_nq_w": _TQ + "inst_interact.norm_q.weight", "tq_nq_b": _TQ + "inst_interact.norm_q.bias",
    "tq_nk_w": _TQ + "inst_interact.norm_k.weight", "tq_nk_b": _TQ + "inst_interact.norm_k.bias",
    "tq_nv_w": _TQ + "inst_interact.norm_v.weight", "tq_nv_b": _TQ + "inst_interact.norm_v.bias",
    "tq_no_w": _TQ + "inst_interact.norm1.weight", "tq_no_b": _TQ + "inst_interact.norm1.bias",
    "tq_lin1_w": _TQ + "linear1.weight", "tq_lin1_b": _TQ + "linear1.bias",
    "tq_lin2_w": _TQ + "linear2.weight", "tq_lin2_b": _TQ + "linear2.bias",
    "tq_norm2_w": _TQ + "norm2.weight", "tq_norm2_b": _TQ + "norm2.bias",
    "tq_norm3_w": _TQ + "norm3.weight", "tq_norm3_b": _TQ + "norm3.bias",
})


class StageParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in STAGE_FIELDS]


class HeadDesc(C.Structure):
    _fields_ = [
        ("n_frames", C.c_int32), ("n_slots", C.c_int32), ("n_levels", C.c_int32),
        ("heads_per_level", C.c_int32 * MAX_LEVELS), ("h", C.c_int32 * MAX_LEVELS), ("w", C.c_int32 * MAX_LEVELS),
        ("num_classes", C.c_int32), ("dim_feedforward", C.c_int32), ("temporal_dim_feedforward", C.c_int32),
        ("nhead", C.c_int32), ("temporal_mask", C.c_int32), ("pos_mode", C.c_int32), ("kernel_path", C.c_int32),
        ("ffn_act", C.c_int32), ("temporal_ffn_act", C.c_int32),
    ]


class HeadOpts(C.Structure):
    _fields_ = [("stage_slots_in", C.POINTER(C.c_void_p)), ("skip_fused_mask", C.c_int32), ("feat_bn_scale", C.c_void_p),
                ("feat_bn_shift", C.c_void_p), ("rnorm_ss", C.c_void_p)]


class FusionCfg(C.Structure):
    _fields_ = [
        ("num_classes", C.c_int32), ("stuff_num", C.c_int32), ("small_area", C.c_int32), ("max_iters", C.c_int32),
        ("threshold", C.c_float), ("pixel_threshold", C.c_float), ("fraction_threshold", C.c_double),
        ("logits_width", C.c_int32), ("reserved", C.c_int32),
    ]


class DcnLayer(C.Structure):
    _fields_ = [("c_in", C.c_int32), ("c_out", C.c_int32), ("offset_w", C.c_void_p), ("offset_b", C.c_void_p),
                ("weight", C.c_void_p), ("gn_w", C.c_void_p), ("gn_b", C.c_void_p)]


# name -> (restype, argtypes): every symbol include/slotvps_b200.h declares
P = C.c_void_p
SYMBOLS = {
    "slotvps_head_workspace_bytes": (C.c_int, [C.POINTER(HeadDesc), C.POINTER(C.c_size_t)]),
    "slotvps_prepared_bytes": (C.c_int, [C.POINTER(HeadDesc), C.POINTER(C.c_size_t)]),
    "slotvps_prepare_weights": (C.c_int, [C.POINTER(HeadDesc), C.POINTER(StageParams), P, P, P, P]),
    "slotvps_prepare_weights_ex": (C.c_int, [C.POINTER(HeadDesc), C.POINTER(StageParams), P, P, P, P, P, P]),
    "slotvps_head_forward": (C.c_int, [C.POINTER(HeadDesc), C.POINTER(StageParams), P, C.POINTER(P), C.POINTER(P),
                                       C.POINTER(P), P, P, C.POINTER(P), P, C.c_size_t, P]),
    "slotvps_head_forward_ex": (C.c_int, [C.POINTER(HeadDesc), C.POINTER(StageParams), P, C.POINTER(P), C.POINTER(P),
                                          C.POINTER(P), P, P, C.POINTER(P), P, C.c_size_t, C.POINTER(HeadOpts), P]),
    "slotvps_head_mask_logits_ex": (C.c_int, [C.POINTER(HeadDesc), P, C.c_size_t, C.c_int, P, P, P, P, P, P, P, P, P, P, C.c_size_t, P]),
    "slotvps_fold_batchnorm": (C.c_int, [P, P, P, P, C.c_int, P, P, P]),
    "slotvps_level_fuse": (C.c_int, [P, P, P, P, P, C.c_int, C.c_int, P, P]),
    "slotvps_slot_attention": (C.c_int, [C.POINTER(StageParams), P, P, P, P, C.c_int, C.c_int, C.c_int, C.c_int, P,
                                         C.c_size_t, P]),
    "slotvps_slot_attention_workspace_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "slotvps_mask_logits_workspace_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "slotvps_mask_logits": (C.c_int, [P, P, P, P, P, P, P, P, C.c_int, C.c_int, C.c_int, P, C.c_size_t, P]),
    "slotvps_head_mask_logits": (C.c_int, [C.POINTER(HeadDesc), P, C.c_size_t, C.c_int, P, P, P, P, P, P, P, P, P, C.c_size_t, P]),
    "slotvps_fusion_workspace_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "slotvps_panoptic_fuse": (C.c_int, [C.POINTER(FusionCfg), P, P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, P, P,
                                        P, C.c_int, P, C.c_size_t, P]),
    "slotvps_panoptic_fuse_resume": (C.c_int, [C.POINTER(FusionCfg), P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, P, P,
                                               P, C.c_int, P, C.c_size_t, C.c_int, P]),
    "slotvps_sine_pos": (C.c_int, [P, C.c_int, C.c_int, P]),
    "slotvps_semantic_argmax": (C.c_int, [P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, P, P]),
    "slotvps_unify_workspace_bytes": (C.c_int, [C.POINTER(C.c_size_t)]),
    "slotvps_unify_reset": (C.c_int, [P, C.c_size_t, P]),
    "slotvps_unify_pan_result": (C.c_int, [P, P, P, C.c_int, P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, P, P, P, C.c_size_t, P]),
    "slotvps_track_scores": (C.c_int, [P, P, C.c_int, P, C.c_int, P, C.c_int, P, P, C.c_size_t, P]),
    "slotvps_track_state_bytes": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "slotvps_track_reset": (C.c_int, [P, C.c_size_t, P]),
    "slotvps_track_step": (C.c_int, [P, P, C.c_int, P, P, C.c_int, P, C.c_size_t, C.c_int, P, P]),
    "slotvps_deform_conv_workspace_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "slotvps_deform_conv_forward": (C.c_int, [P, P, P, P] + [C.c_int] * 16 + [P, C.c_size_t, P]),
    "slotvps_dcn_prepared_bytes": (C.c_int, [C.POINTER(DcnLayer), C.c_int, C.POINTER(C.c_size_t)]),
    "slotvps_dcn_prepare": (C.c_int, [C.POINTER(DcnLayer), C.c_int, P, C.c_size_t, P]),
    "slotvps_dcn_workspace_bytes": (C.c_int, [C.POINTER(DcnLayer), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "slotvps_dcn_subnet_forward": (C.c_int, [C.POINTER(DcnLayer), C.c_int, P, P, P, C.c_int, C.c_int, C.c_int, P, C.c_size_t, P]),
    "slotvps_last_error": (C.c_char_p, []),
    "slotvps_version": (C.c_char_p, []),
    "slotvps_launch_count": (C.c_int64, [C.c_int]),
    "slotvps_profile_begin": (C.c_int, [P]),
    "slotvps_profile_end": (C.c_int, [C.c_char_p, C.c_size_t]),
}

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))) + \
        [os.path.join(ROOT, "include", "slotvps_b200.h")]


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/capi.cu -> libslotvps_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    if not force and os.path.exists(LIB_PATH):
        newest = max(os.path.getmtime(s) for s in sources())
        if os.path.getmtime(LIB_PATH) >= newest:
            return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, os.path.join(CSRC, "capi.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB_PATH


_LIB: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """The loaded library; raises (never falls back) when it is absent."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(slotvps_b200 has no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)           # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _LIB = l
    return _LIB


class SlotVPSError(RuntimeError):
    pass


def check(code: int, what: str):
    if code != 0:
        msg = lib().slotvps_last_error().decode()
        if code == -1:
            raise ValueError(f"{what}: {msg}")
        raise SlotVPSError(f"{what} failed ({code}): {msg}")
