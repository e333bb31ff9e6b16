"""CPU-side checks of the C-ABI boundary: the library builds for sm_100a, loads, exports every
symbol include/slotvps_b200.h declares, validates arguments without touching a GPU, and the host
mirror keeps the reference's parameter names / constructor contract."""
import ctypes as C
import os
import re

import pytest
import torch

import slotvps_b200 as sv
from slotvps_b200 import _lib, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    _lib.build_library()
    return _lib.lib()


def test_header_symbols_exported(L):
    hdr = open(os.path.join(ROOT, "include", "slotvps_b200.h")).read()
    declared = set(re.findall(r"\b(slotvps_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(L, name)
    assert b"sm_100a" in L.slotvps_version()


def test_struct_layout_matches_header():
    hdr = open(os.path.join(ROOT, "include", "slotvps_b200.h")).read()
    body = hdr[hdr.index("typedef struct slotvps_stage_params {"):hdr.index("} slotvps_stage_params;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\*\s*([a-z_0-9]+)", body)
    assert fields == _lib.STAGE_FIELDS
    assert C.sizeof(_lib.StageParams) == 8 * len(fields)
    assert C.sizeof(_lib.HeadDesc) == 4 * (3 + 12 + 9)
    assert C.sizeof(_lib.HeadOpts) == 40
    assert C.sizeof(_lib.FusionCfg) == 40


def test_argument_validation_without_gpu(L):
    d = _lib.HeadDesc()
    n = C.c_size_t()
    assert L.slotvps_head_workspace_bytes(C.byref(d), C.byref(n)) == -1          # n_frames == 0
    assert b"n_frames" in L.slotvps_last_error()
    d.n_frames, d.n_slots, d.n_levels, d.nhead = 2, 100, 4, 8
    d.num_classes, d.dim_feedforward, d.temporal_dim_feedforward = 20, 2048, 1024
    for l, (hp, (h, w)) in enumerate(zip([1, 2, 2, 2], [(32, 64), (64, 128), (128, 256), (256, 512)])):
        d.heads_per_level[l], d.h[l], d.w[l] = hp, h, w
    assert L.slotvps_head_workspace_bytes(C.byref(d), C.byref(n)) == 0
    assert 1 << 20 < n.value < 8 << 30
    assert L.slotvps_prepared_bytes(C.byref(d), C.byref(n)) == 0 and n.value > 7 * 2 * 256 * 256 * 4
    d.w[2] = 250                                                                  # not 2x the previous level
    assert L.slotvps_head_workspace_bytes(C.byref(d), C.byref(n)) == -1
    d.w[2] = 256
    d.n_slots = 513
    assert L.slotvps_head_workspace_bytes(C.byref(d), C.byref(n)) == -1
    assert L.slotvps_fusion_workspace_bytes(100, 1024, 2048, C.byref(n)) == 0 and n.value > 2 * 2 * 1024 * 2048
    assert L.slotvps_fusion_workspace_bytes(0, 4, 4, C.byref(n)) == -1
    assert L.slotvps_launch_count(1) == 0


def test_head_mirrors_reference_contract():
    head = sv.B200DynamicMaskHead(**sv.HEAD_KWARGS)
    sd = synthetic.make_head_state_dict(3)
    assert head.load_state_dict(sd, strict=True).missing_keys == []
    own = head.state_dict()
    assert list(own.keys()) == list(sd.keys())                  # same names, same order
    assert all(own[k].shape == sd[k].shape for k in sd)
    assert sum(p.numel() for p in head.parameters()) == 15_495_052          # 15.50 M
    # temporal heads only on the levels whose first stage is listed (dynamic_mask_head.py:89)
    assert head.head_series_0[0].temporal_query_head is None and head.head_series_1[1].temporal_query_head is None
    assert head.head_series_2[0].temporal_query_head is not None and head.head_series_3[1].temporal_query_head is not None
    with pytest.raises(AssertionError):
        sv.B200DynamicMaskHead(**{**sv.HEAD_KWARGS, "dh_num_heads": 8})
    # both shipped configs construct: r50 (gelu / relu) and swinL (relu stage FFN, gelu temporal FFN; swinL_fpn_slotvps.py:41,56)
    swinl = sv.B200DynamicMaskHead(**{**sv.HEAD_KWARGS, "activation": "relu", "temporal_query_attention_config":
                                      {**sv.HEAD_KWARGS["temporal_query_attention_config"], "activation": "gelu"}})
    assert swinl._desc(2, 100, [(8, 16), (16, 32), (32, 64), (64, 128)], 2).ffn_act == 1
    assert swinl._desc(2, 100, [(8, 16), (16, 32), (32, 64), (64, 128)], 2).temporal_ffn_act == 2
    with pytest.raises(NotImplementedError):
        sv.B200DynamicMaskHead(**{**sv.HEAD_KWARGS, "activation": "glu"})
    # error conventions of the reference forward (dynamic_mask_head.py:167,193-195)
    f = [[torch.zeros(1, 128, 2, 2)] * 4]
    with pytest.raises(AssertionError):
        head(f, [torch.zeros(100, 256)], pad_mask=torch.zeros(1))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        head(f, [torch.zeros(100, 256)], None)


def test_fusion_rejects_unshipped_configs():
    with pytest.raises(NotImplementedError):
        sv.PanopticFusion(apply_mask_removal=False, apply_mask_removal_only_ins=True)
    with pytest.raises(NotImplementedError):
        sv.PanopticFusion()                                  # the reference's own defaults (both switches False) are not the shipped config
    with pytest.raises(NotImplementedError):
        sv.PanopticFusion(**{**sv.FUSION_KWARGS, "pixel_threshold": 0.3})
    with pytest.raises(NotImplementedError):
        sv.PanopticFusion(filter_small_option="4_256")
    sv.PanopticFusion(**sv.FUSION_KWARGS)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "slotvps_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|import_module\([\"']oracle", src, flags=re.M), \
                    f"{f} imports the oracle"


def test_next_row_mirrors_have_no_cpu_path():
    """Tracker / id-map consumer mirrors keep the reference's names and refuse CPU tensors (no fallback)."""
    import torch
    import slotvps_b200 as sv
    th = sv.B200TrackHead(**sv.TRACK_KWARGS)
    assert sorted(k for k, _ in th.named_parameters()) == ["fcs_query.0.bias", "fcs_query.0.weight", "fcs_query.1.bias", "fcs_query.1.weight"]
    with pytest.raises(RuntimeError):
        th(torch.zeros(3, 256), torch.zeros(4, 256))
    with pytest.raises(RuntimeError):
        sv.semantic_argmax(torch.zeros(1, 19, 8, 8), (8, 8))
    with pytest.raises(NotImplementedError):
        sv.B200TrackHead(num_fcs_query=2, in_channels_query=128)
    head = sv.B200DynamicMaskHead(**sv.HEAD_KWARGS)
    head.fold_input_transform(torch.eye(128).reshape(128, 128, 1, 1), torch.zeros(128))
    assert head._in_trans is not None and head._prepared is None
    head.fold_input_transform(None, None)
    assert head._in_trans is None
