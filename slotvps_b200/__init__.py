"""slotvps_b200 -- B200-native (sm_100a) implementation of the Slot-VPS retriever hot path.

Public surface (mirrors of the reference interfaces):
  B200DynamicMaskHead   <- MultiScaleDynamicMaskHead   (dynamic_mask_head.py:36)
  mask_logits           <- generate_final_outputs       (vps_temporal_slots.py:144)
  PanopticFusion        <- PostProcessPanopticInstances (vps_temporal_slots.py:528) + inline fusion (:411-435)
  SlotVPSRetriever      <- the three chained for one clip
  B200DeformSubnet      <- UPSNetFPN.deform_convs[0]    (upsnetFPN.py:36-49), deform_conv <- mmdet.ops.dcn.deform_conv
All compute is hand-written CUDA behind the C ABI of include/slotvps_b200.h; there is no CPU path.
"""
from ._lib import build_library, lib, SlotVPSError  # noqa: F401
from .head import B200DynamicMaskHead  # noqa: F401
from .retriever import (PanopticFusion, SlotVPSRetriever, FusionOutput, GraphedClip, mask_logits, level_fuse,  # noqa: F401
                        slot_attention, sine_position_embedding)
from .tracker import B200TrackHead, SlotTracker  # noqa: F401
from .unify import PanUnifier, get_unified_pan_result, semantic_argmax  # noqa: F401
from .dcn import B200DeformConvWithOffset, B200DeformSubnet, deform_conv, deform_subnet_levels  # noqa: F401

TRACK_KWARGS = dict(num_fcs_query=2, in_channels_query=256, query_matched_weight=1.0)  # r50_fpn_slotvps.py:90-96

HEAD_KWARGS = dict(  # configs/cityscapes/r50_fpn_slotvps.py:27-54
    dh_dim=256, num_classes=20, dim_feedforward=2048, nhead=8, dropout=0.0, activation="gelu", dh_num_heads=7,
    per_dh_num_heads=[1, 2, 2, 2], feat_num_levels=4, merge_operation="concat", trans_in_dim=384,
    return_intermediate=True, use_focal=True, prior_prob=0.01, num_cls=2, num_reg=2, drop_path=0.,
    temporal_query_attention_config=dict(d_model=256, dim_feedforward=1024, dropout=0.0, activation="relu",
                                         softmax_dim="slots", drop_path=0.),
    apply_temporal_query_atten_stages=[3, 4, 5, 6],
)
FUSION_KWARGS = dict(  # configs/cityscapes/r50_fpn_slotvps.py:66-74
    is_thing_map={i: i > 10 for i in range(20)}, threshold=0.85, fraction_threshold=0.03, pixel_threshold=0.4,
    apply_mask_removal=True, apply_mask_removal_only_ins=True, use_mask_low_constant=False,
)
