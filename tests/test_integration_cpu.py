"""Drop-in wiring against the unmodified reference (runs only where /root/reference is mounted)."""
import pytest
import torch

from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not mounted")


def test_patch_level1_builds_from_unchanged_config():
    ref_import.import_reference()
    from slotvps_b200.integration import patch_reference
    import slotvps_b200 as sv
    import mmdet.models.detectors.vps_capsule as caps
    orig = caps.MultiScaleDynamicMaskHead
    try:
        ref_model, cfg = ref_import.build_model(0)
        ref_sd = ref_model.image_model.dynamic_mask_head.state_dict()
        patch_reference(level=1)
        model, _ = ref_import.build_model(0)
        head = model.image_model.dynamic_mask_head
        assert isinstance(head, sv.B200DynamicMaskHead)
        assert list(head.state_dict().keys()) == list(ref_sd.keys())
        head.load_state_dict(ref_sd, strict=True)                     # reference checkpoint loads unchanged
        assert head.per_dh_num_heads == [1, 2, 2, 2] and head.apply_temporal_query_atten_stages == [3, 4, 5, 6]
    finally:
        caps.MultiScaleDynamicMaskHead = orig


def test_patch_level2_repoints_detector():
    ref_import.import_reference()
    from slotvps_b200.integration import patch_reference
    import slotvps_b200 as sv
    import mmdet.models.detectors.vps_capsule as caps
    import mmdet.models.detectors.vps_temporal_slots as vts
    from mmdet.models.registry import DETECTORS
    saved = (caps.MultiScaleDynamicMaskHead, vts.PostProcessPanopticInstances, DETECTORS.module_dict["VPS_Temporal_Slots"])
    try:
        cls = patch_reference(level=2)
        model, _ = ref_import.build_model(0)
        assert type(model) is cls and isinstance(model, saved[2])
        assert isinstance(model.postprocess_panoptic, sv.PanopticFusion)
        assert isinstance(model.image_model.dynamic_mask_head, sv.B200DynamicMaskHead)
        with pytest.raises(RuntimeError, match="CUDA"):               # no CPU fallback
            model.generate_final_outputs([torch.zeros(1, 256, 4, 4)] * 4, torch.zeros(7, 1, 100, 256), generate_aux_output=False)
    finally:
        caps.MultiScaleDynamicMaskHead, vts.PostProcessPanopticInstances, DETECTORS.module_dict["VPS_Temporal_Slots"] = saved
