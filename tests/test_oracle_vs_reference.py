"""Live pinning of the oracle against the UNMODIFIED reference, imported in place from /root/reference by
oracle/ref_import.py (build container only: the GPU box has no reference tree, so everything here skips there; the
committed fixtures of tests/golden carry the same information to it).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import ref_import
from oracle import slotvps_oracle as O
from slotvps_b200 import synthetic

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


@pytest.fixture(scope="module")
def ref_model():
    torch.set_num_threads(8)
    model, _ = ref_import.build_model(0)
    return model


def _ref_pos(model, feat):
    from mmdet.core.utils.misc import nested_tensor_from_tensor_list
    return model.image_model.position_embedding(nested_tensor_from_tensor_list(feat))


def test_stage_by_stage_teacher_forced(ref_model):
    """Every stage of the oracle on the reference's OWN stage inputs (dynamic_mask_head.py:199-212) reproduces the
    reference's stage outputs to fp32 re-association noise, and the mask logits computed from the reference's own
    fused feature / embedding (vps_temporal_slots.py:144-154) agree tightly."""
    T, N, shapes, seed = 2, 100, [(3, 5), (6, 10), (12, 20), (24, 40)], 21
    head = ref_model.image_model.dynamic_mask_head
    sd = synthetic.make_head_state_dict(seed)
    head.load_state_dict(sd, strict=True)
    cap = synthetic.make_capsule_params(seed, N)
    feats = synthetic.make_features(0, 0, T=T, video=seed, frame=0, shapes=shapes)
    pos = [[_ref_pos(ref_model, f) for f in feats[t]] for t in range(T)]
    q = cap["init_mask_query.weight"]
    with torch.no_grad():
        cls, emb, fused = head(features=[list(f) for f in feats], init_masks=[q.clone() for _ in range(T)], pad_mask=None,
                               pos=pos, query_pos=None, gt_non_void_mask=None)
    forced = [[q] * T] + [[emb[t][s, 0] for t in range(T)] for s in range(6)]
    opos = [[O.sine_position_embedding(*s) for s in shapes] for _ in range(T)]
    ocls, oemb, ofused = O.head_forward(sd, feats, [q] * T, opos, stage_slots_in=forced)
    for t in range(T):
        for l in range(4):
            assert rel_l2(ofused[t][l].numpy(), fused[t][l].numpy()) < 2e-6
        for s in range(7):
            assert rel_l2(oemb[t][s].numpy(), emb[t][s].numpy()) < 2e-5, (t, s)
            assert rel_l2(ocls[t][s].numpy(), cls[t][s].numpy()) < 2e-5, (t, s)
    im = ref_model.image_model
    for k in ("weight", "bias", "running_mean", "running_var"):
        getattr(im.feat_bn, k).data.copy_(cap["feat_bn." + k])
        getattr(im.fg_bn, k).data.copy_(cap["fg_bn." + k])
    with torch.no_grad():
        _, ref_pm, _ = ref_model.generate_final_outputs([f.clone() for f in fused[-1]], emb[-1], generate_aux_output=False)
    pm = O.mask_logits(fused[-1][-1][0], emb[-1][-1, 0], cap)
    assert rel_l2(pm.numpy(), ref_pm[0].numpy()) < 2e-6


def test_upsnet_subnet_wiring_against_live_reference(ref_model):
    """SURVEY 8f rank 4 against the reference's own UPSNetFPN instance (upsnetFPN.py:36-49): the mirror loads the reference
    subnet's state_dict strictly, `patch_upsnet_subnet` swaps it in place, and everything of the chain EXCEPT the CUDA-only
    deformable op -- the offset convolutions, GroupNorm(32) eps / grouping, ReLU -- is compared module by module with the oracle
    (the op itself is pinned by tests/golden/dcn_*.npz, frozen from the reference's compiled op on the B200)."""
    from slotvps_b200.dcn import B200DeformSubnet
    from slotvps_b200.integration import patch_upsnet_subnet
    fpn = ref_model.image_model.panopticFPN
    seq = fpn.deform_convs[0]
    sd = synthetic.make_dcn_state_dict(3, fpn.in_channels, fpn.out_channels)
    assert list(seq.state_dict().keys()) == list(sd.keys())
    seq.load_state_dict(sd, strict=True)                                    # the synthetic parameters fit the reference module
    mirror = B200DeformSubnet(fpn.in_channels, fpn.out_channels)
    mirror.load_state_dict(seq.state_dict(), strict=True)
    for (k, a), (k2, b) in zip(mirror.state_dict().items(), seq.state_dict().items()):
        assert k == k2 and torch.equal(a, b)
    x = synthetic.make_fpn_level(3, 1, fpn.in_channels, 9, 11)
    cur = x
    for i in range(3):
        dc, gn, act = seq[3 * i], seq[3 * i + 1], seq[3 * i + 2]
        with torch.no_grad():
            off_ref = dc.conv_offset(cur)                                    # deform_conv_with_offset.py:38-39
        off = torch.nn.functional.conv2d(cur, sd[f"{3 * i}.conv_offset.weight"], sd[f"{3 * i}.conv_offset.bias"], padding=1)
        assert rel_l2(off.numpy(), off_ref.numpy()) < 1e-6
        y = O.deform_conv(cur, off_ref, sd[f"{3 * i}.conv.weight"])
        with torch.no_grad():
            want = act(gn(y.clone()))                                       # the reference's own GroupNorm / ReLU modules
        one = {"0.conv_offset.weight": sd[f"{3 * i}.conv_offset.weight"], "0.conv_offset.bias": sd[f"{3 * i}.conv_offset.bias"],
               "0.conv.weight": sd[f"{3 * i}.conv.weight"], "1.weight": sd[f"{3 * i + 1}.weight"], "1.bias": sd[f"{3 * i + 1}.bias"]}
        got = O.dcn_subnet(one, cur, n_layers=1)
        assert rel_l2(got.numpy(), want.numpy()) < 2e-6, i
        cur = want
    net = patch_upsnet_subnet(ref_model)
    assert fpn.deform_convs[0] is net and isinstance(net, B200DeformSubnet) and not net.training
    with pytest.raises(Exception):                                           # the mirror has no CPU path
        fpn.deform_convs[0](x)
