"""The L2 drop-in run for real: the reference's own ``VPS_Temporal_Slots.simple_test``
(vps_temporal_slots.py:207-469, on the GPU, TF32 off) next to ``B200VPSTemporalSlots.simple_test``
(slotvps_b200/integration.py) over consecutive frames of synthetic videos, both built from the UNCHANGED
configs/cityscapes/r50_fpn_slotvps.py with identical weights; backbone / neck / semantic head are replaced by the same
tensor sources in both (they are out of scope), everything from semantic_trans_ins on is each side's own code.

Needs the reference staged under the git-ignored baseline/_ref/ (scripts/stage_reference.sh); skips otherwise.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "mmdet")), reason="reference not staged under baseline/_ref")]


class _Fn(torch.nn.Module):
    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, *a, **k):
        return self.fn(*a, **k)


def test_dropin_l2_matches_reference_simple_test():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    os.environ["SLOTVPS_REFERENCE_ROOT"] = REF
    import importlib
    import oracle.ref_import as ref_import             # checker-side plumbing: stubs for mmcv & co, imports baseline/_ref in place
    ref_import = importlib.reload(ref_import)          # (re-read SLOTVPS_REFERENCE_ROOT if an earlier test imported the module)
    import slotvps_b200 as sv
    from slotvps_b200 import synthetic
    from slotvps_b200.integration import patch_reference
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda:0")
    H, W, N = 256, 512, 100
    ref_model, _ = ref_import.build_model(0)           # the reference, unpatched
    patch_reference(level=2)
    b200_model, _ = ref_import.build_model(0)          # same config file, class symbols re-pointed (level 2)
    assert type(b200_model.image_model.dynamic_mask_head).__name__ == "B200DynamicMaskHead"
    # identical weights: random init of the reference + a confident classifier in the last stage (random-init heads keep
    # no slot above the 0.85 threshold)
    sd = ref_model.state_dict()
    k = "image_model.dynamic_mask_head.head_series_3.1.class_logits.weight"
    sd[k] = sd[k] * 20.0
    cap = synthetic.make_capsule_params(0, N)
    for name in ("feat_bn", "fg_bn"):
        for p in ("weight", "bias", "running_mean", "running_var"):
            sd[f"image_model.{name}.{p}"] = cap[f"{name}.{p}"].reshape(sd[f"image_model.{name}.{p}"].shape)
    ref_model.load_state_dict(sd, strict=True)
    missing = b200_model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    ref_model, b200_model = ref_model.to(dev).eval(), b200_model.to(dev).eval()
    videos, frames = 2, 4
    feats = {(v, f): [x.to(dev) for x in synthetic.make_features(H, W, T=1, video=40 + v, frame=f)[0]] for v in range(videos) for f in range(frames)}
    fcn = {(v, f): (torch.randn(1, 19, H, W, generator=torch.Generator().manual_seed(900 + 10 * v + f)) * 3).to(dev)
           for v in range(videos) for f in range(frames)}

    def wire(model):
        im = model.image_model
        im.backbone = _Fn(lambda x: x)
        im.neck = None
        model.extract_semantic_feats = lambda x: (fcn[(int(x[0, 0, 0, 0]), int(x[0, 0, 0, 1]))].clone(), None,
                                                  [t.clone() for t in feats[(int(x[0, 0, 0, 0]), int(x[0, 0, 0, 1]))]])
    wire(ref_model)
    wire(b200_model)

    def image(v, f):
        x = torch.zeros(1, 3, H, W, device=dev)
        x[0, 0, 0, 0], x[0, 0, 0, 1] = v, f
        return x
    # spy on the reference's post-processor: the kept masked logits give the decision margin of every pixel
    spy = {}
    pp = ref_model.postprocess_panoptic
    orig_forward = pp.forward

    def spy_forward(outputs, sizes, target_sizes=None, id=None):
        spy["pred_logits"], spy["pred_masks"] = outputs.pred_logits.detach().clone(), outputs.pred_masks.detach().clone()
        spy["embedding"] = outputs.output_embedding.detach().clone()
        res = orig_forward(outputs, sizes, target_sizes, id=id)
        spy["masks"] = res.masks.detach().float().cpu()
        spy["labels"] = [int(x) for x in res.labels]
        return res
    pp.forward = spy_forward
    bspy = {}
    bfuse = b200_model.postprocess_panoptic.fuse

    def spy_fuse(pred_logits, pred_masks, size, **kw):
        bspy["pred_logits"], bspy["pred_masks"] = pred_logits.detach().clone(), pred_masks.detach().clone()
        bspy["fo"] = bfuse(pred_logits, pred_masks, size, **kw)
        return bspy["fo"]
    b200_model.postprocess_panoptic.fuse = spy_fuse

    def kept(lg):
        sc, cl = torch.softmax(lg.float(), -1).max(-1)
        return set(torch.nonzero((cl != 19) & (sc > 0.85)).flatten().tolist()), sc, cl
    lines, worst, worst_hard, worst_same_in = [], 0.0, 0.0, 0.0
    # B200 post-processing chain driven with the REFERENCE's head outputs (identical inputs -> identical decisions)
    fz = sv.PanopticFusion(**sv.FUSION_KWARGS)
    th = sv.B200TrackHead(**sv.TRACK_KWARGS)
    th.load_state_dict(ref_model.temporal_track_head.state_dict(), strict=True)
    trk = sv.SlotTracker(th.to(dev), n_slots=N, capacity=1024, device=dev)
    for v in range(videos):
        for f in range(frames):
            meta = [dict(iid=(v + 1) * 10000 + f + 1, filename="synthetic", ori_shape=(H, W, 3), img_shape=(H, W, 3))]
            img, ref_img = image(v, f), image(v, max(f - 1, 0))
            with torch.no_grad():
                a = ref_model.simple_test(img, meta, rescale=True, ref_img=[ref_img])
                b = b200_model.simple_test(img, meta, rescale=True, ref_img=[ref_img])
            assert set(a.keys()) == set(b.keys())
            pa, pb = a["panoptic_outputs"].cpu().numpy(), b["panoptic_outputs"].cpu().numpy()
            assert pa.shape == pb.shape and pb.dtype == np.int64
            frac = float((pa != pb).mean())
            worst = max(worst, frac)
            # (i) identical inputs: the reference's own pred_logits / pred_masks / embeddings through the B200 fusion + tracker
            if f == 0:
                trk.reset()
            fo = fz.fuse(spy["pred_logits"], spy["pred_masks"], (H, W))
            hst = fo.host()
            rec = sv.SlotTracker.host(trk.step(spy["embedding"].float().contiguous(), fo))
            pc = fo.panoptic.cpu().numpy()[None]
            same_in = float((pa != pc).mean())
            worst_same_in = max(worst_same_in, same_in)
            np.testing.assert_array_equal(hst["cls_inds"], np.asarray(a["panoptic_cls_inds"].cpu()))
            np.testing.assert_array_equal(rec["det_obj_ids"], np.asarray(a["panoptic_det_obj_ids"].cpu()))
            # (ii) conditioning of this frame in the reference itself: smallest class-logit top-2 gap over the slots, against
            #      the head-output difference of the two implementations
            lg_ref = spy["pred_logits"].float()
            t2 = lg_ref.topk(2, dim=1).values
            gap = float((t2[:, 0] - t2[:, 1]).min())
            ka, sca, cla = kept(lg_ref)
            kb, scb, clb = kept(bspy["pred_logits"])
            dl = float((bspy["pred_logits"].float() - lg_ref).abs().max())
            dm = float((bspy["pred_masks"].float() - spy["pred_masks"].float()).norm() / spy["pred_masks"].float().norm())
            flips = sorted(ka ^ kb) + [i for i in sorted(ka & kb) if int(cla[i]) != int(clb[i])]
            # entries that survive mask_removal + the area <= 4 filter; the inline relabel looks stuff labels up by POSITION in
            # the list of ids present in the map (vps_temporal_slots.py:433), so one entry more or less shifts every later label
            # (the surviving SLOTS: from the same-input run for the reference side -- its outputs were just shown identical)
            la, lb = [int(x) for x in hst["keep"]], [int(x) for x in bspy["fo"].host()["keep"]]
            assert spy["labels"] == [int(x) for x in hst["labels"]]
            same_list = la == lb
            # decision margin of the reference at every pixel: top-1 minus top-2 of its kept (masked) logits
            mk = spy["masks"]
            top2 = mk.topk(2, dim=0).values if mk.shape[0] >= 2 else torch.stack([mk[0], mk[0] - 1e9])
            margin = (top2[0] - top2[1]).numpy()
            scale = float(mk.abs().max())
            diff = (pa != pb)[0]
            near = margin < 1e-2 * scale                           # within 10x the stage-6 drift (~1e-3 relative) of the logit scale
            hard = float((diff & ~near).mean())
            worst_hard = max(worst_hard, hard)
            mmax = float(margin[diff].max() / scale) if diff.any() else 0.0
            np.testing.assert_array_equal(np.asarray(a["panoptic_cls_inds"].cpu()), np.asarray(b["panoptic_cls_inds"].cpu()))
            np.testing.assert_array_equal(np.asarray(a["panoptic_det_obj_ids"].cpu()), np.asarray(b["panoptic_det_obj_ids"].cpu()))
            np.testing.assert_allclose(np.asarray(a["panoptic_cls_prob"].cpu()), np.asarray(b["panoptic_cls_prob"].cpu()), rtol=1e-4)
            sa, sb = a["fcn_outputs"].cpu().numpy(), b["fcn_outputs"].cpu().numpy()
            assert float((sa != sb).mean()) < 1e-5
            assert set(np.unique(pa).tolist()) == set(np.unique(pb).tolist())
            lines.append(f"video {v} frame {f}: same-input fusion+tracker vs reference: id-map mismatch {100 * same_in:.4f} %; head outputs B200 vs reference: "
                         f"class logits max abs diff {dl:.1e} (smallest top-2 gap {gap:.1e}), mask logits rel {dm:.1e}, kept slots {len(ka)} vs {len(kb)}, "
                         f"slots whose keep / class decision differs {flips}, surviving slots {la} vs {lb} ({'same list' if same_list else 'lists differ'}) | end to end: "
                         f"ids {sorted(np.unique(pa).tolist())} things {len(a['panoptic_cls_inds'])} "
                         f"id-map mismatch {100 * frac:.4f} % of {pa.size} px, outside near-tie pixels {100 * hard:.4f} % "
                         f"(near-tie pixels {100 * float(near.mean()):.1f} %, largest relative margin at a mismatch {mmax:.1e}); "
                         f"semantic mismatches {int((sa != sb).sum())}")
            print(lines[-1])
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "r2_dropin_L2.txt"), "w") as fh:
            fh.write("reference simple_test (GPU, TF32 off) vs B200VPSTemporalSlots.simple_test, configs/cityscapes/r50_fpn_slotvps.py, "
                     f"{videos} videos x {frames} frames {H}x{W}\n" + "\n".join(lines) + f"\nworst id-map mismatch fraction {worst:.2e}; outside near-tie pixels {worst_hard:.2e}; same-input (reference head outputs -> B200 fusion + tracker) worst mismatch {worst_same_in:.2e}\n")
    # both sides carry their own fp32 drift through 7 stages (~1e-3 on the mask logits): region boundaries may move by a pixel
    # random-init slots collapse: many kept slots carry almost the same mask, so large regions are decided
    # by margins below the drift -- mismatches must be confined to those near-tie pixels
    # -- what is asserted -------------------------------------------------------------------------------------------
    # identical inputs -> identical outputs: the B200 fusion / tracker chain on the reference's own head outputs
    assert worst_same_in < 1e-4, worst_same_in
    # end to end both detectors return the same keys, label sets, thing classes and object ids (asserted per frame above).
    # The pixel maps agree outside near-tie pixels on well-conditioned frames; at random init the reference's own class
    # decisions have top-2 gaps (~5e-3) below its fp32-vs-fp64 logit drift (~4e-2), and ~97 of 100 collapsed slots are kept, so
    # (a) a slot whose keep / class decision flips between the implementations moves its whole region, and (b) one entry more or
    # less surviving the area <= 4 filter shifts every later stuff label through the reference's position-indexed lookup (:433),
    # and (c) with all scores saturated at ~1.0 the score ORDER (np.argsort, :581) that picks which duplicate slot represents a
    # stuff class is decided by the last bits.
    # Frames with either event are reported with the cause; all other frames must agree outside near-tie pixels.
    n_clean = 0
    for l in lines:                      # same slot decisions and same surviving list -> the maps must agree outside near-tie pixels
        if "decision differs []" in l and "(same list)" in l:
            n_clean += 1
            assert float(l.split("outside near-tie pixels ")[1].split(" %")[0]) < 0.5, l
    assert n_clean >= 3, lines


def test_upsnet_forward_with_b200_subnet_matches_reference():
    """SURVEY 8f rank 4 as a drop-in: the reference's own ``UPSNetFPN.forward`` (upsnetFPN.py:64-85) built from the unchanged
    config runs UNMODIFIED on the GPU -- its DeformConv modules call the reference's compiled op (oracle/_ref, in place of the
    extension the reference's setup.py would have built) -- and again after ``patch_upsnet_subnet`` swapped the deformable-conv
    subnet for the B200 kernels; fcn_output, fcn_score and the four feature levels the head consumes must agree."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle import ref_dcn                              # checker: the reference op
    if not ref_dcn.available():
        pytest.skip("oracle/_ref/deform_conv_cuda.so not built")
    os.environ["SLOTVPS_REFERENCE_ROOT"] = REF
    import importlib
    import oracle.ref_import as ref_import
    ref_import = importlib.reload(ref_import)
    from slotvps_b200 import synthetic
    from slotvps_b200.dcn import B200DeformSubnet
    from slotvps_b200.integration import patch_upsnet_subnet
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda:0")
    model, _ = ref_import.build_model(0)
    dc_mod = importlib.import_module("mmdet.ops.dcn.deform_conv")
    dc_mod.deform_conv_cuda = ref_dcn.module()              # the compiled extension deform_conv.py imports (stubbed by ref_import)
    fpn = model.image_model.panopticFPN
    fpn.deform_convs[0].load_state_dict(synthetic.make_dcn_state_dict(7, fpn.in_channels, fpn.out_channels, offset_scale=1.0), strict=True)
    fpn = fpn.to(dev).eval()
    H, W = 256, 512
    inputs = [synthetic.make_fpn_level(30 + l, 1, fpn.in_channels, H // (4 << l), W // (4 << l)).to(dev) for l in range(fpn.num_levels)]
    with torch.no_grad():
        out_r, score_r, feats_r = fpn([x.clone() for x in inputs])
    net = patch_upsnet_subnet(model)
    assert isinstance(fpn.deform_convs[0], B200DeformSubnet) and fpn.deform_convs[0] is net
    with torch.no_grad():
        out_b, score_b, feats_b = fpn([x.clone() for x in inputs])

    def rel(a, b):
        a, b = a.double(), b.double()
        return float((a - b).norm() / b.norm())
    errs = [rel(out_b, out_r), rel(score_b, score_r)] + [rel(a, b) for a, b in zip(feats_b, feats_r)]
    print("UPSNetFPN.forward, reference op vs patch_upsnet_subnet: fcn_output rel %.2e, fcn_score rel %.2e, feature levels %s"
          % (errs[0], errs[1], " ".join("%.2e" % e for e in errs[2:])))
    assert out_b.shape == out_r.shape == (1, fpn.num_classes, H, W) and len(feats_b) == len(feats_r) == 4
    assert max(errs) < 5e-5
