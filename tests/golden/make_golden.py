"""Generate the committed golden fixtures from the UNMODIFIED reference (run in the build container).

    python tests/golden/make_golden.py

The reference ships no tests / golden vectors for the retriever path, so
the oracle is pinned against the reference's own modules, imported in place from
/root/reference by oracle/ref_import.py (nothing of the reference is copied).  Inputs and weights
are regenerated from seeds by slotvps_b200.synthetic, so only OUTPUTS are stored:

  head_*.npz     MultiScaleDynamicMaskHead.forward  (dynamic_mask_head.py:138) cls / emb of all
                 7 stages for every frame + a strided sample of the fused features
  pos_*.npz      PositionEmbeddingSine              (position_encoding.py:236)
  masklogit.npz  generate_final_outputs             (vps_temporal_slots.py:144)
  fusion_*.npz   PostProcessPanopticInstances.forward (:659) + the inline fusion of simple_test
                 (:411-435), obtained by running the reference's simple_test with its
                 backbone / semantic head / retriever head replaced by tensor sources.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_import  # noqa: E402
from slotvps_b200 import synthetic  # noqa: E402

HEAD_CASES = {
    # name: (T, n_slots, level shapes coarse->fine, weight seed, per_dh_num_heads, temporal stages)
    "head_t2_n100": dict(T=2, N=100, shapes=[(2, 4), (4, 8), (8, 16), (16, 32)], seed=0),
    "head_t1_n50": dict(T=1, N=50, shapes=[(3, 5), (6, 10), (12, 20), (24, 40)], seed=1),
    "head_t3_n128": dict(T=3, N=128, shapes=[(2, 3), (4, 6), (8, 12), (16, 24)], seed=2),
    # every level >= 128 pixels: the tensor-core kernels serve ALL levels, and the per-stage outputs double as
    # teacher-forcing inputs (stage s enters with the reference's stage s-1 embedding, dynamic_mask_head.py:210-211)
    "head_t2_n100_big": dict(T=2, N=100, shapes=[(8, 16), (16, 32), (32, 64), (64, 128)], seed=4),
    # the reference's second shipped config (configs/cityscapes/swinL_fpn_slotvps.py:41,56): ReLU stage FFN, GELU temporal FFN
    "head_swinl": dict(T=2, N=100, shapes=[(8, 16), (16, 32), (32, 64), (64, 128)], seed=5,
                       overrides={"dynamic_mask_head.activation": "relu",
                                  "dynamic_mask_head.temporal_query_attention_config.activation": "gelu"}),
}
POS_SHAPES = [(2, 4), (16, 32), (34, 60), (7, 5)]
FUSION_CASES = {
    "fusion_a": dict(seed=0, N=100, h=24, w=40, n_things=12, dup_stuff=2, near_dup_things=3, tiny=2),
    "fusion_b": dict(seed=1, N=100, h=16, w=32, n_things=5, dup_stuff=0, near_dup_things=1, tiny=0),
    "fusion_c": dict(seed=2, N=60, h=32, w=24, n_things=20, dup_stuff=4, near_dup_things=6, tiny=3),
    "fusion_d": dict(seed=3, N=100, h=20, w=20, n_things=0, dup_stuff=3, near_dup_things=0, tiny=1),
    "fusion_e": dict(seed=4, N=100, h=12, w=28, n_things=30, dup_stuff=1, near_dup_things=8, tiny=4),
}


def ref_pos(model, feat):
    from mmdet.core.utils.misc import nested_tensor_from_tensor_list
    return model.image_model.position_embedding(nested_tensor_from_tensor_list(feat))


def gen_head(model, name, T, N, shapes, seed, overrides=None):
    head = model.image_model.dynamic_mask_head
    head.load_state_dict(synthetic.make_head_state_dict(seed), strict=True)
    cap = synthetic.make_capsule_params(seed, N)
    feats = synthetic.make_features(0, 0, T=T, video=seed, frame=0, shapes=shapes)
    pos = [[ref_pos(model, f) for f in feats[t]] for t in range(T)]
    q = cap["init_mask_query.weight"]
    with torch.no_grad():
        cls, emb, fused = head(features=[list(f) for f in feats], init_masks=[q.clone() for _ in range(T)],
                               pad_mask=None, pos=pos, query_pos=None, gt_non_void_mask=None)
    out = {}
    for t in range(T):
        out[f"cls{t}"] = cls[t].numpy()
        out[f"emb{t}"] = emb[t].numpy()
        for l in range(4):
            f = fused[t][l][0]
            out[f"fused{t}_{l}_sample"] = f[::7, ::3, ::5].contiguous().numpy()
            out[f"fused{t}_{l}_sum"] = np.float64(f.double().sum().item())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    return fused, emb, cap


def gen_masklogit(model, fused, emb, cap):
    im = model.image_model
    for k in ("weight", "bias", "running_mean", "running_var"):
        getattr(im.feat_bn, k).data.copy_(cap["feat_bn." + k])
        getattr(im.fg_bn, k).data.copy_(cap["fg_bn." + k])
    with torch.no_grad():
        _, mask_output, _ = model.generate_final_outputs([f.clone() for f in fused[-1]], emb[-1], generate_aux_output=False)
    np.savez_compressed(os.path.join(HERE, "masklogit.npz"), pred_masks=mask_output[0].numpy())


class _Fn(torch.nn.Module):
    """nn.Module shell so a tensor source can replace a child module of the reference model."""

    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, *a, **k):
        return self.fn(*a, **k)


def gen_fusion(model, name, seed, N, h, w, **kw):
    """Drive the reference's simple_test with tensor sources in place of backbone/head."""
    logits, masks, _ = synthetic.make_fusion_case(seed, N, h, w, **kw)
    H, W = 4 * h, 4 * w
    im = model.image_model
    im.backbone = _Fn(lambda x: x)
    im.neck = None
    model.extract_semantic_feats = lambda x: (torch.zeros(1, 19, H, W), None, [torch.zeros(1, 128, 1, 1)] * 4)
    model.semantic_trans_ins = lambda f: f
    model.generate_position_embedding = lambda f: None
    emb = torch.zeros(7, 1, N, 256)
    cls = logits[None, None].repeat(7, 1, 1, 1)
    im.dynamic_mask_head = _Fn(lambda **k: ([cls, cls], [emb, emb], [[None] * 4, [None] * 4]))
    model.generate_final_outputs = lambda feats, om, generate_aux_output=False: (feats, masks[None], [])
    im.init_mask_query = torch.nn.Embedding(N, 256)
    captured = {}
    pp = model.postprocess_panoptic
    orig_forward = pp.forward

    def spy(outputs, sizes, target_sizes=None, id=None):
        res = orig_forward(outputs, sizes, target_sizes, id=id)
        captured["masks"] = res.masks.clone()
        captured["labels"] = res.labels.clone()
        captured["probs"] = res.probs.clone()
        return res
    pp.forward = spy
    torch.Tensor.cuda = lambda self, *a, **k: self
    img = torch.zeros(1, 3, H, W)
    meta = [dict(iid=10001, filename="synthetic", ori_shape=(H, W, 3), img_shape=(H, W, 3))]
    with torch.no_grad():
        res = model.simple_test(img, meta, rescale=True, ref_img=[img])
    pp.forward = orig_forward
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        panoptic=res["panoptic_outputs"][0].numpy().astype(np.int64),
        cls_inds=res["panoptic_cls_inds"].numpy(), cls_prob=res["panoptic_cls_prob"].numpy(),
        labels=captured["labels"].numpy(), probs=captured["probs"].numpy(),
        masks_sum=captured["masks"].double().sum((1, 2)).numpy(),
        masks_nnz=(captured["masks"] != 0).sum((1, 2)).numpy())


def gen_head_intrans(model, name="head_intrans", seed=3, T=2, N=100, shapes=((2, 4), (4, 8), (8, 16), (16, 32))):
    """semantic_trans_ins (VPS_Capsule.conv_trans on every level, vps_temporal_slots.py:129-135) followed by the head,
    from UN-transformed features: the golden for the folded input transform."""
    im = model.image_model
    head = im.dynamic_mask_head
    head.load_state_dict(synthetic.make_head_state_dict(seed), strict=True)
    tp = synthetic.make_in_trans_params(seed)
    im.conv_trans.conv.weight.data.copy_(tp["conv_trans.conv.weight"])
    im.conv_trans.conv.bias.data.copy_(tp["conv_trans.conv.bias"])
    cap = synthetic.make_capsule_params(seed, N)
    raw = synthetic.make_features(0, 0, T=T, video=seed, frame=0, shapes=list(shapes))
    q = cap["init_mask_query.weight"]
    with torch.no_grad():
        feats = [model.semantic_trans_ins(list(fr)) for fr in raw]
        pos = [[ref_pos(model, f) for f in feats[t]] for t in range(T)]
        cls, emb, fused = head(features=feats, init_masks=[q.clone() for _ in range(T)], pad_mask=None, pos=pos,
                               query_pos=None, gt_non_void_mask=None)
    out = {}
    for t in range(T):
        out[f"cls{t}"] = cls[t].numpy()
        out[f"emb{t}"] = emb[t].numpy()
        for l in range(4):
            out[f"fused{t}_{l}_sample"] = fused[t][l][0][::7, ::3, ::5].contiguous().numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


UNIFY_CASES = {
    # name: frames (seeds), size, stuff_area_limit, with object ids, frames whose cls_inds are emptied (the 255 path)
    "unify_a": dict(seeds=[0, 1, 2, 3], H=128, W=256, limit=4096, with_obj=True, empty=[]),
    "unify_b": dict(seeds=[4, 5, 6], H=96, W=160, limit=4 * 64 * 64, with_obj=True, empty=[1]),
    "unify_c": dict(seeds=[7, 8], H=64, W=64, limit=512, with_obj=False, empty=[]),
}
SEMANTIC_CASES = {"semantic_same": dict(seed=0, h=64, w=128, H=64, W=128), "semantic_up4": dict(seed=1, h=32, w=64, H=128, W=256)}


def unify_inputs(seeds, H, W, with_obj, empty, **_):
    segs, pans, cis, ois = [], [], [], []
    for j, sd in enumerate(seeds):
        seg, pan, ci, oi = synthetic.make_unify_case(sd, H, W)
        segs.append(seg.numpy()); pans.append(pan.numpy())
        cis.append(np.zeros((0,), np.int64) if j in empty else ci.numpy())
        ois.append(oi.numpy().astype(np.int32))
    return segs, pans, cis, (ois if with_obj else None)


def gen_unify(name, c):
    """CityscapesVps.get_unified_pan_result (tools/dataset/cityscapes_vps.py:214) on synthetic frames."""
    import contextlib
    import io
    ds = object.__new__(ref_import.import_reference_tools())
    segs, pans, cis, ois = unify_inputs(**c)
    names = ["f%d" % i for i in range(len(segs))]
    with contextlib.redirect_stdout(io.StringIO()):            # the method prints its whole input
        res = ds.get_unified_pan_result([x.copy() for x in segs], [x.copy() for x in pans], cis,
                                        obj_ids=None if ois is None else [x.copy() for x in ois],
                                        stuff_area_limit=c["limit"], names=names)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **{n: res[n] for n in names})


def semantic_input(seed, h, w, **_):
    g = torch.Generator().manual_seed(90_000 + seed)
    coarse = torch.randn((1, 19, max(2, h // 8), max(2, w // 8)), generator=g)
    x = torch.nn.functional.interpolate(coarse, size=(h, w), mode="bilinear", align_corners=False) * 4.0
    return x + torch.randn((1, 19, h, w), generator=g) * 0.5


def gen_semantic(model, name, c):
    """The semantic argmax of simple_test (vps_temporal_slots.py:440-451): drive simple_test with a tensor source for
    the semantic head and keep `fcn_outputs`."""
    logits, masks, _ = synthetic.make_fusion_case(0, 100, c["H"] // 4, c["W"] // 4)
    fcn = semantic_input(**c)
    H, W = c["H"], c["W"]
    im = model.image_model
    im.backbone = _Fn(lambda x: x)
    im.neck = None
    model.extract_semantic_feats = lambda x: (fcn.clone(), None, [torch.zeros(1, 128, 1, 1)] * 4)
    model.semantic_trans_ins = lambda f: f
    model.generate_position_embedding = lambda f: None
    emb = torch.zeros(7, 1, 100, 256)
    cls = logits[None, None].repeat(7, 1, 1, 1)
    im.dynamic_mask_head = _Fn(lambda **k: ([cls, cls], [emb, emb], [[None] * 4, [None] * 4]))
    model.generate_final_outputs = lambda feats, om, generate_aux_output=False: (feats, masks[None], [])
    im.init_mask_query = torch.nn.Embedding(100, 256)
    torch.Tensor.cuda = lambda self, *a, **k: self
    img = torch.zeros(1, 3, H, W)
    meta = [dict(iid=10001, filename="synthetic", ori_shape=(H, W, 3), img_shape=(H, W, 3))]
    with torch.no_grad():
        res = model.simple_test(img, meta, rescale=True, ref_img=[img])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), fcn_outputs=res["fcn_outputs"].numpy().astype(np.uint8))


TRACK_CASES = {
    # name: seed, N, (h, w), [frames per video], FC mode
    "track_a": dict(seed=0, N=100, h=32, w=64, videos=[5, 3], mode="identity"),
    "track_b": dict(seed=1, N=100, h=32, w=64, videos=[4], mode="random"),
    "track_c": dict(seed=2, N=50, h=24, w=40, videos=[4, 2], mode="identity"),
}


def gen_track(model, name, seed, N, h, w, videos, mode):
    """Drive the reference's simple_test over consecutive frames (iid = vid*10000 + fid, fid==1 resets the
    tracker, vps_temporal_slots.py:218-237) with designed logits / masks / slot embeddings per frame, and record
    panoptic_det_obj_ids and the object bank after every frame."""
    H, W = 4 * h, 4 * w
    im = model.image_model
    im.backbone = _Fn(lambda x: x)
    im.neck = None
    model.extract_semantic_feats = lambda x: (torch.zeros(1, 19, H, W), None, [torch.zeros(1, 128, 1, 1)] * 4)
    model.semantic_trans_ins = lambda f: f
    model.generate_position_embedding = lambda f: None
    im.init_mask_query = torch.nn.Embedding(N, 256)
    model.temporal_track_head.load_state_dict(synthetic.make_track_params(seed, mode=mode))
    cur = {}
    im.dynamic_mask_head = _Fn(lambda **k: ([cur["cls"], cur["cls"]], [cur["emb"], cur["emb"]], [[None] * 4, [None] * 4]))
    model.generate_final_outputs = lambda feats, om, generate_aux_output=False: (feats, cur["masks"][None], [])
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.current_device = lambda: "cpu"
    img = torch.zeros(1, 3, H, W)
    out = {}
    fidx = 0
    for v, nf in enumerate(videos):
        embs = synthetic.make_track_sequence(seed * 10 + v, N, nf)
        for f in range(nf):
            logits, masks, _ = synthetic.make_fusion_case(1000 * seed + 100 * v + f, N, h, w)
            cur["cls"] = logits[None, None].repeat(7, 1, 1, 1)
            cur["emb"] = embs[f][None, None].repeat(7, 1, 1, 1)
            cur["masks"] = masks
            meta = [dict(iid=(v + 1) * 10000 + f + 1, filename="synthetic", ori_shape=(H, W, 3), img_shape=(H, W, 3))]
            with torch.no_grad():
                res = model.simple_test(img, meta, rescale=True, ref_img=[img])
            out["ids_%d" % fidx] = res["panoptic_det_obj_ids"].numpy().astype(np.int64)
            out["cls_inds_%d" % fidx] = res["panoptic_cls_inds"].numpy()
            out["bank_sum_%d" % fidx] = model.prev_instances.output_embedding.double().sum(1).numpy()
            fidx += 1
    out["videos"] = np.asarray(videos)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def main():
    """No argument: regenerate every fixture.  `--only a,b`: just those head cases.  `--check`: regenerate everything into a
    temporary directory and compare it with the committed files (bit-identical arrays expected)."""
    global HERE
    only = None
    if "--only" in sys.argv:
        only = sys.argv[sys.argv.index("--only") + 1].split(",")
    committed = HERE
    if "--check" in sys.argv:
        import tempfile
        HERE = tempfile.mkdtemp(prefix="golden_check_")
    torch.set_num_threads(8)
    model, _ = ref_import.build_model(0)
    for shp in POS_SHAPES:
        p = ref_pos(model, torch.zeros(1, 128, *shp))
        np.savez_compressed(os.path.join(HERE, "pos_%dx%d.npz" % shp), pos=p.numpy())
    fused = emb = cap = None
    for name, c in HEAD_CASES.items():
        if only and name not in only:
            continue
        if c["N"] != 100 or c.get("overrides"):
            m2, _ = ref_import.build_model(0, **{"other_config.proposal_num": c["N"], **(c.get("overrides") or {})})
        else:
            m2 = model
        f, e, cp = gen_head(m2, name, **c)
        if name == "head_t2_n100":
            fused, emb, cap = f, e, cp
    if only:
        print("golden fixtures written to", HERE)
        return
    gen_masklogit(model, fused, emb, cap)
    for name, c in FUSION_CASES.items():
        kw = dict(c)
        m3, _ = ref_import.build_model(0, **{"other_config.proposal_num": kw["N"]})
        gen_fusion(m3, name, kw.pop("seed"), kw.pop("N"), kw.pop("h"), kw.pop("w"), **kw)
        print(name, "done")
    gen_head_intrans(model)
    for name, c in UNIFY_CASES.items():
        gen_unify(name, c)
    for name, c in SEMANTIC_CASES.items():
        m5, _ = ref_import.build_model(0)
        gen_semantic(m5, name, c)
    for name, c in TRACK_CASES.items():
        m4, _ = ref_import.build_model(0, **{"other_config.proposal_num": c["N"]})
        gen_track(m4, name, **c)
        print(name, "done")
    print("golden fixtures written to", HERE)
    if committed != HERE:
        bad = 0
        for f in sorted(os.listdir(committed)):
            if not f.endswith(".npz"):
                continue
            a, b = np.load(os.path.join(committed, f)), np.load(os.path.join(HERE, f))
            same = set(a.files) == set(b.files) and all(np.array_equal(a[k], b[k]) for k in a.files)
            print(("identical  " if same else "DIFFERENT  ") + f)
            bad += 0 if same else 1
        sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
