"""Pin the oracle against the golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import slotvps_oracle as O
from slotvps_b200 import synthetic

spec = importlib.util.spec_from_file_location(
    "make_golden_cases", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))


def _cases():
    # read the case tables without importing the reference
    src = open(spec.origin).read()
    ns = {}
    start = src.index("HEAD_CASES = {")
    end = src.index("def ref_pos")
    exec(src[start:end], ns)
    return ns["HEAD_CASES"], ns["POS_SHAPES"], ns["FUSION_CASES"]


HEAD_CASES, POS_SHAPES, FUSION_CASES = _cases()


def head_cfg(case):
    """HeadConfig of a golden head case (the swinL-config case overrides the two FFN activations)."""
    ov = case.get("overrides") or {}
    return O.HeadConfig(activation=ov.get("dynamic_mask_head.activation", "gelu"),
                        temporal_activation=ov.get("dynamic_mask_head.temporal_query_attention_config.activation", "relu"))


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


@pytest.mark.parametrize("shape", POS_SHAPES)
def test_sine_position_embedding(golden_dir, shape):
    g = np.load(os.path.join(golden_dir, "pos_%dx%d.npz" % shape))["pos"]
    p = O.sine_position_embedding(*shape).numpy()
    assert p.shape == g.shape
    np.testing.assert_allclose(p, g, rtol=0, atol=2e-6)


@pytest.mark.parametrize("name", list(HEAD_CASES))
def test_head_forward(golden_dir, name):
    c = HEAD_CASES[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    P = synthetic.make_head_state_dict(c["seed"])
    cap = synthetic.make_capsule_params(c["seed"], c["N"])
    feats = synthetic.make_features(0, 0, T=c["T"], video=c["seed"], frame=0, shapes=c["shapes"])
    pos = [[O.sine_position_embedding(*s) for s in c["shapes"]] for _ in range(c["T"])]
    cls, emb, fused = O.head_forward(P, feats, [cap["init_mask_query.weight"]] * c["T"], pos, head_cfg(c))
    for t in range(c["T"]):
        assert cls[t].shape == g[f"cls{t}"].shape and emb[t].shape == g[f"emb{t}"].shape
        for l in range(4):
            f = fused[t][l][0]
            assert rel_l2(f[::7, ::3, ::5].numpy(), g[f"fused{t}_{l}_sample"]) < 2e-6
        # the 7-stage chain amplifies fp32 re-association noise: per-stage outputs
        # agree to ~1e-5 early and a few 1e-4 at the last stage
        # (measured: 7e-7 at stage 0 growing ~3-4x per stage to 5e-4..2e-3 at stage 6, and the fp64
        # oracle is no closer to the fp32 reference than the fp32 oracle is -> it is the
        # reference's own rounding noise, not a restatement error)
        # (the tight, amplification-free pinning is test_head_teacher_forced_against_golden: ~3e-6 per stage)
        for s in range(7):
            tol = 4e-6 * 3.5 ** s
            assert rel_l2(emb[t][s].numpy(), g[f"emb{t}"][s]) < tol, (t, s)
            assert rel_l2(cls[t][s].numpy(), g[f"cls{t}"][s]) < tol, (t, s)


def test_head_forward_fp64_close_to_reference(golden_dir):
    """fp64 oracle ('truth') vs the fp32 reference: bounded by the reference's own fp32 noise."""
    c = HEAD_CASES["head_t2_n100"]
    g = np.load(os.path.join(golden_dir, "head_t2_n100.npz"))
    P = {k: v.double() for k, v in synthetic.make_head_state_dict(c["seed"]).items()}
    cap = synthetic.make_capsule_params(c["seed"], c["N"])
    feats = [[f.double() for f in fr] for fr in synthetic.make_features(0, 0, T=2, video=c["seed"], shapes=c["shapes"])]
    pos = [[O.sine_position_embedding(*s, dtype=torch.float64) for s in c["shapes"]] for _ in range(2)]
    cls, emb, _ = O.head_forward(P, feats, [cap["init_mask_query.weight"].double()] * 2, pos)
    assert rel_l2(emb[1][-1].numpy(), g["emb1"][-1]) < 5e-3


def test_mask_logits(golden_dir):
    c = HEAD_CASES["head_t2_n100"]
    g = np.load(os.path.join(golden_dir, "masklogit.npz"))["pred_masks"]
    P = synthetic.make_head_state_dict(c["seed"])
    cap = synthetic.make_capsule_params(c["seed"], c["N"])
    feats = synthetic.make_features(0, 0, T=2, video=c["seed"], shapes=c["shapes"])
    pos = [[O.sine_position_embedding(*s) for s in c["shapes"]] for _ in range(2)]
    cls, emb, fused = O.head_forward(P, feats, [cap["init_mask_query.weight"]] * 2, pos)
    pm = O.mask_logits(fused[-1][-1][0], emb[-1][-1, 0], cap)
    assert pm.shape == g.shape
    assert rel_l2(pm.numpy(), g) < 5e-3     # inherits the stage-6 drift of the embedding
    # (the teacher-forced, tight version of this check is test_head_teacher_forced_against_golden below and, live
    #  against the imported reference, tests/test_oracle_vs_reference.py)


@pytest.mark.parametrize("name", ["head_t2_n100_big", "head_swinl", "head_t3_n128"])
def test_head_teacher_forced_against_golden(golden_dir, name):
    """Per-stage pinning without chain amplification: stage s of the oracle is fed the REFERENCE's stage s-1 embedding
    (golden; the reference carries it forward itself, dynamic_mask_head.py:210-211) and must reproduce the reference's
    stage-s outputs to fp32 re-association noise."""
    c = HEAD_CASES[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    cfg = head_cfg(c)
    P = synthetic.make_head_state_dict(c["seed"])
    cap = synthetic.make_capsule_params(c["seed"], c["N"])
    T = c["T"]
    feats = synthetic.make_features(0, 0, T=T, video=c["seed"], frame=0, shapes=c["shapes"])
    pos = [[O.sine_position_embedding(*s) for s in c["shapes"]] for _ in range(T)]
    q = cap["init_mask_query.weight"]
    forced = [[q] * T] + [[torch.from_numpy(g[f"emb{t}"][s, 0]) for t in range(T)] for s in range(6)]
    cls, emb, _ = O.head_forward(P, feats, [q] * T, pos, cfg, stage_slots_in=forced)
    worst = 0.0
    for t in range(T):
        for s in range(7):
            e, cc = rel_l2(emb[t][s].numpy(), g[f"emb{t}"][s]), rel_l2(cls[t][s].numpy(), g[f"cls{t}"][s])
            worst = max(worst, e, cc)
            assert e < 2e-5 and cc < 2e-5, (name, t, s, e, cc)
    print(f"{name}: teacher-forced oracle vs reference golden, worst per-stage rel-L2 {worst:.2e}")


@pytest.mark.parametrize("name", list(FUSION_CASES))
def test_panoptic_fusion_bit_exact(golden_dir, name):
    c = dict(FUSION_CASES[name])
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    seed, N, h, w = c.pop("seed"), c.pop("N"), c.pop("h"), c.pop("w")
    logits, masks, _ = synthetic.make_fusion_case(seed, N, h, w, **c)
    r = O.panoptic_fuse(logits, masks, (4 * h, 4 * w), want_masks=True)
    np.testing.assert_array_equal(r.labels, g["labels"])
    np.testing.assert_allclose(r.probs, g["probs"], rtol=1e-6)
    np.testing.assert_array_equal(r.cls_inds, g["cls_inds"])
    np.testing.assert_array_equal((r.masks != 0).sum((1, 2)), g["masks_nnz"])
    np.testing.assert_allclose(r.masks.astype(np.float64).sum((1, 2)), g["masks_sum"], rtol=1e-9, atol=1e-6)
    np.testing.assert_array_equal(r.panoptic, g["panoptic"])       # bit-identical id map
    assert r.panoptic.dtype == np.int64


TRACK_CASES = {
    "track_a": dict(seed=0, N=100, h=32, w=64, mode="identity"),
    "track_b": dict(seed=1, N=100, h=32, w=64, mode="random"),
    "track_c": dict(seed=2, N=50, h=24, w=40, mode="identity"),
}


def track_inputs(seed, N, h, w, videos):
    """The per-frame inputs tests/golden/make_golden.py::gen_track fed to the reference."""
    for v, nf in enumerate(videos):
        embs = synthetic.make_track_sequence(seed * 10 + v, N, int(nf))
        for f in range(int(nf)):
            logits, masks, _ = synthetic.make_fusion_case(1000 * seed + 100 * v + f, N, h, w)
            yield v, f, logits, masks, embs[f]


@pytest.mark.parametrize("name", sorted(TRACK_CASES))
def test_tracker_matches_reference(name):
    """Oracle tracker (SimpleTrackHead + greedy loop) vs panoptic_det_obj_ids / the object bank the reference's
    simple_test produced over consecutive frames, including the per-video reset."""
    c = TRACK_CASES[name]
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    sd = synthetic.make_track_params(c["seed"], mode=c["mode"])
    fcs = [(sd["fcs_query.%d.weight" % i], sd["fcs_query.%d.bias" % i]) for i in range(2)]
    st = O.TrackerState()
    seen = dict(new=0, matched=0, undone=0, lost=0)
    for i, (v, f, logits, masks, emb) in enumerate(track_inputs(c["seed"], c["N"], c["h"], c["w"], g["videos"])):
        if f == 0:
            st.reset()
        fr = O.panoptic_fuse(logits, masks, (4 * c["h"], 4 * c["w"]))
        ids, _, info = O.track_step(fcs, st, emb[torch.from_numpy(fr.keep.copy())].numpy(), fr.labels)
        np.testing.assert_array_equal(fr.cls_inds, g["cls_inds_%d" % i])
        np.testing.assert_array_equal(ids, g["ids_%d" % i])
        np.testing.assert_allclose(st.bank.astype(np.float64).sum(1), g["bank_sum_%d" % i], rtol=0, atol=1e-9)
        for k in seen:
            seen[k] += info[k]
    if c["mode"] == "identity":
        assert all(v > 0 for v in seen.values()), seen     # every branch of the greedy loop was reached


INTRANS_CASE = dict(seed=3, T=2, N=100, shapes=[(2, 4), (4, 8), (8, 16), (16, 32)])


def test_input_transform_then_head(golden_dir):
    """semantic_trans_ins (1x1 conv on every level) + head vs the reference run on UN-transformed features."""
    c = INTRANS_CASE
    g = np.load(os.path.join(golden_dir, "head_intrans.npz"))
    P = synthetic.make_head_state_dict(c["seed"])
    tp = synthetic.make_in_trans_params(c["seed"])
    cap = synthetic.make_capsule_params(c["seed"], c["N"])
    raw = synthetic.make_features(0, 0, T=c["T"], video=c["seed"], frame=0, shapes=c["shapes"])
    feats = O.input_transform(raw, tp["conv_trans.conv.weight"], tp["conv_trans.conv.bias"])
    pos = [[O.sine_position_embedding(*s) for s in c["shapes"]] for _ in range(c["T"])]
    cls, emb, fused = O.head_forward(P, feats, [cap["init_mask_query.weight"]] * c["T"], pos)
    for t in range(c["T"]):
        for l in range(4):
            assert rel_l2(fused[t][l][0][::7, ::3, ::5].numpy(), g[f"fused{t}_{l}_sample"]) < 2e-6
        for s in range(7):
            tol = 3e-6 * 3.5 ** s
            assert rel_l2(emb[t][s].numpy(), g[f"emb{t}"][s]) < tol, (t, s)
            assert rel_l2(cls[t][s].numpy(), g[f"cls{t}"][s]) < tol, (t, s)


def _golden_tables():
    src = open(spec.origin).read()
    ns = {"np": np, "torch": torch, "synthetic": synthetic}
    a, b = src.index("UNIFY_CASES = {"), src.index("def gen_unify")
    exec(src[a:b], ns)
    a, b = src.index("def semantic_input"), src.index("def gen_semantic")
    exec(src[a:b], ns)
    return ns["UNIFY_CASES"], ns["SEMANTIC_CASES"], ns["unify_inputs"], ns["semantic_input"]


UNIFY_CASES, SEMANTIC_CASES, unify_inputs, semantic_input = _golden_tables()


@pytest.mark.parametrize("name", sorted(UNIFY_CASES))
def test_unify_pan_result_bit_exact(golden_dir, name):
    """Oracle restatement of get_unified_pan_result vs the reference's own output (duplicate object ids with the
    persistent counter, hidden instances, stuff out-voting, the empty-cls_inds 255 path, small-stuff removal)."""
    c = UNIFY_CASES[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    segs, pans, cis, ois = unify_inputs(**c)
    out = O.unify_pan_result(segs, pans, cis, ois, c["limit"])
    for i, o in enumerate(out):
        np.testing.assert_array_equal(o, g["f%d" % i])


@pytest.mark.parametrize("name", sorted(SEMANTIC_CASES))
def test_semantic_argmax_matches_reference(golden_dir, name):
    c = SEMANTIC_CASES[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))["fcn_outputs"]
    got = O.semantic_argmax(semantic_input(**c), (c["H"], c["W"])).numpy()
    np.testing.assert_array_equal(got.astype(np.uint8), g)
