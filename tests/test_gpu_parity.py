"""GPU parity tests (run on the B200 box: ``pytest -m gpu``).  Every call goes through the C ABI
of libslotvps_b200.so; the oracle (CPU restatement pinned to the reference by tests/golden) is
only the checker.

Tolerances (BASELINE.json north_star): slots / mask logits within 1e-3 relative per stage with
teacher forcing (same stage inputs); id maps bit-identical except at pixels whose top-two logits
differ by less than that tolerance (counted and printed).  The 7-stage chain amplifies ANY fp32
rounding difference ~3.5x per stage (tests/test_oracle_golden.py measured it between the fp32
reference and the fp64 oracle), so end-to-end drift is bounded against that envelope, not 1e-3.
"""
import os

import numpy as np
import pytest
import torch

import slotvps_b200 as sv
from oracle import slotvps_oracle as O
from slotvps_b200 import synthetic

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    sv.lib()                                  # raises if the extension is missing: no silent fallback
    return torch.device("cuda:0")


def rel(a, b):
    a = a.detach().double().cpu() if isinstance(a, torch.Tensor) else torch.as_tensor(a).double()
    b = b.detach().double().cpu() if isinstance(b, torch.Tensor) else torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


DRIFT_X = 5.0        # free-running drift allowed relative to the reference arithmetic's own fp32-vs-fp64 drift on the same inputs


def fp32_noise(sd, feats, q, T, shapes, e64, cfg=None, pos=True):
    """Per-stage drift of the fp32 oracle (the reference's arithmetic) against the fp64 oracle on the same inputs: the
    noise floor every free-running comparison is measured against (instead of a fixed 3.5^s envelope)."""
    pos32 = [[O.sine_position_embedding(*s) for s in shapes] for _ in range(T)] if pos else None
    _, e32, _ = O.head_forward(sd, feats, [q] * T, pos32, cfg or O.HeadConfig())
    S = e32[0].shape[0]
    return [max(rel(e32[t][s], e64[t][s]) for t in range(T)) for s in range(S)]


def drift_ok(drift, noise, s):
    return drift <= DRIFT_X * max(noise[s], 2e-6)


def stage_dict(sd, pre):
    return {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}


PATHS = [1, 0]          # kernel_path: 1 = fp32 CUDA cores, 0 = tensor-core kernels where supported


def test_sine_pos(dev):
    for h, w in [(2, 4), (16, 32), (34, 60), (7, 5), (128, 256)]:
        got = sv.sine_position_embedding(h, w, dev).cpu()
        ref = O.sine_position_embedding(h, w)
        assert got.shape == ref.shape
        assert float((got - ref).abs().max()) < 5e-6, (h, w)


@pytest.mark.parametrize("h,w", [(4, 8), (6, 10), (34, 60), (64, 128)])
def test_level_fuse(dev, h, w):
    g = torch.Generator().manual_seed(h * 1000 + w)
    W = torch.randn(256, 384, generator=g) * 0.05
    b = torch.randn(256, generator=g) * 0.1
    x = torch.randn(128, h, w, generator=g)
    prev = torch.randn(256, h // 2, w // 2, generator=g)
    ref0 = O.level_fuse(None, x[None].double(), W.double(), b.double())[0]
    got0 = sv.level_fuse(None, x.to(dev), W.to(dev), b.to(dev))
    assert rel(got0, ref0) < 1e-5
    ref1 = O.level_fuse(prev[None].double(), x[None].double(), W.double(), b.double())[0]
    got1 = sv.level_fuse(prev.to(dev), x.to(dev), W.to(dev), b.to(dev))
    assert rel(got1, ref1) < 1e-5


@pytest.mark.parametrize("kernel_path", PATHS)
@pytest.mark.parametrize("N,h,w,use_pos", [(100, 16, 32, True), (100, 32, 64, True), (50, 12, 20, False),
                                           (128, 24, 40, True), (300, 16, 24, True), (7, 5, 3, True),
                                           (200, 32, 64, True), (209, 24, 40, False), (512, 16, 32, True)])
def test_slot_attention_teacher_forced(dev, kernel_path, N, h, w, use_pos):
    """MaskDynamicConv on given inputs (dynamic_mask_head.py:423-461) vs the fp64 oracle."""
    sd = synthetic.make_head_state_dict(5)
    pre = "head_series_2.0."
    g = torch.Generator().manual_seed(N * 7 + h)
    x = torch.randn(256, h, w, generator=g) * 1.5
    p = torch.randn(N, 256, generator=g)
    pos = O.sine_position_embedding(h, w)[0] if use_pos else None
    P64 = {k: v.double() for k, v in sd.items()}
    ref = O.pixel_attention(p.double(), x.double(), None if pos is None else pos.double(), P64, pre)
    ref32 = O.pixel_attention(p, x, pos, sd, pre)
    sp = {k: v.to(dev) for k, v in stage_dict(sd, pre).items()}
    got = sv.slot_attention(sp, p.to(dev), x.to(dev), None if pos is None else pos.to(dev), kernel_path)
    e, e32 = rel(got, ref), rel(ref32, ref)
    print(f"slot_attention path={kernel_path} N={N} {h}x{w}: rel err vs fp64 {e:.2e} (fp32 oracle itself {e32:.2e})")
    assert e < TOL


def _mk_head(dev, sd, kernel_path, **over):
    kw = {**sv.HEAD_KWARGS, **over, "kernel_path": kernel_path}
    head = sv.B200DynamicMaskHead(**kw)
    head.load_state_dict(sd, strict=True)
    return head.to(dev).eval()


@pytest.mark.parametrize("kernel_path", PATHS)
@pytest.mark.parametrize("temporal", [False, True])
def test_single_stage_teacher_forced(dev, kernel_path, temporal):
    """One MaskRCNNHead stage (+ Video Retriever) on given slots/features: every slot-side op."""
    T, N, shapes = 3, 100, [(8, 16)]
    over = dict(dh_num_heads=1, per_dh_num_heads=[1], feat_num_levels=1,
                apply_temporal_query_atten_stages=[0] if temporal else [5])
    sd = synthetic.make_head_state_dict(7, per_dh_num_heads=[1], temporal_stages=[0] if temporal else [5])
    feats = synthetic.make_features(0, 0, T=T, video=3, shapes=shapes)
    g = torch.Generator().manual_seed(11)
    slots = [torch.randn(N, 256, generator=g) for _ in range(T)]
    pos = [[O.sine_position_embedding(*shapes[0])] for _ in range(T)]
    cfg = O.HeadConfig(per_dh_num_heads=(1,), temporal_stages=(0,) if temporal else ())
    P64 = {k: v.double() for k, v in sd.items()}
    rc, re_, rf = O.head_forward(P64, [[f.double() for f in fr] for fr in feats], [s.double() for s in slots],
                                 [[p.double() for p in pp] for pp in pos], cfg)
    head = _mk_head(dev, sd, kernel_path, **over)
    for pos_arg in ([[p.to(dev) for p in pp] for pp in pos], "sine"):
        c, e, f = head([[f.to(dev) for f in fr] for fr in feats], [s.to(dev) for s in slots], None, pos=pos_arg)
        for t in range(T):
            assert c[t].shape == (1, 1, N, 20) and e[t].shape == (1, 1, N, 256) and f[t][0].shape == (1, 256, 8, 16)
            assert rel(f[t][0], rf[t][0]) < 1e-5
            assert rel(e[t], re_[t]) < TOL, (t, rel(e[t], re_[t]))
            assert rel(c[t], rc[t]) < TOL, (t, rel(c[t], rc[t]))
    print(f"stage path={kernel_path} temporal={temporal}: emb rel {rel(e[0], re_[0]):.2e} cls rel {rel(c[0], rc[0]):.2e}")


@pytest.mark.parametrize("kernel_path", PATHS)
@pytest.mark.parametrize("case", ["head_t2_n100", "head_t1_n50", "head_t3_n128"])
def test_head_chain_vs_oracle_and_golden(dev, kernel_path, case, golden_dir):
    """Full 7-stage chain on the golden cases: drift vs the fp64 oracle stays inside the envelope the
    reference's own fp32 arithmetic shows against fp64, and matches the REFERENCE's golden outputs."""
    from tests.test_oracle_golden import HEAD_CASES
    c = HEAD_CASES[case]
    gold = np.load(os.path.join(golden_dir, case + ".npz"))
    sd = synthetic.make_head_state_dict(c["seed"])
    cap = synthetic.make_capsule_params(c["seed"], c["N"])
    feats = synthetic.make_features(0, 0, T=c["T"], video=c["seed"], frame=0, shapes=c["shapes"])
    q = cap["init_mask_query.weight"]
    P64 = {k: v.double() for k, v in sd.items()}
    pos64 = [[O.sine_position_embedding(*s, dtype=torch.float64) for s in c["shapes"]] for _ in range(c["T"])]
    rc, re_, rf = O.head_forward(P64, [[f.double() for f in fr] for fr in feats], [q.double()] * c["T"], pos64)
    head = _mk_head(dev, sd, kernel_path)
    cl, em, fu = head([[f.to(dev) for f in fr] for fr in feats], [q.to(dev)] * c["T"], None, pos="sine")
    rows = []
    noise = fp32_noise(sd, feats, q, c["T"], c["shapes"], re_)
    gnoise = [max(rel(gold[f"emb{t}"][s], re_[t][s]) for t in range(c["T"])) for s in range(7)]   # the reference itself vs fp64
    for t in range(c["T"]):
        for l in range(4):
            assert rel(fu[t][l], rf[t][l]) < 1e-5
        for s in range(7):
            e64 = rel(em[t][s], re_[t][s])
            eg = rel(em[t][s], gold[f"emb{t}"][s])
            cg = rel(cl[t][s], gold[f"cls{t}"][s])
            rows.append((t, s, e64, eg, cg))
            assert drift_ok(e64, noise, s), (t, s, e64, noise[s])
            # against the reference's golden: both sides carry their own fp32 drift
            assert eg <= DRIFT_X * max(noise[s], gnoise[s], 2e-6) and cg <= DRIFT_X * max(noise[s], gnoise[s], 2e-6), (t, s, eg, cg)
    print(f"{case} path={kernel_path}: per-stage emb rel vs fp64 oracle (frame 0):",
          " ".join(f"{r[2]:.1e}" for r in rows[:7]), "| vs reference golden:", " ".join(f"{r[3]:.1e}" for r in rows[:7]))


def test_mask_logits(dev, golden_dir):
    g = torch.Generator().manual_seed(3)
    cap = synthetic.make_capsule_params(2, 100)
    for N, h, w in [(100, 16, 32), (37, 9, 13), (300, 20, 24)]:
        feat = torch.randn(256, h, w, generator=g) * 2
        emb = torch.relu(torch.randn(N, 256, generator=g))
        ref = O.mask_logits(feat.double(), emb.double(), {k: v.double() for k, v in cap.items()})
        got = sv.mask_logits(feat.to(dev), emb.to(dev), {k: v.to(dev) for k, v in cap.items()})
        assert got.shape == (N, h, w)
        assert rel(got, ref) < 1e-5, (N, h, w, rel(got, ref))
        assert float((got.cpu().double() - ref).abs().max()) < 1e-4


def _fusion_check(dev, logits, masks, size, fz=None):
    fz = fz or sv.PanopticFusion(**sv.FUSION_KWARGS)
    r = O.panoptic_fuse(logits, masks, size, want_masks=True)
    fo = fz.fuse(logits.to(dev), masks.to(dev), size, want_masks=len(r.labels))
    h = fo.host()
    got = fo.panoptic.cpu().numpy()
    assert got.dtype == np.int64 and got.shape == tuple(size)
    assert h["converged"] and h["k"] == len(r.labels)
    np.testing.assert_array_equal(h["keep"], r.keep)
    np.testing.assert_array_equal(h["labels"], r.labels)
    np.testing.assert_allclose(h["probs"], r.probs, rtol=2e-6)
    np.testing.assert_array_equal(h["cls_inds"], r.cls_inds)
    diff = got != r.panoptic
    hard = diff & ~r.near_tie
    gm = fo.masks[:h["k"]].cpu().numpy()
    mask_flip = int(((gm != 0) != (r.masks != 0)).sum())
    return int(diff.sum()), int(hard.sum()), int(r.near_tie.sum()), mask_flip, float(np.abs(gm - r.masks).max())


@pytest.mark.parametrize("case", ["fusion_a", "fusion_b", "fusion_c", "fusion_d", "fusion_e"])
def test_fusion_golden_cases(dev, case, golden_dir):
    """Designed fusion cases: identical to the oracle AND to the reference's own output (golden)."""
    from tests.test_oracle_golden import FUSION_CASES
    c = dict(FUSION_CASES[case])
    seed, N, h, w = c.pop("seed"), c.pop("N"), c.pop("h"), c.pop("w")
    logits, masks, _ = synthetic.make_fusion_case(seed, N, h, w, **c)
    ndiff, nhard, nnear, mflip, merr = _fusion_check(dev, logits, masks, (4 * h, 4 * w))
    print(f"{case}: id-map mismatches {ndiff} (outside near-tie: {nhard}; near-tie pixels {nnear}); "
          f"mask support flips {mflip}; max |mask logit err| {merr:.1e}")
    assert nhard == 0
    gold = np.load(os.path.join(golden_dir, case + ".npz"))
    fz = sv.PanopticFusion(**sv.FUSION_KWARGS)
    fo = fz.fuse(logits.to(dev), masks.to(dev), (4 * h, 4 * w))
    got = fo.panoptic.cpu().numpy()
    r = O.panoptic_fuse(logits, masks, (4 * h, 4 * w))
    assert int(((got != gold["panoptic"]) & ~r.near_tie).sum()) == 0
    np.testing.assert_array_equal(fo.host()["labels"], gold["labels"])


def test_fusion_edge_cases(dev):
    fz = sv.PanopticFusion(**sv.FUSION_KWARGS)
    # non-integer scale (VIPER: 272x480 -> 1080x1920 scaled down 8x here), same-size masks, x4
    for seed, (h, w), size in [(11, (34, 60), (135, 240)), (12, (24, 40), (24, 40)), (13, (16, 16), (64, 64))]:
        logits, masks, _ = synthetic.make_fusion_case(seed, 100, h, w, n_things=8, near_dup_things=2)
        ndiff, nhard, nnear, mflip, merr = _fusion_check(dev, logits, masks, size, fz)
        assert nhard == 0, (seed, ndiff, nhard)
    # single kept stuff slot
    logits = torch.full((100, 20), -4.0)
    logits[:, 19] = 4.0
    logits[5] = -4.0
    logits[5, 3] = 6.0
    masks = torch.randn(100, 8, 8)
    ndiff, nhard, *_ = _fusion_check(dev, logits, masks, (32, 32), fz)
    assert ndiff == 0
    # nothing kept: the reference raises (np.max of an empty array); the mirror raises ValueError
    logits[5] = -4.0
    logits[5, 19] = 4.0

    class Inst:
        pred_logits, pred_masks = logits.to(dev), masks.to(dev)
    with pytest.raises(ValueError):
        fz(Inst(), [(32, 32)])


def test_whole_clip_api_and_determinism(dev):
    """SlotVPSRetriever end to end at config-1 size (512x1024): shapes, determinism, fp32-vs-tc paths."""
    T, N, H, W = 2, 100, 512, 1024
    sd = synthetic.make_head_state_dict(0)
    cap = synthetic.make_capsule_params(0, N)
    feats = [[f.to(dev) for f in fr] for fr in synthetic.make_features(H, W, T=T, video=1, frame=2)]
    outs = []
    for kp in (1, 0, 0):
        m = sv.SlotVPSRetriever({**sv.HEAD_KWARGS, "kernel_path": kp}, N, sv.FUSION_KWARGS)
        m.dynamic_mask_head.load_state_dict(sd)
        m.load_capsule_params(cap)
        m = m.to(dev)
        outs.append(m(feats, (H, W), fuse=False))
    a, b, c = outs
    assert a["pred_masks"].shape == (N, H // 4, W // 4)
    assert a["emb"][1].shape == (7, 1, N, 256) and a["cls"][0].shape == (7, 1, N, 20)
    assert a["feats"][1][3].shape == (1, 256, H // 4, W // 4)
    assert torch.equal(b["pred_masks"], c["pred_masks"]) and torch.equal(b["emb"][1], c["emb"][1])   # run-to-run bitwise
    for s in range(7):
        print(f"512x1024 stage {s}: tc-vs-fp32 path emb rel {rel(b['emb'][1][s], a['emb'][1][s]):.2e}")
    assert rel(b["emb"][1][0], a["emb"][1][0]) < TOL
    # mask logits, teacher-forced on each path's own head outputs (tensor-core path reads the operand planes)
    for name, o in (("fp32", a), ("tc", b)):
        ref = O.mask_logits(o["feats"][1][3][0].double().cpu(), o["emb"][1][-1, 0].double().cpu(), {k: v.double() for k, v in cap.items()})
        e = rel(o["pred_masks"], ref)
        print(f"512x1024 mask logits ({name} path) vs fp64 oracle on the same inputs: rel {e:.2e}, max abs {float((o['pred_masks'].double().cpu() - ref).abs().max()):.2e}")
        assert e < 1e-4


def test_fullsize_fusion_and_properties(dev):
    """BASELINE full size (1024x2048, quarter-res masks 256x512): fusion id map vs the oracle, plus
    size-independent properties: labels come from the kept set, areas add up, re-running is bit-identical."""
    N, h, w = 100, 256, 512
    H, W = 4 * h, 4 * w
    logits, masks, _ = synthetic.make_fusion_case(21, N, h, w, n_things=15, near_dup_things=4, tiny=3)
    fz = sv.PanopticFusion(**sv.FUSION_KWARGS)
    fo = fz.fuse(logits.to(dev), masks.to(dev), (H, W))
    got = fo.panoptic.cpu().numpy()
    hst = fo.host()
    fo2 = fz.fuse(logits.to(dev), masks.to(dev), (H, W))
    assert np.array_equal(got, fo2.panoptic.cpu().numpy())                      # idempotent / deterministic
    r = O.panoptic_fuse(logits, masks, (H, W))
    np.testing.assert_array_equal(hst["keep"], r.keep)
    np.testing.assert_array_equal(hst["labels"], r.labels)
    diff = got != r.panoptic
    print(f"1024x2048 fusion: kept {hst['k']} ({hst['n_things']} things), iters {hst['iters']}, id-map mismatches {int(diff.sum())} "
          f"of {diff.size} (outside near-tie: {int((diff & ~r.near_tie).sum())}; near-tie pixels {int(r.near_tie.sum())})")
    assert int((diff & ~r.near_tie).sum()) == 0
    ids, counts = np.unique(got, return_counts=True)
    n_stuff_labels = set(int(c) for c in r.labels if c <= 10)
    assert set(ids.tolist()) <= n_stuff_labels | {11 + j for j in range(hst["n_things"])}
    assert counts.sum() == H * W and counts.min() > 4                          # every surviving segment has area > 4


def test_fullsize_clip_properties(dev):
    """One 1024x2048 T=2 clip through the whole path: run-to-run bitwise determinism, identical frames give
    identical slots on the frame-independent stages (0..2), fp32-vs-tensor-core agreement at stage 0."""
    T, N, H, W = 2, 100, 1024, 2048
    sd = synthetic.make_head_state_dict(0)
    cap = synthetic.make_capsule_params(0, N)
    one = synthetic.make_features(H, W, T=1, video=5, frame=0)[0]
    feats = [[f.to(dev) for f in one], [f.to(dev) for f in one]]                # the same frame twice
    lg = synthetic.make_fusion_case(0, N, 8, 8)[0].to(dev)
    res = {}
    for kp in (0, 1):
        m = sv.SlotVPSRetriever({**sv.HEAD_KWARGS, "kernel_path": kp}, N, sv.FUSION_KWARGS)
        m.dynamic_mask_head.load_state_dict(sd)
        m.load_capsule_params(cap)
        m = m.to(dev)
        res[kp] = m(feats, (H, W), fusion_logits=lg)
        if kp == 0:
            again = m(feats, (H, W), fusion_logits=lg)
            assert torch.equal(again["fusion"].panoptic, res[0]["fusion"].panoptic)
            assert torch.equal(again["pred_masks"], res[0]["pred_masks"])
    a = res[0]
    assert a["fusion"].panoptic.shape == (H, W) and a["fusion"].panoptic.dtype == torch.int64
    for s in range(3):                                                           # stages without the Video Retriever
        assert torch.equal(a["emb"][0][s], a["emb"][1][s]), s
    e0 = rel(a["emb"][1][0], res[1]["emb"][1][0])
    e6 = rel(a["emb"][1][6], res[1]["emb"][1][6])
    print(f"1024x2048: tensor-core vs fp32 path, stage 0 emb rel {e0:.2e}, stage 6 {e6:.2e}; kept {a['fusion'].host()['k']}")
    assert e0 < 1e-4


@pytest.mark.parametrize("heads,temporal", [([0, 0, 0, 1], [0]), ([0, 1, 1, 1], [1, 2]), ([1, 1, 1, 1], [2, 3]),
                                            ([1, 1, 2, 2], [2, 3, 4, 5])])
def test_config5_iteration_sweep(dev, heads, temporal):
    """BASELINE configs[4]: 1..6 retriever iterations via per_dh_num_heads (the combinations SURVEY.md 8d validated
    on the reference), teacher-free chain vs the fp64 oracle at reduced size."""
    T, N, shapes = 2, 100, [(4, 8), (8, 16), (16, 32), (32, 64)]
    S = sum(heads)
    sd = synthetic.make_head_state_dict(11, per_dh_num_heads=heads, temporal_stages=temporal)
    cap = synthetic.make_capsule_params(11, N)
    feats = synthetic.make_features(0, 0, T=T, video=11, shapes=shapes)
    q = cap["init_mask_query.weight"]
    cfg = O.HeadConfig(per_dh_num_heads=tuple(heads), temporal_stages=tuple(temporal))
    pos64 = [[O.sine_position_embedding(*s, dtype=torch.float64) for s in shapes] for _ in range(T)]
    rc, re_, rf = O.head_forward({k: v.double() for k, v in sd.items()}, [[f.double() for f in fr] for fr in feats], [q.double()] * T, pos64, cfg)
    head = _mk_head(dev, sd, 0, dh_num_heads=S, per_dh_num_heads=heads, apply_temporal_query_atten_stages=temporal)
    cl, em, fu = head([[f.to(dev) for f in fr] for fr in feats], [q.to(dev)] * T, None, pos="sine")
    assert em[0].shape == (S, 1, N, 256) and cl[1].shape == (S, 1, N, 20)
    noise = fp32_noise(sd, feats, q, T, shapes, re_, cfg)
    for t in range(T):
        for l in range(4):
            assert rel(fu[t][l], rf[t][l]) < 1e-5
        for s in range(S):
            assert drift_ok(rel(em[t][s], re_[t][s]), noise, s), (t, s, rel(em[t][s], re_[t][s]), noise[s])
    print(f"heads={heads}: last-stage emb rel {rel(em[1][-1], re_[1][-1]):.2e} (fp32 oracle: {noise[-1]:.2e})")


@pytest.mark.parametrize("N", [50, 200, 300])
def test_config5_slot_sweep(dev, N):
    """BASELINE configs[4]: slot-count sweep (N > 104 runs the tensor-core attention in slot groups of <= 104:
    a denominator pass per group, combined per-pixel softmax statistics, an accumulation pass per group)."""
    T, shapes = 2, [(4, 8), (8, 16), (16, 32), (32, 64)]
    sd = synthetic.make_head_state_dict(12)
    cap = synthetic.make_capsule_params(12, N)
    feats = synthetic.make_features(0, 0, T=T, video=12, shapes=shapes)
    q = cap["init_mask_query.weight"]
    pos64 = [[O.sine_position_embedding(*s, dtype=torch.float64) for s in shapes] for _ in range(T)]
    rc, re_, rf = O.head_forward({k: v.double() for k, v in sd.items()}, [[f.double() for f in fr] for fr in feats], [q.double()] * T, pos64)
    head = _mk_head(dev, sd, 0)
    cl, em, fu = head([[f.to(dev) for f in fr] for fr in feats], [q.to(dev)] * T, None, pos="sine")
    noise = fp32_noise(sd, feats, q, T, shapes, re_)
    for s in range(7):
        assert drift_ok(rel(em[1][s], re_[1][s]), noise, s), (s, rel(em[1][s], re_[1][s]), noise[s])
    print(f"N={N}: stage-0 emb rel {rel(em[1][0], re_[1][0]):.2e}, stage-6 {rel(em[1][6], re_[1][6]):.2e} (fp32 oracle: {noise[6]:.2e})")


def test_slot_groups_n300_t4(dev):
    """N = 300, T = 4: the slot-update kernels keep <= 104 slot rows per cluster, so a frame runs as
    three row groups of 100; the Video Retriever attends over all 1200 slots.  Free-running vs the fp64 oracle within the
    noise-relative bound, both kernel paths."""
    T, N, shapes = 4, 300, [(4, 8), (8, 16), (16, 32), (32, 64)]
    sd = synthetic.make_head_state_dict(14)
    cap = synthetic.make_capsule_params(14, N)
    feats = synthetic.make_features(0, 0, T=T, video=14, shapes=shapes)
    q = cap["init_mask_query.weight"]
    pos64 = [[O.sine_position_embedding(*s, dtype=torch.float64) for s in shapes] for _ in range(T)]
    rc, re_, rf = O.head_forward({k: v.double() for k, v in sd.items()}, [[f.double() for f in fr] for fr in feats], [q.double()] * T, pos64)
    noise = fp32_noise(sd, feats, q, T, shapes, re_)
    for kp in (0, 1):
        head = _mk_head(dev, sd, kp)
        cl, em, fu = head([[f.to(dev) for f in fr] for fr in feats], [q.to(dev)] * T, None, pos="sine")
        for t in range(T):
            for s in range(7):
                assert drift_ok(rel(em[t][s], re_[t][s]), noise, s), (kp, t, s, rel(em[t][s], re_[t][s]), noise[s])
            assert rel(cl[t][0], rc[t][0]) < 1e-4
        print(f"N=300 T=4 path={kp}: stage-0 emb rel {rel(em[3][0], re_[3][0]):.2e}, stage-6 {rel(em[3][6], re_[3][6]):.2e} (fp32 oracle: {noise[6]:.2e})")


def test_config4_viper_shape(dev):
    """BASELINE configs[3]: VIPER-shaped clip (1080x1920 padded to 1088x1920 -> levels 34x60..272x480, pixel counts
    that are not multiples of the 128-pixel tile), T=4 (Video Retriever over 400 slots), at 1/4 linear size vs the
    fp64 oracle, plus the fusion to the UNPADDED target size (non-integer scale)."""
    T, N = 4, 100
    shapes = [(9, 15), (18, 30), (36, 60), (72, 120)]           # 1/4-size analogue: same non-multiple-of-128 structure
    sd = synthetic.make_head_state_dict(13)
    cap = synthetic.make_capsule_params(13, N)
    feats = synthetic.make_features(0, 0, T=T, video=13, shapes=shapes)
    q = cap["init_mask_query.weight"]
    pos64 = [[O.sine_position_embedding(*s, dtype=torch.float64) for s in shapes] for _ in range(T)]
    rc, re_, rf = O.head_forward({k: v.double() for k, v in sd.items()}, [[f.double() for f in fr] for fr in feats], [q.double()] * T, pos64)
    m = sv.SlotVPSRetriever(sv.HEAD_KWARGS, N, sv.FUSION_KWARGS)
    m.dynamic_mask_head.load_state_dict(sd)
    m.load_capsule_params(cap)
    m = m.to(dev)
    lg, _, _ = synthetic.make_fusion_case(13, N, 8, 8)
    size = (270, 480)                                           # unpadded size: 72*4 = 288 rows padded, 270 real -> scale 3.75
    out = m([[f.to(dev) for f in fr] for fr in feats], size, fusion_logits=lg.to(dev))
    noise = fp32_noise(sd, feats, q, T, shapes, re_)
    for t in range(T):
        for s in range(7):
            assert drift_ok(rel(out["emb"][t][s], re_[t][s]), noise, s), (t, s, rel(out["emb"][t][s], re_[t][s]), noise[s])
    r = O.panoptic_fuse(lg, out["pred_masks"].cpu(), size)
    got = out["fusion"].panoptic.cpu().numpy()
    assert got.shape == size
    diff = got != r.panoptic
    print(f"VIPER-shaped T=4: stage-6 emb rel {rel(out['emb'][3][6], re_[3][6]):.2e}; fusion mismatches {int(diff.sum())} "
          f"(outside near-tie {int((diff & ~r.near_tie).sum())})")
    assert int((diff & ~r.near_tie).sum()) == 0


def test_viper_full_size_runs(dev):
    """Full VIPER shape (levels 34x60 .. 272x480, T=4) runs and is deterministic (no oracle at this size)."""
    T, N = 4, 100
    shapes = [(34, 60), (68, 120), (136, 240), (272, 480)]
    sd = synthetic.make_head_state_dict(0)
    cap = synthetic.make_capsule_params(0, N)
    feats = [[f.to(dev) for f in fr] for fr in synthetic.make_features(0, 0, T=T, video=3, shapes=shapes)]
    m = sv.SlotVPSRetriever(sv.HEAD_KWARGS, N, sv.FUSION_KWARGS)
    m.dynamic_mask_head.load_state_dict(sd)
    m.load_capsule_params(cap)
    m = m.to(dev)
    lg = synthetic.make_fusion_case(0, N, 8, 8)[0].to(dev)
    a = m(feats, (1080, 1920), fusion_logits=lg)
    b = m(feats, (1080, 1920), fusion_logits=lg)
    assert a["fusion"].panoptic.shape == (1080, 1920) and a["pred_masks"].shape == (N, 272, 480)
    assert torch.equal(a["fusion"].panoptic, b["fusion"].panoptic) and torch.equal(a["emb"][3], b["emb"][3])
    assert torch.isfinite(a["pred_masks"]).all()


@pytest.mark.parametrize("pos_kind", ["sine", "sine_separable", "tensor", "none"])
def test_stale_workspace_is_never_read(dev, pos_kind, monkeypatch):
    """Regression: scratch buffers pre-filled with NaN bit patterns (SLOTVPS_POISON) must not change the result.
    Pixel counts that are not multiples of the 128-pixel tile make tail tiles reach past the written rows; those
    rows have to come from TMA zero fill, never from stale memory (a NaN there corrupts the MMA even against zero
    weights -- found when a T=4 VIPER-shaped clip ran after a larger clip in the same process)."""
    monkeypatch.setenv("SLOTVPS_POISON", "127")
    monkeypatch.setenv("SLOTVPS_POISON_BYTE", "255")
    if pos_kind == "sine_separable":                    # opt-in mode: x planes only, position terms from tables
        monkeypatch.setenv("SLOTVPS_POS_SEP", "1")
        pos_kind = "sine"
    T, N, shapes = 3, 100, [(9, 15), (18, 30), (36, 60), (72, 120)]
    sd = synthetic.make_head_state_dict(14)
    cap = synthetic.make_capsule_params(14, N)
    feats = synthetic.make_features(0, 0, T=T, video=14, shapes=shapes)
    q = cap["init_mask_query.weight"]
    pos64 = None if pos_kind == "none" else [[O.sine_position_embedding(*s, dtype=torch.float64) for s in shapes] for _ in range(T)]
    rc, re_, rf = O.head_forward({k: v.double() for k, v in sd.items()}, [[f.double() for f in fr] for fr in feats], [q.double()] * T, pos64)
    head = _mk_head(dev, sd, 0)
    pos_arg = {"sine": "sine", "none": None,
               "tensor": None if pos64 is None else [[p.float().to(dev) for p in pp] for pp in pos64]}[pos_kind]
    cl, em, fu = head([[f.to(dev) for f in fr] for fr in feats], [q.to(dev)] * T, None, pos=pos_arg)
    noise = fp32_noise(sd, feats, q, T, shapes, re_, pos=pos64 is not None)
    for t in range(T):
        for s in range(7):
            e = rel(em[t][s], re_[t][s])
            assert drift_ok(e, noise, s), (pos_kind, t, s, e, noise[s])


# ---- tracker -------------------------------------------------------------------------
def _track_head(dev, seed, mode):
    th = sv.B200TrackHead(**sv.TRACK_KWARGS)
    sd = synthetic.make_track_params(seed, mode=mode)
    th.load_state_dict(sd, strict=True)               # the reference's parameter names
    fcs = [(sd["fcs_query.%d.weight" % i], sd["fcs_query.%d.bias" % i]) for i in range(2)]
    return th.to(dev), fcs


@pytest.mark.parametrize("mode", ["identity", "random"])
@pytest.mark.parametrize("k,m", [(1, 1), (22, 37), (100, 333)])
def test_track_head_scores(dev, mode, k, m):
    """SimpleTrackHead.forward: [0 | fc(x) fc(ref)^T] vs the oracle (fp32 CUDA cores, different summation order)."""
    th, fcs = _track_head(dev, 3, mode)
    g = torch.Generator().manual_seed(k * 1000 + m)
    x, r = torch.randn(k, 256, generator=g), torch.randn(m, 256, generator=g)
    got = th(x.to(dev), r.to(dev))
    assert isinstance(got, list) and len(got) == 1 and got[0].shape == (k, 1 + m)
    ref = O.track_match_scores(fcs, x, r)
    assert float(got[0][:, 0].abs().max()) == 0.0
    assert rel(got[0], ref) < 2e-6
    both = th(x.to(dev), [r.to(dev), r[: max(1, m // 2)].to(dev)])     # list of banks, as the reference accepts
    assert len(both) == 2 and both[1].shape == (k, 1 + max(1, m // 2))
    with pytest.raises(RuntimeError):
        th(x, r)                                                     # CPU tensors: no fallback


@pytest.mark.parametrize("name", ["track_a", "track_b", "track_c"])
def test_tracker_golden_videos(dev, name, golden_dir):
    """Fusion + on-device tracker over consecutive frames == the reference's simple_test (golden) and the oracle:
    det_obj_ids, bank length and bank contents after every frame, including the per-video reset."""
    from tests.test_oracle_golden import TRACK_CASES, track_inputs
    c = TRACK_CASES[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    th, fcs = _track_head(dev, c["seed"], c["mode"])
    fz = sv.PanopticFusion(**sv.FUSION_KWARGS)
    trk = sv.SlotTracker(th, n_slots=c["N"], capacity=256, device=dev)
    st = O.TrackerState()
    for i, (v, f, logits, masks, emb) in enumerate(track_inputs(c["seed"], c["N"], c["h"], c["w"], g["videos"])):
        if f == 0:
            trk.reset()
            st.reset()
        fo = fz.fuse(logits.to(dev), masks.to(dev), (4 * c["h"], 4 * c["w"]))
        rec = sv.SlotTracker.host(trk.step(emb.to(dev), fo))
        fr = O.panoptic_fuse(logits, masks, (4 * c["h"], 4 * c["w"]))
        ids_things, ids_all, _ = O.track_step(fcs, st, emb[torch.from_numpy(fr.keep.copy())].numpy(), fr.labels)
        np.testing.assert_array_equal(rec["ids"], ids_all)
        np.testing.assert_array_equal(rec["det_obj_ids"], g["ids_%d" % i])
        assert rec["bank"] == st.bank.shape[0] == g["bank_sum_%d" % i].shape[0]
        bank = trk.bank().cpu().numpy()
        np.testing.assert_array_equal(bank, st.bank)                  # raw embeddings are copied, never recomputed
        np.testing.assert_allclose(bank.astype(np.float64).sum(1), g["bank_sum_%d" % i], rtol=0, atol=1e-9)


def test_tracker_capacity_and_empty(dev):
    """Bank overflow is reported, not silently dropped; a frame with no kept slot leaves the state untouched."""
    th, _ = _track_head(dev, 0, "identity")
    N = 100
    fz = sv.PanopticFusion(**sv.FUSION_KWARGS)
    trk = sv.SlotTracker(th, n_slots=N, capacity=8, device=dev)
    logits, masks, _ = synthetic.make_fusion_case(7, N, 16, 32)
    fo = fz.fuse(logits.to(dev), masks.to(dev), (64, 128))
    out = trk.step(synthetic.make_track_sequence(0, N, 1)[0].to(dev), fo)
    assert int(out[3]) == 1 and int(out[2]) == 8
    with pytest.raises(RuntimeError):
        sv.SlotTracker.host(out)
    trk2 = sv.SlotTracker(th, n_slots=N, capacity=64, device=dev)
    none = torch.full((N, 20), -4.0)
    none[:, 19] = 4.0
    fo0 = fz.fuse(none.to(dev), masks.to(dev), (64, 128))
    rec = sv.SlotTracker.host(trk2.step(torch.zeros(N, 256, device=dev), fo0))
    assert rec["k"] == 0 and rec["bank"] == 0
    rec = sv.SlotTracker.host(trk2.step(synthetic.make_track_sequence(0, N, 1)[0].to(dev), fo))
    assert rec["ids"].tolist() == list(range(rec["k"]))               # still the "first frame" of the video


# ---- folded input transform ------------------------------------------------------------
@pytest.mark.parametrize("kernel_path", PATHS)
def test_folded_input_transform_golden(dev, kernel_path, golden_dir):
    """head.fold_input_transform(conv_trans) on UN-transformed features == the reference's semantic_trans_ins + head."""
    from tests.test_oracle_golden import INTRANS_CASE as c
    gold = np.load(os.path.join(golden_dir, "head_intrans.npz"))
    sd = synthetic.make_head_state_dict(c["seed"])
    tp = synthetic.make_in_trans_params(c["seed"])
    cap = synthetic.make_capsule_params(c["seed"], c["N"])
    raw = synthetic.make_features(0, 0, T=c["T"], video=c["seed"], frame=0, shapes=c["shapes"])
    head = _mk_head(dev, sd, kernel_path)
    head.fold_input_transform(tp["conv_trans.conv.weight"], tp["conv_trans.conv.bias"])
    # noise floor: the fp32 oracle on the transformed features against the reference's golden
    feats_t = O.input_transform(raw, tp["conv_trans.conv.weight"], tp["conv_trans.conv.bias"])
    pos32 = [[O.sine_position_embedding(*s) for s in c["shapes"]] for _ in range(c["T"])]
    _, e32, _ = O.head_forward(sd, feats_t, [cap["init_mask_query.weight"]] * c["T"], pos32)
    noise = [max(rel(e32[t][s], gold[f"emb{t}"][s]) for t in range(c["T"])) for s in range(7)]
    q = cap["init_mask_query.weight"].to(dev)
    cl, em, fu = head([[f.to(dev) for f in fr] for fr in raw], [q] * c["T"], None, pos="sine")
    for t in range(c["T"]):
        for l in range(4):
            assert rel(fu[t][l][0][::7, ::3, ::5], gold[f"fused{t}_{l}_sample"]) < 1e-5
        for s in range(7):
            env = DRIFT_X * max(noise[s], 2e-6)
            assert rel(em[t][s], gold[f"emb{t}"][s]) <= env and rel(cl[t][s], gold[f"cls{t}"][s]) <= env, (t, s)
    # removing the fold restores the plain head (transformed features in)
    head.fold_input_transform(None, None)
    feats = O.input_transform(raw, tp["conv_trans.conv.weight"], tp["conv_trans.conv.bias"])
    cl2, em2, fu2 = head([[f.to(dev) for f in fr] for fr in feats], [q] * c["T"], None, pos="sine")
    assert rel(fu2[0][3], fu[0][3]) < 1e-5 and rel(em2[0][0], em[0][0]) < 3e-5


def test_folded_input_transform_tensor_core_path(dev):
    """Same at a size where every level runs the tensor-core kernels (fuse_tc with folded W0 / Wb and per-level bias)."""
    shapes = [(8, 16), (16, 32), (32, 64), (64, 128)]
    sd = synthetic.make_head_state_dict(5)
    tp = synthetic.make_in_trans_params(5)
    cap = synthetic.make_capsule_params(5, 100)
    raw = synthetic.make_features(0, 0, T=2, video=5, frame=0, shapes=shapes)
    feats64 = O.input_transform([[f.double() for f in fr] for fr in raw], tp["conv_trans.conv.weight"].double(), tp["conv_trans.conv.bias"].double())
    P64 = {k: v.double() for k, v in sd.items()}
    head = _mk_head(dev, sd, 0)
    head.fold_input_transform(tp["conv_trans.conv.weight"], tp["conv_trans.conv.bias"])
    q = cap["init_mask_query.weight"]
    cl, em, fu = head([[f.to(dev) for f in fr] for fr in raw], [q.to(dev)] * 2, None, pos="sine")
    W64 = P64["conv_trans.conv.weight"].reshape(256, 384)
    for t in range(2):
        prev = None
        for l in range(4):
            ref = O.level_fuse(prev, feats64[t][l], W64, P64["conv_trans.conv.bias"])
            assert rel(fu[t][l], ref) < 1e-5, (t, l, rel(fu[t][l], ref))
            prev = ref
    pos64 = [[O.sine_position_embedding(*s, dtype=torch.float64) for s in shapes] for _ in range(2)]
    rc, re_, _ = O.head_forward(P64, feats64, [q.double()] * 2, pos64)
    noise = fp32_noise(sd, [[f.float() for f in fr] for fr in feats64], q, 2, shapes, re_)
    for s in range(7):
        assert drift_ok(rel(em[1][s], re_[1][s]), noise, s), (s, rel(em[1][s], re_[1][s]), noise[s])


# ---- consumers of the id map -------------------------------------------------------------
@pytest.mark.parametrize("name", ["unify_a", "unify_b", "unify_c"])
def test_unify_pan_result_golden(dev, name, golden_dir):
    """get_unified_pan_result on device == the reference's own output, byte for byte, over the frames of a call."""
    from tests.test_oracle_golden import UNIFY_CASES, unify_inputs
    c = UNIFY_CASES[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    segs, pans, cis, ois = unify_inputs(**c)
    names = ["f%d" % i for i in range(len(segs))]
    got = sv.get_unified_pan_result(segs, pans, cis, obj_ids=ois, stuff_area_limit=c["limit"], names=names, device=dev)
    ref = O.unify_pan_result(segs, pans, cis, ois, c["limit"])
    for i, n in enumerate(names):
        assert got[n].dtype == np.uint8 and got[n].shape == g[n].shape
        np.testing.assert_array_equal(got[n], g[n])
        np.testing.assert_array_equal(got[n], ref[i])


def test_unify_full_size_and_errors(dev):
    """1024x2048 frames: equals the oracle; ragged tail (H*W not a multiple of 4/8); error reporting."""
    for seed, (H, W) in [(20, (1024, 2048)), (21, (37, 53))]:
        seg, pan, ci, oi = synthetic.make_unify_case(seed, H, W, n_inst=40 if H > 100 else 6, dup_obj=5 if H > 100 else 1)
        got = sv.get_unified_pan_result([seg.to(dev)], [pan.to(dev)], [ci], obj_ids=[oi], stuff_area_limit=4096, names=["x"], device=dev)["x"]
        ref = O.unify_pan_result([seg.numpy()], [pan.numpy()], [ci.numpy()], [oi.numpy()], 4096)[0]
        np.testing.assert_array_equal(got, ref)
    seg, pan, ci, oi = synthetic.make_unify_case(22, 64, 64, n_inst=6, hidden=0)
    with pytest.raises(IndexError):                    # cls_inds shorter than the ids present: the reference raises too
        sv.get_unified_pan_result([seg], [pan], [ci[:2]], obj_ids=[oi], names=["x"], device=dev)
    with pytest.raises(ValueError):
        sv.get_unified_pan_result([seg], [pan + 300], [ci], obj_ids=[oi], names=["x"], device=dev)
    with pytest.raises(RuntimeError):                  # no CPU path
        sv.PanUnifier(dev).frame(seg, pan, ci, oi)


@pytest.mark.parametrize("name", ["semantic_same", "semantic_up4"])
def test_semantic_argmax_golden(dev, name, golden_dir):
    from tests.test_oracle_golden import SEMANTIC_CASES, semantic_input
    c = SEMANTIC_CASES[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))["fcn_outputs"]
    x = semantic_input(**c)
    got = sv.semantic_argmax(x.to(dev), (c["H"], c["W"])).cpu().numpy()
    assert got.shape == (1, c["H"], c["W"])
    # pixels whose two largest (resized) logits are closer than 1e-4 may resolve differently (exp rounding); none expected
    up = torch.nn.functional.interpolate(x, size=(c["H"], c["W"]), mode="bilinear", align_corners=False) if x.shape[-1] != c["W"] else x
    top2 = up.topk(2, dim=1).values
    near = ((top2[:, 0] - top2[:, 1]) < 1e-4).numpy()
    diff = (got.astype(np.uint8) != g)
    assert int((diff & ~near).sum()) == 0
    print(f"{name}: mismatches {int(diff.sum())} (near-tie pixels {int(near.sum())})")


def test_semantic_argmax_full_size(dev):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 19, 256, 512, generator=g) * 3
    got = sv.semantic_argmax(x.to(dev), (1024, 2048)).cpu()
    ref = O.semantic_argmax(x, (1024, 2048))
    up = torch.nn.functional.interpolate(x, size=(1024, 2048), mode="bilinear", align_corners=False)
    top2 = up.topk(2, dim=1).values
    near = (top2[:, 0] - top2[:, 1]) < 1e-4
    assert int(((got != ref) & ~near).sum()) == 0


@pytest.mark.parametrize("N", [200, 300])
def test_slot_groups_whole_clip(dev, N):
    """N > 104 through the whole tensor-core path at 256x512 (grouped attention + grouped mask logits) vs the fp32 path
    and the fp64 oracle (mask logits teacher-forced)."""
    T, H, W = 2, 256, 512
    sd = synthetic.make_head_state_dict(3)
    cap = synthetic.make_capsule_params(3, N)
    feats = [[f.to(dev) for f in fr] for fr in synthetic.make_features(H, W, T=T, video=4, frame=1)]
    outs = []
    for kp in (1, 0):
        m = sv.SlotVPSRetriever({**sv.HEAD_KWARGS, "kernel_path": kp}, N, sv.FUSION_KWARGS)
        m.dynamic_mask_head.load_state_dict(sd)
        m.load_capsule_params(cap)
        outs.append(m.to(dev)(feats, (H, W), fuse=False))
    a, b = outs
    e0 = rel(b["emb"][1][0], a["emb"][1][0])
    print(f"N={N} 256x512: tc-vs-fp32 path stage-0 emb rel {e0:.2e}, stage-6 {rel(b['emb'][1][6], a['emb'][1][6]):.2e}")
    assert e0 < 3e-5
    ref = O.mask_logits(b["feats"][1][3][0].double().cpu(), b["emb"][1][-1, 0].double().cpu(), {k: v.double() for k, v in cap.items()})
    assert b["pred_masks"].shape == (N, H // 4, W // 4)
    assert rel(b["pred_masks"], ref) < 1e-4


def test_explicit_pos_tensors_match_sine_mode(dev):
    """L1 integration passes the reference's PositionEmbeddingSine tensors (pos_mode 1): the level-fusion epilogue then
    reads them from HBM instead of generating the embedding; results must agree with pos="sine" and with the oracle."""
    shapes = [(8, 16), (16, 32), (32, 64), (64, 128)]
    sd = synthetic.make_head_state_dict(9)
    cap = synthetic.make_capsule_params(9, 100)
    feats = synthetic.make_features(0, 0, T=2, video=9, frame=0, shapes=shapes)
    q = cap["init_mask_query.weight"]
    head = _mk_head(dev, sd, 0)
    f_dev = [[f.to(dev) for f in fr] for fr in feats]
    cl_s, em_s, fu_s = head(f_dev, [q.to(dev)] * 2, None, pos="sine")
    em_s = [e.clone() for e in em_s]
    pos = [[O.sine_position_embedding(*s).to(dev) for s in shapes] for _ in range(2)]
    cl_t, em_t, fu_t = head(f_dev, [q.to(dev)] * 2, None, pos=pos)
    assert rel(em_t[1][0], em_s[1][0]) < 2e-5 and rel(fu_t[1][3], fu_s[1][3]) < 1e-6
    P64 = {k: v.double() for k, v in sd.items()}
    pos64 = [[O.sine_position_embedding(*s, dtype=torch.float64) for s in shapes] for _ in range(2)]
    _, re_, _ = O.head_forward(P64, [[f.double() for f in fr] for fr in feats], [q.double()] * 2, pos64)
    noise = fp32_noise(sd, feats, q, 2, shapes, re_)
    for s in range(7):
        assert drift_ok(rel(em_t[1][s], re_[1][s]), noise, s), (s, rel(em_t[1][s], re_[1][s]), noise[s])


# ---- round 2: parity at the metric's own size, per-stage teacher forcing, drift relative to the reference's own noise ----
def _teacher_forcing_inputs(q, emb_ref, T):
    """Slots entering stage s = the checker's stage s-1 embeddings (dynamic_mask_head.py:210-211); stage 0 = init query."""
    return [[q] * T] + [[emb_ref[t][s, 0] for t in range(T)] for s in range(6)]


def _fullsize_parity(dev, shapes, T, seed, paths, label):
    N = 100
    sd = synthetic.make_head_state_dict(seed)
    cap = synthetic.make_capsule_params(seed, N)
    feats = synthetic.make_features(0, 0, T=T, video=seed, frame=0, shapes=shapes)
    q = cap["init_mask_query.weight"]
    P64 = {k: v.double() for k, v in sd.items()}
    pos32 = [[O.sine_position_embedding(*s) for s in shapes] for _ in range(T)]
    pos64 = [[p.double() for p in pp] for pp in pos32]
    torch.set_num_threads(os.cpu_count() or 1)
    c64, e64, f64 = O.head_forward(P64, [[f.double() for f in fr] for fr in feats], [q.double()] * T, pos64)
    c32, e32, f32 = O.head_forward(sd, feats, [q] * T, pos32)
    noise = [max(rel(e32[t][s], e64[t][s]) for t in range(T)) for s in range(7)]     # the fp32 reference arithmetic's own drift
    forced = _teacher_forcing_inputs(q, e64, T)
    f_dev = [[f.to(dev) for f in fr] for fr in feats]
    for kp in paths:
        head = _mk_head(dev, sd, kp)
        cl, em, fu = head(f_dev, [q.to(dev)] * T, None, pos="sine")
        for t in range(T):
            for l in range(4):
                e = rel(fu[t][l], f64[t][l])
                assert e < 1e-5, (label, kp, t, l, e)
        drift = [max(rel(em[t][s], e64[t][s]) for t in range(T)) for s in range(7)]
        assert drift[0] < 1e-4, (label, kp, drift[0])
        for s in range(7):
            # free-running drift: at most 5x what the reference's own fp32 arithmetic shows against fp64 on the same clip
            assert drift[s] <= 5 * max(noise[s], 2e-6), (label, kp, s, drift[s], noise[s])
        # every stage teacher-forced on the fp64 checker's stage inputs: the <= 1e-3 north-star bound, per stage
        cl_tf, em_tf, _ = head(f_dev, [q.to(dev)] * T, None, pos="sine",
                               stage_slots_in=[[v.float() for v in st] for st in forced])
        tf = [max(max(rel(em_tf[t][s], e64[t][s]), rel(cl_tf[t][s], c64[t][s])) for t in range(T)) for s in range(7)]
        for s in range(7):
            assert tf[s] < TOL, (label, kp, s, tf[s])
        # mask logits, teacher-forced on the checker's finest feature / last embedding (through the operand planes on path 0)
        print(f"{label} path={kp}: fused-feature rel {max(rel(fu[t][l], f64[t][l]) for t in range(T) for l in range(4)):.1e} | "
              f"teacher-forced per-stage rel vs fp64 oracle: {' '.join(f'{v:.1e}' for v in tf)} | free-running drift: "
              f"{' '.join(f'{v:.1e}' for v in drift)} | fp32-oracle-vs-fp64 (reference noise): {' '.join(f'{v:.1e}' for v in noise)} | "
              f"drift/noise max {max(d / max(n, 2e-6) for d, n in zip(drift, noise)):.2f}x")
    return sd, cap, feats, e64, f64


def test_fullsize_head_vs_oracle(dev):
    """BASELINE configs[1] (1024x2048, T=2, N=100): both kernel paths against the fp64 AND fp32 oracle at the metric's
    own size -- fused features, stage 0, all 7 stages teacher-forced (<= 1e-3), free-running drift <= 5x the drift the
    reference's own fp32 arithmetic shows against fp64 on the same clip, and the mask logits teacher-forced."""
    shapes = synthetic.level_shapes(1024, 2048)
    sd, cap, feats, e64, f64 = _fullsize_parity(dev, shapes, 2, 31, (0, 1), "1024x2048 T=2")
    # mask logits through the whole-clip API (operand planes of the finest level), teacher-forced on ITS OWN head outputs
    N, T = 100, 2
    m = sv.SlotVPSRetriever(sv.HEAD_KWARGS, N, sv.FUSION_KWARGS)
    m.dynamic_mask_head.load_state_dict(sd)
    m.load_capsule_params(cap)
    m = m.to(dev)
    f_dev = [[f.to(dev) for f in fr] for fr in feats]
    a = m(f_dev, (1024, 2048), fuse=False)
    ref = O.mask_logits(a["feats"][1][3][0].double().cpu(), a["emb"][1][-1, 0].double().cpu(), {k: v.double() for k, v in cap.items()})
    e = rel(a["pred_masks"], ref)
    b = m(f_dev, (1024, 2048), fuse=False, want_feats=False)           # L2 form: no fp32 features, norms from the fusion epilogue
    assert b["feats"][1][3] is None
    e2 = rel(b["pred_masks"], ref)
    print(f"1024x2048 mask logits teacher-forced vs fp64 oracle: rel {e:.2e}; without fp32 features (epilogue norms): {e2:.2e}; "
          f"emb identical: {torch.equal(a['emb'][1], b['emb'][1])}")
    assert e < 1e-4 and e2 < 1e-4
    assert torch.equal(a["emb"][1], b["emb"][1]) and torch.equal(a["cls"][0], b["cls"][0])


def test_viper_full_size_vs_oracle(dev):
    """BASELINE configs[3] at its real size (1080x1920 padded to 1088x1920 -> levels 34x60 .. 272x480, pixel counts that
    are not multiples of the 128-pixel tile; T=4: Video Retriever over 400 slots) against the fp64 / fp32 oracle."""
    shapes = [(34, 60), (68, 120), (136, 240), (272, 480)]
    _fullsize_parity(dev, shapes, 4, 32, (0,), "VIPER 1088x1920 T=4")


@pytest.mark.parametrize("kernel_path", PATHS)
@pytest.mark.parametrize("case", ["head_t2_n100_big", "head_swinl"])
def test_head_big_golden_teacher_forced(dev, kernel_path, case, golden_dir):
    """Golden cases whose EVERY level has >= 128 pixels (tcgen05 on all levels) against the REFERENCE's own outputs:
    per stage teacher-forced with the reference's stage s-1 embedding, and free-running within 5x of the drift the fp32
    oracle shows against the same golden.  `head_swinl` is the reference's second shipped config (ReLU stage FFN, GELU
    temporal FFN, configs/cityscapes/swinL_fpn_slotvps.py:41,56)."""
    from tests.test_oracle_golden import HEAD_CASES, head_cfg
    c = HEAD_CASES[case]
    gold = np.load(os.path.join(golden_dir, case + ".npz"))
    T = c["T"]
    cfg = head_cfg(c)
    sd = synthetic.make_head_state_dict(c["seed"])
    cap = synthetic.make_capsule_params(c["seed"], c["N"])
    feats = synthetic.make_features(0, 0, T=T, video=c["seed"], frame=0, shapes=c["shapes"])
    q = cap["init_mask_query.weight"]
    over = dict(activation=cfg.activation,
                temporal_query_attention_config={**sv.HEAD_KWARGS["temporal_query_attention_config"], "activation": cfg.temporal_activation})
    head = _mk_head(dev, sd, kernel_path, **over)
    f_dev = [[f.to(dev) for f in fr] for fr in feats]
    forced = [[q] * T] + [[torch.from_numpy(gold[f"emb{t}"][s, 0]) for t in range(T)] for s in range(6)]
    cl, em, fu = head(f_dev, [q.to(dev)] * T, None, pos="sine", stage_slots_in=forced)
    tf = [max(max(rel(em[t][s], gold[f"emb{t}"][s]), rel(cl[t][s], gold[f"cls{t}"][s])) for t in range(T)) for s in range(7)]
    for t in range(T):
        for l in range(4):
            assert rel(fu[t][l][0][::7, ::3, ::5], gold[f"fused{t}_{l}_sample"]) < 1e-5
    assert max(tf) < 5e-5, tf                      # ~1e-5 measured; the north-star bound is 1e-3
    pos32 = [[O.sine_position_embedding(*s) for s in c["shapes"]] for _ in range(T)]
    _, e32, _ = O.head_forward(sd, feats, [q] * T, pos32, cfg)
    cl2, em2, _ = head(f_dev, [q.to(dev)] * T, None, pos="sine")
    noise = [max(rel(e32[t][s], gold[f"emb{t}"][s]) for t in range(T)) for s in range(7)]
    drift = [max(rel(em2[t][s], gold[f"emb{t}"][s]) for t in range(T)) for s in range(7)]
    for s in range(7):
        assert drift[s] <= 5 * max(noise[s], 2e-6), (s, drift[s], noise[s])
    print(f"{case} path={kernel_path}: teacher-forced per-stage rel vs REFERENCE golden: {' '.join(f'{v:.1e}' for v in tf)} | "
          f"free-running: {' '.join(f'{v:.1e}' for v in drift)} | fp32 oracle vs golden: {' '.join(f'{v:.1e}' for v in noise)}")


def test_postprocess_forward_with_instances(dev):
    """PostProcessPanopticInstances.forward's SUCCESS path (vps_temporal_slots.py:659-807) with an Instances-like object
    (structures/instances.py:134-152: boolean / index filtering of every field): the filtered object carries the kept
    slots' fields in the post-processor's order plus .masks / .probs / .labels, as simple_test consumes them (:313-320)."""
    class Inst:                                         # the subset of the reference's Instances the call touches
        def __init__(self, **f):
            self.__dict__["_f"] = dict(f)

        def __getattr__(self, k):
            try:
                return self.__dict__["_f"][k]
            except KeyError:
                raise AttributeError(k)

        def __setattr__(self, k, v):
            self._f[k] = v

        def __getitem__(self, idx):
            return Inst(**{k: v[idx] for k, v in self._f.items()})

    N, h, w = 100, 24, 40
    logits, masks, _ = synthetic.make_fusion_case(5, N, h, w, n_things=10, near_dup_things=2, tiny=1)
    emb = torch.randn(N, 256, generator=torch.Generator().manual_seed(1))
    inst = Inst(pred_logits=logits.to(dev), pred_masks=masks.to(dev), output_embedding=emb.to(dev),
                obj_idxes=torch.full((N,), -1, dtype=torch.long, device=dev))
    fz = sv.PanopticFusion(**sv.FUSION_KWARGS)
    res = fz(inst, [(4 * h, 4 * w)], id=10001)
    r = O.panoptic_fuse(logits, masks, (4 * h, 4 * w), want_masks=True)
    assert res.masks.shape == (len(r.labels), 4 * h, 4 * w)
    np.testing.assert_array_equal(res.labels.cpu().numpy(), r.labels)
    np.testing.assert_allclose(res.probs.cpu().numpy(), r.probs, rtol=2e-6)
    np.testing.assert_array_equal(res.output_embedding.cpu().numpy(), emb[torch.from_numpy(r.keep.copy())].numpy())
    np.testing.assert_array_equal(res.pred_logits.cpu().numpy(), logits[torch.from_numpy(r.keep.copy())].numpy())
    assert int(((res.masks.cpu().numpy() != 0) != (r.masks != 0)).sum()) == 0
    assert float(np.abs(res.masks.cpu().numpy() - r.masks).max()) < 1e-4


def test_fusion_long_removal_cascade_resumes(dev):
    """the reference's small-segment loop is unbounded (vps_temporal_slots.py:761-792).  A cascade that needs
    more passes than the launched ones must not leave a stale id map: meta[3] = 0 + sentinel map, and host() resumes from
    the device state until the fixed point -- same result as the oracle."""
    N, h, w = 40, 16, 16
    logits = torch.full((N, 20), -4.0)
    logits[:, 19] = 4.0
    masks = torch.full((N, h, w), -30.0)
    # slot 0: stuff everywhere (weak); slots 1..8: things, each a 1-source-pixel bump that only wins once the previous
    # (stronger, overlapping) one has been removed for being too small -> one removal per pass
    logits[0] = -4.0; logits[0, 2] = 6.0
    masks[0] = 1.0
    for i in range(1, 9):
        logits[i] = -4.0; logits[i, 11 + (i % 8)] = 6.0 + 0.01 * i
    fz = sv.PanopticFusion(**{**sv.FUSION_KWARGS}, max_iters=1)
    lg, pm, _ = synthetic.make_fusion_case(9, N, h, w, n_things=10, near_dup_things=0, tiny=6)
    fo = fz.fuse(lg.to(dev), pm.to(dev), (4 * h, 4 * w))
    raw = fo.meta.cpu().numpy()
    r = O.panoptic_fuse(lg, pm, (4 * h, 4 * w))
    hst = fo.host()
    assert hst["converged"] and hst["iters"] == r.iters
    if r.iters > 1:
        assert int(raw[3]) == 0                                   # the single launched pass was not enough: resumed on the host
    np.testing.assert_array_equal(hst["keep"], r.keep)
    got = fo.panoptic.cpu().numpy()
    assert int(((got != r.panoptic) & ~r.near_tie).sum()) == 0
    print(f"fusion cascade: oracle iterations {r.iters}, device converged after resume: {hst['converged']}")
