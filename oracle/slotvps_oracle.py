"""CPU oracle for the Slot-VPS retriever hot path (TEST INFRASTRUCTURE, not product code).

A from-the-math restatement (torch-CPU / numpy, dtype selectable fp32|fp64) of the reference's

* ``MultiScaleDynamicMaskHead.forward``        mmdet/models/detectors/dynamic_mask_head.py:138-228
* ``MaskRCNNHead`` stage                        dynamic_mask_head.py:291-400
* ``MaskDynamicConv`` (Panoptic Retriever)      dynamic_mask_head.py:423-461
* ``TemporalSlotsHead``/``SlotsDynamicConv``    dynamic_mask_head.py:494-527, 550-572
* ``PositionEmbeddingSine``                     mmdet/models/detectors/position_encoding.py:236-256
* ``generate_final_outputs`` (mask logits)      mmdet/models/detectors/vps_temporal_slots.py:144-194
* ``PostProcessPanopticInstances.forward``      vps_temporal_slots.py:659-807 (+ mask_removal :564-657)
* inline panoptic fusion of ``simple_test``     vps_temporal_slots.py:411-435

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import this file.  ``slotvps_b200`` never does: the product path has no CPU fallback.

Parity pinning: the reference ships NO tests, golden vectors or fixtures for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference itself, imported in
place by ``oracle/ref_import.py`` and frozen as fixtures by ``tests/golden/make_golden.py``
(``tests/test_oracle_golden.py`` checks them on CPU; ``python tests/golden/make_golden.py --check``
regenerates every fixture from ``/root/reference`` and compares it with the committed file).

Parameters are passed as a flat ``dict[str, Tensor]`` with the reference's own ``state_dict``
key names (``head_series_{l}.{j}.*``, ``conv_trans.conv.*``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

LN_EPS = 1e-5
BN_EPS = 1e-5


@dataclass
class HeadConfig:
    """The knobs of configs/cityscapes/r50_fpn_slotvps.py:27-54 that shape the hot path."""
    dh_dim: int = 256
    num_classes: int = 20
    dim_feedforward: int = 2048
    nhead: int = 8
    per_dh_num_heads: Sequence[int] = (1, 2, 2, 2)
    trans_in_dim: int = 384
    temporal_dim_feedforward: int = 1024
    temporal_stages: Sequence[int] = (3, 4, 5, 6)
    num_cls: int = 2
    num_reg: int = 2
    activation: str = "gelu"             # stage FFN, r50_fpn_slotvps.py:33 ("relu" in swinL_fpn_slotvps.py:41)
    temporal_activation: str = "relu"    # Video Retriever FFN, r50 :49 ("gelu" in swinL :56)


@dataclass
class FusionConfig:
    """configs/cityscapes/r50_fpn_slotvps.py:66-74 + defaults of vps_temporal_slots.py:532-536."""
    num_classes: int = 20
    stuff_num: int = 11
    threshold: float = 0.85
    fraction_threshold: float = 0.03
    pixel_threshold: float = 0.4
    small_area: int = 4          # filter_small_option='4'  -> drop area <= 4


# ------------------------------------------------------------------------------------------
# small helpers
# ------------------------------------------------------------------------------------------
def _ln(x, w, b):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + LN_EPS) * w + b


def _lin(x, w, b=None):
    y = x @ w.t()
    return y if b is None else y + b


def _gelu_erf(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def _act(name):
    """_get_activation_fn, dynamic_mask_head.py:590-597: "relu" -> F.relu, "gelu" -> F.gelu (exact erf form)."""
    return {"relu": torch.relu, "gelu": _gelu_erf}[name]


def sine_position_embedding(h: int, w: int, dtype=torch.float32, num_pos_feats: int = 128,
                            temperature: float = 10000.0) -> torch.Tensor:
    """position_encoding.py:236-256 with normalize=True, scale=2*pi, all-False padding mask.

    Returns [1, 2*num_pos_feats, h, w]; channels [0,128) depend on the row only, [128,256) on
    the column only.  (Intermediates are fp32 in the reference: cumsum(dtype=float32).)
    """
    f32 = torch.float32
    ys = torch.arange(1, h + 1, dtype=f32)
    xs = torch.arange(1, w + 1, dtype=f32)
    ys = ys / (ys[-1] + 1e-6) * (2 * math.pi)
    xs = xs / (xs[-1] + 1e-6) * (2 * math.pi)
    i = torch.arange(num_pos_feats, dtype=f32)
    dim_t = temperature ** (2 * torch.div(i, 2, rounding_mode="floor") / num_pos_feats)
    py = ys[:, None] / dim_t          # [h,128]
    px = xs[:, None] / dim_t          # [w,128]
    even = (torch.arange(num_pos_feats) % 2 == 0)
    py = torch.where(even, py.sin(), py.cos())
    px = torch.where(even, px.sin(), px.cos())
    pos = torch.cat([py.t()[:, :, None].expand(-1, h, w), px.t()[:, None, :].expand(-1, h, w)], 0)
    return pos[None].to(dtype).contiguous()


# ------------------------------------------------------------------------------------------
# retriever head
# ------------------------------------------------------------------------------------------
def level_fuse(prev: Optional[torch.Tensor], x: torch.Tensor, W: torch.Tensor, b: torch.Tensor):
    """dynamic_mask_head.py:172-185.  prev [T,256,h/2,w/2] | None, x [T,128,h,w] -> [T,256,h,w].

    W is conv_trans.conv.weight viewed [256,384], b its bias.
    """
    if prev is None:
        z = torch.cat([x, x, x], 1)
    else:
        up = F.interpolate(prev, scale_factor=2, mode="bilinear", align_corners=False)
        z = torch.cat([up, x], 1)
    return torch.einsum("oc,tchw->tohw", W, z) + b[None, :, None, None]


def self_attention(s, P, pre, nhead):
    """nn.MultiheadAttention(256, 8) on [N,1,256] (dynamic_mask_head.py:346-352), no masks."""
    n, c = s.shape
    d = c // nhead
    qkv = _lin(s, P[pre + "self_attn.in_proj_weight"], P[pre + "self_attn.in_proj_bias"])
    q, k, v = qkv.split(c, -1)
    q = q.view(n, nhead, d).transpose(0, 1) * (1.0 / math.sqrt(d))
    k = k.view(n, nhead, d).transpose(0, 1)
    v = v.view(n, nhead, d).transpose(0, 1)
    a = torch.softmax(q @ k.transpose(1, 2), -1)
    o = (a @ v).transpose(0, 1).reshape(n, c)
    return _lin(o, P[pre + "self_attn.out_proj.weight"], P[pre + "self_attn.out_proj.bias"])


def pixel_attention(p, x, pos, P, pre, return_parts=False):
    """MaskDynamicConv.forward, dynamic_mask_head.py:423-461.

    p [N,C] slots, x [C,h,w] fused feature of one frame, pos [C,h,w] | None.  -> [N,C]
    Softmax is over the SLOT axis; the pixel reduction is a plain sum.
    """
    c = x.shape[0]
    xf = x.reshape(c, -1).t()                               # [P,C]
    xk = xf if pos is None else xf + pos.reshape(c, -1).t()
    pre = pre + "inst_interact."
    q = _ln(_lin(p, P[pre + "to_q.weight"], P[pre + "to_q.bias"]), P[pre + "norm_q.weight"], P[pre + "norm_q.bias"])
    k = _ln(_lin(xk, P[pre + "to_k.weight"], P[pre + "to_k.bias"]), P[pre + "norm_k.weight"], P[pre + "norm_k.bias"])
    v = _ln(_lin(xf, P[pre + "to_v.weight"], P[pre + "to_v.bias"]), P[pre + "norm_v.weight"], P[pre + "norm_v.bias"])
    logits = q @ k.t()                                      # [N,P]
    attn = torch.softmax(logits, 0)
    o = attn @ v                                            # [N,C] plain sum over pixels
    out = torch.relu(_ln(o, P[pre + "norm1.weight"], P[pre + "norm1.bias"]))
    if return_parts:
        return out, dict(q=q, logits=logits, o=o)
    return out


def stage_till_ffn(s, x, pos, P, pre, cfg: HeadConfig):
    """MaskRCNNHead.forward_till_ffn, dynamic_mask_head.py:342-388 (one frame)."""
    p = _ln(s + self_attention(s, P, pre, cfg.nhead), P[pre + "norm1.weight"], P[pre + "norm1.bias"])
    r = pixel_attention(p, x, pos, P, pre)
    p = _ln(p + r, P[pre + "norm2.weight"], P[pre + "norm2.bias"])
    f = _lin(_act(cfg.activation)(_lin(p, P[pre + "linear1.weight"], P[pre + "linear1.bias"])),
             P[pre + "linear2.weight"], P[pre + "linear2.bias"])
    return _ln(p + f, P[pre + "norm3.weight"], P[pre + "norm3.bias"])


def video_retriever(X, P, pre, activation="relu"):
    """TemporalSlotsHead.forward + SlotsDynamicConv.forward, dynamic_mask_head.py:494-527,550-572.

    X [T*N,C] -> [T*N,C] (WITHOUT the outer residual of :317).  softmax_dim="slots" means the
    softmax runs over the QUERY axis (dim=1 of [b,l,u]).
    """
    pre = pre + "temporal_query_head."
    ii = pre + "inst_interact."
    q = _ln(_lin(X, P[ii + "to_q.weight"], P[ii + "to_q.bias"]), P[ii + "norm_q.weight"], P[ii + "norm_q.bias"])
    k = _ln(_lin(X, P[ii + "to_k.weight"], P[ii + "to_k.bias"]), P[ii + "norm_k.weight"], P[ii + "norm_k.bias"])
    v = _ln(_lin(X, P[ii + "to_v.weight"], P[ii + "to_v.bias"]), P[ii + "norm_v.weight"], P[ii + "norm_v.bias"])
    a = torch.softmax(q @ k.t(), 0)                         # normalise over l (queries) per key u
    r = torch.relu(_ln(a @ v, P[ii + "norm1.weight"], P[ii + "norm1.bias"]))
    y = _ln(X + r, P[pre + "norm2.weight"], P[pre + "norm2.bias"])
    f = _lin(_act(activation)(_lin(y, P[pre + "linear1.weight"], P[pre + "linear1.bias"])),
             P[pre + "linear2.weight"], P[pre + "linear2.bias"])
    return _ln(y + f, P[pre + "norm3.weight"], P[pre + "norm3.bias"])


def towers(f, P, pre, cfg: HeadConfig):
    """MaskRCNNHead.forward_after_ffn, dynamic_mask_head.py:390-400.  f [N,C] -> ([N,20],[N,C])."""
    c = f
    for i in range(cfg.num_cls):
        c = torch.relu(_ln(_lin(c, P[pre + f"cls_module.{3 * i}.weight"]),
                           P[pre + f"cls_module.{3 * i + 1}.weight"], P[pre + f"cls_module.{3 * i + 1}.bias"]))
    e = f
    for i in range(cfg.num_reg):
        e = torch.relu(_ln(_lin(e, P[pre + f"reg_module.{3 * i}.weight"]),
                           P[pre + f"reg_module.{3 * i + 1}.weight"], P[pre + f"reg_module.{3 * i + 1}.bias"]))
    return _lin(c, P[pre + "class_logits.weight"], P[pre + "class_logits.bias"]), e


def head_forward(P: Dict[str, torch.Tensor], features: List[List[torch.Tensor]],
                 init_masks: List[torch.Tensor], pos: Optional[List[List[torch.Tensor]]],
                 cfg: HeadConfig = HeadConfig(), capture: Optional[dict] = None,
                 stage_slots_in: Optional[List[Optional[List[torch.Tensor]]]] = None):
    """MultiScaleDynamicMaskHead.forward, dynamic_mask_head.py:138-228 (bs == 1).

    features  T x L x [1,128,h_l,w_l];  init_masks T x [N,C];  pos T x L x [1,C,h_l,w_l] | None
    returns ( T x [S,1,N,num_classes],  T x [S,1,N,C],  T x L x [1,C,h_l,w_l] )
    ``capture`` (optional dict) receives per-stage inputs/outputs for teacher-forced parity.
    ``stage_slots_in[s][t]`` (optional) replaces the slots entering stage s of frame t (teacher forcing).
    """
    T = len(features)
    L = len(cfg.per_dh_num_heads)
    W = P["conv_trans.conv.weight"].reshape(cfg.dh_dim, -1)
    b = P["conv_trans.conv.bias"]
    slots = [m.clone() for m in init_masks]
    cls_all = [[] for _ in range(T)]
    emb_all = [[] for _ in range(T)]
    fused: List[torch.Tensor] = []
    stage = 0
    for l in range(L):
        x128 = torch.cat([features[t][l] for t in range(T)], 0)
        x = level_fuse(fused[l - 1] if l > 0 else None, x128, W, b)      # [T,256,h,w]
        fused.append(x)
        for j in range(cfg.per_dh_num_heads[l]):
            pre = f"head_series_{l}.{j}."
            if stage_slots_in is not None and stage_slots_in[stage] is not None:
                slots = [v.to(slots[t].dtype) if v is not None else slots[t] for t, v in enumerate(stage_slots_in[stage])]
            if capture is not None:
                capture[f"stage{stage}.slots_in"] = [s.clone() for s in slots]
            f = [stage_till_ffn(slots[t], x[t], None if pos is None else pos[t][l][0], P, pre, cfg)
                 for t in range(T)]
            if stage in cfg.temporal_stages:
                X = torch.cat(f, 0)
                Y = X + video_retriever(X, P, pre, cfg.temporal_activation)
                f = list(Y.split(Y.shape[0] // T, 0))
            for t in range(T):
                c, e = towers(f[t], P, pre, cfg)
                cls_all[t].append(c[None])
                emb_all[t].append(e[None])
                slots[t] = e
            if capture is not None:
                capture[f"stage{stage}.ffn_out"] = [v.clone() for v in f]
            stage += 1
    feats = [[fused[l][t:t + 1] for l in range(L)] for t in range(T)]
    return [torch.stack(c) for c in cls_all], [torch.stack(e) for e in emb_all], feats


# ------------------------------------------------------------------------------------------
# mask logits
# ------------------------------------------------------------------------------------------
def mask_logits(feat: torch.Tensor, emb: torch.Tensor, bn: Dict[str, torch.Tensor]) -> torch.Tensor:
    """generate_final_outputs, vps_temporal_slots.py:145-154.

    feat [C,h,w] finest fused feature (current frame), emb [N,C] last-stage mask embedding.
    bn: feat_bn.{weight,bias,running_mean,running_var} [C], fg_bn.{...} [1].  -> [N,h,w]
    """
    s = bn["feat_bn.weight"] / torch.sqrt(bn["feat_bn.running_var"] + BN_EPS)
    g = (feat - bn["feat_bn.running_mean"][:, None, None]) * s[:, None, None] + bn["feat_bn.bias"][:, None, None]
    g = g / g.norm(dim=0, keepdim=True).clamp_min(1e-12)
    m = torch.einsum("chw,nc->nhw", g, emb)
    sg = bn["fg_bn.weight"] / torch.sqrt(bn["fg_bn.running_var"] + BN_EPS)
    return (m - bn["fg_bn.running_mean"]) * sg + bn["fg_bn.bias"]


# ------------------------------------------------------------------------------------------
# panoptic fusion (post-process + inline relabel)
# ------------------------------------------------------------------------------------------
@dataclass
class FusionResult:
    panoptic: np.ndarray                 # [H,W] int64, the reference's panoptic_outputs[0]
    keep: np.ndarray                     # indices into the N slots, final order (stuff..., things...)
    labels: np.ndarray                   # class per kept slot
    probs: np.ndarray                    # score per kept slot (fp32)
    cls_inds: np.ndarray                 # panoptic_cls_inds  (thing classes - 10)
    cls_prob: np.ndarray                 # panoptic_cls_prob
    masks: Optional[np.ndarray] = None   # [K',H,W] fp32 masked logits (Instances.masks)
    near_tie: Optional[np.ndarray] = None  # [H,W] bool: top-2 gap of the final argmax < tol
    iters: int = 0
    info: dict = field(default_factory=dict)


def class_scores(pred_logits: torch.Tensor):
    """vps_temporal_slots.py:685: softmax over classes, max."""
    sc, cl = torch.softmax(pred_logits.float(), -1).max(-1)
    return sc.numpy(), cl.numpy()


def order_desc(scores: np.ndarray) -> np.ndarray:
    """The reference's ordering primitive, vps_temporal_slots.py:581 (argsort ascending, reversed)."""
    return np.argsort(scores)[::-1]


def upsample_masks(masks: torch.Tensor, size: Tuple[int, int]) -> torch.Tensor:
    """vps_temporal_slots.py:697-698: bilinear, align_corners=False, to (H,W)."""
    if tuple(masks.shape[-2:]) == tuple(size):
        return masks
    return F.interpolate(masks[:, None], size=size, mode="bilinear", align_corners=False)[:, 0]


def mask_removal(scores: np.ndarray, masks: np.ndarray, classes: np.ndarray, cfg: FusionConfig):
    """mask_removal, vps_temporal_slots.py:564-657 (apply_mask_removal_only_ins=True,
    use_mask_low_constant=False).  masks [K,H,W] fp32 logits.

    Returns (scores', masks', classes', keep_inds) ordered stuff (by score desc) then the things
    that survive (by score desc); a surviving thing keeps its logits only on the pixels it claims
    (prob >= pixel_threshold and not claimed by an earlier survivor), exactly 0 elsewhere.
    """
    K = masks.shape[0]
    if K == 0:
        raise ValueError("no kept slots (the reference raises here too: np.max of empty)")
    prob = torch.softmax(torch.from_numpy(masks), 0).numpy()
    order = order_desc(scores)
    is_stuff = classes[order] <= cfg.stuff_num - 1
    out_idx = [int(i) for i, s in zip(order, is_stuff) if s]
    out_masks = [masks[i] for i in out_idx]
    claimed = np.zeros(masks.shape[1:], dtype=bool)
    per_class_claim: Dict[int, np.ndarray] = {}
    for i, s in zip(order, is_stuff):
        if s:
            continue
        cand = prob[i] >= cfg.pixel_threshold
        n = int(cand.sum())
        if n == 0 or n == cand.size:                       # constant binarisation
            continue
        same = per_class_claim.get(int(classes[i]))
        ov = 0 if same is None else int(np.logical_and(same, cand).sum())
        if np.int64(ov) / np.float32(n) > cfg.fraction_threshold:   # int64/float32 -> float64, as numpy does at :621-622
            continue
        assign = cand & ~claimed
        m = np.zeros_like(masks[i])
        m[assign] = masks[i][assign]
        out_idx.append(int(i))
        out_masks.append(m)
        claimed |= assign
        per_class_claim[int(classes[i])] = assign if same is None else (same | assign)
    if not out_idx:
        raise ValueError("nothing survives mask_removal (the reference raises: np.stack of empty)")
    out_idx_a = np.asarray(out_idx, dtype=np.int64)
    return scores[out_idx_a], np.stack(out_masks, 0), classes[out_idx_a], out_idx_a


def _argmax_first(masks: np.ndarray) -> np.ndarray:
    """softmax over slots then argmax (vps_temporal_slots.py:728-734 / :417-418), first max wins."""
    if masks.shape[0] == 0:
        return np.zeros(masks.shape[1:], dtype=np.int64)
    p = torch.softmax(torch.from_numpy(masks), 0)
    return p.argmax(0).numpy()


def panoptic_fuse(pred_logits: torch.Tensor, pred_masks: torch.Tensor, size: Tuple[int, int],
                  cfg: FusionConfig = FusionConfig(), tie_tol: float = 1e-3,
                  want_masks: bool = False) -> FusionResult:
    """PostProcessPanopticInstances.forward (:659-807) followed by the inline fusion of
    simple_test (:411-435).  pred_logits [N,num_classes], pred_masks [N,h,w], size=(H,W).
    """
    scores, classes = class_scores(pred_logits)
    keep = np.nonzero((classes != cfg.num_classes - 1) & (scores > cfg.threshold))[0]
    masks = upsample_masks(pred_masks[torch.from_numpy(keep)].float(), size).numpy()
    sc, m, cl, kin = mask_removal(scores[keep], masks, classes[keep], cfg)
    idx = keep[kin]
    H, W = size
    # ---- area filter loop, :724-790 -------------------------------------------------------
    iters = 0
    first = True
    while True:
        ids = _argmax_first(m)
        if first:                                           # dedup only on the first call (:758)
            for c in np.unique(cl[cl <= cfg.stuff_num - 1]):
                same = np.nonzero(cl == c)[0]
                if len(same) > 1:
                    ids[np.isin(ids, same)] = same[0]
            first = False
        area = np.bincount(ids.ravel(), minlength=len(sc))[:len(sc)]
        iters += 1
        small = area <= cfg.small_area
        if len(sc) == 0 or not small.any():
            break
        sc, m, cl, idx = sc[~small], m[~small], cl[~small], idx[~small]
    # ---- inline fusion, :411-435 ------------------------------------------------------------
    thing = cl > cfg.stuff_num - 1
    reorder = np.concatenate([np.nonzero(~thing)[0], np.nonzero(thing)[0]])
    m2, sem = m[reorder], cl[reorder]
    ids = _argmax_first(m2)
    n_inst = int(thing.sum())
    n_all = len(cl)
    present = np.unique(ids)
    out = np.zeros((H, W), dtype=np.int64)
    count = n_inst
    for i in range(len(present) - 1, -1, -1):
        oid = present[i]
        if oid >= n_all - n_inst:
            out[ids == oid] = cfg.stuff_num + count - 1
            count -= 1
        else:
            out[ids == oid] = sem[i]                        # position in the unique list (quirk)
    near = None
    if m2.shape[0] >= 2:
        top2 = np.partition(m2, m2.shape[0] - 2, axis=0)[-2:]
        near = (top2[1] - top2[0]) < tie_tol * np.maximum(1.0, np.abs(top2[1]))
    else:
        near = np.zeros((H, W), dtype=bool)
    return FusionResult(panoptic=out, keep=idx, labels=cl, probs=sc,
                        cls_inds=cl[thing] - (cfg.stuff_num - 1), cls_prob=sc[thing],
                        masks=m if want_masks else None, near_tie=near, iters=iters)


# ------------------------------------------------------------------------------------------
# whole hot path for one clip (used by bench.py's CPU arm and the end-to-end tests)
# ------------------------------------------------------------------------------------------
def clip_forward(P, bn, features, init_query, size, hcfg=HeadConfig(), fcfg=FusionConfig(),
                 pos=None, fuse=True):
    """features T x L x [1,128,h,w]; returns dict(cls, emb, feats, pred_masks, fusion)."""
    T = len(features)
    if pos is None:
        pos = [[sine_position_embedding(f.shape[-2], f.shape[-1], f.dtype) for f in features[t]] for t in range(T)]
    cls, emb, feats = head_forward(P, features, [init_query] * T, pos, hcfg)
    pm = mask_logits(feats[-1][-1][0], emb[-1][-1, 0], bn)
    out = dict(cls=cls, emb=emb, feats=feats, pred_masks=pm)
    if fuse:
        out["fusion"] = panoptic_fuse(cls[-1][-1, 0], pm, size, fcfg)
    return out


# ------------------------------------------------------------------------------------------------
# Tracker (SURVEY.md 8f "next" rank 1): SimpleTrackHead + the greedy assignment loop of simple_test.
# ------------------------------------------------------------------------------------------------
def track_head_embed(fcs, x: torch.Tensor) -> torch.Tensor:
    """simple_track_head.py:63-79: `num_fcs_query` Linear layers, ReLU between (not after the last).
    `fcs` = [(weight [C,C], bias [C]), ...]."""
    x = x.float()
    for i, (w, b) in enumerate(fcs):
        x = x @ w.float().t() + b.float()
        if i < len(fcs) - 1:
            x = torch.relu(x)
    return x


def track_match_scores(fcs, cur: torch.Tensor, bank: torch.Tensor) -> torch.Tensor:
    """simple_track_head.py:88-90: [K, 1+M] = [0 | fc(cur) fc(bank)^T]; column 0 is the "new object" entry."""
    a, b = track_head_embed(fcs, cur), track_head_embed(fcs, bank)
    return torch.cat([torch.zeros(a.shape[0], 1), a @ b.t()], 1)


class TrackerState:
    """`prev_instances.output_embedding` of the reference (vps_temporal_slots.py:232-237): the raw slot embeddings
    of every object seen so far in this video; row index = object id."""

    def __init__(self):
        self.bank: Optional[np.ndarray] = None

    def reset(self):
        self.bank = None


def track_step(fcs, state: TrackerState, embedding: np.ndarray, labels: np.ndarray, stuff_num: int = 11):
    """vps_temporal_slots.py:322-409.  `embedding` [K,256] and `labels` [K] are the post-processed instances
    (stuff and things) in the post-processor's order.  Returns (det_obj_ids of the things [n_things] int32 -- the
    reference's `panoptic_det_obj_ids` -- , ids of all K entries, info)."""
    emb = np.asarray(embedding, np.float32)
    K = emb.shape[0]
    things = np.asarray(labels) > stuff_num - 1
    info = dict(new=0, matched=0, undone=0, lost=0)
    if state.bank is None:                                   # :335-342 first frame of the video
        state.bank = emb.copy()
        ids = np.arange(K, dtype=np.int64)
        return ids[things], ids, info
    score = track_match_scores(fcs, torch.from_numpy(emb), torch.from_numpy(state.bank))
    logp = torch.log_softmax(score, 1)
    lik, mid = logp.max(1)                                   # :348-351
    lik, mid = lik.numpy(), mid.numpy().astype(np.int32)
    M0 = state.bank.shape[0]
    bank = [r for r in state.bank]
    ids = -np.ones(K, np.int32)
    best = -100.0 * np.ones(M0)
    best_id = -np.ones(M0, np.int32)
    for i in range(K):                                       # :359-392
        if mid[i] == 0:
            ids[i] = len(bank)
            bank.append(emb[i])
            info["new"] += 1
        else:
            o = mid[i] - 1
            if lik[i] > best[o]:
                ids[i] = o
                if best_id[o] >= 0:
                    ids[best_id[o]] = -1
                    info["undone"] += 1
                best[o], best_id[o] = lik[i], i
                bank[o] = emb[i]
                info["matched"] += 1
            else:
                info["lost"] += 1
    for i in range(K):                                       # :396-404 redundant matches become new objects
        if ids[i] < 0:
            ids[i] = len(bank)
            bank.append(emb[i])
    state.bank = np.stack(bank).astype(np.float32)
    info["margin"] = float(np.sort(score.numpy(), 1)[:, -1].min() - 0) if K else 0.0
    top2 = np.sort(score.numpy(), 1)[:, -2:] if score.shape[1] > 1 else None
    info["top2_gap"] = float((top2[:, 1] - top2[:, 0]).min()) if top2 is not None and K else float("inf")
    return ids[things], ids, info


def input_transform(features, weight: torch.Tensor, bias: torch.Tensor):
    """semantic_trans_ins, vps_temporal_slots.py:129-135: VPS_Capsule.conv_trans (1x1 conv 128->128 + bias, no norm,
    no activation) on every level.  features T x L x [1,128,h,w]."""
    w = weight.reshape(weight.shape[0], -1)
    return [[torch.einsum("oc,bchw->bohw", w.to(f.dtype), f) + bias.to(f.dtype)[None, :, None, None] for f in fr] for fr in features]


# ------------------------------------------------------------------------------------------------
# Consumers of the id map: semantic argmax of simple_test and the per-frame merge
# CityscapesVps.get_unified_pan_result that produces the evaluation wire format (H x W x 3 uint8).
# ------------------------------------------------------------------------------------------------
def semantic_argmax(fcn_output: torch.Tensor, size: Tuple[int, int]) -> torch.Tensor:
    """vps_temporal_slots.py:440-451: bilinear resize to the image size when needed (align_corners=False), softmax
    over classes, index of the max.  fcn_output [1,Cs,h,w] -> [1,H,W] int64."""
    x = fcn_output.float()
    if tuple(x.shape[-2:]) != tuple(size):
        x = F.interpolate(x, size=tuple(size), mode="bilinear", align_corners=False)
    return torch.softmax(x, 1).max(1)[1]


def dedup_obj_ids(obj_id: np.ndarray, max_oid: int):
    """tools/dataset/cityscapes_vps.py:233-244: of several instances carrying one object id the LAST keeps it, the
    earlier ones (walking backwards) receive fresh ids from a counter that persists over the frames of a call."""
    obj_id = np.array(obj_id).copy()
    uniq, cnt = np.unique(obj_id, return_counts=True)
    if not np.any(cnt > 1):
        return obj_id, max_oid
    rev = obj_id[::-1].copy()
    for red in uniq[cnt > 1]:
        n = int((obj_id == red).sum())
        repl = np.full(n, red, dtype=obj_id.dtype)
        for i in range(1, n):
            repl[i] = max_oid
            max_oid += 1
        rev[rev == red] = repl
    return rev[::-1], max_oid


def unify_pan_result(segs, pans, cls_inds, obj_ids=None, stuff_area_limit: int = 4 * 64 * 64, id_last_stuff: int = 10):
    """CityscapesVps.get_unified_pan_result, tools/dataset/cityscapes_vps.py:214-302, one [H,W,3] uint8 array per frame
    (semantic label, instance index from 1, object id + 1).  id_last_stuff = num_seg_classes - num_classes (:250)."""
    if obj_ids is None:
        obj_ids = [None] * len(cls_inds)
    max_oid = 100
    out = []
    for seg, pan, cls_ind, obj_id in zip(segs, pans, cls_inds, obj_ids):
        seg, pan = np.asarray(seg), np.array(pan).copy()
        if obj_id is not None:
            obj_id, max_oid = dedup_obj_ids(np.asarray(obj_id), max_oid)
        pan_seg = pan.copy()
        if len(cls_ind) == 0:
            pan[pan > id_last_stuff] = 255
        pan_ins, pan_obj = pan.copy(), pan.copy()
        present = np.unique(pan)
        present = present[present > id_last_stuff]
        pan_ins[pan_ins <= id_last_stuff] = 0
        for idx, i in enumerate(present):
            region = pan == i                       # == (pan_ins == i): values written below never collide with a later id
            if i == 255:
                pan_seg[region] = 255
                pan_ins[region] = 0
                continue
            cls, cnt = np.unique(seg[region], return_counts=True)
            major = cls[np.argmax(cnt)]
            inst_cls = cls_ind[i - id_last_stuff - 1] + id_last_stuff
            if major != inst_cls and np.max(cnt) / np.sum(cnt) >= 0.5 and major <= id_last_stuff:
                pan_seg[region] = major             # the semantic head out-votes the instance: becomes stuff
                pan_ins[region] = 0
                pan_obj[region] = 0
            else:
                pan_seg[region] = inst_cls
                pan_ins[region] = idx + 1
                if obj_id is not None:
                    pan_obj[region] = obj_id[idx] + 1
        for v in np.unique(pan_seg):
            if v <= id_last_stuff:
                area = pan_seg == v
                if area.sum() < stuff_area_limit:
                    pan_seg[area] = 255
        o = np.zeros(pan.shape + (3,), dtype=np.uint8)
        o[:, :, 0], o[:, :, 1], o[:, :, 2] = pan_seg.astype(np.uint8), pan_ins.astype(np.uint8), pan_obj.astype(np.uint8)
        out.append(o)
    return out


# ---- UPSNetFPN deformable-convolution subnet ----------------------------------------
# Pinning: the reference op is CUDA-only, so it cannot run in the build container.  It is compiled UNMODIFIED from its own
# two source files by oracle/build_ref_dcn.sh into oracle/_ref/deform_conv_cuda.so (sm_100a), run on the B200 by
# tests/golden/make_golden_dcn.py, and its outputs are frozen as tests/golden/dcn_*.npz; tests/test_oracle_golden.py
# checks this restatement against them on CPU (and against torchvision.ops.deform_conv2d, an independent implementation).
def deform_conv(x: torch.Tensor, offset: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """3x3 / stride 1 / padding 1 / dilation 1 deformable convolution, one group, one deformable group.

    ``deformable_im2col_gpu_kernel`` (mmdet/ops/dcn/src/deform_conv_cuda_kernel.cu:190-236) with
    ``deformable_im2col_bilinear`` (:80-112), then ``weight.flatten(1) @ columns`` (deform_conv_cuda.cpp:214-219).
    x [B,Cin,H,W], offset [B,18,H,W] (channel 2*tap = dy, 2*tap+1 = dx, tap = 3*ky+kx), weight [Cout,Cin,3,3]."""
    B, Cin, H, W = x.shape
    dt = x.dtype
    flat = x.reshape(B, Cin, H * W)
    ys = torch.arange(H, dtype=dt).view(1, H, 1)
    xs = torch.arange(W, dtype=dt).view(1, 1, W)
    cols = []
    for tap in range(9):
        i, j = tap // 3, tap % 3
        h_im = (ys - 1 + i) + offset[:, 2 * tap].to(dt)               # :222-223 (h_in + i * dilation_h + offset_h)
        w_im = (xs - 1 + j) + offset[:, 2 * tap + 1].to(dt)
        inside = (h_im > -1) & (w_im > -1) & (h_im < H) & (w_im < W)  # :224
        h_low, w_low = torch.floor(h_im), torch.floor(w_im)
        lh, lw = h_im - h_low, w_im - w_low
        hh, hw = 1 - lh, 1 - lw
        h_low, w_low = h_low.long(), w_low.long()
        h_high, w_high = h_low + 1, w_low + 1

        def corner(hi, wi, ok):
            idx = (hi.clamp(0, H - 1) * W + wi.clamp(0, W - 1)).view(B, 1, H * W).expand(B, Cin, H * W)
            v = torch.gather(flat, 2, idx)
            return v * (ok & inside).view(B, 1, H * W).to(dt)
        v1 = corner(h_low, w_low, (h_low >= 0) & (w_low >= 0))                     # :93-104
        v2 = corner(h_low, w_high, (h_low >= 0) & (w_high <= W - 1))
        v3 = corner(h_high, w_low, (h_high <= H - 1) & (w_low >= 0))
        v4 = corner(h_high, w_high, (h_high <= H - 1) & (w_high <= W - 1))
        w1, w2, w3, w4 = (hh * hw), (hh * lw), (lh * hw), (lh * lw)                 # :106
        f = lambda t: t.view(B, 1, H * W)
        cols.append(f(w1) * v1 + f(w2) * v2 + f(w3) * v3 + f(w4) * v4)              # :108
    col = torch.stack(cols, dim=2).reshape(B, Cin * 9, H * W)                        # row c * 9 + tap, as :204 (c_col)
    out = torch.matmul(weight.reshape(weight.shape[0], -1).to(dt), col)
    return out.reshape(B, -1, H, W)


def deform_conv_with_offset(x, offset_w, offset_b, weight):
    """DeformConvWithOffset.forward (mmdet/models/utils/deform_conv_with_offset.py:38-39)."""
    return deform_conv(x, F.conv2d(x, offset_w, offset_b, padding=1), weight)


def dcn_subnet(P: Dict[str, torch.Tensor], x: torch.Tensor, n_layers: int = 3, capture: Optional[list] = None) -> torch.Tensor:
    """``UPSNetFPN.deform_convs[0]`` (upsnetFPN.py:36-49): n x [DeformConvWithOffset -> GroupNorm(32) -> ReLU].
    P uses the Sequential's own keys ("0.conv_offset.weight", "0.conv.weight", "1.weight", ...).  ``capture`` collects
    the input of every layer (teacher forcing)."""
    for i in range(n_layers):
        if capture is not None:
            capture.append(x)
        y = deform_conv_with_offset(x, P[f"{3 * i}.conv_offset.weight"], P[f"{3 * i}.conv_offset.bias"], P[f"{3 * i}.conv.weight"])
        x = F.relu(F.group_norm(y, 32, P[f"{3 * i + 1}.weight"], P[f"{3 * i + 1}.bias"], 1e-5))
    return x
