"""Seeded synthetic weights and inputs for the retriever hot path.

Checkpoints and datasets are unavailable offline (reference README.md:24), so every run uses
random-init weights of the architecture named by configs/cityscapes/r50_fpn_slotvps.py and
synthetic Cityscapes-VPS-shaped features at the head boundary.

The state_dict produced here has the reference's exact key names and shapes
(``head_series_{l}.{j}.*`` / ``conv_trans.conv.*``, dynamic_mask_head.py:72-113), so it loads
strictly into the reference head, the oracle and ``B200DynamicMaskHead`` alike.  It is produced
by a CPU ``torch.Generator`` so tests regenerate it from the seed instead of storing 62 MB.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch

C = 256


def _uniform(g, shape, bound):
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound


def _linear(g, sd, name, out_f, in_f, bias=True, gain=1.0):
    bound = gain * math.sqrt(6.0 / (in_f + out_f))       # xavier-uniform, as _reset_parameters (:127-131)
    sd[name + ".weight"] = _uniform(g, (out_f, in_f), bound)
    if bias:
        sd[name + ".bias"] = _uniform(g, (out_f,), 0.05)


def _norm(g, sd, name, dim=C):
    sd[name + ".weight"] = 1.0 + _uniform(g, (dim,), 0.1)
    sd[name + ".bias"] = _uniform(g, (dim,), 0.05)


def _retriever(g, sd, pre):
    for n in ("to_q", "to_k", "to_v"):
        _linear(g, sd, pre + n, C, C)
    for n in ("norm_q", "norm_k", "norm_v", "norm1"):
        _norm(g, sd, pre + n)


def make_head_state_dict(seed: int = 0, per_dh_num_heads: Sequence[int] = (1, 2, 2, 2),
                         temporal_stages: Sequence[int] = (3, 4, 5, 6), num_classes: int = 20,
                         dim_feedforward: int = 2048, temporal_dim_feedforward: int = 1024,
                         trans_in_dim: int = 384, num_cls: int = 2, num_reg: int = 2
                         ) -> Dict[str, torch.Tensor]:
    """Random-init parameters of MultiScaleDynamicMaskHead, reference key names and order."""
    g = torch.Generator().manual_seed(10_000 + seed)
    sd: Dict[str, torch.Tensor] = {}
    stage = 0
    for l, nh in enumerate(per_dh_num_heads):
        # the reference decides per LEVEL whether its stages own a temporal head (:89)
        level_temporal = stage in temporal_stages
        for j in range(nh):
            pre = f"head_series_{l}.{j}."
            sd[pre + "self_attn.in_proj_weight"] = _uniform(g, (3 * C, C), math.sqrt(6.0 / (4 * C)))
            sd[pre + "self_attn.in_proj_bias"] = _uniform(g, (3 * C,), 0.05)
            _linear(g, sd, pre + "self_attn.out_proj", C, C)
            _retriever(g, sd, pre + "inst_interact.")
            _linear(g, sd, pre + "linear1", dim_feedforward, C)
            _linear(g, sd, pre + "linear2", C, dim_feedforward)
            for n in ("norm1", "norm2", "norm3"):
                _norm(g, sd, pre + n)
            if level_temporal:
                tp = pre + "temporal_query_head."
                _retriever(g, sd, tp + "inst_interact.")
                _linear(g, sd, tp + "linear1", temporal_dim_feedforward, C)
                _linear(g, sd, tp + "linear2", C, temporal_dim_feedforward)
                for n in ("norm1", "norm2", "norm3"):
                    _norm(g, sd, tp + n)
            for i in range(num_cls):
                _linear(g, sd, pre + f"cls_module.{3 * i}", C, C, bias=False)
                _norm(g, sd, pre + f"cls_module.{3 * i + 1}")
            for i in range(num_reg):
                _linear(g, sd, pre + f"reg_module.{3 * i}", C, C, bias=False)
                _norm(g, sd, pre + f"reg_module.{3 * i + 1}")
            _linear(g, sd, pre + "class_logits", num_classes, C)
            stage += 1
    sd["conv_trans.conv.weight"] = _uniform(g, (C, trans_in_dim, 1, 1), math.sqrt(6.0 / (trans_in_dim + C)))
    sd["conv_trans.conv.bias"] = _uniform(g, (C,), 0.05)
    return sd


def make_capsule_params(seed: int = 0, n_slots: int = 100) -> Dict[str, torch.Tensor]:
    """Parameters the retriever borrows from VPS_Capsule (vps_capsule.py:71-72, 96-97):
    init_mask_query.weight [N,256], feat_bn (BatchNorm2d(256) eval), fg_bn (BatchNorm2d(1) eval)."""
    g = torch.Generator().manual_seed(20_000 + seed)
    p = {"init_mask_query.weight": _uniform(g, (n_slots, C), math.sqrt(6.0 / (n_slots + C)))}
    p["feat_bn.weight"] = 1.0 + _uniform(g, (C,), 0.1)
    p["feat_bn.bias"] = _uniform(g, (C,), 0.05)
    p["feat_bn.running_mean"] = _uniform(g, (C,), 0.1)
    p["feat_bn.running_var"] = 1.0 + _uniform(g, (C,), 0.2)
    p["fg_bn.weight"] = torch.tensor([0.1]) + _uniform(g, (1,), 0.01)
    p["fg_bn.bias"] = _uniform(g, (1,), 0.05)
    p["fg_bn.running_mean"] = _uniform(g, (1,), 0.05)
    p["fg_bn.running_var"] = 1.0 + _uniform(g, (1,), 0.2)
    return p


def level_shapes(H: int, W: int, n_levels: int = 4) -> List[Tuple[int, int]]:
    """Feature sizes at strides 32,16,8,4 (coarse -> fine) of an HxW input padded to /32."""
    Hp, Wp = (H + 31) // 32 * 32, (W + 31) // 32 * 32
    return [(Hp // s, Wp // s) for s in (32, 16, 8, 4)][:n_levels]


def make_features(H: int, W: int, T: int = 2, video: int = 0, frame: int = 0,
                  shapes: Sequence[Tuple[int, int]] = None) -> List[List[torch.Tensor]]:
    """T x 4 x [1,128,h_l,w_l] N(0,1) features (the 128-ch output of semantic_trans_ins),
    seeded by 1000*video + frame (+ frame offset inside the clip)."""
    shapes = shapes or level_shapes(H, W)
    out = []
    for t in range(T):
        g = torch.Generator().manual_seed(1000 * video + frame + 7919 * t + 1)
        out.append([torch.randn((1, 128, h, w), generator=g, dtype=torch.float32) for (h, w) in shapes])
    return out


def make_fusion_case(seed: int, n_slots: int, h: int, w: int, n_stuff: int = 11, n_things: int = 12,
                     dup_stuff: int = 2, near_dup_things: int = 3, tiny: int = 2,
                     num_classes: int = 20, stuff_num: int = 11):
    """A designed (pred_logits [N,20], pred_masks [N,h,w]) pair for the fusion stage.

    Random-init heads rarely keep any thing slot, so the fusion
    parity cases are built at the fusion boundary: blobby low-frequency mask logits, ``n_stuff``
    stuff slots (``dup_stuff`` of them repeating a class), ``n_things`` confident thing slots of
    which ``near_dup_things`` overlap an earlier same-class thing, ``tiny`` slots whose region is a
    few pixels, and the rest below the 0.85 score threshold or predicted no-object.
    """
    g = torch.Generator().manual_seed(30_000 + seed)
    N = n_slots
    # low-frequency fields: coarse noise upsampled bilinearly
    ch, cw = max(2, h // 8), max(2, w // 8)
    coarse = torch.randn((N, 1, ch, cw), generator=g)
    masks = torch.nn.functional.interpolate(coarse, size=(h, w), mode="bilinear", align_corners=False)[:, 0]
    masks = masks * 6.0 + torch.randn((N, h, w), generator=g) * 0.3
    logits = torch.randn((N, num_classes), generator=g) * 0.5
    logits[:, num_classes - 1] += 3.0                     # default: no-object
    perm = torch.randperm(N, generator=g)
    k = 0
    roles = {}
    # confident-class logits are all distinct and well separated so that the score ORDER (which the
    # reference takes from np.argsort, vps_temporal_slots.py:581) is not decided by 1-ulp differences
    n_conf = n_stuff + n_things + near_dup_things + tiny
    strength = (2.6 + 0.06 * torch.randperm(n_conf, generator=g).float()).tolist()
    for i in range(n_stuff):
        s = int(perm[k]); k += 1
        cls = i % stuff_num if i < n_stuff - dup_stuff else int(torch.randint(0, max(1, n_stuff - dup_stuff), (1,), generator=g))
        logits[s] = -4.0
        logits[s, cls] = strength.pop()
        roles[s] = ("stuff", cls)
    things = []
    for i in range(n_things):
        s = int(perm[k]); k += 1
        cls = stuff_num + int(torch.randint(0, num_classes - 1 - stuff_num, (1,), generator=g))
        logits[s] = -4.0
        logits[s, cls] = strength.pop()
        # give things a compact blob so that prob >= 0.4 somewhere
        cy, cx = int(torch.randint(0, h, (1,), generator=g)), int(torch.randint(0, w, (1,), generator=g))
        yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
        r = 2.0 + float(torch.rand(1, generator=g)) * min(h, w) / 6
        masks[s] = masks[s] * 0.3 + 14.0 * torch.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * r * r)) - 2.0
        things.append((s, cls))
        roles[s] = ("thing", cls)
    for i in range(min(near_dup_things, len(things))):
        s = int(perm[k]); k += 1
        src, cls = things[i]
        logits[s] = -4.0
        logits[s, cls] = strength.pop()
        masks[s] = masks[src] * (0.9 + 0.2 * float(torch.rand(1, generator=g))) + torch.randn((h, w), generator=g) * 0.5
        roles[s] = ("dup", cls)
    for i in range(tiny):
        s = int(perm[k]); k += 1
        cls = stuff_num + int(torch.randint(0, num_classes - 1 - stuff_num, (1,), generator=g))
        logits[s] = -4.0
        logits[s, cls] = strength.pop()
        masks[s] = -8.0
        cy, cx = int(torch.randint(0, h, (1,), generator=g)), int(torch.randint(0, w, (1,), generator=g))
        masks[s, cy, cx] = 30.0
        roles[s] = ("tiny", cls)
    return logits.contiguous(), masks.contiguous(), roles


def make_track_params(seed: int, channels: int = 256, num_fcs: int = 2, mode: str = "identity"):
    """Synthetic SimpleTrackHead weights under the reference's parameter names
    (simple_track_head.py:44-50: ``fcs_query.{i}.weight/bias``).  ``identity``: identity-dominant layers so that
    the designed identities of make_track_sequence survive (every branch of the greedy loop is reached);
    ``random``: dense random layers with an active ReLU (exercises the arithmetic, arbitrary matches)."""
    g = torch.Generator().manual_seed(50_000 + seed)
    sd = {}
    for i in range(num_fcs):
        if mode == "identity":
            w = torch.eye(channels) + torch.randn((channels, channels), generator=g) * 0.001
            b = torch.randn((channels,), generator=g) * 0.002 + (0.1 if i == 0 else -0.1)
        else:
            w = torch.randn((channels, channels), generator=g) * 0.08
            b = torch.randn((channels,), generator=g) * 0.1
        sd["fcs_query.%d.weight" % i] = w
        sd["fcs_query.%d.bias" % i] = b
    return sd


def make_track_sequence(seed: int, n_slots: int, frames: int, channels: int = 256, blk: int = 2,
                        p_new: float = 0.10, p_dup: float = 0.12, amp: float = 10.0, delta: float = 0.04,
                        noise: float = 0.02):
    """Per-frame slot embeddings [frames][N,channels] with designed identities.

    Identity j is ``amp`` on its own block of ``blk`` channels and ``-delta`` elsewhere: self dot ~ blk*amp^2,
    cross dot ~ -2*blk*amp*delta + channels*delta^2 < 0.  A slot re-observing an identity matches its bank row, a
    slot with an unseen identity prefers the all-zero "new object" column (simple_track_head.py:89), and two slots
    sharing an identity in one frame reach the undo branch of the greedy loop (vps_temporal_slots.py:373-381)."""
    g = torch.Generator().manual_seed(60_000 + seed)
    n_ident = channels // blk
    protos = -delta * torch.ones((n_ident, channels))
    for j in range(n_ident):
        protos[j, j * blk:(j + 1) * blk] = amp
    order = torch.randperm(n_ident, generator=g).tolist()
    ident = [order[s % n_ident] for s in range(n_slots)]
    nxt = n_slots
    out = []
    for f in range(frames):
        if f > 0:
            for s in range(n_slots):
                r = float(torch.rand(1, generator=g))
                if r < p_new and nxt < n_ident:
                    ident[s] = order[nxt]
                    nxt += 1
                elif r < p_new + p_dup:
                    ident[s] = ident[int(torch.randint(0, n_slots, (1,), generator=g))]
        e = protos[torch.tensor(ident)] + torch.randn((n_slots, channels), generator=g) * noise
        out.append(e.contiguous())
    return out


def make_in_trans_params(seed: int, channels: int = 128):
    """Synthetic VPS_Capsule.conv_trans (ConvModule(128,128,1, activation=None), vps_capsule.py:74-79) under the
    reference's parameter names: the 1x1 transform semantic_trans_ins applies to every level before the head."""
    g = torch.Generator().manual_seed(70_000 + seed)
    return {"conv_trans.conv.weight": torch.randn((channels, channels, 1, 1), generator=g) * (1.2 / channels ** 0.5),
            "conv_trans.conv.bias": torch.randn((channels,), generator=g) * 0.2}


def make_unify_case(seed: int, H: int, W: int, n_inst: int = 12, dup_obj: int = 3, hidden: int = 2,
                    n_stuff: int = 11, n_sem: int = 19):
    """One frame of inputs for the id-map merge (get_unified_pan_result): semantic argmax ``seg`` [H,W], panoptic ids
    ``pan`` [H,W] (stuff label < n_stuff, instance i -> n_stuff + i), ``cls_ind`` [n_inst] (thing class - 10, 1..8)
    and ``obj_id`` [n_inst].  Designed to reach every branch: instances whose majority semantic class agrees, ones
    out-voted by a stuff class, ones contradicted by another thing class or by a sub-majority, instances with no pixel
    (``hidden``: later ones cover them, so position-in-list and id-based indexing differ), duplicate object ids and
    stuff classes below the area limit."""
    g = torch.Generator().manual_seed(80_000 + seed)
    coarse = torch.randn((1, n_stuff, max(2, H // 32), max(2, W // 32)), generator=g)
    field = torch.nn.functional.interpolate(coarse, size=(H, W), mode="bilinear", align_corners=False)[0]
    field[n_stuff - 2:] -= 1.5                                   # two stuff classes end up small
    pan = field.argmax(0).to(torch.int64)
    seg = pan.clone()
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    cls_ind = torch.randint(1, n_sem - n_stuff + 1, (n_inst,), generator=g)
    centres = []
    for i in range(n_inst):
        cy, cx = int(torch.randint(0, H, (1,), generator=g)), int(torch.randint(0, W, (1,), generator=g))
        ry = 3 + int(torch.randint(0, max(4, H // 6), (1,), generator=g))
        rx = 3 + int(torch.randint(0, max(4, W // 6), (1,), generator=g))
        centres.append((cy, cx, ry, rx))
    hidden_idx = [j for j in (2, 5, 7)[:hidden] if j + 1 < n_inst]
    for j in hidden_idx:
        centres[j] = centres[j + 1]                              # same ellipse as the next instance, painted before it
    order = hidden_idx + [i for i in range(n_inst) if i not in hidden_idx]
    for i in order:
        cy, cx, ry, rx = centres[i]
        region = ((yy - cy).float() / ry) ** 2 + ((xx - cx).float() / rx) ** 2 <= 1.0
        pan[region] = n_stuff + i
    for i in range(n_inst):
        region = pan == n_stuff + i
        if not bool(region.any()):
            continue
        mode = i % 4
        inst_cls = int(cls_ind[i]) + n_stuff - 1
        if mode == 0 or mode == 1:
            seg[region] = inst_cls                                # agreement (plus a stuff fringe for mode 1)
            if mode == 1:
                seg[region & (xx % 3 == 0)] = 2
        elif mode == 2:
            seg[region] = 5                                       # stuff majority >= 0.5: the instance is dropped
            seg[region & (xx % 4 == 0)] = inst_cls
        else:
            other = n_stuff + (int(cls_ind[i]) % (n_sem - n_stuff))
            seg[region] = other                                   # another thing class: instance label wins
            seg[region & (yy % 3 == 0)] = 1
    obj_id = torch.randperm(40, generator=g)[:n_inst].to(torch.int64)
    for k in range(dup_obj):
        a, b = int(torch.randint(0, n_inst, (1,), generator=g)), int(torch.randint(0, n_inst, (1,), generator=g))
        obj_id[a] = obj_id[b]
    return seg, pan, cls_ind.to(torch.int64), obj_id


def make_dcn_state_dict(seed: int = 0, in_channels: int = 256, out_channels: int = 128, offset_scale: float = 1.0
                        ) -> Dict[str, torch.Tensor]:
    """Random-init parameters of ``UPSNetFPN.deform_convs[0]`` (mmdet/models/panoptic/upsnetFPN.py:36-49) under the
    Sequential's own keys.  The reference zero-initialises ``conv_offset`` (deform_conv_with_offset.py:25-26), which would
    make every sampling offset zero; here it gets small weights so that the offsets reach ~``offset_scale`` pixels with
    O(1) inputs (trained checkpoints are unavailable offline), fractional positions, and out-of-map taps at the border."""
    g = torch.Generator().manual_seed(90_000 + seed)
    sd: Dict[str, torch.Tensor] = {}
    chans = [(in_channels, in_channels), (in_channels, out_channels), (out_channels, out_channels)]
    for i, (cin, cout) in enumerate(chans):
        n = cin * 9
        sd[f"{3 * i}.conv_offset.weight"] = _uniform(g, (18, cin, 3, 3), offset_scale * math.sqrt(3.0 / n))
        sd[f"{3 * i}.conv_offset.bias"] = _uniform(g, (18,), 0.5 * offset_scale)
        sd[f"{3 * i}.conv.weight"] = _uniform(g, (cout, cin, 3, 3), 1.0 / math.sqrt(n))      # DeformConv.reset_parameters
        sd[f"{3 * i + 1}.weight"] = 1.0 + _uniform(g, (cout,), 0.1)
        sd[f"{3 * i + 1}.bias"] = _uniform(g, (cout,), 0.05)
    return sd


def make_fpn_level(seed: int, B: int, channels: int, H: int, W: int) -> torch.Tensor:
    """One FPN level as the UPSNetFPN subnet sees it: smooth low-frequency structure plus noise, O(1) magnitude."""
    g = torch.Generator().manual_seed(95_000 + seed)
    coarse = torch.randn((B, channels, max(H // 8, 1), max(W // 8, 1)), generator=g)
    x = torch.nn.functional.interpolate(coarse, size=(H, W), mode="bilinear", align_corners=False)
    return (x + 0.3 * torch.randn((B, channels, H, W), generator=g)).contiguous()
