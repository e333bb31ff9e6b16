"""Mask-logit projection, panoptic fusion and the whole-clip retriever on top of the C ABI.

Host-side mirrors of
* ``VPS_Temporal_Slots.generate_final_outputs``       vps_temporal_slots.py:144-194
* ``PostProcessPanopticInstances``                    vps_temporal_slots.py:528-807
* the inline panoptic fusion of ``simple_test``       vps_temporal_slots.py:411-435
All arithmetic runs in libslotvps_b200.so; nothing here falls back to torch ops or the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import nn

from . import _lib
from .head import B200DynamicMaskHead, _stream_ptr


def _need_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (slotvps_b200 has no CPU fallback)")


def sine_position_embedding(h: int, w: int, device) -> torch.Tensor:
    """PositionEmbeddingSine (position_encoding.py:236-256, normalize=True) -> [1,256,h,w]."""
    out = torch.empty((1, 256, h, w), dtype=torch.float32, device=device)
    _lib.check(_lib.lib().slotvps_sine_pos(out.data_ptr(), h, w, _stream_ptr(out.device)), "slotvps_sine_pos")
    return out


def level_fuse(prev: Optional[torch.Tensor], x: torch.Tensor, conv_w: torch.Tensor, conv_b: torch.Tensor) -> torch.Tensor:
    """dynamic_mask_head.py:172-185 for one frame: prev [256,h/2,w/2] | None, x [128,h,w] -> [256,h,w]."""
    _need_cuda(x, "x")
    h, w = x.shape[-2:]
    out = torch.empty((256, h, w), dtype=torch.float32, device=x.device)
    scratch = torch.empty(256 * max(128, (h // 2) * (w // 2)), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().slotvps_level_fuse(None if prev is None else prev.contiguous().data_ptr(), x.contiguous().data_ptr(),
                                             conv_w.contiguous().data_ptr(), conv_b.contiguous().data_ptr(), out.data_ptr(),
                                             h, w, scratch.data_ptr(), _stream_ptr(x.device)), "slotvps_level_fuse")
    return out


def slot_attention(stage_params: Dict[str, torch.Tensor], slots_p: torch.Tensor, x: torch.Tensor,
                   pos: Optional[torch.Tensor], kernel_path: int = 0) -> torch.Tensor:
    """MaskDynamicConv.forward (dynamic_mask_head.py:423-461) for one frame.

    stage_params: tensors keyed by the state_dict suffixes of one stage (``inst_interact.to_q.weight`` ...).
    slots_p [N,256], x [256,h,w], pos [256,h,w] | None -> [N,256]."""
    _need_cuda(x, "x")
    sp = _lib.StageParams()
    keep = []
    for f in _lib.STAGE_FIELDS:
        t = stage_params.get(_lib.STAGE_KEYS[f])
        if t is not None:
            t = t.contiguous()
            keep.append(t)
            setattr(sp, f, t.data_ptr())
    N = slots_p.shape[0]
    h, w = x.shape[-2:]
    nbytes = C.c_size_t()
    L = _lib.lib()
    _lib.check(L.slotvps_slot_attention_workspace_bytes(N, h, w, C.byref(nbytes)), "slot_attention_workspace_bytes")
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=x.device)
    out = torch.empty((N, 256), dtype=torch.float32, device=x.device)
    _lib.check(L.slotvps_slot_attention(C.byref(sp), slots_p.contiguous().data_ptr(), x.contiguous().data_ptr(),
                                        None if pos is None else pos.contiguous().data_ptr(), out.data_ptr(), N, h, w,
                                        kernel_path, ws.data_ptr(), ws.numel(), _stream_ptr(x.device)), "slotvps_slot_attention")
    return out


def mask_logits(feat: torch.Tensor, emb: torch.Tensor, bn: Dict[str, torch.Tensor], fg_pack: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """generate_final_outputs (vps_temporal_slots.py:145-154): feat [256,h,w], emb [N,256] -> [N,h,w].

    bn: feat_bn.{weight,bias,running_mean,running_var} [256] and fg_bn.{...} [1] (BatchNorm eval)."""
    _need_cuda(feat, "feat")
    feat = feat.reshape(256, *feat.shape[-2:])
    h, w = feat.shape[-2:]
    N = emb.shape[-2]
    if fg_pack is None:
        fg_pack = torch.cat([bn["fg_bn.weight"].reshape(1), bn["fg_bn.bias"].reshape(1),
                             bn["fg_bn.running_mean"].reshape(1), bn["fg_bn.running_var"].reshape(1)]).float().contiguous()
    if out is None:
        out = torch.empty((N, h, w), dtype=torch.float32, device=feat.device)
    L = _lib.lib()
    nbytes = C.c_size_t()
    _lib.check(L.slotvps_mask_logits_workspace_bytes(N, h, w, C.byref(nbytes)), "mask_logits_workspace_bytes")
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=feat.device)
    _lib.check(L.slotvps_mask_logits(feat.contiguous().data_ptr(), emb.reshape(N, 256).contiguous().data_ptr(),
                                     bn["feat_bn.weight"].data_ptr(), bn["feat_bn.bias"].data_ptr(),
                                     bn["feat_bn.running_mean"].data_ptr(), bn["feat_bn.running_var"].data_ptr(),
                                     fg_pack.data_ptr(), out.data_ptr(), N, h, w, ws.data_ptr(), ws.numel(),
                                     _stream_ptr(feat.device)), "slotvps_mask_logits")
    return out


@dataclass
class FusionOutput:
    """Device-side result of the panoptic fusion; ``.host()`` does the one device->host read."""
    panoptic: torch.Tensor                # [H,W] int64 (the reference's panoptic_outputs[0])
    meta: torch.Tensor                    # int32 [4 + 3N], see include/slotvps_b200.h
    masks: Optional[torch.Tensor]         # [cap,H,W] fp32 masked logits when requested
    n_slots: int
    stuff_num: int
    resume: Optional[object] = None       # continues the small-segment loop when the launched passes did not converge

    def host(self):
        m = self.meta.cpu().numpy()
        if int(m[3]) == 0 and self.resume is not None:
            # the reference's loop (vps_temporal_slots.py:761-792) has no bound: keep going from the device state until the
            # fixed point (every pass removes at least one entry, so at most N + 1 passes exist in total)
            for _ in range(self.n_slots // 4 + 2):
                self.resume(4)
                m = self.meta.cpu().numpy()
                if int(m[3]) != 0:
                    break
        k, n_things, iters, conv = int(m[0]), int(m[1]), int(m[2]), int(m[3])
        N = self.n_slots
        keep = m[4:4 + k].astype(np.int64)
        labels = m[4 + N:4 + N + k].astype(np.int64)
        probs = m[4 + 2 * N:4 + 2 * N + k].view(np.float32).copy()
        thing = labels > self.stuff_num - 1
        return dict(k=k, n_things=n_things, iters=iters, converged=bool(conv), keep=keep, labels=labels, probs=probs,
                    cls_inds=labels[thing] - (self.stuff_num - 1), cls_prob=probs[thing])


class PanopticFusion(nn.Module):
    """PostProcessPanopticInstances (vps_temporal_slots.py:528-562 kwargs) + inline fusion (:411-435).

    Only the shipped configuration is implemented on device: apply_mask_removal=True,
    apply_mask_removal_only_ins=True, use_mask_low_constant=False, filter_small_option='4'.
    """

    def __init__(self, is_thing_map=None, threshold=0.85, output_dir="", debug=False, fraction_threshold=0.03,
                 pixel_threshold=0.4, apply_mask_removal=False, apply_mask_removal_only_ins=False,
                 use_mask_low_constant=False, catgories_color=None, filter_small_option='4', num_classes=20,
                 num_stuff=11, max_iters=4):
        super().__init__()
        # defaults as in the reference (vps_temporal_slots.py:532-536: both mask-removal switches default to False); the
        # device kernels implement what configs/cityscapes/*_slotvps.py ship (:66-74: both True), anything else raises
        if not (apply_mask_removal and apply_mask_removal_only_ins) or use_mask_low_constant or filter_small_option != '4':
            raise NotImplementedError("PanopticFusion implements the shipped postprocess_panoptic configuration only "
                                      "(apply_mask_removal=True, apply_mask_removal_only_ins=True, use_mask_low_constant=False, "
                                      "filter_small_option='4')")
        if not pixel_threshold > 1.0 / 3.0:
            raise NotImplementedError("pixel_threshold must be > 1/3: the exact two-pass mask_removal keeps at most two candidates per pixel")
        if is_thing_map is not None:
            for c, t in is_thing_map.items():
                if bool(t) != (c > num_stuff - 1):
                    raise NotImplementedError("is_thing_map must be {c: c > num_stuff-1}")
        self.cfg = _lib.FusionCfg(num_classes=num_classes, stuff_num=num_stuff, small_area=4, max_iters=max_iters,
                                  threshold=threshold, pixel_threshold=pixel_threshold,
                                  fraction_threshold=fraction_threshold, logits_width=0, reserved=0)
        self.num_stuff = num_stuff
        self._ws = None

    @torch.no_grad()
    def fuse(self, pred_logits: torch.Tensor, pred_masks: torch.Tensor, size: Tuple[int, int],
             want_masks: int = 0, out: Optional[torch.Tensor] = None) -> FusionOutput:
        """pred_logits [N,num_classes], pred_masks [N,h,w] -> FusionOutput (all on device, no sync)."""
        _need_cuda(pred_masks, "pred_masks")
        N, h, w = pred_masks.shape
        H, W = int(size[0]), int(size[1])
        dev = pred_masks.device
        L = _lib.lib()
        nbytes = C.c_size_t()
        _lib.check(L.slotvps_fusion_workspace_bytes(N, H, W, C.byref(nbytes)), "fusion_workspace_bytes")
        if self._ws is None or self._ws.numel() < nbytes.value or self._ws.device != dev:
            self._ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        if out is None:
            out = torch.empty((H, W), dtype=torch.int64, device=dev)
        meta = torch.empty(4 + 3 * N, dtype=torch.int32, device=dev)
        masks = torch.empty((want_masks, H, W), dtype=torch.float32, device=dev) if want_masks > 0 else None
        width = int(pred_logits.shape[-1])
        if width not in (self.cfg.num_classes, self.cfg.num_classes - 1):
            raise ValueError("pred_logits must have num_classes or num_classes - 1 columns (vps_temporal_slots.py:688-693)")
        cfg = _lib.FusionCfg.from_buffer_copy(self.cfg)
        cfg.logits_width = width
        pm = pred_masks.float().contiguous()
        ws = self._ws
        _lib.check(L.slotvps_panoptic_fuse(C.byref(cfg), pred_logits.float().contiguous().data_ptr(),
                                           pm.data_ptr(), N, h, w, H, W, out.data_ptr(),
                                           meta.data_ptr(), None if masks is None else masks.data_ptr(), want_masks,
                                           ws.data_ptr(), ws.numel(), _stream_ptr(dev)), "slotvps_panoptic_fuse")

        def resume(iters):
            _lib.check(L.slotvps_panoptic_fuse_resume(C.byref(cfg), pm.data_ptr(), N, h, w, H, W, out.data_ptr(), meta.data_ptr(),
                                                      None if masks is None else masks.data_ptr(), want_masks, ws.data_ptr(),
                                                      ws.numel(), iters, _stream_ptr(dev)), "slotvps_panoptic_fuse_resume")
        return FusionOutput(out, meta, masks, N, self.num_stuff, resume)

    def forward(self, outputs, processed_sizes, target_sizes=None, id=None):
        """Reference signature (vps_temporal_slots.py:659): ``outputs`` is an Instances-like object
        with pred_logits [N,C] / pred_masks [N,h,w]; returns it filtered, with .masks/.probs/.labels."""
        assert target_sizes is None or list(target_sizes) == list(processed_sizes)
        assert len(processed_sizes) == 1
        size = tuple(int(v) for v in processed_sizes[0])
        N = outputs.pred_masks.shape[0]
        fo = self.fuse(outputs.pred_logits, outputs.pred_masks, size, want_masks=N)
        hst = fo.host()
        if hst["k"] == 0:
            raise ValueError("no slot survives the score/class filter (the reference raises here as well)")
        if not hst["converged"]:
            raise RuntimeError("small-segment filter did not converge in max_iters")
        keep = torch.as_tensor(hst["keep"], device=outputs.pred_masks.device)
        res = outputs[keep]
        res.masks = fo.masks[:hst["k"]]
        res.probs = torch.as_tensor(hst["probs"], device=keep.device)
        res.labels = torch.as_tensor(hst["labels"], device=keep.device)
        self.last_fusion = fo
        return res


class SlotVPSRetriever(nn.Module):
    """The whole hot path for one clip: retriever head -> mask logits -> panoptic fusion.

    Owns the parameters the reference keeps on VPS_Capsule for this path (vps_capsule.py:71-72,
    96-97): ``init_mask_query``, ``feat_bn``, ``fg_bn``.
    """

    def __init__(self, head_kwargs: dict, n_slots: int = 100, fusion_kwargs: Optional[dict] = None):
        super().__init__()
        self.dynamic_mask_head = B200DynamicMaskHead(**head_kwargs)
        self.init_mask_query = nn.Embedding(n_slots, 256)
        self.feat_bn = nn.BatchNorm2d(256)
        self.fg_bn = nn.BatchNorm2d(1)
        self.postprocess_panoptic = PanopticFusion(**(fusion_kwargs or {}))
        self.eval()

    def load_capsule_params(self, p: Dict[str, torch.Tensor]):
        with torch.no_grad():
            self.init_mask_query.weight.copy_(p["init_mask_query.weight"])
            for bn, name in ((self.feat_bn, "feat_bn"), (self.fg_bn, "fg_bn")):
                for k in ("weight", "bias", "running_mean", "running_var"):
                    getattr(bn, k).copy_(p[f"{name}.{k}"])

    def _bn_dict(self):
        d = {}
        for bn, name in ((self.feat_bn, "feat_bn"), (self.fg_bn, "fg_bn")):
            for k in ("weight", "bias", "running_mean", "running_var"):
                d[f"{name}.{k}"] = getattr(bn, k)
        return d

    def _mask_logits_after_head(self, feat, emb, frame):
        """Mask logits of clip frame ``frame`` right after the head call: the contraction runs on the tensor
        pipe from the operand planes still resident in the head's workspace; falls back to the generic
        entry point when the head did not take the tensor-core path for the finest level."""
        bn = self._bn_dict()
        last = self.dynamic_mask_head._last_call
        if last is not None:
            d, hws, rnorm_ss = last
            if feat is not None:
                feat = feat.reshape(256, *feat.shape[-2:])
            h, w = d.h[d.n_levels - 1], d.w[d.n_levels - 1]
            N = emb.shape[-2]
            dev = emb.device
            L = _lib.lib()
            fg_pack = torch.cat([bn["fg_bn.weight"].reshape(1), bn["fg_bn.bias"].reshape(1),
                                 bn["fg_bn.running_mean"].reshape(1), bn["fg_bn.running_var"].reshape(1)]).float().contiguous()
            nbytes = C.c_size_t()
            _lib.check(L.slotvps_mask_logits_workspace_bytes(N, h, w, C.byref(nbytes)), "mask_logits_workspace_bytes")
            ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
            out = torch.empty((N, h, w), dtype=torch.float32, device=dev)
            rc = L.slotvps_head_mask_logits_ex(C.byref(d), hws.data_ptr(), hws.numel(), frame,
                                            None if rnorm_ss is not None else feat.contiguous().data_ptr(),
                                            None if rnorm_ss is None else rnorm_ss.data_ptr(),
                                            emb.reshape(N, 256).contiguous().data_ptr(), bn["feat_bn.weight"].data_ptr(),
                                            bn["feat_bn.bias"].data_ptr(), bn["feat_bn.running_mean"].data_ptr(),
                                            bn["feat_bn.running_var"].data_ptr(), fg_pack.data_ptr(), out.data_ptr(),
                                            ws.data_ptr(), ws.numel(), _stream_ptr(dev))
            if rc == 0:
                return out
            if rc != -4 or feat is None:
                _lib.check(rc, "slotvps_head_mask_logits")
        return mask_logits(feat, emb, bn)

    def _feat_bn_affine(self):
        """feat_bn (eval) as scale / shift vectors for the level-fusion epilogue (cached per parameter version)."""
        bn = self.feat_bn
        key = tuple((t.data_ptr(), t._version) for t in (bn.weight, bn.bias, bn.running_mean, bn.running_var))
        if getattr(self, "_bn_fold", None) is None or self._bn_fold[0] != key:
            sc, sh = torch.empty_like(bn.weight.data), torch.empty_like(bn.weight.data)
            _lib.check(_lib.lib().slotvps_fold_batchnorm(bn.weight.data_ptr(), bn.bias.data_ptr(), bn.running_mean.data_ptr(),
                                                         bn.running_var.data_ptr(), 256, sc.data_ptr(), sh.data_ptr(),
                                                         _stream_ptr(sc.device)), "slotvps_fold_batchnorm")
            self._bn_fold = (key, sc, sh)
        return self._bn_fold[1], self._bn_fold[2]

    @torch.no_grad()
    def forward(self, features: List[List[torch.Tensor]], size: Tuple[int, int], pos="sine", fuse: bool = True,
                fusion_logits: Optional[torch.Tensor] = None, panoptic_out: Optional[torch.Tensor] = None,
                want_feats: bool = True, fusion_masks: Optional[torch.Tensor] = None):
        """features T x 4 x [1,128,h,w] (reference frame order: [ref, cur]); size = (H,W) of the image.
        Returns dict(cls, emb, feats, pred_masks [N,h,w], fusion: FusionOutput) -- all on device.
        ``fusion_logits`` [N,num_classes] replaces the head's last-stage class logits at the fusion
        input (random-init heads keep no slot; benchmarks and tests use designed ones).
        ``want_feats=False`` (what simple_test needs: it only consumes the finest level, vps_temporal_slots.py:297-299):
        the fp32 fused features (the head's third return value) are not materialised when every level runs the
        tensor-core path; the mask logits come from the operand planes and the per-pixel norm from the fusion epilogue."""
        T = len(features)
        q = self.init_mask_query.weight
        if want_feats:
            cls, emb, feats = self.dynamic_mask_head(features, [q] * T, None, pos=pos)
        else:
            cls, emb, feats = self.dynamic_mask_head(features, [q] * T, None, pos=pos, skip_fused=True, feat_bn=self._feat_bn_affine())
        f_last = feats[-1][-1]
        pm = self._mask_logits_after_head(None if f_last is None else f_last[0], emb[-1][-1, 0], T - 1)
        out = dict(cls=cls, emb=emb, feats=feats, pred_masks=pm)
        if fuse:
            lg = cls[-1][-1, 0] if fusion_logits is None else fusion_logits
            # ``fusion_masks`` [N,h,w]: designed mask logits for the fusion stage (benchmarks: a random-init head's masks are
            # collapsed, so the fusion would see an unrepresentatively easy input); ``pred_masks`` is still computed and returned
            out["fusion"] = self.postprocess_panoptic.fuse(lg, pm if fusion_masks is None else fusion_masks, size, out=panoptic_out)
        return out


class GraphedClip:
    """One clip shape captured as a CUDA graph (the step is ~400 short kernels; replaying the graph
    removes the per-launch host latency).  ``features`` are the STATIC input tensors: copy a new
    clip's data into them, call ``replay()``, read the static outputs in ``self.out``."""

    def __init__(self, model: "SlotVPSRetriever", features, size, **kw):
        self.model, self.features, self.size, self.kw = model, features, size, kw
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                  # warm-up: workspaces, weight preparation, func attributes
            for _ in range(2):
                model(features, size, **kw)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        L = _lib.lib()
        before = L.slotvps_launch_count(0)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = model(features, size, **kw)
        self.launches = int(L.slotvps_launch_count(0) - before)      # kernels of this library inside the graph

    def replay(self):
        self.graph.replay()
        return self.out
