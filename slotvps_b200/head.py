"""B200DynamicMaskHead -- drop-in for the reference's MultiScaleDynamicMaskHead.

Mirrors mmdet/models/detectors/dynamic_mask_head.py:36-228 of the reference: same constructor
kwargs (configs/cityscapes/r50_fpn_slotvps.py:27-54 + ``other_config``), same parameter names
and shapes (so ``load_state_dict(reference_head.state_dict(), strict=True)`` works), same
``forward`` signature, return structure and assertion behaviour.  The modules below are
PARAMETER CONTAINERS only; all arithmetic runs in libslotvps_b200.so (hand-written sm_100a
kernels) through the C ABI of include/slotvps_b200.h.  There is no eager / CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch
from torch import nn

from . import _lib


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class _MHAParams(nn.Module):
    """Parameter names of nn.MultiheadAttention(256, 8) (dynamic_mask_head.py:243)."""

    def __init__(self, d):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d, d))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d))
        self.out_proj = nn.Linear(d, d)
        nn.init.xavier_uniform_(self.in_proj_weight)


class _RetrieverParams(nn.Module):
    """MaskDynamicConv / SlotsDynamicConv parameters (dynamic_mask_head.py:406-421, 533-548)."""

    def __init__(self, d):
        super().__init__()
        self.to_q, self.to_k, self.to_v = nn.Linear(d, d), nn.Linear(d, d), nn.Linear(d, d)
        self.norm_q, self.norm_k, self.norm_v = nn.LayerNorm(d), nn.LayerNorm(d), nn.LayerNorm(d)
        self.norm1 = nn.LayerNorm(d)


class _TemporalParams(nn.Module):
    """TemporalSlotsHead parameters (dynamic_mask_head.py:468-492).  norm1 exists (unused) there too."""

    def __init__(self, d_model, dim_feedforward=2048, dropout=0.1, activation="relu", softmax_dim="slots", drop_path=0.):
        super().__init__()
        if activation not in ("relu", "gelu") or softmax_dim != "slots":
            raise NotImplementedError("B200 Video Retriever supports activation 'relu' (r50 config) / 'gelu' (swinL config), softmax_dim='slots'")
        self.activation = activation
        self.inst_interact = _RetrieverParams(d_model)
        self.linear1, self.linear2 = nn.Linear(d_model, dim_feedforward), nn.Linear(dim_feedforward, d_model)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(d_model), nn.LayerNorm(d_model), nn.LayerNorm(d_model)
        self.dim_feedforward = dim_feedforward


class _StageParams(nn.Module):
    """MaskRCNNHead parameters (dynamic_mask_head.py:233-289)."""

    def __init__(self, d, num_classes, dim_feedforward, nhead, num_cls, num_reg, temporal_cfg):
        super().__init__()
        self.self_attn = _MHAParams(d)
        self.inst_interact = _RetrieverParams(d)
        self.linear1, self.linear2 = nn.Linear(d, dim_feedforward), nn.Linear(dim_feedforward, d)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(d), nn.LayerNorm(d), nn.LayerNorm(d)
        self.temporal_query_head = _TemporalParams(**temporal_cfg) if temporal_cfg is not None else None
        cls, reg = [], []
        for _ in range(num_cls):
            cls += [nn.Linear(d, d, False), nn.LayerNorm(d), nn.ReLU(inplace=True)]
        for _ in range(num_reg):
            reg += [nn.Linear(d, d, False), nn.LayerNorm(d), nn.ReLU(inplace=True)]
        self.cls_module, self.reg_module = nn.ModuleList(cls), nn.ModuleList(reg)
        self.class_logits = nn.Linear(d, num_classes)


class _ConvParams(nn.Module):
    """ConvModule(384, 256, 1, activation=None) as a parameter container: conv.{weight,bias}."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 1)


class B200DynamicMaskHead(nn.Module):
    """Same kwargs as MultiScaleDynamicMaskHead.__init__ (dynamic_mask_head.py:38-54).

    Extra keyword (not in the reference): ``kernel_path`` 0 = tensor-core kernels where the shape
    allows (default), 1 = force the fp32 CUDA-core kernels.
    """

    def __init__(self, dh_dim=256, num_classes=9, dim_feedforward=2048, nhead=8, dropout=0.0, activation="relu",
                 dh_num_heads=8, per_dh_num_heads=2, feat_num_levels=4, merge_operation="add", trans_in_dim=128,
                 return_intermediate=True, use_focal=True, prior_prob=0.01, num_cls=1, num_reg=3, softmax_dim="slots",
                 drop_path=0., temporal_query_attention_config=None, apply_temporal_query_atten_stages=None,
                 other_config=None, kernel_path=0):
        super().__init__()
        if not isinstance(per_dh_num_heads, (list, tuple)):
            assert per_dh_num_heads * feat_num_levels == dh_num_heads
            per_dh_num_heads = [per_dh_num_heads] * feat_num_levels
        else:
            assert sum(per_dh_num_heads) == dh_num_heads
        # what the sm_100a kernels implement == the shipped configuration of the path
        unsupported = []
        if dh_dim != 256: unsupported.append("dh_dim != 256")
        if nhead != 8: unsupported.append("nhead != 8")
        if activation not in ("gelu", "relu"): unsupported.append("activation not in ('gelu', 'relu')")
        if merge_operation != "concat" or trans_in_dim != 384: unsupported.append("merge_operation/trans_in_dim != concat/384")
        if softmax_dim != "slots": unsupported.append("softmax_dim != 'slots'")
        if dropout != 0.0 or drop_path != 0.0: unsupported.append("dropout/drop_path != 0")
        if num_cls != 2 or num_reg != 2: unsupported.append("num_cls/num_reg != 2")
        if not return_intermediate: unsupported.append("return_intermediate=False")
        if unsupported:
            raise NotImplementedError("B200DynamicMaskHead: " + ", ".join(unsupported))
        self.dh_dim, self.trans_in_dim, self.num_classes = dh_dim, trans_in_dim, num_classes
        self.dim_feedforward, self.nhead = dim_feedforward, nhead
        self.activation = activation                                    # r50 config: "gelu"; swinL config: "relu"
        self.temporal_activation = (temporal_query_attention_config or {}).get("activation", "relu")
        self.per_dh_num_heads = list(per_dh_num_heads)
        self.feat_num_levels = feat_num_levels
        self.apply_temporal_query_atten_stages = apply_temporal_query_atten_stages
        self.other_config = other_config
        self.return_intermediate, self.merge_operation = return_intermediate, merge_operation
        self.kernel_path = kernel_path
        self.temporal_dim_feedforward = (temporal_query_attention_config or {}).get("dim_feedforward", 2048)
        stage = 0
        for i in range(feat_num_levels):
            # the reference decides per LEVEL (first stage index of the level) whether its stages own
            # a Video Retriever (:83-104); run time checks per stage (:197).
            if apply_temporal_query_atten_stages is None or stage in apply_temporal_query_atten_stages:
                tcfg = temporal_query_attention_config
            else:
                tcfg = None
            setattr(self, f"head_series_{i}", nn.ModuleList(
                [_StageParams(dh_dim, num_classes, dim_feedforward, nhead, num_cls, num_reg, tcfg)
                 for _ in range(per_dh_num_heads[i])]))
            stage += per_dh_num_heads[i]
        self.conv_trans = _ConvParams(trans_in_dim, dh_dim)
        self._prepared = None           # (key, buffer, StageParams array)
        self._in_trans = None           # (weight [128,128], bias [128]) of a folded 1x1 input transform
        self._last_call = None
        self._ws = {}
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    # -- parameter table ----------------------------------------------------------------------------
    def stages(self) -> List[_StageParams]:
        return [m for i in range(self.feat_num_levels) for m in getattr(self, f"head_series_{i}")]

    def _param_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _stage_table(self):
        stages = self.stages()
        arr = (_lib.StageParams * len(stages))()
        for s, m in enumerate(stages):
            sd = dict(m.named_parameters())
            for f in _lib.STAGE_FIELDS:
                t = sd.get(_lib.STAGE_KEYS[f])
                if t is not None:
                    if t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
                        raise ValueError(f"parameter {_lib.STAGE_KEYS[f]} must be a contiguous CUDA fp32 tensor")
                    setattr(arr[s], f, t.data_ptr())
                elif not f.startswith("tq_"):
                    raise KeyError(_lib.STAGE_KEYS[f])
        return arr

    def _desc(self, T, N, shapes, pos_mode) -> _lib.HeadDesc:
        d = _lib.HeadDesc()
        d.n_frames, d.n_slots, d.n_levels = T, N, self.feat_num_levels
        for l in range(self.feat_num_levels):
            d.heads_per_level[l] = self.per_dh_num_heads[l]
            d.h[l], d.w[l] = shapes[l]
        d.num_classes, d.dim_feedforward = self.num_classes, self.dim_feedforward
        d.temporal_dim_feedforward, d.nhead = self.temporal_dim_feedforward, self.nhead
        mask = 0
        for s in (self.apply_temporal_query_atten_stages or []):
            mask |= 1 << s
        d.temporal_mask, d.pos_mode, d.kernel_path = mask, pos_mode, self.kernel_path
        d.ffn_act = {"relu": 1, "gelu": 2}[self.activation]
        d.temporal_ffn_act = {"relu": 1, "gelu": 2}[self.temporal_activation]
        return d

    def fold_input_transform(self, weight: Optional[torch.Tensor], bias: Optional[torch.Tensor]):
        """Fold the caller's 1x1 input transform (VPS_Capsule.conv_trans, applied by semantic_trans_ins,
        vps_temporal_slots.py:129-135) into the level fusion: afterwards ``forward`` takes the UN-transformed
        semantic-head features.  weight [128,128(,1,1)], bias [128]; ``None, None`` removes the fold."""
        if weight is None:
            self._in_trans = None
        else:
            dev = next(self.parameters()).device
            w = weight.detach().reshape(128, 128).to(dev, torch.float32).contiguous()
            b = bias.detach().reshape(128).to(dev, torch.float32).contiguous()
            self._in_trans = (w, b)
        self._prepared = None

    def _prepare(self, d, device):
        key = (self._param_key(), self.kernel_path)
        if self._prepared is not None and self._prepared[0] == key:
            return self._prepared[1], self._prepared[2]
        L = _lib.lib()
        nbytes = C.c_size_t()
        _lib.check(L.slotvps_prepared_bytes(C.byref(d), C.byref(nbytes)), "slotvps_prepared_bytes")
        buf = torch.empty(nbytes.value, dtype=torch.uint8, device=device)
        table = self._stage_table()
        w = self.conv_trans.conv.weight
        if self._in_trans is not None:
            self._in_trans = tuple(t.to(device) for t in self._in_trans)
        tw, tb = self._in_trans if self._in_trans is not None else (None, None)
        _lib.check(L.slotvps_prepare_weights_ex(C.byref(d), table, w.data_ptr(), self.conv_trans.conv.bias.data_ptr(),
                                                None if tw is None else tw.data_ptr(), None if tb is None else tb.data_ptr(),
                                                buf.data_ptr(), _stream_ptr(device)), "slotvps_prepare_weights")
        self._prepared = (key, buf, table)
        return buf, table

    def _workspace(self, d, device):
        key = (d.n_frames, d.n_slots, tuple(d.h), tuple(d.w), d.pos_mode, str(device))
        ws = self._ws.get(key)
        if ws is None:
            nbytes = C.c_size_t()
            _lib.check(_lib.lib().slotvps_head_workspace_bytes(C.byref(d), C.byref(nbytes)), "slotvps_head_workspace_bytes")
            ws = torch.empty(nbytes.value, dtype=torch.uint8, device=device)
            self._ws = {key: ws}
        return ws

    # -- forward ------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, features, init_masks, pad_mask, pos=None, query_pos=None, gt_non_void_mask=None, *,
                stage_slots_in=None, skip_fused=False, feat_bn=None):
        """dynamic_mask_head.py:138.  features T x L x [1,128,h,w]; init_masks T x [N,256];
        pos T x L x [1,256,h,w] | None | "sine" (generate PositionEmbeddingSine on device).
        Returns (T x [S,1,N,num_classes], T x [S,1,N,256], T x L x [1,256,h,w]).

        Keyword-only extras (not in the reference; slotvps_head_opts of the C ABI):
        ``stage_slots_in`` S x T x [N,256] | None entries: teacher forcing, the slots entering stage s (parity tests);
        ``skip_fused``: do not materialise the fp32 fused features (third return value becomes None entries) when every
        level runs the tensor-core path -- the L2 integration only needs the finest level, through the operand planes;
        ``feat_bn`` (scale [256], shift [256]): also accumulate the per-pixel squared norm of feat_bn(x) of the finest
        level (kept in ``self._last_call``) for the mask-logit kernel."""
        assert pad_mask is None
        assert query_pos is None
        assert gt_non_void_mask is None
        T, L = len(features), self.feat_num_levels
        assert all(len(f) == L for f in features)
        bs = len(features[0][0])
        assert bs == 1, "imgs_per_gpu must be 1 (vps_temporal_slots.py:484)"
        dev = features[0][0].device
        if dev.type != "cuda":
            raise RuntimeError("B200DynamicMaskHead runs on CUDA only (no CPU fallback)")
        N = init_masks[0].shape[0]
        shapes = [tuple(features[0][l].shape[-2:]) for l in range(L)]
        pos_mode = 0 if pos is None else (2 if isinstance(pos, str) and pos == "sine" else 1)
        d = self._desc(T, N, shapes, pos_mode)
        prepared, table = self._prepare(d, dev)
        ws = self._workspace(d, dev)
        S = sum(self.per_dh_num_heads)

        def f32c(t):
            return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()
        feats = [f32c(features[t][l]) for t in range(T) for l in range(L)]
        queries = [f32c(q) for q in init_masks]
        # the reference mutates the caller's list (dynamic_mask_head.py:152): keep that observable
        for i in range(T):
            init_masks[i] = init_masks[i][None].repeat(bs, 1, 1)
        if pos_mode == 1:
            # one [T,256,h,w] block per level (equal frame strides); identical frames share storage
            pos_l = []
            for l in range(L):
                same = all(pos[t][l].data_ptr() == pos[0][l].data_ptr() for t in range(T))
                blk = f32c(pos[0][l])[None].expand(T, -1, -1, -1, -1) if same else torch.stack([f32c(pos[t][l]) for t in range(T)])
                pos_l.append(blk)
            pos_ptrs = (C.c_void_p * (T * L))(*[pos_l[l][t].data_ptr() for t in range(T) for l in range(L)])
        else:
            pos_l, pos_ptrs = None, None
        all_tc = self.kernel_path == 0 and all(shapes[l][0] * shapes[l][1] >= 128 and self.per_dh_num_heads[l] > 0 for l in range(L))
        skip = bool(skip_fused) and all_tc
        opts = _lib.HeadOpts()
        keep_alive = []
        if skip:
            opts.skip_fused_mask = (1 << L) - 1
        fused = [None if skip else torch.empty((T, 1, self.dh_dim) + shapes[l], dtype=torch.float32, device=dev) for l in range(L)]
        rnorm_ss = None
        if feat_bn is not None and all_tc:
            sc, sh = (f32c(v) for v in feat_bn)
            rnorm_ss = torch.empty((4, T, shapes[-1][0] * shapes[-1][1]), dtype=torch.float32, device=dev)   # four 64-channel partial sums
            opts.feat_bn_scale, opts.feat_bn_shift, opts.rnorm_ss = sc.data_ptr(), sh.data_ptr(), rnorm_ss.data_ptr()
            keep_alive += [sc, sh]
        if stage_slots_in is not None:
            assert len(stage_slots_in) == S
            tf = (C.c_void_p * (S * T))()
            for s_ in range(S):
                for t in range(T):
                    v = None if stage_slots_in[s_] is None else stage_slots_in[s_][t]
                    if v is not None:
                        v = f32c(v.reshape(N, self.dh_dim).to(dev))
                        keep_alive.append(v)
                        tf[s_ * T + t] = v.data_ptr()
            opts.stage_slots_in = C.cast(tf, C.POINTER(C.c_void_p))
        cls = torch.empty((T, S, 1, N, self.num_classes), dtype=torch.float32, device=dev)
        emb = torch.empty((T, S, 1, N, self.dh_dim), dtype=torch.float32, device=dev)
        feat_ptrs = (C.c_void_p * (T * L))(*[t.data_ptr() for t in feats])
        q_ptrs = (C.c_void_p * T)(*[q.data_ptr() for q in queries])
        fused_ptrs = (C.c_void_p * (T * L))(*[None if fused[l] is None else fused[l][t].data_ptr() for t in range(T) for l in range(L)])
        _lib.check(_lib.lib().slotvps_head_forward_ex(
            C.byref(d), table, prepared.data_ptr(), feat_ptrs, pos_ptrs, q_ptrs, cls.data_ptr(), emb.data_ptr(),
            fused_ptrs, ws.data_ptr(), ws.numel(), C.byref(opts), _stream_ptr(dev)), "slotvps_head_forward")
        del pos_l, keep_alive
        self._last_call = (d, ws, rnorm_ss)   # the finest level's operand planes stay valid in ws until the next forward
        return ([cls[t] for t in range(T)], [emb[t] for t in range(T)],
                [[None if fused[l] is None else fused[l][t] for l in range(L)] for t in range(T)])
