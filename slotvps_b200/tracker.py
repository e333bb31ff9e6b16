"""Host mirror of the reference's tracker: SimpleTrackHead (mmdet/models/detectors/simple_track_head.py) and the
per-video id assignment that simple_test runs inline (vps_temporal_slots.py:232-237, :322-409).

`B200TrackHead` keeps SimpleTrackHead's constructor arguments, parameter names (``fcs_query.{i}.weight/bias``) and
``forward(x_query, ref_x_query) -> [match_score]`` contract; `SlotTracker` keeps the object bank
(``prev_instances.output_embedding``) on the device and assigns ids without a host round trip per frame.
There is no CPU path: CPU tensors raise.
"""
import ctypes as C
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .retriever import FusionOutput, _need_cuda, _stream_ptr


class B200TrackHead(nn.Module):
    """SimpleTrackHead (simple_track_head.py:21-92), inference only."""

    def __init__(self, num_fcs_query=0, in_channels_query=0, loss_match=None, query_matched_weight=1.0):
        super().__init__()
        if num_fcs_query > 0 and in_channels_query != 256:
            raise NotImplementedError("the B200 tracker is built for 256-channel slot embeddings")
        self.num_fcs_query = num_fcs_query
        self.query_matched_weight = query_matched_weight
        if num_fcs_query > 0:
            self.fcs_query = nn.ModuleList(nn.Linear(in_channels_query, in_channels_query) for _ in range(num_fcs_query))
        self.init_weights()
        self._packed = None

    def init_weights(self):
        """simple_track_head.py:52-56."""
        if self.num_fcs_query > 0:
            for fc in self.fcs_query:
                nn.init.normal_(fc.weight, 0, 0.01)
                nn.init.constant_(fc.bias, 0)

    def packed(self, device):
        """fc weights as the C ABI takes them: [num_fcs,256,256] and [num_fcs,256], cached per device."""
        if self.num_fcs_query == 0:
            return None, None
        ver = tuple(p._version for p in self.parameters())
        if self._packed is None or self._packed[0] != (ver, device):
            w = torch.stack([fc.weight.detach() for fc in self.fcs_query]).float().contiguous().to(device)
            b = torch.stack([fc.bias.detach() for fc in self.fcs_query]).float().contiguous().to(device)
            self._packed = ((ver, device), w, b)
        return self._packed[1], self._packed[2]

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    @torch.no_grad()
    def forward(self, x_query=None, ref_x_query=None):
        """[match_score [k, 1+m]] (a list, as the reference returns); ref_x_query may be a list of banks."""
        refs = ref_x_query if isinstance(ref_x_query, list) else [ref_x_query]
        _need_cuda(x_query, "x_query")
        dev = x_query.device
        w, b = self.packed(dev)
        L = _lib.lib()
        out = []
        xq = x_query.float().contiguous()
        for r in refs:
            _need_cuda(r, "ref_x_query")
            r = r.float().contiguous()
            k, m = xq.shape[0], r.shape[0]
            score = torch.empty((k, 1 + m), dtype=torch.float32, device=dev)
            ws = torch.empty((k + m) * 256, dtype=torch.float32, device=dev)
            _lib.check(L.slotvps_track_scores(None if w is None else w.data_ptr(), None if b is None else b.data_ptr(),
                                              self.num_fcs_query, xq.data_ptr(), k, r.data_ptr(), m, score.data_ptr(),
                                              ws.data_ptr(), ws.numel() * 4, _stream_ptr(dev)), "slotvps_track_scores")
            out.append(score)
        return out


class SlotTracker:
    """The tracking state of one video: reset() on its first frame (fid == 1), step() once per frame."""

    def __init__(self, track_head: B200TrackHead, n_slots: int = 100, capacity: int = 1024, device="cuda"):
        self.head, self.N, self.capacity = track_head, n_slots, capacity
        self.device = torch.device(device)
        nbytes = C.c_size_t()
        _lib.check(_lib.lib().slotvps_track_state_bytes(capacity, n_slots, C.byref(nbytes)), "track_state_bytes")
        self.state = torch.zeros(nbytes.value, dtype=torch.uint8, device=self.device)
        self.reset()

    def reset(self):
        _lib.check(_lib.lib().slotvps_track_reset(self.state.data_ptr(), self.state.numel(), _stream_ptr(self.device)),
                   "slotvps_track_reset")

    @torch.no_grad()
    def step(self, embedding: torch.Tensor, fusion: FusionOutput, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """embedding [N,256] (last-stage slot embeddings of the current frame), fusion = that frame's FusionOutput.
        Returns the device int32 [4+N] record described in include/slotvps_b200.h; no synchronisation."""
        _need_cuda(embedding, "embedding")
        emb = embedding.reshape(self.N, 256).float().contiguous()
        w, b = self.head.packed(self.device)
        if out is None:
            out = torch.empty(4 + self.N, dtype=torch.int32, device=self.device)
        _lib.check(_lib.lib().slotvps_track_step(None if w is None else w.data_ptr(), None if b is None else b.data_ptr(),
                                                 self.head.num_fcs_query, emb.data_ptr(), fusion.meta.data_ptr(), self.N,
                                                 self.state.data_ptr(), self.state.numel(), self.capacity, out.data_ptr(),
                                                 _stream_ptr(self.device)), "slotvps_track_step")
        return out

    @staticmethod
    def host(track_out: torch.Tensor):
        """One device->host read: dict(k, n_things, bank, ids [k], det_obj_ids = ids of the things)."""
        m = track_out.cpu().numpy()
        k, nt, bank, over = int(m[0]), int(m[1]), int(m[2]), int(m[3])
        if over:
            raise RuntimeError("tracker bank overflowed its capacity; construct SlotTracker with a larger capacity")
        ids = m[4:4 + k].astype(np.int64)
        return dict(k=k, n_things=nt, bank=bank, ids=ids, det_obj_ids=ids[k - nt:])

    def bank(self) -> torch.Tensor:
        """The object bank [count,256] (synchronises)."""
        hdr = self.state[:16].view(torch.int32).cpu()
        return self.state[256:256 + self.capacity * 1024].view(torch.float32).reshape(self.capacity, 256)[:int(hdr[0])]
