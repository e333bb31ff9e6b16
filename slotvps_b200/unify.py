"""Host mirror of the id-map consumers: the semantic argmax of simple_test (vps_temporal_slots.py:440-451) and
CityscapesVps.get_unified_pan_result (tools/dataset/cityscapes_vps.py:214-302).  CUDA only: CPU tensors raise."""
import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .retriever import _need_cuda, _stream_ptr


@torch.no_grad()
def semantic_argmax(fcn_output: torch.Tensor, size) -> torch.Tensor:
    """fcn_output [1,Cs,h,w] (or [Cs,h,w]) -> [1,H,W] int64, as ``torch.max(F.softmax(...), dim=1)[1]`` after the
    bilinear resize of :440-446."""
    _need_cuda(fcn_output, "fcn_output")
    x = fcn_output.reshape(-1, *fcn_output.shape[-2:]).float().contiguous()
    H, W = int(size[0]), int(size[1])
    out = torch.empty((1, H, W), dtype=torch.int64, device=x.device)
    _lib.check(_lib.lib().slotvps_semantic_argmax(x.data_ptr(), x.shape[0], x.shape[1], x.shape[2], H, W, out.data_ptr(),
                                                  _stream_ptr(x.device)), "slotvps_semantic_argmax")
    return out


class PanUnifier:
    """One ``get_unified_pan_result`` call: ``reset()`` (max_oid = 100), then ``frame(...)`` per frame in order."""

    def __init__(self, device="cuda", num_seg_classes: int = 19, num_classes: int = 9):
        self.device = torch.device(device)
        self.id_last_stuff = num_seg_classes - num_classes           # :250
        nbytes = C.c_size_t()
        _lib.check(_lib.lib().slotvps_unify_workspace_bytes(C.byref(nbytes)), "unify_workspace_bytes")
        self.ws = torch.zeros(nbytes.value, dtype=torch.uint8, device=self.device)
        self.status = torch.zeros(2, dtype=torch.int32, device=self.device)
        self.reset()

    def reset(self):
        _lib.check(_lib.lib().slotvps_unify_reset(self.ws.data_ptr(), self.ws.numel(), _stream_ptr(self.device)), "slotvps_unify_reset")

    @torch.no_grad()
    def frame(self, seg: torch.Tensor, pan: torch.Tensor, cls_ind: torch.Tensor, obj_id: Optional[torch.Tensor],
              stuff_area_limit: int = 4 * 64 * 64, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """seg, pan [H,W] int64 (device); cls_ind [n_inst]; obj_id [n_obj] or None -> [H,W,3] uint8 (device)."""
        _need_cuda(pan, "pan")
        _need_cuda(seg, "seg")
        H, W = pan.shape[-2:]
        seg = seg.reshape(H, W).to(torch.int64).contiguous()
        pan = pan.reshape(H, W).to(torch.int64).contiguous()
        ci = torch.as_tensor(cls_ind).to(self.device, torch.int32).contiguous()
        oi = None if obj_id is None else torch.as_tensor(obj_id).to(self.device, torch.int32).contiguous()
        if out is None:
            out = torch.empty((H, W, 3), dtype=torch.uint8, device=self.device)
        _lib.check(_lib.lib().slotvps_unify_pan_result(
            seg.data_ptr(), pan.data_ptr(), ci.data_ptr() if ci.numel() else None, ci.numel(),
            None if oi is None or oi.numel() == 0 else oi.data_ptr(), 0 if oi is None else oi.numel(), H, W, self.id_last_stuff,
            int(stuff_area_limit), out.data_ptr(), self.status.data_ptr(), self.ws.data_ptr(), self.ws.numel(),
            _stream_ptr(self.device)), "slotvps_unify_pan_result")
        return out

    def check(self):
        """One device->host read of the sticky error bits; raises what the reference would have raised."""
        err = int(self.status[0].item())
        if err & 1:
            raise ValueError("panoptic id >= 256 or semantic class >= 32 in the input maps")
        if err & 6:
            raise IndexError("cls_inds / obj_ids shorter than the instance ids present in the map (the reference raises here too)")


def get_unified_pan_result(segs: Sequence, pans: Sequence, cls_inds: Sequence, obj_ids: Optional[Sequence] = None,
                           stuff_area_limit: int = 4 * 64 * 64, names: Optional[List[str]] = None, device="cuda",
                           num_seg_classes: int = 19, num_classes: int = 9) -> Dict[str, np.ndarray]:
    """Same signature and result as CityscapesVps.get_unified_pan_result: {name: [H,W,3] uint8 array}.  Inputs may be
    NumPy arrays (as tools/test_vpq.py passes them) or CUDA tensors; the per-frame work runs on the device."""
    if obj_ids is None:
        obj_ids = [None] * len(cls_inds)
    if names is None:
        names = [str(i) for i in range(len(pans))]
    u = PanUnifier(device, num_seg_classes, num_classes)
    dev = u.device
    outs = []
    for seg, pan, ci, oi in zip(segs, pans, cls_inds, obj_ids):
        seg = torch.as_tensor(np.asarray(seg) if not torch.is_tensor(seg) else seg).to(dev)
        pan = torch.as_tensor(np.asarray(pan) if not torch.is_tensor(pan) else pan).to(dev)
        outs.append(u.frame(seg, pan, torch.as_tensor(np.asarray(ci) if not torch.is_tensor(ci) else ci),
                            None if oi is None else torch.as_tensor(np.asarray(oi) if not torch.is_tensor(oi) else oi),
                            stuff_area_limit))
    u.check()
    return {n: o.cpu().numpy() for n, o in zip(names, outs)}
