"""Host mirror of the UPSNetFPN deformable-convolution subnet: the ``nn.Sequential`` that
``UPSNetFPN.__init__`` builds (mmdet/models/panoptic/upsnetFPN.py:36-49) and applies to every FPN level (:66-70), and the
single op behind it (``DeformConvFunction.forward`` -> ``deform_conv_cuda.deform_conv_forward_cuda``,
mmdet/ops/dcn/deform_conv.py:53-58).  CUDA only: CPU tensors raise, there is no fallback."""
import ctypes as C
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib
from .retriever import _need_cuda, _stream_ptr


@torch.no_grad()
def deform_conv(input: torch.Tensor, offset: Optional[torch.Tensor], weight: torch.Tensor, stride=1, padding=1, dilation=1,
                groups=1, deformable_groups=1, im2col_step=64) -> torch.Tensor:
    """``mmdet.ops.dcn.deform_conv`` (deform_conv.py:13-59) for the UPSNetFPN instance: input [B,c_in,H,W],
    offset [B,18,H,W] (``None``: ordinary convolution), weight [c_out,c_in,3,3] -> [B,c_out,H,W]."""
    _need_cuda(input, "input")
    _need_cuda(weight, "weight")
    if input.dim() != 4:
        raise ValueError(f"Expected 4D tensor as input, got {input.dim()}D tensor instead.")      # deform_conv.py:26-29
    x = input.float().contiguous()
    w = weight.float().contiguous()
    B, cin, H, W = x.shape
    cout = w.shape[0]
    if w.shape[1] != cin:
        raise ValueError("weight / input channel mismatch")
    off = None
    if offset is not None:
        _need_cuda(offset, "offset")
        off = offset.float().contiguous()
        if tuple(off.shape) != (B, 2 * w.shape[2] * w.shape[3] * deformable_groups, H, W):
            raise ValueError(f"offset shape {tuple(off.shape)} does not match input {tuple(x.shape)}")
    pair = lambda v: (v, v) if isinstance(v, int) else tuple(v)
    st, pd, dl = pair(stride), pair(padding), pair(dilation)
    nbytes = C.c_size_t()
    _lib.check(_lib.lib().slotvps_deform_conv_workspace_bytes(B, cin, H, W, C.byref(nbytes)), "deform_conv_workspace_bytes")
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=x.device)
    out = torch.empty((B, cout, H, W), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().slotvps_deform_conv_forward(
        x.data_ptr(), w.data_ptr(), None if off is None else off.data_ptr(), out.data_ptr(), B, cin, cout, H, W,
        w.shape[3], w.shape[2], st[1], st[0], pd[1], pd[0], dl[1], dl[0], groups, deformable_groups, im2col_step,
        ws.data_ptr(), ws.numel(), _stream_ptr(x.device)), "slotvps_deform_conv_forward")
    return out


class _OffsetConv(nn.Module):
    """Parameter container with ``nn.Conv2d``'s names (conv_offset of DeformConvWithOffset)."""

    def __init__(self, cin: int):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(18, cin, 3, 3))       # deform_conv_with_offset.py:25-26: zero-initialised
        self.bias = nn.Parameter(torch.zeros(18))


class _DeformWeight(nn.Module):
    """``DeformConv`` as a parameter container (weight only: bias=False, mmdet/ops/dcn/deform_conv.py:160-184)."""

    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, 3, 3))
        n = cin * 9
        nn.init.uniform_(self.weight, -1.0 / n ** 0.5, 1.0 / n ** 0.5)     # DeformConv.reset_parameters


class B200DeformConvWithOffset(nn.Module):
    """``DeformConvWithOffset`` (mmdet/models/utils/deform_conv_with_offset.py) -- same parameter names; 3x3, padding 1."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 3, stride: int = 1, padding: int = 1,
                 dilation: int = 1, groups: int = 1, deformable_groups: int = 1, bias: bool = True):
        super().__init__()
        if (kernel_size, stride, padding, dilation, groups, deformable_groups) != (3, 1, 1, 1, 1, 1):
            raise NotImplementedError("B200DeformConvWithOffset serves the UPSNetFPN instance: 3x3, stride 1, padding 1, one group")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.conv_offset = _OffsetConv(in_channels)
        self.conv = _DeformWeight(in_channels, out_channels)


class B200DeformSubnet(nn.Sequential):
    """Drop-in for ``UPSNetFPN.deform_convs[0]``: an ``nn.Sequential`` whose entries 0/3/6 are DeformConvWithOffset, 1/4/7
    GroupNorm(32) and 2/5/8 ReLU, so ``load_state_dict(ref.deform_convs[0].state_dict(), strict=True)`` works; ``forward``
    runs the whole chain in one C-ABI call.  [B,c_in,H,W] -> [B,c_out,H,W].  One scratch buffer serves every call of the
    instance (the FPN levels run one after the other on the current stream, as in UPSNetFPN.forward): do not call one instance
    from several streams at once."""

    def __init__(self, in_channels: int = 256, out_channels: int = 128, num_groups: int = 32, channels=None):
        # upsnetFPN.py:37-48; `channels` = explicit [(c_in, c_out), ...] (a single layer for teacher-forced tests)
        chans = channels or [(in_channels, in_channels), (in_channels, out_channels), (out_channels, out_channels)]
        mods: List[nn.Module] = []
        for cin, cout in chans:
            mods += [B200DeformConvWithOffset(cin, cout, 3, padding=1), nn.GroupNorm(num_groups, cout), nn.ReLU(inplace=True)]
        super().__init__(*mods)
        self._prepared = None
        self._key = None
        self._ws = {}            # (B, H, W, device) -> workspace bytes
        self._wsbuf = None

    def _descs(self):
        n = len(self) // 3
        arr = (_lib.DcnLayer * n)()
        keep = []
        for i in range(n):
            dc, gn = self[3 * i], self[3 * i + 1]
            ts = [dc.conv_offset.weight, dc.conv_offset.bias, dc.conv.weight, gn.weight, gn.bias]
            for t in ts:
                _need_cuda(t, "parameter")
                if t.dtype != torch.float32 or not t.is_contiguous():
                    raise ValueError("B200DeformSubnet parameters must be contiguous fp32")
            keep += ts
            if abs(gn.eps - 1e-5) > 1e-12 or gn.num_groups != 32:
                raise NotImplementedError("GroupNorm(32, eps=1e-5) only")
            arr[i] = _lib.DcnLayer(dc.in_channels, dc.out_channels, *[t.data_ptr() for t in ts])
        return arr, n, keep

    def _prepare(self, arr, n, keep, dev):
        key = tuple((t.data_ptr(), t._version) for t in keep)
        if key != self._key:
            nbytes = C.c_size_t()
            _lib.check(_lib.lib().slotvps_dcn_prepared_bytes(arr, n, C.byref(nbytes)), "slotvps_dcn_prepared_bytes")
            self._prepared = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
            _lib.check(_lib.lib().slotvps_dcn_prepare(arr, n, self._prepared.data_ptr(), self._prepared.numel(), _stream_ptr(dev)),
                       "slotvps_dcn_prepare")
            self._key = key

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _need_cuda(x, "x")
        if x.dim() != 4 or x.shape[1] != self[0].in_channels:
            raise ValueError(f"expected [B,{self[0].in_channels},H,W], got {tuple(x.shape)}")
        x = x.float().contiguous()
        arr, n, keep = self._descs()
        self._prepare(arr, n, keep, x.device)
        B, _, H, W = x.shape
        wkey = (B, H, W, x.device)
        if wkey not in self._ws:
            nbytes = C.c_size_t()
            _lib.check(_lib.lib().slotvps_dcn_workspace_bytes(arr, n, B, H, W, C.byref(nbytes)), "slotvps_dcn_workspace_bytes")
            self._ws[wkey] = nbytes.value
        need = self._ws[wkey]
        if self._wsbuf is None or self._wsbuf.device != x.device or self._wsbuf.numel() < need:
            self._wsbuf = None                                   # one scratch buffer, grown to the largest level seen (FPN levels share it)
            self._wsbuf = torch.empty(need, dtype=torch.uint8, device=x.device)
        ws = self._wsbuf
        out = torch.empty((B, self[3 * (n - 1)].out_channels, H, W), dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().slotvps_dcn_subnet_forward(arr, n, self._prepared.data_ptr(), x.data_ptr(), out.data_ptr(), B, H, W,
                                                        ws.data_ptr(), ws.numel(), _stream_ptr(x.device)), "slotvps_dcn_subnet_forward")
        return out


def deform_subnet_levels(subnet: B200DeformSubnet, inputs: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """``fpn_px`` of UPSNetFPN.forward (upsnetFPN.py:66-70): the shared subnet on every FPN level."""
    return [subnet(x) for x in inputs]
