"""Loader of the REFERENCE's compiled deformable-convolution op (TEST INFRASTRUCTURE, checker only).

``oracle/_ref/deform_conv_cuda.so`` is built by ``oracle/build_ref_dcn.sh`` from the unmodified sources under
/root/reference/mmdet/ops/dcn/src (it does not exist in git; the prebuilt file travels to the GPU box).  The op is CUDA-only,
so everything here needs a GPU.  ``ref_deform_conv`` calls it exactly as ``DeformConvFunction.forward`` does
(mmdet/ops/dcn/deform_conv.py:38-59); ``ref_dcn_subnet`` is ``UPSNetFPN.deform_convs[0]`` (upsnetFPN.py:36-49) with that op
and torch's own conv2d / group_norm / relu, i.e. the reference's computation on this GPU."""
import importlib.util
import os

import torch
import torch.nn.functional as F

SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "deform_conv_cuda.so")
_mod = None


def available() -> bool:
    return os.path.exists(SO) and torch.cuda.is_available()


def module():
    global _mod
    if _mod is None:
        spec = importlib.util.spec_from_file_location("deform_conv_cuda", SO)
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod


def ref_deform_conv(x: torch.Tensor, offset: torch.Tensor, weight: torch.Tensor, im2col_step: int = 64) -> torch.Tensor:
    assert x.is_cuda and x.dtype == torch.float32
    B, _, H, W = x.shape
    out = x.new_empty((B, weight.shape[0], H, W))                 # deform_conv.py:41-42 (_output_size: same H, W for 3x3 / pad 1)
    bufs = [x.new_empty(0), x.new_empty(0)]                       # :44 columns, ones
    step = min(im2col_step, B)                                    # :49
    assert B % step == 0
    module().deform_conv_forward_cuda(x.contiguous(), weight.contiguous(), offset.contiguous(), out, bufs[0], bufs[1],
                                      weight.size(3), weight.size(2), 1, 1, 1, 1, 1, 1, 1, 1, step)      # :53-58
    return out


def ref_dcn_subnet(sd, x: torch.Tensor, n_layers: int = 3, capture=None) -> torch.Tensor:
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False                       # fp32 reference arithmetic (matmul TF32 is off by default)
    try:
        for i in range(n_layers):
            if capture is not None:
                capture.append(x)
            off = F.conv2d(x, sd[f"{3 * i}.conv_offset.weight"], sd[f"{3 * i}.conv_offset.bias"], padding=1)
            if capture is not None:
                capture.append(off)
            y = ref_deform_conv(x, off, sd[f"{3 * i}.conv.weight"])
            x = F.relu(F.group_norm(y, 32, sd[f"{3 * i + 1}.weight"], sd[f"{3 * i + 1}.bias"], 1e-5))
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    return x
