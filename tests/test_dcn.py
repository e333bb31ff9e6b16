"""UPSNetFPN deformable-convolution subnet.

CPU (-m "not gpu"): the oracle restatement against (i) outputs of the reference's own compiled op frozen on the B200
(tests/golden/dcn_*.npz, made by tests/golden/make_golden_dcn.py) and (ii) torchvision.ops.deform_conv2d; host-side contract.
GPU (-m gpu): the CUDA path through the C ABI against the fp64 oracle, the reference golden, and -- when oracle/_ref holds the
compiled reference op -- the reference itself running on the same B200, up to the full 1024x2048 FPN level sizes."""
import ctypes as C
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import slotvps_oracle as O
from slotvps_b200 import _lib, synthetic

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
G = importlib.import_module("make_golden_dcn")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

DRIFT_X = 5.0          # free-running error allowed as a multiple of the fp32 oracle's own error against fp64 (as test_gpu_parity)


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm())


def golden(name):
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLDEN, name + ".npz")).items()}


# ---------------------------------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("name", sorted(G.OP_CASES))
def test_oracle_op_matches_reference_golden(name):
    x, off, w = G.op_inputs(*G.OP_CASES[name])
    g = golden(name)["out"]
    e32, e64 = rel(O.deform_conv(x, off, w), g), rel(O.deform_conv(x.double(), off.double(), w.double()), g)
    print(f"{name}: oracle fp32 vs reference op {e32:.2e}, fp64 {e64:.2e}")
    assert e32 < 2e-6 and e64 < 2e-6


@pytest.mark.parametrize("name", sorted(G.NET_CASES))
def test_oracle_subnet_matches_reference_golden(name):
    seed, B, H, W, scale = G.NET_CASES[name]
    sd = synthetic.make_dcn_state_dict(seed, offset_scale=scale)
    x = synthetic.make_fpn_level(seed, B, 256, H, W)
    g = golden(name)
    # teacher-forced: every layer from the REFERENCE's own input of that layer
    ins = [x, g["in1"], g["in2"]]
    for i in range(3):
        off = F.conv2d(ins[i], sd[f"{3 * i}.conv_offset.weight"], sd[f"{3 * i}.conv_offset.bias"], padding=1)
        assert rel(off, g[f"off{i}"]) < 2e-6
        one = {k[len(str(3 * i)):] if k.startswith(f"{3 * i}.") else "1" + k[len(str(3 * i + 1)):]: v for k, v in sd.items()
               if k.startswith((f"{3 * i}.", f"{3 * i + 1}."))}
        one = {("0" + k if k.startswith(".") else k): v for k, v in one.items()}
        y = O.dcn_subnet(one, ins[i], n_layers=1)
        want = g["out"] if i == 2 else ins[i + 1]
        e = rel(y, want)
        print(f"{name}: layer {i} teacher-forced oracle fp32 vs reference {e:.2e}")
        assert e < 5e-6
    e = rel(O.dcn_subnet(sd, x), g["out"])
    print(f"{name}: free-running oracle fp32 vs reference {e:.2e}")
    assert e < 5e-5


def test_oracle_op_matches_torchvision():
    tv = pytest.importorskip("torchvision")
    for name, cfg in G.OP_CASES.items():
        x, off, w = G.op_inputs(*cfg)
        a = O.deform_conv(x.double(), off.double(), w.double())
        b = tv.ops.deform_conv2d(x.double(), off.double(), w.double(), padding=1)
        assert rel(a, b) < 1e-12
    # zero offsets: the ordinary convolution
    assert rel(O.deform_conv(x.double(), torch.zeros_like(off).double(), w.double()), F.conv2d(x.double(), w.double(), padding=1)) < 1e-12


def test_subnet_mirror_contract():
    from slotvps_b200.dcn import B200DeformSubnet, deform_conv
    m = B200DeformSubnet()
    sd = synthetic.make_dcn_state_dict(0)
    assert list(m.state_dict().keys()) == list(sd.keys())                     # nn.Sequential keys of upsnetFPN.py:36-49
    m.load_state_dict(sd, strict=True)
    assert [tuple(v.shape) for v in m.state_dict().values()] == [tuple(v.shape) for v in sd.values()]
    with pytest.raises(Exception):                                            # no CPU path
        m(torch.zeros(1, 256, 8, 8))
    with pytest.raises(Exception):
        deform_conv(torch.zeros(1, 64, 8, 8), torch.zeros(1, 18, 8, 8), torch.zeros(32, 64, 3, 3))


def test_abi_rejects_unserved_instances():
    L = _lib.lib()
    n = C.c_size_t()
    one = C.c_void_p(16)          # non-null dummy pointers: validation happens before any device access
    args = lambda **kw: [one, one, one, one, kw.get("B", 1), kw.get("cin", 64), kw.get("cout", 32), 8, 8, kw.get("k", 3), kw.get("k", 3),
                         kw.get("st", 1), kw.get("st", 1), kw.get("pad", 1), kw.get("pad", 1), kw.get("dil", 1), kw.get("dil", 1),
                         kw.get("group", 1), kw.get("dg", 1), 64, one, 1 << 30, None]
    for bad in (dict(k=5), dict(st=2), dict(pad=0), dict(dil=2), dict(group=2), dict(dg=2), dict(cin=100), dict(cout=40), dict(cin=512)):
        assert L.slotvps_deform_conv_forward(*args(**bad)) == -1, bad
        assert L.slotvps_last_error()
    rows = 2 * 64 * 128
    assert L.slotvps_deform_conv_workspace_bytes(2, 256, 64, 128, C.byref(n)) == 0
    # pixel-major input + raw output + weight planes, and NO column buffer (the reference's `columns` is rows x 9 x 256 floats)
    assert rows * (256 + 256) * 4 <= n.value < rows * 9 * 256 * 4
    lay = (_lib.DcnLayer * 2)(_lib.DcnLayer(256, 256, 16, 16, 16, 16, 16), _lib.DcnLayer(128, 128, 16, 16, 16, 16, 16))
    assert L.slotvps_dcn_prepared_bytes(lay, 2, C.byref(n)) == -1             # c_in != previous c_out


# ---------------------------------------------------------------------------------------------------------- GPU
def _cuda_op(x, off, w):
    from slotvps_b200.dcn import deform_conv
    dev = torch.device("cuda:0")
    return deform_conv(x.to(dev), None if off is None else off.to(dev), w.to(dev))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(G.OP_CASES))
def test_deform_conv_vs_oracle_and_golden(name):
    x, off, w = G.op_inputs(*G.OP_CASES[name])
    y = _cuda_op(x, off, w)
    truth = O.deform_conv(x.double(), off.double(), w.double())
    e, n32, eg = rel(y, truth), rel(O.deform_conv(x, off, w), truth), rel(y, golden(name)["out"])
    print(f"{name}: CUDA vs fp64 oracle {e:.2e} (fp32 oracle {n32:.2e}); vs reference op golden {eg:.2e}")
    assert e < 1e-5 and eg < 1e-5


@pytest.mark.gpu
def test_deform_conv_zero_and_null_offset_is_convolution():
    x, off, w = G.op_inputs(7, 2, 128, 64, 19, 23, 1.0)
    want = F.conv2d(x.double(), w.double(), padding=1)
    assert rel(_cuda_op(x, torch.zeros_like(off), w), want) < 1e-5
    assert rel(_cuda_op(x, None, w), want) < 1e-5


@pytest.mark.gpu
def test_deform_conv_far_offsets_read_zero():
    """Offsets that leave the map entirely (deform_conv_cuda_kernel.cu:224) give exact zeros; a tap landing exactly on row -1 / H too."""
    x, off, w = G.op_inputs(8, 1, 64, 32, 9, 11, 1.0)
    assert float(_cuda_op(x, torch.full_like(off, 1000.0), w).abs().max()) == 0.0
    assert float(_cuda_op(x, torch.full_like(off, -1000.0), w).abs().max()) == 0.0


@pytest.mark.gpu
def test_first_form_matches_implicit_gemm(monkeypatch):
    """SLOTVPS_DCN_IM2COL=1 (bilinear im2col to fp16 planes in HBM + plain tensor-core GEMM, the A/B baseline of the implicit-GEMM
    kernel): same operand values, same products -- the two forms agree to accumulation-order noise."""
    x, off, w = G.op_inputs(9, 2, 256, 128, 21, 37, 2.0)             # 1554 pixels: partial tile
    a = _cuda_op(x, off, w)
    monkeypatch.setenv("SLOTVPS_DCN_IM2COL", "1")
    b = _cuda_op(x, off, w)
    monkeypatch.delenv("SLOTVPS_DCN_IM2COL")
    e = rel(a, b)
    print(f"implicit GEMM vs im2col + GEMM form: rel {e:.2e}")
    assert e < 2e-6 and rel(a, O.deform_conv(x.double(), off.double(), w.double())) < 1e-5


def _subnet(sd, channels=None):
    from slotvps_b200.dcn import B200DeformSubnet
    m = B200DeformSubnet(channels=channels)
    m.load_state_dict(sd, strict=True)
    return m.to("cuda:0")


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(G.NET_CASES))
def test_subnet_vs_oracle_and_golden(name):
    seed, B, H, W, scale = G.NET_CASES[name]
    sd = synthetic.make_dcn_state_dict(seed, offset_scale=scale)
    x = synthetic.make_fpn_level(seed, B, 256, H, W)
    g = golden(name)
    # teacher-forced per layer on the REFERENCE's own layer inputs
    ins, chans = [x, g["in1"], g["in2"]], [(256, 256), (256, 128), (128, 128)]
    for i in range(3):
        one = {("0" + k[len(str(3 * i)):]) if k.startswith(f"{3 * i}.") else ("1" + k[len(str(3 * i + 1)):]): v for k, v in sd.items()
               if k.startswith((f"{3 * i}.", f"{3 * i + 1}."))}
        y = _subnet(one, [chans[i]])(ins[i].to("cuda:0"))
        want = g["out"] if i == 2 else ins[i + 1]
        truth = O.dcn_subnet({k: v.double() for k, v in one.items()}, ins[i].double(), n_layers=1)
        print(f"{name}: layer {i} teacher-forced CUDA vs reference golden {rel(y, want):.2e}, vs fp64 oracle {rel(y, truth):.2e}")
        assert rel(y, want) < 1e-5 and rel(y, truth) < 1e-5
    y = _subnet(sd)(x.to("cuda:0"))
    truth = O.dcn_subnet({k: v.double() for k, v in sd.items()}, x.double())
    # the noise floor of fp32 arithmetic on this chain: the fp32 oracle (CPU kernels) and the REFERENCE's own GPU output, both vs fp64
    e, n32, nref, eg = rel(y, truth), rel(O.dcn_subnet(sd, x), truth), rel(g["out"], truth), rel(y, g["out"])
    print(f"{name}: free-running CUDA vs fp64 oracle {e:.2e} (fp32 oracle {n32:.2e}, reference-on-B200 golden {nref:.2e}); "
          f"vs reference golden {eg:.2e}")
    assert e <= DRIFT_X * max(n32, nref, 2e-6)


@pytest.mark.gpu
def test_subnet_fpn_levels_vs_reference_live():
    """The four FPN levels of the metric's own config (1024x2048, T = 2: 256x512 ... 32x64) against the reference op running on
    the same GPU (oracle/_ref); without the compiled reference the two coarse levels are checked against the fp64 oracle."""
    from oracle import ref_dcn
    sd = synthetic.make_dcn_state_dict(11, offset_scale=1.5)
    net = _subnet(sd)
    dsd = {k: v.to("cuda:0") for k, v in sd.items()}
    for lvl, (H, W) in enumerate([(256, 512), (128, 256), (64, 128), (32, 64)]):
        x = synthetic.make_fpn_level(20 + lvl, 2, 256, H, W)
        y = net(x.to("cuda:0"))
        y2 = net(x.to("cuda:0"))
        assert torch.equal(y, y2) and bool(torch.isfinite(y).all())              # deterministic
        if ref_dcn.available():
            want = ref_dcn.ref_dcn_subnet(dsd, x.to("cuda:0"))
            e = rel(y, want)
            print(f"level {lvl} ({H}x{W}): CUDA subnet vs REFERENCE subnet on this GPU {e:.2e}")
            assert e < 2e-5
        if lvl >= 2:
            truth = O.dcn_subnet({k: v.double() for k, v in sd.items()}, x.double())
            e, n32 = rel(y, truth), rel(O.dcn_subnet(sd, x), truth)
            print(f"level {lvl} ({H}x{W}): CUDA subnet vs fp64 oracle {e:.2e} (fp32 oracle {n32:.2e})")
            assert e <= DRIFT_X * max(n32, 2e-6)
