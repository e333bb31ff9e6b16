"""Freeze outputs of the REFERENCE's deformable-convolution op as fixtures (tests/golden/dcn_*.npz).

The op is CUDA-only, so unlike make_golden.py this script runs ON THE B200 BOX with the reference op compiled unmodified into
oracle/_ref/deform_conv_cuda.so (oracle/build_ref_dcn.sh):

    gpurun -- 'python tests/golden/make_golden_dcn.py gpurun_out/golden_dcn'      # then copy the .npz files to tests/golden/

Inputs are regenerated from seeds by slotvps_b200.synthetic, so the fixtures hold outputs only (fp16-rounded inputs are not
needed).  ``--check DIR`` compares freshly computed outputs with the committed fixtures instead of writing."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_dcn  # noqa: E402
from slotvps_b200 import synthetic  # noqa: E402

# name -> (seed, B, c_in, c_out, H, W, offset_scale)
OP_CASES = {"dcn_op_a": (1, 2, 64, 32, 13, 17, 2.0), "dcn_op_b": (2, 1, 256, 128, 24, 40, 1.0), "dcn_op_c": (3, 2, 128, 128, 16, 32, 6.0)}
# name -> (seed, B, H, W, offset_scale): the shipped subnet 256 -> 256 -> 128 -> 128
NET_CASES = {"dcn_net_a": (4, 2, 16, 32, 1.0), "dcn_net_b": (5, 1, 20, 28, 3.0)}


def op_inputs(seed, B, cin, cout, H, W, scale):
    g = torch.Generator().manual_seed(97_000 + seed)
    x = synthetic.make_fpn_level(seed, B, cin, H, W)
    off = torch.randn((B, 18, H, W), generator=g) * scale
    w = (torch.rand((cout, cin, 3, 3), generator=g) * 2 - 1) / (cin * 9) ** 0.5
    return x, off, w


def main():
    out_dir = sys.argv[-1]
    check = "--check" in sys.argv
    dev = torch.device("cuda:0")
    os.makedirs(out_dir, exist_ok=True)
    for name, cfg in OP_CASES.items():
        x, off, w = op_inputs(*cfg)
        y = ref_dcn.ref_deform_conv(x.to(dev), off.to(dev), w.to(dev)).cpu().numpy()
        save(out_dir, name, check, out=y)
    for name, (seed, B, H, W, scale) in NET_CASES.items():
        sd = synthetic.make_dcn_state_dict(seed, offset_scale=scale)
        x = synthetic.make_fpn_level(seed, B, 256, H, W)
        cap = []
        y = ref_dcn.ref_dcn_subnet({k: v.to(dev) for k, v in sd.items()}, x.to(dev), capture=cap)
        arrs = {"out": y.cpu().numpy()}
        for i in range(3):
            if i:
                arrs[f"in{i}"] = cap[2 * i].cpu().numpy()          # in0 is the seeded input itself
            arrs[f"off{i}"] = cap[2 * i + 1].cpu().numpy()
        save(out_dir, name, check, **arrs)


def save(out_dir, name, check, **arrs):
    path = os.path.join(out_dir, name + ".npz")
    if check:
        old = np.load(path)
        for k, v in arrs.items():
            d = float(np.abs(old[k].astype(np.float64) - v).max())
            print(f"{name}.{k}: max |diff| vs committed {d:.3e}")
    else:
        np.savez_compressed(path, **{k: v.astype(np.float32) for k, v in arrs.items()})
        print("wrote", path, {k: v.shape for k, v in arrs.items()})


if __name__ == "__main__":
    main()
