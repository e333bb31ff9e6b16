"""UPSNetFPN deformable-conv subnet on the metric's FPN levels (1024x2048, T = 2 frames: 256x512 ... 32x64, 256 channels):
this repo's CUDA path (through the C ABI) timed with CUDA events next to the REFERENCE's own op (oracle/_ref, compiled
unmodified) + torch's conv2d / group_norm on the same GPU.  Prints one JSON object; per-kernel breakdown from the library's
launch profiler.  Usage (GPU box): python scripts/dcn_bench.py [--ref] [--levels 0,1,2,3] [--out gpurun_out/dcn_bench.json]"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import slotvps_b200 as sv  # noqa: E402
from slotvps_b200 import synthetic  # noqa: E402

SHAPES = [(256, 512), (128, 256), (64, 128), (32, 64)]


def med(fn, n=7, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(); fn(); a1.record(); a1.synchronize()
        ts.append(a0.elapsed_time(a1))
    return sorted(ts)[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", action="store_true")
    ap.add_argument("--levels", default="0,1,2,3")
    ap.add_argument("--frames", type=int, default=2)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    L = sv.lib()
    levels = [int(v) for v in a.levels.split(",")]
    sd = synthetic.make_dcn_state_dict(11, offset_scale=1.5)
    net = sv.B200DeformSubnet()
    net.load_state_dict(sd)
    net = net.to(dev)
    xs = {l: synthetic.make_fpn_level(20 + l, a.frames, 256, *SHAPES[l]).to(dev) for l in levels}
    res = dict(frames=a.frames, levels={}, chain="256->256->128->128, 3x3 deformable, GroupNorm(32)+ReLU")
    tot = tot_ref = 0.0
    flops_all = 0.0
    for l in levels:
        H, W = SHAPES[l]
        x = xs[l]
        ms = med(lambda: net(x))
        px = a.frames * H * W
        flops = px * 2 * 9 * (256 * 256 + 256 * 128 + 128 * 128) + px * 2 * 9 * 18 * (256 + 256 + 128)   # deformable GEMMs + offset convs
        ent = dict(H=H, W=W, ms=ms, algorithmic_gflop=flops / 1e9, tflops=flops / (ms * 1e-3) / 1e12)
        buf = C.create_string_buffer(1 << 16)
        torch.cuda.synchronize()
        L.slotvps_profile_begin(torch.cuda.current_stream(dev).cuda_stream)
        for _ in range(3):
            net(x)
        L.slotvps_profile_end(buf, len(buf))
        ent["kernels_ms"] = {r.split("\t")[0]: round(float(r.split("\t")[2]) / 3, 4) for r in buf.value.decode().strip().split("\n") if r}
        if a.ref:
            from oracle import ref_dcn            # checker / comparison arm only
            dsd = {k: v.to(dev) for k, v in sd.items()}
            ent["reference_ms"] = med(lambda: ref_dcn.ref_dcn_subnet(dsd, x), n=5, warm=1)
            tot_ref += ent["reference_ms"]
        tot += ms
        flops_all += flops
        res["levels"][str(l)] = ent
    res["total_ms"] = tot
    res["total_algorithmic_gflop"] = flops_all / 1e9
    res["tflops"] = flops_all / (tot * 1e-3) / 1e12
    if a.ref:
        res["reference_total_ms"] = tot_ref
        res["speedup_vs_reference_on_same_gpu"] = tot_ref / tot
    s = json.dumps(res, indent=1)
    print(s)
    if a.out:
        open(a.out, "w").write(s)


if __name__ == "__main__":
    main()
