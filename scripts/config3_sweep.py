#!/usr/bin/env python
"""BASELINE configs[2]: Cityscapes-VPS val-shaped sweep -- V synthetic videos x F frames, one clip per output frame
(frame f pairs with f-1; f=0 with itself, mmdet/datasets/cityscapes_vps.py:262-264), clips sharded over the ranks of
one node, ONE all_gather of the per-clip id maps at the end.  Rank 0 then recomputes a sample of the
clips itself and checks the gathered maps are bit-identical to the single-GPU result.

    python -m torch.distributed.run --nproc-per-node N scripts/config3_sweep.py --videos 50 --frames 6
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import slotvps_b200 as sv  # noqa: E402
from slotvps_b200 import synthetic  # noqa: E402
from slotvps_b200.parallel import gather_id_maps, shard_clips  # noqa: E402


def frame_features(H, W, video, frame):
    return synthetic.make_features(H, W, T=1, video=video, frame=frame)[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--videos", type=int, default=50)
    ap.add_argument("--frames", type=int, default=6)
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--width", type=int, default=2048)
    ap.add_argument("--check", type=int, default=6, help="clips rank 0 recomputes for the bit-identity check")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    H, W, N = a.height, a.width, 100
    model = sv.SlotVPSRetriever(sv.HEAD_KWARGS, N, sv.FUSION_KWARGS)
    model.dynamic_mask_head.load_state_dict(synthetic.make_head_state_dict(0))
    model.load_capsule_params(synthetic.make_capsule_params(0, N))
    model = model.to(dev)
    lg = synthetic.make_fusion_case(0, N, 8, 8)[0].to(dev)
    clips = [(v, f) for v in range(a.videos) for f in range(a.frames)]
    mine = shard_clips(len(clips), rank, world)

    def run_clip(v, f, out):
        cur = [x.to(dev) for x in frame_features(H, W, v, f)]
        ref = cur if f == 0 else [x.to(dev) for x in frame_features(H, W, v, f - 1)]
        return model([ref, cur], (H, W), fusion_logits=lg, panoptic_out=out)

    local_maps = torch.empty((len(mine), H, W), dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i, c in enumerate(mine):
        run_clip(*clips[c], local_maps[i])
    allmaps = gather_id_maps(local_maps, len(clips), dist if world > 1 else None)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        step = max(1, len(clips) // max(1, a.check))
        bad = 0
        for c in list(range(0, len(clips), step))[:a.check]:
            again = torch.empty((H, W), dtype=torch.int64, device=dev)
            run_clip(*clips[c], again)
            bad += int((again != allmaps[c]).sum())
        print(json.dumps(dict(config="configs[2] sweep", videos=a.videos, frames_per_video=a.frames, clips=len(clips), size=[H, W], n_gpus=world,
                              wall_s=dt, output_frames_per_s=len(clips) / dt, retriever_frames_per_s=2 * len(clips) / dt,
                              note="eager, synthetic features generated on the host per clip inside the timed loop (not a kernel benchmark)",
                              checked_clips=a.check, mismatching_pixels_vs_single_gpu=bad, gathered_shape=list(allmaps.shape))))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
