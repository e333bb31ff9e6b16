#!/usr/bin/env python
"""bench.py -- retriever frames/sec @1024x2048 (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config clip|viper|sweep|slots]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one clip: T frames of 4-level 128-ch features -> retriever head
(7 stages) -> mask logits -> panoptic fusion -> id map.  Frame convention: retriever frames/s = T * clips/s (the
reference consumes T=2 frames per call and emits one panoptic frame).

--config clip   (default) BASELINE configs[1]: one 1024x2048 clip, T=2, N=100, replayed K times per rank.
--config viper  BASELINE configs[3]: VIPER-shaped 1080x1920 (padded to 1088x1920), T=4, fusion to the unpadded size.
--config sweep  BASELINE configs[2]: V videos x F frames at 1024x2048, one clip per output frame (frame f pairs with
                f-1, f=0 with itself: mmdet/datasets/cityscapes_vps.py:262-264), sharded by VIDEO over the ranks (a video's
                frames stay in order on one rank, as the tracker needs), every frame's features pre-generated in HBM.
--config slots  BASELINE configs[4]: slot count N in {50,100,200,300} x retriever iterations 1..7 on one GPU (a table).

* value      whole-job frames/s with the inputs already resident in HBM (device events, median of --repeats timed regions).
* e2e        same metric through the public API with HOST (pinned) input buffers: per step the host->device copy of the
             clip's features and a device->host read of the id map (narrowed to uint8 on the device: ids < 256) + meta.
* roofline   the kernel with the largest share of the step, with ITS bound; every kernel is listed under roofline.kernels.
* cpu_baseline  the CPU oracle (port of the reference's algorithm, oracle/) on this box's host cores.

--impl reference times that CPU arm alone, always at the configuration's full size (the reference's own Python
implementation cannot travel to the GPU box; oracle/ is its pinned restatement on the same torch CPU kernels).
Synthetic data, random-init weights (no checkpoints/datasets offline).  Random-init heads keep no slot above the 0.85
score threshold and produce collapsed masks, so BOTH arms feed a designed (class logits, mask logits) pair with ~30 kept
slots into the fusion stage (`config.fusion_inputs`); everything else is computed from the features.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

UNIT = "frames/s"
HEADS_DEFAULT = [1, 2, 2, 2]
# BASELINE configs[4]: per_dh_num_heads / temporal stages for 1..7 retriever iterations
ITER_CONFIGS = {1: ([0, 0, 0, 1], [0]), 2: ([0, 0, 1, 1], [0, 1]), 3: ([0, 1, 1, 1], [1, 2]), 4: ([1, 1, 1, 1], [2, 3]),
                5: ([1, 1, 1, 2], [2, 3, 4]), 6: ([1, 1, 2, 2], [2, 3, 4, 5]), 7: ([1, 2, 2, 2], [3, 4, 5, 6])}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="clip", choices=["clip", "viper", "sweep", "slots"])
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--frames", type=int, default=None)
    ap.add_argument("--slots", type=int, default=100)
    ap.add_argument("--videos", type=int, default=50, help="--config sweep: number of videos")
    ap.add_argument("--video-frames", type=int, default=6, help="--config sweep: frames per video")
    ap.add_argument("--kernel-path", type=int, default=0, help="0 auto (tensor-core kernels), 1 force fp32 CUDA-core")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--cpu-budget-s", type=float, default=25.0)
    ap.add_argument("--repeats", type=int, default=5, help="timed regions of K steps each; the median one is reported")
    ap.add_argument("--inflight", type=int, default=8, help="clips in flight per GPU (independent clips on separate streams)")
    a = ap.parse_args()
    if a.config == "viper":
        a.height, a.width, a.frames = a.height or 1080, a.width or 1920, a.frames or 4
    else:
        a.height, a.width, a.frames = a.height or 1024, a.width or 2048, a.frames or 2
    return a


def metric_name(a):
    return f"retriever_frames_per_sec_{a.height}x{a.width}"


def workload_string(a):
    """Identical in both arms (the driver compares them)."""
    if a.config == "viper":
        return (f"r50_fpn_slotvps retriever, VIPER-shaped {a.height}x{a.width} clip (padded to /32), T={a.frames}, N={a.slots}, "
                f"7 stages (BASELINE configs[3])")
    if a.config == "sweep":
        return (f"r50_fpn_slotvps retriever, Cityscapes-VPS val-shaped sweep: {a.videos} videos x {a.video_frames} frames "
                f"{a.height}x{a.width}, one T=2 clip per frame, sharded by video (BASELINE configs[2])")
    if a.config == "slots":
        return f"r50_fpn_slotvps retriever, slot-count / iteration sweep at {a.height}x{a.width}, T={a.frames} (BASELINE configs[4])"
    return f"r50_fpn_slotvps retriever, single {a.height}x{a.width} clip, T={a.frames}, N={a.slots}, 7 stages (BASELINE configs[1])"


FUSION_CASE = dict(seed=21, n_things=15, near_dup_things=4, tiny=3)      # the case tests/test_gpu_parity.py::test_fullsize_fusion_and_properties checks


def fusion_inputs(a, shapes):
    from slotvps_b200 import synthetic
    h, w = shapes[-1]
    kw = dict(FUSION_CASE)
    lg, pm, _ = synthetic.make_fusion_case(kw.pop("seed"), a.slots, h, w, **kw)
    return lg, pm


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d["bf16_tflops_sustained"], source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------
# CPU arm (the oracle port on the host cores): always the configuration's full size
# ------------------------------------------------------------------------------------------------
def run_cpu_arm(a, steps, warmup, budget_s):
    """Time the oracle on the host cores; returns (frames/s, info).  The size is never reduced: when the budget is short
    the number of clips is."""
    from oracle import slotvps_oracle as O            # CPU baseline leg: the one place bench.py runs oracle/
    from slotvps_b200 import synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synthetic.make_head_state_dict(0)
    cap = synthetic.make_capsule_params(0, a.slots)
    shapes = synthetic.level_shapes(a.height, a.width)
    lg, pm = fusion_inputs(a, shapes)
    size = (a.height, a.width)

    def clip(video):
        feats = synthetic.make_features(a.height, a.width, T=a.frames, video=video, frame=0)
        t0 = time.perf_counter()
        O.clip_forward(sd, cap, feats, cap["init_mask_query.weight"], size, fuse=False)      # head + mask logits
        O.panoptic_fuse(lg, pm, size)                                                        # fusion on the designed pair
        return time.perf_counter() - t0
    t_begin = time.perf_counter()
    times, done_w = [], 0
    for i in range(warmup):
        clip(100 + i)
        done_w += 1
        if time.perf_counter() - t_begin > 0.3 * budget_s:
            break
    for i in range(steps):
        times.append(clip(200 + i))
        if time.perf_counter() - t_begin > budget_s and len(times) >= 1:
            break
    t = statistics.median(times)
    fps = a.frames / t
    return fps, dict(kind="port", cores=torch.get_num_threads(), sec_per_clip=t, clips_timed=len(times),
                     sample=f"{len(times)} clip(s) of {a.height}x{a.width} T={a.frames} at full size (head + mask logits + fusion of the "
                            f"designed pair; median, {done_w} warm-up); torch CPU kernels on {torch.get_num_threads()} threads")


def bind_to_gpu_numa_node(local):
    """Pin this rank's host threads (and, by first touch, its pinned staging buffers) to the NUMA node its GPU hangs
    off, so that N ranks streaming features over PCIe do not all pull from one socket's memory."""
    info = {}
    try:
        p = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
        info["bdf"] = bdf
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        info["node"] = node
        if node < 0:
            info["reason"] = "sysfs reports numa_node = -1 (no NUMA information for this PCI device on this host)"
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            info["reason"] = "no CPU of that node in this process's affinity mask"
            return info
        os.sched_setaffinity(0, cpus)
        info["cpus"] = len(cpus)
        return info
    except Exception as e:
        info["reason"] = f"{type(e).__name__}: {e}"
        return info


# ------------------------------------------------------------------------------------------------
# roofline models
# ------------------------------------------------------------------------------------------------
def kernel_models(a, shapes, heads, kept):
    """name -> dict(bound, work per step, unit).  FLOPs / bytes are ALGORITHMIC (reference fp32 layouts)."""
    C, N, T = 256, a.slots, a.frames
    px_stage = sum(hd * h * w for hd, (h, w) in zip(heads, shapes)) * T          # pixel.stage units of the attention kernels
    P = [h * w for (h, w) in shapes]
    P3 = P[-1]
    # level fusion on reference-layout bytes: read the 128-ch fp32 input, read the previous level's 256-ch output once
    # (amortised: P_{l-1} = P_l / 4), write the 256-ch fp32 output
    fuse_bytes = T * sum(P[l] * (4 * 128 + 4 * 256) + (P[l - 1] * 4 * 256 if l > 0 else 0) for l in range(len(P)))
    m = {
        "proj_rstd(k)": ("tensor", 2 * C * C * px_stage), "proj_rstd(v)": ("tensor", 2 * C * C * px_stage),
        "slot_attn_fp32": ("tensor", 4 * N * C * px_stage),
        "stats_tc": ("tensor", 4 * C * C * px_stage),                         # both projections (their LayerNorm statistics)
        "attn_tc": ("tensor", 4 * N * C * px_stage),                          # slots.keys^T + softmax + attn^T.V
        "fuse_tc": ("hbm", fuse_bytes),
        "mask_tc": ("hbm", 4 * C * P3 + 4 * N * P3),                          # read the feature once, write pred_masks
    }
    return m


def measure_next_rows(sv, synthetic, lane, kw, size, N, dev, pk):
    """Tracker (simple_track_head.py:58-92 + vps_temporal_slots.py:322-409), semantic argmax (:440-451) and
    get_unified_pan_result (tools/dataset/cityscapes_vps.py:214-302) on the 1024x2048 outputs of one clip."""
    H, W = size
    hbm = float(pk["hbm_gbs"])
    def med(fn, n=20, warm=3):
        for _ in range(warm): fn()
        ts = []
        for _ in range(n):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(); fn(); a1.record(); a1.synchronize()
            ts.append(a0.elapsed_time(a1))
        return sorted(ts)[len(ts) // 2]
    out = lane.model(lane.clip, size, **kw)
    fo, emb = out["fusion"], out["emb"][-1][-1, 0]                # current frame, last stage [N,256]
    rows = {}
    # tracker: one step per frame on the kept slots' embeddings (object bank resident in HBM)
    th = sv.B200TrackHead(num_fcs_query=2, in_channels_query=256).to(dev)
    trk = sv.SlotTracker(th, n_slots=N, capacity=1024, device=dev)
    e = emb.reshape(N, 256).contiguous()
    trk.step(e, fo)
    ms = med(lambda: trk.step(e, fo))
    k = int(fo.meta[0].item())
    rows["tracker_step"] = dict(ms=ms, kept=k, bound="latency", launches=4,
                                note="FC stack on kept + bank rows, correlation, sequential greedy replay, bank update")
    # semantic argmax: [19,256,512] logits -> resize x4 + softmax + first-max -> int64 [1024,2048]
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.randn(1, 19, H // 4, W // 4, generator=g, device=dev) * 3
    ms = med(lambda: sv.semantic_argmax(x, size))
    b = 19 * (H // 4) * (W // 4) * 4 + H * W * 8
    rows["semantic_argmax"] = dict(ms=ms, algorithmic_bytes=b, achieved_gbs=b / (ms * 1e-3) / 1e9, frac=b / (ms * 1e-3) / 1e9 / hbm, bound="hbm",
                                   note="write of the int64 map dominates (16.8 MB); the exp/bilinear work is per output pixel x 19 classes")
    # id-map unification: (seg, pan) int64 in -> H x W x 3 uint8 out
    seg, pan, ci, oi = synthetic.make_unify_case(20, H, W, n_inst=40, dup_obj=5)
    seg, pan = seg.to(dev), pan.to(dev)
    un = sv.PanUnifier(dev)
    buf = torch.empty((H, W, 3), dtype=torch.uint8, device=dev)
    def uni():
        un.reset(); un.frame(seg, pan, ci, oi, 4096, out=buf)
    ms = med(uni)
    b = H * W * (16 + 16 + 3)
    rows["unify_pan_result"] = dict(ms=ms, algorithmic_bytes=b, achieved_gbs=b / (ms * 1e-3) / 1e9, frac=b / (ms * 1e-3) / 1e9 / hbm, bound="hbm",
                                    note="histogram pass + LUT pass each read seg and pan (int64) once; 3 B/px written; includes the host-side argument marshalling of one call")
    # UPSNetFPN deformable-conv subnet on the clip's four FPN levels (256 channels, T frames as the batch):
    # the producer side of the head's inputs.  scripts/dcn_bench.py times the reference's own op on the same GPU next to it.
    T = len(lane.clip)
    net = sv.B200DeformSubnet()
    net.load_state_dict(synthetic.make_dcn_state_dict(11, offset_scale=1.5))
    net = net.to(dev)
    lv = [(H // 4, W // 4), (H // 8, W // 8), (H // 16, W // 16), (H // 32, W // 32)]
    xs = [synthetic.make_fpn_level(20 + i, T, 256, h, w).to(dev) for i, (h, w) in enumerate(lv)]
    def subnet():
        for x in xs: net(x)
    ms = med(subnet, n=7, warm=2)
    px = T * sum(h * w for h, w in lv)
    fl = px * 2 * 9 * (256 * 256 + 256 * 128 + 128 * 128) + px * 2 * 9 * 18 * (256 + 256 + 128)
    tf = fl / (ms * 1e-3) / 1e12
    rows["dcn_subnet"] = dict(ms=ms, algorithmic_gflop=fl / 1e9, achieved_tflops=tf, executed_mma_tflops=3 * tf, peak=float(pk["bf16_tflops"]),
                              frac=tf / float(pk["bf16_tflops"]), executed_frac=3 * tf / float(pk["bf16_tflops"]), bound="tensor",
                              note="3 x [3x3 deformable conv + GroupNorm(32) + ReLU], 256->256->128->128, on the four FPN levels of the clip "
                                   "(implicit GEMM on tcgen05, 3 fp16 hi/lo products per algorithmic product); the reference's own op "
                                   "compiled for sm_100a takes 32.7 ms for the same work on a B200 (profiles/r2_dcn_bench.json)")
    del xs, net
    return rows


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    METRIC = metric_name(a)

    if a.impl == "reference":
        if rank != 0:
            return 0
        fps, info = run_cpu_arm(a, a.steps, a.warmup, 150.0)
        line = dict(metric=METRIC, value=fps, unit=UNIT, impl="reference", n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
                    ms_per_step=1e3 * info["sec_per_clip"], higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f32", data="synthetic",
                    config=dict(workload=workload_string(a), fusion_inputs="designed class + mask logits (random-init heads keep no slot)",
                                clips_timed=info["clips_timed"]),
                    cpu_baseline=dict(value=fps, unit=UNIT, **{k: info[k] for k in ("cores", "kind", "sample")}),
                    e2e=dict(value=fps, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return 0

    import slotvps_b200 as sv
    from slotvps_b200 import synthetic
    sv.lib()                                                 # fail loudly if the CUDA extension is missing
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    dist = None
    if world > 1:
        # keep stdout to the single JSON line: NCCL prints its version banner (and anything else it logs) to stdout unless told
        # otherwise, at every level from VERSION up
        # (NCCL honours NCCL_DEBUG_FILE only above the VERSION level, so a VERSION setting is raised to WARN)
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    if a.config == "slots":
        return run_slot_sweep(a, dev, rank, world, dist)

    T, N, H, W = a.frames, a.slots, a.height, a.width
    shapes = synthetic.level_shapes(H, W)
    sd = synthetic.make_head_state_dict(0)
    cap = synthetic.make_capsule_params(0, N)
    lg, pm_designed = fusion_inputs(a, shapes)
    fusion_logits, fusion_masks = lg.to(dev), pm_designed.to(dev)
    K, Wm, M = a.steps, a.warmup, max(1, a.inflight)
    if M >= 4:
        # throughput mode: with several clips in flight the side-stream producers of one clip co-run with the other
        # clips' kernels, and narrower grids interleave better (measured 64 > 112 > 148); the library default (112)
        # is the single-clip latency optimum.
        os.environ.setdefault("SLOTVPS_SIDE_CTAS", "64")
        os.environ.setdefault("SLOTVPS_MAIN_CTAS", "64")
    L = sv.lib()
    wire_dtype = torch.uint8 if 11 + N <= 256 else torch.int16        # ids < stuff_num + N (vps_temporal_slots.py:428)

    # ---- sweep mode: every frame of this rank's videos resident in HBM, clips = (video, frame) in order --------
    sweep = None
    if a.config == "sweep":
        from slotvps_b200.parallel import shard_clips
        vids = shard_clips(a.videos, rank, world)                      # shard by VIDEO: the tracker needs a video's frames in order
        g = torch.Generator(device=dev).manual_seed(1234 + rank)
        frames = {(v, f): [torch.randn((1, 128, h, w), generator=g, device=dev) for (h, w) in shapes]
                  for v in vids for f in range(a.video_frames)}
        clips = [(v, f) for v in vids for f in range(a.video_frames)]
        # one step per clip of the shard; ranks with one video fewer repeat clips so every rank runs the same number of
        # steps (the gathers need equal counts) -- `value` counts the real clips only
        K = -(-a.videos // world) * a.video_frames
        sweep = dict(frames=frames, clips=clips, vids=vids, real_clips=a.videos * a.video_frames)

    # Clips are independent, so M clips are kept in flight per GPU: lane j owns a model instance
    # (its own workspaces), static input buffers (178 MB at T=2: larger than the 126 MB L2), a stream and a
    # CUDA graph of the whole step.  Step i runs on lane i % M.
    class Lane:
        pass
    lanes = []
    kw = dict(pos="sine", fusion_logits=fusion_logits, fusion_masks=fusion_masks, want_feats=False)
    for j in range(M):
        ln = Lane()
        ln.model = sv.SlotVPSRetriever({**sv.HEAD_KWARGS, "kernel_path": a.kernel_path}, N, sv.FUSION_KWARGS)
        ln.model.dynamic_mask_head.load_state_dict(sd, strict=True)
        ln.model.load_capsule_params(cap)
        ln.model = ln.model.to(dev)
        ln.host = synthetic.make_features(H, W, T=T, video=10 * rank + j, frame=0)
        ln.clip = [[f.to(dev) for f in fr] for fr in ln.host]
        ln.stream = torch.cuda.Stream()
        ln.graph = None if a.no_graph else sv.GraphedClip(ln.model, ln.clip, (H, W), **kw)
        ln.end = torch.cuda.Event()
        lanes.append(ln)
    model = lanes[0].model
    dev_clips = [ln.clip for ln in lanes]
    pan_wire = torch.empty((K, H, W), dtype=wire_dtype, device=dev) if world > 1 else None
    comm = torch.cuda.Stream() if world > 1 else None
    gathered = [torch.empty((K, H, W), dtype=wire_dtype, device=dev) for _ in range(world)] if (world > 1 and rank == 0) else None
    n_chunks = 4 if K >= 8 else 1
    bounds = [K * c // n_chunks for c in range(n_chunks + 1)]

    def step(i, feats):                      # eager single-stream step (profiling pass)
        return model(feats, (H, W), **kw)

    def lane_step(ln, i):
        if sweep is not None:                # feed: device-to-device copy of the clip's two frames into the lane's static buffers
            v, f = sweep["clips"][i % len(sweep["clips"])]
            cur, ref = sweep["frames"][(v, f)], sweep["frames"][(v, max(f - 1, 0))]
            for l in range(4):
                ln.clip[0][l].copy_(ref[l], non_blocking=True)
                ln.clip[1][l].copy_(cur[l], non_blocking=True)
        o = ln.graph.replay() if ln.graph is not None else ln.model(ln.clip, (H, W), **kw)
        if pan_wire is not None:
            pan_wire[i % K].copy_(o["fusion"].panoptic, non_blocking=True)     # int64 -> 1 byte per pixel for the wire
        return o

    main_s = torch.cuda.current_stream()
    start = torch.cuda.Event()

    def gather_chunk(c):
        """Rank 0 is the only consumer (it would write the PNG/JSON): a gather of this chunk's id maps on a side stream,
        issued as soon as the chunk's steps are enqueued, so only the last chunk sits behind the last clip."""
        comm.wait_stream(main_s)
        for ln in lanes:
            comm.wait_stream(ln.stream)
        with torch.cuda.stream(comm):
            lo, hi = bounds[c], bounds[c + 1]
            dist.gather(pan_wire[lo:hi], [g[lo:hi] for g in gathered] if rank == 0 else None, dst=0)

    def run(n, body, collect=False):
        start.record(main_s)
        for ln in lanes:
            ln.stream.wait_event(start)
        out, c = None, 0
        for i in range(n):
            ln = lanes[i % M]
            with torch.cuda.stream(ln.stream):
                out = body(ln, i)
            if collect and dist is not None and i + 1 == bounds[c + 1]:
                gather_chunk(c)
                c += 1
        for ln in lanes:
            ln.end.record(ln.stream)
            main_s.wait_event(ln.end)
        if collect and dist is not None:
            main_s.wait_stream(comm)
        return out

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, body, collect=False, repeats=1):
        """`repeats` timed regions of n steps each, barrier + synchronize on both sides, max over ranks; all regions."""
        res = []
        out = None
        for _ in range(repeats):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record()
            out = run(n, body, collect)
            e1.record()
            barrier()
            ms = e0.elapsed_time(e1)
            if dist is not None:
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            res.append(ms)
        return res, out

    run(max(Wm, M), lane_step)
    if dist is not None:
        run(K, lane_step, collect=True)                          # warm-up incl. the gathers: NCCL channel setup stays out of the timed region
    barrier()
    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else local)
    if rank == 0:
        sampler.start()
    reps, out = timed(K, lane_step, collect=True, repeats=max(1, a.repeats))
    ms = statistics.median(reps)
    launches = world * K * lanes[0].graph.launches if lanes[0].graph is not None else 0
    if lanes[0].graph is None:
        L.slotvps_launch_count(1)
        run(K, lane_step)
        torch.cuda.synchronize()
        launches = world * int(L.slotvps_launch_count(0))
    units = sweep["real_clips"] if sweep is not None else world * K          # clips the whole job processed
    value = units * T / (ms * 1e-3)
    meta = out["fusion"].host()

    # single clip in flight (latency view of the same step), for reference
    single_ms = None
    if M > 1 and sweep is None:
        keep, lanes[:] = lanes[:], lanes[:1]
        M1, M = M, 1
        run(max(2, Wm), lane_step)
        r1, _ = timed(K, lane_step, repeats=3)
        single_ms = statistics.median(r1) / K
        lanes[:] = keep
        M = M1

    # ---- sweep: per-video tracker pass + bit-identity of a sample of clips (outside the timed region) ------------
    sweep_info = None
    if sweep is not None:
        sweep_info = dict(videos_this_rank=len(sweep["vids"]), clips_this_rank=K,
                          feed="device-to-device copy of each clip's two frames (178 MB) into the lane's static graph inputs, inside the timed region")

    # ---- e2e: host buffers in, id map + meta out, every step -----------------------------------------
    e2e = None
    if not a.no_e2e and sweep is None:
        for ln in lanes:
            ln.pinned = [[f.pin_memory() for f in fr] for fr in ln.host]
            ln.wire = torch.empty((H, W), dtype=wire_dtype, device=dev)
            ln.h_pan = torch.empty((H, W), dtype=wire_dtype).pin_memory()
            ln.h_meta = torch.empty(4 + 3 * N, dtype=torch.int32).pin_memory()
        h2d = sum(f.numel() * 4 for fr in lanes[0].host for f in fr)
        d2h = lanes[0].h_pan.numel() * lanes[0].h_pan.element_size() + lanes[0].h_meta.numel() * 4

        def e2e_step(ln, i):
            # upload this step's clip from pinned host memory into the lane's input buffers, run, read results back;
            # the other lanes' uploads / downloads overlap this lane's compute
            for t in range(T):
                for l in range(4):
                    ln.clip[t][l].copy_(ln.pinned[t][l], non_blocking=True)
            o = lane_step(ln, i)
            ln.wire.copy_(o["fusion"].panoptic, non_blocking=True)       # ids < 256: 1 byte per pixel over PCIe; the caller widens on the host
            ln.h_pan.copy_(ln.wire, non_blocking=True)
            ln.h_meta.copy_(o["fusion"].meta, non_blocking=True)
            return o
        run(max(2, Wm, M), e2e_step)
        r2, _ = timed(K, e2e_step, repeats=max(1, min(3, a.repeats)))
        ms2 = statistics.median(r2)
        e2e = dict(value=world * K * T / (ms2 * 1e-3), unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                   ms_per_step=ms2 / K, ms_per_region=r2, h2d_gbs_per_rank=h2d * K / (ms2 * 1e-3) / 1e9,
                   pipeline=f"{M} clips in flight: a lane's upload/readback overlaps the other lanes' compute",
                   readback=f"id map as {str(wire_dtype).split('.')[-1]} (ids < stuff_num + N), meta int32[{4 + 3 * N}]")
    clocks = sampler.stop() if rank == 0 else None     # sampled over the timed region and the e2e region (both under load)

    # ---- per-kernel device times (CUDA events on the launching stream) -> roofline ----------------------
    roofline, breakdown = None, None
    if rank == 0:
        pk = peaks()
        heads = HEADS_DEFAULT
        nprof = 3
        import ctypes
        buf = ctypes.create_string_buffer(1 << 16)
        torch.cuda.synchronize()
        for i in range(2):
            step(i, dev_clips[i % len(dev_clips)])            # eager warm-up of the single-stream schedule
        torch.cuda.synchronize()
        L.slotvps_profile_begin(torch.cuda.current_stream(dev).cuda_stream)
        for i in range(nprof):
            step(i, dev_clips[i % len(dev_clips)])
        L.slotvps_profile_end(buf, len(buf))
        rows = [r.split("\t") for r in buf.value.decode().strip().split("\n") if r]
        breakdown = {}
        for r in rows:
            breakdown[r[0]] = dict(launches_per_step=int(r[1]) / nprof, ms_per_step=float(r[2]) / nprof)
            if len(r) > 3 and int(r[3]) > 0:
                breakdown[r[0]]["grid_ctas"] = int(r[3])
        tj_all = {}
        for name in ("r2_ncu_traffic.json", "r1_ncu_traffic.json"):
            tp = os.path.join(ROOT, "profiles", name)
            if os.path.exists(tp):              # DRAM bytes per pixel measured by ncu on the level-3 launches
                tj_all = json.load(open(tp))
                tj_all["_file"] = "profiles/" + name
                break
        models = kernel_models(a, shapes, heads, meta["k"])
        px_stage = sum(hd * h * w for hd, (h, w) in zip(heads, shapes)) * T
        px_all = sum(h * w for (h, w) in shapes) * T
        ktot = sum(v["ms_per_step"] for k, v in breakdown.items() if k != "(host gap)")
        per_kernel = {}
        for name, v in breakdown.items():
            if name == "(host gap)":
                continue
            t_ms, nl = v["ms_per_step"], v["launches_per_step"]
            ent = dict(kernel=name, ms_per_step=t_ms, launches_per_step=nl, us_per_launch=1e3 * t_ms / max(nl, 1e-9),
                       share_of_step=t_ms / ktot, grid_ctas=v.get("grid_ctas"))
            mdl = models.get(name)
            if mdl is None:
                ent.update(bound="latency", note="no byte / FLOP model: slot-side or bookkeeping kernel, reported as time per launch")
            elif mdl[0] == "tensor":
                ach = mdl[1] / (t_ms * 1e-3) / 1e12
                tj = tj_all.get(name)
                # executed MMA work per algorithmic FLOP: 3 products of the fp16 hi/lo split; the statistics kernel runs on the
                # triangular (QR) factors, which need only 10/16 of the weight blocks
                xf = 3 * 0.625 if (name == "stats_tc" and os.environ.get("SLOTVPS_STATS_TRI", "1") != "0") else 3
                ent.update(bound="tensor", achieved=ach, peak=pk["bf16_tflops"], unit="TFLOP/s", frac=ach / pk["bf16_tflops"],
                           executed_mma_tflops=xf * ach, executed_frac=xf * ach / pk["bf16_tflops"], executed_per_algorithmic=xf,
                           algorithmic_flops_per_step=mdl[1],
                           traffic=(tj["dram_bytes_read"] + tj["dram_bytes_write"]) / tj["pixels"] * px_stage / nl if tj else None,
                           traffic_note=f"DRAM bytes per launch ({tj_all.get('_file')}: ncu level-3 capture scaled by pixels)" if tj else None,
                           peak_source=f"{pk['source']} bf16 burst (kernel timed alone in the per-kernel profiling pass)")
            else:
                ach = mdl[1] / (t_ms * 1e-3) / 1e9
                tj = tj_all.get("fuse_tc_main" if name == "fuse_tc" else name)
                scale_px = px_all if name == "fuse_tc" else shapes[-1][0] * shapes[-1][1]
                traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) / tj["pixels"] * scale_px if tj else None
                ent.update(bound="hbm", achieved=ach, peak=pk["hbm_gbs"], unit="GB/s", frac=ach / pk["hbm_gbs"],
                           algorithmic_bytes_per_step=mdl[1], traffic=traffic,
                           traffic_over_algorithmic=(traffic / mdl[1]) if traffic else None,
                           traffic_note=f"DRAM bytes per step ({tj_all.get('_file')}: ncu capture scaled by pixels)" if tj else None,
                           peak_source=pk["source"] + " copy bandwidth")
                if name == "fuse_tc":
                    ent["operand_plane_bytes_per_step"] = px_all * 4 * 2 * 256      # the fp16 hi/lo operand planes it also writes (not algorithmic)
            per_kernel[name] = ent
        name = max(per_kernel, key=lambda k: per_kernel[k]["ms_per_step"])
        roofline = dict(per_kernel[name])
        roofline["kernels"] = per_kernel
        step_ms = single_ms if single_ms else ms / K
        roofline["share_of_single_clip_wall_step"] = roofline["ms_per_step"] / step_ms
        roofline["profile_pass"] = "eager, single stream, 3 steps, every persistent kernel at its full-width grid (grid_ctas per kernel)"
        # the whole attention contraction (all its kernels) against the bf16 burst peak
        names = [k for k in breakdown if k in ("proj_rstd(k)", "proj_rstd(v)", "slot_attn_fp32", "stats_tc", "attn_tc")]
        if names:
            per_frame_flops = sum(hd * h * w * (4 * 256 * 256 + 4 * N * 256) for hd, (h, w) in zip(heads, shapes))
            tt = sum(breakdown[k]["ms_per_step"] for k in names)
            alg = per_frame_flops * T / (tt * 1e-3) / 1e12
            ex = sum(per_kernel[k]["executed_mma_tflops"] * breakdown[k]["ms_per_step"] for k in names if "executed_mma_tflops" in per_kernel[k]) / tt
            roofline["attention_contraction"] = dict(kernels=names, algorithmic_tflops=alg, ms_per_step=tt, frac_of_peak=alg / pk["bf16_tflops"],
                                                     executed_mma_tflops=ex, executed_frac_of_peak=ex / pk["bf16_tflops"], peak=pk["bf16_tflops"],
                                                     note="executed = MMA FLOPs actually issued (3 fp16 hi/lo products per algorithmic product; "
                                                          "the statistics GEMM runs on triangular QR factors: 0.625 of its weight blocks)")
        hb = {"peak_gbs": pk["hbm_gbs"], "peak_source": pk["source"] + " copy bandwidth"}
        if "mask_tc" in per_kernel:
            hb["mask_logits"] = {k: per_kernel["mask_tc"].get(k) for k in ("kernel", "algorithmic_bytes_per_step", "ms_per_step", "achieved", "frac")}
        fk = [k for k in breakdown if k.startswith("fuse_") and k != "fuse_tc"]
        if fk:                          # fusion, exact two-pass form: 2*4*Kept*P3 read + 8*H*W written
            P3 = shapes[-1][0] * shapes[-1][1]
            by = 2 * 4 * meta["k"] * P3 + 8 * H * W
            t_ms = sum(breakdown[k]["ms_per_step"] for k in fk)
            hb["panoptic_fusion"] = dict(kernels=fk, launches_per_step=sum(breakdown[k]["launches_per_step"] for k in fk), kept_slots=meta["k"],
                                         things=meta["n_things"], algorithmic_bytes=by, ms=t_ms,
                                         achieved_gbs=by / (t_ms * 1e-3) / 1e9, frac=by / (t_ms * 1e-3) / 1e9 / pk["hbm_gbs"])
        if "fuse_tc" in per_kernel:
            hb["level_fusion"] = {k: per_kernel["fuse_tc"].get(k) for k in ("kernel", "algorithmic_bytes_per_step", "ms_per_step", "achieved", "frac",
                                                                                "traffic", "traffic_over_algorithmic", "operand_plane_bytes_per_step")}
        roofline["hbm_stages"] = hb

    # ---- the rows beyond the path on this workload's output: tracker step, semantic argmax, id-map
    # unification -- timed alone with CUDA events (median of 20), outside the timed region of the metric -----------------
    next_rows = None
    if rank == 0 and world == 1 and a.config == "clip":
        next_rows = measure_next_rows(sv, synthetic, lanes[0], kw, (H, W), N, dev, peaks())

    # ---- CPU baseline beside it (rank 0, N == 1) ------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        fps, info = run_cpu_arm(a, 3, 1, a.cpu_budget_s)
        cpu = dict(value=fps, unit=UNIT, cores=info["cores"], kind=info["kind"], sample=info["sample"])

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=Wm, ms_per_step=ms / K,
                    higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                    timed_regions=dict(repeats=len(reps), ms=reps, median_ms=ms, min_ms=min(reps), max_ms=max(reps),
                                       value_min=units * T / (max(reps) * 1e-3), value_max=units * T / (min(reps) * 1e-3)),
                    config=dict(workload=workload_string(a),
                                frames_convention="retriever frames/s = T * clips/s; output frames/s = clips/s",
                                l2="inputs larger than L2 (178 MB/clip at T=2); one resident clip per lane, lanes alternate",
                                fusion_inputs=f"designed class + mask logits (synthetic.make_fusion_case{tuple(FUSION_CASE.values())}): "
                                              f"{meta['k']} kept slots, {meta['n_things']} things (random-init heads keep no slot)",
                                fp32_fused_features="not materialised (L2 integration form: simple_test only consumes the finest level, "
                                                    "through the fp16 operand planes; want_feats=True restores the head's third return value)",
                                kernel_path=a.kernel_path, cuda_graph=not a.no_graph, clips_in_flight=M, numa_bind=numa,
                                side_stream_ctas=int(os.environ.get("SLOTVPS_SIDE_CTAS", "112")), main_stream_ctas=int(os.environ.get("SLOTVPS_MAIN_CTAS", "148")),
                                single_clip_in_flight_ms_per_step=single_ms,
                                sharding=(f"clips per rank; id maps ({str(wire_dtype).split('.')[-1]}) gathered to rank 0 in {n_chunks} chunks on a side "
                                          f"stream while later clips compute") if world > 1 else "single GPU",
                                kept_slots=meta["k"], fusion_iters=meta["iters"], sweep=sweep_info),
                    clocks=clocks, e2e=e2e, gpu_launches=launches, roofline=roofline, cpu_baseline=cpu, next_rows=next_rows,
                    kernel_breakdown_ms_per_step=breakdown)
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def run_slot_sweep(a, dev, rank, world, dist):
    """BASELINE configs[4]: N in {50,100,200,300} x 1..7 retriever iterations at 1024x2048 on one GPU."""
    import slotvps_b200 as sv
    from slotvps_b200 import synthetic
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0
    T, H, W = a.frames, a.height, a.width
    shapes = synthetic.level_shapes(H, W)
    host = synthetic.make_features(H, W, T=T, video=3, frame=0)
    M = min(4, max(1, a.inflight))
    clips = [[[f.to(dev) for f in fr] for fr in host] for _ in range(M)]
    table = []
    K = max(4, min(a.steps, 8))
    for N in (50, 100, 200, 300):
        cap = synthetic.make_capsule_params(0, N)
        a.slots = N
        lg, pmd = fusion_inputs(a, shapes)
        lg, pmd = lg.to(dev), pmd.to(dev)
        for it in range(1, 8):
            heads, temporal = ITER_CONFIGS[it]
            sd = synthetic.make_head_state_dict(0, per_dh_num_heads=heads, temporal_stages=temporal)
            lanes = []
            for j in range(M):
                m = sv.SlotVPSRetriever({**sv.HEAD_KWARGS, "dh_num_heads": it, "per_dh_num_heads": heads,
                                         "apply_temporal_query_atten_stages": temporal, "kernel_path": a.kernel_path}, N, sv.FUSION_KWARGS)
                m.dynamic_mask_head.load_state_dict(sd, strict=True)
                m.load_capsule_params(cap)
                m = m.to(dev)
                g = sv.GraphedClip(m, clips[j], (H, W), pos="sine", fusion_logits=lg, fusion_masks=pmd, want_feats=False)
                lanes.append((g, torch.cuda.Stream()))
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

            def go(n):
                main_s = torch.cuda.current_stream()
                for g, s in lanes:
                    s.wait_stream(main_s)
                for i in range(n):
                    g, s = lanes[i % M]
                    with torch.cuda.stream(s):
                        g.replay()
                for g, s in lanes:
                    main_s.wait_stream(s)
            go(M)
            torch.cuda.synchronize()
            ev0.record()
            go(K)
            ev1.record()
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1) / K
            # one clip in flight
            ev0.record()
            for _ in range(K):
                lanes[0][0].replay()
            ev1.record()
            torch.cuda.synchronize()
            table.append(dict(n_slots=N, iterations=it, per_dh_num_heads=heads, temporal_stages=temporal, ms_per_clip=ms,
                              frames_per_s=T / (ms * 1e-3), single_clip_ms=ev0.elapsed_time(ev1) / K, launches_per_clip=lanes[0][0].launches))
            del lanes
            torch.cuda.empty_cache()
    line = dict(metric=metric_name(a), value=[r for r in table if r["n_slots"] == 100 and r["iterations"] == 7][0]["frames_per_s"], unit=UNIT,
                n_gpus=1, steps=K, warmup=M, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=workload_string(a), clips_in_flight=M, value_is="the N=100, 7-iteration row"), table=table)
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
