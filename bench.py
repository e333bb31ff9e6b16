#!/usr/bin/env python
"""bench.py -- retriever frames/sec @1024x2048 (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one clip: T=2 frames of 4-level 128-ch features at
1024x2048 -> retriever head (7 stages) -> mask logits -> panoptic fusion -> int64 id map
(BASELINE.json configs[1]).  Frame convention: retriever frames/s = T * clips/s (the reference
consumes T=2 frames per call and emits one panoptic frame).

* value      whole-job frames/s with the inputs already resident in HBM (device events).
* e2e        same metric through the public API with HOST (pinned) input buffers: per step the
             host->device copy of the clip's features and a device->host read of the id map + meta.
* roofline   dominant contraction kernel: algorithmic FLOP / CUDA-event time of its launches.
* cpu_baseline  the CPU oracle (port of the reference's algorithm, oracle/) on this box's host cores.

--impl reference times that CPU oracle arm alone (the reference's own CPU implementation cannot
travel to the GPU box; oracle/ is its pinned restatement on the same torch CPU kernels).
Synthetic data, random-init weights (no checkpoints/datasets offline).  Random-init heads keep no
slot above the 0.85 score threshold, so BOTH arms feed designed class logits into the fusion stage
(`config.fusion_logits`); everything else is computed from the features.
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

METRIC = "retriever_frames_per_sec_1024x2048"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--width", type=int, default=2048)
    ap.add_argument("--frames", type=int, default=2)
    ap.add_argument("--slots", type=int, default=100)
    ap.add_argument("--kernel-path", type=int, default=0, help="0 auto (tensor-core kernels), 1 force fp32 CUDA-core")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--cpu-budget-s", type=float, default=25.0)
    ap.add_argument("--inflight", type=int, default=8, help="clips in flight per GPU (independent clips on separate streams)")
    return ap.parse_args()


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
def algorithmic_flops(args):
    """attention contraction per frame = sum_l heads_l * P_l * (4 C^2 + 4 N C)."""
    from slotvps_b200 import synthetic
    shapes = synthetic.level_shapes(args.height, args.width)
    heads = [1, 2, 2, 2]
    C, N = 256, args.slots
    per_frame = sum(hd * h * w * (4 * C * C + 4 * N * C) for hd, (h, w) in zip(heads, shapes))
    return per_frame, shapes, heads


def kernel_flops_per_step(name, args, shapes, heads):
    """Algorithmic FLOPs one step (all launches of `name`) performs; None for kernels without a model."""
    C, N, T = 256, args.slots, args.frames
    px = sum(hd * h * w for hd, (h, w) in zip(heads, shapes)) * T
    table = {
        "proj_rstd(k)": 2 * C * C * px, "proj_rstd(v)": 2 * C * C * px,      # K / V projection (LayerNorm statistics)
        "slot_attn_fp32": 4 * N * C * px,                                    # slots.keys^T + attn^T.V
        "stats_tc": 4 * C * C * px,                                          # both projections (their LayerNorm statistics), tcgen05
        "attn_tc": 4 * N * C * px,                                           # slots.keys^T + softmax + attn^T.V, tcgen05
    }
    return table.get(name)


def cpu_clip(args, H, W, sd, cap, fusion_logits, video):
    from oracle import slotvps_oracle as O            # CPU baseline leg: the one place bench.py runs oracle/
    from slotvps_b200 import synthetic
    feats = synthetic.make_features(H, W, T=args.frames, video=video, frame=0)
    t0 = time.perf_counter()
    out = O.clip_forward(sd, cap, feats, cap["init_mask_query.weight"], (H, W), fuse=False)
    O.panoptic_fuse(fusion_logits, out["pred_masks"], (H, W))
    return time.perf_counter() - t0


def run_cpu_arm(args, steps, warmup, budget_s):
    """Time the oracle on the host cores over a bounded sample; returns (frames/s @1024x2048-equivalent, info)."""
    from slotvps_b200 import synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synthetic.make_head_state_dict(0)
    cap = synthetic.make_capsule_params(0, args.slots)
    lg, _, _ = synthetic.make_fusion_case(0, args.slots, 8, 8)
    full_px = args.height * args.width
    est_full = 11.0 * full_px / (1024 * 2048) * (8.0 / max(1, os.cpu_count() or 1)) ** 0.5   # survey probe: ~11 s/clip on 8 vCPU
    H, W = args.height, args.width
    while (steps + warmup) * est_full * (H * W / full_px) > budget_s and H > 128:
        H, W = H // 2, W // 2
    times = []
    for i in range(warmup + steps):
        dt = cpu_clip(args, H, W, sd, cap, lg, video=100 + i)
        if i >= warmup:
            times.append(dt)
    t = statistics.median(times)
    scale = (H * W) / full_px
    fps = args.frames / t * scale
    return fps, dict(kind="port", cores=torch.get_num_threads(),
                     sample=f"{steps} clip(s) of {H}x{W} T={args.frames} (head+mask logits+fusion, median, {warmup} warm-up); "
                            f"frames/s scaled by pixel ratio {scale:g} to {args.height}x{args.width}",
                     sec_per_sample_clip=t), t * (steps + warmup)


def bind_to_gpu_numa_node(local):
    """Pin this rank's host threads (and, by first touch, its pinned staging buffers) to the NUMA node its GPU hangs
    off, so that N ranks streaming features over PCIe do not all pull from one socket's memory."""
    try:
        p = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return dict(node=node, cpus=len(cpus))
    except Exception:
        return None


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        budget = 150.0
        fps, info, _ = run_cpu_arm(args, args.steps, args.warmup, budget)
        line = dict(metric=METRIC, value=fps, unit=UNIT, impl="reference", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                    ms_per_step=1e3 * info["sec_per_sample_clip"], higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f32", data="synthetic",
                    config=dict(workload=f"r50_fpn_slotvps retriever, single {args.height}x{args.width} clip, T={args.frames}, N={args.slots} (BASELINE configs[1])",
                                fusion_logits="designed (random-init heads keep no slot)"),
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
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"            # keep stdout to the single JSON line
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    T, N, H, W = args.frames, args.slots, args.height, args.width
    sd = synthetic.make_head_state_dict(0)
    cap = synthetic.make_capsule_params(0, N)
    fusion_logits = synthetic.make_fusion_case(0, N, 8, 8)[0].to(dev)
    K, Wm, M = args.steps, args.warmup, max(1, args.inflight)
    if M >= 4:
        # throughput mode: with several clips in flight the side-stream producers of one clip co-run with the other
        # clips' kernels, and narrower grids interleave better (measured 64 > 112 > 148); the library default (112)
        # is the single-clip latency optimum.
        os.environ.setdefault("SLOTVPS_SIDE_CTAS", "64")
        os.environ.setdefault("SLOTVPS_MAIN_CTAS", "64")
    pan = torch.empty((K, H, W), dtype=torch.int64, device=dev) if world > 1 else None
    L = sv.lib()

    # Clips are independent, so M clips are kept in flight per GPU: lane j owns a model instance
    # (its own workspaces), a resident input clip (178 MB at T=2: larger than the 126 MB L2), a stream and a
    # CUDA graph of the whole step.  Step i runs on lane i % M.
    class Lane:
        pass
    lanes = []
    for j in range(M):
        ln = Lane()
        ln.model = sv.SlotVPSRetriever({**sv.HEAD_KWARGS, "kernel_path": args.kernel_path}, N, sv.FUSION_KWARGS)
        ln.model.dynamic_mask_head.load_state_dict(sd, strict=True)
        ln.model.load_capsule_params(cap)
        ln.model = ln.model.to(dev)
        ln.host = synthetic.make_features(H, W, T=T, video=10 * rank + j, frame=0)
        ln.clip = [[f.to(dev) for f in fr] for fr in ln.host]
        ln.stream = torch.cuda.Stream()
        ln.graph = None if args.no_graph else sv.GraphedClip(ln.model, ln.clip, (H, W), pos="sine", fusion_logits=fusion_logits)
        ln.end = torch.cuda.Event()
        lanes.append(ln)
    model = lanes[0].model
    dev_clips = [ln.clip for ln in lanes]

    def step(i, feats):                      # eager single-stream step (profiling pass)
        return model(feats, (H, W), pos="sine", fusion_logits=fusion_logits)

    def lane_step(ln, i):
        o = ln.graph.replay() if ln.graph is not None else ln.model(ln.clip, (H, W), pos="sine", fusion_logits=fusion_logits)
        if pan is not None:
            pan[i % K].copy_(o["fusion"].panoptic, non_blocking=True)
        return o

    main = torch.cuda.current_stream()
    start = torch.cuda.Event()

    def run(n, body):
        start.record(main)
        for ln in lanes:
            ln.stream.wait_event(start)
        out = None
        for i in range(n):
            ln = lanes[i % M]
            with torch.cuda.stream(ln.stream):
                out = body(ln, i)
        for ln in lanes:
            ln.end.record(ln.stream)
            main.wait_event(ln.end)
        return out

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    run(max(Wm, M), lane_step)
    if dist is not None:
        from slotvps_b200.parallel import WIRE_DTYPE
        pan_wire = torch.empty((K, H, W), dtype=WIRE_DTYPE, device=dev)
        gathered = torch.empty((world * K, H, W), dtype=WIRE_DTYPE, device=dev)
        dist.all_gather_into_tensor(gathered.view(torch.uint8), pan_wire.view(torch.uint8))     # warm-up: NCCL channel setup stays out of the timed region
    barrier()
    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else local)
    if rank == 0:
        sampler.start()
    L.slotvps_launch_count(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    out = run(K, lane_step)
    if dist is not None:                                    # one all-gather of the shard's id maps,
        pan_wire.copy_(pan)                                 # narrowed to int16 on the wire (ids < stuff_num + N)
        dist.all_gather_into_tensor(gathered.view(torch.uint8), pan_wire.view(torch.uint8))   # raw bytes: NCCL has no int16
    e1.record()
    barrier()
    launches = world * (int(L.slotvps_launch_count(0)) if lanes[0].graph is None else K * lanes[0].graph.launches)
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * K * T / (ms * 1e-3)
    meta = out["fusion"].host()

    # single clip in flight (latency view of the same step), for reference
    single_ms = None
    if M > 1:
        keep, lanes[:] = lanes[:], lanes[:1]
        M1, M = M, 1
        run(max(2, Wm), lane_step)
        barrier()
        e0.record()
        run(K, lane_step)
        e1.record()
        barrier()
        single_ms = e0.elapsed_time(e1) / K
        lanes[:] = keep
        M = M1

    # ---- e2e: host buffers in, id map + meta out, every step -----------------------------------------
    e2e = None
    if not args.no_e2e:
        for ln in lanes:
            ln.pinned = [[f.pin_memory() for f in fr] for fr in ln.host]
            ln.h_pan = torch.empty((H, W), dtype=torch.int64).pin_memory()
            ln.h_meta = torch.empty(4 + 3 * N, dtype=torch.int32).pin_memory()
        h2d = sum(f.numel() * 4 for fr in lanes[0].host for f in fr)
        d2h = lanes[0].h_pan.numel() * 8 + lanes[0].h_meta.numel() * 4

        def e2e_step(ln, i):
            # upload this step's clip from pinned host memory into the lane's input buffers, run, read results back;
            # the other lanes' uploads / downloads overlap this lane's compute
            for t in range(T):
                for l in range(4):
                    ln.clip[t][l].copy_(ln.pinned[t][l], non_blocking=True)
            o = lane_step(ln, i)
            ln.h_pan.copy_(o["fusion"].panoptic, non_blocking=True)
            ln.h_meta.copy_(o["fusion"].meta, non_blocking=True)
            return o
        run(max(2, Wm, M), e2e_step)
        barrier()
        e0.record()
        run(K, e2e_step)
        e1.record()
        barrier()
        ms2 = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms2], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms2 = float(t.item())
        e2e = dict(value=world * K * T / (ms2 * 1e-3), unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                   ms_per_step=ms2 / K, pipeline=f"{M} clips in flight: a lane's upload/readback overlaps the other lanes' compute")
    clocks = sampler.stop() if rank == 0 else None     # sampled over the timed region and the e2e region (both under load)

    # ---- per-kernel device times (CUDA events on the launching stream) -> roofline ----------------------
    roofline, breakdown = None, None
    if rank == 0:
        pk = peaks()
        per_frame_flops, shapes, heads = algorithmic_flops(args)
        nprof = min(K, 3)
        import ctypes
        buf = ctypes.create_string_buffer(1 << 16)
        torch.cuda.synchronize()
        L.slotvps_profile_begin(torch.cuda.current_stream(dev).cuda_stream)
        for i in range(nprof):
            step(i, dev_clips[i % len(dev_clips)])
        L.slotvps_profile_end(buf, len(buf))
        rows = [r.split("\t") for r in buf.value.decode().strip().split("\n") if r]
        breakdown = {r[0]: dict(launches_per_step=int(r[1]) / nprof, ms_per_step=float(r[2]) / nprof) for r in rows}
        total = sum(v["ms_per_step"] for v in breakdown.values())
        tj_all = {}
        tp = os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")
        if os.path.exists(tp):              # DRAM bytes per pixel measured by ncu on the level-3 launches
            tj_all = json.load(open(tp))
        px_step = sum(hd * h * w for hd, (h, w) in zip(heads, shapes)) * T            # pixel.stage units of the attention kernels
        px_all = sum(h * w for (h, w) in shapes) * T                                  # pixels of the level fusion

        def kernel_roofline(name):
            """Roofline entry of one kernel: its own bound, algorithmic work of all its launches in a step / their time."""
            t_ms = breakdown[name]["ms_per_step"]
            nl = breakdown[name]["launches_per_step"]
            fl = kernel_flops_per_step(name, args, shapes, heads)
            if fl:                          # tensor-bound kernels of the attention contraction
                ach = fl / (t_ms * 1e-3) / 1e12
                peak = pk["bf16_tflops"]              # burst figure: the profiling pass times every kernel alone, with host gaps between launches
                tj = tj_all.get(name)
                traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) / tj["pixels"] * px_step / nl if tj else None
                return dict(kernel=name, bound="tensor", achieved=ach, peak=peak, unit="TFLOP/s", frac=ach / peak,
                            executed_mma_tflops=3 * ach, executed_frac=3 * ach / peak,      # fp16 hi/lo split: 3 MMA products per algorithmic product
                            traffic=traffic, traffic_note="DRAM bytes per launch (ncu level-3 capture scaled by pixels)",
                            peak_source=f"{pk['source']} bf16 burst (kernel timed alone in the per-kernel profiling pass)",
                            algorithmic_flops_per_step=fl, ms_per_step=t_ms, launches_per_step=nl)
            if name == "fuse_tc":           # level fusion: read 4*128 B + write 4*256 B (fp32 feature) + 4 fp16 planes (2 KB) per pixel
                by = px_all * (4 * 128 + 4 * 256 + 4 * 2 * 256)
                ach = by / (t_ms * 1e-3) / 1e9
                tj = tj_all.get("fuse_tc_main")
                traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) / tj["pixels"] * px_all / nl if tj else None
                return dict(kernel=name, bound="hbm", achieved=ach, peak=pk["hbm_gbs"], unit="GB/s", frac=ach / pk["hbm_gbs"],
                            traffic=traffic, traffic_note="DRAM bytes per launch (ncu capture of the level-3 main pass scaled by pixels; coarse passes excluded)",
                            peak_source=pk["source"] + " copy bandwidth", algorithmic_bytes_per_step=by, ms_per_step=t_ms, launches_per_step=nl)
            return None

        per_kernel = {k: kernel_roofline(k) for k in breakdown if k in ("stats_tc", "attn_tc", "fuse_tc", "proj_rstd(k)", "proj_rstd(v)", "slot_attn_fp32")}
        per_kernel = {k: v for k, v in per_kernel.items() if v}
        if per_kernel:
            # the dominant kernel = the one with the largest share of the step; the others are listed under "kernels"
            name = max(per_kernel, key=lambda k: per_kernel[k]["ms_per_step"])
            roofline = dict(per_kernel[name])
            roofline["kernels"] = per_kernel
            peak = pk["bf16_tflops"]              # burst figure: the profiling pass times every kernel alone, with host gaps between launches
            # the whole attention contraction (all its kernels) against the same peak
            names = [k for k in breakdown if k in ("proj_rstd(k)", "proj_rstd(v)", "slot_attn_fp32", "stats_tc", "attn_tc")]
            tt = sum(breakdown[k]["ms_per_step"] for k in names)
            roofline["attention_contraction"] = dict(kernels=names, algorithmic_tflops=per_frame_flops * T / (tt * 1e-3) / 1e12,
                                                     ms_per_step=tt, frac_of_peak=per_frame_flops * T / (tt * 1e-3) / 1e12 / peak,
                                                     executed_frac_of_peak=3 * per_frame_flops * T / (tt * 1e-3) / 1e12 / peak)
        # HBM-bound stages: mask logits (read 4*C*P3 + write 4*N*P3) -- SURVEY.md 8d
        P3 = shapes[3][0] * shapes[3][1]
        ml = [k for k in breakdown if k in ("mask_prep", "feat_rnorm", "mask_logits_tc")]
        if roofline is not None:
            step_ms = single_ms if single_ms else ms / K
            # share of the serialised kernel time of one step (the definition an ncu launch list gives: profiles/
            # r1_tc_launches_summary.txt) and, separately, of the wall time of a step with one clip in flight (two streams overlap)
            ktot = sum(v["ms_per_step"] for k, v in breakdown.items() if k != "(host gap)")
            roofline["share_of_step"] = roofline["ms_per_step"] / ktot
            roofline["share_of_single_clip_wall_step"] = roofline["ms_per_step"] / step_ms
            for v in roofline["kernels"].values():
                v["share_of_step"] = v["ms_per_step"] / ktot
            hb = {"peak_gbs": pk["hbm_gbs"], "peak_source": pk["source"] + " copy bandwidth"}
            if "mask_tc" in breakdown:      # mask-logit projection: read 4*C*P3 (as 2 fp16 hi/lo planes = same bytes) + write 4*N*P3
                by = 4 * 256 * P3 + 4 * N * P3
                t_ms = breakdown["mask_tc"]["ms_per_step"]
                hb["mask_logits"] = dict(kernel="mask_tc", algorithmic_bytes=by, ms=t_ms, achieved_gbs=by / (t_ms * 1e-3) / 1e9,
                                         frac=by / (t_ms * 1e-3) / 1e9 / pk["hbm_gbs"])
            fk = [k for k in breakdown if k.startswith("fuse_") and k != "fuse_tc"]
            if fk:                          # fusion, exact two-pass form: 2*4*Kept*P3 read + 8*H*W written
                by = 2 * 4 * meta["k"] * P3 + 8 * H * W
                t_ms = sum(breakdown[k]["ms_per_step"] for k in fk)
                hb["panoptic_fusion"] = dict(kernels=fk, kept_slots=meta["k"], algorithmic_bytes=by, ms=t_ms,
                                             achieved_gbs=by / (t_ms * 1e-3) / 1e9, frac=by / (t_ms * 1e-3) / 1e9 / pk["hbm_gbs"])
            if "fuse_tc" in breakdown:      # level fusion: read 4*128*P + write 4*256*P (fp32 feature) + 4 fp16 planes, per frame & level
                px_all = sum(h * w for (h, w) in shapes) * T
                by = px_all * (4 * 128 + 4 * 256 + 4 * 2 * 256)
                t_ms = breakdown["fuse_tc"]["ms_per_step"]
                hb["level_fusion"] = dict(kernel="fuse_tc", algorithmic_bytes=by, ms=t_ms, achieved_gbs=by / (t_ms * 1e-3) / 1e9,
                                          frac=by / (t_ms * 1e-3) / 1e9 / pk["hbm_gbs"])
            roofline["hbm_stages"] = hb

    # ---- CPU baseline beside it (rank 0, N == 1) ------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fps, info, _ = run_cpu_arm(args, 2, 1, args.cpu_budget_s)
        cpu = dict(value=fps, unit=UNIT, cores=info["cores"], kind=info["kind"], sample=info["sample"])

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=Wm, ms_per_step=ms / K,
                    higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                    config=dict(workload=f"r50_fpn_slotvps retriever, single {H}x{W} clip, T={T}, N={N}, 7 stages (BASELINE configs[1])",
                                frames_convention="retriever frames/s = T * clips/s; output frames/s = clips/s",
                                l2="inputs larger than L2 (178 MB/clip); one resident clip per lane, lanes alternate",
                                fusion_logits="designed (random-init heads keep no slot)",
                                kernel_path=args.kernel_path, cuda_graph=not args.no_graph, clips_in_flight=M, numa_bind=numa, side_stream_ctas=int(os.environ.get("SLOTVPS_SIDE_CTAS", "112")), main_stream_ctas=int(os.environ.get("SLOTVPS_MAIN_CTAS", "148")),
                                single_clip_in_flight_ms_per_step=single_ms, sharding="clips per rank, one all_gather of id maps" if world > 1 else "single GPU",
                                kept_slots=meta["k"], fusion_iters=meta["iters"]),
                    clocks=clocks, e2e=e2e, gpu_launches=launches, roofline=roofline, cpu_baseline=cpu,
                    kernel_breakdown_ms_per_step=breakdown)
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
