"""the LayerNorm-statistics GEMM (stats_tri_kernel) with 1, 2 and 3 fp16 products per algorithmic product.
Measures, on one 1024x2048 T=2 clip against the fp64 oracle: every stage teacher-forced, the free-running drift next to the fp32
oracle's own drift, and the kernel time.  SLOTVPS_STATS_PRODUCTS is a measurement switch; the library ships 3.
Run on the GPU box:  python scripts/stats_products.py > gpurun_out/stats_products.txt"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import slotvps_b200 as sv  # noqa: E402
from slotvps_b200 import synthetic  # noqa: E402
from oracle import slotvps_oracle as O  # noqa: E402  (checker)


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())


def main():
    H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1024, 2048)
    dev = torch.device("cuda:0")
    T, N, seed = 2, 100, 31
    shapes = synthetic.level_shapes(H, W)
    sd = synthetic.make_head_state_dict(seed)
    cap = synthetic.make_capsule_params(seed, N)
    feats = synthetic.make_features(0, 0, T=T, video=seed, frame=0, shapes=shapes)
    q = cap["init_mask_query.weight"]
    pos32 = [[O.sine_position_embedding(*s) for s in shapes] for _ in range(T)]
    torch.set_num_threads(os.cpu_count() or 1)
    c64, e64, _ = O.head_forward({k: v.double() for k, v in sd.items()}, [[f.double() for f in fr] for fr in feats], [q.double()] * T,
                                 [[p.double() for p in pp] for pp in pos32])
    _, e32, _ = O.head_forward(sd, feats, [q] * T, pos32)
    noise = [max(rel(e32[t][s], e64[t][s]) for t in range(T)) for s in range(7)]
    forced = [[q] * T] + [[e64[t][s, 0].float() for t in range(T)] for s in range(6)]
    f_dev = [[f.to(dev) for f in fr] for fr in feats]
    head = sv.B200DynamicMaskHead(**{**sv.HEAD_KWARGS, "kernel_path": 0})
    head.load_state_dict(sd, strict=True)
    head = head.to(dev).eval()
    L = sv.lib()
    print(f"{H}x{W} T={T} N={N}; fp32-oracle-vs-fp64 drift per stage (reference noise): {' '.join(f'{v:.1e}' for v in noise)}")
    for prod in (3, 2, 1):
        os.environ["SLOTVPS_STATS_PRODUCTS"] = str(prod)
        cl, em, _ = head(f_dev, [q.to(dev)] * T, None, pos="sine")
        drift = [max(rel(em[t][s], e64[t][s]) for t in range(T)) for s in range(7)]
        cl_tf, em_tf, _ = head(f_dev, [q.to(dev)] * T, None, pos="sine", stage_slots_in=forced)
        tf = [max(max(rel(em_tf[t][s], e64[t][s]), rel(cl_tf[t][s], c64[t][s])) for t in range(T)) for s in range(7)]
        buf = C.create_string_buffer(1 << 16)
        torch.cuda.synchronize()
        L.slotvps_profile_begin(torch.cuda.current_stream(dev).cuda_stream)
        for _ in range(3):
            head(f_dev, [q.to(dev)] * T, None, pos="sine")
        L.slotvps_profile_end(buf, len(buf))
        ms = {r.split("\t")[0]: float(r.split("\t")[2]) / 3 for r in buf.value.decode().strip().split("\n") if r}
        ratio = max(d / max(n, 2e-6) for d, n in zip(drift, noise))
        print(f"products={prod}: stats_tc {ms.get('stats_tc', float('nan')):.3f} ms/clip | teacher-forced per stage: {' '.join(f'{v:.1e}' for v in tf)} | "
              f"free-running drift: {' '.join(f'{v:.1e}' for v in drift)} | drift/noise max {ratio:.1f}x "
              f"({'passes' if ratio <= 5 and max(tf) < 1e-3 else 'FAILS'} the <=5x drift gate; teacher-forced <=1e-3: {max(tf) < 1e-3})")
    os.environ.pop("SLOTVPS_STATS_PRODUCTS")


if __name__ == "__main__":
    main()
