import torch, sys, os
sys.path.insert(0, os.getcwd())
import slotvps_b200 as sv
from oracle import slotvps_oracle as O
from slotvps_b200 import synthetic
dev = torch.device("cuda:0")
def rel(a, b):
    a = a.double().cpu(); b = b.double()
    return float((a - b).norm() / b.norm())
for T, shapes in ((4, [(9, 15), (18, 30), (36, 60), (72, 120)]), (2, [(9, 15), (18, 30), (36, 60), (72, 120)]), (4, [(8, 16), (16, 32), (32, 64), (64, 128)])):
    N = 100
    sd = synthetic.make_head_state_dict(13); cap = synthetic.make_capsule_params(13, N)
    feats = synthetic.make_features(0, 0, T=T, video=13, shapes=shapes)
    q = cap["init_mask_query.weight"]
    pos64 = [[O.sine_position_embedding(*s, dtype=torch.float64) for s in shapes] for _ in range(T)]
    rc, re_, rf = O.head_forward({k: v.double() for k, v in sd.items()}, [[f.double() for f in fr] for fr in feats], [q.double()] * T, pos64)
    for kp in (0, 1):
        head = sv.B200DynamicMaskHead(**{**sv.HEAD_KWARGS, "kernel_path": kp}); head.load_state_dict(sd); head = head.to(dev)
        cl, em, fu = head([[f.to(dev) for f in fr] for fr in feats], [q.to(dev)] * T, None, pos="sine")
        print("T", T, shapes[0], "path", kp, "noovl" if os.environ.get("SLOTVPS_NO_OVERLAP") else "ovl", [" ".join("%.1e" % rel(em[t][s], re_[t][s]) for s in range(7)) for t in range(min(T, 2))])
