import torch, sys, os
sys.path.insert(0, os.getcwd())
import slotvps_b200 as sv
from oracle import slotvps_oracle as O
from slotvps_b200 import synthetic
dev = torch.device("cuda:0")
def rel(a, b):
    a = a.double().cpu(); b = b.double()
    return float((a - b).norm() / b.norm())
# predecessor: a 512x1024 clip (as test_whole_clip)
sd0 = synthetic.make_head_state_dict(0); cap0 = synthetic.make_capsule_params(0, 100)
feats0 = [[f.to(dev) for f in fr] for fr in synthetic.make_features(512, 1024, T=2, video=1, frame=2)]
for kp in (1, 0):
    m = sv.SlotVPSRetriever({**sv.HEAD_KWARGS, "kernel_path": kp}, 100, sv.FUSION_KWARGS)
    m.dynamic_mask_head.load_state_dict(sd0); m.load_capsule_params(cap0); m = m.to(dev)
    m(feats0, (512, 1024), fuse=False)
del m
torch.cuda.synchronize()
T, shapes, N = 4, [(9, 15), (18, 30), (36, 60), (72, 120)], 100
sd = synthetic.make_head_state_dict(13); cap = synthetic.make_capsule_params(13, N)
feats = synthetic.make_features(0, 0, T=T, video=13, shapes=shapes)
q = cap["init_mask_query.weight"]
pos64 = [[O.sine_position_embedding(*s, dtype=torch.float64) for s in shapes] for _ in range(T)]
rc, re_, rf = O.head_forward({k: v.double() for k, v in sd.items()}, [[f.double() for f in fr] for fr in feats], [q.double()] * T, pos64)
for kp in (0, 1, 0):
    head = sv.B200DynamicMaskHead(**{**sv.HEAD_KWARGS, "kernel_path": kp}); head.load_state_dict(sd); head = head.to(dev)
    for rep in range(2):
        cl, em, fu = head([[f.to(dev) for f in fr] for fr in feats], [q.to(dev)] * T, None, pos="sine")
        print("path", kp, "rep", rep, [" ".join("%.1e" % rel(em[t][s], re_[t][s]) for s in range(7)) for t in range(T)], "fused", ["%.1e" % rel(fu[0][l], rf[0][l]) for l in range(4)])
