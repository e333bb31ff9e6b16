import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import slotvps_b200 as sv
from slotvps_b200 import synthetic
from oracle import slotvps_oracle as O
dev = torch.device("cuda:0")
seg, pan, ci, oi = synthetic.make_unify_case(0, 128, 256)
u = sv.PanUnifier(dev)
out = u.frame(seg.to(dev), pan.to(dev), ci, oi, 4096)
torch.cuda.synchronize()
print("status", u.status.tolist())
ref = O.unify_pan_result([seg.numpy()], [pan.numpy()], [ci.numpy()], [oi.numpy()], 4096)[0]
got = out.cpu().numpy()
print("mismatch", (got != ref).sum(), [np.unique(got[:, :, c]).tolist() for c in range(3)])
print([np.unique(ref[:, :, c]).tolist() for c in range(3)])
hist = u.ws[256:256 + 256 * 32 * 4].view(torch.int32).reshape(256, 32).cpu()
print("totals", {i: int(hist[i].sum()) for i in range(256) if hist[i].sum() > 0})
