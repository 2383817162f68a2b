"""Small end-to-end forward for compute-sanitizer (memcheck / racecheck): 64x96 window + one FullModel_VMD pass."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, tcvom_b200
from tcvom_b200 import synthetic
from helpers import fixture_sd, golden
m = tcvom_b200.EvalModel(model="vmn_gca", agg_window=7, dilate_kernel=2)
m.NET.load_state_dict(fixture_sd(), strict=True)
m = m.cuda().eval()
m.NET.engine().use_graphs = False
imgs, tris = synthetic.make_window(64, 96, seed=5)
with torch.no_grad():
    a = m(torch.from_numpy(imgs).cuda(), torch.from_numpy(tris).cuda())
torch.cuda.synchronize()
g = golden("train_s5.npz")
f = tcvom_b200.FullModel_VMD(model="vmn_gca", agg_window=7, dilate_kernel=3)
f.NET.load_state_dict(fixture_sd(), strict=True)
f = f.cuda().eval()
f.NET.engine().use_graphs = False
with torch.no_grad():
    out = f(*(torch.from_numpy(g[k]).float().cuda() for k in ("a", "fg", "bg")))
torch.cuda.synchronize()
print("ok", float(a.mean()), [round(float(o), 4) for o in out[:5]])
