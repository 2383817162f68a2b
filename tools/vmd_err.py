import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, tcvom_b200
from helpers import fixture_sd, golden
g = golden("train_s5.npz")
m = tcvom_b200.FullModel_VMD(model="vmn_gca", agg_window=7, dilate_kernel=3)
m.NET.load_state_dict(fixture_sd(), strict=True)
m = m.cuda().eval()
a, fg, bg = (torch.from_numpy(g[k]).float().cuda() for k in ("a", "fg", "bg"))
with torch.no_grad():
    out = m(a, fg, bg)
print(os.environ.get("TCV_TC_CONV"), os.environ.get("TCV_TC_ATTN"), os.environ.get("TCV_PV_MODE"),
      "alpha err", float(np.abs(out[7].cpu().numpy() - g["alphas"]).max()), "losses", [round(float(o), 5) for o in out[:5]])
