"""Quick device timing of one EvalModel forward (not the bench): python tools/time_window.py H W [reps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import tcvom_b200
from tcvom_b200 import synthetic, _cabi
from helpers import fixture_sd

H, W = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
graphs = os.environ.get("TCV_GRAPHS", "1") == "1"
_cabi.lib().tcv_set_debug_flags(int(os.environ.get("TCV_DEBUG_FLAGS", "0")))
m = tcvom_b200.EvalModel(model="vmn_gca", agg_window=7)
m.NET.load_state_dict(fixture_sd(), strict=True)
m = m.cuda().eval()
imgs, tris = synthetic.make_window(H, W, seed=7)
ti, tt = torch.from_numpy(imgs).float().cuda(), torch.from_numpy(tris).float().cuda()
with torch.no_grad():
    eng = None
    t0 = time.time(); out = m(ti, tt); torch.cuda.synchronize(); print("first call (record) s", time.time() - t0)
    m.NET.engine().use_graphs = graphs
    out = m(ti, tt); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        out = m(ti, tt)
    e1.record(); torch.cuda.synchronize()
    print(f"{H}x{W}: {e0.elapsed_time(e1)/reps:.3f} ms/window  graphs={graphs}  launches/plan={list(m.NET.engine().plans.values())[0].n_launch}")
    print("alpha mean", float(out[:,1].mean()), "finite", bool(torch.isfinite(out).all()))
    print("mem GB", torch.cuda.max_memory_allocated()/2**30)
