"""One eager (CUDA graphs off) EvalModel('vmn_fba') forward on a 1088x1920 3-frame window, for an ncu launch list:
  TCV_GRAPHS=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      --csv --log-file gpurun_out/fba_launches.csv python tools/fba_window_once.py
  python tools/launch_summary.py gpurun_out/fba_launches.csv "title" trimask_raw_kernel > profiles/r01_fba_launches_summary.md"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("TCV_GRAPHS", "0")
import torch
import tcvom_b200
from tcvom_b200 import synthetic
from helpers import fixture_sd_fba

H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1088, 1920)
m = tcvom_b200.EvalModel(model="vmn_fba", agg_window=7)
m.NET.load_state_dict(fixture_sd_fba(), strict=True)
m = m.cuda().eval()
imgs, tris = synthetic.make_window(H, W, seed=7)
with torch.no_grad():
    out = m(torch.from_numpy(imgs).cuda(), torch.from_numpy(tris).cuda())     # records = runs every kernel once
torch.cuda.synchronize()
print("alpha mean", float(out[0][:, 1].mean()))
