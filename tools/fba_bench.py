"""Device timing of one EvalModel('vmn_fba') forward with a per-kernel-kind breakdown (BASELINE configs[4]; not the
headline bench):  python tools/fba_bench.py [H W [reps]]   -> one JSON line (also written to gpurun_out/ when present)."""
import collections, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import tcvom_b200
from tcvom_b200 import synthetic
from helpers import fixture_sd_fba

H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1088, 1920)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
GFLOP_256 = 223.1                               # SURVEY.md section 8d config 5 (reference FlopCounterMode, conv only)
from tcvom_b200 import _cabi
_cabi.lib().tcv_set_debug_flags(int(os.environ.get("TCV_DEBUG_FLAGS", "0")))   # measurement switches
m = tcvom_b200.EvalModel(model="vmn_fba", agg_window=7)
m.NET.load_state_dict(fixture_sd_fba(), strict=True)
m = m.cuda().eval()
imgs, tris = synthetic.make_window(H, W, seed=7)
ti, tt = torch.from_numpy(imgs).cuda(), torch.from_numpy(tris).cuda()
with torch.no_grad():
    t0 = time.time(); out = m(ti, tt); torch.cuda.synchronize(); rec_s = time.time() - t0
    for _ in range(2):
        out = m(ti, tt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        out = m(ti, tt)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    plan = list(m.NET.engine().plans.values())[0]
    per = plan.replay_timed(ti.device)
    per = plan.replay_timed(ti.device)
kinds = collections.OrderedDict()
for meta, t in zip(plan.meta, per):
    k = kinds.setdefault(meta["kind"], dict(ms=0.0, n=0, gflop=0.0, gb=0.0))
    k["ms"] += t; k["n"] += 1; k["gflop"] += meta.get("flops", 0) / 1e9; k["gb"] += meta.get("bytes", 0) / 1e9
gflop = GFLOP_256 * H * W / 65536.0
res = dict(workload=f"FBA+TAM forward {H}x{W} 3-frame window, batch 1 (configs[4])", ms_per_window=ms,
           windows_per_s=1000.0 / ms, algorithmic_tflops=gflop / ms, launches=plan.n_launch, record_s=rec_s,
           peak_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30, finite=bool(torch.isfinite(out[0]).all()),
           alpha_mean=float(out[0][:, 1].mean()),
           breakdown={k: dict(ms=round(v["ms"], 3), launches=v["n"], tflops=round(v["gflop"] / max(v["ms"], 1e-9), 1),
                              gbs=round(v["gb"] / max(v["ms"], 1e-9) * 1e3, 1))
                      for k, v in sorted(kinds.items(), key=lambda kv: -kv[1]["ms"])})
line = json.dumps(res)
print(line)
od = os.path.join(ROOT, "gpurun_out")
if os.path.isdir(od):
    open(os.path.join(od, "fba_bench.json"), "w").write(line + "\n")
