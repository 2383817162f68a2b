"""Full-size parity of EvalModel('vmn_fba') against the CPU oracle on one 1088x1920 window (BASELINE configs[4]), and
the oracle's wall time on the host cores (the CPU baseline of that config):  python tools/fba_parity_full.py [H W]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import tcvom_b200
from tcvom_b200 import synthetic
from helpers import fixture_sd_fba
from oracle import vmn_fba_oracle as O

H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1088, 1920)
torch.set_num_threads(os.cpu_count() or 1)
sd = fixture_sd_fba()
imgs, tris = synthetic.make_window(H, W, seed=7)
ti, tt = torch.from_numpy(imgs), torch.from_numpy(tris)
variants = {}
with torch.no_grad():
    t0 = time.time()
    ra, rF, rB = O.eval_forward(sd, ti.float(), tt.float())
    cpu_s = time.time() - t0
    for split in (os.environ.get("TCV_FBA_SPLIT_K_VARIANTS", "1")).split(","):
        os.environ["TCV_FBA_SPLIT_K"] = split
        m = tcvom_b200.EvalModel(model="vmn_fba", agg_window=7)
        m.NET.load_state_dict(sd, strict=True)
        m = m.cuda().eval()
        a, F, B = m(ti.cuda(), tt.cuda())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(3):
            m(ti.cuda(), tt.cuda())
        e1.record(); torch.cuda.synchronize()
        variants[split] = dict(alpha_max_abs_err=float((a.cpu() - ra).abs().max()), ms=e0.elapsed_time(e1) / 3,
                               alpha_err_p9999=float((a.cpu() - ra).abs().flatten()[::7].quantile(0.9999)))
        del m
        torch.cuda.empty_cache()
unk = (tt[:, 1] == 128)
res = dict(size=[H, W], split_k_variants=variants, alpha_max_abs_err=float((a.cpu() - ra).abs().max()), F_max_abs_err=float((F.cpu() - rF).abs().max()),
           B_max_abs_err=float((B.cpu() - rB).abs().max()), alpha_mean_abs_err=float((a.cpu() - ra).abs().mean()),
           unknown_fraction=float(unk.float().mean()), alpha_unknown_mean=float(ra[:, 1][unk].mean()),
           alpha_unknown_std=float(ra[:, 1][unk].std()), cpu_oracle_s=cpu_s, cpu_cores=os.cpu_count(),
           cpu_windows_per_s=1.0 / cpu_s)
line = json.dumps(res)
print(line)
od = os.path.join(ROOT, "gpurun_out")
if os.path.isdir(od):
    open(os.path.join(od, "fba_parity_full.json"), "w").write(line + "\n")
