"""Alpha-matte parity (max abs err vs reference golden vectors / CPU oracle) for each kernel-path
combination.  python tests/parity_matrix.py [--big]   (development aid, not collected by pytest)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = ["ring64", "ring96x128", "allunk64", "nounk64", "dil64x96", "batch2_64"]


def child(big):
    import numpy as np, torch
    import tcvom_b200
    from helpers import fixture_sd, golden
    from tcvom_b200 import synthetic
    errs = []
    for case in CASES:
        g = golden(f"eval_{case}.npz")
        dil = int(g["dilate"])
        m = tcvom_b200.EvalModel(model="vmn_gca", agg_window=7, dilate_kernel=None if dil < 0 else dil)
        m.NET.load_state_dict(fixture_sd(), strict=True)
        m = m.cuda().eval()
        with torch.no_grad():
            a = m(torch.from_numpy(g["imgs"]).float().cuda(), torch.from_numpy(g["tris"]).float().cuda())
        errs.append(float(np.abs(a.cpu().numpy() - g["alphas"]).max()))
    line = " ".join(f"{c}={e:.2e}" for c, e in zip(CASES, errs))
    if big:
        from oracle import vmn_gca_oracle as O
        for hw in ((256, 256), (384, 512)):
            imgs, tris = synthetic.make_window(*hw, seed=7)
            ti, tt = torch.from_numpy(imgs).float(), torch.from_numpy(tris).float()
            ref = O.eval_forward(fixture_sd(), ti, tt)
            m = tcvom_b200.EvalModel(model="vmn_gca", agg_window=7)
            m.NET.load_state_dict(fixture_sd(), strict=True)
            m = m.cuda().eval()
            with torch.no_grad():
                out = m(ti.cuda(), tt.cuda())
            line += f" {hw[0]}x{hw[1]}={float((out.cpu() - ref).abs().max()):.2e}"
    print(line)


if __name__ == "__main__":
    if os.environ.get("PM_CHILD") == "1":
        child("--big" in sys.argv)
        sys.exit(0)
    combos = [("0", "0", "fp16"), ("1", "0", "fp16"), ("0", "1", "bf16"), ("0", "1", "fp16"), ("0", "1", "bf16x3"),
              ("1", "1", "fp16"), ("1", "1", "bf16x3")]
    for conv, attn, pv in combos:
        env = dict(os.environ, PM_CHILD="1", TCV_TC_CONV=conv, TCV_TC_ATTN=attn, TCV_PV_MODE=pv)
        r = subprocess.run([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=env, capture_output=True,
                           text=True, timeout=600)
        out = (r.stdout + r.stderr).strip().splitlines()
        print(f"conv_tc={conv} attn_tc={attn} pv={pv if attn == '1' else '-'} rc={r.returncode} :: {out[-1] if out else ''}", flush=True)
