"""Run-to-run reproducibility of the native training step: two fresh models, same inputs.  Prints where the two
runs first differ (forward outputs, per-layer gradients in backward order)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import train_check as tc
from helpers import golden

g = golden("train_step_s5.npz")
a, fg, bg = (torch.from_numpy(g[k]).float().cuda() for k in ("a", "fg", "bg"))
runs = []
for r in range(2):
    m = tc.make_net()
    out = m(a, fg, bg)
    loss = sum(w * o.mean() for w, o in zip(tc.LOSS_WEIGHTS, out[:5]))
    loss.backward()
    torch.cuda.synchronize()
    runs.append(dict(losses=[float(o) for o in out[:5]], alphas=out[7].detach().clone(),
                     grads={n: p.grad.detach().clone() for n, p in m.NET.named_parameters() if p.grad is not None},
                     state={k: v.detach().clone() for k, v in m.NET.state_dict().items()}))
A, B = runs
print("losses", A["losses"], B["losses"])
print("alpha max abs diff", float((A["alphas"] - B["alphas"]).abs().max()))
sd = max(float((A["state"][k].float() - B["state"][k].float()).abs().max()) for k in A["state"])
print("state max abs diff", sd)
rows = []
for n in A["grads"]:
    ga, gb = A["grads"][n], B["grads"][n]
    rows.append((float((ga - gb).double().norm()) / max(float(gb.double().norm()), 1e-20), n))
for e, n in rows:
    if n.startswith(("decoder.conv2", "decoder.bn1", "decoder.conv1", "decoder.layer4.1", "decoder.fam", "decoder.layer3.0.conv1",
                     "decoder.gca", "decoder.layer1.0.conv1", "encoder.layer_bottleneck.1.conv2", "encoder.conv1", "encoder.shortcut.0")):
        print(f"  {e:.3e} {n}")
rows.sort(reverse=True)
print("worst", rows[:3], "median", rows[len(rows) // 2])
print("---- state tensors that differ between the two runs (state_dict order)")
cnt = 0
for k in A["state"]:
    d = float((A["state"][k].float() - B["state"][k].float()).abs().max())
    sc = max(float(A["state"][k].float().abs().max()), 1e-12)
    if d / sc > 1e-7 and cnt < 40:
        print(f"  {d / sc:.2e} {k}")
        cnt += 1
