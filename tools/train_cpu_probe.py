"""Is the native training step CPU-launch-bound?  Wall time the host needs to ENQUEUE one step (no synchronisation) next to
the device time of the step:  python tools/train_cpu_probe.py [B S H W]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import tcvom_b200
from tcvom_b200 import synthetic, _cabi
from helpers import fixture_sd
B, S, H, W = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (4, 5, 512, 512)
dev = torch.device("cuda:0")
model = tcvom_b200.FullModel_VMD(model="vmn_gca", agg_window=7, dilate_kernel=None)
model.NET.load_state_dict(fixture_sd(), strict=True)
model = model.to(dev).train()
a, fg, bg = (torch.from_numpy(t).float().to(dev) for t in synthetic.make_train_batch(B, S, H, W, seed=21))
opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-5, weight_decay=1e-4)
W5 = (1.0, 1.0, 1.0, 0.5, 0.25)
def step():
    out = model(a, fg, bg)
    loss = sum(w * o.mean() for w, o in zip(W5, out[:5]))
    model.zero_grad()
    loss.backward()
    opt.step()
for _ in range(3):
    step()
torch.cuda.synchronize()
enq, tot = [], []
for _ in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    enq.append((t1 - t0) * 1e3); tot.append((t2 - t0) * 1e3)
print(f"B={B} S={S} {H}x{W}: host enqueue {min(enq):.1f} ms, step (enqueue + drain) {min(tot):.1f} ms -> "
      f"{'CPU-launch-bound' if min(enq) > 0.9 * min(tot) else 'GPU-bound'}")
import cProfile, pstats, io
pr = cProfile.Profile(); pr.enable(); step(); pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
