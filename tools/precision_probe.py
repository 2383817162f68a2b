"""CPU experiment behind DESIGN.md's precision section: how many of the split-bf16 product terms do the convolutions
need for the 1e-3 alpha contract?  Every conv of the vmn_gca oracle is evaluated as
  S3 = Ahi.Bhi + Ahi.Blo + Alo.Bhi (shipped),  S2a = Ahi.Bhi + Ahi.Blo,  S2b = Ahi.Bhi + Alo.Bhi,  S1 = Ahi.Bhi
(fp32 accumulation) and the alpha matte is compared with the plain fp32 oracle:   python tools/precision_probe.py 256 256
Measured: S3 1.1e-4 / 1.9e-4 (128^2 / 256^2), S2a 3.3e-2 / 6.9e-2, S2b 3.2e-2 / 6.6e-2, S1 6.7e-2 / 9.9e-2."""
import sys, numpy as np, torch, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import fixture_sd
from oracle import vmn_gca_oracle as O
from tcvom_b200 import synthetic
import torch.nn.functional as F
torch.set_num_threads(8)
sd = fixture_sd()
H, W = int(sys.argv[1]), int(sys.argv[2])
imgs, tris = synthetic.make_window(H, W, seed=7)
ti, tt = torch.from_numpy(imgs).float(), torch.from_numpy(tris).float()
ref = O.eval_forward(sd, ti, tt)
orig_conv, orig_convT = F.conv2d, F.conv_transpose2d
def split(t):
    h = t.bfloat16().float(); l = (t - h).bfloat16().float(); return h, l
MODE = {"m": "S3"}
def conv_emul(fn):
    def f(x, w, b=None, *a, **k):
        if MODE["m"] == "fp32" or w.shape[-1] * w.shape[-2] * w.shape[1] < 16:
            return fn(x, w, b, *a, **k)
        xh, xl = split(x); wh, wl = split(w)
        m = MODE["m"]
        y = fn(xh, wh, None, *a, **k)
        if m in ("S3", "S2a"): y = y + fn(xh, wl, None, *a, **k)
        if m in ("S3", "S2b"): y = y + fn(xl, wh, None, *a, **k)
        if b is not None: y = y + b.reshape(1, -1, 1, 1)
        return y
    return f
F.conv2d = conv_emul(orig_conv); F.conv_transpose2d = conv_emul(orig_convT)
unk = (tt[:, 1] == 128)
for m in ("fp32", "S3", "S2a", "S2b", "S1"):
    MODE["m"] = m
    out = O.eval_forward(sd, ti, tt)
    d = (out - ref).abs()
    print(f"{H}x{W} {m:5s} alpha max abs err {float(d.max()):.2e}  mean over unknown {float(d[:,1][unk].mean()):.2e}")
