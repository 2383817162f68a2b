"""Companion of precision_probe.py: drop the Alo.Bhi term only in the convs with at least N input channels (the
tensor-bound layers).  Measured at 128x128: none 1.1e-4, cin>=512 2.0e-3, >=256 5.6e-3, >=128 1.4e-2, >=64 2.8e-2, all 3.3e-2
-> no subset of layers can run with two terms under the 1e-3 contract.   python tools/precision_probe_layers.py 128 128"""
import sys, os, numpy as np, torch
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
CFG = {"pred": None}
def conv_emul(fn, transposed):
    def f(x, w, b=None, *a, **k):
        cin = w.shape[0] if transposed else w.shape[1]
        cout = w.shape[1] if transposed else w.shape[0]
        xh, xl = split(x); wh, wl = split(w)
        y = fn(xh, wh, None, *a, **k) + fn(xh, wl, None, *a, **k)
        if not CFG["pred"](cin, cout, x.shape[-1]):
            y = y + fn(xl, wh, None, *a, **k)
        if b is not None: y = y + b.reshape(1, -1, 1, 1)
        return y
    return f
F.conv2d = conv_emul(orig_conv, False); F.conv_transpose2d = conv_emul(orig_convT, True)
unk = (tt[:, 1] == 128)
for name, pred in (("none", lambda ci, co, w: False), ("cin>=512", lambda ci, co, w: ci >= 512), ("cin>=256", lambda ci, co, w: ci >= 256),
                   ("cin>=128", lambda ci, co, w: ci >= 128), ("cin>=64", lambda ci, co, w: ci >= 64), ("all", lambda ci, co, w: True)):
    CFG["pred"] = pred
    out = O.eval_forward(sd, ti, tt)
    d = (out - ref).abs()
    print(f"{H}x{W} drop Alo.Bhi where {name:9s}: alpha max abs err {float(d.max()):.2e}  mean over unknown {float(d[:,1][unk].mean()):.2e}")
