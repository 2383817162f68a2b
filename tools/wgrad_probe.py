"""Bring-up probe for the tensor-core weight-gradient GEMM (tcv_transpose_pad + tcv_wgrad_tc) vs torch autograd.
usage: python tools/wgrad_probe.py [case]   (each case in its own process: a device trap must not hide the others)"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

CASES = {
    "1x1_64_64": dict(cin=64, cout=64, k=1, n=2, h=16, w=24),
    "1x1_128_128": dict(cin=128, cout=128, k=1, n=2, h=16, w=24),
    "3x3_64_64": dict(cin=64, cout=64, k=3, n=2, h=16, w=24),
    "3x3_32_32": dict(cin=32, cout=32, k=3, n=2, h=32, w=32),
    "3x3_256_128": dict(cin=256, cout=128, k=3, n=2, h=8, w=12),
    "3x3_8_32": dict(cin=8, cout=32, k=3, n=2, h=32, w=32),
    "3x3_32_8": dict(cin=32, cout=8, k=3, n=2, h=32, w=32),
    # single-tap probes: which K offsets does TMA accept?  (w+2 = 24: row shifts are 16-byte aligned)
    "tap_pos_unaligned": dict(cin=64, cout=64, k=3, n=2, h=16, w=22, taps=[(0, 1)]),
    "tap_neg_unaligned": dict(cin=64, cout=64, k=3, n=2, h=16, w=22, taps=[(0, -1)]),
    "tap_pos_aligned": dict(cin=64, cout=64, k=3, n=2, h=16, w=22, taps=[(1, 0)]),
    "tap_neg_aligned": dict(cin=64, cout=64, k=3, n=2, h=16, w=22, taps=[(-1, 0)]),
}


def run(name):
    import torch
    import torch.nn.functional as F
    from tcvom_b200 import _cabi
    from train_check import to_act, from_act, rnd, rel
    c = CASES[name]
    L = _cabi.lib()
    st = torch.cuda.current_stream().cuda_stream
    x = rnd((c["n"], c["cin"], c["h"], c["w"]), 1)
    dz = rnd((c["n"], c["cout"], c["h"], c["w"]), 2)
    xa, za = to_act(x), to_act(dz)
    row = (c["w"] + 2 + 7) // 8 * 8
    ktot = c["n"] * (c["h"] + 2) * row
    xt = torch.empty((2, c["cin"], ktot), dtype=torch.bfloat16, device="cuda")
    zt = torch.empty((2, c["cout"], ktot), dtype=torch.bfloat16, device="cuda")
    _cabi.check(L.tcv_transpose_pad(xa.ptr, xa.plane, xa.n, xa.h, xa.w, xa.c, 1, 0, 0, row, 0, xt.data_ptr(), c["cin"] * ktot, ktot, st), "tp")
    _cabi.check(L.tcv_transpose_pad(za.ptr, za.plane, za.n, za.h, za.w, za.c, 1, 0, 0, row, 0, zt.data_ptr(), c["cout"] * ktot, ktot, st), "tp")
    torch.cuda.synchronize()
    # check the transpose
    xr = from_act(xa)
    xt32 = (xt[0].float() + xt[1].float()).reshape(c["cin"], c["n"], c["h"] + 2, row)
    print(name, "transpose err", rel(xt32[:, :, 1:-1, 1:c["w"] + 1].permute(1, 0, 2, 3), xr), "ring", float(xt32[:, :, 0].abs().max()))
    k = c["k"]
    all_taps = [(ky - k // 2, kx - k // 2) for ky in range(k) for kx in range(k)]
    taps = c.get("taps", all_taps)
    nt = len(taps)
    Arr = C.c_int * nt
    dy, dx, wt = Arr(*[t[0] for t in taps]), Arr(*[t[1] for t in taps]), Arr(*range(nt))
    nsplit = 3
    partial = torch.empty((nsplit, c["cin"], 3 * c["cout"]), device="cuda")
    dw = torch.zeros((nt, c["cin"], c["cout"]), device="cuda")
    for sx in sorted({t[1] for t in taps}):
        _cabi.check(L.tcv_transpose_pad(za.ptr, za.plane, za.n, za.h, za.w, za.c, 1, 0, 0, row, sx, zt.data_ptr(), c["cout"] * ktot, ktot, st), "tp")
        ts = [i for i, t in enumerate(taps) if t[1] == sx]
        A2 = C.c_int * len(ts)
        _cabi.check(L.tcv_wgrad_tc(xt.data_ptr(), c["cin"] * ktot, zt.data_ptr(), c["cout"] * ktot, c["cin"], c["cout"], ktot,
                                   row, len(ts), A2(*[taps[i][0] for i in ts]), A2(*([0] * len(ts))), A2(*ts),
                                   partial.data_ptr(), nsplit, dw.data_ptr(), c["cout"], st), "wgrad_tc")
    torch.cuda.synchronize()
    # NHWC-direct kernel on the same problem
    from tcvom_b200._cabi import ConvDesc
    d = ConvDesc()
    d.x, d.x_plane, d.x_img_stride = xa.ptr, xa.plane, xa.img_elems
    d.n, d.ih, d.iw, d.cin = xa.n, xa.h, xa.w, xa.c
    d.ntaps = nt
    for i, tp in enumerate(taps):
        d.dy[i], d.dx[i], d.wtap[i] = tp[0], tp[1], i
    d.stride, d.pad_mode = 1, 0
    d.oh, d.ow, d.cout, d.gh, d.gw = xa.h, xa.w, c["cout"], xa.h, xa.w
    d.oy_mul, d.oy_off, d.ox_mul, d.ox_off = 1, 0, 1, 0
    dw2 = torch.zeros((nt, c["cin"], c["cout"]), device="cuda")
    _cabi.check(L.tcv_conv2d_wgrad_nhwc_tc(C.byref(d), za.ptr, za.plane, c["cout"], dw2.data_ptr(), st), "wgrad_nhwc")
    torch.cuda.synchronize()
    w = torch.zeros((c["cout"], c["cin"], k, k), device="cuda", requires_grad=True)
    y = F.conv2d(xr, w, None, 1, k // 2)
    (gw,) = torch.autograd.grad(y, w, from_act(za))
    ref = gw.permute(2, 3, 1, 0).reshape(len(all_taps), c["cin"], c["cout"])[[all_taps.index(t) for t in taps]]
    print(name, "wgrad err", rel(dw, ref), "nhwc-direct err", rel(dw2, ref))


if __name__ == "__main__":
    if len(sys.argv) == 2:
        run(sys.argv[1])
    else:
        for n in (sys.argv[2:] if len(sys.argv) > 2 else CASES):
            r = subprocess.run([sys.executable, __file__, n], capture_output=True, text=True, timeout=120)
            out = (r.stdout + r.stderr).strip().splitlines()
            print("\n".join(l for l in out if n in l or "rror" in l or "timed" in l)[:1500], flush=True)
