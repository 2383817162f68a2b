"""Times one conv layer shape under the kernel's measurement switches: which pipeline stage bounds it?"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tcvom_b200 import _cabi
from tcvom_b200._cabi import ConvDesc
sys.path.insert(0, os.path.join(ROOT, "tools"))
from tc_check import split

L = _cabi.lib()
dev = "cuda"
st = torch.cuda.current_stream().cuda_stream


def run(cin, cout, h, w, n, res=True, sweep=None):
    torch.manual_seed(1)
    x = split(torch.randn(n, h, w, cin, device=dev))
    taps = [(ky - 1, kx - 1) for ky in range(3) for kx in range(3)]
    wf = (torch.randn(9, cin, cout, device=dev) / (cin * 9) ** 0.5).contiguous()
    wtc = torch.empty((2, 9, cout, cin), dtype=torch.bfloat16, device=dev)
    _cabi.check(L.tcv_pack_weight_tc(wf.data_ptr(), 9, cin, cout, wtc.data_ptr(), st), "pack")
    s1 = torch.rand(cout, device=dev) + 0.5; b1 = torch.randn(cout, device=dev)
    res1 = split(torch.randn(n, h, w, cout, device=dev))
    y = torch.zeros((2, n, h, w, cout), dtype=torch.bfloat16, device=dev)
    d = ConvDesc()
    d.x = x.data_ptr(); d.n, d.ih, d.iw, d.cin = n, h, w, cin
    d.w = wf.data_ptr(); d.ntaps = 9; d.w_tc = wtc.data_ptr(); d.w_tc_taps = 9
    for i, (dy, dx) in enumerate(taps):
        d.dy[i], d.dx[i], d.wtap[i] = dy, dx, i
    d.stride, d.pad_mode = 1, 0
    d.y = y.data_ptr()
    d.oh, d.ow, d.cout, d.gh, d.gw = h, w, cout, h, w
    d.oy_mul, d.oy_off, d.ox_mul, d.ox_off = 1, 0, 1, 0
    d.s1, d.b1 = s1.data_ptr(), b1.data_ptr()
    if res:
        d.res1 = res1.data_ptr()
    d.act = 1
    flops = 2 * n * h * w * 9 * cin * cout
    for ver in ((int(os.environ.get("TCV_PROBE_VER", "2")),) if sweep else (3, 2)):
        L.tcv_set_conv_tc_version(ver)
        for flags in (sweep if sweep else ((0, 64, 1) if ver == 3 else (0,))):
            L.tcv_set_debug_flags(flags)
            for _ in range(3):
                _cabi.check(L.tcv_conv2d(C.byref(d), st), "conv")
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(10):
                _cabi.check(L.tcv_conv2d(C.byref(d), st), "conv")
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"{cin}->{cout} {h}x{w} n{n} v{ver} flags={flags:6d}: {ms*1e3:8.1f} us  {flops/ms/1e9:7.1f} TF/s(alg)", flush=True)
    L.tcv_set_debug_flags(0); L.tcv_set_conv_tc_version(2)


if len(sys.argv) >= 7:      # python tools/conv_probe.py cin cout h w n flagA,flagB,... : A/B of tcv_set_debug_flags values
    cin, cout, h, w, n = map(int, sys.argv[1:6])
    for rep in range(2):
        run(cin, cout, h, w, n, sweep=tuple(int(v) for v in sys.argv[6].split(",")))
elif len(sys.argv) >= 6:      # python tools/conv_probe.py cin cout h w n : conv_tc2 under its measurement switches
    cin, cout, h, w, n = map(int, sys.argv[1:6])
    # 1 no MMA, 2 no epilogue memory ops, 4 activations loaded once, 8 weights loaded once
    run(cin, cout, h, w, n, sweep=(0, 1, 2, 3, 4, 8, 12, 15))
else:
    run(32, 32, 1088, 1920, 3, res=False)
    run(32, 32, 544, 960, 3)
