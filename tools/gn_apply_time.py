"""tcv_gn_apply: 1 / 2 / 4 work items per thread (debug flags 1 << 23, 1 << 24): identical bits, achieved bandwidth."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tcvom_b200 import _cabi

L = _cabi.lib()
st = torch.cuda.current_stream().cuda_stream
for (n, pixels, c, with_res) in ((3, 272 * 480, 256, False), (3, 272 * 480, 256, True), (3, 136 * 240, 512, True),
                                 (3, 136 * 240, 2048, False), (3, 544 * 960, 64, False), (1, 1000, 64, True)):
    torch.manual_seed(c)
    x = torch.randn(2, n, pixels, c, device="cuda").to(torch.bfloat16)
    res = torch.randn(2, n, pixels, c, device="cuda").to(torch.bfloat16) if with_res else None
    sc, sh = torch.rand(n, c, device="cuda") + 0.5, torch.randn(n, c, device="cuda")
    row = dict(shape=[n, pixels, c], res=with_res)
    ys = []
    for name, flags in (("u1", 0), ("u2", 1 << 23), ("u4", 1 << 24)):
        L.tcv_set_debug_flags(flags)
        y = torch.empty_like(x)
        call = lambda: _cabi.check(L.tcv_gn_apply(x.data_ptr(), 0, n, pixels, c, sc.data_ptr(), sh.data_ptr(),
                                                  res.data_ptr() if with_res else None, 0, 1, y.data_ptr(), 0, c, 0, st), "gn_apply")
        call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            call()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        gb = x.numel() * 2 * (3 if with_res else 2) / 1e9
        row[name + "_us"] = round(us, 1)
        row[name + "_GB_s"] = round(gb / (us * 1e-6))
        ys.append(y)
    L.tcv_set_debug_flags(0)
    row["equal"] = bool(torch.equal(ys[0], ys[1]) and torch.equal(ys[0], ys[2]))
    print(json.dumps(row))
    assert row["equal"]
