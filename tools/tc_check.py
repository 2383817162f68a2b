"""Standalone checks of the tcgen05 kernels against fp32/fp64 references (each case in its own
process so a trapped kernel cannot poison the others).  python tools/tc_check.py [case ...]"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def split(x):
    import torch
    hi = x.bfloat16()
    lo = (x - hi.float()).bfloat16()
    return torch.stack([hi, lo]).contiguous()


def case_gemm(nsplit, M, N, K, batch, out_bf16=0):
    import torch
    from tcvom_b200 import _cabi
    L = _cabi.lib()
    torch.manual_seed(0)
    dev = "cuda"
    A = torch.randn(batch, M, K, device=dev)
    B = torch.randn(batch, N, K, device=dev)
    ldc = (N + 63) // 64 * 64
    Cm = torch.full((batch, M, ldc), 7.0, device=dev, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    st = torch.cuda.current_stream().cuda_stream
    if nsplit == 6:
        def split3(x):
            h = x.bfloat16(); m = (x - h.float()).bfloat16(); l = (x - h.float() - m.float()).bfloat16()
            return torch.stack([h, m, l]).contiguous()
        As, Bs = split3(A), split3(B)
        ref = torch.matmul(A.double(), B.double().transpose(1, 2))
        rc = L.tcv_gemm_tn_tc(As.data_ptr(), batch * M * K, Bs.data_ptr(), batch * N * K, Cm.data_ptr(), M, N, K, ldc,
                              M * ldc, batch, 6, out_bf16, 0, st)
    elif nsplit == 3:
        As, Bs = split(A), split(B)
        ref = torch.matmul(A.double(), B.double().transpose(1, 2))
        rc = L.tcv_gemm_tn_tc(As.data_ptr(), batch * M * K, Bs.data_ptr(), batch * N * K, Cm.data_ptr(), M, N, K, ldc,
                              M * ldc, batch, 3, out_bf16, 0, st)
    else:
        Ab, Bb = A.bfloat16(), B.bfloat16()
        ref = torch.matmul(Ab.double(), Bb.double().transpose(1, 2))
        rc = L.tcv_gemm_tn_tc(Ab.data_ptr(), 0, Bb.data_ptr(), 0, Cm.data_ptr(), M, N, K, ldc, M * ldc, batch, 1,
                              out_bf16, 0, st)
    _cabi.check(rc, "gemm_tn_tc")
    torch.cuda.synchronize()
    got = Cm[:, :, :N].double()
    err = (got - ref).abs().max().item()
    rel = err / ref.abs().max().item()
    pad_ok = bool((Cm[:, :, N:].float() == 7.0).all()) if ldc > N else True
    print(f"gemm nsplit={nsplit} M={M} N={N} K={K} b={batch} bf16out={out_bf16}: max abs err {err:.3e} rel {rel:.3e} pad_untouched={pad_ok}")
    tol = 2e-2 if out_bf16 else (1e-4 if nsplit == 3 else (1e-5 if nsplit == 6 else 1e-5))
    assert rel < tol and pad_ok


def conv_ref_fp64(x, wf, taps, wt, stride, gh, gw, oh, ow, mul, offs, s1, b1, res1, s2, b2, res2):
    """fp64 torch restatement of what tcv_conv2d computes (include/tcvom_b200.h): zero-padded taps, BatchNorm affine,
    nearest-upsampled residual, LeakyReLU(0.2), second affine, second residual, strided placement."""
    import torch
    n, h, w, cin = x.shape
    P = 4
    xp = torch.nn.functional.pad(x.double(), (0, 0, P, P + 2, P, P + 2))
    acc = torch.zeros((n, gh, gw, wf.shape[2]), dtype=torch.float64, device=x.device)
    for (dy, dx), k in zip(taps, wt):
        sl = xp[:, P + dy: P + dy + (gh - 1) * stride + 1: stride, P + dx: P + dx + (gw - 1) * stride + 1: stride, :]
        acc += torch.einsum("nhwc,cd->nhwd", sl, wf[k].double())
    t = acc * s1.double() + b1.double()
    oy = torch.arange(gh, device=x.device) * mul + offs[0]
    ox = torch.arange(gw, device=x.device) * mul + offs[1]
    t = t + res1.double()[:, (oy // 2)][:, :, (ox // 2)]
    t = torch.where(t > 0, t, 0.2 * t)
    t = t * s2.double() + b2.double()
    t = t + res2.double()[:, oy][:, :, ox]
    out = torch.zeros((n, oh, ow, wf.shape[2]), dtype=torch.float64, device=x.device)
    out[:, oy[:, None], ox[None, :]] = t
    return out, (oy, ox)


def case_wgrad(cin, cout, h, w, n, prepad=False, flags=0, perm=False):
    """tcv_conv2d_wgrad_nhwc_tc (3x3, stride 1) against torch fp64 autograd.  prepad: x carries its own 1-pixel border and
    the taps are (0..2, 0..2) (the reflect-padded layers); perm: the nine taps in a shuffled order / shuffled weight slots."""
    import torch
    import torch.nn.functional as F
    from tcvom_b200 import _cabi
    from tcvom_b200._cabi import ConvDesc
    L = _cabi.lib()
    torch.manual_seed(3)
    dev = "cuda"
    st = torch.cuda.current_stream().cuda_stream
    ih, iw = (h + 2, w + 2) if prepad else (h, w)
    x = split(torch.randn(n, ih, iw, cin, device=dev))
    dz = split(torch.randn(n, h, w, cout, device=dev))
    off = 0 if prepad else -1
    taps = [(ky + off, kx + off) for ky in range(3) for kx in range(3)]
    order = list(range(9))
    slots = list(range(9))
    if perm:
        order = [4, 0, 8, 2, 6, 1, 3, 5, 7]
        slots = [8, 7, 6, 5, 4, 3, 2, 1, 0]
    d = ConvDesc()
    d.x, d.x_plane, d.x_img_stride = x.data_ptr(), x[0].numel(), ih * iw * cin
    d.n, d.ih, d.iw, d.cin = n, ih, iw, cin
    d.ntaps = 9
    for i, t in enumerate(order):
        d.dy[i], d.dx[i], d.wtap[i] = taps[t][0], taps[t][1], slots[t]
    d.stride, d.pad_mode = 1, 0
    d.oh, d.ow, d.cout, d.gh, d.gw = h, w, cout, h, w
    d.oy_mul, d.oy_off, d.ox_mul, d.ox_off = 1, 0, 1, 0
    dw = torch.zeros((9, cin, cout), device=dev)
    L.tcv_set_debug_flags(flags)
    try:
        _cabi.check(L.tcv_conv2d_wgrad_nhwc_tc(C.byref(d), dz.data_ptr(), dz[0].numel(), cout, dw.data_ptr(), st), "wgrad_nhwc")
        torch.cuda.synchronize()
    finally:
        L.tcv_set_debug_flags(0)
    x64 = (x[0].double() + x[1].double()).permute(0, 3, 1, 2)
    z64 = (dz[0].double() + dz[1].double()).permute(0, 3, 1, 2)
    wz = torch.zeros((cout, cin, 3, 3), device=dev, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x64, wz, None, 1, 0 if prepad else 1)
    (gw,) = torch.autograd.grad(y, wz, z64)
    ref = gw.permute(2, 3, 1, 0).reshape(9, cin, cout)          # [tap][ci][co]
    got = torch.empty_like(ref)
    for t in range(9):
        got[t] = dw[slots[t]].double()
    err = float((got - ref).abs().max() / ref.abs().max())
    print(f"wgrad cin={cin} cout={cout} {h}x{w} n={n} prepad={prepad} flags={flags} perm={perm}: rel err {err:.2e}")
    assert err < 1e-4, err
    return err


def case_gemm_ex(M, N, K, batch, a_mn, b_mn):
    """tcv_gemm_tc_ex: C = A . B^T on the CTA-pair kernel with K-major / MN-major operands (padded pitches, ragged K)."""
    import torch
    from tcvom_b200 import _cabi
    L = _cabi.lib()
    torch.manual_seed(4)
    dev = "cuda"
    st = torch.cuda.current_stream().cuda_stream
    A = torch.randn(batch, M, K, device=dev, dtype=torch.float64)
    B = torch.randn(batch, N, K, device=dev, dtype=torch.float64)

    def operand(X, mn):      # -> split planes with a padded pitch, filled with NaN outside the matrix
        rows = X.shape[1]
        if mn:
            ld = (rows + 63) // 64 * 64 + 64
            buf = torch.full((batch, K, ld), float("nan"), device=dev)
            buf[:, :, :rows] = X.transpose(1, 2).float()
        else:
            ld = (K + 7) // 8 * 8 + 8
            buf = torch.full((batch, rows, ld), float("nan"), device=dev)
            buf[:, :, :K] = X.float()
        planes = split(buf)
        return planes, ld, buf.shape[1] * ld

    a, a_ld, a_bs = operand(A, a_mn)
    b, b_ld, b_bs = operand(B, b_mn)
    ldc = (N + 3) // 4 * 4
    C_ = torch.zeros((batch, M, ldc), device=dev)
    _cabi.check(L.tcv_gemm_tc_ex(a.data_ptr(), a[0].numel(), a_ld, a_bs, int(a_mn), b.data_ptr(), b[0].numel(), b_ld, b_bs,
                                 int(b_mn), C_.data_ptr(), M, N, K, ldc, M * ldc, batch, st), "gemm_tc_ex")
    torch.cuda.synchronize()
    a64 = (a[0].double() + a[1].double())
    b64 = (b[0].double() + b[1].double())
    A64 = a64[:, :, :M].transpose(1, 2) if a_mn else a64[:, :, :K]
    B64 = b64[:, :, :N].transpose(1, 2) if b_mn else b64[:, :, :K]
    ref = A64 @ B64.transpose(1, 2)
    err = float((C_[:, :, :N].double() - ref).abs().max() / ref.abs().max())
    print(f"gemm_ex M={M} N={N} K={K} batch={batch} a_mn={a_mn} b_mn={b_mn}: rel err {err:.2e}")
    assert err < 3e-5, err
    return err


def case_conv_stats(cin, cout, h, w, n, groups, dil=1):
    """tcv_conv_desc.stats: per-(image group, channel) sum / sum of squares of the conv output accumulated by the epilogue
    of the CTA-pair kernel, against the same sums taken from the stored output tensor."""
    import torch
    from tcvom_b200 import _cabi
    from tcvom_b200._cabi import ConvDesc
    L = _cabi.lib()
    torch.manual_seed(2)
    dev = "cuda"
    st = torch.cuda.current_stream().cuda_stream
    x = split(torch.randn(n, h, w, cin, device=dev))
    taps = [((ky - 1) * dil, (kx - 1) * dil) for ky in range(3) for kx in range(3)]
    wf = (torch.randn(9, cin, cout, device=dev) / (cin * 9) ** 0.5).contiguous()
    wtc = torch.empty((2, 9, cout, cin), dtype=torch.bfloat16, device=dev)
    _cabi.check(L.tcv_pack_weight_tc(wf.data_ptr(), 9, cin, cout, wtc.data_ptr(), st), "pack")
    bias = torch.randn(cout, device=dev)
    outs = []
    for with_stats in (False, True):
        y = torch.zeros((2, n, h, w, cout), dtype=torch.bfloat16, device=dev)
        d = ConvDesc()
        d.x = x.data_ptr(); d.n, d.ih, d.iw, d.cin = n, h, w, cin
        d.w = wf.data_ptr(); d.w_tc = wtc.data_ptr(); d.w_tc_taps = 9; d.ntaps = 9
        for i, (dy, dx) in enumerate(taps):
            d.dy[i], d.dx[i], d.wtap[i] = dy, dx, i
        d.stride, d.pad_mode = 1, 0
        d.y = y.data_ptr()
        d.oh, d.ow, d.cout, d.gh, d.gw = h, w, cout, h, w
        d.oy_mul, d.oy_off, d.ox_mul, d.ox_off = 1, 0, 1, 0
        d.b1 = bias.data_ptr()
        assert L.tcv_conv2d_path(C.byref(d)) == 4
        copies = 5
        sums = torch.full((copies, groups, cout, 2), 7.0, dtype=torch.float64, device=dev)
        if with_stats:
            _cabi.check(L.tcv_zero_bytes(sums.data_ptr(), sums.numel() * 8, st), "zero_bytes")
            d.stats, d.stats_groups, d.stats_copies = sums.data_ptr(), groups, copies
        _cabi.check(L.tcv_conv2d(C.byref(d), st), "conv2d")
        torch.cuda.synchronize()
        outs.append((y, sums))
    (y0, _), (y1, sums) = outs
    assert all(float(sums[k].abs().sum()) > 0 for k in range(sums.shape[0]))     # every accumulator copy was used
    sums = sums.sum(dim=0)
    assert torch.equal(y0, y1)                                  # the statistics do not change the output
    v = (y1[0].double() + y1[1].double()).reshape(n // groups, groups, h * w, cout)
    ref = torch.stack([v.sum(dim=(0, 2)), (v * v).sum(dim=(0, 2))], dim=-1)
    err = float(((sums - ref).abs() / ref.abs().clamp(min=1.0)).max())
    print(f"conv stats cin={cin} cout={cout} {h}x{w} n={n} groups={groups} dil={dil}: rel err {err:.2e}")
    assert err < 2e-4, err                                      # fp32 accumulator vs the split-bf16 value it is stored as
    return err


def case_conv(cin, cout, h, w, n, kind, f32_side=True, expect_path=None):
    import torch
    from tcvom_b200 import _cabi
    from tcvom_b200._cabi import ConvDesc
    L = _cabi.lib()
    torch.manual_seed(1)
    dev = "cuda"
    st = torch.cuda.current_stream().cuda_stream
    x = split(torch.randn(n, h, w, cin, device=dev))
    if kind == "3x3":
        taps = [(ky - 1, kx - 1) for ky in range(3) for kx in range(3)]; wt = list(range(9)); ntw = 9
        oh, ow, gh, gw, mul, offs = h, w, h, w, 1, (0, 0)
    elif kind == "3x3s2":
        taps = [(ky - 1, kx - 1) for ky in range(3) for kx in range(3)]; wt = list(range(9)); ntw = 9
        oh, ow = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
        gh, gw, mul, offs = oh, ow, 1, (0, 0)
    elif kind == "1x1":
        taps = [(0, 0)]; wt = [0]; ntw = 1
        oh, ow, gh, gw, mul, offs = h, w, h, w, 1, (0, 0)
    else:  # one phase (py=1, px=0) of the 4x4 stride-2 transposed conv
        kys = [(0, 1), (2, 0)]; kxs = [(1, 0), (3, -1)]
        taps = [(dy, dx) for ky, dy in kys for kx, dx in kxs]
        wt = [ky * 4 + kx for ky, dy in kys for kx, dx in kxs]; ntw = 16
        oh, ow, gh, gw, mul, offs = 2 * h, 2 * w, h, w, 2, (1, 0)
    wf = (torch.randn(ntw, cin, cout, device=dev) / (cin * len(taps)) ** 0.5).contiguous()
    wtc = torch.empty((2, ntw, cout, cin), dtype=torch.bfloat16, device=dev)
    _cabi.check(L.tcv_pack_weight_tc(wf.data_ptr(), ntw, cin, cout, wtc.data_ptr(), st), "pack")
    wfold = None
    if cin == 8 and kind == "3x3":
        wfold = torch.empty((2, 3, cout, 32), dtype=torch.bfloat16, device=dev)
        _cabi.check(L.tcv_pack_weight_fold(wf.data_ptr(), cout, wfold.data_ptr(), st), "pack_fold")
    s1 = torch.rand(cout, device=dev) + 0.5; b1 = torch.randn(cout, device=dev)
    s2 = torch.rand(cout, device=dev) + 0.5; b2 = torch.randn(cout, device=dev)
    res1 = split(torch.randn(n, oh // 2, ow // 2, cout, device=dev))
    res2 = split(torch.randn(n, oh, ow, cout, device=dev))
    outs = []
    for use_tc in (0, 1):
        y = torch.zeros((2, n, oh, ow, cout), dtype=torch.bfloat16, device=dev)
        yf = torch.zeros((n, oh, ow, cout), dtype=torch.float32, device=dev)
        d = ConvDesc()
        d.x = x.data_ptr(); d.n, d.ih, d.iw, d.cin = n, h, w, cin
        d.w = wf.data_ptr(); d.ntaps = len(taps)
        if use_tc:
            if cin % 32 == 0:
                d.w_tc = wtc.data_ptr(); d.w_tc_taps = ntw
            if wfold is not None:
                d.w_tc_fold = wfold.data_ptr()
        for i, (dy, dx) in enumerate(taps):
            d.dy[i], d.dx[i], d.wtap[i] = dy, dx, wt[i]
        d.stride, d.pad_mode = (2 if kind == "3x3s2" else 1), 0
        d.y = y.data_ptr()
        if cout != 32 and f32_side:
            d.y_f32 = yf.data_ptr()   # (the narrow-layer and the CTA-pair kernels have no fp32 side output)
        d.oh, d.ow, d.cout, d.gh, d.gw = oh, ow, cout, gh, gw
        d.oy_mul, d.oy_off, d.ox_mul, d.ox_off = mul, offs[0], mul, offs[1]
        d.s1, d.b1 = s1.data_ptr(), b1.data_ptr()
        d.res1, d.res1_shift = res1.data_ptr(), 1
        d.act = 2
        d.s2, d.b2 = s2.data_ptr(), b2.data_ptr()
        d.res2 = res2.data_ptr()
        path = L.tcv_conv2d_path(C.byref(d))
        assert (path > 0) == bool(use_tc), "dispatch did not pick the expected path"
        if use_tc and expect_path is not None:
            assert path == expect_path, f"expected kernel path {expect_path}, dispatch chose {path}"
        _cabi.check(L.tcv_conv2d(C.byref(d), st), "conv2d")
        torch.cuda.synchronize()
        ys = y[0].float() + y[1].float()
        outs.append((ys, yf.clone() if (cout != 32 and f32_side) else ys))
    err = (outs[0][1] - outs[1][1]).abs().max().item()
    errs = (outs[0][0] - outs[1][0]).abs().max().item()
    mag = outs[0][1].abs().max().item()
    # independent reference: fp64 torch on the same (split-bf16 exact) operands
    xs = x[0].float().double() + x[1].float().double()
    r1 = res1[0].float().double() + res1[1].float().double()
    r2 = res2[0].float().double() + res2[1].float().double()
    ref, (oy, ox) = conv_ref_fp64(xs, wf, taps, wt, 2 if kind == "3x3s2" else 1, gh, gw, oh, ow, mul, offs, s1, b1, r1, s2, b2, r2)
    sel = lambda t: t.double()[:, oy[:, None], ox[None, :]]
    e_tc = (sel(outs[1][0]) - sel(ref)).abs().max().item()
    e_dir = (sel(outs[0][0]) - sel(ref)).abs().max().item()
    untouched = True
    if mul == 2:      # a transposed-conv phase must leave the other three phases alone
        mask = torch.ones((oh, ow), dtype=torch.bool, device=dev)
        mask[oy[:, None], ox[None, :]] = False
        untouched = bool((outs[1][0][:, mask] == 0).all())
    print(f"conv {kind} {cin}->{cout} {h}x{w} n={n} path={path}: tc vs direct max abs err f32 {err:.3e} split {errs:.3e}; "
          f"vs fp64 torch: tc {e_tc:.3e} direct {e_dir:.3e} (max |y| {mag:.2f})")
    assert err < 2e-4 * max(1.0, mag) and errs < 2e-4 * max(1.0, mag)
    assert e_tc < 1e-4 * max(1.0, mag) and e_dir < 1e-4 * max(1.0, mag) and untouched


CASES = {
    "gemm1_small": lambda: case_gemm(1, 128, 256, 64, 1),
    "gemm1": lambda: case_gemm(1, 300, 520, 192, 2),
    "gemm1_bf16out": lambda: case_gemm(1, 300, 520, 192, 2, 1),
    "gemm3_small": lambda: case_gemm(3, 128, 128, 64, 1),
    "gemm3": lambda: case_gemm(3, 1000, 1000, 576, 2),
    "gemm6": lambda: case_gemm(6, 1000, 1000, 576, 2),
    "conv3x3_64_128": lambda: case_conv(64, 128, 20, 28, 2, "3x3"),
    "conv3x3_32_32": lambda: case_conv(32, 32, 16, 48, 1, "3x3"),
    "conv3x3_256_256": lambda: case_conv(256, 256, 10, 12, 3, "3x3"),
    "conv3x3_128_128_big": lambda: case_conv(128, 128, 136, 240, 3, "3x3"),
    "conv3x3_32_32_big": lambda: case_conv(32, 32, 272, 480, 2, "3x3"),
    "conv3x3_512_512": lambda: case_conv(512, 512, 34, 60, 3, "3x3"),
    "pair_3x3_128_128_big": lambda: case_conv(128, 128, 136, 240, 3, "3x3", False, 4),
    "pair_3x3_256_256": lambda: case_conv(256, 256, 68, 120, 3, "3x3", False, 4),
    "pair_3x3_512_512": lambda: case_conv(512, 512, 34, 60, 3, "3x3", False, 4),
    "pair_3x3_256_128_ragged": lambda: case_conv(256, 128, 10, 12, 3, "3x3", False, 4),
    "pair_1x1_128_256": lambda: case_conv(128, 256, 18, 22, 2, "1x1", False, 4),
    "pair_3x3_64_64_stacked": lambda: case_conv(64, 64, 40, 56, 2, "3x3", False, 4),
    "pair_3x3_128_64_stacked": lambda: case_conv(128, 64, 272, 480, 1, "3x3", False, 4),
    "pair_deconv_64_64_stacked": lambda: case_conv(64, 64, 9, 14, 2, "deconv", False, 4),
    "pair_deconv_256_256": lambda: case_conv(256, 256, 9, 14, 2, "deconv", False, 4),
    "conv3x3s2_64_128": lambda: case_conv(64, 128, 40, 56, 2, "3x3s2"),
    "conv3x3s2_32_64": lambda: case_conv(32, 64, 36, 52, 1, "3x3s2"),
    "conv3x3_8_32_fold": lambda: case_conv(8, 32, 40, 56, 2, "3x3"),
    "conv3x3_64_32": lambda: case_conv(64, 32, 36, 44, 2, "3x3"),
    "deconv_32_32": lambda: case_conv(32, 32, 18, 20, 1, "deconv"),
    "conv1x1_64_32": lambda: case_conv(64, 32, 20, 24, 1, "1x1"),
    "conv1x1_128_64": lambda: case_conv(128, 64, 16, 16, 2, "1x1"),
    "deconv_64_64": lambda: case_conv(64, 64, 9, 14, 2, "deconv"),
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    if len(names) == 1 and os.environ.get("TC_CHECK_CHILD") == "1":
        CASES[names[0]]()
        sys.exit(0)
    bad = 0
    for nme in names:
        env = dict(os.environ, TC_CHECK_CHILD="1")
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), nme], env=env, capture_output=True,
                               text=True, timeout=120)
            out = (r.stdout + r.stderr).strip().splitlines()
            print(f"[{nme}] rc={r.returncode} :: " + " | ".join(out[-4:]), flush=True)
            bad += r.returncode != 0
        except subprocess.TimeoutExpired:
            print(f"[{nme}] TIMEOUT", flush=True)
            bad += 1
    sys.exit(1 if bad else 0)
