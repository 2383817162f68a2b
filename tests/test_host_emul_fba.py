"""Host-logic tests of the FBA path in the GPU-less build container.

The HOST program of ``FbaVmnEngine`` (which C-ABI calls, in which order, on which buffers and descriptors) is run on
host memory against a *test double* of the C ABI (tests/host_emul/emul.cpp, compiled here with g++) and compared
with the oracle / the reference golden vectors.  The double executes the very same per-work-item bodies as the
CUDA kernels for the FBA element-wise ops (tcvom_b200/csrc/fba_body.h) and a naive restatement of tcv_conv2d /
tcv_tam_attend.  This is test infrastructure: the product cannot load it (tcvom_b200._cabi only knows
libtcvom_b200.so) and the real parity tests remain the ``-m gpu`` ones.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import ROOT, fixture_sd_fba, golden

EMU_DIR = os.path.join(ROOT, "tests", "host_emul")
EMU_SO = os.path.join(EMU_DIR, "_emul.so")


@pytest.fixture(scope="module")
def emu():
    from tcvom_b200 import _cabi
    src = os.path.join(EMU_DIR, "emul.cpp")
    body = os.path.join(ROOT, "tcvom_b200", "csrc", "fba_body.h")
    if not os.path.exists(EMU_SO) or os.path.getmtime(EMU_SO) < max(os.path.getmtime(src), os.path.getmtime(body)):
        subprocess.check_call(["g++", "-O3", "-march=native", "-fopenmp", "-shared", "-fPIC", "-std=c++17",
                               "-o", EMU_SO, src])
    L = ctypes.CDLL(EMU_SO)
    for name, (res, args) in _cabi.SIGNATURES.items():
        if hasattr(L, name):
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
    old = _cabi._lib
    _cabi._lib = L
    yield L
    _cabi._lib = old


def make_engine(window=7):
    from tcvom_b200.fba_engine import FbaVmnEngine

    class HostEmuEngine(FbaVmnEngine):
        """FbaVmnEngine on host memory: no device check, null stream."""

        @staticmethod
        def _check_device(dev):
            pass

        def _stream_ptr(self):
            return 0

    eng = HostEmuEngine(window)
    eng.device = torch.device("cpu")
    return eng


def to_act(t: torch.Tensor, c_pad=None):
    """fp32 NCHW -> split-bf16 NHWC Act on the host."""
    from tcvom_b200.engine import Act
    n, c, h, w = t.shape
    c_pad = c_pad or c
    a = Act.empty(n, h, w, c_pad, torch.device("cpu"))
    x = torch.zeros(n, h, w, c_pad)
    x[..., :c] = t.permute(0, 2, 3, 1)
    hi = x.bfloat16()
    a.buf[0].copy_(hi)
    a.buf[1].copy_((x - hi.float()).bfloat16())
    return a


def from_act(a, c=None):
    return a.float()[..., : (c or a.c)].permute(0, 3, 1, 2).contiguous()


# ------------------------------------------------------------------------------------------ kernel bodies vs torch
def test_groupnorm_bodies_match_torch(emu):
    eng = make_engine()
    torch.manual_seed(0)
    x = torch.randn(2, 64, 5, 7) * 2 + 0.3
    res = torch.randn(2, 64, 5, 7)
    g, b = torch.rand(64) + 0.5, torch.randn(64) * 0.1
    eng.gn_params["p"] = (g, b)
    xa, ra = to_act(x), to_act(res)
    y = eng.gn(xa, "p", 1, res=ra)
    ref = F.relu(F.group_norm(from_act(xa), 32, g, b, 1e-5) + from_act(ra))
    assert float((from_act(y) - ref).abs().max()) < 6e-5      # split-bf16 storage: 2^-17 relative
    # write into a channel slice of a wider tensor, leaky 0.01
    from tcvom_b200.engine import Act
    wide = Act.empty(2, 5, 7, 96, torch.device("cpu"))
    wide.buf.zero_()
    eng.gn(xa, "p", 4, out=wide, out_off=16)
    ref = F.leaky_relu(F.group_norm(from_act(xa), 32, g, b, 1e-5), 0.01)
    got = wide.float().permute(0, 3, 1, 2)
    assert float((got[:, 16:80] - ref).abs().max()) < 6e-5
    assert float(got[:, :16].abs().max()) == 0 and float(got[:, 80:].abs().max()) == 0


def test_pool_resize_copy_bodies_match_torch(emu):
    eng = make_engine()
    from tcvom_b200.engine import Act
    torch.manual_seed(1)
    x = torch.randn(2, 16, 9, 12)
    xa = to_act(x)
    xf = from_act(xa)
    assert torch.equal(from_act(eng.maxpool(xa)), F.max_pool2d(xf, 3, 2, 1))
    for (oh, ow) in ((18, 24), (17, 5), (9, 12), (3, 40)):
        out = Act.empty(2, oh, ow, 32, torch.device("cpu"))
        out.buf.zero_()
        eng.bilinear(xa, oh, ow, out, 8)
        ref = F.interpolate(xf, (oh, ow), mode="bilinear", align_corners=False)
        assert float((from_act(out)[:, 8:24] - ref).abs().max()) < 3e-5, (oh, ow)
    out = Act.empty(2, 18, 24, 16, torch.device("cpu"))
    eng.bilinear(xa, 18, 24, out, 0)
    assert float((from_act(out) - F.interpolate(xf, scale_factor=2, mode="bilinear", align_corners=False)).abs().max()) < 3e-5
    # adaptive average pooling of a channel slice (incl. more bins than pixels)
    x = torch.randn(1, 128, 4, 7)
    xa = to_act(x)
    for s in (1, 2, 3, 6):
        pooled = Act.empty(1, s, s, 64, torch.device("cpu"))
        eng._call("tcv_adaptive_avgpool", xa.ptr, xa.plane, 1, 4, 7, 64, 128, 64, s, pooled.ptr)
        ref = F.adaptive_avg_pool2d(from_act(xa)[:, 64:], s)
        assert float((from_act(pooled) - ref).abs().max()) < 2e-5, s
    dst = Act.empty(1, 4, 7, 192, torch.device("cpu"))
    dst.buf.zero_()
    eng.copy_channels(xa, 32, 64, dst, 128)
    assert torch.equal(from_act(dst)[:, 128:], from_act(xa)[:, 32:96])


def test_weight_standardisation_pack(emu):
    from oracle import vmn_fba_oracle as O
    eng = make_engine()
    w = torch.randn(7, 11, 3, 3)
    from tcvom_b200 import _cabi
    eng._pack_fba(_cabi.lib(), 0, "k", w, True)
    ent = eng.w["k"]
    assert ent["cin"] == 16 and ent["cout"] == 8 and tuple(ent["w"].shape) == (9, 16, 8)
    ref = O.ws_weight(w).permute(2, 3, 1, 0).reshape(9, 11, 7)
    assert float((ent["w"][:, :11, :7] - ref).abs().max()) < 1e-5
    assert float(ent["w"][:, 11:].abs().max()) == 0 and float(ent["w"][:, :, 7:].abs().max()) == 0


def test_conv_double_matches_torch(emu):
    """The naive tcv_conv2d of the test double itself (dilation, stride, bias, the chained 7x7 stem)."""
    eng = make_engine()
    from tcvom_b200 import _cabi
    torch.manual_seed(2)
    x = torch.randn(2, 32, 9, 10)
    xa = to_act(x)
    xf = from_act(xa)
    for (cout, k, stride, dil) in ((64, 3, 1, 1), (64, 3, 1, 2), (32, 3, 1, 4), (64, 3, 2, 1), (64, 1, 2, 1)):
        w = torch.randn(cout, 32, k, k) * 0.1
        b = torch.randn(cout)
        eng._pack_fba(_cabi.lib(), 0, "c", w, False)
        eng.w["c"]["w"] = eng.w["c"]["w"].clone()
        eng.bias["c"] = b
        y = eng.convf(xa, "c", stride=stride, dilation=dil, bias=True, act=4)
        ref = F.leaky_relu(F.conv2d(xf, w, b, stride, dil * (k // 2), dil), 0.01)
        assert float((from_act(y) - ref).abs().max()) < 1e-4, (cout, k, stride, dil)
        del eng.w["c"]
    x = torch.randn(1, 11, 12, 14)
    xa = to_act(x, 16)
    w = torch.randn(64, 11, 7, 7) * 0.05
    eng._pack_fba(_cabi.lib(), 0, "stem", w, False)
    y = eng.conv7x7s2(xa, "stem")
    ref = F.conv2d(from_act(xa, 11), w, None, 2, 3)
    assert float((from_act(y) - ref).abs().max()) < 1e-4


def test_stem_space_to_depth_matches_torch(emu):
    """7x7 / stride 2 / pad 3 == 4x4 / stride 1 (taps -2..1) over the 2x2 space-to-depth image (the tcgen05 stem)."""
    from tcvom_b200 import _cabi
    from tcvom_b200.fba_engine import STEM
    eng = make_engine()
    torch.manual_seed(3)
    x = torch.randn(2, 11, 12, 18)
    xa = to_act(x, 16)
    w = torch.randn(64, 11, 7, 7) * 0.05
    eng._pack_fba(_cabi.lib(), 0, STEM, w, True)
    y = eng.stem_s2d(xa, STEM)
    from oracle import vmn_fba_oracle as O
    ref = F.conv2d(from_act(xa, 11), O.ws_weight(w), None, 2, 3)
    assert tuple(y.buf.shape[1:]) == (2, 6, 9, 64)
    tol = 2e-5 * max(1.0, float(ref.abs().max()))           # split-bf16 storage: 2^-17 relative
    assert float((from_act(y) - ref).abs().max()) < tol
    assert float((from_act(eng.conv7x7s2(xa, STEM)) - ref).abs().max()) < tol


# ------------------------------------------------------------------------------------------ input encoding
@pytest.mark.parametrize("name", ["fba_ring64.npz"])
def test_input_encoding_matches_reference(emu, name):
    g = golden(name)
    eng = make_engine()
    from tcvom_b200.engine import Act
    for u8 in (True, False):
        imgs = torch.from_numpy(g["imgs"])
        tris = torch.from_numpy(g["tris"])
        if not u8:
            imgs, tris = imgs.float(), tris.float()
        B, S, _, H, W = imgs.shape
        x16 = Act.empty(B * S, H, W, 16, torch.device("cpu"))
        eng.encode_inputs(imgs.contiguous(), tris.contiguous(), B * S, H, W, x16)
        got = from_act(x16)
        ref = torch.from_numpy(g["x11"]).reshape(B * S, 11, H, W)
        assert float((got[:, :11] - ref).abs().max()) < 2e-5       # incl. the six distance-transform channels
        scaled = imgs.float().flip([2]).reshape(B * S, 3, H, W) / 255
        assert float((got[:, 11:14] - scaled).abs().max()) < 1e-5
        assert float(got[:, 14:].abs().max()) == 0


def test_distance_transform_edge_cases(emu):
    """no seed of one kind in a frame -> zeros (cv2 returns +huge, exp -> 0); seeds everywhere -> ones."""
    from oracle import vmn_fba_oracle as O
    from tcvom_b200.engine import Act
    eng = make_engine()
    H, W = 32, 40
    tris = torch.zeros(3, 1, H, W, dtype=torch.uint8)
    tris[0] = 128                                            # all unknown: no bg, no fg
    tris[1, :, :, :7] = 255                                  # fg stripe, bg elsewhere
    tris[2] = 255                                            # all fg
    tris[2, 0, 5, 9] = 0                                     # a single bg pixel
    imgs = torch.zeros(3, 3, H, W, dtype=torch.uint8)
    x16 = Act.empty(3, H, W, 16, torch.device("cpu"))
    eng.encode_inputs(imgs, tris, 3, H, W, x16)
    got = from_act(x16)
    _, x11, _, _ = O.eval_preprocess(imgs[None].float(), tris[None].float())
    assert float((got[:, 3:11] - x11[0, :, 3:11]).abs().max()) < 2e-5
    assert float(got[0, 3:9].abs().max()) == 0
    # distances beyond the capped search radius (d^2 > 145000: every feature < 1e-12) are written as exact zeros
    H, W = 8, 1312
    tris = torch.full((1, 1, H, W), 128, dtype=torch.uint8)
    tris[0, 0, 3, 5] = 255
    tris[0, 0, 6, 1300] = 0
    imgs = torch.zeros(1, 3, H, W, dtype=torch.uint8)
    x16 = Act.empty(1, H, W, 16, torch.device("cpu"))
    eng.encode_inputs(imgs, tris, 1, H, W, x16)
    got = from_act(x16)
    _, x11, _, _ = O.eval_preprocess(imgs[None].float(), tris[None].float())
    assert float((got[:, 3:11] - x11[0, :, 3:11]).abs().max()) < 2e-5
    assert float(got[0, 6:9, :, 600:].abs().max()) == 0 and float(x11[0, 0, 6:9, :, 600:].max()) < 1e-12


# ------------------------------------------------------------------------------------------ the whole program
@pytest.mark.parametrize("name", ["ring64", "dil64x96", "batch2_64"])
def test_eval_program_matches_reference_golden(emu, name):
    import tcvom_b200
    g = golden(f"fba_{name}.npz")
    net = tcvom_b200.get_VMN_models("vmn_fba", agg_window=7)
    net.load_state_dict(fixture_sd_fba(), strict=True)
    net.eval()
    eng = make_engine()
    eng.refresh_weights(net)
    imgs, tris = torch.from_numpy(g["imgs"]), torch.from_numpy(g["tris"])
    B, S, _, H, W = imgs.shape
    # record once on zero inputs (as EvalModel._plan_fba does), then fill the inputs and replay the recorded calls
    from tcvom_b200.engine import Plan
    plan = Plan()
    eng._rec = plan
    io = eng.eval_program(B, S, H, W, int(g["dilate"]), True)
    eng._rec = None
    io["imgs"].copy_(imgs); io["tris"].copy_(tris)
    plan.replay(0)
    assert np.array_equal(io["trimask"].reshape(B, S, 1, H, W).numpy().astype(np.uint8), g["trimask"])
    assert np.array_equal(io["small_mask"][:, 0].numpy().astype(bool), g["small_mask1"])
    err = lambda a, b: float(np.abs(a - b.astype(np.float32)).max())
    if "feat1" in g.files:
        feat = (io["feat"][0].float() + io["feat"][1].float())[1:2].permute(0, 3, 1, 2).numpy()
        assert err(feat, g["feat1"]) < 2e-3
    assert err(io["pred"][:, 0].numpy(), g["pred1"]) < 1e-3
    assert err(io["alphas"].numpy(), g["alphas"]) < 1e-3          # north_star bar: 1e-3 on the alpha matte
    assert err(io["Fs"][:, 1].numpy(), g["Fs"]) < 2e-3            # golden stored as fp16
    assert err(io["Bs"][:, 1].numpy(), g["Bs"]) < 2e-3
    for k, ref in (("attb", g["attb1"]), ("attf", g["attf1"])):
        assert err(io[k][:, 0].numpy(), ref) < 3e-3 * max(1.0, float(np.abs(ref.astype(np.float32)).max()))
    assert float(io["alphas"][:, 0].abs().max()) == 0 and float(io["alphas"][:, -1].abs().max()) == 0


def test_frame_stream_equals_windowed_program(emu):
    """tcvom_b200.FrameStream (per-frame feature reuse across sliding windows) against the windowed
    program on every window of a 5-frame clip: same kernels on the same values."""
    import tcvom_b200
    from tcvom_b200 import synthetic
    from tcvom_b200.engine import Plan
    from tcvom_b200.stream import FrameStream
    net = tcvom_b200.get_VMN_models("vmn_fba", agg_window=7)
    net.load_state_dict(fixture_sd_fba(), strict=True)
    m = tcvom_b200.EvalModel(model="vmn_fba", agg_window=7, dilate_kernel=2)
    m.NET = net
    m.eval()
    eng = make_engine()
    eng.refresh_weights(net)
    H, W = 32, 64
    imgs, tris = synthetic.make_window(H, W, seed=4, frames=5)
    imgs, tris = torch.from_numpy(imgs), torch.from_numpy(tris)

    class HostStream(FrameStream):
        def _run(self, plan):
            plan.replay(0)

        @staticmethod
        def _check_input(img):
            pass

    stream = HostStream(m, H, W, u8=True, engine=eng)
    outs = [stream.push(imgs[0, t], tris[0, t]) for t in range(5)]
    assert outs[0] is None and outs[1] is None
    plan = Plan()
    eng._rec = plan
    io = eng.eval_program(1, 3, H, W, 2, True)
    eng._rec = None
    for t in range(1, 4):
        io["imgs"].copy_(imgs[:, t - 1:t + 2]); io["tris"].copy_(tris[:, t - 1:t + 2])
        plan.replay(0)
        a, Fg, Bg = outs[t + 1]
        assert float((a - io["alphas"][0, 1]).abs().max()) < 1e-6
        assert float((Fg - io["Fs"][0, 1]).abs().max()) < 1e-6 and float((Bg - io["Bs"][0, 1]).abs().max()) < 1e-6
    stream.reset()
    assert stream.push(imgs[0, 0], tris[0, 0]) is None


def test_eval_program_five_frame_samples_match_oracle(emu):
    """S = 5 (three centre frames per sample, tail launched with ncen = 3) and B = 2 against the CPU oracle."""
    import tcvom_b200
    from oracle import vmn_fba_oracle as O
    from tcvom_b200 import synthetic
    from tcvom_b200.engine import Plan
    sd = fixture_sd_fba()
    net = tcvom_b200.get_VMN_models("vmn_fba", agg_window=7)
    net.load_state_dict(sd, strict=True)
    net.eval()
    eng = make_engine()
    eng.refresh_weights(net)
    B, S, H, W = 2, 5, 32, 32
    imgs, tris = synthetic.make_window(H, W, seed=21, frames=S, batch=B)
    imgs, tris = torch.from_numpy(imgs), torch.from_numpy(tris)
    plan = Plan()
    eng._rec = plan
    io = eng.eval_program(B, S, H, W, -1, False)
    eng._rec = None
    io["imgs"].copy_(imgs.float()); io["tris"].copy_(tris.float())
    plan.replay(0)
    with torch.no_grad():
        ra, rF, rB = O.eval_forward(sd, imgs.float(), tris.float())
    assert float((io["alphas"] - ra).abs().max()) < 1e-3
    assert float((io["Fs"] - rF).abs().max()) < 1e-3 and float((io["Bs"] - rB).abs().max()) < 1e-3
    assert float(io["alphas"][:, 0].abs().max()) == 0 and float(io["alphas"][:, -1].abs().max()) == 0
    assert float(io["alphas"][:, 1:4].abs().max()) > 0


def test_weight_cache_follows_load_state_dict(emu):
    """SURVEY 8b state / ownership: packed (weight-standardised) weights are derived caches, re-derived in place when
    the parameters change, so a RECORDED plan picks up a new checkpoint without being re-recorded."""
    import tcvom_b200
    from oracle import vmn_fba_oracle as O
    from tcvom_b200 import synthetic
    from tcvom_b200.engine import Plan
    sd = fixture_sd_fba()
    net = tcvom_b200.get_VMN_models("vmn_fba", agg_window=7)
    net.load_state_dict(sd, strict=True)
    net.eval()
    eng = make_engine()
    eng.refresh_weights(net)
    fp0 = eng._fingerprint
    eng.refresh_weights(net)
    assert eng._fingerprint == fp0                         # nothing changed: no re-packing
    H = W = 32
    imgs, tris = synthetic.make_window(H, W, seed=2)
    imgs, tris = torch.from_numpy(imgs), torch.from_numpy(tris)
    plan = Plan()
    eng._rec = plan
    io = eng.eval_program(1, 3, H, W, -1, True)
    eng._rec = None
    io["imgs"].copy_(imgs); io["tris"].copy_(tris)
    plan.replay(0)
    a0 = io["alphas"].clone()
    sd2 = {k: v.clone() for k, v in sd.items()}
    g = torch.Generator().manual_seed(9)
    for k in ("decoder.conv_up4.0.weight", "encoder.layer2.1.conv2.weight", "decoder.conv_up1.1.weight"):
        sd2[k] = sd2[k] * (1.0 + 0.2 * torch.randn(sd2[k].shape, generator=g))
    net.load_state_dict(sd2, strict=True)                  # in-place copy_: parameter versions change
    eng.refresh_weights(net)
    assert eng._fingerprint != fp0
    plan.replay(0)                                         # same recorded calls, same buffers
    with torch.no_grad():
        ra, _, _ = O.eval_forward(sd2, imgs.float(), tris.float())
    assert float((io["alphas"] - ra).abs().max()) < 1e-3
    assert float((io["alphas"] - a0).abs().max()) > 1e-3   # and the result did change


def test_trimap_transform_operator_matches_reference_semantics(emu):
    """tcvom_b200.trimap_transform (drop-in for utils/utils.py:25-39) against the oracle's restatement (scipy EDT, pinned
    to the reference's cv2 output by the fba goldens)."""
    from oracle import vmn_fba_oracle as O
    from tcvom_b200.model import _trimap_transform_impl
    rng = np.random.default_rng(3)
    t = np.zeros((2, 3, 2, 24, 40), np.float32)
    u = rng.uniform(size=(2, 3, 24, 40))
    t[:, :, 0] = u < 0.05
    t[:, :, 1] = u > 0.97
    t[1, 2] = 0                                              # a frame without any seed
    trimap = torch.from_numpy(t)
    got = _trimap_transform_impl(trimap, 0)
    ref = O.trimap_transform(trimap)
    assert got.shape == ref.shape == (2, 3, 6, 24, 40)
    assert float((got - ref).abs().max()) < 2e-5
    assert float(got[1, 2].abs().max()) == 0
