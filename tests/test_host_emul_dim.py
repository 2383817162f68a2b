"""Host-logic tests of the DIM path in the GPU-less build container: ``DimVmnEngine``'s program on host memory against the
test double of the C ABI (tests/host_emul/emul.cpp; the pooling / unpooling / input-fix bodies are the very same inline
functions the CUDA kernels wrap, tcvom_b200/csrc/fba_body.h) compared with the oracle and the reference golden vectors.
Test infrastructure only; the parity tests proper are the ``-m gpu`` ones."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import fixture_sd_dim, golden
from test_host_emul_fba import emu, from_act, to_act  # noqa: F401  (fixture + layout helpers)


def make_engine(window=7):
    from tcvom_b200.dim_engine import DimVmnEngine

    class HostEmuEngine(DimVmnEngine):
        @staticmethod
        def _check_device(dev):
            pass

        def _stream_ptr(self):
            return 0

    eng = HostEmuEngine(window)
    eng.device = torch.device("cpu")
    return eng


def test_pool_unpool_bodies_match_torch(emu):
    eng = make_engine()
    torch.manual_seed(0)
    x = torch.randn(2, 16, 8, 12)
    x[:, :, 2:4, 4:8] = 0.0                      # ties: torch keeps the first position of the window
    x[0, 3, 0, 0] = x[0, 3, 1, 1] = 7.0
    x = F.relu(x)
    a = to_act(x)
    xs = from_act(a)                             # the values the kernels see (split-bf16 rounding)
    y, idx = eng.maxpool2(a)
    ref, ridx = F.max_pool2d(xs, (2, 2), 2, return_indices=True)
    assert torch.equal(from_act(y), ref)
    # torch's index is the flat position in the input plane; ours is ky*2 + kx inside the window
    oy = torch.arange(4).view(1, 1, 4, 1) * 2
    ox = torch.arange(6).view(1, 1, 1, 6) * 2
    ky, kx = ridx // 12 - oy, ridx % 12 - ox
    assert torch.equal(idx.permute(0, 3, 1, 2).long(), ky * 2 + kx)
    u = eng.unpool2(y, idx.data_ptr())
    assert torch.equal(from_act(u), F.max_unpool2d(ref, ridx, (2, 2), 2))


def test_conv_k_tap_groups_match_torch(emu):
    """5x5 and 7x7 convolutions as chains of <= 3x3 tap groups (partial sums through the residual input)."""
    eng = make_engine()
    torch.manual_seed(1)
    for k, cin, cout in ((5, 32, 64), (7, 32, 32)):
        net = torch.nn.Sequential()
        net.add_module("c", torch.nn.Conv2d(cin, cout, k, padding=k // 2))
        eng.w.clear(); eng.bias.clear(); eng._fingerprint = None; eng._tensors = None
        from tcvom_b200.engine import GcaVmnEngine
        GcaVmnEngine.refresh_weights(eng, net)
        x = torch.randn(2, cin, 9, 11)
        a = to_act(x)
        y = eng.conv_k(a, "c", act=1)
        ref = F.relu(net(from_act(a)))
        assert float((from_act(y) - ref).abs().max()) < 2e-4, k


@pytest.mark.parametrize("name", ["ring64x96", "dil96x64", "batch2_64"])
def test_eval_program_matches_reference_golden(emu, name):
    import tcvom_b200
    from tcvom_b200.engine import Plan
    g = golden(f"dim_{name}.npz")
    net = tcvom_b200.get_VMN_models("vmn_dim", agg_window=7)
    net.load_state_dict(fixture_sd_dim(), strict=True)
    net.eval()
    eng = make_engine()
    eng.refresh_weights(net)
    imgs, tris = torch.from_numpy(g["imgs"]), torch.from_numpy(g["tris"])
    B, S, _, H, W = imgs.shape
    dil = int(g["dilate"])
    # the calls EvalModel._plan records for method 'dim', on host buffers
    plan = Plan()
    eng._rec = plan
    x8 = eng._act(B * S, H, W, 8)
    trimask = eng._empty((B * S, H, W))
    tmp = eng._empty((2 * B * S * H * W,), torch.uint8)
    alphas = eng._empty((B, S, 1, H, W))
    im, tr = imgs.clone(), tris.clone()
    eng._call("tcv_preprocess_eval_u8", im.data_ptr(), tr.data_ptr(), B * S, H, W, dil, x8.ptr, trimask.data_ptr(),
              tmp.data_ptr())
    eng._call("tcv_dim_fix_inputs", tr.data_ptr(), 1, B * S, H, W, x8.ptr)
    out = eng.window_program(x8, trimask, B, S, H, W)
    eng._call("tcv_postprocess_eval_u8", out["pred"].data_ptr(), tr.data_ptr(), trimask.data_ptr(), B, S, H, W,
              alphas.data_ptr())
    eng._rec = None
    assert np.array_equal(trimask.reshape(B, S, 1, H, W).numpy().astype(np.uint8), g["trimask"])
    assert np.array_equal(out["small_mask"][:, 0].numpy().astype(bool), g["small_mask1"])
    err = lambda a, b: float(np.abs(a - b.astype(np.float32)).max())
    if "x4" in g.files:
        x4 = from_act(x8, 4).reshape(B, S, 4, H, W).numpy()
        assert err(x4, g["x4"]) < 3e-5                 # split-bf16 storage: 16 mantissa bits
        assert float(from_act(x8)[:, 4:].abs().max()) == 0
        feat = from_act(out["feat"])[1:2].numpy()
        assert err(feat, g["feat1"]) < 1e-3
    assert float(alphas[:, 0].abs().max()) == 0 and float(alphas[:, -1].abs().max()) == 0
    # Max-unpooling makes the matte discontinuous in the arg-max routing: against the reference's OWN routing (the golden
    # file) a handful of near-tie windows land on a neighbouring pixel and move the matte around them ...
    d = np.abs(out["pred"][:, 0].numpy() - g["pred1"])
    assert float(np.median(d)) < 1e-3 and float(d.max()) < 0.1       # (the 5x5 convs spread one re-routed value widely)
    # ... and with the oracle following the routing of the implementation under test the 1e-3 bar holds everywhere, while
    # every routing difference is a genuine near-tie (gap at rounding level)
    from oracle import vmn_dim_oracle as O
    idxs = out["pf"]["idxs"]
    force = [[O.window_idx_to_torch(t.reshape(B, S, *t.shape[1:])[:, i]) for t in idxs] for i in range(S)]
    ties = []
    ref_alphas, aux = O.eval_forward(fixture_sd_dim(), imgs.float(), tris.float(), None if dil < 0 else dil, 7, True,
                                     force_idx=force, ties=ties)
    assert max(ties) < 5e-5, ties                                    # split-bf16 storage resolves 2^-17 per value
    assert err(out["pred"][:, 0].numpy(), aux["preds"][1].numpy()) < 1e-3
    for k in ("attb", "attf"):
        assert err(out[k][:, 0].numpy(), aux[k][1].numpy()) < 1e-3
    assert err(alphas.numpy(), ref_alphas.numpy()) < 1e-3            # north_star bar: 1e-3 on the alpha matte
    # replaying the recorded calls reproduces the result bit for bit (what the CUDA graph does on the GPU)
    first = alphas.clone()
    alphas.zero_()
    plan.replay(0)
    assert torch.equal(alphas, first)


def test_frame_stream_equals_windowed_program(emu):
    """tcvom_b200.FrameStream (per-frame feature reuse across sliding windows) against the windowed program on
    every window of a 5-frame clip: same kernels on the same values."""
    import tcvom_b200
    from tcvom_b200 import synthetic
    from tcvom_b200.engine import Plan
    from tcvom_b200.stream import FrameStream
    net = tcvom_b200.get_VMN_models("vmn_dim", agg_window=7)
    net.load_state_dict(fixture_sd_dim(), strict=True)
    m = tcvom_b200.EvalModel(model="vmn_dim", agg_window=7, dilate_kernel=2)
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
    x8 = eng._act(3, H, W, 8)
    trimask = eng._empty((3, H, W))
    tmp = eng._empty((2 * 3 * H * W,), torch.uint8)
    alphas = eng._empty((1, 3, 1, H, W))
    im, tr = imgs[:, :3].clone(), tris[:, :3].clone()
    eng._call("tcv_preprocess_eval_u8", im.data_ptr(), tr.data_ptr(), 3, H, W, 2, x8.ptr, trimask.data_ptr(), tmp.data_ptr())
    eng._call("tcv_dim_fix_inputs", tr.data_ptr(), 1, 3, H, W, x8.ptr)
    out = eng.window_program(x8, trimask, 1, 3, H, W)
    eng._call("tcv_postprocess_eval_u8", out["pred"].data_ptr(), tr.data_ptr(), trimask.data_ptr(), 1, 3, H, W,
              alphas.data_ptr())
    eng._rec = None
    for t in range(1, 4):
        im.copy_(imgs[:, t - 1:t + 2]); tr.copy_(tris[:, t - 1:t + 2])
        plan.replay(0)
        assert float((outs[t + 1] - alphas[0, 1]).abs().max()) < 1e-6
