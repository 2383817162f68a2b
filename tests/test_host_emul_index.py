"""Host-logic tests of the IndexNet path in the GPU-less build container: ``IndexVmnEngine``'s program on host memory against
the test double of the C ABI (tests/host_emul/emul.cpp; the depthwise-conv / index kernels' bodies are the very same inline
functions the CUDA kernels wrap, tcvom_b200/csrc/fba_body.h) compared with the oracle and the reference golden vectors.
Test infrastructure only; the parity tests proper are the ``-m gpu`` ones."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import fixture_sd_index, golden
from test_host_emul_fba import emu, from_act, to_act  # noqa: F401  (fixture + layout helpers)


def make_engine(window=7):
    from tcvom_b200.index_engine import IndexVmnEngine

    class HostEmuEngine(IndexVmnEngine):
        @staticmethod
        def _check_device(dev):
            pass

        def _stream_ptr(self):
            return 0

    eng = HostEmuEngine(window)
    eng.device = torch.device("cpu")
    return eng


def _net_engine():
    import tcvom_b200
    net = tcvom_b200.get_VMN_models("vmn_index", agg_window=7)
    net.load_state_dict(fixture_sd_index(), strict=True)
    net.eval()
    eng = make_engine()
    eng.refresh_weights(net)
    return net, eng


def test_index_bodies_match_torch(emu):
    from oracle import vmn_index_oracle as O
    net, eng = _net_engine()
    sd = fixture_sd_index()
    torch.manual_seed(0)
    # index block + index pooling on 32 channels
    x = torch.randn(2, 32, 8, 12)
    a = to_act(x)
    xs = from_act(a)
    masked, pooled, idx_de = eng.index_pool(a, "encoder.index0")
    with torch.no_grad():
        ien, ide = O.index_block(xs, sd, "encoder.index0")
    assert float((from_act(idx_de) - ide).abs().max()) < 2e-4
    assert float((from_act(masked) - ien * xs).abs().max()) < 2e-4
    assert float((from_act(pooled) - 4 * F.avg_pool2d(ien * xs, 2, 2)).abs().max()) < 5e-4
    # inverted residual blocks: t = 1 (zero border) and t = 6 with a residual (border = relu6(BN shift)), padded 144 -> 160
    x = F.relu(torch.randn(2, 32, 6, 10))
    a = to_act(x)
    y = eng.inverted_residual(a, "encoder.layer1.0", 32, 16, 1)
    with torch.no_grad():
        ref = O.inverted_residual(from_act(a), sd, "encoder.layer1.0", 32, 16, 1)
    assert y.c == 32 and float((from_act(y, 16) - ref).abs().max()) < 3e-4 and float(from_act(y)[:, 16:].abs().max()) == 0
    x24 = torch.randn(2, 24, 6, 10)
    a = to_act(x24, 32)
    y = eng.inverted_residual(a, "encoder.layer2.1", 24, 24, 6)
    with torch.no_grad():
        ref = O.inverted_residual(from_act(a, 24), sd, "encoder.layer2.1", 24, 24, 6)
    assert float((from_act(y, 24) - ref).abs().max()) < 5e-4 and float(from_act(y)[:, 24:].abs().max()) == 0
    # ASPP
    x = F.relu(torch.randn(1, 320, 3, 4))
    a = to_act(x)
    y = eng.aspp(a, "encoder.dconv_pp")
    with torch.no_grad():
        ref = O.aspp(from_act(a), sd, "encoder.dconv_pp")
    assert float((from_act(y) - ref).abs().max()) < 5e-4


@pytest.mark.parametrize("name", ["ring64x96", "dil96x64", "batch2_64"])
def test_eval_program_matches_reference_golden(emu, name):
    from tcvom_b200.engine import Plan
    g = golden(f"index_{name}.npz")
    net, eng = _net_engine()
    imgs, tris = torch.from_numpy(g["imgs"]), torch.from_numpy(g["tris"])
    B, S, _, H, W = imgs.shape
    dil = int(g["dilate"])
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
    assert np.array_equal(out["small_mask"][:, 0].numpy().astype(bool), g["small_mask1"])
    err = lambda a, b: float(np.abs(a - b.astype(np.float32)).max())
    if "feat1" in g.files:
        assert err(from_act(out["feat"])[1:2].numpy(), g["feat1"]) < 2e-3
    assert err(out["pred"][:, 0].numpy(), g["pred1"]) < 1e-3
    assert err(alphas.numpy(), g["alphas"]) < 1e-3                  # north_star bar: 1e-3 on the alpha matte
    for k, ref in (("attb", g["attb1"]), ("attf", g["attf1"])):
        assert err(out[k][:, 0].numpy(), ref) < 3e-3 * max(1.0, float(np.abs(ref.astype(np.float32)).max()))
    assert float(alphas[:, 0].abs().max()) == 0 and float(alphas[:, -1].abs().max()) == 0
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
    net = tcvom_b200.get_VMN_models("vmn_index", agg_window=7)
    net.load_state_dict(fixture_sd_index(), strict=True)
    m = tcvom_b200.EvalModel(model="vmn_index", agg_window=7, dilate_kernel=2)
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
