"""Host-logic tests of the vmn_gca inference path in the GPU-less build container: the program of GcaVmnEngine (weight
folding / packing, which C-ABI calls run on which buffers, the recorded plan) is executed on host memory against the
test double of the C ABI (tests/host_emul/emul.cpp: naive restatements of every entry point the eval path uses, incl. the
bf16 operand planes of the attention GEMMs) and compared with the golden vectors produced by the unmodified reference.
Test infrastructure only; the real parity tests are the ``-m gpu`` ones."""
import numpy as np
import pytest
import torch

from helpers import fixture_sd, golden
from test_host_emul_fba import emu  # noqa: F401  (fixture)

CASES = ["ring64", "ring96x128", "allunk64", "nounk64", "dil64x96", "batch2_64"]


def make_gca_engine(window=7):
    from tcvom_b200.engine import GcaVmnEngine

    class HostEmuEngine(GcaVmnEngine):
        """GcaVmnEngine on host memory: no device check, null stream, fp32 attention operands."""

        @staticmethod
        def _check_device(dev):
            pass

        def _stream_ptr(self):
            return 0

    eng = HostEmuEngine(window)
    eng.device = torch.device("cpu")
    return eng                         # default program: bf16x3 operand planes for both attention GEMMs (use_tc_attn)


def record_eval(eng, B, S, H, W, dil, u8):
    """The three steps of tcvom_b200.EvalModel._plan (preprocess -> window_program -> postprocess) on static buffers."""
    from tcvom_b200.engine import Plan
    plan = Plan()
    eng._rec = plan
    try:
        in_dt = torch.uint8 if u8 else torch.float32
        sfx = "_u8" if u8 else ""
        imgs = eng._empty((B, S, 3, H, W), in_dt)
        tris = eng._empty((B, S, 1, H, W), in_dt)
        x8 = eng._act(B * S, H, W, 8)
        trimask = eng._empty((B * S, H, W))
        tmp = eng._empty((2 * B * S * H * W,), torch.uint8)
        alphas = eng._empty((B, S, 1, H, W))
        imgs.zero_(); tris.zero_()
        eng._call("tcv_preprocess_eval" + sfx, imgs.data_ptr(), tris.data_ptr(), B * S, H, W, dil, x8.ptr,
                  trimask.data_ptr(), tmp.data_ptr())
        out = eng.window_program(x8, trimask, B, S, H, W)
        eng._call("tcv_postprocess_eval" + sfx, out["pred"].data_ptr(), tris.data_ptr(), trimask.data_ptr(), B, S, H, W,
                  alphas.data_ptr())
    finally:
        eng._rec = None
    return plan, dict(imgs=imgs, tris=tris, alphas=alphas, trimask=trimask, **out)


def _net():
    import tcvom_b200
    net = tcvom_b200.get_VMN_models("vmn_gca", agg_window=7)
    net.load_state_dict(fixture_sd(), strict=True)
    return net.eval()


@pytest.mark.parametrize("case", CASES)
def test_window_program_matches_reference_golden(emu, case):
    g = golden(f"eval_{case}.npz")
    eng = make_gca_engine()
    eng.refresh_weights(_net())
    imgs, tris = torch.from_numpy(g["imgs"]), torch.from_numpy(g["tris"])
    B, S, _, H, W = imgs.shape
    plan, io = record_eval(eng, B, S, H, W, int(g["dilate"]), True)
    io["imgs"].copy_(imgs); io["tris"].copy_(tris)
    plan.replay(0)
    assert np.array_equal(io["trimask"].reshape(g["trimask"].shape).numpy().astype(np.uint8), g["trimask"])
    assert np.abs(io["alphas"].numpy() - g["alphas"]).max() < 1e-3
    assert np.abs(io["pred"][:, 0].numpy() - g["pred1"]).max() < 1e-3
    assert np.array_equal(io["small_mask"][:, 0].numpy().astype(bool), g["small_mask1"])
    for mine, ref in ((io["attb"][:, 0], g["attb1"]), (io["attf"][:, 0], g["attf1"])):
        assert np.abs(mine.numpy() - ref).max() <= 2e-3 * max(1.0, np.abs(ref).max())
    assert float(io["alphas"][:, 0].abs().max()) == 0 and float(io["alphas"][:, -1].abs().max()) == 0


@pytest.mark.parametrize("variant", ["fp32_attn", "six_term_scores", "bf16_pv"])
def test_attention_program_variants(emu, variant):
    """The other operand formats of the attention program: exact fp32 GEMMs (TCV_TC_ATTN=0), three-plane scores
    (TCV_SCORE_PLANES=3), single-plane bf16 aggregation (TCV_PV_MODE=bf16; known to miss 1e-3: only its wiring)."""
    g = golden("eval_ring64.npz")
    eng = make_gca_engine()
    if variant == "fp32_attn":
        eng.use_tc_attn = False
    elif variant == "six_term_scores":
        eng.score_planes = 3
    else:
        eng.pv_mode = "bf16"
    eng.refresh_weights(_net())
    imgs, tris = torch.from_numpy(g["imgs"]), torch.from_numpy(g["tris"])
    plan, io = record_eval(eng, 1, 3, 64, 64, -1, True)
    io["imgs"].copy_(imgs); io["tris"].copy_(tris)
    plan.replay(0)
    assert np.abs(io["alphas"].numpy() - g["alphas"]).max() < (2e-2 if variant == "bf16_pv" else 1e-3)


def test_cuda_core_conv_program_matches_too(emu):
    """TCV_TC_CONV=0 program variant (reflect-padded stride-2 guidance convs and the 1-channel head through
    tcv_conv2d itself instead of pad_reflect1 / head32)."""
    g = golden("eval_ring64.npz")
    eng = make_gca_engine()
    eng.use_tc_conv = False
    eng.refresh_weights(_net())
    imgs, tris = torch.from_numpy(g["imgs"]).float(), torch.from_numpy(g["tris"]).float()
    plan, io = record_eval(eng, 1, 3, 64, 64, -1, False)
    io["imgs"].copy_(imgs); io["tris"].copy_(tris)
    plan.replay(0)
    assert np.abs(io["alphas"].numpy() - g["alphas"]).max() < 1e-3


def test_frame_stream_gca_equals_windowed_program(emu):
    """tcvom_b200.FrameStream on the vmn_gca engine: streamed mattes == windowed program on every window of a clip."""
    import tcvom_b200
    from tcvom_b200 import synthetic
    from tcvom_b200.stream import FrameStream
    net = _net()
    m = tcvom_b200.EvalModel(model="vmn_gca", agg_window=7, dilate_kernel=None)
    m.NET = net
    m.eval()
    eng = make_gca_engine()
    eng.refresh_weights(net)
    H, W = 64, 64
    imgs, tris = synthetic.make_window(H, W, seed=6, frames=5)
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
    plan, io = record_eval(eng, 1, 3, H, W, -1, True)
    for t in range(1, 4):
        io["imgs"].copy_(imgs[:, t - 1:t + 2]); io["tris"].copy_(tris[:, t - 1:t + 2])
        plan.replay(0)
        assert float((outs[t + 1] - io["alphas"][0, 1]).abs().max()) < 1e-6


@pytest.mark.parametrize("case", ["ring64", "dil64x96"])
def test_space_to_depth_stride2_program(emu, case):
    """Opt-in TCV_S2D_STRIDE2 program: encoder.conv1 and guidance_head.1 / .5 (3x3, stride 2; zero resp. reflect padding)
    as 2x2-tap stride-1 convolutions over the 2x2 space-to-depth image == the default program, and the goldens."""
    g = golden(f"eval_{case}.npz")
    imgs, tris = torch.from_numpy(g["imgs"]), torch.from_numpy(g["tris"])
    B, S, _, H, W = imgs.shape
    res = {}
    for s2d in (False, True):
        eng = make_gca_engine()
        eng.s2d_stride2 = eng.s2d_guidance = s2d            # all three layers (the default program converts conv1 only)
        eng.refresh_weights(_net())
        plan, io = record_eval(eng, B, S, H, W, int(g["dilate"]), True)
        io["imgs"].copy_(imgs); io["tris"].copy_(tris)
        plan.replay(0)
        res[s2d] = io["alphas"].clone()
        kinds = [m["kind"] for m in plan.meta]
        assert ("tcv_space_to_depth2" in kinds) == s2d
    assert np.abs(res[True].numpy() - g["alphas"]).max() < 1e-3
    assert float((res[True] - res[False]).abs().max()) < 1e-4


@pytest.mark.parametrize("case", ["ring64", "ring96x128", "allunk64"])
def test_shift_sum_aggregation_equals_fold_program(emu, case):
    """Default program (aggregation as A2.F on the padded grid, csrc/gca.cu) == the unfold-values + overlap-add program
    (TCV_GCA_SHIFT_SUM=0) and the reference goldens: GuidedCxtAtten's fold(A.V)/4 (GCA/ops.py:112-118,204) re-associated."""
    g = golden(f"eval_{case}.npz")
    imgs, tris = torch.from_numpy(g["imgs"]), torch.from_numpy(g["tris"])
    B, S, _, H, W = imgs.shape
    res = {}
    for ss in (False, True, "consumer"):
        eng = make_gca_engine()
        eng.gca_shift_sum = bool(ss)
        eng.gca_softmax_in_consumer = ss == "consumer"      # exponentials in the shift kernel / a normalise pass (default)
        eng.refresh_weights(_net())
        plan, io = record_eval(eng, B, S, H, W, int(g["dilate"]), True)
        io["imgs"].copy_(imgs); io["tris"].copy_(tris)
        plan.replay(0)
        res[ss] = io["alphas"].clone()
        kinds = [m["kind"] for m in plan.meta]
        eng_default = make_gca_engine()
        assert eng_default.gca_shift_sum and not eng_default.gca_softmax_in_consumer  # what ships (measured faster)
        assert ("tcv_gca_shift_add" in kinds) == (ss is True) and ("tcv_gca_softmax_shift" in kinds) == (ss == "consumer")
        assert ("tcv_gca_fold" in kinds) == (not ss)
    assert np.abs(res[True].numpy() - g["alphas"]).max() < 1e-3
    assert float((res[True] - res[False]).abs().max()) < 1e-4
    assert float((res[True] - res["consumer"]).abs().max()) < 5e-5


def test_recorded_pointers_survive_replaced_parameters(emu):
    """Plans / CUDA graphs bake in device pointers: conv biases must be engine-owned copies updated in place, because a
    module parameter's storage is not stable (nn.DataParallel hands every forward freshly broadcast replica tensors).
    A plan recorded against net A must give net B's result after refresh_weights(net B) -- without re-recording."""
    g = golden("eval_ring64.npz")
    imgs, tris = torch.from_numpy(g["imgs"]), torch.from_numpy(g["tris"])
    B, S, _, H, W = imgs.shape
    eng = make_gca_engine()
    net_a = _net()
    with torch.no_grad():
        for k, v in net_a.state_dict().items():
            if k.endswith(".bias") and ("fam" in k or "guidance_conv" in k or k == "decoder.conv2.bias"):
                v.add_(0.3)                                     # net A: different conv biases than the fixture
    eng.refresh_weights(net_a)
    ptrs = {k: v.data_ptr() for k, v in eng.bias.items()}
    plan, io = record_eval(eng, B, S, H, W, int(g["dilate"]), True)
    io["imgs"].copy_(imgs); io["tris"].copy_(tris)
    plan.replay(0)
    assert np.abs(io["alphas"].numpy() - g["alphas"]).max() > 1e-3      # the biases matter
    net_b = _net()                                                      # fresh tensors, fixture values
    del net_a
    eng.refresh_weights(net_b)
    assert {k: v.data_ptr() for k, v in eng.bias.items()} == ptrs
    assert all(v.data_ptr() not in {t.data_ptr() for t in net_b.state_dict().values()} for v in eng.bias.values())
    plan.replay(0)                                                      # the SAME recorded plan
    assert np.abs(io["alphas"].numpy() - g["alphas"]).max() < 1e-3
