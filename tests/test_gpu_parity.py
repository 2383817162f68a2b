"""GPU parity tests: the sm_100a path (through the C ABI) against the golden vectors produced by
the unmodified reference and against the CPU oracle on the same seeded inputs.

Tolerance: north_star states 1e-3 max abs on the alpha matte (fp32 reference)."""
import numpy as np
import pytest
import torch

from helpers import fixture_sd, golden, op_inputs
from oracle import vmn_gca_oracle as O

pytestmark = pytest.mark.gpu
ALPHA_TOL = 1e-3
CASES = ["ring64", "ring96x128", "allunk64", "nounk64", "dil64x96", "batch2_64"]


def _model(dilate=None):
    import tcvom_b200
    m = tcvom_b200.EvalModel(model="vmn_gca", agg_window=7, dilate_kernel=dilate)
    m.NET.load_state_dict(fixture_sd(), strict=True)
    return m.cuda().eval()


def _prefixed(prefix):
    return {k[len(prefix):]: v for k, v in fixture_sd().items() if k.startswith(prefix)}


@pytest.mark.parametrize("case", CASES)
def test_eval_forward_matches_reference_golden(case):
    g = golden(f"eval_{case}.npz")
    dil = int(g["dilate"])
    m = _model(None if dil < 0 else dil)
    imgs = torch.from_numpy(g["imgs"]).float().cuda()
    tris = torch.from_numpy(g["tris"]).float().cuda()
    with torch.no_grad():
        alphas = m(imgs, tris)
    plan = list(m.NET.engine().plans.values())[0]
    assert np.array_equal(plan.io["trimask"].reshape(g["trimask"].shape).cpu().numpy().astype(np.uint8), g["trimask"])
    err = np.abs(alphas.cpu().numpy() - g["alphas"]).max()
    print(case, "alpha max abs err", err)
    assert err < ALPHA_TOL
    pred = plan.io["pred"][:, 0].cpu().numpy()
    assert np.abs(pred - g["pred1"]).max() < ALPHA_TOL
    assert np.array_equal(plan.io["small_mask"][:, 0].bool().cpu().numpy(), g["small_mask1"])
    for mine, ref in ((plan.io["attb"][:, 0], g["attb1"]), (plan.io["attf"][:, 0], g["attf1"])):
        assert np.abs(mine.cpu().numpy() - ref).max() <= 2e-3 * max(1.0, np.abs(ref).max())


def test_eval_forward_second_call_and_graph_replay():
    g = golden("eval_ring64.npz")
    m = _model()
    imgs = torch.from_numpy(g["imgs"]).float().cuda()
    tris = torch.from_numpy(g["tris"]).float().cuda()
    with torch.no_grad():
        a1 = m(imgs, tris).clone()
        a2 = m(imgs, tris).clone()         # CUDA-graph replay of the recorded plan
        a3 = m(imgs.flip(-1).contiguous(), tris.flip(-1).contiguous())
        a4 = m(imgs, tris)
    assert torch.equal(a1, a2) and torch.equal(a1, a4)
    assert not torch.equal(a1, a3)
    assert np.abs(a1.cpu().numpy() - g["alphas"]).max() < ALPHA_TOL


def test_vmn_seam_matches_oracle():
    """Plugin seam models.VMN.get_VMN_models(...).forward(images, masks)."""
    import tcvom_b200
    g = golden("eval_batch2_64.npz")
    sd = fixture_sd()
    imgs = torch.from_numpy(g["imgs"]).float()
    tris = torch.from_numpy(g["tris"]).float()
    x6, trimask = O.eval_preprocess(imgs, tris)
    S = imgs.shape[1]
    ref_preds, ref_attb, ref_attf, ref_small, _ = O.vmn_forward(
        sd, [x6[:, i] for i in range(S)], [trimask[:, i] for i in range(S)], 7)
    net = tcvom_b200.get_VMN_models("vmn_gca", agg_window=7)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    images = list(x6.cuda().split(1, dim=1))
    masks = trimask.cuda().split(1, dim=1)
    with torch.no_grad():
        preds, attb, attf, small = net(images, masks)
    assert preds[0].abs().max() == 0 and preds[-1].abs().max() == 0 and attb[0] is None
    assert (preds[1].cpu() - ref_preds[1]).abs().max() < ALPHA_TOL
    assert torch.equal(small[1].cpu(), ref_small[1])
    assert (attb[1].cpu() - ref_attb[1]).abs().max() <= 2e-3 * max(1.0, float(ref_attb[1].abs().max()))


def test_tam_operator_matches_reference_golden():
    import tcvom_b200
    g = golden("op_tam.npz")
    fam = tcvom_b200.FeatureAggregationModule(128, 1, 7)
    fam.load_state_dict(_prefixed("decoder.fam."), strict=True)
    fam = fam.cuda().eval()
    x = torch.from_numpy(op_inputs("tam_x", (2, 128, 12, 16))).cuda()
    b = torch.from_numpy(op_inputs("tam_b", (2, 128, 12, 16))).cuda()
    f = torch.from_numpy(op_inputs("tam_f", (2, 128, 12, 16))).cuda()
    mask = torch.from_numpy((op_inputs("tam_m", (2, 1, 96, 128)) > 0.5).astype(np.float32)).cuda()
    with torch.no_grad():
        feat, attb, attf, sm = fam(x, b, f, mask)
    assert np.array_equal(sm.cpu().numpy(), g["small_mask"])
    # inputs are rounded to split-bf16 (16 mantissa bits) at the seam: tolerance is relative
    for mine, ref in ((feat, g["feat"]), (attb, g["attb"]), (attf, g["attf"])):
        assert np.abs(mine.cpu().numpy() - ref).max() <= 1e-3 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("chn", [32, 256])
def test_tam_operator_other_widths_match_oracle(chn):
    """FeatureAggregationModule(32 / 256, ...) -- the TAM widths of the IndexNet and the DIM / FBA base networks
    (VMN_Index.py:10, VMN_DIM.py:99, VMN_FBA.py:9) -- against the oracle's dense-then-mask restatement."""
    import tcvom_b200
    from oracle import vmn_gca_oracle as O
    torch.manual_seed(chn)
    fam = tcvom_b200.FeatureAggregationModule(chn, 1, 7)
    sd = {k: torch.randn_like(v) * (0.3 / (chn * 9) ** 0.5 if k.endswith("weight") else 0.1) for k, v in fam.state_dict().items()}
    fam.load_state_dict(sd, strict=True)
    fam = fam.cuda().eval()
    x, b, f = (torch.randn(2, chn, 10, 14) for _ in range(3))
    mask = (torch.rand(2, 1, 80, 112) > 0.4).float()
    with torch.no_grad():
        feat, attb, attf, sm = fam(x.cuda(), b.cuda(), f.cuda(), mask.cuda())
        rfeat, rb, rf, rm = O.tam({"fam." + k: v for k, v in sd.items()}, "fam", x, b, f, mask, 7)
    assert torch.equal(sm.cpu(), rm)
    for mine, ref in ((feat, rfeat), (attb, rb), (attf, rf)):
        assert float((mine.cpu() - ref).abs().max()) <= 1e-3 * max(1.0, float(ref.abs().max()))


def test_gca_operator_matches_reference_golden():
    import tcvom_b200
    g = golden("op_gca.npz")
    gca = tcvom_b200.GuidedCxtAtten(128, 128)
    gca.load_state_dict(_prefixed("decoder.gca."), strict=True)
    gca = gca.cuda().eval()
    f = torch.from_numpy(op_inputs("gca_f", (2, 128, 16, 24))).cuda()
    al = torch.from_numpy(op_inputs("gca_alpha", (2, 128, 16, 24))).cuda()
    unk = torch.from_numpy((op_inputs("gca_unk", (2, 1, 16, 24)) > 0.3).astype(np.float32)).cuda()
    with torch.no_grad():
        y, (offsets, scale) = gca(f, al, unk)
    assert np.abs(scale.cpu().numpy() - g["scale"]).max() < 1e-5
    assert np.abs(y.cpu().numpy() - g["y"]).max() <= 1e-3 * max(1.0, np.abs(g["y"]).max())


def test_eval_forward_256_matches_oracle():
    """BASELINE config 1 shape (256x256 window) against the CPU oracle on the same seeded input."""
    from tcvom_b200 import synthetic
    imgs, tris = synthetic.make_window(256, 256, seed=7)
    ti, tt = torch.from_numpy(imgs).float(), torch.from_numpy(tris).float()
    ref = O.eval_forward(fixture_sd(), ti, tt)
    m = _model()
    with torch.no_grad():
        out = m(ti.cuda(), tt.cuda())
    err = (out.cpu() - ref).abs().max().item()
    print("256x256 alpha max abs err", err)
    assert err < ALPHA_TOL


def test_no_fallback_on_cpu_tensor():
    m = _model()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 3, 64, 64), torch.zeros(1, 3, 1, 64, 64))


def test_eval_forward_1080p_matches_oracle():
    """BASELINE configs[1] at its full size (1088x1920 window): direct comparison with the CPU oracle
    (a few seconds of host time), plus size-independent properties of the output."""
    from tcvom_b200 import synthetic
    imgs, tris = synthetic.make_window(1088, 1920, seed=7)
    ti, tt = torch.from_numpy(imgs).float(), torch.from_numpy(tris).float()
    m = _model()
    with torch.no_grad():
        out = m(ti.cuda(), tt.cuda()).cpu()
    # properties: end frames are zero; known-region pixels reproduce the trimap exactly
    assert out[:, 0].abs().max() == 0 and out[:, 2].abs().max() == 0
    known = (tt[:, 1] == 0) | (tt[:, 1] == 255)
    assert torch.equal(out[:, 1][known], (tt[:, 1] * (1.0 / 255))[known])
    assert out.min() >= 0 and out.max() <= 1
    ref = O.eval_forward(fixture_sd(), ti, tt)
    err = (out - ref).abs().max().item()
    print("1088x1920 alpha max abs err", err)
    assert err < ALPHA_TOL


def test_full_vmd_forward_matches_reference_golden():
    """FullModel_VMD eval-mode forward (pred_vmn.py path), S=5: losses L_im / L_tc / L_af + visual outputs."""
    import tcvom_b200
    g = golden("train_s5.npz")
    m = tcvom_b200.FullModel_VMD(model="vmn_gca", agg_window=7, dilate_kernel=3)
    m.NET.load_state_dict(fixture_sd(), strict=True)
    m = m.cuda().eval()
    a, fg, bg = (torch.from_numpy(g[k]).float().cuda() for k in ("a", "fg", "bg"))
    with torch.no_grad():
        out = m(a, fg, bg)
    assert len(out) == 12
    losses = np.array([float(o) for o in out[:5]])
    print("losses", losses, "ref", g["losses"])
    # L_alpha / L_dt are means of |alpha - gt| (alpha within 1e-3); L_att is a BCE over logits of magnitude ~1e2
    assert abs(losses[0] - g["losses"][0]) < 1e-3 and abs(losses[3] - g["losses"][3]) < 1e-3
    assert losses[1] == 0 and losses[2] == 0
    assert abs(losses[4] - g["losses"][4]) <= 2e-3 * abs(g["losses"][4])
    assert np.abs(out[5].cpu().numpy() - g["scaled_imgs"]).max() < 1e-6
    assert np.abs(out[6].cpu().numpy() - g["tris_vis"]).max() < 1e-6
    assert np.abs(out[7].cpu().numpy() - g["alphas"]).max() < ALPHA_TOL
    assert np.abs(out[8].cpu().numpy() - g["comps"]).max() < ALPHA_TOL


def test_full_vmd_s3_and_random_dilation_against_oracle():
    """S=3 (L_tc == 0, model.py:344-345) and the per-sample random trimap width (model.py:62)."""
    import tcvom_b200
    g = golden("train_s5.npz")
    a, fg, bg = (torch.from_numpy(g[k][:, :3]).float() for k in ("a", "fg", "bg"))
    a2 = torch.cat([a, a.flip(-1)], 0); fg2 = torch.cat([fg, fg.flip(-1)], 0); bg2 = torch.cat([bg, bg.flip(-1)], 0)
    m = tcvom_b200.FullModel_VMD(model="vmn_gca", agg_window=7, dilate_kernel=None)
    m.NET.load_state_dict(fixture_sd(), strict=True)
    m = m.cuda().eval()
    torch.manual_seed(5)
    radii = [int(torch.randint(0, 26, size=())) for _ in range(2)]
    torch.manual_seed(5)
    with torch.no_grad():
        out = m(a2.cuda(), fg2.cuda(), bg2.cuda())
    ref = O.full_vmd_forward(fixture_sd(), a2, fg2, bg2, radii)
    assert float(out[3]) == 0.0 and float(ref[3]) == 0.0
    assert abs(float(out[0]) - float(ref[0])) < 1e-3
    assert abs(float(out[4]) - float(ref[4])) <= 2e-3 * abs(float(ref[4])) + 1e-4
    assert (out[6].cpu() - ref[6]).abs().max() < 1e-6          # tris_vis: same dilation radii were drawn
    assert (out[7].cpu() - ref[7]).abs().max() < ALPHA_TOL


def test_uint8_ingest_equals_float_ingest():
    """uint8 frames/trimaps (cv2.imread dtype) give bit-identical alphas to the float path."""
    g = golden("eval_dil64x96.npz")
    m = _model(int(g["dilate"]))
    iu, tu = torch.from_numpy(g["imgs"]).cuda(), torch.from_numpy(g["tris"]).cuda()
    with torch.no_grad():
        a_u8 = m(iu, tu).clone()
        a_f = m(iu.float(), tu.float())
    assert torch.equal(a_u8, a_f)
    assert np.abs(a_u8.cpu().numpy() - g["alphas"]).max() < ALPHA_TOL


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_data_parallel_wrapper_matches_single_gpu():
    """pred_vmn.py:85 wraps the model in nn.DataParallel (batch = #GPUs)."""
    g = golden("eval_batch2_64.npz")
    m = _model()
    imgs = torch.from_numpy(g["imgs"]).float().cuda()
    tris = torch.from_numpy(g["tris"]).float().cuda()
    with torch.no_grad():
        single = m(imgs, tris).clone()
        dp = torch.nn.DataParallel(m)
        out1 = dp(imgs, tris).clone()
        out2 = dp(imgs, tris)            # second call: replicas are rebuilt, plans/graphs are reused per device
    assert torch.equal(out1, out2)
    assert (out1 - single).abs().max() < 1e-6
    assert np.abs(out1.cpu().numpy() - g["alphas"]).max() < ALPHA_TOL


def test_plan_cache_is_bounded_and_shapes_can_alternate():
    """Plans own GBs of buffers: only the most recent shapes are kept, evicted shapes are re-recorded."""
    from tcvom_b200 import synthetic
    m = _model()
    outs = {}
    for hw in ((64, 64), (64, 96), (96, 64), (64, 64)):
        imgs, tris = synthetic.make_window(*hw, seed=3)
        with torch.no_grad():
            a = m(torch.from_numpy(imgs).cuda(), torch.from_numpy(tris).cuda()).clone()
        if hw in outs:
            assert torch.equal(outs[hw], a)
        outs[hw] = a
        assert len(m.NET.engine().plans) <= 2


def test_shape_validation():
    m = _model()
    with pytest.raises(ValueError):
        m(torch.zeros(1, 3, 3, 70, 64).cuda(), torch.zeros(1, 3, 1, 70, 64).cuda())
