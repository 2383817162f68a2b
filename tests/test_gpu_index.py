"""GPU parity tests of the IndexNet+TAM path (``vmn_index``) through the C ABI:
``tcvom_b200.EvalModel('vmn_index')`` / ``get_VMN_models('vmn_index')`` against the golden vectors of the unmodified reference
and the CPU oracle (oracle/vmn_index_oracle.py, pinned to the same goldens by tests/test_oracle_index.py)."""
import numpy as np
import pytest
import torch

from helpers import fixture_sd_index, golden

pytestmark = pytest.mark.gpu
ALPHA_TOL = 1e-3          # north_star bar on the alpha matte


@pytest.fixture(scope="module")
def model():
    import tcvom_b200
    m = tcvom_b200.EvalModel(model="vmn_index", agg_window=7, dilate_kernel=None)
    m.NET.load_state_dict(fixture_sd_index(), strict=True)
    return m.cuda().eval()


@pytest.mark.parametrize("case", ["ring64x96", "dil96x64", "batch2_64"])
def test_eval_forward_matches_reference_golden(model, case):
    g = golden(f"index_{case}.npz")
    dil = int(g["dilate"])
    model.DILATION_KERNEL = None if dil < 0 else dil
    imgs, tris = torch.from_numpy(g["imgs"]).cuda(), torch.from_numpy(g["tris"]).cuda()      # uint8 ingest
    with torch.no_grad():
        alphas = model(imgs, tris)
    B, S, _, H, W = imgs.shape
    plan = model._plan(B, S, H, W, imgs.device, True)
    err = lambda a, b: float(np.abs(a.cpu().numpy() - b.astype(np.float32)).max())
    ea = err(alphas, g["alphas"])
    print(case, "alpha max abs err", ea)
    assert ea < ALPHA_TOL
    assert err(plan.io["pred"][:, 0], g["pred1"]) < ALPHA_TOL
    assert np.array_equal(plan.io["small_mask"][:, 0].bool().cpu().numpy(), g["small_mask1"])
    for k, ref in (("attb", g["attb1"]), ("attf", g["attf1"])):
        assert err(plan.io[k][:, 0], ref) <= 3e-3 * max(1.0, float(np.abs(ref.astype(np.float32)).max()))
    assert float(alphas[:, 0].abs().max()) == 0 and float(alphas[:, -1].abs().max()) == 0
    model.DILATION_KERNEL = None


def test_five_frame_window_matches_oracle(model):
    from oracle import vmn_index_oracle as O
    from tcvom_b200 import synthetic
    imgs, tris = synthetic.make_window(128, 192, seed=21, frames=5)
    imgs, tris = torch.from_numpy(imgs).float(), torch.from_numpy(tris).float()            # float ingest
    with torch.no_grad():
        alphas = model(imgs.cuda(), tris.cuda())
    ref = O.eval_forward(fixture_sd_index(), imgs, tris, None, 7)
    err = float((alphas.cpu() - ref).abs().max())
    print("5-frame 128x192 alpha max abs err", err)
    assert err < ALPHA_TOL and float(ref[:, 1:4].std()) > 0.03


def test_vmn_seam_equals_eval_model(model):
    """get_VMN_models('vmn_index')(images, masks) -- the plugin seam (models/VMN/__init__.py:22-24)."""
    from oracle import vmn_index_oracle as O
    from tcvom_b200 import synthetic
    imgs, tris = synthetic.make_window(64, 96, seed=3)
    imgs, tris = torch.from_numpy(imgs), torch.from_numpy(tris)
    with torch.no_grad():
        model(imgs.cuda(), tris.cuda())
        plan = model._plan(1, 3, 64, 96, torch.device("cuda:0"), True)
        want = plan.io["pred"][:, 0].clone()
        x4, trimask = O.eval_preprocess(imgs.float(), tris.float())
        images = [x4[:, i:i + 1].cuda() for i in range(3)]
        masks = [trimask[:, i:i + 1].cuda() for i in range(3)]
        preds, attb, attf, small = model.NET(images, masks)
    assert len(preds) == 3 and float(preds[0].abs().max()) == 0 and float(preds[2].abs().max()) == 0
    assert float((preds[1] - want).abs().max()) < 1e-4
    assert attb[0] is None and attb[1].shape == (1, 49, 8 * 12) and small[1].dtype == torch.bool


def test_full_hd_window_runs(model):
    """One 1088x1920 window: finite matte, zero end frames, trimap passthrough outside the unknown band."""
    import time
    from tcvom_b200 import synthetic
    imgs, tris = synthetic.make_window(1088, 1920, seed=7)
    ti, tt = torch.from_numpy(imgs).cuda(), torch.from_numpy(tris).cuda()
    with torch.no_grad():
        a = model(ti, tt)
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(3):
            a = model(ti, tt)
        torch.cuda.synchronize()
    print(f"vmn_index 1088x1920 window: {(time.time() - t0) / 3 * 1e3:.1f} ms")
    assert torch.isfinite(a).all()
    assert float(a[:, 0].abs().max()) == 0 and float(a[:, 2].abs().max()) == 0
    known = tt[:, 1] != 128
    assert torch.equal(a[:, 1][known], tt[:, 1][known].float() * (1.0 / 255))
    model.NET.engine().plans.clear()
    torch.cuda.empty_cache()
