"""Pins oracle/vmn_gca_oracle.py against outputs of the unmodified reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from helpers import fixture_sd, golden, op_inputs
from oracle import vmn_gca_oracle as O

CASES = ["ring64", "ring96x128", "allunk64", "nounk64", "dil64x96", "batch2_64"]


@pytest.mark.parametrize("case", CASES)
def test_eval_forward_matches_reference(case):
    g = golden(f"eval_{case}.npz")
    sd = fixture_sd()
    dil = int(g["dilate"])
    imgs = torch.from_numpy(g["imgs"]).float()
    tris = torch.from_numpy(g["tris"]).float()
    alphas, aux = O.eval_forward(sd, imgs, tris, None if dil < 0 else dil, 7, return_aux=True)
    _, trimask = O.eval_preprocess(imgs, tris, None if dil < 0 else dil)
    assert np.array_equal(trimask.numpy().astype(np.uint8), g["trimask"])
    assert np.array_equal(aux["small_mask"][1].numpy(), g["small_mask1"])
    # oracle and reference are both torch-CPU fp32; differences are re-association only
    assert np.abs(alphas.numpy() - g["alphas"]).max() < 2e-5
    assert np.abs(aux["preds"][1].numpy() - g["pred1"]).max() < 2e-5
    for a, b in ((aux["attb"][1], g["attb1"]), (aux["attf"][1], g["attf1"])):
        assert np.abs(a.numpy() - b).max() <= 1e-4 * max(1.0, np.abs(b).max())


def test_gca_operator_matches_reference():
    g = golden("op_gca.npz")
    sd = fixture_sd()
    f = torch.from_numpy(op_inputs("gca_f", (2, 128, 16, 24)))
    al = torch.from_numpy(op_inputs("gca_alpha", (2, 128, 16, 24)))
    unk = torch.from_numpy((op_inputs("gca_unk", (2, 1, 16, 24)) > 0.3).astype(np.float32))
    y = O.gca_attention(sd, "decoder.gca", f, al, unk)
    assert np.abs(y.numpy() - g["y"]).max() < 1e-4


def test_tam_operator_matches_reference():
    g = golden("op_tam.npz")
    sd = fixture_sd()
    x = torch.from_numpy(op_inputs("tam_x", (2, 128, 12, 16)))
    b = torch.from_numpy(op_inputs("tam_b", (2, 128, 12, 16)))
    f = torch.from_numpy(op_inputs("tam_f", (2, 128, 12, 16)))
    mask = torch.from_numpy((op_inputs("tam_m", (2, 1, 96, 128)) > 0.5).astype(np.float32))
    feat, attb, attf, sm = O.tam(sd, "decoder.fam", x, b, f, mask, 7)
    assert np.array_equal(sm.numpy(), g["small_mask"])
    assert np.abs(feat.numpy() - g["feat"]).max() < 1e-4
    assert np.abs(attb.numpy() - g["attb"]).max() < 1e-4
    assert np.abs(attf.numpy() - g["attf"]).max() < 1e-4


def test_full_vmd_forward_matches_reference():
    """FullModel_VMD.forward (eval mode, S=5, fixed dilation 3): losses + visual outputs."""
    g = golden("train_s5.npz")
    sd = fixture_sd()
    a, fg, bg = (torch.from_numpy(g[k]).float() for k in ("a", "fg", "bg"))
    out = O.full_vmd_forward(sd, a, fg, bg, radii=[3])
    losses = np.array([float(o) for o in out[:5]])
    assert np.allclose(losses, g["losses"], rtol=2e-4, atol=1e-6), (losses, g["losses"])
    assert np.abs(out[5].numpy() - g["scaled_imgs"]).max() < 1e-6
    assert np.abs(out[6].numpy() - g["tris_vis"]).max() < 1e-6
    assert np.abs(out[7].numpy() - g["alphas"]).max() < 5e-5
    assert np.abs(out[8].numpy() - g["comps"]).max() < 5e-5
