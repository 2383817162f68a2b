"""The FBA+TAM CPU oracle (oracle/vmn_fba_oracle.py) against golden vectors produced by the unmodified reference
(tests/golden/make_golden_fba.py).  CPU only."""
import numpy as np
import pytest
import torch

from helpers import fixture_sd_fba, golden
from oracle import vmn_fba_oracle as O

CASES = ["ring64", "ring96x128", "allunk64", "nounk64", "dil64x96", "batch2_64"]


@pytest.mark.parametrize("name", CASES)
def test_eval_forward_matches_reference(name):
    g = golden(f"fba_{name}.npz")
    sd = fixture_sd_fba()
    dil = int(g["dilate"])
    imgs, tris = torch.from_numpy(g["imgs"]).float(), torch.from_numpy(g["tris"]).float()
    alphas, Fs, Bs, aux = O.eval_forward(sd, imgs, tris, None if dil < 0 else dil, 7, return_aux=True)
    assert np.array_equal(aux["trimask"].numpy().astype(np.uint8), g["trimask"])
    assert np.abs(alphas.numpy() - g["alphas"]).max() < 5e-5
    assert np.abs(aux["preds"][1].numpy() - g["pred1"]).max() < 5e-5
    assert np.abs(Fs[:, 1].numpy() - g["Fs"].astype(np.float32)).max() < 1e-3      # stored as fp16
    assert np.abs(Bs[:, 1].numpy() - g["Bs"].astype(np.float32)).max() < 1e-3
    assert np.array_equal(aux["small"][1].numpy(), g["small_mask1"])
    for k, ref in (("attb", g["attb1"]), ("attf", g["attf1"])):
        ref = ref.astype(np.float32)
        assert np.abs(aux[k][1].numpy() - ref).max() < 2e-3 * max(1.0, np.abs(ref).max())   # fp16 storage
    assert float(alphas[:, 0].abs().max()) == 0 and float(alphas[:, -1].abs().max()) == 0


def test_preprocess_and_head_match_reference():
    g = golden("fba_ring64.npz")
    sd = fixture_sd_fba()
    imgs, tris = torch.from_numpy(g["imgs"]).float(), torch.from_numpy(g["tris"]).float()
    scaled, x11, trimask, trimap2 = O.eval_preprocess(imgs, tris, None)
    assert np.abs(x11.numpy() - g["x11"]).max() < 1e-6             # incl. the 6 distance-transform channels
    with torch.no_grad():
        co = O.encoder(x11[:, 1], sd)
        feat = O.decoder_head(co, sd)
    assert np.abs(co[-1].numpy()[:, ::64] - g["conv5_sample"]).max() < 1e-3      # values up to ~5: fp32 summation-order noise through 16 blocks
    assert np.abs(feat.numpy() - g["feat1"]).max() < 5e-4


def test_weight_standardisation_properties():
    w = torch.randn(8, 5, 3, 3)
    s = O.ws_weight(w)
    assert float(s.mean(dim=(1, 2, 3)).abs().max()) < 1e-6
    assert float((s.reshape(8, -1).var(dim=1) - 1).abs().max()) < 1e-3
