"""The DIM+TAM CPU oracle (oracle/vmn_dim_oracle.py) against golden vectors produced by the unmodified reference
(tests/golden/make_golden_dim.py), and the state_dict layout of the native ``vmn_dim`` module.  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, fixture_sd_dim, golden
from oracle import vmn_dim_oracle as O

CASES = ["ring64x96", "allunk64", "dil96x64", "batch2_64"]


@pytest.mark.parametrize("name", CASES)
def test_eval_forward_matches_reference(name):
    g = golden(f"dim_{name}.npz")
    sd = fixture_sd_dim()
    dil = int(g["dilate"])
    imgs, tris = torch.from_numpy(g["imgs"]).float(), torch.from_numpy(g["tris"]).float()
    alphas, aux = O.eval_forward(sd, imgs, tris, None if dil < 0 else dil, 7, return_aux=True)
    assert np.abs(alphas.numpy() - g["alphas"]).max() < 2e-5
    assert np.abs(aux["preds"][1].numpy() - g["pred1"]).max() < 2e-5
    assert np.array_equal(aux["small_mask"][1].numpy(), g["small_mask1"])
    for k, ref in (("attb", g["attb1"]), ("attf", g["attf1"])):
        ref = ref.astype(np.float32)
        assert np.abs(aux[k][1].numpy() - ref).max() < 2e-3 * max(1.0, np.abs(ref).max())   # fp16 storage
    assert float(alphas[:, 0].abs().max()) == 0 and float(alphas[:, -1].abs().max()) == 0
    p = aux["preds"][1]
    assert float(((p > 1e-3) & (p < 1 - 1e-3)).float().mean()) > 0.9, "vacuous fixture: matte sits on the clamp"


def test_preprocess_and_head_match_reference():
    g = golden("dim_ring64x96.npz")
    sd = fixture_sd_dim()
    imgs, tris = torch.from_numpy(g["imgs"]).float(), torch.from_numpy(g["tris"]).float()
    x4, trimask = O.eval_preprocess(imgs, tris, None)
    assert np.abs(x4.numpy() - g["x4"]).max() < 1e-6
    assert np.array_equal(trimask.numpy().astype(np.uint8), g["trimask"])
    with torch.no_grad():
        idxs, x6 = O.encoder(x4[:, 1], sd)
        feat = O.decoder_head(idxs, x6, sd)
    assert np.abs(feat.numpy() - g["feat1"]).max() < 2e-5


def test_native_module_has_the_reference_state_dict_layout():
    import tcvom_b200
    with open(os.path.join(GOLDEN, "vmn_dim_keys.json")) as f:
        want = [(k, tuple(s)) for k, s in json.load(f)["state_dict"]]
    net = tcvom_b200.get_VMN_models("vmn_dim", agg_window=7)
    got = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    assert got == want
    net.load_state_dict(fixture_sd_dim(), strict=True)
