"""The IndexNet+TAM CPU oracle (oracle/vmn_index_oracle.py) against golden vectors produced by the unmodified reference
(tests/golden/make_golden_index.py), and the state_dict layout of the native ``vmn_index`` module.  CPU only."""
import numpy as np
import pytest
import torch

from helpers import fixture_sd_index, golden, key_table_index
from oracle import vmn_index_oracle as O

CASES = ["ring64x96", "dil96x64", "batch2_64"]


@pytest.mark.parametrize("name", CASES)
def test_eval_forward_matches_reference(name):
    g = golden(f"index_{name}.npz")
    sd = fixture_sd_index()
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
    assert float(aux["preds"][1].std()) > 0.03, "vacuous fixture"


def test_head_feature_matches_reference():
    g = golden("index_ring64x96.npz")
    sd = fixture_sd_index()
    x4 = torch.from_numpy(g["x4"])
    with torch.no_grad():
        feat = O.decoder_head(O.encoder(x4[:, 1], sd), sd)
    assert np.abs(feat.numpy() - g["feat1"]).max() < 2e-5


def test_native_module_has_the_reference_state_dict_layout():
    import tcvom_b200
    want = [(k, tuple(s)) for k, s in key_table_index()["state_dict"]]
    net = tcvom_b200.get_VMN_models("vmn_index", agg_window=7)
    got = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    assert got == want
    net.load_state_dict(fixture_sd_index(), strict=True)
