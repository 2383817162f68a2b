"""CPU-side checks: C-ABI library exports, state_dict contract, host logic (no GPU needed)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from helpers import ROOT, fixture_sd, key_table


def test_library_exports_every_declared_symbol():
    from tcvom_b200 import _cabi
    hdr = open(os.path.join(ROOT, "include", "tcvom_b200.h")).read()
    declared = set(re.findall(r"\b(tcv_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"tcv_conv_desc"}
    assert declared, "no declarations parsed"
    assert os.path.exists(_cabi.LIB_PATH), "native library not built (run __graft_entry__.build())"
    L = ctypes.CDLL(_cabi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/tcvom_b200.h but not exported"
    assert declared == set(_cabi.SIGNATURES), declared ^ set(_cabi.SIGNATURES)
    assert L.tcv_version() >= 100          # pure host call, no CUDA needed


def test_conv_desc_layout_matches_header():
    from tcvom_b200 import _cabi
    # 3 pointers/ll + 4 int + ptr + int + 48 int + 2 int + 2 ptr + 9 int + ... : check against C via sizeof probe
    assert ctypes.sizeof(_cabi.ConvDesc) % 8 == 0
    assert _cabi.ConvDesc.wtap.offset == _cabi.ConvDesc.dx.offset + 4 * _cabi.MAX_TAPS


def test_state_dict_contract_matches_reference_layout():
    import tcvom_b200
    net = tcvom_b200.get_VMN_models("vmn_gca", agg_window=7)
    kt = key_table()
    mine = [[k, list(v.shape)] for k, v in net.state_dict().items()]
    assert mine == kt["state_dict"]                      # names, shapes AND order (584 keys)
    assert [n for n, p in net.named_parameters() if p.requires_grad] == kt["trainable"]
    net.load_state_dict(fixture_sd(), strict=True)       # pred_test.py:92 contract
    m = tcvom_b200.EvalModel(model="vmn_gca", agg_window=7, dilate_kernel=None)
    assert hasattr(m, "NET") and len(m.NET.state_dict()) == 584


def test_get_vmn_models_error_behaviour():
    import tcvom_b200
    with pytest.raises(ValueError):
        tcvom_b200.get_VMN_models("nope", agg_window=7)   # VMN/__init__.py:26-27
    for arch in ("vmn_gca", "vmn_fba", "vmn_dim", "vmn_index"):          # the four base networks of models/VMN/__init__.py
        assert tcvom_b200.get_VMN_models(arch, agg_window=7) is not None


def test_fba_state_dict_contract_matches_reference_layout():
    import tcvom_b200
    from helpers import fixture_sd_fba, key_table_fba
    net = tcvom_b200.get_VMN_models("vmn_fba", agg_window=7)
    kt = key_table_fba()
    mine = [[k, list(v.shape)] for k, v in net.state_dict().items()]
    assert mine == kt["state_dict"]                      # names, shapes AND order (203 keys)
    assert [n for n, p in net.named_parameters() if p.requires_grad] == kt["trainable"]
    net.load_state_dict(fixture_sd_fba(), strict=True)
    m = tcvom_b200.EvalModel(model="vmn_fba", agg_window=7, dilate_kernel=None)
    assert m.TRIMAP_CHANNEL == 8 and len(m.NET.state_dict()) == 203
    with pytest.raises(RuntimeError):                    # CPU tensors: no fallback
        m.eval()(torch.zeros(1, 3, 3, 64, 64), torch.zeros(1, 3, 1, 64, 64))


def test_cpu_module_refuses_to_run():
    import tcvom_b200
    m = tcvom_b200.EvalModel(model="vmn_gca", agg_window=7).eval()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 3, 64, 64), torch.zeros(1, 3, 1, 64, 64))


def test_synthetic_window_is_deterministic():
    from tcvom_b200 import synthetic
    a = synthetic.make_window(64, 96, seed=3)
    b = synthetic.make_window(64, 96, seed=3)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert set(np.unique(a[1])) <= {0, 128, 255}
    from helpers import golden
    g = golden("eval_ring64.npz")
    imgs, tris = synthetic.make_window(64, 64, seed=7)
    assert np.array_equal(tris, g["tris"])


def test_missing_native_library_fails_loudly(monkeypatch):
    """No CPU / PyTorch fallback: without the .so the product path raises instead of computing elsewhere."""
    from tcvom_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", "/nonexistent/libtcvom_b200.so")
    with pytest.raises(RuntimeError, match="native library not built"):
        _cabi.lib()


def test_product_package_does_not_import_the_oracle():
    import pathlib
    for f in pathlib.Path(ROOT, "tcvom_b200").rglob("*.py"):
        src = f.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f
