"""Shared helpers for the test-suite (oracle access, fixture weights, golden vectors)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tcvom_b200 import synthetic  # noqa: E402


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def key_table():
    with open(os.path.join(GOLDEN, "vmn_gca_keys.json")) as f:
        return json.load(f)


_SD = {}


def fixture_sd():
    """The calibrated fixture checkpoint (584 keys, CPU fp32)."""
    if "sd" not in _SD:
        shapes = {k: tuple(s) for k, s in key_table()["state_dict"]}
        _SD["sd"] = synthetic.fixture_state_dict(shapes, 0)
    return _SD["sd"]


def op_inputs(tag, shape, seed=11):
    return synthetic._rng(tag, seed).standard_normal(size=shape).astype(np.float32)


def key_table_fba():
    with open(os.path.join(GOLDEN, "vmn_fba_keys.json")) as f:
        return json.load(f)


def fixture_sd_fba():
    """The seeded ``vmn_fba`` fixture checkpoint (203 keys, CPU fp32)."""
    if "fba" not in _SD:
        shapes = {k: tuple(s) for k, s in key_table_fba()["state_dict"]}
        _SD["fba"] = synthetic.fixture_state_dict_fba(shapes, 0)
    return _SD["fba"]


def fixture_sd_dim():
    """The seeded ``vmn_dim`` fixture checkpoint (113 keys / 134 M parameters, CPU fp32; regenerated, never committed)."""
    if "dim" not in _SD:
        from oracle.vmn_dim_oracle import fixture_sd_dim as make
        _SD["dim"] = make()
    return _SD["dim"]


def key_table_index():
    with open(os.path.join(GOLDEN, "vmn_index_keys.json")) as f:
        return json.load(f)


def fixture_sd_index():
    """The seeded ``vmn_index`` fixture checkpoint (555 keys, CPU fp32)."""
    if "index" not in _SD:
        from oracle.vmn_index_oracle import fixture_sd_index as make
        _SD["index"] = make({k: tuple(s) for k, s in key_table_index()["state_dict"]})
    return _SD["index"]
