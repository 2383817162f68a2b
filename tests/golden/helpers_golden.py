"""Loads the committed fixture checkpoint for partial golden regeneration (make_golden.py --only-train)."""
import json
import os

from tcvom_b200 import synthetic

HERE = os.path.dirname(os.path.abspath(__file__))


def load_full():
    kt = json.load(open(os.path.join(HERE, "vmn_gca_keys.json")))
    return synthetic.fixture_state_dict({k: tuple(s) for k, s in kt["state_dict"]}, 0)
