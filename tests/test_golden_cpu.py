"""The oracle reproduces the committed golden fixtures (tests/golden, made by scripts/make_golden.py)."""
import os

import numpy as np
import pytest

from conftest import ROOT, random_tree_net
from mapdn_b200 import cases
from oracle.pandapower_nr import PandapowerEquivalent

GOLD = os.path.join(ROOT, "tests", "golden")


def _net(name):
    if name == "baran_wu":
        return cases.baran_wu_nominal()[0]
    if name == "rand23":
        return random_tree_net(23, 4, seed=11)
    return cases.make_case(name)


@pytest.mark.parametrize("name", ["case33", "case141", "case322", "baran_wu", "rand23"])
def test_oracle_matches_golden(name):
    g = np.load(os.path.join(GOLD, f"solve_{name}.npz"))
    pf = PandapowerEquivalent(_net(name))
    for e in range(g["p_load"].shape[0]):
        r = pf.runpp(g["p_load"][e], g["q_load"][e], g["p_pv"][e], g["q"][e])
        assert r.converged == bool(g["converged"][e]) and r.iterations == g["iterations"][e]
        assert np.abs(r.vm_pu - g["vm_pu"][e]).max() < 1e-12
        assert np.abs(r.va_degree - g["va_degree"][e]).max() < 1e-10
        assert np.abs(r.pl_mw - g["pl_mw"][e]).max() < 1e-10


def test_baran_wu_golden_is_the_published_case():
    g = np.load(os.path.join(GOLD, "solve_baran_wu.npz"))
    assert abs(g["vm_pu"][0].min() - 0.9131) < 1e-4 and abs(g["pl_mw"][0].sum() - 0.20268) < 5e-5
