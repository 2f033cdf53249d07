"""Parity against the REAL pandapower - active as soon as tests/golden/pp_*.npz exist.

Those files are written by scripts/pin_with_pandapower.py from `pp.runpp` (needs pandapower, which this image does
not have; until somebody runs the script the tests below skip and the oracle stays "parity unpinned", see DESIGN.md
§1). Tolerances are the north-star ones: |dV| <= 1e-6 p.u.; result columns to 1e-5.
"""
import glob
import os

import numpy as np
import pytest

from conftest import ROOT

PINS = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "pp_*.npz")))
needs_pins = pytest.mark.skipif(not PINS, reason="no tests/golden/pp_*.npz - run scripts/pin_with_pandapower.py where "
                                                 "pandapower is installed")


def _net(name):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from pin_with_pandapower import nets
    return nets()[name][0]


@needs_pins
@pytest.mark.parametrize("path", PINS or [None])
def test_oracle_matches_pandapower(path):
    from oracle.pandapower_nr import PandapowerEquivalent
    g = np.load(path)
    pf = PandapowerEquivalent(_net(os.path.basename(path)[3:-4]))
    for e in range(g["p_load"].shape[0]):
        r = pf.runpp(g["p_load"][e], g["q_load"][e], g["p_pv"][e], g["q"][e])
        assert r.converged == bool(g["converged"][e])
        if not r.converged:
            continue
        assert np.abs(r.vm_pu - g["vm_pu"][e]).max() < 1e-6
        assert np.abs(r.va_degree - g["va_degree"][e]).max() < 1e-4
        assert np.abs(r.p_mw - g["p_mw"][e]).max() < 1e-5 and np.abs(r.q_mvar - g["q_mvar"][e]).max() < 1e-5
        assert np.abs(r.pl_mw - g["pl_mw"][e]).max() < 1e-5
        if g["iterations"][e] >= 0:
            assert r.iterations == g["iterations"][e]


@needs_pins
@pytest.mark.gpu
@pytest.mark.parametrize("path", PINS or [None])
def test_cuda_path_matches_pandapower(path):
    import torch
    from mapdn_b200.env import BatchedVoltageControl
    g = np.load(path)
    env = BatchedVoltageControl(_net(os.path.basename(path)[3:-4]), None, dict(voltage_barrier_type="l1"), batch=1)
    out = env.solve(g["p_load"], g["q_load"], g["p_pv"], g["q"])
    torch.cuda.synchronize()
    conv = out["converged"].cpu().numpy().astype(bool)
    assert (conv == g["converged"].astype(bool)).all()
    assert np.abs(out["vm"].cpu().numpy() - g["vm_pu"])[conv].max() < 1e-6
    assert np.abs(out["va_deg"].cpu().numpy() - g["va_degree"])[conv].max() < 1e-4
    assert np.abs(out["p_bus"].cpu().numpy() - g["p_mw"])[conv].max() < 1e-5
    assert np.abs(out["pl"].cpu().numpy() - g["pl_mw"])[conv].max() < 1e-5
    env.close()
