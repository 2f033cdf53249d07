"""Known-answer tests from the literature - numbers that were NOT produced by this repository.

They pin the oracle (CPU tests) and the CUDA path (``-m gpu`` tests) to published results:

* the complete base-case voltage profile of the IEEE 33-bus feeder (Baran & Wu 1989; the 33 values every
  reconfiguration / DG-placement paper reprints, identical to MATPOWER ``case33bw``), to the 4 printed decimals;
* Saadat, *Power System Analysis*, Example 6.7 (3-bus meshed system, slack 1.05 p.u., 100 MVA base): the exact solution
  V2 = 0.98 - j0.06, V3 = 1.00 - j0.05, slack power 409.5 MW + j189 Mvar - exercises the meshed (dense-LU) path;
* an off-nominal tap transformer feeding one load, solved in closed form (the pi model of PYPOWER's makeYbus puts the
  tap on the from side: the load bus sees the source V1 / tap behind the series impedance).
"""
import numpy as np
import pytest

from mapdn_b200 import cases
from mapdn_b200.network import NetDesc
from oracle.pandapower_nr import PandapowerEquivalent

# IEEE 33-bus base case, |V| in p.u. at buses 1..33 (12.66 kV, 3715 kW + j2300 kvar)
BARAN_WU_VM = np.array([
    1.0000, 0.9970, 0.9829, 0.9755, 0.9681, 0.9497, 0.9462, 0.9413, 0.9351, 0.9292, 0.9284,
    0.9269, 0.9208, 0.9185, 0.9171, 0.9157, 0.9137, 0.9131, 0.9965, 0.9929, 0.9922, 0.9916,
    0.9794, 0.9727, 0.9694, 0.9477, 0.9452, 0.9337, 0.9255, 0.9220, 0.9178, 0.9169, 0.9166])
ROUNDING = 0.5e-4 + 1e-6        # the profile is printed with 4 decimals


def saadat_6_7():
    z = np.array([0.02 + 0.04j, 0.01 + 0.03j, 0.0125 + 0.025j])
    net = NetDesc(base_mva=100.0, n_bus=3, slack_bus=0, slack_vm=1.05, br_from=np.array([0, 0, 1]),
                  br_to=np.array([1, 2, 2]), br_r=z.real, br_x=z.imag, load_bus=np.array([1, 2]),
                  sgen_bus=np.array([2]), sgen_zone=np.array([1]), bus_zone=np.array([0, 1, 1]), name="saadat_6_7")
    return net, np.array([256.6, 138.6]), np.array([110.2, 45.2])


SAADAT_V = np.array([1.05 + 0j, 0.98 - 0.06j, 1.00 - 0.05j])


def tap_two_bus(tap=0.96, v1=1.02, r=0.01, x=0.06, p=0.35, q=0.12):
    """(net, p, q, exact V2 magnitude): a load S = p + jq behind z = r + jx fed from V1 / tap."""
    net = NetDesc(base_mva=1.0, n_bus=2, slack_bus=0, slack_vm=v1, br_from=np.array([0]), br_to=np.array([1]),
                  br_r=np.array([r]), br_x=np.array([x]), br_tap=np.array([tap]), br_is_line=np.array([0], np.uint8),
                  load_bus=np.array([1]), sgen_bus=np.array([1]), sgen_zone=np.array([1]), bus_zone=np.array([0, 1]),
                  name="tap2")
    # |V2|^4 + (2 (pr + qx) - V0^2) |V2|^2 + (p^2 + q^2)(r^2 + x^2) = 0 with V0 = V1 / tap (upper root)
    v0 = v1 / tap
    b = 2.0 * (p * r + q * x) - v0 * v0
    c = (p * p + q * q) * (r * r + x * x)
    v2 = np.sqrt((-b + np.sqrt(b * b - 4.0 * c)) / 2.0)
    return net, np.array([p]), np.array([q]), v2


# ------------------------------------------------------------------ oracle (CPU) ---------------------------------
def test_oracle_baran_wu_full_voltage_profile():
    net, p, q = cases.baran_wu_nominal()
    r = PandapowerEquivalent(net).runpp(p, q, np.zeros(6), np.zeros(6))
    assert r.converged
    assert np.abs(r.vm_pu - BARAN_WU_VM).max() < ROUNDING
    assert int(np.argmin(r.vm_pu)) == 17                     # bus 18


def test_c_oracle_baran_wu_full_voltage_profile():
    from oracle import c_oracle
    net, p, q = cases.baran_wu_nominal()
    vm, _, conv, _ = c_oracle.COracle(net).runpp(p, q, np.zeros(6), np.zeros(6))
    assert conv and np.abs(vm - BARAN_WU_VM).max() < ROUNDING


def test_oracle_saadat_example_6_7():
    net, p, q = saadat_6_7()
    r = PandapowerEquivalent(net).runpp(p, q, np.zeros(1), np.zeros(1))
    assert r.converged and np.abs(r.V - SAADAT_V).max() < 1e-9
    assert abs(r.p_ext_mw - 409.5) < 1e-6 and abs(r.q_ext_mvar - 189.0) < 1e-6


def test_oracle_tap_transformer_closed_form():
    for tap in (0.92, 1.0, 1.07):
        net, p, q, v2 = tap_two_bus(tap=tap)
        r = PandapowerEquivalent(net).runpp(p, q, np.zeros(1), np.zeros(1))
        assert r.converged and abs(r.vm_pu[1] - v2) < 1e-10


# ------------------------------------------------------------------ CUDA path ------------------------------------
def _solve_gpu(net, p, q):
    import torch
    from mapdn_b200.env import BatchedVoltageControl
    env = BatchedVoltageControl(net, None, dict(voltage_barrier_type="l1"), batch=1)
    z = np.zeros((1, net.n_sgen))
    out = env.solve(p[None], q[None], z, z)
    torch.cuda.synchronize()
    res = {k: (None if v is None else v.cpu().numpy()[0]) for k, v in out.items()}
    env.close()
    return res


@pytest.mark.gpu
def test_gpu_baran_wu_full_voltage_profile():
    net, p, q = cases.baran_wu_nominal()
    out = _solve_gpu(net, p, q)
    assert out["converged"] == 1
    assert np.abs(out["vm"] - BARAN_WU_VM).max() < ROUNDING
    assert abs(out["pl"].sum() - 0.20268) < 5e-5              # published losses 202.68 kW


@pytest.mark.gpu
def test_gpu_saadat_example_6_7_meshed():
    net, p, q = saadat_6_7()
    out = _solve_gpu(net, p, q)
    V = out["vm"] * np.exp(1j * np.deg2rad(out["va_deg"]))
    assert out["converged"] == 1 and np.abs(V - SAADAT_V).max() < 1e-9
    assert abs(-out["p_bus"][0] - 409.5) < 1e-6 and abs(-out["q_bus"][0] - 189.0) < 1e-6   # res_bus at the slack = -infeed


@pytest.mark.gpu
def test_gpu_tap_transformer_closed_form():
    for tap in (0.92, 1.0, 1.07):
        net, p, q, v2 = tap_two_bus(tap=tap)
        out = _solve_gpu(net, p, q)
        assert out["converged"] == 1 and abs(out["vm"][1] - v2) < 1e-10
