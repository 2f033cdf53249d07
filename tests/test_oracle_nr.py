"""Pins the oracle's power-flow restatement (PARITY UNPINNED against pandapower itself - see
oracle/pandapower_nr.py): closed form, literature values, independent solver, KCL residual."""
import numpy as np
import pytest

from mapdn_b200 import cases
from mapdn_b200.network import NetDesc
from oracle.independent import backward_forward_sweep, two_bus_closed_form
from oracle.pandapower_nr import PandapowerEquivalent, bus_demand, make_ybus
from conftest import random_tree_net


def test_two_bus_closed_form():
    r, x, p, q, v0 = 0.02, 0.04, 0.8, 0.3, 1.03
    net = NetDesc(base_mva=1.0, n_bus=2, slack_bus=0, slack_vm=v0, br_from=[0], br_to=[1], br_r=[r], br_x=[x],
                  load_bus=[1], sgen_bus=[1], sgen_zone=[1], bus_zone=[0, 1])
    res = PandapowerEquivalent(net).runpp([p], [q], [0.0], [0.0])
    assert res.converged
    assert abs(res.vm_pu[1] - two_bus_closed_form(v0, r, x, p, q)) < 1e-9  # NR stops at ||F|| < 1e-8
    assert abs(res.vm_pu[0] - v0) < 1e-15
    # loss = slack infeed - load
    assert abs(res.pl_mw[0] - (res.p_ext_mw - p)) < 1e-10


def test_baran_wu_published_results():
    """IEEE 33-bus (Baran & Wu 1989): min voltage 0.9131 p.u. at bus 18, losses 202.7 kW / 135.1 kvar."""
    net, p, q = cases.baran_wu_nominal()
    res = PandapowerEquivalent(net).runpp(p, q, np.zeros(6), np.zeros(6))
    assert res.converged and res.iterations in (3, 4)
    assert abs(res.vm_pu.min() - 0.9131) < 1e-4
    assert int(res.vm_pu.argmin()) == 17
    assert abs(res.pl_mw.sum() * 1e3 - 202.68) < 0.05
    assert abs((res.q_ext_mvar - q.sum()) * 1e3 - 135.14) < 0.05
    assert abs(res.p_ext_mw - (p.sum() + res.pl_mw.sum())) < 1e-10


@pytest.mark.parametrize("name", ["case33", "case141", "case322"])
def test_newton_vs_backward_forward_sweep(name):
    net = cases.make_case(name)
    inp = cases.synthetic_inputs(name, 3, seed=3)
    pf = PandapowerEquivalent(net)
    for e in range(3):
        qs = inp["action"][e] * np.sqrt(inp["s_max"] ** 2 - inp["p_pv"][e] ** 2)
        res = pf.runpp(inp["p_load"][e], inp["q_load"][e], inp["p_pv"][e], qs)
        assert res.converged and res.iterations <= 5
        PD, QD = bus_demand(net, inp["p_load"][e], inp["q_load"][e], inp["p_pv"][e], qs)
        V = backward_forward_sweep(net, PD, QD)
        assert np.abs(np.abs(V) - res.vm_pu).max() < 1e-9
        assert np.abs(np.angle(V) - np.angle(res.V)).max() < 1e-9
        # KCL at every PQ bus
        S = -(PD + 1j * QD) / net.base_mva
        mis = res.V * np.conj(pf.Ybus @ res.V) - S
        assert np.abs(np.delete(mis, net.slack_bus)).max() < 1e-8


def test_ybus_tap_shunt_formulas():
    """makeYbus with an off-nominal tap, phase shift, charging and a bus shunt, against the
    textbook pi-model written out by hand."""
    net = NetDesc(base_mva=100.0, n_bus=2, slack_bus=0, slack_vm=1.0, br_from=[0], br_to=[1], br_r=[0.01],
                  br_x=[0.05], br_b=[0.2], br_g=[0.02], br_tap=[0.97], br_shift=[5.0], load_bus=[1],
                  sgen_bus=[1], sgen_zone=[1], bus_zone=[0, 1], bus_gs=[0.0, 3.0], bus_bs=[0.0, -7.0])
    Y = make_ybus(net)[0].toarray()
    ys = 1 / (0.01 + 0.05j)
    tau = 0.97 * np.exp(1j * np.deg2rad(5.0))
    ytt = ys + (0.02 + 0.2j) / 2
    assert np.allclose(Y[0, 0], ytt / abs(tau) ** 2, rtol=1e-14)
    assert np.allclose(Y[0, 1], -ys / np.conj(tau), rtol=1e-14)
    assert np.allclose(Y[1, 0], -ys / tau, rtol=1e-14)
    assert np.allclose(Y[1, 1], ytt + (3.0 - 7.0j) / 100.0, rtol=1e-14)


def test_general_net_self_consistency():
    net = random_tree_net(23, 4, seed=11)
    rng = np.random.default_rng(0)
    pf = PandapowerEquivalent(net)
    pl, ql = rng.uniform(0, 0.3, net.n_load), rng.uniform(0, 0.1, net.n_load)
    pg, qg = rng.uniform(0, 0.5, net.n_sgen), rng.uniform(-0.2, 0.2, net.n_sgen)
    res = pf.runpp(pl, ql, pg, qg)
    assert res.converged
    PD, QD = bus_demand(net, pl, ql, pg, qg)
    S = -(PD + 1j * QD) / net.base_mva
    mis = res.V * np.conj(pf.Ybus @ res.V) - S
    assert np.abs(np.delete(mis, net.slack_bus)).max() < 1e-8
    assert abs(res.vm_pu[net.slack_bus] - 1.02) < 1e-13 and abs(res.va_degree[net.slack_bus] - 3.0) < 1e-12
    # total balance: ext infeed = demand + branch losses (all branches) + shunt consumption
    Sf = res.V[net.br_from] * np.conj(pf.Yf @ res.V) * net.base_mva
    St = res.V[net.br_to] * np.conj(pf.Yt @ res.V) * net.base_mva
    shunt = (np.abs(res.V) ** 2 * net.bus_gs).sum()
    assert abs(res.p_ext_mw - (PD.sum() + (Sf + St).real.sum() + shunt)) < 1e-8
    assert res.pl_mw.shape == (int(net.br_is_line.sum()),)


def test_divergence_is_reported():
    net, p, q = cases.baran_wu_nominal()
    res = PandapowerEquivalent(net).runpp(p * 40, q * 40, np.zeros(6), np.zeros(6))
    assert not res.converged and res.iterations == 10
