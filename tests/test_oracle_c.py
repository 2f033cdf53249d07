"""The plain-C restatement (oracle/c/nr_dense.c: dense Jacobian, angle-then-magnitude unknown order, LU with partial
pivoting) against the NumPy restatement (sparse dS/dV + spsolve): two independently written versions of pandapower's
newtonpf must produce the same iterates."""
import numpy as np
import pytest

from conftest import random_tree_net
from mapdn_b200 import cases
from oracle.c_oracle import COracle
from oracle.pandapower_nr import PandapowerEquivalent


@pytest.mark.parametrize("name", ["case33", "case141", "rand23", "baran_wu"])
def test_c_and_numpy_oracles_agree(name):
    if name == "rand23":
        net = random_tree_net(23, 4, seed=11)
        rng = np.random.default_rng(0)
        pl, ql = rng.uniform(0, 0.3, (4, net.n_load)), rng.uniform(0, 0.1, (4, net.n_load))
        pv, q = rng.uniform(0, 0.5, (4, net.n_sgen)), rng.uniform(-0.2, 0.2, (4, net.n_sgen))
    elif name == "baran_wu":
        net, p, qq = cases.baran_wu_nominal()
        pl, ql, pv, q = p[None], qq[None], np.zeros((1, 6)), np.zeros((1, 6))
    else:
        net = cases.make_case(name)
        inp = cases.synthetic_inputs(name, 4, seed=8)
        pl, ql, pv = inp["p_load"], inp["q_load"], inp["p_pv"]
        q = inp["action"] * np.sqrt(inp["s_max"] ** 2 - pv ** 2)
    a, b = COracle(net), PandapowerEquivalent(net)
    for e in range(pl.shape[0]):
        vm, va, conv, it = a.runpp(pl[e], ql[e], pv[e], q[e])
        r = b.runpp(pl[e], ql[e], pv[e], q[e])
        assert conv and r.converged and it == r.iterations
        assert np.abs(vm - r.vm_pu).max() < 1e-11 and np.abs(va - r.va_degree).max() < 1e-9
    if name == "baran_wu":
        assert abs(vm.min() - 0.9131) < 1e-4


def test_c_oracle_reports_divergence():
    net, p, q = cases.baran_wu_nominal()
    vm, va, conv, it = COracle(net).runpp(p * 40, q * 40, np.zeros(6), np.zeros(6))
    assert not conv and it == 10


@pytest.mark.parametrize("seed", range(16))
def test_random_radial_nets_three_independent_solvers(seed):
    """Random feeders (3..60 buses; taps, charging, shunts, scaling, a parallel twin, an open branch): the NumPy and
    the C restatement of newtonpf agree iterate for iterate; on the plain variant of the same feeder (no shunt
    elements, no taps) both also agree with a current-summation backward/forward sweep, which shares no code and no
    algorithm with them."""
    import dataclasses
    from oracle.independent import backward_forward_sweep
    from oracle.pandapower_nr import bus_demand
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(3, 61))
    ng = int(rng.integers(1, min(6, n)))
    net = random_tree_net(n, ng, seed=seed, with_taps=bool(seed % 2))
    pl, ql = rng.uniform(0, 0.25, net.n_load), rng.uniform(0, 0.08, net.n_load)
    pv, q = rng.uniform(0, 0.6, ng), rng.uniform(-0.3, 0.3, ng)
    vm, va, conv, it = COracle(net).runpp(pl, ql, pv, q)
    r = PandapowerEquivalent(net).runpp(pl, ql, pv, q)
    assert conv and r.converged and it == r.iterations
    assert np.abs(vm - r.vm_pu).max() < 1e-11 and np.abs(va - r.va_degree).max() < 1e-9

    status = net.br_status.copy(); status[-2] = 0                      # drop the parallel twin for the sweep
    plain = dataclasses.replace(net, br_b=None, br_g=None, br_tap=None, bus_gs=None, bus_bs=None, br_status=status)
    vm2, va2, conv2, _ = COracle(plain).runpp(pl, ql, pv, q)
    r2 = PandapowerEquivalent(plain).runpp(pl, ql, pv, q)
    PD, QD = bus_demand(plain, pl, ql, pv, q)
    V = backward_forward_sweep(plain, PD, QD) * np.exp(1j * np.deg2rad(plain.slack_va_deg))
    assert conv2 and r2.converged
    assert np.abs(np.abs(V) - r2.vm_pu).max() < 1e-9 and np.abs(np.abs(V) - vm2).max() < 1e-9
    assert np.abs(np.angle(V * np.conj(r2.V))).max() < 1e-9
