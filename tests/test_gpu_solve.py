"""Parity of the CUDA Newton-Raphson (through the C-ABI) with the oracle / golden fixtures.
Tolerances: BASELINE.json asks for 1e-6 p.u. on voltages; both sides converge the same Newton
iteration to ||F|| < 1e-8, so the tests hold the CUDA path to 1e-9 and identical iteration counts."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, random_tree_net
from mapdn_b200 import cases
from mapdn_b200.network import NetDesc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")
VTOL = 1e-9


def _env(net, batch=1, **kw):
    from mapdn_b200.env import BatchedVoltageControl
    return BatchedVoltageControl(net, None, None, batch=batch, **kw)


def _net(name):
    if name == "baran_wu":
        return cases.baran_wu_nominal()[0]
    if name == "rand23":
        return random_tree_net(23, 4, seed=11)
    return cases.make_case(name)


@pytest.mark.parametrize("name", ["case33", "case141", "case322", "baran_wu", "rand23"])
@pytest.mark.parametrize("lanes", [0, 4, 8, 16, 32, 64, 128])
def test_solve_matches_golden(name, lanes):
    from mapdn_b200._capi import MapdnError
    g = np.load(os.path.join(GOLD, f"solve_{name}.npz"))
    try:
        env = _env(_net(name), lanes_per_env=lanes)
    except MapdnError as ex:
        assert "too large" in str(ex) and name == "case322" and lanes in (4, 8)
        return
    out = env.solve(g["p_load"], g["q_load"], g["p_pv"], g["q"])
    torch.cuda.synchronize()
    assert np.array_equal(out["converged"].cpu().numpy(), g["converged"])
    assert np.array_equal(out["iterations"].cpu().numpy(), g["iterations"])
    assert np.abs(out["vm"].cpu().numpy() - g["vm_pu"]).max() < VTOL
    assert np.abs(out["va_deg"].cpu().numpy() - g["va_degree"]).max() < 1e-8
    assert np.abs(out["p_bus"].cpu().numpy() - g["p_mw"]).max() < 1e-8
    assert np.abs(out["q_bus"].cpu().numpy() - g["q_mvar"]).max() < 1e-8
    assert np.abs(out["pl"].cpu().numpy() - g["pl_mw"]).max() < 1e-9
    env.close()


def test_baran_wu_literature_values_on_gpu():
    net, p, q = cases.baran_wu_nominal()
    env = _env(net)
    out = env.solve(p[None], q[None], np.zeros((1, 6)), np.zeros((1, 6)))
    vm = out["vm"].cpu().numpy()[0]
    assert abs(vm.min() - 0.9131) < 1e-4 and int(vm.argmin()) == 17
    assert abs(out["pl"].cpu().numpy().sum() * 1e3 - 202.68) < 0.05


def test_ybus_assembly_matches_makeYbus():
    from oracle.pandapower_nr import make_ybus
    for net in (random_tree_net(23, 4, seed=11), random_tree_net(40, 5, seed=3), cases.case141()):
        env = _env(net)
        Y, Yo = env.ybus_dense(), make_ybus(net)[0].toarray()
        assert np.abs(Y - Yo).max() <= 1e-12 * np.abs(Yo).max()
        env.close()


@pytest.mark.parametrize("name,batch", [("case33", 4096), ("case141", 2048), ("case322", 1024)])
def test_full_size_batches_satisfy_kcl(name, batch):
    """BASELINE.json batch sizes: size-independent property - every solved env satisfies the power-flow
    equations (||V conj(Ybus V) - Sbus||inf < tol at the PQ buses) and the bus-demand bookkeeping."""
    from oracle.pandapower_nr import make_ybus
    net = cases.make_case(name)
    inp = cases.synthetic_inputs(name, batch, seed=9)
    q = inp["action"] * np.sqrt(inp["s_max"] ** 2 - inp["p_pv"] ** 2)
    env = _env(net)
    out = env.solve(inp["p_load"], inp["q_load"], inp["p_pv"], q)
    torch.cuda.synchronize()
    assert bool(out["converged"].all()) and int(out["iterations"].max()) <= 6
    dev = out["vm"].device
    Y = torch.tensor(make_ybus(net)[0].toarray(), dtype=torch.complex128, device=dev)
    V = torch.polar(out["vm"], torch.deg2rad(out["va_deg"]))
    S = V * torch.conj(V @ Y.T)                                          # injection into the network, p.u.
    PD = torch.zeros(batch, net.n_bus, dtype=torch.float64, device=dev)
    QD = torch.zeros_like(PD)
    PD.index_add_(1, torch.tensor(net.load_bus, device=dev, dtype=torch.long), torch.tensor(inp["p_load"], device=dev))
    QD.index_add_(1, torch.tensor(net.load_bus, device=dev, dtype=torch.long), torch.tensor(inp["q_load"], device=dev))
    PD.index_add_(1, torch.tensor(net.sgen_bus, device=dev, dtype=torch.long), -torch.tensor(inp["p_pv"], device=dev))
    QD.index_add_(1, torch.tensor(net.sgen_bus, device=dev, dtype=torch.long), -torch.tensor(q, device=dev))
    mis = S + torch.complex(PD, QD) / net.base_mva
    pq = [b for b in range(net.n_bus) if b != net.slack_bus]
    assert float(torch.view_as_real(mis[:, pq]).abs().max()) < 1e-8      # ||F||inf over Re/Im parts
    assert float((out["p_bus"][:, pq] - PD[:, pq]).abs().max()) < 1e-12
    # slack row = -(ext-grid infeed); losses = infeed - demand
    loss = -out["p_bus"][:, net.slack_bus] - PD[:, pq].sum(1)
    assert float((out["pl"].sum(1) - loss).abs().max()) < 1e-8 * net.n_bus      # sum of per-bus mismatches
    # spot-check 4 envs against the oracle
    from oracle.pandapower_nr import PandapowerEquivalent
    pf = PandapowerEquivalent(net)
    for e in (0, 1, batch // 2, batch - 1):
        r = pf.runpp(inp["p_load"][e], inp["q_load"][e], inp["p_pv"][e], q[e])
        assert np.abs(out["vm"][e].cpu().numpy() - r.vm_pu).max() < VTOL
        assert int(out["iterations"][e]) == r.iterations


@pytest.mark.parametrize("nb", [1, 2, 3, 15, 17, 33, 257])
def test_ragged_batch_sizes(nb):
    net = cases.case33()
    from oracle.pandapower_nr import PandapowerEquivalent
    inp = cases.synthetic_inputs("case33", nb, seed=nb)
    q = inp["action"] * np.sqrt(inp["s_max"] ** 2 - inp["p_pv"] ** 2)
    env = _env(net, batch=1)
    out = env.solve(inp["p_load"], inp["q_load"], inp["p_pv"], q)
    pf = PandapowerEquivalent(net)
    vm = out["vm"].cpu().numpy()
    for e in range(nb):
        assert np.abs(vm[e] - pf.runpp(inp["p_load"][e], inp["q_load"][e], inp["p_pv"][e], q[e]).vm_pu).max() < VTOL


def test_divergence_is_flagged_per_env():
    net, p, q = cases.baran_wu_nominal()
    pl = np.stack([p, p * 40, p]); ql = np.stack([q, q * 40, q])
    env = _env(net)
    out = env.solve(pl, ql, np.zeros((3, 6)), np.zeros((3, 6)))
    assert out["converged"].cpu().tolist() == [1, 0, 1]
    assert out["iterations"].cpu().tolist()[1] == 10
    assert np.abs(out["vm"][0].cpu().numpy() - out["vm"][2].cpu().numpy()).max() == 0.0


def _case33_meshed():
    """IEEE 33-bus with its five tie lines closed (Baran & Wu): a weakly meshed feeder."""
    net = cases.case33()
    zb = 12.66 ** 2
    ties = [(8, 21, 2.0, 2.0), (9, 15, 2.0, 2.0), (12, 22, 2.0, 2.0), (18, 33, 0.5, 0.5), (25, 29, 0.5, 0.5)]
    f = np.r_[net.br_from, [t[0] - 1 for t in ties]]; t = np.r_[net.br_to, [t[1] - 1 for t in ties]]
    r = np.r_[net.br_r, [t[2] / zb for t in ties]]; x = np.r_[net.br_x, [t[3] / zb for t in ties]]
    return NetDesc(base_mva=1.0, n_bus=33, slack_bus=0, slack_vm=1.0, br_from=f, br_to=t, br_r=r, br_x=x,
                   load_bus=net.load_bus, sgen_bus=net.sgen_bus, sgen_zone=net.sgen_zone, bus_zone=net.bus_zone)


def test_meshed_network_uses_the_dense_fallback():
    """pandapower solves any topology; a meshed net takes the dense-LU Newton path: same iterates, same results."""
    from oracle.pandapower_nr import PandapowerEquivalent
    net = _case33_meshed()
    inp = cases.synthetic_inputs("case33", 40, seed=12)
    q = inp["action"] * np.sqrt(inp["s_max"] ** 2 - inp["p_pv"] ** 2)
    env = _env(net, batch=1)
    assert env.dims["lanes_per_env"] == 32
    out = env.solve(inp["p_load"], inp["q_load"], inp["p_pv"], q)
    pf = PandapowerEquivalent(net)
    for e in range(40):
        r = pf.runpp(inp["p_load"][e], inp["q_load"][e], inp["p_pv"][e], q[e])
        assert r.converged and bool(out["converged"][e]) and int(out["iterations"][e]) == r.iterations
        assert np.abs(out["vm"][e].cpu().numpy() - r.vm_pu).max() < VTOL
        assert np.abs(out["va_deg"][e].cpu().numpy() - r.va_degree).max() < 1e-8
        assert np.abs(out["p_bus"][e].cpu().numpy() - r.p_mw).max() < 1e-8
        assert np.abs(out["pl"][e].cpu().numpy() - r.pl_mw).max() < 1e-9
    # a small meshed square as well
    sq = NetDesc(base_mva=1.0, n_bus=4, slack_bus=0, slack_vm=1.0, load_bus=[1, 2, 3], sgen_bus=[2], sgen_zone=[1],
                 bus_zone=[0, 1, 1, 1], br_from=[0, 1, 2, 3], br_to=[1, 2, 3, 1], br_r=[.01] * 4, br_x=[.02] * 4)
    o2 = _env(sq).solve(np.full((1, 3), 0.1), np.full((1, 3), 0.03), np.full((1, 1), 0.2), np.zeros((1, 1)))
    r2 = PandapowerEquivalent(sq).runpp(np.full(3, 0.1), np.full(3, 0.03), [0.2], [0.0])
    assert np.abs(o2["vm"][0].cpu().numpy() - r2.vm_pu).max() < VTOL


def test_topology_errors():
    from mapdn_b200._capi import MapdnError
    base = dict(base_mva=1.0, n_bus=4, slack_bus=0, slack_vm=1.0, load_bus=[1], sgen_bus=[2], sgen_zone=[1],
                bus_zone=[0, 1, 1, 1])
    island = NetDesc(br_from=[0, 2], br_to=[1, 3], br_r=[.01] * 2, br_x=[.01] * 2, **base)
    with pytest.raises(MapdnError, match="not connected"):
        _env(island)
    n = 300                                              # a large ring: meshed and beyond the dense fallback
    ring = NetDesc(base_mva=1.0, n_bus=n, slack_bus=0, slack_vm=1.0, br_from=np.arange(n), br_to=(np.arange(n) + 1) % n,
                   br_r=[.001] * n, br_x=[.001] * n, load_bus=[1], sgen_bus=[2], sgen_zone=[1], bus_zone=[0] + [1] * (n - 1))
    with pytest.raises(MapdnError, match="meshed"):
        _env(ring)


def test_degenerate_topologies():
    """2-bus net, a star (every PQ bus adjacent to the slack: forests of single-bus trees, no back sweep) and a
    pure chain, against the oracle."""
    from oracle.pandapower_nr import PandapowerEquivalent
    rng = np.random.default_rng(5)
    nets = []
    nets.append(NetDesc(base_mva=1.0, n_bus=2, slack_bus=1, slack_vm=1.01, br_from=[0], br_to=[1], br_r=[0.01], br_x=[0.02],
                        load_bus=[0], sgen_bus=[0], sgen_zone=[1], bus_zone=[1, 0]))
    n = 9
    nets.append(NetDesc(base_mva=1.0, n_bus=n, slack_bus=4, slack_vm=1.0, br_from=[4] * (n - 1),
                        br_to=[b for b in range(n) if b != 4], br_r=rng.uniform(0.005, 0.02, n - 1),
                        br_x=rng.uniform(0.005, 0.02, n - 1), load_bus=np.arange(n), sgen_bus=[0, 8], sgen_zone=[1, 2],
                        bus_zone=[1, 1, 1, 1, 0, 2, 2, 2, 2]))
    n = 40
    nets.append(NetDesc(base_mva=1.0, n_bus=n, slack_bus=0, slack_vm=1.0, br_from=np.arange(n - 1), br_to=np.arange(1, n),
                        br_r=rng.uniform(0.001, 0.004, n - 1), br_x=rng.uniform(0.001, 0.004, n - 1),
                        load_bus=np.arange(1, n), sgen_bus=[n - 1, n // 2], sgen_zone=[1, 1], bus_zone=[0] + [1] * (n - 1)))
    for net in nets:
        for lanes in (0, 4, 32):
            env = _env(net, lanes_per_env=lanes)
            B = 5
            pl = rng.uniform(0.0, 0.05, (B, net.n_load)); ql = 0.3 * pl
            pv = rng.uniform(0.0, 0.2, (B, net.n_sgen)); q = rng.uniform(-0.05, 0.05, (B, net.n_sgen))
            out = env.solve(pl, ql, pv, q)
            pf = PandapowerEquivalent(net)
            for e in range(B):
                r = pf.runpp(pl[e], ql[e], pv[e], q[e])
                assert r.converged and bool(out["converged"][e]) and int(out["iterations"][e]) == r.iterations
                assert np.abs(out["vm"][e].cpu().numpy() - r.vm_pu).max() < VTOL
                assert np.abs(out["va_deg"][e].cpu().numpy() - r.va_degree).max() < 1e-8
                assert np.abs(out["p_bus"][e].cpu().numpy() - r.p_mw).max() < 1e-8
                assert np.abs(out["pl"][e].cpu().numpy() - r.pl_mw).max() < 1e-9
            env.close()


def test_high_degree_hub():
    """A bus with six children (the > 2 children path of the sweeps: extras gathered from shared memory) and a
    second hub further down, all lane counts."""
    from oracle.pandapower_nr import PandapowerEquivalent
    rng = np.random.default_rng(9)
    f, t = [0], [1]                        # slack 0 - hub 1
    nxt = 2
    for k in range(6):                     # six feeders of three buses off the hub
        f += [1, nxt, nxt + 1]; t += [nxt, nxt + 1, nxt + 2]; nxt += 3
    hub2 = 4                               # a second hub on the first feeder: four more leaves
    for k in range(4):
        f.append(hub2); t.append(nxt); nxt += 1
    n = nxt
    net = NetDesc(base_mva=1.0, n_bus=n, slack_bus=0, slack_vm=1.0, br_from=f, br_to=t,
                  br_r=rng.uniform(0.002, 0.01, n - 1), br_x=rng.uniform(0.002, 0.01, n - 1), load_bus=np.arange(1, n),
                  sgen_bus=[3, 10, n - 1], sgen_zone=[1, 1, 1], bus_zone=[0] + [1] * (n - 1))
    pf = PandapowerEquivalent(net)
    B = 6
    pl = rng.uniform(0.0, 0.06, (B, n - 1)); ql = 0.4 * pl
    pv = rng.uniform(0.0, 0.3, (B, 3)); q = rng.uniform(-0.1, 0.1, (B, 3))
    for lanes in (4, 8, 32, 64):
        env = _env(net, lanes_per_env=lanes)
        out = env.solve(pl, ql, pv, q)
        for e in range(B):
            r = pf.runpp(pl[e], ql[e], pv[e], q[e])
            assert r.converged and bool(out["converged"][e]) and int(out["iterations"][e]) == r.iterations
            assert np.abs(out["vm"][e].cpu().numpy() - r.vm_pu).max() < VTOL
            assert np.abs(out["pl"][e].cpu().numpy() - r.pl_mw).max() < 1e-9
        env.close()
