"""More known-answer tests from the literature (numbers NOT produced by this repository), typed in from the printed data:

* **IEEE 69-bus radial feeder** (Baran & Wu 1989, "Optimal capacitor placement on radial distribution systems";
  12.66 kV, 68 branches, 48 loads): the data carries two checksums of its own - total demand 3802.19 kW + j2694.6 kvar -
  and the base-case results every reconfiguration / DG paper quotes: real-power loss 225.0 kW, reactive loss 102.2 kvar,
  minimum voltage 0.9092 p.u. at bus 65. 68 PQ buses: the CUDA path takes the 32-lanes-per-env instantiation of the tree
  solver (case33 takes the 8-lane one).
* **Stagg & El-Abiad 5-bus system** (*Computer Methods in Power System Analysis*, 1968; 100 MVA, slack 1.06 p.u., bus 2
  generates 40 MW + j30 Mvar as a fixed P/Q injection, lines with charging): the meshed network of the book with the
  load-flow solution reprinted in most power-system courses - |V| = 1.0474, 1.0242, 1.0236, 1.0179 p.u. at
  -2.81, -5.00, -5.33, -6.15 degrees, slack generation 129.6 MW - j7.4 Mvar. Exercises the dense-LU path with
  line charging, an sgen with non-zero P and Q, and a slack voltage other than 1.
"""
import numpy as np
import pytest

from mapdn_b200.network import NetDesc
from oracle.pandapower_nr import PandapowerEquivalent

# from, to, R [ohm], X [ohm]
_B69_BRANCH = """
1 2 .0005 .0012|2 3 .0005 .0012|3 4 .0015 .0036|4 5 .0251 .0294|5 6 .3660 .1864|6 7 .3811 .1941|7 8 .0922 .0470|
8 9 .0493 .0251|9 10 .8190 .2707|10 11 .1872 .0619|11 12 .7114 .2351|12 13 1.0300 .3400|13 14 1.0440 .3450|
14 15 1.0580 .3496|15 16 .1966 .0650|16 17 .3744 .1238|17 18 .0047 .0016|18 19 .3276 .1083|19 20 .2106 .0690|
20 21 .3416 .1129|21 22 .0140 .0046|22 23 .1591 .0526|23 24 .3463 .1145|24 25 .7488 .2475|25 26 .3089 .1021|
26 27 .1732 .0572|3 28 .0044 .0108|28 29 .0640 .1565|29 30 .3978 .1315|30 31 .0702 .0232|31 32 .3510 .1160|
32 33 .8390 .2816|33 34 1.7080 .5646|34 35 1.4740 .4873|3 36 .0044 .0108|36 37 .0640 .1565|37 38 .1053 .1230|
38 39 .0304 .0355|39 40 .0018 .0021|40 41 .7283 .8509|41 42 .3100 .3623|42 43 .0410 .0478|43 44 .0092 .0116|
44 45 .1089 .1373|45 46 .0009 .0012|4 47 .0034 .0084|47 48 .0851 .2083|48 49 .2898 .7091|49 50 .0822 .2011|
8 51 .0928 .0473|51 52 .3319 .1114|9 53 .1740 .0886|53 54 .2030 .1034|54 55 .2842 .1447|55 56 .2813 .1433|
56 57 1.5900 .5337|57 58 .7837 .2630|58 59 .3042 .1006|59 60 .3861 .1172|60 61 .5075 .2585|61 62 .0974 .0496|
62 63 .1450 .0738|63 64 .7105 .3619|64 65 1.0410 .5302|11 66 .2012 .0611|66 67 .0047 .0014|12 68 .7394 .2444|
68 69 .0047 .0016"""
# bus, P [kW], Q [kvar]
_B69_LOAD = """
6 2.6 2.2|7 40.4 30|8 75 54|9 30 22|10 28 19|11 145 104|12 145 104|13 8 5.5|14 8 5.5|16 45.5 30|17 60 35|18 60 35|
20 1 .6|21 114 81|22 5.3 3.5|24 28 20|26 14 10|27 14 10|28 26 18.6|29 26 18.6|33 14 10|34 19.5 14|35 6 4|
36 26 18.55|37 26 18.55|39 24 17|40 24 17|41 1.2 1|43 6 4.3|45 39.22 26.3|46 39.22 26.3|48 79 56.4|49 384.7 274.5|
50 384.7 274.5|51 40.5 28.3|52 3.6 2.7|53 4.35 3.5|54 26.4 19|55 24 17.2|59 100 72|61 1244 888|62 32 23|64 227 162|
65 59 42|66 18 13|67 18 13|68 28 20|69 28 20"""


def _rows(text):
    return [r.split() for r in text.replace("\n", "").split("|")]


def ieee69():
    """(net, p_load_mw, q_load_mvar): 1 MVA base, one (idle) sgen at the weakest bus so that the env shape is valid."""
    br, ld = _rows(_B69_BRANCH), _rows(_B69_LOAD)
    zb = 12.66 ** 2 / 1.0
    f = np.array([int(b[0]) - 1 for b in br]); t = np.array([int(b[1]) - 1 for b in br])
    r = np.array([float(b[2]) for b in br]) / zb; x = np.array([float(b[3]) for b in br]) / zb
    zone = np.ones(69, np.int32); zone[0] = 0
    net = NetDesc(base_mva=1.0, n_bus=69, slack_bus=0, slack_vm=1.0, br_from=f, br_to=t, br_r=r, br_x=x,
                  load_bus=np.array([int(l[0]) - 1 for l in ld]), sgen_bus=np.array([64]), sgen_zone=np.array([1]),
                  bus_zone=zone, name="ieee69")
    return net, np.array([float(l[1]) for l in ld]) * 1e-3, np.array([float(l[2]) for l in ld]) * 1e-3


def stagg5():
    """(net, p_load_mw, q_load_mvar, p_sgen_mw, q_sgen_mvar); line charging given as y/2 per end in the book."""
    br = [(1, 2, .02, .06, .030), (1, 3, .08, .24, .025), (2, 3, .06, .18, .020), (2, 4, .06, .18, .020),
          (2, 5, .04, .12, .015), (3, 4, .01, .03, .010), (4, 5, .08, .24, .025)]
    net = NetDesc(base_mva=100.0, n_bus=5, slack_bus=0, slack_vm=1.06, br_from=np.array([b[0] - 1 for b in br]),
                  br_to=np.array([b[1] - 1 for b in br]), br_r=np.array([b[2] for b in br]),
                  br_x=np.array([b[3] for b in br]), br_b=np.array([2.0 * b[4] for b in br]),
                  load_bus=np.array([1, 2, 3, 4]), sgen_bus=np.array([1]), sgen_zone=np.array([1]),
                  bus_zone=np.array([0, 1, 1, 1, 1]), name="stagg5")
    return net, np.array([20., 45., 40., 60.]), np.array([10., 15., 5., 10.]), np.array([40.]), np.array([30.])


STAGG_VM = np.array([1.06, 1.0474, 1.0242, 1.0236, 1.0179])
STAGG_VA = np.array([0.0, -2.81, -5.00, -5.33, -6.15])


def _check69(vm, loss_kw, loss_kvar=None):
    assert abs(vm.min() - 0.9092) < 0.5e-4 + 1e-6 and int(np.argmin(vm)) == 64          # bus 65
    assert abs(loss_kw - 225.0) < 0.05
    if loss_kvar is not None:
        assert abs(loss_kvar - 102.2) < 0.05


def _check_stagg(vm, va_deg, p_slack=None, q_slack=None):
    assert np.abs(vm - STAGG_VM).max() < 0.5e-4 + 1e-6
    assert np.abs(va_deg - STAGG_VA).max() < 0.5e-2 + 1e-6
    if p_slack is not None:
        assert abs(p_slack - 129.6) < 0.05 and abs(q_slack + 7.4) < 0.05


# ------------------------------------------------------------------ oracle (CPU) ---------------------------------
def test_ieee69_data_checksums():
    net, p, q = ieee69()
    assert len(net.br_from) == 68 and len(p) == 48
    assert abs(p.sum() * 1e3 - 3802.19) < 1e-6 and abs(q.sum() * 1e3 - 2694.6) < 1e-6


def test_oracle_ieee69_published_base_case():
    net, p, q = ieee69()
    r = PandapowerEquivalent(net).runpp(p, q, np.zeros(1), np.zeros(1))
    assert r.converged
    _check69(r.vm_pu, (r.p_ext_mw - p.sum()) * 1e3, (r.q_ext_mvar - q.sum()) * 1e3)


def test_c_oracle_ieee69_published_base_case():
    from oracle import c_oracle
    net, p, q = ieee69()
    vm, _, conv, _ = c_oracle.COracle(net).runpp(p, q, np.zeros(1), np.zeros(1))
    assert conv and abs(vm.min() - 0.9092) < 0.5e-4 + 1e-6 and int(np.argmin(vm)) == 64


def test_oracle_stagg_el_abiad_5_bus():
    net, p, q, ps, qs = stagg5()
    r = PandapowerEquivalent(net).runpp(p, q, ps, qs)
    assert r.converged
    _check_stagg(r.vm_pu, np.rad2deg(np.angle(r.V)), r.p_ext_mw, r.q_ext_mvar)


def test_c_oracle_stagg_el_abiad_5_bus():
    from oracle import c_oracle
    net, p, q, ps, qs = stagg5()
    vm, va, conv, _ = c_oracle.COracle(net).runpp(p, q, ps, qs)
    assert conv
    _check_stagg(vm, va)                                  # COracle.runpp returns degrees


# ------------------------------------------------------------------ CUDA path ------------------------------------
def _solve_gpu(net, p, q, ps, qs, **cfg):
    import torch
    from mapdn_b200.env import BatchedVoltageControl
    env = BatchedVoltageControl(net, None, dict(voltage_barrier_type="l1"), batch=1, **cfg)
    out = env.solve(p[None], q[None], ps[None], qs[None])
    torch.cuda.synchronize()
    res = {k: (None if v is None else v.cpu().numpy()[0]) for k, v in out.items()}
    env.close()
    return res


@pytest.mark.gpu
def test_gpu_ieee69_published_base_case():
    net, p, q = ieee69()
    out = _solve_gpu(net, p, q, np.zeros(1), np.zeros(1))
    assert out["converged"] == 1
    _check69(out["vm"], out["pl"].sum() * 1e3)
    assert abs(-out["p_bus"][0] * 1e3 - (3802.19 + 225.0)) < 0.05            # res_bus at the slack = -infeed
    assert abs(-out["q_bus"][0] * 1e3 - (2694.6 + 102.2)) < 0.05
    # and to rounding error against the oracle
    r = PandapowerEquivalent(net).runpp(p, q, np.zeros(1), np.zeros(1))
    assert np.abs(out["vm"] - r.vm_pu).max() < 1e-9 and int(out["iterations"]) == r.iterations


@pytest.mark.gpu
def test_gpu_stagg_el_abiad_5_bus_meshed():
    net, p, q, ps, qs = stagg5()
    out = _solve_gpu(net, p, q, ps, qs)
    assert out["converged"] == 1
    _check_stagg(out["vm"], out["va_deg"], -out["p_bus"][0], -out["q_bus"][0])
    r = PandapowerEquivalent(net).runpp(p, q, ps, qs)
    assert np.abs(out["vm"] - r.vm_pu).max() < 1e-9
