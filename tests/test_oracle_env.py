"""Pins the oracle's restatement of the reference env logic (voltage_control_env.py) with
hand-written expectations: barrier tables, a 5-bus / 2-zone net whose observation, state and
reward are written out by hand following SURVEY Appendix B, and the sequencing quirks."""
import math

import numpy as np
import pytest

from mapdn_b200.network import NetDesc, ProfileDesc
from oracle.voltage_control_ref import (INFO_KEYS, VOLTAGE_BARRIER, VoltageControlOracle, bowl, bump,
                                        courant_beltrami, l1, l2)


def test_barrier_tables():
    v = [0.90, 0.96, 1.0, 1.04, 1.10]
    assert np.allclose(l1(v), [0.10, 0.04, 0.0, 0.04, 0.10])
    assert np.allclose(l2(v), [0.02, 0.0032, 0.0, 0.0032, 0.02])
    assert np.allclose(courant_beltrami(v), [0.0025, 0, 0, 0, 0.0025])
    pdf = lambda x: math.exp(-0.5 * (x - 1) ** 2 / 0.01) / math.sqrt(2 * math.pi * 0.01)
    assert np.allclose(bowl(v), [2 * 0.10 - 0.095, -0.01 * pdf(0.96) + 0.04, -0.01 * pdf(1.0) + 0.04,
                                 -0.01 * pdf(1.04) + 0.04, 2 * 0.10 - 0.095])
    # bowl is continuous-ish at the 0.05 switch: 2*0.05-0.095 = 0.005 vs -0.01*pdf(1.05)+0.04
    assert abs(bowl([1.05 + 1e-12])[0] - 0.005) < 1e-9
    assert np.allclose(bump([0.5, 1.0, 1.5, 2.0, 3.5]),
                       [math.exp(-1 / (1 - 0.5 ** 4)), 0.0, math.exp(-1 / (1 - 0.5 ** 4)), math.exp(-1.0), 0.0])
    assert set(VOLTAGE_BARRIER) == {"l1", "l2", "bowl", "bump", "courant_beltrami"}


def five_bus():
    """slack 0 - 1 - 2(zone1, PV a) ; 1 - 3(zone2) - 4(zone2, PVs b and c on bus 4)."""
    net = NetDesc(base_mva=1.0, n_bus=5, slack_bus=0, slack_vm=1.0,
                  br_from=[0, 1, 1, 3], br_to=[1, 2, 3, 4], br_r=[0.01, 0.02, 0.015, 0.02],
                  br_x=[0.02, 0.03, 0.02, 0.03], load_bus=[1, 2, 3, 4], sgen_bus=[2, 4, 4],
                  sgen_zone=[1, 2, 2], bus_zone=[0, 0, 1, 2, 2])
    T = 481 * 2
    t = np.arange(T)
    pv = np.stack([0.3 + 0.001 * t, 0.2 + 0.0005 * t, 0.1 + 0.0002 * t], 1)
    lp = np.stack([0.2 + 0.0001 * t, 0.3 + 0 * t, 0.25 + 0.0002 * t, 0.1 + 0 * t], 1)
    prof = ProfileDesc(pv=pv, load_p=lp, load_q=0.3 * lp, steps_per_hour=20, n_days=2)
    return net, prof


def test_five_bus_obs_state_reward_by_hand():
    net, prof = five_bus()
    cfg = dict(voltage_barrier_type="l1", episode_limit=10, reset_action=False, action_scale=0.8)
    env = VoltageControlOracle(net, prof, cfg)
    obs, state = env.reset(start=(0, 1, 2), add_noise=False)       # start row = 2 + 1*20 = 22
    assert env.start == 22 and env.steps == 1
    row1 = 23                                                      # row 1 of the window (row 0 never used)
    assert np.allclose(env.g.sgen_p, prof.pv[row1]) and np.allclose(env.g.load_p, prof.load_p[row1])
    assert env.obs_dim == 4 * 2 + 2
    a = np.array([0.5, -0.25, 0.1])
    pv_old = env.g.sgen_p.copy()
    reward, term, info = env.step(a, add_noise=False)
    # B.1: the first step re-loads row 1; B.2: q uses the old pv
    assert np.allclose(env.g.sgen_p, prof.pv[row1]) and env.steps == 2
    q = a * np.sqrt(prof.s_max ** 2 - pv_old ** 2)
    assert np.allclose(env.g.sgen_q, q)
    res = env.g.res
    # hand-built reward: all 5 buses incl. slack in the barrier mean
    v = res.vm_pu
    assert abs(reward + (0.1 * np.mean(np.abs(q)) + np.mean(np.abs(v - 1.0)))) < 1e-15
    assert set(info) == set(INFO_KEYS) and info["destroy"] == 0.0 and not term
    # res_bus.p_mw = load - sgen(old) on PQ buses; slack row = -(ext grid infeed)
    assert np.allclose(res.p_mw[1:], [prof.load_p[row1, 0], prof.load_p[row1, 1] - pv_old[0],
                                      prof.load_p[row1, 2], prof.load_p[row1, 3] - pv_old[1] - pv_old[2]])
    assert abs(res.p_mw[0] + res.p_ext_mw) < 1e-15
    # second step loads row 2 -> obs mixes old-solve bus P with the NEW pv (B.3)
    reward, term, info = env.step(a, add_noise=False)
    res, pv_new = env.g.res, prof.pv[row1 + 1]
    obs = env.get_obs()
    o0 = np.array([res.p_mw[2] + pv_new[0], res.q_mvar[2] + env.g.sgen_q[0], pv_new[0], env.g.sgen_q[0],
                   res.vm_pu[2], np.deg2rad(res.va_degree[2]), 0, 0, 0, 0])
    assert np.allclose(obs[0], o0, atol=1e-15)
    # zone2 = buses 3,4; both PVs of the zone sit on bus 4 and both are added back for each agent
    pz = [res.p_mw[3], res.p_mw[4] + pv_new[1] + pv_new[2]]
    qz = [res.q_mvar[3], res.q_mvar[4] + env.g.sgen_q[1] + env.g.sgen_q[2]]
    for ag in (1, 2):
        exp = np.array(pz + qz + [pv_new[ag], env.g.sgen_q[ag]] + list(res.vm_pu[3:5]) +
                       list(np.deg2rad(res.va_degree[3:5])))
        assert np.allclose(obs[ag], exp, atol=1e-15)
    st = env.get_state()
    assert st.shape == (4 * 5 + 2 * 3,)
    assert np.allclose(st, np.r_[res.p_mw, res.q_mvar, pv_new, env.g.sgen_q, res.vm_pu, res.va_degree])


def test_termination_and_failure_branch():
    net, prof = five_bus()
    env = VoltageControlOracle(net, prof, dict(episode_limit=4, reset_action=False))
    env.reset(start=(0, 0, 0), add_noise=False)
    terms = [env.step(np.zeros(3), add_noise=False)[1] for _ in range(3)]
    assert terms == [False, False, True]                          # steps 2,3,4 -> 4 >= episode_limit
    # failure: blow the demand up so that the power flow diverges
    env = VoltageControlOracle(net, prof, dict(episode_limit=10, reset_action=False))
    env.reset(start=(0, 0, 0), add_noise=False)
    r_ok, _, info_ok = env.step(np.zeros(3), add_noise=False)
    v_prev = env.g.res.vm_pu.copy()
    env.g.load_p = env.g.load_p * 500.0
    a = np.array([0.3, 0.3, 0.3])
    q_try = a * np.sqrt(prof.s_max ** 2 - env.g.sgen_p ** 2)
    r, term, info = env.step(a, add_noise=False)
    assert term and info["destroy"] == 1.0 and info["totally_controllable_ratio"] == 0.0
    assert abs(info["q_loss"] - np.mean(np.abs(q_try))) < 1e-15
    assert np.array_equal(env.g.res.vm_pu, v_prev)                # rolled back to the previous net
    assert np.allclose(env.g.sgen_q, 0.0)
    assert abs(r - (-(0.1 * 0.0 + np.mean(np.abs(v_prev - 1.0))) - 200.0)) < 1e-12


def test_noise_is_half_normal_and_keyed():
    net, prof = five_bus()
    e0 = VoltageControlOracle(net, prof, dict(seed=3), env_id=0)
    e1 = VoltageControlOracle(net, prof, dict(seed=3), env_id=1)
    e0b = VoltageControlOracle(net, prof, dict(seed=3), env_id=0)
    for e in (e0, e1, e0b):
        e.reset(start=(0, 2, 3), add_noise=True)
    assert np.array_equal(e0.g.load_p, e0b.g.load_p) and not np.array_equal(e0.g.load_p, e1.g.load_p)
    row = e0.start + 1
    assert np.all(e0.g.sgen_p >= prof.pv[row]) and np.all(e0.g.load_q >= prof.load_q[row])
