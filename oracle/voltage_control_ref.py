"""ORACLE - TEST INFRASTRUCTURE ONLY. CPU restatement of MAPDN's ``VoltageControl`` env
(reference ``environments/var_voltage_control/voltage_control_env.py``) on top of
``oracle.pandapower_nr.PandapowerEquivalent`` - one env, fp64, plain NumPy.

Restated with the *intended* pandas-1.1.3 semantics of ``get_obs`` (the chained in-place ``+=``
at reference :239-244 mutates the zone frames there; under pandas 3 it is a no-op).
Sequencing quirks follow SURVEY Appendix B. Randomness comes from ``oracle.philox_ref`` (the
reference's global ``np.random`` stream cannot be reproduced by construction).

PINNED: ``tests/test_reference_golden.py`` replays trajectories that the reference's own, unmodified
``VoltageControl`` code produced in this container (``oracle/ref_harness.py``: substitute pandapower whose ``runpp`` is
``PandapowerEquivalent``, pandas-1.1.3 table semantics, ``np.random`` fed with the Philox draws) and this class
reproduces them to 1e-11 - reset incl. retries, noisy and noise-free steps, all barriers, ``line_weight``,
``state_space`` subsets, the divergence branch.
"""
from __future__ import annotations

import numpy as np

from . import philox_ref as rng
from .pandapower_nr import PandapowerEquivalent

INFO_KEYS = ("percentage_of_v_out_of_control", "percentage_of_lower_than_lower_v",
             "percentage_of_higher_than_upper_v", "totally_controllable_ratio",
             "average_voltage_deviation", "average_voltage", "max_voltage_drop_deviation",
             "max_voltage_rise_deviation", "total_line_loss", "q_loss", "destroy")


# --- voltage barriers: literal transcriptions of reference voltage_barrier/*.py --------------
def l1(vs, v_ref=1.0):                       # l1.py:5-8
    return np.array([np.abs(v - v_ref) for v in vs])


def l2(vs, v_ref=1.0):                       # l2.py:5-8
    return np.array([2 * np.square(v - v_ref) for v in vs])


def bowl(vs, v_ref=1.0, scale=.1):           # bowl.py:5-13
    def normal(v, loc, scale):
        return 1 / np.sqrt(2 * np.pi * scale ** 2) * np.exp(- 0.5 * np.square(v - loc) / scale ** 2)

    def _bowl(v):
        if np.abs(v - v_ref) > 0.05:
            return 2 * np.abs(v - v_ref) - 0.095
        else:
            return - 0.01 * normal(v, v_ref, scale) + 0.04
    return np.array([_bowl(v) for v in vs])


def bump(vs):                                # bump.py:5-13
    def _bump(v):
        if np.abs(v) < 1:
            return np.exp(- 1 / (1 - v ** 4))
        elif 1 < v < 3:
            return np.exp(- 1 / (1 - (v - 2) ** 4))
        else:
            return 0.0
    return np.array([_bump(v) for v in vs])


def courant_beltrami(vs, v_lower=0.95, v_upper=1.05):   # courant_beltrami.py:5-8
    return np.array([np.square(max(0, v - v_upper)) + np.square(max(0, v_lower - v)) for v in vs])


VOLTAGE_BARRIER = dict(l1=l1, l2=l2, bowl=bowl, bump=bump, courant_beltrami=courant_beltrami)


class _NetState:
    """The mutable slice of the pandapower net the env touches."""
    __slots__ = ("load_p", "load_q", "sgen_p", "sgen_q", "res")

    def copy(self):
        c = _NetState()
        c.load_p, c.load_q = self.load_p.copy(), self.load_q.copy()
        c.sgen_p, c.sgen_q = self.sgen_p.copy(), self.sgen_q.copy()
        c.res = self.res
        return c


class VoltageControlOracle:
    """One environment. ``cfg`` keys as in reference args/env_args/var_voltage_control.yaml."""

    def __init__(self, net, profiles, cfg, env_id=0):
        self.net, self.prof = net, profiles
        self.cfg = dict(voltage_barrier_type="l1", voltage_weight=1.0, q_weight=0.1, line_weight=None,
                        v_upper=1.05, v_lower=0.95, episode_limit=240, history=1, action_scale=0.8,
                        action_bias=0.0, reset_action=True, seed=0,
                        state_space=["pv", "demand", "reactive", "vm_pu", "va_degree"])
        self.cfg.update(cfg)
        c = self.cfg
        self.env_id = env_id
        self.pf = PandapowerEquivalent(net)
        self.episode_limit = c["episode_limit"]
        self.v_upper, self.v_lower = c["v_upper"], c["v_lower"]
        self.pv_std, self.active_std, self.reactive_std = profiles.pv_std, profiles.load_p_std, profiles.load_q_std
        self.s_max = profiles.s_max                                   # :515-521
        self.low = -c["action_scale"] + c["action_bias"]              # :76
        self.high = c["action_scale"] + c["action_bias"]
        self.n_agents = net.n_sgen
        self.barrier = VOLTAGE_BARRIER[c["voltage_barrier_type"]]
        self.episode = 0
        self.zones = [net.zone_buses(i) for i in range(net.n_sgen)]
        ss = self.cfg["state_space"]
        self.obs_dim = max(len(z) * (2 * ("demand" in ss) + ("vm_pu" in ss) + ("va_degree" in ss)) + ("pv" in ss)
                           + ("reactive" in ss) for z in self.zones)

    # ---- profile access (:440-513) --------------------------------------------------------
    def _row(self, t):
        return self.start + t

    def _set_demand_and_pv(self, add_noise, c1):
        r = self._row(self.steps)
        pv = self.prof.pv[r].copy()
        lp = self.prof.load_p[r].copy()
        lq = self.prof.load_q[r].copy()
        if add_noise:
            ng, nl = self.net.n_sgen, self.net.n_load
            z = rng.half_normal(self.cfg["seed"], self.env_id, self.episode, c1, np.arange(ng + 2 * nl))
            pv += self.pv_std * z[:ng]
            lp += self.active_std * z[ng:ng + nl]
            lq += self.reactive_std * z[ng + nl:]
        self.g.sgen_p, self.g.load_p, self.g.load_q = pv, lp, lq

    def _clip_reactive_power(self, a, p):                             # :568-572
        return np.sqrt(self.s_max ** 2 - p ** 2) * a

    def _runpp(self):
        res = self.pf.runpp(self.g.load_p, self.g.load_q, self.g.sgen_p, self.g.sgen_q)
        if res.converged:
            self.g.res = res
        return res.converged

    # ---- reset (:96-176) ------------------------------------------------------------------
    def reset(self, start=None, add_noise=True):
        """start=None: sample (hour, day, interval); else (day, hour, interval) like manual_reset."""
        self.steps = 1
        self.sum_rewards = 0.0
        self.episode += 1
        spd_h = self.prof.steps_per_hour
        episode_days = self.episode_limit // (24 * spd_h) + 1
        attempt = 0
        self.g = _NetState()
        self.g.sgen_q = np.zeros(self.net.n_sgen)
        self.g.res = None
        while True:
            if start is None:
                hour, day, interval = rng.start_time(self.cfg["seed"], self.env_id, self.episode, attempt,
                                                     self.prof.n_days - episode_days, spd_h)
            else:
                day, hour, interval = start
            self.start = interval + hour * spd_h + day * 24 * spd_h
            self._set_demand_and_pv(add_noise, rng.RESET_FLAG | attempt)
            if self.cfg["reset_action"]:
                a = rng.uniform_action(self.cfg["seed"], self.env_id, self.episode, attempt,
                                       self.net.n_sgen, self.low, self.high)
                self.g.sgen_q = self._clip_reactive_power(a, self.g.sgen_p)
            if self._runpp():
                break
            attempt += 1
            if attempt >= 16:
                raise RuntimeError("reset: power flow does not converge")
        return self.get_obs(), self.get_state()

    # ---- step (:178-211) ------------------------------------------------------------------
    def step(self, actions, add_noise=True):
        last = self.g.copy()
        self.g.sgen_q = self._clip_reactive_power(np.asarray(actions, np.float64), self.g.sgen_p)
        solvable = self._runpp()
        if solvable:
            reward, info = self._calc_reward()
        else:
            q_loss = np.mean(np.abs(self.g.sgen_q))
            self.g = last
            reward, info = self._calc_reward()
            reward -= 200.
            info["destroy"] = 1.
            info["totally_controllable_ratio"] = 0.
            info["q_loss"] = q_loss
        self._set_demand_and_pv(add_noise, self.steps)
        self.steps += 1
        self.sum_rewards += reward
        terminated = bool(self.steps >= self.episode_limit or not solvable)
        return reward, terminated, info

    # ---- reward (:574-623) ----------------------------------------------------------------
    def _calc_reward(self):
        info = {}
        res = self.g.res
        v = res.vm_pu
        lo, hi = np.sum(v < self.v_lower), np.sum(v > self.v_upper)
        pct = (lo + hi) / v.shape[0]
        info["percentage_of_v_out_of_control"] = pct
        info["percentage_of_lower_than_lower_v"] = lo / v.shape[0]
        info["percentage_of_higher_than_upper_v"] = hi / v.shape[0]
        info["totally_controllable_ratio"] = 0. if pct > 1e-3 else 1.
        v_ref = 0.5 * (self.v_lower + self.v_upper)
        info["average_voltage_deviation"] = np.mean(np.abs(v - v_ref))
        info["average_voltage"] = np.mean(v)
        info["max_voltage_drop_deviation"] = np.max((v < self.v_lower) * (self.v_lower - v))
        info["max_voltage_rise_deviation"] = np.max((v > self.v_upper) * (v - self.v_upper))
        line_loss = np.sum(res.pl_mw)
        avg_line_loss = np.mean(res.pl_mw)
        info["total_line_loss"] = line_loss
        q = self.g.sgen_q * self.net.sgen_scaling                    # res_sgen.q_mvar
        q_loss = np.mean(np.abs(q))
        info["q_loss"] = q_loss
        v_loss = np.mean(self.barrier(v)) * self.cfg["voltage_weight"]
        if self.cfg["line_weight"] is not None:
            loss = avg_line_loss * self.cfg["line_weight"] + v_loss
        elif self.cfg["q_weight"] is not None:
            loss = q_loss * self.cfg["q_weight"] + v_loss
        else:
            raise NotImplementedError
        info["destroy"] = 0.0
        return -loss, info

    # ---- observations (:213-316, :523-546) --------------------------------------------------
    def get_state(self):                                              # :213-230
        r, ss = self.g.res, self.cfg["state_space"]
        state = []
        if "demand" in ss:
            state += list(r.p_mw) + list(r.q_mvar)
        if "pv" in ss:
            state += list(self.g.sgen_p)
        if "reactive" in ss:
            state += list(self.g.sgen_q)
        if "vm_pu" in ss:
            state += list(r.vm_pu)
        if "va_degree" in ss:
            state += list(r.va_degree)
        return np.array(state)

    def get_obs(self):
        r, net = self.g.res, self.net
        obs = []
        for i in range(net.n_sgen):
            zb = self.zones[i]
            p = r.p_mw[zb].copy()
            q = r.q_mvar[zb].copy()
            # :238-244 - every sgen of the same zone adds its p/q back on its own bus row
            for j in range(net.n_sgen):
                if net.sgen_zone[j] == net.sgen_zone[i]:
                    k = np.nonzero(zb == net.sgen_bus[j])[0]
                    p[k] += self.g.sgen_p[j]
                    q[k] += self.g.sgen_q[j]
            ss, o = self.cfg["state_space"], []                       # :254-266
            if "demand" in ss:
                o += list(p) + list(q)
            if "pv" in ss:
                o.append(self.g.sgen_p[i])
            if "reactive" in ss:
                o.append(self.g.sgen_q[i])
            if "vm_pu" in ss:
                o += list(r.vm_pu[zb])
            if "va_degree" in ss:
                o += list(r.va_degree[zb] * np.pi / 180)
            o = np.array(o)
            obs.append(np.concatenate([o, np.zeros(self.obs_dim - o.shape[0])]))
        return obs
