#!/usr/bin/env python
"""Writes tests/golden/ref_env_*.npz: trajectories produced by the REFERENCE's own env code
(/root/reference/environments/var_voltage_control/voltage_control_env.py, unmodified, imported at run time) for the
scenarios of oracle/ref_scenarios.py. Only runs where /root/reference exists (this container); pandapower is substituted
as described in oracle/ref_harness.py (the Newton-Raphson behind pp.runpp is oracle/pandapower_nr.py - these fixtures pin
the ENV LOGIC, the literature KATs and scripts/pin_with_pandapower.py pin the power flow).

    python scripts/make_reference_golden.py [scenario ...]
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as H            # noqa: E402
from oracle import ref_scenarios as S          # noqa: E402


def main():
    names = sys.argv[1:] or list(S.SCENARIOS)
    for name in names:
        sc = S.SCENARIOS[name]
        res = record(sc)
        path = S.fixture_path(ROOT, name)
        np.savez_compressed(path, **res)
        print(f"{name}: {os.path.getsize(path) / 1024:.1f} kB, {res['reward'].shape[0]} steps x {len(sc['env_ids'])} envs")


def record(sc, view_rows=True):
    """Runs the reference env for every env id of the scenario. Arrays: obs [n_ops, E, n_agents, obs_dim],
    state [n_ops, E, state_dim] (one snapshot after every operation), reward / term [n_steps, E], info [n_steps, E, 11],
    alive [n_steps, E] (0 once an env has terminated: the reference would be reset by its caller), start [n_resets, E, 3]
    = (day, hour, interval) the reference used, actions [n_steps, E, n_agents]."""
    net, prof = sc["build"]()
    ids = sc["env_ids"]
    E, n_steps = len(ids), S.n_steps_of(sc)
    obs, state, start = [], [], []
    reward, term = np.zeros((n_steps, E)), np.zeros((n_steps, E), np.uint8)
    info, alive = np.zeros((n_steps, E, 11)), np.ones((n_steps, E), np.uint8)
    with tempfile.TemporaryDirectory() as d:
        H.write_reference_data(d, net, prof)
        runs = [H.ReferenceRun(d, net, sc["args"], env_id=e, view_rows=view_rows) for e in ids]
        lo, hi = runs[0].env.action_space.low, runs[0].env.action_space.high
        acts = S.action_stream(sc["name"], n_steps, E, net.n_sgen, lo, hi)
        dead = [False] * E
        t = 0
        for k_op, op in enumerate(sc["ops"]):
            o_all, s_all = [], []
            for k, r in enumerate(runs):
                if op[0] == "init":
                    o, s = r.initial()
                elif op[0] == "reset":
                    o, s = r.reset()
                elif op[0] == "manual":
                    o, s = r.manual_reset(*S.manual_of(sc, op, k))
                elif op[0] == "reset_keep":
                    o, s = r.reset_keep_time()
                elif dead[k]:
                    alive[t, k] = 0
                    o, s = obs[-1][k], state[-1][k]
                else:
                    reward[t, k], tm, info[t, k], o, s = r.step(acts[t, k], add_noise=op[1])
                    term[t, k] = tm
                    dead[k] = tm
                o_all.append(o); s_all.append(s)
            if op[0] == "step":
                t += 1
            else:
                start.append([r.start[0] for r in runs])
                dead = [False] * E
            obs.append(np.array(o_all)); state.append(np.array(s_all))
    return dict(env_ids=np.array(ids), ops=np.array(json.dumps(sc["ops"])), actions=acts, obs=np.array(obs),
                state=np.array(state), reward=reward, term=term, info=info, alive=alive,
                start=np.array(start, np.int64).reshape(len(start), E, 3))


if __name__ == "__main__":
    main()
