"""Generates tests/golden/*.npz from the ORACLE (oracle/ - the NumPy restatement of pandapower
2.7.0's runpp + the reference env logic). The reference itself cannot run in this image
(pandapower absent), so these are regression pins of the oracle, not outputs of the reference;
the literature-anchored entry is `baran_wu` (IEEE-33 nominal case, published Vmin / losses).

    python scripts/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from mapdn_b200 import cases  # noqa: E402
from oracle.pandapower_nr import PandapowerEquivalent  # noqa: E402
from oracle.voltage_control_ref import INFO_KEYS, VoltageControlOracle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def solve_fixture(net, p_load, q_load, p_pv, q):
    pf = PandapowerEquivalent(net)
    keys = ("vm_pu", "va_degree", "p_mw", "q_mvar", "pl_mw")
    out = {k: [] for k in keys}
    it, cv = [], []
    for e in range(p_load.shape[0]):
        r = pf.runpp(p_load[e], q_load[e], p_pv[e], q[e])
        for k in keys:
            out[k].append(getattr(r, k))
        it.append(r.iterations)
        cv.append(r.converged)
    return dict(p_load=p_load, q_load=q_load, p_pv=p_pv, q=q, iterations=np.array(it, np.int32),
                converged=np.array(cv, np.uint8), **{k: np.array(v) for k, v in out.items()})


def main():
    os.makedirs(OUT, exist_ok=True)
    for name in ("case33", "case141", "case322"):
        net = cases.make_case(name)
        inp = cases.synthetic_inputs(name, 6, seed=2024)
        q = inp["action"] * np.sqrt(inp["s_max"] ** 2 - inp["p_pv"] ** 2)
        np.savez_compressed(os.path.join(OUT, f"solve_{name}.npz"),
                            **solve_fixture(net, inp["p_load"], inp["q_load"], inp["p_pv"], q))
    net, p, q = cases.baran_wu_nominal()
    np.savez_compressed(os.path.join(OUT, "solve_baran_wu.npz"),
                        **solve_fixture(net, p[None], q[None], np.zeros((1, 6)), np.zeros((1, 6))))
    from conftest import random_tree_net
    net = random_tree_net(23, 4, seed=11)
    rng = np.random.default_rng(0)
    np.savez_compressed(os.path.join(OUT, "solve_rand23.npz"),
                        **solve_fixture(net, rng.uniform(0, 0.3, (5, net.n_load)), rng.uniform(0, 0.1, (5, net.n_load)),
                                        rng.uniform(0, 0.5, (5, net.n_sgen)), rng.uniform(-0.2, 0.2, (5, net.n_sgen))))
    # env trajectory: case33, bowl barrier, noise on, env ids 0 and 5, 6 steps
    net, prof = cases.make_case("case33"), cases.make_profiles("case33", n_days=4)
    cfg = dict(voltage_barrier_type="bowl", action_scale=0.8, seed=11, episode_limit=240)
    rng = np.random.default_rng(1)
    acts = rng.uniform(-0.8, 0.8, (6, 2, net.n_sgen))
    tr = dict(actions=acts, env_ids=np.array([0, 5]), reward=[], info=[], obs=[], state=[], obs0=[], state0=[], start=[])
    for k, eid in enumerate((0, 5)):
        env = VoltageControlOracle(net, prof, cfg, env_id=eid)
        o, s = env.reset()
        tr["obs0"].append(np.array(o)); tr["state0"].append(s); tr["start"].append(env.start)
        rew, inf, ob, st = [], [], [], []
        for t in range(6):
            r, term, info = env.step(acts[t, k])
            rew.append(r); inf.append([info[x] for x in INFO_KEYS]); ob.append(np.array(env.get_obs())); st.append(env.get_state())
        tr["reward"].append(rew); tr["info"].append(inf); tr["obs"].append(ob); tr["state"].append(st)
    np.savez_compressed(os.path.join(OUT, "traj_case33.npz"), **{k: np.array(v) for k, v in tr.items()})
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
