"""Turns "parity unpinned" into pinned: regenerates golden vectors from the REAL pandapower.

    python scripts/pin_with_pandapower.py [--out tests/golden]

Needs `import pandapower` (the reference pins 2.7.0, environment.yml:133) - not available in this image, so this
script has not been run yet; it is committed so that whoever has pandapower can pin the oracle and the CUDA path with
one command. For every net below it builds the pandapower net (oracle/pp_bridge.py), runs `pp.runpp(net)` with
default arguments - the reference's call, voltage_control_env.py:124,165,557 - on seeded element values and writes
tests/golden/pp_<name>.npz (inputs + res_bus / res_line columns + iteration counts + the pandapower version).
tests/test_pandapower_pins.py then holds the oracle (CPU) and the CUDA path (-m gpu) to these files:
voltages to 1e-6 p.u. (north-star tolerance; expected agreement ~1e-9), identical convergence flags.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.append(os.path.join(ROOT, "baseline", "_ref"))


def nets():
    from mapdn_b200 import cases
    from conftest import random_tree_net
    out = {}
    for name in ("case33", "case141", "case322"):
        inp = cases.synthetic_inputs(name, 8, seed=2025)
        q = inp["action"] * np.sqrt(inp["s_max"] ** 2 - inp["p_pv"] ** 2)
        out[name] = (cases.make_case(name), inp["p_load"], inp["q_load"], inp["p_pv"], q)
    net, p, q = cases.baran_wu_nominal()
    out["baran_wu"] = (net, p[None], q[None], np.zeros((1, 6)), np.zeros((1, 6)))
    # taps + bus shunts + scaling (no charging on the tap branches, slack angle 0: representable in pandapower tables)
    net = random_tree_net(19, 3, seed=5)
    net.br_b[net.br_tap != 1.0] = 0.0
    net.br_g[net.br_tap != 1.0] = 0.0
    net.br_b[net.br_is_line == 0] = 0.0
    net.br_g[net.br_is_line == 0] = 0.0
    net.slack_va_deg = 0.0
    rng = np.random.default_rng(3)
    out["taps19"] = (net, rng.uniform(0, 0.3, (6, net.n_load)), rng.uniform(0, 0.1, (6, net.n_load)),
                     rng.uniform(0, 0.5, (6, net.n_sgen)), rng.uniform(-0.2, 0.2, (6, net.n_sgen)))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    args = ap.parse_args()
    try:
        import pandapower as pp
    except Exception as ex:
        raise SystemExit(f"pandapower is not importable here ({ex}); nothing written - parity stays unpinned")
    from oracle.pp_bridge import PandapowerBackend
    for name, (net, pl, ql, pv, q) in nets().items():
        be = PandapowerBackend(net)
        keys = ("vm_pu", "va_degree", "p_mw", "q_mvar", "pl_mw")
        cols = {k: [] for k in keys}
        it, cv = [], []
        for e in range(pl.shape[0]):
            r = be.runpp(pl[e], ql[e], pv[e], q[e])
            for k in keys:
                cols[k].append(getattr(r, k))
            it.append(r.iterations); cv.append(r.converged)
        np.savez_compressed(os.path.join(args.out, f"pp_{name}.npz"), p_load=pl, q_load=ql, p_pv=pv, q=q,
                            iterations=np.array(it, np.int32), converged=np.array(cv, np.uint8),
                            pandapower_version=np.array(pp.__version__), **{k: np.array(v) for k, v in cols.items()})
        print(f"pp_{name}.npz: {pl.shape[0]} solves, iterations {it}")


if __name__ == "__main__":
    main()
