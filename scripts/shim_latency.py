"""Wall-clock cost of ONE transition through the drop-in class (B = 1) - the call pattern of the reference's own loop
(`models/model.py:204-221`: env.step(actions) followed by env.get_obs()), next to the CPU restatement of the reference
path on one core of this box:
    python scripts/shim_latency.py [steps]
"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from mapdn_b200 import cases                                   # noqa: E402
from mapdn_b200.env import VoltageControl                      # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
for name in ("case33", "case141", "case322"):
    scale = cases.SCENARIOS[name]["action_scale"]
    env = VoltageControl(dict(scenario=name, voltage_barrier_type="bowl", action_scale=scale, seed=0))
    rng = np.random.default_rng(0)
    acts = rng.uniform(-scale, scale, (steps, env.n_agents))

    def run(n):
        t0 = time.perf_counter()
        for t in range(n):
            _, done, _ = env.step(acts[t])
            env.get_obs()
            if done:
                env.reset()
        return (time.perf_counter() - t0) / n

    run(20)
    dt = run(steps)
    # the CPU restatement of the same transition (oracle/: test infrastructure, here only as the yardstick)
    from oracle.voltage_control_ref import VoltageControlOracle
    o = VoltageControlOracle(cases.make_case(name), cases.make_profiles(name),
                             dict(voltage_barrier_type="bowl", action_scale=scale, seed=0))
    o.reset()
    n_cpu = max(10, min(60, steps))
    t0 = time.perf_counter()
    for t in range(n_cpu):
        _, done, _ = o.step(acts[t])
        o.get_obs()
        if done:
            o.reset()
    dt_cpu = (time.perf_counter() - t0) / n_cpu
    print(f"{name}: drop-in VoltageControl.step + get_obs (B = 1, NumPy in / out): {dt * 1e6:.1f} us per transition "
          f"({1 / dt:.0f} /s); CPU restatement on one core: {dt_cpu * 1e3:.2f} ms ({1 / dt_cpu:.0f} /s); ratio {dt_cpu / dt:.0f}x")
    env.close()
