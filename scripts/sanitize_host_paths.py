"""compute-sanitizer driver: a few steps of every host path (padded zero-copy, compact direct, compact DMA, fp32) on the
three scenarios at small batches, plus reset and droop; prints a checksum."""
import sys
import numpy as np
sys.path.insert(0, ".")
from mapdn_b200 import cases
from mapdn_b200.env import BatchedVoltageControl

tot = 0.0
for sc, B in (("case33", 30), ("case141", 9), ("case322", 5)):
    net, prof = cases.make_case(sc), cases.make_profiles(sc, n_days=3)
    env = BatchedVoltageControl(net, prof, dict(seed=3), batch=B)
    env.reset()
    rng = np.random.default_rng(1)
    for t in range(2):
        a = rng.uniform(-0.8, 0.8, (B, env.n_agents))
        for kw in (dict(), dict(layout="compact"), dict(layout="compact", staged=False),
                   dict(layout="compact", obs_dtype=np.float32), dict(obs_dtype=np.float32), dict(staged=True)):
            r, d, i, o = env.step_host(a, **kw)
            tot += float(r.sum()) + float(o.sum())
    env.close()
print("done", tot)
