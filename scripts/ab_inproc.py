"""A/B of two builds of the library in ONE process (one `import torch`): step-kernel time of the BASELINE.json single-GPU
configs, alternating runs, CUDA events per step with the L2 flushed in between; then (optionally) pytest on the in-tree build.
    python scripts/ab_inproc.py scripts/variants/base.so [--pytest tests/test_gpu_solve.py ...]
"""
import os
import sys
import time

t_start = time.time()
import torch                                                     # noqa: E402

sys.path.insert(0, ".")
from mapdn_b200 import _capi, cases                              # noqa: E402
from mapdn_b200.env import BatchedVoltageControl                 # noqa: E402

other = os.path.abspath(sys.argv[1])
cur = os.path.abspath("mapdn_b200/libmapdn_b200.so")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
print(f"import + setup {time.time() - t_start:.1f} s", flush=True)


def use(path):
    _capi._lib = None
    _capi.LIB_PATH = path


def measure(name, B, K=150, W=8):
    net, prof = cases.make_case(name), cases.make_profiles(name)
    env = BatchedVoltageControl(net, prof, dict(voltage_barrier_type=cases.SCENARIOS[name]["barrier"],
                                                action_scale=cases.SCENARIOS[name]["action_scale"]), batch=B)
    g = torch.Generator(device=env.device); g.manual_seed(1234)
    lo, hi = env.action_space.low, env.action_space.high
    acts = lo + (hi - lo) * torch.rand(8, B, env.n_agents, dtype=torch.float64, device=env.device, generator=g)
    env.reset()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    for i in range(W):
        env.step(acts[i % 8]); flush.zero_()
    torch.cuda.synchronize()
    for i in range(K):
        if (i + W) % 200 == 199:
            env.reset()
        ev0[i].record(); env.step(acts[i % 8]); ev1[i].record(); flush.zero_()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in zip(ev0, ev1))
    it = float(env.get_field("nr_iters").mean())
    env.close()
    return sum(ts) / K * 1e3, ts[K // 2] * 1e3, it


for name, B in (("case33", 4096), ("case322", 1024), ("case141", 2048)):
    for rep in range(1 if name == "case141" else 2):
        for tag, path in (("base", other), ("new ", cur)):
            use(path)
            us, med, it = measure(name, B)
            print(f"{name} x {B} {tag}: {us:.2f} us/step (median {med:.2f}), mean NR iterations {it:.3f}", flush=True)
use(cur)
if "--pytest" in sys.argv:
    import pytest
    args = sys.argv[sys.argv.index("--pytest") + 1:]
    print(f"timing done at {time.time() - t_start:.1f} s; pytest {' '.join(args)}", flush=True)
    rc = pytest.main(args)
    print(f"pytest rc={int(rc)} at {time.time() - t_start:.1f} s", flush=True)
