"""New CTA-shape heuristic vs the previous fixed shape (MAPDN_EPB / MAPDN_HELPERS overrides) over batch sizes."""
import os, sys, torch
sys.path.insert(0, ".")
from mapdn_b200 import cases
from mapdn_b200.env import BatchedVoltageControl

def run(name, B, old):
    sc = cases.SCENARIOS[name]
    for k in ("MAPDN_EPB", "MAPDN_HELPERS"):
        os.environ.pop(k, None)
    if old:
        os.environ["MAPDN_EPB"], os.environ["MAPDN_HELPERS"] = str(old), "1"
    env = BatchedVoltageControl(NET[name], PROF[name], dict(voltage_barrier_type=sc["barrier"]), batch=B)
    env.reset()
    a = torch.zeros(B, env.n_agents, dtype=torch.float64, device=env.device).uniform_(-0.6, 0.6)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=env.device)
    for _ in range(3): env.step(a)
    n, tot = 20, 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); env.step(a); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    us = tot / n * 1e3
    print(f"{name} B={B} {'old' if old else 'new'} epb={env.dims['envs_per_block']}: {us:.1f} us/step {B/us:.2f} M/s", flush=True)
    env.close()

NET = {n: cases.make_case(n) for n in ("case33", "case141", "case322")}
PROF = {n: cases.make_profiles(n) for n in NET}
for B in (512, 1024, 2048, 8192, 12288):
    run("case33", B, 0); run("case33", B, 8)
for B in (512, 1024, 4096):
    run("case141", B, 0); run("case141", B, 4)
for B in (256, 512, 2048):
    run("case322", B, 0); run("case322", B, 3)
