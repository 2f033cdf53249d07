"""Droop baseline: one fused launch (mapdn_droop) against the host-driven loop of round 1 (one mapdn_solve launch +
PyTorch elementwise kernels + a host sync per iteration).   python scripts/droop_timing.py [case33] [batch]"""
import sys
import time
import torch
sys.path.insert(0, ".")
from mapdn_b200 import cases
from mapdn_b200.baselines import droop_control, droop_control_host_loop
from mapdn_b200.env import BatchedVoltageControl

name = sys.argv[1] if len(sys.argv) > 1 else "case33"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
net = cases.make_case(name)
inp = cases.synthetic_inputs(name, B, seed=1)
env = BatchedVoltageControl(net, None, None, batch=1)
t = lambda x: torch.as_tensor(x, dtype=torch.float64, device=env.device)
args = [t(inp[k]) for k in ("p_load", "q_load", "p_pv", "s_max")]
for fn in (droop_control, droop_control_host_loop):
    out = fn(env, *args)
    torch.cuda.synchronize()
    n0, t0 = env.launch_count, time.perf_counter()
    for _ in range(5):
        out = fn(env, *args)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"{name} B={B} {fn.__name__}: {dt * 1e3:.2f} ms per control instant, {(env.launch_count - n0) // 5} launches of this "
          f"library, power flows per env: mean {out['iterations'].double().mean():.1f} max {int(out['iterations'].max())}")
