"""Kernel duration vs batch size (single-warp latency vs throughput regime)."""
import sys, torch
sys.path.insert(0, ".")
from mapdn_b200 import cases
from mapdn_b200.env import BatchedVoltageControl
name = sys.argv[1] if len(sys.argv) > 1 else "case33"
net, prof = cases.make_case(name), cases.make_profiles(name)
for G in [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "8,16").split(",")]:
    for B in (max(1, 32 // G), 128, 1024, 4096, 16384, 65536, 262144):
        try:
            env = BatchedVoltageControl(net, prof, dict(voltage_barrier_type=cases.SCENARIOS[name]["barrier"]), batch=B, lanes_per_env=G)
        except Exception as e:
            print(name, G, B, "failed", e); continue
        env.reset()
        a = torch.zeros(B, env.n_agents, dtype=torch.float64, device=env.device).uniform_(-0.8, 0.8)
        for _ in range(3): env.step(a)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 30 if B <= 16384 else 8
        ev0.record()
        for _ in range(n): env.step(a)
        ev1.record(); torch.cuda.synchronize()
        us = ev0.elapsed_time(ev1) / n * 1e3
        print(f"{name} G={G} B={B}: {us:.1f} us/step  {B/us:.1f} M env-steps/s  epb={env.dims['envs_per_block']} smem={env.dims['smem_bytes']}")
        env.close()
