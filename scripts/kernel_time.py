"""Step-kernel time of one BASELINE.json config (CUDA events per step, L2 flushed between steps, 200 steps):
    python scripts/kernel_time.py case33 [batch] [lanes] [barrier]
Much cheaper than a full bench.py run (no e2e / CPU legs); used by scripts/ab_run.sh and scripts/bench3.sh."""
import sys
import torch
sys.path.insert(0, ".")
from mapdn_b200 import cases
from mapdn_b200.env import BatchedVoltageControl

name = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) > 0 else {"case33": 4096, "case141": 2048, "case322": 1024}[name]
G = int(sys.argv[3]) if len(sys.argv) > 3 else 0
barrier = sys.argv[4] if len(sys.argv) > 4 else cases.SCENARIOS[name]["barrier"]
net, prof = cases.make_case(name), cases.make_profiles(name)
env = BatchedVoltageControl(net, prof, dict(voltage_barrier_type=barrier, action_scale=cases.SCENARIOS[name]["action_scale"]),
                            batch=B, lanes_per_env=G)
dev = env.device
g = torch.Generator(device=dev); g.manual_seed(1234)
lo, hi = env.action_space.low, env.action_space.high
acts = lo + (hi - lo) * torch.rand(8, B, env.n_agents, dtype=torch.float64, device=dev, generator=g)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
env.reset()
K, W = 200, 10
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
us = sum(ts) / K * 1e3
print(f"{name} B={B} G={env.dims['lanes_per_env']} epb={env.dims['envs_per_block']} smem={env.dims['smem_bytes']}: "
      f"{us:.2f} us/step (median {ts[K // 2] * 1e3:.2f}) {B / us:.2f} M env-steps/s")
