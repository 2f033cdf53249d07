"""Tiny workload for compute-sanitizer (memcheck / racecheck): reset + 2 steps + solve on a few envs."""
import sys, torch
sys.path.insert(0, ".")
from mapdn_b200 import cases
from mapdn_b200.env import BatchedVoltageControl
name, G, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
net, prof = cases.make_case(name), cases.make_profiles(name, n_days=3)
env = BatchedVoltageControl(net, prof, dict(voltage_barrier_type="bowl"), batch=B, lanes_per_env=G)
env.reset()
a = torch.zeros(B, env.n_agents, dtype=torch.float64, device=env.device).uniform_(-0.8, 0.8)
for _ in range(2):
    env.step(a)
inp = cases.synthetic_inputs(name, B, seed=0)
env.solve(inp["p_load"], inp["q_load"], inp["p_pv"], inp["action"] * 0.1)
torch.cuda.synchronize()
print("done", float(env.reward.sum()))
