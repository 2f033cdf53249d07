import sys, torch
sys.path.insert(0, ".")
from mapdn_b200 import cases
from mapdn_b200.env import BatchedVoltageControl
name = sys.argv[1]; G = int(sys.argv[2]); B = int(sys.argv[3])
net, prof = cases.make_case(name), cases.make_profiles(name)
env = BatchedVoltageControl(net, prof, dict(voltage_barrier_type=cases.SCENARIOS[name]["barrier"]), batch=B, lanes_per_env=G)
env.reset()
a = torch.zeros(B, env.n_agents, dtype=torch.float64, device=env.device).uniform_(-0.8, 0.8)
for _ in range(4): env.step(a)
torch.cuda.synchronize()
