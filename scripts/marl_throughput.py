"""Learner-inclusive throughput of the batched runner: B lock-step envs, a GRU actor + MLP critic with the
reference's model interface, one critic update on a device batch per lock-step (the reference updates once per 60
transitions of ONE env).   python scripts/marl_throughput.py [B] [steps]"""
import sys
import time
from collections import namedtuple
import torch
sys.path.insert(0, ".")
from mapdn_b200 import cases
from mapdn_b200.env import BatchedVoltageControl
from mapdn_b200.marl_runner import BatchedMarlRunner, DeviceTransitionBuffer, attach

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 60


class ActorCritic(torch.nn.Module):
    def __init__(self, n, obs_dim, hid, dev):
        super().__init__()
        self.args = namedtuple("A", "max_steps action_scale action_bias hid_size num_eval_episodes gamma")(T, 0.8, 0.0, hid, 4, 0.99)
        self.n, self.hid = n, hid
        self.fc1 = torch.nn.Linear(obs_dim + n, hid); self.rnn = torch.nn.GRUCell(hid, hid); self.fc2 = torch.nn.Linear(hid, 1)
        self.q = torch.nn.Sequential(torch.nn.Linear((obs_dim + 1) * n, hid), torch.nn.ReLU(), torch.nn.Linear(hid, hid),
                                     torch.nn.ReLU(), torch.nn.Linear(hid, 1))
        self.policy_dicts = [self]
        self.eye = torch.eye(n, device=dev)
        self.to(dev)

    def init_hidden(self):
        return self.fc1.weight.new_zeros(1, self.n, self.hid)

    def get_actions(self, state, status, exploration, actions_avail, target=False, last_hid=None):
        Bn = state.shape[0]
        x = torch.cat([state, self.eye.expand(Bn, -1, -1)], dim=-1).reshape(Bn * self.n, -1)
        h = self.rnn(torch.relu(self.fc1(x)), last_hid.reshape(-1, self.hid))
        mean = self.fc2(h).view(Bn, self.n, 1)
        act = torch.tanh(mean + (0.3 * torch.randn_like(mean) if exploration else 0.0))
        return act, act, torch.zeros_like(act), mean, h.view(Bn, self.n, self.hid)

    def value(self, obs, act):
        Bn = obs.shape[0]
        return self.q(torch.cat([obs.reshape(Bn, -1), act.reshape(Bn, -1)], dim=-1)).view(Bn, 1, 1).expand(-1, self.n, -1)

    def unpack_data(self, batch):
        raise AssertionError


net, prof = cases.make_case("case33"), cases.make_profiles("case33")
env = BatchedVoltageControl(net, prof, dict(voltage_barrier_type="bowl", seed=0), batch=B)
model = attach(ActorCritic(env.n_agents, env.obs_size, 64, env.device))
buf = DeviceTransitionBuffer(64, B, env.n_agents, env.obs_size, 1, 64, env.device)
opt = torch.optim.RMSprop(model.q.parameters(), lr=1e-4)


def update(runner, stat):
    if runner.buffer.count < 2:
        return
    st, ac, _, _, nv, rw, ns, dn, *_ = model.unpack_data(runner.buffer.get_batch(32, n_windows=32))      # 1024 transitions
    target = rw[:, :, None] + model.args.gamma * (1 - dn[:, :, None]) * nv
    loss = (model.value(st, ac) - target.detach()).pow(2).mean()
    opt.zero_grad(); loss.backward(); opt.step()


for label, fn in (("rollout only (actor + critic forward, no update)", None), ("with one critic update per lock-step", update)):
    runner = BatchedMarlRunner(env, model, buf, update_fn=fn)
    runner.train_process({})                       # warm-up
    torch.cuda.synchronize(); t0 = time.perf_counter()
    runner.steps = 0
    stat = runner.train_process({})
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"B={B} T={T} {label}: {runner.steps / dt / 1e6:.2f} M env-steps/s ({dt / T * 1e3:.2f} ms per lock-step), "
          f"mean_train_reward {stat['mean_train_reward']:.4f}")
