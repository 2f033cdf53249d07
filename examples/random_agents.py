"""Random-agent rollouts through the drop-in ``VoltageControl`` (the call pattern of the reference's
``code_examples.py``) and, next to it, the same number of env-steps through the batched API.

    python examples/random_agents.py [case33|case141|case322]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mapdn_b200 import cases                                   # noqa: E402
from mapdn_b200.env import BatchedVoltageControl, VoltageControl  # noqa: E402


def single_env(scenario, n_episodes=3, max_steps=100):
    env = VoltageControl(dict(scenario=scenario, mode="distributed", voltage_barrier_type="l1",
                              action_scale=cases.SCENARIOS[scenario]["action_scale"], action_bias=0.0,
                              data_path="unused", seed=0))
    n_agents, n_actions = env.get_num_of_agents(), env.get_total_actions()
    t0 = time.perf_counter()
    for e in range(n_episodes):
        env.reset()
        episode_reward = 0.0
        for _ in range(max_steps):
            env.get_obs(); env.get_state()
            actions = np.concatenate([np.random.normal(0, 0.5, n_actions)[np.nonzero(env.get_avail_agent_actions(a))[0]]
                                      for a in range(n_agents)])
            reward, done, info = env.step(np.clip(actions, env.action_space.low, env.action_space.high))
            episode_reward += reward
        print(f"[B=1 shim] total reward in episode {e} = {episode_reward:.2f}")
    dt = time.perf_counter() - t0
    print(f"[B=1 shim] {n_episodes * max_steps / dt:.0f} env-steps/s (one env, NumPy in/out, host sync every call)")
    env.close()


def batched(scenario, batch=4096, steps=100):
    net, prof = cases.make_case(scenario), cases.make_profiles(scenario)
    env = BatchedVoltageControl(net, prof, dict(voltage_barrier_type="l1",
                                                action_scale=cases.SCENARIOS[scenario]["action_scale"]), batch=batch)
    env.reset()
    ret = torch.zeros(batch, dtype=torch.float64, device=env.device)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        a = torch.randn(batch, env.n_agents, dtype=torch.float64, device=env.device).mul_(0.5)
        a.clamp_(env.action_space.low, env.action_space.high)
        r, done, info = env.step(a)
        ret += r
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"[batched] {batch} envs x {steps} steps: mean return {ret.mean().item():.2f}, "
          f"{batch * steps / dt / 1e6:.1f} M env-steps/s including the action sampling (eager launches)")
    # the same inner loop captured once in a CUDA graph (mapdn_step is stream-ordered and allocation-free)
    a = torch.zeros(batch, env.n_agents, dtype=torch.float64, device=env.device)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        a.normal_(0, 0.5).clamp_(env.action_space.low, env.action_space.high)
        env.step(a)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            a.normal_(0, 0.5).clamp_(env.action_space.low, env.action_space.high)
            env.step(a)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps // 10):
        g.replay()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"[batched, CUDA graph of 10 steps] {batch * (steps // 10) * 10 / dt / 1e6:.1f} M env-steps/s")
    env.close()


if __name__ == "__main__":
    sc = sys.argv[1] if len(sys.argv) > 1 else "case33"
    single_env(sc)
    batched(sc)
