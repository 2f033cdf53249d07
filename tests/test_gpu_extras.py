"""GPU tests of the widened rows: droop / no-control baselines (f4) and the batched rollout adapter (f2)."""
import numpy as np
import pytest
import torch

from mapdn_b200 import cases

pytestmark = pytest.mark.gpu


def _oracle_droop(net, pl, ql, pv, s_rated, max_ite=100, gain=0.1):
    """Literal restatement of reference traditional_control/pf_droop_matpower_all.m:121-152,196-231 (one instant)."""
    from oracle.pandapower_nr import PandapowerEquivalent
    pf = PandapowerEquivalent(net)

    def droop(p, s, v, qmm):
        q_max = min(np.sqrt(s * s - p * p), qmm)
        if v <= 0.95: return q_max
        if v > 1.05: return -q_max
        if 1.0 <= v <= 1.0: return 0.0
        if v < 1.0: return (q_max - 0) / (0.95 - 1.0) * (v - 1.0)
        return (0 - q_max) / (1.0 - 1.05) * (1.0 - v)
    q_last = np.zeros(net.n_sgen); v_last = 100 * np.ones(net.n_sgen)
    for i in range(max_ite):
        res = pf.runpp(pl, ql, pv, q_last)
        v = res.vm_pu[net.sgen_bus]
        if np.linalg.norm(v_last - v) < 1e-4:
            break
        v_last = v
        q_new = np.array([droop(pv[j], s_rated[j], v[j], s_rated[j]) for j in range(net.n_sgen)])
        q_last = (1 - gain) * q_last + gain * q_new
    return res.vm_pu, q_last, res.pl_mw.sum(), i + 1


def test_droop_and_no_control_match_the_script():
    from mapdn_b200.baselines import droop_control, no_control
    from mapdn_b200.env import BatchedVoltageControl
    from oracle.pandapower_nr import PandapowerEquivalent
    net = cases.case33()
    inp = cases.synthetic_inputs("case33", 12, seed=4)
    env = BatchedVoltageControl(net, None, None, batch=1)
    s_rated = inp["s_max"]
    out = droop_control(env, inp["p_load"], inp["q_load"], inp["p_pv"], s_rated)
    nc = no_control(env, torch.tensor(inp["p_load"], device=env.device), torch.tensor(inp["q_load"], device=env.device),
                    torch.tensor(inp["p_pv"], device=env.device))
    pf = PandapowerEquivalent(net)
    for e in (0, 5, 11):
        vm, q, loss, it = _oracle_droop(net, inp["p_load"][e], inp["q_load"][e], inp["p_pv"][e], s_rated)
        assert int(out["iterations"][e]) == it
        assert np.abs(out["vm"][e].cpu().numpy() - vm).max() < 1e-8
        assert np.abs(out["q"][e].cpu().numpy() - q).max() < 1e-8
        assert abs(float(out["loss"][e]) - loss) < 1e-8
        r0 = pf.runpp(inp["p_load"][e], inp["q_load"][e], inp["p_pv"][e], np.zeros(6))
        assert np.abs(nc["vm"][e].cpu().numpy() - r0.vm_pu).max() < 1e-9
    # droop pulls the voltages of the PV buses towards 1.0
    dev0 = (nc["vm"][:, net.sgen_bus] - 1).abs().mean()
    dev1 = (out["vm"][:, net.sgen_bus] - 1).abs().mean()
    assert float(dev1) < float(dev0)


@pytest.mark.parametrize("name,lanes", [("case33", 0), ("case141", 0), ("case322", 64)])
def test_droop_kernel_is_one_launch_and_equals_the_host_loop(name, lanes):
    """mapdn_droop (the whole relaxed loop inside the fused kernel) against the round-1 host-driven loop."""
    from mapdn_b200.baselines import droop_control, droop_control_host_loop
    from mapdn_b200.env import BatchedVoltageControl
    net = cases.make_case(name)
    B = 37
    inp = cases.synthetic_inputs(name, B, seed=9)
    env = BatchedVoltageControl(net, None, None, batch=1, lanes_per_env=lanes)
    n0 = env.launch_count
    out = droop_control(env, inp["p_load"], inp["q_load"], inp["p_pv"], inp["s_max"])
    torch.cuda.synchronize()
    assert env.launch_count - n0 == 1
    ref = droop_control_host_loop(env, inp["p_load"], inp["q_load"], inp["p_pv"], inp["s_max"])
    assert torch.equal(out["iterations"], ref["iterations"])
    assert int(out["iterations"].min()) >= 2 and int(out["iterations"].max()) < 100
    assert float((out["vm"] - ref["vm"]).abs().max()) < 1e-10
    assert float((out["q"] - ref["q"]).abs().max()) < 1e-10
    assert float((out["loss"] - ref["loss"]).abs().max()) < 1e-10
    # the iteration cap: q of the last power flow, iterations == cap
    capped = droop_control(env, inp["p_load"], inp["q_load"], inp["p_pv"], inp["s_max"], max_ite=3)
    assert int(capped["iterations"].max()) == 3
    env.close()


def test_batched_rollout_with_replay_buffer():
    from mapdn_b200.env import BatchedVoltageControl
    from mapdn_b200.rollout import BatchedRollout, DeviceReplayBuffer
    net, prof = cases.make_case("case33"), cases.make_profiles("case33", n_days=4)
    B = 64
    env = BatchedVoltageControl(net, prof, dict(episode_limit=6, voltage_barrier_type="bowl", seed=3), batch=B)
    torch.manual_seed(0)
    mlp = torch.nn.Sequential(torch.nn.Linear(env.obs_size, 32), torch.nn.Tanh(), torch.nn.Linear(32, 1)).to(env.device)
    policy = lambda obs: mlp(obs).squeeze(-1)                      # shared-parameter actor, [B, n_agents]
    buf = DeviceReplayBuffer(1000, env.n_agents, env.obs_size, env.device)
    ro = BatchedRollout(env, policy, buf)
    obs = ro.run(12)
    torch.cuda.synchronize()
    assert obs.shape == (B, env.n_agents, env.obs_size) and len(buf) == 12 * B
    rets = ro.completed_episode_returns()
    assert rets.numel() == 2 * B                                   # episode_limit 6 -> 5 steps per episode -> 2 episodes in 12 steps
    s = buf.sample(256)
    assert s["obs"].shape == (256, env.n_agents, env.obs_size) and s["done"].dtype == torch.bool
    assert torch.isfinite(s["reward"]).all() and bool((s["action"].abs() < 10).all())
    # after a masked reset the step counter restarts at 1 for every env that terminated
    steps = env.get_field("steps")[:, 0]
    assert float(steps.min()) >= 1 and float(steps.max()) <= 5
