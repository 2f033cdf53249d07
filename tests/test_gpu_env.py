"""Parity of the fused env step (reset / step / obs / state / reward / info) with the oracle's
restatement of reference voltage_control_env.py. Reward tolerance of BASELINE.json: 1e-5; held to 1e-9."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, random_tree_net
from mapdn_b200 import cases
from mapdn_b200.network import ProfileDesc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")
TOL = 1e-9


def _make(net, prof, args, batch, **kw):
    from mapdn_b200.env import BatchedVoltageControl
    return BatchedVoltageControl(net, prof, args, batch=batch, **kw)


def _cmp_step(env, oracles, ids, a, add_noise=True):
    from oracle.voltage_control_ref import INFO_KEYS
    r, term, info = env.step(torch.tensor(a, device=env.device), add_noise=add_noise)
    st = env.get_state()
    obs2 = env.obs.clone()
    obs3 = env.get_obs().clone()                     # standalone get_obs kernel == fused output
    torch.cuda.synchronize()
    assert torch.equal(obs2, obs3)
    for o, i in zip(oracles, ids):
        ro, to, io = o.step(a[i], add_noise=add_noise)
        assert abs(ro - r[i].item()) < TOL
        assert to == bool(term[i].item())
        for j, k in enumerate(INFO_KEYS):
            assert abs(io[k] - info[i, j].item()) < TOL, k
        assert np.abs(np.array(o.get_obs()) - obs2[i].cpu().numpy()).max() < TOL
        assert np.abs(o.get_state() - st[i].cpu().numpy()).max() < 1e-8


def test_golden_trajectory_case33():
    from oracle.voltage_control_ref import INFO_KEYS
    g = np.load(os.path.join(GOLD, "traj_case33.npz"))
    net, prof = cases.make_case("case33"), cases.make_profiles("case33", n_days=4)
    env = _make(net, prof, dict(voltage_barrier_type="bowl", action_scale=0.8, seed=11), batch=6)
    obs, state = env.reset()
    ids = g["env_ids"]
    torch.cuda.synchronize()
    assert env.get_field("start_row")[ids, 0].cpu().tolist() == g["start"].tolist()
    assert np.abs(obs[ids].cpu().numpy() - g["obs0"]).max() < TOL
    assert np.abs(state[ids].cpu().numpy() - g["state0"]).max() < 1e-8
    for t in range(6):
        a = np.zeros((6, net.n_sgen)); a[ids] = g["actions"][t]
        r, term, info = env.step(torch.tensor(a, device=env.device))
        torch.cuda.synchronize()
        assert np.abs(r[ids].cpu().numpy() - g["reward"][:, t]).max() < TOL
        assert np.abs(info[ids].cpu().numpy() - g["info"][:, t]).max() < TOL
        assert np.abs(env.obs[ids].cpu().numpy() - g["obs"][:, t]).max() < TOL
        assert np.abs(env.get_state()[ids].cpu().numpy() - g["state"][:, t]).max() < 1e-8
    assert len(INFO_KEYS) == info.shape[1]


@pytest.mark.parametrize("name,barrier,lanes", [("case33", "bowl", 0), ("case33", "bump", 32), ("case33", "l2", 4),
                                                 ("case141", "l1", 0), ("case322", "courant_beltrami", 0),
                                                 ("case141", "bowl", 64), ("case322", "l2", 128), ("case33", "l1", 64),
                                                 ("case322", "bowl", 0)])      # BASELINE.json config 5
@pytest.mark.parametrize("add_noise", [True, False])
def test_trajectory_matches_oracle(name, barrier, lanes, add_noise):
    from oracle.voltage_control_ref import VoltageControlOracle
    net, prof = cases.make_case(name), cases.make_profiles(name, n_days=4)
    args = dict(voltage_barrier_type=barrier, action_scale=cases.SCENARIOS[name]["action_scale"], seed=5)
    B = 21
    env = _make(net, prof, args, batch=B, lanes_per_env=lanes, env_id_offset=1000)
    ids = [0, 7, B - 1] if name == "case322" else [0, 3, 7, 12, 16, B - 1]      # the case322 oracle is the slow part
    oracles = [VoltageControlOracle(net, prof, env.args, env_id=1000 + i) for i in ids]
    if add_noise:
        obs, state = env.reset()
        for o, i in zip(oracles, ids):
            oo, os_ = o.reset()
            assert np.abs(np.array(oo) - obs[i].cpu().numpy()).max() < TOL
            assert np.abs(os_ - state[i].cpu().numpy()).max() < 1e-8
    else:
        start = torch.tensor([[i % 3, (5 * i) % 24, i % 20] for i in range(B)], dtype=torch.int32, device=env.device)
        obs, state = env.reset(start, add_noise=False)
        for o, i in zip(oracles, ids):
            oo, os_ = o.reset(start=tuple(start[i].tolist()), add_noise=False)
            assert np.abs(np.array(oo) - obs[i].cpu().numpy()).max() < TOL
    rng = np.random.default_rng(3)
    for t in range(5):
        a = rng.uniform(env.action_space.low, env.action_space.high, (B, env.n_agents))
        _cmp_step(env, oracles, ids, a, add_noise)
    # second episode: the episode counter re-keys the RNG
    env.reset()
    for o in oracles:
        o.reset()
    a = rng.uniform(env.action_space.low, env.action_space.high, (B, env.n_agents))
    _cmp_step(env, oracles, ids, a, True)


def test_line_weight_reward_and_general_net():
    """line_weight branch of the reward (reference :612-613) on a net with transformers (res_line excludes
    them), shunts, scaling factors and a non-zero slack angle."""
    from oracle.voltage_control_ref import VoltageControlOracle
    net = random_tree_net(23, 4, seed=11)
    rng = np.random.default_rng(2)
    T = 3 * 480 + 1
    prof = ProfileDesc(pv=rng.uniform(0.1, 0.5, (T, net.n_sgen)), load_p=rng.uniform(0.0, 0.3, (T, net.n_load)),
                       load_q=rng.uniform(0.0, 0.1, (T, net.n_load)), steps_per_hour=20, n_days=3)
    args = dict(voltage_barrier_type="l1", line_weight=0.7, q_weight=None, action_scale=0.5, action_bias=0.1, seed=9)
    env = _make(net, prof, args, batch=5)
    oracles = [VoltageControlOracle(net, prof, env.args, env_id=i) for i in range(5)]
    obs, _ = env.reset()
    for o, i in zip(oracles, range(5)):
        assert np.abs(np.array(o.reset()[0]) - obs[i].cpu().numpy()).max() < TOL
    for t in range(4):
        a = rng.uniform(-0.4, 0.6, (5, net.n_sgen))
        _cmp_step(env, oracles, range(5), a)


def test_divergence_branch_matches_reference_semantics():
    """reference :188-196,204 - reward from the previous net - 200, destroy=1, q_loss=attempted, terminate,
    state rolled back; SURVEY Appendix B.6."""
    from oracle.voltage_control_ref import VoltageControlOracle
    net = cases.case33()
    prof = cases.make_profiles("case33", n_days=4)
    # a profile whose rows >= 60 carry a demand the feeder cannot serve
    lp = prof.load_p.copy(); lp[60:] *= 60.0
    prof = ProfileDesc(pv=prof.pv, load_p=lp, load_q=prof.load_q, steps_per_hour=20, n_days=prof.n_days)
    args = dict(voltage_barrier_type="l1", seed=1)
    B = 9
    env = _make(net, prof, args, batch=B, lanes_per_env=8)
    start = torch.tensor([[0, 2, 15 + (i % 4)] for i in range(B)], dtype=torch.int32, device=env.device)  # rows 55..58
    env.reset(start, add_noise=False)
    oracles = [VoltageControlOracle(net, prof, env.args, env_id=i) for i in range(B)]
    for o, i in zip(oracles, range(B)):
        o.reset(start=tuple(start[i].tolist()), add_noise=False)
    rng = np.random.default_rng(0)
    seen_fail = False
    for t in range(7):
        a = rng.uniform(-0.8, 0.8, (B, 6))
        alive = [i for i in range(B) if not getattr(oracles[i], "dead", False)]
        r, term, info = env.step(torch.tensor(a, device=env.device), add_noise=False)
        torch.cuda.synchronize()
        for i in alive:
            ro, to, io = oracles[i].step(a[i], add_noise=False)
            assert abs(ro - r[i].item()) < 1e-8 and to == bool(term[i].item())
            assert io["destroy"] == info[i, 10].item()
            assert abs(io["q_loss"] - info[i, 9].item()) < TOL and io["totally_controllable_ratio"] == info[i, 3].item()
            assert np.abs(np.array(oracles[i].get_obs()) - env.obs[i].cpu().numpy()).max() < TOL
            if io["destroy"] == 1.0:
                seen_fail = True
                oracles[i].dead = True
                assert ro < -200.0
    assert seen_fail


def test_masked_reset_and_episode_termination():
    net, prof = cases.make_case("case33"), cases.make_profiles("case33", n_days=4)
    env = _make(net, prof, dict(episode_limit=5, seed=2), batch=8)
    env.reset()
    a = torch.zeros(8, 6, dtype=torch.float64, device=env.device)
    terms = []
    for _ in range(4):
        _, term, _ = env.step(a)
        terms.append(term.clone())
    torch.cuda.synchronize()
    assert [int(t.sum()) for t in terms] == [0, 0, 0, 8]              # steps 2,3,4,5 -> 5 >= episode_limit
    assert env.get_field("steps")[:, 0].cpu().tolist() == [5.0] * 8
    before = env.get_field("start_row").clone()
    mask = torch.tensor([1, 0, 1, 0, 0, 0, 0, 1], dtype=torch.uint8, device=env.device)
    env.reset(mask=mask)
    torch.cuda.synchronize()
    steps = env.get_field("steps")[:, 0].cpu().numpy()
    assert steps.tolist() == [1, 5, 1, 5, 5, 5, 5, 1]
    assert torch.equal(env.get_field("start_row")[mask == 0], before[mask == 0])
    assert env.get_field("sum_rewards")[0, 0].item() == 0.0 and env.get_field("sum_rewards")[1, 0].item() != 0.0


def test_step_host_equals_device_path():
    net, prof = cases.make_case("case33"), cases.make_profiles("case33", n_days=4)
    e1 = _make(net, prof, dict(seed=4, voltage_barrier_type="bowl"), batch=32)
    e2 = _make(net, prof, dict(seed=4, voltage_barrier_type="bowl"), batch=32)
    e1.reset(); e2.reset()
    rng = np.random.default_rng(0)
    for _ in range(3):
        a = rng.uniform(-0.8, 0.8, (32, 6))
        r1, t1, i1 = e1.step(torch.tensor(a, device=e1.device))
        r2, t2, i2, o2 = e2.step_host(a)
        torch.cuda.synchronize()
        assert np.array_equal(r1.cpu().numpy(), r2) and np.array_equal(t1.cpu().numpy(), t2)
        assert np.array_equal(i1.cpu().numpy(), i2) and np.array_equal(e1.obs.cpu().numpy(), o2)
    # fp32 observations: the same values rounded once
    e3 = _make(net, prof, dict(seed=4, voltage_barrier_type="bowl"), batch=32)
    e3.reset()
    a = np.random.default_rng(0).uniform(-0.8, 0.8, (32, 6))
    e1b = _make(net, prof, dict(seed=4, voltage_barrier_type="bowl"), batch=32); e1b.reset()
    e1b.step(torch.tensor(a, device=e1b.device))
    r3, t3, i3, o3 = e3.step_host(a, obs_dtype=np.float32)
    torch.cuda.synchronize()
    assert o3.dtype == np.float32 and np.array_equal(o3, e1b.obs.cpu().numpy().astype(np.float32))
    assert np.array_equal(r3, e1b.reward.cpu().numpy())


def test_voltage_control_shim_api():
    """The reference call pattern (code_examples.py:34-58, models/model.py:204-221, utilities/tester.py:34-56)."""
    from mapdn_b200.env import VoltageControl
    from oracle.voltage_control_ref import INFO_KEYS, VoltageControlOracle
    env = VoltageControl(dict(scenario="case33", voltage_barrier_type="bowl", action_scale=0.8, action_bias=0.0,
                              mode="distributed", episode_limit=240, seed=0, data_path="unused"))
    assert env.get_num_of_agents() == 6 and env.get_total_actions() == 1
    assert env.get_obs_size() == 50 and env.get_state_size() == 144
    info = env.get_env_info()
    assert info == dict(state_shape=144, obs_shape=50, n_actions=1, n_agents=6, episode_limit=240)
    obs, state = env.manual_reset(1, 23, 2)
    assert len(obs) == 6 and obs[0].shape == (50,) and state.shape == (144,)
    net, prof = cases.make_case("case33"), cases.make_profiles("case33")
    o = VoltageControlOracle(net, prof, dict(voltage_barrier_type="bowl", action_scale=0.8, seed=0))
    o.reset()                                   # the constructor's reset() is episode 1 (reference :85)
    oo, os_ = o.reset(start=(1, 23, 2), add_noise=False)
    assert np.abs(np.array(oo) - np.array(obs)).max() < TOL and np.abs(os_ - state).max() < 1e-8
    assert env.get_avail_actions().shape == (1, 6, 1)
    for t in range(3):
        a = env.get_action()
        assert a.shape == (6,) and a.min() >= -0.8 and a.max() <= 0.8
        reward, done, info = env.step(a, add_noise=False)
        ro, to, io = o.step(a, add_noise=False)
        assert isinstance(reward, float) and isinstance(done, bool) and set(info) == set(INFO_KEYS)
        assert abs(reward - ro) < TOL and done == to
        assert np.abs(np.array(env.get_obs()) - np.array(o.get_obs())).max() < TOL
        assert np.abs(env._get_res_bus_v() - o.g.res.vm_pu).max() < TOL
        assert np.abs(env._get_sgen_active() - o.g.sgen_p).max() < 1e-12
        assert np.abs(env._get_sgen_reactive() - o.g.sgen_q).max() < 1e-12
        assert np.abs(env._get_res_line_loss() - o.g.res.pl_mw).max() < TOL
        assert np.abs(env._get_res_bus_active() - o.g.res.p_mw).max() < 1e-8
    assert env.steps == 4
    obs, state = env.reset()
    assert env.steps == 1 and env.sum_rewards == 0
    env.close()


def test_step_is_cuda_graph_capturable():
    """mapdn_step is stream-ordered (no sync, no allocation): a rollout inner loop can be captured in a CUDA graph."""
    net, prof = cases.make_case("case33"), cases.make_profiles("case33", n_days=4)
    e1 = _make(net, prof, dict(seed=8, voltage_barrier_type="bowl"), batch=128)
    e2 = _make(net, prof, dict(seed=8, voltage_barrier_type="bowl"), batch=128)
    e1.reset(); e2.reset()
    a = torch.zeros(128, 6, dtype=torch.float64, device=e1.device).uniform_(-0.8, 0.8)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        e1.step(a)                                   # warm-up on the side stream
    torch.cuda.current_stream().wait_stream(s)
    e2.step(a)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(4):
            e1.step(a)
    g.replay()
    for _ in range(4):
        e2.step(a)
    torch.cuda.synchronize()
    assert torch.equal(e1.reward, e2.reward) and torch.equal(e1.obs, e2.obs)
    assert torch.equal(e1.get_field("steps"), e2.get_field("steps"))


def test_shim_history_and_reset_time():
    """history > 1 stacking (reference :303-315) and reset(reset_time=False) keeping the episode start (:110-113)."""
    from mapdn_b200.env import VoltageControl
    env = VoltageControl(dict(scenario="case33", voltage_barrier_type="l1", history=3, seed=1, data_path="unused"))
    assert env.get_obs_size() == 3 * 50
    o1 = env.get_obs()
    assert o1[0].shape == (150,)
    env.step(np.zeros(6))
    o2 = env.get_obs()
    assert np.array_equal(o2[0][50:100], o1[0][100:150])          # the previous newest frame moved back one slot
    start = (env._episode_start_day, env._episode_start_hour, env._episode_start_interval)
    env.reset(reset_time=False)
    assert (env._episode_start_day, env._episode_start_hour, env._episode_start_interval) == start
    env.reset()
    env.close()


def test_meshed_env_trajectory_matches_oracle():
    """Full env step (reward / obs / next row / divergence-free path) on a meshed feeder = dense fallback solver."""
    from oracle.voltage_control_ref import VoltageControlOracle
    from test_gpu_solve import _case33_meshed
    net, prof = _case33_meshed(), cases.make_profiles("case33", n_days=4)
    env = _make(net, prof, dict(voltage_barrier_type="bowl", seed=6), batch=7)
    ids = [0, 3, 6]
    oracles = [VoltageControlOracle(net, prof, env.args, env_id=i) for i in ids]
    obs, state = env.reset()
    for o, i in zip(oracles, ids):
        oo, os_ = o.reset()
        assert np.abs(np.array(oo) - obs[i].cpu().numpy()).max() < TOL and np.abs(os_ - state[i].cpu().numpy()).max() < 1e-8
    rng = np.random.default_rng(1)
    for t in range(3):
        _cmp_step(env, oracles, ids, rng.uniform(-0.8, 0.8, (7, 6)))


@pytest.mark.parametrize("ss", [["vm_pu"], ["pv", "reactive"], ["demand", "va_degree"], ["vm_pu", "va_degree", "pv"]])
def test_state_space_subsets(ss):
    """state_space selects the blocks of obs / state (reference :78, :217-228, :254-266)."""
    from oracle.voltage_control_ref import VoltageControlOracle
    net, prof = cases.make_case("case33"), cases.make_profiles("case33", n_days=4)
    env = _make(net, prof, dict(state_space=ss, seed=2), batch=5)
    nz = 12
    assert env.obs_size == nz * (2 * ("demand" in ss) + ("vm_pu" in ss) + ("va_degree" in ss)) + ("pv" in ss) + ("reactive" in ss)
    assert env.state_size == 33 * (2 * ("demand" in ss) + ("vm_pu" in ss) + ("va_degree" in ss)) + 6 * (("pv" in ss) + ("reactive" in ss))
    oracles = [VoltageControlOracle(net, prof, env.args, env_id=i) for i in range(5)]
    obs, state = env.reset()
    for o, i in zip(oracles, range(5)):
        oo, os_ = o.reset()
        assert np.abs(np.array(oo) - obs[i].cpu().numpy()).max() < TOL and np.abs(os_ - state[i].cpu().numpy()).max() < 1e-8
    rng = np.random.default_rng(0)
    for t in range(2):
        _cmp_step(env, oracles, range(5), rng.uniform(-0.8, 0.8, (5, 6)))


def test_batched_history_stacking():
    net, prof = cases.make_case("case33"), cases.make_profiles("case33", n_days=4)
    env = _make(net, prof, dict(history=3, seed=1, episode_limit=4), batch=6)
    env.reset()
    o0 = env.obs.clone()
    st = env.get_obs_stacked()
    assert st.shape == (6, 6, 150) and torch.equal(st[:, :, 100:], o0) and float(st[:, :, :100].abs().max()) == 0.0
    a = torch.zeros(6, 6, dtype=torch.float64, device=env.device)
    env.step(a); o1 = env.obs.clone()
    st = env.get_obs_stacked()
    assert torch.equal(st[:, :, 50:100], o0) and torch.equal(st[:, :, 100:], o1)
    env.step(a); _, done, _ = env.step(a)
    assert bool(done.all())
    mask = torch.tensor([1, 0, 0, 1, 0, 0], dtype=torch.uint8, device=env.device)
    env.reset(mask=mask)
    st = env.get_obs_stacked()
    assert float(st[0, :, :100].abs().max()) == 0.0 and float(st[1, :, :100].abs().max()) > 0.0


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_calls_leave_the_current_device_alone():
    """A handle bound to cuda:1 works while the caller's current device stays cuda:0 (every C-ABI entry point
    restores the device it found)."""
    net, prof = cases.make_case("case33"), cases.make_profiles("case33", n_days=4)
    torch.cuda.set_device(0)
    e0 = _make(net, prof, dict(seed=2), batch=8)
    from mapdn_b200.env import BatchedVoltageControl
    e1 = BatchedVoltageControl(net, prof, dict(seed=2), batch=8, device=1)
    assert torch.cuda.current_device() == 0
    e0.reset(); e1.reset()
    a = torch.full((8, 6), 0.3, dtype=torch.float64)
    r0, _, _ = e0.step(a.to("cuda:0"))
    r1, _, _ = e1.step(a.to("cuda:1"))
    assert torch.cuda.current_device() == 0
    assert torch.equal(r0.cpu(), r1.cpu())
    e1.close()
    assert torch.cuda.current_device() == 0


def _barrier_torch(kind, v):
    """Reference voltage_barrier/{l1,l2,bowl}.py restated on tensors (independent of the oracle and of the kernel)."""
    d = (v - 1.0).abs()
    if kind == "l1":
        return d
    if kind == "l2":
        return 2.0 * (v - 1.0) ** 2
    pdf = torch.exp(-0.5 * ((v - 1.0) / 0.1) ** 2) / (0.1 * (2.0 * torch.pi) ** 0.5)
    return torch.where(d > 0.05, 2.0 * d - 0.095, -0.01 * pdf + 0.04)


@pytest.mark.parametrize("name,batch", [("case33", 4096), ("case141", 2048), ("case322", 1024)])
def test_full_size_step_properties(name, batch):
    """BASELINE.json configs 1-3 at full batch: the fused step's outputs satisfy the reference's definitions when
    recomputed from the solved voltages with plain tensor code (reference _clip_reactive_power :568-572,
    _calc_reward :574-623), and the trajectories do not depend on how the batch is sharded."""
    sc = cases.SCENARIOS[name]
    net, prof = cases.make_case(name), cases.make_profiles(name)
    args = dict(voltage_barrier_type=sc["barrier"], action_scale=sc["action_scale"], seed=11)
    env = _make(net, prof, args, batch=batch)
    half = [_make(net, prof, args, batch=batch // 2), _make(net, prof, args, batch=batch // 2, env_id_offset=batch // 2)]
    env.reset()
    for h in half:
        h.reset()
    s_max = torch.tensor(prof.s_max, device=env.device)
    g = torch.Generator(device="cpu").manual_seed(3)
    for _ in range(3):
        a = (torch.rand(batch, env.n_agents, generator=g, dtype=torch.float64) * 2 - 1) * sc["action_scale"]
        a = a.to(env.device)
        p_before = env.get_field("p_sgen")
        r, term, info = env.step(a)
        vm, q = env.get_field("vm"), env.get_field("q_sgen")
        ok = info[:, 10] == 0                                     # destroy flag: the reward definition below holds
        assert ok.float().mean().item() > 0.99
        assert torch.allclose(q, a * torch.sqrt(s_max ** 2 - p_before ** 2), rtol=0, atol=1e-12)
        want = -(0.1 * q.abs().mean(dim=1) + _barrier_torch(sc["barrier"], vm).mean(dim=1))
        assert (r - want)[ok].abs().max().item() < 1e-10
        out = ((vm < 0.95) | (vm > 1.05)).double().mean(dim=1)
        assert (info[:, 0] - out)[ok].abs().max().item() < 1e-12
        assert (info[:, 5] - vm.mean(dim=1))[ok].abs().max().item() < 1e-12
        assert (info[:, 8] - env.get_field("line_loss").sum(dim=1))[ok].abs().max().item() < 1e-12
        assert (info[:, 9] - q.abs().mean(dim=1))[ok].abs().max().item() < 1e-12
        assert not term.any()
        rh = torch.cat([h.step(a[i * (batch // 2):(i + 1) * (batch // 2)].contiguous())[0] for i, h in enumerate(half)])
        assert torch.equal(rh, r)
        assert torch.equal(torch.cat([h.obs for h in half]), env.obs)


def test_handles_with_different_footprints_coexist():
    """The dynamic shared-memory limit is a per-function attribute: creating a handle with a smaller per-CTA footprint
    (smaller batch -> fewer envs per CTA) must not break the launches of an older, larger one."""
    net, prof = cases.make_case("case33"), cases.make_profiles("case33", n_days=4)
    big = _make(net, prof, dict(seed=1), batch=4096)
    small = _make(net, prof, dict(seed=1), batch=8)
    assert big.dims["smem_bytes"] > small.dims["smem_bytes"]
    big.reset(); small.reset()
    a = torch.zeros(4096, 6, dtype=torch.float64, device=big.device)
    r_big, _, _ = big.step(a)
    r_small, _, _ = small.step(a[:8].contiguous())
    torch.cuda.synchronize()
    assert torch.equal(r_big[:8], r_small)             # same env ids, same seed: the CTA shape does not change results
