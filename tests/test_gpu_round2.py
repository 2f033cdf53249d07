"""GPU tests of the round-2 additions: zero-copy host path, reset status, NR-iteration field, manual-start
validation, multi-warp groups on tiny nets, the batched MARL runner on the device."""
import numpy as np
import pytest
import torch

from mapdn_b200 import cases
from mapdn_b200.network import NetDesc, ProfileDesc

pytestmark = pytest.mark.gpu


def _make(net, prof, args, batch, **kw):
    from mapdn_b200.env import BatchedVoltageControl
    return BatchedVoltageControl(net, prof, args, batch=batch, **kw)


def test_zero_copy_host_path_equals_staged_path_bit_for_bit():
    net, prof = cases.make_case("case33"), cases.make_profiles("case33", n_days=4)
    args = dict(seed=4, voltage_barrier_type="bowl")
    e1, e2, e3 = (_make(net, prof, args, batch=96) for _ in range(3))
    for e in (e1, e2, e3):
        e.reset()
    rng = np.random.default_rng(0)
    n0 = e2.launch_count
    for k in range(4):
        a = rng.uniform(-0.8, 0.8, (96, 6))
        r1, t1, i1, o1 = (x.copy() for x in e1.step_host(a, staged=True))
        r2, t2, i2, o2 = e2.step_host(a)                                  # kernel writes straight into pinned memory
        assert np.array_equal(r1, r2) and np.array_equal(t1, t2) and np.array_equal(i1, i2) and np.array_equal(o1, o2)
        r3, t3, i3, o3 = e3.step_host(a, sync=False)                      # asynchronous: valid after wait()
        e3.wait()
        assert np.array_equal(r1, r3) and np.array_equal(o1, o3)
    assert e2.launch_count - n0 == 4                                      # one launch per step, no copy kernels
    pad = o2.reshape(96, 6, -1)[:, 2, 18:]                                # agent 2's zone has 4 buses: 18 of 50 slots used
    assert pad.shape[-1] == 32 and not pad.any()
    assert e2.host_obs_bytes_per_env == 8 * (50 + 50 + 18 + 14 + 34 + 34)
    # fp32 observations through the zero-copy path
    a = rng.uniform(-0.8, 0.8, (96, 6))
    _, _, _, o32 = e2.step_host(a, obs_dtype=np.float32)
    _, _, _, o64 = e1.step_host(a, staged=True)
    assert o32.dtype == np.float32 and np.array_equal(o32, o64.astype(np.float32))


def test_pinned_entry_point_rejects_pageable_memory():
    import ctypes as C
    from mapdn_b200 import _capi
    net, prof = cases.make_case("case33"), cases.make_profiles("case33", n_days=4)
    env = _make(net, prof, None, batch=4)
    env.reset()
    a = np.zeros((4, 6)); r = np.zeros(4); t = np.zeros(4, np.uint8)
    st = env._L.mapdn_step_host_pinned(env._h, a.ctypes.data_as(C.c_void_p), 1, r.ctypes.data_as(C.c_void_p),
                                       t.ctypes.data_as(C.c_void_p), None, None, 0, 0, 1, None)
    assert st == 1 and b"page-locked" in env._L.mapdn_last_error()


def _unservable_profiles(net, n_days=3):
    T = n_days * 480 + 1
    return ProfileDesc(pv=np.full((T, net.n_sgen), 0.1), load_p=np.full((T, net.n_load), 40.0),
                       load_q=np.full((T, net.n_load), 20.0), steps_per_hour=20, n_days=n_days)


def test_reset_reports_envs_that_never_converge():
    from mapdn_b200._capi import MapdnError
    from mapdn_b200.env import VoltageControl
    from oracle.voltage_control_ref import VoltageControlOracle
    net = cases.case33()
    prof = _unservable_profiles(net)
    env = _make(net, prof, dict(seed=1), batch=5)
    env.reset()
    torch.cuda.synchronize()
    assert env.reset_ok.cpu().tolist() == [0] * 5                        # 16 draws each, none solvable
    assert env.get_field("nr_iters")[:, 0].cpu().tolist() == [10.0] * 5
    with pytest.raises(MapdnError, match="no solvable start"):
        env.reset(check=True, max_retries=1)
    with pytest.raises(RuntimeError):                                     # the oracle gives up at the same point
        VoltageControlOracle(net, prof, env.args, env_id=0).reset()
    with pytest.raises(MapdnError):                                       # the B = 1 shim raises instead of returning junk
        VoltageControl(dict(net=net, profiles=prof, seed=1))
    # a healthy store: every env solved, and a masked reset only reports the selected envs
    good = _make(net, cases.make_profiles("case33", n_days=4), dict(seed=1), batch=5)
    good.reset(check=True)
    assert good.reset_ok.cpu().tolist() == [1] * 5


def test_nr_iteration_field_matches_the_oracle():
    from oracle.voltage_control_ref import VoltageControlOracle
    net, prof = cases.make_case("case141"), cases.make_profiles("case141", n_days=4)
    env = _make(net, prof, dict(seed=3, voltage_barrier_type="l1", action_scale=0.6), batch=9)
    env.reset()
    ors = [VoltageControlOracle(net, prof, env.args, env_id=i) for i in range(9)]
    for o in ors:
        o.reset()
    assert env.get_field("nr_iters")[:, 0].cpu().tolist() == [float(o.g.res.iterations) for o in ors]
    a = np.random.default_rng(1).uniform(-0.6, 0.6, (9, net.n_sgen))
    env.step(torch.tensor(a, device=env.device))
    for o, ai in zip(ors, a):
        o.step(ai)
    assert env.get_field("nr_iters")[:, 0].cpu().tolist() == [float(o.g.res.iterations) for o in ors]


def test_manual_start_outside_the_store_is_clamped_on_the_device_and_rejected_by_the_shim():
    from mapdn_b200.env import VoltageControl
    net, prof = cases.make_case("case33"), cases.make_profiles("case33", n_days=4)
    env = _make(net, prof, dict(seed=0), batch=3)
    start = torch.tensor([[-5, 0, 0], [0, 0, 0], [900, 23, 19]], dtype=torch.int32, device=env.device)
    env.reset(start, add_noise=False)
    torch.cuda.synchronize()
    rows = env.get_field("start_row")[:, 0].cpu().tolist()
    assert rows == [0.0, 0.0, float(prof.n_rows - 1 - env.episode_limit)]
    assert bool(torch.isfinite(env.obs).all())
    shim = VoltageControl(dict(net=net, profiles=prof, seed=0))
    with pytest.raises(ValueError, match="does not fit"):
        shim.manual_reset(900, 23, 19)
    with pytest.raises(ValueError):
        shim.manual_reset(0, 24, 0)
    shim.manual_reset(1, 2, 3)
    assert shim._env.get_field("start_row").item() == 3 + 2 * 20 + 480


@pytest.mark.parametrize("lanes", [64, 128])
def test_multi_warp_groups_on_a_tiny_net(lanes):
    """The group reduction's scratch has its own region (it used to alias n_sgen + 2 n_load doubles, too few on tiny nets)."""
    from oracle.voltage_control_ref import INFO_KEYS, VoltageControlOracle
    net = NetDesc(base_mva=1.0, n_bus=4, slack_bus=0, slack_vm=1.0, br_from=np.array([0, 1, 1]), br_to=np.array([1, 2, 3]),
                  br_r=np.array([0.01, 0.02, 0.015]), br_x=np.array([0.02, 0.03, 0.02]), load_bus=np.array([2, 3]),
                  sgen_bus=np.array([3]), sgen_zone=np.array([1]), bus_zone=np.array([0, 1, 1, 1]), name="tiny4")
    T = 3 * 480 + 1
    rng = np.random.default_rng(0)
    prof = ProfileDesc(pv=rng.uniform(0.1, 0.3, (T, 1)), load_p=rng.uniform(0.1, 0.4, (T, 2)),
                       load_q=rng.uniform(0.0, 0.1, (T, 2)), steps_per_hour=20, n_days=3)
    env = _make(net, prof, dict(seed=2, voltage_barrier_type="bowl"), batch=7, lanes_per_env=lanes)
    env.reset()
    ors = [VoltageControlOracle(net, prof, env.args, env_id=i) for i in range(7)]
    for o in ors:
        o.reset()
    for _ in range(3):
        a = rng.uniform(-0.8, 0.8, (7, 1))
        r, term, info = env.step(torch.tensor(a, device=env.device))
        for i, o in enumerate(ors):
            ro, _, io = o.step(a[i])
            assert abs(ro - r[i].item()) < 1e-9
            assert max(abs(io[k] - info[i, j].item()) for j, k in enumerate(INFO_KEYS)) < 1e-9


class _TinyActorCritic(torch.nn.Module):
    """Stand-in with the reference Model's interface (models/model.py) - the reference tree is not on the GPU box; the
    CPU test tests/test_marl_runner_cpu.py runs the reference's own MADDPG through the same runner."""

    def __init__(self, n, obs_dim, hid, device):
        super().__init__()
        from collections import namedtuple
        self.args = namedtuple("A", "max_steps action_scale action_bias hid_size num_eval_episodes gamma")(8, 0.8, 0.0, hid, 4, 0.99)
        self.n, self.hid = n, hid
        self.fc1 = torch.nn.Linear(obs_dim, hid); self.rnn = torch.nn.GRUCell(hid, hid); self.fc2 = torch.nn.Linear(hid, 1)
        self.q = torch.nn.Sequential(torch.nn.Linear((obs_dim + 1) * n, hid), torch.nn.ReLU(), torch.nn.Linear(hid, 1))
        self.policy_dicts = [self]
        self.to(device)

    def init_hidden(self):
        return self.fc1.weight.new_zeros(1, self.n, self.hid)

    def get_actions(self, state, status, exploration, actions_avail, target=False, last_hid=None):
        B = state.shape[0]
        h = self.rnn(torch.relu(self.fc1(state.reshape(B * self.n, -1))), last_hid.reshape(-1, self.hid))
        mean = self.fc2(h).view(B, self.n, 1)
        act = torch.tanh(mean + (0.3 * torch.randn_like(mean) if exploration else 0.0))
        return act, act, torch.zeros_like(act), mean, h.view(B, self.n, self.hid)

    def value(self, obs, act):
        B = obs.shape[0]
        return self.q(torch.cat([obs.reshape(B, -1), act.reshape(B, -1)], dim=-1)).view(B, 1, 1).expand(-1, self.n, -1)

    def unpack_data(self, batch):
        raise AssertionError("host batches are not used")


def test_marl_runner_on_device():
    import time
    from mapdn_b200.marl_runner import BatchedMarlRunner, DeviceTransitionBuffer, attach
    net, prof = cases.make_case("case33"), cases.make_profiles("case33", n_days=4)
    B = 256
    env = _make(net, prof, dict(voltage_barrier_type="bowl", seed=3, episode_limit=6), batch=B)
    torch.manual_seed(0)
    model = attach(_TinyActorCritic(env.n_agents, env.obs_size, 32, env.device))
    buf = DeviceTransitionBuffer(16, B, env.n_agents, env.obs_size, 1, 32, env.device)
    opt = torch.optim.RMSprop(model.q.parameters(), lr=1e-3)
    losses = []

    def update(runner, stat):            # a DDPG-style critic step on a device batch in the reference's field layout
        if runner.buffer.count < 2:
            return
        st, ac, _, _, nv, rw, ns, dn, ls, *_ = model.unpack_data(runner.buffer.get_batch(2, n_windows=64))
        target = rw[:, :, None] + model.args.gamma * (1 - dn[:, :, None]) * nv
        loss = (model.value(st, ac) - target.detach()).pow(2).mean()
        opt.zero_grad(); loss.backward(); opt.step()
        losses.append(loss.detach())
    runner = BatchedMarlRunner(env, model, buf, update_fn=update)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    stat = runner.train_process({})
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    assert runner.steps == 8 * B and buf.count == 8 and len(losses) == 7
    assert np.isfinite(stat["mean_train_reward"]) and 0.0 <= stat["mean_train_percentage_of_v_out_of_control"] <= 1.0
    b = buf.latest(8)
    st, ac, lp, v, nv, rw, ns, dn, ls, av, lh, h = b.unpacked()
    dn, ls = dn.view(8, B), ls.view(8, B)
    assert bool((dn[4] == 1).all()) and float(dn.sum()) == B            # episode_limit 6: terminated after 5 steps
    assert bool((ls[7] == 1).all()) and float(ls.sum()) == 2 * B        # + the cut at t = max_steps - 1
    assert float(lh.view(8, B, env.n_agents, -1)[5].abs().max()) == 0.0  # hidden state restarts with the new episode
    assert float(ac.abs().max()) <= 1.0 and bool(torch.isfinite(torch.stack(losses)).all())
    ev = runner.evaluation({}, num_eval_episodes=B)
    assert np.isfinite(ev["mean_test_reward"]) and ev["mean_test_destroy"] == 0.0
    print(f"learner-inclusive throughput: {runner.steps / dt / 1e3:.1f} k env-steps/s (B={B}, tiny actor-critic, "
          f"one critic update per lock-step)")


def test_ingest_verify_report(tmp_path):
    """`python -m mapdn_b200.ingest --verify <dir>`: the numbers somebody with the real files compares with pandapower."""
    from mapdn_b200 import ingest
    from oracle.pandapower_nr import PandapowerEquivalent
    net, prof = cases.make_case("case33"), cases.make_profiles("case33", n_days=3)
    ingest.save_scenario_npz(str(tmp_path), net, prof)
    rep = ingest.verify_report(str(tmp_path))
    r = PandapowerEquivalent(net).runpp(prof.load_p[0], prof.load_q[0], prof.pv[0], np.zeros(net.n_sgen))
    assert rep["n_bus"] == 33 and rep["n_load"] == 32 and rep["n_sgen"] == 6 and rep["obs_dim"] == 50
    assert rep["zone_sizes"] == {"main": 6, "zone1": 12, "zone2": 4, "zone3": 3, "zone4": 8}
    assert rep["converged"] and rep["newton_iterations"] == r.iterations
    assert abs(rep["v_min_pu"] - r.vm_pu.min()) < 1e-9 and rep["v_min_bus"] == int(r.vm_pu.argmin())
    assert abs(rep["total_line_loss_mw"] - r.pl_mw.sum()) < 1e-9 and abs(rep["ext_grid_p_mw"] - r.p_ext_mw) < 1e-9
    assert rep["kcl_residual_pu"] < 1e-8 and rep["solver"].startswith("radial")


def test_out_of_service_zero_impedance_branch_contributes_nothing():
    """makeYbus multiplies by the status: a disconnected branch with r = x = 0 must not poison the diagonal with NaN."""
    net = NetDesc(base_mva=1.0, n_bus=3, slack_bus=0, slack_vm=1.0, br_from=np.array([0, 1, 0]), br_to=np.array([1, 2, 2]),
                  br_r=np.array([0.01, 0.02, 0.0]), br_x=np.array([0.02, 0.03, 0.0]),
                  br_status=np.array([1, 1, 0], np.uint8), load_bus=np.array([1, 2]), sgen_bus=np.array([2]),
                  sgen_zone=np.array([1]), bus_zone=np.array([0, 1, 1]), name="open3")
    from oracle.pandapower_nr import PandapowerEquivalent
    env = _make(net, None, None, batch=1)
    Y = env.ybus_dense()
    assert np.isfinite(Y.real).all() and np.isfinite(Y.imag).all()
    pl, ql = np.array([[0.3, 0.2]]), np.array([[0.1, 0.05]])
    out = env.solve(pl, ql, np.array([[0.1]]), np.array([[0.02]]))
    torch.cuda.synchronize()
    with np.errstate(all="ignore"):
        r = PandapowerEquivalent(net).runpp(pl[0], ql[0], np.array([0.1]), np.array([0.02]))
    assert int(out["converged"][0]) == 1 and np.abs(out["vm"][0].cpu().numpy() - r.vm_pu).max() < 1e-10


def test_node_relabelling_changes_addresses_not_results(monkeypatch):
    """The bank-conflict-aware node ids (mapdn_create, section 2b) only move records inside shared memory: rewards, info,
    observations and voltages are bit-identical with and without them."""
    net, prof = cases.make_case("case33"), cases.make_profiles("case33", n_days=4)
    args = dict(seed=6, voltage_barrier_type="bowl")
    monkeypatch.setenv("MAPDN_NO_RELABEL", "1")
    e0 = _make(net, prof, args, batch=40)
    monkeypatch.delenv("MAPDN_NO_RELABEL")
    e1 = _make(net, prof, args, batch=40)
    o0, s0 = e0.reset(); o1, s1 = e1.reset()
    assert torch.equal(o0, o1) and torch.equal(s0, s1)
    rng = np.random.default_rng(0)
    for _ in range(3):
        a = torch.tensor(rng.uniform(-0.8, 0.8, (40, 6)), device=e0.device)
        r0, t0, i0 = e0.step(a); r1, t1, i1 = e1.step(a)
        assert torch.equal(r0, r1) and torch.equal(t0, t1) and torch.equal(i0, i1) and torch.equal(e0.obs, e1.obs)
        assert torch.equal(e0.get_field("vm"), e1.get_field("vm")) and torch.equal(e0.get_state(), e1.get_state())


def test_droop_on_a_meshed_net_uses_the_dense_solver():
    """MODE_DROOP is instantiated for the dense-LU fallback too: same loop, same results as the host-driven loop."""
    from mapdn_b200.baselines import droop_control, droop_control_host_loop
    z = np.array([0.02 + 0.04j, 0.01 + 0.03j, 0.0125 + 0.025j, 0.015 + 0.03j])
    net = NetDesc(base_mva=100.0, n_bus=4, slack_bus=0, slack_vm=1.03, br_from=np.array([0, 0, 1, 2]), br_to=np.array([1, 2, 2, 3]),
                  br_r=z.real, br_x=z.imag, load_bus=np.array([1, 2, 3]), sgen_bus=np.array([2, 3]), sgen_zone=np.array([1, 1]),
                  bus_zone=np.array([0, 1, 1, 1]), name="mesh4")
    env = _make(net, None, None, batch=1)
    assert env.dims["n_levels"] == 1                                   # meshed: no elimination forest
    rng = np.random.default_rng(2)
    B = 9
    pl = rng.uniform(80, 260, (B, 3)); ql = pl * rng.uniform(0.2, 0.5, (B, 3))
    pv = rng.uniform(0, 60, (B, 2)); s_rated = np.array([80.0, 80.0])
    out = droop_control(env, pl, ql, pv, s_rated)
    ref = droop_control_host_loop(env, pl, ql, pv, s_rated)
    torch.cuda.synchronize()
    assert torch.equal(out["iterations"], ref["iterations"]) and int(out["iterations"].max()) > 2
    assert float((out["vm"] - ref["vm"]).abs().max()) < 1e-10 and float((out["q"] - ref["q"]).abs().max()) < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("scenario,batch", [("case33", 37), ("case141", 9), ("case322", 5)])
def test_compact_host_rows_equal_the_padded_observations(scenario, batch):
    """mapdn_step_host_compact: agent a's block of a compact row == its reference row without the zero padding
    (voltage_control_env.py:254-274); the padded layout is rebuilt exactly by expand_obs; fp32 rows are the rounded
    fp64 ones."""
    from mapdn_b200 import cases
    from mapdn_b200.env import BatchedVoltageControl
    net, prof = cases.make_case(scenario), cases.make_profiles(scenario, n_days=3)
    mk = lambda: BatchedVoltageControl(net, prof, dict(seed=11), batch=batch)
    a, b, c = mk(), mk(), mk()
    for e in (a, b, c):
        e.reset()
    slices = a.obs_slices
    assert len(slices) == a.n_agents and a.obs_row_len % 4 == 0
    assert sum(n for _, n in slices) <= a.obs_row_len < sum(n for _, n in slices) + 4
    assert a.host_obs_bytes_per_env == 8 * sum(n for _, n in slices)
    rng = np.random.default_rng(5)
    for t in range(3):
        act = rng.uniform(-0.8, 0.8, (batch, a.n_agents))
        r0, d0, i0, o0 = [x.copy() for x in a.step_host(act, add_noise=True)]
        r1, d1, i1, o1 = [x.copy() for x in b.step_host(act, add_noise=True, layout="compact", staged=False if t & 1 else None)]
        r2, d2, i2, o2 = [x.copy() for x in c.step_host(act, add_noise=True, layout="compact", obs_dtype=np.float32,
                                                        staged=None if t & 1 else False)]
        assert o1.shape == (batch, a.obs_row_len) and o2.dtype == np.float32
        np.testing.assert_array_equal(r0, r1); np.testing.assert_array_equal(d0, d1); np.testing.assert_array_equal(i0, i1)
        np.testing.assert_array_equal(r0, r2)
        np.testing.assert_array_equal(a.expand_obs(o1), o0)                       # bit-exact, padding included
        np.testing.assert_array_equal(a.expand_obs(o2), o0.astype(np.float32))
        for ag, (off, n) in enumerate(slices):
            assert not np.any(o0[:, ag, n:])                                      # what is left out is zero padding
        np.testing.assert_array_equal(o1[:, sum(n for _, n in slices):], 0.0)     # row tail
    # non-blocking form
    act = rng.uniform(-0.8, 0.8, (batch, a.n_agents))
    ref = [x.copy() for x in a.step_host(act)]
    out = b.step_host(act, layout="compact", sync=False)
    b.wait()
    np.testing.assert_array_equal(a.expand_obs(out[3]), ref[3])
    for e in (a, b, c):
        e.close()


@pytest.mark.gpu
def test_compact_host_path_rejects_pageable_memory():
    import ctypes as C
    from mapdn_b200 import cases, _capi
    from mapdn_b200.env import BatchedVoltageControl
    net, prof = cases.make_case("case33"), cases.make_profiles("case33", n_days=3)
    env = BatchedVoltageControl(net, prof, dict(seed=1), batch=4)
    env.reset()
    hb = env._host_buffers()
    pageable = np.zeros((4, env.obs_row_len))
    st = env._L.mapdn_step_host_compact(env._h, hb["actions"].data_ptr(), 1, hb["reward"].data_ptr(),
                                        hb["terminated"].data_ptr(), hb["info"].data_ptr(),
                                        pageable.ctypes.data, 0, 0, 1, None)
    assert st != 0 and b"obs_host" in env._L.mapdn_last_error()
    env.close()


@pytest.mark.gpu
def test_compact_rows_longer_than_the_padded_ones():
    """One agent, 14 entries: no padding at all, and the compact row (rounded up to 16 entries) is LONGER than the padded
    one - the device staging buffer has to be sized for it."""
    net = NetDesc(base_mva=1.0, n_bus=4, slack_bus=0, slack_vm=1.0, br_from=np.array([0, 1, 1]), br_to=np.array([1, 2, 3]),
                  br_r=np.array([0.01, 0.02, 0.015]), br_x=np.array([0.02, 0.03, 0.02]), load_bus=np.array([2, 3]),
                  sgen_bus=np.array([3]), sgen_zone=np.array([1]), bus_zone=np.array([0, 1, 1, 1]), name="tiny4")
    T = 3 * 480 + 1
    rng = np.random.default_rng(0)
    prof = ProfileDesc(pv=rng.uniform(0.1, 0.3, (T, 1)), load_p=rng.uniform(0.1, 0.4, (T, 2)),
                       load_q=rng.uniform(0.0, 0.1, (T, 2)), steps_per_hour=20, n_days=3)
    a, b = _make(net, prof, dict(seed=2), batch=7), _make(net, prof, dict(seed=2), batch=7)
    a.reset(); b.reset()
    assert a.obs_size == 14 and b.obs_slices == [(0, 14)] and b.obs_row_len == 16
    for t in range(4):
        act = rng.uniform(-0.8, 0.8, (7, 1))
        ref = [x.copy() for x in a.step_host(act)]
        out = [x.copy() for x in b.step_host(act, layout="compact", staged=False if t & 1 else None,
                                             obs_dtype=np.float32 if t & 2 else np.float64)]
        np.testing.assert_array_equal(out[0], ref[0])
        np.testing.assert_array_equal(out[3][:, :14], ref[3][:, 0, :].astype(out[3].dtype))
        np.testing.assert_array_equal(out[3][:, 14:], 0.0)
    a.close(); b.close()
