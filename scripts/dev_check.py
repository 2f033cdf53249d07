"""Developer smoke check on a GPU box: CUDA path vs the oracle on a few envs (not a test)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from mapdn_b200 import cases
from mapdn_b200.env import BatchedVoltageControl
from oracle.pandapower_nr import PandapowerEquivalent, make_ybus
from oracle.voltage_control_ref import VoltageControlOracle, INFO_KEYS

for name in ["case33", "case141", "case322"]:
    net = cases.make_case(name); prof = cases.make_profiles(name, n_days=4)
    B = 37
    inp = cases.synthetic_inputs(name, B, seed=1)
    for G in (0, 4, 8, 16, 32):
        try:
            env = BatchedVoltageControl(net, prof, dict(voltage_barrier_type=cases.SCENARIOS[name]["barrier"],
                                        action_scale=cases.SCENARIOS[name]["action_scale"]), batch=B, lanes_per_env=G)
        except Exception as ex:
            print(name, "G", G, "create failed:", ex); continue
        if G == 0:
            Y = env.ybus_dense(); Yo = make_ybus(net)[0].toarray()
            print(name, "dims", env.dims, "ybus err", np.abs(Y - Yo).max())
        q = inp["action"] * np.sqrt(inp["s_max"] ** 2 - inp["p_pv"] ** 2)
        out = env.solve(inp["p_load"], inp["q_load"], inp["p_pv"], q)
        torch.cuda.synchronize()
        pf = PandapowerEquivalent(net)
        errs = dict(vm=0, va=0, p=0, q=0, pl=0); itbad = 0
        for e in range(B):
            r = pf.runpp(inp["p_load"][e], inp["q_load"][e], inp["p_pv"][e], q[e])
            errs["vm"] = max(errs["vm"], np.abs(out["vm"][e].cpu().numpy() - r.vm_pu).max())
            errs["va"] = max(errs["va"], np.abs(out["va_deg"][e].cpu().numpy() - r.va_degree).max())
            errs["p"] = max(errs["p"], np.abs(out["p_bus"][e].cpu().numpy() - r.p_mw).max())
            errs["q"] = max(errs["q"], np.abs(out["q_bus"][e].cpu().numpy() - r.q_mvar).max())
            errs["pl"] = max(errs["pl"], np.abs(out["pl"][e].cpu().numpy() - r.pl_mw).max())
            itbad += int(out["iterations"][e].item() != r.iterations) + int(out["converged"][e].item() != r.converged)
        print(name, "G", G, "solve errs", {k: float(f"{v:.2e}") for k, v in errs.items()}, "iter mismatches", itbad)
        # env trajectory parity, 2 envs, with noise
        if G in (0, 32):
            obs, state = env.reset()
            torch.cuda.synchronize()
            ors = [VoltageControlOracle(net, prof, env.args, env_id=i) for i in (0, B - 1)]
            mx = 0
            for o, i in zip(ors, (0, B - 1)):
                oo, os_ = o.reset()
                mx = max(mx, np.abs(np.array(oo) - obs[i].cpu().numpy()).max(), np.abs(os_ - state[i].cpu().numpy()).max())
            print(name, "G", G, "reset obs/state err", mx)
            rng = np.random.default_rng(5)
            mx = dict(rew=0, info=0, obs=0, state=0)
            for t in range(6):
                a = rng.uniform(env.action_space.low, env.action_space.high, (B, env.n_agents))
                r, term, info = env.step(torch.tensor(a, device=env.device))
                st = env.get_state()
                torch.cuda.synchronize()
                for o, i in zip(ors, (0, B - 1)):
                    ro, to, io = o.step(a[i])
                    mx["rew"] = max(mx["rew"], abs(ro - r[i].item()))
                    mx["info"] = max(mx["info"], max(abs(io[k] - info[i, j].item()) for j, k in enumerate(INFO_KEYS)))
                    mx["obs"] = max(mx["obs"], np.abs(np.array(o.get_obs()) - env.obs[i].cpu().numpy()).max())
                    mx["state"] = max(mx["state"], np.abs(o.get_state() - st[i].cpu().numpy()).max())
                    assert to == bool(term[i].item())
            print(name, "G", G, "step errs", mx)
        env.close()

# quick timing
name = "case33"; net = cases.make_case(name); prof = cases.make_profiles(name)
for G in (4, 8, 16, 32):
    B = 4096
    env = BatchedVoltageControl(net, prof, dict(voltage_barrier_type="bowl"), batch=B, lanes_per_env=G)
    env.reset()
    a = torch.zeros(B, env.n_agents, dtype=torch.float64, device=env.device).uniform_(-0.8, 0.8)
    for _ in range(5): env.step(a)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(50): env.step(a)
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 50
    print(f"case33 B={B} G={G}: {ms*1e3:.1f} us/step, {B/ms*1e3/1e6:.1f} M env-steps/s", env.dims["smem_bytes"], env.dims["envs_per_block"])
    env.close()
