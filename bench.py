#!/usr/bin/env python
"""Benchmark of the MAPDN var_voltage_control env step (BASELINE.json metric: env-steps/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): case33 (33-bus feeder), 4096 env instances per GPU, Bowl
voltage barrier, noise on, synthetic load/PV profiles and random actions (the reference's data
files are not in the repo). One "step" = one fused launch advancing every env by one transition:
action clip -> fp64 Newton-Raphson power flow -> reward/info -> next profile row + noise -> obs.

Prints ONE JSON line (rank 0). `value` = env-steps/s with actions already resident in HBM, timed
with CUDA events around each step (L2 flushed between steps, outside the event pairs), max over
ranks. `e2e` = the same through the host-buffer API (H2D of actions, D2H of reward / done / info /
obs inside the timed region). `--impl reference` times the CPU restatement of the reference path
(oracle/: pandapower-2.7.0-equivalent NR + env logic, NumPy/SciPy; pandapower itself cannot be
installed in this image) on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCENARIO_DEFAULT = "case33"
BATCH_DEFAULT = {"case33": 4096, "case141": 2048, "case322": 1024}
METRIC = "env-steps/sec (batched power-flow solves)"


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (port of the reference path) on the host cores
# ------------------------------------------------------------------------------------------------
_W = {}


def _cpu_init(scenario, barrier, seed):
    from mapdn_b200 import cases
    from oracle.voltage_control_ref import VoltageControlOracle
    import multiprocessing as mp
    wid = mp.current_process()._identity[0] if mp.current_process()._identity else 0
    net, prof = cases.make_case(scenario), cases.make_profiles(scenario)
    env = VoltageControlOracle(net, prof, dict(voltage_barrier_type=barrier, seed=seed,
                                               action_scale=cases.SCENARIOS[scenario]["action_scale"]), env_id=wid)
    env.reset()
    _W["env"], _W["rng"] = env, np.random.default_rng(wid)


def _cpu_work(n_steps):
    env, rng = _W["env"], _W["rng"]
    t0 = time.perf_counter()
    for _ in range(n_steps):
        a = rng.uniform(env.low, env.high, env.n_agents)
        _, term, _ = env.step(a)            # reference hot path: step() ...
        env.get_obs()                       # ... followed by get_obs() (models/model.py:216-219)
        if term:
            env.reset()
    return time.perf_counter() - t0


class CpuArm:
    def __init__(self, scenario, barrier, seed=0, cores=None):
        import multiprocessing as mp
        self.cores = cores or os.cpu_count() or 1
        os.environ.setdefault("OMP_NUM_THREADS", "1")
        self.pool = mp.get_context("spawn").Pool(self.cores, initializer=_cpu_init, initargs=(scenario, barrier, seed))
        self.pool.map(_cpu_work, [1] * self.cores)      # spin-up

    def run(self, env_steps):
        per = max(1, env_steps // self.cores)
        t0 = time.perf_counter()
        self.pool.map(_cpu_work, [per] * self.cores)
        dt = time.perf_counter() - t0
        return per * self.cores, dt

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference(args):
    from mapdn_b200 import cases
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sc = args.scenario
    barrier = cases.SCENARIOS[sc]["barrier"]
    arm = CpuArm(sc, barrier)
    # env-steps per "step" of this arm: a bounded sample of the 4096-env batch, sized from the rate seen in the first
    # warm-up pass so that warm-up + K timed steps take about two minutes whatever K is (at least one env-step per core)
    sample = args.cpu_sample or 2 * arm.cores
    n0, dt0 = arm.run(sample)
    if not args.cpu_sample:
        budget_steps = (n0 / dt0) * 100.0 / max(1, args.steps + args.warmup)
        sample = int(min(8 * arm.cores, max(arm.cores, budget_steps // arm.cores * arm.cores)))
    for _ in range(max(0, args.warmup - 1)):
        arm.run(sample)
    done, t = 0, 0.0
    for _ in range(args.steps):
        n, dt = arm.run(sample)
        done += n
        t += dt
    arm.close()
    val = done / t
    net = cases.make_case(sc)
    line = dict(impl="reference", metric=METRIC, value=val, unit="env-steps/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=t / args.steps * 1e3, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=f"{sc} x {BATCH_DEFAULT[sc]} envs, {barrier} barrier (sampled: "
                                     f"{sample} env-steps per step)", n_bus=net.n_bus, n_agents=net.n_sgen),
                cpu_baseline=dict(value=val, unit="env-steps/s", cores=arm.cores, kind="port",
                                  sample=f"{done} env-steps of {sc} ({sample} per step x {args.steps} steps), "
                                         "oracle/ NumPy+SciPy restatement of pandapower 2.7.0 runpp + env logic "
                                         "(pandapower not installable in this image), one process per core"),
                e2e=dict(value=val, unit="env-steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons DURING the timed region: NVML polled every ~2 ms from a thread (the timed region of
    the default run is only tens of ms long); `nvidia-smi -lms 50` is the fallback when NVML is not usable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASON_BITS = (("sw_power_cap", 0x4), ("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20),
                   ("hw_thermal_slowdown", 0x40))            # nvmlClocksEventReason* bit masks

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml, self.samples = index, [], None, None, []
        self._stop = threading.Event()

    def _nvml_open(self):
        import pynvml
        import torch
        pynvml.nvmlInit()
        h = None
        try:
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)); int(reasons(h))      # probe once
        return pynvml, h, reasons, mx

    def _poll(self):
        pynvml, h, reasons, _ = self.nvml
        while not self._stop.is_set():
            try:
                self.samples.append((time.perf_counter(), float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                                     int(reasons(h))))
            except Exception:
                break
            time.sleep(0.002)

    def start(self):
        try:
            self.nvml = self._nvml_open()
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True,
                                         stderr=subprocess.DEVNULL)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.perf_counter(), ln.strip()))

    def stop(self, t0, t1):
        if self.nvml is not None:
            self._stop.set()
            self.th.join(timeout=1.0)
            rows = [r for r in self.samples if t0 <= r[0] <= t1] or self.samples[-3:]
            if rows:
                bits = 0
                for r in rows:
                    bits |= r[2]
                return dict(sm_mhz=float(np.median([r[1] for r in rows])), sm_max_mhz=self.nvml[3],
                            reasons=sorted(nm for nm, b in self.REASON_BITS if bits & b), samples=len(rows), source="nvml")
            return None
        if self.proc is None:
            return None
        time.sleep(0.12)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.06] or [r for _, r in self.rows[-3:]]
        sm, mx, reasons = [], [], set()
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except Exception:
                continue
            for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return None
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm),
                    source="nvidia-smi")


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from mapdn_b200 import cases
    from mapdn_b200.distributed import ShardedVoltageControl

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    sc = args.scenario
    B = args.batch or BATCH_DEFAULT[sc]
    barrier = cases.SCENARIOS[sc]["barrier"]
    net, prof = cases.make_case(sc), cases.make_profiles(sc)
    env_args = dict(voltage_barrier_type=barrier, action_scale=cases.SCENARIOS[sc]["action_scale"], seed=0)
    env = ShardedVoltageControl(net, prof, env_args, global_batch=B * world, device=local, lanes_per_env=args.lanes)
    assert env.count == B
    K, W = args.steps, args.warmup
    n_sets = 8
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    lo, hi = env.action_space.low, env.action_space.high
    acts = lo + (hi - lo) * torch.rand(n_sets, B, env.n_agents, dtype=torch.float64, device=dev, generator=g)
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    ep_len = env.episode_limit - 1
    state = dict(t=0)

    def one_step(i):
        if state["t"] == ep_len:            # every env terminated at episode_limit: start a new episode
            env.reset()
            state["t"] = 0
        env.step(acts[i % n_sets])
        state["t"] += 1

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    env.reset()
    for i in range(W):
        one_step(i)
        if flush is not None:
            flush.zero_()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.15)
    sync_all()
    launches0 = env.launch_count
    w0 = time.perf_counter()
    for i in range(K):
        ev0[i].record()
        one_step(W + i)
        ev1[i].record()
        if flush is not None:
            flush.zero_()
    returns = env.gather_episode_returns()           # the path's only collective
    sync_all()
    w1 = time.perf_counter()
    launches = env.launch_count - launches0
    clk = clocks.stop(w0, w1) if rank == 0 else None
    dev_ms = sum(a.elapsed_time(b) for a, b in zip(ev0, ev1))
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t.item())
    value = B * world * K / (dev_ms_max * 1e-3)
    assert returns.shape[0] == B * world and bool(torch.isfinite(returns).all())

    # ---- end to end through the host-buffer API (numpy in / numpy out) ----
    rng = np.random.default_rng(rank)
    host_acts = [rng.uniform(lo, hi, (B, env.n_agents)) for _ in range(4)]
    Ke = max(3, min(K, args.e2e_steps))
    env.reset(); state["t"] = 0
    for i in range(3):
        env.step_host(host_acts[i % 4])
    sync_all()
    e0 = time.perf_counter()
    for i in range(Ke):
        if state["t"] == ep_len:
            env.reset(); state["t"] = 0
        r_h, t_h, i_h, o_h = env.step_host(host_acts[i % 4])
        state["t"] += 1
    sync_all()
    e_dt = time.perf_counter() - e0
    te = torch.tensor([e_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = B * world * Ke / float(te.item())
    h2d = B * env.n_agents * 8
    d2h = B * (8 + 1 + 11 * 8 + env.n_agents * env.obs_size * 8)
    # variant: observations delivered in fp32 (what the reference's learners consume after prep_obs)
    env.reset(); state["t"] = 0
    for i in range(3):
        env.step_host(host_acts[i % 4], obs_dtype=np.float32)
    sync_all()
    e0 = time.perf_counter()
    for i in range(Ke):
        if state["t"] == ep_len:
            env.reset(); state["t"] = 0
        env.step_host(host_acts[i % 4], obs_dtype=np.float32)
        state["t"] += 1
    sync_all()
    te32 = torch.tensor([time.perf_counter() - e0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te32, op=dist.ReduceOp.MAX)
    e2e32_val = B * world * Ke / float(te32.item())
    d2h32 = B * (8 + 1 + 11 * 8 + env.n_agents * env.obs_size * 4)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant (only) kernel of the step ----
    peaks, peak_src = None, "fallback"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, peak_src = float(peaks["hbm_gbs"]), "measured"
    except Exception:
        peak = 6650.0
    alg_bytes = env.dims["algorithmic_bytes_per_env_step"] * B           # per launch
    ms_kernel = dev_ms_max / K
    achieved = alg_bytes / (ms_kernel * 1e-3) / 1e9
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj.get(f"{sc}_B{B}")
    except Exception:
        pass
    roofline = dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic,
                    peak_source=f"{peak_src} (MEASURED_PEAKS.json hbm_gbs)" if peak_src == "measured" else "fallback 6.65 TB/s",
                    kernel=f"env_kernel<{env.dims['lanes_per_env']},STEP>", algorithmic_bytes_per_launch=alg_bytes,
                    note="fused kernel is fp64-issue/latency bound, not HBM bound (DESIGN.md §roofline)")

    # ---- CPU baseline: the oracle port on the host cores, bounded sample ----
    cpu = None
    if world == 1 and not args.no_cpu:
        arm = CpuArm(sc, barrier)
        n, dt = arm.run(args.cpu_sample or 200 * arm.cores)
        arm.close()
        cpu = dict(value=n / dt, unit="env-steps/s", cores=arm.cores, kind="port",
                   sample=f"{n} env-steps of {sc} ({barrier} barrier, noise on) in {dt:.1f} s: oracle/ NumPy+SciPy "
                          "restatement of pandapower 2.7.0 runpp + reference env logic, one process per core")

    line = dict(metric=METRIC, value=value, unit="env-steps/s", n_gpus=world, steps=K, warmup=W,
                ms_per_step=ms_kernel, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                data="synthetic",
                config=dict(workload=f"{sc} x {B} envs per GPU, {barrier} barrier, noise on, fused step "
                                     "(runpp-equivalent NR + reward + next row + obs)",
                            n_bus=net.n_bus, n_agents=net.n_sgen, obs_dim=env.obs_size, global_batch=B * world,
                            lanes_per_env=env.dims["lanes_per_env"], parallelism=f"envs sharded over {world} GPU(s)",
                            l2="flushed between steps (256 MiB memset, outside the event pairs)" if flush is not None
                            else "not flushed", timing="sum of per-step CUDA-event pairs, max over ranks"),
                clocks=clk, gpu_launches=int(launches),
                e2e=dict(value=e2e_val, unit="env-steps/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, steps=Ke,
                         obs_dtype="f64"),
                e2e_obs_f32=dict(value=e2e32_val, unit="env-steps/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h32,
                                 steps=Ke, note="same call with observations delivered in fp32 (opt-in API)"),
                roofline=roofline, cpu_baseline=cpu, wall_ms_per_step=(w1 - w0) / K * 1e3)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scenario", default=SCENARIO_DEFAULT, choices=list(BATCH_DEFAULT))
    ap.add_argument("--batch", type=int, default=0, help="envs per GPU (default: BASELINE.json config)")
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=50)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
