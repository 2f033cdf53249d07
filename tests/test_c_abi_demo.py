"""The C-ABI from plain C: examples/c_abi_demo.c compiles as strict C99 against include/mapdn_b200.h, links against
the in-tree library, and (on a GPU) reproduces what the Python binding computes for the same feeder."""
import math
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT

SRC = os.path.join(ROOT, "examples", "c_abi_demo.c")
OUT_DIR = os.path.join(ROOT, "examples", "_build")
EXE = os.path.join(OUT_DIR, "c_abi_demo")


def build_demo(built_lib):
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    os.makedirs(OUT_DIR, exist_ok=True)
    lib_dir = os.path.dirname(built_lib)
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-O2", "-I", os.path.join(ROOT, "include"),
           SRC, "-o", EXE, "-L", lib_dir, "-lmapdn_b200", "-Wl,-rpath," + lib_dir, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return EXE


def test_demo_compiles_as_c99_and_fails_loudly_without_a_gpu(built_lib):
    exe = build_demo(built_lib)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 2 and "mapdn_create failed" in r.stderr      # no CPU fallback behind the C-ABI


def _demo_inputs():
    """The feeder and profile store of examples/c_abi_demo.c, expression for expression (libm sin/cos)."""
    from mapdn_b200.network import NetDesc, ProfileDesc
    net = NetDesc(base_mva=10.0, n_bus=6, slack_bus=0, slack_vm=1.0, br_from=np.array([0, 1, 2, 2, 4]),
                  br_to=np.array([1, 2, 3, 4, 5]), br_r=np.array([0.01 * (k + 1) for k in range(5)]),
                  br_x=np.array([0.02 * (k + 1) for k in range(5)]), load_bus=np.array([1, 2, 3, 4, 5]),
                  sgen_bus=np.array([3, 5]), sgen_zone=np.array([1, 2]), bus_zone=np.array([0, 0, 1, 1, 2, 2]),
                  vm_init=1.0)
    T, day = 5 * 480, 480
    pv, lp, lq = np.zeros((T, 2)), np.zeros((T, 5)), np.zeros((T, 5))
    for t in range(T):
        sun = math.sin(math.pi * (t % day) / day)
        for g in range(2):
            pv[t, g] = 0.3 * (sun if sun > 0.0 else 0.0) * (1.0 + 0.1 * g)
        for l in range(5):
            lp[t, l] = 0.2 + 0.05 * l + 0.05 * math.cos(2.0 * math.pi * t / day)
            lq[t, l] = 0.3 * lp[t, l]

    class DemoProfiles(ProfileDesc):          # the demo passes round constants instead of the data statistics
        pv_std = np.full(2, 0.001)
        load_p_std = np.full(5, 0.002)
        load_q_std = np.full(5, 0.0005)
        s_max = np.array([1.2 * 0.3 * (1.0 + 0.1 * g) for g in range(2)])

    return net, DemoProfiles(pv=pv, load_p=lp, load_q=lq, steps_per_hour=20, n_days=4)


@pytest.mark.gpu
def test_demo_matches_the_python_binding(built_lib):
    exe = build_demo(built_lib)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[0].startswith("dims n_agents 2 obs_dim ")
    got = [ln.split() for ln in lines if ln.startswith("step ")]
    assert len(got) == 10

    from mapdn_b200.env import BatchedVoltageControl
    net, prof = _demo_inputs()
    env = BatchedVoltageControl(net, prof, dict(voltage_barrier_type="bowl", seed=2024, action_scale=0.8), batch=8)
    assert int(lines[0].split()[4]) == env.obs_size
    env.reset()
    for k in range(10):
        a = np.array([[-0.8 + 1.6 * ((e * 7 + g * 3 + k) % 11) / 10.0 for g in range(2)] for e in range(8)])
        rew, term, _ = env.step(torch.tensor(a, device=env.device))
        torch.cuda.synchronize()
        rs, os_ = float(got[k][2]), float(got[k][3])
        assert abs(rs - float(rew.cpu().numpy().sum())) < 1e-12 * max(1.0, abs(rs))
        assert abs(os_ - float(env.obs.cpu().numpy().sum())) < 1e-12 * max(1.0, abs(os_))
        assert int(got[k][4]) == int(term.sum().item())
    assert lines[-1] == f"launches {env.launch_count}"   # Ybus assembly at create + one reset + ten fused steps
