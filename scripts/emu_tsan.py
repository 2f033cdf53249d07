"""Race check of the kernel source on the CPU: the emulated library (tests/emu) built with -fsanitize=thread. Every CUDA
thread is an OS thread and every barrier (__syncwarp, __syncthreads, bar.sync / arrive / red) a mutex-protected counting
barrier, so ThreadSanitizer's happens-before analysis covers shared AND global memory (compute-sanitizer's racecheck only
sees shared memory) and is stricter than the hardware (no implicit warp-synchronous ordering).
    python scripts/emu_tsan.py            # builds /tmp/mapdn_emu_tsan/libemu_tsan.so, runs reset / step / solve / droop
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = "/tmp/mapdn_emu_tsan"
LIB = os.path.join(OUT, "libemu_tsan.so")
CHILD = r'''
import sys, numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests/emu")
import emu_env
emu_env.LIB = %(lib)r; emu_env.build = lambda force=False: emu_env.LIB
from emu_env import EmuEnv
from mapdn_b200 import cases
name, lanes, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
net, prof = cases.make_case(name), cases.make_profiles(name, n_days=3)
scale = cases.SCENARIOS[name]["action_scale"]
env = EmuEnv(net, prof, dict(seed=3, voltage_barrier_type="bowl", action_scale=scale), batch=B, lanes_per_env=lanes)
env.reset()
rng = np.random.default_rng(0)
for t in range(2):
    env.step(rng.uniform(-scale, scale, (B, net.n_sgen)))
env.step_host(rng.uniform(-scale, scale, (B, net.n_sgen)), path="compact")
inp = cases.synthetic_inputs(name, 3, seed=1)
env.solve(inp["p_load"], inp["q_load"], inp["p_pv"], np.zeros_like(inp["p_pv"]))
env.droop(inp["p_load"], inp["q_load"], inp["p_pv"], inp["s_max"], max_ite=6)
print("ran", name, lanes, B)
'''


def main():
    os.makedirs(OUT, exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-pthread", "-fsanitize=thread", "-DMAPDN_HOST_EMU",
                    "-DMAPDN_EMU_DEFINE_GLOBALS", "-I", os.path.join(ROOT, "tests", "emu"), "-x", "c++",
                    os.path.join(ROOT, "mapdn_b200", "csrc", "mapdn_b200.cu"), "-o", LIB], check=True, stderr=subprocess.DEVNULL)
    tsan = subprocess.run(["g++", "-print-file-name=libtsan.so"], capture_output=True, text=True).stdout.strip()
    child = os.path.join(OUT, "child.py")
    open(child, "w").write(CHILD % dict(root=ROOT, lib=LIB))
    env = dict(os.environ, LD_PRELOAD=os.path.realpath(tsan), TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=0")
    total = 0
    for cfg in (("case33", "8", "9"), ("case33", "4", "5"), ("case33", "64", "2"), ("case141", "32", "3"), ("case322", "64", "2")):
        res = subprocess.run([sys.executable, child, *cfg], env=env, capture_output=True, text=True)
        n = res.stderr.count("WARNING: ThreadSanitizer")
        total += n
        ok = "ran" in res.stdout
        print(f"{cfg[0]} lanes/env={cfg[1]} B={cfg[2]} (reset, 2 steps, compact host step, solve, droop): "
              f"{'completed' if ok else 'FAILED'}, ThreadSanitizer reports: {n}")
        if n:
            sys.stderr.write(res.stderr[:6000])
    return 1 if total else 0


if __name__ == "__main__":
    sys.exit(main())
