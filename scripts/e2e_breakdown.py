"""Where the end-to-end step time goes (case33 x 4096, zero-copy host path): device time of the launch that writes to
pinned host memory (CUDA events) vs wall time of the whole step_host call."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from mapdn_b200 import cases
from mapdn_b200.env import BatchedVoltageControl

B = 4096
net, prof = cases.make_case("case33"), cases.make_profiles("case33")
env = BatchedVoltageControl(net, prof, dict(voltage_barrier_type="bowl", seed=0), batch=B)
env.reset()
rng = np.random.default_rng(0)
acts = [rng.uniform(-0.8, 0.8, (B, 6)) for _ in range(4)]
for f32, layout in ((False, "padded"), (True, "padded"), (False, "compact"), (True, "compact")):
    dt = np.float32 if f32 else np.float64
    for i in range(5):
        env.step_host(acts[i % 4], obs_dtype=dt, layout=layout)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(50)]
    t_wall, t_copy = [], []
    for i in range(50):
        t0 = time.perf_counter()
        hb = env._host_buffers(); hb["actions"].numpy()[...] = acts[i % 4]
        t1 = time.perf_counter()
        ev[i][0].record()
        env.step_host(acts[i % 4], obs_dtype=dt, sync=False, layout=layout)
        ev[i][1].record()
        env.wait()
        t_wall.append(time.perf_counter() - t0); t_copy.append(t1 - t0)
    torch.cuda.synchronize()
    dev = np.median([a.elapsed_time(b) for a, b in ev]) * 1e3
    print(f"obs {'fp32' if f32 else 'fp64'} {layout}: device time of the zero-copy launch {dev:.1f} us; wall per step {np.median(t_wall) * 1e6:.1f} us "
          f"(of which copying the actions into the pinned buffer {np.median(t_copy) * 1e6:.1f} us)")
# device-only step for reference
a = torch.tensor(acts[0], device=env.device)
for _ in range(5): env.step(a)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): env.step(a)
e1.record(); torch.cuda.synchronize()
print(f"device-resident step (L2 warm, no flush): {e0.elapsed_time(e1) / 50 * 1e3:.1f} us")
