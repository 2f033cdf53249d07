"""bench.py contract checks that need no GPU: the configurations are BASELINE.json's, the reference arm prints the
required JSON line, the core count comes from affinity / cgroup quota."""
import json
import os
import re
import subprocess
import sys

from conftest import ROOT


def _bench():
    sys.path.insert(0, ROOT)
    import bench
    return bench


def test_configs_are_the_baseline_json_configs():
    bench = _bench()
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    texts = base["configs"][1:]                       # configs[0] is the reference's 1-env CPU correctness run
    assert len(texts) == len(bench.CONFIGS) == 4
    for cid, txt in zip(sorted(bench.CONFIGS), texts):
        c = bench.CONFIGS[cid]
        assert c["scenario"] in txt and str(c["batch"]) in txt
        assert re.search(c["barrier"], txt, re.I)
        assert f"{c['n_gpus']}×B200" in txt or f"{c['n_gpus']}xB200" in txt
    assert bench.CONFIGS[bench.HEADLINE]["scenario"] == "case33" and bench.CONFIGS[bench.HEADLINE]["batch"] == 4096


def test_usable_cores_respects_affinity():
    bench = _bench()
    n, src = bench.usable_cores()
    assert 1 <= n <= len(os.sched_getaffinity(0)) and src in ("sched_getaffinity", "cgroup cpu quota")


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3",
                          "--cpu-sample", "4", "--cpu-cores", "2"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "env-steps/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "pandapower") and cb["cores"] == 2 and cb["one_core"]["value"] > 0
    assert "case33 x 4096 envs per GPU" in line["config"]["workload"]
    # rank != 0 of a torchrun launch exits quietly
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
