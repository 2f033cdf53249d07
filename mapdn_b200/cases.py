"""Deterministic stand-ins for the MAPDN scenarios case33 / case141 / case322.

The reference's data (``model.p`` + three CSVs per scenario) is a Google-Drive download
that is not part of the repository (reference README.md:98-107), so the benchmark and the
parity tests run on synthetic networks with the published *sizes* of the scenarios
(reference README.md:299-303: 32 loads / 4 zones / 6 PVs; 84 / 9 / 22; 337 / 22 / 38) and
SURVEY.md §8d's recipe:

* ``case33``  - the IEEE 33-bus feeder of Baran & Wu (1989) that MAPDN's case33 is derived
  from (topology of reference ``img/case33.png``): published branch impedances and nominal
  loads, 12.66 kV, zones / PV buses as in the reference figure.
* ``case141`` / ``case322`` - random recursive radial feeders (bus k attaches to one of the
  previous min(k, 8) buses) with the scenario's bus / load / PV / zone counts.

Profiles are synthetic 3-minute series (PV bell curve x cloud factor, double-peak demand).
Everything is seeded; nothing is read from disk.
"""
from __future__ import annotations

import numpy as np

from .network import NetDesc, ProfileDesc

__all__ = ["case33", "case141", "case322", "make_case", "make_profiles", "synthetic_inputs",
           "SCENARIOS"]

# --- IEEE 33-bus (Baran & Wu 1989), 1-based bus numbers, ohms, kW / kvar -------------------
_BW_BRANCH = [
    (1, 2, 0.0922, 0.0470), (2, 3, 0.4930, 0.2511), (3, 4, 0.3660, 0.1864),
    (4, 5, 0.3811, 0.1941), (5, 6, 0.8190, 0.7070), (6, 7, 0.1872, 0.6188),
    (7, 8, 0.7114, 0.2351), (8, 9, 1.0300, 0.7400), (9, 10, 1.0440, 0.7400),
    (10, 11, 0.1966, 0.0650), (11, 12, 0.3744, 0.1238), (12, 13, 1.4680, 1.1550),
    (13, 14, 0.5416, 0.7129), (14, 15, 0.5910, 0.5260), (15, 16, 0.7463, 0.5450),
    (16, 17, 1.2890, 1.7210), (17, 18, 0.7320, 0.5740), (2, 19, 0.1640, 0.1565),
    (19, 20, 1.5042, 1.3554), (20, 21, 0.4095, 0.4784), (21, 22, 0.7089, 0.9373),
    (3, 23, 0.4512, 0.3083), (23, 24, 0.8980, 0.7091), (24, 25, 0.8960, 0.7011),
    (6, 26, 0.2030, 0.1034), (26, 27, 0.2842, 0.1447), (27, 28, 1.0590, 0.9337),
    (28, 29, 0.8042, 0.7006), (29, 30, 0.5075, 0.2585), (30, 31, 0.9744, 0.9630),
    (31, 32, 0.3105, 0.3619), (32, 33, 0.3410, 0.5302),
]
_BW_LOAD = {  # bus: (kW, kvar)
    2: (100, 60), 3: (90, 40), 4: (120, 80), 5: (60, 30), 6: (60, 20), 7: (200, 100),
    8: (200, 100), 9: (60, 20), 10: (60, 20), 11: (45, 30), 12: (60, 35), 13: (60, 35),
    14: (120, 80), 15: (60, 10), 16: (60, 20), 17: (60, 20), 18: (90, 40), 19: (90, 40),
    20: (90, 40), 21: (90, 40), 22: (90, 40), 23: (90, 50), 24: (420, 200), 25: (420, 200),
    26: (60, 25), 27: (60, 25), 28: (60, 20), 29: (120, 70), 30: (200, 600), 31: (150, 70),
    32: (210, 100), 33: (60, 40),
}
_BW_KV = 12.66


def baran_wu_nominal():
    """(NetDesc with nominal loads as metadata, p_load_mw, q_load_mvar) of the IEEE-33 feeder."""
    base_mva = 1.0
    zb = _BW_KV ** 2 / base_mva
    f = np.array([b[0] - 1 for b in _BW_BRANCH], np.int32)
    t = np.array([b[1] - 1 for b in _BW_BRANCH], np.int32)
    r = np.array([b[2] for b in _BW_BRANCH]) / zb
    x = np.array([b[3] for b in _BW_BRANCH]) / zb
    load_bus = np.array(sorted(_BW_LOAD), np.int32) - 1
    p = np.array([_BW_LOAD[b + 1][0] for b in load_bus]) * 1e-3
    q = np.array([_BW_LOAD[b + 1][1] for b in load_bus]) * 1e-3
    # zones of reference img/case33.png: 1-based bus numbers
    zone = np.zeros(33, np.int32)
    zone[7 - 1:18] = 1       # buses 7..18
    zone[19 - 1:22] = 2      # 19..22
    zone[23 - 1:25] = 3      # 23..25
    zone[26 - 1:33] = 4      # 26..33
    sgen_bus = np.array([13, 18, 22, 25, 29, 33], np.int32) - 1
    net = NetDesc(base_mva=base_mva, n_bus=33, slack_bus=0, slack_vm=1.0,
                  br_from=f, br_to=t, br_r=r, br_x=x, load_bus=load_bus,
                  sgen_bus=sgen_bus, sgen_zone=zone[sgen_bus], bus_zone=zone,
                  zone_names=["main", "zone1", "zone2", "zone3", "zone4"], name="case33")
    return net, p, q


def case33() -> NetDesc:
    return baran_wu_nominal()[0]


def _random_feeder(n_bus, n_load, n_zone, n_sgen, seed, name, r_scale=1.0):
    """Random recursive radial feeder: a trunk (zone "main") plus one random
    recursive sub-tree per zone (bus k attaches to one of the previous min(k, 8) buses of its
    zone), uneven zone sizes, PVs at the deepest buses of each zone."""
    rng = np.random.default_rng(seed)
    n_trunk = max(4, n_bus // 16)
    parent = np.full(n_bus, -1, np.int64)
    zone = np.zeros(n_bus, np.int32)
    for k in range(1, n_trunk):
        parent[k] = k - 1
    w = rng.uniform(0.5, 1.5, n_zone)
    sizes = np.maximum(4, np.floor(w / w.sum() * (n_bus - n_trunk)).astype(int))
    while sizes.sum() > n_bus - n_trunk:
        sizes[np.argmax(sizes)] -= 1
    while sizes.sum() < n_bus - n_trunk:
        sizes[np.argmin(sizes)] += 1
    attach = np.linspace(1, n_trunk - 1, n_zone).round().astype(int)
    k = n_trunk
    for z in range(n_zone):
        first = k
        parent[k] = attach[z]
        zone[k] = z + 1
        k += 1
        for _ in range(sizes[z] - 1):
            parent[k] = int(rng.integers(max(first, k - 8), k))
            zone[k] = z + 1
            k += 1
    assert k == n_bus
    f = parent[1:].astype(np.int32)
    t = np.arange(1, n_bus, dtype=np.int32)
    r = rng.uniform(6e-4, 9e-3, n_bus - 1) * 33.0 / n_bus * r_scale
    x = r * rng.uniform(0.3, 1.0, n_bus - 1)
    depth = np.zeros(n_bus, np.int64)
    for b in range(1, n_bus):
        depth[b] = depth[parent[b]] + 1
    sgen_bus = []
    per_zone = np.full(n_zone, n_sgen // n_zone)
    per_zone[: n_sgen - per_zone.sum()] += 1
    for z in range(n_zone):
        zb = np.nonzero(zone == z + 1)[0]
        order = zb[np.argsort(-depth[zb], kind="stable")]
        sgen_bus += list(order[:per_zone[z]])
    sgen_bus = np.array(sorted(sgen_bus), np.int32)
    assert len(sgen_bus) == n_sgen
    # loads: spread over the non-slack buses; case322 has more loads than buses
    cand = np.arange(1, n_bus)
    if n_load <= len(cand):
        load_bus = np.sort(rng.choice(cand, n_load, replace=False))
    else:
        load_bus = np.sort(np.concatenate([cand, rng.choice(cand, n_load - len(cand))]))
    return NetDesc(base_mva=1.0, n_bus=n_bus, slack_bus=0, slack_vm=1.0, br_from=f, br_to=t,
                   br_r=r, br_x=x, load_bus=load_bus.astype(np.int32), sgen_bus=sgen_bus,
                   sgen_zone=zone[sgen_bus], bus_zone=zone,
                   zone_names=["main"] + [f"zone{z + 1}" for z in range(n_zone)], name=name)


def case141() -> NetDesc:
    return _random_feeder(141, 84, 9, 22, seed=141, name="case141", r_scale=0.35)


def case322() -> NetDesc:
    return _random_feeder(322, 337, 22, 38, seed=322, name="case322")


# peak total demand / total PV capacity (MW) and action scale of each scenario
# (reference README.md:301-303; train.py:34-42)
SCENARIOS = {
    "case33": dict(make=case33, peak_load=3.5, pv_cap=8.75, action_scale=0.8, barrier="bowl"),
    "case141": dict(make=case141, peak_load=20.0, pv_cap=80.0, action_scale=0.6, barrier="l1"),
    "case322": dict(make=case322, peak_load=1.5, pv_cap=3.75, action_scale=0.8, barrier="l2"),
}


def make_case(name: str) -> NetDesc:
    return SCENARIOS[name]["make"]()


def _load_weights(net: NetDesc, name: str):
    if name == "case33":
        _, p, q = baran_wu_nominal()
        return p / p.sum(), q / p
    rng = np.random.default_rng(net.n_bus + 7)
    w = rng.uniform(0.5, 1.5, net.n_load)
    return w / w.sum(), rng.uniform(0.2, 0.6, net.n_load)


def make_profiles(name: str, n_days: int = 12, seed: int = 0) -> ProfileDesc:
    """Synthetic 3-minute profiles, ``n_days`` whole days + 1 row (480 rows/day)."""
    net = make_case(name)
    sc = SCENARIOS[name]
    rng = np.random.default_rng(seed)
    spd = 480
    T = n_days * spd + 1
    tod = (np.arange(T) % spd) / spd * 24.0                       # hour of day
    day = np.arange(T) // spd
    # PV: daylight bell, per-day cloud factor, slow per-PV flicker
    bell = np.clip(np.sin((tod - 6.0) / 12.0 * np.pi), 0.0, None) ** 1.5
    cloud = rng.uniform(0.35, 1.0, n_days + 1)[day]
    cap = rng.uniform(0.7, 1.3, net.n_sgen)
    cap *= sc["pv_cap"] / cap.sum()
    flick = 1.0 - 0.15 * np.abs(np.sin(np.outer(np.arange(T), rng.uniform(0.01, 0.05, net.n_sgen))))
    pv = (bell * cloud)[:, None] * cap[None, :] * flick
    # demand: double-peak curve in [0.3, 1.0], per-load jitter
    curve = 0.3 + 0.7 * (0.55 * np.exp(-0.5 * ((tod - 8.5) / 2.2) ** 2)
                         + 1.0 * np.exp(-0.5 * ((tod - 19.0) / 2.8) ** 2)
                         + 0.25)
    curve = np.clip(curve / curve.max(), 0.3, 1.0)
    w, pf = _load_weights(net, name)
    jit = 1.0 + 0.1 * np.sin(np.outer(np.arange(T), rng.uniform(0.02, 0.09, net.n_load))
                             + rng.uniform(0, 6.28, net.n_load)[None, :])
    load_p = curve[:, None] * (w * sc["peak_load"])[None, :] * jit
    load_q = load_p * pf[None, :]
    return ProfileDesc(pv=pv, load_p=load_p, load_q=load_q, steps_per_hour=20, n_days=n_days)


def synthetic_inputs(name: str, batch: int, seed: int = 0):
    """SURVEY §8d per-env inputs: dict(p_load, q_load, p_pv, action, s_max), fp64 [batch, n]."""
    net = make_case(name)
    sc = SCENARIOS[name]
    rng = np.random.default_rng(seed)
    w, pf = _load_weights(net, name)
    p_bar = w * sc["peak_load"]
    p_load = rng.uniform(0.3, 1.0, (batch, net.n_load)) * p_bar[None, :]
    q_load = p_load * (rng.uniform(0.2, 0.6, (batch, net.n_load)) if name != "case33"
                       else pf[None, :] * rng.uniform(0.8, 1.2, (batch, net.n_load)))
    cap = np.full(net.n_sgen, sc["pv_cap"] / net.n_sgen)
    p_pv = rng.uniform(0.0, 1.0, (batch, net.n_sgen)) * cap[None, :]
    s_max = 1.2 * cap
    a = rng.uniform(-sc["action_scale"], sc["action_scale"], (batch, net.n_sgen))
    return dict(p_load=p_load, q_load=q_load, p_pv=p_pv, action=a, s_max=s_max)
