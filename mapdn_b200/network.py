"""Static network / profile descriptions consumed by the C-ABI (``include/mapdn_b200.h``).

The reference keeps the network as a pandapower ``net`` (dozens of DataFrames) that is
re-converted to the solver's array form on *every* ``pp.runpp`` call
(reference ``environments/var_voltage_control/voltage_control_env.py:124,165,557``).
Here the conversion happens once: a :class:`NetDesc` is the *ppci-level* description
(per-unit branch table, element->bus maps, zones) that ``mapdn_create`` copies to HBM.

Nothing in this module computes a power flow; it only validates and packs arrays.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

__all__ = ["NetDesc", "ProfileDesc", "population_std"]


def _f64(x, n=None, fill=None):
    if x is None:
        x = np.full(n, fill, dtype=np.float64)
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    if n is not None and a.shape != (n,):
        raise ValueError(f"expected shape ({n},), got {a.shape}")
    return a


def _i32(x, n=None):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.int32))
    if n is not None and a.shape != (n,):
        raise ValueError(f"expected shape ({n},), got {a.shape}")
    return a


@dataclass
class NetDesc:
    """Per-unit network description (the arrays pandapower's ``_pd2ppc`` would produce).

    Buses are numbered 0..n_bus-1 in ascending pandapower bus index (the order
    ``res_bus.sort_index()`` yields, reference voltage_control_env.py:584).
    Branch quantities are in p.u. on ``base_mva``: series ``r + jx``,
    total charging susceptance ``b``, total shunt conductance ``g``, off-nominal ``tap``
    (0 or 1 = nominal), ``shift`` in degrees. ``is_line`` marks rows that appear in
    ``res_line`` (transformers do not; reference voltage_control_env.py:599).
    """

    base_mva: float
    n_bus: int
    slack_bus: int
    slack_vm: float
    br_from: np.ndarray
    br_to: np.ndarray
    br_r: np.ndarray
    br_x: np.ndarray
    load_bus: np.ndarray
    sgen_bus: np.ndarray
    sgen_zone: np.ndarray          # zone id of each sgen (== sgen.name in the reference)
    bus_zone: np.ndarray           # zone id of each bus (bus.zone in the reference); 0 = "main"
    slack_va_deg: float = 0.0
    vm_init: Optional[float] = None  # flat-start magnitude; None -> slack_vm (pandapower init="auto")
    br_b: Optional[np.ndarray] = None
    br_g: Optional[np.ndarray] = None
    br_tap: Optional[np.ndarray] = None
    br_shift: Optional[np.ndarray] = None
    br_status: Optional[np.ndarray] = None
    br_is_line: Optional[np.ndarray] = None
    bus_gs: Optional[np.ndarray] = None   # MW consumed at 1 p.u.
    bus_bs: Optional[np.ndarray] = None   # MVAr injected at 1 p.u.
    load_scaling: Optional[np.ndarray] = None
    sgen_scaling: Optional[np.ndarray] = None
    zone_names: List[str] = field(default_factory=list)
    name: str = "net"

    def __post_init__(self):
        self.n_bus = int(self.n_bus)
        self.slack_bus = int(self.slack_bus)
        self.base_mva = float(self.base_mva)
        self.slack_vm = float(self.slack_vm)
        self.slack_va_deg = float(self.slack_va_deg)
        if self.vm_init is None:
            self.vm_init = self.slack_vm
        self.vm_init = float(self.vm_init)
        self.br_from = _i32(self.br_from)
        nbr = self.br_from.shape[0]
        self.br_to = _i32(self.br_to, nbr)
        self.br_r = _f64(self.br_r, nbr)
        self.br_x = _f64(self.br_x, nbr)
        self.br_b = _f64(self.br_b, nbr, 0.0)
        self.br_g = _f64(self.br_g, nbr, 0.0)
        self.br_tap = _f64(self.br_tap, nbr, 1.0)
        self.br_shift = _f64(self.br_shift, nbr, 0.0)
        self.br_status = np.ascontiguousarray(
            np.ones(nbr, np.uint8) if self.br_status is None else np.asarray(self.br_status, np.uint8))
        self.br_is_line = np.ascontiguousarray(
            np.ones(nbr, np.uint8) if self.br_is_line is None else np.asarray(self.br_is_line, np.uint8))
        self.bus_gs = _f64(self.bus_gs, self.n_bus, 0.0)
        self.bus_bs = _f64(self.bus_bs, self.n_bus, 0.0)
        self.bus_zone = _i32(self.bus_zone, self.n_bus)
        self.load_bus = _i32(self.load_bus)
        self.load_scaling = _f64(self.load_scaling, self.load_bus.shape[0], 1.0)
        self.sgen_bus = _i32(self.sgen_bus)
        ng = self.sgen_bus.shape[0]
        self.sgen_zone = _i32(self.sgen_zone, ng)
        self.sgen_scaling = _f64(self.sgen_scaling, ng, 1.0)
        for nm, a in (("br_from", self.br_from), ("br_to", self.br_to),
                      ("load_bus", self.load_bus), ("sgen_bus", self.sgen_bus)):
            if a.size and (a.min() < 0 or a.max() >= self.n_bus):
                raise ValueError(f"{nm} has a bus index outside [0, {self.n_bus})")
        if not (0 <= self.slack_bus < self.n_bus):
            raise ValueError("slack_bus out of range")

    # sizes ------------------------------------------------------------------
    @property
    def n_branch(self) -> int:
        return int(self.br_from.shape[0])

    @property
    def n_load(self) -> int:
        return int(self.load_bus.shape[0])

    @property
    def n_sgen(self) -> int:
        return int(self.sgen_bus.shape[0])

    @property
    def n_line(self) -> int:
        return int(self.br_is_line.sum())

    def zone_buses(self, agent: int) -> np.ndarray:
        """Ascending bus indices with ``bus.zone == sgen.name[agent]`` (reference :536)."""
        return np.nonzero(self.bus_zone == self.sgen_zone[agent])[0].astype(np.int32)

    @property
    def obs_dim(self) -> int:
        """4*z_max + 2 with the default state_space (reference :254-274)."""
        return int(max(4 * self.zone_buses(i).shape[0] + 2 for i in range(self.n_sgen)))

    @property
    def state_dim(self) -> int:
        return 4 * self.n_bus + 2 * self.n_sgen


def population_std(x: np.ndarray) -> np.ndarray:
    """``data.values.std(axis=0)`` (ddof=0), reference voltage_control_env.py:70-72."""
    return np.asarray(x, dtype=np.float64).std(axis=0)


@dataclass
class ProfileDesc:
    """Load / PV profile store: the three CSVs of the reference after scaling
    (reference voltage_control_env.py:407-438): row = one sensor interval, col = one element."""

    pv: np.ndarray        # [T, n_sgen]  MW
    load_p: np.ndarray    # [T, n_load]  MW
    load_q: np.ndarray    # [T, n_load]  MVAr
    steps_per_hour: int = 20          # 60 // time_delta, 3-min data (reference :389,396)
    n_days: Optional[int] = None      # (index[-1]-index[0]).days  (reference :395)

    def __post_init__(self):
        self.pv = np.ascontiguousarray(self.pv, dtype=np.float64)
        self.load_p = np.ascontiguousarray(self.load_p, dtype=np.float64)
        self.load_q = np.ascontiguousarray(self.load_q, dtype=np.float64)
        T = self.pv.shape[0]
        if self.load_p.shape[0] != T or self.load_q.shape[0] != T:
            raise ValueError("profile tables must have the same number of rows")
        if self.load_p.shape != self.load_q.shape:
            raise ValueError("load_p / load_q shape mismatch")
        self.steps_per_hour = int(self.steps_per_hour)
        if self.n_days is None:
            # timestamps are equally spaced: (T-1) intervals span this many whole days
            self.n_days = int((T - 1) // (24 * self.steps_per_hour))
        self.n_days = int(self.n_days)

    @property
    def n_rows(self) -> int:
        return int(self.pv.shape[0])

    # statistics the reference derives in __init__ -----------------------------
    @property
    def pv_std(self) -> np.ndarray:
        return population_std(self.pv) / 100.0          # reference :72

    @property
    def load_p_std(self) -> np.ndarray:
        return population_std(self.load_p) / 100.0      # reference :70

    @property
    def load_q_std(self) -> np.ndarray:
        return population_std(self.load_q) / 100.0      # reference :71

    @property
    def s_max(self) -> np.ndarray:
        return 1.2 * self.pv.max(axis=0)                # reference :515-521
