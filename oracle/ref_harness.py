"""ORACLE - TEST INFRASTRUCTURE ONLY. Runs the reference's OWN, UNMODIFIED env code
(``/root/reference/environments/var_voltage_control/voltage_control_env.py``) in a container that has no pandapower,
to produce golden trajectories for ``tests/golden/ref_env_*.npz`` (``scripts/make_reference_golden.py``).

What is real and what is substituted:

* REAL: every line of ``VoltageControl`` - ``__init__`` (CSV loading, std / s_max derivation, action space),
  ``reset`` / ``manual_reset`` (retry loop, start selection, episode slicing), ``step`` (deepcopy, failure branch,
  sequencing of reward / next row / step counter), ``_take_action`` / ``_clip_reactive_power``, ``_calc_reward`` (the 11
  info entries), ``_set_demand_and_pv``, ``get_obs`` (clusters, the chained ``+=``, padding), ``get_state``, and the five
  ``voltage_barrier/*.py`` functions - imported from ``/root/reference`` and executed as is.
* SUBSTITUTED (this module): the ``pandapower`` import. ``pp.runpp`` is ``oracle.pandapower_nr.PandapowerEquivalent``
  (the NumPy restatement of pandapower 2.7.0's Newton-Raphson; pinned separately by the literature KATs), ``pp.from_pickle``
  reads a small pickle written by :func:`write_reference_data` (NOT pandapower's own pickle layout), and the handful of
  table operations the env performs on ``net.bus / load / sgen / res_bus / res_line / res_sgen`` are served by
  :class:`Frame`, a minimal table with the **pandas 1.1.3** behaviour the reference was written against
  (``environment.yml:134``): ``frame.loc[label]`` of a single-dtype table is a VIEW, so the chained
  ``zone_buses.loc[bus]["p_mw"] += pv`` of ``get_obs`` (``voltage_control_env.py:239-244``) writes through
  (pandas' ``DataFrame.xs`` -> ``BlockManager.fast_xs`` returns a view of the single float block; the write only triggers a
  ``SettingWithCopyWarning``). The pandas in this image (3.x, copy-on-write) would silently drop that write - the
  reference's observations depend on the pandas version; the product follows the pinned one. ``Frame(view_rows=False)``
  reproduces the copy-on-write behaviour for the test that documents the difference.
* The process-global ``np.random`` stream of the reference is replaced, while the reference code runs, by the draws of the
  device RNG's NumPy mirror (``oracle.philox_ref``) for the same (seed, env id, episode, step / attempt): the reference then
  consumes exactly the random numbers the CUDA env consumes, which makes whole noisy trajectories comparable
  (``np.random.choice`` x3 per start, ``np.random.randn`` x3 per profile row, ``np.random.uniform`` per reset action).

Nothing here is imported by the product; only ``scripts/make_reference_golden.py`` and ``tests/`` use it, and only where
``/root/reference`` exists (this container, not the GPU box).
"""
from __future__ import annotations

import contextlib
import io
import os
import pickle
import sys
import types

import numpy as np
import pandas as pd

from . import philox_ref as rng
from .pandapower_nr import PandapowerEquivalent

REFERENCE_ROOT = os.environ.get("MAPDN_REFERENCE_ROOT", "/root/reference")
INFO_KEYS = ("percentage_of_v_out_of_control", "percentage_of_lower_than_lower_v",
             "percentage_of_higher_than_upper_v", "totally_controllable_ratio",
             "average_voltage_deviation", "average_voltage", "max_voltage_drop_deviation",
             "max_voltage_rise_deviation", "total_line_loss", "q_loss", "destroy")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "environments", "var_voltage_control", "voltage_control_env.py"))


# ----------------------------------------------------------------------------------------------------------------
# minimal table (pandas 1.1.3 semantics of the operations voltage_control_env.py performs)
# ----------------------------------------------------------------------------------------------------------------
class _Row:
    def __init__(self, frame, pos, view):
        self._f, self._pos, self._view = frame, pos, view
        self._own = None if view else {k: v[pos] for k, v in frame._cols.items()}

    def __getitem__(self, col):
        return self._f._cols[col][self._pos] if self._view else self._own[col]

    def __setitem__(self, col, value):
        if self._view:
            self._f._cols[col][self._pos] = value          # writes through to the table (pandas 1.1.3: fast_xs view)
        else:
            self._own[col] = value                         # copy-on-write pandas: the temporary row absorbs the write


class _Loc:
    def __init__(self, frame):
        self._f = frame

    def __getitem__(self, key):
        f = self._f
        if isinstance(key, tuple):                         # .loc[:, "col"]
            rows, col = key
            if not (isinstance(rows, slice) and rows == slice(None)):
                raise NotImplementedError("Frame.loc[rows, col]: only ':' rows")
            return f[col]
        if isinstance(key, pd.Series) and key.dtype == bool:     # boolean mask, aligned on the index like pandas
            mask = key.reindex(f.index).to_numpy(dtype=bool)
            return Frame(f.index[mask], {k: v[mask] for k, v in f._cols.items()}, view_rows=f._view_rows)
        if isinstance(key, np.ndarray) and key.dtype == bool:
            return Frame(f.index[key], {k: v[key] for k, v in f._cols.items()}, view_rows=f._view_rows)
        pos = np.nonzero(f.index == key)[0]
        if pos.size != 1:
            raise KeyError(key)
        return _Row(f, int(pos[0]), f._view_rows and f._single_dtype())


class Frame:
    def __init__(self, index, cols, view_rows=True):
        self.index = np.asarray(index).copy()
        self._cols = {k: np.array(v, copy=True) for k, v in cols.items()}
        self._view_rows = view_rows
        for k, v in self._cols.items():
            if v.shape != (len(self.index),):
                raise ValueError(f"column {k}: shape {v.shape} for {len(self.index)} rows")

    @property
    def columns(self):
        return list(self._cols)

    def _single_dtype(self):
        return len({v.dtype for v in self._cols.values()}) == 1

    def __len__(self):
        return len(self.index)

    def __getitem__(self, col):
        return pd.Series(self._cols[col].copy(), index=self.index.copy(), name=col)

    def __setitem__(self, col, value):
        if isinstance(value, pd.Series):
            value = value.reindex(self.index).to_numpy()
        arr = np.asarray(value)
        if arr.ndim == 0:
            arr = np.full(len(self), arr)
        if arr.shape != (len(self),):
            raise ValueError(f"column {col}: shape {arr.shape} for {len(self)} rows")
        self._cols[col] = arr.copy()

    def sort_index(self):
        order = np.argsort(self.index, kind="stable")
        return Frame(self.index[order], {k: v[order] for k, v in self._cols.items()}, view_rows=self._view_rows)

    @property
    def loc(self):
        return _Loc(self)

    def __repr__(self):
        return repr(pd.DataFrame(self._cols, index=self.index))


class _Static:
    """Per-network constants; shared by every deepcopy of the net (the reference deep-copies the net on every step)."""

    def __init__(self, desc):
        self.desc, self.pf = desc, PandapowerEquivalent(desc)

    def __deepcopy__(self, memo):
        return self


class Net:
    """The slice of a pandapowerNet the env touches (attribute and item access)."""

    def __init__(self, desc, zone_names, view_rows=True):
        n, nl, ng = desc.n_bus, desc.n_load, desc.n_sgen
        names = np.array([zone_names[z] for z in desc.bus_zone], dtype=object)
        self._static = _Static(desc)
        self._view_rows = view_rows
        self.bus = Frame(np.arange(n), dict(zone=names), view_rows)
        self.load = Frame(np.arange(nl), dict(bus=desc.load_bus.astype(np.int64), p_mw=np.zeros(nl), q_mvar=np.zeros(nl),
                                              scaling=desc.load_scaling), view_rows)
        self.sgen = Frame(np.arange(ng), dict(name=np.array([zone_names[z] for z in desc.sgen_zone], dtype=object),
                                              bus=desc.sgen_bus.astype(np.int64), p_mw=np.zeros(ng), q_mvar=np.zeros(ng),
                                              scaling=desc.sgen_scaling), view_rows)
        self.res_bus = Frame(np.arange(n), {k: np.full(n, np.nan) for k in ("vm_pu", "va_degree", "p_mw", "q_mvar")}, view_rows)
        self.res_line = Frame(np.arange(desc.n_line), dict(pl_mw=np.full(desc.n_line, np.nan)), view_rows)
        self.res_sgen = Frame(np.arange(ng), dict(p_mw=np.full(ng, np.nan), q_mvar=np.full(ng, np.nan)), view_rows)
        self.converged = False

    def __getitem__(self, k):
        return getattr(self, k)

    def __setitem__(self, k, v):
        setattr(self, k, v)


# ----------------------------------------------------------------------------------------------------------------
# the substitute ``pandapower`` package
# ----------------------------------------------------------------------------------------------------------------
class ppException(Exception):
    pass


class LoadflowNotConverged(ppException):
    pass


VIEW_ROWS = True        # module switch read by from_pickle: False = copy-on-write pandas behaviour (see the module docstring)


def runpp(net, **kwargs):
    """pp.runpp(net) with default arguments on the substitute net: fills res_bus / res_line / res_sgen, raises
    LoadflowNotConverged (a ppException) when Newton-Raphson does not converge (results reset to NaN, like pandapower's
    reset_results before the solve)."""
    st = net._static
    d = st.desc
    res = st.pf.runpp(net.load["p_mw"].to_numpy(), net.load["q_mvar"].to_numpy(),
                      net.sgen["p_mw"].to_numpy(), net.sgen["q_mvar"].to_numpy())
    vr = net._view_rows
    if not res.converged:
        n = d.n_bus
        net.res_bus = Frame(np.arange(n), {k: np.full(n, np.nan) for k in ("vm_pu", "va_degree", "p_mw", "q_mvar")}, vr)
        net.res_line = Frame(np.arange(d.n_line), dict(pl_mw=np.full(d.n_line, np.nan)), vr)
        net.converged = False
        raise LoadflowNotConverged("Power Flow nr did not converge after 10 iterations!")
    net.res_bus = Frame(np.arange(d.n_bus), dict(vm_pu=res.vm_pu, va_degree=res.va_degree, p_mw=res.p_mw,
                                                 q_mvar=res.q_mvar), vr)
    net.res_line = Frame(np.arange(d.n_line), dict(pl_mw=res.pl_mw), vr)
    net.res_sgen = Frame(np.arange(d.n_sgen), dict(p_mw=net.sgen["p_mw"].to_numpy() * d.sgen_scaling,
                                                   q_mvar=net.sgen["q_mvar"].to_numpy() * d.sgen_scaling), vr)
    net.converged = True


def from_pickle(path):
    with open(path, "rb") as f:
        blob = pickle.load(f)
    from mapdn_b200.network import NetDesc
    return Net(NetDesc(**blob["net_desc"]), blob["zone_names"], view_rows=VIEW_ROWS)


def install_substitute_pandapower():
    """Registers the substitute ``pandapower`` package (and the plotting sub-modules ``pf_res_plot.py`` imports at module
    level) in ``sys.modules``. Refuses to shadow a real pandapower."""
    if "pandapower" in sys.modules and not getattr(sys.modules["pandapower"], "_MAPDN_SUBSTITUTE", False):
        raise RuntimeError("a real pandapower is already imported - use scripts/pin_with_pandapower.py instead")
    pp = types.ModuleType("pandapower")
    pp._MAPDN_SUBSTITUTE = True
    pp.ppException, pp.LoadflowNotConverged, pp.runpp, pp.from_pickle = ppException, LoadflowNotConverged, runpp, from_pickle
    pp.__path__ = []
    mods = {"pandapower": pp}

    def sub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__path__ = []
        mods[name] = m
        return m

    def _unavailable(*a, **k):
        raise NotImplementedError("plotting is out of scope of the substitute pandapower")

    sub("pandapower.plotting")
    sub("pandapower.plotting.generic_geodata", create_generic_coordinates=_unavailable)
    sub("pandapower.plotting.plotly")
    sub("pandapower.plotting.plotly.mapbox_plot")
    sub("pandapower.plotting.plotly.traces", create_bus_trace=_unavailable, create_line_trace=_unavailable,
        create_trafo_trace=_unavailable, draw_traces=_unavailable, version_check=_unavailable)
    sub("pandapower.run", runpp=runpp)
    sys.modules.update(mods)
    return pp


def import_reference_env():
    """The reference's VoltageControl class, imported from /root/reference with the substitute pandapower."""
    if not reference_available():
        raise FileNotFoundError(f"{REFERENCE_ROOT} is not available here")
    install_substitute_pandapower()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from environments.var_voltage_control.voltage_control_env import VoltageControl
    return VoltageControl


# ----------------------------------------------------------------------------------------------------------------
# data directory in the reference's layout
# ----------------------------------------------------------------------------------------------------------------
def zone_names_of(net):
    nz = int(max(net.bus_zone.max(), net.sgen_zone.max())) + 1
    names = list(net.zone_names) if getattr(net, "zone_names", None) else []
    return names if len(names) >= nz else ["main"] + [f"zone{k}" for k in range(1, nz)]


def write_reference_data(path, net, prof, start="2012-01-01 00:00:00"):
    """model.p (substitute layout, see module docstring) + pv_active.csv / load_active.csv / load_reactive.csv in the
    reference's CSV layout (first column = timestamp, one column per element; voltage_control_env.py:407-438)."""
    import dataclasses
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "model.p"), "wb") as f:
        pickle.dump(dict(net_desc=dataclasses.asdict(net), zone_names=zone_names_of(net)), f)
    minutes = 60 // prof.steps_per_hour
    t = pd.date_range(start, periods=prof.n_rows, freq=f"{minutes}min").strftime("%Y-%m-%d %H:%M:%S")
    for name, arr in (("pv_active.csv", prof.pv), ("load_active.csv", prof.load_p), ("load_reactive.csv", prof.load_q)):
        with open(os.path.join(path, name), "w") as f:
            f.write("time," + ",".join(str(k) for k in range(arr.shape[1])) + "\n")
            for ts, row in zip(t, arr):
                f.write(ts + "," + ",".join(repr(float(v)) for v in row) + "\n")     # repr: exact fp64 round trip


# ----------------------------------------------------------------------------------------------------------------
# np.random replaced by the device RNG's mirror while the reference code runs
# ----------------------------------------------------------------------------------------------------------------
class _Draws:
    def __init__(self, seed, env_id, n_sgen, n_load):
        self.seed, self.env_id, self.ng, self.nl = int(seed), int(env_id), n_sgen, n_load
        self.episode = 0
        self.env = None
        self.begin_reset()

    def begin_reset(self):
        self.episode += 1
        self.mode, self.n_choice, self.n_randn, self.n_uniform = "reset", 0, 0, 0

    def begin_step(self):
        self.mode, self.n_randn = "step", 0

    # the three np.random entry points the reference uses
    def choice(self, n):
        # voltage_control_env.py:381-398 asks for hour (24), day (pv_days - episode_days), interval (60 // time_delta) one
        # by one; the device draws the three from ONE Philox call, so the ranges are derived from the env's data here
        # and the call's own argument is checked against them
        attempt, which = divmod(self.n_choice, 3)
        self.n_choice += 1
        env = self.env
        pv = env.pv_data
        pv_days = (pv.index[-1] - pv.index[0]).days
        time_delta = (pv.index[1] - pv.index[0]).seconds // 60
        episode_days = (env.episode_limit // (24 * (60 // time_delta))) + 1
        ranges = (24, pv_days - episode_days, 60 // time_delta)
        assert n == ranges[which], (n, ranges, which)
        return rng.start_time(self.seed, self.env_id, self.episode, attempt, ranges[1], ranges[2])[which]

    def randn(self, *shape):
        if self.mode == "reset":
            attempt, part = divmod(self.n_randn, 3)
            c1 = rng.RESET_FLAG | attempt
        else:
            part = self.n_randn % 3
            c1 = int(self.env.steps)            # _set_demand_and_pv runs before `self.steps += 1` (:199-202)
        self.n_randn += 1
        z = rng.half_normal(self.seed, self.env_id, self.episode, c1, np.arange(self.ng + 2 * self.nl))
        lo, hi = ((0, self.ng), (self.ng, self.ng + self.nl), (self.ng + self.nl, self.ng + 2 * self.nl))[part]
        out = z[lo:hi]
        assert tuple(shape) == out.shape, (shape, out.shape)
        return out

    def uniform(self, low=0.0, high=1.0, size=None):
        attempt = self.n_uniform
        self.n_uniform += 1
        n = int(np.prod(size))
        assert n == self.ng
        return rng.uniform_action(self.seed, self.env_id, self.episode, attempt, n, low, high).reshape(size)


@contextlib.contextmanager
def _patched_np_random(draws):
    saved = {k: getattr(np.random, k) for k in ("choice", "randn", "uniform", "seed")}
    np.random.choice, np.random.randn, np.random.uniform = draws.choice, draws.randn, draws.uniform
    np.random.seed = lambda *a, **k: None
    try:
        yield
    finally:
        for k, v in saved.items():
            setattr(np.random, k, v)


# ----------------------------------------------------------------------------------------------------------------
# driver
# ----------------------------------------------------------------------------------------------------------------
REF_DEFAULT_ARGS = dict(voltage_barrier_type="l1", voltage_weight=1.0, q_weight=0.1, line_weight=None, dq_dv_weight=None,
                        history=1, pv_scale=1.0, demand_scale=1.0,
                        state_space=["pv", "demand", "reactive", "vm_pu", "va_degree"], v_upper=1.05, v_lower=0.95,
                        episode_limit=240, action_scale=0.8, action_bias=0.0, mode="distributed", reset_action=True,
                        seed=0)        # reference args/env_args/var_voltage_control.yaml + train.py:34-46


class ReferenceRun:
    """One reference env instance driven with the device RNG's draws of global env id ``env_id``."""

    def __init__(self, data_path, net, env_args, env_id=0, quiet=True, view_rows=True):
        global VIEW_ROWS
        VoltageControl = import_reference_env()
        args = dict(REF_DEFAULT_ARGS)
        args.update(env_args)
        args["data_path"] = data_path
        self.args, self.quiet = args, quiet
        self.draws = _Draws(args["seed"], env_id, net.n_sgen, net.n_load)
        VIEW_ROWS = view_rows
        try:
            with self._ctx():
                # __init__ ends with self.reset() (:85): that is episode 1 of this env
                env = VoltageControl.__new__(VoltageControl)
                self.draws.env = env
                env.__init__(args)
        finally:
            VIEW_ROWS = True
        self.env = env

    @contextlib.contextmanager
    def _ctx(self):
        with _patched_np_random(self.draws):
            if self.quiet:
                with contextlib.redirect_stdout(io.StringIO()):
                    yield
            else:
                yield

    def _snapshot(self):
        e = self.env
        return np.array(e.get_obs()), np.array(e.get_state())

    def initial(self):
        """Observations / state after the reset that ``__init__`` performs."""
        with self._ctx():
            return self._snapshot()

    def reset(self):
        self.draws.begin_reset()
        with self._ctx():
            obs, state = self.env.reset()
        return np.array(obs), np.array(state)

    def reset_keep_time(self):
        """``reset(reset_time=False)``: a new episode at the previous start (noise and reset action re-drawn, :109-122)."""
        self.draws.begin_reset()
        with self._ctx():
            obs, state = self.env.reset(reset_time=False)
        return np.array(obs), np.array(state)

    def manual_reset(self, day, hour, interval):
        self.draws.begin_reset()
        with self._ctx():
            obs, state = self.env.manual_reset(day, hour, interval)
        return np.array(obs), np.array(state)

    def step(self, actions, add_noise=True):
        self.draws.begin_step()
        with self._ctx():
            reward, terminated, info = self.env.step(np.asarray(actions, np.float64), add_noise=add_noise)
            info = dict(info)                  # _calc_reward's mutable default dict is shared between calls (:574)
            obs, state = self._snapshot()
        return float(reward), bool(terminated), np.array([float(info[k]) for k in INFO_KEYS]), obs, state

    @property
    def start(self):
        e = self.env
        spd = 60 // e.time_delta if hasattr(e, "time_delta") else None
        return (int(e._episode_start_day), int(e._episode_start_hour), int(e._episode_start_interval)), spd
