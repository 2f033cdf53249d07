"""Reference file formats -> :class:`NetDesc` / :class:`ProfileDesc` without pandapower.

A MAPDN scenario directory holds ``model.p`` (a pickled pandapower net, reference
``voltage_control_env.py:400-405``) and ``pv_active.csv`` / ``load_active.csv`` / ``load_reactive.csv``
(``:407-438``: column 0 = timestamp, one column per sgen / load in table order, 3-minute rows).

``net_from_tables`` restates pandapower 2.7.0's ``_pd2ppc`` element -> per-unit conversion for the element
types MAPDN uses: lines, two-winding transformers (tap changer, ``trafo_model="t"``
T -> pi conversion, phase shift ignored as with ``calculate_voltage_angles="auto"`` below 70 kV), shunts,
loads, sgens, one ext_grid. PARITY UNPINNED for this module: neither pandapower nor a sample ``model.p`` is
available in this image; the formulas are checked against hand-computed cases only (tests/test_ingest.py).
Unsupported content raises ``NotImplementedError`` instead of guessing.
"""
from __future__ import annotations

import io
import os
import pickle
from typing import Dict, Tuple

import numpy as np

from .network import NetDesc, ProfileDesc

__all__ = ["load_profiles", "load_network", "load_scenario", "net_from_tables", "read_model_pickle",
           "save_scenario_npz", "load_scenario_npz", "net_to_json", "net_from_json", "SCENARIO_NPZ"]


# --------------------------------------------------------------------------------------------------
# CSV profiles
# --------------------------------------------------------------------------------------------------
def _read_profile_csv(path: str):
    import pandas as pd
    df = pd.read_csv(path, index_col=None)
    t = pd.to_datetime(df.iloc[:, 0])
    return t, df.iloc[:, 1:].to_numpy(dtype=np.float64)


def load_profiles(data_path: str, pv_scale: float = 1.0, demand_scale: float = 1.0) -> ProfileDesc:
    """reference ``_load_pv_data`` / ``_load_active_demand_data`` / ``_load_reactive_demand_data``."""
    t, pv = _read_profile_csv(os.path.join(data_path, "pv_active.csv"))
    _, lp = _read_profile_csv(os.path.join(data_path, "load_active.csv"))
    _, lq = _read_profile_csv(os.path.join(data_path, "load_reactive.csv"))
    time_delta = int((t.iloc[1] - t.iloc[0]).seconds // 60)              # reference :396
    n_days = int((t.iloc[-1] - t.iloc[0]).days)                          # reference :395
    return ProfileDesc(pv=pv * pv_scale, load_p=lp * demand_scale, load_q=lq * demand_scale,
                       steps_per_hour=60 // time_delta, n_days=n_days)


# --------------------------------------------------------------------------------------------------
# model.p
# --------------------------------------------------------------------------------------------------
class _Stub(dict):
    """Stand-in for pandapower classes (pandapowerNet is a dict subclass) while unpickling."""

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split(".")[0] == "pandapower":
            return type(name, (_Stub,), {})
        return super().find_class(module, name)


def read_model_pickle(path: str) -> Dict[str, object]:
    """``model.p`` -> dict of element tables (pandas DataFrames) + scalars (``sn_mva``, ``f_hz``).

    Handles both layouts pandapower writes: ``pp.to_pickle`` (tables stored as
    ``{"DF": df.to_dict("split"), "dtypes": ...}``) and a plain pickle of the net object."""
    import pandas as pd
    with open(path, "rb") as f:
        raw = f.read()
    try:
        obj = _Unpickler(io.BytesIO(raw)).load()
    except Exception as first:
        try:
            obj = pd.read_pickle(io.BytesIO(raw))
        except Exception:
            raise RuntimeError(f"cannot unpickle {path} without pandapower: {first!r}") from first
    out = {}
    for k, v in dict(obj).items():
        if isinstance(v, dict) and "DF" in v:
            d = v["DF"]
            df = pd.DataFrame(data=d.get("data"), index=d.get("index"), columns=d.get("columns"))
            for col, dt in (v.get("dtypes") or {}).items():
                try:
                    df[col] = df[col].astype(dt)
                except Exception:
                    pass
            out[k] = df
        else:
            out[k] = v
    return out


def _col(df, name, default):
    if df is None or name not in df.columns:
        return np.full(0 if df is None else len(df), default, dtype=np.float64)
    return np.nan_to_num(df[name].to_numpy(dtype=np.float64), nan=default)


def net_from_tables(t: Dict[str, object], name: str = "net") -> NetDesc:
    """pandapower element tables -> per-unit :class:`NetDesc`."""
    def tab(k):
        v = t.get(k)
        return v if v is not None and len(v) else None

    bus = tab("bus")
    if bus is None:
        raise ValueError("net has no bus table")
    bus = bus.sort_index()
    if not bool(np.all(bus["in_service"].to_numpy(dtype=bool))):
        raise NotImplementedError("out-of-service buses")
    if tab("gen") is not None:
        raise NotImplementedError("`gen` elements (PV buses) - MAPDN nets use sgen only")
    for k in ("trafo3w", "impedance", "ward", "xward", "dcline", "storage", "motor", "asymmetric_load"):
        if tab(k) is not None:
            raise NotImplementedError(f"`{k}` elements")
    base = float(t.get("sn_mva", 1.0) or 1.0)
    f_hz = float(t.get("f_hz", 50.0) or 50.0)
    idx = {int(b): i for i, b in enumerate(bus.index)}
    nb = len(bus)
    vn = bus["vn_kv"].to_numpy(dtype=np.float64)
    zones = [str(z) for z in bus["zone"].tolist()] if "zone" in bus.columns else ["main"] * nb
    zone_names = ["main"] + sorted({z for z in zones if z != "main"})
    zid = {z: i for i, z in enumerate(zone_names)}

    open_lines = set()
    sw = tab("switch")
    if sw is not None:
        for _, r in sw.iterrows():
            if r["et"] == "b" and bool(r["closed"]):
                raise NotImplementedError("closed bus-bus switches (bus fusion)")
            if r["et"] == "l" and not bool(r["closed"]):
                open_lines.add(int(r["element"]))
            if r["et"] == "t" and not bool(r["closed"]):
                raise NotImplementedError("open transformer switches")

    f, to, r_, x_, b_, g_, tap, status, is_line = [], [], [], [], [], [], [], [], []
    line = tab("line")
    if line is not None:
        line = line.sort_index()
        fb = np.array([idx[int(b)] for b in line["from_bus"]]); tb = np.array([idx[int(b)] for b in line["to_bus"]])
        length, par = _col(line, "length_km", 1.0), _col(line, "parallel", 1.0)
        base_r = vn[fb] ** 2 / base                                       # A.2: baseR = vn_kv(from_bus)^2 / baseMVA
        f += list(fb); to += list(tb)
        r_ += list(_col(line, "r_ohm_per_km", 0.0) * length / par / base_r)
        x_ += list(_col(line, "x_ohm_per_km", 0.0) * length / par / base_r)
        b_ += list(2 * np.pi * f_hz * _col(line, "c_nf_per_km", 0.0) * 1e-9 * length * par * base_r)
        g_ += list(_col(line, "g_us_per_km", 0.0) * 1e-6 * length * par * base_r)
        tap += [1.0] * len(line)
        ins = line["in_service"].to_numpy(dtype=bool)
        status += [bool(s) and (int(i) not in open_lines) for s, i in zip(ins, line.index)]
        is_line += [1] * len(line)
    trafo = tab("trafo")
    if trafo is not None:
        trafo = trafo.sort_index()
        hv = np.array([idx[int(b)] for b in trafo["hv_bus"]]); lv = np.array([idx[int(b)] for b in trafo["lv_bus"]])
        vn_hv, vn_lv = _col(trafo, "vn_hv_kv", 0.0).copy(), _col(trafo, "vn_lv_kv", 0.0).copy()
        sn = _col(trafo, "sn_mva", 1.0); par = _col(trafo, "parallel", 1.0)
        pos = trafo["tap_pos"].to_numpy(dtype=np.float64) if "tap_pos" in trafo.columns else np.full(len(trafo), np.nan)
        neu = _col(trafo, "tap_neutral", 0.0); step = _col(trafo, "tap_step_percent", 0.0)
        side = trafo["tap_side"].tolist() if "tap_side" in trafo.columns else [None] * len(trafo)
        for k in range(len(trafo)):                                       # tap changer (A.2)
            if not np.isnan(pos[k]) and side[k] in ("hv", "lv"):
                fac = 1.0 + (pos[k] - neu[k]) * step[k] / 100.0
                if side[k] == "hv":
                    vn_hv[k] *= fac
                else:
                    vn_lv[k] *= fac
        ratio = (vn_hv / vn_lv) / (vn[hv] / vn[lv])
        tap_lv = (vn_lv / vn[lv]) ** 2 * base
        z = _col(trafo, "vk_percent", 0.0) / 100.0 / sn * tap_lv
        r = _col(trafo, "vkr_percent", 0.0) / 100.0 / sn * tap_lv
        x = np.sign(z) * np.sqrt(np.maximum(z ** 2 - r ** 2, 0.0))
        r, x = r / par, x / par
        # magnetising branch, pandapower convention y = b - j g (p.u.)
        base_r = vn[lv] ** 2 / base
        vnl2 = _col(trafo, "vn_lv_kv", 1.0) ** 2
        pfe = _col(trafo, "pfe_kw", 0.0) * 1e-3
        i0 = _col(trafo, "i0_percent", 0.0)
        b_real = pfe / vnl2 * base_r
        b_img = np.sqrt(np.maximum((i0 / 100.0 * sn) ** 2 - pfe ** 2, 0.0)) * base_r / vnl2
        y = (-b_real * 1j - b_img * np.sign(i0)) / (vn_lv / vn[lv]) ** 2 * par
        # trafo_model="t": T (z/2, y, z/2) -> pi by wye-delta where y != 0
        zs = r + 1j * x
        yb = y.copy()
        nz = y != 0
        if np.any(nz):
            za = zs[nz] / 2.0
            zc = -1j / y[nz]
            sd = za * za + 2.0 * za * zc
            zs[nz] = sd / zc
            yb[nz] = -2j / (sd / za)
        f += list(hv); to += list(lv)
        r_ += list(zs.real); x_ += list(zs.imag); b_ += list(yb.real); g_ += list(-yb.imag)
        tap += list(ratio)
        status += list(trafo["in_service"].to_numpy(dtype=bool))
        is_line += [0] * len(trafo)

    gs, bs = np.zeros(nb), np.zeros(nb)
    sh = tab("shunt")
    if sh is not None:
        for _, rr in sh.iterrows():
            if not bool(rr["in_service"]):
                continue
            b = idx[int(rr["bus"])]
            k = float(rr.get("step", 1) or 1) * (vn[b] / float(rr.get("vn_kv", vn[b]) or vn[b])) ** 2
            gs[b] += float(rr["p_mw"]) * k
            bs[b] -= float(rr["q_mvar"]) * k

    load = tab("load")
    if load is not None:
        load = load.sort_index()
        if np.any(_col(load, "const_z_percent", 0.0) != 0) or np.any(_col(load, "const_i_percent", 0.0) != 0):
            raise NotImplementedError("voltage-dependent (ZIP) loads")
        load_bus = np.array([idx[int(b)] for b in load["bus"]], np.int32)
        load_scaling = _col(load, "scaling", 1.0) * load["in_service"].to_numpy(dtype=np.float64)
    else:
        load_bus, load_scaling = np.zeros(0, np.int32), np.zeros(0)
    sgen = tab("sgen")
    if sgen is None:
        raise ValueError("net has no sgen (no agents)")
    sgen = sgen.sort_index()
    sgen_bus = np.array([idx[int(b)] for b in sgen["bus"]], np.int32)
    sgen_scaling = _col(sgen, "scaling", 1.0) * sgen["in_service"].to_numpy(dtype=np.float64)
    sgen_zone = np.array([zid.get(str(nm), -1) for nm in sgen["name"]], np.int32)   # sgen.name == zone (ref :532)
    eg = tab("ext_grid")
    if eg is None or int(np.sum(eg["in_service"].to_numpy(dtype=bool))) != 1:
        raise NotImplementedError("exactly one in-service ext_grid is supported")
    eg = eg[eg["in_service"].to_numpy(dtype=bool)].iloc[0]
    vm = float(eg["vm_pu"])
    return NetDesc(base_mva=base, n_bus=nb, slack_bus=idx[int(eg["bus"])], slack_vm=vm,
                   slack_va_deg=float(eg.get("va_degree", 0.0) or 0.0), vm_init=vm,
                   br_from=np.array(f, np.int32), br_to=np.array(to, np.int32), br_r=np.array(r_), br_x=np.array(x_),
                   br_b=np.array(b_), br_g=np.array(g_), br_tap=np.array(tap), br_status=np.array(status, np.uint8),
                   br_is_line=np.array(is_line, np.uint8), bus_gs=gs, bus_bs=bs,
                   bus_zone=np.array([zid[z] for z in zones], np.int32), load_bus=load_bus, load_scaling=load_scaling,
                   sgen_bus=sgen_bus, sgen_zone=sgen_zone, sgen_scaling=sgen_scaling, zone_names=zone_names, name=name)


def load_network(path: str) -> NetDesc:
    return net_from_tables(read_model_pickle(path), name=os.path.basename(os.path.dirname(path)) or "net")


# --------------------------------------------------------------------------------------------------
# Export format: the per-unit network as JSON, a whole scenario as one NPZ
# --------------------------------------------------------------------------------------------------
SCENARIO_NPZ = "scenario.npz"
_NET_SCALARS = ("base_mva", "n_bus", "slack_bus", "slack_vm", "slack_va_deg", "vm_init", "name")
_NET_ARRAYS = ("br_from", "br_to", "br_r", "br_x", "br_b", "br_g", "br_tap", "br_shift", "br_status", "br_is_line",
               "bus_gs", "bus_bs", "bus_zone", "load_bus", "load_scaling", "sgen_bus", "sgen_zone", "sgen_scaling")


def net_to_json(net: NetDesc) -> str:
    """The static network (what ``_pd2ppc`` derives from ``model.p``) as a JSON document; floats round-trip exactly
    (``repr`` precision)."""
    import json
    d = {k: getattr(net, k) for k in _NET_SCALARS}
    d.update({k: np.asarray(getattr(net, k)).tolist() for k in _NET_ARRAYS})
    d["zone_names"] = list(net.zone_names)
    d["format"] = "mapdn_b200.net/1"
    return json.dumps(d)


def net_from_json(text: str) -> NetDesc:
    import json
    d = json.loads(text)
    if d.pop("format", None) != "mapdn_b200.net/1":
        raise ValueError("not a mapdn_b200 network document")
    return NetDesc(**d)


def save_scenario_npz(path: str, net: NetDesc, prof: ProfileDesc) -> str:
    """One compressed file with the network (JSON) and the three profile tables *before* ``pv_scale`` /
    ``demand_scale`` - parsing three years of 3-minute CSV rows takes tens of seconds, this loads in well under one.
    ``path`` may be a directory (-> ``<path>/scenario.npz``, picked up by :func:`load_scenario`)."""
    if os.path.isdir(path):
        path = os.path.join(path, SCENARIO_NPZ)
    np.savez_compressed(path, net_json=np.array(net_to_json(net)), pv=prof.pv, load_p=prof.load_p, load_q=prof.load_q,
                        steps_per_hour=np.int64(prof.steps_per_hour), n_days=np.int64(prof.n_days))
    return path


def load_scenario_npz(path: str, pv_scale: float = 1.0, demand_scale: float = 1.0) -> Tuple[NetDesc, ProfileDesc]:
    with np.load(path, allow_pickle=False) as z:
        net = net_from_json(str(z["net_json"]))
        prof = ProfileDesc(pv=z["pv"] * pv_scale, load_p=z["load_p"] * demand_scale, load_q=z["load_q"] * demand_scale,
                           steps_per_hour=int(z["steps_per_hour"]), n_days=int(z["n_days"]))
    return net, prof


def load_scenario(data_path: str, pv_scale: float = 1.0, demand_scale: float = 1.0) -> Tuple[NetDesc, ProfileDesc]:
    """``data_path`` as in the reference's env_args (a directory with model.p and the three CSVs). A
    ``scenario.npz`` written by :func:`save_scenario_npz` (``python -m mapdn_b200.ingest <data_path>``) takes
    precedence when it is at least as new as the four source files."""
    npz = os.path.join(data_path, SCENARIO_NPZ)
    sources = [os.path.join(data_path, f) for f in ("model.p", "pv_active.csv", "load_active.csv", "load_reactive.csv")]
    if os.path.isfile(npz) and all((not os.path.exists(f)) or os.path.getmtime(f) <= os.path.getmtime(npz) for f in sources):
        net, prof = load_scenario_npz(npz, pv_scale, demand_scale)
        if prof.pv.shape[1] != net.n_sgen or prof.load_p.shape[1] != net.n_load:
            raise ValueError("profile columns do not match the net's sgen / load tables")
        return net, prof
    net = load_network(os.path.join(data_path, "model.p"))
    prof = load_profiles(data_path, pv_scale, demand_scale)
    if prof.pv.shape[1] != net.n_sgen or prof.load_p.shape[1] != net.n_load:
        raise ValueError("profile columns do not match the net's sgen / load tables")
    return net, prof


def verify_report(data_path: str, device: int = 0) -> dict:
    """Self-check for whoever has the real scenario files (reference README.md:98-107): ingest ``data_path``, solve the
    net on the GPU at the first profile row with q = 0 (no control) and return the numbers a pandapower user can compare
    with ``pp.runpp(pp.from_pickle("model.p"))`` on the same row: sizes, zone sizes, V_min / V_max and their buses, total
    line loss, ext-grid infeed, Newton iterations, KCL residual of the solution against the ingested Ybus."""
    import torch
    from .env import BatchedVoltageControl
    net, prof = load_scenario(data_path)
    env = BatchedVoltageControl(net, None, None, batch=1, device=device)
    row = 0
    pl, ql, pv = prof.load_p[row][None], prof.load_q[row][None], prof.pv[row][None]
    out = env.solve(pl, ql, pv, np.zeros_like(pv))
    torch.cuda.synchronize()
    vm = out["vm"].cpu().numpy()[0]
    va = np.deg2rad(out["va_deg"].cpu().numpy()[0])
    V = vm * np.exp(1j * va)
    Y = env.ybus_dense()
    pd = np.zeros(net.n_bus); qd = np.zeros(net.n_bus)
    np.add.at(pd, net.load_bus, pl[0] * net.load_scaling); np.add.at(qd, net.load_bus, ql[0] * net.load_scaling)
    np.add.at(pd, net.sgen_bus, -pv[0] * net.sgen_scaling)
    mis = V * np.conj(Y @ V) + (pd + 1j * qd) / net.base_mva
    mis[net.slack_bus] = 0.0
    zones = {str(net.zone_names[z]) if z < len(net.zone_names) else int(z): int((net.bus_zone == z).sum())
             for z in sorted(set(net.bus_zone.tolist()))}
    rep = dict(data_path=os.path.abspath(data_path), n_bus=int(net.n_bus), n_branch=int(net.n_branch),
               n_line=int(net.br_is_line.sum()), n_trafo=int(net.n_branch - net.br_is_line.sum()), n_load=int(net.n_load),
               n_sgen=int(net.n_sgen), base_mva=float(net.base_mva), slack_bus=int(net.slack_bus), slack_vm=float(net.slack_vm),
               zone_sizes=zones, obs_dim=int(env.obs_size), state_dim=int(env.state_size),
               profile_rows=int(prof.n_rows), steps_per_hour=int(prof.steps_per_hour), n_days=int(prof.n_days),
               s_max_mva=[float(x) for x in prof.s_max], row=row, converged=bool(out["converged"][0]),
               newton_iterations=int(out["iterations"][0]), v_min_pu=float(vm.min()), v_min_bus=int(vm.argmin()),
               v_max_pu=float(vm.max()), v_max_bus=int(vm.argmax()), total_line_loss_mw=float(out["pl"].sum()),
               ext_grid_p_mw=float(-out["p_bus"][0, net.slack_bus]), ext_grid_q_mvar=float(-out["q_bus"][0, net.slack_bus]),
               kcl_residual_pu=float(np.abs(mis).max()), solver="meshed (dense LU)" if env.dims["n_levels"] == 1 and net.n_bus > 2
               else "radial (tree elimination)")
    env.close()
    return rep


if __name__ == "__main__":          # python -m mapdn_b200.ingest <data_path> [out.npz] | --verify <data_path>
    import json
    import sys
    if len(sys.argv) >= 3 and sys.argv[1] == "--verify":
        print(json.dumps(verify_report(sys.argv[2]), indent=1))
        sys.exit(0)
    if len(sys.argv) < 2:
        sys.exit("usage: python -m mapdn_b200.ingest <scenario directory> [out.npz]\n"
                 "       python -m mapdn_b200.ingest --verify <scenario directory>   (needs a GPU)")
    _net, _prof = load_scenario(sys.argv[1])
    _out = save_scenario_npz(sys.argv[2] if len(sys.argv) > 2 else sys.argv[1], _net, _prof)
    print(f"{_out}: {_net.n_bus} buses, {_net.n_load} loads, {_net.n_sgen} sgens, {_prof.n_rows} rows")
