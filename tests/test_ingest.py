"""Reference file formats (model.p, pv_active.csv, load_*.csv) -> NetDesc / ProfileDesc, checked against
hand-computed per-unit values (PARITY UNPINNED: no pandapower / sample model.p in this image)."""
import os
import pickle

import numpy as np
import pandas as pd
import pytest

from mapdn_b200 import ingest
from oracle.pandapower_nr import PandapowerEquivalent, make_ybus


def tables(with_trafo=True, magnetising=False):
    bus = pd.DataFrame(dict(name=["hv", "b0", "b1", "b2"], vn_kv=[110.0, 20.0, 20.0, 20.0],
                            zone=["main", "main", "zone1", "zone2"], in_service=True), index=[10, 0, 1, 2])
    line = pd.DataFrame(dict(from_bus=[0, 1], to_bus=[1, 2], length_km=[2.0, 1.5], r_ohm_per_km=[0.3, 0.4],
                             x_ohm_per_km=[0.35, 0.3], c_nf_per_km=[10.0, 0.0], parallel=[1, 2], in_service=True))
    trafo = pd.DataFrame(dict(hv_bus=[10], lv_bus=[0], sn_mva=[25.0], vn_hv_kv=[110.0], vn_lv_kv=[20.0],
                              vk_percent=[12.0], vkr_percent=[0.41], pfe_kw=[14.0 if magnetising else 0.0],
                              i0_percent=[0.07 if magnetising else 0.0], shift_degree=[150.0], tap_side=["hv"],
                              tap_neutral=[0], tap_pos=[2], tap_step_percent=[1.5], parallel=[1], in_service=True))
    load = pd.DataFrame(dict(bus=[1, 2, 2], p_mw=[1.0, 0.5, 0.2], q_mvar=[0.3, 0.1, 0.0], scaling=[1.0, 0.9, 1.0],
                             in_service=[True, True, False], const_z_percent=0.0, const_i_percent=0.0))
    sgen = pd.DataFrame(dict(bus=[1, 2], p_mw=[0.4, 0.6], q_mvar=0.0, scaling=1.0, in_service=True, name=["zone1", "zone2"]))
    eg = pd.DataFrame(dict(bus=[10 if with_trafo else 0], vm_pu=[1.02], va_degree=[0.0], in_service=True))
    t = dict(bus=bus, line=line, load=load, sgen=sgen, ext_grid=eg, sn_mva=10.0, f_hz=50.0)
    if with_trafo:
        t["trafo"] = trafo
    else:
        t["bus"] = bus.drop(index=10)
    return t


def test_line_and_element_conversion():
    net = ingest.net_from_tables(tables(with_trafo=False))
    assert net.n_bus == 3 and net.slack_bus == 0 and net.slack_vm == 1.02 and net.base_mva == 10.0
    zb = 20.0 ** 2 / 10.0
    assert np.allclose(net.br_r, [0.3 * 2.0 / zb, 0.4 * 1.5 / 2 / zb])
    assert np.allclose(net.br_x, [0.35 * 2.0 / zb, 0.3 * 1.5 / 2 / zb])
    assert np.allclose(net.br_b, [2 * np.pi * 50 * 10e-9 * 2.0 * zb, 0.0])
    assert list(net.br_is_line) == [1, 1] and list(net.br_tap) == [1.0, 1.0]
    assert list(net.load_bus) == [1, 2, 2] and np.allclose(net.load_scaling, [1.0, 0.9, 0.0])
    assert list(net.sgen_zone) == [1, 2] and list(net.bus_zone) == [0, 1, 2]
    assert net.obs_dim == 4 * 1 + 2


def test_trafo_series_and_tap():
    net = ingest.net_from_tables(tables(with_trafo=True))
    # bus order by pandapower index: 0,1,2,10 -> trafo hv bus is internal 3
    assert net.slack_bus == 3 and net.n_bus == 4
    k = 2
    assert (net.br_from[k], net.br_to[k]) == (3, 0) and net.br_is_line[k] == 0
    vn_hv = 110.0 * (1 + 2 * 1.5 / 100)
    assert np.isclose(net.br_tap[k], (vn_hv / 20.0) / (110.0 / 20.0))
    tap_lv = (20.0 / 20.0) ** 2 * 10.0
    z = 12.0 / 100 / 25.0 * tap_lv
    r = 0.41 / 100 / 25.0 * tap_lv
    assert np.isclose(net.br_r[k], r) and np.isclose(net.br_x[k], np.sqrt(z * z - r * r))
    assert net.br_b[k] == 0 and net.br_g[k] == 0 and net.br_shift[k] == 0.0      # phase shift ignored below 70 kV
    res = PandapowerEquivalent(net).runpp([1.0, 0.5, 0.2], [0.3, 0.1, 0.0], [0.4, 0.6], [0.0, 0.0])
    assert res.converged and 0.9 < res.vm_pu.min() < 1.03


def test_trafo_t_to_pi_conversion():
    """The pi equivalent must reproduce the T circuit's two-port exactly."""
    net = ingest.net_from_tables(tables(with_trafo=True, magnetising=True))
    k = 2
    tap_lv = 10.0
    z = 12.0 / 100 / 25.0 * tap_lv; r = 0.41 / 100 / 25.0 * tap_lv; x = np.sqrt(z * z - r * r)
    base_r = 20.0 ** 2 / 10.0
    pfe = 14e-3
    ymag = (pfe / 400.0 * base_r) + 1j * (-np.sqrt((0.07 / 100 * 25.0) ** 2 - pfe ** 2) * base_r / 400.0)   # g + jb
    za = (r + 1j * x) / 2
    # T circuit two-port admittance (no tap): Y11 = 1/(za + 1/(ymag + 1/za))
    y11_t = 1 / (za + 1 / (ymag + 1 / za))
    ys = 1 / (net.br_r[k] + 1j * net.br_x[k])
    y11_pi = ys + (net.br_g[k] + 1j * net.br_b[k]) / 2
    assert np.isclose(y11_pi, y11_t, rtol=1e-12)
    y12_t = -(1 / za) * (1 / za) / (2 / za + ymag)             # transfer admittance of the T
    assert np.isclose(-ys, y12_t, rtol=1e-12)


def test_model_pickle_layouts_and_csv(tmp_path):
    t = tables(with_trafo=False)
    # layout 1: pp.to_pickle style ({"DF": split-dict, "dtypes": ...})
    d1 = {k: ({"DF": v.to_dict("split"), "dtypes": {c: str(dt) for c, dt in zip(v.columns, v.dtypes)}}
              if isinstance(v, pd.DataFrame) else v) for k, v in t.items()}
    p1 = tmp_path / "s1"; p1.mkdir()
    pickle.dump(d1, open(p1 / "model.p", "wb"), protocol=2)
    # layout 2: plain pickle of a dict-like net object holding DataFrames
    p2 = tmp_path / "s2"; p2.mkdir()
    pickle.dump(t, open(p2 / "model.p", "wb"))
    ts = pd.date_range("2012-01-01", periods=3 * 480 + 1, freq="3min")
    rng = np.random.default_rng(0)
    for p in (p1, p2):
        for fn, ncol in (("pv_active.csv", 2), ("load_active.csv", 3), ("load_reactive.csv", 3)):
            df = pd.DataFrame(rng.uniform(0, 1, (len(ts), ncol)))
            df.insert(0, "time", ts)
            df.to_csv(p / fn, index=False)
    nets = []
    for p in (p1, p2):
        net, prof = ingest.load_scenario(str(p), pv_scale=2.0, demand_scale=0.5)
        nets.append(net)
        assert prof.steps_per_hour == 20 and prof.n_days == 3 and prof.n_rows == 3 * 480 + 1
        assert prof.pv.shape[1] == net.n_sgen == 2 and prof.load_p.shape[1] == net.n_load == 3
    assert np.allclose(make_ybus(nets[0])[0].toarray(), make_ybus(nets[1])[0].toarray())
    raw = pd.read_csv(p1 / "pv_active.csv").iloc[:, 1:].to_numpy()
    assert np.allclose(ingest.load_profiles(str(p1), pv_scale=2.0).pv, 2.0 * raw)
    # one-off conversion (python -m mapdn_b200.ingest <dir>): load_scenario then reads scenario.npz, as long as it is
    # not older than the source files
    import subprocess, sys, time
    from conftest import ROOT
    net0, prof0 = ingest.load_scenario(str(p1))
    r = subprocess.run([sys.executable, "-m", "mapdn_b200.ingest", str(p1)], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0 and "scenario.npz" in r.stdout, r.stderr
    os.rename(p1 / "model.p", p1 / "model.p.off")                       # proves the NPZ is what gets read
    net1, prof1 = ingest.load_scenario(str(p1))
    assert np.array_equal(prof1.pv, prof0.pv) and np.array_equal(net1.br_r, net0.br_r)
    os.rename(p1 / "model.p.off", p1 / "model.p")
    later = time.time() + 10
    os.utime(p1 / "pv_active.csv", (later, later))                      # a newer source invalidates the NPZ
    df = pd.read_csv(p1 / "pv_active.csv"); df.iloc[:, 1] *= 3.0; df.to_csv(p1 / "pv_active.csv", index=False)
    os.utime(p1 / "pv_active.csv", (later, later))
    _, prof2 = ingest.load_scenario(str(p1))
    assert np.allclose(prof2.pv[:, 0], 3.0 * prof0.pv[:, 0])


def test_unsupported_content_is_refused():
    t = tables(with_trafo=False)
    t["gen"] = pd.DataFrame(dict(bus=[1], p_mw=[1.0], vm_pu=[1.0], in_service=True))
    with pytest.raises(NotImplementedError):
        ingest.net_from_tables(t)
    t = tables(with_trafo=False)
    t["load"].loc[0, "const_z_percent"] = 30.0
    with pytest.raises(NotImplementedError):
        ingest.net_from_tables(t)


def test_export_format_round_trips_exactly(tmp_path):
    """JSON network document + one-file NPZ scenario; bit-exact round trip, picked up by load_scenario."""
    from mapdn_b200 import cases, ingest
    net, prof = cases.make_case("case141"), cases.make_profiles("case141", n_days=3)
    back = ingest.net_from_json(ingest.net_to_json(net))
    for k in ingest._NET_ARRAYS:
        a, b = getattr(net, k), getattr(back, k)
        assert a.dtype == b.dtype and np.array_equal(a, b), k
    for k in ingest._NET_SCALARS:
        assert getattr(net, k) == getattr(back, k), k
    assert back.zone_names == net.zone_names and back.obs_dim == net.obs_dim
    with pytest.raises(ValueError):
        ingest.net_from_json('{"format": "something else"}')

    out = ingest.save_scenario_npz(str(tmp_path), net, prof)
    assert out.endswith("scenario.npz")
    n2, p2 = ingest.load_scenario(str(tmp_path), pv_scale=0.5, demand_scale=2.0)     # no model.p / CSVs needed
    assert np.array_equal(n2.br_x, net.br_x) and np.array_equal(n2.sgen_zone, net.sgen_zone)
    assert np.array_equal(p2.pv, prof.pv * 0.5) and np.array_equal(p2.load_q, prof.load_q * 2.0)
    assert p2.steps_per_hour == prof.steps_per_hour and p2.n_days == prof.n_days
    assert np.array_equal(p2.s_max, 1.2 * (prof.pv * 0.5).max(axis=0))
