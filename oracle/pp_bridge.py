"""ORACLE - TEST INFRASTRUCTURE ONLY. Bridge to the REAL pandapower, for the day it is importable.

pandapower (pinned 2.7.0 by the reference, ``environment.yml:133``) is not installed in this image, not in the
offline wheelhouse and not vendored, so nothing in here has been executed yet: the pandapower API calls below are
written from the 2.7.0 documentation [RECALLED] and are exercised only by ``scripts/pin_with_pandapower.py`` and by
``bench.py``'s CPU arm when ``import pandapower`` succeeds (kind = "pandapower"). They turn "parity unpinned" into a
one-command check:

    python scripts/pin_with_pandapower.py        # writes tests/golden/pp_*.npz from pp.runpp
    python -m pytest tests -q                    # test_pandapower_pins.py then holds oracle AND CUDA path to them

``build_pp_net`` maps a per-unit :class:`mapdn_b200.network.NetDesc` back onto pandapower element tables so that
pandapower's own ``_pd2ppc`` reproduces exactly the branch table the NetDesc holds:

* lines: ``vn_kv`` of every bus = ``vn_kv`` (any value - the per-unit conversion cancels), ``length_km = 1``,
  ``r_ohm_per_km = r * baseR``, ``x`` likewise, ``c_nf_per_km = b / (2 pi f 1e-9 baseR)``,
  ``g_us_per_km = g / (1e-6 baseR)``, ``baseR = vn_kv^2 / sn_mva``;
* branches with an off-nominal ratio (``br_tap != 1``) or ``br_is_line = 0``: a two-winding transformer without
  magnetising branch (T and pi models coincide): ``sn_mva = net.sn_mva``, ``vn_lv_kv = vn_kv``,
  ``vn_hv_kv = tap * vn_kv`` (hv = from bus), ``vk_percent = 100 |z|``, ``vkr_percent = 100 r``,
  ``pfe_kw = i0_percent = 0``; line charging on such a branch is not representable -> ValueError;
* bus shunts: ``create_shunt(p_mw = GS, q_mvar = -BS)``; loads / sgens with their ``scaling``; one ``ext_grid``.
"""
from __future__ import annotations

import numpy as np

F_HZ = 50.0


def build_pp_net(nd, vn_kv: float = 12.66):
    import pandapower as pp

    net = pp.create_empty_network(sn_mva=float(nd.base_mva), f_hz=F_HZ)
    for b in range(nd.n_bus):
        pp.create_bus(net, vn_kv=vn_kv, index=b, zone=int(nd.bus_zone[b]) if nd.bus_zone is not None else None)
    base_r = vn_kv ** 2 / nd.base_mva
    for k in range(nd.n_branch):
        f, t = int(nd.br_from[k]), int(nd.br_to[k])
        r, x, b, g = float(nd.br_r[k]), float(nd.br_x[k]), float(nd.br_b[k]), float(nd.br_g[k])
        tap = float(nd.br_tap[k]) or 1.0
        in_service = bool(nd.br_status[k])
        if tap == 1.0 and nd.br_is_line[k] and float(nd.br_shift[k]) == 0.0:
            pp.create_line_from_parameters(
                net, f, t, length_km=1.0, r_ohm_per_km=r * base_r, x_ohm_per_km=x * base_r,
                c_nf_per_km=b / (2.0 * np.pi * F_HZ * 1e-9 * base_r), g_us_per_km=g / (1e-6 * base_r),
                max_i_ka=1e3, in_service=in_service)
        else:
            if b != 0.0 or g != 0.0:
                raise ValueError(f"branch {k}: charging on a tap-changing branch has no pandapower transformer equivalent")
            z = abs(complex(r, x))
            pp.create_transformer_from_parameters(
                net, hv_bus=f, lv_bus=t, sn_mva=float(nd.base_mva), vn_hv_kv=tap * vn_kv, vn_lv_kv=vn_kv,
                vkr_percent=100.0 * r, vk_percent=100.0 * z, pfe_kw=0.0, i0_percent=0.0,
                shift_degree=float(nd.br_shift[k]), in_service=in_service)
    for b in range(nd.n_bus):
        if nd.bus_gs[b] != 0.0 or nd.bus_bs[b] != 0.0:
            pp.create_shunt(net, b, q_mvar=-float(nd.bus_bs[b]), p_mw=float(nd.bus_gs[b]))
    for l in range(nd.n_load):
        pp.create_load(net, int(nd.load_bus[l]), p_mw=0.0, q_mvar=0.0, scaling=float(nd.load_scaling[l]))
    for j in range(nd.n_sgen):
        pp.create_sgen(net, int(nd.sgen_bus[j]), p_mw=0.0, q_mvar=0.0, scaling=float(nd.sgen_scaling[j]),
                       name=str(nd.zone_names[int(nd.sgen_zone[j])]) if getattr(nd, "zone_names", None) else None)
    pp.create_ext_grid(net, int(nd.slack_bus), vm_pu=float(nd.slack_vm), va_degree=float(nd.slack_va_deg))
    return net


class PandapowerBackend:
    """Same ``runpp(p_load, q_load, p_sgen, q_sgen) -> PowerFlowResult`` as ``PandapowerEquivalent``, but the numbers come
    from ``pandapower.runpp(net)`` with default arguments - the call the reference makes
    (``voltage_control_env.py:124,165,557``)."""

    def __init__(self, nd, vn_kv: float = 12.66):
        import pandapower as pp
        self.pp, self.nd = pp, nd
        self.net = build_pp_net(nd, vn_kv)
        # res_line rows of the NetDesc's "line" branches, in branch order
        self.n_line = len(self.net.line)

    def runpp(self, p_load, q_load, p_sgen, q_sgen):
        from .pandapower_nr import PowerFlowResult
        net, pp = self.net, self.pp
        net.load["p_mw"] = np.asarray(p_load, float)
        net.load["q_mvar"] = np.asarray(q_load, float)
        net.sgen["p_mw"] = np.asarray(p_sgen, float)
        net.sgen["q_mvar"] = np.asarray(q_sgen, float)
        res = PowerFlowResult()
        try:
            pp.runpp(net)
            res.converged = True
        except pp.powerflow.LoadflowNotConverged:
            res.converged = False
        res.iterations = int(net._ppc.get("iterations", -1)) if getattr(net, "_ppc", None) is not None else -1
        rb = net.res_bus.sort_index()
        res.vm_pu = rb.vm_pu.values.copy()
        res.va_degree = rb.va_degree.values.copy()
        res.p_mw, res.q_mvar = rb.p_mw.values.copy(), rb.q_mvar.values.copy()
        res.pl_mw = net.res_line.sort_index().pl_mw.values.copy()
        res.V = res.vm_pu * np.exp(1j * np.deg2rad(res.va_degree))
        res.p_ext_mw = float(net.res_ext_grid.p_mw.values[0])
        res.q_ext_mvar = float(net.res_ext_grid.q_mvar.values[0])
        return res
