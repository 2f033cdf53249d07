"""ORACLE - TEST INFRASTRUCTURE ONLY. Not part of the product; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this package.

PARITY UNPINNED: the arithmetic of the hot path lives in the third-party dependency
**pandapower 2.7.0** (pinned at reference ``environment.yml:133``; called with default
arguments at reference ``environments/var_voltage_control/voltage_control_env.py:124,165,557``).
pandapower is neither vendored under /root/reference nor installable in this image, and the
reference ships no tests / golden vectors. This file therefore *restates the published
algorithm* of pandapower 2.7.0's default ``runpp`` (PYPOWER ``makeYbus`` / ``makeSbus`` /
``newtonpf`` / ``dSbus_dV`` / ``pfsoln``) in NumPy/SciPy fp64.
What pins it instead: closed-form 2-bus solution, the published IEEE-33 (Baran & Wu) results,
and an independent backward/forward-sweep solver (``oracle/independent.py``).

Input: any object with the attributes of ``mapdn_b200.network.NetDesc`` (duck-typed; this
package never imports the product).
"""
from __future__ import annotations

import numpy as np
from scipy.sparse import csr_matrix, hstack, vstack
from scipy.sparse.linalg import spsolve

TOL_DEFAULT = 1e-8      # pandapower runpp tolerance_mva default, applied to ||F||inf in p.u.
MAX_IT_DEFAULT = 10     # runpp max_iteration="auto" -> 10 for algorithm="nr"


def make_ybus(net):
    """PYPOWER ``makeYbus`` (pandapower/pypower/makeYbus.py), SURVEY Appendix A.3.

    Returns (Ybus csr [nb,nb], Yf csr [nbr,nb], Yt csr [nbr,nb])."""
    nb, nbr = net.n_bus, net.br_from.shape[0]
    stat = net.br_status.astype(np.float64)
    # pandapower drops out-of-service branches in _ppc2ppci before makeYbus; here they stay as all-zero rows (so that Yf / Yt
    # keep the branch numbering) - also when r = x = 0
    z = net.br_r + 1j * net.br_x
    Ys = np.where(stat > 0, stat / np.where(stat > 0, z, 1.0), 0.0)
    Bc = stat * (net.br_b - 1j * net.br_g)           # complex "b": b - j g
    tap = np.where(net.br_tap == 0.0, 1.0, net.br_tap).astype(np.complex128)
    tap = tap * np.exp(1j * np.pi / 180.0 * net.br_shift)
    Ytt = Ys + 1j * Bc / 2.0
    Yff = Ytt / (tap * np.conj(tap))
    Yft = -Ys / np.conj(tap)
    Ytf = -Ys / tap
    Ysh = (net.bus_gs + 1j * net.bus_bs) / net.base_mva
    f, t = net.br_from.astype(np.int64), net.br_to.astype(np.int64)
    i = np.arange(nbr)
    Yf = csr_matrix((np.r_[Yff, Yft], (np.r_[i, i], np.r_[f, t])), (nbr, nb))
    Yt = csr_matrix((np.r_[Ytf, Ytt], (np.r_[i, i], np.r_[f, t])), (nbr, nb))
    Cf = csr_matrix((np.ones(nbr), (i, f)), (nbr, nb))
    Ct = csr_matrix((np.ones(nbr), (i, t)), (nbr, nb))
    Ybus = Cf.T @ Yf + Ct.T @ Yt + csr_matrix((Ysh, (np.arange(nb), np.arange(nb))), (nb, nb))
    return Ybus.tocsr(), Yf, Yt


def bus_demand(net, p_load, q_load, p_sgen, q_sgen):
    """pandapower ``_calc_pq_elements_and_add_on_ppc``: PD/QD per bus in MW/MVAr,
    load convention: ``PD_b = sum(load.p*scaling) - sum(sgen.p*scaling)``."""
    PD = np.zeros(net.n_bus)
    QD = np.zeros(net.n_bus)
    np.add.at(PD, net.load_bus, np.asarray(p_load) * net.load_scaling)
    np.add.at(QD, net.load_bus, np.asarray(q_load) * net.load_scaling)
    np.add.at(PD, net.sgen_bus, -np.asarray(p_sgen) * net.sgen_scaling)
    np.add.at(QD, net.sgen_bus, -np.asarray(q_sgen) * net.sgen_scaling)
    return PD, QD


def dSbus_dV(Ybus, V):
    """PYPOWER ``dSbus_dV`` sparse branch (pandapower/pypower/dSbus_dV.py)."""
    nb = V.shape[0]
    ib = np.arange(nb)
    Ibus = Ybus @ V
    diagV = csr_matrix((V, (ib, ib)))
    diagIbus = csr_matrix((Ibus, (ib, ib)))
    diagVnorm = csr_matrix((V / np.abs(V), (ib, ib)))
    dS_dVm = diagV @ np.conj(Ybus @ diagVnorm) + np.conj(diagIbus) @ diagVnorm
    dS_dVa = 1j * diagV @ np.conj(diagIbus - Ybus @ diagV)
    return dS_dVm.tocsr(), dS_dVa.tocsr()


def _evaluate_Fx(Ybus, V, Sbus, pv, pq):
    mis = V * np.conj(Ybus @ V) - Sbus
    return np.r_[mis[pv].real, mis[pq].real, mis[pq].imag]


def newtonpf(Ybus, Sbus, V0, pv, pq, tol=TOL_DEFAULT, max_it=MAX_IT_DEFAULT):
    """pandapower/pypower/newtonpf.py (2.7.0), SURVEY Appendix A.4.

    Polar Newton-Raphson; convergence is tested *before* the first solve; at most ``max_it``
    solves. Returns (V, converged, iterations)."""
    pv = np.asarray(pv, np.int64)
    pq = np.asarray(pq, np.int64)
    pvpq = np.r_[pv, pq]
    npv, npq = len(pv), len(pq)
    V = V0.astype(np.complex128).copy()
    Va, Vm = np.angle(V), np.abs(V)
    i = 0
    F = _evaluate_Fx(Ybus, V, Sbus, pv, pq)
    converged = np.linalg.norm(F, np.inf) < tol
    while (not converged) and i < max_it:
        i += 1
        dS_dVm, dS_dVa = dSbus_dV(Ybus, V)
        J11 = dS_dVa[pvpq, :][:, pvpq].real
        J12 = dS_dVm[pvpq, :][:, pq].real
        J21 = dS_dVa[pq, :][:, pvpq].imag
        J22 = dS_dVm[pq, :][:, pq].imag
        J = vstack([hstack([J11, J12]), hstack([J21, J22])], format="csr")
        dx = -1 * spsolve(J, F)
        if npv:
            Va[pv] = Va[pv] + dx[:npv]
        if npq:
            Va[pq] = Va[pq] + dx[npv:npv + npq]
            Vm[pq] = Vm[pq] + dx[npv + npq:]
        V = Vm * np.exp(1j * Va)
        Vm, Va = np.abs(V), np.angle(V)          # re-wrap, as newtonpf does
        F = _evaluate_Fx(Ybus, V, Sbus, pv, pq)
        converged = np.linalg.norm(F, np.inf) < tol
    return V, bool(converged), i


class PowerFlowResult:
    __slots__ = ("converged", "iterations", "vm_pu", "va_degree", "p_mw", "q_mvar",
                 "pl_mw", "V", "p_ext_mw", "q_ext_mvar")


class PandapowerEquivalent:
    """``pp.runpp(net)`` with default arguments for a MAPDN-style net (one ext_grid slack, no
    ``gen`` elements, PVs as ``sgen``): flat start every call, tol 1e-8 p.u., <= 10 iterations.

    Ybus is cached because the topology never changes (the reference rebuilds it each call);
    this does not alter any result."""

    def __init__(self, net, tol=TOL_DEFAULT, max_it=MAX_IT_DEFAULT):
        self.net = net
        self.tol, self.max_it = tol, max_it
        self.Ybus, self.Yf, self.Yt = make_ybus(net)
        self.ref = int(net.slack_bus)
        self.pq = np.array([b for b in range(net.n_bus) if b != self.ref], np.int64)
        self.pv = np.zeros(0, np.int64)
        self.lines = np.nonzero(net.br_is_line)[0]

    def flat_start(self):
        n = self.net
        V0 = np.full(n.n_bus, n.vm_init, np.complex128)       # init="auto": mean slack vm, angle 0
        V0[self.ref] = n.slack_vm * np.exp(1j * np.pi / 180.0 * n.slack_va_deg)
        return V0

    def runpp(self, p_load, q_load, p_sgen, q_sgen) -> PowerFlowResult:
        n = self.net
        PD, QD = bus_demand(n, p_load, q_load, p_sgen, q_sgen)
        Sbus = -(PD + 1j * QD) / n.base_mva                    # makeSbus; slack gen irrelevant to F
        V, conv, it = newtonpf(self.Ybus, Sbus, self.flat_start(), self.pv, self.pq,
                               self.tol, self.max_it)
        res = PowerFlowResult()
        res.converged, res.iterations, res.V = conv, it, V
        res.vm_pu = np.abs(V)
        res.va_degree = np.angle(V) * 180.0 / np.pi
        # pfsoln + results_bus: res_bus.p_mw = demand at the bus (load convention);
        # at the slack bus the ext_grid's infeed is subtracted.
        Sinj = V * np.conj(self.Ybus @ V) * n.base_mva         # injection into the network, MVA
        p, q = PD.copy(), QD.copy()
        p_ext = Sinj[self.ref].real + PD[self.ref]
        q_ext = Sinj[self.ref].imag + QD[self.ref]
        p[self.ref] = PD[self.ref] - p_ext
        q[self.ref] = QD[self.ref] - q_ext
        # bus shunts appear in res_bus as demand at the solved voltage (pandapower results_bus._get_shunt_results:
        # p += p_shunt * vm^2, q += q_shunt * vm^2 with ppc GS = p_shunt, BS = -q_shunt) [RECALLED]
        p = p + n.bus_gs * res.vm_pu ** 2
        q = q - n.bus_bs * res.vm_pu ** 2
        res.p_mw, res.q_mvar = p, q
        res.p_ext_mw, res.q_ext_mvar = p_ext, q_ext
        Sf = V[n.br_from] * np.conj(self.Yf @ V) * n.base_mva
        St = V[n.br_to] * np.conj(self.Yt @ V) * n.base_mva
        res.pl_mw = (Sf + St).real[self.lines]                 # res_line.pl_mw (lines only)
        return res
