"""ORACLE - TEST INFRASTRUCTURE ONLY. ctypes wrapper of the plain-C Newton-Raphson restatement
(``oracle/c/nr_dense.c``), compiled on demand with gcc into ``oracle/c/_build/`` (git-ignored)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "c", "nr_dense.c")
LIB = os.path.join(_HERE, "c", "_build", "libnr_dense.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"])
    return LIB


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        pd = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
        L.mapdn_oracle_nr_dense.argtypes = [C.c_int, C.c_int, pd, pd, pd, pd, C.c_double, C.c_double, C.c_double,
                                            C.c_double, C.c_int, pd, pd, C.POINTER(C.c_int)]
        L.mapdn_oracle_nr_dense.restype = C.c_int
        _lib = L
    return _lib


class COracle:
    """``runpp`` through the C restatement: (vm_pu, va_degree, converged, iterations)."""

    def __init__(self, net, tol=1e-8, max_it=10):
        from .pandapower_nr import make_ybus
        self.net, self.tol, self.max_it = net, tol, max_it
        Y = make_ybus(net)[0].toarray()
        self.G, self.B = np.ascontiguousarray(Y.real), np.ascontiguousarray(Y.imag)

    def runpp(self, p_load, q_load, p_sgen, q_sgen):
        from .pandapower_nr import bus_demand
        n = self.net
        PD, QD = bus_demand(n, p_load, q_load, p_sgen, q_sgen)
        P, Q = np.ascontiguousarray(-PD / n.base_mva), np.ascontiguousarray(-QD / n.base_mva)
        vm, va = np.zeros(n.n_bus), np.zeros(n.n_bus)
        conv = C.c_int(0)
        it = lib().mapdn_oracle_nr_dense(n.n_bus, n.slack_bus, self.G, self.B, P, Q, n.vm_init, n.slack_vm,
                                         np.deg2rad(n.slack_va_deg), self.tol, self.max_it, vm, va, C.byref(conv))
        return vm, np.rad2deg(va), bool(conv.value), it
