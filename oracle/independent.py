"""ORACLE - TEST INFRASTRUCTURE ONLY. Independent solutions used to pin ``pandapower_nr``
(SURVEY §8c items 1-3): a closed-form 2-bus case and a backward/forward-sweep solver for radial
feeders. Neither shares code or formulation with the Newton-Raphson restatement."""
import numpy as np


def two_bus_closed_form(v0, r, x, p, q):
    """Receiving-end voltage magnitude of slack(v0) --(r+jx)-- PQ bus drawing p+jq (p.u.).

    |V|^4 + (2(pr+qx) - v0^2)|V|^2 + (p^2+q^2)(r^2+x^2) = 0, upper root."""
    b = 2.0 * (p * r + q * x) - v0 * v0
    c = (p * p + q * q) * (r * r + x * x)
    v2 = 0.5 * (-b + np.sqrt(b * b - 4.0 * c))
    return np.sqrt(v2)


def backward_forward_sweep(net, PD, QD, tol=1e-14, max_it=200):
    """Current-summation sweep on a radial net without shunts/taps. PD/QD: bus demand MW/MVAr.
    Returns complex V [n_bus]."""
    n = net.n_bus
    assert np.all(net.br_b == 0) and np.all(net.br_g == 0)
    assert np.all((net.br_tap == 1.0) | (net.br_tap == 0.0))
    adj = [[] for _ in range(n)]
    z = net.br_r + 1j * net.br_x
    for k, (f, t) in enumerate(zip(net.br_from, net.br_to)):
        if net.br_status[k]:
            adj[f].append((t, k))
            adj[t].append((f, k))
    parent = np.full(n, -1)
    pbr = np.full(n, -1)
    order = [int(net.slack_bus)]
    seen = {int(net.slack_bus)}
    for u in order:
        for v, k in adj[u]:
            if v not in seen:
                seen.add(v)
                parent[v], pbr[v] = u, k
                order.append(v)
    assert len(order) == n, "network is not connected"
    S = (PD + 1j * QD) / net.base_mva
    V = np.full(n, net.slack_vm, np.complex128)
    for _ in range(max_it):
        I = np.conj(S / V)
        I[net.slack_bus] = 0.0
        J = I.copy()
        for u in reversed(order[1:]):
            J[parent[u]] += J[u]
        Vn = V.copy()
        for u in order[1:]:
            Vn[u] = Vn[parent[u]] - z[pbr[u]] * J[u]
        d = np.max(np.abs(Vn - V))
        V = Vn
        if d < tol:
            break
    return V
