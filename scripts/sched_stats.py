"""Depth / width statistics of the elimination forest the step kernel sweeps (planning tool, CPU only):
PQ forest re-rooted at its centre, levels by height (leaves first), steps = sum over levels of ceil(width / G)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from mapdn_b200 import cases


def forest_stats(net, G):
    n, slack = net.n_bus, net.slack_bus
    adj = [[] for _ in range(n)]
    seen_pairs = set()
    for f, t, st in zip(net.br_from, net.br_to, net.br_status):
        if st and (min(f, t), max(f, t)) not in seen_pairs and f != t:
            seen_pairs.add((min(f, t), max(f, t)))
            adj[f].append(t); adj[t].append(f)
    pq = [b for b in range(n) if b != slack]
    comp, stats = {}, []
    for s in pq:
        if s in comp:
            continue
        nodes, stack = [], [s]
        comp[s] = s
        while stack:
            u = stack.pop(); nodes.append(u)
            for v in adj[u]:
                if v != slack and v not in comp:
                    comp[v] = s; stack.append(v)

        def far(src):
            dist = {src: 0}; order = [src]
            for u in order:
                for v in adj[u]:
                    if v != slack and v not in dist:
                        dist[v] = dist[u] + 1; order.append(v)
            return order[-1], dist
        a, _ = far(nodes[0]); b, da = far(a); _, db = far(b)
        diam = da[b]
        centre = min(nodes, key=lambda u: max(da[u], db[u]))
        _, dc = far(centre)
        height = {u: 0 for u in nodes}
        for u in sorted(nodes, key=lambda u: -dc[u]):          # leaves first
            for v in adj[u]:
                if v != slack and dc.get(v, -1) == dc[u] - 1:
                    height[v] = max(height[v], height[u] + 1)
        stats.append((len(nodes), diam, max(dc.values()), height))
    levels = {}
    for _, _, _, height in stats:
        for u, hgt in height.items():
            levels[hgt] = levels.get(hgt, 0) + 1
    widths = [levels[k] for k in sorted(levels)]
    steps = sum(-(-w // G) for w in widths)
    return dict(trees=len(stats), diameter=max(s[1] for s in stats), depth=max(s[2] for s in stats) + 1,
                widths=widths, steps=steps, lane_use=sum(widths) / (steps * G))


if __name__ == "__main__":
    for name, G in (("case33", 8), ("case141", 32), ("case322", 64)):
        st = forest_stats(cases.make_case(name), G)
        print(f"{name} G={G}: {st['trees']} tree(s), diameter {st['diameter']} edges, {st['depth']} levels, "
              f"{st['steps']} sweep steps, lane use {st['lane_use']:.0%}, widths {st['widths']}")
