"""TEST INFRASTRUCTURE — numpy/pure-Python restatement of the elimination orderings that
`kb2_symbolic` (kinetica.jl_b200/csrc/kb2_symbolic.cpp, `banded_order`) offers for networks with
locality: hub species last, the others by reverse Cuthill-McKee or by Sloan's profile reduction on
the graph without the hubs.  Only tests/ may import this module; the product computes its orderings
in C++.  The reference (Kinetica.jl) leaves the ordering to the linear solver the user picks
(docs/src/getting-started.md:66-70, KLU), so there is no reference ordering to match: what the
tests pin is that the C++ implements the published algorithms (the Cuthill-McKee numbering is also
compared with scipy's) and their tie-breaking rules, stated here:
  * hub species: symmetrised degree > max(32, 8 * median degree); appended last by ascending
    (degree, index);
  * neighbours are visited, roots are tried and ties are broken by ascending (degree in the hub-less
    graph, index);
  * start vertex of a component: George-Liu pseudo-peripheral vertex (repeat a breadth-first search
    from the smallest-(degree, index) vertex of the last level while the eccentricity grows);
  * Sloan: priority = W2 * distance to the end vertex - W1 * (degree + 1); numbering a preactive
    vertex raises its neighbours by W1; a neighbour that turns active is raised by W1 and raises
    its own neighbours by W1; highest priority first, ties to the smaller index.
"""
import heapq

import numpy as np


def hubless_graph(S, colptr, rowval):
    """Symmetrised pattern without self loops -> (adjacency lists without hubs sorted by (degree, index),
    hub-less degrees, hub mask, hubs in their final order)."""
    nb = [set() for _ in range(S)]
    for l in range(S):
        for p in range(colptr[l], colptr[l + 1]):
            i = int(rowval[p])
            if i != l:
                nb[i].add(l); nb[l].add(i)
    deg = np.array([len(a) for a in nb])
    thr = max(32, 8 * int(np.sort(deg)[S // 2]))
    hub = deg > thr
    hubs = sorted((int(deg[v]), v) for v in range(S) if hub[v])
    adj = [[w for w in nb[v] if not hub[w]] if not hub[v] else [] for v in range(S)]
    sdeg = np.array([len(a) for a in adj])
    adj = [sorted(a, key=lambda x: (sdeg[x], x)) for a in adj]
    return adj, sdeg, hub, [v for _, v in hubs]


def _bfs(adj, start, done):
    lev = {start: 0}
    order = [start]
    for v in order:
        for w in adj[v]:
            if w not in lev and not done[w]:
                lev[w] = lev[v] + 1
                order.append(w)
    return order, lev


def _pseudo_peripheral(adj, sdeg, start, done):
    order, lev = _bfs(adj, start, done)
    ecc = lev[order[-1]]
    while True:
        cand = min((v for v in order if lev[v] == ecc), key=lambda x: (sdeg[x], x))
        o2, l2 = _bfs(adj, cand, done)
        if l2[o2[-1]] > ecc:
            start, order, lev, ecc = cand, o2, l2, l2[o2[-1]]
        else:
            return start, order, lev


def banded_order(S, colptr, rowval, kind, weights=(1, 2)):
    """kind: 'natural', 'rcm' or 'sloan' (weights = (W1, W2)) -> perm (perm[a] = species at pivot position a)."""
    adj, sdeg, hub, hubs = hubless_graph(S, colptr, rowval)
    if kind == "natural":
        return np.array([v for v in range(S) if not hub[v]] + hubs)
    done = hub.copy()
    seq = []
    W1, W2 = weights
    for v0 in sorted((v for v in range(S) if not hub[v]), key=lambda x: (sdeg[x], x)):
        if done[v0]:
            continue
        s, order, lev = _pseudo_peripheral(adj, sdeg, v0, done)
        if kind == "rcm":
            for v in order:
                done[v] = True
            seq += order
            continue
        ecc = lev[order[-1]]
        e = min((v for v in order if lev[v] == ecc), key=lambda x: (sdeg[x], x))
        _, dist = _bfs(adj, e, done)
        status = {v: 0 for v in order}          # 0 inactive, 1 preactive, 2 active, 3 numbered
        prio = {v: W2 * dist[v] - W1 * (int(sdeg[v]) + 1) for v in order}
        status[s] = 1
        heap = [(-prio[s], s)]
        while heap:
            p, v = heapq.heappop(heap)
            if status[v] == 3 or -p != prio[v]:
                continue                           # stale entry
            if status[v] == 1:
                for w in adj[v]:
                    if status[w] != 3:
                        prio[w] += W1
                        if status[w] == 0:
                            status[w] = 1
                        heapq.heappush(heap, (-prio[w], w))
            status[v] = 3
            done[v] = True
            seq.append(v)
            for w in adj[v]:
                if status[w] == 1:
                    status[w] = 2
                    prio[w] += W1
                    heapq.heappush(heap, (-prio[w], w))
                    for x in adj[w]:
                        if status[x] != 3:
                            prio[x] += W1
                            if status[x] == 0:
                                status[x] = 1
                            heapq.heappush(heap, (-prio[x], x))
    if kind == "rcm":
        seq.reverse()
    return np.array(seq + hubs)
