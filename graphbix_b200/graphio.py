"""Host-side graph preparation shared by the sbm.jl / mod.jl drivers (the part of the reference that stays on the
host): edge-list reading, `simplegraph`, giant component, `LinksConnected`, `UniformGraphError`, `sortblock`.

Each function cites the reference lines it mirrors.  Arrays follow the reference: `links` is K x 2, 1-based,
both directions of every edge."""
from __future__ import annotations

import math

import numpy as np


def read_edgelist(path: str):
    """sbm.jl:783-800: one edge per line, "u v" separated by a single space; returns (links, Ntotinput)."""
    rows = []
    with open(path) as fp:
        for line in fp:
            line = line.rstrip("\n")
            if not line.strip():
                continue
            u = line.split(" ")
            rows.append((int(u[0]), int(u[1])))
    links = np.array(rows, dtype=np.int64).reshape(-1, 2)
    return links, int(links.max()) if links.size else 0


def simplegraph(links: np.ndarray) -> np.ndarray:
    """sbm.jl:414-425: bidirect, drop duplicates (first occurrence kept, order preserved), drop self loops."""
    links = np.asarray(links, dtype=np.int64)
    both = np.vstack([links, links[:, ::-1]])
    _, first = np.unique(both, axis=0, return_index=True)
    both = both[np.sort(first)]
    return both[both[:, 0] != both[:, 1]]


def connected_component_of(links: np.ndarray, n: int, root: int) -> np.ndarray:
    """sbm.jl:427-441 `DFS(nb, root)`: the vertices (1-based) reachable from `root`.  The reference walks an explicit
    stack with O(N) membership tests; the set of visited vertices is the same."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    A = coo_matrix((np.ones(len(links), dtype=np.int8), (links[:, 0] - 1, links[:, 1] - 1)), shape=(n, n))
    _, comp = connected_components(A, directed=False)
    return np.flatnonzero(comp == comp[root - 1]) + 1


def links_connected(links: np.ndarray, Ntotinput: int, cc: np.ndarray):
    """sbm.jl:443-475 `LinksConnected`: keep the links of component `cc`, renumber its vertices 1..Ntot in order."""
    cc = np.sort(np.asarray(cc, dtype=np.int64))
    keep = np.zeros(Ntotinput + 1, dtype=bool)
    keep[cc] = True
    newid = np.cumsum(keep)  # defects below a vertex are subtracted (sbm.jl:469-472)
    sel = keep[links[:, 0]] | keep[links[:, 1]]
    out = links[sel]
    return int(cc.size), np.column_stack([newid[out[:, 0]], newid[out[:, 1]]]).astype(np.int64)


def giant_component(links: np.ndarray, Ntotinput: int, threshold: float, log=print):
    """sbm.jl:835-847: try roots 1, 2, ... until a component with at least `threshold` vertices is found."""
    Ntot, out = 0, links
    for nn in range(1, max(int(Ntotinput / 2), 1) + 1):
        cc = connected_component_of(links, Ntotinput, nn)
        log("connected component identified...")
        Ntot, newlinks = links_connected(links, Ntotinput, cc)
        log(f"nodes & links updated... : Ntot = {Ntot} 2L = {newlinks.shape[0]}")
        if Ntot < threshold:
            log("This is not a giant component... try again.")
        else:
            out = newlinks
            break
    return Ntot, out


def degree_sequence(links: np.ndarray, Ntot: int) -> np.ndarray:
    """sbm.jl:407-412."""
    return np.bincount(links[:, 0] - 1, minlength=Ntot).astype(np.float64)


def uniform_graph_error(links: np.ndarray, Ntot: int, degreecorrection: bool) -> float:
    """sbm.jl:477-492 (note `Ntot*Ntot-1`: operator precedence of the reference kept)."""
    Ktot = links.shape[0]
    if degreecorrection:
        d = degree_sequence(links, Ntot)
        d = d[d > 0]
        return 1 + math.log(Ktot) - 2 * float(np.sum(d * np.log(d))) / Ktot
    return 1 - math.log(Ktot / (Ntot * Ntot - 1))


def sortblock(block: np.ndarray, blockprev: np.ndarray) -> np.ndarray:
    """sbm.jl:519-545: relabel the blocks of this q so that they line up with the previous q (alluvial diagram)."""
    block = np.asarray(block, dtype=np.int64).copy()
    blockprev = np.asarray(blockprev, dtype=np.int64)
    Bnow, Bprev = int(block.max()), int(blockprev.max()) if blockprev.size else 0
    if Bprev > 1:
        M = np.zeros((Bnow, Bprev), dtype=np.int64)
        np.add.at(M, (block - 1, blockprev - 1), 1)
        maxind = np.argmax(M, axis=1) + 1
        varM = np.var(M, axis=1, ddof=1) if Bprev > 1 else np.zeros(Bnow)
        order = sorted(range(Bnow), key=lambda r: (maxind[r], -varM[r]))
        relabel = np.zeros(Bnow + 1, dtype=np.int64)
        for k, r in enumerate(order):
            relabel[r + 1] = k + 1
        block = relabel[block]
    return block


def parse_q_list(Blist: str):
    """sbm.jl:764-774: "a:b", "a:s:b" or "a,b,c"."""
    parts = Blist.split(":")
    if len(parts) == 2:
        return list(range(int(parts[0]), int(parts[1]) + 1))
    if len(parts) == 3:
        return list(range(int(parts[0]), int(parts[2]) + 1, int(parts[1])))
    return [int(b) for b in Blist.split(",")]


def julia_float(x: float) -> str:
    """How Julia prints a Float64 inside string interpolation (shortest round-trip, `e` notation beyond 1e-5..1e16)."""
    if isinstance(x, (int, np.integer)):
        return str(int(x))
    x = float(x)
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "Inf" if x > 0 else "-Inf"
    r = repr(x)
    if "e" in r or "E" in r:
        m, e = r.lower().split("e")
        if "." not in m:
            m += ".0"
        return f"{m}e{int(e)}"
    a = abs(x)
    if a != 0 and (a < 1e-5 or a >= 1e16):
        m, e = f"{x:.17e}".split("e")
        m = repr(float(m))
        return f"{m}e{int(e)}"
    return r


def julia_float16(x: float) -> str:
    """`$(Float16(x))` as the .smap writer uses it (sbm.jl:576-594)."""
    s = str(np.float16(x))
    if "e" in s:
        m, e = s.split("e")
        if "." not in m:
            m += ".0"
        return f"{m}e{int(e)}"
    return s if "." in s else s + ".0"
