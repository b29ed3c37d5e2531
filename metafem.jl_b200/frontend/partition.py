"""Element-block partitioning for multi-GPU runs (new relative to the reference, which is single-GPU:
README.md:4, src/misc/04_GPU_Utils.jl:86-87,102).

Each rank owns a contiguous block of elements and holds every node its elements touch. Matrices and
residuals are kept *unassembled* across ranks: a node shared by several ranks carries a partial row on
each of them, and every vector produced by an element loop or an SpMV is completed by an
interface exchange-add (sum over the sharing ranks, in rank order so all copies are bit-identical).
Dot products count each node once, on its owner (the lowest rank that holds it).
"""
import numpy as np

from ..api import MeshTables


class Subdomain:
    """Local mesh tables of one rank + the maps back to the global (reference) numbering."""

    def __init__(self, rank, tables, node_l2g, elem_l2g, neighbors, shared, owned):
        self.rank = rank
        self.tables = tables          # MeshTables in LOCAL numbering (1-based like the reference)
        self.node_l2g = node_l2g      # local node (0-based) -> global node id (1-based)
        self.elem_l2g = elem_l2g      # local element (0-based) -> global element id (1-based)
        self.neighbors = neighbors    # sorted list of neighbour ranks
        self.shared = shared          # {neighbour rank: local node ids (1-based), ordered by global id}
        self.owned = owned            # uint8 [N_local]: 1 if this rank owns the node


def _bisect(cen, idx, lo, n_parts, part):
    """Recursive coordinate bisection: split the elements `idx` into n_parts boxes of (nearly) equal element count, cutting
    the longest extent of the current box each time; ranks lo .. lo + n_parts - 1."""
    if n_parts == 1:
        part[idx] = lo
        return
    c = cen[:, idx]
    axis = int(np.argmax(c.max(axis=1) - c.min(axis=1)))
    left = n_parts // 2
    order = idx[np.argsort(c[axis], kind="stable")]
    cut = (len(idx) * left) // n_parts
    _bisect(cen, order[:cut], lo, left, part)
    _bisect(cen, order[cut:], lo + left, n_parts - left, part)


def split_elements(tables, n_parts, method="slab"):
    """Element -> rank. 'slab': contiguous blocks of the element list sorted by centroid x (then y, z) -- what bench.py and the
    GPU tests use (two neighbours per rank, every interface node shared by exactly two ranks). 'box': recursive coordinate
    bisection into n_parts boxes (smaller interfaces, up to 26 neighbours, edge / corner nodes shared by 4 / 8 ranks; host logic
    tested on CPU, tests/test_multi_gpu.py). 'contiguous': blocks of the element list as it is."""
    cp = tables.controlpoint_IDs
    n_el = cp.shape[1]
    if method == "box":
        cen = tables.x[:, cp - 1].mean(axis=1)
        part = np.empty(n_el, np.int32)
        _bisect(cen, np.arange(n_el), 0, n_parts, part)
        return part
    if method == "contiguous":
        order = np.arange(n_el)
    else:
        cen = tables.x[:, cp - 1].mean(axis=1)                       # (3, n_el)
        order = np.lexsort((cen[2], cen[1], cen[0]))
    part = np.empty(n_el, np.int32)
    bounds = np.linspace(0, n_el, n_parts + 1).astype(np.int64)
    for r in range(n_parts):
        part[order[bounds[r]:bounds[r + 1]]] = r
    return part


def make_subdomains(tables, part, ranks=None):
    """Build the Subdomain of every rank in `ranks` (default: all) from the global tables and an element->rank map."""
    cp = tables.controlpoint_IDs
    n_parts = int(part.max()) + 1
    N = tables.x.shape[1]
    # ranks holding each node, as a bit mask (n_parts <= 64)
    assert n_parts <= 64
    held = np.zeros(N, np.uint64)
    for r in range(n_parts):
        nodes = np.unique(cp[:, part == r]) - 1
        held[nodes] |= np.uint64(1) << np.uint64(r)
    lowest = np.zeros(N, np.int32)
    rem = held.copy()
    # index of the lowest set bit
    for r in range(n_parts - 1, -1, -1):
        lowest[(held >> np.uint64(r)) & np.uint64(1) == 1] = r
    out = {}
    for r in (range(n_parts) if ranks is None else ranks):
        els = np.nonzero(part == r)[0]
        nodes = np.unique(cp[:, els]) - 1                             # sorted global ids (0-based)
        g2l = np.zeros(N, np.int32)
        g2l[nodes] = np.arange(1, len(nodes) + 1)
        lcp = g2l[cp[:, els] - 1]
        e_g2l = np.zeros(cp.shape[1], np.int32)
        e_g2l[els] = np.arange(1, len(els) + 1)
        t = tables
        fe = fi = None
        bg = {}
        if t.facet_element_ID is not None:
            keep = np.nonzero(part[t.facet_element_ID - 1] == r)[0]   # facets of this rank's elements
            f_g2l = np.zeros(len(t.facet_element_ID), np.int32)
            f_g2l[keep] = np.arange(1, len(keep) + 1)
            fe = e_g2l[t.facet_element_ID[keep] - 1]
            fi = t.facet_element_eindex[keep]
            for g, ids in t.bg_fIDs.items():
                loc = f_g2l[np.asarray(ids) - 1]
                bg[g] = loc[loc > 0]
        lt = MeshTables(controlpoint_IDs=lcp, x=t.x[:, nodes], ref_itp_vals=t.ref_itp_vals, itg_weight=t.itg_weight,
                        bdy_ref_itp_vals=t.bdy_ref_itp_vals, bdy_itg_weights=t.bdy_itg_weights,
                        bdy_tangent_directions=t.bdy_tangent_directions, facet_element_ID=fe, facet_element_eindex=fi,
                        bg_fIDs=bg)
        mine = held[nodes]
        shared, neighbors = {}, []
        for q in range(n_parts):
            if q == r:
                continue
            m = (mine >> np.uint64(q)) & np.uint64(1) == 1
            if m.any():
                neighbors.append(q)
                shared[q] = (np.nonzero(m)[0] + 1).astype(np.int32)   # local ids; `nodes` is sorted by global id
        owned = (lowest[nodes] == r).astype(np.uint8)
        out[r] = Subdomain(r, lt, (nodes + 1).astype(np.int32), (els + 1).astype(np.int32), neighbors, shared, owned)
    return out


def scatter_field(sub, global_field):
    """Global nodal array -> local nodal array of a subdomain."""
    return np.ascontiguousarray(np.asarray(global_field)[..., sub.node_l2g - 1])


def gather_owned(subs, local_fields, N):
    """Owned entries of per-rank local nodal arrays -> one global nodal array."""
    out = np.zeros(N)
    for r, sub in subs.items():
        m = sub.owned.astype(bool)
        out[sub.node_l2g[m] - 1] = np.asarray(local_fields[r])[m]
    return out
