"""Synthetic inline rectangle / brick meshes and the host-side data a MrHyDE run would hand to the
C ABI for them: element vertex coordinates, HGRAD-Q1 LIDs, the overlapped CSR graph, the strong
Dirichlet mask, reference cubature and basis tables.

In a real integration these arrays come from the reference's own objects (panzer_stk mesh, Panzer
DOFManager, Tpetra CrsGraph, Intrepid2 tabulation); this module only exists so bench.py and the
multi-GPU path can build benchmark-scale inputs without the reference.  Numbering follows
SimpleMeshManager_Brick (src/tools/simplemeshmanager.hpp:1460-1512): nodes x-fastest, Hex8
connectivity in Shards order; the graph is the union over elements of gids x gids with ascending
columns (linearAlgebraInterface_construct.hpp:213-265).
"""
import numpy as np


def brick(dim, n, lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0)):
    """nodes (N, dim) float64, conn (E, 2**dim) int32."""
    n = [int(v) for v in n[:dim]]
    nn = [v + 1 for v in n]
    ax = [lo[d] + np.arange(nn[d], dtype=np.float64) * ((hi[d] - lo[d]) / n[d]) for d in range(dim)]
    if dim == 2:
        Y, X = np.meshgrid(ax[1], ax[0], indexing="ij")
        nodes = np.stack([X.ravel(), Y.ravel()], axis=1)
        j, i = np.meshgrid(np.arange(n[1]), np.arange(n[0]), indexing="ij")
        base = (j * nn[0] + i).ravel()
        conn = np.stack([base, base + 1, base + nn[0] + 1, base + nn[0]], axis=1)
    else:
        Z, Y, X = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")
        nodes = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
        k, j, i = np.meshgrid(np.arange(n[2]), np.arange(n[1]), np.arange(n[0]), indexing="ij")
        nxy = nn[0] * nn[1]
        base = (k * nxy + j * nn[0] + i).ravel()
        q = np.stack([base, base + 1, base + nn[0] + 1, base + nn[0]], axis=1)
        conn = np.concatenate([q, q + nxy], axis=1)
    return np.ascontiguousarray(nodes), np.ascontiguousarray(conn.astype(np.int32))


def q1_graph(dim, n):
    """CSR pattern of the nodal Q1 operator on the brick: rowptr int64, colind int32 (ascending)."""
    nn = [int(v) + 1 for v in n[:dim]] + [1] * (3 - dim)
    N = nn[0] * nn[1] * nn[2]
    idx = np.arange(N, dtype=np.int64)
    i = idx % nn[0]
    j = (idx // nn[0]) % nn[1]
    k = idx // (nn[0] * nn[1])
    counts = np.zeros(N, dtype=np.int64)
    cols, valid = [], []
    for dk in ((-1, 0, 1) if dim == 3 else (0,)):
        for dj in (-1, 0, 1):
            for di in (-1, 0, 1):
                ok = (i + di >= 0) & (i + di < nn[0]) & (j + dj >= 0) & (j + dj < nn[1]) & (k + dk >= 0) & (k + dk < nn[2])
                cols.append((idx + di + dj * nn[0] + dk * nn[0] * nn[1]).astype(np.int32))
                valid.append(ok)
                counts += ok
    rowptr = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    colind = np.empty(int(rowptr[-1]), dtype=np.int32)
    fill = rowptr[:-1].copy()
    for c, ok in zip(cols, valid):
        colind[fill[ok]] = c[ok]
        fill += ok
    return rowptr, colind


def graph_from_lids(lids, n_rows):
    """General pattern: union over elements of lids x lids, ascending columns."""
    import scipy.sparse as sp
    E, nd = lids.shape
    r = np.repeat(lids, nd, axis=1).ravel()
    c = np.tile(lids, (1, nd)).ravel()
    A = sp.csr_matrix((np.ones(len(r), dtype=np.int8), (r, c)), shape=(n_rows, n_rows))
    A.sum_duplicates()
    A.sort_indices()
    return A.indptr.astype(np.int64), A.indices.astype(np.int32)


def boundary_mask(dim, n):
    """1 on every node of the brick's boundary (all-sides strong Dirichlet)."""
    nn = [int(v) + 1 for v in n[:dim]] + [1] * (3 - dim)
    idx = np.arange(nn[0] * nn[1] * nn[2])
    i = idx % nn[0]
    j = (idx // nn[0]) % nn[1]
    k = idx // (nn[0] * nn[1])
    m = (i == 0) | (i == nn[0] - 1) | (j == 0) | (j == nn[1] - 1)
    if dim == 3:
        m |= (k == 0) | (k == nn[2] - 1)
    return m.astype(np.uint8)


_SIGNS = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], dtype=np.float64)


def q1_reference(dim):
    """2-point tensor Gauss rule (x fastest) and the HGRAD C1 table at its points, Shards vertex order.
    Returns pts (nqp, dim), wts (nqp), val (card, nqp, 1), grad (card, nqp, dim)."""
    g = 1.0 / np.sqrt(3.0)
    nq = 2 ** dim
    pts = np.zeros((nq, dim))
    for q in range(nq):
        r = q
        for d in range(dim):
            pts[q, d] = -g if (r % 2 == 0) else g
            r //= 2
    wts = np.ones(nq)
    nv = 2 ** dim
    val = np.zeros((nv, nq, 1))
    grad = np.zeros((nv, nq, dim))
    for v in range(nv):
        s = _SIGNS[v, :dim]
        f = 0.5 * (1.0 + s[None, :] * pts)          # (nq, dim)
        val[v, :, 0] = np.prod(f, axis=1)
        for d in range(dim):
            gd = 0.5 * s[d] * np.ones(nq)
            for o in range(dim):
                if o != d:
                    gd = gd * f[:, o]
            grad[v, :, d] = gd
    return pts, wts, val, grad


def partition_slabs(dim, n, nparts):
    """Element ranges of `nparts` slabs along the last axis (Zprocs = nparts).
    Returns a list of (elem_lo, elem_hi) in the global x-fastest element numbering."""
    n = [int(v) for v in n[:dim]]
    layers = n[-1]
    per = np.prod(n[:-1], dtype=np.int64)
    cuts = [(layers * p) // nparts for p in range(nparts + 1)]
    return [(int(cuts[p] * per), int(cuts[p + 1] * per)) for p in range(nparts)]
