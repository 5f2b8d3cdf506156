"""Set-up helpers that play the role of the reference host for synthetic inline meshes: they build
the arrays of mrhyde_b200.inline_mesh and feed them through the C ABI exactly as INTEGRATION.md
describes for a MrHyDE host (mesh -> set_mesh, DOF manager -> lids, CrsGraph -> set_graph,
Functions/Physics sublists -> set_function/set_option)."""
import numpy as np

from . import inline_mesh as im
from .capi import AssemblyPlan

# regression/thermal/3D_verification/input.yaml and 2D_verification/input.yaml source terms
THERMAL_SOURCE = {2: "8*(pi*pi)*sin(2*pi*x)*sin(2*pi*y)", 3: "12*(pi*pi)*sin(2*pi*x)*sin(2*pi*y)*sin(2*pi*z)"}


class NodeSlab:
    """Node-level data of rank `rank`'s slab of `nranks` bricks stacked along the last axis (Zprocs = nranks):
    owned node rows first, the ghost plane (owned by rank+1) last; the plane below a non-zero rank appears as
    column-only ghosts (the column map of the owned Tpetra matrix that exportMatrixFromOverlapped fills)."""

    def __init__(self, dim, n, rank=0, nranks=1):
        self.dim, self.n, self.rank, self.nranks = dim, [int(v) for v in n[:dim]], rank, nranks
        n = self.n
        lo = [0.0] * 3
        hi = [1.0] * 3
        lo[dim - 1] = float(rank)
        hi[dim - 1] = float(rank + 1)
        nodes, conn = im.brick(dim, n, lo, hi)
        # keep the global problem on the unit cube so the source term matches the regression deck
        nodes[:, dim - 1] /= float(nranks)
        self.nodes, self.conn = nodes, conn
        nn = [v + 1 for v in n]
        plane = int(np.prod(nn[:-1]))
        self.n_rows = int(np.prod(nn))
        self.n_owned = self.n_rows if rank == nranks - 1 else self.n_rows - plane
        self.row_gids = np.arange(self.n_rows, dtype=np.int64) + rank * n[-1] * plane
        if rank == 0:
            self.rowptr, self.colind = im.q1_graph(dim, n)
            self.col_gids = self.row_gids
        else:
            # The bottom plane is owned here but shared with rank-1's elements: its rows also couple to the plane
            # below, which no local element touches.  Those nodes become column-only ghosts (local column ids
            # >= n_rows), i.e. the column map of the owned matrix that Tpetra's export fills.
            n_ext = list(n)
            n_ext[-1] += 1
            rp, ci = im.q1_graph(dim, n_ext)           # box extended by one plane below; ext id = local id + plane
            rp = rp[plane:] - rp[plane]
            ci = ci[int(im.q1_graph(dim, n_ext)[0][plane]):].astype(np.int64) - plane
            ci = np.where(ci < 0, self.n_rows + (ci + plane), ci)   # plane -1 -> column-only ghost ids
            rows = np.repeat(np.arange(self.n_rows), np.diff(rp))
            order = np.lexsort((ci, rows))             # ascending local column ids within each row
            self.rowptr, self.colind = rp.astype(np.int64), ci[order].astype(np.int32)
            self.col_gids = np.concatenate([self.row_gids, np.arange(plane, dtype=np.int64) + (rank * n[-1] - 1) * plane])
        # strong Dirichlet on the GLOBAL boundary only (partition planes are interior)
        idx = np.arange(self.n_rows)
        fixed = np.zeros(self.n_rows, dtype=bool)
        stride = 1
        for d in range(dim):
            c = (idx // stride) % nn[d]
            if d < dim - 1:
                fixed |= (c == 0) | (c == nn[d] - 1)
            else:
                fixed |= ((c == 0) & (rank == 0)) | ((c == nn[d] - 1) & (rank == nranks - 1))
            stride *= nn[d]
        fixed = fixed.astype(np.uint8)
        self.is_fixed = fixed
        self.n_elem = conn.shape[0]


class ThermalBrick:
    """Steady thermal, HGRAD Q1, all-boundary strong Dirichlet on an inline brick; with nranks > 1 the
    global mesh is `nranks` bricks stacked along the last axis (weak scaling, Zprocs = nranks) and this
    object holds rank `rank`'s slab: owned rows first, the ghost plane (owned by rank+1) last."""

    def __init__(self, dim, n, device=0, rank=0, nranks=1, functions=None, options=None, perturb=0.0):
        self.dim, self.n, self.rank, self.nranks = dim, [int(v) for v in n[:dim]], rank, nranks
        ns = NodeSlab(dim, n, rank, nranks)
        for k in ("nodes", "conn", "n_rows", "n_owned", "row_gids", "rowptr", "colind", "col_gids", "is_fixed"):
            setattr(self, k, getattr(ns, k))
        conn = self.conn
        if perturb:
            # general (non-parallelepiped) cells: every node not on the global boundary moves by up to perturb * h per axis; the
            # displacement is a function of the global node id, so the replicas of a shared node agree across ranks
            h = np.array([1.0 / v for v in self.n]); h[dim - 1] /= nranks
            g = self.row_gids.astype(np.float64)
            d = np.stack([np.modf(np.sin(g * (12.9898 + 7.233 * a)) * 43758.5453)[0] for a in range(dim)], axis=1)
            self.nodes = self.nodes + (self.is_fixed == 0)[:, None] * (perturb * h)[None, :] * d
        nodes = self.nodes
        self.lids = conn  # Q1 scalar field: local dof id == local node id (owned planes first by construction)
        self.n_elem = conn.shape[0]
        self.nnz = int(self.rowptr[-1])
        pts, wts, val, grad = im.q1_reference(dim)
        nv = 2 ** dim
        self.plan = AssemblyPlan("thermal", dim, ["T"], [0], [dict(type="HGRAD", order=1, card=nv, val=val, grad=grad)], nv,
                                 np.arange(nv, dtype=np.int32).reshape(1, nv), pts, wts, device=device)
        fn = {"thermal source": THERMAL_SOURCE[dim]}
        fn.update(functions or {})
        for k, v in fn.items():
            self.plan.set_function(k, v)
        for k, v in (options or {}).items():
            self.plan.set_option(k, v)
        self.plan.set_mesh_indexed(nodes, conn, self.lids)
        self.plan.set_graph(self.rowptr, self.colind, self.is_fixed, n_owned=self.n_owned)
        self.plan.finalize()

    def state(self, seed=20261017):
        """u = manufactured solution + 1e-3 U(-1,1), same values on shared rows of every rank."""
        x = self.nodes
        u = np.prod(np.sin(2.0 * np.pi * x), axis=1)
        if self.nranks == 1:
            noise = np.random.default_rng(seed).uniform(-1.0, 1.0, size=self.n_rows)
        else:  # a function of the global id, so the replicas of a shared row agree across ranks
            noise = np.modf(np.sin(self.row_gids * 12.9898) * 43758.5453)[0]
        return u + 1e-3 * noise

    # algorithmic bytes per element, SURVEY 8(d): LIDs + unique vertices + state + residual + J values
    def algorithmic_bytes(self):
        nv = 2 ** self.dim
        return 4.0 * nv * self.n_elem + 8.0 * self.dim * self.nodes.shape[0] + 8.0 * self.n_rows + 8.0 * self.n_rows + 8.0 * self.nnz


def _expand_graph(rowptr, colind, nvar):
    """Node graph -> dof graph for `nvar` fields interleaved per node (dof = nvar * node + field), ascending columns."""
    n = len(rowptr) - 1
    cnt = np.diff(rowptr)
    rp = np.zeros(n * nvar + 1, dtype=np.int64)
    np.cumsum(np.repeat(cnt * nvar, nvar), out=rp[1:])
    # every dof row of a node repeats the node's column list with each column expanded to nvar consecutive dofs
    base = (colind.astype(np.int64)[:, None] * nvar + np.arange(nvar)[None, :]).reshape(-1)     # per node row: len*nvar entries
    seg = np.repeat(np.arange(n), cnt * nvar)                                                     # node row of each entry in `base`
    # rows are [node0 x nvar copies, node1 x nvar copies, ...]; build by repeating each node's segment nvar times
    starts = np.concatenate([[0], np.cumsum(cnt * nvar)])
    idx = np.concatenate([np.tile(np.arange(starts[r], starts[r + 1]), nvar) for r in range(n)]) if n < 4096 else None
    if idx is None:
        # vectorised for large meshes: position p in the output belongs to dof row rho = searchsorted(rp, p); node = rho // nvar
        total = int(rp[-1])
        out = np.empty(total, dtype=np.int32)
        chunk = 1 << 24
        for lo in range(0, total, chunk):
            p = np.arange(lo, min(total, lo + chunk), dtype=np.int64)
            rho = np.searchsorted(rp, p, side="right") - 1
            node = rho // nvar
            out[lo:lo + len(p)] = base[starts[node] + (p - rp[rho])]
        return rp, out
    del seg
    return rp, base[idx].astype(np.int32)


class SystemBrick:
    """Steady multi-field HGRAD-Q1 system on an inline brick through the general path: "linearelasticity"
    (dx,dy[,dz], all-boundary strong Dirichlet) or "navier stokes" (ux,pr,uy[,uz], velocity fixed on the boundary,
    SUPG+PSPG).  DOFs are interleaved per node in the module's variable order (oracle/mesh.hpp)."""

    VARS = {"linearelasticity": {2: ["dx", "dy"], 3: ["dx", "dy", "dz"]}, "navier stokes": {2: ["ux", "pr", "uy"], 3: ["ux", "pr", "uy", "uz"]}}

    def __init__(self, physics, dim, n, device=0, rank=0, nranks=1, functions=None, options=None):
        self.physics, self.dim, self.n, self.rank, self.nranks = physics, dim, [int(v) for v in n[:dim]], rank, nranks
        names = self.VARS[physics][dim]
        nvar = len(names)
        ns = NodeSlab(dim, n, rank, nranks)
        nodes, conn = ns.nodes, ns.conn
        self.nodes, self.conn = nodes, conn
        nv = 2 ** dim
        self.lids = np.ascontiguousarray((conn.astype(np.int64)[:, :, None] * nvar + np.arange(nvar)[None, None, :]).reshape(conn.shape[0], nv * nvar).astype(np.int32))
        # dofs interleaved per node: rows (owned nodes first) and column-only ghosts expand node by node
        self.rowptr, self.colind = _expand_graph(ns.rowptr, ns.colind, nvar)
        self.n_rows = ns.n_rows * nvar
        self.n_owned = ns.n_owned * nvar
        self.row_gids = (ns.row_gids[:, None] * nvar + np.arange(nvar)[None, :]).reshape(-1)
        self.col_gids = (ns.col_gids[:, None] * nvar + np.arange(nvar)[None, :]).reshape(-1)
        fixed = np.zeros((ns.n_rows, nvar), dtype=np.uint8)
        for v, name in enumerate(names):
            if name != "pr":
                fixed[ns.is_fixed.astype(bool), v] = 1
        self.is_fixed = fixed.reshape(-1)
        self.n_elem = conn.shape[0]
        self.nnz = int(self.rowptr[-1])
        pts, wts, val, grad = im.q1_reference(dim)
        offsets = np.ascontiguousarray((np.arange(nv)[None, :] * nvar + np.arange(nvar)[:, None]).astype(np.int32))
        self.plan = AssemblyPlan(physics, dim, names, [0] * nvar, [dict(type="HGRAD", order=1, card=nv, val=val, grad=grad)], nv * nvar,
                                 offsets, pts, wts, device=device)
        if physics == "linearelasticity":
            fn = {"lambda": "1.0", "mu": "1.0", "source dx": "sin(pi*x)*sin(pi*y)", "source dy": "sin(2*pi*x)*sin(2*pi*y)"}
            if dim == 3:
                fn["source dz"] = "sin(3*pi*x)*sin(3*pi*y)*sin(3*pi*z)"
            opts = {}
        else:
            fn = {"source ux": "1.0", "viscosity": "1.0", "density": "1.0"}
            opts = {"useSUPG": "true", "usePSPG": "true"}
        fn.update(functions or {})
        opts.update(options or {})
        for k, v in fn.items():
            self.plan.set_function(k, v)
        for k, v in opts.items():
            self.plan.set_option(k, v)
        self.plan.set_mesh_indexed(nodes, conn, self.lids)
        self.plan.set_graph(self.rowptr, self.colind, self.is_fixed, n_owned=self.n_owned)
        self.plan.finalize()

    def state(self, seed=20261017):
        x = self.nodes
        nvar = self.n_rows // x.shape[0]
        u = np.stack([np.prod(np.sin((v + 1) * np.pi * x), axis=1) for v in range(nvar)], axis=1).reshape(-1)
        if self.nranks == 1:
            noise = np.random.default_rng(seed).uniform(-1.0, 1.0, size=self.n_rows)
        else:  # a function of the global id, so the replicas of a shared row agree across ranks
            noise = np.modf(np.sin(self.row_gids * 12.9898) * 43758.5453)[0]
        return u + 1e-3 * noise

    def algorithmic_bytes(self):
        nd = self.lids.shape[1]
        return 4.0 * nd * self.n_elem + 8.0 * self.dim * self.nodes.shape[0] + 8.0 * self.n_rows + 8.0 * self.n_rows + 8.0 * self.nnz


def q2_reference_3d():
    """3-point tensor Gauss rule (degree 4 = 2 * order, x fastest) and the HGRAD hex C2 (equispaced Lagrange) table at its points,
    tensor ordinal a + 3 b + 9 c: pts (27, 3), wts (27), val (27, 27, 1), grad (27, 27, 3)."""
    g = np.sqrt(0.6)
    x1, w1 = np.array([-g, 0.0, g]), np.array([5.0, 8.0, 5.0]) / 9.0
    l = np.stack([x1 * (x1 - 1.0) / 2.0, 1.0 - x1 * x1, x1 * (x1 + 1.0) / 2.0])       # l[a][point]
    dl = np.stack([x1 - 0.5, -2.0 * x1, x1 + 0.5])
    pts, wts = np.zeros((27, 3)), np.zeros(27)
    val, grad = np.zeros((27, 27, 1)), np.zeros((27, 27, 3))
    for q in range(27):
        qi = (q % 3, (q // 3) % 3, q // 9)
        pts[q] = [x1[qi[0]], x1[qi[1]], x1[qi[2]]]
        wts[q] = w1[qi[0]] * w1[qi[1]] * w1[qi[2]]
        for o in range(27):
            a = (o % 3, (o // 3) % 3, o // 9)
            val[o, q, 0] = l[a[0], qi[0]] * l[a[1], qi[1]] * l[a[2], qi[2]]
            grad[o, q, 0] = dl[a[0], qi[0]] * l[a[1], qi[1]] * l[a[2], qi[2]]
            grad[o, q, 1] = l[a[0], qi[0]] * dl[a[1], qi[1]] * l[a[2], qi[2]]
            grad[o, q, 2] = l[a[0], qi[0]] * l[a[1], qi[1]] * dl[a[2], qi[2]]
    return pts, wts, val, grad


def q2_node_graph(n):
    """CSR pattern of the nodal hex-Q2 operator on a brick of n = (nx, ny, nz) elements (an int means a cube): a lattice node couples to
    every node of the elements that contain it, i.e. per axis the index range [I-2, I+2] for an even (vertex-like) index and
    [I-1, I+1] for an odd one."""
    nn = [int(n)] * 3 if np.isscalar(n) else [int(v) for v in n]
    M = [2 * v + 1 for v in nn]
    lo, hi, cnt1 = [], [], []
    for d in range(3):
        ax = np.arange(M[d])
        lo.append(np.where(ax % 2 == 0, np.maximum(ax - 2, 0), ax - 1))
        hi.append(np.where(ax % 2 == 0, np.minimum(ax + 2, M[d] - 1), ax + 1))
        cnt1.append(hi[d] - lo[d] + 1)
    K, J, I = np.meshgrid(np.arange(M[2]), np.arange(M[1]), np.arange(M[0]), indexing="ij")
    Iv, Jv, Kv = I.ravel(), J.ravel(), K.ravel()
    cnt = (cnt1[0][Iv] * cnt1[1][Jv] * cnt1[2][Kv]).astype(np.int64)
    rowptr = np.zeros(M[0] * M[1] * M[2] + 1, dtype=np.int64)
    np.cumsum(cnt, out=rowptr[1:])
    colind = np.empty(int(rowptr[-1]), dtype=np.int32)
    # rows grouped by their (cx, cy, cz) extents so that each group is a dense block
    for cz in (3, 4, 5):
        for cy in (3, 4, 5):
            for cx in (3, 4, 5):
                sel = np.nonzero((cnt1[0][Iv] == cx) & (cnt1[1][Jv] == cy) & (cnt1[2][Kv] == cz))[0]
                if len(sel) == 0:
                    continue
                dz, dy, dx = np.meshgrid(np.arange(cz), np.arange(cy), np.arange(cx), indexing="ij")
                for a in range(0, len(sel), 1 << 18):   # bounded temporaries
                    ss = sel[a:a + (1 << 18)]
                    cols = ((lo[0][Iv[ss]][:, None] + dx.ravel()[None, :]) + M[0] * ((lo[1][Jv[ss]][:, None] + dy.ravel()[None, :]) + M[1] * (lo[2][Kv[ss]][:, None] + dz.ravel()[None, :])))
                    dst = rowptr[ss][:, None] + np.arange(cx * cy * cz)[None, :]
                    colind[dst.ravel()] = cols.ravel().astype(np.int32)
    return rowptr, colind


class ElasticityQ2Brick:
    """BASELINE configs[2]: 3-D linear elasticity, hex Q2 (27 nodes x 3 displacement dofs = 81 per element, 3^3 Gauss points),
    lambda = mu = 1, all-boundary strong Dirichlet, on an n^3 inline brick (Hex8 geometry: the cell topology stays
    Hexahedron_8, discretizationInterface_basis.hpp:244-252).  One rank."""

    def __init__(self, n, device=0, options=None, physics="linearelasticity", rank=0, nranks=1, nz=None):
        """nranks > 1: the n x n x (nz * nranks) brick is cut into z-slabs of nz element layers (Zprocs = nranks).  Rank `rank` numbers
        the lattice nodes of its slab x-fastest, so the owned planes come first and the top plane (owned by rank + 1) last as ghost
        rows; the two lattice planes under a non-zero rank are column-only ghosts of its bottom plane."""
        self.n, self.rank, self.nranks = n, rank, nranks
        nz = n if nz is None else int(nz)
        nzt = nz * nranks
        nodes, conn = im.brick(3, [n, n, nz], [0.0, 0.0, float(rank)], [1.0, 1.0, float(rank + 1)])
        nodes[:, 2] /= float(nranks)
        self.nodes, self.conn = nodes, conn
        nvar = 3 if physics == "linearelasticity" else 1      # "thermal": scalar hex-Q2 (thermal/2D_verification_highorder's 3-D analogue)
        M, Mz = 2 * n + 1, 2 * nz + 1
        k, j, i = np.meshgrid(np.arange(nz), np.arange(n), np.arange(n), indexing="ij")
        base = (2 * i + M * (2 * j + M * 2 * k)).ravel().astype(np.int64)
        o = np.arange(27)
        offs = (o % 3) + M * ((o // 3) % 3) + M * M * (o // 9)
        lat = base[:, None] + offs[None, :]                                                      # (E, 27) lattice node of every basis function
        self.lids = np.ascontiguousarray((lat[:, :, None] * nvar + np.arange(nvar)[None, None, :]).reshape(len(base), 27 * nvar).astype(np.int32))
        plane = M * M
        n_nodes = plane * Mz
        if rank == 0:
            rp, ci = q2_node_graph((n, n, nz))
            node_gids_cols = np.arange(n_nodes, dtype=np.int64)
        else:
            # box extended by one element layer (two lattice planes) below: ext id = local id + 2 planes; the rows of those planes are
            # dropped and their ids become column-only ghosts (local column ids >= n_nodes)
            rpe, cie = q2_node_graph((n, n, nz + 1))
            a = int(rpe[2 * plane])
            rp = rpe[2 * plane:] - a
            ci = cie[a:].astype(np.int64) - 2 * plane
            ci = np.where(ci < 0, n_nodes + (ci + 2 * plane), ci)
            rows = np.repeat(np.arange(n_nodes), np.diff(rp))
            order = np.lexsort((ci, rows))
            ci = ci[order].astype(np.int32)
            node_gids_cols = np.concatenate([np.arange(n_nodes, dtype=np.int64), np.arange(2 * plane, dtype=np.int64) - 2 * plane])
        node_gid0 = rank * 2 * nz * plane                       # global lattice id of local node 0
        node_gids_cols = node_gids_cols + node_gid0
        self.rowptr, self.colind = _expand_graph(rp, ci, nvar) if nvar > 1 else (rp.astype(np.int64), ci.astype(np.int32))
        self.n_rows = nvar * n_nodes
        self.n_owned = self.n_rows if rank == nranks - 1 else nvar * (n_nodes - plane)
        self.col_gids = (node_gids_cols[:, None] * nvar + np.arange(nvar)[None, :]).reshape(-1)
        self.row_gids = self.col_gids[: self.n_rows]
        idx = np.arange(n_nodes)
        li, lj, lk = idx % M, (idx // M) % M, idx // plane + rank * 2 * nz
        Mzt = 2 * nzt + 1
        bnd = (li == 0) | (li == M - 1) | (lj == 0) | (lj == M - 1) | (lk == 0) | (lk == Mzt - 1)
        self.is_fixed = np.repeat(bnd.astype(np.uint8), nvar)
        self.lattice = np.stack([li / float(M - 1), lj / float(M - 1), lk / float(Mzt - 1)], axis=1)
        self.n_elem = conn.shape[0]
        self.nnz = int(self.rowptr[-1])
        pts, wts, val, grad = q2_reference_3d()
        offsets = np.ascontiguousarray((np.arange(27)[None, :] * nvar + np.arange(nvar)[:, None]).astype(np.int32))
        names = ["dx", "dy", "dz"] if nvar == 3 else ["T"]
        self.plan = AssemblyPlan(physics, 3, names, [0] * nvar, [dict(type="HGRAD", order=2, card=27, val=val, grad=grad)], 27 * nvar,
                                 offsets, pts, wts, device=device)
        fns = {"lambda": "1.0", "mu": "1.0", "source dx": "sin(pi*x)*sin(pi*y)*sin(pi*z)", "source dy": "sin(2*pi*x)*sin(2*pi*y)*sin(2*pi*z)",
               "source dz": "sin(3*pi*x)*sin(3*pi*y)*sin(3*pi*z)"} if nvar == 3 else {"thermal source": THERMAL_SOURCE[3]}
        for kf, vf in fns.items():
            self.plan.set_function(kf, vf)
        for kf, vf in (options or {}).items():
            self.plan.set_option(kf, vf)
        self.plan.set_mesh_indexed(nodes, conn, self.lids)
        self.plan.set_graph(self.rowptr, self.colind, self.is_fixed, n_owned=self.n_owned)
        self.plan.finalize()

    def state(self, seed=20261017):
        x = self.lattice
        nvar = self.n_rows // x.shape[0]
        u = np.stack([np.prod(np.sin((v + 1) * np.pi * x), axis=1) for v in range(nvar)], axis=1).reshape(-1)
        if self.nranks == 1:
            return u + 1e-3 * np.random.default_rng(seed).uniform(-1.0, 1.0, size=self.n_rows)
        return u + 1e-3 * np.modf(np.sin(self.row_gids * 12.9898) * 43758.5453)[0]   # same values on shared rows of every rank

    def algorithmic_bytes(self):
        return 4.0 * self.lids.shape[1] * self.n_elem + 24.0 * self.nodes.shape[0] + 16.0 * self.n_rows + 8.0 * self.nnz


def hcurl_hdiv_reference_3d():
    """2-point tensor Gauss rule and the lowest-order hex HCURL (12 edge functions: x-, y-, z-directed, tensor order) and
    HDIV (6 face functions: -x,+x,-y,+y,-z,+z normals) tables at its points.
    Returns pts, wts, (curl_val (12,8,3), curl_curl (12,8,3)), (div_val (6,8,3), div_div (6,8))."""
    pts, wts, _, _ = im.q1_reference(3)
    nq = 8
    ev, ec = np.zeros((12, nq, 3)), np.zeros((12, nq, 3))
    fv, fd = np.zeros((6, nq, 3)), np.zeros((6, nq))
    dl = (-0.5, 0.5)
    for p in range(nq):
        x, y, z = pts[p]
        lx, ly, lz = (0.5 * (1 - x), 0.5 * (1 + x)), (0.5 * (1 - y), 0.5 * (1 + y)), (0.5 * (1 - z), 0.5 * (1 + z))
        f = 0
        for k in range(2):
            for j in range(2):          # x-directed: (ly_j lz_k, 0, 0)
                ev[f, p, 0] = ly[j] * lz[k]; ec[f, p, 1] = ly[j] * dl[k]; ec[f, p, 2] = -dl[j] * lz[k]; f += 1
        for k in range(2):
            for i in range(2):          # y-directed: (0, lx_i lz_k, 0)
                ev[f, p, 1] = lx[i] * lz[k]; ec[f, p, 0] = -lx[i] * dl[k]; ec[f, p, 2] = dl[i] * lz[k]; f += 1
        for j in range(2):
            for i in range(2):          # z-directed: (0, 0, lx_i ly_j)
                ev[f, p, 2] = lx[i] * ly[j]; ec[f, p, 0] = lx[i] * dl[j]; ec[f, p, 1] = -dl[i] * ly[j]; f += 1
        c = (x, y, z)
        f = 0
        for d in range(3):
            for i in range(2):          # d-normal: l_i(x_d) e_d
                fv[f, p, d] = 0.5 * (1 - c[d]) if i == 0 else 0.5 * (1 + c[d]); fd[f, p] = dl[i]; f += 1
    return pts, wts, (ev, ec), (fv, fd)


def slab_partition(lids_mine, lids_below, lids_above, fixed_of_gid=None):
    """Overlapped numbering of one rank of an element-wise z-slab partition from GLOBAL element dof lists (SURVEY 8(e);
    discretizationInterface_dof.hpp:129-137: owned dofs first, then ghosts): `lids_mine` are the rank's elements, `lids_below` the
    element layer under the slab (rank - 1's top layer, or None) and `lids_above` the layer over it (rank + 1's first layer, or None).
    A dof shared with the rank above is a GHOST row here (owned there); dofs that only the layer below touches are column-only
    ghosts of the owned interface rows.  Returns dict(lids local, n_rows, n_owned, col_gids, rowptr, colind, is_fixed)."""
    mine = np.unique(lids_mine)
    ghost_mask = np.isin(mine, np.unique(lids_above)) if lids_above is not None and len(lids_above) else np.zeros(len(mine), dtype=bool)
    owned, ghost = mine[~ghost_mask], mine[ghost_mask]
    colonly = np.setdiff1d(np.unique(lids_below), mine) if lids_below is not None and len(lids_below) else np.zeros(0, dtype=mine.dtype)
    col_gids = np.concatenate([owned, ghost, colonly]).astype(np.int64)
    order = np.argsort(col_gids, kind="stable")
    sorted_g = col_gids[order]

    def g2l(a):
        return order[np.searchsorted(sorted_g, a)].astype(np.int32)
    n_rows, n_owned = len(owned) + len(ghost), len(owned)
    local = g2l(lids_mine)
    ext = local if len(colonly) == 0 and (lids_below is None or not len(lids_below)) else np.concatenate([local, g2l(lids_below)], axis=0)
    rp, ci = im.graph_from_lids(ext, len(col_gids))
    rp, ci = rp[: n_rows + 1], ci[: rp[n_rows]]
    fixed = np.zeros(n_rows, dtype=np.uint8) if fixed_of_gid is None else fixed_of_gid(col_gids[:n_rows]).astype(np.uint8)
    return dict(lids=np.ascontiguousarray(local), n_rows=n_rows, n_owned=n_owned, col_gids=col_gids, row_gids=col_gids[:n_rows],
                rowptr=rp.astype(np.int64), colind=ci.astype(np.int32), is_fixed=fixed)


def maxwell_lids(n, nzt, k0, k1):
    """Global dof lists of the elements of layers [k0, k1) of an n x n x nzt brick: E edges (x-, y-, z-directed lattices), then B
    faces (x-, y-, z-normal lattices)."""
    if k1 <= k0:
        return np.zeros((0, 18), dtype=np.int64)
    k, j, i = [a.ravel().astype(np.int64) for a in np.meshgrid(np.arange(k0, k1), np.arange(n), np.arange(n), indexing="ij")]
    n1, nz1 = n + 1, nzt + 1
    nxe, nye, nze = n * n1 * nz1, n1 * n * nz1, n1 * n1 * nzt
    cols = []
    for d in range(12):
        if d < 4:      # x-directed: d = j + 2k
            jj, kk = j + (d % 2), k + (d // 2)
            cols.append(i + n * (jj + n1 * kk))
        elif d < 8:    # y-directed: d-4 = i + 2k
            ii, kk = i + ((d - 4) % 2), k + ((d - 4) // 2)
            cols.append(nxe + ii + n1 * (j + n * kk))
        else:          # z-directed: d-8 = i + 2j
            ii, jj = i + ((d - 8) % 2), j + ((d - 8) // 2)
            cols.append(nxe + nye + ii + n1 * (jj + n1 * k))
    ne_dofs = nxe + nye + nze
    nxf, nyf = n1 * n * nzt, n * n1 * nzt
    for d in range(6):
        dr, s_ = d // 2, d % 2
        ii, jj, kk = i + (s_ if dr == 0 else 0), j + (s_ if dr == 1 else 0), k + (s_ if dr == 2 else 0)
        if dr == 0:
            cols.append(ne_dofs + ii + n1 * (jj + n * kk))
        elif dr == 1:
            cols.append(ne_dofs + nxf + ii + n * (jj + n1 * kk))
        else:
            cols.append(ne_dofs + nxf + nyf + ii + n * (jj + n * kk))
    return np.stack(cols, axis=1)


class MaxwellBrick:
    """BASELINE configs[4]: 3-D Maxwell, lowest-order HCURL E (12 edge dofs) + HDIV B (6 face dofs) per hex, eps = mu = n = 1,
    sigma = 0, on an inline brick; transient stages are supplied per call.  DOF numbering: E edges (x-, y-, z-directed
    lattices), then B faces (x-, y-, z-normal lattices), all orientation signs +1 on a lexicographic brick.  With nranks > 1 the
    n x n x (nz * nranks) brick is cut into z-slabs of nz layers (`slab_partition`): rank `rank` holds its owned edge / face dofs
    first, the dofs of its top plane (owned by rank + 1) as ghost rows, and the layer below as column-only ghosts."""

    def __init__(self, n, device=0, functions=None, options=None, rank=0, nranks=1, nz=None):
        self.n, self.rank, self.nranks = n, rank, nranks
        nz = n if nz is None else int(nz)
        nzt = nz * nranks
        if nranks > 1:
            lo, hi = [0.0, 0.0, float(rank)], [1.0, 1.0, float(rank + 1)]
            nodes, conn = im.brick(3, [n, n, nz], lo, hi)
            nodes[:, 2] /= float(nranks)
            k0 = rank * nz
            part = slab_partition(maxwell_lids(n, nzt, k0, k0 + nz), maxwell_lids(n, nzt, k0 - 1, k0) if rank > 0 else None,
                                  maxwell_lids(n, nzt, k0 + nz, k0 + nz + 1) if rank < nranks - 1 else None)
            self.nodes, self.conn = nodes, conn
            for key in ("lids", "n_rows", "n_owned", "col_gids", "row_gids", "rowptr", "colind", "is_fixed"):
                setattr(self, key, part[key])
            self.n_elem = conn.shape[0]
            self.nnz = int(self.rowptr[-1])
            self._finish(device, functions, options)
            return
        nodes, conn = im.brick(3, [n, n, nz])
        self.nodes, self.conn = nodes, conn
        if nz != n:
            self.lids = np.ascontiguousarray(maxwell_lids(n, nz, 0, nz).astype(np.int32))
            self.n_rows = int(self.lids.max()) + 1
            self.n_owned = self.n_rows
            self.row_gids = self.col_gids = np.arange(self.n_rows, dtype=np.int64)
            self.rowptr, self.colind = im.graph_from_lids(self.lids, self.n_rows)
            self.is_fixed = np.zeros(self.n_rows, dtype=np.uint8)
            self.n_elem = conn.shape[0]
            self.nnz = int(self.rowptr[-1])
            self._finish(device, functions, options)
            return
        k, j, i = [a.ravel().astype(np.int64) for a in np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")]
        n1 = n + 1
        nxe, nye, nze = n * n1 * n1, n1 * n * n1, n1 * n1 * n
        cols = []
        for d in range(12):
            if d < 4:      # x-directed: d = j + 2k
                jj, kk = j + (d % 2), k + (d // 2)
                cols.append(i + n * (jj + n1 * kk))
            elif d < 8:    # y-directed: d-4 = i + 2k
                ii, kk = i + ((d - 4) % 2), k + ((d - 4) // 2)
                cols.append(nxe + ii + n1 * (j + n * kk))
            else:          # z-directed: d-8 = i + 2j
                ii, jj = i + ((d - 8) % 2), j + ((d - 8) // 2)
                cols.append(nxe + nye + ii + n1 * (jj + n1 * k))
        ne_dofs = nxe + nye + nze
        nxf, nyf = n1 * n * n, n * n1 * n
        for d in range(6):
            dr, s = d // 2, d % 2
            ii, jj, kk = i + (s if dr == 0 else 0), j + (s if dr == 1 else 0), k + (s if dr == 2 else 0)
            if dr == 0:
                cols.append(ne_dofs + ii + n1 * (jj + n * kk))
            elif dr == 1:
                cols.append(ne_dofs + nxf + ii + n * (jj + n1 * kk))
            else:
                cols.append(ne_dofs + nxf + nyf + ii + n * (jj + n * kk))
        self.lids = np.ascontiguousarray(np.stack(cols, axis=1).astype(np.int32))
        self.n_rows = ne_dofs + nxf + nyf + n * n * n1
        self.n_owned = self.n_rows
        self.row_gids = self.col_gids = np.arange(self.n_rows, dtype=np.int64)
        self.rowptr, self.colind = im.graph_from_lids(self.lids, self.n_rows)
        self.is_fixed = np.zeros(self.n_rows, dtype=np.uint8)
        self.n_elem = conn.shape[0]
        self.nnz = int(self.rowptr[-1])
        self._finish(device, functions, options)

    def _finish(self, device, functions, options):
        nodes, conn = self.nodes, self.conn
        pts, wts, (ev, ec), (fv, fd) = hcurl_hdiv_reference_3d()
        offsets = np.full((2, 12), -1, dtype=np.int32)
        offsets[0, :] = np.arange(12)
        offsets[1, :6] = 12 + np.arange(6)
        self.plan = AssemblyPlan("maxwell", 3, ["E", "B"], [0, 1], [dict(type="HCURL", order=1, card=12, val=ev, curl=ec), dict(type="HDIV", order=1, card=6, val=fv, div=fd)],
                                 18, offsets, pts, wts, device=device)
        for kf, vf in dict({"current x": "sin(2*pi*z)"}, **(functions or {})).items():
            self.plan.set_function(kf, vf)
        for kf, vf in (options or {}).items():
            self.plan.set_option(kf, vf)
        self.plan.set_mesh_indexed(nodes, conn, self.lids)
        self.plan.set_graph(self.rowptr, self.colind, self.is_fixed, n_owned=self.n_owned)
        self.plan.finalize()

    def state(self, seed=20261017):
        if self.nranks == 1:
            return 0.3 * np.sin(1.0 + 0.7 * np.arange(self.n_rows) / self.n_rows) + 1e-3 * np.random.default_rng(seed).uniform(-1.0, 1.0, self.n_rows)
        g = self.row_gids.astype(np.float64)   # a function of the global id: replicas of a shared row agree across ranks
        return 0.3 * np.sin(1.0 + 0.7e-6 * g) + 1e-3 * np.modf(np.sin(g * 12.9898) * 43758.5453)[0]

    def algorithmic_bytes(self):
        return 4.0 * 18 * self.n_elem + 24.0 * self.nodes.shape[0] + 16.0 * self.n_rows + 8.0 * self.nnz
