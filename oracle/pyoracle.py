"""ORACLE (test infrastructure, never shipped, never on the product path).

ctypes wrapper around oracle/liboracle.so, the CPU restatement of the reference's
AssemblyManager::assembleJacRes path (see oracle/assembly.hpp for the file:line map).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compile liboracle.so (g++, a few tens of seconds with -j)."""
    so = os.path.join(_HERE, "liboracle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-s", "-j8", "-C", _HERE])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.c_char_p]
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_print_tree.restype = C.c_char_p
        L.oracle_print_tree.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_get_string.restype = C.c_char_p
        L.oracle_get_string.argtypes = [C.c_void_p, C.c_char_p]
        _LIB = L
    return _LIB


def flatten(d, prefix=""):
    """Nested dict (the reference's input YAML minus the ANONYMOUS root) -> 'a/b/key' map."""
    out = {}
    for k, v in d.items():
        key = prefix + str(k)
        if isinstance(v, dict):
            out.update(flatten(v, key + "/"))
        else:
            if isinstance(v, bool):
                v = "true" if v else "false"
            out[key] = str(v)
    return out


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class OracleProblem:
    def __init__(self, config):
        self.L = lib()
        flat = flatten(config)
        text = "".join("%s\t%s\n" % kv for kv in flat.items())
        self.h = self.L.oracle_create(text.encode())
        if not self.h:
            raise RuntimeError(self.L.oracle_last_error().decode())
        self.h = C.c_void_p(self.h)
        sz = np.zeros(16, dtype=np.int64)
        self.L.oracle_sizes(self.h, _p(sz, C.c_int64))
        (self.dim, self.num_nodes, self.num_elems, self.nverts, self.ndof_elem, self.num_dofs, self.nnz, self.nqp,
         self.nqp_side, self.num_groups, self.num_bgroups, self.type_AD, self.nvars, self.nbases, self.workset) = [int(x) for x in sz[:15]]
        self.nodes = np.zeros((self.num_nodes, self.dim))
        self.conn = np.zeros((self.num_elems, self.nverts), dtype=np.int32)
        self.lids = np.zeros((self.num_elems, self.ndof_elem), dtype=np.int32)
        self.L.oracle_get_mesh(self.h, _p(self.nodes, C.c_double), _p(self.conn, C.c_int32), _p(self.lids, C.c_int32))
        self.rowptr = np.zeros(self.num_dofs + 1, dtype=np.int64)
        self.colind = np.zeros(self.nnz, dtype=np.int32)
        self.is_fixed = np.zeros(self.num_dofs, dtype=np.uint8)
        self.L.oracle_get_graph(self.h, _p(self.rowptr, C.c_int64), _p(self.colind, C.c_int32), _p(self.is_fixed, C.c_uint8))
        md = C.c_int32(0)
        offs = np.zeros((self.nvars, 128), dtype=np.int32)
        self.numdof = np.zeros(self.nvars, dtype=np.int32)
        self.usebasis = np.zeros(self.nvars, dtype=np.int32)
        tmp = np.zeros(self.nvars * 128, dtype=np.int32)
        self.L.oracle_get_offsets(self.h, _p(tmp, C.c_int32), _p(self.numdof, C.c_int32), _p(self.usebasis, C.c_int32), C.byref(md))
        self.maxdof = md.value
        self.offsets = tmp[: self.nvars * self.maxdof].reshape(self.nvars, self.maxdof).copy()
        self.qpts = np.zeros((self.nqp, self.dim))
        self.qwts = np.zeros(self.nqp)
        self.L.oracle_get_quadrature(self.h, _p(self.qpts, C.c_double), _p(self.qwts, C.c_double))
        self.nsides = 2 * self.dim
        self.bc_codes = np.zeros((self.nvars, self.nsides), dtype=np.int32)
        self.L.oracle_get_bcs(self.h, _p(self.bc_codes, C.c_int32))
        self.elem_nodes = np.ascontiguousarray(self.nodes[self.conn])  # (E, nverts, dim)

    def __del__(self):
        try:
            self.L.oracle_destroy(self.h)
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.L.oracle_last_error().decode())

    # ---- metadata -----------------------------------------------------------------------
    def _s(self, key):
        return self.L.oracle_get_string(self.h, key.encode()).decode()

    def modules(self):
        return self._s("modules")

    def var_names(self):
        return self._s("var_names").split(",")

    def basis_type(self, b):
        return self._s("basis_types").split(",")[b]

    def basis_order(self, b):
        return int(self._s("basis_orders").split(",")[b])

    def bc_exprs(self):
        return [[self._s("bc_expr %d %d" % (v, s)) for s in range(self.nsides)] for v in range(self.nvars)]

    # ---- reference tables -------------------------------------------------------------
    def ref_basis(self, b):
        sizes = np.zeros(5, dtype=np.int64)
        self.L.oracle_get_ref_basis(self.h, b, _p(sizes, C.c_int64), None, None, None, None)
        card, vdim = int(sizes[0]), int(sizes[1])
        val = np.zeros((card, self.nqp, vdim))
        grad = np.zeros((card, self.nqp, self.dim)) if sizes[2] else None
        curl = np.zeros((card, self.nqp, self.dim)) if sizes[3] else None
        div = np.zeros((card, self.nqp)) if sizes[4] else None
        self.L.oracle_get_ref_basis(self.h, b, _p(sizes, C.c_int64), _p(val, C.c_double), _p(grad, C.c_double),
                                    _p(curl, C.c_double), _p(div, C.c_double))
        return dict(card=card, vdim=vdim, val=val, grad=grad, curl=curl, div=div)

    def side_rule(self, side):
        pts = np.zeros((self.nqp_side, self.dim))
        wts = np.zeros(self.nqp_side)
        tu = np.zeros(3)
        tv = np.zeros(3)
        self.L.oracle_get_side_rule(self.h, side, _p(pts, C.c_double), _p(wts, C.c_double), _p(tu, C.c_double), _p(tv, C.c_double))
        return pts, wts, tu, tv

    def ref_basis_side(self, side, b, card, has_grad=True, vdim=1):
        val = np.zeros((card, self.nqp_side, vdim))
        grad = np.zeros((card, self.nqp_side, self.dim)) if has_grad else None
        self.L.oracle_get_ref_basis_side(self.h, side, b, _p(val, C.c_double), _p(grad, C.c_double))
        return val, grad

    def bgroup(self, g):
        out = np.zeros(3, dtype=np.int64)
        self.L.oracle_get_bgroup(self.h, g, _p(out, C.c_int64), None)
        ids = np.zeros(int(out[0]), dtype=np.int32)
        self.L.oracle_get_bgroup(self.h, g, _p(out, C.c_int64), _p(ids, C.c_int32))
        return dict(numElem=int(out[0]), sideset=int(out[1]), local_side=int(out[2]), elem_ids=ids)

    # ---- time integration data --------------------------------------------------------
    def set_time(self, transient, time=0.0, dt=1.0, stage=0, A=((1.0,),), b=(1.0,), c=(1.0,), bdf=(1.0, -1.0)):
        A = np.ascontiguousarray(A, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        c = np.ascontiguousarray(c, dtype=np.float64)
        bdf = np.ascontiguousarray(bdf, dtype=np.float64)
        self._nsteps = max(1, len(bdf) - 1)
        self._nstages = len(b)
        self._transient = bool(transient)
        self.L.oracle_set_time(self.h, int(transient), C.c_double(time), C.c_double(dt), int(stage), len(b),
                               _p(A, C.c_double), _p(b, C.c_double), _p(c, C.c_double), len(bdf), _p(bdf, C.c_double))

    def set_seeding(self, seedwhat=1, seedindex=0):
        """What the Jacobian differentiates with respect to: 1 the stage solution (default), 2 previous step `seedindex`
        (compute_previous_jac, assemblyManager_jacres.hpp:176-190), 3 previous stage `seedindex` (workset.cpp:727-785)."""
        self.L.oracle_set_seeding(self.h, int(seedwhat), int(seedindex))

    def set_point_dofs(self, dofs):
        """disc->point_dofs: dofConstraints turns their Jacobian rows into identity rows (assemblyManager_constraints.hpp:97-116, 261-266)."""
        d = np.ascontiguousarray(dofs, dtype=np.int32)
        self.L.oracle_set_point_dofs(self.h, len(d), d.ctypes.data_as(C.POINTER(C.c_int)))

    def set_adjoint(self, useadjoint):
        """assembleJacRes(..., useadjoint = true): the Jacobian is filled transposed (and thermal's weak-Dirichlet sides use sf = 1)."""
        self.L.oracle_set_adjoint(self.h, int(bool(useadjoint)))

    def _ptrs(self, vecs, n):
        if not getattr(self, "_transient", False):
            return None, None
        arr = (C.POINTER(C.c_double) * n)()
        keep = []
        for i in range(n):
            v = np.ascontiguousarray(vecs[i], dtype=np.float64)
            keep.append(v)
            arr[i] = _p(v, C.c_double)
        return arr, keep

    # ---- assembly -------------------------------------------------------------------
    def assemble_jacres(self, sol, sol_prev=None, sol_stage=None, compute_jacobian=True, res=None, jac=None):
        sol = np.ascontiguousarray(sol, dtype=np.float64)
        res = np.zeros(self.num_dofs) if res is None else res
        jac = np.zeros(self.nnz) if jac is None else jac
        pp, k1 = self._ptrs(sol_prev, getattr(self, "_nsteps", 1))
        ps, k2 = self._ptrs(sol_stage, getattr(self, "_nstages", 1))
        self._chk(self.L.oracle_assemble_jacres(self.h, _p(sol, C.c_double), pp, ps, int(compute_jacobian), _p(res, C.c_double), _p(jac, C.c_double)))
        return res, jac

    def assemble_res(self, sol, sol_prev=None, sol_stage=None, res=None):
        sol = np.ascontiguousarray(sol, dtype=np.float64)
        res = np.zeros(self.num_dofs) if res is None else res
        pp, k1 = self._ptrs(sol_prev, getattr(self, "_nsteps", 1))
        ps, k2 = self._ptrs(sol_stage, getattr(self, "_nstages", 1))
        self._chk(self.L.oracle_assemble_res(self.h, _p(sol, C.c_double), pp, ps, _p(res, C.c_double)))
        return res

    def apply_mass(self, mass_wts, x):
        """applyMassMatrixFree: y = M x (y starts at zero)."""
        w = np.ascontiguousarray(mass_wts, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(self.num_dofs)
        self._chk(self.L.oracle_apply_mass(self.h, _p(w, C.c_double), _p(x, C.c_double), _p(y, C.c_double)))
        return y

    def project_initial(self):
        """setInitial: right-hand side of the L2 projection of the `Initial conditions` (sum_q u0 phi_i w per row)."""
        rhs = np.zeros(self.num_dofs)
        self._chk(self.L.oracle_project_initial(self.h, _p(rhs, C.c_double)))
        return rhs

    def weighted_mass(self, mass_wts, lump=False):
        """getWeightedMass: (mass values in graph order, diagonal vector)."""
        w = np.ascontiguousarray(mass_wts, dtype=np.float64)
        M, d = np.zeros(self.nnz), np.zeros(self.num_dofs)
        self._chk(self.L.oracle_weighted_mass(self.h, _p(w, C.c_double), int(lump), _p(M, C.c_double), _p(d, C.c_double)))
        return M, d

    def set_initial(self):
        """setInitial as a whole (assemblyManager_initial.hpp:36-133): the projection's right-hand side, getMass (unit weights) and the
        routine's own fix_zero_rows loop (`bool fix_zero_rows = true`, :48, :114-131): rows with sum |M(row,:)| < 1e-14 get M(row,row) = 1."""
        rhs = self.project_initial()
        M, _ = self.weighted_mass(np.ones(len(self.var_names())))
        for row in range(self.num_dofs):
            a, b = self.rowptr[row], self.rowptr[row + 1]
            if np.abs(M[a:b]).sum() < 1.0e-14:
                M[a:b][self.colind[a:b] == row] = 1.0
        return rhs, M

    # ---- evaluation helpers (postprocess-style L2 errors, expression trees) -------------
    def group_info(self, grp, boundary=False):
        n = C.c_int64(0)
        self.L.oracle_group_info(self.h, grp, int(boundary), C.byref(n), None, None, None, None)
        ne = n.value
        npts = self.nqp_side if boundary else self.nqp
        wts = np.zeros((ne, npts))
        ip = [np.zeros((ne, npts)) for _ in range(3)]
        self.L.oracle_group_info(self.h, grp, int(boundary), C.byref(n), _p(wts, C.c_double), _p(ip[0], C.c_double),
                                 _p(ip[1], C.c_double), _p(ip[2], C.c_double))
        return wts, ip

    def eval_function(self, name, grp, loc="ip"):
        ne = self.group_info(grp, loc == "side ip")[0].shape[0]
        npts = self.nqp_side if loc == "side ip" else self.nqp
        out = np.zeros((self.workset, npts))
        self._chk(self.L.oracle_eval_function(self.h, name.encode(), loc.encode(), grp, _p(out, C.c_double)))
        return out[:ne]

    def eval_field(self, label, sol, grp):
        ne = self.group_info(grp)[0].shape[0]
        sol = np.ascontiguousarray(sol, dtype=np.float64)
        out = np.zeros((self.workset, self.nqp))
        self._chk(self.L.oracle_eval_field(self.h, label.encode(), _p(sol, C.c_double), grp, _p(out, C.c_double)))
        return out[:ne]

    def l2_error(self, labels, sol):
        """sqrt(sum_e sum_q (u_h - u_true)^2 w), summed over the given field labels
        (postprocessManager_error_estimation.hpp:266-330)."""
        err = 0.0
        for grp in range(self.num_groups):
            wts, _ = self.group_info(grp)
            for lab in labels:
                d = self.eval_field(lab, sol, grp) - self.eval_function("true " + lab, grp)
                err += float(np.sum(d * d * wts))
        return np.sqrt(err)

    def print_tree(self, name, loc="ip"):
        s = self.L.oracle_print_tree(self.h, name.encode(), loc.encode())
        if s is None:
            raise RuntimeError(self.L.oracle_last_error().decode())
        return s.decode()

    def csr(self, vals):
        import scipy.sparse as sp
        return sp.csr_matrix((vals, self.colind, self.rowptr), shape=(self.num_dofs, self.num_dofs))
