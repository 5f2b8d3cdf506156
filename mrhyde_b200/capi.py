"""ctypes binding of include/mrhyde_b200.h (the C ABI a MrHyDE host links against).

This module is plumbing: it loads mrhyde_b200/libmrhyde_b200.so, declares every entry point of the
header, and wraps a plan in a small Python class so tests and bench.py can drive the same calls the
reference's C++ host would make.  There is no fallback: if the library is missing
or a call fails an exception is raised.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MRHYDE_B200_LIB: an alternative build of the same library (tools/asan_check.sh points it at an AddressSanitizer build)
LIB_PATH = os.environ.get("MRHYDE_B200_LIB") or os.path.join(_HERE, "libmrhyde_b200.so")
_LIB = None
_EMU = None

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_PARSE, ERR_CUDA, ERR_STATE, ERR_NCCL = range(7)

# every symbol include/mrhyde_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "mrhyde_b200_version", "mrhyde_b200_last_error", "mrhyde_b200_plan_create", "mrhyde_b200_plan_destroy",
    "mrhyde_b200_plan_set_function", "mrhyde_b200_plan_set_option", "mrhyde_b200_plan_set_mesh",
    "mrhyde_b200_plan_set_mesh_indexed", "mrhyde_b200_plan_set_graph", "mrhyde_b200_plan_set_sidesets",
    "mrhyde_b200_plan_set_bc", "mrhyde_b200_plan_add_boundary_group", "mrhyde_b200_plan_finalize",
    "mrhyde_b200_assemble_jacres", "mrhyde_b200_assemble_res", "mrhyde_b200_assemble_jacres_host",
    "mrhyde_b200_comm_unique_id", "mrhyde_b200_plan_comm_init", "mrhyde_b200_plan_set_halo", "mrhyde_b200_halo_sum",
    "mrhyde_b200_plan_stat", "mrhyde_b200_plan_kernel_time", "mrhyde_b200_plan_eval_function",
    "mrhyde_b200_expr_disassemble", "mrhyde_b200_expr_eval_host", "mrhyde_b200_plan_debug_scatter_host", "mrhyde_b200_plan_debug_jit", "mrhyde_b200_plan_debug_metric_host", "mrhyde_b200_plan_debug_class_host", "mrhyde_b200_plan_debug_stage_map", "mrhyde_b200_plan_debug_chain_rows", "mrhyde_b200_project_initial", "mrhyde_b200_plan_debug_emulate_initial",
    "mrhyde_b200_plan_debug_emulate", "mrhyde_b200_debug_set_emulator", "mrhyde_b200_plan_warmup", "mrhyde_b200_assemble_jacres_adjoint", "mrhyde_b200_plan_owned_extent", "mrhyde_b200_plan_set_point_dofs", "mrhyde_b200_set_initial", "mrhyde_b200_assemble_mass", "mrhyde_b200_plan_debug_emulate_mass",
    "mrhyde_b200_apply_mass", "mrhyde_b200_plan_debug_emulate_apply_mass",
]


class MrhydeB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("mrhyde_b200 error %d: %s" % (code, msg))
        self.code = code
        self.message = msg


class Basis(C.Structure):
    _fields_ = [("type", C.c_char_p), ("order", C.c_int32), ("card", C.c_int32), ("val", C.POINTER(C.c_double)),
                ("grad", C.POINTER(C.c_double)), ("curl", C.POINTER(C.c_double)), ("div", C.POINTER(C.c_double))]


class Desc(C.Structure):
    _fields_ = [("physics", C.c_char_p), ("dim", C.c_int32), ("nvars", C.c_int32), ("var_names", C.POINTER(C.c_char_p)),
                ("var_basis", C.POINTER(C.c_int32)), ("nbases", C.c_int32), ("bases", C.POINTER(Basis)),
                ("ndof_elem", C.c_int32), ("offsets", C.POINTER(C.c_int32)), ("max_card", C.c_int32), ("nqp", C.c_int32),
                ("qp_pts", C.POINTER(C.c_double)), ("qp_wts", C.POINTER(C.c_double))]


class TimeData(C.Structure):
    _fields_ = [("time", C.c_double), ("deltat", C.c_double), ("stage", C.c_int32), ("nstages", C.c_int32),
                ("butcher_A", C.POINTER(C.c_double)), ("butcher_b", C.POINTER(C.c_double)), ("butcher_c", C.POINTER(C.c_double)),
                ("nbdf", C.c_int32), ("bdf_wts", C.POINTER(C.c_double)),
                ("sol_prev", C.POINTER(C.c_void_p)), ("sol_stage", C.POINTER(C.c_void_p)), ("seed_what", C.c_int32), ("seed_index", C.c_int32)]


class BoundaryGroup(C.Structure):
    _fields_ = [("sideset", C.c_int32), ("local_side", C.c_int32), ("n_elem", C.c_int32), ("elem_ids", C.POINTER(C.c_int32)),
                ("nqp_side", C.c_int32), ("side_pts", C.POINTER(C.c_double)), ("side_wts", C.POINTER(C.c_double)),
                ("tangent_u", C.POINTER(C.c_double)), ("tangent_v", C.POINTER(C.c_double)), ("side_bases", C.POINTER(Basis))]


def lib():
    """Loads the shared library (fails loudly if it was not built: run __graft_entry__.build())."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("mrhyde_b200: %s is missing; build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        # test-only companion (host replay of the general kernel's stages for debug_emulate on host-only plans): registered when present
        emu = os.path.join(os.path.dirname(LIB_PATH), "libmrhyde_b200_emulate.so")
        if os.path.exists(emu):
            global _EMU
            _EMU = C.CDLL(emu)
            L.mrhyde_b200_debug_set_emulator.argtypes = [C.c_void_p]
            L.mrhyde_b200_debug_set_emulator(C.cast(_EMU.mrhyde_b200_emulator_lookup, C.c_void_p))
        L.mrhyde_b200_version.restype = C.c_char_p
        L.mrhyde_b200_last_error.restype = C.c_char_p
        L.mrhyde_b200_plan_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Desc), C.c_int]
        L.mrhyde_b200_plan_destroy.argtypes = [C.c_void_p]
        L.mrhyde_b200_plan_destroy.restype = None
        L.mrhyde_b200_plan_set_function.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.mrhyde_b200_plan_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.mrhyde_b200_plan_set_mesh.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mrhyde_b200_plan_set_mesh_indexed.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mrhyde_b200_plan_set_graph.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mrhyde_b200_plan_set_sidesets.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_char_p)]
        L.mrhyde_b200_plan_set_bc.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p]
        L.mrhyde_b200_plan_add_boundary_group.argtypes = [C.c_void_p, C.POINTER(BoundaryGroup)]
        L.mrhyde_b200_plan_finalize.argtypes = [C.c_void_p]
        L.mrhyde_b200_assemble_jacres.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(TimeData), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mrhyde_b200_assemble_jacres_adjoint.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(TimeData), C.c_void_p, C.c_void_p, C.c_void_p]
        L.mrhyde_b200_plan_warmup.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.mrhyde_b200_assemble_res.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(TimeData), C.c_void_p, C.c_void_p]
        L.mrhyde_b200_assemble_jacres_host.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(TimeData), C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.mrhyde_b200_comm_unique_id.argtypes = [C.c_void_p]
        L.mrhyde_b200_plan_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.mrhyde_b200_plan_set_halo.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.mrhyde_b200_halo_sum.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mrhyde_b200_plan_stat.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64)]
        L.mrhyde_b200_plan_kernel_time.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        L.mrhyde_b200_plan_eval_function.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_void_p, C.c_double, C.c_void_p]
        L.mrhyde_b200_expr_disassemble.argtypes = [C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_char_p, C.c_char_p, C.c_size_t]
        L.mrhyde_b200_expr_eval_host.argtypes = [C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_char_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.mrhyde_b200_plan_debug_scatter_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
        L.mrhyde_b200_plan_debug_jit.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t]
        L.mrhyde_b200_project_initial.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        L.mrhyde_b200_set_initial.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mrhyde_b200_plan_debug_emulate_initial.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
        L.mrhyde_b200_plan_debug_chain_rows.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.mrhyde_b200_plan_debug_stage_map.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.mrhyde_b200_plan_debug_metric_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.mrhyde_b200_plan_debug_class_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.mrhyde_b200_assemble_mass.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mrhyde_b200_plan_debug_emulate_mass.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.mrhyde_b200_apply_mass.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mrhyde_b200_plan_debug_emulate_apply_mass.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mrhyde_b200_plan_debug_emulate.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(TimeData), C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def _fn_arrays(functions):
    names = (C.c_char_p * max(1, len(functions)))(*[k.encode() for k in functions])
    exprs = (C.c_char_p * max(1, len(functions)))(*[str(v).encode() for v in functions.values()])
    return names, exprs


def expr_disassemble(functions, which):
    """Flattened device program of `which` out of a {name: expression} dict (host-side compile)."""
    L = lib()
    names, exprs = _fn_arrays(functions)
    buf = C.create_string_buffer(1 << 16)
    rc = L.mrhyde_b200_expr_disassemble(len(functions), names, exprs, which.encode(), buf, len(buf))
    if rc != 0:
        raise MrhydeB200Error(rc, L.mrhyde_b200_last_error().decode())
    return buf.value.decode()


def expr_eval_host(functions, which, vars7):
    """Runs the flattened program on the host at rows of (x, y, z, t, n[x], n[y], n[z])."""
    L = lib()
    names, exprs = _fn_arrays(functions)
    v = np.ascontiguousarray(vars7, dtype=np.float64).reshape(-1, 7)
    out = np.zeros(v.shape[0])
    rc = L.mrhyde_b200_expr_eval_host(len(functions), names, exprs, which.encode(), v.shape[0], C.c_void_p(v.ctypes.data), C.c_void_p(out.ctypes.data))
    if rc != 0:
        raise MrhydeB200Error(rc, L.mrhyde_b200_last_error().decode())
    return out


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ptr(x):
    """Device (torch tensor) or host (numpy) buffer -> raw address."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return C.c_void_p(x.ctypes.data)
    return C.c_void_p(x.data_ptr())


def _make_basis(b, keep):
    B = Basis()
    B.type = b["type"].encode()
    B.order = int(b.get("order", 1))
    B.card = int(b["card"])
    for k in ("val", "grad", "curl", "div"):
        a = b.get(k)
        if a is not None:
            a = np.ascontiguousarray(a, dtype=np.float64)
            keep.append(a)
            setattr(B, k, _dp(a))
    return B


class TimeSpec:
    """Python-side holder of mrhyde_b200_time (keeps the arrays alive)."""

    def __init__(self, time=0.0, deltat=1.0, stage=0, A=None, b=None, c=None, bdf=None, sol_prev=(), sol_stage=(), seed_what=1, seed_index=0):
        self.s = TimeData()
        self.s.seed_what, self.s.seed_index = seed_what, seed_index
        self.s.time = time
        self.s.deltat = deltat
        self.s.stage = stage
        self._keep = []
        if b is None:
            self.s.nstages = 0
            self.s.nbdf = 0
            return
        A = np.ascontiguousarray(A, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        c = np.ascontiguousarray(c, dtype=np.float64)
        bdf = np.ascontiguousarray(bdf, dtype=np.float64)
        self._keep += [A, b, c, bdf, list(sol_prev), list(sol_stage)]
        self.s.nstages = len(b)
        self.s.butcher_A, self.s.butcher_b, self.s.butcher_c = _dp(A), _dp(b), _dp(c)
        self.s.nbdf = len(bdf)
        self.s.bdf_wts = _dp(bdf)
        pp = (C.c_void_p * max(1, len(sol_prev)))(*[_ptr(v) for v in sol_prev])
        ps = (C.c_void_p * max(1, len(sol_stage)))(*[_ptr(v) for v in sol_stage])
        self._keep += [pp, ps]
        self.s.sol_prev = C.cast(pp, C.POINTER(C.c_void_p))
        self.s.sol_stage = C.cast(ps, C.POINTER(C.c_void_p))

    def ref(self):
        return C.byref(self.s)


class AssemblyPlan:
    """One block of one physics set behind the C ABI (mrhyde_b200_plan)."""

    def __init__(self, physics, dim, var_names, var_basis, bases, ndof_elem, offsets, qp_pts, qp_wts, device=0):
        self.L = lib()
        self._keep = []
        d = Desc()
        d.physics = physics.encode()
        d.dim = dim
        d.nvars = len(var_names)
        names = (C.c_char_p * len(var_names))(*[v.encode() for v in var_names])
        vb = np.ascontiguousarray(var_basis, dtype=np.int32)
        offs = np.ascontiguousarray(offsets, dtype=np.int32)
        qp = np.ascontiguousarray(qp_pts, dtype=np.float64)
        qw = np.ascontiguousarray(qp_wts, dtype=np.float64)
        barr = (Basis * len(bases))(*[_make_basis(b, self._keep) for b in bases])
        self._keep += [names, vb, offs, qp, qw, barr]
        d.var_names = C.cast(names, C.POINTER(C.c_char_p))
        d.var_basis = vb.ctypes.data_as(C.POINTER(C.c_int32))
        d.nbases = len(bases)
        d.bases = C.cast(barr, C.POINTER(Basis))
        d.ndof_elem = ndof_elem
        d.offsets = offs.ctypes.data_as(C.POINTER(C.c_int32))
        d.max_card = offs.shape[1] if offs.ndim == 2 else offs.size // max(1, len(var_names))
        d.nqp = len(qw)
        d.qp_pts = _dp(qp)
        d.qp_wts = _dp(qw)
        self.dim = dim
        self.ndof_elem = ndof_elem
        self.h = C.c_void_p()
        self._chk(self.L.mrhyde_b200_plan_create(C.byref(self.h), C.byref(d), device))
        self.n_rows = 0
        self.nnz = 0

    def _chk(self, rc):
        if rc != 0:
            raise MrhydeB200Error(rc, self.L.mrhyde_b200_last_error().decode())

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.mrhyde_b200_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- set-up ------------------------------------------------------------------------
    def set_function(self, name, expr):
        self._chk(self.L.mrhyde_b200_plan_set_function(self.h, name.encode(), str(expr).encode()))

    def set_option(self, key, value):
        if isinstance(value, bool):
            value = "true" if value else "false"
        self._chk(self.L.mrhyde_b200_plan_set_option(self.h, key.encode(), str(value).encode()))

    def set_mesh(self, elem_nodes, lids, orient_sign=None):
        en = np.ascontiguousarray(elem_nodes, dtype=np.float64)
        ld = np.ascontiguousarray(lids, dtype=np.int32)
        os_ = None if orient_sign is None else np.ascontiguousarray(orient_sign, dtype=np.int8)
        self._chk(self.L.mrhyde_b200_plan_set_mesh(self.h, en.shape[0], _ptr(en), _ptr(ld), _ptr(os_)))

    def set_mesh_indexed(self, vert_coords, conn, lids, orient_sign=None):
        vc = np.ascontiguousarray(vert_coords, dtype=np.float64)
        cn = np.ascontiguousarray(conn, dtype=np.int32)
        ld = np.ascontiguousarray(lids, dtype=np.int32)
        os_ = None if orient_sign is None else np.ascontiguousarray(orient_sign, dtype=np.int8)
        self._chk(self.L.mrhyde_b200_plan_set_mesh_indexed(self.h, vc.shape[0], _ptr(vc), cn.shape[0], _ptr(cn), _ptr(ld), _ptr(os_)))

    def set_graph(self, rowptr, colind, is_fixed=None, n_owned=None):
        rp = np.ascontiguousarray(rowptr, dtype=np.int64)
        ci = np.ascontiguousarray(colind, dtype=np.int32)
        fx = None if is_fixed is None else np.ascontiguousarray(is_fixed, dtype=np.uint8)
        self.n_rows = len(rp) - 1
        self.nnz = len(ci)
        self._chk(self.L.mrhyde_b200_plan_set_graph(self.h, self.n_rows, self.n_rows if n_owned is None else n_owned, _ptr(rp), _ptr(ci), _ptr(fx)))

    def set_sidesets(self, names):
        arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        self._chk(self.L.mrhyde_b200_plan_set_sidesets(self.h, len(names), C.cast(arr, C.POINTER(C.c_char_p))))

    def set_bc(self, var, side, bctype, expr="0.0"):
        self._chk(self.L.mrhyde_b200_plan_set_bc(self.h, var.encode(), side.encode(), bctype.encode(), str(expr).encode()))

    def add_boundary_group(self, sideset, local_side, elem_ids, side_pts, side_wts, tangent_u, tangent_v, side_bases):
        keep = []
        g = BoundaryGroup()
        ids = np.ascontiguousarray(elem_ids, dtype=np.int32)
        pts = np.ascontiguousarray(side_pts, dtype=np.float64)
        wts = np.ascontiguousarray(side_wts, dtype=np.float64)
        tu = np.ascontiguousarray(tangent_u, dtype=np.float64)
        tv = np.ascontiguousarray(tangent_v, dtype=np.float64)
        barr = (Basis * len(side_bases))(*[_make_basis(b, keep) for b in side_bases])
        g.sideset, g.local_side, g.n_elem = sideset, local_side, len(ids)
        g.elem_ids = ids.ctypes.data_as(C.POINTER(C.c_int32))
        g.nqp_side = len(wts)
        g.side_pts, g.side_wts, g.tangent_u, g.tangent_v = _dp(pts), _dp(wts), _dp(tu), _dp(tv)
        g.side_bases = C.cast(barr, C.POINTER(Basis))
        self._chk(self.L.mrhyde_b200_plan_add_boundary_group(self.h, C.byref(g)))

    def finalize(self):
        self._chk(self.L.mrhyde_b200_plan_finalize(self.h))

    # ---- hot path ----------------------------------------------------------------------
    def assemble_jacres(self, sol, res, jac, time=None, compute_jacobian=True, compute_residual=True, stream=None):
        self._chk(self.L.mrhyde_b200_assemble_jacres(self.h, _ptr(sol), time.ref() if time is not None else None,
                                                   int(compute_jacobian), int(compute_residual), _ptr(res), _ptr(jac),
                                                   C.c_void_p(stream) if stream else None))

    def assemble_res(self, sol, res, time=None, stream=None):
        self._chk(self.L.mrhyde_b200_assemble_res(self.h, _ptr(sol), time.ref() if time is not None else None, _ptr(res),
                                                C.c_void_p(stream) if stream else None))

    def assemble_jacres_adjoint(self, sol, res, jac, time=None, stream=None):
        """useadjoint = true: forward residual, transposed local Jacobians (device buffers)."""
        self._chk(self.L.mrhyde_b200_assemble_jacres_adjoint(self.h, _ptr(sol), time.ref() if time is not None else None, _ptr(res), _ptr(jac),
                                                              C.c_void_p(stream) if stream else None))

    def warmup(self, transient=False, compute_jacobian=True, compute_residual=True):
        """Builds the specialised kernel variant of an upcoming call now (no NVRTC compile inside the first assemble call)."""
        self._chk(self.L.mrhyde_b200_plan_warmup(self.h, int(transient), int(compute_jacobian), int(compute_residual)))

    def assemble_jacres_host(self, sol, res, jac, time=None, compute_jacobian=True, compute_residual=True):
        for a in (sol, res, jac):
            assert a is None or (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous) or hasattr(a, "data_ptr")
        self._chk(self.L.mrhyde_b200_assemble_jacres_host(self.h, _ptr(sol), time.ref() if time is not None else None,
                                                        int(compute_jacobian), int(compute_residual), _ptr(res), _ptr(jac)))

    def assemble_mass(self, mass_wts, mass_values, diag, lump=False, stream=None):
        """getWeightedMass: weighted mass values (graph order) and the Jacobi / lumped diagonal vector (device buffers)."""
        w = np.ascontiguousarray(mass_wts, dtype=np.float64)
        self._chk(self.L.mrhyde_b200_assemble_mass(self.h, _ptr(w), int(lump), _ptr(mass_values), _ptr(diag), C.c_void_p(stream) if stream else None))

    def apply_mass(self, mass_wts, x, y, stream=None):
        """applyMassMatrixFree: y (+)= M x (device buffers)."""
        w = np.ascontiguousarray(mass_wts, dtype=np.float64)
        self._chk(self.L.mrhyde_b200_apply_mass(self.h, _ptr(w), _ptr(x), _ptr(y), C.c_void_p(stream) if stream else None))

    def debug_emulate_apply_mass(self, mass_wts, x, y):
        w = np.ascontiguousarray(mass_wts, dtype=np.float64)
        self._chk(self.L.mrhyde_b200_plan_debug_emulate_apply_mass(self.h, _ptr(w), _ptr(x), _ptr(y)))

    def debug_emulate_mass(self, mass_wts, mass_values, diag, lump=False):
        w = np.ascontiguousarray(mass_wts, dtype=np.float64)
        self._chk(self.L.mrhyde_b200_plan_debug_emulate_mass(self.h, _ptr(w), int(lump), _ptr(mass_values), _ptr(diag)))

    # ---- multi-GPU -----------------------------------------------------------------------
    def comm_unique_id(self):
        buf = np.zeros(128, dtype=np.uint8)
        self._chk(self.L.mrhyde_b200_comm_unique_id(_ptr(buf)))
        return buf

    def comm_init(self, unique_id, rank, nranks):
        uid = np.ascontiguousarray(unique_id, dtype=np.uint8)
        self._chk(self.L.mrhyde_b200_plan_comm_init(self.h, _ptr(uid), rank, nranks))

    def set_halo(self, col_gids):
        """col_gids: global ids of the local column ids (rows first, then column-only ghosts)."""
        g = np.ascontiguousarray(col_gids, dtype=np.int64)
        self._chk(self.L.mrhyde_b200_plan_set_halo(self.h, len(g), _ptr(g)))

    def set_point_dofs(self, lids):
        """Point constraints: after every assembly the Jacobian rows of these local dofs are identity rows (the residual is left as assembled)."""
        d = np.ascontiguousarray(lids, dtype=np.int32)
        self._chk(self.L.mrhyde_b200_plan_set_point_dofs(self.h, len(d), _ptr(d)))

    def owned_extent(self):
        """(owned rows, non-zeros of the owned rows): the owned matrix after halo_sum is the prefix of the caller's CSR arrays."""
        a, b = C.c_int64(0), C.c_int64(0)
        self._chk(self.L.mrhyde_b200_plan_owned_extent(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def halo_sum(self, res, jac, stream=None):
        self._chk(self.L.mrhyde_b200_halo_sum(self.h, _ptr(res), _ptr(jac), C.c_void_p(stream) if stream else None))

    # ---- introspection ---------------------------------------------------------------
    def stat(self, key):
        v = C.c_int64(0)
        self._chk(self.L.mrhyde_b200_plan_stat(self.h, key.encode(), C.byref(v)))
        return v.value

    def kernel_time(self, reset=False):
        ms = C.c_double(0.0)
        n = C.c_int64(0)
        self._chk(self.L.mrhyde_b200_plan_kernel_time(self.h, int(reset), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def debug_scatter_host(self, stage, accumulate, res, jac):
        st = np.ascontiguousarray(stage, dtype=np.float64)
        self._chk(self.L.mrhyde_b200_plan_debug_scatter_host(self.h, _ptr(st), st.shape[1], int(accumulate), _ptr(res), _ptr(jac)))

    def debug_emulate(self, sol, res, jac, time=None, compute_jacobian=True, compute_residual=True):
        """Host replay of the general path's kernel stages (host-only plans, option kernel=general): a debugging aid,
        never an assembly path (mrhyde_b200_plan_debug_emulate)."""
        self._chk(self.L.mrhyde_b200_plan_debug_emulate(self.h, _ptr(sol), time.ref() if time is not None else None,
                                                      int(compute_jacobian), int(compute_residual), _ptr(res), _ptr(jac)))

    def project_initial(self, rhs, time=0.0, stream=0):
        """setInitial: rhs (+)= sum_q initial(x_q) phi_i w (device vector); functions "initial <var>[...]" come from set_function."""
        self._chk(self.L.mrhyde_b200_project_initial(self.h, float(time), _ptr(rhs), C.c_void_p(stream)))

    def set_initial(self, rhs, mass_values, time=0.0, stream=0):
        """setInitial as a whole: projection right-hand side, unit-weight mass matrix, the routine's own fix_zero_rows pass (device buffers)."""
        self._chk(self.L.mrhyde_b200_set_initial(self.h, float(time), _ptr(rhs), _ptr(mass_values), C.c_void_p(stream)))

    def debug_emulate_initial(self, rhs, time=0.0):
        """Host replay of project_initial on a host-only plan (debugging aid, never an assembly path)."""
        self._chk(self.L.mrhyde_b200_plan_debug_emulate_initial(self.h, float(time), _ptr(rhs)))

    def debug_chain_rows(self, chain_begin, chain_end, n_rows):
        """Boolean mask of the rows the sweep chains [chain_begin, chain_end) write."""
        mask = np.zeros(n_rows, dtype=np.uint8)
        self._chk(self.L.mrhyde_b200_plan_debug_chain_rows(self.h, int(chain_begin), int(chain_end), _ptr(mask)))
        return mask.astype(bool)

    def debug_stage_map(self, ndof):
        """(kmap[ndof, ndof], rmap[ndof]): where local-matrix entry (i, j) / residual entry i sit in the staged element vector."""
        kmap, rmap = np.zeros((ndof, ndof), dtype=np.int32), np.zeros(ndof, dtype=np.int32)
        self._chk(self.L.mrhyde_b200_plan_debug_stage_map(self.h, _ptr(kmap), _ptr(rmap)))
        return kmap, rmap

    def debug_metric_host(self, sol, accumulate, res, jac, time=None):
        """Host replay of the sweep kernel's metric ring (plans with stat("metric_ring") > 0): plan-analysis check, never an
        assembly path (mrhyde_b200_plan_debug_metric_host)."""
        self._chk(self.L.mrhyde_b200_plan_debug_metric_host(self.h, _ptr(sol), time.ref() if time is not None else None, int(accumulate), _ptr(res), _ptr(jac)))

    def debug_class_host(self, sol, accumulate, res, jac, time=None):
        """Host replay of the sweep kernel's class ring (plans with stat("class_ring") > 0): plan-analysis check, never an assembly path."""
        self._chk(self.L.mrhyde_b200_plan_debug_class_host(self.h, _ptr(sol), time.ref() if time is not None else None, int(accumulate), _ptr(res), _ptr(jac)))

    def debug_jit(self, source_path=None, cubin_path=None):
        """Generates + NVRTC-compiles the plan-specialised kernel (no device needed); returns the compiler log."""
        buf = C.create_string_buffer(1 << 16)
        self._chk(self.L.mrhyde_b200_plan_debug_jit(self.h, source_path.encode() if source_path else None,
                                                  cubin_path.encode() if cubin_path else None, buf, len(buf)))
        return buf.value.decode()

    def eval_function(self, name, xyz, time=0.0):
        p = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        out = np.zeros(p.shape[0])
        self._chk(self.L.mrhyde_b200_plan_eval_function(self.h, name.encode(), p.shape[0], _ptr(p), time, _ptr(out)))
        return out
