"""wavebem_b200 -- B200-native collocation-BEM hot path of mathLab/WaveBEM.

The product is `lib/libwbem.so` (hand-written sm_100a CUDA behind the C ABI of
include/wbem.h).  This package is the thin host side used by the tests and the benchmark:

* `lib()`        ctypes handle of libwbem.so -- raises if the library is not built; there is
                 NO CPU fallback and nothing here ever imports oracle/;
* `BEMProblem`   Python mirror of the reference's `BEMProblem<3>` interface
                 (include/bem_problem.h:87-180): reinit, assemble_system, compute_alpha, vmult,
                 compute_rhs, compute_constraints, assemble_preconditioner, solve_system, solve,
                 residual -- same names, argument meaning and error behaviour (a GMRES that
                 hits `Max steps` raises NoConvergence, like deal.II's SolverControl);
* `meshgen`      synthetic tank/Wigley/cube/sphere meshes (the reference needs OpenCASCADE);
* `constraints`  host restatement of compute_constraints (bem_problem.cc:990-1105), used by callers
                 that hand their own ConstraintMatrix lines over (the library has its own, csrc/constraints.cu).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import constraints as _constraints
from . import meshgen  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
# WBEM_LIB: development override (kernel tuning variants built by build.build_variant)
LIB_PATH = os.environ.get("WBEM_LIB") or os.path.join(_HERE, "lib", "libwbem.so")
_lib = None


class WbemError(RuntimeError):
    pass


class NoConvergence(WbemError):
    """GMRES reached `Max steps` (deal.II SolverControl::NoConvergence)."""

    def __init__(self, last_step, last_residual):
        super().__init__(f"GMRES did not converge: step {last_step}, residual {last_residual:g}")
        self.last_step = last_step
        self.last_residual = last_residual


class Params(C.Structure):
    _fields_ = [("quad_order", C.c_int), ("sing_order", C.c_int), ("gmres_tol", C.c_double),
                ("gmres_max_steps", C.c_int), ("gmres_n_tmp_vectors", C.c_int),
                ("preconditioner_band", C.c_int), ("device", C.c_int), ("rank", C.c_int),
                ("world_size", C.c_int), ("assemble_variant", C.c_int), ("precond_on_host", C.c_int),
                ("precond_kind", C.c_int), ("auto_constraints", C.c_int), ("n_gpus", C.c_int),
                ("fused_gather_on_shared_device", C.c_int), ("reserved", C.c_int), ("devices", C.c_int * 16)]


class Timings(C.Structure):
    _fields_ = [("geometry_ms", C.c_double), ("assemble_regular_ms", C.c_double),
                ("assemble_singular_ms", C.c_double), ("alpha_ms", C.c_double),
                ("assemble_total_ms", C.c_double), ("rhs_ms", C.c_double),
                ("precond_setup_ms", C.c_double), ("gmres_ms", C.c_double),
                ("gemv_ms_sum", C.c_double), ("precond_apply_ms_sum", C.c_double),
                ("allgather_ms_sum", C.c_double), ("solve_system_total_ms", C.c_double),
                ("gemv_calls", C.c_int), ("gmres_iters", C.c_int), ("kernel_launches", C.c_longlong),
                ("gemv_bytes_last", C.c_double), ("constraints_ms", C.c_double),
                ("reserved", C.c_double * 5)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


# every symbol include/wbem.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "wbem_default_params", "wbem_create", "wbem_destroy", "wbem_last_error", "wbem_version",
    "wbem_set_topology", "wbem_set_geometry", "wbem_set_geometry_dev", "wbem_assemble",
    "wbem_compute_alpha", "wbem_get_alpha", "wbem_get_rows", "wbem_row_block", "wbem_set_masks",
    "wbem_set_constraints", "wbem_vmult", "wbem_constrained_vmult", "wbem_distribute_rhs",
    "wbem_compute_rhs", "wbem_assemble_preconditioner", "wbem_precond_vmult", "wbem_get_band",
    "wbem_solve_system", "wbem_solve", "wbem_solve_dev", "wbem_solve_system_dev", "wbem_residual",
    "wbem_get_system_rhs", "wbem_get_sol", "wbem_get_timings", "wbem_reset_counters",
    "wbem_comm_unique_id", "wbem_comm_init", "wbem_measure_fp64_peak", "wbem_measure_copy_bw",
    "wbem_time_operator", "wbem_time_assemble", "wbem_selftest_rsqrt", "wbem_plan_check", "wbem_stream_order_check",
    "wbem_timer_start", "wbem_timer_stop", "wbem_comm_ipc_export", "wbem_comm_ipc_import",
    "wbem_comm_ipc_close", "wbem_get_spai", "wbem_spai_pattern_check", "wbem_set_precond_kind",
    "wbem_compute_normals", "wbem_compute_surface_gradients", "wbem_set_hanging_constraints",
    "wbem_compute_constraints", "wbem_get_constraints", "wbem_mass_cg_iterations", "wbem_gmres",
    "wbem_set_fevalues", "wbem_generate_double_nodes_set", "wbem_internal_velocities", "wbem_pressure_force",
    "wbem_solve_system_multi", "wbem_constrained_vmult_multi", "wbem_compute_rhs_multi",
]


def lib():
    """Load libwbem.so.  Fails loudly when it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise WbemError(f"{LIB_PATH} is missing: run `python -m wavebem_b200.build` "
                            "(__graft_entry__.build()).  There is no CPU fallback.")
        _lib = C.CDLL(LIB_PATH)
        _lib.wbem_last_error.restype = C.c_char_p
        _lib.wbem_last_error.argtypes = [C.c_void_p]
    return _lib


def generate_double_nodes_set(xyz, boundary_dofs=None, tol=1e-8):
    """ComputationalDomain::generate_double_nodes_set (computational_domain.cc:258-307) through the
    library's host helper: returns the CSR (dn_ptr, dn_idx) that Context.set_topology takes."""
    xyz = _f64(xyz)
    n = xyz.shape[0]
    b = None if boundary_dofs is None else np.ascontiguousarray(boundary_dofs, dtype=np.uint8)
    ptr = np.zeros(n + 1, dtype=np.uint32)
    needed = C.c_uint64(0)
    f = lib().wbem_generate_double_nodes_set
    f.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_uint64,
                  C.POINTER(C.c_uint64)]
    rc = f(n, _dp(xyz), None if b is None else _dp(b), float(tol), _dp(ptr), None, 0, C.byref(needed))
    if rc < 0:
        raise WbemError("wbem_generate_double_nodes_set: bad argument")
    idx = np.zeros(needed.value, dtype=np.uint32)
    rc = f(n, _dp(xyz), None if b is None else _dp(b), float(tol), _dp(ptr), _dp(idx), idx.size, C.byref(needed))
    if rc != 0:
        raise WbemError("wbem_generate_double_nodes_set failed")
    return ptr, idx


def default_params(**kw) -> Params:
    p = Params()
    lib().wbem_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError(f"unknown wbem_params field {k}")
        if k == "devices":   # device ordinal of every row block of a single-process context (n_gpus > 1)
            v = list(v)
            v = (C.c_int * 16)(*(v + [-1] * (16 - len(v))))
        setattr(p, k, v)
    return p


def _dp(a):
    return a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Context:
    """Owns one wbem_ctx: one GPU and one block of matrix rows -- or, with n_gpus=P (and optionally
    devices=[...]), ALL P row blocks of a single-process multi-GPU context."""

    def __init__(self, params: Params | None = None, **kw):
        self._h = C.c_void_p()
        self.params = params if params is not None else default_params(**kw)
        rc = lib().wbem_create(C.byref(self.params), C.byref(self._h))
        if rc != 0:
            msg = lib().wbem_last_error(None).decode()
            self._h = C.c_void_p()
            raise WbemError(f"wbem_create failed ({rc}): {msg}")
        self.n = 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().wbem_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc, allow_positive=False):
        if rc < 0 or (rc > 0 and not allow_positive):
            raise WbemError(f"libwbem error {rc}: {lib().wbem_last_error(self._h).decode()}")
        return rc

    # --- flattening / uploads ---
    def set_topology(self, n_dofs, cells, dir_flag, dn_ptr, dn_idx):
        cells = np.ascontiguousarray(cells, dtype=np.uint32)
        dir_flag = np.ascontiguousarray(dir_flag, dtype=np.uint8)
        dn_ptr = np.ascontiguousarray(dn_ptr, dtype=np.uint32)
        dn_idx = np.ascontiguousarray(dn_idx, dtype=np.uint32)
        assert cells.ndim == 2 and cells.shape[1] == 4 and len(dn_ptr) == n_dofs + 1
        self._chk(lib().wbem_set_topology(self._h, C.c_uint32(n_dofs), C.c_uint32(cells.shape[0]),
                                          _dp(cells), _dp(dir_flag), _dp(dn_ptr), _dp(dn_idx)))
        self.n = int(n_dofs)
        r0, r1 = C.c_uint32(), C.c_uint32()
        lib().wbem_row_block(self._h, C.byref(r0), C.byref(r1))
        self.row0, self.row1 = r0.value, r1.value

    def set_geometry(self, xyz):
        xyz = _f64(xyz)
        assert xyz.shape == (self.n, 3)
        self._chk(lib().wbem_set_geometry(self._h, _dp(xyz)))

    def set_masks(self, surface_nodes, other_nodes):
        s, o = _f64(surface_nodes), _f64(other_nodes)
        assert s.shape == (self.n,) and o.shape == (self.n,)
        self._chk(lib().wbem_set_masks(self._h, _dp(s), _dp(o)))

    def set_constraints(self, cl):
        ptr = cl.ptr if cl.n_lines else np.zeros(1, dtype=np.uint32)
        self._chk(lib().wbem_set_constraints(self._h, C.c_uint32(cl.n_lines), _dp(cl.lines), _dp(ptr),
                                             _dp(cl.col), _dp(cl.val), _dp(cl.inhom)))

    # --- compute ---
    def set_fevalues(self, q_points, normals, jxw):
        self._chk(lib().wbem_set_fevalues(self._h, _dp(_f64(q_points)), _dp(_f64(normals)), _dp(_f64(jxw))))

    def gmres(self, rhs, raise_on_no_convergence=True):
        """solver.solve(cc, sol, system_rhs, preconditioner) for a prepared right-hand side."""
        sol = np.empty(self.n)
        it, res = C.c_int(0), C.c_double(0)
        rc = self._chk(lib().wbem_gmres(self._h, _dp(_f64(rhs)), _dp(sol), C.byref(it), C.byref(res)),
                       allow_positive=True)
        if rc > 0 and raise_on_no_convergence:
            raise NoConvergence(it.value, res.value)
        return sol, it.value, res.value

    def assemble(self):
        self._chk(lib().wbem_assemble(self._h))

    def compute_alpha(self):
        self._chk(lib().wbem_compute_alpha(self._h))

    def get_alpha(self):
        a = np.empty(self.n)
        self._chk(lib().wbem_get_alpha(self._h, _dp(a)))
        return a

    def get_rows(self, which, r0=None, r1=None):
        r0 = self.row0 if r0 is None else r0
        r1 = self.row1 if r1 is None else r1
        out = np.empty((r1 - r0, self.n))
        self._chk(lib().wbem_get_rows(self._h, int(which), C.c_uint32(r0), C.c_uint32(r1), _dp(out)))
        return out

    def _apply(self, fn, src):
        src = _f64(src)
        assert src.shape == (self.n,)
        dst = np.empty(self.n)
        self._chk(fn(self._h, _dp(dst), _dp(src)))
        return dst

    def vmult(self, src):
        return self._apply(lib().wbem_vmult, src)

    def constrained_vmult(self, src):
        return self._apply(lib().wbem_constrained_vmult, src)

    def constrained_vmult_multi(self, src, fn="wbem_constrained_vmult_multi"):
        src = _f64(src).reshape(-1, self.n)
        dst = np.empty_like(src)
        self._chk(getattr(lib(), fn)(self._h, C.c_int(src.shape[0]), _dp(dst), _dp(src)))
        return dst

    def compute_rhs_multi(self, src):
        return self.constrained_vmult_multi(src, "wbem_compute_rhs_multi")

    def compute_rhs(self, src):
        return self._apply(lib().wbem_compute_rhs, src)

    def precond_vmult(self, src):
        return self._apply(lib().wbem_precond_vmult, src)

    def distribute_rhs(self, rhs):
        rhs = np.array(rhs, dtype=np.float64)
        self._chk(lib().wbem_distribute_rhs(self._h, _dp(rhs)))
        return rhs

    def assemble_preconditioner(self):
        self._chk(lib().wbem_assemble_preconditioner(self._h))

    def get_band(self):
        band = self.params.preconditioner_band
        out = np.empty((self.row1 - self.row0, band))
        self._chk(lib().wbem_get_band(self._h, _dp(out)))
        return out

    def compute_normals(self):
        out = np.empty((self.n, 3))
        self._chk(lib().wbem_compute_normals(self._h, _dp(out)))
        return out

    def compute_surface_gradients(self, tmp_rhs):
        out = np.empty((self.n, 3))
        self._chk(lib().wbem_compute_surface_gradients(self._h, _dp(_f64(tmp_rhs)), _dp(out)))
        return out

    def set_hanging_constraints(self, hanging):
        """hanging: list of (dof, [(master, weight), ...]) -- DoFTools::make_hanging_node_constraints"""
        lines = np.array([d for d, _ in hanging], dtype=np.uint32)
        ptr = np.zeros(len(hanging) + 1, dtype=np.uint32)
        col, val = [], []
        for k, (_, ent) in enumerate(hanging):
            for m, w in ent:
                col.append(m)
                val.append(w)
            ptr[k + 1] = len(col)
        col = np.array(col, dtype=np.uint32)
        val = np.array(val, dtype=np.float64)
        self._chk(lib().wbem_set_hanging_constraints(self._h, C.c_uint32(len(lines)), _dp(lines), _dp(ptr), _dp(col),
                                                     _dp(val)))

    def compute_constraints(self, tmp_rhs):
        """BEMProblem<3>::compute_constraints inside the library; returns the installed lines."""
        self._chk(lib().wbem_compute_constraints(self._h, _dp(_f64(tmp_rhs))))
        return self.get_constraints()

    def get_constraints(self):
        nl, nnz = C.c_uint32(0), C.c_uint32(0)
        self._chk(lib().wbem_get_constraints(self._h, C.byref(nl), C.byref(nnz), None, None, None, None, None))
        lines = np.empty(nl.value, dtype=np.uint32)
        ptr = np.empty(nl.value + 1, dtype=np.uint32)
        col = np.empty(nnz.value, dtype=np.uint32)
        val = np.empty(nnz.value)
        inhom = np.empty(nl.value)
        self._chk(lib().wbem_get_constraints(self._h, None, None, _dp(lines), _dp(ptr), _dp(col), _dp(val), _dp(inhom)))
        return _constraints.ConstraintLines(self.n, lines, ptr, col, val, inhom)

    def mass_cg_iterations(self):
        return lib().wbem_mass_cg_iterations(self._h)

    def set_precond_kind(self, kind):
        self._chk(lib().wbem_set_precond_kind(self._h, C.c_int(int(kind))))
        self.params.precond_kind = int(kind)

    def get_spai(self):
        """precond_kind = 1: (nbr[N,k] uint32 with 0xffffffff pads, val[N,k], n_singular)."""
        k = C.c_uint32(0)
        self._chk(lib().wbem_get_spai(self._h, C.byref(k), None, None, None))
        nbr = np.empty((self.n, k.value), dtype=np.uint32)
        val = np.empty((self.n, k.value))
        ns = C.c_int(0)
        self._chk(lib().wbem_get_spai(self._h, C.byref(k), _dp(nbr), _dp(val), C.byref(ns)))
        return nbr, val, ns.value

    def solve_system(self, phi, dphi_dn, tmp_rhs, raise_on_no_convergence=True):
        phi, dphi_dn, tmp_rhs = np.array(phi, dtype=np.float64), np.array(dphi_dn, dtype=np.float64), _f64(tmp_rhs)
        it, res = C.c_int(0), C.c_double(0)
        rc = self._chk(lib().wbem_solve_system(self._h, _dp(phi), _dp(dphi_dn), _dp(tmp_rhs), C.byref(it),
                                               C.byref(res)), allow_positive=True)
        if rc > 0 and raise_on_no_convergence:
            raise NoConvergence(it.value, res.value)
        return phi, dphi_dn, it.value, res.value

    def solve_system_multi(self, phi, dphi_dn, tmp_rhs, raise_on_no_convergence=True):
        """nrhs solve_system calls on the same matrices (the J.v pattern) sharing each pass over them.
        tmp_rhs: (nrhs, N); phi / dphi_dn: (N,) broadcast or (nrhs, N).  Returns phi, dphi_dn, iters, residuals."""
        tmp_rhs = _f64(tmp_rhs).reshape(-1, self.n)
        nrhs = tmp_rhs.shape[0]
        # order="C": the default order="K" would copy a broadcast view into an interleaved layout
        phi = np.array(np.broadcast_to(_f64(phi), (nrhs, self.n)), dtype=np.float64, order="C")
        dphi_dn = np.array(np.broadcast_to(_f64(dphi_dn), (nrhs, self.n)), dtype=np.float64, order="C")
        it = np.zeros(nrhs, dtype=np.int32)
        res = np.zeros(nrhs)
        rc = self._chk(lib().wbem_solve_system_multi(self._h, C.c_int(nrhs), _dp(phi), _dp(dphi_dn), _dp(tmp_rhs),
                                                     _dp(it), _dp(res)), allow_positive=True)
        if rc > 0 and raise_on_no_convergence:
            k = int(np.argmax(res))
            raise NoConvergence(int(it[k]), float(res[k]))
        return phi, dphi_dn, it, res

    def solve(self, xyz, phi, dphi_dn, tmp_rhs, raise_on_no_convergence=True):
        xyz = _f64(xyz)
        phi, dphi_dn, tmp_rhs = np.array(phi, dtype=np.float64), np.array(dphi_dn, dtype=np.float64), _f64(tmp_rhs)
        it, res = C.c_int(0), C.c_double(0)
        rc = self._chk(lib().wbem_solve(self._h, _dp(xyz), _dp(phi), _dp(dphi_dn), _dp(tmp_rhs),
                                        C.byref(it), C.byref(res)), allow_positive=True)
        if rc > 0 and raise_on_no_convergence:
            raise NoConvergence(it.value, res.value)
        return phi, dphi_dn, it.value, res.value

    def residual(self, phi, dphi_dn):
        phi, dphi_dn = _f64(phi), _f64(dphi_dn)
        res = np.empty(self.n)
        self._chk(lib().wbem_residual(self._h, _dp(res), _dp(phi), _dp(dphi_dn)))
        return res

    def internal_velocities(self, phi, dphi_dn, points):
        """FreeSurface<3>::compute_internal_velocities (free_surface.cc:10426-10537) on the device."""
        pts = _f64(points).reshape(-1, 3)
        out = np.empty_like(pts)
        self._chk(lib().wbem_internal_velocities(self._h, _dp(_f64(phi)), _dp(_f64(dphi_dn)), C.c_uint32(len(pts)),
                                                 _dp(pts), _dp(out)))
        return out

    def pressure_force(self, phi, dphi_dn, cell_marked, vinf, rho=1025.1, g=9.81, baricenter=(0.0, 0.0, 0.0)):
        """The hull integrals of compute_pressure (free_surface.cc:9534-9598): 11 numbers, see wbem.h."""
        out = np.empty(11)
        mk = np.ascontiguousarray(cell_marked, dtype=np.uint8)
        self._chk(lib().wbem_pressure_force(self._h, _dp(_f64(phi)), _dp(_f64(dphi_dn)), _dp(mk), _dp(_f64(vinf)),
                                            C.c_double(rho), C.c_double(g), _dp(_f64(baricenter)), _dp(out)))
        return out

    def get_system_rhs(self):
        a = np.empty(self.n)
        self._chk(lib().wbem_get_system_rhs(self._h, _dp(a)))
        return a

    def get_sol(self):
        a = np.empty(self.n)
        self._chk(lib().wbem_get_sol(self._h, _dp(a)))
        return a

    def timings(self) -> dict:
        t = Timings()
        self._chk(lib().wbem_get_timings(self._h, C.byref(t)))
        return t.as_dict()

    def timer_start(self):
        self._chk(lib().wbem_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_double(0)
        self._chk(lib().wbem_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def reset_counters(self):
        self._chk(lib().wbem_reset_counters(self._h))

    # --- device-pointer entry points (bench `value` leg) ---
    def solve_dev(self, d_xyz, d_phi, d_dphi, d_bc):
        it, res = C.c_int(0), C.c_double(0)
        rc = self._chk(lib().wbem_solve_dev(self._h, C.c_void_p(d_xyz), C.c_void_p(d_phi), C.c_void_p(d_dphi),
                                            C.c_void_p(d_bc), C.byref(it), C.byref(res)), allow_positive=True)
        return rc, it.value, res.value

    def solve_system_dev(self, d_phi, d_dphi, d_bc):
        it, res = C.c_int(0), C.c_double(0)
        rc = self._chk(lib().wbem_solve_system_dev(self._h, C.c_void_p(d_phi), C.c_void_p(d_dphi),
                                                   C.c_void_p(d_bc), C.byref(it), C.byref(res)),
                       allow_positive=True)
        return rc, it.value, res.value

    # --- multi-GPU ---
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        rc = lib().wbem_comm_unique_id(buf)
        if rc:
            raise WbemError("wbem_comm_unique_id failed: " + lib().wbem_last_error(None).decode())
        return buf.raw

    def comm_init(self, uid: bytes):
        assert len(uid) == 128
        self._chk(lib().wbem_comm_init(self._h, C.c_char_p(uid)))

    def ipc_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        self._chk(lib().wbem_comm_ipc_export(self._h, buf))
        return buf.raw

    def ipc_import(self, handles: bytes):
        assert len(handles) == 64 * self.params.world_size
        self._chk(lib().wbem_comm_ipc_import(self._h, C.c_char_p(handles)))

    def ipc_close(self):
        self._chk(lib().wbem_comm_ipc_close(self._h))

    # --- diagnostics ---
    def measure_fp64_peak(self):
        v = C.c_double(0)
        self._chk(lib().wbem_measure_fp64_peak(self._h, C.byref(v)))
        return v.value

    def measure_copy_bw(self):
        v = C.c_double(0)
        self._chk(lib().wbem_measure_copy_bw(self._h, C.byref(v)))
        return v.value

    def time_operator(self, reps=5, flush_l2=True):
        ms, by = C.c_double(0), C.c_double(0)
        self._chk(lib().wbem_time_operator(self._h, int(reps), int(flush_l2), C.byref(ms), C.byref(by)))
        return ms.value, by.value

    def time_assemble(self, reps=3):
        ms = C.c_double(0)
        self._chk(lib().wbem_time_assemble(self._h, int(reps), C.byref(ms)))
        return ms.value

    def selftest_rsqrt(self, x):
        x = _f64(x)
        out = np.empty_like(x)
        f = lib().wbem_selftest_rsqrt
        self._chk(f(self._h, _dp(x), _dp(out), C.c_int(x.size)))
        return out


class BEMProblem:
    """Mirror of the reference's BEMProblem<3> (include/bem_problem.h:87-180) over the C ABI.

    `comp_dom` plays ComputationalDomain<3>: it must expose xyz (support points), cells,
    dir_flag, dn_ptr/dn_idx (double_nodes_set), surface_nodes, other_nodes; optionally `hanging`
    (make_hanging_node_constraints lines).  compute_constraints runs inside the library
    (compute_normals + compute_surface_gradients on the GPU, like the reference's solve_system,
    bem_problem.cc:845); a domain that carries its own `nodes_normals` gets the host restatement
    with those normals instead.
    """

    def __init__(self, comp_dom, **params):
        self.comp_dom = comp_dom
        self._host_constraints = getattr(comp_dom, "nodes_normals", None) is not None
        if not self._host_constraints:
            params.setdefault("auto_constraints", 1)
        self.ctx = Context(**params)
        self.constraints = None
        self.alpha = None
        self.system_rhs = None
        self.sol = None
        self.last_step = 0
        self.last_residual = 0.0

    # BEMProblem::reinit (source/bem_problem.cc:55-71)
    def reinit(self):
        d = self.comp_dom
        self.ctx.set_topology(d.xyz.shape[0], d.cells, d.dir_flag, d.dn_ptr, d.dn_idx)
        if getattr(d, "hanging", None) and not self._host_constraints:
            self.ctx.set_hanging_constraints(d.hanging)
        self.constraints = None
        self._have_geometry = False

    # BEMProblem::assemble_system (source/bem_problem.cc:106-590)
    def assemble_system(self):
        self.ctx.set_geometry(self.comp_dom.xyz)
        self._have_geometry = True
        self.ctx.assemble()
        self.alpha = self.ctx.get_alpha()

    # BEMProblem::compute_alpha (source/bem_problem.cc:594-618)
    def compute_alpha(self):
        self.ctx.compute_alpha()
        self.alpha = self.ctx.get_alpha()

    def _masks(self):
        self.ctx.set_masks(self.comp_dom.surface_nodes, self.comp_dom.other_nodes)

    # BEMProblem::vmult / compute_rhs (source/bem_problem.cc:620-707): dst is filled in place
    def vmult(self, dst, src):
        self._masks()
        dst[:] = self.ctx.vmult(src)

    def compute_rhs(self, dst, src):
        self._masks()
        dst[:] = self.ctx.compute_rhs(src)

    # BEMProblem::compute_constraints (source/bem_problem.cc:990-1105)
    def compute_constraints(self, tmp_rhs):
        d = self.comp_dom
        if self._host_constraints:
            self.constraints = _constraints.compute_constraints(
                d.dn_ptr, d.dn_idx, d.surface_nodes, tmp_rhs,
                nodes_normals=getattr(d, "nodes_normals", None),
                node_surface_gradients=getattr(d, "node_surface_gradients", None),
                hanging=getattr(d, "hanging", None))
            self.ctx.set_constraints(self.constraints)
        else:
            self._masks()
            if not self._have_geometry:
                self.ctx.set_geometry(d.xyz)
                self._have_geometry = True
            self.constraints = self.ctx.compute_constraints(tmp_rhs)
        return self.constraints

    # ComputationalDomain::compute_normals / BEMProblem::compute_surface_gradients
    def compute_normals(self):
        if not self._have_geometry:
            self.ctx.set_geometry(self.comp_dom.xyz)
            self._have_geometry = True
        return self.ctx.compute_normals()

    def compute_surface_gradients(self, tmp_rhs):
        self._masks()
        if not self._have_geometry:
            self.ctx.set_geometry(self.comp_dom.xyz)
            self._have_geometry = True
        return self.ctx.compute_surface_gradients(tmp_rhs)

    # BEMProblem::assemble_preconditioner (source/bem_problem.cc:1107-1149)
    def assemble_preconditioner(self):
        self._masks()
        self.ctx.assemble_preconditioner()

    # BEMProblem::solve_system (source/bem_problem.cc:821-895): phi / dphi_dn updated in place
    def solve_system(self, phi, dphi_dn, tmp_rhs):
        self._masks()
        if self._host_constraints:
            self.compute_constraints(tmp_rhs)
        p, d, it, res = self.ctx.solve_system(phi, dphi_dn, tmp_rhs, raise_on_no_convergence=False)
        self._finish(phi, dphi_dn, p, d, it, res)

    # BEMProblem::solve (source/bem_problem.cc:969-987)
    def solve(self, phi, dphi_dn, tmp_rhs):
        self._masks()
        if self._host_constraints:
            self.compute_constraints(tmp_rhs)
        p, d, it, res = self.ctx.solve(self.comp_dom.xyz, phi, dphi_dn, tmp_rhs, raise_on_no_convergence=False)
        self._have_geometry = True
        self._finish(phi, dphi_dn, p, d, it, res)

    def _finish(self, phi, dphi_dn, p, d, it, res):
        phi[:] = p
        dphi_dn[:] = d
        self.last_step, self.last_residual = it, res
        if not self._host_constraints:
            self.constraints = self.ctx.get_constraints()   # what solve_system's compute_constraints built
        self.alpha = self.ctx.get_alpha()
        self.system_rhs = self.ctx.get_system_rhs()
        self.sol = self.ctx.get_sol()
        tol = self.ctx.params.gmres_tol
        if not (res <= tol):
            raise NoConvergence(it, res)

    # BEMProblem::residual (source/bem_problem.cc:903-961)
    def residual(self, res, phi, dphi_dn):
        self._masks()
        d = self.comp_dom
        tmp = np.asarray(dphi_dn) * d.other_nodes + np.asarray(phi) * d.surface_nodes
        if self._host_constraints:
            self.compute_constraints(tmp)
        res[:] = self.ctx.residual(phi, dphi_dn)

    # public members of the reference class (include/bem_problem.h:153-154)
    def neumann_matrix(self, r0=None, r1=None):
        return self.ctx.get_rows(0, r0, r1)

    def dirichlet_matrix(self, r0=None, r1=None):
        return self.ctx.get_rows(1, r0, r1)
