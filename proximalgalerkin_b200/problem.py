"""Problem-setup and solve API of the reference's ``lvpp`` package on top of liblvpp_b200.so.

Mirrors src/lvpp/problem.py of the reference:

* :class:`SNESProblem` (``problem.py:14-77``) -- ``.L``, ``.a``, ``.bcs``, ``.u``; ``F(snes, x, F)``
  assembles the residual with lifting and ``set_bc`` into ``F``; ``J(snes, x, J, P)`` assembles the
  Jacobian into ``J``.
* :class:`SNESSolver` (``problem.py:80-127``) -- ``solve() -> (converged_reason, iterations)``; the
  caller's function is overwritten only on convergence (``:121-123``).
* :class:`NonlinearProblem` -- the ``dolfinx.fem.petsc.NonlinearProblem`` call shape every example
  uses (examples/01_obstacle_problem/obstacle_pg.py:140-142,190-192).

UFL is not available, so the "forms" are descriptor objects for the obstacle formulation
(:func:`obstacle_residual`, :func:`derivative`).  All arithmetic happens in the CUDA library; PyTorch
only owns device buffers.  There is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _capi
from .fem import Constant, Function, FunctionSpace, QuadratureFunction

# PETSc SNESConvergedReason / KSPConvergedReason values used by the examples
SNES_CONVERGED_FNORM_ABS = 2
SNES_CONVERGED_FNORM_RELATIVE = 3
SNES_CONVERGED_SNORM_RELATIVE = 4
SNES_DIVERGED_LINEAR_SOLVE = -3
SNES_DIVERGED_FNORM_NAN = -4
SNES_DIVERGED_MAX_IT = -5
SNES_DIVERGED_DTOL = -9


def _torch():
    import torch

    return torch


# ---------------------------------------------------------------------------------------------
# forms
class ObstacleResidual:
    """The residual form of obstacle_pg.py:116-124,

        alpha (grad u, grad v) + (psi, v) + (u, w) - (exp(psi), w) - (phi, w) - alpha (f, v) - (psi_k, v),

    as a descriptor: the unknown ``sol``, the previous proximal iterate ``sol_k``, the ``Constant``\\ s
    ``alpha`` and ``f`` and the quadrature-space obstacle ``phi``."""

    def __init__(self, sol: Function, sol_k: Function, alpha: Constant, f: Constant, phi: QuadratureFunction):
        if sol.function_space.mesh is not sol_k.function_space.mesh:
            raise ValueError("sol and sol_k live on different meshes")
        self.sol, self.sol_k, self.alpha, self.f, self.phi = sol, sol_k, alpha, f, phi
        self.function_space = sol.function_space


class ObstacleJacobian:
    """``ufl.derivative(F, sol)`` (obstacle_pg.py:125): [[alpha K, M], [M, -D(psi)]]."""

    def __init__(self, F: ObstacleResidual):
        self.F = F
        self.function_space = F.function_space


def obstacle_residual(sol, sol_k, alpha, f, phi):
    return ObstacleResidual(sol, sol_k, alpha, f, phi)


def derivative(F, u=None, du=None):
    if not isinstance(F, ObstacleResidual):
        raise TypeError("derivative() is defined for ObstacleResidual forms")
    if u is not None and u is not F.sol:
        raise ValueError("derivative must be taken with respect to the form's unknown")
    return ObstacleJacobian(F)


# ---------------------------------------------------------------------------------------------
# device objects
class DeviceVector:
    """A mixed vector in HBM (fp64, torch-owned)."""

    def __init__(self, n, device):
        torch = _torch()
        self.tensor = torch.zeros(n, dtype=torch.float64, device=device)

    @property
    def ptr(self):
        return self.tensor.data_ptr()

    def numpy(self):
        return self.tensor.cpu().numpy()

    def set(self, host_array):
        torch = _torch()
        self.tensor.copy_(torch.from_numpy(np.ascontiguousarray(host_array, dtype=np.float64)))
        torch.cuda.current_stream().synchronize()


class DeviceMatrix:
    """The Jacobian held by the library (sliced-ELL block storage); CSR views for inspection."""

    def __init__(self, dev):
        self._dev = dev

    def pattern(self):
        return self._dev.csr_pattern()

    def values(self):
        return self._dev.jacobian_values()

    def mult(self, x: DeviceVector, y: DeviceVector):
        self._dev.spmv(x, y)

    def zeroEntries(self):  # storage is overwritten by assembly
        return None

    def assemble(self):
        return None


class DeviceProblem:
    """Owns the library handle for one (space, bcs, obstacle) triple on the current CUDA device."""

    def __init__(self, F: ObstacleResidual, bcs, device=None):
        torch = _torch()
        self.lib = _capi.load()
        if not torch.cuda.is_available():
            raise _capi.LvppError(_capi.E_NOGPU, "no CUDA device: the LVPP path has no CPU fallback")
        self.F = F
        V: FunctionSpace = F.function_space
        self.V = V
        mesh = V.mesh
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        torch.cuda.set_device(self.device)
        d = _capi.ObstacleDesc()
        keep = []  # arrays that must outlive lvpp_create

        def arr(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return a

        d.tdim, d.nld = mesh.tdim, V.nld
        d.num_nodes, d.num_owned = V.num_nodes, V.num_owned_nodes
        d.num_cells, d.num_owned_cells = mesh.num_cells, mesh.num_owned_cells
        d.node_coords = _capi.as_ptr(arr(V.node_coords, np.float64), C.c_double)
        d.cell_nodes = _capi.as_ptr(arr(V.cell_nodes, np.int32), C.c_int32)
        d.nq = V.qweights.size
        d.qweights = _capi.as_ptr(arr(V.qweights, np.float64), C.c_double)
        d.phi_tab = _capi.as_ptr(arr(V.phi_tab, np.float64), C.c_double)
        d.dphi_tab = _capi.as_ptr(arr(V.dphi_tab, np.float64), C.c_double)
        d.qpoints = _capi.as_ptr(arr(V.qpoints, np.float64), C.c_double)
        bc_nodes = np.zeros(0, dtype=np.int32)
        bc_vals = np.zeros(0)
        self._bcs = list(bcs or [])
        for bc in bcs or []:
            bc_nodes = np.concatenate([bc_nodes, bc.nodes])
            bc_vals = np.concatenate([bc_vals, bc.values])
        d.num_bc = bc_nodes.size
        if bc_nodes.size:
            d.bc_nodes = _capi.as_ptr(arr(bc_nodes, np.int32), C.c_int32)
            d.bc_values = _capi.as_ptr(arr(bc_vals, np.float64), C.c_double)
        phi = F.phi
        if phi.builtin == "phi_set":
            d.obstacle_kind = _capi.OBSTACLE_PHI_SET
            d.obstacle_period, d.obstacle_origin = phi.period or 0.0, phi.origin
        elif phi.values is not None:
            d.obstacle_kind = _capi.OBSTACLE_ARRAY
            d.phi_obs_q = _capi.as_ptr(arr(phi.values, np.float64), C.c_double)
        else:
            raise ValueError("the obstacle function has not been interpolated")
        d.f = float(F.f.value)
        halo = mesh.halo
        d.num_neighbors = len(halo.neighbors)
        if halo.neighbors:
            sp = np.cumsum([0] + [s.size for s in halo.send]).astype(np.int64)
            rp = np.cumsum([0] + [r.size for r in halo.recv]).astype(np.int64)
            d.neighbor_ranks = _capi.as_ptr(arr(halo.neighbors, np.int32), C.c_int32)
            d.send_ptr = _capi.as_ptr(arr(sp, np.int64), C.c_int64)
            d.recv_ptr = _capi.as_ptr(arr(rp, np.int64), C.c_int64)
            d.send_nodes = _capi.as_ptr(arr(np.concatenate(halo.send), np.int32), C.c_int32)
            d.recv_nodes = _capi.as_ptr(arr(np.concatenate(halo.recv), np.int32), C.c_int32)
        h = _capi.H()
        _capi.check(self.lib.lvpp_create(C.byref(d), C.byref(h)))
        self.h = h
        del keep
        self.n = V.num_rows  # local vector length (owned + ghost)
        self.n_owned = 2 * V.num_owned_nodes
        self.x = DeviceVector(self.n, self.device)  # current iterate
        self.work = DeviceVector(self.n, self.device)
        self._alpha = None
        if mesh.nranks > 1:
            self._init_comm(mesh.rank, mesh.nranks)

    # -- communication -------------------------------------------------------------------------
    def _init_comm(self, rank, nranks):
        torch = _torch()
        import torch.distributed as dist

        if not dist.is_initialized() or dist.get_world_size() != nranks:
            raise RuntimeError("torch.distributed must be initialised with one rank per mesh part")
        ident = (C.c_uint8 * 128)()
        if rank == 0:
            _capi.check(self.lib.lvpp_comm_unique_id(ident))
        dev = self.device if dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor(list(ident), dtype=torch.uint8, device=dev)
        dist.broadcast(t, src=0)
        ident = (C.c_uint8 * 128)(*t.cpu().tolist())
        _capi.check(self.lib.lvpp_comm_init(self.h, ident, rank, nranks))

    # -- state ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.lvpp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stats(self):
        s = _capi.Stats()
        _capi.check(self.lib.lvpp_get_stats(self.h, C.byref(s)))
        return s.as_dict()

    def set_alpha(self, alpha):
        _capi.check(self.lib.lvpp_set_alpha(self.h, float(alpha)))
        self._alpha = float(alpha)

    def set_previous(self, xk):
        """sol_k <- xk; ``xk`` a DeviceVector (device copy) or a host array (H2D copy)."""
        if isinstance(xk, DeviceVector):
            _torch().cuda.current_stream().synchronize()
            _capi.check(self.lib.lvpp_set_previous(self.h, xk.ptr))
        else:
            a = np.ascontiguousarray(xk, dtype=np.float64)
            _capi.check(self.lib.lvpp_set_previous_host(self.h, _capi.as_ptr(a, C.c_double)))

    def sync_coefficients(self):
        """Read the form's mutable inputs the way a dolfinx assembly would at call time."""
        self.set_alpha(self.F.alpha.value)
        _capi.check(self.lib.lvpp_set_forcing(self.h, float(self.F.f.value)))
        for bc in self._bcs:  # Dirichlet values changed since the last push (u_bc.x.array[...] = ...)
            vals = bc.current_values()
            if not np.array_equal(vals, bc.values):
                nodes = np.ascontiguousarray(bc.nodes, dtype=np.int32)
                vals = np.ascontiguousarray(vals, dtype=np.float64)
                _capi.check(self.lib.lvpp_set_bc_values(self.h, nodes.size, _capi.as_ptr(nodes, C.c_int32), _capi.as_ptr(vals, C.c_double)))
                bc.values = vals
        self.set_previous(self.F.sol_k.x.array)

    # -- assembly / linear algebra ---------------------------------------------------------------
    def csr_pattern(self):
        st = self.stats()
        indptr = np.empty(st["local_rows"] + 1, dtype=np.int64)
        indices = np.empty(st["nnz"], dtype=np.int32)
        _capi.check(self.lib.lvpp_get_csr_pattern(self.h, _capi.as_ptr(indptr, C.c_int64), _capi.as_ptr(indices, C.c_int32)))
        return indptr, indices

    def assemble_residual(self, x: DeviceVector, F: DeviceVector):
        _torch().cuda.current_stream().synchronize()
        fn = C.c_double()
        _capi.check(self.lib.lvpp_assemble_residual(self.h, x.ptr, F.ptr, C.byref(fn)))
        return fn.value

    def assemble_jacobian(self, x: DeviceVector):
        _torch().cuda.current_stream().synchronize()
        _capi.check(self.lib.lvpp_assemble_jacobian(self.h, x.ptr))

    def jacobian_values(self):
        torch = _torch()
        vals = torch.empty(self.stats()["nnz"], dtype=torch.float64, device=self.device)
        torch.cuda.current_stream().synchronize()
        _capi.check(self.lib.lvpp_get_jacobian_values(self.h, vals.data_ptr()))
        return vals

    def spmv(self, v: DeviceVector, y: DeviceVector):
        _torch().cuda.current_stream().synchronize()
        _capi.check(self.lib.lvpp_spmv(self.h, v.ptr, y.ptr))

    def linear_solve(self, rhs: DeviceVector, y: DeviceVector, opts):
        _torch().cuda.current_stream().synchronize()
        its, reason, rn = C.c_int32(), C.c_int32(), C.c_double()
        _capi.check(self.lib.lvpp_linear_solve(self.h, rhs.ptr, y.ptr, C.byref(opts), C.byref(its), C.byref(reason), C.byref(rn)))
        return its.value, reason.value, rn.value

    # -- Newton ----------------------------------------------------------------------------------
    def newton_solve(self, x: DeviceVector, opts):
        _torch().cuda.current_stream().synchronize()
        its, reason, fn, lin = C.c_int32(), C.c_int32(), C.c_double(), C.c_int32()
        _capi.check(self.lib.lvpp_newton_solve(self.h, x.ptr, C.byref(opts), C.byref(its), C.byref(reason), C.byref(fn), C.byref(lin)))
        return reason.value, its.value, fn.value, lin.value

    def newton_solve_host(self, x_host, opts):
        """Host-buffer entry point: H2D of ``x_host``, Newton solve, D2H on convergence."""
        its, reason, fn, lin = C.c_int32(), C.c_int32(), C.c_double(), C.c_int32()
        _capi.check(
            self.lib.lvpp_newton_solve_host(
                self.h, _capi.as_ptr(x_host, C.c_double), C.byref(opts), C.byref(its), C.byref(reason), C.byref(fn), C.byref(lin)
            )
        )
        return reason.value, its.value, fn.value, lin.value

    def newton_begin(self, x: DeviceVector, same_iterate=False):
        """``same_iterate``: ``x`` is where the previous Newton solve on this handle ended and only alpha / f / the
        Dirichlet values / the previous iterate changed since -- D(psi) is kept (lvpp_newton_begin_same_iterate)."""
        _torch().cuda.current_stream().synchronize()
        fn = C.c_double()
        f = self.lib.lvpp_newton_begin_same_iterate if same_iterate else self.lib.lvpp_newton_begin
        _capi.check(f(self.h, x.ptr, C.byref(fn)))
        return fn.value

    def newton_step(self, x: DeviceVector, opts):
        norms = (C.c_double * 3)()
        kits, kreason = C.c_int32(), C.c_int32()
        _capi.check(self.lib.lvpp_newton_step(self.h, x.ptr, C.byref(opts), norms, C.byref(kits), C.byref(kreason)))
        return (norms[0], norms[1], norms[2]), kits.value, kreason.value

    def observables(self, x: DeviceVector):
        _torch().cuda.current_stream().synchronize()
        out = (C.c_double * 6)()
        _capi.check(self.lib.lvpp_observables(self.h, x.ptr, out))
        return np.array(out[:])

    def timer_start(self):
        _capi.check(self.lib.lvpp_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double()
        _capi.check(self.lib.lvpp_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def time_spmv(self, v: DeviceVector, y: DeviceVector, reps=20, flush_l2=False):
        _torch().cuda.current_stream().synchronize()
        ms = C.c_double()
        _capi.check(self.lib.lvpp_time_spmv(self.h, v.ptr, y.ptr, reps, int(flush_l2), C.byref(ms)))
        return ms.value

    def time_assembly(self, x: DeviceVector, F: DeviceVector, reps=5):
        _torch().cuda.current_stream().synchronize()
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        _capi.check(self.lib.lvpp_time_assembly(self.h, x.ptr, F.ptr, reps, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value


# ---------------------------------------------------------------------------------------------
# PETSc-style options
def newton_options(options, generic=False):
    """Translate a PETSc options dict (keys without prefix, obstacle_pg.py:128-139) to lvpp_newton_opts.

    ``ksp_type preonly`` + ``pc_type lu`` (the reference's direct solve) maps to MINRES converged to
    ``ksp_rtol`` (default 1e-12 on the preconditioned residual), which reproduces the Newton
    iterates of an exact solve to rounding."""
    o = _capi.NewtonOpts.defaults()
    opts = dict(options or {})
    # PETSc's default for newtonls is "bt"; the obstacle driver always sets "none" (obstacle_pg.py:136)
    ls = opts.get("snes_linesearch_type", "bt" if generic else "none")
    if ls in ("none", "basic"):
        o.snes_linesearch = _capi.LINESEARCH_NONE
    elif ls == "bt":
        # mixed-form engine: inside the library (forms.cu); obstacle engine: the host loop of linesearch.py over
        # the library's assembly / Krylov / J*v entry points
        o.snes_linesearch = _capi.LINESEARCH_BT
    elif ls == "l2" and not generic:
        # PETSc's secant line search (the reference's examples 03, 07-10): obstacle engine only, host loop of linesearch.py
        o.snes_linesearch = _capi.LINESEARCH_L2
    else:
        raise NotImplementedError(f"snes_linesearch_type {ls!r}")
    # snes_linesearch_maxlambda (PETSc >= 3.23; fracture_dolfinx.py:135) / snes_linesearch_maxstep (older name)
    o.linesearch_maxstep = float(opts.get("snes_linesearch_maxlambda", opts.get("snes_linesearch_maxstep", 1e8)))
    # not a PETSc option: bound on the growth of psi per Newton step of the obstacle engine (lvpp_newton_opts.psi_increase_max;
    # off by default = the reference's full step)
    if opts.get("lvpp_psi_increase_max") is not None:
        o.psi_increase_max = float(opts["lvpp_psi_increase_max"])
    if opts.get("lvpp_psi_free_below") is not None:
        o.psi_free_below = float(opts["lvpp_psi_free_below"])
    if opts.get("ksp_gmres_restart") is not None:
        o.ksp_restart = int(opts["ksp_gmres_restart"])
    st = opts.get("snes_type", "newtonls")
    if st != "newtonls":
        raise NotImplementedError(f"snes_type {st!r}")
    kt = opts.get("ksp_type", "preonly")
    if kt not in ("preonly", "minres", "gmres", "fgmres"):
        raise NotImplementedError(f"ksp_type {kt!r}: the Newton system is solved with MINRES or GMRES")
    pc = opts.get("pc_type", "lu")
    if generic:
        # the mixed-form engine has one Krylov solver: block-Jacobi preconditioned restarted GMRES
        if pc not in ("lu", "bjacobi", "pbjacobi", "jacobi", "none"):
            raise NotImplementedError(f"pc_type {pc!r}")
        o.ksp_max_it = 20000
        # these examples push alpha to 1e5 and stop on increments of 1e-8: the stand-in for the reference's LU
        # has to be converged to rounding for the proximal iteration counts to agree
        o.ksp_rtol = 1e-14
    elif pc in ("mg", "gamg") or (pc == "lu" and kt in ("gmres", "fgmres")):
        # monolithic aggregation multigrid + restarted GMRES
        o.pc_type = _capi.PC_MG
        o.ksp_max_it = 2000
    elif pc in ("lu", "jacobi", "fieldsplit", "none"):
        if kt in ("gmres", "fgmres"):
            raise NotImplementedError("GMRES is paired with pc_type mg")
        o.pc_type = _capi.PC_JACOBI
    else:
        raise NotImplementedError(f"pc_type {pc!r}")
    if opts.get("pc_mg_smoothing_sweeps") is not None:
        o.pc_degree = int(opts["pc_mg_smoothing_sweeps"])
    for key, field, cast in (
        ("snes_rtol", "snes_rtol", float),
        ("snes_atol", "snes_atol", float),
        ("snes_stol", "snes_stol", float),
        ("snes_divergence_tolerance", "snes_divtol", float),
        ("snes_max_it", "snes_max_it", int),
        ("ksp_rtol", "ksp_rtol", float),
        ("ksp_atol", "ksp_atol", float),
        ("ksp_max_it", "ksp_max_it", int),
    ):
        if opts.get(key) is not None:
            setattr(o, field, cast(opts[key]))
    return o


def _flag(options, key):
    """PETSc flag semantics: present with value None/True means set (obstacle_pg.py:132-135)."""
    if key not in (options or {}):
        return False
    v = options[key]
    return v is None or bool(v)


# ---------------------------------------------------------------------------------------------
# the reference's classes
class SNESProblem:
    """src/lvpp/problem.py:14-77."""

    def __init__(self, F, u, J=None, bcs=None, form_compiler_options=None, jit_options=None):
        if not isinstance(F, ObstacleResidual):
            raise TypeError("F must be an obstacle residual form (see obstacle_residual())")
        if u is not F.sol:
            raise ValueError("u must be the unknown the form F was written in")
        self.L = F
        self.a = derivative(F, u) if J is None else J
        self.bcs = bcs
        self._F, self._J = None, None
        self.u = u
        self._dev = None

    @property
    def device_problem(self) -> DeviceProblem:
        if self._dev is None:
            self._dev = DeviceProblem(self.L, self.bcs)
        return self._dev

    def create_vector(self):
        dev = self.device_problem
        return DeviceVector(dev.n, dev.device)

    def create_matrix(self):
        return DeviceMatrix(self.device_problem)

    def F(self, snes, x, F):
        """Assemble the residual at ``x`` into ``F`` (ghost update, assemble_vector, apply_lifting
        with scale -1 and x0 = x, reverse scatter, set_bc -- problem.py:54-67 -- all on the device)."""
        dev = self.device_problem
        dev.sync_coefficients()
        self._mirror(x)
        return dev.assemble_residual(x, F)

    def J(self, snes, x, J, P=None):
        """Assemble the Jacobian at ``x`` (zeroEntries, assemble_matrix with bcs, assemble --
        problem.py:69-77)."""
        dev = self.device_problem
        dev.sync_coefficients()
        self._mirror(x)
        dev.assemble_jacobian(x)

    #: the reference's callbacks copy the iterate into ``u`` at every call (problem.py:57,72: ``x.copy(self.u.x.petsc_vec)``).
    #: Here the iterate lives in HBM and ``u.x.array`` is a host vector, so that copy is a device-to-host transfer of the
    #: whole vector per callback; it is done only when this flag is set (a caller that reads ``u`` inside a monitor).
    mirror_iterate = False

    def _mirror(self, x):
        if self.mirror_iterate:
            self.u.x.array[:] = x.numpy()


class NotConvergedError(RuntimeError):
    """Raised for ``snes_error_if_not_converged`` / ``ksp_error_if_not_converged`` (PETSc raises there); a device,
    argument or communication failure is a ``_capi.LvppError`` instead and must not be mistaken for it."""

    def __init__(self, message, reason, iterations):
        super().__init__(message)
        self.reason, self.iterations = reason, iterations


class SNESSolver:
    """src/lvpp/problem.py:80-127."""

    def __init__(self, problem: SNESProblem, options):
        self.problem = problem
        self.options = dict(options or {})
        self.create_solver()
        self.create_data_structures()
        self.converged_reason = 0
        self.iterations = 0
        self.linear_iterations = 0
        self.fnorm = float("nan")

    def create_solver(self):
        self._opts = newton_options(self.options)

    def create_data_structures(self):
        self._A = None
        self._b = None
        self._x = None

    def solve(self, copy_on_failure=False):
        """``copy_on_failure``: also write the last Newton iterate into ``u`` when SNES did not converge -- what
        dolfinx.fem.petsc.NonlinearProblem.solve does (NonlinearProblem below passes True); the reference's own
        SNESSolver.solve only writes ``u`` on convergence (src/lvpp/problem.py:121-123)."""
        dev = self.problem.device_problem
        self.converged_reason, self.iterations, self.linear_iterations = 0, 0, 0  # never report a previous solve
        dev.sync_coefficients()
        xh = self.problem.u.x.array
        if self._opts.snes_linesearch in (_capi.LINESEARCH_BT, _capi.LINESEARCH_L2):
            from . import linesearch

            o = self._opts
            dev.x.set(xh)
            nb = linesearch.NewtonBT(linesearch.DeviceBackend(dev, o), rtol=o.snes_rtol, atol=o.snes_atol, stol=o.snes_stol,
                                     max_it=o.snes_max_it, divtol=o.snes_divtol,
                                     linesearch="l2" if o.snes_linesearch == _capi.LINESEARCH_L2 else "bt",
                                     maxstep=getattr(o, "linesearch_maxstep", 1e8))
            reason, its = nb.solve(dev.x)
            fnorm, lin = nb.fnorm, nb.linear_its
            if reason > 0 or copy_on_failure:  # SNESSolver.solve only overwrites the caller's function on convergence (problem.py:121-123)
                xh[:] = dev.x.numpy()
        else:
            reason, its, fnorm, lin = dev.newton_solve_host(xh, self._opts)  # writes xh only if reason > 0
            if reason <= 0 and copy_on_failure:
                _capi.check(dev.lib.lvpp_get_last_iterate_host(dev.h, _capi.as_ptr(xh, C.c_double)))
        self.converged_reason, self.iterations, self.fnorm, self.linear_iterations = reason, its, fnorm, lin
        if _flag(self.options, "snes_monitor"):
            print(f"  SNES: {its} Newton steps, ||F|| = {fnorm:.6e}, {lin} Krylov iterations, reason {reason}")
        return reason, its


class _KSPView:
    def __init__(self, solver):
        self._s = solver

    def getConvergedReason(self):
        return -3 if self._s.converged_reason == SNES_DIVERGED_LINEAR_SOLVE else 2

    def getIterationNumber(self):
        return self._s.linear_iterations


class _SNESView:
    """What the drivers touch through ``problem.solver`` (obstacle_pg.py:191-192,
    signorini_dolfinx.py:332, gradient_constraint_dolfinx.py:183)."""

    def __init__(self, solver: SNESSolver):
        self._s = solver
        self.ksp = _KSPView(solver)

    def getConvergedReason(self):
        return self._s.converged_reason

    def getIterationNumber(self):
        return self._s.iterations

    def getFunctionNorm(self):
        return self._s.fnorm

    def getLinearSolveIterations(self):
        return self._s.linear_iterations

    def setTolerances(self, rtol=None, atol=None, stol=None, max_it=None):
        o = self._s._opts
        if rtol is not None:
            o.snes_rtol = float(rtol)
        if atol is not None:
            o.snes_atol = float(atol)
        if stol is not None:
            o.snes_stol = float(stol)
        if max_it is not None:
            o.snes_max_it = int(max_it)


class NonlinearProblem:
    """``dolfinx.fem.petsc.NonlinearProblem(F, u, bcs=, J=, petsc_options=, petsc_options_prefix=)``
    as called at obstacle_pg.py:140-142; ``solve()`` runs SNES and returns ``u``."""

    def __init__(self, F, u, bcs=None, J=None, petsc_options=None, petsc_options_prefix="", entity_maps=None,
                 kind=None, jit_options=None, form_compiler_options=None):
        if isinstance(u, (list, tuple)):
            # blocked problem (signorini_dolfinx.py:283-291: F = extract_blocks(residual), u = [u, psi], entity_maps,
            # kind="mpi"): the mixed-form engine, forms.BlockedNonlinearProblem
            from .forms import BlockedForm, BlockedNonlinearProblem

            if not isinstance(F, BlockedForm):
                raise TypeError("a list of unknowns needs a blocked form (forms.BlockedForm, e.g. signorini.setup()['F'])")
            self._blocked = BlockedNonlinearProblem(F, u, petsc_options)
            self._options = dict(petsc_options or {})
            self.solver = self._blocked.solver
            self.u = list(u)
            self.prefix = petsc_options_prefix
            return
        self._blocked = None
        self._problem = SNESProblem(F, u, J=J, bcs=bcs)
        self._options = dict(petsc_options or {})
        self._snes = SNESSolver(self._problem, self._options)
        self.solver = _SNESView(self._snes)
        self.u = u
        self.prefix = petsc_options_prefix

    @property
    def device_problem(self):
        return self._blocked.F.dev if self._blocked is not None else self._problem.device_problem

    def solve(self):
        if self._blocked is not None:
            return self._blocked.solve()
        reason, its = self._snes.solve(copy_on_failure=True)  # dolfinx leaves the last iterate in u whatever the reason
        if reason == SNES_DIVERGED_LINEAR_SOLVE and _flag(self._options, "ksp_error_if_not_converged"):
            raise NotConvergedError("KSP did not converge (ksp_error_if_not_converged)", reason, its)
        if reason <= 0 and _flag(self._options, "snes_error_if_not_converged"):
            raise NotConvergedError(f"SNES did not converge: reason {reason} after {its} iterations", reason, its)
        return self.u
