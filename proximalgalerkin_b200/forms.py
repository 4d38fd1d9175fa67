"""Host side of the generic mixed-form engine (``lvpp_form_*`` in include/lvpp_b200.h).

What the reference gets from dolfinx for the gradient-constraint, multiphase and Signorini examples --
mixed dofmaps, ``create_matrix`` sparsity (union over integration entities of dofs x dofs), the
cell-to-nnz map of ``MatSetValuesLocal``, Dirichlet dof lists -- is built here with numpy and handed
to the library as plain arrays; residuals, Jacobians, the Krylov solve and the Newton loop run on the
GPU.  There is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _capi
from .problem import DeviceVector, _torch, newton_options


def csr_pattern(num_dofs, dof_lists):
    """Union over entities of dofs x dofs: (indptr int64, indices int32, [to_nnz int64 per list])."""
    keys = []
    for cd in dof_lists:
        cd = np.asarray(cd, dtype=np.int64)
        n = cd.shape[1]
        rows = np.repeat(cd, n, axis=1)
        cols = np.tile(cd, (1, n))
        keys.append((rows * num_dofs + cols).ravel())
    uniq, inv = np.unique(np.concatenate(keys), return_inverse=True)
    rows = uniq // num_dofs
    indices = (uniq % num_dofs).astype(np.int32)
    indptr = np.zeros(num_dofs + 1, dtype=np.int64)
    np.add.at(indptr, rows + 1, 1)
    indptr = np.cumsum(indptr)
    maps, off = [], 0
    for k in keys:
        maps.append(np.ascontiguousarray(inv[off : off + k.size], dtype=np.int64))
        off += k.size
    return indptr, indices, maps


class Integral:
    """One integration block of a form: entity dof lists, geometry vertices and tables."""

    def __init__(self, dofs, vertices, qweights=None, tab_a=None, dtab_a=None, tab_b=None):
        self.dofs = np.ascontiguousarray(dofs, dtype=np.int32)
        self.vertices = np.ascontiguousarray(vertices, dtype=np.int32)
        c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        self.qweights, self.tab_a, self.dtab_a, self.tab_b = c(qweights), c(tab_a), c(dtab_a), c(tab_b)


class FormProblem:
    """Owns one ``lvpp_form_handle``."""

    def __init__(self, form, gdim, num_dofs, vertex_coords, integrals, params, bc_dofs=(), bc_values=None,
                 coef0=None, coef1=None, blocks=None, device=None):
        torch = _torch()
        if not torch.cuda.is_available():
            raise _capi.LvppError(_capi.E_NOGPU, "no CUDA device: the LVPP path has no CPU fallback")
        self.lib = _capi.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        torch.cuda.set_device(self.device)
        self.n = int(num_dofs)
        self.integrals = integrals
        self.indptr, self.indices, maps = csr_pattern(self.n, [i.dofs for i in integrals])
        self.nnz = int(self.indptr[-1])
        coords = np.ascontiguousarray(vertex_coords, dtype=np.float64)
        self.bc_dofs = np.ascontiguousarray(bc_dofs, dtype=np.int64)
        bcv = np.zeros(self.bc_dofs.size) if bc_values is None else np.ascontiguousarray(bc_values, dtype=np.float64)
        if blocks is None:
            blocks = [np.array([i]) for i in range(self.n)]
        bptr = np.zeros(len(blocks) + 1, dtype=np.int64)
        bptr[1:] = np.cumsum([len(b) for b in blocks])
        bdofs = np.ascontiguousarray(np.concatenate(blocks), dtype=np.int32)
        params = np.ascontiguousarray(params, dtype=np.float64)
        descs = (_capi.IntegralDesc * len(integrals))()
        keep = [coords, bcv, bptr, bdofs, params, maps]
        P = _capi.as_ptr
        for k, itg in enumerate(integrals):
            d = descs[k]
            d.num_entities, d.nld, d.nv = itg.dofs.shape[0], itg.dofs.shape[1], itg.vertices.shape[1]
            d.dofs, d.vertices, d.to_nnz = P(itg.dofs, C.c_int32), P(itg.vertices, C.c_int32), P(maps[k], C.c_int64)
            d.nq = 0 if itg.qweights is None else itg.qweights.size
            for name in ("qweights", "tab_a", "dtab_a", "tab_b"):
                a = getattr(itg, name)
                setattr(d, name, None if a is None else P(a, C.c_double))
        fd = _capi.FormDesc()
        fd.form, fd.gdim, fd.num_dofs, fd.num_vertices = form, gdim, self.n, coords.shape[0]
        fd.vertex_coords, fd.indptr, fd.indices = P(coords, C.c_double), P(self.indptr, C.c_int64), P(self.indices, C.c_int32)
        fd.num_bc = self.bc_dofs.size
        fd.bc_dofs = P(self.bc_dofs, C.c_int64) if self.bc_dofs.size else None
        fd.bc_values = P(bcv, C.c_double) if self.bc_dofs.size else None
        fd.num_integrals, fd.num_params, fd.integrals, fd.params = len(integrals), params.size, descs, P(params, C.c_double)
        for name, a in (("coef0", coef0), ("coef1", coef1)):
            if a is not None:
                a = np.ascontiguousarray(a, dtype=np.float64)
                if a.size != self.n:
                    raise ValueError(f"{name} must have one entry per dof")
                keep.append(a)
                setattr(fd, name, P(a, C.c_double))
        fd.num_blocks, fd.block_ptr, fd.block_dofs = len(blocks), P(bptr, C.c_int64), P(bdofs, C.c_int32)
        h = _capi.H()
        _capi.check(self.lib.lvpp_form_create(C.byref(fd), C.byref(h)))
        self.handle = h
        del keep

    def __del__(self):
        h = getattr(self, "handle", None)
        if h is not None and getattr(self, "lib", None) is not None:
            self.lib.lvpp_form_destroy(h)
            self.handle = None

    # -- state ------------------------------------------------------------------------------
    def vector(self, host=None):
        v = DeviceVector(self.n, self.device)
        if host is not None:
            v.set(host)
        return v

    def set_param(self, index, value):
        _capi.check(self.lib.lvpp_form_set_param(self.handle, int(index), float(value)))

    def set_aux(self, which, v):
        if not isinstance(v, DeviceVector):
            v = self.vector(v)
        _capi.check(self.lib.lvpp_form_set_aux(self.handle, int(which), v.ptr))

    def set_bc_values(self, values):
        values = np.ascontiguousarray(values, dtype=np.float64)
        if values.size != self.bc_dofs.size:
            raise ValueError("one value per Dirichlet dof")
        _capi.check(self.lib.lvpp_form_set_bc_values(self.handle, _capi.as_ptr(values, C.c_double)))

    # -- assembly / algebra -------------------------------------------------------------------
    def assemble_residual(self, X, F):
        fn = C.c_double()
        _capi.check(self.lib.lvpp_form_assemble_residual(self.handle, X.ptr, F.ptr, C.byref(fn)))
        return fn.value

    def jacobian_values(self):
        torch = _torch()
        vals = torch.empty(self.nnz, dtype=torch.float64, device=self.device)
        _capi.check(self.lib.lvpp_form_get_jacobian_values(self.handle, vals.data_ptr()))
        return vals.cpu().numpy()

    def spmv(self, X, Y):
        _capi.check(self.lib.lvpp_form_spmv(self.handle, X.ptr, Y.ptr))

    def linear_solve(self, R, Y, opts):
        its, reason, rn = C.c_int32(), C.c_int32(), C.c_double()
        _capi.check(self.lib.lvpp_form_linear_solve(self.handle, R.ptr, Y.ptr, C.byref(opts), C.byref(its), C.byref(reason), C.byref(rn)))
        return its.value, reason.value, rn.value

    def newton_solve(self, X, opts):
        """(reason, iterations, fnorm, linear iterations); X is updated in place."""
        its, reason, fn, lin = C.c_int32(), C.c_int32(), C.c_double(), C.c_int32()
        _capi.check(self.lib.lvpp_form_newton_solve(self.handle, X.ptr, C.byref(opts), C.byref(its), C.byref(reason), C.byref(fn), C.byref(lin)))
        return reason.value, its.value, fn.value, lin.value

    def increment_sq(self, X, X0):
        out = C.c_double()
        _capi.check(self.lib.lvpp_form_increment_sq(self.handle, X.ptr, X0.ptr, C.byref(out)))
        return out.value

    def stats(self):
        s = _capi.Stats()
        _capi.check(self.lib.lvpp_form_get_stats(self.handle, C.byref(s)))
        return s.as_dict()

    def time_kernels(self, X, reps=5):
        a, b = C.c_double(), C.c_double()
        _capi.check(self.lib.lvpp_form_time_kernels(self.handle, X.ptr, int(reps), C.byref(a), C.byref(b)))
        return a.value, b.value


class _Solver:
    """``problem.solver``: getIterationNumber / getConvergedReason / setTolerances / ksp."""

    class _KSP:
        def __init__(self):
            self.reason, self.its = 0, 0

        def getConvergedReason(self):
            return self.reason

        def getIterationNumber(self):
            return self.its

    def __init__(self, opts):
        self.opts = opts
        self.reason, self.its, self.fnorm = 0, 0, 0.0
        self.ksp = _Solver._KSP()

    def getIterationNumber(self):
        return self.its

    def getConvergedReason(self):
        return self.reason

    def getFunctionNorm(self):
        return self.fnorm

    def setTolerances(self, rtol=None, atol=None, stol=None, max_it=None):
        for name, v in (("snes_rtol", rtol), ("snes_atol", atol), ("snes_stol", stol), ("snes_max_it", max_it)):
            if v is not None:
                setattr(self.opts, name, v)


class FormNonlinearProblem:
    """The ``dolfinx.fem.petsc.NonlinearProblem(F, u, bcs, petsc_options=...)`` call shape over a
    :class:`FormProblem`: ``x`` is the host solution array the driver mutates; ``solve()`` copies it to
    the device, runs the Newton loop there and copies the result back (PETSc semantics: the iterate is
    returned whatever the reason; ``snes_error_if_not_converged`` raises)."""

    def __init__(self, dev: FormProblem, x, petsc_options=None):
        self.dev = dev
        self.x = x
        self.options = dict(petsc_options or {})
        self.solver = _Solver(newton_options(self.options, generic=True))
        self.X = dev.vector()

    def solve(self):
        self.X.set(self.x)
        reason, its, fnorm, lin = self.dev.newton_solve(self.X, self.solver.opts)
        s = self.solver
        s.reason, s.its, s.fnorm = reason, its, fnorm
        s.ksp.its = lin
        s.ksp.reason = -3 if reason == -3 else 2
        self.x[:] = self.X.numpy()
        err = self.options.get("snes_error_if_not_converged", False)
        if reason <= 0 and (err is None or err):
            from .problem import NotConvergedError

            raise NotConvergedError(f"SNES did not converge: reason {reason} after {its} iterations", reason, its)
        return self.x



# ---------------------------------------------------------------------------------------------
# blocked problems: NonlinearProblem(F, [u, psi], bcs=..., J=..., entity_maps=..., kind="mpi")
# (examples/02_signorini/signorini_dolfinx.py:283-291)
class _Vec:
    def __init__(self, n):
        self.array = np.zeros(n)


class BlockFunction:
    """One unknown of a blocked problem (``fem.Function(V)`` / ``fem.Function(W)``, signorini_dolfinx.py:227-228):
    ``.x.array`` is the host vector the driver reads and mutates."""

    def __init__(self, n, name="f"):
        self.x = _Vec(n)
        self.name = name


class BlockedForm:
    """What ``ufl.extract_blocks(residual)`` (signorini_dolfinx.py:252) is to the reference: the residual blocks of a
    :class:`FormProblem` whose unknown is the concatenation of ``unknowns``.  ``sync`` pushes the form's mutable inputs
    (``alpha.value``, ``psi_k.x.array``, Dirichlet values) to the device the way a dolfinx assembly reads them at
    call time."""

    def __init__(self, dev: FormProblem, unknowns, sync=None):
        self.dev = dev
        self.unknowns = list(unknowns)
        self.sizes = [f.x.array.size for f in self.unknowns]
        if sum(self.sizes) != dev.n:
            raise ValueError("block sizes do not add up to the number of dofs of the form")
        self.sync = sync or (lambda: None)


class BlockedNonlinearProblem:
    """``NonlinearProblem(F, [u, psi], ...)``: gathers the blocks into the mixed vector, runs the Newton loop on the
    device, scatters the iterate back into the blocks (whatever the reason, like dolfinx)."""

    def __init__(self, F: BlockedForm, u, petsc_options=None):
        if [id(f) for f in u] != [id(f) for f in F.unknowns]:
            raise ValueError("u must be the list of unknowns the blocked form was written in")
        self.F = F
        self._x = np.zeros(F.dev.n)
        self._inner = FormNonlinearProblem(F.dev, self._x, petsc_options)
        self.solver = self._inner.solver
        self.u = list(u)

    def _gather(self):
        off = 0
        for f, n in zip(self.F.unknowns, self.F.sizes):
            self._x[off:off + n] = f.x.array
            off += n

    def _scatter(self):
        off = 0
        for f, n in zip(self.F.unknowns, self.F.sizes):
            f.x.array[:] = self._x[off:off + n]
            off += n

    def solve(self):
        self.F.sync()
        self._gather()
        try:
            self._inner.solve()
        finally:
            self._scatter()
        return self.u
