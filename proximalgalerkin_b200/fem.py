"""Host-side finite element bookkeeping with the call shape of ``dolfinx.fem`` as the reference uses it.

Mirrors, for the LVPP obstacle path, what examples/01_obstacle_problem/obstacle_pg.py:68-111 builds
with dolfinx/basix: the mixed P_p x P_p space (:68-70), ``Constant`` (:73-74), boundary dof location
and ``dirichletbc`` on the u sub-space (:76-83), ``Function`` with a host ``.x.array`` the driver
mutates in place (:86-87,157-158,226) and the quadrature-space obstacle function (:106-111).

Unknown numbering of the mixed space: node-interleaved, u at scalar node n -> 2n, psi -> 2n + 1
(a recorded permutation of dolfinx's numbering).
"""
import numpy as np

from . import quadrature as _quad
from .mesh import Mesh

# basix reference topology: edge e joins these local vertices
EDGES = {
    2: ((1, 2), (0, 2), (0, 1)),
    3: ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)),
}


def tabulate_lagrange(degree, points):
    """Lagrange P1/P2 basis on the reference simplex: (phi [nq, nld], dphi [nq, nld, tdim]).
    Node order: vertices, then (P2) edge midpoints in basix edge order."""
    points = np.asarray(points, dtype=np.float64)
    nq, tdim = points.shape
    lam = np.empty((nq, tdim + 1))
    lam[:, 0] = 1.0 - points.sum(axis=1)
    lam[:, 1:] = points
    dlam = np.vstack([-np.ones((1, tdim)), np.eye(tdim)])
    if degree == 1:
        return lam, np.tile(dlam[None], (nq, 1, 1))
    if degree != 2:
        raise ValueError("only Lagrange degree 1 and 2 are supported")
    edges = EDGES[tdim]
    nld = tdim + 1 + len(edges)
    phi = np.empty((nq, nld))
    dphi = np.empty((nq, nld, tdim))
    for v in range(tdim + 1):
        phi[:, v] = lam[:, v] * (2.0 * lam[:, v] - 1.0)
        dphi[:, v] = (4.0 * lam[:, v] - 1.0)[:, None] * dlam[v]
    for e, (a, b) in enumerate(edges):
        k = tdim + 1 + e
        phi[:, k] = 4.0 * lam[:, a] * lam[:, b]
        dphi[:, k] = 4.0 * (lam[:, [a]] * dlam[b] + lam[:, [b]] * dlam[a])
    return phi, dphi


def _host_zeros(n):
    """Zeroed float64 host vector; page-locked when a CUDA device is present, so that the H2D / D2H copies of
    ``sol`` and ``sol_k`` around every proximal step (``lvpp_newton_solve_host``, ``lvpp_set_previous_host``) run
    at PCIe speed instead of through the driver's pageable-memory staging.  The numpy view keeps the tensor alive."""
    try:
        import torch

        if torch.cuda.is_available():
            return torch.zeros(n, dtype=torch.float64, pin_memory=True).numpy()
    except Exception:
        pass
    return np.zeros(n, dtype=np.float64)


class _Array:
    """``Function.x``: owns the host array (``.array``), like dolfinx.la.Vector."""

    def __init__(self, n):
        self.array = _host_zeros(n)

    def scatter_forward(self):  # ghosts are refreshed on the device by the solver
        return None


class FunctionSpace:
    """Mixed (u, psi) Lagrange space P_p x P_p on a :class:`Mesh` (obstacle_pg.py:68-70)."""

    def __init__(self, mesh: Mesh, degree=1, quadrature_degree=6, rule=None, _sub=None):
        self.mesh = mesh
        self.degree = int(degree)
        self._sub = _sub
        tdim = mesh.tdim
        if rule is None:
            rule = _quad.make_quadrature(mesh.cell_name, quadrature_degree)
        self.qpoints = np.ascontiguousarray(rule[0], dtype=np.float64)
        self.qweights = np.ascontiguousarray(rule[1], dtype=np.float64)
        self.quadrature_degree = quadrature_degree
        phi, dphi = tabulate_lagrange(self.degree, self.qpoints)
        self.phi_tab = np.ascontiguousarray(phi)
        self.dphi_tab = np.ascontiguousarray(dphi)
        if self.degree == 1:
            self.cell_nodes = mesh.cells
            self.num_nodes = mesh.num_vertices
            self.num_owned_nodes = mesh.num_owned_vertices
            self.node_coords = mesh.coords
            self.boundary_nodes = mesh.boundary_vertices
        else:
            if mesh.nranks != 1:
                raise NotImplementedError("P2 spaces are single-rank for now")
            self._build_p2()
        self.nld = self.cell_nodes.shape[1]

    def _build_p2(self):
        mesh = self.mesh
        cells = mesh.cells.astype(np.int64)
        nv = mesh.num_vertices
        edges = EDGES[mesh.tdim]
        lo = np.stack([np.minimum(cells[:, a], cells[:, b]) for a, b in edges], axis=1)
        hi = np.stack([np.maximum(cells[:, a], cells[:, b]) for a, b in edges], axis=1)
        key = (lo * nv + hi).ravel()
        uniq, inv = np.unique(key, return_inverse=True)
        edge_nodes = (inv.reshape(cells.shape[0], len(edges)) + nv).astype(np.int32)
        self.cell_nodes = np.ascontiguousarray(np.concatenate([mesh.cells, edge_nodes], axis=1))
        self.num_nodes = nv + uniq.size
        self.num_owned_nodes = self.num_nodes
        e0, e1 = uniq // nv, uniq % nv
        self.node_coords = np.vstack([mesh.coords, 0.5 * (mesh.coords[e0] + mesh.coords[e1])])
        # boundary edge nodes: edges of exterior facets.  An edge lies on the boundary iff it is an
        # edge of a facet that belongs to a single cell.
        nvc = cells.shape[1]
        facets = np.concatenate([np.delete(cells, i, axis=1) for i in range(nvc)], axis=0)
        facets.sort(axis=1)
        fu, cnt = np.unique(facets, axis=0, return_counts=True)
        ext = fu[cnt == 1]
        bkeys = []
        for i in range(ext.shape[1]):
            for j in range(i + 1, ext.shape[1]):
                bkeys.append(ext[:, i] * nv + ext[:, j])
        bkeys = np.unique(np.concatenate(bkeys))
        bedge = np.searchsorted(uniq, bkeys) + nv
        self.boundary_nodes = np.unique(np.concatenate([mesh.boundary_vertices, bedge])).astype(np.int32)

    # -- dolfinx-like surface ---------------------------------------------------------------
    @property
    def num_rows(self):
        return 2 * self.num_nodes

    def sub(self, i):
        if i not in (0, 1):
            raise IndexError(i)
        s = object.__new__(FunctionSpace)
        s.__dict__.update(self.__dict__)
        s._sub = i
        return s

    def collapse(self):
        """(collapsed scalar space, map from its dofs to the mixed numbering), obstacle_pg.py:78"""
        if self._sub is None:
            raise ValueError("collapse() is for sub-spaces")
        return self, 2 * np.arange(self.num_nodes, dtype=np.int64) + self._sub

    def dof(self, field, nodes):
        return 2 * np.asarray(nodes, dtype=np.int64) + field


def functionspace(mesh, element=("Lagrange", 1), quadrature_degree=6, rule=None):
    """``fem.functionspace(msh, mixed_element([P, P]))`` for P = Lagrange degree 1 or 2."""
    family, degree = element
    if family not in ("Lagrange", "P", "CG"):
        raise ValueError(f"unsupported element family {family}")
    return FunctionSpace(mesh, degree, quadrature_degree, rule)


class Constant:
    """``fem.Constant``: mutable scalar (``alpha.value = ...``, obstacle_pg.py:176-183)."""

    def __init__(self, mesh, value):
        self.mesh = mesh
        self.value = float(value)


class Function:
    """``fem.Function`` on the mixed space; ``.x.array`` is the host vector the driver mutates."""

    def __init__(self, V: FunctionSpace, name="f"):
        self.function_space = V
        self.name = name
        self.x = _Array(V.num_rows)

    def sub(self, i):
        return self.x.array[i::2]

    def interpolate(self, other):
        if isinstance(other, Function):
            self.x.array[:] = other.x.array
        else:
            raise TypeError("mixed Functions interpolate from Functions only")


class QuadratureFunction:
    """Function in a quadrature-element space (obstacle_pg.py:106-111): one value per cell and
    quadrature point, ``interpolate(callable)`` evaluates at the physical quadrature points."""

    def __init__(self, V: FunctionSpace, name="phi"):
        self.function_space = V
        self.name = name
        self.values = None  # [num_cells, nq]
        self.builtin = None
        self.period, self.origin = None, 0.0

    def interpolate(self, expr):
        V = self.function_space
        mesh = V.mesh
        lam, _ = tabulate_lagrange(1, V.qpoints)
        xq = np.einsum("qv,cvg->gcq", lam, mesh.coords[mesh.cells])
        vals = np.asarray(expr(xq.reshape(mesh.gdim, -1)), dtype=np.float64)
        self.values = np.ascontiguousarray(vals.reshape(mesh.num_cells, V.qpoints.shape[0]))
        self.builtin = None

    def interpolate_phi_set(self, period=None, origin=0.0):
        """Use the closed-form obstacle of obstacle_pg.py:92-104, evaluated on the device (no
        cells x nq host array; needed at the 20 M-row configuration).  ``period``: tile the obstacle
        along the last axis (one copy per period starting at ``origin``) -- the weak-scaling workload
        stacks one [-1, 1]^3 problem per GPU."""
        self.values = None
        self.builtin = "phi_set"
        self.period, self.origin = period, float(origin)


def phi_set(x):
    """The obstacle of obstacle_pg.py:92-104 (r = Euclidean norm of the point)."""
    r = np.sqrt(np.sum(np.asarray(x, dtype=np.float64) ** 2, axis=0))
    r0, beta = 0.5, 0.9
    b = r0 * beta
    tmp = np.sqrt(r0**2 - b**2)
    B = tmp + b * b / tmp
    Cc = -b / tmp
    return np.where(r > b, B + r * Cc, np.sqrt(np.maximum(r0**2 - r**2, 0.0)))


def locate_dofs_boundary(Vsub):
    """Mixed dof numbers of the sub-space's nodes on the exterior boundary
    (mesh.exterior_facet_indices + fem.locate_dofs_topological, obstacle_pg.py:76-79)."""
    if Vsub._sub is None:
        raise ValueError("pass a sub-space, e.g. V.sub(0)")
    return Vsub.dof(Vsub._sub, Vsub.boundary_nodes)


class DirichletBC:
    def __init__(self, value, dofs, V):
        if V._sub != 0:
            raise NotImplementedError("Dirichlet conditions are supported on the u sub-space only")
        self.function_space = V
        self.dofs = np.asarray(dofs, dtype=np.int64)
        if np.any(self.dofs % 2 != 0):
            raise ValueError("dofs do not belong to V.sub(0)")
        self.nodes = (self.dofs // 2).astype(np.int32)
        self._source = value  # read again at every solve, like dolfinx reads the bc's Function / Constant at assembly time
        self.values = self.current_values()

    def current_values(self):
        """The Dirichlet values as the source object holds them now (a ``Function``, a ``Constant``, an array over the
        scalar nodes or a scalar)."""
        value = self._source
        if isinstance(value, Function):
            return value.x.array[self.dofs].copy()
        if isinstance(value, Constant):
            return np.full(self.nodes.size, float(value.value))
        v = np.asarray(value, dtype=np.float64)
        return np.full(self.nodes.size, float(v)) if v.ndim == 0 else v[self.nodes].copy()


def dirichletbc(value, dofs, V):
    """``fem.dirichletbc(value=u_bc, dofs=dofs, V=V.sub(0))`` (obstacle_pg.py:81-83).  ``value``: a
    scalar, an array over the scalar nodes, or a mixed ``Function``."""
    return DirichletBC(value, dofs, V)
