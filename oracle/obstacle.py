"""Obstacle-problem LVPP forms, assembled cell by cell (oracle; see oracle/__init__.py).

Literal restatement of what dolfinx + FFCx evaluate for the forms at
examples/01_obstacle_problem/obstacle_pg.py:116-125 (residual F, Jacobian J = derivative(F, sol)),
the Dirichlet conventions of src/lvpp/problem.py:54-77 (assemble_vector, apply_lifting with
scale -1 and x0 = x, set_bc with scale -1; assemble_matrix with bcs: BC rows/cols zeroed, unit
diagonal), the obstacle interpolated into a quadrature space (:92-111) and the six observables
(:145-152).  Everything is evaluated by quadrature at every point, exactly as the generated
kernels do -- no use is made of the K/M/D block structure here (the GPU path does; that is what
this file checks).
"""
import numpy as np
import scipy.sparse as sp

from . import elements
from .mesh import boundary_vertices, exterior_facets
from .quadrature import make_quadrature


def phi_set(x):
    """Obstacle of obstacle_pg.py:92-104; ``x`` is [gdim, npts]; r is the Euclidean norm of x."""
    r = np.sqrt(np.sum(np.asarray(x) ** 2, axis=0))
    r0 = 0.5
    beta = 0.9
    b = r0 * beta
    tmp = np.sqrt(r0**2 - b**2)
    B = tmp + b * b / tmp
    C = -b / tmp
    cond_true = B + r * C
    cond_false = np.sqrt(np.maximum(r0**2 - r**2, 0.0))
    return np.where(r > b, cond_true, cond_false)


class ObstacleOracle:
    """Mixed (u, psi) P_p x P_p obstacle problem on a simplicial mesh.

    ``layout``: "interleaved" (u_n -> 2n, psi_n -> 2n+1), "blocked" (u_n -> n, psi_n -> N+n), or a
    pair of integer arrays (dof_u, dof_psi) -- e.g. a dofmap exported from dolfinx.
    """

    def __init__(self, mesh, degree=1, quadrature_degree=6, scheme="default", f=0.0, layout="interleaved",
                 phi=phi_set):
        self.mesh = mesh
        self.degree = degree
        self.f = float(f)
        cell = mesh.cell_name
        if isinstance(scheme, (tuple, list)):  # an explicit rule (points [nq, tdim], weights [nq]), e.g. basix's, exported
            self.qpts, self.qwts = (np.asarray(a, dtype=np.float64) for a in scheme)  # by tools/export_from_dolfinx.py
        else:
            self.qpts, self.qwts = make_quadrature(cell, quadrature_degree, scheme)
        self.phi_tab, self.dphi_tab = elements.tabulate(degree, self.qpts)
        self.cell_nodes, self.num_nodes, self.node_coords = elements.build_nodes(mesh, degree)
        self.nld = self.cell_nodes.shape[1]
        N = self.num_nodes
        if isinstance(layout, str):
            if layout == "interleaved":
                self.dof_u = 2 * np.arange(N, dtype=np.int64)
                self.dof_psi = self.dof_u + 1
            elif layout == "blocked":
                self.dof_u = np.arange(N, dtype=np.int64)
                self.dof_psi = self.dof_u + N
            else:
                raise ValueError(layout)
        else:
            self.dof_u, self.dof_psi = (np.asarray(a, dtype=np.int64) for a in layout)
        self.num_rows = 2 * N
        self.detJ, self.Jinv = elements.geometry(mesh)
        self.scale = np.abs(self.detJ)
        # physical gradients of the basis: [C, nq, nld, gdim]
        self.gphi = np.einsum("qad,cdg->cqag", self.dphi_tab, self.Jinv)
        # physical quadrature points [C, nq, gdim] through the affine (P1) geometry map
        lam_v, _ = elements.tabulate(1, self.qpts)
        self.xq = np.einsum("qv,cvg->cqg", lam_v, mesh.coords[mesh.cells])
        # obstacle at quadrature points: the quadrature-element Function of obstacle_pg.py:106-111
        self.phi_q = phi(np.moveaxis(self.xq, -1, 0).reshape(mesh.gdim, -1)).reshape(self.xq.shape[:2])
        # mixed cell dofmap [C, 2*nld]: u block then psi block
        self.cell_dofs = np.concatenate(
            [self.dof_u[self.cell_nodes], self.dof_psi[self.cell_nodes]], axis=1
        )
        # Dirichlet: u = 0 on exterior facets, located topologically (obstacle_pg.py:76-83)
        self.bc_nodes = self._boundary_nodes()
        self.bc_dofs = np.sort(self.dof_u[self.bc_nodes])
        self.bc_values = np.zeros(self.num_rows)  # g, stored on all rows, used on bc_dofs only
        self.is_bc = np.zeros(self.num_rows, dtype=bool)
        self.is_bc[self.bc_dofs] = True
        self._build_pattern()

    def _boundary_nodes(self):
        mesh = self.mesh
        bv = boundary_vertices(mesh)
        if self.degree == 1:
            return bv
        # P2: edge nodes of exterior facets as well
        isb = np.zeros(mesh.num_vertices, dtype=bool)
        ef = exterior_facets(mesh)
        edge_nodes = set()
        nv = mesh.tdim + 1
        # map (v0, v1) -> node through cell_nodes
        emap = {}
        for e, (a, b) in enumerate(elements.EDGES[mesh.tdim]):
            va = self.mesh.cells[:, a]
            vb = self.mesh.cells[:, b]
            lo, hi = np.minimum(va, vb), np.maximum(va, vb)
            for l, h, n in zip(lo, hi, self.cell_nodes[:, nv + e]):
                emap[(int(l), int(h))] = int(n)
        for fct in ef:
            for i in range(len(fct)):
                for j in range(i + 1, len(fct)):
                    edge_nodes.add(emap[(int(min(fct[i], fct[j])), int(max(fct[i], fct[j])))])
        del isb
        return np.unique(np.concatenate([bv, np.array(sorted(edge_nodes), dtype=np.int64)]))

    # ---------------------------------------------------------------- sparsity (create_matrix)
    def _build_pattern(self):
        """CSR pattern = union over cells of cell_dofs x cell_dofs, columns sorted
        (dolfinx.fem.petsc.create_matrix, reference call site src/lvpp/problem.py:110)."""
        cd = self.cell_dofs
        n = cd.shape[1]
        rows = np.repeat(cd, n, axis=1).ravel()
        cols = np.tile(cd, (1, n)).ravel()
        key = rows * self.num_rows + cols
        uniq, inv = np.unique(key, return_inverse=True)
        self.csr_rows = (uniq // self.num_rows).astype(np.int64)
        self.indices = (uniq % self.num_rows).astype(np.int32)
        self.indptr = np.zeros(self.num_rows + 1, dtype=np.int64)
        np.add.at(self.indptr, self.csr_rows + 1, 1)
        self.indptr = np.cumsum(self.indptr)
        self.cell_to_nnz = inv.reshape(cd.shape[0], n, n)  # the scatter map
        self.nnz = uniq.size

    # ---------------------------------------------------------------- element kernels
    def _local(self, x):
        """u and psi local coefficient arrays [C, nld]."""
        return x[self.dof_u[self.cell_nodes]], x[self.dof_psi[self.cell_nodes]]

    def element_residual(self, x, xk, alpha):
        """tabulate_tensor of F (obstacle_pg.py:116-124): [C, 2*nld]."""
        ul, pl = self._local(x)
        _, pkl = self._local(xk)
        phi, w, s = self.phi_tab, self.qwts, self.scale
        uq = ul @ phi.T
        pq = pl @ phi.T
        pkq = pkl @ phi.T
        gu = np.einsum("ca,cqag->cqg", ul, self.gphi)
        ws = w[None, :] * s[:, None]  # [C, nq]
        Fu = alpha * np.einsum("cqg,cqag->ca", ws[:, :, None] * gu, self.gphi)
        Fu += np.einsum("cq,qa->ca", ws * (pq - alpha * self.f - pkq), phi)
        Fp = np.einsum("cq,qa->ca", ws * (uq - np.exp(pq) - self.phi_q), phi)
        return np.concatenate([Fu, Fp], axis=1)

    def element_jacobian(self, x, alpha, cells=None):
        """tabulate_tensor of J = derivative(F, sol) (obstacle_pg.py:125): [C, 2*nld, 2*nld]
        (``cells``: optional subset, as apply_lifting only visits cells with a Dirichlet dof)."""
        sel = slice(None) if cells is None else cells
        pl = x[self.dof_psi[self.cell_nodes[sel]]]
        phi, w, s = self.phi_tab, self.qwts, self.scale[sel]
        gphi = self.gphi[sel]
        pq = pl @ phi.T
        ws = w[None, :] * s[:, None]
        # the three quadrature sums  sum_q ws[c,q] * (.)_a(q) * (.)_b(q)  written as batched matmuls
        nc, nq, na, ng = gphi.shape
        G = np.transpose(gphi, (0, 1, 3, 2)).reshape(nc, nq * ng, na)  # [(q,g), a]
        K = np.matmul(np.transpose(G * np.repeat(ws, ng, axis=1)[:, :, None], (0, 2, 1)), G)
        M = np.matmul(np.transpose(ws[:, :, None] * phi[None], (0, 2, 1)), phi)
        D = np.matmul(np.transpose((ws * np.exp(pq))[:, :, None] * phi[None], (0, 2, 1)), phi)
        n = self.nld
        A = np.empty((K.shape[0], 2 * n, 2 * n))
        A[:, :n, :n] = alpha * K
        A[:, :n, n:] = M
        A[:, n:, :n] = M
        A[:, n:, n:] = -D
        return A

    # ---------------------------------------------------------------- global assembly
    def assemble_residual(self, x, xk, alpha):
        """SNESProblem.F (src/lvpp/problem.py:54-67): assemble_vector, lifting, set_bc."""
        Fe = self.element_residual(x, xk, alpha)
        b = np.bincount(self.cell_dofs.ravel(), weights=Fe.ravel(), minlength=self.num_rows)
        # apply_lifting(b, [a], bcs, x0=[x], scale=-1):  b -= scale * A_e (g - x0) over BC columns
        touched = np.flatnonzero(self.is_bc[self.cell_dofs].any(axis=1))
        Ae = self.element_jacobian(x, alpha, cells=touched)
        gmx = np.where(self.is_bc, self.bc_values - x, 0.0)[self.cell_dofs[touched]]  # [Ct, n]
        lift = np.einsum("cij,cj->ci", Ae, gmx)
        b += np.bincount(self.cell_dofs[touched].ravel(), weights=lift.ravel(), minlength=self.num_rows)
        # set_bc(b, bcs, x0=x, scale=-1): b[d] = -(g[d] - x[d])
        b[self.bc_dofs] = -(self.bc_values[self.bc_dofs] - x[self.bc_dofs])
        return b

    def assemble_jacobian_values(self, x, alpha):
        """SNESProblem.J (src/lvpp/problem.py:69-77): CSR values on the create_matrix pattern."""
        Ae = self.element_jacobian(x, alpha)
        bcl = self.is_bc[self.cell_dofs]  # [C, n]
        Ae = np.where(bcl[:, :, None] | bcl[:, None, :], 0.0, Ae)
        vals = np.bincount(self.cell_to_nnz.ravel(), weights=Ae.ravel(), minlength=self.nnz)
        # unit diagonal on BC rows
        diag_pos = self._diag_positions()
        vals[diag_pos[self.bc_dofs]] = 1.0
        return vals

    def _diag_positions(self):
        if not hasattr(self, "_diag"):
            rows = self.csr_rows
            d = np.flatnonzero(rows == self.indices)
            self._diag = np.empty(self.num_rows, dtype=np.int64)
            self._diag[rows[d]] = d
        return self._diag

    def jacobian(self, x, alpha):
        vals = self.assemble_jacobian_values(x, alpha)
        return sp.csr_matrix((vals, self.indices, self.indptr), shape=(self.num_rows, self.num_rows))

    # ---------------------------------------------------------------- observables
    def observables(self, x, xk, alpha):
        """The six scalar forms of obstacle_pg.py:145-152, in the order they are evaluated at
        :196-201: energy, complementarity (signed), feasibility, dual feasibility,
        H1-increment^2, latent-increment^2."""
        ul, pl = self._local(x)
        ukl, pkl = self._local(xk)
        phi, w, s = self.phi_tab, self.qwts, self.scale
        ws = w[None, :] * s[:, None]
        uq, pq, ukq, pkq = ul @ phi.T, pl @ phi.T, ukl @ phi.T, pkl @ phi.T
        gu = np.einsum("ca,cqag->cqg", ul, self.gphi)
        gd = np.einsum("ca,cqag->cqg", ul - ukl, self.gphi)
        energy = np.sum(ws * (0.5 * np.sum(gu * gu, axis=2) - self.f * uq))
        comp = np.sum(ws * ((pkq - pq) / alpha * uq))
        feas = np.sum(ws * np.where(uq < 0, -uq, 0.0))
        dual = np.sum(ws * np.where(pkq < pq, (pq - pkq) / alpha, 0.0))
        h1 = np.sum(ws * (np.sum(gd * gd, axis=2) + (uq - ukq) ** 2))
        l2 = np.sum(ws * (np.exp(pq) - np.exp(pkq)) ** 2)
        return np.array([energy, comp, feas, dual, h1, l2])
