"""The other LVPP formulations of SURVEY.md section 8a (rows a13-a18), assembled cell by cell (oracle;
TEST INFRASTRUCTURE, see oracle/__init__.py -- nothing under proximalgalerkin_b200/ may import this).

Literal numpy restatements of what dolfinx + FFCx evaluate for

* the gradient-constraint forms, examples/06_gradient_constraints/gradient_constraint_dolfinx.py:34-107
  (u in P2, psi in (P1)^2, quadrature degree 10, Hellinger map psi / sqrt(1 + |psi|^2)),
* the multiphase Cahn-Hilliard forms, examples/04_multiphase/multiphase_dolfinx.py:32-90
  (u, z, psi in (P1)^4 each, softmax latent map, crossed-diagonal mesh, no Dirichlet data),
* the Signorini contact forms, examples/02_signorini/signorini_dolfinx.py:199-291
  (u in (P1)^3 on the tetrahedra, psi in P1 on the contact facets, exp latent map),

with the Dirichlet conventions of src/lvpp/problem.py:54-77 / dolfinx NonlinearProblem: residual =
assemble_vector + apply_lifting(x0 = x, scale -1) + set_bc(x, -1); Jacobian = assemble_matrix with
bcs (rows and columns zeroed, unit diagonal); sparsity = union over integration entities of
dofs x dofs (create_matrix).  Jacobians are the hand-written Gateaux derivatives of the residuals
(``ufl.derivative``); tests/test_oracle_forms.py checks each against finite differences of the residual.

Parity unpinned: the reference ships no golden vectors for these forms.
"""
import numpy as np
import scipy.sparse as sp

from . import elements
from .mesh import exterior_facets
from .quadrature import make_quadrature


class Integral:
    """One integration block: entities (cells or facets) with their mixed dof lists."""

    def __init__(self, dofs):
        self.dofs = np.ascontiguousarray(dofs, dtype=np.int64)  # [E, n]


class MixedOracle:
    """Shared assembly plumbing.  Subclasses define ``integrals`` (list of Integral) and
    ``element_residual(k, x)`` / ``element_jacobian(k, x, entities=None)`` for block k."""

    def _finish_setup(self, num_rows, bc_dofs, bc_values):
        self.num_rows = int(num_rows)
        self.bc_dofs = np.sort(np.asarray(bc_dofs, dtype=np.int64))
        self.bc_values = np.zeros(self.num_rows)
        self.bc_values[np.asarray(bc_dofs, dtype=np.int64)] = bc_values
        self.is_bc = np.zeros(self.num_rows, dtype=bool)
        self.is_bc[self.bc_dofs] = True
        # create_matrix: union over all integrals of dofs x dofs, columns sorted
        keys = []
        for itg in self.integrals:
            cd = itg.dofs
            n = cd.shape[1]
            rows = np.repeat(cd, n, axis=1).ravel()
            cols = np.tile(cd, (1, n)).ravel()
            keys.append(rows * self.num_rows + cols)
        uniq, inv = np.unique(np.concatenate(keys), return_inverse=True)
        self.csr_rows = (uniq // self.num_rows).astype(np.int64)
        self.indices = (uniq % self.num_rows).astype(np.int32)
        self.indptr = np.zeros(self.num_rows + 1, dtype=np.int64)
        np.add.at(self.indptr, self.csr_rows + 1, 1)
        self.indptr = np.cumsum(self.indptr)
        self.nnz = uniq.size
        off = 0
        for itg in self.integrals:
            E, n = itg.dofs.shape
            itg.to_nnz = inv[off : off + E * n * n].reshape(E, n, n)
            off += E * n * n
        d = np.flatnonzero(self.csr_rows == self.indices)
        self.diag_pos = np.full(self.num_rows, -1, dtype=np.int64)
        self.diag_pos[self.csr_rows[d]] = d

    def assemble_residual(self, x):
        b = np.zeros(self.num_rows)
        for k, itg in enumerate(self.integrals):
            Fe = self.element_residual(k, x)
            b += np.bincount(itg.dofs.ravel(), weights=Fe.ravel(), minlength=self.num_rows)
            # apply_lifting(b, [a], bcs, x0=[x], scale=-1): b += A_e (g - x) over the Dirichlet columns
            touched = np.flatnonzero(self.is_bc[itg.dofs].any(axis=1))
            if touched.size:
                Ae = self.element_jacobian(k, x, entities=touched)
                gmx = np.where(self.is_bc, self.bc_values - x, 0.0)[itg.dofs[touched]]
                lift = np.einsum("cij,cj->ci", Ae, gmx)
                b += np.bincount(itg.dofs[touched].ravel(), weights=lift.ravel(), minlength=self.num_rows)
        b[self.bc_dofs] = -(self.bc_values[self.bc_dofs] - x[self.bc_dofs])  # set_bc(b, bcs, x, -1)
        return b

    def assemble_jacobian_values(self, x):
        vals = np.zeros(self.nnz)
        for k, itg in enumerate(self.integrals):
            Ae = self.element_jacobian(k, x)
            bcl = self.is_bc[itg.dofs]
            Ae = np.where(bcl[:, :, None] | bcl[:, None, :], 0.0, Ae)
            vals += np.bincount(itg.to_nnz.ravel(), weights=Ae.ravel(), minlength=self.nnz)
        vals[self.diag_pos[self.bc_dofs]] = 1.0
        return vals

    def jacobian(self, x):
        return sp.csr_matrix((self.assemble_jacobian_values(x), self.indices, self.indptr),
                             shape=(self.num_rows, self.num_rows))


def _cell_tables(mesh, degree, qpts):
    phi, dphi = elements.tabulate(degree, qpts)
    detJ, Jinv = elements.geometry(mesh)
    gphi = np.einsum("qad,cdg->cqag", dphi, Jinv)  # physical gradients [C, nq, nld, gdim]
    return phi, gphi, np.abs(detJ)


# ====================================================================================================
class GradientConstraintOracle(MixedOracle):
    """gradient_constraint_dolfinx.py:34-107.  dofs: u_n -> n (P2 nodes), psi_{v,c} -> N2 + 2 v + c."""

    def __init__(self, mesh, quadrature_degree=10, scheme="default",
                 phi_func=lambda x: 0.1 + 0.2 * x[0] + 0.4 * x[1],
                 f_func=lambda x: 15.0 * np.sin(np.pi * x[0]) * np.sin(np.pi * x[0])):
        self.mesh = mesh
        self.qpts, self.qwts = make_quadrature(mesh.cell_name, quadrature_degree, scheme)
        self.cell_nodes, self.N2, self.node_coords = elements.build_nodes(mesh, 2)
        self.N1 = mesh.num_vertices
        self.phi2, self.g2, self.scale = _cell_tables(mesh, 2, self.qpts)
        self.phi1, _, _ = _cell_tables(mesh, 1, self.qpts)
        gd = mesh.gdim
        self.gd = gd
        self.dof_u = np.arange(self.N2, dtype=np.int64)
        self.dof_psi = self.N2 + gd * np.arange(self.N1, dtype=np.int64)[:, None] + np.arange(gd)[None, :]  # [N1, gd]
        cu = self.dof_u[self.cell_nodes]  # [C, 6]
        cp = self.dof_psi[mesh.cells].reshape(mesh.num_cells, -1)  # [C, 3*gd], vertex-major, component fastest
        self.integrals = [Integral(np.concatenate([cu, cp], axis=1))]
        self.n2 = cu.shape[1]
        self.n1 = mesh.cells.shape[1]
        # phi, f interpolated into the P2 space (:56-62)
        self.phi_h = phi_func(self.node_coords.T)
        self.f_h = f_func(self.node_coords.T)
        self.alpha = 1.0
        self.w0 = np.zeros(self.N2 + gd * self.N1)
        # u = 0 on the boundary: every P2 node on an exterior facet (:64-70,109-111)
        ef = exterior_facets(mesh)
        bnodes = set(np.unique(ef).tolist())
        vpair = {}
        nv = mesh.tdim + 1
        for e, (a, b) in enumerate(elements.EDGES[mesh.tdim]):
            lo = np.minimum(mesh.cells[:, a], mesh.cells[:, b])
            hi = np.maximum(mesh.cells[:, a], mesh.cells[:, b])
            for l, h, n in zip(lo, hi, self.cell_nodes[:, nv + e]):
                vpair[(int(l), int(h))] = int(n)
        for fct in ef:
            for i in range(len(fct)):
                for j in range(i + 1, len(fct)):
                    bnodes.add(vpair[(int(min(fct[i], fct[j])), int(max(fct[i], fct[j])))])
        self.bc_nodes = np.array(sorted(bnodes), dtype=np.int64)
        self._finish_setup(self.N2 + gd * self.N1, self.dof_u[self.bc_nodes], 0.0)

    def _local(self, x, sel=slice(None)):
        ul = x[self.dof_u[self.cell_nodes[sel]]]  # [C, 6]
        pl = x[self.dof_psi[self.mesh.cells[sel]]]  # [C, 3, gd]
        return ul, pl

    def element_residual(self, k, x):
        ul, pl = self._local(x)
        _, p0l = self._local(self.w0)
        ws = self.qwts[None, :] * self.scale[:, None]
        gu = np.einsum("ca,cqag->cqg", ul, self.g2)  # grad u at q
        psi = np.einsum("qv,cvg->cqg", self.phi1, pl)
        psi0 = np.einsum("qv,cvg->cqg", self.phi1, p0l)
        fq = self.f_h[self.cell_nodes] @ self.phi2.T
        phq = self.phi_h[self.cell_nodes] @ self.phi2.T
        a = self.alpha
        Fu = np.einsum("cq,cqg,cqag->ca", ws, a * gu + psi - psi0, self.g2) - a * np.einsum("cq,qa->ca", ws * fq, self.phi2)
        s = np.sqrt(1.0 + np.sum(psi * psi, axis=2))
        Fp = np.einsum("cq,cqg,qv->cvg", ws, gu - (phq / s)[:, :, None] * psi, self.phi1)
        return np.concatenate([Fu, Fp.reshape(Fp.shape[0], -1)], axis=1)

    def element_jacobian(self, k, x, entities=None):
        sel = slice(None) if entities is None else entities
        ul, pl = self._local(x, sel)
        g2, scale = self.g2[sel], self.scale[sel]
        ws = self.qwts[None, :] * scale[:, None]
        psi = np.einsum("qv,cvg->cqg", self.phi1, pl)
        phq = self.phi_h[self.cell_nodes[sel]] @ self.phi2.T
        s = np.sqrt(1.0 + np.sum(psi * psi, axis=2))
        n2, n1, gd = self.n2, self.n1, self.gd
        C = ws.shape[0]
        A = np.zeros((C, n2 + n1 * gd, n2 + n1 * gd))
        A[:, :n2, :n2] = self.alpha * np.einsum("cq,cqag,cqbg->cab", ws, g2, g2)
        B = np.einsum("cq,cqag,qv->cavg", ws, g2, self.phi1).reshape(C, n2, n1 * gd)  # d F_u[a] / d psi_{v,g}
        A[:, :n2, n2:] = B
        A[:, n2:, :n2] = np.transpose(B, (0, 2, 1))
        # d/dpsi [phi psi / s] = phi (I / s - psi psi^T / s^3)
        T = (phq / s)[:, :, None, None] * np.eye(gd)[None, None] - (phq / s**3)[:, :, None, None] * psi[:, :, :, None] * psi[:, :, None, :]
        Dpp = np.einsum("cq,qv,qw,cqgh->cvgwh", ws, self.phi1, self.phi1, T).reshape(C, n1 * gd, n1 * gd)
        A[:, n2:, n2:] = -Dpp
        return A

    def l2_increment_sq(self, x, x0):
        """assemble_scalar of dot(u - u0, u - u0) * dx (:166-168)."""
        d = (x - x0)[self.dof_u[self.cell_nodes]] @ self.phi2.T
        return float(np.sum(self.qwts[None, :] * self.scale[:, None] * d * d))


# ====================================================================================================
class MultiphaseOracle(MixedOracle):
    """multiphase_dolfinx.py:32-90.  dofs at vertex n: u_m -> 12 n + m, z_m -> 12 n + 4 + m,
    psi_m -> 12 n + 8 + m.  Test functions (v, y, w) pair with (u, z, psi) slots: EQ2 is tested with v
    (u rows), EQ1 with y (z rows), EQ3 with w (psi rows) (:46,64-90).  The nonlinear term is integrated
    with the rule passed in (UFL's estimated degree for it is 7 [3P-mem]); all other terms are
    polynomials of degree <= 2."""

    NS = 4

    def __init__(self, mesh, quadrature_degree=7, scheme="default", tau=1e-5, eps0=1e-9):
        self.mesh = mesh
        self.qpts, self.qwts = make_quadrature(mesh.cell_name, quadrature_degree, scheme)
        self.phi, self.gphi, self.scale = _cell_tables(mesh, 1, self.qpts)
        self.N = mesh.num_vertices
        ns = self.NS
        self.tau, self.eps0, self.alpha = float(tau), float(eps0), 1.0
        # epsilon = 2 h, h = 2 * circumradius (:52-53)
        self.eps_cell = 2.0 * 2.0 * circumradius(mesh)
        base = 3 * ns * np.arange(self.N, dtype=np.int64)
        self.dof = {"u": base[:, None] + np.arange(ns), "z": base[:, None] + ns + np.arange(ns),
                    "psi": base[:, None] + 2 * ns + np.arange(ns)}  # [N, ns]
        cells = mesh.cells
        nv = cells.shape[1]
        # local order: node-major, (u0..3, z0..3, psi0..3) per node
        self.integrals = [Integral((base[cells][:, :, None] + np.arange(3 * ns)[None, None, :]).reshape(cells.shape[0], -1))]
        self.nv = nv
        self.u_prev = np.zeros((self.N, ns))
        self.lvpp_old = np.zeros(3 * ns * self.N)
        self._finish_setup(3 * ns * self.N, np.zeros(0, dtype=np.int64), 0.0)

    def _fields(self, x, sel=slice(None)):
        loc = x.reshape(self.N, 3, self.NS)[self.mesh.cells[sel]]  # [C, nv, 3, ns]
        return loc[:, :, 0, :], loc[:, :, 1, :], loc[:, :, 2, :]

    def element_residual(self, k, x):
        ul, zl, pl = self._fields(x)
        _, _, pol = self._fields(self.lvpp_old)
        upl = self.u_prev[self.mesh.cells]
        phi, g, ws = self.phi, self.gphi, self.qwts[None, :] * self.scale[:, None]
        a, tau, e2 = self.alpha, self.tau, self.eps_cell**2
        q = lambda loc: np.einsum("qv,cvm->cqm", phi, loc)
        uq, zq, pq, poq, upq = q(ul), q(zl), q(pl), q(pol), q(upl)
        gu = np.einsum("cvm,cqvg->cqmg", ul, g)
        gz = np.einsum("cvm,cqvg->cqmg", zl, g)
        # EQ1, test y (z rows)
        R_z = np.einsum("cq,cqm,qv->cvm", ws, a * zq - 2.0 * a * uq + pq - poq - a, phi)
        R_z += a * e2[:, None, None] * np.einsum("cq,cqmg,cqvg->cvm", ws, gu, g)
        # EQ2, test v (u rows)
        R_u = np.einsum("cq,cqm,qv->cvm", ws, uq - upq, phi) - tau * np.einsum("cq,cqmg,cqvg->cvm", ws, gz, g)
        # EQ3, test w (psi rows)
        e = np.exp(pq)
        sm = e / np.sum(e, axis=2, keepdims=True)
        R_p = np.einsum("cq,cqm,qv->cvm", ws, uq - sm - self.eps0 * pq, phi)
        return np.stack([R_u, R_z, R_p], axis=2).reshape(ul.shape[0], -1)

    def element_jacobian(self, k, x, entities=None):
        sel = slice(None) if entities is None else entities
        _, _, pl = self._fields(x, sel)
        phi, g = self.phi, self.gphi[sel]
        ws = self.qwts[None, :] * self.scale[sel][:, None]
        a, tau, e2 = self.alpha, self.tau, self.eps_cell[sel] ** 2
        ns, nv = self.NS, self.nv
        M = np.einsum("cq,qv,qw->cvw", ws, phi, phi)
        K = np.einsum("cq,cqvg,cqwg->cvw", ws, g, g)
        pq = np.einsum("qv,cvm->cqm", phi, pl)
        e = np.exp(pq)
        sm = e / np.sum(e, axis=2, keepdims=True)
        S = sm[:, :, :, None] * np.eye(ns)[None, None] - sm[:, :, :, None] * sm[:, :, None, :]  # d softmax_m / d psi_n
        N = np.einsum("cq,qv,qw,cqmn->cvmwn", ws, phi, phi, S)
        C = M.shape[0]
        A = np.zeros((C, nv, 3, ns, nv, 3, ns))
        I = np.eye(ns)
        mm = M[:, :, None, :, None] * I[None, None, :, None, :]  # [C, v, m, w, n]
        kk = K[:, :, None, :, None] * I[None, None, :, None, :]
        # u rows (EQ2): u.v - tau grad z : grad v
        A[:, :, 0, :, :, 0, :] = mm
        A[:, :, 0, :, :, 1, :] = -tau * kk
        # z rows (EQ1): alpha z.y + eps^2 alpha grad u : grad y - 2 alpha u.y + psi.y
        A[:, :, 1, :, :, 0, :] = a * e2[:, None, None, None, None] * kk - 2.0 * a * mm
        A[:, :, 1, :, :, 1, :] = a * mm
        A[:, :, 1, :, :, 2, :] = mm
        # psi rows (EQ3): (u - softmax(psi)).w - eps0 psi.w
        A[:, :, 2, :, :, 0, :] = mm
        A[:, :, 2, :, :, 2, :] = -N - self.eps0 * mm
        return A.reshape(C, nv * 3 * ns, nv * 3 * ns)

    def l2_increment_sq(self, x, u_old):
        """assemble_scalar of dot(sol.sub(0) - u_old, .) * dx (:166-169); u_old is [N, ns]."""
        d = x.reshape(self.N, 3, self.NS)[:, 0, :] - u_old
        dq = np.einsum("qv,cvm->cqm", self.phi, d[self.mesh.cells])
        return float(np.sum(self.qwts[None, :, None] * self.scale[:, None, None] * dq * dq))


def circumradius(mesh):
    """ufl.Circumradius of affine simplices."""
    x = mesh.coords[mesh.cells]
    if mesh.tdim == 2:
        a = np.linalg.norm(x[:, 1] - x[:, 2], axis=1)
        b = np.linalg.norm(x[:, 0] - x[:, 2], axis=1)
        c = np.linalg.norm(x[:, 0] - x[:, 1], axis=1)
        area = 0.5 * np.abs(np.linalg.det(np.stack([x[:, 1] - x[:, 0], x[:, 2] - x[:, 0]], axis=1)))
        return a * b * c / (4.0 * area)
    # tetrahedron: R = sqrt((aA + bB + cC)(aA + bB - cC)(aA - bB + cC)(-aA + bB + cC)) / (24 V)
    e = lambda i, j: np.linalg.norm(x[:, i] - x[:, j], axis=1)
    la, lb, lc = e(0, 1) * e(2, 3), e(0, 2) * e(1, 3), e(0, 3) * e(1, 2)
    vol = np.abs(np.linalg.det(x[:, 1:] - x[:, :1])) / 6.0
    return np.sqrt((la + lb + lc) * (la + lb - lc) * (la - lb + lc) * (-la + lb + lc)) / (24.0 * vol)


# ====================================================================================================
class SignoriniOracle(MixedOracle):
    """signorini_dolfinx.py:199-291 on a tetrahedral mesh.  dofs: u_{n,c} -> 3 n + c, psi on the
    contact-facet vertices (submesh numbering = ascending mesh vertex number) -> 3 N + s.
    Contact boundary: exterior facets with all vertices on x_last = ``contact_coord``; Dirichlet
    boundary u = (0, 0, disp): facets on x_last = ``disp_coord``."""

    def __init__(self, mesh, E=2.0e4, nu=0.3, gap=0.0, disp=-0.25, quadrature_degree=4, scheme="default",
                 contact_coord=None, disp_coord=None, tol=1e-12):
        self.mesh = mesh
        gd = mesh.gdim
        self.gd = gd
        self.N = mesh.num_vertices
        self.mu = E / (2.0 * (1.0 + nu))
        self.lmbda = E * nu / ((1.0 + nu) * (1.0 - 2.0 * nu))
        self.gap, self.disp, self.alpha = float(gap), float(disp), 1.0
        last = mesh.coords[:, gd - 1]
        contact_coord = last.min() if contact_coord is None else contact_coord
        disp_coord = last.max() if disp_coord is None else disp_coord
        ef = exterior_facets(mesh)
        on = lambda c: np.all(np.abs(last[ef] - c) < tol, axis=1)
        self.contact_facets = ef[on(contact_coord)]  # sorted vertex tuples
        bc_vertices = np.unique(ef[on(disp_coord)])
        self.sub_vertices = np.unique(self.contact_facets)  # submesh vertex s <-> mesh vertex
        self.NS = self.sub_vertices.size
        sub_of = np.full(self.N, -1, dtype=np.int64)
        sub_of[self.sub_vertices] = np.arange(self.NS)
        self.dof_u = gd * np.arange(self.N, dtype=np.int64)[:, None] + np.arange(gd)[None, :]
        self.dof_psi = gd * self.N + np.arange(self.NS, dtype=np.int64)
        # volume integral: dx with the estimated degree (grad P1 . grad P1 is constant per cell)
        self.qpts, self.qwts = make_quadrature(mesh.cell_name, 1, scheme)
        _, self.gphi, self.scale = _cell_tables(mesh, 1, self.qpts)
        self.gphi = self.gphi[:, 0]  # constant gradients [C, nv, gd]
        self.vol = self.scale * np.sum(self.qwts)
        cells = mesh.cells
        cd = self.dof_u[cells].reshape(cells.shape[0], -1)  # vertex-major, component fastest
        # facet integral: ds(contact) with quadrature degree 4 (:211-218)
        fname = {3: "triangle", 2: "interval"}[gd]
        if gd != 3:
            raise NotImplementedError("tetrahedral meshes only")
        self.fq, self.fw = make_quadrature(fname, quadrature_degree, scheme)
        self.fphi, _ = elements.tabulate(1, self.fq)  # [nq, 3]
        fx = mesh.coords[self.contact_facets]  # [F, 3, gd]
        self.farea2 = np.linalg.norm(np.cross(fx[:, 1] - fx[:, 0], fx[:, 2] - fx[:, 0]), axis=1)  # |J| of the facet map
        self.fxq_last = np.einsum("qv,fv->fq", self.fphi, fx[:, :, gd - 1])
        fu = self.dof_u[self.contact_facets].reshape(self.contact_facets.shape[0], -1)
        fp = self.dof_psi[sub_of[self.contact_facets]]
        self.integrals = [Integral(cd), Integral(np.concatenate([fu, fp], axis=1))]
        self.sub_of = sub_of
        self.psi_k = np.zeros(self.NS)
        self.n_g = np.zeros(gd)
        self.n_g[-1] = -1.0
        bc_dofs = self.dof_u[bc_vertices].ravel()
        bc_vals = np.tile(np.array([0.0] * (gd - 1) + [self.disp]), bc_vertices.size)
        self._finish_setup(gd * self.N + self.NS, bc_dofs, bc_vals)

    def _elasticity(self, sel=slice(None)):
        """Element stiffness of sigma(u) : eps(v), [C, 4*gd, 4*gd], vertex-major / component fastest."""
        g = self.gphi[sel]  # [C, nv, gd]
        vol = self.vol[sel]
        mu, lm = self.mu, self.lmbda
        # a(phi_b e_j, phi_a e_i) = vol * (mu (g_a.g_b delta_ij + g_a[j] g_b[i]) + lambda g_a[i] g_b[j])
        gg = np.einsum("cag,cbg->cab", g, g)
        I = np.eye(self.gd)
        A = mu * (gg[:, :, None, :, None] * I[None, None, :, None, :] + np.einsum("caj,cbi->caibj", g, g))
        A += lm * np.einsum("cai,cbj->caibj", g, g)
        A *= vol[:, None, None, None, None]
        n = g.shape[1] * self.gd
        return A.reshape(g.shape[0], n, n)

    def element_residual(self, k, x):
        if k == 0:
            ul = x[self.integrals[0].dofs]
            return self.alpha * np.einsum("cij,cj->ci", self._elasticity(), ul)  # f = 0 (:239)
        F = self.contact_facets
        ul = x[self.dof_u[F]]  # [F, 3, gd]
        pl = x[self.dof_psi[self.sub_of[F]]]  # [F, 3]
        pkl = self.psi_k[self.sub_of[F]]
        ws = self.fw[None, :] * self.farea2[:, None]
        un = np.einsum("qv,fvg,g->fq", self.fphi, ul, self.n_g)
        pq = pl @ self.fphi.T
        pkq = pkl @ self.fphi.T
        g = self.fxq_last - self.gap
        Ru = -np.einsum("fq,qv->fv", ws * (pq - pkq), self.fphi)[:, :, None] * self.n_g[None, None, :]
        Rp = np.einsum("fq,qv->fv", ws * (un + np.exp(pq) - g), self.fphi)
        return np.concatenate([Ru.reshape(Ru.shape[0], -1), Rp], axis=1)

    def element_jacobian(self, k, x, entities=None):
        sel = slice(None) if entities is None else entities
        if k == 0:
            return self.alpha * self._elasticity(sel)
        F = self.contact_facets[sel]
        pl = x[self.dof_psi[self.sub_of[F]]]
        ws = self.fw[None, :] * self.farea2[sel][:, None]
        pq = pl @ self.fphi.T
        M = np.einsum("fq,qv,qw->fvw", ws, self.fphi, self.fphi)
        D = np.einsum("fq,qv,qw->fvw", ws * np.exp(pq), self.fphi, self.fphi)
        nf, gd = F.shape[0], self.gd
        A = np.zeros((nf, 3 * gd + 3, 3 * gd + 3))
        Bup = -(M[:, :, None, :] * self.n_g[None, None, :, None]).reshape(nf, 3 * gd, 3)  # d R_u[(v,g)] / d psi_w
        A[:, : 3 * gd, 3 * gd :] = Bup
        A[:, 3 * gd :, : 3 * gd] = (M[:, :, :, None] * self.n_g[None, None, None, :]).reshape(nf, 3, 3 * gd)
        A[:, 3 * gd :, 3 * gd :] = D
        return A
