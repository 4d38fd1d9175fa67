#!/usr/bin/env python
"""CPU prototype (numpy / scipy) of the library's multigrid-preconditioned GMRES -- a RESEARCH TOOL for studying
iteration counts off the GPU: it mirrors csrc/multigrid.cu (2^d coordinate-box aggregation with Dirichlet nodes
aggregated separately, piecewise-constant prolongation with over-correction, Galerkin K/M/D per field, node-block
Jacobi sweeps with plain or Chebyshev-root dampings, dense coarsest solve, right-preconditioned restarted GMRES) on
the oracle's matrices, at sizes where everything fits a laptop.  Not product code, not a test oracle.

  python tools/mg_prototype.py --size 16 --outer 3                 # counts per Newton step: plain vs Chebyshev
  python tools/mg_prototype.py --size 16 --outer 3 --slabs 4       # aggregates confined to 4 z-slabs (the partition)
  python tools/mg_prototype.py --size 12 --spectrum                # extreme eigenvalues of Binv J along the solve
"""
import argparse
import sys
from pathlib import Path

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import lvpp_driver, mesh as omesh, obstacle as oobs  # noqa: E402


# ------------------------------------------------------------------------------------------------ operators
def scalar_blocks(orc, x):
    """Node-by-node K, M, D(psi) (unmasked CSR) from the oracle's element tensors."""
    Ae = orc.element_jacobian(x, 1.0)
    n = orc.nld
    cn = orc.cell_nodes
    rows = np.repeat(cn, n, axis=1).ravel()
    cols = np.tile(cn, (1, n)).ravel()
    N = orc.num_nodes

    def asm(block):
        return sp.coo_matrix((block.ravel(), (rows, cols)), shape=(N, N)).tocsr()

    return asm(Ae[:, :n, :n]), asm(Ae[:, :n, n:]), asm(-Ae[:, n:, n:])


class Level:
    def __init__(self, K, M, D, bc, box, slab):
        self.K, self.M, self.D, self.bc, self.box, self.slab = K, M, D, bc, box, slab
        self.N = K.shape[0]
        self.P = None

    def build(self, alpha):
        """Masked 2N x 2N operator (blocked [u; psi]) exactly as k_block_op applies it, and the node-block inverses."""
        N, bc = self.N, self.bc
        free = sp.diags((~bc).astype(float))
        ident_bc = sp.diags(bc.astype(float))
        Kuu = free @ (alpha * self.K) @ free + ident_bc       # Dirichlet rows = identity, Dirichlet columns masked
        Kup = free @ self.M                                    # u rows of Dirichlet nodes have no psi coupling
        Kpu = self.M @ free                                    # Dirichlet columns of u masked in the psi rows
        self.J = sp.bmat([[Kuu, Kup], [Kpu, -self.D]], format="csr")
        a, m, d = alpha * self.K.diagonal(), self.M.diagonal(), self.D.diagonal()
        det = -a * d - m * m
        B00, B01, B11 = -d / det, -m / det, a / det
        B00 = np.where(bc, 1.0, B00)
        B01 = np.where(bc, 0.0, B01)
        B11 = np.where(bc, -1.0 / d, B11)
        self.Binv = sp.bmat([[sp.diags(B00), sp.diags(B01)], [sp.diags(B01), sp.diags(B11)]], format="csr")
        self.is_bc2 = np.concatenate([bc, np.zeros(N, dtype=bool)])

    def coarsen(self, smooth=0.0):
        half = self.box >> 1
        key = np.stack([self.slab, self.bc.astype(np.int64), half[:, 2], half[:, 1], half[:, 0]], axis=1)
        uniq, agg = np.unique(key, axis=0, return_inverse=True)
        agg = agg.ravel()
        Nc = uniq.shape[0]
        P = sp.csr_matrix((np.ones(self.N), (np.arange(self.N), agg)), shape=(self.N, Nc))
        if smooth > 0.0:
            # smoothed aggregation (experiment): one damped Jacobi step on the stiffness block applied to the tentative
            # prolongator, free nodes only -- denser coarse operators, better interpolation
            free = sp.diags((~self.bc).astype(float))
            Kf = free @ self.K @ free
            dinv = sp.diags(np.where(self.bc, 0.0, 1.0 / self.K.diagonal()))
            P = (P - smooth * (dinv @ (Kf @ P))).tocsr()
        self.P = P
        c = Level(P.T @ self.K @ P, P.T @ self.M @ P, P.T @ self.D @ P, uniq[:, 1].astype(bool), uniq[:, [4, 3, 2]], uniq[:, 0])
        return c


class Multigrid:
    def __init__(self, orc, x, alpha, slabs=1, over=1.8, coarse_target=96, max_levels=12, smooth=0.0, slab_levels=99):
        K, M, D = scalar_blocks(orc, x)
        coords = orc.node_coords
        h0 = np.min(np.diff(np.unique(np.round(coords[:, 0], 12))))
        box = np.floor((coords - coords.min(axis=0)) / h0 + 0.25).astype(np.int64)
        if box.shape[1] == 2:
            box = np.concatenate([box, np.zeros((box.shape[0], 1), dtype=np.int64)], axis=1)
        bc = np.zeros(orc.num_nodes, dtype=bool)
        bc[orc.bc_nodes] = True
        # the partition: aggregates never cross slab boundaries (slabs along the last axis, owned planes split evenly);
        # every slab aggregates from its own first plane, as every rank does
        ax = orc.mesh.gdim - 1
        nplanes = box[:, ax].max() + 1
        slab = np.minimum(box[:, ax] * slabs // nplanes, slabs - 1)
        if slabs > 1:
            first = np.array([box[slab == s, ax].min() for s in range(slabs)])
            box = box.copy()
            box[:, ax] -= first[slab]
        self.levels = [Level(K, M, D, bc, box, slab)]
        while self.levels[-1].N > coarse_target and len(self.levels) < max_levels:
            if len(self.levels) - 1 == slab_levels and slabs > 1:
                # the replicated part of the hierarchy: from here down aggregates ignore the rank
                # boundaries; box coordinates go back to the global grid of this level
                L = self.levels[-1]
                shift = np.array([L.box[L.slab == s_, ax].max() + 1 for s_ in range(slabs)])
                off = np.concatenate([[0], np.cumsum(shift)[:-1]])
                L.box = L.box.copy()
                L.box[:, ax] += off[L.slab]
                L.slab = np.zeros_like(L.slab)
            c = self.levels[-1].coarsen(smooth)
            if c.N >= self.levels[-1].N:
                self.levels[-1].P = None
                break
            self.levels.append(c)
        self.over = over
        self.set_alpha(alpha)

    def set_alpha(self, alpha):
        for L in self.levels:
            L.build(alpha)
        Lc = self.levels[-1]
        self.coarse = np.linalg.inv(Lc.J.toarray())

    def lambda_max(self, L, exact=False, its=30):
        A = L.Binv @ L.J
        if exact and L.N > 40:
            return float(np.max(np.abs(spla.eigs(A, k=1, which="LM", return_eigenvectors=False, maxiter=5000, tol=1e-6))))
        if exact:
            return float(np.max(np.abs(np.linalg.eigvals(A.toarray()))))
        rng = np.random.default_rng(0)
        v = rng.random(2 * L.N) - 0.5
        lam = 0.0
        for _ in range(its):
            v /= np.linalg.norm(v)
            w = v + A @ v
            w[L.is_bc2] = 0.0
            lam = np.linalg.norm(w) - 1.0
            v = w
        return lam

    LAMBDA_HISTORY = {}  # level index -> largest estimate so far (the library's estimate is monotone per handle)

    def set_smoother(self, cheb=10.0, margin=1.10, npre=2, npost=2, exact_lambda=False, monotone=True):
        self.npre, self.npost = npre, npost
        for li, L in enumerate(self.levels[:-1]):
            if not hasattr(L, "lam_raw"):
                L.lam_raw = self.lambda_max(L, exact=exact_lambda)
            lam = L.lam_raw
            if monotone:
                lam = max(lam, Multigrid.LAMBDA_HISTORY.get(li, 0.0))
                Multigrid.LAMBDA_HISTORY[li] = lam
            L.lam = lam

            def roots(m):
                if m == 0:
                    return []
                if cheb > 1.0:
                    b = margin * lam
                    a = b / cheb
                    return [1.0 / (0.5 * (b + a) + 0.5 * (b - a) * np.cos(np.pi * (2 * k + 1) / (2.0 * m))) for k in range(m)]
                return [min(1.0, 2.0 / (margin * lam))] * m

            L.om_pre, L.om_post = roots(npre), roots(npost)

    @staticmethod
    def _sweep(L, x, b, om):
        r = L.Binv @ (b - L.J @ x)
        w = np.where(L.is_bc2, 1.0, om)
        return x + w * r

    def cycle(self, b, l=0, gamma=1):
        L = self.levels[l]
        if l == len(self.levels) - 1:
            return self.coarse @ b
        x = np.zeros_like(b)
        for om in L.om_pre:
            x = self._sweep(L, x, b, om)
        r = b - L.J @ x if L.om_pre else b
        C = self.levels[l + 1]
        P2 = sp.block_diag([L.P, L.P], format="csr")
        rc = P2.T @ r
        rc[: C.N][C.bc] = 0.0
        xc = self.cycle(rc, l + 1, gamma)
        for _ in range(gamma - 1):  # W-cycle: second visit from the current coarse iterate
            xc = xc + self.cycle(rc - C.J @ xc, l + 1, gamma)
        e = self.over * (P2 @ xc)
        e[L.is_bc2] = 0.0
        x = x + e
        for om in L.om_post:
            x = self._sweep(L, x, b, om)
        return x


def gmres_right(J, prec, b, rtol=1e-12, restart=50, maxit=400):
    """Right-preconditioned restarted GMRES (modified Gram-Schmidt); returns (y, iterations)."""
    n = b.size
    y = np.zeros(n)
    bnorm = np.linalg.norm(b)
    total = 0
    while total < maxit:
        r = b - J @ y
        beta = np.linalg.norm(r)
        if beta <= rtol * bnorm:
            return y, total
        V = np.zeros((restart + 1, n))
        H = np.zeros((restart + 1, restart))
        V[0] = r / beta
        g = np.zeros(restart + 1)
        g[0] = beta
        cs, sn = np.zeros(restart), np.zeros(restart)
        k = 0
        for j in range(restart):
            w = J @ prec(V[j])
            for i in range(j + 1):
                H[i, j] = V[i] @ w
                w -= H[i, j] * V[i]
            H[j + 1, j] = np.linalg.norm(w)
            if H[j + 1, j] > 0:
                V[j + 1] = w / H[j + 1, j]
            for i in range(j):
                t = cs[i] * H[i, j] + sn[i] * H[i + 1, j]
                H[i + 1, j] = -sn[i] * H[i, j] + cs[i] * H[i + 1, j]
                H[i, j] = t
            den = np.hypot(H[j, j], H[j + 1, j])
            cs[j], sn[j] = H[j, j] / den, H[j + 1, j] / den
            H[j, j], H[j + 1, j] = den, 0.0
            g[j + 1] = -sn[j] * g[j]
            g[j] = cs[j] * g[j]
            total += 1
            k = j + 1
            if abs(g[j + 1]) <= rtol * bnorm or total >= maxit:
                break
        c = np.linalg.solve(np.triu(H[:k, :k]), g[:k])
        y = y + prec(V[:k].T @ c)
    return y, total


# ------------------------------------------------------------------------------------------------ driver
def newton_states(orc, outer):
    """(x, xk, alpha, F) at the start of every Newton step of the first `outer` proximal steps (exact LU Newton)."""
    x = np.zeros(orc.num_rows)
    xk = x.copy()
    alpha_k, alpha = 1, 1.0
    out = []
    for k in range(outer):
        alpha, alpha_k = lvpp_driver.alpha_schedule("double_exponential", k, alpha_k, 1e2, alpha_current=alpha)
        F = orc.assemble_residual(x, xk, alpha)
        f0 = np.linalg.norm(F)
        for it in range(50):
            out.append((k, it, x.copy(), xk.copy(), alpha, F.copy()))
            y = spla.splu(orc.jacobian(x, alpha).tocsc()).solve(F)
            x = x - y
            F = orc.assemble_residual(x, xk, alpha)
            if np.linalg.norm(F) <= 1e-6 * f0:
                break
        xk = x.copy()
    return out


def to_blocked(orc, v):
    return np.concatenate([v[orc.dof_u], v[orc.dof_psi]])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=12)
    ap.add_argument("--dim", type=int, default=3)
    ap.add_argument("--outer", type=int, default=3)
    ap.add_argument("--slabs", type=int, default=1)
    ap.add_argument("--over", type=float, default=1.8)
    ap.add_argument("--spectrum", action="store_true")
    ap.add_argument("--slab-levels", dest="slab_levels", type=int, default=99,
                    help="confine aggregates to the slabs only for the first L coarsenings (the replicated hierarchy below)")
    ap.add_argument("--smooth", type=float, default=0.0, help="smoothed-aggregation damping (0 = plain aggregation)")
    ap.add_argument("--configs", default="plain,cheb10", help="comma list: plain, chebR, wplain, wchebR (w = W-cycle)")
    args = ap.parse_args()
    n = args.size
    msh = omesh.box_kuhn(n, n, n) if args.dim == 3 else omesh.rectangle(n, n)
    orc = oobs.ObstacleOracle(msh)
    states = newton_states(orc, args.outer)
    print(f"# {args.dim}-D n={n}: {orc.num_rows} rows, {len(states)} Newton steps in {args.outer} proximal steps, slabs={args.slabs}")
    for (k, it, x, xk, alpha, F) in states:
        mg = Multigrid(orc, x, alpha, slabs=args.slabs, over=args.over, smooth=args.smooth, slab_levels=args.slab_levels)
        L0 = mg.levels[0]
        rhs = to_blocked(orc, F)
        line = f"outer {k} alpha {alpha:.3g} newton {it}: levels {[L.N for L in mg.levels]}"
        mg.set_smoother()
        line += f" lam0 {mg.levels[0].lam_raw:.2f}->{mg.levels[0].lam:.2f}"
        if args.spectrum:
            A = (L0.Binv @ L0.J).toarray() if L0.N <= 1500 else None
            if A is not None:
                ev = np.linalg.eigvals(A)
                line += (f" | eig(Binv J): re [{ev.real.min():.3f}, {ev.real.max():.3f}] max|im| {np.abs(ev.imag).max():.3f} "
                         f"#re<0 {int((ev.real < -1e-9).sum())} #|im|>1e-6 {int((np.abs(ev.imag) > 1e-6).sum())}")
        for cfg in args.configs.split(","):
            gamma = 2 if cfg.startswith("w") else 1
            name = cfg[1:] if cfg.startswith("w") else cfg
            cheb = float(name[4:]) if name.startswith("cheb") else 0.0
            mg.set_smoother(cheb=cheb)
            _, its = gmres_right(L0.J, lambda v: mg.cycle(v, 0, gamma), rhs)
            line += f" | {cfg}: {its}"
        print(line, flush=True)


if __name__ == "__main__":
    main()
