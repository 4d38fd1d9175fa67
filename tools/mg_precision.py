#!/usr/bin/env python
"""How much precision does the multigrid *cycle* need?  (CPU experiment on tools/mg_prototype.py.)

The library's cycle streams a packed {col u32, alpha K f32, M f32, D f32} record per slot (k_packed_op, 16 bytes) and
fp64 vectors; the Krylov operator, residuals and Gram-Schmidt stay fp64.  This tool rounds what the *cycle* reads
-- the stored alpha K, M, D of every level, the node-block inverses, optionally the cycle's vectors -- to a smaller
format and counts the Krylov iterations of right-preconditioned GMRES(50) to 1e-12 at the Newton states of the first
proximal steps (exact-LU Newton trajectory, like mg_prototype.py).

  python tools/mg_precision.py --size 20 --outer 4
"""
import argparse
import importlib.util
import sys
from pathlib import Path

import numpy as np
import scipy.sparse as sp

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import mesh as omesh, obstacle as oobs  # noqa: E402

spec = importlib.util.spec_from_file_location("mg_prototype", ROOT / "tools" / "mg_prototype.py")
mp = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mp)


def q_bf16(a):
    u = np.asarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000  # round to nearest even on the upper 16 bits
    return u.astype(np.uint32).view(np.float32).astype(np.float64)


def q_fp16(a):
    return np.clip(np.asarray(a, dtype=np.float64), -65504.0, 65504.0).astype(np.float16).astype(np.float64)


def q_fp32(a):
    return np.clip(np.asarray(a, dtype=np.float64), -3.0e38, 3.0e38).astype(np.float32).astype(np.float64)


QUANT = {"fp64": lambda a: np.asarray(a, dtype=np.float64), "fp32": q_fp32, "bf16": q_bf16, "fp16": q_fp16}


def qmat(A, q):
    A = A.tocsr().copy()
    A.data = q(A.data)
    return A


def quantised_cycle_operators(mg, alpha, q, qb):
    """Per level: the masked operator rebuilt from q(alpha K), q(M), q(D) -- what the packed records hold -- and the
    node-block inverses (computed in fp64 from the fp64 diagonals, as build_binv does) stored through qb."""
    out = []
    for L in mg.levels:
        N, bc = L.N, L.bc
        free = sp.diags((~bc).astype(float))
        ident_bc = sp.diags(bc.astype(float))
        aK, M, D = qmat(alpha * L.K, q), qmat(L.M, q), qmat(L.D, q)
        J = sp.bmat([[free @ aK @ free + ident_bc, free @ M], [M @ free, -D]], format="csr")
        out.append((J, qmat(L.Binv, qb)))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=16)
    ap.add_argument("--dim", type=int, default=3)
    ap.add_argument("--outer", type=int, default=4)
    ap.add_argument("--cheb", type=float, default=6.0)
    ap.add_argument("--formats", default="fp64,fp32,bf16,bf16+v32,fp16",
                    help="comma list; +v32 also rounds the cycle's vectors to fp32, +b32 only the block inverses, +fg uses flexible GMRES")
    args = ap.parse_args()
    n = args.size
    orc = oobs.ObstacleOracle(omesh.box_kuhn(n, n, n) if args.dim == 3 else omesh.rectangle(n, n))
    states = mp.newton_states(orc, args.outer)
    print(f"# {args.dim}-D n={n}: {orc.num_rows} rows, {len(states)} Newton steps in {args.outer} proximal steps, Chebyshev ratio {args.cheb}")
    totals = {f: 0 for f in args.formats.split(",")}
    for (k, it, x, xk, alpha, F) in states:
        mg = mp.Multigrid(orc, x, alpha)
        mg.set_smoother(cheb=args.cheb)
        J0 = mg.levels[0].J
        exact = [(L.J, L.Binv) for L in mg.levels]
        rhs = mp.to_blocked(orc, F)
        line = f"outer {k} alpha {alpha:.3g} newton {it}:"
        for fmt in args.formats.split(","):
            name, *mods = fmt.split("+")
            q = QUANT[name]
            qb = q_fp32 if (name != "fp64" or "b32" in mods) else QUANT["fp64"]
            for L, (J, B) in zip(mg.levels, quantised_cycle_operators(mg, alpha, q, qb)):
                L.J, L.Binv = J, B
            if "v32" in mods:
                prec = lambda v: q_fp32(cycle32(mg, q_fp32(v)))  # noqa: E731
            else:
                prec = lambda v: mg.cycle(v, 0, 1)  # noqa: E731
            _, its = (fgmres if "fg" in mods else mp.gmres_right)(J0, prec, rhs)
            for L, (J, B) in zip(mg.levels, exact):
                L.J, L.Binv = J, B
            totals[fmt] += its
            line += f" {fmt} {its}"
        print(line, flush=True)
    print("# total Krylov iterations: " + ", ".join(f"{f} {t}" for f, t in totals.items()))


def fgmres(J, prec, b, rtol=1e-12, restart=50, maxit=400):
    """Flexible GMRES (Saad): keeps Z_j = prec(V_j), so the preconditioner may change from one application to the
    next -- e.g. a cycle whose vectors are rounded to single precision.  Returns (y, iterations)."""
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
        Z = np.zeros((restart, n))
        H = np.zeros((restart + 1, restart))
        V[0] = r / beta
        g = np.zeros(restart + 1)
        g[0] = beta
        cs, sn = np.zeros(restart), np.zeros(restart)
        k = 0
        for j in range(restart):
            Z[j] = prec(V[j])
            w = J @ Z[j]
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
        y = y + Z[:k].T @ c
    return y, total


def cycle32(mg, b, l=0):
    """mg.cycle with every vector the cycle stores rounded to fp32 (accumulation inside a sweep stays fp64)."""
    L = mg.levels[l]
    if l == len(mg.levels) - 1:
        return q_fp32(mg.coarse @ b)
    x = np.zeros_like(b)
    for om in L.om_pre:
        x = q_fp32(mg._sweep(L, x, b, om))
    r = q_fp32(b - L.J @ x) if L.om_pre else b
    C = mg.levels[l + 1]
    P2 = sp.block_diag([L.P, L.P], format="csr")
    rc = q_fp32(P2.T @ r)
    rc[: C.N][C.bc] = 0.0
    xc = cycle32(mg, rc, l + 1)
    e = mg.over * (P2 @ xc)
    e[L.is_bc2] = 0.0
    x = q_fp32(x + e)
    for om in L.om_post:
        x = q_fp32(mg._sweep(L, x, b, om))
    return x


if __name__ == "__main__":
    main()
