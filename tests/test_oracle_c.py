"""The compiled part of the oracle (oracle/c/lvpp_cpu.c, C + OpenMP) against the numpy restatement it mirrors."""
import numpy as np
import pytest

from oracle import cpu_kernels as ck
from oracle import mesh as omesh
from oracle import obstacle as oobs


@pytest.mark.parametrize("kind,n,degree", [("tet", 6, 1), ("tri", 9, 1), ("tri", 7, 2), ("tet", 3, 2)])
def test_c_jacobian_matches_numpy_oracle(kind, n, degree):
    msh = omesh.box_kuhn(n, n, n) if kind == "tet" else omesh.rectangle(n, n)
    orc = oobs.ObstacleOracle(msh, degree=degree)
    rng = np.random.default_rng(4)
    x = 0.4 * rng.standard_normal(orc.num_rows)
    x[1::2] -= 5.0 * (rng.random(orc.num_rows // 2) > 0.5)  # exp(psi) over several orders of magnitude
    ja = ck.JacobianAssembler(orc)
    for alpha in (1.0, 7.5):
        v = ja.assemble(x, alpha)
        vo = orc.assemble_jacobian_values(x, alpha)
        assert np.abs(v - vo).max() <= 1e-13 * np.abs(vo).max()
    # a second assembly overwrites (J.zeroEntries) instead of accumulating
    assert np.array_equal(ja.assemble(x, 7.5), v)


def test_c_spmv_matches_scipy():
    orc = oobs.ObstacleOracle(omesh.box_kuhn(5, 5, 5))
    rng = np.random.default_rng(1)
    x = rng.standard_normal(orc.num_rows)
    J = orc.jacobian(0.1 * x, 2.0)
    y = ck.csr_spmv(J.indptr, J.indices, J.data, x)
    assert np.abs(y - J @ x).max() <= 1e-13 * np.abs(y).max()
    assert ck.num_threads() >= 1


@pytest.mark.parametrize("kind,n", [("tet", 8), ("tri", 24)])
def test_c_minres_matches_sparse_lu(kind, n):
    """The CPU Krylov baseline (C + OpenMP MINRES with the block-Jacobi / Schur-diagonal preconditioner of ex40.cpp:261-274
    reduced to its diagonals) solves the Newton system of the oracle to the accuracy of a sparse LU."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    orc = oobs.ObstacleOracle(omesh.box_kuhn(n, n, n) if kind == "tet" else omesh.rectangle(n, n))
    rng = np.random.default_rng(7)
    x = 0.2 * rng.standard_normal(orc.num_rows)
    x[1::2] -= 6.0 * (rng.random(orc.num_rows // 2) > 0.6)
    x[orc.bc_dofs] = 0.0
    vals = ck.JacobianAssembler(orc).assemble(x, 3.0).copy()
    pinv = ck.BlockJacobiDiagonal(orc)(vals)
    assert np.all(pinv > 0)
    rhs = rng.standard_normal(orc.num_rows)
    y, its, rn = ck.minres(orc.indptr, orc.indices, vals, pinv, rhs, rtol=1e-12)
    J = sp.csr_matrix((vals, orc.indices, orc.indptr), shape=(orc.num_rows, orc.num_rows))
    ye = spla.splu(J.tocsc()).solve(rhs)
    assert 0 < its < 5000
    assert np.linalg.norm(y - ye) <= 1e-9 * np.linalg.norm(ye)


def test_cpu_krylov_newton_sample_runs():
    r = ck.time_newton_steps_krylov(8, 2)
    assert r["value"] > 0 and len(r["minres_iterations"]) == 2 and r["threads"] >= 1
