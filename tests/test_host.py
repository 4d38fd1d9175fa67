"""Host-side logic of the product (no GPU): meshes and slab partitions, spaces, tables, option
parsing, the alpha schedule, and that the product never touches the oracle."""
import re
from pathlib import Path

import numpy as np
import pytest

import proximalgalerkin_b200 as lvpp
from oracle import elements as oelem
from oracle import mesh as omesh
from oracle import obstacle as oobs
from oracle import quadrature as oquad

ROOT = Path(__file__).resolve().parents[1]


def test_meshes_match_oracle_numbering():
    m, o = lvpp.mesh.create_box(3, 4, 5), omesh.box_kuhn(3, 4, 5)
    assert np.array_equal(m.cells, o.cells) and np.allclose(m.coords, o.coords, rtol=0, atol=1e-15)
    assert np.array_equal(m.boundary_vertices, omesh.boundary_vertices(o))
    m, o = lvpp.mesh.create_rectangle(6, 5), omesh.rectangle(6, 5)
    assert np.array_equal(m.cells, o.cells) and np.allclose(m.coords, o.coords, rtol=0, atol=1e-15)
    assert np.array_equal(m.boundary_vertices, omesh.boundary_vertices(o))
    g = lvpp.mesh.from_arrays(o.coords, o.cells)
    assert np.array_equal(g.boundary_vertices, omesh.boundary_vertices(o))


@pytest.mark.parametrize("make,shape", [(lvpp.mesh.create_box, (4, 3, 9)), (lvpp.mesh.create_rectangle, (5, 11))])
@pytest.mark.parametrize("nranks", [2, 3, 4])
def test_slab_partition(make, shape, nranks):
    whole = make(*shape)
    parts = [make(*shape, rank=r, nranks=nranks) for r in range(nranks)]
    # owned vertices partition the global numbering; coordinates agree with the global mesh
    owned = np.concatenate([p.global_vertex[: p.num_owned_vertices] for p in parts])
    assert np.array_equal(np.sort(owned), np.arange(whole.num_vertices))
    for p in parts:
        assert np.allclose(p.coords, whole.coords[p.global_vertex], rtol=0, atol=1e-15)
        assert np.array_equal(np.sort(p.global_vertex[p.boundary_vertices]),
                              np.intersect1d(whole.boundary_vertices, p.global_vertex))
    # owned cells partition the global cells (as global vertex tuples)
    gc = np.concatenate([p.global_vertex[p.cells[: p.num_owned_cells]] for p in parts])
    assert np.array_equal(np.sort(gc.view([("", gc.dtype)] * gc.shape[1]).ravel()),
                          np.sort(whole.cells.astype(np.int64).view([("", np.int64)] * gc.shape[1]).ravel()))
    # every cell incident to an owned vertex is present locally
    for p in parts:
        own_g = set(p.global_vertex[: p.num_owned_vertices].tolist())
        need = {tuple(c) for c in whole.cells.astype(np.int64).tolist() if own_g.intersection(c)}
        have = {tuple(c) for c in p.global_vertex[p.cells].tolist()}
        assert need <= have
    # halo lists: what r sends to s is what s receives from r, in the same order
    for p in parts:
        for nb, send in zip(p.halo.neighbors, p.halo.send):
            q = parts[nb]
            recv = q.halo.recv[q.halo.neighbors.index(p.rank)]
            assert np.all(send < p.num_owned_vertices) and np.all(recv >= q.num_owned_vertices)
            assert np.array_equal(p.global_vertex[send], q.global_vertex[recv])
        ghosts = np.concatenate(p.halo.recv) if p.halo.recv else np.zeros(0, dtype=np.int32)
        assert np.array_equal(np.sort(ghosts), np.arange(p.num_owned_vertices, p.num_vertices))


def test_tables_and_spaces_match_oracle():
    for cell, tdim in (("triangle", 2), ("tetrahedron", 3)):
        for deg in (1, 2, 4, 6, 9):
            p, w = lvpp.quadrature.make_quadrature(cell, deg)
            po, wo = oquad.make_quadrature(cell, deg)
            assert np.allclose(p, po, rtol=0, atol=1e-15) and np.allclose(w, wo, rtol=0, atol=1e-16)
        for degree in (1, 2):
            phi, dphi = lvpp.fem.tabulate_lagrange(degree, p)
            phio, dphio = oelem.tabulate(degree, p)
            assert np.array_equal(phi, phio) and np.array_equal(dphi, dphio)
    msh, om = lvpp.mesh.create_box(3, 3, 2), omesh.box_kuhn(3, 3, 2)
    for degree in (1, 2):
        V = lvpp.fem.functionspace(msh, ("Lagrange", degree))
        orc = oobs.ObstacleOracle(om, degree=degree)
        assert np.array_equal(V.cell_nodes, orc.cell_nodes)
        assert np.allclose(V.node_coords, orc.node_coords)
        assert np.array_equal(np.sort(V.boundary_nodes), orc.bc_nodes)
        dofs = lvpp.fem.locate_dofs_boundary(V.sub(0))
        assert np.array_equal(np.sort(dofs), orc.bc_dofs)
    phi = lvpp.fem.QuadratureFunction(V)
    phi.interpolate(lvpp.fem.phi_set)
    assert np.allclose(phi.values, orc.phi_q, rtol=1e-15, atol=1e-17)


def test_alpha_update_known_answers():
    expect = [1.0, 1.0, 1.4900343193257237, 2.439200639104808, 5.349445965312164, 16.387223352344883,
              84.95478289516922, 100.0, 100.0]
    a, ak, got = 1.0, 1, []
    for k in range(9):
        a, ak = lvpp.obstacle_pg.alpha_update("double_exponential", k, a, ak, 1e2)
        got.append(a)
    assert np.allclose(got, expect, rtol=1e-14)
    assert ak > 100.0  # alpha_k is stored before clamping (obstacle_pg.py:182-183)


def test_option_parsing():
    o = lvpp.newton_options(lvpp.obstacle_pg.PETSC_OPTIONS)
    assert (o.snes_rtol, o.snes_max_it, o.snes_atol, o.snes_stol, o.snes_divtol) == (1e-6, 100, 1e-50, 1e-8, 1e4)
    assert o.ksp_rtol == 1e-12
    o = lvpp.newton_options({})
    assert (o.snes_rtol, o.snes_max_it) == (1e-8, 50)  # PETSc defaults
    assert lvpp.newton_options({"ksp_type": "gmres"}).pc_type == lvpp._capi.PC_MG
    assert lvpp.newton_options({"ksp_type": "gmres", "pc_type": "mg", "pc_mg_smoothing_sweeps": 3}).pc_degree == 3
    assert lvpp.newton_options({"ksp_type": "minres", "pc_type": "jacobi"}).pc_type == lvpp._capi.PC_JACOBI
    assert lvpp.newton_options({"snes_linesearch_type": "bt"}).snes_linesearch == lvpp._capi.LINESEARCH_BT
    assert lvpp.newton_options({"snes_linesearch_type": "none"}).snes_linesearch == lvpp._capi.LINESEARCH_NONE
    for bad in ({"snes_linesearch_type": "cp"}, {"ksp_type": "gmres", "pc_type": "jacobi"}, {"ksp_type": "cg"},
                {"pc_type": "hypre"}, {"snes_type": "vinewtonssls"}):
        with pytest.raises(NotImplementedError):
            lvpp.newton_options(bad)


def test_no_cpu_fallback_and_no_oracle_in_product():
    import torch

    for path in (ROOT / "proximalgalerkin_b200").rglob("*"):
        if path.suffix in (".py", ".cu", ".cuh", ".h"):
            text = path.read_text()
            assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), path
            assert "/root/reference" not in text, path
    if not torch.cuda.is_available():
        msh = lvpp.mesh.create_rectangle(4, 4)
        s = lvpp.obstacle_pg.setup(msh)
        with pytest.raises(lvpp._capi.LvppError):
            s["problem"].solve()


def test_bench_weak_scaling_mesh():
    """bench.py's N-GPU workload: the refined one-obstacle cube keeps ~n^3 cubes per GPU and nz divisible by N."""
    import bench

    assert bench.weak_scaling_mesh(215, 1) == (215, 215, (-1.0, -1.0, -1.0), (1.0, 1.0, 1.0), 1)
    for world, expect in ((2, (271, 272)), (4, (341, 344)), (8, (430, 432))):
        nxy, nz, lo, hi, slabs = bench.weak_scaling_mesh(215, world)
        assert (nxy, nz) == expect and nz % (2 * world) == 0 and slabs == 1  # even number of planes per rank
        assert (lo, hi) == ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))
        per_gpu = nxy * nxy * nz / world
        assert abs(per_gpu / 215**3 - 1.0) < 0.015
    assert bench.weak_scaling_mesh(215, 8, "stack") == (215, 1720, (-1.0, -1.0, -8.0), (1.0, 1.0, 8.0), 8)
    assert bench.weak_scaling_mesh(215, 1, "refine", 2) == (215, 430, (-1.0, -1.0, -2.0), (1.0, 1.0, 2.0), 2)


@pytest.mark.parametrize("cell,tdim", [("triangle", 2), ("tetrahedron", 3)])
@pytest.mark.parametrize("deg", [1, 2, 4, 6, 10])
def test_product_quadrature_integrates_monomials_exactly(cell, tdim, deg):
    """The product's quadrature tables checked on their own (not against the oracle's copy of the same numbers): every
    monomial of total degree <= deg is integrated exactly over the reference simplex,
    int x^a y^b z^c = a! b! c! / (a + b + c + d)!, weights positive, points inside."""
    from itertools import product as iproduct
    from math import factorial

    p, w = lvpp.quadrature.make_quadrature(cell, deg)
    assert p.shape[1] == tdim and np.all(w > 0) and np.all(p > -1e-14) and np.all(p.sum(axis=1) < 1 + 1e-14)
    for e in iproduct(range(deg + 1), repeat=tdim):
        if sum(e) > deg:
            continue
        exact = np.prod([factorial(k) for k in e]) / factorial(sum(e) + tdim)
        val = float(np.sum(w * np.prod(p ** np.array(e), axis=1)))
        assert abs(val - exact) <= 2e-15 * max(1.0, 1.0 / exact) * exact + 1e-17, (e, val, exact)


def test_clamped_stack_mesh_marks_the_planes_between_the_copies():
    """bench.py's weak-scaling workload: copies of one box stacked along z, every copy clamped on all six faces --
    ``create_box(clamp_every=n)`` adds the vertex planes z-index = 0, n, 2n, ... to the Dirichlet set, identically on one
    rank and on the slabs of a partition."""
    n, copies = 4, 3
    kw = dict(lo=(-1.0, -1.0, -float(copies)), hi=(1.0, 1.0, float(copies)), clamp_every=n)
    whole = lvpp.mesh.create_box(n, n, n * copies, **kw)
    open_ = lvpp.mesh.create_box(n, n, n * copies, lo=kw["lo"], hi=kw["hi"])
    extra = np.setdiff1d(whole.boundary_vertices, open_.boundary_vertices)
    assert extra.size == (copies - 1) * (n - 1) ** 2  # the interior nodes of the two planes between three copies
    assert np.allclose(np.unique(whole.coords[extra][:, 2]), [-1.0, 1.0])
    marked = set(whole.global_vertex[whole.boundary_vertices].tolist())
    for r in range(copies):
        part = lvpp.mesh.create_box(n, n, n * copies, rank=r, nranks=copies, **kw)
        assert set(part.global_vertex[part.boundary_vertices].tolist()) <= marked
        own = part.global_vertex[: part.num_owned_vertices]
        onb = np.isin(own, part.global_vertex[part.boundary_vertices])
        assert np.array_equal(np.isin(own, list(marked)), onb)


def test_bench_workload_config_is_shared_by_both_arms():
    """bench.py: the reference arm reports the b200 arm's workload (it times a bounded sample of it)."""
    import argparse
    import importlib

    bench = importlib.import_module("bench")
    a = argparse.Namespace(workload="obstacle", n=215, n2d=1000, weak="stack", slabs=1, alpha_scheme="constant", alpha_max=1e5,
                           tol_exit=1e-6)
    c1 = bench.workload_config(a, 1)
    assert c1["rows"] == 20155392 and c1["primal_dofs"] == 10077696 and "215x215x215" in c1["workload"]
    c8 = bench.workload_config(a, 8)
    assert c8["rows"] == 2 * 216 * 216 * 1721 and "215x215x1720" in c8["workload"] and "u = 0 on the planes" in c8["obstacle"]
    a.workload = "obstacle2d"
    assert bench.workload_config(a, 1)["rows"] == 2004002
