"""Structured simplicial meshes (oracle; see oracle/__init__.py).

The reference reads gmsh/XDMF meshes (examples/01_obstacle_problem/obstacle_pg.py:64-65,
generate_mesh_gmsh.py:12-48); gmsh and HDF5 are unavailable, so the synthetic configurations of
SURVEY.md section 8d are generated in-process: the square [lo,hi]^2 with right or crossed diagonals
(the reference's multiphase example uses ``DiagonalType.crossed``,
examples/04_multiphase/multiphase_dolfinx.py:34-36) and the cube [lo,hi]^3 split into six Kuhn
tetrahedra per cube.  Vertices are numbered lexicographically (x fastest).
"""
import itertools

import numpy as np


class Mesh:
    def __init__(self, coords, cells, cell_name):
        self.coords = np.ascontiguousarray(coords, dtype=np.float64)
        self.cells = np.ascontiguousarray(cells, dtype=np.int32)
        self.cell_name = cell_name
        self.tdim = {"triangle": 2, "tetrahedron": 3}[cell_name]
        self.gdim = self.coords.shape[1]

    @property
    def num_vertices(self):
        return self.coords.shape[0]

    @property
    def num_cells(self):
        return self.cells.shape[0]


def rectangle(nx, ny, diagonal="right", lo=-1.0, hi=1.0):
    xs = np.linspace(lo, hi, nx + 1)
    ys = np.linspace(lo, hi, ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="xy")  # Y slow, X fast
    coords = np.stack([X.ravel(), Y.ravel()], axis=1)
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    v00 = (j * (nx + 1) + i).ravel()
    v10, v01, v11 = v00 + 1, v00 + nx + 1, v00 + nx + 2
    if diagonal == "right":
        cells = np.stack(
            [np.stack([v00, v10, v11], 1), np.stack([v00, v11, v01], 1)], axis=1
        ).reshape(-1, 3)
    elif diagonal == "crossed":
        mid = (nx + 1) * (ny + 1) + np.arange(nx * ny)
        cm = 0.25 * (coords[v00] + coords[v10] + coords[v01] + coords[v11])
        coords = np.vstack([coords, cm])
        cells = np.stack(
            [
                np.stack([v00, v10, mid], 1),
                np.stack([v10, v11, mid], 1),
                np.stack([v11, v01, mid], 1),
                np.stack([v01, v00, mid], 1),
            ],
            axis=1,
        ).reshape(-1, 3)
    else:
        raise ValueError(diagonal)
    return Mesh(coords, cells, "triangle")


def box_kuhn(nx, ny, nz, lo=(-1.0, -1.0, -1.0), hi=(1.0, 1.0, 1.0)):
    xs = np.linspace(lo[0], hi[0], nx + 1)
    ys = np.linspace(lo[1], hi[1], ny + 1)
    zs = np.linspace(lo[2], hi[2], nz + 1)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    coords = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    sx, sy, sz = 1, nx + 1, (nx + 1) * (ny + 1)
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    base = (k * sz + j * sy + i * sx).ravel()
    strides = (sx, sy, sz)
    tets = []
    for perm in itertools.permutations(range(3)):
        a = base + strides[perm[0]]
        b = a + strides[perm[1]]
        c = b + strides[perm[2]]
        tets.append(np.stack([base, a, b, c], 1))
    cells = np.stack(tets, axis=1).reshape(-1, 4)
    return Mesh(coords, cells, "tetrahedron")


def facets_of_cells(cells):
    """All (cell, local facet) vertex tuples, facet i opposite local vertex i."""
    nv = cells.shape[1]
    out = []
    for i in range(nv):
        keep = [j for j in range(nv) if j != i]
        out.append(cells[:, keep])
    return np.stack(out, axis=1)  # [C, nv, nv-1]


def exterior_facets(mesh):
    """Sorted vertex tuples of facets that belong to exactly one cell
    (dolfinx.mesh.exterior_facet_indices, reference call site obstacle_pg.py:76-77)."""
    f = np.sort(facets_of_cells(mesh.cells).reshape(-1, mesh.cells.shape[1] - 1), axis=1)
    uniq, counts = np.unique(f, axis=0, return_counts=True)
    return uniq[counts == 1]


def boundary_vertices(mesh):
    return np.unique(exterior_facets(mesh).ravel())
