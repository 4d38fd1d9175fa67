"""Simplicial meshes for the LVPP path, with slab partitioning for one-process-per-GPU runs.

The reference reads gmsh/XDMF meshes through dolfinx (examples/01_obstacle_problem/obstacle_pg.py:64-65)
and lets dolfinx partition them over MPI ranks.  Here the synthetic configurations of SURVEY.md
section 8d are generated in-process, directly in partitioned form: rank ``r`` of ``nranks`` owns a
contiguous slab of vertex layers along the last axis, numbers its owned vertices first and the ghost
vertices (one layer below, one above) last -- the dolfinx index-map convention that
src/lvpp/problem.py:56-73 relies on -- and holds every cell incident to an owned vertex, owned cells
first.  :func:`from_arrays` wraps any (coords, cells) pair, e.g. arrays exported from dolfinx.
"""
from dataclasses import dataclass, field
from itertools import permutations

import numpy as np


@dataclass
class Halo:
    """Neighbour ranks and the local vertex lists exchanged with each (owned -> send, ghost -> recv)."""

    neighbors: list = field(default_factory=list)
    send: list = field(default_factory=list)  # list of int32 arrays of owned local vertices
    recv: list = field(default_factory=list)  # list of int32 arrays of ghost local vertices


@dataclass
class Mesh:
    coords: np.ndarray  # [num_vertices, gdim] float64 (owned first, ghosts last)
    cells: np.ndarray  # [num_cells, tdim + 1] int32 local vertex numbers (owned cells first)
    cell_name: str
    num_owned_vertices: int
    num_owned_cells: int
    boundary_vertices: np.ndarray  # local numbers of vertices on the exterior boundary (owned + ghost)
    global_vertex: np.ndarray  # [num_vertices] int64 global number of each local vertex
    num_global_vertices: int
    halo: Halo = field(default_factory=Halo)
    rank: int = 0
    nranks: int = 1
    shape: tuple = ()  # structured meshes: global cube counts

    @property
    def tdim(self):
        return {"triangle": 2, "tetrahedron": 3}[self.cell_name]

    @property
    def gdim(self):
        return self.coords.shape[1]

    @property
    def num_vertices(self):
        return self.coords.shape[0]

    @property
    def num_cells(self):
        return self.cells.shape[0]


def _slab(nlayers, rank, nranks):
    """Cube layers [k0, k1) of ``rank`` and its owned vertex planes [p0, p1)."""
    if nlayers < nranks:
        raise ValueError(f"{nlayers} layers cannot be split over {nranks} ranks")
    k0 = (nlayers * rank) // nranks
    k1 = (nlayers * (rank + 1)) // nranks
    p0, p1 = k0, (k1 if rank < nranks - 1 else k1 + 1)
    return k0, k1, p0, p1


def _structured(n, lo, hi, rank, nranks, cube_to_simplices, cell_name, clamp_every=None):
    """Common slab machinery.  ``n`` = cube counts per axis (last axis is partitioned).  ``clamp_every``: vertex planes
    of the last axis whose index is a multiple of it count as boundary as well (a stack of clamped membranes)."""
    dim = len(n)
    nv = [m + 1 for m in n]
    plane = int(np.prod(nv[:-1]))  # vertices per layer of the last axis
    k0, k1, p0, p1 = _slab(n[-1], rank, nranks)
    ghost_below = rank > 0
    ghost_above = rank < nranks - 1
    planes = list(range(p0, p1))
    if ghost_below:
        planes.append(p0 - 1)
    if ghost_above:
        planes.append(p1)
    planes = np.array(planes, dtype=np.int64)
    n_owned = (p1 - p0) * plane
    # global number of every local vertex (lexicographic, first axis fastest)
    gv = (planes[:, None] * plane + np.arange(plane, dtype=np.int64)[None, :]).ravel()
    # coordinates
    axes = [np.linspace(lo[d], hi[d], nv[d]) for d in range(dim)]
    inplane = np.meshgrid(*axes[:-1][::-1], indexing="ij")[::-1]  # first axis fastest
    inplane = [a.ravel() for a in inplane]
    coords = np.empty((gv.size, dim))
    for d in range(dim - 1):
        coords[:, d] = np.tile(inplane[d], planes.size)
    coords[:, dim - 1] = np.repeat(axes[-1][planes], plane)
    # local index of a vertex plane
    plane_slot = {int(p): s for s, p in enumerate(planes)}
    # cells: own layers first, then the ghost layer below
    layers = list(range(k0, k1)) + ([k0 - 1] if ghost_below else [])
    strides = [int(np.prod(nv[:d])) for d in range(dim)]  # global-style strides within a slab of planes
    idx = np.meshgrid(*[np.arange(m) for m in n[:-1]][::-1], indexing="ij")[::-1]
    base_inplane = sum(i.ravel().astype(np.int64) * strides[d] for d, i in enumerate(idx))
    blocks = []
    for k in layers:
        lo_off = plane_slot[k] * plane
        hi_off = plane_slot[k + 1] * plane
        blocks.append(cube_to_simplices(base_inplane, strides, lo_off, hi_off))
    cells = np.concatenate(blocks, axis=0).astype(np.int32)
    cells_per_layer = blocks[0].shape[0]
    n_owned_cells = (k1 - k0) * cells_per_layer
    # boundary vertices of the global box
    onb = np.zeros(gv.size, dtype=bool)
    for d in range(dim - 1):
        c = (gv % (strides[d] * nv[d])) // strides[d]
        onb |= (c == 0) | (c == n[d])
    kk = gv // plane
    onb |= (kk == 0) | (kk == n[-1])
    if clamp_every:
        onb |= (kk % int(clamp_every)) == 0
    # halo
    halo = Halo()
    ar = np.arange(plane, dtype=np.int32)
    if ghost_below:
        halo.neighbors.append(rank - 1)
        halo.send.append(ar + plane_slot[p0] * plane)
        halo.recv.append(ar + plane_slot[p0 - 1] * plane)
    if ghost_above:
        halo.neighbors.append(rank + 1)
        halo.send.append(ar + plane_slot[p1 - 1] * plane)
        halo.recv.append(ar + plane_slot[p1] * plane)
    return Mesh(
        coords=coords,
        cells=np.ascontiguousarray(cells),
        cell_name=cell_name,
        num_owned_vertices=n_owned,
        num_owned_cells=n_owned_cells,
        boundary_vertices=np.flatnonzero(onb).astype(np.int32),
        global_vertex=gv,
        num_global_vertices=int(np.prod(nv)),
        halo=halo,
        rank=rank,
        nranks=nranks,
        shape=tuple(n),
    )


def create_rectangle(nx, ny, lo=(-1.0, -1.0), hi=(1.0, 1.0), rank=0, nranks=1, diagonal="right"):
    """[lo, hi] rectangle, nx x ny squares, each split by its "right" diagonal (dolfinx
    DiagonalType.right); partitioned into slabs along y.  ``diagonal="crossed"`` (single rank) splits
    every square into four triangles around its centre (DiagonalType.crossed,
    examples/04_multiphase/multiphase_dolfinx.py:34-36); the centres are numbered after the grid vertices."""
    if diagonal == "crossed":
        if nranks != 1:
            raise NotImplementedError("crossed meshes are single-rank")
        xs, ys = np.linspace(lo[0], hi[0], nx + 1), np.linspace(lo[1], hi[1], ny + 1)
        X, Y = np.meshgrid(xs, ys, indexing="xy")
        grid = np.stack([X.ravel(), Y.ravel()], axis=1)
        i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
        v00 = (j * (nx + 1) + i).ravel()
        v10, v01, v11 = v00 + 1, v00 + nx + 1, v00 + nx + 2
        mid = (nx + 1) * (ny + 1) + np.arange(nx * ny)
        centres = 0.25 * (grid[v00] + grid[v10] + grid[v01] + grid[v11])
        cells = np.stack([np.stack([v00, v10, mid], 1), np.stack([v10, v11, mid], 1), np.stack([v11, v01, mid], 1),
                          np.stack([v01, v00, mid], 1)], axis=1).reshape(-1, 3)
        return from_arrays(np.vstack([grid, centres]), cells)
    if diagonal != "right":
        raise ValueError(diagonal)

    def split(base, strides, lo_off, hi_off):
        v00 = base + lo_off
        v10 = v00 + strides[0]
        v01 = base + hi_off
        v11 = v01 + strides[0]
        return np.stack([np.stack([v00, v10, v11], 1), np.stack([v00, v11, v01], 1)], axis=1).reshape(-1, 3)

    return _structured((nx, ny), lo, hi, rank, nranks, split, "triangle")


def create_box(nx, ny, nz, lo=(-1.0, -1.0, -1.0), hi=(1.0, 1.0, 1.0), rank=0, nranks=1, clamp_every=None):
    """[lo, hi] box, nx x ny x nz cubes, each split into six Kuhn tetrahedra sharing the main
    diagonal; partitioned into slabs along z.  ``clamp_every = m`` adds the vertex planes z-index = 0, m, 2m, ... to
    ``boundary_vertices`` (bench.py's weak-scaling workload: copies of one box stacked along z, each clamped on all
    six faces)."""

    def split(base, strides, lo_off, hi_off):
        # corner (dx, dy, dz) of the cube -> local vertex number
        def corner(dx, dy, dz):
            return base + (hi_off if dz else lo_off) + dx * strides[0] + dy * strides[1]

        tets = []
        for perm in permutations(range(3)):
            c = [0, 0, 0]
            verts = [corner(*c)]
            for axis in perm:
                c[axis] = 1
                verts.append(corner(*c))
            tets.append(np.stack(verts, 1))
        return np.stack(tets, axis=1).reshape(-1, 4)

    return _structured((nx, ny, nz), lo, hi, rank, nranks, split, "tetrahedron", clamp_every)


def from_arrays(coords, cells, boundary_vertices=None):
    """Single-rank mesh from vertex coordinates and cell connectivity (any simplicial mesh).

    ``boundary_vertices`` are located topologically when not given: vertices of facets that belong
    to exactly one cell (dolfinx.mesh.exterior_facet_indices, obstacle_pg.py:76-77)."""
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    cells = np.ascontiguousarray(cells, dtype=np.int32)
    name = {3: "triangle", 4: "tetrahedron"}[cells.shape[1]]
    if boundary_vertices is None:
        nvc = cells.shape[1]
        facets = np.concatenate([np.delete(cells, i, axis=1) for i in range(nvc)], axis=0)
        facets.sort(axis=1)
        uniq, counts = np.unique(facets, axis=0, return_counts=True)
        boundary_vertices = np.unique(uniq[counts == 1])
    return Mesh(
        coords=coords,
        cells=cells,
        cell_name=name,
        num_owned_vertices=coords.shape[0],
        num_owned_cells=cells.shape[0],
        boundary_vertices=np.asarray(boundary_vertices, dtype=np.int32),
        global_vertex=np.arange(coords.shape[0], dtype=np.int64),
        num_global_vertices=coords.shape[0],
    )
