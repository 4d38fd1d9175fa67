"""Lagrange P1/P2 on simplices (oracle; see oracle/__init__.py).

Stands in for ``basix.ufl.element("Lagrange", cell, p)`` (reference call site
examples/01_obstacle_problem/obstacle_pg.py:68).  Local node order: vertices first, then (P2) edge
midpoints with edge i numbered as in basix's published reference topology: triangle edges
(1,2),(0,2),(0,1); tetrahedron edges (2,3),(1,3),(1,2),(0,3),(0,2),(0,1).
"""
import numpy as np

EDGES = {
    2: [(1, 2), (0, 2), (0, 1)],
    3: [(2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)],
}


def tabulate(degree, points):
    """Return (phi [nq, nld], dphi [nq, nld, tdim]) on the reference simplex."""
    points = np.asarray(points, dtype=np.float64)
    nq, tdim = points.shape
    lam = np.concatenate([1.0 - points.sum(axis=1, keepdims=True), points], axis=1)  # [nq, tdim+1]
    dlam = np.concatenate([-np.ones((1, tdim)), np.eye(tdim)], axis=0)  # [tdim+1, tdim]
    if degree == 1:
        return lam.copy(), np.broadcast_to(dlam, (nq, tdim + 1, tdim)).copy()
    if degree == 2:
        nv = tdim + 1
        edges = EDGES[tdim]
        nld = nv + len(edges)
        phi = np.empty((nq, nld))
        dphi = np.empty((nq, nld, tdim))
        for a in range(nv):
            phi[:, a] = lam[:, a] * (2.0 * lam[:, a] - 1.0)
            dphi[:, a, :] = (4.0 * lam[:, a] - 1.0)[:, None] * dlam[a][None, :]
        for e, (a, b) in enumerate(edges):
            phi[:, nv + e] = 4.0 * lam[:, a] * lam[:, b]
            dphi[:, nv + e, :] = 4.0 * (lam[:, a, None] * dlam[b][None, :] + lam[:, b, None] * dlam[a][None, :])
        return phi, dphi
    raise ValueError("degree must be 1 or 2")


def build_nodes(mesh, degree):
    """Scalar Lagrange node numbering: vertices keep their mesh numbers; P2 edge nodes follow.

    Returns (cell_nodes [C, nld] int32, num_nodes, node_coords [num_nodes, gdim])."""
    cells = mesh.cells
    if degree == 1:
        return cells.copy(), mesh.num_vertices, mesh.coords.copy()
    edges = EDGES[mesh.tdim]
    ev = np.stack([np.sort(cells[:, list(e)], axis=1) for e in edges], axis=1)  # [C, ne, 2]
    flat = ev.reshape(-1, 2).astype(np.int64)
    key = flat[:, 0] * mesh.num_vertices + flat[:, 1]
    uniq, inv = np.unique(key, return_inverse=True)
    edge_ids = inv.reshape(cells.shape[0], len(edges)).astype(np.int32) + mesh.num_vertices
    v0 = (uniq // mesh.num_vertices).astype(np.int64)
    v1 = (uniq % mesh.num_vertices).astype(np.int64)
    mid = 0.5 * (mesh.coords[v0] + mesh.coords[v1])
    return (
        np.concatenate([cells, edge_ids], axis=1).astype(np.int32),
        mesh.num_vertices + len(uniq),
        np.vstack([mesh.coords, mid]),
    )


def geometry(mesh):
    """Affine cell geometry: (detJ [C], Jinv [C, tdim, gdim]) with J[:, :, k] = x_{k+1} - x_0."""
    x = mesh.coords[mesh.cells]  # [C, nv, gdim]
    J = np.transpose(x[:, 1:, :] - x[:, :1, :], (0, 2, 1))  # [C, gdim, tdim]
    detJ = np.linalg.det(J)
    Jinv = np.linalg.inv(J)
    return detJ, Jinv
