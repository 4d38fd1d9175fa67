"""Mesh and function I/O beside the hot path -- pure Python, no gmsh / h5py / adios2 needed.

The reference reads its meshes through ``dolfinx.io.XDMFFile.read_mesh`` (examples/01_obstacle_problem/obstacle_pg.py:64-65)
or builds them with the gmsh API (``src/lvpp/mesh_generation.py:11-168``, ``generate_mesh_gmsh.py:12-48``) and writes results
as VTX / XDMF (``obstacle_pg.py:242-243``, ``signorini_dolfinx.py:294-299,407-411``) and the iteration history as CSV
(``obstacle_pg.py:245-259``).  Here:

* :func:`read_msh`   gmsh MSH 2.2 and 4.1 ASCII files (what ``gmsh -format msh2/msh4`` writes): triangles / tetrahedra, the
  physical tags of the boundary elements; second-order elements are read as their straight-sided first-order simplices;
* :func:`read_xdmf`  XDMF files whose heavy data is inline (``Format="XML"``); HDF5-backed files need ``h5py`` and say so;
* :func:`write_xdmf` the inline-XML twin, so that meshes round-trip without HDF5;
* :func:`write_vtu`  one ASCII ``.vtu`` (VTK XML unstructured grid, opens in ParaView) with point data, e.g. u and psi;
* :func:`write_history_csv` the per-outer-step table of ``obstacle_pg.py:245-259``.

Everything returns / takes plain arrays; ``mesh.from_arrays(coords, cells)`` turns a mesh read here into the object the
drivers take.
"""
import csv
import xml.etree.ElementTree as ET
from pathlib import Path

import numpy as np

# gmsh element type -> (name, nodes per element, vertices of the first-order simplex)
_GMSH = {1: ("line", 2, 2), 2: ("triangle", 3, 3), 4: ("tetrahedron", 4, 4), 8: ("line", 3, 2), 9: ("triangle", 6, 3),
         11: ("tetrahedron", 10, 4), 15: ("point", 1, 1)}
_VTK_TYPE = {"triangle": 5, "tetrahedron": 10, "line": 3}
_XDMF_TOPOLOGY = {"triangle": "Triangle", "tetrahedron": "Tetrahedron"}


def _sections(text):
    out, name, buf = {}, None, []
    for line in text.splitlines():
        s = line.strip()
        if s.startswith("$End"):
            out[name] = buf
            name, buf = None, []
        elif s.startswith("$"):
            name, buf = s[1:], []
        elif name is not None and s:
            buf.append(s)
    return out


def read_msh(path):
    """Returns ``(coords [N, 3], cells [C, nv] int32, cell_name, boundary)`` where ``boundary`` is a dict
    ``physical/entity tag -> [F, nv - 1] int32`` of the boundary elements (facets) carrying that tag.  Node numbers are
    compacted to 0 .. N-1 in ascending gmsh tag order; nodes not used by any cell are dropped."""
    sec = _sections(Path(path).read_text())
    version = float(sec["MeshFormat"][0].split()[0])
    tags, xyz, elems = [], [], []  # elems: (type, tag, node tags)
    if version < 3.0:
        for ln in sec["Nodes"][1:]:
            t = ln.split()
            tags.append(int(t[0]))
            xyz.append([float(v) for v in t[1:4]])
        for ln in sec["Elements"][1:]:
            t = [int(v) for v in ln.split()]
            etype, ntags = t[1], t[2]
            elems.append((etype, t[3] if ntags > 0 else 0, t[3 + ntags:]))
    else:
        lines = sec["Nodes"]
        nblocks = int(lines[0].split()[0])
        p = 1
        for _ in range(nblocks):
            _, _, parametric, nn = (int(v) for v in lines[p].split())
            p += 1
            tags += [int(lines[p + i]) for i in range(nn)]
            p += nn
            xyz += [[float(v) for v in lines[p + i].split()[:3]] for i in range(nn)]
            p += nn
        lines = sec["Elements"]
        nblocks = int(lines[0].split()[0])
        p = 1
        for _ in range(nblocks):
            _, etag, etype, ne = (int(v) for v in lines[p].split())
            p += 1
            for i in range(ne):
                t = [int(v) for v in lines[p + i].split()]
                elems.append((etype, etag, t[1:]))
            p += ne
    for etype, _, _ in elems:
        if etype not in _GMSH:
            raise NotImplementedError(f"gmsh element type {etype} (simplicial meshes only)")
    top = max(_GMSH[e[0]][2] for e in elems)
    cell_name = {3: "triangle", 4: "tetrahedron"}.get(top)
    if cell_name is None:
        raise ValueError("no triangles or tetrahedra in the file")
    cells = np.array([e[2][:top] for e in elems if _GMSH[e[0]][2] == top], dtype=np.int64)
    tags = np.asarray(tags, dtype=np.int64)
    order = np.argsort(tags)
    used = np.unique(cells)
    lookup = {int(t): i for i, t in enumerate(used)}
    remap = np.vectorize(lookup.__getitem__, otypes=[np.int64])
    pos = order[np.searchsorted(tags[order], used)]
    coords = np.asarray(xyz, dtype=np.float64)[pos]
    boundary = {}
    for etype, etag, nodes in elems:
        if _GMSH[etype][2] == top - 1 and all(int(v) in lookup for v in nodes[:top - 1]):
            boundary.setdefault(etag, []).append([lookup[int(v)] for v in nodes[:top - 1]])
    boundary = {k: np.asarray(v, dtype=np.int32) for k, v in boundary.items()}
    return coords, remap(cells).astype(np.int32), cell_name, boundary


def write_xdmf(path, coords, cells, cell_name):
    """XDMF with the heavy data inline (``Format="XML"``): the layout ``XDMFFile.write_mesh`` produces, minus HDF5."""
    coords, cells = np.asarray(coords, dtype=np.float64), np.asarray(cells, dtype=np.int64)
    root = ET.Element("Xdmf", Version="3.0")
    grid = ET.SubElement(ET.SubElement(root, "Domain"), "Grid", Name="mesh", GridType="Uniform")
    topo = ET.SubElement(grid, "Topology", TopologyType=_XDMF_TOPOLOGY[cell_name], NumberOfElements=str(cells.shape[0]),
                         NodesPerElement=str(cells.shape[1]))
    ET.SubElement(topo, "DataItem", Dimensions=f"{cells.shape[0]} {cells.shape[1]}", NumberType="Int", Format="XML").text = \
        "\n".join(" ".join(str(v) for v in row) for row in cells)
    geo = ET.SubElement(grid, "Geometry", GeometryType="XYZ" if coords.shape[1] == 3 else "XY")
    ET.SubElement(geo, "DataItem", Dimensions=f"{coords.shape[0]} {coords.shape[1]}", Format="XML").text = \
        "\n".join(" ".join(repr(float(v)) for v in row) for row in coords)
    ET.ElementTree(root).write(str(path), xml_declaration=True, encoding="utf-8")


def read_xdmf(path, name=None):
    """``XDMFFile(...).read_mesh(name=...)``: returns ``(coords, cells int32, cell_name)`` of the first (or named) uniform
    grid.  Inline (``Format="XML"``) data is parsed here; ``Format="HDF"`` needs h5py."""
    root = ET.parse(str(path)).getroot()
    grids = [g for g in root.iter("Grid") if g.get("GridType", "Uniform") == "Uniform" and (name is None or g.get("Name") == name)]
    if not grids:
        raise ValueError(f"no uniform grid{'' if name is None else ' named ' + name} in {path}")
    grid = grids[0]

    def data(item, dtype):
        dims = [int(v) for v in item.get("Dimensions").split()]
        if item.get("Format", "XML").upper() == "XML":
            return np.array(item.text.split(), dtype=dtype).reshape(dims)
        try:
            import h5py
        except ImportError as e:
            raise ImportError("this XDMF file keeps its data in HDF5; reading it needs h5py (not in this image) -- "
                              "re-export with inline data (io.write_xdmf) or as gmsh .msh") from e
        fname, dset = item.text.strip().split(":")
        with h5py.File(Path(path).parent / fname, "r") as f:
            return np.asarray(f[dset], dtype=dtype).reshape(dims)

    topo = grid.find("Topology")
    ttype = topo.get("TopologyType", "").lower()
    cell_name = {"triangle": "triangle", "tetrahedron": "tetrahedron", "triangle_6": "triangle", "tetrahedron_10": "tetrahedron"}.get(ttype)
    if cell_name is None:
        raise NotImplementedError(f"XDMF topology {ttype!r} (simplicial meshes only)")
    cells = data(topo.find("DataItem"), np.int64)
    cells = cells[:, : (3 if cell_name == "triangle" else 4)]  # second-order geometry: keep the vertices
    coords = data(grid.find("Geometry").find("DataItem"), np.float64)
    used = np.unique(cells)
    if used.size != coords.shape[0]:  # the mid-side nodes of a second-order mesh are dropped
        lookup = np.full(coords.shape[0], -1, dtype=np.int64)
        lookup[used] = np.arange(used.size)
        cells, coords = lookup[cells], coords[used]
    return coords, cells.astype(np.int32), cell_name


def write_vtu(path, coords, cells, cell_name, point_data=None):
    """ASCII VTK XML unstructured grid with ``point_data = {name: [N] or [N, k] array}`` (e.g. u and psi at the vertices)."""
    coords = np.asarray(coords, dtype=np.float64)
    if coords.shape[1] == 2:
        coords = np.concatenate([coords, np.zeros((coords.shape[0], 1))], axis=1)
    cells = np.asarray(cells, dtype=np.int64)

    def arr(parent, a, **attrs):
        a = np.asarray(a)
        ET.SubElement(parent, "DataArray", format="ascii", **attrs).text = \
            " ".join((repr(float(v)) if a.dtype.kind == "f" else str(int(v))) for v in a.ravel())

    root = ET.Element("VTKFile", type="UnstructuredGrid", version="1.0", byte_order="LittleEndian")
    piece = ET.SubElement(ET.SubElement(root, "UnstructuredGrid"), "Piece", NumberOfPoints=str(coords.shape[0]),
                          NumberOfCells=str(cells.shape[0]))
    pd = ET.SubElement(piece, "PointData")
    for name, a in (point_data or {}).items():
        a = np.asarray(a, dtype=np.float64)
        if a.shape[0] != coords.shape[0]:
            raise ValueError(f"point data {name!r}: {a.shape[0]} values for {coords.shape[0]} points")
        arr(pd, a, type="Float64", Name=name, NumberOfComponents=str(1 if a.ndim == 1 else a.shape[1]))
    arr(ET.SubElement(piece, "Points"), coords, type="Float64", NumberOfComponents="3")
    c = ET.SubElement(piece, "Cells")
    arr(c, cells, type="Int64", Name="connectivity")
    arr(c, np.arange(1, cells.shape[0] + 1) * cells.shape[1], type="Int64", Name="offsets")
    arr(c, np.full(cells.shape[0], _VTK_TYPE[cell_name]), type="UInt8", Name="types")
    ET.ElementTree(root).write(str(path), xml_declaration=True, encoding="utf-8")


def write_solution(path, V, x):
    """u and psi of a P1 obstacle solution ``x`` (node-interleaved mixed vector of ``fem.functionspace`` V) as a .vtu."""
    msh = V.mesh
    x = np.asarray(x)
    nv = msh.coords.shape[0]
    write_vtu(path, msh.coords, msh.cells, msh.cell_name, {"u": x[0:2 * nv:2], "psi": x[1:2 * nv:2]})


def write_history_csv(path, history, dofs=None):
    """The table of obstacle_pg.py:245-259: one row per outer step (energy, complementarity, feasibility, dual feasibility,
    Newton steps, step size, primal and latent increments; ``dofs`` = primal dofs, the reference's column of that name)."""
    cols = [("energy", "Energy"), ("complementarity", "Complementarity"), ("feasibility", "Feasibility"),
            ("dual_feasibility", "Dual Feasibility"), ("newton_steps", "Newton steps"), ("alpha", "Step size"),
            ("primal_increment", "Primal increments"), ("latent_increment", "Latent increments")]
    cols = [(k, label) for k, label in cols if k in history]
    n = len(history[cols[0][0]])
    with open(path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([label for _, label in cols] + (["dofs"] if dofs is not None else []))
        for i in range(n):
            w.writerow([history[k][i] for k, _ in cols] + ([dofs] if dofs is not None else []))
