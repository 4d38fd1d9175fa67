"""Mesh / function I/O beside the hot path (proximalgalerkin_b200/io.py): gmsh MSH 2.2 / 4.1 ASCII, inline
XDMF round trip, VTU and CSV output.  CPU only."""
import csv
import xml.etree.ElementTree as ET

import numpy as np
import pytest

import proximalgalerkin_b200 as lvpp
from proximalgalerkin_b200 import io as lio

MSH2 = """$MeshFormat
2.2 0 8
$EndMeshFormat
$Nodes
5
1 0 0 0
2 1 0 0
3 1 1 0
4 0 1 0
7 0.5 0.5 0
$EndNodes
$Elements
8
1 1 2 10 1 1 2
2 1 2 10 1 2 3
3 1 2 11 2 3 4
4 1 2 11 2 4 1
5 2 2 1 1 1 2 7
6 2 2 1 1 2 3 7
7 2 2 1 1 3 4 7
8 2 2 1 1 4 1 7
$EndElements
"""

MSH4 = """$MeshFormat
4.1 0 8
$EndMeshFormat
$Nodes
2 5 1 5
2 1 0 4
1
2
3
4
0 0 0
1 0 0
0 1 0
0 0 1
3 1 0 1
5
1 1 1
$EndNodes
$Elements
2 3 1 3
3 1 4 2
1 1 2 3 4
2 2 3 4 5
2 7 2 1
3 1 2 3
$EndElements
"""


def test_read_msh_2_and_4(tmp_path):
    p = tmp_path / "square.msh"
    p.write_text(MSH2)
    coords, cells, name, bnd = lio.read_msh(p)
    assert name == "triangle" and coords.shape == (5, 3) and cells.shape == (4, 3)
    assert np.array_equal(cells[0], [0, 1, 4])  # gmsh tag 7 is the fifth node
    assert sorted(bnd) == [10, 11] and bnd[10].shape == (2, 2)
    msh = lvpp.mesh.from_arrays(coords[:, :2], cells)
    assert msh.num_cells == 4 and msh.cell_name == "triangle"
    p4 = tmp_path / "tets.msh"
    p4.write_text(MSH4)
    coords, cells, name, bnd = lio.read_msh(p4)
    assert name == "tetrahedron" and coords.shape == (5, 3) and cells.shape == (2, 4)
    assert np.array_equal(cells[1], [1, 2, 3, 4]) and np.array_equal(bnd[7], [[0, 1, 2]])


@pytest.mark.parametrize("dim", [2, 3])
def test_xdmf_round_trip_and_vtu(tmp_path, dim):
    msh = lvpp.mesh.create_rectangle(3, 2) if dim == 2 else lvpp.mesh.create_box(2, 2, 2)
    f = tmp_path / "mesh.xdmf"
    lio.write_xdmf(f, msh.coords, msh.cells, msh.cell_name)
    coords, cells, name = lio.read_xdmf(f, name="mesh")
    assert name == msh.cell_name and np.array_equal(cells, msh.cells) and np.array_equal(coords, msh.coords)
    back = lvpp.mesh.from_arrays(coords, cells)
    assert np.array_equal(np.sort(back.boundary_vertices), np.sort(msh.boundary_vertices))
    V = lvpp.fem.functionspace(msh, ("Lagrange", 1))
    x = np.arange(V.num_rows, dtype=np.float64)
    v = tmp_path / "sol.vtu"
    lio.write_solution(v, V, x)
    root = ET.parse(v).getroot()
    piece = root.find("UnstructuredGrid/Piece")
    assert int(piece.get("NumberOfPoints")) == msh.coords.shape[0] and int(piece.get("NumberOfCells")) == msh.cells.shape[0]
    data = {d.get("Name"): np.array(d.text.split(), dtype=float) for d in piece.find("PointData")}
    assert np.array_equal(data["u"], x[0::2]) and np.array_equal(data["psi"], x[1::2])
    types = [d for d in piece.find("Cells") if d.get("Name") == "types"][0]
    assert set(types.text.split()) == {"5" if dim == 2 else "10"}


def test_xdmf_hdf5_backed_file_says_what_it_needs(tmp_path):
    f = tmp_path / "h5.xdmf"
    f.write_text('<Xdmf><Domain><Grid Name="mesh" GridType="Uniform"><Topology TopologyType="Triangle" NumberOfElements="1">'
                 '<DataItem Dimensions="1 3" Format="HDF">mesh.h5:/Mesh/mesh/topology</DataItem></Topology>'
                 '<Geometry GeometryType="XY"><DataItem Dimensions="3 2" Format="HDF">mesh.h5:/Mesh/mesh/geometry</DataItem>'
                 '</Geometry></Grid></Domain></Xdmf>')
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="h5py"):
            lio.read_xdmf(f)


def test_history_csv(tmp_path):
    h = {"energy": [1.0, 0.5], "complementarity": [0.1, 0.01], "feasibility": [0.0, 0.0], "dual_feasibility": [0.0, 0.0],
         "newton_steps": [6, 4], "alpha": [1.0, 1.0], "primal_increment": [1.1, 0.3], "latent_increment": [2.0, 0.5]}
    f = tmp_path / "h.csv"
    lio.write_history_csv(f, h, dofs=121)
    rows = list(csv.reader(open(f)))
    assert rows[0][:2] == ["Energy", "Complementarity"] and rows[0][-1] == "dofs" and len(rows) == 3
    assert rows[2][4] == "4" and rows[2][-1] == "121"
