"""CPU oracle for the LVPP Newton inner loop -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

**Parity unpinned**: the reference (METHODS-Group/ProximalGalerkin) holds no golden vectors,
known-answer tests or fixtures for this path (its CI only checks exit codes), and the stack that
does its arithmetic (DOLFINx v0.10.0.post1, Basix v0.10.0, UFL 2025.2.0, FFCx v0.10.0, PETSc +
MUMPS from ghcr.io/fenics/dolfinx/dev-env:v0.10.0-openmpi; pinned in docker/Dockerfile:1,60-63 of
the reference) is not vendored and not installable here.  This package restates the published
algorithms those libraries implement for the path (Lagrange P1/P2 on simplices, cell-wise
quadrature assembly, dolfinx Dirichlet lifting conventions, PETSc SNES ``newtonls`` convergence
tests, sparse LU) and anchors on the reference's own call sites:

* forms, options, outer loop, observables: examples/01_obstacle_problem/obstacle_pg.py:68-227
* Dirichlet conventions / callback order:  src/lvpp/problem.py:54-77,114-124
* explicit block structure of the Newton system: examples/01_obstacle_problem/obstacle_finite_difference.jl:29-43

What *is* pinned: the alpha-schedule known answers, the obstacle closed form, exact P1 element
matrices and the structural identities listed in SURVEY.md section 8c (tests/test_oracle_*.py).

``oracle/c/lvpp_cpu.c`` (bound by ``oracle/cpu_kernels.py``) is the compiled part: the Jacobian assembly
and MatMult of the reference on its own data layout in C + OpenMP, checked against the numpy functions
here, used only as the all-cores CPU baseline of bench.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (proximalgalerkin_b200) never does.
"""
