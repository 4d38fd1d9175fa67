/*
 * lvpp_b200.h -- C ABI of liblvpp_b200.so: the B200-native LVPP (latent variable proximal point)
 * Newton inner loop for the obstacle problem.
 *
 * This is the drop-in boundary for the path that the reference (METHODS-Group/ProximalGalerkin)
 * hands to dolfinx/PETSc.  The reference has no FFI of its own for this path (it is Python calling
 * dolfinx C++ / PETSc C through their Python bindings); each entry point below names the reference
 * call it replaces.  Citations are relative to the reference repository root.
 *
 * Conventions
 *   - plain C, no exceptions; every function returns 0 on success or a negative LVPP_E_* code;
 *     lvpp_last_error() returns a message for the calling thread's last failure.
 *   - one handle = one GPU (the CUDA device current at lvpp_create) = one CUDA stream.
 *   - "host" pointers in lvpp_obstacle_desc are read during lvpp_create and not retained.
 *   - vector arguments named d_* are DEVICE pointers (fp64); h_* are HOST pointers.
 *   - unknown layout: node-interleaved mixed vector, row 2n = u at scalar node n, row 2n+1 = psi at
 *     node n (a recorded permutation of dolfinx's mixed-element numbering).
 *     Local numbering is "owned nodes first, ghost nodes last" (dolfinx index-map convention).
 */
#ifndef LVPP_B200_H
#define LVPP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LVPP_OK 0
#define LVPP_E_INVALID (-1)   /* bad argument / unsupported combination */
#define LVPP_E_CUDA (-2)      /* CUDA runtime failure */
#define LVPP_E_CAPACITY (-3)  /* an internal fixed capacity was exceeded (row length, index width) */
#define LVPP_E_COMM (-4)      /* NCCL failure / communicator not initialised */
#define LVPP_E_NOGPU (-5)     /* no CUDA device */

/* SNESConvergedReason values returned by lvpp_newton_solve (PETSc petscsnes.h numbering, which the
 * reference reads back at examples/01_obstacle_problem/obstacle_pg.py:191). */
#define LVPP_SNES_CONVERGED_FNORM_ABS 2
#define LVPP_SNES_CONVERGED_FNORM_RELATIVE 3
#define LVPP_SNES_CONVERGED_SNORM_RELATIVE 4
#define LVPP_SNES_DIVERGED_LINEAR_SOLVE (-3)
#define LVPP_SNES_DIVERGED_FNORM_NAN (-4)
#define LVPP_SNES_DIVERGED_MAX_IT (-5)
#define LVPP_SNES_DIVERGED_DTOL (-9)
/* KSPConvergedReason subset */
#define LVPP_KSP_CONVERGED_RTOL 2
#define LVPP_KSP_CONVERGED_ATOL 3
#define LVPP_KSP_DIVERGED_ITS (-3)
#define LVPP_KSP_DIVERGED_NANORINF (-9)
#define LVPP_KSP_DIVERGED_BREAKDOWN (-5)

typedef struct lvpp_problem* lvpp_handle;

/* Obstacle description: closed form of examples/01_obstacle_problem/obstacle_pg.py:92-104 evaluated
 * on the device at the physical quadrature points (r = Euclidean norm of x), or an explicit array. */
#define LVPP_OBSTACLE_ARRAY 0
#define LVPP_OBSTACLE_PHI_SET 1

/* Preconditioner for the saddle-point Krylov solve (recipe of examples/09_eikonal/ex40.cpp:261-274:
 * diagonal on the (0,0) block, diag Schur approximation D + M diag(A)^-1 M on the (1,1) block). */
#define LVPP_PC_JACOBI 0
#define LVPP_PC_CHEBYSHEV 1 /* reserved */
/* monolithic aggregation multigrid (node-block Jacobi smoother) preconditioning restarted GMRES */
#define LVPP_PC_MG 2

/* Replaces, for the obstacle forms of obstacle_pg.py:68-125: fem.functionspace(mixed P_p x P_p)
 * (:68-70), the Dirichlet data (:76-83), the quadrature-space obstacle Function (:106-111), the
 * compiled forms F and J (:116-125) and dolfinx.fem.petsc.create_matrix (src/lvpp/problem.py:110). */
typedef struct lvpp_obstacle_desc {
  int32_t tdim;            /* 2 = triangles, 3 = tetrahedra (affine geometry) */
  int32_t nld;             /* scalar Lagrange nodes per cell: P1 tdim+1; P2 6 / 10 */
  int64_t num_nodes;       /* local scalar nodes, owned + ghost */
  int64_t num_owned;       /* owned scalar nodes (== num_nodes on one GPU) */
  int64_t num_cells;       /* local cells (every cell incident to an owned node must be present) */
  int64_t num_owned_cells; /* cells [0, num_owned_cells) are integrated by this rank's observables */
  const double* node_coords;  /* host [num_nodes * tdim]; only the vertex nodes are read */
  const int32_t* cell_nodes;  /* host [num_cells * nld]; first tdim+1 entries are the vertices */
  int32_t nq;              /* quadrature points per cell (<= 64) */
  const double* qweights;  /* host [nq], sum = reference cell volume */
  const double* phi_tab;   /* host [nq * nld]        basis values at the quadrature points */
  const double* dphi_tab;  /* host [nq * nld * tdim] reference gradients at the quadrature points */
  const double* qpoints;   /* host [nq * tdim] reference coordinates (used by LVPP_OBSTACLE_PHI_SET) */
  int64_t num_bc;          /* Dirichlet nodes of the u field */
  const int32_t* bc_nodes; /* host [num_bc] local scalar node numbers */
  const double* bc_values; /* host [num_bc] prescribed u values (NULL = 0) */
  int32_t obstacle_kind;   /* LVPP_OBSTACLE_* */
  const double* phi_obs_q; /* host [num_cells * nq] when obstacle_kind == LVPP_OBSTACLE_ARRAY */
  double f;                /* constant forcing (obstacle_pg.py:74) */
  /* LVPP_OBSTACLE_PHI_SET on stacked domains (weak scaling): when obstacle_period > 0 the last coordinate is
   * wrapped, x_last <- fmod(x_last - obstacle_origin, period) - period / 2, before r is taken, so every period
   * holds its own copy of the reference's obstacle; 0 = the reference's single obstacle at the origin. */
  double obstacle_period;
  double obstacle_origin;
  /* halo description (all zero / NULL on one GPU): neighbour ranks and the local node lists
   * exchanged with each; send lists hold owned nodes, recv lists hold ghost nodes. */
  int32_t num_neighbors;
  const int32_t* neighbor_ranks;  /* host [num_neighbors] */
  const int64_t* send_ptr;        /* host [num_neighbors + 1] */
  const int32_t* send_nodes;      /* host [send_ptr[num_neighbors]] */
  const int64_t* recv_ptr;        /* host [num_neighbors + 1] */
  const int32_t* recv_nodes;      /* host [recv_ptr[num_neighbors]] */
} lvpp_obstacle_desc;

/* Options of the Newton solve: the PETSc options dictionary of obstacle_pg.py:128-139 plus the
 * Krylov controls that take the place of ksp_type preonly / pc_type lu / mumps (:129-131). */
typedef struct lvpp_newton_opts {
  double snes_rtol;   /* PETSc default 1e-8; obstacle_pg.py:137 sets 1e-6 */
  double snes_atol;   /* 1e-50 */
  double snes_stol;   /* 1e-8 */
  double snes_divtol; /* 1e4 */
  int32_t snes_max_it;/* 50; obstacle_pg.py:138 sets 100 */
  double ksp_rtol;    /* relative decrease of the preconditioned residual norm */
  double ksp_atol;
  int32_t ksp_max_it;
  int32_t pc_type;    /* LVPP_PC_* */
  int32_t pc_degree;  /* LVPP_PC_MG: smoothing sweeps before and after the coarse correction (0 = 2) */
  int32_t snes_linesearch; /* LVPP_LINESEARCH_*: "none"/"basic" (obstacle_pg.py:136) or PETSc's default "bt"
                              (examples/04_multiphase/multiphase_dolfinx.py:128-143 sets none); lvpp_form_* only */
  int32_t ksp_restart;     /* GMRES restart length of the lvpp_form_* solver (0 = 200) */
  double psi_increase_max; /* 0 = off (the reference: full Newton step, obstacle_pg.py:136).  > 0: no component of the
                              latent variable grows by more than this in one Newton step of the obstacle engine -- a
                              safeguard that is NOT in the reference, for meshes on which its full step overshoots
                              (exp(psi) next to the contact boundary).  Steps whose psi increments stay
                              below the bound are the reference's steps, bit for bit. */
  double psi_free_below;   /* with psi_increase_max > 0: psi may rise freely up to this value and by at most
                              psi_increase_max per step beyond it, psi_new <= max(psi_old, psi_free_below) +
                              psi_increase_max (a node that leaves the contact set comes up from -100 in one step and
                              still cannot overshoot).  -1e300 = none: the bound applies to every increment. */
} lvpp_newton_opts;

#define LVPP_LINESEARCH_NONE 0
#define LVPP_LINESEARCH_BT 1
/* 2 = PETSc's "l2" (examples/03_fracture/fracture_dolfinx.py:132-138): carried in the options for the obstacle engine's
   host loop (proximalgalerkin_b200/linesearch.py over lvpp_assemble_residual / lvpp_linear_solve); the library's own
   Newton loops read it as "none" */
#define LVPP_LINESEARCH_L2 2
#define LVPP_SNES_DIVERGED_LINE_SEARCH (-6)

typedef struct lvpp_stats {
  int64_t num_rows;          /* global rows of the mixed system (2 * owned nodes summed over ranks) */
  int64_t local_rows;        /* 2 * num_owned */
  int64_t nnz;               /* nnz of the monolithic local CSR matrix (4 * scalar nnz) */
  int64_t scalar_nnz;        /* nnz of the scalar node pattern */
  int64_t sell_slots;        /* padded slots of the internal sliced-ELL storage */
  int64_t krylov_iterations; /* cumulative */
  int64_t newton_steps;      /* cumulative */
  int64_t residual_evals;    /* cumulative */
  int64_t kernel_launches;   /* cumulative count of this library's kernel launches */
  int64_t device_bytes;      /* device memory held by the handle */
  double t_assembly_ms;      /* cumulative CUDA-event time in cell + gather + residual kernels */
  double t_krylov_ms;        /* cumulative CUDA-event time in the Krylov solve */
  double last_spmv_ms;       /* mean J*v kernel time of the last lvpp_time_spmv call */
  double spmv_sampled_ms;    /* cumulative CUDA-event time of the J*v launches sampled inside Krylov solves */
  int64_t spmv_samples;      /* number of sampled J*v launches (one per convergence poll) */
  int64_t fine_op_launches;  /* cumulative launches of the block operator kernel on the fine level (J*v, residual
                                and smoother sweeps of the multigrid cycle all run the same kernel) */
  int64_t vcycles;           /* cumulative multigrid V-cycles */
  int32_t mg_levels;         /* levels of the multigrid hierarchy (0 until first used) */
  int32_t reserved0;
  double smooth_sampled_ms;  /* cumulative CUDA-event time of the sampled fine-level smoother sweeps of the multigrid
                                cycle (packed single-precision operator, one sample per V-cycle of a Krylov solve) */
  int64_t smooth_samples;    /* number of sampled smoother sweeps */
  int64_t packed_op_launches;/* cumulative fine-level launches of the cycle's packed operator kernel */
} lvpp_stats;

const char* lvpp_last_error(void);
int lvpp_version(void);
/* number of CUDA devices visible, or LVPP_E_NOGPU */
int lvpp_device_count(void);

/* ---- construction (create_matrix + form compilation, see lvpp_obstacle_desc) ---- */
int lvpp_create(const lvpp_obstacle_desc* desc, lvpp_handle* out);
int lvpp_destroy(lvpp_handle h);
int lvpp_get_stats(lvpp_handle h, lvpp_stats* out);

/* Monolithic CSR pattern of the owned rows (dolfinx.fem.petsc.create_matrix, problem.py:110):
 * h_indptr [2*num_owned + 1], h_indices [nnz] local column numbers, sorted within each row. */
int lvpp_get_csr_pattern(lvpp_handle h, int64_t* h_indptr, int32_t* h_indices);

/* ---- state mutated by the outer proximal loop ---- */
/* alpha.value = ... (obstacle_pg.py:176-183) */
int lvpp_set_alpha(lvpp_handle h, double alpha);
/* f.value = ... : the forcing Constant of obstacle_pg.py:74,122 (dolfinx reads constants at assembly time) */
int lvpp_set_forcing(lvpp_handle h, double f);
/* new Dirichlet values g on (a subset of) the Dirichlet nodes given to lvpp_create: a dolfinx DirichletBC reads its
 * Function / Constant at assembly time (u_bc.x.array[...] = ..., signorini_dolfinx.py:322).  Host arrays. */
int lvpp_set_bc_values(lvpp_handle h, int64_t num_bc, const int32_t* h_bc_nodes, const double* h_bc_values);
/* sol_k.x.array[:] = ... (obstacle_pg.py:158,226); d_xk [2*num_nodes] */
int lvpp_set_previous(lvpp_handle h, const double* d_xk);

/* ---- assembly (SNESProblem.F / SNESProblem.J, src/lvpp/problem.py:54-77) ---- */
/* residual with lifting and set_bc; d_x, d_F [2*num_nodes] (owned rows of d_F written); also leaves
 * the Jacobian at d_x assembled.  h_fnorm (optional) receives ||F||_2 over all ranks. */
int lvpp_assemble_residual(lvpp_handle h, const double* d_x, double* d_F, double* h_fnorm);
/* Jacobian at d_x into the handle's internal storage (zeroEntries + assemble_matrix + assemble). */
int lvpp_assemble_jacobian(lvpp_handle h, const double* d_x);
/* values of the last assembled Jacobian on the pattern of lvpp_get_csr_pattern, Dirichlet rows and
 * columns zeroed with unit diagonal; d_values [nnz] */
int lvpp_get_jacobian_values(lvpp_handle h, double* d_values);

/* ---- linear algebra (PETSc MatMult / KSPSolve behind SNES, obstacle_pg.py:129-131) ---- */
/* y = J v with the last assembled Jacobian (halo exchange of v included) */
int lvpp_spmv(lvpp_handle h, const double* d_v, double* d_y);
/* block-preconditioned MINRES on J y = rhs, y0 = 0 */
int lvpp_linear_solve(lvpp_handle h, const double* d_rhs, double* d_y, const lvpp_newton_opts* opts,
                      int32_t* its, int32_t* reason, double* h_rnorm);

/* ---- the Newton loop (SNESSolver.solve, problem.py:114-124; PETSc newtonls, line search none) */
int lvpp_newton_solve(lvpp_handle h, double* d_x, const lvpp_newton_opts* opts, int32_t* its,
                      int32_t* reason, double* h_fnorm, int32_t* linear_its);
/* one Newton step from the residual state left by the previous call (bench granularity):
 * begin: evaluates F(x) and the Jacobian; step: solve, update, re-evaluate; returns norms
 * h_norms = {fnorm, ynorm, xnorm} */
int lvpp_newton_begin(lvpp_handle h, const double* d_x, double* h_fnorm);
/* same, when d_x is the iterate the previous Newton solve on this handle ended at and only alpha, f, the Dirichlet values
 * or the previous iterate changed since (the start of the next proximal step, obstacle_pg.py:175-190,226): D(psi) is
 * kept instead of being re-assembled -- the result is bit-identical to lvpp_newton_begin */
int lvpp_newton_begin_same_iterate(lvpp_handle h, const double* d_x, double* h_fnorm);
int lvpp_newton_step(lvpp_handle h, double* d_x, const lvpp_newton_opts* opts, double* h_norms,
                     int32_t* ksp_its, int32_t* ksp_reason);

/* ---- observables of the outer loop (obstacle_pg.py:145-152,196-201): energy, complementarity
 * (signed), feasibility, dual feasibility, H1 increment squared, latent increment squared;
 * summed over ranks.  d_x is the current iterate, the previous one is lvpp_set_previous's. */
int lvpp_observables(lvpp_handle h, const double* d_x, double* h_out6);

/* ---- host-buffer entry points (what a ctypes/cffi binding of the reference calls with numpy
 * arrays): same as above with the host<->device copies inside. */
int lvpp_newton_solve_host(lvpp_handle h, double* h_x, const lvpp_newton_opts* opts, int32_t* its,
                           int32_t* reason, double* h_fnorm, int32_t* linear_its);
int lvpp_set_previous_host(lvpp_handle h, const double* h_xk);
/* the last Newton iterate of lvpp_newton_solve_host whatever its reason (dolfinx.fem.petsc.NonlinearProblem.solve leaves
 * it in u, obstacle_pg.py:190; SNESSolver.solve of src/lvpp/problem.py:121-123 keeps u on failure); h_x [2 * num_nodes] */
int lvpp_get_last_iterate_host(lvpp_handle h, double* h_x);

/* ---- measurement helpers ---- */
/* CUDA events on the handle's stream: start, then stop returns the elapsed device time in ms */
int lvpp_timer_start(lvpp_handle h);
int lvpp_timer_stop(lvpp_handle h, double* h_ms);
/* runs the J*v kernel `reps` times on the handle's stream between CUDA events and returns the mean
 * kernel time in ms; flush_l2 != 0 writes a >L2 scratch buffer between repetitions (untimed). */
int lvpp_time_spmv(lvpp_handle h, const double* d_v, double* d_y, int32_t reps, int32_t flush_l2,
                   double* h_ms);
/* same for the cell kernel + row gather (Jacobian assembly) and the residual kernel */
int lvpp_time_assembly(lvpp_handle h, const double* d_x, double* d_F, int32_t reps,
                       double* h_ms_cells, double* h_ms_gather, double* h_ms_residual);

/* ---- multi-GPU (replaces the MPI communicator of the reference) ---- */
/* 128-byte NCCL unique id, generated on one rank and broadcast by the host side */
int lvpp_comm_unique_id(uint8_t* h_id128);
int lvpp_comm_init(lvpp_handle h, const uint8_t* h_id128, int32_t rank, int32_t nranks);
/* owner -> ghost update of a mixed vector (Vec.ghostUpdate(INSERT, FORWARD), problem.py:56) */
int lvpp_halo_forward(lvpp_handle h, double* d_v);

/* =====================================================================================================
 * The other LVPP formulations of the reference behind one generic
 * mixed-form engine: per-formulation element kernels, an atomic-free cell-to-nnz gather into a CSR
 * matrix (the sparsity of dolfinx.fem.petsc.create_matrix: union over integration entities of
 * dofs x dofs), fp64 CSR SpMV, dof-block-Jacobi preconditioned restarted GMRES and the SNES newtonls
 * loop with line search none or bt.  Single GPU.
 *
 *   LVPP_FORM_GRADIENT    examples/06_gradient_constraints/gradient_constraint_dolfinx.py:100-107
 *       u in P2, psi in (P1)^2 on triangles.  One integral (cells).  Local dof order: 6 u nodes, then
 *       (psi_x, psi_y) of the 3 vertices.  params = {alpha}.  aux0 = w0 (previous iterate, :205),
 *       coef0 = phi and coef1 = f interpolated into the primal space (:56-62), indexed by u dof.
 *       tab_a = P2 basis [nq*6], dtab_a = its reference gradients [nq*6*2], tab_b = P1 basis [nq*3].
 *   LVPP_FORM_MULTIPHASE  examples/04_multiphase/multiphase_dolfinx.py:64-90
 *       u, z, psi in (P1)^4 on triangles; local dof order vertex-major, (u0..3, z0..3, psi0..3) per vertex.
 *       params = {alpha, tau, eps0, h_scale} with epsilon = h_scale * circumradius (:52-53: 4).
 *       aux0 = lvpp_old (:60, psi slots read), aux1 = u_prev in the u slots (:57).  tab_a = P1 basis [nq*3].
 *   LVPP_FORM_SIGNORINI   examples/02_signorini/signorini_dolfinx.py:244-249
 *       integral 0: tetrahedra, u in (P1)^3, local order vertex-major / component fastest;
 *       integral 1: contact facets (triangles), 9 u dofs then the 3 psi dofs of the facet submesh.
 *       params = {alpha, mu, lambda, gap, n_g[0], n_g[1], n_g[2]}.  aux0 = psi_k in the psi slots (:344).
 *       integral 1: tab_a = P1 facet basis [nq*3].
 * Vectors are plain fp64 device arrays of length num_dofs.
 * ===================================================================================================== */
#define LVPP_FORM_GRADIENT 1
#define LVPP_FORM_MULTIPHASE 2
#define LVPP_FORM_SIGNORINI 3

typedef struct lvpp_form_problem* lvpp_form_handle;

typedef struct lvpp_integral_desc {
  int64_t num_entities;
  int32_t nld;               /* mixed local dofs per entity (element matrix is nld x nld) */
  int32_t nv;                /* geometry vertices per entity */
  const int32_t* dofs;       /* host [num_entities * nld] */
  const int32_t* vertices;   /* host [num_entities * nv] */
  const int64_t* to_nnz;     /* host [num_entities * nld * nld]: CSR position of every element-matrix entry */
  int32_t nq;
  int32_t reserved0;
  const double* qweights;    /* host [nq] */
  const double* tab_a;       /* see the formulation list above */
  const double* dtab_a;
  const double* tab_b;
} lvpp_integral_desc;

typedef struct lvpp_form_desc {
  int32_t form;              /* LVPP_FORM_* */
  int32_t gdim;
  int64_t num_dofs;
  int64_t num_vertices;
  const double* vertex_coords;  /* host [num_vertices * gdim] */
  const int64_t* indptr;     /* host [num_dofs + 1]  CSR pattern, columns sorted within a row */
  const int32_t* indices;    /* host [nnz] */
  int64_t num_bc;
  const int64_t* bc_dofs;    /* host [num_bc] */
  const double* bc_values;   /* host [num_bc] */
  int32_t num_integrals;
  int32_t num_params;
  const lvpp_integral_desc* integrals;
  const double* params;      /* host [num_params] */
  const double* coef0;       /* host [num_dofs] or NULL */
  const double* coef1;       /* host [num_dofs] or NULL */
  int64_t num_blocks;        /* dof blocks of the Jacobi preconditioner (e.g. the dofs of one mesh node) */
  const int64_t* block_ptr;  /* host [num_blocks + 1] */
  const int32_t* block_dofs; /* host [block_ptr[num_blocks]], every dof in exactly one block, block size <= 16 */
} lvpp_form_desc;

int lvpp_form_create(const lvpp_form_desc* desc, lvpp_form_handle* out);
int lvpp_form_destroy(lvpp_form_handle h);
/* alpha.value = ... and the other constants of the form */
int lvpp_form_set_param(lvpp_form_handle h, int32_t index, double value);
/* previous-iterate data read by the residual (w0 / lvpp_old / u_prev / psi_k): which = 0, 1; d_v [num_dofs] */
int lvpp_form_set_aux(lvpp_form_handle h, int32_t which, const double* d_v);
/* Dirichlet values g (u_bc.x.array[...] = disp, signorini_dolfinx.py:322): h_values [num_bc] in bc_dofs order */
int lvpp_form_set_bc_values(lvpp_form_handle h, const double* h_values);
/* residual with lifting and set_bc (src/lvpp/problem.py:54-67); leaves the Jacobian at d_x assembled */
int lvpp_form_assemble_residual(lvpp_form_handle h, const double* d_x, double* d_F, double* h_fnorm);
/* CSR values of the last assembled Jacobian (assemble_matrix with bcs, problem.py:75-77); d_values [nnz] */
int lvpp_form_get_jacobian_values(lvpp_form_handle h, double* d_values);
int lvpp_form_spmv(lvpp_form_handle h, const double* d_v, double* d_y);
int lvpp_form_linear_solve(lvpp_form_handle h, const double* d_rhs, double* d_y, const lvpp_newton_opts* opts,
                           int32_t* its, int32_t* reason, double* h_rnorm);
int lvpp_form_newton_solve(lvpp_form_handle h, double* d_x, const lvpp_newton_opts* opts, int32_t* its,
                           int32_t* reason, double* h_fnorm, int32_t* linear_its);
/* assemble_scalar of |primal(x) - primal(x0)|^2 dx (gradient_constraint_dolfinx.py:166-168,
 * multiphase_dolfinx.py:166-169) or, for LVPP_FORM_SIGNORINI, the discrete sum over the u dofs of
 * (x - x0)^2 (signorini_dolfinx.py:337-339) */
int lvpp_form_increment_sq(lvpp_form_handle h, const double* d_x, const double* d_x0, double* h_out);
int lvpp_form_get_stats(lvpp_form_handle h, lvpp_stats* out);
/* mean device time (ms) of `reps` launches of the element kernels + gathers (assembly) and of the CSR SpMV */
int lvpp_form_time_kernels(lvpp_form_handle h, const double* d_x, int32_t reps, double* h_ms_assembly,
                           double* h_ms_spmv);

#ifdef __cplusplus
}
#endif
#endif /* LVPP_B200_H */
