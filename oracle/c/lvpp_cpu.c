/* CPU restatement (plain C + OpenMP) of two kernels of the reference path, on the reference's own data layout --
 * TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/__init__.py; only tests/ and bench.py's cpu_baseline
 * leg load this library).
 *
 *   lvpp_cpu_assemble_jacobian   SNESProblem.J, src/lvpp/problem.py:69-77: J.zeroEntries(); assemble_matrix(J, a, bcs);
 *                                cell loop calling the tabulate_tensor of J = derivative(F, sol)
 *                                (examples/01_obstacle_problem/obstacle_pg.py:116-125), rows / columns of Dirichlet
 *                                dofs zeroed in the element tensor, MatSetValuesLocal(ADD) through the cell-to-nnz
 *                                map of create_matrix's pattern, unit diagonal on Dirichlet rows.
 *   lvpp_cpu_csr_spmv            PETSc MatMult on the assembled AIJ matrix.
 *
 * dolfinx runs one thread per MPI rank (docker/Dockerfile:320, OMP_NUM_THREADS=1) and partitions the cells over
 * ranks; here the cell loop is partitioned over OpenMP threads and the ADD is an atomic update -- the shared-memory
 * equivalent of the same algorithm, so that all host cores of the GPU box can be used for the baseline.
 * Same arithmetic, in the same order per cell, as oracle/obstacle.py:element_jacobian / assemble_jacobian_values.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAX_NLD 10
#define MAX_GDIM 3

int lvpp_cpu_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* cell_nodes [C][nld] scalar node of every local basis function; dof_u / dof_psi [N] rows of the mixed system;
 * phi [nq][nld], gphi [C][nq][nld][gdim] physical gradients, ws = |det J| (scale [C]) * w [nq];
 * is_bc [rows]; cell_to_nnz [C][2 nld][2 nld] position of every element entry in the CSR value array;
 * diag_pos [rows]; vals [nnz] (overwritten). */
void lvpp_cpu_assemble_jacobian(int64_t C, int nld, int nq, int gdim, const int64_t* cell_nodes, const int64_t* dof_u,
                                const int64_t* dof_psi, const double* phi, const double* gphi, const double* scale,
                                const double* w, const uint8_t* is_bc, const int64_t* cell_to_nnz,
                                const int64_t* diag_pos, const int64_t* bc_dofs, int64_t num_bc, const double* x,
                                double alpha, int64_t nnz, double* vals) {
  const int n2 = 2 * nld;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nnz; ++i) vals[i] = 0.0; /* J.zeroEntries() */
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < C; ++c) {
    double K[MAX_NLD][MAX_NLD], M[MAX_NLD][MAX_NLD], D[MAX_NLD][MAX_NLD], psi[MAX_NLD];
    int64_t rows[2 * MAX_NLD];
    for (int a = 0; a < nld; ++a) {
      const int64_t node = cell_nodes[c * nld + a];
      psi[a] = x[dof_psi[node]];
      rows[a] = dof_u[node];
      rows[nld + a] = dof_psi[node];
      for (int b = 0; b < nld; ++b) K[a][b] = M[a][b] = D[a][b] = 0.0;
    }
    const double* g = gphi + (size_t)c * nq * nld * gdim;
    for (int q = 0; q < nq; ++q) {
      const double wsq = w[q] * scale[c];
      double pq = 0.0;
      for (int a = 0; a < nld; ++a) pq += psi[a] * phi[q * nld + a];
      const double e = wsq * exp(pq);
      for (int a = 0; a < nld; ++a) {
        const double pa = phi[q * nld + a];
        for (int b = 0; b < nld; ++b) {
          double gg = 0.0;
          for (int d = 0; d < gdim; ++d) gg += g[(q * nld + a) * gdim + d] * g[(q * nld + b) * gdim + d];
          K[a][b] += wsq * gg;
          M[a][b] += wsq * pa * phi[q * nld + b];
          D[a][b] += e * pa * phi[q * nld + b];
        }
      }
    }
    const int64_t* map = cell_to_nnz + (size_t)c * n2 * n2;
    for (int i = 0; i < n2; ++i) {
      for (int j = 0; j < n2; ++j) {
        if (is_bc[rows[i]] || is_bc[rows[j]]) continue; /* assemble_matrix(..., bcs): zeroed in the element tensor */
        const int a = i < nld ? i : i - nld, b = j < nld ? j : j - nld;
        double v;
        if (i < nld) v = j < nld ? alpha * K[a][b] : M[a][b];
        else v = j < nld ? M[a][b] : -D[a][b];
#pragma omp atomic
        vals[map[i * n2 + j]] += v;
      }
    }
  }
  for (int64_t k = 0; k < num_bc; ++k) vals[diag_pos[bc_dofs[k]]] = 1.0; /* diagonal 1.0 on Dirichlet rows */
}

/* y = A x, CSR (MatMult_SeqAIJ) */
void lvpp_cpu_csr_spmv(int64_t n, const int64_t* indptr, const int32_t* indices, const double* vals, const double* x,
                       double* y) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    double s = 0.0;
    for (int64_t p = indptr[i]; p < indptr[i + 1]; ++p) s += vals[p] * x[indices[p]];
    y[i] = s;
  }
}

/* ------------------------------------------------------------------------------------------------------------------
 * lvpp_cpu_minres: the Krylov solve the north star puts in place of the reference's MUMPS LU (KSP preonly + PC lu,
 * examples/01_obstacle_problem/obstacle_pg.py:129-131), restated on the CPU for a like-for-like baseline on all host
 * cores: preconditioned MINRES (Paige & Saunders 1975; the recurrences as in Choi, Paige & Saunders' MINRES-QLP
 * report, section 2) on the assembled monolithic CSR with a positive diagonal preconditioner -- the recipe of
 * examples/09_eikonal/ex40.cpp:261-274 reduced to its diagonals: 1 / (alpha K_ii) on the u rows and
 * 1 / (D_ii + M_ii^2 / (alpha K_ii)) on the psi rows (passed in as pinv).  Stops on the preconditioned residual norm
 * <= rtol * initial.  Returns the iteration count; y0 = 0. */
static double dot_omp(int64_t n, const double* a, const double* b) {
  double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

int64_t lvpp_cpu_minres(int64_t n, const int64_t* indptr, const int32_t* indices, const double* vals, const double* pinv,
                        const double* rhs, double* y, double rtol, int64_t maxit, double* work /* [7 n] */,
                        double* rnorm_out) {
  double *r1 = work, *r2 = work + n, *v = work + 2 * n, *z = work + 3 * n, *w = work + 4 * n, *w1 = work + 5 * n,
         *w2 = work + 6 * n;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    y[i] = 0.0; r1[i] = rhs[i]; r2[i] = rhs[i]; z[i] = pinv[i] * rhs[i]; w[i] = 0.0; w2[i] = 0.0;
  }
  double beta1 = dot_omp(n, rhs, z);
  if (!(beta1 > 0.0)) { if (rnorm_out) *rnorm_out = 0.0; return 0; }
  beta1 = sqrt(beta1);
  double oldb = 0.0, beta = beta1, dbar = 0.0, epsln = 0.0, phibar = beta1, cs = -1.0, sn = 0.0;
  int64_t it = 0;
  while (it < maxit) {
    ++it;
    const double s = 1.0 / beta;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) v[i] = s * z[i];
    /* z <- A v (reuse z as the product) */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      double a = 0.0;
      for (int64_t p = indptr[i]; p < indptr[i + 1]; ++p) a += vals[p] * v[indices[p]];
      z[i] = a;
    }
    if (it >= 2) {
      const double c = beta / oldb;
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < n; ++i) z[i] -= c * r1[i];
    }
    const double alfa = dot_omp(n, v, z);
    {
      const double c = alfa / beta;
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < n; ++i) {
        const double t = z[i] - c * r2[i];
        r1[i] = r2[i];
        r2[i] = t;
        z[i] = pinv[i] * t;
      }
    }
    oldb = beta;
    beta = dot_omp(n, r2, z);
    if (beta < 0.0) break; /* preconditioner not positive */
    beta = sqrt(beta);
    /* previous rotation */
    const double oldeps = epsln;
    const double delta = cs * dbar + sn * alfa;
    const double gbar = sn * dbar - cs * alfa;
    epsln = sn * beta;
    dbar = -cs * beta;
    /* next rotation */
    double gamma = sqrt(gbar * gbar + beta * beta);
    if (gamma < 1e-300) gamma = 1e-300;
    cs = gbar / gamma;
    sn = beta / gamma;
    const double phi = cs * phibar;
    phibar = sn * phibar;
    const double denom = 1.0 / gamma;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      const double t1 = w2[i];
      w1[i] = t1;
      const double t2 = w[i];
      w2[i] = t2;
      const double wn = (v[i] - oldeps * t1 - delta * t2) * denom;
      w[i] = wn;
      y[i] += phi * wn;
    }
    if (phibar <= rtol * beta1 || beta == 0.0) break;
  }
  if (rnorm_out) *rnorm_out = phibar;
  return it;
}
