// Newton loop: PETSc SNES "newtonls" with line search "none" as configured at
// examples/01_obstacle_problem/obstacle_pg.py:128-139 and driven by SNESSolver.solve
// (src/lvpp/problem.py:114-124): F(x0); repeat { J(x); solve J y = F; x <- x - y; F(x); norms; test }.
// Convergence logic = SNESConvergedDefault (atol, rtol relative to ||F0||, stol on the step, divtol).
#include <cmath>

#include "lvpp_internal.cuh"

// x <- x - y on the owned rows; partial ||y||^2 and ||x_new||^2  (the fused Newton update + norms).
// cap > 0 (lvpp_newton_opts.psi_increase_max / psi_free_below, not in the reference):
// psi_new <= max(psi_old, free_below) + cap.
__global__ void __launch_bounds__(256)
k_newton_update(int64_t Vown, double2* __restrict__ x, const double2* __restrict__ y, int nparts,
                double* __restrict__ partials, double cap, double free_below) {
  __shared__ double s_red[32];
  double py = 0.0, px = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < Vown;
       i += (int64_t)gridDim.x * blockDim.x) {
    double2 a = x[i];
    double2 b = y[i];
    if (cap > 0.0) {
      const double top = fmax(a.y, free_below) + cap;  // the highest psi this step may reach
      if (a.y - b.y > top) b.y = a.y - top;
    }
    a.x -= b.x;
    a.y -= b.y;
    x[i] = a;
    py += b.x * b.x + b.y * b.y;
    px += a.x * a.x + a.y * a.y;
  }
  const double ry = lvpp_block_sum<256>(py, s_red);
  if (threadIdx.x == 0) partials[blockIdx.x] = ry;
  const double rx = lvpp_block_sum<256>(px, s_red);
  if (threadIdx.x == 0) partials[nparts + blockIdx.x] = rx;
}

static int snes_converged(int it, double xnorm, double snorm, double fnorm, double ttol, double fnorm0,
                          const lvpp_newton_opts* o) {
  if (std::isnan(fnorm) || std::isinf(fnorm)) return LVPP_SNES_DIVERGED_FNORM_NAN;
  if (fnorm < o->snes_atol) return LVPP_SNES_CONVERGED_FNORM_ABS;
  if (it) {
    if (fnorm <= ttol) return LVPP_SNES_CONVERGED_FNORM_RELATIVE;
    if (snorm < o->snes_stol * xnorm) return LVPP_SNES_CONVERGED_SNORM_RELATIVE;
    if (o->snes_divtol > 0 && fnorm > o->snes_divtol * fnorm0) return LVPP_SNES_DIVERGED_DTOL;
  }
  return 0;
}

static int check_opts(const lvpp_newton_opts* o) {
  if (!o) { lvpp_set_error("null options"); return LVPP_E_INVALID; }
  if (!(o->ksp_rtol >= 0) || !(o->ksp_atol >= 0) || o->snes_max_it < 0) {
    lvpp_set_error("invalid solver options");
    return LVPP_E_INVALID;
  }
  return 0;
}

static int newton_begin(lvpp_handle h, const double* d_x, double* h_fnorm, bool same_iterate);

extern "C" int lvpp_newton_begin(lvpp_handle h, const double* d_x, double* h_fnorm) {
  return newton_begin(h, d_x, h_fnorm, false);
}

// The Newton solve of the previous proximal step ended at d_x and nothing but alpha, f, the Dirichlet values or the
// previous iterate changed since (obstacle_pg.py:175-190,226: `alpha.value = ...; sol_k <- sol; problem.solve()`):
// D(psi) -- the only part of the Jacobian that depends on the iterate -- is the one of the last residual evaluation and
// is kept; only the residual is evaluated.  (PETSc re-assembles J(x0) there; the result is bit-identical.)  Falls back
// to the full evaluation when no Jacobian is valid.
extern "C" int lvpp_newton_begin_same_iterate(lvpp_handle h, const double* d_x, double* h_fnorm) {
  return newton_begin(h, d_x, h_fnorm, true);
}

static int newton_begin(lvpp_handle h, const double* d_x, double* h_fnorm, bool same_iterate) {
  if (!h || !d_x) { lvpp_set_error("null argument"); return LVPP_E_INVALID; }
  CK(cudaSetDevice(h->device));
  CK(cudaEventRecord(h->ev0, h->stream));
  CKR(lvpp_eval_residual(h, d_x, h->F, true, same_iterate));
  CK(cudaMemcpyAsync(h->red_host, h->scal->red, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaEventRecord(h->ev1, h->stream));
  CKR(lvpp_sync_check_comm(h));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->t_assembly_ms += ms;
  h->fnorm = sqrt(h->red_host[0]);
  if (h_fnorm) *h_fnorm = h->fnorm;
  return LVPP_OK;
}

extern "C" int lvpp_newton_step(lvpp_handle h, double* d_x, const lvpp_newton_opts* opts, double* h_norms,
                                int32_t* ksp_its, int32_t* ksp_reason) {
  if (!h || !d_x) { lvpp_set_error("null argument"); return LVPP_E_INVALID; }
  CKR(check_opts(opts));
  CK(cudaSetDevice(h->device));
  if (!h->jac_valid) { lvpp_set_error("lvpp_newton_begin has not been called"); return LVPP_E_INVALID; }
  int32_t kits = 0, kreason = 0;
  h->gm_warm_next = h->y_is_newton_correction;  // (multigrid.cu: lvpp_gmres_mg uses it only if it beats the zero guess)
  CKR(lvpp_solve_linear(h, h->F, h->y, opts, &kits, &kreason, nullptr));
  h->y_is_newton_correction = kreason > 0;
  if (ksp_its) *ksp_its = kits;
  if (ksp_reason) *ksp_reason = kreason;
  CK(cudaEventRecord(h->ev0, h->stream));
  LAUNCH(h, k_newton_update, h->npartials, 256, 0, h->Vown, (double2*)d_x, (const double2*)h->y, h->npartials,
         h->partials, opts->psi_increase_max, opts->psi_free_below);
  CK(cudaGetLastError());
  CKR(lvpp_reduce_partials(h, 2, h->scal->red + 1));
  CKR(lvpp_eval_residual(h, d_x, h->F, true));  // writes red[0]
  CK(cudaMemcpyAsync(h->red_host, h->scal->red, 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaEventRecord(h->ev1, h->stream));
  CKR(lvpp_sync_check_comm(h));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->t_assembly_ms += ms;
  h->fnorm = sqrt(h->red_host[0]);
  h->newton_steps++;
  if (h_norms) {
    h_norms[0] = h->fnorm;
    h_norms[1] = sqrt(h->red_host[1]);
    h_norms[2] = sqrt(h->red_host[2]);
  }
  return LVPP_OK;
}

extern "C" int lvpp_newton_solve(lvpp_handle h, double* d_x, const lvpp_newton_opts* opts, int32_t* its,
                                 int32_t* reason, double* h_fnorm, int32_t* linear_its) {
  if (!h || !d_x) { lvpp_set_error("null argument"); return LVPP_E_INVALID; }
  CKR(check_opts(opts));
  double fnorm = 0.0;
  CKR(lvpp_newton_begin(h, d_x, &fnorm));
  const double fnorm0 = fnorm, ttol = fnorm * opts->snes_rtol;
  int r = snes_converged(0, 0.0, 0.0, fnorm, ttol, fnorm0, opts);
  int it = 0, lin = 0;
  while (!r) {
    if (it >= opts->snes_max_it) { r = LVPP_SNES_DIVERGED_MAX_IT; break; }
    double norms[3];
    int32_t kits = 0, kreason = 0;
    CKR(lvpp_newton_step(h, d_x, opts, norms, &kits, &kreason));
    lin += kits;
    ++it;
    if (kreason < 0) { r = LVPP_SNES_DIVERGED_LINEAR_SOLVE; fnorm = norms[0]; break; }
    fnorm = norms[0];
    r = snes_converged(it, norms[2], norms[1], fnorm, ttol, fnorm0, opts);
  }
  if (its) *its = it;
  if (reason) *reason = r;
  if (h_fnorm) *h_fnorm = fnorm;
  if (linear_its) *linear_its = lin;
  return LVPP_OK;
}

extern "C" int lvpp_newton_solve_host(lvpp_handle h, double* h_x, const lvpp_newton_opts* opts, int32_t* its,
                                      int32_t* reason, double* h_fnorm, int32_t* linear_its) {
  if (!h || !h_x) { lvpp_set_error("null argument"); return LVPP_E_INVALID; }
  CK(cudaSetDevice(h->device));
  const size_t bytes = sizeof(double) * 2 * h->V;
  CK(cudaMemcpyAsync(h->xhost_stage, h_x, bytes, cudaMemcpyHostToDevice, h->stream));
  int32_t r = 0;
  CKR(lvpp_newton_solve(h, h->xhost_stage, opts, its, &r, h_fnorm, linear_its));
  if (reason) *reason = r;
  // SNESSolver.solve only overwrites the caller's function on convergence (problem.py:121-123)
  if (r > 0) {
    CK(cudaMemcpyAsync(h_x, h->xhost_stage, bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return LVPP_OK;
}

// The last iterate of lvpp_newton_solve_host whatever its reason: dolfinx.fem.petsc.NonlinearProblem.solve leaves the
// last Newton iterate in u (obstacle_pg.py:190; the recovery loops of examples 03 / 07 / 08 restore it themselves),
// while lvpp's SNESSolver.solve writes u only on convergence (src/lvpp/problem.py:121-123).
extern "C" int lvpp_get_last_iterate_host(lvpp_handle h, double* h_x) {
  if (!h || !h_x) { lvpp_set_error("null argument"); return LVPP_E_INVALID; }
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(h_x, h->xhost_stage, sizeof(double) * 2 * h->V, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return LVPP_OK;
}
