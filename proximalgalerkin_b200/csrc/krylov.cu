// Block-preconditioned MINRES for the symmetric indefinite Newton system
//   J = [[alpha K, M], [M, -D(psi)]]
// taking the place of the reference's direct solve (ksp_type preonly / pc_type lu / MUMPS,
// examples/01_obstacle_problem/obstacle_pg.py:129-131).  The preconditioner follows the only
// iterative saddle-point recipe in the reference, examples/09_eikonal/ex40.cpp:261-274: a
// (smoothed) diagonal for the (0,0) block and the diagonal of D + M diag(alpha K)^-1 M for the Schur
// complement.  All recurrence scalars live on the device (KryScal); the host only polls a
// convergence flag every few iterations, so an iteration is five back-to-back launches.
#include <cstring>

#include "lvpp_internal.cuh"

// pinv_u = 1 / (alpha K_ii)  (1 on Dirichlet rows);  pinv_psi = 1 / (D_ii + M_ii^2 / (alpha K_ii))
__global__ void k_build_pinv(int64_t Vown, const int64_t* __restrict__ slice_ptr,
                             const uint8_t* __restrict__ diag_k, const double* __restrict__ K,
                             const double* __restrict__ M, const double* __restrict__ D,
                             const uint8_t* __restrict__ bc_flag, double alpha, double2* __restrict__ pinv) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < Vown;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t idx = slice_ptr[i >> 5] + (i & 31) + (int64_t)diag_k[i] * LVPP_SLICE;
    const double a = alpha * K[idx], m = M[idx], d = D[idx];
    double2 p;
    if (bc_flag[i]) {
      p.x = 1.0;
      p.y = 1.0 / d;
    } else {
      p.x = 1.0 / a;
      p.y = 1.0 / (d + m * m / a);
    }
    pinv[i] = p;
  }
}

// v1 = rhs, z1 = P^-1 v1, partial z1.v1; zero v0, w0, w1, y
__global__ void __launch_bounds__(256)
k_minres_init(int64_t Vown, const double2* __restrict__ rhs, const double2* __restrict__ pinv,
              double2* __restrict__ v0, double2* __restrict__ v1, double2* __restrict__ z1,
              double2* __restrict__ w0, double2* __restrict__ w1, double2* __restrict__ y,
              double* __restrict__ partials) {
  __shared__ double s_red[32];
  double part = 0.0;
  const double2 zero = make_double2(0.0, 0.0);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < Vown;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double2 r = rhs[i], p = pinv[i];
    const double2 z = make_double2(p.x * r.x, p.y * r.y);
    v1[i] = r; z1[i] = z; v0[i] = zero; w0[i] = zero; w1[i] = zero; y[i] = zero;
    part += z.x * r.x + z.y * r.y;
  }
  const double r = lvpp_block_sum<256>(part, s_red);
  if (threadIdx.x == 0) partials[blockIdx.x] = r;
}

__global__ void k_minres_scal_init(KryScal* s, double rtol, double atol, int maxit) {
  const double g = sqrt(fmax(s->red[0], 0.0));
  s->gamma0 = 1.0; s->gamma1 = g; s->eta = g; s->s0 = 0.0; s->s1 = 0.0; s->c0 = 1.0; s->c1 = 1.0;
  s->r0 = g;
  s->inv_gamma1 = g > 0.0 ? 1.0 / g : 0.0;
  s->rtol = rtol; s->atol = atol;
  s->tol = fmax(rtol * g, atol);
  s->its = 0; s->maxit = maxit; s->skip = 0;
  if (!(g == g) || isinf(g)) { s->conv = 1; s->reason = LVPP_KSP_DIVERGED_NANORINF; }
  else if (g <= atol || g == 0.0) { s->conv = 1; s->reason = LVPP_KSP_CONVERGED_ATOL; }
  else { s->conv = 0; s->reason = 0; }
}

// after J z: delta = (J z).z  -> Lanczos coefficients
__global__ void k_minres_scal1(KryScal* s) {
  s->skip = s->conv;
  if (s->skip) return;
  const double delta = s->red[0];
  s->delta = delta;
  s->cv1 = delta / s->gamma1;
  s->cv0 = s->gamma1 / s->gamma0;
}

// v2 = Az - cv1 v1 - cv0 v0 (written over v0); z2 = P^-1 v2; partial z2.v2
__global__ void __launch_bounds__(256, 6)
k_minres_lanczos(int64_t Vown, const KryScal* __restrict__ s, const double2* __restrict__ Az,
                 const double2* __restrict__ v1, double2* __restrict__ v0, const double2* __restrict__ pinv,
                 double2* __restrict__ z2, double* __restrict__ partials) {
  __shared__ double s_red[32];
  if (s->skip) return;
  const double c1 = s->cv1, c0 = s->cv0;
  double part = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < Vown;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double2 a = Az[i], b = v1[i], c = v0[i], p = pinv[i];
    double2 v;
    v.x = a.x - c1 * b.x - c0 * c.x;
    v.y = a.y - c1 * b.y - c0 * c.y;
    const double2 z = make_double2(p.x * v.x, p.y * v.y);
    v0[i] = v;
    z2[i] = z;
    part += z.x * v.x + z.y * v.y;
  }
  const double r = lvpp_block_sum<256>(part, s_red);
  if (threadIdx.x == 0) partials[blockIdx.x] = r;
}

// Givens rotation, residual-norm update, convergence test
__global__ void k_minres_scal2(KryScal* s) {
  if (s->skip) return;
  const double g2 = sqrt(fmax(s->red[0], 0.0));
  const double delta = s->delta, g1 = s->gamma1;
  const double a0 = s->c1 * delta - s->c0 * s->s1 * g1;
  const double a1 = sqrt(a0 * a0 + g2 * g2);
  const double a2 = s->s1 * delta + s->c0 * s->c1 * g1;
  const double a3 = s->s0 * g1;
  const double c2 = a1 != 0.0 ? a0 / a1 : 1.0;
  const double s2 = a1 != 0.0 ? g2 / a1 : 0.0;
  s->a1 = a1 != 0.0 ? a1 : 1.0;
  s->a2 = a2; s->a3 = a3;
  s->tau = c2 * s->eta;
  s->eta = -s2 * s->eta;
  s->inv_gamma_w = s->inv_gamma1;
  s->gamma0 = g1; s->gamma1 = g2;
  s->inv_gamma1 = g2 > 0.0 ? 1.0 / g2 : 0.0;
  s->c0 = s->c1; s->c1 = c2; s->s0 = s->s1; s->s1 = s2;
  s->its += 1;
  const double res = fabs(s->eta);
  if (!(res == res) || isinf(res) || !(a1 == a1)) { s->conv = 1; s->reason = LVPP_KSP_DIVERGED_NANORINF; }
  else if (res <= s->tol) { s->conv = 1; s->reason = res <= s->atol ? LVPP_KSP_CONVERGED_ATOL : LVPP_KSP_CONVERGED_RTOL; }
  else if (g2 == 0.0) { s->conv = 1; s->reason = LVPP_KSP_DIVERGED_BREAKDOWN; }
  else if (s->its >= s->maxit) { s->conv = 1; s->reason = LVPP_KSP_DIVERGED_ITS; }
}

// w2 = (z1 / gamma1 - a3 w0 - a2 w1) / a1 (written over w0);  y += tau w2
__global__ void __launch_bounds__(256)
k_minres_update(int64_t Vown, const KryScal* __restrict__ s, const double2* __restrict__ z1,
                double2* __restrict__ w0, const double2* __restrict__ w1, double2* __restrict__ y) {
  if (s->skip) return;
  const double ig = s->inv_gamma_w, a3 = s->a3, a2 = s->a2, ia1 = 1.0 / s->a1, tau = s->tau;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < Vown;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double2 z = z1[i], a = w0[i], b = w1[i];
    double2 w, yy = y[i];
    w.x = (z.x * ig - a3 * a.x - a2 * b.x) * ia1;
    w.y = (z.y * ig - a3 * a.y - a2 * b.y) * ia1;
    yy.x += tau * w.x;
    yy.y += tau * w.y;
    w0[i] = w;
    y[i] = yy;
  }
}

// J*v wrapper that honours the convergence flag without a host round trip
__global__ void k_noop() {}

int lvpp_build_preconditioner(lvpp_problem* h, const lvpp_newton_opts* o) {
  if (o && o->pc_type != LVPP_PC_JACOBI) {
    lvpp_set_error("pc_type %d not available (only LVPP_PC_JACOBI)", o->pc_type);
    return LVPP_E_INVALID;
  }
  LAUNCH(h, k_build_pinv, lvpp_grid(h->Vown, 256, 8), 256, 0, h->Vown, h->slice_ptr, h->diag_k, h->K, h->M,
         h->D, h->bc_flag, h->alpha, (double2*)h->pinv);
  CK(cudaGetLastError());
  return 0;
}

int lvpp_minres(lvpp_problem* h, const double* d_rhs, double* d_y, const lvpp_newton_opts* o, int32_t* its,
                int32_t* reason, double* rnorm) {
  const int64_t Vown = h->Vown;
  const int nb = h->npartials;
  double2 *v0 = (double2*)h->va, *v1 = (double2*)h->vb, *z1 = (double2*)h->za, *z2 = (double2*)h->zb,
          *w0 = (double2*)h->wa, *w1 = (double2*)h->wb;
  double2* y = (double2*)d_y;
  const int maxit = o->ksp_max_it > 0 ? o->ksp_max_it : 10000;
  CK(cudaEventRecord(h->ev0, h->stream));
  LAUNCH(h, k_minres_init, nb, 256, 0, Vown, (const double2*)d_rhs, (const double2*)h->pinv, v0, v1, z1, w0,
         w1, y, h->partials);
  CK(cudaGetLastError());
  CKR(lvpp_reduce_partials(h, 1, h->scal->red));
  LAUNCH(h, k_minres_scal_init, 1, 1, 0, h->scal, o->ksp_rtol, o->ksp_atol, maxit);
  CK(cudaGetLastError());
  const int check_every = 32;
  int launched = 0;
  bool done = false;
  while (!done) {
    for (int k = 0; k < check_every; ++k) {
      // 1. Az = J (z1 / gamma1), partial delta
      if (h->nranks > 1) CKR(lvpp_halo_forward_impl(h, (double*)z1));
      if (k == 0) CK(cudaEventRecord(h->evs0, h->stream));  // sample one J*v launch per poll
      CKR(lvpp_apply_jacobian(h, (const double*)z1, h->Az, &h->scal->inv_gamma1, h->partials, &h->scal->conv));
      if (k == 0) CK(cudaEventRecord(h->evs1, h->stream));
      CKR(lvpp_reduce_partials(h, 1, h->scal->red));
      LAUNCH(h, k_minres_scal1, 1, 1, 0, h->scal);
      // 2. Lanczos vector + preconditioner, partial gamma2^2
      LAUNCH(h, k_minres_lanczos, nb, 256, 0, Vown, h->scal, (const double2*)h->Az, (const double2*)v1, v0,
             (const double2*)h->pinv, z2, h->partials);
      CKR(lvpp_reduce_partials(h, 1, h->scal->red));
      LAUNCH(h, k_minres_scal2, 1, 1, 0, h->scal);
      // 3. direction and solution update
      LAUNCH(h, k_minres_update, nb, 256, 0, Vown, h->scal, (const double2*)z1, w0, (const double2*)w1, y);
      CK(cudaGetLastError());
      // rotate: (v0, v1) <- (v1, v2 = old v0 buffer); z1 <- z2; (w0, w1) <- (w1, w2 = old w0 buffer)
      std::swap(v0, v1);
      std::swap(z1, z2);
      std::swap(w0, w1);
      ++launched;
    }
    CK(cudaMemcpyAsync(h->scal_host, h->scal, sizeof(KryScal), cudaMemcpyDeviceToHost, h->stream));
    CKR(lvpp_sync_check_comm(h));
    {
      float sms = 0.f;
      CK(cudaEventElapsedTime(&sms, h->evs0, h->evs1));
      h->spmv_sampled_ms += sms;
      h->spmv_samples++;
    }
    done = h->scal_host->conv != 0 || launched >= maxit + check_every;
  }
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaEventSynchronize(h->ev1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->t_krylov_ms += ms;
  h->krylov_its += h->scal_host->its;
  if (its) *its = h->scal_host->its;
  if (reason) *reason = h->scal_host->reason;
  if (rnorm) *rnorm = fabs(h->scal_host->eta);
  return 0;
}

int lvpp_solve_linear(lvpp_problem* h, const double* d_rhs, double* d_y, const lvpp_newton_opts* o, int32_t* its,
                      int32_t* reason, double* rnorm) {
  if (o->pc_type == LVPP_PC_MG) {
    if (o->pc_degree > 0 && (h->mg_npre != o->pc_degree || h->mg_npost != o->pc_degree)) {
      h->mg_nsmooth = h->mg_npre = h->mg_npost = o->pc_degree;
      h->mg_alpha_est = -1.0;  // the sweep dampings depend on the degree
    }
    CKR(lvpp_mg_update(h));
    int32_t its1 = 0, reason1 = 0;
    CKR(lvpp_gmres_mg(h, d_rhs, d_y, o, &its1, &reason1, rnorm));
    // The Chebyshev dampings pay in the first proximal step (26 - 30 Krylov iterations per Newton step at n = 215
    // against 32 - 35 with plain damping) and lose later: once the contact set has developed (exp(psi) -> 0 on it)
    // the solves of the second proximal step take 49 - 60 iterations with ratio 10 against 38 - 39 with plain damping,
    // and on the hardest systems of a CPU emulation of the whole solve (tools/full_solve_cpu.py, 40^3 mesh, second
    // Newton step at alpha = 5.3) plain damping needs 72 iterations where ratios 10 / 6 / 4 need 104 / 209 / > 300 --
    // no ratio is uniformly safe there.  So: start with the default ratio and fall back to plain damping, for the
    // rest of the handle's life, the first time a solve needs 1.5 times the best count seen.  Deterministic;
    // identical on every rank.
    if (h->mg_cheb_adapt && h->mg_cheb > 1.0 && reason1 > 0 && its1 >= 8) {  // (a solve that converges at once says nothing)
      if (h->mg_best_its == 0 || its1 < h->mg_best_its) h->mg_best_its = its1;
      else if (2 * its1 > 3 * h->mg_best_its && its1 > h->mg_best_its + 8) {
        h->mg_cheb = 0.0;
        if (getenv("LVPP_MG_VERBOSE") && h->rank == 0)
          fprintf(stderr, "[lvpp mg] %d Krylov iterations (best %d): plain damping from here on\n", (int)its1, (int)h->mg_best_its);
      }
    }
    if (its) *its = its1;
    if (reason) *reason = reason1;
    return 0;
  }
  CKR(lvpp_build_preconditioner(h, o));
  return lvpp_minres(h, d_rhs, d_y, o, its, reason, rnorm);
}

extern "C" int lvpp_linear_solve(lvpp_handle h, const double* d_rhs, double* d_y, const lvpp_newton_opts* opts,
                                 int32_t* its, int32_t* reason, double* h_rnorm) {
  if (!h || !d_rhs || !d_y || !opts) { lvpp_set_error("null argument"); return LVPP_E_INVALID; }
  CK(cudaSetDevice(h->device));
  if (!h->jac_valid) { lvpp_set_error("no Jacobian assembled yet"); return LVPP_E_INVALID; }
  CKR(lvpp_solve_linear(h, d_rhs, d_y, opts, its, reason, h_rnorm));
  return LVPP_OK;
}
