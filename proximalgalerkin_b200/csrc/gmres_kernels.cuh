// Vector kernels of the restarted GMRES (classical Gram-Schmidt with fused multi-dot products and
// deterministic two-stage reductions), shared by multigrid.cu (obstacle path, vectors of double2 per
// node) and forms.cu (generic mixed forms: plain fp64 vectors padded to an even length).
#pragma once
#include "lvpp_internal.cuh"

#define GM_CHUNK 8
#define GM_RING 8   // host-visible status slots / event pairs: the host polls iteration j - 1 while iteration j is queued

// Device-resident state of one restarted-GMRES solve (multigrid.cu: lvpp_gmres_mg).  The Hessenberg column, the Givens
// rotations, the residual estimate and the convergence decision live on the device, so an iteration is a fixed list of
// launches with no host round trip; the host reads a copy of this struct one iteration late (GM_RING pinned slots) and
// stops queueing when `conv` is set -- everything queued after that returns at once (gm_skip).
struct GmState {
  double rtol, atol, tol, bnorm, rnorm, hsq, inv_beta, eta2;
  int conv;         // non-zero: the solve is over (converged, broke down, diverged or out of iterations)
  int reason;       // KSP reason when conv
  int total;        // iterations of this solve
  int maxit;
  int need_reorth;  // the projection of pass 0 removed most of w: pass 1 runs
  int ncols;        // columns finished in this restart cycle
  int first;        // the next cycle start defines bnorm and tol
  int pad;
};

__device__ __forceinline__ bool gm_skip(const GmState* st, int pass) {
  return st && (st->conv || (pass == 1 && !st->need_reorth));
}

// ------------------------------------------------------------------------------------------------
// GMRES kernels
// partial sums of V_k . w for k = k0 .. k0+nv-1 (nv <= GM_CHUNK); partials[(k0 + k) * nparts + block].
// wy weighs the second component of every pair (1 = Euclidean; the obstacle path's equilibrated residual norm
// weighs the psi rows, multigrid.cu)
static __global__ void __launch_bounds__(256)
k_multi_dot(int64_t Vown, const double2* __restrict__ Vb, int64_t stride2, int k0, int nv,
            const double2* __restrict__ w, int nparts, double* __restrict__ partials, double wy,
            const GmState* __restrict__ st = nullptr, int pass = 0) {
  __shared__ double s_red[32];
  if (gm_skip(st, pass)) return;
  double acc[GM_CHUNK];
#pragma unroll
  for (int k = 0; k < GM_CHUNK; ++k) acc[k] = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < Vown; i += (int64_t)gridDim.x * blockDim.x) {
    double2 wi = w[i];
    wi.y *= wy;
#pragma unroll
    for (int k = 0; k < GM_CHUNK; ++k)
      if (k < nv) {
        const double2 v = Vb[(int64_t)(k0 + k) * stride2 + i];
        acc[k] += v.x * wi.x + v.y * wi.y;
      }
  }
#pragma unroll
  for (int k = 0; k < GM_CHUNK; ++k)
    if (k < nv) {
      const double r = lvpp_block_sum<256>(acc[k], s_red);
      if (threadIdx.x == 0) partials[(int64_t)(k0 + k) * nparts + blockIdx.x] = r;
    }
}

// w -= sum_k hc[k] V_k (k < nv); partial ||w_new||^2 into partials[slot * nparts + block]
static __global__ void __launch_bounds__(256)
k_gmres_update(int64_t Vown, const double2* __restrict__ Vb, int64_t stride2, int nv, const double* __restrict__ hc,
               double2* __restrict__ w, int nparts, int slot, double* __restrict__ partials, double wy,
               const GmState* __restrict__ st = nullptr, int pass = 0) {
  __shared__ double s_red[32];
  if (gm_skip(st, pass)) return;
  double part = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < Vown; i += (int64_t)gridDim.x * blockDim.x) {
    double2 wi = w[i];
    for (int k = 0; k < nv; ++k) {
      const double2 v = Vb[(int64_t)k * stride2 + i];
      const double c = hc[k];
      wi.x -= c * v.x;
      wi.y -= c * v.y;
    }
    w[i] = wi;
    part += wi.x * wi.x + wy * (wi.y * wi.y);
  }
  const double r = lvpp_block_sum<256>(part, s_red);
  if (threadIdx.x == 0) partials[(int64_t)slot * nparts + blockIdx.x] = r;
}

// out = sum_k yc[k] V_k
static __global__ void __launch_bounds__(256)
k_lincomb(int64_t Vown, const double2* __restrict__ Vb, int64_t stride2, int nv, const double* __restrict__ yc,
          double2* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < Vown; i += (int64_t)gridDim.x * blockDim.x) {
    double2 s = make_double2(0.0, 0.0);
    for (int k = 0; k < nv; ++k) {
      const double2 v = Vb[(int64_t)k * stride2 + i];
      const double c = yc[k];
      s.x += c * v.x;
      s.y += c * v.y;
    }
    out[i] = s;
  }
}
// y = a * x (+ y if accumulate)
static __global__ void k_axpby(int64_t Vown, double a, const double2* __restrict__ x, int accumulate, double2* __restrict__ y) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < Vown; i += (int64_t)gridDim.x * blockDim.x) {
    const double2 v = x[i];
    double2 o = accumulate ? y[i] : make_double2(0.0, 0.0);
    o.x += a * v.x;
    o.y += a * v.y;
    y[i] = o;
  }
}
static __global__ void __launch_bounds__(256) k_reduce_multi(int nparts, int nvals, const double* __restrict__ partials,
                                                       double* __restrict__ out, const GmState* __restrict__ st = nullptr,
                                                       int pass = 0) {
  __shared__ double s_red[32];
  if (gm_skip(st, pass)) return;
  for (int v = blockIdx.x; v < nvals; v += gridDim.x) {
    double s = 0.0;
    for (int i = threadIdx.x; i < nparts; i += 256) s += partials[(int64_t)v * nparts + i];
    const double r = lvpp_block_sum<256>(s, s_red);
    if (threadIdx.x == 0) out[v] = r;
  }
}


// ------------------------------------------------------------------------------------------------
// Device-side control of the GMRES recurrence (one thread each; the data is a Hessenberg column).
// Start of a restart cycle: red[0] = ||r||^2 (all-reduced).  Defines tol on the first cycle, sets g = ||r|| e_0.
static __global__ void k_gm_cycle_begin(GmState* s, const double* __restrict__ red, double* __restrict__ g, int m) {
  if (s->conv) return;
  const double rn = sqrt(red[0]);
  s->rnorm = rn;
  if (s->first) {
    s->bnorm = rn;
    s->tol = fmax(s->rtol * rn, s->atol);
    s->first = 0;
  }
  for (int i = 0; i <= m; ++i) g[i] = 0.0;
  g[0] = rn;
  s->ncols = 0;
  s->need_reorth = 0;
  if (!isfinite(rn)) { s->conv = 1; s->reason = LVPP_KSP_DIVERGED_NANORINF; }
  else if (rn <= s->tol) { s->conv = 1; s->reason = rn <= s->atol ? LVPP_KSP_CONVERGED_ATOL : LVPP_KSP_CONVERGED_RTOL; }
  s->inv_beta = rn > 0.0 ? 1.0 / rn : 0.0;
}

// After a Gram-Schmidt pass of column j: hc[0..j] = V_k . w (all-reduced), nrm2[0] = ||w - sum hc_k V_k||^2.
// Pass 0 decides whether the projection has to be repeated (Daniel et al.: less than eta2 of ||w||^2 survived); the
// pass that ends the column applies the stored rotations, makes the new one and tests the residual estimate.
static __global__ void k_gm_after_pass(GmState* s, int j, int pass, const double* __restrict__ hc, const double* __restrict__ nrm2,
                                       double* __restrict__ Hcol, double* __restrict__ cs, double* __restrict__ sn,
                                       double* __restrict__ g) {
  if (gm_skip(s, pass)) return;
  double hsq = 0.0;
  for (int k = 0; k <= j; ++k) {
    const double c = hc[k];
    Hcol[k] = (pass ? Hcol[k] : 0.0) + c;
    hsq += c * c;
  }
  const double beta2 = nrm2[0];
  if (pass == 0 && !(beta2 > s->eta2 * (hsq + beta2)) && isfinite(beta2)) {
    s->need_reorth = 1;
    return;
  }
  s->need_reorth = 0;
  const double beta = sqrt(beta2);
  Hcol[j + 1] = beta;
  s->total += 1;
  s->ncols = j + 1;
  for (int k = 0; k < j; ++k) {
    const double t = cs[k] * Hcol[k] + sn[k] * Hcol[k + 1];
    Hcol[k + 1] = -sn[k] * Hcol[k] + cs[k] * Hcol[k + 1];
    Hcol[k] = t;
  }
  const double den = hypot(Hcol[j], Hcol[j + 1]);
  cs[j] = den > 0.0 ? Hcol[j] / den : 1.0;
  sn[j] = den > 0.0 ? Hcol[j + 1] / den : 0.0;
  Hcol[j] = den;
  Hcol[j + 1] = 0.0;
  g[j + 1] = -sn[j] * g[j];
  g[j] = cs[j] * g[j];
  const double rn = fabs(g[j + 1]);
  s->rnorm = rn;
  s->inv_beta = beta > 0.0 ? 1.0 / beta : 0.0;
  if (!isfinite(rn)) { s->conv = 1; s->reason = LVPP_KSP_DIVERGED_NANORINF; }
  else if (rn <= s->tol) { s->conv = 1; s->reason = rn <= s->atol ? LVPP_KSP_CONVERGED_ATOL : LVPP_KSP_CONVERGED_RTOL; }
  else if (beta == 0.0) { s->conv = 1; s->reason = LVPP_KSP_CONVERGED_RTOL; }  // happy breakdown: exact solution in the space
  else if (s->total >= s->maxit) { s->conv = 1; s->reason = LVPP_KSP_DIVERGED_ITS; }
}

// v *= inv_beta (the new basis vector); nothing once the solve is over
static __global__ void k_gm_scale(int64_t Vown, const GmState* __restrict__ s, double2* __restrict__ v) {
  if (s->conv) return;
  const double a = s->inv_beta;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < Vown; i += (int64_t)gridDim.x * blockDim.x) {
    double2 t = v[i];
    t.x *= a;
    t.y *= a;
    v[i] = t;
  }
}

// back substitution H yv = g over the k finished columns of the cycle (H upper triangular after the rotations,
// column c at H + c * (m + 1)); coefficients beyond k are zeroed so that a combination over a fixed count is safe
static __global__ void k_gm_backsolve(int k, int m, const double* __restrict__ H, const double* __restrict__ g,
                                      double* __restrict__ yv) {
  for (int i = k; i < m; ++i) yv[i] = 0.0;
  for (int i = k - 1; i >= 0; --i) {
    double t = g[i];
    for (int c = i + 1; c < k; ++c) t -= H[(size_t)c * (m + 1) + i] * yv[c];
    yv[i] = t / H[(size_t)i * (m + 1) + i];
  }
}
