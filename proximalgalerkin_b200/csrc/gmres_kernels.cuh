// Vector kernels of the restarted GMRES (classical Gram-Schmidt with fused multi-dot products and
// deterministic two-stage reductions), shared by multigrid.cu (obstacle path, vectors of double2 per
// node) and forms.cu (generic mixed forms: plain fp64 vectors padded to an even length).
#pragma once
#include "lvpp_internal.cuh"

#define GM_CHUNK 8

// ------------------------------------------------------------------------------------------------
// GMRES kernels
// partial sums of V_k . w for k = k0 .. k0+nv-1 (nv <= GM_CHUNK); partials[(k0 + k) * nparts + block].
// wy weighs the second component of every pair (1 = Euclidean; the obstacle path's equilibrated residual norm
// weighs the psi rows, multigrid.cu)
static __global__ void __launch_bounds__(256)
k_multi_dot(int64_t Vown, const double2* __restrict__ Vb, int64_t stride2, int k0, int nv,
            const double2* __restrict__ w, int nparts, double* __restrict__ partials, double wy) {
  __shared__ double s_red[32];
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
               double2* __restrict__ w, int nparts, int slot, double* __restrict__ partials, double wy) {
  __shared__ double s_red[32];
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
                                                       double* __restrict__ out) {
  __shared__ double s_red[32];
  for (int v = blockIdx.x; v < nvals; v += gridDim.x) {
    double s = 0.0;
    for (int i = threadIdx.x; i < nparts; i += 256) s += partials[(int64_t)v * nparts + i];
    const double r = lvpp_block_sum<256>(s, s_red);
    if (threadIdx.x == 0) out[v] = r;
  }
}

