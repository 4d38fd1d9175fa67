// Monolithic aggregation multigrid for the saddle-point Newton system and the GMRES that it
// preconditions ("block-preconditioned MINRES/GMRES with a Jacobi ... smoother", BASELINE.json).
//
// J = [[alpha K, M], [M, -D(psi)]] is symmetric quasi-definite with one scalar node pattern and a 2x2
// block per entry.  Block-diagonal preconditioners (the ex40.cpp:261-274 recipe used by the MINRES
// path in krylov.cu) need O(1/h) iterations here because the diagonal Schur approximation is poor on
// the contact set.  Instead both fields are coarsened together:
//   * aggregates = boxes of 2^d neighbouring nodes found from the node coordinates (Dirichlet nodes
//     are aggregated separately, so a coarse node is either free or Dirichlet and every level has
//     exactly the fine level's structure: same kernels, same masks);
//   * prolongation = piecewise constant per field; Galerkin coarse operators K_c, M_c, D_c are sums of
//     fine entries through a precomputed, sorted (deterministic, atomic-free) slot map; only D_c is
//     recomputed per Newton step;
//   * smoother = damped node-block Jacobi: the 2x2 node blocks [[alpha K_ii, M_ii], [M_ii, -D_ii]]
//     are inverted exactly, which locally eliminates psi (eigenvalues of Binv J are real, in (0, 2]);
//   * coarsest level: explicit dense inverse (Gauss-Jordan with partial pivoting, one CTA).
// The V-cycle is not SPD, so the Krylov method is right-preconditioned restarted GMRES (classical
// Gram-Schmidt with selective re-orthogonalisation; Hessenberg/Givens on the host, two small D2H
// reads per iteration; all reductions deterministic).
#include <cub/cub.cuh>
#include <cuda_bf16.h>

#include <cmath>
#include <cstring>

#include "block_op.cuh"
#include "gmres_kernels.cuh"
#include "lvpp_internal.cuh"

#define MG_MAX_LEVELS 14
#define MG_COARSE_TARGET 96    // stop coarsening at <= this many coarse nodes
#define MG_COARSE_MAX 384      // dense inverse limit (unknowns = 2 * nodes)

// ------------------------------------------------------------------------------------------------
// setup kernels
__global__ void k_box0(int64_t Vown, int tdim, const double* __restrict__ coords, double x0, double y0,
                       double z0, double inv_h, int32_t* __restrict__ box) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < Vown; i += (int64_t)gridDim.x * blockDim.x) {
    const double o[3] = {x0, y0, z0};
    for (int d = 0; d < 3; ++d) {
      int32_t b = 0;
      if (d < tdim) {
        const double q = (coords[i * tdim + d] - o[d]) * inv_h + 0.25;
        b = q < 0 ? 0 : (q > 1048574.0 ? 1048574 : (int32_t)q);
      }
      box[i * 3 + d] = b;
    }
  }
}

__global__ void k_agg_keys(int64_t Vown, const int32_t* __restrict__ box, const uint8_t* __restrict__ bc,
                           uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < Vown; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t bx = (uint64_t)(box[i * 3 + 0] >> 1), by = (uint64_t)(box[i * 3 + 1] >> 1),
                   bz = (uint64_t)(box[i * 3 + 2] >> 1);
    keys[i] = ((uint64_t)(bc[i] ? 1 : 0) << 62) | (bz << 40) | (by << 20) | bx;
    vals[i] = (uint32_t)i;
  }
}

__global__ void k_heads(int64_t n, const uint64_t* __restrict__ skeys, int64_t* __restrict__ head) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x)
    head[p] = (p == 0 || skeys[p] != skeys[p - 1]) ? 1 : 0;
}

// after the inclusive scan of head: id[p] = scan[p] - 1
__global__ void k_agg_fill(int64_t Vown, const uint64_t* __restrict__ skeys, const uint32_t* __restrict__ sidx,
                           const int64_t* __restrict__ scan, int32_t* __restrict__ agg,
                           int32_t* __restrict__ members, int64_t* __restrict__ agg_ptr,
                           int32_t* __restrict__ cbox, uint8_t* __restrict__ cbc) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < Vown; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t id = scan[p] - 1;
    agg[sidx[p]] = (int32_t)id;
    members[p] = (int32_t)sidx[p];
    if (p == 0 || skeys[p] != skeys[p - 1]) {
      agg_ptr[id] = p;
      const uint64_t k = skeys[p];
      cbox[id * 3 + 0] = (int32_t)(k & 0xfffff);
      cbox[id * 3 + 1] = (int32_t)((k >> 20) & 0xfffff);
      cbox[id * 3 + 2] = (int32_t)((k >> 40) & 0xfffff);
      cbc[id] = (uint8_t)((k >> 62) & 1);
    }
    if (p == Vown - 1) agg_ptr[id + 1] = Vown;
  }
}

// key of every stored slot of the fine level: (coarse row << 32) | coarse column
__global__ void k_gal_keys(int64_t Vown, const int64_t* __restrict__ slice_ptr, const uint32_t* __restrict__ col,
                           const int32_t* __restrict__ agg, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < Vown; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b0 = slice_ptr[i >> 5];
    const int w = (int)((slice_ptr[(i >> 5) + 1] - b0) >> 5);
    const uint64_t I = (uint64_t)agg[i];
    for (int k = 0; k < w; ++k) {
      const int64_t s = b0 + (i & 31) + (int64_t)k * LVPP_SLICE;
      const uint32_t j = col[s] & ~LVPP_COL_BC;
      keys[s] = (I << 32) | (uint64_t)(uint32_t)agg[j];
      vals[s] = (uint32_t)s;
    }
  }
}

__global__ void k_compact_heads(int64_t n, const uint64_t* __restrict__ skeys, const int64_t* __restrict__ scan,
                                uint64_t* __restrict__ ukeys, int64_t* __restrict__ gal_ptr, int64_t nvalid) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nvalid; p += (int64_t)gridDim.x * blockDim.x)
    if (p == 0 || skeys[p] != skeys[p - 1]) {
      ukeys[scan[p] - 1] = skeys[p];
      gal_ptr[scan[p] - 1] = p;
    }
}

// first unique entry of every coarse row (lower bound of row << 32)
__global__ void k_coarse_rowstart(int64_t Vc, int64_t nnzc, const uint64_t* __restrict__ ukeys,
                                  int64_t* __restrict__ rowstart, int32_t* __restrict__ rowlen) {
  for (int64_t I = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; I <= Vc; I += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t target = (uint64_t)I << 32;
    int64_t lo = 0, hi = nnzc;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (ukeys[mid] < target) lo = mid + 1; else hi = mid;
    }
    rowstart[I] = lo;
  }
}
__global__ void k_coarse_rowlen(int64_t Vc, const int64_t* __restrict__ rowstart, int32_t* __restrict__ rowlen,
                                int* __restrict__ err) {
  for (int64_t I = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; I < Vc; I += (int64_t)gridDim.x * blockDim.x) {
    const int64_t len = rowstart[I + 1] - rowstart[I];
    if (len > LVPP_MAX_ROW) atomicExch(err, 1);
    rowlen[I] = (int32_t)len;
  }
}
__global__ void k_slice_slots(int64_t nslices, int64_t Vc, const int32_t* __restrict__ rowlen, int64_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  for (int64_t s = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; s <= nslices;
       s += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    const int64_t i = s * LVPP_SLICE + lane;
    int w = (s < nslices && i < Vc) ? rowlen[i] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
    if (lane == 0) out[s] = (int64_t)w * LVPP_SLICE;
  }
}
__global__ void k_coarse_cols(int64_t Vc, const int64_t* __restrict__ rowstart, const uint64_t* __restrict__ ukeys,
                              const int64_t* __restrict__ slice_ptr, const uint8_t* __restrict__ cbc,
                              uint32_t* __restrict__ col, uint8_t* __restrict__ diag_k, int64_t* __restrict__ gal_dst) {
  for (int64_t I = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; I < Vc; I += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b0 = slice_ptr[I >> 5];
    const int w = (int)((slice_ptr[(I >> 5) + 1] - b0) >> 5);
    const int64_t u0 = rowstart[I];
    const int len = (int)(rowstart[I + 1] - u0);
    for (int k = 0; k < w; ++k) {
      const int64_t s = b0 + (I & 31) + (int64_t)k * LVPP_SLICE;
      if (k < len) {
        const uint32_t J = (uint32_t)(ukeys[u0 + k] & 0xffffffffu);
        col[s] = J | (cbc[J] ? LVPP_COL_BC : 0u);
        gal_dst[u0 + k] = s;
        if ((int64_t)J == I) diag_k[I] = (uint8_t)k;
      } else {
        col[s] = (uint32_t)I | (cbc[I] ? LVPP_COL_BC : 0u);
      }
    }
  }
}

// Galerkin sums through the sorted slot map: Xc[dst[u]] = sum_{p in seg(u)} X[src[p]] (fixed order)
__global__ void k_galerkin(int64_t nnzc, const int64_t* __restrict__ gal_ptr, const uint32_t* __restrict__ gal_src,
                           const int64_t* __restrict__ gal_dst, int narr, const double* __restrict__ X0,
                           const double* __restrict__ X1, double* __restrict__ Y0, double* __restrict__ Y1, double scale0) {
  for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < nnzc; u += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p1 = gal_ptr[u + 1];
    double a0 = 0.0, a1 = 0.0;
    for (int64_t p = gal_ptr[u]; p < p1; ++p) {
      const uint32_t s = gal_src[p];
      a0 += X0[s];
      if (narr > 1) a1 += X1[s];
    }
    const int64_t d = gal_dst[u];
    Y0[d] = a0 * scale0;
    if (narr > 1) Y1[d] = a1;
  }
}

// ------------------------------------------------------------------------------------------------
// cycle kernels
__global__ void k_build_binv(int64_t Vown, const int64_t* __restrict__ slice_ptr, const uint8_t* __restrict__ diag_k,
                             const double* __restrict__ K, const double* __restrict__ M, const double* __restrict__ D,
                             const uint8_t* __restrict__ bc, double alpha, double* __restrict__ binv,
                             float4* __restrict__ binv32) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < Vown; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t idx = slice_ptr[i >> 5] + (i & 31) + (int64_t)diag_k[i] * LVPP_SLICE;
    const double a = alpha * K[idx], m = M[idx], d = D[idx];
    double* B = binv + 4 * i;
    if (bc[i]) {  // identity row for u (applied undamped by the sweeps), psi decoupled from the masked column
      B[0] = 1.0; B[1] = 0.0; B[2] = 0.0; B[3] = -1.0 / d;
    } else {
      const double idet = 1.0 / (-a * d - m * m);
      B[0] = -d * idet; B[1] = -m * idet; B[2] = -m * idet; B[3] = a * idet;
    }
    if (binv32) {  // single-precision copy read by k_packed2_op (clamped: 1 / D leaves the range at psi < -70)
      const double big = 3.0e38;
      binv32[i] = make_float4((float)fmin(fmax(B[0], -big), big), (float)fmin(fmax(B[1], -big), big),
                              (float)fmin(fmax(B[2], -big), big), (float)fmin(fmax(B[3], -big), big));
    }
  }
}

__global__ void k_smooth_first(int64_t Vown, const double2* __restrict__ b, const double* __restrict__ binv,
                               const uint8_t* __restrict__ bc, double omega, double2* __restrict__ x) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < Vown; i += (int64_t)gridDim.x * blockDim.x) {
    const double2 r = b[i];
    const double* B = binv + 4 * i;
    const double wu = bc[i] ? 1.0 : omega;  // Dirichlet rows are solved exactly
    x[i] = make_double2(wu * (B[0] * r.x + B[1] * r.y), omega * (B[2] * r.x + B[3] * r.y));
  }
}

__global__ void k_restrict(int64_t Vc, const int64_t* __restrict__ agg_ptr, const int32_t* __restrict__ members,
                           const double2* __restrict__ r, const uint8_t* __restrict__ cbc, double2* __restrict__ bc_out) {
  for (int64_t I = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; I < Vc; I += (int64_t)gridDim.x * blockDim.x) {
    double su = 0.0, sp = 0.0;
    const int64_t p1 = agg_ptr[I + 1];
    for (int64_t p = agg_ptr[I]; p < p1; ++p) {
      const double2 v = r[members[p]];
      su += v.x;
      sp += v.y;
    }
    bc_out[I] = make_double2(cbc[I] ? 0.0 : su, sp);
  }
}

__global__ void k_prolong_add(int64_t Vown, const int32_t* __restrict__ agg, const double2* __restrict__ xc,
                              const uint8_t* __restrict__ bc, double over, double2* __restrict__ x) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < Vown; i += (int64_t)gridDim.x * blockDim.x) {
    const double2 c = xc[agg[i]];
    double2 v = x[i];
    if (!bc[i]) v.x += over * c.x;
    v.y += over * c.y;
    x[i] = v;
  }
}

// dense coarsest operator [n x 2n] = [J_c | I], n = 2 * Vc, Dirichlet masks applied
__global__ void k_coarse_dense(int64_t Vc, int64_t n, const int32_t* __restrict__ gmap,
                               const int64_t* __restrict__ slice_ptr, const int32_t* __restrict__ rowlen,
                               const uint32_t* __restrict__ col, const double* __restrict__ K,
                               const double* __restrict__ M, const double* __restrict__ D,
                               const uint8_t* __restrict__ bc, double alpha, double* __restrict__ A) {
  // n = global unknowns; rows of the owned nodes only (the other ranks' rows stay zero and are summed in)
  const int64_t ld = 2 * n;
  for (int64_t il = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; il < Vc; il += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = il, gi = gmap[il];
    const int64_t base = slice_ptr[i >> 5] + (i & 31);
    const bool rbc = bc[i] != 0;
    const int len = rowlen[i];
    for (int k = 0; k < len; ++k) {
      const int64_t idx = base + (int64_t)k * LVPP_SLICE;
      const uint32_t c = col[idx];
      const bool cbc = (c & LVPP_COL_BC) != 0;
      const int64_t j = c & ~LVPP_COL_BC;
      double kuu = alpha * K[idx], kup = M[idx], kpu = M[idx], kpp = -D[idx];
      if (rbc) { kuu = (j == i) ? 1.0 : 0.0; kup = 0.0; }
      if (cbc) { if (!rbc) kuu = 0.0; kpu = 0.0; }
      const int64_t gj = gmap[j];
      A[(2 * gi) * ld + 2 * gj] = kuu;
      A[(2 * gi) * ld + 2 * gj + 1] = kup;
      A[(2 * gi + 1) * ld + 2 * gj] = kpu;
      A[(2 * gi + 1) * ld + 2 * gj + 1] = kpp;
    }
    A[(2 * gi) * ld + n + 2 * gi] = 1.0;
    A[(2 * gi + 1) * ld + n + 2 * gi + 1] = 1.0;
  }
}

// Gauss-Jordan with partial pivoting on [A | I] in one CTA; writes inv column-major: inv[c * n + r]
__global__ void __launch_bounds__(1024) k_gauss_jordan(int n, double* __restrict__ A, double* __restrict__ inv,
                                                        int* __restrict__ err) {
  __shared__ double s_val[32];
  __shared__ int s_idx[32];
  __shared__ int s_piv;
  const int ld = 2 * n, tid = threadIdx.x, nt = blockDim.x;
  for (int k = 0; k < n; ++k) {
    // pivot search in column k, rows k..n-1
    double best = -1.0;
    int bi = k;
    for (int r = k + tid; r < n; r += nt) {
      const double v = fabs(A[(int64_t)r * ld + k]);
      if (v > best) { best = v; bi = r; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_down_sync(0xffffffffu, best, o);
      const int oi = __shfl_down_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((tid & 31) == 0) { s_val[tid >> 5] = best; s_idx[tid >> 5] = bi; }
    __syncthreads();
    if (tid < 32) {
      best = tid < (nt >> 5) ? s_val[tid] : -1.0;
      bi = tid < (nt >> 5) ? s_idx[tid] : n;
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, best, o);
        const int oi = __shfl_down_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      if (tid == 0) {
        s_piv = bi;
        if (!(best > 0.0)) *err = 1;
      }
    }
    __syncthreads();
    const int p = s_piv;
    if (p != k)
      for (int c = k + tid; c < ld; c += nt) {
        const double t = A[(int64_t)k * ld + c];
        A[(int64_t)k * ld + c] = A[(int64_t)p * ld + c];
        A[(int64_t)p * ld + c] = t;
      }
    __syncthreads();
    const double ipiv = 1.0 / A[(int64_t)k * ld + k];
    __syncthreads();
    for (int c = k + tid; c < ld; c += nt) A[(int64_t)k * ld + c] *= ipiv;
    __syncthreads();
    // eliminate column k from all other rows (row swaps move the identity part: all columns > k)
    const int c0 = k + 1, c1 = ld, ncol = c1 - c0;
    for (int64_t e = tid; e < (int64_t)n * ncol; e += nt) {
      const int r = (int)(e / ncol), c = c0 + (int)(e % ncol);
      if (r == k) continue;
      const double f = A[(int64_t)r * ld + k];
      if (f != 0.0) A[(int64_t)r * ld + c] -= f * A[(int64_t)k * ld + c];
    }
    __syncthreads();
    for (int r = tid; r < n; r += nt)
      if (r != k) A[(int64_t)r * ld + k] = 0.0;
    __syncthreads();
  }
  for (int64_t e = tid; e < (int64_t)n * n; e += nt) {
    const int r = (int)(e / n), c = (int)(e % n);
    inv[(int64_t)c * n + r] = A[(int64_t)r * ld + n + c];
  }
}

// x[r] = (inv b)[row0 + r] for the nloc owned unknowns of this rank; b holds all ranks' unknowns
__global__ void k_coarse_apply(int n, int nloc, int row0, const double* __restrict__ inv, const double* __restrict__ b,
                               double* __restrict__ x) {
  extern __shared__ double s_b[];
  for (int c = threadIdx.x; c < n; c += blockDim.x) s_b[c] = b[c];
  __syncthreads();
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nloc; r += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int c = 0; c < n; ++c) s += inv[(int64_t)c * n + row0 + r] * s_b[c];
    x[r] = s;
  }
}

// packed single-precision copy of a level's operator for the cycle (block_op.cuh: k_packed_op)
__global__ void k_pack_op(int64_t n, const uint32_t* __restrict__ col, const double* __restrict__ K,
                          const double* __restrict__ M, const double* __restrict__ D, double alpha,
                          uint4* __restrict__ P) {
  // D = int exp(psi) phi_i phi_j can exceed the single-precision range while Newton overshoots (psi > 88 + ln(1/M_ii)):
  // clamp instead of producing inf (inf * 0 = NaN would poison the whole cycle)
  const double big = 3.0e38;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    P[i] = make_uint4(col[i], __float_as_uint((float)fmin(fmax(alpha * K[i], -big), big)), __float_as_uint((float)M[i]),
                      __float_as_uint((float)fmin(D[i], big)));
}

// bf16 pair records of the experimental cycle operator (block_op.cuh: k_packed2_op); one warp per slice
__device__ __forceinline__ uint32_t lvpp_bf16_pair(float hi, float lo) {
  return ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(hi)) << 16) | (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(lo));
}
__global__ void __launch_bounds__(256) k_pack_op_bf16(int64_t nslices, const int64_t* __restrict__ slice_ptr,
                                                      const uint32_t* __restrict__ col, const double* __restrict__ K,
                                                      const double* __restrict__ M, const double* __restrict__ D, double alpha,
                                                      uint4* __restrict__ P2, uint32_t* __restrict__ Pd) {
  const double big = 3.0e38;  // as k_pack_op: no inf in the records (bf16 has the exponent range of fp32)
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t s = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; s < nslices; s += nwarps) {
    const int64_t b0 = slice_ptr[s];
    const int w = (int)((slice_ptr[s + 1] - b0) >> 5);
    const int64_t po = ((b0 + 32 * s) >> 1) + lane;
    for (int k = 0; k < w; k += 2) {
      const int64_t i0 = b0 + lane + (int64_t)k * LVPP_SLICE;
      const bool two = k + 1 < w;
      const int64_t i1 = two ? i0 + LVPP_SLICE : i0;  // rows of odd width: a zero second slot on the first one's column
      const float k0 = (float)fmin(fmax(alpha * K[i0], -big), big), m0 = (float)M[i0], d0 = (float)fmin(D[i0], big);
      const float k1 = two ? (float)fmin(fmax(alpha * K[i1], -big), big) : 0.f, m1 = two ? (float)M[i1] : 0.f,
                  d1 = two ? (float)fmin(D[i1], big) : 0.f;
      P2[po + (int64_t)(k >> 1) * LVPP_SLICE] = make_uint4(col[i0], col[i1], lvpp_bf16_pair(k0, m0), lvpp_bf16_pair(k1, m1));
      Pd[po + (int64_t)(k >> 1) * LVPP_SLICE] = lvpp_bf16_pair(d0, d1);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// hierarchy construction
static int reduce_to_host(lvpp_problem* h, double* partials, int nvals, double* dst_dev, double* dst_host);

// sum over ranks of a few host doubles (setup only)
static int host_allreduce_sum(lvpp_problem* h, double* vals, int n) {
  if (h->nranks <= 1) return 0;
  if (n > 64) { lvpp_set_error("host_allreduce_sum: too many values"); return LVPP_E_INVALID; }
  CK(cudaMemcpyAsync(h->gm_h, vals, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
  CKR(lvpp_allreduce_sum(h, h->gm_h, n));
  CK(cudaMemcpyAsync(vals, h->gm_h, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

__global__ void k_pack_agg(int64_t n, const int32_t* __restrict__ nodes, const int32_t* __restrict__ agg,
                           const uint8_t* __restrict__ bc, int32_t* __restrict__ out) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const int32_t i = nodes[p];
    out[p] = (int32_t)((uint32_t)agg[i] | (bc[i] ? 0x80000000u : 0u));
  }
}
__global__ void k_scatter_i32(int64_t n, const int32_t* __restrict__ nodes, const int32_t* __restrict__ vals,
                              int32_t* __restrict__ dst) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x)
    dst[nodes[p]] = vals[p];
}
__global__ void k_ev_init(int64_t Vown, double2* __restrict__ v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < Vown; i += (int64_t)gridDim.x * blockDim.x) {
    // fixed pseudo-random start vector (integer hash of the node number)
    uint32_t a = (uint32_t)i * 2654435761u + 12345u;
    a ^= a >> 15; a *= 2246822519u; a ^= a >> 13;
    uint32_t b = a * 3266489917u + 1u;
    b ^= b >> 16;
    v[i] = make_double2((double)(a & 0xffff) / 65536.0 - 0.5, (double)(b & 0xffff) / 65536.0 - 0.5);
  }
}

static int level_alloc_vectors(lvpp_problem* h, MgLevel& L) {
  CKR(lvpp_dalloc(h, &L.binv, (size_t)4 * L.Vown));
  CKR(lvpp_dalloc(h, &L.b, (size_t)2 * L.V));
  CKR(lvpp_dalloc(h, &L.x, (size_t)2 * L.V));
  CKR(lvpp_dalloc(h, &L.t, (size_t)2 * L.V));
  CKR(lvpp_dalloc(h, &L.ev, (size_t)2 * L.V));
  return 0;
}

// Coarse halo of level l + 1 from the aggregates of level l: the owner sends the aggregate number (and
// Dirichlet bit) of every node on its send lists; both sides take the sorted unique aggregate numbers
// per neighbour, so the coarse send list of the owner and the coarse recv list of the receiver agree
// entry by entry.  Coarse ghosts are numbered after the owned coarse nodes, neighbour by neighbour.
static int build_coarse_halo(lvpp_problem* h, LevelHalo& FH, MgLevel& F, MgLevel& C, int64_t Vc) {
  LevelHalo& CH = C.halo;
  C.Vown = Vc;
  C.V = Vc;
  if (h->nranks <= 1) return 0;
  if (FH.num_neighbors == 0) return lvpp_halo_p2p_setup(h, CH);  // collective: take part without neighbours
  const int64_t ns = FH.send_ptr.back(), nr = FH.recv_ptr.back();
  int32_t *d_s = nullptr, *d_r = nullptr;
  CKR(lvpp_dalloc(h, &d_s, (size_t)ns, false));
  CKR(lvpp_dalloc(h, &d_r, (size_t)nr, false));
  if (ns > 0) {
    LAUNCH(h, k_pack_agg, lvpp_grid(ns, 256, 4), 256, 0, ns, FH.send_nodes, F.agg, F.bc_flag, d_s);
    CK(cudaGetLastError());
  }
  CKR(lvpp_halo_exchange_i32(h, FH, d_s, d_r));
  std::vector<uint32_t> hs((size_t)ns), hr((size_t)nr);
  CK(cudaMemcpyAsync(hs.data(), d_s, sizeof(int32_t) * ns, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(hr.data(), d_r, sizeof(int32_t) * nr, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CH.num_neighbors = FH.num_neighbors;
  CH.neighbor_ranks = FH.neighbor_ranks;
  CH.send_ptr.assign(1, 0);
  CH.recv_ptr.assign(1, 0);
  std::vector<int32_t> csend, crecv, aggghost((size_t)nr);
  std::vector<uint8_t> cbcg;
  int64_t ncg = 0;
  for (int b = 0; b < FH.num_neighbors; ++b) {
    std::vector<uint32_t> u;
    for (int64_t p = FH.send_ptr[b]; p < FH.send_ptr[b + 1]; ++p) u.push_back(hs[p] & 0x7fffffffu);
    std::sort(u.begin(), u.end());
    u.erase(std::unique(u.begin(), u.end()), u.end());
    for (uint32_t v : u) csend.push_back((int32_t)v);
    CH.send_ptr.push_back((int64_t)csend.size());
    std::vector<uint32_t> w;
    for (int64_t p = FH.recv_ptr[b]; p < FH.recv_ptr[b + 1]; ++p) w.push_back(hr[p] & 0x7fffffffu);
    std::sort(w.begin(), w.end());
    w.erase(std::unique(w.begin(), w.end()), w.end());
    cbcg.resize((size_t)(ncg + (int64_t)w.size()), 0);
    for (int64_t p = FH.recv_ptr[b]; p < FH.recv_ptr[b + 1]; ++p) {
      const int64_t id = std::lower_bound(w.begin(), w.end(), hr[p] & 0x7fffffffu) - w.begin();
      aggghost[p] = (int32_t)(Vc + ncg + id);
      cbcg[ncg + id] = (uint8_t)(hr[p] >> 31);
    }
    for (size_t id = 0; id < w.size(); ++id) {
      crecv.push_back((int32_t)(Vc + ncg + (int64_t)id));
      CH.ghost_owner_local.push_back((int32_t)w[id]);
      CH.ghost_nbr.push_back(b);
    }
    ncg += (int64_t)w.size();
    CH.recv_ptr.push_back(ncg);
  }
  CH.recv_nodes_host = crecv;
  C.V = Vc + ncg;
  CKR(lvpp_dalloc(h, &CH.send_nodes, csend.size(), false));
  CKR(lvpp_dalloc(h, &CH.recv_nodes, crecv.size(), false));
  CKR(lvpp_dalloc(h, &CH.send_buf, 2 * csend.size()));
  CKR(lvpp_dalloc(h, &CH.recv_buf, 2 * crecv.size()));
  int32_t* d_ag = nullptr;
  CKR(lvpp_dalloc(h, &d_ag, (size_t)nr, false));
  if (!csend.empty()) CK(cudaMemcpyAsync(CH.send_nodes, csend.data(), sizeof(int32_t) * csend.size(), cudaMemcpyHostToDevice, h->stream));
  if (!crecv.empty()) CK(cudaMemcpyAsync(CH.recv_nodes, crecv.data(), sizeof(int32_t) * crecv.size(), cudaMemcpyHostToDevice, h->stream));
  if (nr > 0) {
    CK(cudaMemcpyAsync(d_ag, aggghost.data(), sizeof(int32_t) * nr, cudaMemcpyHostToDevice, h->stream));
    LAUNCH(h, k_scatter_i32, lvpp_grid(nr, 256, 4), 256, 0, nr, FH.recv_nodes, d_ag, F.agg);
    CK(cudaGetLastError());
  }
  if (ncg > 0) CK(cudaMemcpyAsync(C.bc_flag + Vc, cbcg.data(), (size_t)ncg, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CKR(lvpp_dfree(h, d_s));
  CKR(lvpp_dfree(h, d_r));
  CKR(lvpp_dfree(h, d_ag));
  CKR(lvpp_halo_p2p_setup(h, CH));
  return 0;
}

// builds level l + 1 from level l; every rank aggregates its owned nodes only
static int build_next_level(lvpp_problem* h, int l, bool* stop) {
  MgLevel& F = h->levels[l];
  *stop = false;
  const int64_t n = F.Vown;
  // ---- aggregation: sort owned nodes by (Dirichlet bit, box / 2)
  uint64_t *keys = nullptr, *skeys = nullptr;
  uint32_t *vals = nullptr, *svals = nullptr;
  int64_t* scan = nullptr;
  CKR(lvpp_dalloc(h, &keys, (size_t)n, false));
  CKR(lvpp_dalloc(h, &skeys, (size_t)n, false));
  CKR(lvpp_dalloc(h, &vals, (size_t)n, false));
  CKR(lvpp_dalloc(h, &svals, (size_t)n, false));
  CKR(lvpp_dalloc(h, &scan, (size_t)n, false));
  LAUNCH(h, k_agg_keys, lvpp_grid(n, 256, 16), 256, 0, n, F.box, F.bc_flag, keys, vals);
  CK(cudaGetLastError());
  size_t tb = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, skeys, vals, svals, n, 0, 63, h->stream));
  void* tmp = nullptr;
  CKR(lvpp_dalloc(h, (char**)&tmp, tb, false));
  CK(cub::DeviceRadixSort::SortPairs(tmp, tb, keys, skeys, vals, svals, n, 0, 63, h->stream));
  LAUNCH(h, k_heads, lvpp_grid(n, 256, 16), 256, 0, n, skeys, scan);
  CK(cudaGetLastError());
  size_t sb = 0;
  CK(cub::DeviceScan::InclusiveSum(nullptr, sb, scan, scan, n, h->stream));
  void* stmp = nullptr;
  CKR(lvpp_dalloc(h, (char**)&stmp, sb, false));
  CK(cub::DeviceScan::InclusiveSum(stmp, sb, scan, scan, n, h->stream));
  int64_t Vc = 0;
  CK(cudaMemcpyAsync(&Vc, scan + (n - 1), sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  double tot[2] = {(double)Vc, (double)n};
  CKR(host_allreduce_sum(h, tot, 2));  // the decision is collective
  if (tot[0] >= tot[1] || tot[0] > 0.75 * tot[1]) {  // no useful coarsening
    *stop = true;
    for (void* p : {(void*)keys, (void*)skeys, (void*)vals, (void*)svals, (void*)scan, tmp, stmp}) CKR(lvpp_dfree(h, p));
    return 0;
  }
  MgLevel C;
  CKR(lvpp_dalloc(h, &F.agg, (size_t)F.V));
  CKR(lvpp_dalloc(h, &F.agg_members, (size_t)n, false));
  CKR(lvpp_dalloc(h, &F.agg_ptr, (size_t)Vc + 1));
  CKR(lvpp_dalloc(h, &C.box, (size_t)3 * Vc));
  CKR(lvpp_dalloc(h, &C.bc_flag, (size_t)(Vc + (F.V - F.Vown))));  // coarse ghosts <= fine ghosts
  LAUNCH(h, k_agg_fill, lvpp_grid(n, 256, 16), 256, 0, n, skeys, svals, scan, F.agg, F.agg_members, F.agg_ptr,
         C.box, C.bc_flag);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  for (void* p : {(void*)keys, (void*)skeys, (void*)vals, (void*)svals, (void*)scan, tmp, stmp}) CKR(lvpp_dfree(h, p));
  CKR(build_coarse_halo(h, l == 0 ? h->halo : F.halo, F, C, Vc));

  // ---- coarse pattern and the Galerkin slot map: sort every fine slot by (coarse row, coarse col)
  const int64_t S = F.slots;
  if (S >= (int64_t)0xffffffffLL) { lvpp_set_error("too many slots for the 32-bit Galerkin map"); return LVPP_E_CAPACITY; }
  CKR(lvpp_dalloc(h, &keys, (size_t)S, false));
  CKR(lvpp_dalloc(h, &skeys, (size_t)S, false));
  CKR(lvpp_dalloc(h, &vals, (size_t)S, false));
  CKR(lvpp_dalloc(h, &F.gal_src, (size_t)S, false));
  CK(cudaMemsetAsync(keys, 0xff, sizeof(uint64_t) * S, h->stream));  // slots of non-existent rows sort last
  LAUNCH(h, k_gal_keys, lvpp_grid(n, 256, 16), 256, 0, n, F.slice_ptr, F.col, F.agg, keys, vals);
  CK(cudaGetLastError());
  tb = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, skeys, vals, F.gal_src, S, 0, 64, h->stream));
  CKR(lvpp_dalloc(h, (char**)&tmp, tb, false));
  CK(cub::DeviceRadixSort::SortPairs(tmp, tb, keys, skeys, vals, F.gal_src, S, 0, 64, h->stream));
  // number of valid slots = slots of existing rows
  int64_t sp_last[2];
  CK(cudaMemcpyAsync(sp_last, F.slice_ptr + (F.nslices - 1), 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  const int64_t w_last = (sp_last[1] - sp_last[0]) / LVPP_SLICE;
  const int64_t nvalid = S - (F.nslices * LVPP_SLICE - n) * w_last;
  int64_t* scan2 = (int64_t*)keys;  // keys are no longer needed after the sort
  LAUNCH(h, k_heads, lvpp_grid(nvalid, 256, 16), 256, 0, nvalid, skeys, scan2);
  CK(cudaGetLastError());
  sb = 0;
  CK(cub::DeviceScan::InclusiveSum(nullptr, sb, scan2, scan2, nvalid, h->stream));
  CKR(lvpp_dalloc(h, (char**)&stmp, sb, false));
  CK(cub::DeviceScan::InclusiveSum(stmp, sb, scan2, scan2, nvalid, h->stream));
  int64_t nnzc = 0;
  CK(cudaMemcpyAsync(&nnzc, scan2 + (nvalid - 1), sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  C.nnz = nnzc;
  uint64_t* ukeys = nullptr;
  CKR(lvpp_dalloc(h, &ukeys, (size_t)nnzc, false));
  CKR(lvpp_dalloc(h, &F.gal_ptr, (size_t)nnzc + 1, false));
  CKR(lvpp_dalloc(h, &F.gal_dst, (size_t)nnzc, false));
  LAUNCH(h, k_compact_heads, lvpp_grid(nvalid, 256, 16), 256, 0, nvalid, skeys, scan2, ukeys, F.gal_ptr, nvalid);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(F.gal_ptr + nnzc, &nvalid, sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
  // coarse rows
  int64_t* rowstart = nullptr;
  int* d_err = nullptr;
  CKR(lvpp_dalloc(h, &rowstart, (size_t)Vc + 1));
  CKR(lvpp_dalloc(h, &d_err, 1));
  CKR(lvpp_dalloc(h, &C.rowlen, (size_t)Vc));
  LAUNCH(h, k_coarse_rowstart, lvpp_grid(Vc + 1, 256, 16), 256, 0, Vc, nnzc, ukeys, rowstart, C.rowlen);
  LAUNCH(h, k_coarse_rowlen, lvpp_grid(Vc, 256, 16), 256, 0, Vc, rowstart, C.rowlen, d_err);
  CK(cudaGetLastError());
  C.nslices = (Vc + LVPP_SLICE - 1) / LVPP_SLICE;
  int64_t* slots = nullptr;
  CKR(lvpp_dalloc(h, &slots, (size_t)C.nslices + 1));
  CKR(lvpp_dalloc(h, &C.slice_ptr, (size_t)C.nslices + 1));
  LAUNCH(h, k_slice_slots, lvpp_grid((C.nslices + 1) * 32, 256, 16), 256, 0, C.nslices, Vc, C.rowlen, slots);
  CK(cudaGetLastError());
  size_t sb2 = 0;
  CK(cub::DeviceScan::ExclusiveSum(nullptr, sb2, slots, C.slice_ptr, C.nslices + 1, h->stream));
  void* stmp2 = nullptr;
  CKR(lvpp_dalloc(h, (char**)&stmp2, sb2, false));
  CK(cub::DeviceScan::ExclusiveSum(stmp2, sb2, slots, C.slice_ptr, C.nslices + 1, h->stream));
  int herr = 0;
  CK(cudaMemcpyAsync(&C.slots, C.slice_ptr + C.nslices, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(&herr, d_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (herr) { lvpp_set_error("a coarse row has more than %d entries", LVPP_MAX_ROW); return LVPP_E_CAPACITY; }
  CKR(lvpp_dalloc(h, &C.col, (size_t)C.slots));
  CKR(lvpp_dalloc(h, &C.diag_k, (size_t)Vc));
  CKR(lvpp_dalloc(h, &C.K, (size_t)C.slots));
  CKR(lvpp_dalloc(h, &C.M, (size_t)C.slots));
  CKR(lvpp_dalloc(h, &C.D, (size_t)C.slots));
  LAUNCH(h, k_coarse_cols, lvpp_grid(Vc, 256, 16), 256, 0, Vc, rowstart, ukeys, C.slice_ptr, C.bc_flag, C.col,
         C.diag_k, F.gal_dst);
  CK(cudaGetLastError());
  // constant operators of the coarse level
  // The Galerkin stiffness of a piecewise-constant prolongation over-estimates the energy of smooth functions by about
  // the aggregate width (2 per coarsening): K_c is divided by mg_kscale here, level after level, instead of (or
  // together with) multiplying the coarse correction by mg_over.  The two are the same thing for the elliptic block of a
  // two-level method; on the contact set, where the roles of the blocks are exchanged (psi_c = M_c^-1 (r - alpha K_c u_c)),
  // only the rescaled stiffness gives a coarse correction of the right size.  M_c and D_c are exact for constants.
  LAUNCH(h, k_galerkin, lvpp_grid(nnzc, 256, 16), 256, 0, nnzc, F.gal_ptr, F.gal_src, F.gal_dst, 2, F.K, F.M, C.K, C.M,
         1.0 / h->mg_kscale);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  for (void* p : {(void*)keys, (void*)skeys, (void*)vals, tmp, stmp, (void*)ukeys, (void*)rowstart, (void*)d_err,
                  (void*)slots, stmp2})
    CKR(lvpp_dfree(h, p));
  CKR(level_alloc_vectors(h, C));
  h->levels.push_back(C);
  return 0;
}

static double env_double(const char* name, double dflt) {
  const char* v = getenv(name);
  return v && *v ? atof(v) : dflt;
}

int lvpp_mg_setup(lvpp_problem* h) {
  if (h->mg_ready) return 0;
  if (h->nranks > 64) { lvpp_set_error("multigrid preconditioner: more than 64 ranks"); return LVPP_E_INVALID; }
  // tunables (defaults are the tested values)
  h->mg_over = env_double("LVPP_MG_OVER", h->mg_over);
  h->mg_kscale = env_double("LVPP_MG_KSCALE", h->mg_kscale);
  if (!(h->mg_kscale > 0.0)) { lvpp_set_error("bad LVPP_MG_KSCALE"); return LVPP_E_INVALID; }
  h->mg_omega = env_double("LVPP_MG_OMEGA", h->mg_omega);
  h->mg_nsmooth = (int)env_double("LVPP_MG_NSMOOTH", h->mg_nsmooth);
  if (h->mg_nsmooth < 1 || h->mg_nsmooth > MG_MAX_SWEEPS) { lvpp_set_error("bad LVPP_MG_NSMOOTH"); return LVPP_E_INVALID; }
  h->mg_npre = (int)env_double("LVPP_MG_NPRE", getenv("LVPP_MG_NSMOOTH") ? h->mg_nsmooth : h->mg_npre);
  h->mg_npost = (int)env_double("LVPP_MG_NPOST", getenv("LVPP_MG_NSMOOTH") ? h->mg_nsmooth : h->mg_npost);
  if (h->mg_npre < 0 || h->mg_npre > MG_MAX_SWEEPS || h->mg_npost < 1 || h->mg_npost > MG_MAX_SWEEPS) {
    lvpp_set_error("bad LVPP_MG_NPRE / LVPP_MG_NPOST");
    return LVPP_E_INVALID;
  }
  h->mg_cheb = env_double("LVPP_MG_CHEB", h->mg_cheb);
  h->mg_cheb_adapt = env_double("LVPP_MG_CHEB_ADAPT", 0.0) != 0.0;
  h->mg_margin = env_double("LVPP_MG_MARGIN", h->mg_margin);
  h->mg_power_its = (int)env_double("LVPP_MG_POWER_ITS", h->mg_power_its);
  h->gm_eta2 = env_double("LVPP_GMRES_ETA2", h->gm_eta2);
  if (const char* gw = getenv("LVPP_GMRES_WEIGHT")) h->gm_weight_auto = strcmp(gw, "off") != 0;
  h->gm_warm = env_double("LVPP_GMRES_WARM", 1.0) != 0.0;
  // GMRES workspace (also the scratch of the collective decisions below)
  h->gm_restart = (int)env_double("LVPP_GMRES_RESTART", 50);
  if (h->gm_restart < 2 || h->gm_restart > 200) { lvpp_set_error("bad LVPP_GMRES_RESTART"); return LVPP_E_INVALID; }
  CKR(lvpp_dalloc(h, &h->gm_V, (size_t)(h->gm_restart + 1) * 2 * h->V));
  CKR(lvpp_dalloc(h, &h->gm_h, (size_t)h->gm_restart + 72));
  {
    const size_t m = (size_t)h->gm_restart;
    CKR(lvpp_dalloc(h, &h->gm_state, 1));
    CKR(lvpp_dalloc(h, &h->gm_H, (m + 1) * m));
    CKR(lvpp_dalloc(h, &h->gm_cs, m));
    CKR(lvpp_dalloc(h, &h->gm_sn, m));
    CKR(lvpp_dalloc(h, &h->gm_g, m + 1));
    CKR(lvpp_dalloc(h, &h->gm_yv, m));
    CKR(lvpp_dalloc(h, &h->gm_red, 8));
    CK(cudaMallocHost((void**)&h->gm_state_host, sizeof(GmState) * (GM_RING + 2)));
    for (int i = 0; i < GM_RING; ++i) {
      CK(cudaEventCreateWithFlags(&h->gm_ev[i], cudaEventDisableTiming));
      CK(cudaEventCreate(&h->evs0_ring[i]));
      CK(cudaEventCreate(&h->evs1_ring[i]));
      CK(cudaEventCreate(&h->evp0_ring[i]));
      CK(cudaEventCreate(&h->evp1_ring[i]));
    }
  }
  CKR(lvpp_dalloc(h, &h->gm_part, (size_t)(h->gm_restart + 2) * h->npartials));
  CK(cudaMallocHost((void**)&h->gm_h_host, sizeof(double) * (h->gm_restart + 72)));
  h->levels.clear();
  MgLevel L0;
  L0.V = h->V; L0.Vown = h->Vown; L0.nslices = h->nslices; L0.slots = h->sell_slots; L0.nnz = h->scalar_nnz;
  L0.slice_ptr = h->slice_ptr; L0.col = h->col; L0.rowlen = h->rowlen; L0.diag_k = h->diag_k;
  L0.K = h->K; L0.M = h->M; L0.D = h->D; L0.bc_flag = h->bc_flag;
  // level 0 uses h->halo itself (the peer-memory halo keeps a sequence counter per level)
  if (h->nranks > 1 && h->halo.num_neighbors > 0) {
    // number of every ghost node on its owner (needed when the fine level is also the coarsest)
    LevelHalo& H = h->halo;
    const int64_t nr = H.recv_ptr.back();
    int32_t* d_r = nullptr;
    CKR(lvpp_dalloc(h, &d_r, (size_t)nr, false));
    CKR(lvpp_halo_exchange_i32(h, H, H.send_nodes, d_r));
    H.ghost_owner_local.resize((size_t)nr);
    H.ghost_nbr.resize((size_t)nr);
    CK(cudaMemcpyAsync(H.ghost_owner_local.data(), d_r, sizeof(int32_t) * nr, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int b = 0; b < H.num_neighbors; ++b)
      for (int64_t p = H.recv_ptr[b]; p < H.recv_ptr[b + 1]; ++p) H.ghost_nbr[p] = b;
    CKR(lvpp_dfree(h, d_r));
  }
  CKR(lvpp_dalloc(h, &L0.box, (size_t)3 * L0.Vown));
  LAUNCH(h, k_box0, lvpp_grid(L0.Vown, 256, 16), 256, 0, L0.Vown, h->tdim, h->coords, h->xmin[0], h->xmin[1],
         h->xmin[2], 1.0 / h->h0, L0.box);
  CK(cudaGetLastError());
  CKR(level_alloc_vectors(h, L0));
  h->levels.push_back(L0);
  for (int l = 0; l < MG_MAX_LEVELS - 1; ++l) {
    double tot = (double)h->levels[l].Vown;
    CKR(host_allreduce_sum(h, &tot, 1));
    if (tot <= (double)MG_COARSE_TARGET) break;
    bool stop = false;
    CKR(build_next_level(h, l, &stop));
    if (stop) break;
  }
  h->mg_fp32 = env_double("LVPP_MG_FP32", 1.0) != 0.0;
  // records of the cycle's operator copy: bf16 pairs (10 bytes per slot; the default from round 2: same Newton and
  // Krylov counts as the single-precision records over the whole n = 215 solve, 15 % less time --
  // profiles/r02_scan_const_alpha.txt) or LVPP_MG_PACK=fp32 (16 bytes per slot)
  h->mg_bf16 = h->mg_fp32;
  if (const char* pk = getenv("LVPP_MG_PACK")) h->mg_bf16 = h->mg_fp32 && strcmp(pk, "fp32") != 0;
  if (h->mg_fp32)
    for (size_t l = 0; l + 1 < h->levels.size(); ++l) {
      MgLevel& L = h->levels[l];
      if (h->mg_bf16) {  // pair records: slots / 2 + 16 per slice (block_op.cuh)
        const size_t npairs = (size_t)(L.slots >> 1) + 16 * (size_t)L.nslices;
        CKR(lvpp_dalloc(h, &L.P2, npairs, false));
        CKR(lvpp_dalloc(h, &L.Pd, npairs, false));
        CKR(lvpp_dalloc(h, &L.binv32, (size_t)L.Vown, false));
      } else {
        CKR(lvpp_dalloc(h, &L.P, (size_t)L.slots, false));
      }
    }
  // coarsest level: global numbering of all ranks' coarse nodes, rank by rank
  const MgLevel& Lc = h->levels.back();
  std::vector<double> cnt((size_t)h->nranks, 0.0);
  cnt[h->rank] = (double)Lc.Vown;
  CKR(host_allreduce_sum(h, cnt.data(), h->nranks));
  std::vector<int64_t> off((size_t)h->nranks + 1, 0);
  for (int r = 0; r < h->nranks; ++r) off[r + 1] = off[r] + (int64_t)(cnt[r] + 0.5);
  if (off[h->nranks] > MG_COARSE_MAX) {
    lvpp_set_error("multigrid: coarsening stalled at %lld nodes (limit %d)", (long long)off[h->nranks], MG_COARSE_MAX);
    return LVPP_E_CAPACITY;
  }
  h->coarse_n = (int)(2 * off[h->nranks]);
  h->coarse_off = off[h->rank];
  {
    std::vector<int32_t> gmap((size_t)Lc.V, -1);
    for (int64_t i = 0; i < Lc.Vown; ++i) gmap[i] = (int32_t)(h->coarse_off + i);
    const LevelHalo& H = h->levels.size() == 1 ? h->halo : Lc.halo;
    for (size_t p = 0; p < H.recv_nodes_host.size() && h->nranks > 1; ++p)
      gmap[H.recv_nodes_host[p]] = (int32_t)(off[H.neighbor_ranks[H.ghost_nbr[p]]] + H.ghost_owner_local[p]);
    for (int64_t i = 0; i < Lc.V; ++i)
      if (gmap[i] < 0) { lvpp_set_error("multigrid: ghost node %lld of the coarsest level has no owner", (long long)i); return LVPP_E_INVALID; }
    CKR(lvpp_dalloc(h, &h->coarse_gmap, (size_t)Lc.V, false));
    CK(cudaMemcpyAsync(h->coarse_gmap, gmap.data(), sizeof(int32_t) * Lc.V, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  CKR(lvpp_dalloc(h, &h->coarse_lu, (size_t)h->coarse_n * h->coarse_n));
  CKR(lvpp_dalloc(h, &h->coarse_bg, (size_t)h->coarse_n));
  CK(cudaStreamSynchronize(h->stream));
  if (getenv("LVPP_MG_VERBOSE") && h->rank == 0) {
    fprintf(stderr, "[lvpp mg] %d levels:", (int)h->levels.size());
    for (const MgLevel& L : h->levels) fprintf(stderr, " %lld(+%lld)", (long long)L.Vown, (long long)(L.V - L.Vown));
    fprintf(stderr, "; coarsest dense n = %d; halo %s\n", h->coarse_n,
            h->nranks > 1 ? (h->halo.p2p ? "peer memory" : "nccl send/recv") : "none");
  }
  h->mg_ready = true;
  return 0;
}

static int level_op_local(lvpp_problem* h, MgLevel& L, int epi, double omega, const double* v, const double* b,
                         double* y, bool f32 = false) {
  const int grid = lvpp_grid(L.Vown, 256, 6);
  if (&L == &h->levels[0]) h->fine_op_launches++;
  if (f32 && L.P2) {
    Packed2OpArgs q;
    q.Vown = L.Vown; q.slice_ptr = L.slice_ptr; q.P2 = L.P2; q.Pd = L.Pd; q.bc_flag = L.bc_flag;
    q.v = (const double2*)v; q.y = (double2*)y; q.epi = epi; q.b = (const double2*)b; q.binv = L.binv32; q.omega = omega;
    q.skip = h->gm_skip;
    const bool sample = &L == &h->levels[0] && epi == EPI_JACOBI && h->smooth_sample_pending;
    if (sample) CK(cudaEventRecord(h->evp0_cur, h->stream));
    LAUNCH(h, (k_packed2_op<2, 4>), lvpp_grid(L.Vown, 256, 4), 256, 0, q);
    if (&L == &h->levels[0]) h->packed_op_launches++;
    if (sample) {
      CK(cudaEventRecord(h->evp1_cur, h->stream));
      h->smooth_sample_pending = false;
      h->smooth_sample_recorded = true;
    }
  } else if (f32 && L.P) {
    PackedOpArgs q;
    q.Vown = L.Vown; q.slice_ptr = L.slice_ptr; q.P = L.P; q.bc_flag = L.bc_flag;
    q.v = (const double2*)v; q.y = (double2*)y; q.epi = epi; q.b = (const double2*)b; q.binv = L.binv; q.omega = omega;
    q.skip = h->gm_skip;
    const bool sample = &L == &h->levels[0] && epi == EPI_JACOBI && h->smooth_sample_pending;
    if (sample) CK(cudaEventRecord(h->evp0_cur, h->stream));
    LAUNCH(h, (k_packed_op<4, 4>), lvpp_grid(L.Vown, 256, 4), 256, 0, q);
    if (&L == &h->levels[0]) h->packed_op_launches++;
    if (sample) {
      CK(cudaEventRecord(h->evp1_cur, h->stream));
      h->smooth_sample_pending = false;
      h->smooth_sample_recorded = true;
    }
  } else {
    OpArgs p = lvpp_level_op(h, L);
    p.v = (const double2*)v;
    p.y = (double2*)y;
    p.epi = epi;
    p.b = (const double2*)b;
    p.binv = L.binv;
    p.omega = omega;
    p.skip_flag = h->gm_skip;
    LAUNCH(h, (k_block_op<0>), grid, 256, 0, p);
  }
  CK(cudaGetLastError());
  return 0;
}

// ghost update of v, then the operator (smoother / cycle residual: single-precision values when enabled)
static int level_op(lvpp_problem* h, MgLevel& L, int epi, double omega, const double* v, const double* b, double* y,
                    bool f32 = true) {
  if (h->nranks > 1) CKR(lvpp_halo_forward_level(h, &L == &h->levels[0] ? h->halo : L.halo, const_cast<double*>(v)));
  return level_op_local(h, L, epi, omega, v, b, y, f32 && h->mg_fp32);
}

static int build_binv(lvpp_problem* h, MgLevel& L) {
  LAUNCH(h, k_build_binv, lvpp_grid(L.Vown, 256, 8), 256, 0, L.Vown, L.slice_ptr, L.diag_k, L.K, L.M, L.D,
         L.bc_flag, h->alpha, L.binv, L.binv32);
  CK(cudaGetLastError());
  return 0;
}

// lambda_max(Binv J) of one level by power iteration on I + Binv J (the Jacobi epilogue with b = 0 and
// omega = -1).  The norm growth of an unconverged vector is a LOWER estimate, so the first call on a level runs
// three times as long from a fixed pseudo-random vector and later calls (alpha changed, the operator moved a
// little) restart from the previous vector.  Everything is deterministic: the same solve gives the same estimates.
static int estimate_lambda(lvpp_problem* h, MgLevel& L) {
  const int nb = h->npartials;
  const int nit = (L.ev_valid ? h->mg_power_its : 3 * h->mg_power_its) * h->mg_power_boost;
  if (!L.ev_valid) {
    LAUNCH(h, k_ev_init, lvpp_grid(L.Vown, 256, 8), 256, 0, L.Vown, (double2*)L.ev);
    CK(cudaGetLastError());
  }
  L.ev_valid = true;
  CK(cudaMemsetAsync(L.x, 0, sizeof(double) * 2 * L.V, h->stream));
  double lam = 0.0;
  for (int it = 0; it <= nit; ++it) {
    // normalise ev, then t = ev + Binv J ev
    LAUNCH(h, k_multi_dot, nb, 256, 0, L.Vown, (const double2*)L.ev, L.V, 0, 1, (const double2*)L.ev, nb, h->gm_part, 1.0);
    CK(cudaGetLastError());
    CKR(reduce_to_host(h, h->gm_part, 1, h->gm_h, h->gm_h_host));
    const double nrm = sqrt(h->gm_h_host[0]);
    if (!(nrm > 0.0) || !std::isfinite(nrm)) { lvpp_set_error("multigrid: eigenvalue estimate failed"); return LVPP_E_INVALID; }
    if (it > 0) lam = nrm - 1.0;  // ||(I + Binv J) ev|| with ||ev|| = 1
    if (it == nit) {
      LAUNCH(h, k_axpby, nb, 256, 0, L.Vown, 1.0 / nrm, (const double2*)L.ev, 0, (double2*)L.ev);
      CK(cudaGetLastError());
      break;
    }
    LAUNCH(h, k_axpby, nb, 256, 0, L.Vown, 1.0 / nrm, (const double2*)L.ev, 0, (double2*)L.ev);
    CK(cudaGetLastError());
    CKR(level_op(h, L, EPI_JACOBI, -1.0, L.ev, L.x, L.t));
    std::swap(L.ev, L.t);
  }
  // lambda_max(Binv J) is set by the stiffness block -- mesh and element -- and moves by a few per cent over a
  // whole LVPP solve (2.37 - 2.49 on a 12 x 12 x 16 mesh with 200 iterations per estimate), while the power
  // iteration converges ever more slowly once exp(psi) varies over many orders of magnitude (the estimate at the
  // second proximal step of the n = 368 solve fell from 2.41 to 1.34 and the Chebyshev sweeps diverged).  The
  // estimate is therefore monotone per handle: an over-estimate only makes the polynomial a little less sharp.
  L.lambda = std::max(L.lambda, lam);
  return 0;
}

// after a failed Krylov solve: forget the eigenvalue estimates, redo them from the fixed start vector with ten
// times the iterations and fall back to plain damping (krylov.cu retries the solve once)
int lvpp_mg_reestimate(lvpp_problem* h) {
  for (MgLevel& L : h->levels) {
    L.ev_valid = false;
    L.lambda = 0.0;
  }
  h->mg_alpha_est = -1.0;
  h->mg_cheb = 0.0;  // the retry (and the rest of this handle's life) uses plain damping: robust to eigenvalues off the interval
  h->mg_power_boost = 10;
  const int r = lvpp_mg_update(h);
  h->mg_power_boost = 1;
  h->mg_retries++;
  if (getenv("LVPP_MG_VERBOSE") && h->rank == 0) fprintf(stderr, "[lvpp mg] Krylov solve failed: eigenvalue estimates redone, retrying\n");
  return r;
}

// per Newton step: coarse D, damping, node-block inverses, dense inverse of the coarsest operator
int lvpp_mg_update(lvpp_problem* h) {
  CKR(lvpp_mg_setup(h));
  const int nl = (int)h->levels.size();
  for (int l = 0; l + 1 < nl; ++l) {
    MgLevel& F = h->levels[l];
    MgLevel& C = h->levels[l + 1];
    LAUNCH(h, k_galerkin, lvpp_grid(C.nnz, 256, 16), 256, 0, C.nnz, F.gal_ptr, F.gal_src, F.gal_dst, 1, F.D, F.D, C.D, C.D, 1.0);
    CK(cudaGetLastError());
  }
  if (h->mg_fp32)
    for (int l = 0; l + 1 < nl; ++l) {
      MgLevel& L = h->levels[l];
      if (L.P2)
        LAUNCH(h, k_pack_op_bf16, lvpp_grid(32 * L.nslices, 256, 16), 256, 0, L.nslices, L.slice_ptr, L.col, L.K, L.M, L.D,
               h->alpha, L.P2, L.Pd);
      else
        LAUNCH(h, k_pack_op, lvpp_grid(L.slots, 256, 16), 256, 0, L.slots, L.col, L.K, L.M, L.D, h->alpha, L.P);
      CK(cudaGetLastError());
    }
  // lambda_max(Binv J) is set by the stiffness block (mesh and element, much less by psi), so it is estimated when
  // alpha changes (once per proximal step), with a margin for the estimate's deficit and its drift over the Newton
  // steps.  The dampings are recomputed from it at every update (the Chebyshev ratio may have been adapted by
  // lvpp_solve_linear in between).
  const bool estimate = !(h->mg_alpha_est == h->alpha);
  for (int l = 0; l + 1 < nl; ++l) {
    MgLevel& L = h->levels[l];
    CKR(build_binv(h, L));
    if (estimate) CKR(estimate_lambda(h, L));
    // sweep k of a smoothing step is damped by omega_k.  Plain damped Jacobi: omega_k = omega.  Chebyshev
    // (mg_cheb = ratio > 1): 1 / omega_k are the roots of the degree-nsmooth Chebyshev polynomial of the
    // interval [b / ratio, b], b = margin * lambda_max -- the sweeps are the same kernel at the same cost, their
    // product is the polynomial that is smallest on the upper part of the spectrum.  An eigenvalue above b is
    // amplified, which is why the estimate is monotone, warm-started and carries a margin.
    L.omega = h->mg_omega * std::min(1.0, 2.0 / (h->mg_margin * L.lambda));
    for (int side = 0; side < 2; ++side) {  // pre- and post-smoothing polynomials may have different degrees
      const int m = side == 0 ? h->mg_npre : h->mg_npost;
      double* om = side == 0 ? L.sweep_omega_pre : L.sweep_omega;
      for (int k = 0; k < m; ++k) {
        if (h->mg_cheb > 1.0) {
          const double b = h->mg_margin * L.lambda, a = b / h->mg_cheb;
          om[k] = 1.0 / (0.5 * (b + a) + 0.5 * (b - a) * cos(M_PI * (2 * k + 1) / (2.0 * m)));
        } else {
          om[k] = L.omega;
        }
      }
    }
  }
  h->mg_alpha_est = h->alpha;
  if (estimate && getenv("LVPP_MG_VERBOSE") && h->rank == 0) {
    fprintf(stderr, "[lvpp mg] lambda_max / omega:");
    for (int l = 0; l + 1 < nl; ++l)
      fprintf(stderr, " %.3f/%.3f,%.3f", h->levels[l].lambda, h->levels[l].sweep_omega[0], h->levels[l].sweep_omega[h->mg_npost - 1]);
    fprintf(stderr, "\n");
  }
  MgLevel& Lc = h->levels.back();
  const int n = h->coarse_n;
  double* aug = nullptr;
  int* d_err = nullptr;
  CKR(lvpp_dalloc(h, &aug, (size_t)n * 2 * n));
  CKR(lvpp_dalloc(h, &d_err, 1));
  LAUNCH(h, k_coarse_dense, lvpp_grid(Lc.Vown, 128, 4), 128, 0, Lc.Vown, n, h->coarse_gmap, Lc.slice_ptr, Lc.rowlen,
         Lc.col, Lc.K, Lc.M, Lc.D, Lc.bc_flag, h->alpha, aug);
  CK(cudaGetLastError());
  if (h->nranks > 1) {  // every rank holds the whole coarsest operator and inverts it redundantly
    const size_t total = (size_t)n * 2 * n;
    for (size_t o = 0; o < total; o += (size_t)1 << 30)
      CKR(lvpp_allreduce_sum(h, aug + o, (int)std::min<size_t>((size_t)1 << 30, total - o)));
  }
  LAUNCH(h, k_gauss_jordan, 1, 1024, 0, n, aug, h->coarse_lu, d_err);
  CK(cudaGetLastError());
  int herr = 0;
  CK(cudaMemcpyAsync(&herr, d_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CKR(lvpp_dfree(h, aug));
  CKR(lvpp_dfree(h, d_err));
  if (herr) { lvpp_set_error("multigrid: singular coarsest operator"); return LVPP_E_INVALID; }
  return 0;
}

// z = V-cycle(b) with zero initial guess on level 0; returns the buffer that holds z (owned entries)
int lvpp_mg_vcycle(lvpp_problem* h, const double* b_in, double** z_out) {
  const int nl = (int)h->levels.size();
  std::vector<double*> cur(nl), oth(nl);
  std::vector<const double*> rhs(nl);
  const int npre = h->mg_npre, npost = h->mg_npost;
  rhs[0] = b_in;
  for (int l = 0; l < nl - 1; ++l) {
    MgLevel& L = h->levels[l];
    cur[l] = L.x; oth[l] = L.t;
    const double* res = rhs[l];  // no pre-smoothing: the iterate is zero and the residual is the right-hand side
    if (npre == 0) {
      CK(cudaMemsetAsync(cur[l], 0, sizeof(double) * 2 * L.Vown, h->stream));
    } else {
      LAUNCH(h, k_smooth_first, lvpp_grid(L.Vown, 256, 6), 256, 0, L.Vown, (const double2*)rhs[l], L.binv, L.bc_flag,
             L.sweep_omega_pre[0], (double2*)cur[l]);
      CK(cudaGetLastError());
      for (int s = 1; s < npre; ++s) {
        CKR(level_op(h, L, EPI_JACOBI, L.sweep_omega_pre[s], cur[l], rhs[l], oth[l]));
        std::swap(cur[l], oth[l]);
      }
      CKR(level_op(h, L, EPI_RESID, L.omega, cur[l], rhs[l], oth[l]));  // residual into the spare buffer
      res = oth[l];
    }
    MgLevel& C = h->levels[l + 1];
    LAUNCH(h, k_restrict, lvpp_grid(C.Vown, 256, 6), 256, 0, C.Vown, L.agg_ptr, L.agg_members, (const double2*)res,
           C.bc_flag, (double2*)C.b);
    CK(cudaGetLastError());
    rhs[l + 1] = C.b;
  }
  {  // coarsest: gather the right-hand side of all ranks, x = inv * b for the owned unknowns
    MgLevel& L = h->levels[nl - 1];
    cur[nl - 1] = L.x; oth[nl - 1] = L.t;
    const int n = h->coarse_n;
    const double* bg = rhs[nl - 1];
    if (h->nranks > 1) {
      CK(cudaMemsetAsync(h->coarse_bg, 0, sizeof(double) * n, h->stream));
      CK(cudaMemcpyAsync(h->coarse_bg + 2 * h->coarse_off, rhs[nl - 1], sizeof(double) * 2 * L.Vown,
                         cudaMemcpyDeviceToDevice, h->stream));
      CKR(lvpp_allreduce_sum(h, h->coarse_bg, n));
      bg = h->coarse_bg;
    }
    const int nloc = (int)(2 * L.Vown);
    LAUNCH(h, k_coarse_apply, (nloc + 127) / 128, 128, sizeof(double) * n, n, nloc, (int)(2 * h->coarse_off), h->coarse_lu,
           bg, cur[nl - 1]);
    CK(cudaGetLastError());
  }
  for (int l = nl - 2; l >= 0; --l) {
    MgLevel& L = h->levels[l];
    LAUNCH(h, k_prolong_add, lvpp_grid(L.Vown, 256, 6), 256, 0, L.Vown, L.agg, (const double2*)cur[l + 1], L.bc_flag,
           h->mg_over, (double2*)cur[l]);
    CK(cudaGetLastError());
    for (int s = 0; s < npost; ++s) {
      CKR(level_op(h, L, EPI_JACOBI, L.sweep_omega[s], cur[l], rhs[l], oth[l]));
      std::swap(cur[l], oth[l]);
    }
  }
  h->vcycles++;
  *z_out = cur[0];
  return 0;
}

// partial sums of K_ii and M_ii over the owned free nodes: partials[0 * nparts + block], partials[1 * nparts + block]
__global__ void __launch_bounds__(256) k_diag_sums(int64_t Vown, const int64_t* __restrict__ slice_ptr,
                                                   const uint8_t* __restrict__ diag_k, const double* __restrict__ K,
                                                   const double* __restrict__ M, const uint8_t* __restrict__ bc, int nparts,
                                                   double* __restrict__ partials) {
  __shared__ double s_red[32];
  double sk = 0.0, sm = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < Vown; i += (int64_t)gridDim.x * blockDim.x) {
    if (bc[i]) continue;
    const int64_t idx = slice_ptr[i >> 5] + (i & 31) + (int64_t)diag_k[i] * LVPP_SLICE;
    sk += K[idx];
    sm += M[idx];
  }
  const double rk = lvpp_block_sum<256>(sk, s_red);
  if (threadIdx.x == 0) partials[blockIdx.x] = rk;
  const double rm = lvpp_block_sum<256>(sm, s_red);
  if (threadIdx.x == 0) partials[(int64_t)nparts + blockIdx.x] = rm;
}

// ------------------------------------------------------------------------------------------------
// right-preconditioned restarted GMRES:  J M^-1 (M y) = rhs
static int reduce_to_host(lvpp_problem* h, double* partials, int nvals, double* dst_dev, double* dst_host) {
  LAUNCH(h, k_reduce_multi, nvals < 64 ? nvals : 64, 256, 0, h->npartials, nvals, partials, dst_dev);
  CK(cudaGetLastError());
  if (h->nranks > 1) CKR(lvpp_allreduce_sum(h, dst_dev, nvals));
  CK(cudaMemcpyAsync(dst_host, dst_dev, sizeof(double) * nvals, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int lvpp_gmres_mg(lvpp_problem* h, const double* d_rhs, double* d_y, const lvpp_newton_opts* o, int32_t* its_out,
                  int32_t* reason_out, double* rnorm_out) {
  // Restarted GMRES with the whole recurrence on the device (gmres_kernels.cuh: GmState, k_gm_*): an iteration is a
  // fixed list of launches -- cycle, J*v, two guarded Gram-Schmidt passes, the Givens update of the Hessenberg column,
  // the scaling of the new basis vector -- with NO host round trip.  The host reads a copy of the state one iteration
  // late (it waits for the copy of iteration j - 1 only after iteration j is queued, so the stream never drains) and
  // stops queueing once `conv` is set; whatever was queued behind the deciding iteration returns at once.
  const int m = h->gm_restart;
  const int64_t Vown = h->Vown, stride2 = h->V;  // basis vectors are 2V doubles = V double2
  const int nb = h->npartials;
  double2* Vb = (double2*)h->gm_V;
  double* gpart = h->gm_part;  // partial sums [(m + 2) * nb]
  const int maxit = o->ksp_max_it > 0 ? o->ksp_max_it : 1000;
  CK(cudaEventRecord(h->ev0, h->stream));
  auto vec = [&](int k) { return (double*)(Vb + (int64_t)k * stride2); };
  MgLevel& L0 = h->levels[0];
  // Equilibrated residual norm (the default; LVPP_GMRES_WEIGHT=off for the Euclidean one): the rows of the u equation
  // carry entries of size alpha K_ii, those of the psi equation entries of size M_ii = O(h^2 / alpha) times that, and a
  // Euclidean residual norm all but ignores the latter.  With the weight wy = sum alpha K_ii / sum M_ii on the psi
  // component of every inner product GMRES runs on S J S, S = diag(1, sqrt(wy)), without touching the operator or the
  // cycle.  Measured at n = 215 (profiles/r02_diag215_*.txt): 22 - 29 Krylov iterations per Newton step through the
  // first two proximal steps against 27 - 57, identical Newton iterates (both norms leave residuals of 1e-15 of the
  // right-hand side in BOTH row blocks at ksp_rtol 1e-12).
  double wy = 1.0;
  if (h->gm_weight_auto) {
    if (!(h->gm_weight_alpha == h->alpha)) {
      LAUNCH(h, k_diag_sums, nb, 256, 0, Vown, L0.slice_ptr, L0.diag_k, L0.K, L0.M, L0.bc_flag, nb, gpart);
      CK(cudaGetLastError());
      CKR(reduce_to_host(h, gpart, 2, h->gm_h, h->gm_h_host));
      if (!(h->gm_h_host[0] > 0.0) || !(h->gm_h_host[1] > 0.0)) { lvpp_set_error("GMRES weight: bad diagonal sums"); return LVPP_E_INVALID; }
      h->gm_weight = h->alpha * h->gm_h_host[0] / h->gm_h_host[1];
      h->gm_weight_alpha = h->alpha;
    }
    wy = h->gm_weight;
  }
  GmState* st = h->gm_state;
  GmState* ring = h->gm_state_host;            // [0, GM_RING): lagged copies; [GM_RING]: cycle-end copy; [GM_RING + 1]: upload
  GmState init;
  memset(&init, 0, sizeof(init));
  init.rtol = o->ksp_rtol; init.atol = o->ksp_atol; init.eta2 = h->gm_eta2; init.maxit = maxit; init.first = 1;
  // Warm start (h->gm_warm_next, set by lvpp_newton_step; LVPP_GMRES_WARM=0 disables it): d_y still holds the Newton
  // correction of the previous solve on this handle.  Late in an LVPP solve consecutive proximal steps take one Newton
  // step each and their corrections are nearly the same vector (psi falls by alpha lambda on the contact set every
  // time), so r0 = rhs - J y_prev is orders of magnitude below ||rhs||.  The stopping test stays ||r|| <= rtol ||rhs||
  // (PETSc's KSPConvergedDefault with a non-zero initial guess); the guess is used only if it is better than zero.
  bool have_r0 = false;
  if (h->gm_warm_next && h->gm_warm) {
    CKR(level_op(h, L0, EPI_RESID, 1.0, d_y, d_rhs, vec(0), false));
    LAUNCH(h, k_multi_dot, nb, 256, 0, Vown, Vb, stride2, 0, 1, (const double2*)vec(0), nb, gpart, wy);
    LAUNCH(h, k_multi_dot, nb, 256, 0, Vown, (const double2*)d_rhs, stride2, 0, 1, (const double2*)d_rhs, nb, gpart + nb, wy);
    CK(cudaGetLastError());
    CKR(reduce_to_host(h, gpart, 2, h->gm_h, h->gm_h_host));
    const double r0n = sqrt(h->gm_h_host[0]), bn = sqrt(h->gm_h_host[1]);
    if (std::isfinite(r0n) && r0n < bn) {
      have_r0 = true;
      init.first = 0;
      init.bnorm = bn;
      init.tol = std::max(o->ksp_rtol * bn, o->ksp_atol);
      h->gm_warm_used++;
    }
  }
  h->gm_warm_next = false;
  ring[GM_RING + 1] = init;
  CK(cudaMemcpyAsync(st, &ring[GM_RING + 1], sizeof(GmState), cudaMemcpyHostToDevice, h->stream));
  const int* skip = &st->conv;
  if (!have_r0) CK(cudaMemsetAsync(d_y, 0, sizeof(double) * 2 * h->V, h->stream));
  bool smooth_rec[GM_RING] = {false};
  auto take_samples = [&](int slot) -> int {  // the event pairs of a finished iteration
    float sms = 0.f;
    CK(cudaEventElapsedTime(&sms, h->evs0_ring[slot], h->evs1_ring[slot]));
    h->spmv_sampled_ms += sms;
    h->spmv_samples++;
    if (smooth_rec[slot]) {
      CK(cudaEventElapsedTime(&sms, h->evp0_ring[slot], h->evp1_ring[slot]));
      h->smooth_sampled_ms += sms;
      h->smooth_samples++;
    }
    return 0;
  };
  int reason = 0;
  bool first = true;
  GmState fin;
  memset(&fin, 0, sizeof(fin));
  while (true) {
    // r = rhs - J y  (first cycle: y = 0)
    if (first) {
      if (!have_r0)  // (warm start: v_0 already holds r0 = rhs - J y_prev)
        CK(cudaMemcpyAsync(vec(0), d_rhs, sizeof(double) * 2 * Vown, cudaMemcpyDeviceToDevice, h->stream));
      first = false;
    } else {
      CKR(level_op(h, L0, EPI_RESID, 1.0, d_y, d_rhs, vec(0), false));
    }
    LAUNCH(h, k_multi_dot, nb, 256, 0, Vown, Vb, stride2, 0, 1, (const double2*)vec(0), nb, gpart, wy);
    LAUNCH(h, k_reduce_multi, 1, 256, 0, nb, 1, gpart, h->gm_red);
    CK(cudaGetLastError());
    if (h->nranks > 1) CKR(lvpp_allreduce_sum(h, h->gm_red, 1));
    LAUNCH(h, k_gm_cycle_begin, 1, 1, 0, st, h->gm_red, h->gm_g, m);
    LAUNCH(h, k_gm_scale, nb, 256, 0, Vown, st, (double2*)vec(0));
    CK(cudaGetLastError());
    int j = 0, polled = 0;  // iterations queued / iterations whose state copy the host has looked at
    bool over = false;
    for (; j < m && !over; ++j) {
      const int slot = j % GM_RING;
      // w = J M^-1 v_j  -> stored in v_{j+1}
      double* z = nullptr;
      h->gm_skip = skip;
      h->evp0_cur = h->evp0_ring[slot];
      h->evp1_cur = h->evp1_ring[slot];
      h->smooth_sample_pending = true;
      h->smooth_sample_recorded = false;
      int rc = lvpp_mg_vcycle(h, vec(j), &z);
      h->smooth_sample_pending = false;
      smooth_rec[slot] = h->smooth_sample_recorded;
      if (rc) { h->gm_skip = nullptr; return rc; }
      if (h->nranks > 1) { rc = lvpp_halo_forward_level(h, h->halo, z); if (rc) { h->gm_skip = nullptr; return rc; } }
      CK(cudaEventRecord(h->evs0_ring[slot], h->stream));
      rc = level_op_local(h, L0, EPI_NONE, 1.0, z, nullptr, vec(j + 1));
      h->gm_skip = nullptr;
      if (rc) return rc;
      CK(cudaEventRecord(h->evs1_ring[slot], h->stream));
      double* Hcol = h->gm_H + (size_t)j * (m + 1);
      for (int pass = 0; pass < 2; ++pass) {
        // classical Gram-Schmidt; pass 1 (selective re-orthogonalisation) returns at once unless pass 0 asked for it
        for (int k0 = 0; k0 <= j; k0 += GM_CHUNK) {
          const int nv = std::min(GM_CHUNK, j + 1 - k0);
          LAUNCH(h, k_multi_dot, nb, 256, 0, Vown, Vb, stride2, k0, nv, (const double2*)vec(j + 1), nb, gpart, wy, st, pass);
        }
        LAUNCH(h, k_reduce_multi, (j + 1) < 64 ? (j + 1) : 64, 256, 0, nb, j + 1, gpart, h->gm_h, st, pass);
        CK(cudaGetLastError());
        if (h->nranks > 1) CKR(lvpp_allreduce_sum(h, h->gm_h, j + 1));
        LAUNCH(h, k_gmres_update, nb, 256, 0, Vown, Vb, stride2, j + 1, h->gm_h, (double2*)vec(j + 1), nb, m + 1, gpart, wy, st, pass);
        LAUNCH(h, k_reduce_multi, 1, 256, 0, nb, 1, gpart + (size_t)(m + 1) * nb, h->gm_red + 1, st, pass);
        CK(cudaGetLastError());
        if (h->nranks > 1) CKR(lvpp_allreduce_sum(h, h->gm_red + 1, 1));
        LAUNCH(h, k_gm_after_pass, 1, 1, 0, st, j, pass, h->gm_h, h->gm_red + 1, Hcol, h->gm_cs, h->gm_sn, h->gm_g);
        CK(cudaGetLastError());
      }
      LAUNCH(h, k_gm_scale, nb, 256, 0, Vown, st, (double2*)vec(j + 1));
      CK(cudaGetLastError());
      CK(cudaMemcpyAsync(&ring[slot], st, sizeof(GmState), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaEventRecord(h->gm_ev[slot], h->stream));
      if (j >= 1) {  // look at iteration j - 1 while iteration j keeps the device busy
        const int ps = (j - 1) % GM_RING;
        CK(cudaEventSynchronize(h->gm_ev[ps]));
        over = ring[ps].conv != 0;
        if (ring[ps].ncols == j) CKR(take_samples(ps));  // (not a skipped iteration)
        polled = j;
      }
    }
    // end of the restart cycle (or of the solve): the state after everything queued so far
    CK(cudaMemcpyAsync(&ring[GM_RING], st, sizeof(GmState), cudaMemcpyDeviceToHost, h->stream));
    CKR(lvpp_sync_check_comm(h));
    fin = ring[GM_RING];
    for (int q = polled; q < j; ++q)
      if (q < fin.ncols) CKR(take_samples(q % GM_RING));
    const int k = fin.ncols;  // columns used
    if (fin.conv && fin.reason == LVPP_KSP_DIVERGED_NANORINF) { reason = fin.reason; break; }
    if (k > 0) {
      // back substitution on the device, y += M^-1 (V yv); the combination goes to v_k (v_0..v_{k-1} are the basis)
      LAUNCH(h, k_gm_backsolve, 1, 1, 0, k, m, h->gm_H, h->gm_g, h->gm_yv);
      double* comb = vec(k);
      LAUNCH(h, k_lincomb, nb, 256, 0, Vown, Vb, stride2, k, h->gm_yv, (double2*)comb);
      CK(cudaGetLastError());
      double* z = nullptr;
      CKR(lvpp_mg_vcycle(h, comb, &z));
      LAUNCH(h, k_axpby, nb, 256, 0, Vown, 1.0, (const double2*)z, 1, (double2*)d_y);
      CK(cudaGetLastError());
    }
    if (fin.conv) { reason = fin.reason; break; }
  }
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaEventSynchronize(h->ev1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->t_krylov_ms += ms;
  h->krylov_its += fin.total;
  if (its_out) *its_out = fin.total;
  if (reason_out) *reason_out = reason;
  if (rnorm_out) *rnorm_out = fin.rnorm;
  return 0;
}
